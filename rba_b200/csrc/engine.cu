// Engine: MaskFormer.forward (eval branch) + RbA score as one stream of kernel launches.
//
// Mirrors the reference call stack:
//   maskformer_model.py:255-257  normalise + pad            -> patch_embed kernel
//   modeling/backbone/swin.py:651-678                       -> swin_stage()
//   modeling/pixel_decoder/msdeformattn.py:323-367          -> pixel_decoder()
//   modeling/transformer_decoder/mask2former_transformer_decoder.py:398-470 -> transformer_decoder()
//   maskformer_model.py:294-299,381-386 + evaluate_ood.py:148-150 -> rba_score_fused
//
// Data layout in HBM: activations are token-major (B, H*W, C) — NHWC — end to end, so every Linear and 1x1
// conv is a K-contiguous GEMM and no NCHW<->NHWC transposes exist.  The residual stream stays fp32; every
// tensor whose only consumer is a GEMM is stored as bf16 split planes (same bytes as fp32).
// Memory: one device arena, bump-allocated in a fixed order (pointers are identical for identical shapes, so the
// whole forward is CUDA-graph capturable); per-block temporaries are released stack-wise.
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "kernels.cuh"

extern "C" int rba_score_fused(const float*, const float*, int, int, int, int, int, int, int, float*, float*, void*);

namespace rba {

struct DevTensor {
  float* d = nullptr;
  std::vector<int64_t> shape;
  int64_t numel = 0;
};
struct Planes {
  uint16_t* hi = nullptr;
  uint16_t* lo = nullptr;
  Planes offset(int64_t e) const { return Planes{hi + e, lo + e}; }
};

struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  bool dry = false;
  bool overflow = false;
  void* alloc(size_t bytes) {
    off = (off + 255) & ~size_t(255);
    void* p = dry ? reinterpret_cast<void*>(uintptr_t(0x1000) + off) : (void*)(base + off);
    off += bytes;
    if (off > peak) peak = off;
    if (!dry && off > cap) overflow = true;
    return p;
  }
  float* f32(int64_t n) { return (float*)alloc((size_t)n * 4); }
  Planes planes(int64_t n) {
    Planes p;
    p.hi = (uint16_t*)alloc((size_t)n * 2);
    p.lo = (uint16_t*)alloc((size_t)n * 2);
    return p;
  }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }
};

__global__ void permute_conv3x3_kernel(const float* __restrict__ w, int O, int I, float* __restrict__ out) {
  // (O, I, 3, 3) -> [O][tap][I]
  int64_t total = (int64_t)O * I * 9;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int ci = (int)(i % I);
    int tap = (int)((i / I) % 9);
    int64_t o = i / (9 * (int64_t)I);
    out[i] = w[(o * I + ci) * 9 + tap];
  }
}
__global__ void transpose_kernel(const float* __restrict__ w, int R, int Cc, float* __restrict__ out) {
  int64_t total = (int64_t)R * Cc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % Cc);
    int64_t r = i / Cc;
    out[(int64_t)c * R + r] = w[i];
  }
}

// DenseHybrid head: BatchNorm2d (eval) of mask_features folded into the mask_features 1x1 conv:
//   relu(bn(Wmf y + bmf)) = relu((s . Wmf) y + s (bmf - mean) + beta),  s = gamma / sqrt(var + eps)
__global__ void bn_fold_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var,
                               float eps, int O, int I, float* __restrict__ w_out, float* __restrict__ b_out) {
  int64_t total = (int64_t)O * I;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i / I);
    const float sc = gamma[o] / sqrtf(var[o] + eps);
    w_out[i] = sc * w[i];
    if (i % I == 0) b_out[o] = sc * (b[o] - mean[o]) + beta[o];
  }
}

}  // namespace rba

using namespace rba;

struct rba_model {
  rba_config cfg;
  int device = 0;
  bool finalized = false;
  bool taps_enabled = false;
  int attn_backend = 2;             // 2: tcgen05 + TMA window attention (bf16x3), 1: mma.sync bf16x3 kernel, 0: fp32 CUDA-core kernel
  int gemm_backend = RBA_GEMM_TC;   // tcgen05 bf16x3; RBA_GEMM_BACKEND=ffma selects the exact fp32 FMA kernels
  int fused_score = 1;              // 1: last mask einsum + score in one kernel (score_fused.cu) when pred_masks is not asked for
  int score_func = RBA_SCORE_RBA;   // per-pixel reduction written to the score output (evaluate_ood.py:143-159)
  int include_void = 0;             // 1: semantic_inference_with_void (maskformer_model.py:388-392): K+1 sem_seg planes
  bool has_ood_pred = false;        // DenseHybrid head weights present (predictor.ood_pred.*)
  std::unordered_map<std::string, DevTensor> w;
  std::unordered_map<std::string, Planes> wp;      // split planes of GEMM weights, by key
  std::vector<void*> owned;                         // cudaMalloc'ed blocks (weights, derived)
  std::map<std::pair<int, int>, float*> pos_cache;  // sine position embeddings per (h, w), token-major (h*w, D)
  Arena arena;
  std::vector<void*> retired_arenas;                // outgrown arenas that a captured CUDA graph may still reference
  bool arena_captured = false;                      // a forward into the current arena ran under stream capture
  int64_t arena_generation = 0;                     // bumped whenever the arena is re-allocated
  std::map<int, const uint8_t*> am_force;           // parity aid: per prediction head, decisions to use instead of our own
  std::map<int, uint8_t*> am_dump;                  // parity aid: per prediction head, where to copy our own decisions
  int rB = 0, rH = 0, rW = 0;
  struct Tap { const float* p; int64_t n; };
  std::map<std::string, Tap> taps;

  ~rba_model() {
    for (void* p : owned) cudaFree(p);
    for (void* p : retired_arenas) cudaFree(p);
    if (arena.base) cudaFree(arena.base);
  }
  template <typename T>
  int dmalloc(T** out, size_t count) {
    void* p = nullptr;
    RBA_CUDA(cudaMalloc(&p, count * sizeof(T)));
    owned.push_back(p);
    *out = (T*)p;
    return RBA_OK;
  }
  const DevTensor* get(const std::string& k) const {
    auto it = w.find(k);
    return it == w.end() ? nullptr : &it->second;
  }
};

namespace {

#define RBA_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != RBA_OK) return _rc; \
  } while (0)

int need(rba_model* m, const std::string& k, std::initializer_list<int64_t> shape, const DevTensor** out = nullptr) {
  const DevTensor* t = m->get(k);
  if (!t) return fail(RBA_ERR_STATE, "missing tensor '%s' in state_dict", k.c_str());
  int64_t n = 1;
  for (auto s : shape) n *= s;
  if (t->numel != n) return fail(RBA_ERR_STATE, "tensor '%s' has %lld elements, expected %lld", k.c_str(), (long long)t->numel, (long long)n);
  if (out) *out = t;
  return RBA_OK;
}

// fp32 [rows, cols] device matrix -> planes registered under `key`
int make_planes(rba_model* m, const std::string& key, const float* src, int64_t rows, int cols) {
  Planes p;
  RBA_TRY(m->dmalloc(&p.hi, (size_t)rows * cols));
  RBA_TRY(m->dmalloc(&p.lo, (size_t)rows * cols));
  RBA_TRY(rba_k_split(src, rows, cols, cols, p.hi, p.lo, cols, nullptr));
  m->wp[key] = p;
  return RBA_OK;
}

int linear_planes(rba_model* m, const std::string& key, int64_t N, int64_t K) {
  const DevTensor* t;
  RBA_TRY(need(m, key, {N, K}, &t));
  return make_planes(m, key, t->d, N, (int)K);
}

// channels of res2..res5 as the pixel decoder sees them
int feat_channels(const rba_config& c, int i) { return c.backbone_type == 1 ? (256 << i) : (c.embed_dim << i); }
int resnet_blocks(const rba_config& c, int stage) {
  static const int r50[4] = {3, 4, 6, 3}, r101[4] = {3, 4, 23, 3};
  return (c.resnet_depth == 101 ? r101 : r50)[stage];
}

struct Levels {
  int n;
  int H[4], W[4], start[4];
  int S;
};

}  // namespace

// ---------------------------------------------------------------------------------------------
// creation / weights
// ---------------------------------------------------------------------------------------------
extern "C" int rba_model_create(const rba_config* cfg, int device, rba_model** out) {
  RBA_CHECK(cfg && out, "rba_model_create: null pointer");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(RBA_ERR_CUDA, "rba_model_create: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
  RBA_CHECK(device >= 0 && device < ndev, "rba_model_create: bad device %d", device);
  RBA_CHECK(cfg->backbone_type == 0 || cfg->backbone_type == 1, "unknown backbone type %d", cfg->backbone_type);
  if (cfg->backbone_type == 0) {
    RBA_CHECK(cfg->window_size == 12, "only MODEL.SWIN.WINDOW_SIZE 12 is built (got %d)", cfg->window_size);
    RBA_CHECK(cfg->embed_dim % 32 == 0 && cfg->embed_dim <= 256, "unsupported EMBED_DIM %d", cfg->embed_dim);
    for (int i = 0; i < 4; ++i)
      RBA_CHECK(cfg->num_heads[i] * 32 == (cfg->embed_dim << i), "stage %d: head_dim must be 32", i);
  } else {
    RBA_CHECK(cfg->resnet_depth == 50 || cfg->resnet_depth == 101, "only bottleneck ResNet-50 / -101 are built (got depth %d)",
              cfg->resnet_depth);
  }
  RBA_CHECK(cfg->conv_dim == 256 && cfg->mask_dim == 256 && cfg->nheads == 8, "CONVS_DIM/MASK_DIM 256 and NHEADS 8 expected");
  RBA_CHECK(cfg->num_enc_levels == 1 || cfg->num_enc_levels == 3, "1 or 3 encoder levels supported");
  RBA_CHECK(cfg->size_divisibility > 0 && cfg->size_divisibility % 32 == 0, "SIZE_DIVISIBILITY must be a multiple of 32");
  RBA_CHECK(cfg->num_classes + 1 <= 20, "NUM_CLASSES + 1 must be <= 20");
  RBA_CUDA(cudaSetDevice(device));
  rba_model* m = new rba_model();
  m->cfg = *cfg;
  m->device = device;
  const char* be = getenv("RBA_GEMM_BACKEND");
  if (be && std::string(be) == "tc") m->gemm_backend = RBA_GEMM_TC;
  if (be && std::string(be) == "ffma") m->gemm_backend = RBA_GEMM_FFMA;
  const char* fs = getenv("RBA_FUSED_SCORE");
  if (fs && fs[0] == '0') m->fused_score = 0;
  *out = m;
  return RBA_OK;
}

extern "C" void rba_model_destroy(rba_model* m) { delete m; }

extern "C" int rba_model_set_option(rba_model* m, const char* name, int value) {
  RBA_CHECK(m && name, "rba_model_set_option: null pointer");
  std::string n(name);
  if (n == "taps") { m->taps_enabled = value != 0; m->rB = m->rH = m->rW = 0; }
  else if (n == "gemm_backend") {
    RBA_CHECK(value == RBA_GEMM_FFMA || value == RBA_GEMM_TC, "bad gemm backend %d", value);
    m->gemm_backend = value;
  } else if (n == "attn_backend") {
    RBA_CHECK(value >= 0 && value <= 2, "bad attn backend %d", value);
    m->attn_backend = value;
    m->rB = m->rH = m->rW = 0;
  } else if (n == "score_func") {
    RBA_CHECK(value == RBA_SCORE_RBA || value == RBA_SCORE_ENERGY || value == RBA_SCORE_DENSEHYBRID, "bad score_func %d", value);
    m->score_func = value;
  } else if (n == "include_void") {
    RBA_CHECK(value == 0 || value == 1, "bad include_void %d", value);
    m->include_void = value;
  } else if (n == "fused_score") {
    RBA_CHECK(value == 0 || value == 1, "bad fused_score %d", value);
    m->fused_score = value;
    m->rB = m->rH = m->rW = 0;
  } else return fail(RBA_ERR_INVALID, "unknown option '%s'", name);
  return RBA_OK;
}

extern "C" int rba_model_load_tensor(rba_model* m, const char* key, const float* data, const int64_t* shape, int ndim) {
  RBA_CHECK(m && key && data && (shape || ndim == 0), "rba_model_load_tensor: null pointer");
  RBA_CHECK(!m->finalized, "rba_model_load_tensor: model already finalized");
  RBA_CUDA(cudaSetDevice(m->device));
  DevTensor t;
  t.numel = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); t.numel *= shape[i]; }
  RBA_CHECK(t.numel > 0, "rba_model_load_tensor: empty tensor '%s'", key);
  RBA_TRY(m->dmalloc(&t.d, (size_t)t.numel));
  RBA_CUDA(cudaMemcpy(t.d, data, (size_t)t.numel * 4, cudaMemcpyHostToDevice));
  m->w[key] = t;
  return RBA_OK;
}

// conv (bias-free) + eval-mode batch norm -> folded fp32 weight [O][I] (3x3: re-laid out to [O][tap][Cin]) + bias; planes under
// `key + "weight"` when `as_planes`
static int fold_conv_bn(rba_model* m, const std::string& key, int O, int Cin, int k, bool as_planes) {
  const DevTensor *w, *g, *be, *mu, *va;
  RBA_TRY(need(m, key + "weight", {O, Cin, k, k}, &w));
  RBA_TRY(need(m, key + "norm.weight", {O}, &g));
  RBA_TRY(need(m, key + "norm.bias", {O}, &be));
  RBA_TRY(need(m, key + "norm.running_mean", {O}, &mu));
  RBA_TRY(need(m, key + "norm.running_var", {O}, &va));
  const int64_t I = (int64_t)Cin * k * k;
  const float* src = w->d;
  if (k == 3) {
    float* perm;
    RBA_TRY(m->dmalloc(&perm, (size_t)O * I));
    permute_conv3x3_kernel<<<256, 256>>>(w->d, O, Cin, perm);
    src = perm;
  }
  float *fw, *fb;
  RBA_TRY(m->dmalloc(&fw, (size_t)O * I));
  RBA_TRY(m->dmalloc(&fb, (size_t)O));
  RBA_TRY(bn_fold_conv(src, g->d, be->d, mu->d, va->d, 1e-5f, O, I, fw, fb, nullptr));
  DevTensor wt; wt.d = fw; wt.numel = (int64_t)O * I; wt.shape = {O, I};
  DevTensor bt; bt.d = fb; bt.numel = O; bt.shape = {O};
  m->w[key + "folded.weight"] = wt;
  m->w[key + "folded.bias"] = bt;
  if (as_planes) RBA_TRY(make_planes(m, key + "folded.weight", fw, O, (int)I));
  return RBA_OK;
}

static int finalize_resnet(rba_model* m) {
  const rba_config& c = m->cfg;
  RBA_TRY(fold_conv_bn(m, "backbone.stem.conv1.", 64, 3, 7, false));
  int cin = 64, width = 64;
  for (int i = 0; i < 4; ++i) {
    const int cout = width * 4;
    for (int j = 0; j < resnet_blocks(c, i); ++j) {
      const std::string p = "backbone.res" + std::to_string(i + 2) + "." + std::to_string(j) + ".";
      if (cin != cout) RBA_TRY(fold_conv_bn(m, p + "shortcut.", cout, cin, 1, true));
      RBA_TRY(fold_conv_bn(m, p + "conv1.", width, cin, 1, true));
      RBA_TRY(fold_conv_bn(m, p + "conv2.", width, width, 3, true));
      RBA_TRY(fold_conv_bn(m, p + "conv3.", cout, width, 1, true));
      cin = cout;
    }
    width *= 2;
  }
  return RBA_OK;
}

extern "C" int rba_model_finalize(rba_model* m) {
  RBA_CHECK(m, "rba_model_finalize: null model");
  if (m->finalized) return RBA_OK;
  RBA_CUDA(cudaSetDevice(m->device));
  const rba_config& c = m->cfg;
  const int D = c.conv_dim;
  // ---- backbone ----
  if (c.backbone_type == 1) {
    RBA_TRY(finalize_resnet(m));
  } else {
  RBA_TRY(need(m, "backbone.patch_embed.proj.weight", {c.embed_dim, 3, 4, 4}));
    RBA_TRY(need(m, "backbone.patch_embed.proj.bias", {c.embed_dim}));
    RBA_TRY(need(m, "backbone.patch_embed.norm.weight", {c.embed_dim}));
    RBA_TRY(need(m, "backbone.patch_embed.norm.bias", {c.embed_dim}));
    for (int i = 0; i < 4; ++i) {
      const int64_t C = (int64_t)c.embed_dim << i;
      for (int j = 0; j < c.depths[i]; ++j) {
        std::string p = "backbone.layers." + std::to_string(i) + ".blocks." + std::to_string(j) + ".";
        RBA_TRY(need(m, p + "norm1.weight", {C}));
        RBA_TRY(need(m, p + "norm1.bias", {C}));
        const DevTensor* rpb;
        RBA_TRY(need(m, p + "attn.relative_position_bias_table", {23 * 23, c.num_heads[i]}, &rpb));
        {  // head-major, log2(e)-scaled copy for the tensor-core attention kernel (one contiguous cp.async burst per CTA)
          float* prep;
          RBA_TRY(m->dmalloc(&prep, (size_t)window_attn_bias_floats(c.num_heads[i])));
          RBA_TRY(window_attn_prepare_bias(rpb->d, c.num_heads[i], prep, nullptr));
          DevTensor pt = *rpb;
          pt.d = prep;
          pt.numel = window_attn_bias_floats(c.num_heads[i]);
          m->w[p + "attn.relative_position_bias_prepared"] = pt;
        }
        RBA_TRY(linear_planes(m, p + "attn.qkv.weight", 3 * C, C));
        RBA_TRY(need(m, p + "attn.qkv.bias", {3 * C}));
        RBA_TRY(linear_planes(m, p + "attn.proj.weight", C, C));
        RBA_TRY(need(m, p + "attn.proj.bias", {C}));
        RBA_TRY(need(m, p + "norm2.weight", {C}));
        RBA_TRY(need(m, p + "norm2.bias", {C}));
        RBA_TRY(linear_planes(m, p + "mlp.fc1.weight", 4 * C, C));
        RBA_TRY(need(m, p + "mlp.fc1.bias", {4 * C}));
        RBA_TRY(linear_planes(m, p + "mlp.fc2.weight", C, 4 * C));
        RBA_TRY(need(m, p + "mlp.fc2.bias", {C}));
      }
      std::string n = "backbone.norm" + std::to_string(i) + ".";
      RBA_TRY(need(m, n + "weight", {C}));
      RBA_TRY(need(m, n + "bias", {C}));
      if (i < 3) {
        std::string p = "backbone.layers." + std::to_string(i) + ".downsample.";
        RBA_TRY(need(m, p + "norm.weight", {4 * C}));
        RBA_TRY(need(m, p + "norm.bias", {4 * C}));
        RBA_TRY(linear_planes(m, p + "reduction.weight", 2 * C, 4 * C));
      }
    }
  }
  // ---- pixel decoder ----
  const std::string pd = "sem_seg_head.pixel_decoder.";
  const int L = c.num_enc_levels;
  for (int idx = 0; idx < L; ++idx) {       // input_proj[idx]: low-res first (res5, res4, res3), msdeformattn.py:221-235
    const int64_t Cin = feat_channels(c, 3 - idx);
    std::string p = pd + "input_proj." + std::to_string(idx) + ".";
    RBA_TRY(linear_planes(m, p + "0.weight", D, Cin));
    RBA_TRY(need(m, p + "0.bias", {D}));
    RBA_TRY(need(m, p + "1.weight", {D}));
    RBA_TRY(need(m, p + "1.bias", {D}));
  }
  RBA_TRY(need(m, pd + "transformer.level_embed", {L, D}));
  const int LP = L * c.enc_points, M = c.nheads;
  for (int i = 0; i < c.enc_layers; ++i) {
    std::string p = pd + "transformer.encoder.layers." + std::to_string(i) + ".";
    const DevTensor *so_w, *so_b, *aw_w, *aw_b;
    RBA_TRY(need(m, p + "self_attn.sampling_offsets.weight", {(int64_t)M * LP * 2, D}, &so_w));
    RBA_TRY(need(m, p + "self_attn.sampling_offsets.bias", {(int64_t)M * LP * 2}, &so_b));
    RBA_TRY(need(m, p + "self_attn.attention_weights.weight", {(int64_t)M * LP, D}, &aw_w));
    RBA_TRY(need(m, p + "self_attn.attention_weights.bias", {(int64_t)M * LP}, &aw_b));
    // merged [offsets | attention logits] projection: one GEMM with N = M*LP*3
    float *mw, *mb;
    const int64_t No = (int64_t)M * LP * 2, Na = (int64_t)M * LP;
    RBA_TRY(m->dmalloc(&mw, (size_t)(No + Na) * D));
    RBA_TRY(m->dmalloc(&mb, (size_t)(No + Na)));
    RBA_CUDA(cudaMemcpy(mw, so_w->d, (size_t)No * D * 4, cudaMemcpyDeviceToDevice));
    RBA_CUDA(cudaMemcpy(mw + No * D, aw_w->d, (size_t)Na * D * 4, cudaMemcpyDeviceToDevice));
    RBA_CUDA(cudaMemcpy(mb, so_b->d, (size_t)No * 4, cudaMemcpyDeviceToDevice));
    RBA_CUDA(cudaMemcpy(mb + No, aw_b->d, (size_t)Na * 4, cudaMemcpyDeviceToDevice));
    RBA_TRY(make_planes(m, p + "self_attn.oa.weight", mw, No + Na, D));
    DevTensor bt; bt.d = mb; bt.numel = No + Na; bt.shape = {No + Na};
    m->w[p + "self_attn.oa.bias"] = bt;
    RBA_TRY(linear_planes(m, p + "self_attn.value_proj.weight", D, D));
    RBA_TRY(need(m, p + "self_attn.value_proj.bias", {D}));
    RBA_TRY(linear_planes(m, p + "self_attn.output_proj.weight", D, D));
    RBA_TRY(need(m, p + "self_attn.output_proj.bias", {D}));
    RBA_TRY(need(m, p + "norm1.weight", {D}));
    RBA_TRY(need(m, p + "norm1.bias", {D}));
    RBA_TRY(linear_planes(m, p + "linear1.weight", c.enc_ffn, D));
    RBA_TRY(need(m, p + "linear1.bias", {c.enc_ffn}));
    RBA_TRY(linear_planes(m, p + "linear2.weight", D, c.enc_ffn));
    RBA_TRY(need(m, p + "linear2.bias", {D}));
    RBA_TRY(need(m, p + "norm2.weight", {D}));
    RBA_TRY(need(m, p + "norm2.bias", {D}));
  }
  const int num_fpn = (L == 1) ? 3 : 1;     // log2(min transformer stride) - log2(4), msdeformattn.py:267-268
  for (int k = 1; k <= num_fpn; ++k) {      // adapter_k / layer_k act on res(k+1)
    const int64_t Cin = feat_channels(c, k - 1);
    std::string a = pd + "adapter_" + std::to_string(k) + ".", l = pd + "layer_" + std::to_string(k) + ".";
    RBA_TRY(linear_planes(m, a + "weight", D, Cin));
    RBA_TRY(need(m, a + "norm.weight", {D}));
    RBA_TRY(need(m, a + "norm.bias", {D}));
    const DevTensor* cw;
    RBA_TRY(need(m, l + "weight", {D, D, 3, 3}, &cw));
    float* perm;
    RBA_TRY(m->dmalloc(&perm, (size_t)D * D * 9));
    permute_conv3x3_kernel<<<256, 256>>>(cw->d, D, D, perm);
    RBA_TRY(make_planes(m, l + "weight", perm, D, 9 * D));
    RBA_TRY(need(m, l + "norm.weight", {D}));
    RBA_TRY(need(m, l + "norm.bias", {D}));
  }
  {
    // mask_features 1x1 conv folded into the mask einsum: pred = (E.Wmf) y + E.bmf  (msdeformattn.py:254-260,367;
    // mask2former_transformer_decoder.py:479).  Needs Wmf^T [c_in][o] as the "weight" of E' = E Wmf.
    const DevTensor *mw, *mb;
    RBA_TRY(need(m, pd + "mask_features.weight", {c.mask_dim, D, 1, 1}, &mw));
    RBA_TRY(need(m, pd + "mask_features.bias", {c.mask_dim}, &mb));
    float* wt;
    RBA_TRY(m->dmalloc(&wt, (size_t)c.mask_dim * D));
    transpose_kernel<<<256, 256>>>(mw->d, c.mask_dim, D, wt);
    RBA_TRY(make_planes(m, pd + "mask_features.weightT", wt, D, c.mask_dim));
    RBA_TRY(make_planes(m, pd + "mask_features.bias_row", mb->d, 1, c.mask_dim));
  }
  // ---- transformer decoder ----
  const std::string pr = "sem_seg_head.predictor.";
  if (m->get(pr + "ood_pred.conv.weight")) {
    // DenseHybrid head BNReluConv(hidden_dim, 2, k=1) on mask_features (mask2former_transformer_decoder.py:216-230,365-366)
    RBA_CHECK(c.mask_dim == D, "ood_pred head needs MASK_DIM == HIDDEN_DIM");
    const DevTensor *mw, *mb, *g, *be, *mu, *va;
    RBA_TRY(need(m, pd + "mask_features.weight", {c.mask_dim, D, 1, 1}, &mw));
    RBA_TRY(need(m, pd + "mask_features.bias", {c.mask_dim}, &mb));
    RBA_TRY(need(m, pr + "ood_pred.norm.weight", {D}, &g));
    RBA_TRY(need(m, pr + "ood_pred.norm.bias", {D}, &be));
    RBA_TRY(need(m, pr + "ood_pred.norm.running_mean", {D}, &mu));
    RBA_TRY(need(m, pr + "ood_pred.norm.running_var", {D}, &va));
    float *fw, *fb;
    RBA_TRY(m->dmalloc(&fw, (size_t)c.mask_dim * D));
    RBA_TRY(m->dmalloc(&fb, (size_t)c.mask_dim));
    bn_fold_kernel<<<256, 256>>>(mw->d, mb->d, g->d, be->d, mu->d, va->d, 1e-5f, c.mask_dim, D, fw, fb);
    RBA_TRY(make_planes(m, pr + "ood_pred.fold.weight", fw, c.mask_dim, D));
    DevTensor fbt = *mb;
    fbt.d = fb;
    m->w[pr + "ood_pred.fold.bias"] = fbt;
    const DevTensor* cw;
    RBA_TRY(need(m, pr + "ood_pred.conv.weight", {2, c.mask_dim, 1, 1}, &cw));
    RBA_TRY(make_planes(m, pr + "ood_pred.conv.weight", cw->d, 2, c.mask_dim));
    RBA_TRY(need(m, pr + "ood_pred.conv.bias", {2}));
    m->has_ood_pred = true;
  }
  RBA_CHECK(m->get(pr + "input_proj.0.weight") == nullptr, "predictor.input_proj (CONVS_DIM != HIDDEN_DIM) is not supported");
  for (int i = 0; i < c.dec_layers; ++i) {
    std::string ca = pr + "transformer_cross_attention_layers." + std::to_string(i) + ".";
    std::string sa = pr + "transformer_self_attention_layers." + std::to_string(i) + ".";
    std::string ff = pr + "transformer_ffn_layers." + std::to_string(i) + ".";
    RBA_TRY(linear_planes(m, ca + "multihead_attn.in_proj_weight", 3 * D, D));
    RBA_TRY(need(m, ca + "multihead_attn.in_proj_bias", {3 * D}));
    RBA_TRY(linear_planes(m, ca + "multihead_attn.out_proj.weight", D, D));
    RBA_TRY(need(m, ca + "multihead_attn.out_proj.bias", {D}));
    RBA_TRY(need(m, ca + "norm.weight", {D}));
    RBA_TRY(need(m, ca + "norm.bias", {D}));
    RBA_TRY(linear_planes(m, sa + "self_attn.in_proj_weight", 3 * D, D));
    RBA_TRY(need(m, sa + "self_attn.in_proj_bias", {3 * D}));
    RBA_TRY(linear_planes(m, sa + "self_attn.out_proj.weight", D, D));
    RBA_TRY(need(m, sa + "self_attn.out_proj.bias", {D}));
    RBA_TRY(need(m, sa + "norm.weight", {D}));
    RBA_TRY(need(m, sa + "norm.bias", {D}));
    RBA_TRY(linear_planes(m, ff + "linear1.weight", c.dim_feedforward, D));
    RBA_TRY(need(m, ff + "linear1.bias", {c.dim_feedforward}));
    RBA_TRY(linear_planes(m, ff + "linear2.weight", D, c.dim_feedforward));
    RBA_TRY(need(m, ff + "linear2.bias", {D}));
    RBA_TRY(need(m, ff + "norm.weight", {D}));
    RBA_TRY(need(m, ff + "norm.bias", {D}));
  }
  RBA_TRY(need(m, pr + "decoder_norm.weight", {D}));
  RBA_TRY(need(m, pr + "decoder_norm.bias", {D}));
  RBA_TRY(need(m, pr + "query_feat.weight", {c.num_queries, D}));
  RBA_TRY(need(m, pr + "query_embed.weight", {c.num_queries, D}));
  RBA_TRY(need(m, pr + "level_embed.weight", {L, D}));
  RBA_TRY(linear_planes(m, pr + "class_embed.weight", c.num_classes + 1, D));
  RBA_TRY(need(m, pr + "class_embed.bias", {c.num_classes + 1}));
  for (int i = 0; i < 3; ++i) {
    std::string p = pr + "mask_embed.layers." + std::to_string(i) + ".";
    RBA_TRY(linear_planes(m, p + "weight", i == 2 ? c.mask_dim : D, D));
    RBA_TRY(need(m, p + "bias", {i == 2 ? c.mask_dim : D}));
  }
  RBA_CUDA(cudaDeviceSynchronize());
  m->finalized = true;
  return RBA_OK;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
namespace {

struct Fwd {
  rba_model* m;
  Arena& A;
  cudaStream_t st;
  bool dry;
  int backend;

  const float* W(const std::string& k) { return m->w.at(k).d; }
  Planes P(const std::string& k) { return m->wp.at(k); }

#define RBA_RUN(expr)            \
  do {                           \
    if (!dry) RBA_TRY(expr);     \
  } while (0)

  // C = act(A W^T + bias) (+res)
  int lin(Planes a, int64_t lda, int64_t M, int K, Planes w, int N, const float* bias, int act, const float* res,
          float* c, int64_t ldc, Planes cp = Planes(), int64_t ldcp = 0) {
    rba_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.a_hi = a.hi; g.a_lo = a.lo; g.lda = lda;
    g.w_hi = w.hi; g.w_lo = w.lo; g.ldw = K;
    g.M = (int)M; g.N = N; g.K = K; g.batch = 1;
    g.bias = bias; g.act = act; g.residual = res;
    g.c = c; g.ldc = ldc; g.c_hi = cp.hi; g.c_lo = cp.lo; g.ldcp = ldcp;
    g.backend = backend;
    RBA_RUN(gemm(g, st));
    return RBA_OK;
  }

  int tap(const char* name, const float* p, int64_t n) {
    if (!dry && m->taps_enabled) m->taps[name] = rba_model::Tap{p, n};
    return RBA_OK;
  }

  // sine position embedding (position_encoding.py:29-52, normalize=True), token-major (h*w, D), cached per size
  int pos_embed(int h, int w, int D, const float** out) {
    auto key = std::make_pair(h, w);
    auto it = m->pos_cache.find(key);
    if (it != m->pos_cache.end()) { *out = it->second; return RBA_OK; }
    if (dry) { *out = nullptr; return RBA_OK; }
    const int npf = D / 2;
    std::vector<float> host((size_t)h * w * D);
    std::vector<float> dim_t(npf);
    for (int c = 0; c < npf; ++c) dim_t[c] = powf(10000.0f, (float)(2 * (c / 2)) / (float)npf);
    const float scale = 6.283185307179586f, eps = 1e-6f;
    for (int i = 0; i < h; ++i)
      for (int j = 0; j < w; ++j) {
        const float ye = (float)(i + 1) / ((float)h + eps) * scale;
        const float xe = (float)(j + 1) / ((float)w + eps) * scale;
        float* o = host.data() + ((size_t)i * w + j) * D;
        for (int c = 0; c < npf; ++c) {
          const float py = ye / dim_t[c], px = xe / dim_t[c];
          o[c] = (c % 2 == 0) ? sinf(py) : cosf(py);          // pos = cat(pos_y, pos_x)
          o[npf + c] = (c % 2 == 0) ? sinf(px) : cosf(px);
        }
      }
    float* d;
    RBA_TRY(m->dmalloc(&d, host.size()));
    RBA_CUDA(cudaMemcpy(d, host.data(), host.size() * 4, cudaMemcpyHostToDevice));
    m->pos_cache[key] = d;
    *out = d;
    return RBA_OK;
  }
};

}  // namespace

static int forward_impl(rba_model* m, const void* images, int img_dtype, int B, int H, int W, float* rba_out,
                        float* sem_seg, float* pred_logits_out, float* pred_masks_out, float* ood_pred_out, cudaStream_t st,
                        bool dry) {
  const rba_config& c = m->cfg;
  Arena& A = m->arena;
  A.off = 0;
  A.peak = 0;
  A.dry = dry;
  A.overflow = false;
  if (!dry) m->taps.clear();
  Fwd F{m, A, st, dry, m->gemm_backend};
  const int sd = c.size_divisibility;
  const int Hp = (H + sd - 1) / sd * sd, Wp = (W + sd - 1) / sd * sd;
  const int D = c.conv_dim, Q = c.num_queries, K1 = c.num_classes + 1, ws = c.window_size;
  const float eps = 1e-5f;

  Planes resp[4];
  int resH[4], resW[4], resC[4];
  if (c.backbone_type == 1) {
    // ================= ResNet backbone (detectron2 build_resnet_backbone; bottleneck blocks, stride in the 3x3) =================
    const int H2 = Hp / 2, W2 = Wp / 2;
    int Hs = Hp / 4, Wsz = Wp / 4;
    float* x = A.f32((int64_t)B * Hs * Wsz * 64);
    Planes xp = A.planes((int64_t)B * Hs * Wsz * 64);
    {
      const size_t mk = A.mark();
      float* t0 = A.f32((int64_t)B * H2 * W2 * 64);
      RBA_RUN(stem_conv(images, img_dtype, B, H, W, Hp, Wp, c.pixel_mean, c.pixel_std, F.W("backbone.stem.conv1.folded.weight"),
                        F.W("backbone.stem.conv1.folded.bias"), t0, st));
      RBA_RUN(maxpool3x3s2(t0, B, H2, W2, 64, x, xp.hi, xp.lo, st));
      A.release(mk);
    }
    int cin = 64, width = 64;
    for (int i = 0; i < 4; ++i) {
      const int cout = width * 4;
      for (int j = 0; j < resnet_blocks(c, i); ++j) {
        const std::string p = "backbone.res" + std::to_string(i + 2) + "." + std::to_string(j) + ".";
        const int stride = (j == 0 && i > 0) ? 2 : 1;
        const int Ho = Hs / stride, Wo = Wsz / stride;
        const int64_t T = (int64_t)B * Hs * Wsz, To = (int64_t)B * Ho * Wo;
        // outputs of the block live below the mark; temporaries above it
        float* xo = A.f32(To * cout);
        Planes xop = A.planes(To * cout);
        const size_t mk = A.mark();
        Planes t1 = A.planes(T * width);
        RBA_TRY(F.lin(xp, cin, T, cin, F.P(p + "conv1.folded.weight"), width, F.W(p + "conv1.folded.bias"), RBA_ACT_RELU, nullptr,
                      nullptr, 0, t1, width));
        float* y2 = A.f32(T * width);
        Planes w2 = F.P(p + "conv2.folded.weight");
        RBA_RUN(conv3x3(t1.hi, t1.lo, w2.hi, w2.lo, B, Hs, Wsz, width, width, y2, F.backend, st));
        Planes t2 = A.planes(To * width);
        RBA_RUN(bias_act_sub(y2, F.W(p + "conv2.folded.bias"), B, Hs, Wsz, width, stride, 1, nullptr, t2.hi, t2.lo, st));
        const float* sc = x;
        if (cin != cout) {
          Planes xs = xp;
          if (stride == 2) {
            xs = A.planes(To * cin);
            RBA_RUN(bias_act_sub(x, nullptr, B, Hs, Wsz, cin, 2, 0, nullptr, xs.hi, xs.lo, st));
          }
          float* scf = A.f32(To * cout);
          RBA_TRY(F.lin(xs, cin, To, cin, F.P(p + "shortcut.folded.weight"), cout, F.W(p + "shortcut.folded.bias"), RBA_ACT_NONE,
                        nullptr, scf, cout));
          sc = scf;
        }
        RBA_TRY(F.lin(t2, width, To, width, F.P(p + "conv3.folded.weight"), cout, F.W(p + "conv3.folded.bias"), RBA_ACT_NONE, sc, xo,
                      cout));
        RBA_RUN(bias_act_sub(xo, nullptr, B, Ho, Wo, cout, 1, 1, xo, xop.hi, xop.lo, st));   // ReLU after the residual add
        A.release(mk);
        x = xo; xp = xop; cin = cout; Hs = Ho; Wsz = Wo;
      }
      resp[i] = xp; resH[i] = Hs; resW[i] = Wsz; resC[i] = cout;
      F.tap(("res" + std::to_string(i + 2)).c_str(), x, (int64_t)B * Hs * Wsz * cout);
      width *= 2;
    }
  } else {
  // ================= backbone (swin.py:651-678) =================
  int Hs = Hp / 4, Wsz = Wp / 4;
  int C = c.embed_dim;
  float* x = A.f32((int64_t)B * Hs * Wsz * C);
  RBA_RUN(patch_embed(images, img_dtype, B, H, W, Hp, Wp, c.pixel_mean, c.pixel_std, F.W("backbone.patch_embed.proj.weight"),
                      F.W("backbone.patch_embed.proj.bias"), F.W("backbone.patch_embed.norm.weight"),
                      F.W("backbone.patch_embed.norm.bias"), C, x, st));
  for (int i = 0; i < 4; ++i) {
    const int64_t N = (int64_t)Hs * Wsz, T = (int64_t)B * N;
    const int heads = c.num_heads[i];
    SwinGeom g = make_swin_geom(Hs, Wsz, ws, 0);
    const int64_t RW = (int64_t)B * g.nWh * g.nWw * ws * ws;
    for (int j = 0; j < c.depths[i]; ++j) {
      const std::string p = "backbone.layers." + std::to_string(i) + ".blocks." + std::to_string(j) + ".";
      const int shift = (j % 2 == 0) ? 0 : ws / 2;
      const size_t mk = A.mark();
      Planes a1 = A.planes(RW * C);
      RBA_RUN(layernorm(x, F.W(p + "norm1.weight"), F.W(p + "norm1.bias"), 1, B, Hs, Wsz, C, ws, shift, eps, nullptr, a1.hi,
                        a1.lo, st));
      Planes ao = A.planes(RW * C);
      if (m->attn_backend == 2) {
        // q | k | v leave the QKV GEMM already in the (window, part, head) tiled layout: one bulk copy per operand tile
        Planes qkv = A.planes(RW * 3 * C);
        const bool tiled = F.backend == RBA_GEMM_TC;
        {
          rba_gemm_args ga;
          memset(&ga, 0, sizeof(ga));
          Planes w = F.P(p + "attn.qkv.weight");
          ga.a_hi = a1.hi; ga.a_lo = a1.lo; ga.lda = C;
          ga.w_hi = w.hi; ga.w_lo = w.lo; ga.ldw = C;
          ga.M = (int)RW; ga.N = 3 * C; ga.K = C; ga.batch = 1;
          ga.bias = F.W(p + "attn.qkv.bias");
          ga.c_hi = qkv.hi; ga.c_lo = qkv.lo; ga.ldcp = 3 * C;
          ga.qkv_tile_heads = tiled ? heads : 0;
          ga.backend = F.backend;
          RBA_RUN(gemm(ga, st));
        }
        RBA_RUN(window_attn_tc(qkv.hi, qkv.lo, F.W(p + "attn.relative_position_bias_prepared"), B, Hs, Wsz, C, heads, ws, shift,
                               ao.hi, ao.lo, st, tiled ? 1 : 0));
      } else if (m->attn_backend == 1) {
        Planes qkv = A.planes(RW * 3 * C);
        RBA_TRY(F.lin(a1, C, RW, C, F.P(p + "attn.qkv.weight"), 3 * C, F.W(p + "attn.qkv.bias"), RBA_ACT_NONE, nullptr, nullptr, 0,
                      qkv, 3 * C));
        RBA_RUN(window_attn_planes(qkv.hi, qkv.lo, F.W(p + "attn.relative_position_bias_table"),
                                   F.W(p + "attn.relative_position_bias_prepared"), B, Hs, Wsz, C, heads, ws, shift,
                                   ao.hi, ao.lo, st));
      } else {
        float* qkv = A.f32(RW * 3 * C);
        RBA_TRY(F.lin(a1, C, RW, C, F.P(p + "attn.qkv.weight"), 3 * C, F.W(p + "attn.qkv.bias"), RBA_ACT_NONE, nullptr, qkv,
                      3 * C));
        RBA_RUN(window_attn(qkv, F.W(p + "attn.relative_position_bias_table"), B, Hs, Wsz, C, heads, ws, shift, ao.hi, ao.lo,
                            st));
      }
      {  // x = shortcut + window_reverse(proj(attn))  (swin.py:169,277-292): scatter epilogue, in place
        rba_gemm_args ga;
        memset(&ga, 0, sizeof(ga));
        Planes w = F.P(p + "attn.proj.weight");
        ga.a_hi = ao.hi; ga.a_lo = ao.lo; ga.lda = C;
        ga.w_hi = w.hi; ga.w_lo = w.lo; ga.ldw = C;
        ga.M = (int)RW; ga.N = C; ga.K = C; ga.batch = 1;
        ga.bias = F.W(p + "attn.proj.bias"); ga.residual = x; ga.c = x; ga.ldc = C;
        ga.swin_map = 1; ga.sw_H = Hs; ga.sw_W = Wsz; ga.sw_ws = ws; ga.sw_shift = shift;
        ga.backend = F.backend;
        RBA_RUN(gemm(ga, st));
      }
      Planes a2 = A.planes(T * C);
      RBA_RUN(layernorm(x, F.W(p + "norm2.weight"), F.W(p + "norm2.bias"), 0, B, Hs, Wsz, C, 0, 0, eps, nullptr, a2.hi, a2.lo,
                        st));
      Planes hid = A.planes(T * 4 * C);
      RBA_TRY(F.lin(a2, C, T, C, F.P(p + "mlp.fc1.weight"), 4 * C, F.W(p + "mlp.fc1.bias"), RBA_ACT_GELU, nullptr, nullptr, 0,
                    hid, 4 * C));
      RBA_TRY(F.lin(hid, 4 * C, T, 4 * C, F.P(p + "mlp.fc2.weight"), C, F.W(p + "mlp.fc2.bias"), RBA_ACT_NONE, x, x, C));
      A.release(mk);
    }
    // out norm (swin.py:671-676) -> planes feeding the pixel decoder's 1x1 convs (+ fp32 tap)
    resp[i] = A.planes(T * C);
    resH[i] = Hs; resW[i] = Wsz; resC[i] = C;
    float* tapbuf = (m->taps_enabled) ? A.f32(T * C) : nullptr;
    const std::string n = "backbone.norm" + std::to_string(i) + ".";
    RBA_RUN(layernorm(x, F.W(n + "weight"), F.W(n + "bias"), 0, B, Hs, Wsz, C, 0, 0, eps, tapbuf, resp[i].hi, resp[i].lo, st));
    if (tapbuf) F.tap(("res" + std::to_string(i + 2)).c_str(), tapbuf, T * C);
    if (i < 3) {
      RBA_CHECK(Hs % 2 == 0 && Wsz % 2 == 0, "odd token grid %dx%d at stage %d", Hs, Wsz, i);
      const std::string p = "backbone.layers." + std::to_string(i) + ".downsample.";
      const int64_t T2 = T / 4;
      float* xn = A.f32(T2 * 2 * C);
      const size_t mk = A.mark();
      Planes mg = A.planes(T2 * 4 * C);
      RBA_RUN(layernorm(x, F.W(p + "norm.weight"), F.W(p + "norm.bias"), 2, B, Hs, Wsz, C, 0, 0, eps, nullptr, mg.hi, mg.lo, st));
      RBA_TRY(F.lin(mg, 4 * C, T2, 4 * C, F.P(p + "reduction.weight"), 2 * C, nullptr, RBA_ACT_NONE, nullptr, xn, 2 * C));
      A.release(mk);
      x = xn;
      Hs /= 2; Wsz /= 2; C *= 2;
    }
  }

  }

  // ================= pixel decoder (msdeformattn.py:323-367) =================
  const std::string pd = "sem_seg_head.pixel_decoder.";
  const int L = c.num_enc_levels;
  Levels lv;
  lv.n = L; lv.S = 0;
  for (int l = 0; l < L; ++l) {             // level l <- res(5-l): low-res first
    lv.H[l] = resH[3 - l]; lv.W[l] = resW[3 - l]; lv.start[l] = lv.S; lv.S += lv.H[l] * lv.W[l];
  }
  const int S = lv.S;
  const int64_t BS = (int64_t)B * S;
  float* src = A.f32(BS * D);               // encoder state (B, S, D)
  float* lvl_pos = A.f32((int64_t)S * D);   // pos + level_embed, shared by all images
  double* gnws = nullptr;
  {
    int64_t need_ws = 0;
    for (int i = 0; i < 4; ++i) need_ws = std::max(need_ws, groupnorm_ws_doubles(B, resH[i], resW[i], D, 32));
    gnws = (double*)A.alloc((size_t)need_ws * 8);
  }
  for (int l = 0; l < L; ++l) {
    const int ri = 3 - l;
    const std::string p = pd + "input_proj." + std::to_string(l) + ".";
    const int64_t Nl = (int64_t)lv.H[l] * lv.W[l];
    const size_t mk = A.mark();
    float* t = A.f32((int64_t)B * Nl * D);
    RBA_TRY(F.lin(resp[ri], resC[ri], (int64_t)B * Nl, resC[ri], F.P(p + "0.weight"), D, F.W(p + "0.bias"), RBA_ACT_NONE,
                  nullptr, t, D));
    RBA_RUN(groupnorm(t, Nl * D, F.W(p + "1.weight"), F.W(p + "1.bias"), B, lv.H[l], lv.W[l], D, 32, eps, nullptr, 0, 0, 0, 0,
                      src + (int64_t)lv.start[l] * D, nullptr, nullptr, (int64_t)S * D, gnws, st));
    const float* pe;
    RBA_TRY(F.pos_embed(lv.H[l], lv.W[l], D, &pe));
    RBA_RUN(ew_add(pe, F.W(pd + "transformer.level_embed") + (int64_t)l * D, Nl, D, 1, lvl_pos + (int64_t)lv.start[l] * D,
                   nullptr, nullptr, st));
    A.release(mk);
  }
  {
    const int M = c.nheads, P = c.enc_points, LP = L * P, NOA = M * LP * 3;
    Planes srcp = A.planes(BS * D);
    RBA_RUN(ew_add(src, nullptr, BS, D, 1, nullptr, srcp.hi, srcp.lo, st));
    for (int i = 0; i < c.enc_layers; ++i) {
      const std::string p = pd + "transformer.encoder.layers." + std::to_string(i) + ".";
      const size_t mk = A.mark();
      Planes qp = A.planes(BS * D);
      RBA_RUN(ew_add(src, lvl_pos, BS, D, S, nullptr, qp.hi, qp.lo, st));
      float* value = A.f32(BS * D);
      RBA_TRY(F.lin(srcp, D, BS, D, F.P(p + "self_attn.value_proj.weight"), D, F.W(p + "self_attn.value_proj.bias"),
                    RBA_ACT_NONE, nullptr, value, D));
      float* oa = A.f32(BS * NOA);
      RBA_TRY(F.lin(qp, D, BS, D, F.P(p + "self_attn.oa.weight"), NOA, F.W(p + "self_attn.oa.bias"), RBA_ACT_NONE, nullptr, oa,
                    NOA));
      Planes ms = A.planes(BS * D);
      RBA_RUN(msda_fused(value, lv.H, lv.W, oa, B, S, M, D / M, L, P, ms.hi, ms.lo, st));
      float* t = A.f32(BS * D);
      RBA_TRY(F.lin(ms, D, BS, D, F.P(p + "self_attn.output_proj.weight"), D, F.W(p + "self_attn.output_proj.bias"),
                    RBA_ACT_NONE, src, t, D));
      Planes s1p = A.planes(BS * D);
      float* s1 = A.f32(BS * D);
      RBA_RUN(layernorm(t, F.W(p + "norm1.weight"), F.W(p + "norm1.bias"), 0, 1, 1, (int)BS, D, 0, 0, eps, s1, s1p.hi, s1p.lo, st));
      Planes hid = A.planes(BS * c.enc_ffn);
      RBA_TRY(F.lin(s1p, D, BS, D, F.P(p + "linear1.weight"), c.enc_ffn, F.W(p + "linear1.bias"), RBA_ACT_RELU, nullptr, nullptr,
                    0, hid, c.enc_ffn));
      RBA_TRY(F.lin(hid, c.enc_ffn, BS, c.enc_ffn, F.P(p + "linear2.weight"), D, F.W(p + "linear2.bias"), RBA_ACT_NONE, s1, t, D));
      RBA_RUN(layernorm(t, F.W(p + "norm2.weight"), F.W(p + "norm2.bias"), 0, 1, 1, (int)BS, D, 0, 0, eps, src, srcp.hi, srcp.lo,
                        st));
      A.release(mk);
    }
  }
  F.tap("enc_out", src, BS * D);
  // FPN top-down (msdeformattn.py:352-360)
  const int num_fpn = (L == 1) ? 3 : 1;
  const float* prev = src + (int64_t)lv.start[L - 1] * D;   // out[-1]: highest-resolution encoder level
  int64_t prev_bs = (int64_t)S * D;
  int ph = lv.H[L - 1], pw = lv.W[L - 1];
  Planes ypl;                                               // planes of the last FPN output (feeds the mask einsum)
  int mH = ph, mW = pw;
  for (int idx = 0; idx < num_fpn; ++idx) {
    const int k = num_fpn - idx;                            // adapter_k / layer_k <-> res(k+1)
    const int ri = k - 1;
    const std::string a = pd + "adapter_" + std::to_string(k) + ".", l = pd + "layer_" + std::to_string(k) + ".";
    const int fh = resH[ri], fw = resW[ri];
    const int64_t T = (int64_t)B * fh * fw;
    const bool last = idx == num_fpn - 1;
    float* outf = A.f32(T * D);
    if (last) ypl = A.planes(T * D);
    const size_t mk = A.mark();
    float* lat = A.f32(T * D);
    RBA_TRY(F.lin(resp[ri], resC[ri], T, resC[ri], F.P(a + "weight"), D, nullptr, RBA_ACT_NONE, nullptr, lat, D));
    Planes yp = A.planes(T * D);
    RBA_RUN(groupnorm(lat, (int64_t)fh * fw * D, F.W(a + "norm.weight"), F.W(a + "norm.bias"), B, fh, fw, D, 32, eps, prev,
                      prev_bs, ph, pw, 0, nullptr, yp.hi, yp.lo, (int64_t)fh * fw * D, gnws, st));
    float* z = lat;                                         // lateral is dead: reuse for the conv output
    Planes lw = F.P(l + "weight");
    RBA_RUN(conv3x3(yp.hi, yp.lo, lw.hi, lw.lo, B, fh, fw, D, D, z, F.backend, st));
    RBA_RUN(groupnorm(z, (int64_t)fh * fw * D, F.W(l + "norm.weight"), F.W(l + "norm.bias"), B, fh, fw, D, 32, eps, nullptr, 0,
                      0, 0, 1, outf, last ? ypl.hi : nullptr, last ? ypl.lo : nullptr, (int64_t)fh * fw * D, gnws, st));
    A.release(mk);
    F.tap(("fpn_res" + std::to_string(k + 1)).c_str(), outf, T * D);
    prev = outf; prev_bs = (int64_t)fh * fw * D; ph = fh; pw = fw;
    mH = fh; mW = fw;
  }
  const int64_t HWm = (int64_t)mH * mW;

  // ================= transformer decoder (mask2former_transformer_decoder.py:398-470) =================
  const std::string pr = "sem_seg_head.predictor.";
  const int64_t BQ = (int64_t)B * Q;
  // memory per level: src_l = enc_l + level_embed[l]; keys use src_l + pos_l
  Planes kin[3], vin[3];
  for (int l = 0; l < L; ++l) {
    const int64_t Nl = (int64_t)lv.H[l] * lv.W[l];
    kin[l] = A.planes((int64_t)B * Nl * D);
    vin[l] = A.planes((int64_t)B * Nl * D);
    const float* pe;
    RBA_TRY(F.pos_embed(lv.H[l], lv.W[l], D, &pe));
    const size_t mk = A.mark();
    float* pl = A.f32(Nl * D);
    RBA_RUN(ew_add(pe, F.W(pr + "level_embed.weight") + (int64_t)l * D, Nl, D, 1, pl, nullptr, nullptr, st));
    // gather level l of every image into a contiguous (B, Nl, D) tensor, adding level_embed
    for (int b = 0; b < B; ++b) {
      RBA_RUN(ew_add(src + ((int64_t)b * S + lv.start[l]) * D, F.W(pr + "level_embed.weight") + (int64_t)l * D, Nl, D, 1,
                     nullptr, vin[l].hi + (int64_t)b * Nl * D, vin[l].lo + (int64_t)b * Nl * D, st));
      RBA_RUN(ew_add(src + ((int64_t)b * S + lv.start[l]) * D, pl, Nl, D, Nl, nullptr, kin[l].hi + (int64_t)b * Nl * D,
                     kin[l].lo + (int64_t)b * Nl * D, st));
    }
    A.release(mk);
  }
  float* out = A.f32(BQ * D);
  RBA_RUN(ew_add(nullptr, F.W(pr + "query_feat.weight"), BQ, D, Q, out, nullptr, nullptr, st));
  float* cls = A.f32(BQ * K1);
  float* masks = A.f32(BQ * HWm);
  int maxS = 0;
  for (int l = 0; l < L; ++l) maxS = std::max(maxS, lv.H[l] * lv.W[l]);
  uint8_t* am = (uint8_t*)A.alloc((size_t)BQ * maxS);

  // The last prediction head feeds only the score: its mask einsum is fused into the score kernel (score_fused.cu)
  // unless the caller wants pred_masks itself.
  const bool special = m->score_func != RBA_SCORE_RBA || m->include_void;   // only the fused kernel implements these
  const bool fuse_last = (m->fused_score || special) && (!pred_masks_out || special) && F.backend == RBA_GEMM_TC &&
                         einsum_score_supported(Q, c.num_classes, D) && (rba_out || sem_seg);
  if (special && (rba_out || sem_seg))
    RBA_CHECK(fuse_last, "score_func / include_void need the tensor-core backend and Q <= 104, K + 1 <= 24");
  Planes ef = A.planes(BQ * D);                 // E' = mask_embed . Wmf (kept for the fused kernel)
  float* bq = A.f32(BQ);                        // b' = mask_embed . bmf

  int head_idx = 0;
  auto heads = [&](int target_level, bool last) -> int {   // forward_prediction_heads (:472-489)
    const size_t mk = A.mark();
    const int hd = head_idx++;
    Planes dn = A.planes(BQ * D);
    RBA_RUN(layernorm(out, F.W(pr + "decoder_norm.weight"), F.W(pr + "decoder_norm.bias"), 0, 1, 1, (int)BQ, D, 0, 0, eps,
                      nullptr, dn.hi, dn.lo, st));
    RBA_TRY(F.lin(dn, D, BQ, D, F.P(pr + "class_embed.weight"), K1, F.W(pr + "class_embed.bias"), RBA_ACT_NONE, nullptr, cls, K1));
    Planes m1 = A.planes(BQ * D), m2 = A.planes(BQ * D), me = A.planes(BQ * c.mask_dim);
    RBA_TRY(F.lin(dn, D, BQ, D, F.P(pr + "mask_embed.layers.0.weight"), D, F.W(pr + "mask_embed.layers.0.bias"), RBA_ACT_RELU,
                  nullptr, nullptr, 0, m1, D));
    RBA_TRY(F.lin(m1, D, BQ, D, F.P(pr + "mask_embed.layers.1.weight"), D, F.W(pr + "mask_embed.layers.1.bias"), RBA_ACT_RELU,
                  nullptr, nullptr, 0, m2, D));
    RBA_TRY(F.lin(m2, D, BQ, D, F.P(pr + "mask_embed.layers.2.weight"), c.mask_dim, F.W(pr + "mask_embed.layers.2.bias"),
                  RBA_ACT_NONE, nullptr, nullptr, 0, me, c.mask_dim));
    // fold mask_features: E' = E Wmf, b' = E bmf
    RBA_TRY(F.lin(me, c.mask_dim, BQ, c.mask_dim, F.P(pd + "mask_features.weightT"), D, nullptr, RBA_ACT_NONE, nullptr, nullptr,
                  0, ef, D));
    RBA_TRY(F.lin(me, c.mask_dim, BQ, c.mask_dim, F.P(pd + "mask_features.bias_row"), 1, nullptr, RBA_ACT_NONE, nullptr, bq, 1));
    if (!(last && fuse_last && !pred_masks_out)) {  // masks[b] (Q, HW) = E'[b] (Q, D) . y[b]^T (HW, D) + b'[b]   (einsum "bqc,bchw->bqhw", :479)
      rba_gemm_args ga;
      memset(&ga, 0, sizeof(ga));
      ga.a_hi = ef.hi; ga.a_lo = ef.lo; ga.lda = D; ga.a_bstride = (int64_t)Q * D;
      ga.w_hi = ypl.hi; ga.w_lo = ypl.lo; ga.ldw = D; ga.w_bstride = HWm * D;
      ga.M = Q; ga.N = (int)HWm; ga.K = D; ga.batch = B;
      ga.bias = bq; ga.bias_per_row = 1; ga.bias_bstride = Q;
      ga.c = masks; ga.ldc = HWm; ga.c_bstride = (int64_t)Q * HWm;
      ga.backend = F.backend;
      RBA_RUN(gemm(ga, st));
      if (!last) {
        RBA_RUN(attn_mask(masks, B, Q, mH, mW, lv.H[target_level], lv.W[target_level], am, st));
        if (!dry) {   // parity aids (rba_model_debug_attn_mask): export our decisions / run on the reference's
          const size_t nb = (size_t)BQ * lv.H[target_level] * lv.W[target_level];
          auto du = m->am_dump.find(hd);
          if (du != m->am_dump.end() && du->second) RBA_CUDA(cudaMemcpyAsync(du->second, am, nb, cudaMemcpyDeviceToDevice, st));
          auto fo = m->am_force.find(hd);
          if (fo != m->am_force.end() && fo->second) RBA_CUDA(cudaMemcpyAsync(am, fo->second, nb, cudaMemcpyDeviceToDevice, st));
        }
      }
    }
    A.release(mk);
    return RBA_OK;
  };

  RBA_TRY(heads(0, c.dec_layers == 0));
  const float* qe = F.W(pr + "query_embed.weight");
  for (int i = 0; i < c.dec_layers; ++i) {
    const int l = i % L;
    const int64_t Nl = (int64_t)lv.H[l] * lv.W[l];
    const size_t mk = A.mark();
    const std::string ca = pr + "transformer_cross_attention_layers." + std::to_string(i) + ".";
    const std::string sa = pr + "transformer_self_attention_layers." + std::to_string(i) + ".";
    const std::string ff = pr + "transformer_ffn_layers." + std::to_string(i) + ".";
    // ---- masked cross attention (:435-440) ----
    {
      Planes w = F.P(ca + "multihead_attn.in_proj_weight");
      const float* bi = F.W(ca + "multihead_attn.in_proj_bias");
      Planes qin = A.planes(BQ * D);
      RBA_RUN(ew_add(out, qe, BQ, D, Q, nullptr, qin.hi, qin.lo, st));
      float* qp = A.f32(BQ * D);
      RBA_TRY(F.lin(qin, D, BQ, D, w, D, bi, RBA_ACT_NONE, nullptr, qp, D));
      float* kp = A.f32((int64_t)B * Nl * D);
      float* vp = A.f32((int64_t)B * Nl * D);
      RBA_TRY(F.lin(kin[l], D, (int64_t)B * Nl, D, w.offset((int64_t)D * D), D, bi + D, RBA_ACT_NONE, nullptr, kp, D));
      RBA_TRY(F.lin(vin[l], D, (int64_t)B * Nl, D, w.offset((int64_t)2 * D * D), D, bi + 2 * D, RBA_ACT_NONE, nullptr, vp, D));
      Planes ao = A.planes(BQ * D);
      float* mws = A.f32(mha_workspace_floats(B, Q, (int)Nl, c.nheads));
      RBA_RUN(mha(qp, D, kp, D, vp, D, am, B, Q, (int)Nl, D, c.nheads, ao.hi, ao.lo, mws, st));
      float* t = A.f32(BQ * D);
      RBA_TRY(F.lin(ao, D, BQ, D, F.P(ca + "multihead_attn.out_proj.weight"), D, F.W(ca + "multihead_attn.out_proj.bias"),
                    RBA_ACT_NONE, out, t, D));
      RBA_RUN(layernorm(t, F.W(ca + "norm.weight"), F.W(ca + "norm.bias"), 0, 1, 1, (int)BQ, D, 0, 0, eps, out, nullptr, nullptr, st));
    }
    // ---- self attention (:442-446) ----
    {
      Planes w = F.P(sa + "self_attn.in_proj_weight");
      const float* bi = F.W(sa + "self_attn.in_proj_bias");
      Planes qk = A.planes(BQ * D), vi = A.planes(BQ * D);
      RBA_RUN(ew_add(out, qe, BQ, D, Q, nullptr, qk.hi, qk.lo, st));
      RBA_RUN(ew_add(out, nullptr, BQ, D, 1, nullptr, vi.hi, vi.lo, st));
      float* qkp = A.f32(BQ * 2 * D);       // [q | k] in one GEMM (rows 0..2D of in_proj)
      RBA_TRY(F.lin(qk, D, BQ, D, w, 2 * D, bi, RBA_ACT_NONE, nullptr, qkp, 2 * D));
      float* vp = A.f32(BQ * D);
      RBA_TRY(F.lin(vi, D, BQ, D, w.offset((int64_t)2 * D * D), D, bi + 2 * D, RBA_ACT_NONE, nullptr, vp, D));
      Planes ao = A.planes(BQ * D);
      float* mws = A.f32(mha_workspace_floats(B, Q, Q, c.nheads));
      RBA_RUN(mha(qkp, 2 * D, qkp + D, 2 * D, vp, D, nullptr, B, Q, Q, D, c.nheads, ao.hi, ao.lo, mws, st));
      float* t = A.f32(BQ * D);
      RBA_TRY(F.lin(ao, D, BQ, D, F.P(sa + "self_attn.out_proj.weight"), D, F.W(sa + "self_attn.out_proj.bias"), RBA_ACT_NONE, out,
                    t, D));
      RBA_RUN(layernorm(t, F.W(sa + "norm.weight"), F.W(sa + "norm.bias"), 0, 1, 1, (int)BQ, D, 0, 0, eps, out, nullptr, nullptr, st));
    }
    // ---- FFN (:449-451) ----
    {
      Planes oi = A.planes(BQ * D), hid = A.planes(BQ * c.dim_feedforward);
      RBA_RUN(ew_add(out, nullptr, BQ, D, 1, nullptr, oi.hi, oi.lo, st));
      RBA_TRY(F.lin(oi, D, BQ, D, F.P(ff + "linear1.weight"), c.dim_feedforward, F.W(ff + "linear1.bias"), RBA_ACT_RELU, nullptr,
                    nullptr, 0, hid, c.dim_feedforward));
      float* t = A.f32(BQ * D);
      RBA_TRY(F.lin(hid, c.dim_feedforward, BQ, c.dim_feedforward, F.P(ff + "linear2.weight"), D, F.W(ff + "linear2.bias"),
                    RBA_ACT_NONE, out, t, D));
      RBA_RUN(layernorm(t, F.W(ff + "norm.weight"), F.W(ff + "norm.bias"), 0, 1, 1, (int)BQ, D, 0, 0, eps, out, nullptr, nullptr, st));
    }
    A.release(mk);
    RBA_TRY(heads((i + 1) % L, i == c.dec_layers - 1));
  }
  F.tap("dec_out", out, BQ * D);

  // ================= outputs =================
  RBA_CHECK(mH * 4 == Hp && mW * 4 == Wp, "mask resolution %dx%d is not 1/4 of the padded image %dx%d", mH, mW, Hp, Wp);
  if (pred_logits_out) RBA_RUN((cudaMemcpyAsync(pred_logits_out, cls, (size_t)BQ * K1 * 4, cudaMemcpyDeviceToDevice, st) == cudaSuccess) ? RBA_OK : fail(RBA_ERR_CUDA, "copy pred_logits failed"));
  if (pred_masks_out) RBA_RUN((cudaMemcpyAsync(pred_masks_out, masks, (size_t)BQ * HWm * 4, cudaMemcpyDeviceToDevice, st) == cudaSuccess) ? RBA_OK : fail(RBA_ERR_CUDA, "copy pred_masks failed"));
  if (rba_out || sem_seg) {
    float* ro = rba_out;
    if (!ro) ro = A.f32((int64_t)B * H * W);
    if (fuse_last)
      RBA_RUN(einsum_score_launch(ef.hi, ef.lo, bq, ypl.hi, ypl.lo, cls, B, Q, c.num_classes, D, mH, mW, H, W,
                                  m->score_func == RBA_SCORE_DENSEHYBRID ? RBA_SCORE_ENERGY : m->score_func, m->include_void,
                                  ro, sem_seg, st));
    else
      RBA_RUN(rba_score_fused(masks, cls, B, Q, c.num_classes, mH, mW, H, W, ro, sem_seg, (void*)st));
  }
  // ---- DenseHybrid head: ood_pred = conv1x1(relu(bn(mask_features))), resized with align_corners=True ----
  const bool dh_score = m->score_func == RBA_SCORE_DENSEHYBRID && rba_out;
  if (dry ? m->has_ood_pred : (ood_pred_out || dh_score)) {
    RBA_CHECK(m->has_ood_pred, "return_ood_pred / densehybrid score need the ood_pred head (state_dict has no "
                               "sem_seg_head.predictor.ood_pred.*; MODEL.MASK_FORMER.DENSE_HYBRID_LOSS)");
    const int64_t T = (int64_t)B * HWm;
    Planes t1 = A.planes(T * c.mask_dim);
    RBA_TRY(F.lin(ypl, D, T, D, F.P(pr + "ood_pred.fold.weight"), c.mask_dim, F.W(pr + "ood_pred.fold.bias"), RBA_ACT_RELU, nullptr,
                  nullptr, 0, t1, c.mask_dim));
    float* ol = A.f32(T * 2);
    RBA_TRY(F.lin(t1, c.mask_dim, T, c.mask_dim, F.P(pr + "ood_pred.conv.weight"), 2, F.W(pr + "ood_pred.conv.bias"), RBA_ACT_NONE,
                  nullptr, ol, 2));
    RBA_RUN(ood_pred_resize(ol, B, mH, mW, H, W, ood_pred_out, dh_score ? rba_out : nullptr, st));
  }
  if (!dry && A.overflow) return fail(RBA_ERR_STATE, "workspace overflow (reserved %zu, needed %zu)", A.cap, A.peak);
  return RBA_OK;
}

static int ensure_workspace(rba_model* m, int B, int H, int W, cudaStream_t st) {
  if (B == m->rB && H == m->rH && W == m->rW && m->arena.base) return RBA_OK;   // validated shape
  int rc = forward_impl(m, nullptr, RBA_IMG_U8, B, H, W, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, true);
  if (rc != RBA_OK) return rc;
  // + room for the optional internal rba buffer (sem_seg-only calls) and slack
  const size_t need_bytes = m->arena.peak + (size_t)B * H * W * 4 + (1 << 20);
  if (need_bytes > m->arena.cap) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (st) cudaStreamIsCapturing(st, &cs);
    RBA_CHECK(cs == cudaStreamCaptureStatusNone,
              "workspace too small during stream capture; call rba_model_reserve(batch, height, width) first");
    RBA_CUDA(cudaDeviceSynchronize());
    // A CUDA graph captured on the old arena has its pointers baked in: keep that arena alive (retired) instead of
    // freeing it under the graph; rba_model_release_retired() frees retired arenas once the caller dropped its graphs.
    if (m->arena.base) {
      if (m->arena_captured) m->retired_arenas.push_back(m->arena.base);
      else RBA_CUDA(cudaFree(m->arena.base));
    }
    m->arena_captured = false;
    m->arena_generation++;
    m->arena.base = nullptr;
    m->arena.cap = 0;
    void* p = nullptr;
    RBA_CUDA(cudaMalloc(&p, need_bytes));
    m->arena.base = (char*)p;
    m->arena.cap = need_bytes;
  }
  m->rB = B; m->rH = H; m->rW = W;
  return RBA_OK;
}

extern "C" int rba_model_reserve(rba_model* m, int batch, int height, int width) {
  RBA_CHECK(m && m->finalized, "rba_model_reserve: model not finalized");
  RBA_CHECK(batch > 0 && height > 0 && width > 0, "rba_model_reserve: bad shape");
  RBA_CUDA(cudaSetDevice(m->device));
  return ensure_workspace(m, batch, height, width, nullptr);
}

extern "C" int rba_forward_ex(rba_model* m, const void* images, int img_dtype, int B, int H, int W, const rba_outputs* out,
                              void* stream) {
  RBA_CHECK(m && m->finalized, "rba_forward: model not finalized");
  RBA_CHECK(images && out, "rba_forward: null pointer");
  RBA_CHECK(B > 0 && H > 0 && W > 0, "rba_forward: bad shape B=%d H=%d W=%d", B, H, W);
  RBA_CHECK(img_dtype == RBA_IMG_U8 || img_dtype == RBA_IMG_F32, "rba_forward: bad image dtype");
  RBA_CUDA(cudaSetDevice(m->device));
  RBA_TRY(ensure_workspace(m, B, H, W, (cudaStream_t)stream));
  if (stream) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing((cudaStream_t)stream, &cs);
    if (cs != cudaStreamCaptureStatusNone) m->arena_captured = true;
  }
  return forward_impl(m, images, img_dtype, B, H, W, out->rba, out->sem_seg, out->pred_logits, out->pred_masks, out->ood_pred,
                      (cudaStream_t)stream, false);
}

extern "C" int rba_forward(rba_model* m, const void* images, int img_dtype, int B, int H, int W, float* rba_out,
                           float* sem_seg, float* pred_logits, float* pred_masks, void* stream) {
  rba_outputs o;
  o.rba = rba_out; o.sem_seg = sem_seg; o.pred_logits = pred_logits; o.pred_masks = pred_masks; o.ood_pred = nullptr;
  return rba_forward_ex(m, images, img_dtype, B, H, W, &o, stream);
}

extern "C" int64_t rba_model_arena_generation(rba_model* m) { return m ? m->arena_generation : -1; }

extern "C" int rba_model_release_retired(rba_model* m) {
  RBA_CHECK(m, "rba_model_release_retired: null model");
  RBA_CUDA(cudaSetDevice(m->device));
  RBA_CUDA(cudaDeviceSynchronize());
  for (void* p : m->retired_arenas) cudaFree(p);
  m->retired_arenas.clear();
  return RBA_OK;
}

extern "C" int rba_model_debug_attn_mask(rba_model* m, int head, const uint8_t* force, uint8_t* dump) {
  RBA_CHECK(m, "rba_model_debug_attn_mask: null model");
  RBA_CHECK(head >= 0 && head < m->cfg.dec_layers, "rba_model_debug_attn_mask: head %d outside [0, %d)", head, m->cfg.dec_layers);
  m->am_force[head] = force;
  m->am_dump[head] = dump;
  return RBA_OK;
}

extern "C" int rba_model_get_tap(rba_model* m, const char* name, float* dst, int64_t capacity, int64_t* count, void* stream) {
  RBA_CHECK(m && name && count, "rba_model_get_tap: null pointer");
  auto it = m->taps.find(name);
  if (it == m->taps.end()) return fail(RBA_ERR_INVALID, "no tap named '%s' (enable option 'taps' before forward)", name);
  *count = it->second.n;
  if (!dst) return RBA_OK;
  RBA_CHECK(capacity >= it->second.n, "rba_model_get_tap: capacity %lld < %lld", (long long)capacity, (long long)it->second.n);
  RBA_CUDA(cudaMemcpyAsync(dst, it->second.p, (size_t)it->second.n * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return RBA_OK;
}
