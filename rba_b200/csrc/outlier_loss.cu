// Training-side RbA outlier loss, forward and backward.
//
// Replaces SetCriterion.outlier_loss (mask2former/modeling/criterion.py:435-553) for the configurations the reference
// ships (configs/.../*_coco_mix_finetune.yaml: OUTLIER_LOSS_TARGET nls + SCORE_NORM tanh, pebal: energy;
// OUTLIER_LOSS_FUNC squared_hinge):
//   P[b,q,c]   = softmax(pred_logits[b,q,:])[c],  c < K                      (:450)
//   S[b,q,y,x] = sigmoid(pred_masks[b,q,y,x])                                (:451)
//   L[b,c,y,x] = sum_q P[b,q,c] S[b,q,y,x]                                   (:453, einsum "bqc,bqhw->bchw")
//   s[b,y,x]   = -sum_c f(L)  (f = id / tanh / sigmoid)   or   -logsumexp_c L (:455-466)
//   u          = bilinear(s -> label size, align_corners=True)               (:474-475)
//   loss       = mean_{gt==0} relu(u - t_in)^2 [ + mean_{gt==1} relu(t_out - u)^2, all * 0.5 if any gt==1 ]   (:480-487)
// The reference materialises S (B,Q,h,w), L (B,K,h,w) and their autograd copies; here the (Q,K) contraction lives in
// registers: one thread per low-res pixel loops over the queries with the class probabilities broadcast from shared memory.
// Backward recomputes the contraction (cheaper than storing S and L), writes d pred_masks directly and reduces
// d P = S dL^T per 128-pixel tile in shared memory before ONE atomicAdd per (q, c) and tile.
#include "common.cuh"

namespace rba {

constexpr int OL_TP = 128;                     // low-res pixels per CTA (one per thread)
enum { OL_NLS_NONE = 0, OL_NLS_TANH = 1, OL_NLS_SIGMOID = 2, OL_ENERGY = 3 };

__device__ __forceinline__ float ol_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// class probabilities of image b into shared memory, [Q][KP], columns >= K zero
template <int KP>
__device__ __forceinline__ void ol_load_probs(const float* __restrict__ logits, int b, int Q, int K, float* sP) {
  for (int q = threadIdx.x; q < Q; q += blockDim.x) {
    const float* lg = logits + ((int64_t)b * Q + q) * (K + 1);
    float m = lg[0];
    for (int c = 1; c <= K; ++c) m = fmaxf(m, lg[c]);
    float sum = 0.f;
    for (int c = 0; c <= K; ++c) sum += expf(lg[c] - m);
    const float inv = 1.0f / sum;
    for (int c = 0; c < KP; ++c) sP[q * KP + c] = c < K ? expf(lg[c] - m) * inv : 0.f;
  }
}

// score of one pixel from its class sums; also d score / d L_c (in place) when GRAD
template <int KP, bool GRAD>
__device__ __forceinline__ float ol_score(float (&L)[KP], int K, int mode) {
  float s = 0.f;
  if (mode == OL_ENERGY) {
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < KP; ++c)
      if (c < K) m = fmaxf(m, L[c]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < KP; ++c)
      if (c < K) sum += expf(L[c] - m);
    s = -(m + logf(sum));
    if (GRAD) {
      const float inv = 1.0f / sum;
#pragma unroll
      for (int c = 0; c < KP; ++c) L[c] = c < K ? -expf(L[c] - m) * inv : 0.f;
    }
  } else {
#pragma unroll
    for (int c = 0; c < KP; ++c) {
      if (c < K) {
        float f, d;
        if (mode == OL_NLS_TANH) { f = tanhf(L[c]); d = 1.0f - f * f; }
        else if (mode == OL_NLS_SIGMOID) { f = ol_sigmoid(L[c]); d = f * (1.0f - f); }
        else { f = L[c]; d = 1.0f; }
        s -= f;
        if (GRAD) L[c] = -d;
      } else if (GRAD) {
        L[c] = 0.f;
      }
    }
  }
  return s;
}

template <int KP>
__global__ void __launch_bounds__(OL_TP)
ol_score_kernel(const float* __restrict__ masks, const float* __restrict__ logits, int Q, int K, int hw, int mode,
                float* __restrict__ score) {
  extern __shared__ float ol_smem[];
  float* sP = ol_smem;                                  // [Q][KP]
  const int b = blockIdx.y;
  ol_load_probs<KP>(logits, b, Q, K, sP);
  __syncthreads();
  const int pix = blockIdx.x * OL_TP + threadIdx.x;
  if (pix >= hw) return;
  float L[KP];
#pragma unroll
  for (int c = 0; c < KP; ++c) L[c] = 0.f;
  const float* mp = masks + (int64_t)b * Q * hw + pix;
  for (int q = 0; q < Q; ++q) {
    const float s = ol_sigmoid(mp[(int64_t)q * hw]);
    const float4* pr = reinterpret_cast<const float4*>(sP + q * KP);
#pragma unroll
    for (int c4 = 0; c4 < KP / 4; ++c4) {
      const float4 p4 = pr[c4];
      L[4 * c4] = fmaf(p4.x, s, L[4 * c4]); L[4 * c4 + 1] = fmaf(p4.y, s, L[4 * c4 + 1]);
      L[4 * c4 + 2] = fmaf(p4.z, s, L[4 * c4 + 2]); L[4 * c4 + 3] = fmaf(p4.w, s, L[4 * c4 + 3]);
    }
  }
  score[(int64_t)b * hw + pix] = ol_score<KP, false>(L, K, mode);
}

// Label-resolution pass: bilinear align_corners=True resize of the score, squared hinge, sums / counts (double atomics) and
// the un-normalised adjoint of the resize scattered to two low-res maps (in-distribution and outlier part separately:
// their 1/n factors are only known once every pixel has been counted).
template <typename LT>
__global__ void __launch_bounds__(256)
ol_loss_kernel(const float* __restrict__ score, const LT* __restrict__ labels, int B, int h, int w, int H, int W, float sy,
               float sx, float t_in, float t_out, double* __restrict__ sums /*[4]: sum_id, sum_ood, n_id, n_ood*/,
               float* __restrict__ g_id, float* __restrict__ g_ood) {
  double s_id = 0.0, s_ood = 0.0, n_id = 0.0, n_ood = 0.0;
  const int64_t total = (int64_t)B * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const LT lab = labels[i];
    if (lab != 0 && lab != 1) continue;
    const int X = (int)(i % W);
    const int Y = (int)((i / W) % H);
    const int64_t b = i / ((int64_t)W * H);
    const float fy = sy * (float)Y, fx = sx * (float)X;
    int y0 = (int)fy, x0 = (int)fx;
    y0 = y0 > h - 1 ? h - 1 : y0;
    x0 = x0 > w - 1 ? w - 1 : x0;
    const int y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const int64_t base = b * (int64_t)h * w;
    const int64_t i00 = base + (int64_t)y0 * w + x0, i01 = base + (int64_t)y0 * w + x1;
    const int64_t i10 = base + (int64_t)y1 * w + x0, i11 = base + (int64_t)y1 * w + x1;
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const float u = w00 * score[i00] + w01 * score[i01] + w10 * score[i10] + w11 * score[i11];
    if (lab == 0) {
      n_id += 1.0;
      const float d = u - t_in;
      if (d > 0.f) {
        s_id += (double)d * d;
        const float g = 2.0f * d;
        atomicAdd(g_id + i00, w00 * g); atomicAdd(g_id + i01, w01 * g);
        atomicAdd(g_id + i10, w10 * g); atomicAdd(g_id + i11, w11 * g);
      }
    } else {
      n_ood += 1.0;
      const float d = t_out - u;
      if (d > 0.f) {
        s_ood += (double)d * d;
        const float g = -2.0f * d;
        atomicAdd(g_ood + i00, w00 * g); atomicAdd(g_ood + i01, w01 * g);
        atomicAdd(g_ood + i10, w10 * g); atomicAdd(g_ood + i11, w11 * g);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_id += __shfl_xor_sync(0xffffffffu, s_id, o);
    s_ood += __shfl_xor_sync(0xffffffffu, s_ood, o);
    n_id += __shfl_xor_sync(0xffffffffu, n_id, o);
    n_ood += __shfl_xor_sync(0xffffffffu, n_ood, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (n_id != 0.0) { atomicAdd(sums + 0, s_id); atomicAdd(sums + 2, n_id); }
    if (n_ood != 0.0) { atomicAdd(sums + 1, s_ood); atomicAdd(sums + 3, n_ood); }
  }
}

// loss value + the low-res gradient map g = a g_id + b g_ood (criterion.py:480-487: the 0.5 only when outliers exist;
// an empty in-distribution set gives mean(empty) = NaN exactly as torch does)
__global__ void ol_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ g_id,
                                   const float* __restrict__ g_ood, int64_t n, float* __restrict__ loss, float* __restrict__ g) {
  const double n_id = sums[2], n_ood = sums[3];
  const bool any_ood = n_ood > 0.0;
  const double half = any_ood ? 0.5 : 1.0;
  const float a = (float)(half / n_id), bb = any_ood ? (float)(0.5 / n_ood) : 0.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double l = sums[0] / n_id;
    if (any_ood) l = 0.5 * (l + sums[1] / n_ood);
    *loss = (float)l;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    g[i] = a * g_id[i] + bb * g_ood[i];
}

template <int KP>
__global__ void __launch_bounds__(OL_TP)
ol_backward_kernel(const float* __restrict__ masks, const float* __restrict__ logits, const float* __restrict__ g, int Q,
                   int K, int hw, int mode, float* __restrict__ d_masks, float* __restrict__ d_probs /*(B,Q,KP), zeroed*/) {
  extern __shared__ float ol_smem[];
  float* sP = ol_smem;                                  // [Q][KP]
  float* sS = sP + Q * KP;                              // [Q][OL_TP]
  float* sD = sS + Q * OL_TP;                           // [KP][OL_TP + 1]
  const int b = blockIdx.y, t = threadIdx.x;
  ol_load_probs<KP>(logits, b, Q, K, sP);
  __syncthreads();
  const int pix = blockIdx.x * OL_TP + t;
  const bool ok = pix < hw;
  float L[KP];
#pragma unroll
  for (int c = 0; c < KP; ++c) L[c] = 0.f;
  const float* mp = masks + (int64_t)b * Q * hw + pix;
  for (int q = 0; q < Q; ++q) {
    const float s = ok ? ol_sigmoid(mp[(int64_t)q * hw]) : 0.f;
    sS[q * OL_TP + t] = s;
    const float4* pr = reinterpret_cast<const float4*>(sP + q * KP);
#pragma unroll
    for (int c4 = 0; c4 < KP / 4; ++c4) {
      const float4 p4 = pr[c4];
      L[4 * c4] = fmaf(p4.x, s, L[4 * c4]); L[4 * c4 + 1] = fmaf(p4.y, s, L[4 * c4 + 1]);
      L[4 * c4 + 2] = fmaf(p4.z, s, L[4 * c4 + 2]); L[4 * c4 + 3] = fmaf(p4.w, s, L[4 * c4 + 3]);
    }
  }
  ol_score<KP, true>(L, K, mode);                       // L[c] <- d score / d L_c
  const float gp = ok ? g[(int64_t)b * hw + pix] : 0.f;
#pragma unroll
  for (int c = 0; c < KP; ++c) {
    L[c] *= gp;                                         // d loss / d L_c
    sD[c * (OL_TP + 1) + t] = L[c];
  }
  // d pred_masks = (sum_c P[q,c] dL_c) * S (1 - S)
  if (ok) {
    float* dm = d_masks + (int64_t)b * Q * hw + pix;
    for (int q = 0; q < Q; ++q) {
      const float4* pr = reinterpret_cast<const float4*>(sP + q * KP);
      float ds = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < KP / 4; ++c4) {
        const float4 p4 = pr[c4];
        ds = fmaf(p4.x, L[4 * c4], ds); ds = fmaf(p4.y, L[4 * c4 + 1], ds);
        ds = fmaf(p4.z, L[4 * c4 + 2], ds); ds = fmaf(p4.w, L[4 * c4 + 3], ds);
      }
      const float s = sS[q * OL_TP + t];
      dm[(int64_t)q * hw] = ds * s * (1.0f - s);
    }
  }
  __syncthreads();
  // d P[q,c] += sum_pix S[q,pix] dL[c,pix]   (this tile's part)
  for (int o = t; o < Q * KP; o += OL_TP) {
    const int q = o / KP, c = o - q * KP;
    if (c >= K) continue;
    const float* sr = sS + q * OL_TP;
    const float* dr = sD + c * (OL_TP + 1);
    float acc = 0.f;
#pragma unroll 8
    for (int i = 0; i < OL_TP; ++i) acc = fmaf(sr[i], dr[i], acc);
    atomicAdd(d_probs + ((int64_t)b * Q + q) * KP + c, acc);
  }
}

// softmax backward over the K+1 logits of every (b, q) row; the void column has no upstream gradient (:450 drops it)
__global__ void ol_softmax_backward_kernel(const float* __restrict__ logits, const float* __restrict__ d_probs, int64_t rows,
                                           int K, int KP, float* __restrict__ d_logits) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* lg = logits + r * (K + 1);
  const float* dp = d_probs + r * KP;
  float m = lg[0];
  for (int c = 1; c <= K; ++c) m = fmaxf(m, lg[c]);
  float sum = 0.f;
  for (int c = 0; c <= K; ++c) sum += expf(lg[c] - m);
  const float inv = 1.0f / sum;
  float dot = 0.f;
  for (int c = 0; c < K; ++c) dot = fmaf(expf(lg[c] - m) * inv, dp[c], dot);
  for (int c = 0; c <= K; ++c) {
    const float p = expf(lg[c] - m) * inv;
    d_logits[r * (K + 1) + c] = p * ((c < K ? dp[c] : 0.f) - dot);
  }
}

template <int KP>
static int ol_run(const float* masks, const float* logits, const void* labels, int label_bytes, int B, int Q, int K, int h,
                  int w, int H, int W, int mode, float t_in, float t_out, float* loss, float* d_masks, float* d_logits,
                  float* ws_f, double* ws_d, cudaStream_t st) {
  const int hw = h * w;
  const int64_t n_lr = (int64_t)B * hw;
  float* score = ws_f;
  float* g_id = ws_f + n_lr;
  float* g_ood = ws_f + 2 * n_lr;
  float* g = ws_f + 3 * n_lr;
  float* d_probs = ws_f + 4 * n_lr;                       // (B,Q,KP)
  RBA_CUDA(cudaMemsetAsync(g_id, 0, (size_t)2 * n_lr * sizeof(float), st));
  RBA_CUDA(cudaMemsetAsync(d_probs, 0, (size_t)B * Q * KP * sizeof(float), st));
  RBA_CUDA(cudaMemsetAsync(ws_d, 0, 4 * sizeof(double), st));
  const dim3 grid((unsigned)cdiv(hw, OL_TP), (unsigned)B);
  const size_t sm_f = (size_t)Q * KP * 4;
  const size_t sm_b = sm_f + (size_t)Q * OL_TP * 4 + (size_t)KP * (OL_TP + 1) * 4;
  RBA_CHECK(sm_b <= 200 * 1024, "outlier_loss: Q=%d too large for the shared-memory tile", Q);
  static PerDeviceOnce once;
  if (once.needed()) {
    RBA_CUDA(cudaFuncSetAttribute(ol_backward_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    RBA_CUDA(cudaFuncSetAttribute(ol_score_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    once.done();
  }
  RBA_CHECK(sm_f <= 64 * 1024, "outlier_loss: Q=%d too large", Q);
  ol_score_kernel<KP><<<grid, OL_TP, sm_f, st>>>(masks, logits, Q, K, hw, mode, score);
  RBA_LAUNCHED();
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const unsigned lgrid = (unsigned)std::min<int64_t>(cdiv((int64_t)B * H * W, 256), 148 * 8);
  if (label_bytes == 1)
    ol_loss_kernel<uint8_t><<<lgrid, 256, 0, st>>>(score, (const uint8_t*)labels, B, h, w, H, W, sy, sx, t_in, t_out, ws_d, g_id, g_ood);
  else
    ol_loss_kernel<int64_t><<<lgrid, 256, 0, st>>>(score, (const int64_t*)labels, B, h, w, H, W, sy, sx, t_in, t_out, ws_d, g_id, g_ood);
  RBA_LAUNCHED();
  ol_finalize_kernel<<<(unsigned)std::min<int64_t>(cdiv(n_lr, 256), 148 * 8), 256, 0, st>>>(ws_d, g_id, g_ood, n_lr, loss, g);
  RBA_LAUNCHED();
  if (d_masks) {
    ol_backward_kernel<KP><<<grid, OL_TP, sm_b, st>>>(masks, logits, g, Q, K, hw, mode, d_masks, d_probs);
    RBA_LAUNCHED();
    ol_softmax_backward_kernel<<<(unsigned)cdiv((int64_t)B * Q, 128), 128, 0, st>>>(logits, d_probs, (int64_t)B * Q, K, KP, d_logits);
    RBA_LAUNCHED();
  }
  return RBA_OK;
}

}  // namespace rba

extern "C" int64_t rba_outlier_loss_workspace_floats(int B, int Q, int K, int h, int w) {
  const int KP = (K + 7) / 8 * 8;
  return 4 * (int64_t)B * h * w + (int64_t)B * Q * KP + 16;     // + 4 doubles (16-byte aligned tail)
}

extern "C" int rba_outlier_loss(const float* pred_masks, const float* pred_logits, const void* outlier_masks,
                                int label_dtype_bytes, int B, int Q, int K, int h, int w, int H, int W, int score_mode,
                                float inlier_upper_threshold, float outlier_lower_threshold, float* loss, float* d_pred_masks,
                                float* d_pred_logits, float* workspace, void* stream) {
  using namespace rba;
  RBA_CHECK(pred_masks && pred_logits && outlier_masks && loss && workspace, "rba_outlier_loss: null pointer");
  RBA_CHECK((d_pred_masks == nullptr) == (d_pred_logits == nullptr), "rba_outlier_loss: pass both gradient buffers or neither");
  RBA_CHECK(B > 0 && Q > 0 && K > 0 && K <= 32 && h > 0 && w > 0 && H > 0 && W > 0, "rba_outlier_loss: bad shape (K <= 32)");
  RBA_CHECK(label_dtype_bytes == 1 || label_dtype_bytes == 8, "rba_outlier_loss: labels must be uint8 or int64");
  RBA_CHECK(score_mode >= OL_NLS_NONE && score_mode <= OL_ENERGY, "rba_outlier_loss: unknown score mode %d", score_mode);
  const int KP = (K + 7) / 8 * 8;
  const int64_t nf = 4 * (int64_t)B * h * w + (int64_t)B * Q * KP;
  double* ws_d = reinterpret_cast<double*>(workspace + ((nf + 3) / 4) * 4);
  RBA_CHECK(((uintptr_t)ws_d & 7) == 0, "rba_outlier_loss: workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
#define OL_GO(KPV) return ol_run<KPV>(pred_masks, pred_logits, outlier_masks, label_dtype_bytes, B, Q, K, h, w, H, W, score_mode, \
                                      inlier_upper_threshold, outlier_lower_threshold, loss, d_pred_masks, d_pred_logits, workspace, ws_d, st)
  switch (KP) {
    case 8: OL_GO(8);
    case 16: OL_GO(16);
    case 24: OL_GO(24);
    default: OL_GO(32);
  }
#undef OL_GO
}
