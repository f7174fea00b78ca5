// Attention kernels: Swin window attention core, decoder multi-head attention core, decoder attention mask.
#include "common.cuh"

namespace rba {

// ------------------------------------------------------------------------------------------------
// Swin window attention (swin.py:145-168) for window 12x12 (N = 144 tokens), head_dim 32.
// ------------------------------------------------------------------------------------------------
// One CTA per (window, head): S = (q*scale) k^T + rel-pos bias (+ -100 shift mask), row softmax, O = P v.
// The 144x144 attention matrix lives only in registers / shared memory (the reference materialises
// nW*heads*144*144 floats per block: 314 MB/img at stage 0).
//   phase 1: 288 threads = 18 row-blocks (8 rows) x 16 col-blocks (9 cols); 72 accumulators per thread;
//            operands read as float4 along head_dim from padded smem (pitch 36 floats).
//   phase 2: softmax: the 16 threads sharing a row reduce max / sum with xor-shuffles.
//   phase 3: P (pitch 148) overwrites the Q/K region; thread -> (row, 16 of the 32 output dims).
constexpr int WA_N = 144, WA_WS = 12, WA_D = 32, WA_QP = 36, WA_PP = 148, WA_THREADS = 288;
constexpr int WA_SMEM_FLOATS = WA_N * WA_PP /*P, aliases Q|K*/ + WA_N * WA_D /*V*/ + 23 * 23 /*bias*/;
static_assert(2 * WA_N * WA_QP <= WA_N * WA_PP, "Q|K must fit under P");

__global__ void __launch_bounds__(WA_THREADS, 1)
window_attn_kernel(const float* __restrict__ qkv, const float* __restrict__ bias_table, int C, int heads, int nWh,
                   int nWw, int shift, float scale, uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo) {
  extern __shared__ __align__(16) float smem[];
  float* sQ = smem;                          // [144][36]
  float* sK = smem + WA_N * WA_QP;           // [144][36]
  float* sP = smem;                          // [144][148] (after phase 1)
  float* sV = smem + WA_N * WA_PP;           // [144][32]
  float* sB = sV + WA_N * WA_D;              // [529]
  const int tid = threadIdx.x;
  const int head = blockIdx.y;
  const int64_t win = blockIdx.x;            // b*nW + wh*nWw + ww
  const int ww = (int)(win % nWw), wh = (int)((win / nWw) % nWh);
  const float* base = qkv + win * WA_N * (int64_t)(3 * C) + head * WA_D;

  for (int e = tid; e < WA_N * 8 * 3; e += WA_THREADS) {
    int part = e / (WA_N * 8);               // 0 q, 1 k, 2 v
    int rem = e - part * (WA_N * 8);
    int r = rem >> 3, v4 = rem & 7;
    float4 t = *reinterpret_cast<const float4*>(base + (int64_t)r * 3 * C + part * C + v4 * 4);
    if (part == 0) {
      t.x *= scale; t.y *= scale; t.z *= scale; t.w *= scale;   // q = q * self.scale (swin.py:145)
      *reinterpret_cast<float4*>(sQ + r * WA_QP + v4 * 4) = t;
    } else if (part == 1) {
      *reinterpret_cast<float4*>(sK + r * WA_QP + v4 * 4) = t;
    } else {
      *reinterpret_cast<float4*>(sV + r * WA_D + v4 * 4) = t;
    }
  }
  for (int e = tid; e < 23 * 23; e += WA_THREADS) sB[e] = bias_table[(int64_t)e * heads + head];
  __syncthreads();

  const int rb = tid >> 4, cb = tid & 15;    // rows rb*8.., cols cb*9..
  float acc[8][9];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 9; ++j) acc[i][j] = 0.f;
#pragma unroll 2
  for (int d4 = 0; d4 < 8; ++d4) {
    float4 q[8], k[9];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = *reinterpret_cast<const float4*>(sQ + (rb * 8 + i) * WA_QP + d4 * 4);
#pragma unroll
    for (int j = 0; j < 9; ++j) k[j] = *reinterpret_cast<const float4*>(sK + (cb * 9 + j) * WA_QP + d4 * 4);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        acc[i][j] = fmaf(q[i].x, k[j].x, acc[i][j]);
        acc[i][j] = fmaf(q[i].y, k[j].y, acc[i][j]);
        acc[i][j] = fmaf(q[i].z, k[j].z, acc[i][j]);
        acc[i][j] = fmaf(q[i].w, k[j].w, acc[i][j]);
      }
  }
  // bias + shift mask (swin.py:148-161; mask regions of swin.py:416-440 evaluated analytically)
  const int Hp = nWh * WA_WS, Wp = nWw * WA_WS;
  int cid[9], ci_[9], cj_[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    int c = cb * 9 + j;
    ci_[j] = c / WA_WS; cj_[j] = c - ci_[j] * WA_WS;
    int hs = wh * WA_WS + ci_[j], wsx = ww * WA_WS + cj_[j];
    int rh = hs < Hp - WA_WS ? 0 : (hs < Hp - shift ? 1 : 2);
    int rw = wsx < Wp - WA_WS ? 0 : (wsx < Wp - shift ? 1 : 2);
    cid[j] = rh * 3 + rw;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = rb * 8 + i;
    int ri = r / WA_WS, rj = r - ri * WA_WS;
    int hs = wh * WA_WS + ri, wsx = ww * WA_WS + rj;
    int rh = hs < Hp - WA_WS ? 0 : (hs < Hp - shift ? 1 : 2);
    int rw = wsx < Wp - WA_WS ? 0 : (wsx < Wp - shift ? 1 : 2);
    int rid = rh * 3 + rw;
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      float s = acc[i][j] + sB[(ri - ci_[j] + WA_WS - 1) * (2 * WA_WS - 1) + (rj - cj_[j] + WA_WS - 1)];
      if (shift > 0 && rid != cid[j]) s += -100.0f;
      acc[i][j] = s;
      mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) { acc[i][j] = expf(acc[i][j] - mx); sum += acc[i][j]; }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < 9; ++j) acc[i][j] *= inv;
  }
  __syncthreads();                           // everyone is done reading Q/K
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 9; ++j) sP[(rb * 8 + i) * WA_PP + cb * 9 + j] = acc[i][j];
  __syncthreads();

  const int r = tid >> 1, dh = (tid & 1) * 16;
  float o[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) o[d] = 0.f;
  for (int j4 = 0; j4 < WA_N / 4; ++j4) {
    float4 p4 = *reinterpret_cast<const float4*>(sP + r * WA_PP + j4 * 4);
    const float pj[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float* vr = sV + (j4 * 4 + jj) * WA_D + dh;
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4) {
        float4 vv = *reinterpret_cast<const float4*>(vr + v4 * 4);
        o[v4 * 4 + 0] = fmaf(pj[jj], vv.x, o[v4 * 4 + 0]);
        o[v4 * 4 + 1] = fmaf(pj[jj], vv.y, o[v4 * 4 + 1]);
        o[v4 * 4 + 2] = fmaf(pj[jj], vv.z, o[v4 * 4 + 2]);
        o[v4 * 4 + 3] = fmaf(pj[jj], vv.w, o[v4 * 4 + 3]);
      }
    }
  }
  const int64_t orow = win * WA_N + r;
  const int64_t ob = orow * C + head * WA_D + dh;   // (attn @ v).transpose(1,2).reshape(B_, N, C), swin.py:168
#pragma unroll
  for (int v4 = 0; v4 < 4; ++v4)
    store_split4(out_hi, out_lo, ob + v4 * 4, o[v4 * 4], o[v4 * 4 + 1], o[v4 * 4 + 2], o[v4 * 4 + 3]);
}

int window_attn(const float* qkv, const float* bias_table, int B, int H, int W, int C, int heads, int ws, int shift,
                uint16_t* out_hi, uint16_t* out_lo, cudaStream_t st) {
  RBA_CHECK(qkv && bias_table && out_hi && out_lo, "window_attn: null pointer");
  RBA_CHECK(ws == WA_WS, "window_attn: only window_size 12 is built (got %d)", ws);
  RBA_CHECK(heads > 0 && C == heads * WA_D, "window_attn: head_dim must be 32 (C=%d heads=%d)", C, heads);
  RBA_CHECK(shift >= 0 && shift < ws, "window_attn: bad shift %d", shift);
  SwinGeom g = make_swin_geom(H, W, ws, shift);
  const int64_t nwin = (int64_t)B * g.nWh * g.nWw;
  if (nwin == 0) return RBA_OK;
  RBA_CHECK(nwin < (1LL << 31), "window_attn: too many windows");
  const size_t smem = WA_SMEM_FLOATS * sizeof(float);
  RBA_CUDA(cudaFuncSetAttribute(window_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)nwin, (unsigned)heads);
  window_attn_kernel<<<grid, WA_THREADS, smem, st>>>(qkv, bias_table, C, heads, g.nWh, g.nWw, shift,
                                                     1.0f / sqrtf((float)WA_D), out_hi, out_lo);
  RBA_LAUNCHED();
  return RBA_OK;
}

// ------------------------------------------------------------------------------------------------
// Decoder multi-head attention core (nn.MultiheadAttention, mask2former_transformer_decoder.py:52-53,110-113)
// ------------------------------------------------------------------------------------------------
// One warp per (b, head, query); keys are streamed in chunks of 32 (lane <-> key for the scores, lane <-> dim for
// P.V) with an online softmax, so Lk is unbounded (2048 at 1dl, 32768 for the 3-level decoder).  head_dim = 32.
__global__ void __launch_bounds__(256)
mha_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
           const uint8_t* __restrict__ mask, int B, int Lq, int Lk, int E, int heads, int64_t ldq, int64_t ldk,
           int64_t ldv, float scale, uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t total = (int64_t)B * heads * Lq;
  if (wid >= total) return;
  const int qi = (int)(wid % Lq);
  const int head = (int)((wid / Lq) % heads);
  const int b = (int)(wid / ((int64_t)Lq * heads));
  const float* qp = q + ((int64_t)b * Lq + qi) * ldq + head * 32;
  float qr[32];
#pragma unroll
  for (int d4 = 0; d4 < 8; ++d4) {
    float4 t = *reinterpret_cast<const float4*>(qp + d4 * 4);
    qr[d4 * 4] = t.x * scale; qr[d4 * 4 + 1] = t.y * scale; qr[d4 * 4 + 2] = t.z * scale; qr[d4 * 4 + 3] = t.w * scale;
  }
  const float* kb = k + (int64_t)b * Lk * ldk + head * 32;
  const float* vb = v + (int64_t)b * Lk * ldv + head * 32;
  const uint8_t* mrow = mask ? mask + ((int64_t)b * Lq + qi) * Lk : nullptr;
  float m = -INFINITY, l = 0.f, acc = 0.f;
  for (int j0 = 0; j0 < Lk; j0 += 32) {
    const int j = j0 + lane;
    float s = -INFINITY;
    if (j < Lk && !(mrow && mrow[j])) {
      const float* kr = kb + (int64_t)j * ldk;
      float dot = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < 8; ++d4) {
        float4 t = *reinterpret_cast<const float4*>(kr + d4 * 4);
        dot = fmaf(qr[d4 * 4], t.x, dot);
        dot = fmaf(qr[d4 * 4 + 1], t.y, dot);
        dot = fmaf(qr[d4 * 4 + 2], t.z, dot);
        dot = fmaf(qr[d4 * 4 + 3], t.w, dot);
      }
      s = dot;
    }
    const float cm = warp_max(s);
    if (cm == -INFINITY) continue;           // whole chunk masked
    const float mn = fmaxf(m, cm);
    const float corr = (m == -INFINITY) ? 0.f : expf(m - mn);
    const float p = (s == -INFINITY) ? 0.f : expf(s - mn);
    l = l * corr + warp_sum(p);
    acc *= corr;
    const int nk = min(32, Lk - j0);
    for (int t = 0; t < nk; ++t) {
      const float pt = __shfl_sync(0xffffffffu, p, t);
      acc = fmaf(pt, vb[(int64_t)(j0 + t) * ldv + lane], acc);
    }
    m = mn;
  }
  const float o = (l > 0.f) ? acc / l : 0.f;
  store_split1(out_hi, out_lo, ((int64_t)b * Lq + qi) * E + head * 32 + lane, o);
}

int mha(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, const uint8_t* mask, int B,
        int Lq, int Lk, int E, int heads, uint16_t* out_hi, uint16_t* out_lo, cudaStream_t st) {
  RBA_CHECK(q && k && v && out_hi && out_lo, "mha: null pointer");
  RBA_CHECK(heads > 0 && E == heads * 32, "mha: head_dim must be 32 (E=%d heads=%d)", E, heads);
  RBA_CHECK(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0, "mha: pitches must be multiples of 4");
  const int64_t total = (int64_t)B * heads * Lq;
  if (total == 0 || Lk == 0) return RBA_OK;
  mha_kernel<<<(unsigned)cdiv(total, 8), 256, 0, st>>>(q, k, v, mask, B, Lq, Lk, E, heads, ldq, ldk, ldv,
                                                      1.0f / sqrtf(32.0f), out_hi, out_lo);
  RBA_LAUNCHED();
  return RBA_OK;
}

// ------------------------------------------------------------------------------------------------
// Decoder attention mask (mask2former_transformer_decoder.py:483-486 + :433)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilin_coeff2(int o, int in, int out, int& i0, int& i1, float& l1) {
  float scale = (float)in / (float)out;
  float src = ((float)o + 0.5f) * scale - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = min(i0 + 1, in - 1);
  l1 = src - (float)i0;
}

__global__ void __launch_bounds__(256)
attn_mask_kernel(const float* __restrict__ masks, int h, int w, int th, int tw, uint8_t* __restrict__ out) {
  // one CTA per (b, q) row
  const int64_t row = blockIdx.x;
  const float* mp = masks + row * (int64_t)h * w;
  uint8_t* op = out + row * (int64_t)th * tw;
  __shared__ int s_open;
  if (threadIdx.x == 0) s_open = 0;
  __syncthreads();
  int open = 0;
  for (int e = threadIdx.x; e < th * tw; e += blockDim.x) {
    const int oy = e / tw, ox = e - oy * tw;
    int y0, y1, x0, x1; float ly, lx;
    bilin_coeff2(oy, h, th, y0, y1, ly);
    bilin_coeff2(ox, w, tw, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float val = hy * (hx * mp[y0 * w + x0] + lx * mp[y0 * w + x1]) + ly * (hx * mp[y1 * w + x0] + lx * mp[y1 * w + x1]);
    const float sg = 1.0f / (1.0f + expf(-val));
    const uint8_t blocked = sg < 0.5f ? 1 : 0;
    op[e] = blocked;
    open |= !blocked;
  }
  if (open) atomicOr(&s_open, 1);
  __syncthreads();
  if (!s_open)                                 // attn_mask[where(sum == L)] = False
    for (int e = threadIdx.x; e < th * tw; e += blockDim.x) op[e] = 0;
}

int attn_mask(const float* masks, int B, int Q, int h, int w, int th, int tw, uint8_t* out, cudaStream_t st) {
  RBA_CHECK(masks && out, "attn_mask: null pointer");
  RBA_CHECK(h > 0 && w > 0 && th > 0 && tw > 0, "attn_mask: bad size");
  if ((int64_t)B * Q == 0) return RBA_OK;
  attn_mask_kernel<<<(unsigned)((int64_t)B * Q), 256, 0, st>>>(masks, h, w, th, tw, out);
  RBA_LAUNCHED();
  return RBA_OK;
}

}  // namespace rba

extern "C" int rba_k_window_attn(const float* qkv, const float* bias_table, int B, int H, int W, int C, int heads, int ws,
                                 int shift, uint16_t* out_hi, uint16_t* out_lo, void* stream) {
  return rba::window_attn(qkv, bias_table, B, H, W, C, heads, ws, shift, out_hi, out_lo, (cudaStream_t)stream);
}

extern "C" int rba_k_mha(const float* q, const float* k, const float* v, const uint8_t* mask, int B, int Lq, int Lk, int E,
                         int heads, uint16_t* out_hi, uint16_t* out_lo, void* stream) {
  return rba::mha(q, E, k, E, v, E, mask, B, Lq, Lk, E, heads, out_hi, out_lo, (cudaStream_t)stream);
}

extern "C" int rba_k_attn_mask(const float* masks, int B, int Q, int h, int w, int th, int tw, uint8_t* out, void* stream) {
  return rba::attn_mask(masks, B, Q, h, w, th, tw, out, (cudaStream_t)stream);
}
