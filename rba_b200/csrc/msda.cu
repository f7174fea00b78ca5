// Multi-scale deformable attention, forward sampling step.
//
// Drop-in for the reference FFI `MultiScaleDeformableAttention.ms_deform_attn_forward`
// (ops/src/vision.cpp:18-21 -> ops/src/cuda/ms_deform_attn_cuda.cu:25-84 -> kernel
// ops/src/cuda/ms_deform_im2col_cuda.cuh:242-304).  Semantics (== grid_sample bilinear / zeros /
// align_corners=False, ops/functions/ms_deform_attn_func.py:52-72):
//   out[b,q,m,:] = sum_{l,p} w[b,q,m,l,p] * bilinear(value_l[b,:,m,:], (x*W_l - 0.5, y*H_l - 0.5))
// with taps outside the map contributing zero and samples outside (-1,H)x(-1,W) skipped.
//
// Layout / mapping: value is (B,S,M,D) with D contiguous, so one thread per output element with d fastest
// makes every tap a D*4-byte coalesced gather (128 B for D=32) and the location/weight loads warp-uniform
// broadcasts.  The op is gather/latency bound (16 taps x 128 B per (q,head) at 1 level x 4 points): the grid is
// sized to cover all B*Lq*M*D elements at 256 threads/CTA, no shared memory, no reuse to exploit.
#include <cstring>

#include "common.cuh"

namespace rba {

constexpr int MSDA_MAX_LEVELS = 8;
struct MsdaLevels {
  int H[MSDA_MAX_LEVELS], W[MSDA_MAX_LEVELS], start[MSDA_MAX_LEVELS];
};

// The *_dev entry points keep spatial_shapes / level_start_index on the DEVICE, where the reference's own kernel reads
// them (ms_deform_im2col_cuda.cuh:242-250): no host copy, no synchronisation, graph-capturable.  The level table is
// then filled from global memory (L*3 uniform, L1-resident loads) instead of the kernel parameter.
__device__ __forceinline__ void load_device_levels(MsdaLevels& lv, const int64_t* __restrict__ dshapes,
                                                   const int64_t* __restrict__ dstart, int L) {
  if (dshapes == nullptr) return;
  for (int l = 0; l < L; ++l) {
    lv.H[l] = (int)__ldg(dshapes + 2 * l);
    lv.W[l] = (int)__ldg(dshapes + 2 * l + 1);
    lv.start[l] = (int)__ldg(dstart + l);
  }
}

__global__ void __launch_bounds__(256)
msda_forward_kernel(const float* __restrict__ value, MsdaLevels lv, const int64_t* __restrict__ dshapes,
                    const int64_t* __restrict__ dstart, const float* __restrict__ loc,
                    const float* __restrict__ attw, int64_t total, int S, int M, int D, int Lq, int L, int P,
                    float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  load_device_levels(lv, dshapes, dstart, L);
  const int d = (int)(idx % D);
  int64_t t = idx / D;
  const int m = (int)(t % M);
  t /= M;                                   // t = b*Lq + q
  const int64_t b = t / Lq;
  const float* vb = value + (b * S) * (int64_t)M * D + (int64_t)m * D + d;
  const int64_t vstride = (int64_t)M * D;   // between spatial positions
  const float* lp = loc + (t * M + m) * (int64_t)L * P * 2;
  const float* wp = attw + (t * M + m) * (int64_t)L * P;
  float acc = 0.f;
  for (int l = 0; l < L; ++l) {
    const int H = lv.H[l], W = lv.W[l];
    const float* vl = vb + (int64_t)lv.start[l] * vstride;
    for (int p = 0; p < P; ++p) {
      const float x = lp[(l * P + p) * 2 + 0];
      const float y = lp[(l * P + p) * 2 + 1];
      const float wgt = wp[l * P + p];
      const float h_im = y * H - 0.5f;
      const float w_im = x * W - 0.5f;
      if (h_im > -1.f && w_im > -1.f && h_im < H && w_im < W) {
        const int h0 = (int)floorf(h_im), w0 = (int)floorf(w_im);
        const float lh = h_im - h0, lw = w_im - w0;
        const float hh = 1.f - lh, hw = 1.f - lw;
        const int h1 = h0 + 1, w1 = w0 + 1;
        float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
        if (h0 >= 0 && w0 >= 0) v00 = __ldg(vl + ((int64_t)h0 * W + w0) * vstride);
        if (h0 >= 0 && w1 <= W - 1) v01 = __ldg(vl + ((int64_t)h0 * W + w1) * vstride);
        if (h1 <= H - 1 && w0 >= 0) v10 = __ldg(vl + ((int64_t)h1 * W + w0) * vstride);
        if (h1 <= H - 1 && w1 <= W - 1) v11 = __ldg(vl + ((int64_t)h1 * W + w1) * vstride);
        const float s = hh * hw * v00 + hh * lw * v01 + lh * hw * v10 + lh * lw * v11;
        acc = fmaf(wgt, s, acc);
      }
    }
  }
  out[idx] = acc;
}

// Same op, one thread per 4 channels (D % 4 == 0): 16-byte gathers, the location / weight loads shared by D/4 lanes.
__global__ void __launch_bounds__(256)
msda_forward_vec4_kernel(const float* __restrict__ value, MsdaLevels lv, const int64_t* __restrict__ dshapes,
                         const int64_t* __restrict__ dstart, const float* __restrict__ loc,
                         const float* __restrict__ attw, int64_t total, int S, int M, int D, int Lq, int L, int P,
                         float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;        // over B*Lq*M*(D/4)
  if (idx >= total) return;
  load_device_levels(lv, dshapes, dstart, L);
  const int D4 = D >> 2;
  const int d = (int)(idx % D4) * 4;
  int64_t t = idx / D4;
  const int m = (int)(t % M);
  t /= M;                                   // t = b*Lq + q
  const int64_t b = t / Lq;
  const float* vb = value + (b * S) * (int64_t)M * D + (int64_t)m * D + d;
  const int64_t vstride = (int64_t)M * D;
  const float* lp = loc + (t * M + m) * (int64_t)L * P * 2;
  const float* wp = attw + (t * M + m) * (int64_t)L * P;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = 0; l < L; ++l) {
    const int H = lv.H[l], W = lv.W[l];
    const float* vl = vb + (int64_t)lv.start[l] * vstride;
    for (int p = 0; p < P; ++p) {
      const float x = lp[(l * P + p) * 2 + 0];
      const float y = lp[(l * P + p) * 2 + 1];
      const float wgt = wp[l * P + p];
      const float h_im = y * H - 0.5f;
      const float w_im = x * W - 0.5f;
      if (h_im > -1.f && w_im > -1.f && h_im < H && w_im < W) {
        const int h0 = (int)floorf(h_im), w0 = (int)floorf(w_im);
        const float lh = h_im - h0, lw = w_im - w0;
        const float hh = 1.f - lh, hw = 1.f - lw;
        const int h1 = h0 + 1, w1 = w0 + 1;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 v00 = z, v01 = z, v10 = z, v11 = z;
        if (h0 >= 0 && w0 >= 0) v00 = __ldg(reinterpret_cast<const float4*>(vl + ((int64_t)h0 * W + w0) * vstride));
        if (h0 >= 0 && w1 <= W - 1) v01 = __ldg(reinterpret_cast<const float4*>(vl + ((int64_t)h0 * W + w1) * vstride));
        if (h1 <= H - 1 && w0 >= 0) v10 = __ldg(reinterpret_cast<const float4*>(vl + ((int64_t)h1 * W + w0) * vstride));
        if (h1 <= H - 1 && w1 <= W - 1) v11 = __ldg(reinterpret_cast<const float4*>(vl + ((int64_t)h1 * W + w1) * vstride));
        const float w00 = hh * hw, w01 = hh * lw, w10 = lh * hw, w11 = lh * lw;
        acc.x = fmaf(wgt, w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x, acc.x);
        acc.y = fmaf(wgt, w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y, acc.y);
        acc.z = fmaf(wgt, w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z, acc.z);
        acc.w = fmaf(wgt, w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w, acc.w);
      }
    }
  }
  *reinterpret_cast<float4*>(out + idx * 4) = acc;
}

// Engine variant (ops/modules/ms_deform_attn.py:102-109 fused in): takes the raw output of the merged
// sampling_offsets | attention_weights Linear, `oa` [B*Lq, M*L*P*3] = offsets (M,L,P,2) then logits (M,L,P),
// applies softmax over L*P, builds sampling locations from the encoder reference points of
// msdeformattn.py:150-162 (valid_ratios == 1: ref = ((j+0.5)/W_q, (i+0.5)/H_q) of the query's own level),
// samples, and writes the result as split planes for the output_proj GEMM.
// One thread per (b, q, head, 4 channels): the softmax over the L*P logits and the sampling locations are shared by the
// D/4 = 8 lanes of a head (the first version used one thread per channel, i.e. 32 lanes repeating that scalar work and
// 4-byte gathers; at 3 levels -- 43008 queries x 12 samples -- it took 11 ms per encoder layer), every tap is one 16-byte
// gather per lane = a contiguous 128-byte row segment per head.
template <int MAXLP>
__global__ void __launch_bounds__(256)
msda_fused_kernel(const float* __restrict__ value, MsdaLevels lv, const float* __restrict__ oa, int64_t total, int S,
                  int M, int D, int L, int P, uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;        // over B*S*M*(D/4)
  if (idx >= total) return;
  const int D4 = D >> 2;
  const int d = (int)(idx % D4) * 4;
  int64_t t = idx / D4;
  const int m = (int)(t % M);
  t /= M;                                   // b*S + q  (Lq == S for encoder self-attention)
  const int64_t b = t / S;
  const int q = (int)(t - b * S);
  int lq = 0;
  for (int l = 1; l < L; ++l)
    if (q >= lv.start[l]) lq = l;
  const int qi = (q - lv.start[lq]) / lv.W[lq], qj = (q - lv.start[lq]) % lv.W[lq];
  const float ref_x = (qj + 0.5f) / (float)lv.W[lq], ref_y = (qi + 0.5f) / (float)lv.H[lq];
  const int LP = L * P;
  const float* row = oa + t * (int64_t)(M * LP * 3);
  const float* offp = row + (int64_t)m * LP * 2;
  const float* lgp = row + (int64_t)M * LP * 2 + (int64_t)m * LP;
  float e[MAXLP];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < MAXLP; ++i)
    if (i < LP) { e[i] = lgp[i]; mx = fmaxf(mx, e[i]); }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXLP; ++i)
    if (i < LP) { e[i] = expf(e[i] - mx); sum += e[i]; }
  const float inv = 1.f / sum;
  const float* vb = value + (b * S) * (int64_t)M * D + (int64_t)m * D + d;
  const int64_t vstride = (int64_t)M * D;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < MAXLP; ++i) {
    if (i < LP) {
      const int l = i / P;
      const int H = lv.H[l], W = lv.W[l];
      const float* vl = vb + (int64_t)lv.start[l] * vstride;
      const float x = ref_x + offp[2 * i] / (float)W;
      const float y = ref_y + offp[2 * i + 1] / (float)H;
      const float h_im = y * H - 0.5f;
      const float w_im = x * W - 0.5f;
      if (h_im > -1.f && w_im > -1.f && h_im < H && w_im < W) {
        const int h0 = (int)floorf(h_im), w0 = (int)floorf(w_im);
        const float lh = h_im - h0, lw = w_im - w0;
        const float hh = 1.f - lh, hw = 1.f - lw;
        const int h1 = h0 + 1, w1 = w0 + 1;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 v00 = z, v01 = z, v10 = z, v11 = z;
        if (h0 >= 0 && w0 >= 0) v00 = __ldg(reinterpret_cast<const float4*>(vl + ((int64_t)h0 * W + w0) * vstride));
        if (h0 >= 0 && w1 <= W - 1) v01 = __ldg(reinterpret_cast<const float4*>(vl + ((int64_t)h0 * W + w1) * vstride));
        if (h1 <= H - 1 && w0 >= 0) v10 = __ldg(reinterpret_cast<const float4*>(vl + ((int64_t)h1 * W + w0) * vstride));
        if (h1 <= H - 1 && w1 <= W - 1) v11 = __ldg(reinterpret_cast<const float4*>(vl + ((int64_t)h1 * W + w1) * vstride));
        const float w00 = hh * hw, w01 = hh * lw, w10 = lh * hw, w11 = lh * lw, aw = e[i] * inv;
        // same association as the scalar kernel: s = w00 v00 + w01 v01 + w10 v10 + w11 v11; acc = fma(aw, s, acc)
        acc.x = fmaf(aw, w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x, acc.x);
        acc.y = fmaf(aw, w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y, acc.y);
        acc.z = fmaf(aw, w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z, acc.z);
        acc.w = fmaf(aw, w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w, acc.w);
      }
    }
  }
  store_split4(out_hi, out_lo, idx * 4, acc.x, acc.y, acc.z, acc.w);
}

// Second form of the fused kernel for head_dim 32 (every shipped configuration): the 8 lanes of a head SHARE the per-sample
// scalar work instead of repeating it.  `ncu --set full` of the form above at 3 levels (profiles/r2j: 43008 queries x 12
// samples) showed it issue-bound, not memory-bound: 507 M warp instructions per 2 images (~245 per lane and sample: expf,
// two float divisions, 64-bit index arithmetic and four bounds tests, all eight times per head), instruction-cache misses
// from the 16-way unrolled body, L1 hit rate 67 %, L2 13 % busy.  Here lane j of a head owns samples j and j + 8: it forms
// their softmax terms (max / sum by three xor-shuffles inside the 8-lane group), sampling location, the four CLAMPED 32-bit
// tap offsets and the four bilinear weights (zero for a tap outside the map: same sum as the skipped tap of the reference,
// ms_deform_im2col_cuda.cuh:242-304) and parks them in shared memory; after a __syncwarp every lane walks the samples with
// three broadcast LDS, four unconditional 16-byte gathers and 20 FP32 instructions each.  Same arithmetic, same association.
// lv.X[l] for a run-time l without spilling the kernel-parameter struct to local memory (constant-bank selects)
__device__ __forceinline__ int msda_sel(const int* a, int l) {
  int v = a[0];
#pragma unroll
  for (int k = 1; k < MSDA_MAX_LEVELS; ++k) v = l == k ? a[k] : v;
  return v;
}

template <int MAXLP>
__global__ void __launch_bounds__(256)
msda_fused_d32_kernel(const float* __restrict__ value, MsdaLevels lv, const float* __restrict__ oa, uint32_t ngroups, int S, int M,
                      int L, int P, uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo) {
  constexpr int D = 32;
  __shared__ int4 sOff[8][4][MAXLP];
  __shared__ float4 sWgt[8][4][MAXLP];
  __shared__ float sAw[8][4][MAXLP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane & 7, gw = lane >> 3;
  const uint32_t g0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;       // (b, q, head) of this 8-lane group
  const bool live = g0 < ngroups;
  const uint32_t g = live ? g0 : ngroups - 1;                             // dead groups shadow the last one (warp-wide shuffles)
  const uint32_t t = g / (uint32_t)M, m = g - t * (uint32_t)M;            // t = b * S + q
  const uint32_t b = t / (uint32_t)S, q = t - b * (uint32_t)S;
  int lq = 0;
#pragma unroll
  for (int l = 1; l < MSDA_MAX_LEVELS; ++l)
    if (l < L && (int)q >= lv.start[l]) lq = l;
  const int Wq = msda_sel(lv.W, lq), Hq = msda_sel(lv.H, lq), sq = msda_sel(lv.start, lq);
  const int qi = ((int)q - sq) / Wq, qj = ((int)q - sq) - qi * Wq;
  const float ref_x = (qj + 0.5f) / (float)Wq, ref_y = (qi + 0.5f) / (float)Hq;
  const int LP = L * P;
  const float* row = oa + (size_t)t * (size_t)(M * LP * 3);
  const float* offp = row + (size_t)m * LP * 2;
  const float* lgp = row + (size_t)M * LP * 2 + (size_t)m * LP;
  // ---- softmax over the L*P logits: this lane's terms are samples sub and sub + 8 ----
  const bool has0 = sub < LP, has1 = sub + 8 < LP;
  float e0 = has0 ? lgp[sub] : -INFINITY, e1 = has1 ? lgp[sub + 8] : -INFINITY;
  float mx = fmaxf(e0, e1);
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  e0 = has0 ? expf(e0 - mx) : 0.f;
  e1 = has1 ? expf(e1 - mx) : 0.f;
  float sum = e0 + e1;
  // same left-to-right order as the serial sum of the first form would need a serial loop; the 8-lane tree is used instead and
  // differs from it by fp32 rounding of the denominator only (<= 1 ulp of the attention weights)
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  // ---- per-sample tap offsets and weights ----
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = sub + 8 * k;
    if (i < LP) {
      const int l = i / P;
      const int H = msda_sel(lv.H, l), W = msda_sel(lv.W, l);
      const float x = ref_x + offp[2 * i] / (float)W;
      const float y = ref_y + offp[2 * i + 1] / (float)H;
      const float h_im = y * H - 0.5f;
      const float w_im = x * W - 0.5f;
      const bool inside = h_im > -1.f && w_im > -1.f && h_im < H && w_im < W;
      const int h0 = (int)floorf(h_im), w0 = (int)floorf(w_im);
      const float lh = h_im - h0, lw = w_im - w0;
      const float hh = 1.f - lh, hw = 1.f - lw;
      const int h1 = h0 + 1, w1 = w0 + 1;
      const bool okh0 = inside && h0 >= 0, okh1 = inside && h1 <= H - 1, okw0 = w0 >= 0, okw1 = w1 <= W - 1;
      const int h0c = min(max(h0, 0), H - 1), h1c = min(max(h1, 0), H - 1), w0c = min(max(w0, 0), W - 1), w1c = min(max(w1, 0), W - 1);
      const int base = msda_sel(lv.start, l);
      const int stride = M * D;
      sOff[warp][gw][i] = make_int4((base + h0c * W + w0c) * stride, (base + h0c * W + w1c) * stride, (base + h1c * W + w0c) * stride,
                                    (base + h1c * W + w1c) * stride);
      sWgt[warp][gw][i] = make_float4(okh0 && okw0 ? hh * hw : 0.f, okh0 && okw1 ? hh * lw : 0.f, okh1 && okw0 ? lh * hw : 0.f,
                                      okh1 && okw1 ? lh * lw : 0.f);
      sAw[warp][gw][i] = (k ? e1 : e0) * inv;
    }
  }
  __syncwarp();
  const float* vb = value + ((size_t)b * S * M + m) * D + 4 * sub;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int i = 0; i < LP; ++i) {
    const int4 o = sOff[warp][gw][i];
    const float4 w = sWgt[warp][gw][i];
    const float aw = sAw[warp][gw][i];
    const float4 v00 = __ldg(reinterpret_cast<const float4*>(vb + o.x));
    const float4 v01 = __ldg(reinterpret_cast<const float4*>(vb + o.y));
    const float4 v10 = __ldg(reinterpret_cast<const float4*>(vb + o.z));
    const float4 v11 = __ldg(reinterpret_cast<const float4*>(vb + o.w));
    acc.x = fmaf(aw, w.x * v00.x + w.y * v01.x + w.z * v10.x + w.w * v11.x, acc.x);
    acc.y = fmaf(aw, w.x * v00.y + w.y * v01.y + w.z * v10.y + w.w * v11.y, acc.y);
    acc.z = fmaf(aw, w.x * v00.z + w.y * v01.z + w.z * v10.z + w.w * v11.z, acc.z);
    acc.w = fmaf(aw, w.x * v00.w + w.y * v01.w + w.z * v10.w + w.w * v11.w, acc.w);
  }
  if (live) store_split4(out_hi, out_lo, ((int64_t)g * 8 + sub) * 4, acc.x, acc.y, acc.z, acc.w);
}

int msda_fused(const float* value, const int* Hs, const int* Ws, const float* oa, int B, int S, int M, int D, int L,
               int P, uint16_t* out_hi, uint16_t* out_lo, cudaStream_t st) {
  RBA_CHECK(L <= MSDA_MAX_LEVELS && L * P <= 16, "msda_fused: L=%d P=%d unsupported", L, P);
  RBA_CHECK(D % 4 == 0 && (((uintptr_t)value) & 15) == 0, "msda_fused: head_dim must be a multiple of 4, value 16-byte aligned");
  MsdaLevels lv;
  int start = 0;
  for (int l = 0; l < L; ++l) { lv.H[l] = Hs[l]; lv.W[l] = Ws[l]; lv.start[l] = start; start += Hs[l] * Ws[l]; }
  RBA_CHECK(start == S, "msda_fused: level sizes do not sum to S");
  const int64_t total = (int64_t)B * S * M * (D / 4);
  static const bool first_form = getenv("RBA_MSDA_FORM1") != nullptr;      // profiling: the one-thread-does-everything form
  if (D == 32 && !first_form && (int64_t)B * S * M < (1LL << 28) && (int64_t)S * M * D < (1LL << 31)) {
    msda_fused_d32_kernel<16><<<(unsigned)cdiv(total, 256), 256, 0, st>>>(value, lv, oa, (uint32_t)((int64_t)B * S * M), S, M, L, P,
                                                                            out_hi, out_lo);
    RBA_LAUNCHED();
    return RBA_OK;
  }
  msda_fused_kernel<16><<<(unsigned)cdiv(total, 256), 256, 0, st>>>(value, lv, oa, total, S, M, D, L, P, out_hi, out_lo);
  RBA_LAUNCHED();
  return RBA_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward of the sampling step: drop-in for `MultiScaleDeformableAttention.ms_deform_attn_backward`
// (ops/src/vision.cpp:20, ops/src/cuda/ms_deform_attn_cuda.cu:87-153, kernels
// ops/src/cuda/ms_deform_im2col_cuda.cuh:92-239 col2im_bilinear + :306-925 reduction variants).
//   grad_value[b, tap, m, d]      += w_tap * attn_w * grad_out[b,q,m,d]                       (atomic: taps collide)
//   grad_attn_weight[b,q,m,l,p]    = sum_d grad_out[b,q,m,d] * bilinear(value)[d]
//   grad_sampling_loc[...,0 / 1]   = W_l / H_l * sum_d attn_w * grad_out[d] * d bilinear / d w_im / d h_im
// The reference picks one of seven kernels by channel count to reduce over d in shared memory; here ONE warp owns one
// (b, q, m) row: lanes stride over d (a tap is a coalesced D*4-byte gather / red.global.add), the three per-sample sums are
// reduced with xor-shuffles and written once -- no atomics except the unavoidable grad_value scatter, any D.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
msda_backward_kernel(const float* __restrict__ value, MsdaLevels lv, const int64_t* __restrict__ dshapes,
                     const int64_t* __restrict__ dstart, const float* __restrict__ loc,
                     const float* __restrict__ attw, const float* __restrict__ grad_out, int64_t rows, int S, int M, int D,
                     int Lq, int L, int P, float* __restrict__ grad_value, float* __restrict__ grad_loc,
                     float* __restrict__ grad_attw) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // (b*Lq + q)*M + m
  if (row >= rows) return;
  load_device_levels(lv, dshapes, dstart, L);
  const int lane = threadIdx.x & 31;
  const int m = (int)(row % M);
  const int64_t t = row / M;
  const int64_t b = t / Lq;
  const int64_t vstride = (int64_t)M * D;
  const int64_t vbase = (b * S) * vstride + (int64_t)m * D;
  const float* lp = loc + row * (int64_t)L * P * 2;
  const float* wp = attw + row * (int64_t)L * P;
  const float* go = grad_out + row * (int64_t)D;
  for (int l = 0; l < L; ++l) {
    const int H = lv.H[l], W = lv.W[l];
    const int64_t lbase = vbase + (int64_t)lv.start[l] * vstride;
    for (int p = 0; p < P; ++p) {
      const float x = lp[(l * P + p) * 2 + 0];
      const float y = lp[(l * P + p) * 2 + 1];
      const float wgt = wp[l * P + p];
      const float h_im = y * H - 0.5f;
      const float w_im = x * W - 0.5f;
      float gx = 0.f, gy = 0.f, gw = 0.f;
      if (h_im > -1.f && w_im > -1.f && h_im < H && w_im < W) {
        const int h0 = (int)floorf(h_im), w0 = (int)floorf(w_im);
        const float lh = h_im - h0, lw = w_im - w0;
        const float hh = 1.f - lh, hw = 1.f - lw;
        const int h1 = h0 + 1, w1 = w0 + 1;
        const bool ok00 = h0 >= 0 && w0 >= 0, ok01 = h0 >= 0 && w1 <= W - 1;
        const bool ok10 = h1 <= H - 1 && w0 >= 0, ok11 = h1 <= H - 1 && w1 <= W - 1;
        const int64_t o00 = lbase + ((int64_t)h0 * W + w0) * vstride, o01 = o00 + vstride;
        const int64_t o10 = o00 + (int64_t)W * vstride, o11 = o10 + vstride;
        for (int d = lane; d < D; d += 32) {
          const float g = go[d];
          const float top = g * wgt;
          float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
          if (ok00) { v00 = __ldg(value + o00 + d); atomicAdd(grad_value + o00 + d, hh * hw * top); }
          if (ok01) { v01 = __ldg(value + o01 + d); atomicAdd(grad_value + o01 + d, hh * lw * top); }
          if (ok10) { v10 = __ldg(value + o10 + d); atomicAdd(grad_value + o10 + d, lh * hw * top); }
          if (ok11) { v11 = __ldg(value + o11 + d); atomicAdd(grad_value + o11 + d, lh * lw * top); }
          gw = fmaf(g, hh * hw * v00 + hh * lw * v01 + lh * hw * v10 + lh * lw * v11, gw);
          gx = fmaf(top, hh * (v01 - v00) + lh * (v11 - v10), gx);
          gy = fmaf(top, hw * (v10 - v00) + lw * (v11 - v01), gy);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gx += __shfl_xor_sync(0xffffffffu, gx, o);
        gy += __shfl_xor_sync(0xffffffffu, gy, o);
        gw += __shfl_xor_sync(0xffffffffu, gw, o);
      }
      if (lane == 0) {
        const int64_t i = row * (int64_t)L * P + l * P + p;
        grad_loc[2 * i] = gx * (float)W;
        grad_loc[2 * i + 1] = gy * (float)H;
        grad_attw[i] = gw;
      }
    }
  }
}

static int msda_levels(const char* who, const int64_t* spatial_shapes, const int64_t* level_start_index, int L, int S,
                       MsdaLevels* lv) {
  RBA_CHECK(L > 0 && L <= MSDA_MAX_LEVELS, "%s: L=%d outside 1..%d levels", who, L, MSDA_MAX_LEVELS);
  int64_t total_s = 0;
  for (int l = 0; l < L; ++l) {
    lv->H[l] = (int)spatial_shapes[2 * l];
    lv->W[l] = (int)spatial_shapes[2 * l + 1];
    lv->start[l] = (int)level_start_index[l];
    RBA_CHECK(lv->H[l] > 0 && lv->W[l] > 0, "%s: empty level %d", who, l);
    RBA_CHECK(lv->start[l] >= 0 && (int64_t)lv->start[l] + (int64_t)lv->H[l] * lv->W[l] <= S,
              "%s: level %d exceeds value length", who, l);
    total_s += (int64_t)lv->H[l] * lv->W[l];
  }
  RBA_CHECK(total_s == S, "%s: sum(H*W)=%lld != S=%d", who, (long long)total_s, S);
  return RBA_OK;
}

}  // namespace rba

static int msda_forward_impl(const char* who, const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                             bool shapes_on_device, const float* sampling_loc, const float* attn_weight, int B, int S, int M,
                             int D, int Lq, int L, int P, int im2col_step, float* out, void* stream) {
  using namespace rba;
  RBA_CHECK(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && out, "%s: null pointer", who);
  RBA_CHECK(B >= 0 && S > 0 && M > 0 && D > 0 && Lq > 0 && L > 0 && P > 0, "%s: bad shape", who);
  RBA_CHECK(L <= MSDA_MAX_LEVELS, "%s: L=%d > %d levels", who, L, MSDA_MAX_LEVELS);
  if (B == 0) return RBA_OK;
  // same precondition as the reference host function (ms_deform_attn_cuda.cu:55-57)
  RBA_CHECK(im2col_step > 0, "%s: im2col_step must be positive", who);
  const int step = B < im2col_step ? B : im2col_step;
  RBA_CHECK(B % step == 0, "batch(%d) must divide im2col_step(%d)", B, step);
  MsdaLevels lv;
  memset(&lv, 0, sizeof(lv));
  const int64_t *dsh = nullptr, *dst = nullptr;
  if (shapes_on_device) { dsh = spatial_shapes; dst = level_start_index; }
  else RBA_TRY_(msda_levels(who, spatial_shapes, level_start_index, L, S, &lv));
  if (D % 4 == 0 && ((((uintptr_t)value) | ((uintptr_t)out)) & 15) == 0) {
    const int64_t total = (int64_t)B * Lq * M * (D / 4);
    msda_forward_vec4_kernel<<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        value, lv, dsh, dst, sampling_loc, attn_weight, total, S, M, D, Lq, L, P, out);
  } else {
    const int64_t total = (int64_t)B * Lq * M * D;
    msda_forward_kernel<<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        value, lv, dsh, dst, sampling_loc, attn_weight, total, S, M, D, Lq, L, P, out);
  }
  RBA_LAUNCHED();
  return RBA_OK;
}

// spatial_shapes / level_start_index are HOST int64 arrays here (validated: sum(H*W) == S)
extern "C" int rba_msda_forward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                const float* sampling_loc, const float* attn_weight, int B, int S, int M, int D, int Lq,
                                int L, int P, int im2col_step, float* out, void* stream) {
  return msda_forward_impl("rba_msda_forward", value, spatial_shapes, level_start_index, false, sampling_loc, attn_weight, B, S,
                           M, D, Lq, L, P, im2col_step, out, stream);
}
// ... and DEVICE int64 arrays here, exactly the tensors the reference FFI receives (no host copy, no sync)
extern "C" int rba_msda_forward_dev(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                    const float* sampling_loc, const float* attn_weight, int B, int S, int M, int D, int Lq,
                                    int L, int P, int im2col_step, float* out, void* stream) {
  return msda_forward_impl("rba_msda_forward_dev", value, spatial_shapes, level_start_index, true, sampling_loc, attn_weight, B,
                           S, M, D, Lq, L, P, im2col_step, out, stream);
}

static int msda_backward_impl(const char* who, const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                              bool shapes_on_device, const float* sampling_loc, const float* attn_weight, const float* grad_output,
                              int B, int S, int M, int D, int Lq, int L, int P, int im2col_step, float* grad_value,
                              float* grad_sampling_loc, float* grad_attn_weight, void* stream) {
  using namespace rba;
  RBA_CHECK(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && grad_output && grad_value &&
                grad_sampling_loc && grad_attn_weight,
            "%s: null pointer", who);
  RBA_CHECK(B >= 0 && S > 0 && M > 0 && D > 0 && Lq > 0 && L > 0 && P > 0, "%s: bad shape", who);
  RBA_CHECK(L <= MSDA_MAX_LEVELS, "%s: L=%d > %d levels", who, L, MSDA_MAX_LEVELS);
  if (B == 0) return RBA_OK;
  RBA_CHECK(im2col_step > 0, "%s: im2col_step must be positive", who);
  const int step = B < im2col_step ? B : im2col_step;
  RBA_CHECK(B % step == 0, "batch(%d) must divide im2col_step(%d)", B, step);       // ms_deform_attn_cuda.cu:117-119
  MsdaLevels lv;
  memset(&lv, 0, sizeof(lv));
  const int64_t *dsh = nullptr, *dst = nullptr;
  if (shapes_on_device) { dsh = spatial_shapes; dst = level_start_index; }
  else RBA_TRY_(msda_levels(who, spatial_shapes, level_start_index, L, S, &lv));
  cudaStream_t st = (cudaStream_t)stream;
  RBA_CUDA(cudaMemsetAsync(grad_value, 0, (size_t)B * S * M * D * sizeof(float), st));
  const int64_t rows = (int64_t)B * Lq * M;
  RBA_CHECK(cdiv(rows, 8) < (1LL << 31), "%s: grid too large", who);
  msda_backward_kernel<<<(unsigned)cdiv(rows, 8), 256, 0, st>>>(value, lv, dsh, dst, sampling_loc, attn_weight, grad_output, rows,
                                                               S, M, D, Lq, L, P, grad_value, grad_sampling_loc, grad_attn_weight);
  RBA_LAUNCHED();
  return RBA_OK;
}

// grad_value (B,S,M,D) is zeroed here (the reference allocates it with at::zeros_like, ms_deform_attn_cuda.cu:123);
// grad_sampling_loc (B,Lq,M,L,P,2) and grad_attn_weight (B,Lq,M,L,P) are fully written.
extern "C" int rba_msda_backward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                 const float* sampling_loc, const float* attn_weight, const float* grad_output, int B, int S,
                                 int M, int D, int Lq, int L, int P, int im2col_step, float* grad_value,
                                 float* grad_sampling_loc, float* grad_attn_weight, void* stream) {
  return msda_backward_impl("rba_msda_backward", value, spatial_shapes, level_start_index, false, sampling_loc, attn_weight,
                            grad_output, B, S, M, D, Lq, L, P, im2col_step, grad_value, grad_sampling_loc, grad_attn_weight, stream);
}
extern "C" int rba_msda_backward_dev(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                     const float* sampling_loc, const float* attn_weight, const float* grad_output, int B, int S,
                                     int M, int D, int Lq, int L, int P, int im2col_step, float* grad_value,
                                     float* grad_sampling_loc, float* grad_attn_weight, void* stream) {
  return msda_backward_impl("rba_msda_backward_dev", value, spatial_shapes, level_start_index, true, sampling_loc, attn_weight,
                            grad_output, B, S, M, D, Lq, L, P, im2col_step, grad_value, grad_sampling_loc, grad_attn_weight, stream);
}
