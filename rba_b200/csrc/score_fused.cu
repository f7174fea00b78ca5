// Fused mask einsum + RbA score: ONE pass from the pixel decoder's feature planes to the score map
// (SURVEY §8d "Variant A"; the kernel BASELINE.json's metric names).
//
// Replaces, per image (mask2former_transformer_decoder.py:479, maskformer_model.py:294-299,381-386,
// evaluate_ood.py:148-150; the mask_features 1x1 conv of msdeformattn.py:254-260 is already folded into E'):
//   m[q,i,j]   = sum_c E'[q,c] y[i,j,c] + b'[q]                          einsum "bqc,bchw->bqhw"   (tcgen05, bf16x3)
//   u[q,Y,X]   = bilinear x4 (align_corners=False) of m                   F.interpolate             (mma.sync tf32)
//   s[k,Y,X]   = sum_q softmax(logits[q,:])[k] * sigmoid(u[q,Y,X])        semantic_inference        (mma.sync f16 hi/lo)
//   rba[Y,X]   = -sum_k tanh(s[k,Y,X])                                    get_RbA
// Neither the (Q,h,w) mask logits, nor the (Q,4h,4w) upsampled masks, nor (unless asked for) the (K,4h,4w) sem_seg
// ever reach HBM.  Algorithmic HBM bytes per image: 4*D*h*w (feature planes) + 4*Q*D + 4*Q*(K+2) + 4*H*W.
//
// Persistent kernel, one CTA per SM, 18 warps.  A tile is an 8 x 16 patch of low-resolution pixels (M = 128 rows of the
// einsum GEMM) whose 7 x 15 interior cells each own the 4 x 4 output pixels that interpolate between the cell's four
// corner taps; neighbouring tiles overlap by one low-res row/column (22 % redundant GEMM work, no halo exchange).
//   warp 0      TMA producer: feature planes through a 4-D NHWC map (out-of-bounds rows/cols zero-filled) and E' planes,
//               K blocks of 64 channels into a 2-stage mbarrier ring.
//   warp 1      MMA issuer: 3 tcgen05.mma (hi*hi, hi*lo, lo*hi) per K=16 step into a 128 x 112 fp32 TMEM accumulator.
//   warps 2-17  drain TMEM -> shared memory as patch[pixel][query] (scaled by -log2 e, bias added), then per cell:
//               interpolation as a 16x8x8 tf32 MMA (A = the 16 pixels' tap weights, exact in tf32; B = the four taps of
//               8 queries as tf32 hi|lo), sigmoid on the MUFU pipe (ex2 + rcp), the C fragment re-used in place as the A
//               fragment of the (pixel x query) x (query x class) contraction on f16 hi/lo MMAs, tanh + class sum with
//               quad shuffles.  The MMAs of tile i+1 overlap the score phase of tile i (the accumulator is drained first).
#include <cuda_fp16.h>

#include "kernels.cuh"
#include "tcgen05.cuh"

namespace rba {

constexpr int FS_NQ = 112;                                  // MMA N: queries padded to a multiple of 16
constexpr int FS_QP = 104;                                  // patch pitch (words): queries kept; = 8 (mod 32)
constexpr int FS_ROWSTRIDE = TC_CONV_TW * FS_QP + 16;       // 1680 words: = 16 (mod 32) -> the 4 taps hit 4 bank groups
constexpr int FS_BR = TC_CONV_TH - 1, FS_BC = TC_CONV_TW - 1, FS_NBLK = FS_BR * FS_BC;   // 7 x 15 cells
constexpr int FS_STAGES = 2;
constexpr int FS_A_BYTES = TC_BM * TC_BK * 2;               // 16 KB: one plane of the feature tile per K block
constexpr int FS_E_BYTES = FS_NQ * TC_BK * 2;               // 14 KB: one plane of E'
constexpr int FS_STAGE_BYTES = 2 * FS_A_BYTES + 2 * FS_E_BYTES;
constexpr int FS_PATCH_BYTES = TC_CONV_TH * FS_ROWSTRIDE * 4;
constexpr int FS_KS = 7;                                    // k16 steps in the probability layout
constexpr int FS_NT = 3;                                    // n8 class tiles (K <= 24)
constexpr int FS_P_BYTES = FS_NT * 8 * FS_KS * 4 * 16;      // [class][k16 step][tq] x {b0_hi, b1_hi, b0_lo, b1_lo}
constexpr int FS_BIAS_BYTES = 512;
#ifndef RBA_FS_CW
#define RBA_FS_CW 16
#endif
constexpr int FS_CW = RBA_FS_CW;                            // compute warps (a multiple of 4: TMEM lane quadrants)
constexpr int FS_CWQ = FS_CW / 4;                            // compute warps per TMEM lane quadrant
constexpr int FS_THREADS = (2 + FS_CW) * 32;
constexpr int FS_SMEM = FS_STAGES * FS_STAGE_BYTES + FS_PATCH_BYTES + FS_P_BYTES + FS_BIAS_BYTES + 128 + 1024;
constexpr uint32_t FS_TMEM_COLS = 128;

struct FsParams {
  const float* logits;   // (B, Q, K+1)
  const float* bias;     // (B, Q) or null
  float* rba;            // (B, H, W)
  float* sem;            // (B, K, H, W) or null
  int B, Q, K, h, w, H, W;
  int Kc;                // class columns kept: K (semantic_inference) or K+1 (semantic_inference_with_void)
  int score_func;        // RBA_SCORE_RBA: -sum_c tanh(s_c); RBA_SCORE_ENERGY: -logsumexp_c(s_c)
  int nkb;               // D / 64
  int tilesX, tilesY, ntiles;
  int debug;             // RBA_FS_DEBUG (profiling aid): 16 = skip the score phase (times TMA + einsum GEMM + drain alone)
};

__device__ __forceinline__ float fs_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fs_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// u = -x log2(e): sigmoid(x) = 1 / (1 + 2^u); 2^u -> +inf gives rcp(inf) = 0
__device__ __forceinline__ float fs_sigmoid_scaled(float u) { return fs_rcp(1.0f + fs_ex2(u)); }
// tanh(x), x >= 0: 1 - 2 / (1 + e^(2x))
__device__ __forceinline__ float fs_tanh_pos(float x) { return 1.0f - 2.0f * fs_rcp(1.0f + fs_ex2(2.8853900817779268f * x)); }

__device__ __forceinline__ uint32_t fs_pack_f16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
__device__ __forceinline__ void fs_split_f16x2(float e0, float e1, uint32_t& hi, uint32_t& lo) {
  hi = fs_pack_f16x2(e0, e1);
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = fs_pack_f16x2(e0 - f.x, e1 - f.y);
}
// Compile-time ablation (profiling aid, RBA_FS_ABL): 1 no contraction MMAs, 2 sigmoid -> FMA, 4 no f16 split, 8 no tanh
template <int ABL>
__device__ __forceinline__ float fs_sig(float u) { return (ABL & 2) ? fmaf(u, 0.01f, 0.5f) : fs_sigmoid_scaled(u); }
template <int ABL>
__device__ __forceinline__ void fs_split(float e0, float e1, uint32_t& hi, uint32_t& lo) {
  if (ABL & 4) { hi = __float_as_uint(e0); lo = __float_as_uint(e1); return; }
  fs_split_f16x2(e0, e1, hi, lo);
}
// lo = x - float(h) in ONE instruction: sm_100 mixed-precision FMA (SASS FHFMA), h = one half of a packed f16x2 register
__device__ __forceinline__ float fs_residual(float x, uint16_t h) {
  float r;
  const uint16_t m1 = 0xBC00;                               // -1.0h
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r) : "h"(h), "h"(m1), "f"(x));
  return r;
}
__device__ __forceinline__ void fs_split_fast(float e0, float e1, uint32_t& hi, uint32_t& lo) {
  hi = fs_pack_f16x2(e0, e1);
  lo = fs_pack_f16x2(fs_residual(e0, (uint16_t)(hi & 0xffffu)), fs_residual(e1, (uint16_t)(hi >> 16)));
}
// Four sigmoids 1 / (1 + 2^u) with ONE reciprocal (the MUFU pipe is the binding resource of this kernel): the four
// denominators are multiplied up, inverted once and divided back out (9 FMUL on the idle FMA pipe for 3 MUFU.RCP).
// u is clamped at 30 so the product stays below 2^121; the clamp changes a sigmoid by < 2^-30.
__device__ __forceinline__ void fs_sigmoid4(const float* u, float* s) {
  const float a0 = 1.0f + fs_ex2(fminf(u[0], 30.f)), a1 = 1.0f + fs_ex2(fminf(u[1], 30.f));
  const float a2 = 1.0f + fs_ex2(fminf(u[2], 30.f)), a3 = 1.0f + fs_ex2(fminf(u[3], 30.f));
  const float ab = a0 * a1, cd = a2 * a3;
  const float r = fs_rcp(ab * cd);
  const float rab = r * cd, rcd = r * ab;
  s[0] = rab * a1; s[1] = rab * a0; s[2] = rcd * a3; s[3] = rcd * a2;
}
// Two sigmoids with one reciprocal (u clamped at 60)
__device__ __forceinline__ void fs_sigmoid2(float u0, float u1, float& s0, float& s1) {
  const float a0 = 1.0f + fs_ex2(fminf(u0, 60.f)), a1 = 1.0f + fs_ex2(fminf(u1, 60.f));
  const float r = fs_rcp(a0 * a1);
  s0 = r * a1; s1 = r * a0;
}
// RCPM (RBA_FS_RCP): 0 one reciprocal per sigmoid, 1 per pair, 2 per quad.  Produces the f16 hi/lo A fragments of one k16 step.
template <int ABL, int RCPM>
__device__ __forceinline__ void fs_sig_frag(const float* u0, const float* u1, uint32_t* ah, uint32_t* al) {
  if (ABL != 0 || RCPM == 0) {
    fs_split<ABL>(fs_sig<ABL>(u0[0]), fs_sig<ABL>(u0[1]), ah[0], al[0]);
    fs_split<ABL>(fs_sig<ABL>(u0[2]), fs_sig<ABL>(u0[3]), ah[1], al[1]);
    fs_split<ABL>(fs_sig<ABL>(u1[0]), fs_sig<ABL>(u1[1]), ah[2], al[2]);
    fs_split<ABL>(fs_sig<ABL>(u1[2]), fs_sig<ABL>(u1[3]), ah[3], al[3]);
  } else {
    float s0[4], s1[4];
    if (RCPM == 3) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { s0[e] = fs_sigmoid_scaled(u0[e]); s1[e] = fs_sigmoid_scaled(u1[e]); }
    } else if (RCPM == 1) {
      fs_sigmoid2(u0[0], u0[1], s0[0], s0[1]); fs_sigmoid2(u0[2], u0[3], s0[2], s0[3]);
      fs_sigmoid2(u1[0], u1[1], s1[0], s1[1]); fs_sigmoid2(u1[2], u1[3], s1[2], s1[3]);
    } else {
      fs_sigmoid4(u0, s0); fs_sigmoid4(u1, s1);
    }
    fs_split_fast(s0[0], s0[1], ah[0], al[0]); fs_split_fast(s0[2], s0[3], ah[1], al[1]);
    fs_split_fast(s1[0], s1[1], ah[2], al[2]); fs_split_fast(s1[2], s1[3], ah[3], al[3]);
  }
}
__device__ __forceinline__ void fs_mma_tf32(float* d, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%4,%5}, {%6,%7}, {%8,%8,%8,%8};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "f"(0.0f));
}
__device__ __forceinline__ void fs_mma_f16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void fs_mma_f16_k8(float* c, uint32_t a0, uint32_t a1, uint32_t b0) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void fs_bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(FS_CW * 32) : "memory"); }

// Interpolated, scaled logits of 16 pixels x 8 queries: one LDS (this lane's tap of query g), tf32 hi|lo, one MMA.
__device__ __forceinline__ void fs_interp8(const float* tp, uint32_t a0, uint32_t a1, float* u) {
  const float v = *tp;
  const uint32_t hi = __float_as_uint(v) & 0xffffe000u;
  const float lo = v - __uint_as_float(hi);
  fs_mma_tf32(u, a0, a1, hi, __float_as_uint(lo));
}

// ---- unclamped batched sigmoids (kept for the profiling variants; the launch path clamps the interpolated value, see
// fs_sig_frag_u): valid only when the caller guarantees u <= FS_UMAX ----
constexpr float FS_UMAX = 30.0f;      // sigmoid(x) for -x log2(e) > 30 is < 2^-30
__device__ __forceinline__ void fs_sigmoid2u(float u0, float u1, float& s0, float& s1) {
  const float a0 = 1.0f + fs_ex2(u0), a1 = 1.0f + fs_ex2(u1);
  const float r = fs_rcp(a0 * a1);
  s0 = r * a1; s1 = r * a0;
}
__device__ __forceinline__ void fs_sigmoid4u(const float* u, float* s) {
  const float a0 = 1.0f + fs_ex2(u[0]), a1 = 1.0f + fs_ex2(u[1]);
  const float a2 = 1.0f + fs_ex2(u[2]), a3 = 1.0f + fs_ex2(u[3]);
  const float ab = a0 * a1, cd = a2 * a3;
  const float r = fs_rcp(ab * cd);
  const float rab = r * cd, rcd = r * ab;
  s[0] = rab * a1; s[1] = rab * a0; s[2] = rcd * a3; s[3] = rcd * a2;
}
// sum_i 1 / (1 + 2^v_i) over four / two values with ONE reciprocal (v clamped to FS_UMAX: 1/(1+2^30) < 1e-9)
__device__ __forceinline__ float fs_rsum4(float v0, float v1, float v2, float v3) {
  const float a0 = 1.0f + fs_ex2(fminf(v0, FS_UMAX)), a1 = 1.0f + fs_ex2(fminf(v1, FS_UMAX));
  const float a2 = 1.0f + fs_ex2(fminf(v2, FS_UMAX)), a3 = 1.0f + fs_ex2(fminf(v3, FS_UMAX));
  const float ab = a0 * a1, cd = a2 * a3;
  return fmaf(cd, a0 + a1, ab * (a2 + a3)) * fs_rcp(ab * cd);
}
__device__ __forceinline__ float fs_rsum2(float v0, float v1) {
  const float a0 = 1.0f + fs_ex2(fminf(v0, FS_UMAX)), a1 = 1.0f + fs_ex2(fminf(v1, FS_UMAX));
  return (a0 + a1) * fs_rcp(a0 * a1);
}
// A fragments (f16 hi / lo) of one k16 step from 8 + 8 interpolated logits.  The clamp that keeps the batched reciprocals
// finite is applied to the INTERPOLATED value (fs_sigmoid2 / fs_sigmoid4): clamping the taps instead (an earlier revision did,
// to save one FMNMX per sigmoid) changes the interpolated logit wherever a tap beyond the clamp sits next to a small one.
template <int ABL, int RCPM>
__device__ __forceinline__ void fs_sig_frag_u(const float* u0, const float* u1, uint32_t* ah, uint32_t* al) {
  fs_sig_frag<ABL, RCPM>(u0, u1, ah, al);
}

// per-lane constants of the score phase
struct FsLane {
  int g, tq, tr, tcn, dyA, dx;
  float wyA_in, wyB_in, wx_in;
  uint32_t a0_in, a1_in;
  const uint4* bp;
  int nfull, tail;
};

// One k16 step (16 queries) of NC cells: prefetch the interpolated logits of the NEXT step into (n0, n1), then sigmoid,
// f16 hi/lo split and the 9 contraction MMAs of the CURRENT step's logits (u0, u1).
template <int NC, int ABL, int RCPM>
__device__ __forceinline__ void fs_ks_step(const FsLane& L, const float* const* tp, uint32_t a0, uint32_t a1, int ks,
                                           float (*u0)[4], float (*u1)[4], float (*n0)[4], float (*n1)[4],
                                           float (*acc)[FS_NT][4]) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    fs_interp8(tp[c] + ks * 16 + 16, a0, a1, n0[c]);     // may run past the last query group: finite garbage, unused
    fs_interp8(tp[c] + ks * 16 + 24, a0, a1, n1[c]);
  }
  uint4 bv[FS_NT];
#pragma unroll
  for (int nt = 0; nt < FS_NT; ++nt) bv[nt] = L.bp[(size_t)nt * 8 * FS_KS * 4 + ks * 4];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    uint32_t ah[4], al[4];
    fs_sig_frag_u<ABL, RCPM>(u0[c], u1[c], ah, al);
    // consecutive MMAs target different accumulators (no back-to-back dependent HMMAs)
    if (ABL & 1) {
      acc[c][0][0] += __uint_as_float(ah[0] ^ al[1] ^ bv[0].x ^ bv[1].y ^ bv[2].z);
      acc[c][1][0] += __uint_as_float(ah[2] ^ al[3] ^ ah[1] ^ ah[3] ^ al[0] ^ al[2]);
    } else {
#pragma unroll
      for (int nt = 0; nt < FS_NT; ++nt) fs_mma_f16(acc[c][nt], ah, bv[nt].x, bv[nt].y);
#pragma unroll
      for (int nt = 0; nt < FS_NT; ++nt) fs_mma_f16(acc[c][nt], ah, bv[nt].z, bv[nt].w);
#pragma unroll
      for (int nt = 0; nt < FS_NT; ++nt) fs_mma_f16(acc[c][nt], al, bv[nt].x, bv[nt].y);
    }
  }
}

// Score of NC 4x4 output cells (cell index blk[c] inside the tile) by one warp.  NC = 2 runs two independent
// LDS -> tf32 MMA -> MUFU -> f16 split -> HMMA chains per warp (the kernel is bound by the latency of that chain at 4 warps
// per scheduler, not by any pipe) and shares the class-probability B fragments between them.
template <int NC, bool INTERIOR, bool WRITE_SEM, int ABL, int RCPM>
__device__ __forceinline__ void fs_score_cells(const FsParams& p, const FsLane& L, const float* __restrict__ sPatch,
                                               float* __restrict__ rba_b, int b, int r0, int c0, const int* blk) {
  static_assert(INTERIOR || NC == 1, "border tiles are scored one cell at a time");
  const int g = L.g, tq = L.tq, tr = L.tr, tcn = L.tcn;
  int y0[NC], x0[NC];
  const float* tp[NC];
  uint32_t a0 = L.a0_in, a1 = L.a1_in;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int br = blk[c] / FS_BC, bc = blk[c] - br * FS_BC;
    const int i = r0 + br, j = c0 + bc;                  // low-res coordinates of the cell's top-left tap
    y0[c] = 4 * i + 2; x0[c] = 4 * j + 2;                // the cell's 4x4 output pixels
    if (!INTERIOR) {
      if (i > p.h - 1 || j > p.w - 1) return;
      if (y0[c] >= p.H || x0[c] >= p.W) return;
      // tap weights: interior cells use the phase weights; at the image border the clamped source index puts
      // all the weight on the in-range tap (the out-of-range tap was zero-filled by TMA)
      float wyA = L.wyA_in, wyB = L.wyB_in, wx = L.wx_in;
      if (i < 0) wyA = wyB = tr ? 1.f : 0.f;
      if (i == p.h - 1) wyA = wyB = tr ? 0.f : 1.f;
      if (j < 0) wx = tcn ? 1.f : 0.f;
      if (j == p.w - 1) wx = tcn ? 0.f : 1.f;
      a0 = __float_as_uint(wyA * wx);
      a1 = __float_as_uint(wyB * wx);
    }
    tp[c] = sPatch + (br + tr) * FS_ROWSTRIDE + (bc + tcn) * FS_QP + g;
  }
  float acc[NC][FS_NT][4];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int nt = 0; nt < FS_NT; ++nt) acc[c][nt][0] = acc[c][nt][1] = acc[c][nt][2] = acc[c][nt][3] = 0.f;
  // software pipeline: the interpolation (LDS -> tf32 MMA) of step ks+1 is issued before the sigmoid / split /
  // contraction of step ks, so the MUFU chain of one step overlaps the MMA latency of the next
  float u0[NC][4], u1[NC][4];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    fs_interp8(tp[c], a0, a1, u0[c]);                    // queries 16ks + {2tq, 2tq+1}: rows g (u[0..1]), g+8 (u[2..3])
    fs_interp8(tp[c] + 8, a0, a1, u1[c]);                // queries 16ks + 8 + {2tq, 2tq+1}
  }
  // two k16 steps per trip with the "current" and "next" logits swapping roles (no register moves between steps)
  float n0[NC][4], n1[NC][4];
  int ks = 0;
#pragma unroll 1
  for (; ks + 1 < L.nfull; ks += 2) {
    fs_ks_step<NC, ABL, RCPM>(L, tp, a0, a1, ks, u0, u1, n0, n1, acc);
    fs_ks_step<NC, ABL, RCPM>(L, tp, a0, a1, ks + 1, n0, n1, u0, u1, acc);
  }
  if (ks < L.nfull) {
    fs_ks_step<NC, ABL, RCPM>(L, tp, a0, a1, ks, u0, u1, n0, n1, acc);
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) { u0[c][e] = n0[c][e]; u1[c][e] = n1[c][e]; }
  }
  if (L.tail) {
    uint4 bv[FS_NT];
#pragma unroll
    for (int nt = 0; nt < FS_NT; ++nt) bv[nt] = L.bp[(size_t)nt * 8 * FS_KS * 4 + L.nfull * 4];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      uint32_t ah0, al0, ah1, al1;
      if (ABL != 0 || RCPM == 0) {
        fs_split<ABL>(fs_sig<ABL>(u0[c][0]), fs_sig<ABL>(u0[c][1]), ah0, al0);
        fs_split<ABL>(fs_sig<ABL>(u0[c][2]), fs_sig<ABL>(u0[c][3]), ah1, al1);
      } else {
        float s0[4];
        if (RCPM == 3) {
#pragma unroll
          for (int e = 0; e < 4; ++e) s0[e] = fs_sigmoid_scaled(u0[c][e]);
        } else {
          fs_sigmoid4(u0[c], s0);
        }
        fs_split_fast(s0[0], s0[1], ah0, al0);
        fs_split_fast(s0[2], s0[3], ah1, al1);
      }
#pragma unroll
      for (int nt = 0; nt < FS_NT; ++nt) fs_mma_f16_k8(acc[c][nt], ah0, ah1, bv[nt].x);
#pragma unroll
      for (int nt = 0; nt < FS_NT; ++nt) fs_mma_f16_k8(acc[c][nt], ah0, ah1, bv[nt].z);
#pragma unroll
      for (int nt = 0; nt < FS_NT; ++nt) fs_mma_f16_k8(acc[c][nt], al0, al1, bv[nt].x);
    }
  }
  // ---- epilogue: acc[nt][e] = sem_seg[class 8nt+2tq+e] of pixel A (row g), acc[nt][2+e] of pixel B (row g+8).
  // sum_c tanh(s_c) = n - 2 sum_c 1/(1 + e^(2 s_c)); padded classes hold exactly 0 and contribute tanh(0) = 0 ----
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int yA = y0[c] + L.dyA, yB = yA + 2, x = x0[c] + L.dx;
    const bool okx = INTERIOR || (x >= 0 && x < p.W);
    const bool okA = INTERIOR || (okx && yA >= 0 && yA < p.H), okB = INTERIOR || (okx && yB >= 0 && yB < p.H);
    if (WRITE_SEM) {
#pragma unroll
      for (int nt = 0; nt < FS_NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int cc = nt * 8 + tq * 2 + e;
          if (cc < p.Kc) {
            if (okA) p.sem[(((size_t)b * p.Kc + cc) * p.H + yA) * p.W + x] = acc[c][nt][e];
            if (okB) p.sem[(((size_t)b * p.Kc + cc) * p.H + yB) * p.W + x] = acc[c][nt][2 + e];
          }
        }
    }
    if (p.score_func == RBA_SCORE_ENERGY) {
      // -logsumexp over the kept classes (evaluate_ood.py:152-159): max and sum reduced over the quad
      float ma = -1e30f, mb = -1e30f;
#pragma unroll
      for (int nt = 0; nt < FS_NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (nt * 8 + tq * 2 + e < p.Kc) { ma = fmaxf(ma, acc[c][nt][e]); mb = fmaxf(mb, acc[c][nt][2 + e]); }
      ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1));
      ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
      mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
      mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int nt = 0; nt < FS_NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (nt * 8 + tq * 2 + e < p.Kc) { sa += __expf(acc[c][nt][e] - ma); sb += __expf(acc[c][nt][2 + e] - mb); }
      sa += __shfl_xor_sync(0xffffffffu, sa, 1);
      sa += __shfl_xor_sync(0xffffffffu, sa, 2);
      sb += __shfl_xor_sync(0xffffffffu, sb, 1);
      sb += __shfl_xor_sync(0xffffffffu, sb, 2);
      if (tq == 0 && okA) rba_b[(size_t)yA * p.W + x] = -(ma + logf(sa));
      if (tq == 1 && okB) rba_b[(size_t)yB * p.W + x] = -(mb + logf(sb));
    } else {
      float ra = 0.f, rb = 0.f;
      if (ABL & 8) {
#pragma unroll
        for (int nt = 0; nt < FS_NT; ++nt) { ra += acc[c][nt][0] + acc[c][nt][1]; rb += acc[c][nt][2] + acc[c][nt][3]; }
      } else if (WRITE_SEM) {
#pragma unroll
        for (int nt = 0; nt < FS_NT; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            ra += fs_rcp(1.0f + fs_ex2(2.8853900817779268f * acc[c][nt][e]));
            rb += fs_rcp(1.0f + fs_ex2(2.8853900817779268f * acc[c][nt][2 + e]));
          }
      } else {
        // class sums arrive pre-scaled by 2 log2(e); six per pixel and lane: one quad + one pair, 2 reciprocals instead of 6
        ra = fs_rsum4(acc[c][0][0], acc[c][0][1], acc[c][1][0], acc[c][1][1]) + fs_rsum2(acc[c][2][0], acc[c][2][1]);
        rb = fs_rsum4(acc[c][0][2], acc[c][0][3], acc[c][1][2], acc[c][1][3]) + fs_rsum2(acc[c][2][2], acc[c][2][3]);
      }
      ra += __shfl_xor_sync(0xffffffffu, ra, 1);
      ra += __shfl_xor_sync(0xffffffffu, ra, 2);
      rb += __shfl_xor_sync(0xffffffffu, rb, 1);
      rb += __shfl_xor_sync(0xffffffffu, rb, 2);
      // rba = -sum tanh = 2 sum r - 24
      if (tq == 0 && okA) rba_b[(size_t)yA * p.W + x] = fmaf(2.0f, ra, -(float)(FS_NT * 8));
      if (tq == 1 && okB) rba_b[(size_t)yB * p.W + x] = fmaf(2.0f, rb, -(float)(FS_NT * 8));
    }
  }
}

template <bool WRITE_SEM, int ABL = 0, int RCPM = 2, int NCELL = 2>
__global__ void __launch_bounds__(FS_THREADS, 1)
rba_einsum_score_kernel(const __grid_constant__ CUtensorMap tmY_hi, const __grid_constant__ CUtensorMap tmY_lo,
                        const __grid_constant__ CUtensorMap tmE_hi, const __grid_constant__ CUtensorMap tmE_lo,
                        const FsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* sPatch = reinterpret_cast<float*>(smem + FS_STAGES * FS_STAGE_BYTES);
  uint4* sP = reinterpret_cast<uint4*>(smem + FS_STAGES * FS_STAGE_BYTES + FS_PATCH_BYTES);
  float* sBias = reinterpret_cast<float*>(smem + FS_STAGES * FS_STAGE_BYTES + FS_PATCH_BYTES + FS_P_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FS_STAGES * FS_STAGE_BYTES + FS_PATCH_BYTES + FS_P_BYTES + FS_BIAS_BYTES);
  uint64_t* full = bars;                    // [FS_STAGES]
  uint64_t* empty = bars + FS_STAGES;       // [FS_STAGES]
  uint64_t* acc_full = bars + 2 * FS_STAGES;
  uint64_t* acc_empty = bars + 2 * FS_STAGES + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * FS_STAGES + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmY_hi); prefetch_tmap(&tmY_lo); prefetch_tmap(&tmE_hi); prefetch_tmap(&tmE_lo);
    for (int s = 0; s < FS_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, FS_CW);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(FS_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const int tx = t % p.tilesX;
        const int r = t / p.tilesX;
        const int ty = r % p.tilesY, b = r / p.tilesY;
        const int r0 = FS_BR * ty - 1, c0 = FS_BC * tx - 1;
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          const int s = it % FS_STAGES;
          const uint32_t ph = (it / FS_STAGES) & 1;
          mbar_wait_sleep(&empty[s], ph ^ 1);
          uint8_t* st = smem + s * FS_STAGE_BYTES;
          mbar_expect_tx(&full[s], FS_STAGE_BYTES);
          tma_load_4d(st, &tmY_hi, &full[s], kb * TC_BK, c0, r0, b);
          tma_load_4d(st + FS_A_BYTES, &tmY_lo, &full[s], kb * TC_BK, c0, r0, b);
          tma_load_3d(st + 2 * FS_A_BYTES, &tmE_hi, &full[s], kb * TC_BK, 0, b);
          tma_load_3d(st + 2 * FS_A_BYTES + FS_E_BYTES, &tmE_lo, &full[s], kb * TC_BK, 0, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TC_BM, FS_NQ);
      uint32_t it = 0, lt = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++lt) {
        mbar_wait_sleep(acc_empty, (lt & 1) ^ 1);              // the previous tile's accumulator has been drained
        tc_fence_after();
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          const int s = it % FS_STAGES;
          const uint32_t ph = (it / FS_STAGES) & 1;
          mbar_wait_sleep(&full[s], ph);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + s * FS_STAGE_BYTES);
          const uint64_t a_hi = make_sdesc(base), a_lo = make_sdesc(base + FS_A_BYTES);
          const uint64_t e_hi = make_sdesc(base + 2 * FS_A_BYTES), e_lo = make_sdesc(base + 2 * FS_A_BYTES + FS_E_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            umma_bf16(tmem_base, a_hi + adv, e_hi + adv, idesc, (kb | k) != 0);
            umma_bf16(tmem_base, a_hi + adv, e_lo + adv, idesc, 1);
            umma_bf16(tmem_base, a_lo + adv, e_hi + adv, idesc, 1);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(acc_full);
      }
    }
  } else {
    // ===================== drain + score: warps 2..2+FS_CW-1 =====================
    const int cw = warp - 2;
    const int ctid = cw * 32 + lane;
    const int qd = warp & 3, jq = cw >> 2;                 // TMEM lane quadrant of this warp; index among its 4 warps
    const int g = lane >> 2, tq = lane & 3;
    const int tr = tq >> 1, tcn = tq & 1;                  // tap row / column select of this lane
    const float SCALE = -1.4426950408889634f;
    const int dbg = p.debug;
    const int ngroups = (p.Q + 7) >> 3, nfull = ngroups >> 1, tail = ngroups & 1;
    // interior tap weights (align_corners=False, scale 4): l1 = 1/8 + d/4 on the lower / right tap
    const int dyA = g >> 2, dx = g & 3;
    const float lyA = 0.125f + 0.25f * (float)dyA, lyB = lyA + 0.5f, lxx = 0.125f + 0.25f * (float)dx;
    const float wyA_in = tr ? lyA : 1.f - lyA, wyB_in = tr ? lyB : 1.f - lyB, wx_in = tcn ? lxx : 1.f - lxx;
    const uint32_t a0_in = __float_as_uint(wyA_in * wx_in), a1_in = __float_as_uint(wyB_in * wx_in);
    FsLane L;
    L.g = g; L.tq = tq; L.tr = tr; L.tcn = tcn; L.dyA = dyA; L.dx = dx;
    L.wyA_in = wyA_in; L.wyB_in = wyB_in; L.wx_in = wx_in; L.a0_in = a0_in; L.a1_in = a1_in;
    L.bp = sP + (size_t)g * FS_KS * 4 + tq;
    L.nfull = nfull; L.tail = tail;
    int cur_b = -1;
    uint32_t lt = 0;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++lt) {
      const int tx = t % p.tilesX;
      const int rr = t / p.tilesX;
      const int ty = rr % p.tilesY, b = rr / p.tilesY;
      const int r0 = FS_BR * ty - 1, c0 = FS_BC * tx - 1;
      if (b != cur_b) {
        // ---- per image: class probabilities as f16 hi/lo MMA B fragments, scaled bias ----
        cur_b = b;
        for (int e = ctid; e < FS_P_BYTES / 16; e += FS_CW * 32) sP[e] = make_uint4(0u, 0u, 0u, 0u);
        if (ctid < FS_NQ) sBias[ctid] = (p.bias && ctid < p.Q) ? p.bias[(size_t)b * p.Q + ctid] * SCALE : 0.f;
        fs_bar_compute();
        if (ctid < p.Q) {
          const int q = ctid;
          const float* lg = p.logits + ((size_t)b * p.Q + q) * (p.K + 1);
          float m = lg[0];
          for (int c = 1; c <= p.K; ++c) m = fmaxf(m, lg[c]);
          float ssum = 0.f;
          for (int c = 0; c <= p.K; ++c) ssum += expf(lg[c] - m);
          const float inv = 1.0f / ssum;
          const int ks = q >> 4, r = q & 15;
          const int hoff = (r >> 3) * 2 + (r & 1);         // half index inside the 16-byte entry (hi); lo = +4
          __half* base = reinterpret_cast<__half*>(sP) + ((size_t)ks * 4 + ((r & 7) >> 1)) * 8 + hoff;
          // RbA-only launches fold the 2 log2(e) of tanh(s) = 1 - 2 / (1 + 2^(2 log2(e) s)) into the probabilities, so the
          // epilogue feeds the class sums straight into ex2 (one FMUL less per class and pixel)
          const float pscale = (!WRITE_SEM && p.score_func == RBA_SCORE_RBA) ? 2.8853900817779268f : 1.0f;
          for (int c = 0; c < p.Kc; ++c) {
            const float pv = expf(lg[c] - m) * inv * pscale;
            const __half hh = __float2half_rn(pv);
            __half* d = base + (size_t)c * FS_KS * 4 * 8;
            d[0] = hh;
            d[4] = __float2half_rn(pv - __half2float(hh));
          }
        }
        fs_bar_compute();
      }
      // ---- drain the accumulator: TMEM lane = low-res pixel, column = query ----
      mbar_wait(acc_full, lt & 1);
      tc_fence_after();
      {
        const int m = qd * 32 + lane;
        float* prow = sPatch + (m >> 4) * FS_ROWSTRIDE + (m & 15) * FS_QP;
        for (int chunk = jq; chunk * 16 < FS_QP; chunk += FS_CWQ) {
          const int q0 = chunk * 16;
          uint32_t v[16];
          const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)q0;
          if (q0 + 16 <= FS_QP) {
            tmem_ld16(taddr, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b4 = *reinterpret_cast<const float4*>(sBias + q0 + 4 * j);
              float4 o;
              o.x = fmaf(__uint_as_float(v[4 * j]), SCALE, b4.x);
              o.y = fmaf(__uint_as_float(v[4 * j + 1]), SCALE, b4.y);
              o.z = fmaf(__uint_as_float(v[4 * j + 2]), SCALE, b4.z);
              o.w = fmaf(__uint_as_float(v[4 * j + 3]), SCALE, b4.w);
              *reinterpret_cast<float4*>(prow + q0 + 4 * j) = o;
            }
          } else {
            tmem_ld8(taddr, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float4 b4 = *reinterpret_cast<const float4*>(sBias + q0 + 4 * j);
              float4 o;
              o.x = fmaf(__uint_as_float(v[4 * j]), SCALE, b4.x);
              o.y = fmaf(__uint_as_float(v[4 * j + 1]), SCALE, b4.y);
              o.z = fmaf(__uint_as_float(v[4 * j + 2]), SCALE, b4.z);
              o.w = fmaf(__uint_as_float(v[4 * j + 3]), SCALE, b4.w);
              *reinterpret_cast<float4*>(prow + q0 + 4 * j) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);               // the MMAs of the next tile may start
      fs_bar_compute();                                    // patch complete

      // ---- score phase: 4x4 output cells, two per warp iteration on interior tiles (cells cw + 16 k) ----
      float* const rba_b = p.rba + (size_t)b * p.H * p.W;
      // interior tile (89 % of them at 1024 x 2048): every cell has four in-range taps and sixteen in-range output pixels,
      // so the per-cell range checks, border weights and store predicates are skipped (CTA-uniform branch)
      const bool interior = r0 >= 0 && c0 >= 0 && r0 + FS_BR <= p.h - 1 && c0 + FS_BC <= p.w - 1 &&
                            4 * (r0 + FS_BR - 1) + 5 < p.H && 4 * (c0 + FS_BC - 1) + 5 < p.W;
      if (!(dbg & 16)) {
        if (interior) {
          int blk = cw;
          if (NCELL == 2) {
            for (; blk + FS_CW < FS_NBLK; blk += 2 * FS_CW) {
              const int pair[2] = {blk, blk + FS_CW};
              fs_score_cells<2, true, WRITE_SEM, ABL, RCPM>(p, L, sPatch, rba_b, b, r0, c0, pair);
            }
          }
          for (; blk < FS_NBLK; blk += FS_CW) fs_score_cells<1, true, WRITE_SEM, ABL, RCPM>(p, L, sPatch, rba_b, b, r0, c0, &blk);
        } else {
          for (int blk = cw; blk < FS_NBLK; blk += FS_CW)
            fs_score_cells<1, false, WRITE_SEM, ABL, RCPM>(p, L, sPatch, rba_b, b, r0, c0, &blk);
        }
      }
      fs_bar_compute();                                    // patch (and, at an image change, sP) free for the next tile
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(FS_TMEM_COLS) : "memory");
  }
}

static std::atomic<int> g_fs_variant{-1};
int fused_score_variant() {
  int v = g_fs_variant.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("RBA_FS_VARIANT");
    v = e ? atoi(e) : 3;   // third generation (score_fused3.cu) by default; 1 / 2 select the earlier kernels
    g_fs_variant.store(v, std::memory_order_relaxed);
  }
  return v;
}

int einsum_score_supported(int Q, int K, int D) { return Q > 0 && Q <= FS_QP && K > 0 && K + 1 <= FS_NT * 8 && D % TC_BK == 0; }

int einsum_score_launch(const uint16_t* e_hi, const uint16_t* e_lo, const float* bias, const uint16_t* y_hi,
                        const uint16_t* y_lo, const float* logits, int B, int Q, int K, int D, int h, int w, int H, int W,
                        int score_func, int include_void, float* rba, float* sem, cudaStream_t st) {
  RBA_CHECK(einsum_score_supported(Q, K, D), "einsum_score: unsupported Q=%d (<= %d) K=%d (<= %d) D=%d (multiple of %d)", Q,
            FS_QP, K, FS_NT * 8 - 1, D, TC_BK);
  RBA_CHECK(score_func == RBA_SCORE_RBA || score_func == RBA_SCORE_ENERGY, "einsum_score: unknown score function %d", score_func);
  // RbA-only launches (the hot path) run on the third-generation kernel (score_fused3.cu: runs in registers, 1.33 ms per
  // 8 images against 1.45 here); RBA_FS_VARIANT / rba_k_set_fused_score_variant select 2 (score_fused2.cu, tcgen05 score
  // phase, 1.54 ms) or 1 (this file); sem_seg / energy launches always run here
  if (!sem && score_func == RBA_SCORE_RBA && fused_score_variant() == 3)
    return einsum_score3_launch(e_hi, e_lo, bias, y_hi, y_lo, logits, B, Q, K, D, h, w, H, W, include_void, rba, st);
  if (!sem && score_func == RBA_SCORE_RBA && fused_score_variant() == 2)
    return einsum_score2_launch(e_hi, e_lo, bias, y_hi, y_lo, logits, B, Q, K, D, h, w, H, W, include_void, rba, st);
  RBA_CHECK(((uintptr_t)e_hi & 15) == 0 && ((uintptr_t)e_lo & 15) == 0 && ((uintptr_t)y_hi & 15) == 0 && ((uintptr_t)y_lo & 15) == 0,
            "einsum_score: operand planes must be 16-byte aligned");
  FsParams p;
  memset(&p, 0, sizeof(p));
  p.logits = logits; p.bias = bias; p.rba = rba; p.sem = sem;
  p.B = B; p.Q = Q; p.K = K; p.h = h; p.w = w; p.H = H; p.W = W;
  p.Kc = include_void ? K + 1 : K;
  p.score_func = score_func;
  p.nkb = D / TC_BK;
  { const char* e = getenv("RBA_FS_DEBUG"); p.debug = e ? atoi(e) : 0; }
  p.tilesX = (int)cdiv(w + 1, FS_BC); p.tilesY = (int)cdiv(h + 1, FS_BR);
  const int64_t nt = (int64_t)B * p.tilesX * p.tilesY;
  RBA_CHECK(nt < (1LL << 31), "einsum_score: too many tiles");
  p.ntiles = (int)nt;
  CUtensorMap ty_hi, ty_lo, te_hi, te_lo;
  RBA_TRY_(make_map_nhwc(&ty_hi, y_hi, B, h, w, D));
  RBA_TRY_(make_map_nhwc(&ty_lo, y_lo, B, h, w, D));
  RBA_TRY_(make_map_3d(&te_hi, e_hi, D, Q, D, B, (int64_t)Q * D, FS_NQ));
  RBA_TRY_(make_map_3d(&te_lo, e_lo, D, Q, D, B, (int64_t)Q * D, FS_NQ));
  dim3 grid((unsigned)std::min<int64_t>(nt, num_sms()));
#define RBA_FS_LAUNCH(S, A, R, N)                                                                                          \
  do {                                                                                                                     \
    static PerDeviceOnce once;                                                                                             \
    if (once.needed()) { RBA_CUDA(cudaFuncSetAttribute(rba_einsum_score_kernel<S, A, R, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM)); once.done(); } \
    rba_einsum_score_kernel<S, A, R, N><<<grid, FS_THREADS, FS_SMEM, st>>>(ty_hi, ty_lo, te_hi, te_lo, p);                 \
  } while (0)
  if (sem) {
    RBA_FS_LAUNCH(true, 0, 2, 1);     // sem_seg written: HBM-store-bound, one cell per warp iteration keeps registers low
  } else {
    // profiling knobs: RBA_FS_ABL (ablations), RBA_FS_RCP (0 one reciprocal per sigmoid, 1 per pair [default], 2 per quad),
    // RBA_FS_NCELL (1 [default]: one cell per warp iteration, 2: two).  Measured (profiles/r2b_fused_score_ab.txt): no variant
    // moves the time by more than 5 %: the kernel sits at ~45 % of THREE co-equal floors (XU 0.66 ms, mma.sync 0.73 ms,
    // issue 0.78 ms per 8 images) with math_pipe_throttle as the top stall, so neither ILP nor fewer reciprocals help
    static const int abl = []() { const char* e = getenv("RBA_FS_ABL"); return e ? atoi(e) : 0; }();
    static const int rcpm = []() { const char* e = getenv("RBA_FS_RCP"); return e ? atoi(e) : 1; }();
    static const int ncell = []() { const char* e = getenv("RBA_FS_NCELL"); return e ? atoi(e) : 1; }();
    if (abl == 1) RBA_FS_LAUNCH(false, 1, 0, 2);
    else if (abl == 2) RBA_FS_LAUNCH(false, 2, 0, 2);
    else if (abl == 4) RBA_FS_LAUNCH(false, 4, 0, 2);
    else if (abl == 8) RBA_FS_LAUNCH(false, 8, 0, 2);
    else if (abl == 15) RBA_FS_LAUNCH(false, 15, 0, 2);
    else if (ncell == 1 && rcpm == 1) RBA_FS_LAUNCH(false, 0, 1, 1);
    else if (ncell == 1) RBA_FS_LAUNCH(false, 0, 2, 1);
    else if (rcpm == 0) RBA_FS_LAUNCH(false, 0, 0, 2);
    else if (rcpm == 1) RBA_FS_LAUNCH(false, 0, 1, 2);
    else RBA_FS_LAUNCH(false, 0, 2, 2);
  }
#undef RBA_FS_LAUNCH
  RBA_LAUNCHED();
  return RBA_OK;
}

}  // namespace rba

// mask_embed (B,Q,D) and features (B,h,w,D) as bf16 split planes; bias (B,Q) fp32 or NULL; pred_logits (B,Q,K+1);
// sem_seg (B, K or K+1, H, W) or NULL.
// test / profiling hook: 1 = first-generation kernel (mma.sync cells), 2 = second (tcgen05 score phase), 3 = third (runs in
// registers), 0 = default (RBA_FS_VARIANT, else 3)
extern "C" int rba_k_set_fused_score_variant(int v) {
  rba::g_fs_variant.store(v == 0 ? -1 : (v == 1 ? 1 : (v == 3 ? 3 : 2)), std::memory_order_relaxed);   // 0: back to the default / RBA_FS_VARIANT
  return RBA_OK;
}

extern "C" int rba_einsum_score_fused(const uint16_t* embed_hi, const uint16_t* embed_lo, const float* bias,
                                      const uint16_t* feat_hi, const uint16_t* feat_lo, const float* pred_logits, int B, int Q,
                                      int K, int D, int h, int w, int H, int W, int score_func, int include_void,
                                      float* rba_out, float* sem_seg, void* stream) {
  using namespace rba;
  if (B == 0) return RBA_OK;
  RBA_CHECK(embed_hi && embed_lo && feat_hi && feat_lo && pred_logits && rba_out, "rba_einsum_score_fused: null pointer");
  RBA_CHECK(B > 0 && h > 0 && w > 0, "rba_einsum_score_fused: bad shape B=%d h=%d w=%d", B, h, w);
  RBA_CHECK(H > 0 && W > 0 && H <= 4 * h && W <= 4 * w, "rba_einsum_score_fused: output (%d,%d) exceeds 4x(%d,%d)", H, W, h, w);
  return einsum_score_launch(embed_hi, embed_lo, bias, feat_hi, feat_lo, pred_logits, B, Q, K, D, h, w, H, W, score_func,
                             include_void, rba_out, sem_seg, (cudaStream_t)stream);
}
