// Fused mask einsum + RbA score, third generation (SURVEY §8d "Variant A"; the RbA-only launch of the kernel BASELINE.json's
// metric names).
//
// Per image (mask2former_transformer_decoder.py:479, maskformer_model.py:294-299,381-386, evaluate_ood.py:148-150):
//   m[q,i,j]   = sum_c E'[q,c] y[i,j,c] + b'[q]                          einsum "bqc,bchw->bqhw"   (tcgen05, bf16x3)
//   u[q,Y,X]   = bilinear x4 (align_corners=False) of m                   F.interpolate             (in the thread, fp32)
//   s[k,Y,X]   = sum_q softmax(logits[q,:])[k] * sigmoid(u[q,Y,X])        semantic_inference        (mma.sync f16 hi/lo)
//   rba[Y,X]   = -sum_k tanh(s[k,Y,X])                                    get_RbA
//
// The first generation (score_fused.cu) spends 15.5 thread-instructions and 1.5 MUFU per sigmoid (interpolation as tf32 MMAs,
// one cell of 16 pixels per warp); the second (score_fused2.cu) brought that to ~10 instructions / 1 MUFU with the RUN form
// below but pays ~40 % of its time synchronising the sigmoids through tensor memory.  This one keeps the run form and keeps the
// contraction in registers:
//   * a thread owns a RUN -- the 4 horizontally adjacent output pixels of one output row of one interpolation cell -- and, per
//     k16 step, 4 queries.  Along a run the interpolated logit is linear, u_j = x0 + j d, so 2^u_j = 2^x0 (2^d)^j: two ex2 and
//     two paired reciprocals give the four sigmoids (1 MUFU each); x0 and d cost 7 FP32 instructions per (run, query).
//   * the rows of the m16n8k16 MMA tiles are assigned so that a thread's fragment registers ARE its run: lane (g, t) of a warp
//     holds rows g and g + 8 of two M tiles = the 4 pixels of run g, and columns {2t, 2t+1, 2t+8, 2t+9} = queries 4t .. 4t+3 of
//     the step (the class-probability B fragments are laid out to match).  A warp unit is 8 runs = 2 cells x 4 output rows;
//     sigmoid -> f16 hi/lo split -> 18 HMMA per step, nothing passes through shared or tensor memory.
//   Range: the product form needs every tap |u| <= 60 (|mask logit| <= 41.6); the drain records the tile's max |u| and a tile
//   beyond it takes the exact path (four ex2 of individually clamped u_j) -- no clamp ever touches a tap.
//   Image borders: out-of-range taps are replicated from the edge in the patch (= align_corners=False's clamped source index).
//
// Persistent kernel, one CTA per SM, 18 warps at 96 registers (with 18+ warps of equal size five warps share a sub-partition's
// 16 K registers; `setmaxnreg` can only move registers that the CTA's own warps release).  Tile = 7 x 17 low-resolution pixels
// (a 17 x 7 TMA box = 119 of the M = 128 rows of the einsum) = 6 x 16 cells = 48 units, a multiple of the 8 (or 16) warps that
// share them: no warp waits at the tile barrier for a mostly empty last round (an 8 x 16 tile has 105 cells = 52.5 units; with 16
// warps that cost 17 % of the warp cycles).  The queries beyond the last full k16 step (Q = 100: 4) take a k8 step with one (or
// two) queries per lane instead of a seventh full step that would compute 12 padded queries.
//   warp 0      TMA producer (feature planes NHWC + E' planes, K blocks of 32 channels, SWIZZLE_64B ring) and issuer of the einsum
//               MMAs of the tiles to come (they run under the score phases)
//   warp 1      tensor-memory allocation only
//   warps 2-17  two GROUPS of 8 compute warps (RBA_F3_GROUPS=1: one group of 16).  The CTA's tiles alternate between the groups;
//               each group has its own tensor-memory accumulator, patch, class-probability tables and named barrier, drains D1
//               -> fp32 patch[pixel][query] (scaled by -log2 e, bias added) and scores the 48 units of its tile, 6 per warp.
//               While one group drains / waits at its tile barrier (7 % of a tile) the other keeps the schedulers busy:
//               1.329 -> 1.303 ms per 8 images.
#include <cuda_fp16.h>

#include "kernels.cuh"
#include "tcgen05.cuh"

namespace rba {

#ifndef RBA_F3_PROFILE
#define RBA_F3_PROFILE 0                                     // 1: also build the ablation variants of the unit (RBA_FS_DEBUG bits 4, 8)
#endif
constexpr int F3_NQ = 112;                                  // einsum N: queries padded to a multiple of 16
constexpr int F3_BK = 32;
#ifndef RBA_F3_GROUPS
#define RBA_F3_GROUPS 2                                      // 1: one group of 16 compute warps; 2: two groups of 8 on alternating tiles
#endif
constexpr int F3_NG = RBA_F3_GROUPS;
constexpr int F3_STAGES = F3_NG == 2 ? 2 : 4;
constexpr int F3_A_BYTES = TC_BM * F3_BK * 2;               // 8 KB: one plane of the feature tile per K block (119 rows loaded)
constexpr int F3_E_BYTES = F3_NQ * F3_BK * 2;               // 7 KB: one plane of E'
constexpr int F3_STAGE_BYTES = 2 * F3_A_BYTES + 2 * F3_E_BYTES;   // 30 KB
constexpr int F3_PITCH = 132;                               // patch pitch (words): = 4 (mod 32) -> 8 lanes x 16 B hit 8 bank groups
constexpr int F3_PATCH_ROWS = 120;                           // 119 pixels + the tap prefetch past the last one
constexpr int F3_PATCH_BYTES = F3_PATCH_ROWS * F3_PITCH * 4;  // 62 KB per group
constexpr int F3_KS = F3_NQ / 16;                           // 7 k16 steps
constexpr int F3_NT = 3;                                    // n8 class tiles (K + 1 <= 24)
constexpr int F3_P_BYTES = F3_NT * F3_KS * 32 * 16;         // [class tile][k16 step][lane] x {b0_hi, b1_hi, b0_lo, b1_lo}
constexpr int F3_CW = 16;                                   // compute warps
constexpr int F3_GW = F3_CW / F3_NG;                         // compute warps per group
static_assert(F3_NG == 1 || F3_NG == 2, "one or two groups");
constexpr int F3_SW = 2;                                    // service warps: TMA + einsum issue; tensor-memory allocation
constexpr int F3_THREADS = (F3_SW + F3_CW) * 32;
// per group: patch | sP | sPt | bias
constexpr int F3_OFF_PATCH = F3_STAGES * F3_STAGE_BYTES;
constexpr int F3_PT_BYTES = F3_NT * 32 * 8;                  // tail step: [class tile][lane] x {hi pair, lo pair}
constexpr int F3_GOFF_P = F3_PATCH_BYTES, F3_GOFF_PT = F3_GOFF_P + F3_P_BYTES, F3_GOFF_BIAS = F3_GOFF_PT + F3_PT_BYTES;
constexpr int F3_GROUP_BYTES = F3_GOFF_BIAS + 512;
constexpr int F3_OFF_BARS = F3_OFF_PATCH + F3_NG * F3_GROUP_BYTES;
constexpr int F3_SMEM = F3_OFF_BARS + 256 + 1024;
constexpr uint32_t F3_TMEM_COLS = F3_NG == 2 ? 256 : 128;     // one 112-column accumulator per group (columns 0 / 128)
static_assert(F3_SMEM <= 227 * 1024, "shared memory");
constexpr int F3_TW = 17, F3_TH = 7;                        // tile = 7 x 17 low-resolution pixels (119 of the 128 einsum rows)
constexpr int F3_CELLS_X = F3_TW - 1, F3_CELLS_Y = F3_TH - 1;   // 16 x 6 cells
constexpr int F3_NCELL = F3_CELLS_X * F3_CELLS_Y;           // 96
constexpr int F3_NUNIT = F3_NCELL / 2;                      // 48 units of two cells = 3 per compute warp, no remainder
static_assert(F3_TW * F3_TH <= TC_BM && F3_TW * F3_TH < F3_PATCH_ROWS && F3_NCELL % 2 == 0 && F3_NUNIT % F3_GW == 0, "tile geometry");
constexpr int F3_A_TX = F3_TW * F3_TH * F3_BK * 2;             // bytes one feature-plane box delivers
constexpr int F3_STAGE_TX = 2 * F3_A_TX + 2 * F3_E_BYTES;
constexpr float F3_UFAST = 60.0f;                           // product form valid while every tap |u| <= 60

struct F3Params {
  const float* logits;   // (B, Q, K+1)
  const float* bias;     // (B, Q) or null
  float* rba;            // (B, H, W)
  int B, Q, K, h, w, H, W;
  int Kc;                // class columns kept: K or K+1
  int nkb;               // D / 32
  int tilesX, tilesY, ntiles;
  long long* tl;         // profiling aid (RBA_FS_TIMELINE): clock64 stamps of CTA 0, warp 2: [tile < 32][8 events]
  int debug;             // RBA_FS_DEBUG (profiling aid): 1 = always the exact path, 2 = no score phase, 4 = no sigmoid math, 8 = no MMAs,
                         // 16 = no epilogue math (4 and 8 need a build with -DRBA_F3_PROFILE=1)
};

#define F3_STAMP(tile, ev)                                                                          \
  do {                                                                                              \
    if (p.tl && blockIdx.x == 0 && cw == 0 && lane == 0 && (tile) < 32) p.tl[(tile) * 8 + (ev)] = clock64(); \
  } while (0)

struct F3Bars {
  uint64_t full[F3_STAGES], empty[F3_STAGES];
  uint64_t acc_full[2], acc_empty[2];   // per group: D1 of a tile complete (tcgen05.commit) / drained by the group's warps
  uint32_t tmem_slot;
  uint32_t amax[2][2];                  // max |u| of the tile's taps (float bits), by group and tile parity
};

__device__ __forceinline__ float f3_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float f3_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// two values -> packed f16x2 hi word and f16x2 residual word (element a in the low half); the residual is one FHFMA each
__device__ __forceinline__ void f3_split_f16(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  float ra, rb;
  const uint16_t m1 = 0xBC00;                               // -1.0h
  asm("{.reg .b16 l, h; mov.b32 {l, h}, %2; fma.rn.f32.f16 %0, l, %3, %4; fma.rn.f32.f16 %1, h, %3, %5;}"
      : "=f"(ra), "=f"(rb)
      : "r"(hi), "h"(m1), "f"(a), "f"(b));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}
__device__ __forceinline__ void f3_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void f3_mma_k8(float* c, uint32_t a0, uint32_t a1, uint32_t b0) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
// barrier of one group's compute warps (named barrier 1 + group)
__device__ __forceinline__ void f3_bar_group(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(F3_GW * 32) : "memory"); }

// The four sigmoids 1 / (1 + 2^(x0 + j d)), j = 0..3, of one run and one query.
//   FAST: 2^x0 and 2^d once, the other three exponentials by multiplication (every intermediate is 2^(u_j) or (2^d)^j with
//         |u_j| <= 60, |d| <= 30: no overflow, no underflow that matters); 4 MUFU per 4 sigmoids.
//   else: four ex2 of the individually formed and clamped u_j (exact for any tap magnitude); 6 MUFU.
template <bool FAST>
__device__ __forceinline__ void f3_sig4(float x0, float d, float* s) {
  float a0, a1, a2, a3;
  if (FAST) {
    const float E = f3_ex2(x0), R = f3_ex2(d);
    const float E1 = E * R, R2 = R * R;
    a0 = 1.0f + E; a1 = 1.0f + E1; a2 = fmaf(E, R2, 1.0f); a3 = fmaf(E1, R2, 1.0f);
  } else {
    a0 = 1.0f + f3_ex2(fminf(x0, F3_UFAST));
    a1 = 1.0f + f3_ex2(fminf(x0 + d, F3_UFAST));
    a2 = 1.0f + f3_ex2(fminf(fmaf(2.0f, d, x0), F3_UFAST));
    a3 = 1.0f + f3_ex2(fminf(fmaf(3.0f, d, x0), F3_UFAST));
  }
  const float r01 = f3_rcp(a0 * a1), r23 = f3_rcp(a2 * a3);
  s[0] = r01 * a1; s[1] = r01 * a0; s[2] = r23 * a3; s[3] = r23 * a2;
}
// sum_i 1 / (1 + 2^v_i) over four / two values with ONE reciprocal (class sums, v <= ~3 in practice; the clamp keeps the product finite)
__device__ __forceinline__ float f3_rsum4(float v0, float v1, float v2, float v3) {
  const float a0 = 1.0f + f3_ex2(fminf(v0, 30.f)), a1 = 1.0f + f3_ex2(fminf(v1, 30.f));
  const float a2 = 1.0f + f3_ex2(fminf(v2, 30.f)), a3 = 1.0f + f3_ex2(fminf(v3, 30.f));
  const float ab = a0 * a1, cd = a2 * a3;
  return fmaf(cd, a0 + a1, ab * (a2 + a3)) * f3_rcp(ab * cd);
}
__device__ __forceinline__ float f3_rsum2(float v0, float v1) {
  const float a0 = 1.0f + f3_ex2(fminf(v0, 30.f)), a1 = 1.0f + f3_ex2(fminf(v1, 30.f));
  return (a0 + a1) * f3_rcp(a0 * a1);
}

// x0 and d of one (run, query): taps a b / c d, L = ly1 a + ly c, R = ly1 b + ly d, d = (R - L) / 4, x0 = L + (R - L) / 8
// (k1 = ly1 / 4, k2 = ly / 4)
__device__ __forceinline__ void f3_line(float ta, float tb, float tc, float td, float ly, float ly1, float k1, float k2, float& x0,
                                        float& d) {
  d = fmaf(k2, td - tc, k1 * (tb - ta));
  x0 = fmaf(0.5f, d, fmaf(ly, tc, ly1 * ta));
}

// denominators 1 + 2^(x0 + j d), j = 0..3
template <bool FAST>
__device__ __forceinline__ void f3_den4(float x0, float d, float* a) {
  if (FAST) {
    const float E = f3_ex2(x0), R = f3_ex2(d);
    const float E1 = E * R, R2 = R * R;
    a[0] = 1.0f + E; a[1] = 1.0f + E1; a[2] = fmaf(E, R2, 1.0f); a[3] = fmaf(E1, R2, 1.0f);
  } else {
    a[0] = 1.0f + f3_ex2(fminf(x0, F3_UFAST));
    a[1] = 1.0f + f3_ex2(fminf(x0 + d, F3_UFAST));
    a[2] = 1.0f + f3_ex2(fminf(fmaf(2.0f, d, x0), F3_UFAST));
    a[3] = 1.0f + f3_ex2(fminf(fmaf(3.0f, d, x0), F3_UFAST));
  }
}
// the 9 MMAs of one M tile; consecutive MMAs target different accumulators
template <int ABL>
__device__ __forceinline__ void f3_mma9(const uint32_t* ah, const uint32_t* al, const uint4* bv, float (*acc)[4]) {
  if (ABL & 2) {   // profiling aid: no MMAs, the fragments stay live
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e & 1][e] += __uint_as_float((ah[e] ^ al[e] ^ bv[e % F3_NT].x ^ bv[(e + 1) % F3_NT].w) & 0x3fffffffu);
    return;
  }
#pragma unroll
  for (int nt = 0; nt < F3_NT; ++nt) f3_mma(acc[nt], ah, bv[nt].x, bv[nt].y);
#pragma unroll
  for (int nt = 0; nt < F3_NT; ++nt) f3_mma(acc[nt], ah, bv[nt].z, bv[nt].w);
#pragma unroll
  for (int nt = 0; nt < F3_NT; ++nt) f3_mma(acc[nt], al, bv[nt].x, bv[nt].y);
}

// One k16 step of a unit: this lane's 4 queries x 4 pixels -> A fragments of the two M tiles -> 18 MMAs.
//   M tile mt holds pixels 2 mt (row g) and 2 mt + 1 (row g + 8) of run g; fragment register order a0 (row g, k 2t..), a1 (row
//   g + 8, k 2t..), a2 (row g, k 2t + 8..), a3 (row g + 8, k 2t + 8..) with k 2t, 2t+1 = queries 4t, 4t+1 and k 2t+8, 2t+9 =
//   queries 4t+2, 4t+3 of the step.
// Order: the four denominators of every query, then pixels 0, 1 (reciprocal of a0 a1, f16 split, the 9 MMAs of M tile 0), then
// pixels 2, 3 and M tile 1 -- the MMAs of tile 0 are in flight under the reciprocal / split work of tile 1, and a warp holds the
// tensor pipe for 9, not 18, back-to-back MMAs (math_pipe_throttle on the HMMAs was the top stall of the all-at-the-end order).
template <bool FAST, int ABL>
__device__ __forceinline__ void f3_step(const float4& a4, const float4& b4, const float4& c4, const float4& d4, const uint4* bv,
                                        float ly, float ly1, float k1, float k2, float (*acc)[F3_NT][4]) {
  const float ta[4] = {a4.x, a4.y, a4.z, a4.w}, tb[4] = {b4.x, b4.y, b4.z, b4.w};
  const float tc[4] = {c4.x, c4.y, c4.z, c4.w}, td[4] = {d4.x, d4.y, d4.z, d4.w};
  float den[4][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float x0, d0;
    f3_line(ta[q], tb[q], tc[q], td[q], ly, ly1, k1, k2, x0, d0);
    if (ABL & 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) den[q][j] = fmaf((float)j, d0, x0);
    } else {
      f3_den4<FAST>(x0, d0, den[q]);
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    float s0[4], s1[4];                                   // pixels 2 mt, 2 mt + 1 of the four queries
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (ABL & 1) {
        s0[q] = den[q][2 * mt]; s1[q] = den[q][2 * mt + 1];
      } else {
        const float r = f3_rcp(den[q][2 * mt] * den[q][2 * mt + 1]);
        s0[q] = r * den[q][2 * mt + 1]; s1[q] = r * den[q][2 * mt];
      }
    }
    uint32_t ah[4], al[4];
    f3_split_f16(s0[0], s0[1], ah[0], al[0]);
    f3_split_f16(s1[0], s1[1], ah[1], al[1]);
    f3_split_f16(s0[2], s0[3], ah[2], al[2]);
    f3_split_f16(s1[2], s1[3], ah[3], al[3]);
    f3_mma9<ABL>(ah, al, bv, acc[mt]);
  }
}

// The k8 tail step: columns 2t, 2t + 1 of the fragment = queries base + t and (NTQ == 2) base + 4 + t.
// tq -> patch[top-left tap][base + t]; bt -> this lane's {hi pair, lo pair} entries, one per class tile.
template <bool FAST, int ABL, int NTQ>
__device__ __forceinline__ void f3_tail(const float* __restrict__ tq, const uint2* __restrict__ bt, float ly, float ly1, float k1,
                                        float k2, float (*acc)[F3_NT][4]) {
  float sa[4], sb[4] = {0.f, 0.f, 0.f, 0.f};
  {
    float x0, d0;
    f3_line(tq[0], tq[F3_PITCH], tq[F3_TW * F3_PITCH], tq[(F3_TW + 1) * F3_PITCH], ly, ly1, k1, k2, x0, d0);
    if (ABL & 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) sa[j] = fmaf((float)j, d0, x0);
    } else {
      f3_sig4<FAST>(x0, d0, sa);
    }
  }
  if (NTQ == 2) {
    float x0, d0;
    f3_line(tq[4], tq[F3_PITCH + 4], tq[F3_TW * F3_PITCH + 4], tq[(F3_TW + 1) * F3_PITCH + 4], ly, ly1, k1, k2, x0, d0);
    if (ABL & 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) sb[j] = fmaf((float)j, d0, x0);
    } else {
      f3_sig4<FAST>(x0, d0, sb);
    }
  }
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) f3_split_f16(sa[j], sb[j], h[j], l[j]);
  uint2 b[F3_NT];
#pragma unroll
  for (int nt = 0; nt < F3_NT; ++nt) b[nt] = bt[nt * 32];
  if (ABL & 2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j >> 1][j & 1][j] += __uint_as_float((h[j] ^ l[j] ^ b[j % F3_NT].x ^ b[(j + 1) % F3_NT].y) & 0x3fffffffu);
    return;
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < F3_NT; ++nt) f3_mma_k8(acc[mt][nt], h[2 * mt], h[2 * mt + 1], b[nt].x);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < F3_NT; ++nt) f3_mma_k8(acc[mt][nt], h[2 * mt], h[2 * mt + 1], b[nt].y);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < F3_NT; ++nt) f3_mma_k8(acc[mt][nt], l[2 * mt], l[2 * mt + 1], b[nt].x);
}

// One unit: the 8 runs of cells 2u, 2u + 1 of the tile against every query; the class sums stay in the accumulators.
template <bool FAST, int ABL>
__device__ __forceinline__ void f3_unit(const float* __restrict__ tap, const uint4* __restrict__ bp, int nks, int ntq,
                                        const float* __restrict__ tq, const uint2* __restrict__ bt, float ly, float ly1,
                                        float k1, float k2, float (*acc)[F3_NT][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < F3_NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
  float4 a4 = *reinterpret_cast<const float4*>(tap);
  float4 b4 = *reinterpret_cast<const float4*>(tap + F3_PITCH);
  float4 c4 = *reinterpret_cast<const float4*>(tap + F3_TW * F3_PITCH);
  float4 d4 = *reinterpret_cast<const float4*>(tap + (F3_TW + 1) * F3_PITCH);
#pragma unroll 1
  for (int ks = 0; ks < nks; ++ks) {
    uint4 bv[F3_NT];
#pragma unroll
    for (int nt = 0; nt < F3_NT; ++nt) bv[nt] = bp[(nt * F3_KS + ks) * 32];
    // the next step's taps (the pitch leaves 20 spare words after the last step: finite garbage, unused)
    const float* tn = tap + 16 * (ks + 1);
    const float4 an = *reinterpret_cast<const float4*>(tn);
    const float4 bn = *reinterpret_cast<const float4*>(tn + F3_PITCH);
    const float4 cn = *reinterpret_cast<const float4*>(tn + F3_TW * F3_PITCH);
    const float4 dn = *reinterpret_cast<const float4*>(tn + (F3_TW + 1) * F3_PITCH);
    f3_step<FAST, ABL>(a4, b4, c4, d4, bv, ly, ly1, k1, k2, acc);
    a4 = an; b4 = bn; c4 = cn; d4 = dn;
  }
  if (ntq == 1) f3_tail<FAST, ABL, 1>(tq, bt, ly, ly1, k1, k2, acc);
  else if (ntq == 2) f3_tail<FAST, ABL, 2>(tq, bt, ly, ly1, k1, k2, acc);
}

struct F3Tile {
  int b, r0, c0;
};
__device__ __forceinline__ F3Tile f3_tile(const F3Params& p, int t) {
  const int tx = t % p.tilesX;
  const int rr = t / p.tilesX;
  F3Tile T;
  T.b = rr / p.tilesY;
  T.r0 = F3_CELLS_Y * (rr - T.b * p.tilesY) - 1;
  T.c0 = F3_CELLS_X * tx - 1;
  return T;
}

__global__ void __launch_bounds__(F3_THREADS, 1)
rba_einsum_score3_kernel(const __grid_constant__ CUtensorMap tmY_hi, const __grid_constant__ CUtensorMap tmY_lo,
                         const __grid_constant__ CUtensorMap tmE_hi, const __grid_constant__ CUtensorMap tmE_lo,
                         const F3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  F3Bars* bars = reinterpret_cast<F3Bars*>(smem + F3_OFF_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmY_hi); prefetch_tmap(&tmY_lo); prefetch_tmap(&tmE_hi); prefetch_tmap(&tmE_lo);
    for (int s = 0; s < F3_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int g2 = 0; g2 < 2; ++g2) {
      mbar_init(&bars->acc_full[g2], 1); mbar_init(&bars->acc_empty[g2], F3_GW);
      bars->amax[g2][0] = 0u; bars->amax[g2][1] = 0u;
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)), "r"(F3_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer + einsum issuer (one warp, software-pipelined over the stage ring) =====================
    constexpr uint32_t idE = make_idesc(TC_BM, F3_NQ);              // bf16, both operands K-major
    const uint32_t smem0 = smem_u32(smem);
    const int my_tiles = p.ntiles > (int)blockIdx.x ? (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total = my_tiles * p.nkb;
    int g_load = 0, lt_load = 0, kb_load = 0;              // K blocks whose loads have been issued; tile / K block of the next one
    uint32_t ls = 0, lph = 1;                              // fresh "empty" barriers pass a wait on parity 1
    uint32_t ms = 0, mph = 0;
    auto issue_loads_upto = [&](int limit) {
      while (g_load < total && g_load < limit) {
        mbar_wait(&bars->empty[ls], lph);                  // the MMAs that read this stage were issued by this warp: short wait
        if (elect_one()) {
          const F3Tile T = f3_tile(p, (int)blockIdx.x + lt_load * (int)gridDim.x);
          uint8_t* st = smem + ls * F3_STAGE_BYTES;
          mbar_expect_tx(&bars->full[ls], F3_STAGE_TX);
          tma_load_4d(st, &tmY_hi, &bars->full[ls], kb_load * F3_BK, T.c0, T.r0, T.b);
          tma_load_4d(st + F3_A_BYTES, &tmY_lo, &bars->full[ls], kb_load * F3_BK, T.c0, T.r0, T.b);
          tma_load_3d(st + 2 * F3_A_BYTES, &tmE_hi, &bars->full[ls], kb_load * F3_BK, 0, T.b);
          tma_load_3d(st + 2 * F3_A_BYTES + F3_E_BYTES, &tmE_lo, &bars->full[ls], kb_load * F3_BK, 0, T.b);
        }
        __syncwarp();
        ++g_load;
        if (++kb_load == p.nkb) { kb_load = 0; ++lt_load; }
        if (++ls == F3_STAGES) { ls = 0; lph ^= 1; }
      }
    };
    int g = 0;
    for (int lt = 0; lt < my_tiles; ++lt) {
      issue_loads_upto(g + F3_STAGES);                     // the first stages of this tile load under the previous tile's score phase
      const int pg = F3_NG == 2 ? (lt & 1) : 0;            // group that scores this tile; its accumulator
      const int plt = F3_NG == 2 ? (lt >> 1) : lt;         // tile count of that group
      const uint32_t dcol = tmem_base + (uint32_t)(pg * 128);
      if (plt > 0) {
        mbar_wait_sleep(&bars->acc_empty[pg], (uint32_t)(plt - 1) & 1);   // the group has drained its previous tile's D1
        tc_fence_after();
      }
      for (int kb = 0; kb < p.nkb; ++kb, ++g) {
        issue_loads_upto(g + F3_STAGES);
        mbar_wait(&bars->full[ms], mph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t base = smem0 + ms * F3_STAGE_BYTES;
          const uint64_t a_hi = make_sdesc64(base), a_lo = make_sdesc64(base + F3_A_BYTES);
          const uint64_t e_hi = make_sdesc64(base + 2 * F3_A_BYTES), e_lo = make_sdesc64(base + 2 * F3_A_BYTES + F3_E_BYTES);
#pragma unroll
          for (int k = 0; k < F3_BK / 16; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            umma_bf16(dcol, a_hi + adv, e_hi + adv, idE, (kb | k) != 0);
            umma_bf16(dcol, a_hi + adv, e_lo + adv, idE, 1);
            umma_bf16(dcol, a_lo + adv, e_hi + adv, idE, 1);
          }
          umma_commit(&bars->empty[ms]);
          if (kb == p.nkb - 1) umma_commit(&bars->acc_full[pg]);
        }
        __syncwarp();
        if (++ms == F3_STAGES) { ms = 0; mph ^= 1; }
      }
    }
  } else if (warp >= F3_SW) {
    // ===================== drain + score: warps 2..17 =====================
    const int cw = warp - F3_SW;
    const int gi = cw / F3_GW, gw = cw - gi * F3_GW;       // group; warp inside the group
    const int ctid = gw * 32 + lane;                       // thread inside the group
    const int qd = warp & 3, grp = gw >> 2;                // TMEM lane quadrant; chunk group of the drain
    uint8_t* gsm = smem + F3_OFF_PATCH + gi * F3_GROUP_BYTES;
    float* sPatch = reinterpret_cast<float*>(gsm);         // [pixel][query], pitch F3_PITCH
    uint4* sP = reinterpret_cast<uint4*>(gsm + F3_GOFF_P);
    uint2* sPt = reinterpret_cast<uint2*>(gsm + F3_GOFF_PT);
    float* sBias = reinterpret_cast<float*>(gsm + F3_GOFF_BIAS);
    const uint32_t dcol = tmem_base + (uint32_t)(gi * 128);
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const int m = qd * 32 + lane;                          // TMEM lane = low-res pixel of the drain
    const int g = lane >> 2, t = lane & 3;                 // run of the unit; query quad of the step / output pixel of the store
    const int half = g >> 2, dy = g & 3;                   // cell of the unit, output row inside the cell
    const float ly = 0.125f + 0.25f * (float)dy, ly1 = 1.0f - ly, k1 = 0.25f * ly1, k2 = 0.25f * ly;
    const float SCALE = -1.4426950408889634f;
    const int dbg = p.debug;
    // k16 steps, and the queries left for the k8 tail step: 1..4 -> one per lane, 5..8 -> two per lane, more -> a full step
    const int qrem = p.Q & 15;
    const int nks = (p.Q >> 4) + (qrem > 8 ? 1 : 0);
    const int ntq = qrem == 0 || qrem > 8 ? 0 : (qrem <= 4 ? 1 : 2);
    const uint4* bp = sP + lane;
    const uint2* bt = sPt + lane;
    int cur_b = -1;
    uint32_t lt = 0;
    for (int tt = blockIdx.x + gi * gridDim.x; tt < p.ntiles; tt += F3_NG * gridDim.x, ++lt) {   // the CTA's tiles alternate between the groups
      const F3Tile T = f3_tile(p, tt);
      F3_STAMP(lt, 0);
      if (ctid == 0) bars->amax[gi][lt & 1] = 0u;               // last read one tile ago, before that tile's closing barrier
      f3_bar_group(gi);                                     // every warp has finished the previous tile: patch, bias and sP are free
      if (T.b != cur_b) {
        // ---- per image: scaled bias; class probabilities as f16 hi/lo B fragments, scaled by 2 log2(e)
        // (tanh(s) = 1 - 2 / (1 + 2^(2 log2(e) s))): entry [class tile][k16 step][lane (g, t)] = {b0_hi, b1_hi, b0_lo, b1_lo},
        // class 8 nt + g, b0 = queries 16 ks + 4t + {0, 1}, b1 = queries 16 ks + 4t + {2, 3} ----
        cur_b = T.b;
        if (ctid < F3_NQ) sBias[ctid] = (p.bias && ctid < p.Q) ? p.bias[(size_t)T.b * p.Q + ctid] * SCALE : 0.f;
        for (int e4 = ctid; e4 < (F3_P_BYTES + F3_PT_BYTES) / 16; e4 += F3_GW * 32) sP[e4] = make_uint4(0u, 0u, 0u, 0u);   // sP and sPt
        f3_bar_group(gi);
        if (ctid < p.Q) {
          const int q = ctid;
          const float* lg = p.logits + ((size_t)T.b * p.Q + q) * (p.K + 1);
          float mx = lg[0];
          for (int c = 1; c <= p.K; ++c) mx = fmaxf(mx, lg[c]);
          float ssum = 0.f;
          for (int c = 0; c <= p.K; ++c) ssum += expf(lg[c] - mx);
          const float inv = 2.8853900817779268f / ssum;
          const int ks = q >> 4, r = q & 15, tq = r >> 2, e = r & 3;
          for (int c = 0; c < p.Kc; ++c) {
            const float pv = expf(lg[c] - mx) * inv;
            const __half hh = __float2half_rn(pv);
            __half* ent = reinterpret_cast<__half*>(sP + (((c >> 3) * F3_KS + ks) * 32 + (c & 7) * 4 + tq));
            const __half hl = __float2half_rn(pv - __half2float(hh));
            if (ks < nks) {
              ent[e] = hh;
              ent[4 + e] = hl;
            } else {
              // tail step: column 2 jt = query 16 nks + jt, column 2 jt + 1 = query 16 nks + 4 + jt
              __half* et = reinterpret_cast<__half*>(sPt + ((c >> 3) * 32 + (c & 7) * 4 + (r & 3)));
              et[r >> 2] = hh;
              et[2 + (r >> 2)] = hl;
            }
          }
        }
        f3_bar_group(gi);
      }
      // ---- drain the accumulator: TMEM lane = low-res pixel, column = query -> patch[pixel][query] ----
      F3_STAMP(lt, 1);
      mbar_wait(&bars->acc_full[gi], lt & 1);
      tc_fence_after();
      F3_STAMP(lt, 2);
      float am = 0.f;
      for (int chunk = grp; chunk * 16 < F3_NQ; chunk += F3_GW / 4) {
        const int q0 = chunk * 16;
        uint32_t v[16];
        tmem_ld16(dcol + lane_addr + (uint32_t)q0, v);
        tmem_ld_wait();
        float* prow = sPatch + (m < F3_PATCH_ROWS ? m : F3_PATCH_ROWS - 1) * F3_PITCH + q0;   // rows >= 119: garbage, parked on the spare row
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sBias + q0 + 4 * j4);
          float4 o;
          o.x = fmaf(__uint_as_float(v[4 * j4]), SCALE, b4.x);
          o.y = fmaf(__uint_as_float(v[4 * j4 + 1]), SCALE, b4.y);
          o.z = fmaf(__uint_as_float(v[4 * j4 + 2]), SCALE, b4.z);
          o.w = fmaf(__uint_as_float(v[4 * j4 + 3]), SCALE, b4.w);
          am = fmaxf(fmaxf(am, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
          *reinterpret_cast<float4*>(prow + 4 * j4) = o;
        }
      }
      if (m >= F3_TW * F3_TH) am = 0.f;                     // rows beyond the box: whatever the stage held before
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[gi]);         // the einsum of the next tile may start
      {
        const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(am));   // non-negative floats order like their bits
        if (lane == 0) atomicMax(&bars->amax[gi][lt & 1], wm);
      }
      F3_STAMP(lt, 3);
      f3_bar_group(gi);                                     // every pixel of the patch has been written
      F3_STAMP(lt, 4);
      // ---- image borders: replicate the edge into the out-of-range taps (CTA-uniform) ----
      {
        const int rtop = T.r0 < 0 ? 0 : -1;                                   // patch row above the image
        const int rbot = p.h - T.r0 <= F3_TH - 1 ? p.h - T.r0 : -1;      // first patch row below the image
        const int cleft = T.c0 < 0 ? 0 : -1;
        const int cright = p.w - T.c0 <= F3_TW - 1 ? p.w - T.c0 : -1;
        if (rtop >= 0 || rbot >= 0 || cleft >= 0 || cright >= 0) {
          // rows first, then columns (so that the corners pick up the diagonal neighbour); F3_NQ / 4 = 28 float4 per pixel
          for (int e4 = ctid; e4 < 2 * F3_TW * (F3_NQ / 4); e4 += F3_GW * 32) {
            const int which = e4 / (F3_TW * (F3_NQ / 4)), r2 = e4 - which * (F3_TW * (F3_NQ / 4));
            const int col = r2 / (F3_NQ / 4), q4 = r2 - col * (F3_NQ / 4);
            const int dstrow = which ? rbot : rtop;
            if (dstrow < 0) continue;
            const int srcrow = which ? rbot - 1 : 1;
            *reinterpret_cast<float4*>(sPatch + (dstrow * F3_TW + col) * F3_PITCH + 4 * q4) =
                *reinterpret_cast<const float4*>(sPatch + (srcrow * F3_TW + col) * F3_PITCH + 4 * q4);
          }
          f3_bar_group(gi);
          for (int e4 = ctid; e4 < 2 * F3_TH * (F3_NQ / 4); e4 += F3_GW * 32) {
            const int which = e4 / (F3_TH * (F3_NQ / 4)), r2 = e4 - which * (F3_TH * (F3_NQ / 4));
            const int row = r2 / (F3_NQ / 4), q4 = r2 - row * (F3_NQ / 4);
            const int dstcol = which ? cright : cleft;
            if (dstcol < 0) continue;
            const int srccol = which ? cright - 1 : 1;
            *reinterpret_cast<float4*>(sPatch + (row * F3_TW + dstcol) * F3_PITCH + 4 * q4) =
                *reinterpret_cast<const float4*>(sPatch + (row * F3_TW + srccol) * F3_PITCH + 4 * q4);
          }
          f3_bar_group(gi);
        }
      }
      const bool fast = !(dbg & 1) && bars->amax[gi][lt & 1] <= __float_as_uint(F3_UFAST);
      F3_STAMP(lt, 5);
      if (dbg & 2) continue;
      // ---- score phase: unit u = cells 2u, 2u + 1 (row-major over the 6 x 16 cells) ----
      float* const rba_b = p.rba + (size_t)T.b * p.H * p.W;
#pragma unroll 1
      for (int u = gw; u < F3_NUNIT; u += F3_GW) {
        const int cell = 2 * u + half;
        const int br = cell / F3_CELLS_X, bc = cell - br * F3_CELLS_X;
        const float* tap = sPatch + (br * F3_TW + bc) * F3_PITCH + 4 * t;
        const float* tq = tap + 16 * nks - 3 * t;            // -> patch[top-left tap][16 nks + t]
        float acc[2][F3_NT][4];
#if RBA_F3_PROFILE
        if ((dbg & 12) == 12) f3_unit<true, 3>(tap, bp, nks, ntq, tq, bt, ly, ly1, k1, k2, acc);
        else if (dbg & 4) f3_unit<true, 1>(tap, bp, nks, ntq, tq, bt, ly, ly1, k1, k2, acc);
        else if (dbg & 8) f3_unit<true, 2>(tap, bp, nks, ntq, tq, bt, ly, ly1, k1, k2, acc);
        else
#endif
        if (fast) f3_unit<true, 0>(tap, bp, nks, ntq, tq, bt, ly, ly1, k1, k2, acc);
        else f3_unit<false, 0>(tap, bp, nks, ntq, tq, bt, ly, ly1, k1, k2, acc);
        // ---- epilogue: acc[mt][nt][e] = scaled class sum of class 8 nt + 2t + (e & 1), pixel 2 mt + (e >> 1) of the run.
        // -sum_c tanh(s_c) = -n + 2 sum_c 1 / (1 + 2^(s'_c)); padded classes hold exactly 0 and contribute tanh(0) = 0 ----
        float r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int mt = j >> 1, h2 = (j & 1) * 2;
          if (dbg & 16) { r[j] = acc[mt][0][h2] + acc[mt][0][h2 + 1] + acc[mt][1][h2] + acc[mt][1][h2 + 1] + acc[mt][2][h2] + acc[mt][2][h2 + 1]; continue; }
          r[j] = f3_rsum4(acc[mt][0][h2], acc[mt][0][h2 + 1], acc[mt][1][h2], acc[mt][1][h2 + 1]) +
                 f3_rsum2(acc[mt][2][h2], acc[mt][2][h2 + 1]);
        }
        // transpose-reduce over the quad: lane t ends with the total of pixel t
        const bool odd = (t & 1) != 0, up = (t & 2) != 0;
        const float x0 = __shfl_xor_sync(0xffffffffu, odd ? r[0] : r[1], 1);
        const float x1 = __shfl_xor_sync(0xffffffffu, odd ? r[2] : r[3], 1);
        const float s01 = (odd ? r[1] : r[0]) + x0;           // pixel (t & 1)
        const float s23 = (odd ? r[3] : r[2]) + x1;           // pixel 2 + (t & 1)
        const float x2 = __shfl_xor_sync(0xffffffffu, up ? s01 : s23, 2);
        const float tot = (up ? s23 : s01) + x2;
        const int i = T.r0 + br, j = T.c0 + bc;               // low-res coordinates of the cell's top-left tap
        const int y = 4 * i + 2 + dy, x = 4 * j + 2 + t;
        const bool ok = i <= p.h - 1 && j <= p.w - 1 && y >= 0 && y < p.H && x >= 0 && x < p.W;
        if (ok) rba_b[(size_t)y * p.W + x] = fmaf(2.0f, tot, -(float)(F3_NT * 8));
      }
      F3_STAMP(lt, 6);
    }
    F3_STAMP(lt < 32 ? lt : 31, 0);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(F3_TMEM_COLS) : "memory");
  }
}

int einsum_score3_supported(int Q, int K, int D) { return Q > 0 && Q <= 104 && K > 0 && K + 1 <= F3_NT * 8 && D % 64 == 0; }

// RbA-only launch (no sem_seg, score function RbA); same operands as einsum_score_launch
int einsum_score3_launch(const uint16_t* e_hi, const uint16_t* e_lo, const float* bias, const uint16_t* y_hi, const uint16_t* y_lo,
                         const float* logits, int B, int Q, int K, int D, int h, int w, int H, int W, int include_void, float* rba,
                         cudaStream_t st) {
  RBA_CHECK(einsum_score3_supported(Q, K, D), "einsum_score3: unsupported Q=%d (<= 104) K=%d (<= 23) D=%d (multiple of 64)", Q, K, D);
  RBA_CHECK(((uintptr_t)e_hi & 15) == 0 && ((uintptr_t)e_lo & 15) == 0 && ((uintptr_t)y_hi & 15) == 0 && ((uintptr_t)y_lo & 15) == 0,
            "einsum_score3: operand planes must be 16-byte aligned");
  F3Params p;
  memset(&p, 0, sizeof(p));
  p.logits = logits; p.bias = bias; p.rba = rba;
  p.B = B; p.Q = Q; p.K = K; p.h = h; p.w = w; p.H = H; p.W = W;
  p.Kc = include_void ? K + 1 : K;
  p.nkb = D / F3_BK;
  { const char* e = getenv("RBA_FS_DEBUG"); p.debug = e ? atoi(e) : 0; }
  p.tilesX = (int)cdiv(w + 1, F3_CELLS_X); p.tilesY = (int)cdiv(h + 1, F3_CELLS_Y);
  const int64_t nt = (int64_t)B * p.tilesX * p.tilesY;
  RBA_CHECK(nt < (1LL << 31), "einsum_score3: too many tiles");
  p.ntiles = (int)nt;
  CUtensorMap ty_hi, ty_lo, te_hi, te_lo;
  RBA_TRY_(make_map_nhwc_k32(&ty_hi, y_hi, B, h, w, D, F3_TW, F3_TH));
  RBA_TRY_(make_map_nhwc_k32(&ty_lo, y_lo, B, h, w, D, F3_TW, F3_TH));
  RBA_TRY_(make_map_embed_k32(&te_hi, e_hi, B, Q, D));
  RBA_TRY_(make_map_embed_k32(&te_lo, e_lo, B, Q, D));
  static PerDeviceOnce once;
  if (once.needed()) {
    RBA_CUDA(cudaFuncSetAttribute(rba_einsum_score3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F3_SMEM));
    once.done();
  }
  dim3 grid((unsigned)std::min<int64_t>(nt, num_sms()));
  static const bool timeline = getenv("RBA_FS_TIMELINE") != nullptr;       // profiling aid: prints CTA 0's clock stamps
  static long long* tl_dev = nullptr;
  if (timeline) {
    if (!tl_dev) RBA_CUDA(cudaMalloc(&tl_dev, 32 * 8 * sizeof(long long)));
    RBA_CUDA(cudaMemsetAsync(tl_dev, 0, 32 * 8 * sizeof(long long), st));
    p.tl = tl_dev;
  }
  rba_einsum_score3_kernel<<<grid, F3_THREADS, F3_SMEM, st>>>(ty_hi, ty_lo, te_hi, te_lo, p);
  RBA_LAUNCHED();
  if (timeline) {
    static long long h2[32 * 8];
    RBA_CUDA(cudaStreamSynchronize(st));
    RBA_CUDA(cudaMemcpy(h2, tl_dev, sizeof(h2), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[score3 timeline, CTA 0 warp 2, clocks] tile | top barrier (+ image setup) | acc wait | drain | patch barrier | border + flag | units | tile total\n");
    for (int i = 0; i + 1 < 24; ++i) {
      const long long* a = h2 + i * 8;
      const long long nxt = h2[(i + 1) * 8];
      fprintf(stderr, "%4d | %6lld %6lld %6lld %6lld %6lld %6lld | %6lld\n", i, a[1] - a[0], a[2] - a[1], a[3] - a[2], a[4] - a[3], a[5] - a[4],
              a[6] ? a[6] - a[5] : -1, nxt - a[0]);
    }
  }
  return RBA_OK;
}

}  // namespace rba
