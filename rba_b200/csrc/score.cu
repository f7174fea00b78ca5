// Fused RbA score kernel.
//
// Replaces, in ONE pass over the low-resolution mask logits:
//   F.interpolate(pred_masks, x4, bilinear, align_corners=False)      maskformer_model.py:294-299
//   softmax(pred_logits)[..., :-1]; sigmoid(mask); einsum("qc,qhw->chw") maskformer_model.py:381-386
//   sem_seg_postprocess crop to the un-padded size                     maskformer_model.py:330-333
//   -tanh(sem_seg).sum(0)                                              evaluate_ood.py:148-150
// The (Q,4h,4w) upsampled tensor (839 MB/img at 1024x2048) and, unless requested, the (K,H,W) sem_seg
// tensor are never materialised.  Algorithmic HBM bytes per image: 4*Q*h*w + 4*Q*(K+1) + 4*H*W.
//
// Work decomposition: one CTA = 8 output rows x 128 output columns; one warp per output row, one lane
// per 4 consecutive output pixels (they share the 3 low-res columns j-1, j, j+1).  The low-res patch
// (4 rows x 34 cols per query) is staged through shared memory with cp.async, double buffered over
// chunks of QC queries.  Class probabilities (Q x K, padded to 20 floats/row) live in shared memory and
// are read as broadcast float4s.  Accumulators: 4 pixels x K classes per thread in registers.
#include <cuda_fp16.h>

#include "common.cuh"

namespace rba {

constexpr int SC_TW = 128;   // output tile width
constexpr int SC_TH = 8;     // output tile height
constexpr int SC_PW = 36;    // patch row pitch (34 used)
constexpr int SC_PR = 4;     // patch rows
constexpr int SC_QC = 20;    // queries per pipeline stage
constexpr int SC_KP = 20;    // padded classes per query row in smem (K <= 20)

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Single-MUFU helpers: ex2.approx / rcp.approx are accurate to ~2 ulp, far inside the 1e-3 parity budget.
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// sigmoid(u) = 1 / (1 + 2^(-u*log2e)); the exponent is clamped so 1+e stays finite (u > -87)
__device__ __forceinline__ float fast_sigmoid(float u) {
  return fast_rcp(1.0f + fast_ex2(fminf(-1.4426950408889634f * u, 126.0f)));
}
// tanh(x) for x >= 0 (sem_seg is a sum of non-negative terms): 1 - 2/(1 + e^(2x)); abs error ~1e-7
__device__ __forceinline__ float fast_tanh_pos(float x) {
  return 1.0f - 2.0f * fast_rcp(1.0f + fast_ex2(fminf(2.8853900817779268f * x, 126.0f)));
}

// PyTorch area_pixel_compute_source_index (align_corners=False, scale 1/4) for output index o:
// src = max((o+0.5)/4 - 0.5, 0); i0 = floor(src); l1 = src - i0.
__device__ __forceinline__ void up4_coeff(int o, int& i0, float& l1) {
  float src = (o + 0.5f) * 0.25f - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  l1 = src - (float)i0;
}

template <int K, bool WRITE_SEM>
__global__ void __launch_bounds__(256, 2)
rba_score_kernel(const float* __restrict__ masks, const float* __restrict__ logits, int Q, int h, int w, int H, int W,
                 float* __restrict__ rba, float* __restrict__ sem) {
  extern __shared__ __align__(16) float smem[];
  float* sP = smem;                                   // [Q][SC_KP]
  float* sPatch = smem + (size_t)Q * SC_KP;           // [2][SC_QC][SC_PR][SC_PW]
  const int tid = threadIdx.x, lane = tid & 31, ry = tid >> 5;
  const int b = blockIdx.z;
  const int X0 = blockIdx.x * SC_TW, Y0 = blockIdx.y * SC_TH;
  const int lx0 = X0 / 4 - 1, ly0 = Y0 / 4 - 1;       // low-res origin of the patch (may be -1)
  const float* mb = masks + (size_t)b * Q * h * w;

  auto load_chunk = [&](int chunk, int buf) {
    const int q0 = chunk * SC_QC;
    float* dst = sPatch + (size_t)buf * SC_QC * SC_PR * SC_PW;
    for (int e = tid; e < SC_QC * SC_PR * 34; e += 256) {
      int pc = e % 34;
      int t = e / 34;
      int pr = t % SC_PR;
      int qq = t / SC_PR;
      int q = q0 + qq;
      if (q < Q) {
        int gy = min(max(ly0 + pr, 0), h - 1);
        int gx = min(max(lx0 + pc, 0), w - 1);
        cp_async4(dst + (qq * SC_PR + pr) * SC_PW + pc, mb + ((size_t)q * h + gy) * w + gx);
      }
    }
    cp_async_commit();
  };

  const int nchunks = (Q + SC_QC - 1) / SC_QC;
  load_chunk(0, 0);

  // class probabilities: softmax over K+1 logits, keep the first K (maskformer_model.py:382)
  for (int q = tid; q < Q; q += 256) {
    const float* lg = logits + ((size_t)b * Q + q) * (K + 1);
    float m = lg[0];
#pragma unroll
    for (int c = 1; c <= K; ++c) m = fmaxf(m, lg[c]);
    float e[K + 1];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c <= K; ++c) { e[c] = expf(lg[c] - m); s += e[c]; }
    float inv = 1.0f / s;
#pragma unroll
    for (int c = 0; c < SC_KP; ++c) sP[q * SC_KP + c] = (c < K) ? e[c] * inv : 0.f;
  }

  // per-thread interpolation coefficients
  const int y = Y0 + ry;
  int iy0; float wy1;
  up4_coeff(y, iy0, wy1);
  const float wy0 = 1.f - wy1;
  const int pr0 = iy0 - ly0;                           // patch row of the upper tap (0..2)
  const int pr1 = min(iy0 + 1, h - 1) - ly0;
  // Pixels 4j, 4j+1 interpolate low-res columns (j-1, j); 4j+2, 4j+3 use (j, j+1): patch columns
  // (lane, lane+1) and (lane+1, lane+2).  The loader clamps columns to [0, w-1], which reproduces
  // PyTorch's i1 = min(i0+1, w-1); at x < 2 the clamped source index gives weights (1, 0).
  float wx0[4], wx1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int ix0; float l1;
    up4_coeff(X0 + 4 * lane + i, ix0, l1);
    wx1[i] = l1; wx0[i] = 1.f - l1;
  }

  float acc[4][K];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < K; ++c) acc[i][c] = 0.f;

  for (int chunk = 0; chunk < nchunks; ++chunk) {
    if (chunk + 1 < nchunks) {
      load_chunk(chunk + 1, (chunk + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* pb = sPatch + (size_t)(chunk & 1) * SC_QC * SC_PR * SC_PW;
    const int qn = min(SC_QC, Q - chunk * SC_QC);
    for (int qq = 0; qq < qn; ++qq) {
      const float* r0 = pb + (qq * SC_PR + pr0) * SC_PW + lane;
      const float* r1 = pb + (qq * SC_PR + pr1) * SC_PW + lane;
      float col[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) col[j] = wy0 * r0[j] + wy1 * r1[j];
      float sg[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float u = wx0[i] * col[i >> 1] + wx1[i] * col[(i >> 1) + 1];
        sg[i] = fast_sigmoid(u);                       // sigmoid, maskformer_model.py:383
      }
      const float4* pq = reinterpret_cast<const float4*>(sP + (size_t)(chunk * SC_QC + qq) * SC_KP);
      float p[SC_KP];
#pragma unroll
      for (int v = 0; v < SC_KP / 4; ++v) {
        float4 t = pq[v];
        p[4 * v] = t.x; p[4 * v + 1] = t.y; p[4 * v + 2] = t.z; p[4 * v + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < K; ++c) acc[i][c] = fmaf(p[c], sg[i], acc[i][c]);
    }
    __syncthreads();
  }

  // epilogue: crop, optional sem_seg planes, -sum tanh
  if (y >= H) return;
  const int x0 = X0 + 4 * lane;
  if (x0 >= W) return;
  float r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < K; ++c) s += fast_tanh_pos(acc[i][c]);   // evaluate_ood.py:150
    r[i] = -s;
  }
  const bool vec = ((W & 3) == 0) && (x0 + 3 < W);
  float* ro = rba + ((size_t)b * H + y) * W + x0;
  if (vec) {
    *reinterpret_cast<float4*>(ro) = make_float4(r[0], r[1], r[2], r[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (x0 + i < W) ro[i] = r[i];
  }
  if (WRITE_SEM) {
#pragma unroll
    for (int c = 0; c < K; ++c) {
      float* so = sem + (((size_t)b * K + c) * H + y) * W + x0;
      if (vec) {
        *reinterpret_cast<float4*>(so) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (x0 + i < W) so[i] = acc[i][c];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core variant: the (Q x K) contraction runs on mma.sync (m16n8k16, fp16 hi/lo split = fp32-class accuracy:
// sigma = s_hi + s_lo, p = p_hi + p_lo, products s_hi*p_hi + s_hi*p_lo + s_lo*p_hi, fp32 accumulate), which frees the
// FMA pipe: what remains per (pixel, query) is the bilinear tap combination, two MUFU ops (ex2, rcp) and the fp16
// split.  The sigma values are produced directly in the A-fragment layout (no shared-memory staging of the operand):
// MMA row g <-> pixel (x0+g, y), row g+8 <-> pixel (x0+g, y+1) — two vertically adjacent pixels that share all four
// low-resolution taps — and the k index is the query.  Class probabilities (B operand, [K pad 24] x [Q pad 112]) live in
// registers for the whole CTA.  CTA = 4 output rows x 128 columns, 8 warps = 2 row pairs x 4 column groups of 32 px.
// ------------------------------------------------------------------------------------------------
constexpr int SM_KS = 7;                    // k16 steps: Q padded to 112
constexpr int SM_QP = SM_KS * 16;
constexpr int SM_NT = 3;                    // n8 tiles: K padded to 24
constexpr int SM_QS = 120;                  // shared-memory pitch (words) of one query row: = 24 mod 32, so the 8-word
                                            // spans read by the 4 lanes of a quad never collide across quads
constexpr int SM_PR = 3, SM_PC = 34;        // patch rows / columns per CTA
constexpr int SM_TW = 128, SM_TH = 4;
constexpr int SM_PS = 122;                  // patch pitch (words) per (row, col): even (8-byte loads), = 26 mod 32 so the
                                            // column-major staging writes are at most 2-way conflicted
constexpr int SM_PATCH_WORDS = SM_PR * SM_PC * SM_PS;       // patch [row][col][query] fp32
constexpr int SM_B_WORDS = 24 * SM_QS;                      // probabilities [class][query pair]{hi.x2, lo.x2} (2 words/pair)

__device__ __forceinline__ uint32_t pack_f16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
__device__ __forceinline__ void split_f16x2(float e0, float e1, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16x2(e0, e1);
  const __half2 h = *reinterpret_cast<const __half2*>(&hi);
  const float2 f = __half22float2(h);
  lo = pack_f16x2(e0 - f.x, e1 - f.y);
}
__device__ __forceinline__ void mma_f16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int K, bool WRITE_SEM>
__global__ void __launch_bounds__(256, 3)
rba_score_mma_kernel(const float* __restrict__ masks, const float* __restrict__ logits, int Q, int h, int w, int H, int W,
                     float* __restrict__ rba, float* __restrict__ sem) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  float* sPatch = reinterpret_cast<float*>(sm_raw);                       // [3][34][SM_PS]: query innermost
  uint32_t* sB = reinterpret_cast<uint32_t*>(sm_raw) + SM_PATCH_WORDS;    // [24][SM_QS]: word 2*(q/2) = hi pair, +1 = lo pair
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int b = blockIdx.z;
  const int X0 = blockIdx.x * SM_TW, Y0 = blockIdx.y * SM_TH;
  const int lx0 = X0 / 4 - 1, ly0 = Y0 / 4 - 1;
  const float* mb = masks + (size_t)b * Q * h * w;

  // ---- stage the low-res patch transposed to [row][col][query] (padded queries read as 0) ----
  for (int e = tid; e < SM_QP * SM_PR * SM_PC; e += 256) {
    const int pc = e % SM_PC;
    const int t = e / SM_PC;
    const int pr = t % SM_PR, q = t / SM_PR;
    float* dst = sPatch + (pr * SM_PC + pc) * SM_PS + q;
    if (q < Q) {
      const int gy = min(max(ly0 + pr, 0), h - 1), gx = min(max(lx0 + pc, 0), w - 1);
      cp_async4(dst, mb + ((size_t)q * h + gy) * w + gx);
    } else {
      *dst = 0.f;
    }
  }
  cp_async_commit();
  // ---- class probabilities -> fp16 hi/lo pairs, [class][query pair] ----
  for (int e = tid; e < SM_B_WORDS; e += 256) sB[e] = 0u;
  __syncthreads();
  for (int q = tid; q < Q; q += 256) {
    const float* lg = logits + ((size_t)b * Q + q) * (K + 1);
    float m = lg[0];
#pragma unroll
    for (int c = 1; c <= K; ++c) m = fmaxf(m, lg[c]);
    float e[K + 1];
    float ssum = 0.f;
#pragma unroll
    for (int c = 0; c <= K; ++c) { e[c] = expf(lg[c] - m); ssum += e[c]; }
    const float inv = 1.0f / ssum;
    __half* sBh = reinterpret_cast<__half*>(sB);
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const float pv = e[c] * inv;
      const __half hh = __float2half_rn(pv);
      // halves of class row c: pair p = q/2 occupies halves [4p .. 4p+3] = {hi(q even), hi(q odd), lo(even), lo(odd)}
      const int base = (c * SM_QS + (q >> 1) * 2) * 2 + (q & 1);
      sBh[base] = hh;
      sBh[base + 2] = __float2half_rn(pv - __half2float(hh));
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- per-warp geometry ----
  const int rp = warp >> 2, xg = warp & 3;
  const int y = Y0 + 2 * rp;                              // rows y, y+1 share their two low-res rows
  int iy0, iyb; float l1a, l1b;
  up4_coeff(y, iy0, l1a);
  up4_coeff(y + 1, iyb, l1b);
  const int prA = iy0 - ly0, prB = min(iy0 + 1, h - 1) - ly0;
  // weights pre-multiplied by -log2(e): sigmoid(u) = 1 / (1 + 2^(-u log2 e))
  const float NL2E = -1.4426950408889634f;
  const float wyA0 = (1.f - l1a) * NL2E, wyA1 = l1a * NL2E;   // row y
  const float wyB0 = (1.f - l1b) * NL2E, wyB1 = l1b * NL2E;   // row y + 1
  const uint32_t* bbase = sB + g * SM_QS + tq * 2;             // + nt*8*SM_QS + ks*16 (+8)

#pragma unroll 1
  for (int ti = 0; ti < 4; ++ti) {
    const int x = X0 + 32 * xg + 8 * ti + g;
    int ix0; float lx1;
    up4_coeff(min(x, 4 * w - 1), ix0, lx1);
    ix0 = min(ix0, w - 1);
    const float wx0 = 1.f - lx1, wx1 = lx1;
    const int pcA = ix0 - lx0, pcB = min(ix0 + 1, w - 1) - lx0;
    // the four taps of this pixel pair, query-contiguous rows
    const float* t00 = sPatch + (prA * SM_PC + pcA) * SM_PS + tq * 2;
    const float* t01 = sPatch + (prA * SM_PC + pcB) * SM_PS + tq * 2;
    const float* t10 = sPatch + (prB * SM_PC + pcA) * SM_PS + tq * 2;
    const float* t11 = sPatch + (prB * SM_PC + pcB) * SM_PS + tq * 2;
    float acc[SM_NT][4];
#pragma unroll
    for (int nt = 0; nt < SM_NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll 1
    for (int ks = 0; ks < SM_KS; ++ks) {
      uint32_t ah[4], al[4];
#pragma unroll
      for (int hq = 0; hq < 2; ++hq) {                    // queries 16ks + 2t + {0,1} (+8 for hq = 1)
        const int qo = ks * 16 + hq * 8;
        const float2 a = *reinterpret_cast<const float2*>(t00 + qo);
        const float2 bq = *reinterpret_cast<const float2*>(t01 + qo);
        const float2 c = *reinterpret_cast<const float2*>(t10 + qo);
        const float2 d = *reinterpret_cast<const float2*>(t11 + qo);
        const float h0x = wx0 * a.x + wx1 * bq.x, h0y = wx0 * a.y + wx1 * bq.y;
        const float h1x = wx0 * c.x + wx1 * d.x, h1y = wx0 * c.y + wx1 * d.y;
        const float sAx = fast_rcp(1.0f + fast_ex2(wyA0 * h0x + wyA1 * h1x));   // row y
        const float sAy = fast_rcp(1.0f + fast_ex2(wyA0 * h0y + wyA1 * h1y));
        const float sBx = fast_rcp(1.0f + fast_ex2(wyB0 * h0x + wyB1 * h1x));   // row y + 1
        const float sBy = fast_rcp(1.0f + fast_ex2(wyB0 * h0y + wyB1 * h1y));
        split_f16x2(sAx, sAy, ah[2 * hq], al[2 * hq]);             // MMA row g
        split_f16x2(sBx, sBy, ah[2 * hq + 1], al[2 * hq + 1]);     // MMA row g + 8
      }
#pragma unroll
      for (int nt = 0; nt < SM_NT; ++nt) {
        const uint2 b0 = *reinterpret_cast<const uint2*>(bbase + nt * 8 * SM_QS + ks * 16);       // {hi, lo} of k = 2t, 2t+1
        const uint2 b1 = *reinterpret_cast<const uint2*>(bbase + nt * 8 * SM_QS + ks * 16 + 8);   // k = 2t+8, 2t+9
        mma_f16(acc[nt], ah, b0.x, b1.x);
        mma_f16(acc[nt], ah, b0.y, b1.y);
        mma_f16(acc[nt], al, b0.x, b1.x);
      }
    }
    // ---- epilogue: acc[nt][j] = sem_seg[class 8nt+2t+j] of pixel (x, y); acc[nt][2+j] of pixel (x, y+1) ----
    float ra = 0.f, rb = 0.f;
#pragma unroll
    for (int nt = 0; nt < SM_NT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = nt * 8 + tq * 2 + j;
        if (c < K) {
          ra += fast_tanh_pos(acc[nt][j]);
          rb += fast_tanh_pos(acc[nt][2 + j]);
          if (WRITE_SEM && x < W) {
            if (y < H) sem[(((size_t)b * K + c) * H + y) * W + x] = acc[nt][j];
            if (y + 1 < H) sem[(((size_t)b * K + c) * H + y + 1) * W + x] = acc[nt][2 + j];
          }
        }
      }
    ra += __shfl_xor_sync(0xffffffffu, ra, 1);
    ra += __shfl_xor_sync(0xffffffffu, ra, 2);
    rb += __shfl_xor_sync(0xffffffffu, rb, 1);
    rb += __shfl_xor_sync(0xffffffffu, rb, 2);
    if (x < W) {
      if (tq == 0 && y < H) rba[((size_t)b * H + y) * W + x] = -ra;
      if (tq == 1 && y + 1 < H) rba[((size_t)b * H + y + 1) * W + x] = -rb;
    }
  }
}

template <int K>
static int launch_score_mma(const float* masks, const float* logits, int B, int Q, int h, int w, int H, int W, float* rba,
                            float* sem, cudaStream_t st) {
  dim3 grid((unsigned)cdiv(4 * w, SM_TW), (unsigned)cdiv(4 * h, SM_TH), (unsigned)B);
  const size_t smem = (size_t)(SM_PATCH_WORDS + SM_B_WORDS) * 4;
  if (sem) {
    RBA_CUDA(cudaFuncSetAttribute(rba_score_mma_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rba_score_mma_kernel<K, true><<<grid, 256, smem, st>>>(masks, logits, Q, h, w, H, W, rba, sem);
  } else {
    RBA_CUDA(cudaFuncSetAttribute(rba_score_mma_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rba_score_mma_kernel<K, false><<<grid, 256, smem, st>>>(masks, logits, Q, h, w, H, W, rba, sem);
  }
  RBA_LAUNCHED();
  return RBA_OK;
}

template <int K>
static int launch_score(const float* masks, const float* logits, int B, int Q, int h, int w, int H, int W, float* rba,
                        float* sem, cudaStream_t st) {
  dim3 grid((unsigned)cdiv(4 * w, SC_TW), (unsigned)cdiv(4 * h, SC_TH), (unsigned)B);
  size_t smem = ((size_t)Q * SC_KP + 2 * SC_QC * SC_PR * SC_PW) * sizeof(float);
  if (sem) {
    RBA_CUDA(cudaFuncSetAttribute(rba_score_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rba_score_kernel<K, true><<<grid, 256, smem, st>>>(masks, logits, Q, h, w, H, W, rba, sem);
  } else {
    RBA_CUDA(cudaFuncSetAttribute(rba_score_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rba_score_kernel<K, false><<<grid, 256, smem, st>>>(masks, logits, Q, h, w, H, W, rba, sem);
  }
  RBA_LAUNCHED();
  return RBA_OK;
}

}  // namespace rba

extern "C" int rba_score_fused(const float* pred_masks, const float* pred_logits, int B, int Q, int K, int h, int w,
                               int H, int W, float* rba_out, float* sem_seg, void* stream) {
  using namespace rba;
  if (B == 0) return RBA_OK;   // empty batch: nothing to do (empty tensors carry null pointers)
  RBA_CHECK(pred_masks && pred_logits && rba_out, "rba_score_fused: null pointer");
  RBA_CHECK(B >= 0 && Q > 0 && h > 0 && w > 0, "rba_score_fused: bad shape B=%d Q=%d h=%d w=%d", B, Q, h, w);
  RBA_CHECK(H > 0 && W > 0 && H <= 4 * h && W <= 4 * w, "rba_score_fused: output (%d,%d) exceeds 4x(%d,%d)", H, W, h, w);
  RBA_CHECK(Q <= 2048, "rba_score_fused: Q=%d too large", Q);
  cudaStream_t st = (cudaStream_t)stream;
  static const bool use_mma = []() { const char* e = getenv("RBA_SCORE_MMA"); return !(e && e[0] == '0'); }();
  if (use_mma && K == 19 && Q > SM_QP - 16 && Q <= SM_QP)      // tensor-core path: Q padded to 112 (100 queries)
    return launch_score_mma<19>(pred_masks, pred_logits, B, Q, h, w, H, W, rba_out, sem_seg, st);
  switch (K) {
    case 19: return launch_score<19>(pred_masks, pred_logits, B, Q, h, w, H, W, rba_out, sem_seg, st);
    case 13: return launch_score<13>(pred_masks, pred_logits, B, Q, h, w, H, W, rba_out, sem_seg, st);  // StreetHazards
    case 3: return launch_score<3>(pred_masks, pred_logits, B, Q, h, w, H, W, rba_out, sem_seg, st);    // test size
    default: return fail(RBA_ERR_INVALID, "rba_score_fused: K=%d not instantiated (19, 13, 3)", K);
  }
}

// ------------------------------------------------------------------------------------------------
// DenseHybrid head: `ood_pred` logits (B, h, w, 2) from the BN-ReLU-1x1 head
// (mask2former_transformer_decoder.py:216-230,365-366,467-468) are resized to the image size with bilinear
// align_corners=TRUE (maskformer_model.py:303-305) and, optionally, folded into the score
//   score += log(softmax(ood_pred)[1] + 1e-9)        (evaluate_ood.py:161-173, score holds -logsumexp(sem_seg))
// ------------------------------------------------------------------------------------------------
namespace rba {
__global__ void __launch_bounds__(256)
ood_pred_resize_kernel(const float* __restrict__ logits, int B, int h, int w, int H, int W, float sy, float sx,
                       float* __restrict__ ood_pred, float* __restrict__ score) {
  const int64_t total = (int64_t)B * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % W);
    const int Y = (int)((i / W) % H);
    const int64_t b = i / ((int64_t)W * H);
    const float fy = sy * (float)Y, fx = sx * (float)X;                 // align_corners=True: src = dst * (in-1)/(out-1)
    int y0 = (int)fy, x0 = (int)fx;
    y0 = y0 > h - 1 ? h - 1 : y0;
    x0 = x0 > w - 1 ? w - 1 : x0;
    const int y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float2* base = reinterpret_cast<const float2*>(logits) + b * (int64_t)h * w;
    const float2 v00 = base[(int64_t)y0 * w + x0], v01 = base[(int64_t)y0 * w + x1];
    const float2 v10 = base[(int64_t)y1 * w + x0], v11 = base[(int64_t)y1 * w + x1];
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const float l0 = w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x;
    const float l1 = w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y;
    if (ood_pred) {
      ood_pred[(b * 2 + 0) * (int64_t)H * W + (int64_t)Y * W + X] = l0;
      ood_pred[(b * 2 + 1) * (int64_t)H * W + (int64_t)Y * W + X] = l1;
    }
    if (score) {
      const float p2 = 1.0f / (1.0f + expf(l0 - l1));                   // softmax over the two planes, plane 1
      score[i] += logf(p2 + 1e-9f);
    }
  }
}

int ood_pred_resize(const float* logits, int B, int h, int w, int H, int W, float* ood_pred, float* score, cudaStream_t st) {
  RBA_CHECK(logits && (ood_pred || score), "ood_pred_resize: null pointer");
  RBA_CHECK(B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "ood_pred_resize: bad shape");
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const int64_t total = (int64_t)B * H * W;
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(total, 256), (int64_t)148 * 16);   // grid-stride, 16 CTAs per SM
  ood_pred_resize_kernel<<<grid, 256, 0, st>>>(logits, B, h, w, H, W, sy, sx, ood_pred, score);
  RBA_LAUNCHED();
  return RBA_OK;
}
}  // namespace rba

extern "C" int rba_k_ood_pred_resize(const float* logits, int B, int h, int w, int H, int W, float* ood_pred, float* score,
                                     void* stream) {
  return rba::ood_pred_resize(logits, B, h, w, H, W, ood_pred, score, (cudaStream_t)stream);
}
