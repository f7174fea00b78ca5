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
  switch (K) {
    case 19: return launch_score<19>(pred_masks, pred_logits, B, Q, h, w, H, W, rba_out, sem_seg, st);
    case 13: return launch_score<13>(pred_masks, pred_logits, B, Q, h, w, H, W, rba_out, sem_seg, st);  // StreetHazards
    case 3: return launch_score<3>(pred_masks, pred_logits, B, Q, h, w, H, W, rba_out, sem_seg, st);    // test size
    default: return fail(RBA_ERR_INVALID, "rba_score_fused: K=%d not instantiated (19, 13, 3)", K);
  }
}
