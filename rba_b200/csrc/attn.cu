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
// Swin window attention on the legacy tensor path (mma.sync m16n8k16, bf16x3 split precision)
// ------------------------------------------------------------------------------------------------
// Same contract as window_attn_kernel, but q/k/v arrive as bf16 split planes (written by the QKV GEMM epilogue) and
// both contractions run on tensor cores: S = q k^T and O = P v each as hi*hi + hi*lo + lo*hi with fp32 accumulation.
// 144 = 9 x 16 query rows = 18 x 8 key columns, so the warp-level m16n8k16 shape tiles a 12x12 window exactly (the
// tcgen05 shapes, M in {64,128,256}, do not).  One CTA per (window, head), 9 warps, warp w owns query rows 16w..16w+15:
// S fragments (72 fp32/thread) stay in registers through bias/mask/softmax and are re-packed in place as the A operand
// of P.v (FlashAttention-2 register reuse).  K and V^T fragments come from shared memory via ldmatrix(.trans).
constexpr int WM_PITCH = 40;                         // bf16 per smem row (32 used): 80 B pitch, conflict-free ldmatrix
constexpr int WM_PLANE = WA_N * WM_PITCH;            // one 144 x 32 operand plane
constexpr int WM_THREADS = 288;
constexpr int WM_BIAS_PITCH = 532;                   // floats per head in the prepared bias table (529 padded to 16 bytes)
#ifndef WM_CTAS_PER_SM
#define WM_CTAS_PER_SM 2
#endif

__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void ldsm_x4_a(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t_a(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ int lds_i32(uint32_t addr) {
  int v;
  asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// Per-lane shared-memory BYTE addresses (32-bit shared space; everything else is a compile-time immediate):
//   ka  K_hi plane + this lane's ldmatrix row/column; K_lo = + WM_PLANE * 2
//   va  V_hi plane + this lane's transposed-ldmatrix row/column; V_lo = + WM_PLANE * 2
//   ca  column table entry of key column 2 tq (bias byte offset); region ids at + WA_N * 4
struct WmAddr { uint32_t ka, va, ca; };

// One key chunk [NB0*8, (NB0+NBN)*8) of the online-softmax attention for this warp's 16 query rows.
// m / l: running row max (log2 domain) and row sum for rows g and g+8; oacc: unnormalised output accumulators.
template <int NB0, int NBN, bool MASK>
__device__ __forceinline__ void wm_chunk(const WmAddr& ad, const uint32_t (&sBrow)[2], const uint32_t (&qh)[2][4],
                                         const uint32_t (&ql)[2][4], const int (&rid)[2], float scale2, float (&m)[2],
                                         float (&l)[2], float (&oacc)[4][4]) {
  static_assert(NBN % 2 == 0, "chunks are whole k16 steps");
  float sacc[NBN][4];
#pragma unroll
  for (int nb = 0; nb < NBN; ++nb) {
    sacc[nb][0] = sacc[nb][1] = sacc[nb][2] = sacc[nb][3] = 0.f;
    uint32_t kh[4], kl[4];
    ldsm_x4_a(kh[0], kh[1], kh[2], kh[3], ad.ka + (NB0 + nb) * 8 * WM_PITCH * 2);
    ldsm_x4_a(kl[0], kl[1], kl[2], kl[3], ad.ka + (NB0 + nb) * 8 * WM_PITCH * 2 + WM_PLANE * 2);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      mma_bf16(sacc[nb], qh[ks], kh[2 * ks], kh[2 * ks + 1]);
      mma_bf16(sacc[nb], qh[ks], kl[2 * ks], kl[2 * ks + 1]);
      mma_bf16(sacc[nb], ql[ks], kh[2 * ks], kh[2 * ks + 1]);
    }
  }
  // scale + relative position bias (+ shift mask), all in the log2 domain; online softmax update.
  // bias: table entry (ri - ci + 11) * 23 + (rj - cj + 11) = this row's base minus the column's byte offset
#pragma unroll
  for (int nb = 0; nb < NBN; ++nb)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t centry = ad.ca + ((NB0 + nb) * 8 + j) * 4;
      const uint32_t coff = (uint32_t)lds_i32(centry);
      int creg = 0;
      if (MASK) creg = lds_i32(centry + WA_N * 4);
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        float bias;
        asm("ld.shared.f32 %0, [%1];" : "=f"(bias) : "r"(sBrow[hrow] - coff));
        float sv = fmaf(sacc[nb][hrow * 2 + j], scale2, bias);
        if (MASK && rid[hrow] != creg) sv += -100.0f * 1.4426950408889634f;
        sacc[nb][hrow * 2 + j] = sv;
      }
    }
#pragma unroll
  for (int hrow = 0; hrow < 2; ++hrow) {
    float mx = m[hrow];
#pragma unroll
    for (int nb = 0; nb < NBN; ++nb) mx = fmaxf(mx, fmaxf(sacc[nb][hrow * 2], sacc[nb][hrow * 2 + 1]));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float alpha = ex2_approx(m[hrow] - mx);               // 0 on the first chunk (m = -inf)
    float sum = 0.f;
#pragma unroll
    for (int nb = 0; nb < NBN; ++nb)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float e = ex2_approx(sacc[nb][hrow * 2 + j] - mx);
        sacc[nb][hrow * 2 + j] = e;
        sum += e;
      }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    l[hrow] = l[hrow] * alpha + sum;
    m[hrow] = mx;
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) { oacc[nd][hrow * 2] *= alpha; oacc[nd][hrow * 2 + 1] *= alpha; }
  }
  // O += P v
#pragma unroll
  for (int kk = 0; kk < NBN / 2; ++kk) {
    uint32_t ph[4], pl[4];
    split_pack2(sacc[2 * kk][0], sacc[2 * kk][1], ph[0], pl[0]);          // row g,   keys 16kk + 2t, +1
    split_pack2(sacc[2 * kk][2], sacc[2 * kk][3], ph[1], pl[1]);          // row g+8
    split_pack2(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1], ph[2], pl[2]);  // row g,   keys 16kk + 8 + 2t, +1
    split_pack2(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3], ph[3], pl[3]);  // row g+8
#pragma unroll
    for (int np = 0; np < 2; ++np) {                                     // pairs of 8-wide d blocks
      uint32_t vh[4], vl[4];
      const uint32_t voff = ((NB0 / 2 + kk) * 16 * WM_PITCH + np * 16) * 2;
      ldsm_x4_t_a(vh[0], vh[1], vh[2], vh[3], ad.va + voff);
      ldsm_x4_t_a(vl[0], vl[1], vl[2], vl[3], ad.va + voff + WM_PLANE * 2);
#pragma unroll
      for (int q2 = 0; q2 < 2; ++q2) {
        float* o = oacc[np * 2 + q2];
        mma_bf16(o, ph, vh[2 * q2], vh[2 * q2 + 1]);
        mma_bf16(o, ph, vl[2 * q2], vl[2 * q2 + 1]);
        mma_bf16(o, pl, vh[2 * q2], vh[2 * q2 + 1]);
      }
    }
  }
}

__global__ void __launch_bounds__(WM_THREADS, WM_CTAS_PER_SM)
window_attn_mma_kernel(const uint16_t* __restrict__ qkv_hi, const uint16_t* __restrict__ qkv_lo,
                       const float* __restrict__ bias_table, const float* __restrict__ bias_t, int C, int heads, int nWh,
                       int nWw, int shift, float scale, uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo) {
  extern __shared__ __align__(16) uint8_t wm_smem[];
  // planes: 0 q_hi, 1 q_lo, 2 k_hi, 3 k_lo, 4 v_hi, 5 v_lo
  uint16_t* sOp = reinterpret_cast<uint16_t*>(wm_smem);
  float* sB = reinterpret_cast<float*>(wm_smem + 6 * WM_PLANE * 2);     // [529] relative position bias * log2(e)
  int* sCol = reinterpret_cast<int*>(sB + 532);                          // [144] byte offset 4 * (ci*23 + cj) into the bias table
  int* sReg = sCol + WA_N;                                               // [144] shift-mask region id of the column
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // heads fastest: the CTAs of one window run together, so the 128-byte lines that hold two heads' 64-byte q/k/v
  // segments are fetched from HBM once
  const int head = blockIdx.x % heads;
  const int64_t win = blockIdx.x / heads;
  const int ww = (int)(win % nWw), wh = (int)((win / nWw) % nWh);
  const int Hp = nWh * WA_WS, Wp = nWw * WA_WS;
  constexpr float LOG2E = 1.4426950408889634f;

  // ---- stage q/k/v planes: 144 rows x 64 B per operand plane = 576 16-byte chunks = 2 per thread and plane; the
  // (row, chunk) of a thread is the same for all six planes, so the address arithmetic is done twice, not 3456 / 288 times ----
  {
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sOp);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int idx = tid + half * WM_THREADS;                 // 0..575
      const int r = idx >> 2, chunk = idx & 3;
      const int64_t goff = (win * WA_N + r) * (int64_t)(3 * C) + head * WA_D + chunk * 8;
      const uint32_t soff = sbase + (r * WM_PITCH + chunk * 8) * 2;
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(soff + (2 * part) * WM_PLANE * 2), "l"(qkv_hi + goff + part * C));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(soff + (2 * part + 1) * WM_PLANE * 2), "l"(qkv_lo + goff + part * C));
      }
    }
    asm volatile("cp.async.commit_group;" ::);
  }
  // relative-position bias of this head, in the log2 domain.  Prepared form (engine): [heads][WM_BIAS_PITCH] already scaled
  // by log2(e), contiguous per head -> 133 16-byte cp.async in the same group as the operand planes.  Raw form (the
  // reference's (529, heads) table, per-kernel entry point): strided loads, which stall every CTA for a global round trip
  // (ncu: 16 % of the kernel's stall samples sat on this loop).
  if (bias_t) {
    for (int c = tid; c < WM_BIAS_PITCH / 4; c += WM_THREADS) cp_async16(sB + c * 4, bias_t + (int64_t)head * WM_BIAS_PITCH + c * 4);
    asm volatile("cp.async.commit_group;" ::);
  } else {
    for (int e = tid; e < 23 * 23; e += WM_THREADS) sB[e] = bias_table[(int64_t)e * heads + head] * LOG2E;
  }
  for (int c = tid; c < WA_N; c += WM_THREADS) {
    const int ci = c / WA_WS, cj = c - ci * WA_WS;
    const int hs = wh * WA_WS + ci, wsx = ww * WA_WS + cj;
    const int rh = hs < Hp - WA_WS ? 0 : (hs < Hp - shift ? 1 : 2);
    const int rw = wsx < Wp - WA_WS ? 0 : (wsx < Wp - shift ? 1 : 2);
    sCol[c] = 4 * (ci * 23 + cj);
    sReg[c] = rh * 3 + rw;
  }
  asm volatile("cp.async.wait_group 0;" ::);
  __syncthreads();
  if (scale < 0.f) {                                   // RBA_WA_DEBUG=1 (profiling aid): loads + tables only, no attention
    if (tid == 0) out_hi[win * WA_N * (int64_t)C + head * WA_D] = sOp[0];
    return;
  }

  const uint16_t* sQh = sOp, *sQl = sOp + WM_PLANE;
  WmAddr ad;
  {
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sOp);
    ad.ka = sbase + (2 * WM_PLANE + (lane & 7) * WM_PITCH + (lane >> 3) * 8) * 2;
    ad.va = sbase + (4 * WM_PLANE + ((lane & 7) + ((lane >> 3) & 1) * 8) * WM_PITCH + (lane >> 4) * 8) * 2;
    ad.ca = (uint32_t)__cvta_generic_to_shared(sCol) + (lane & 3) * 2 * 4;
  }
  const int g = lane >> 2, tq = lane & 3;
  const int r0 = warp * 16;

  // ---- Q fragments (2 k-steps x hi/lo) ----
  uint32_t qh[2][4], ql[2][4];
  {
    const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int col = (lane >> 4) * 8;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      ldsm_x4(qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], sQh + row * WM_PITCH + ks * 16 + col);
      ldsm_x4(ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3], sQl + row * WM_PITCH + ks * 16 + col);
    }
  }
  int rid[2];
  uint32_t sBrow[2];                                  // shared-space byte address of this row's bias-table base
#pragma unroll
  for (int hrow = 0; hrow < 2; ++hrow) {
    const int r = r0 + g + hrow * 8;
    const int ri = r / WA_WS, rj = r - ri * WA_WS;
    sBrow[hrow] = (uint32_t)__cvta_generic_to_shared(sB + (ri * 23 + rj + (WA_WS - 1) * 23 + (WA_WS - 1)));
    const int hs = wh * WA_WS + ri, wsx = ww * WA_WS + rj;
    rid[hrow] = (hs < Hp - WA_WS ? 0 : (hs < Hp - shift ? 1 : 2)) * 3 + (wsx < Wp - WA_WS ? 0 : (wsx < Wp - shift ? 1 : 2));
  }
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  float oacc[4][4];
#pragma unroll
  for (int nd = 0; nd < 4; ++nd) oacc[nd][0] = oacc[nd][1] = oacc[nd][2] = oacc[nd][3] = 0.f;
  const float scale2 = scale * LOG2E;
  // only windows in the last window row / column see more than one shift region (swin.py:416-431)
  const bool need_mask = shift > 0 && (wh == nWh - 1 || ww == nWw - 1);
  // three key chunks of 48 keys (6 n-blocks): 24 score registers per chunk keep the kernel at 96 registers without spills
  if (need_mask) {
    wm_chunk<0, 6, true>(ad, sBrow, qh, ql, rid, scale2, m, l, oacc);
    wm_chunk<6, 6, true>(ad, sBrow, qh, ql, rid, scale2, m, l, oacc);
    wm_chunk<12, 6, true>(ad, sBrow, qh, ql, rid, scale2, m, l, oacc);
  } else {
    wm_chunk<0, 6, false>(ad, sBrow, qh, ql, rid, scale2, m, l, oacc);
    wm_chunk<6, 6, false>(ad, sBrow, qh, ql, rid, scale2, m, l, oacc);
    wm_chunk<12, 6, false>(ad, sBrow, qh, ql, rid, scale2, m, l, oacc);
  }
  // ---- store (attn @ v).transpose(1,2).reshape(B_, N, C) as split planes ----
#pragma unroll
  for (int hrow = 0; hrow < 2; ++hrow) {
    const float inv = 1.0f / l[hrow];
    const int64_t orow = win * WA_N + r0 + g + hrow * 8;
    const int64_t ob = orow * C + head * WA_D + tq * 2;
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) {
      uint32_t hi, lo;
      split_pack2(oacc[nd][hrow * 2] * inv, oacc[nd][hrow * 2 + 1] * inv, hi, lo);
      *reinterpret_cast<uint32_t*>(out_hi + ob + nd * 8) = hi;
      *reinterpret_cast<uint32_t*>(out_lo + ob + nd * 8) = lo;
    }
  }
}

__global__ void window_attn_prepare_bias_kernel(const float* __restrict__ table, int heads, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= heads * WM_BIAS_PITCH) return;
  const int h = i / WM_BIAS_PITCH, e = i - h * WM_BIAS_PITCH;
  out[i] = e < 23 * 23 ? table[(int64_t)e * heads + h] * 1.4426950408889634f : 0.f;
}

int window_attn_bias_floats(int heads) { return heads * WM_BIAS_PITCH; }

int window_attn_prepare_bias(const float* table, int heads, float* out, cudaStream_t st) {
  RBA_CHECK(table && out && heads > 0, "window_attn_prepare_bias: bad arguments");
  window_attn_prepare_bias_kernel<<<(unsigned)cdiv(heads * WM_BIAS_PITCH, 256), 256, 0, st>>>(table, heads, out);
  return RBA_OK;
}

int window_attn_planes(const uint16_t* qkv_hi, const uint16_t* qkv_lo, const float* bias_table, const float* bias_prepared,
                       int B, int H, int W, int C, int heads, int ws, int shift, uint16_t* out_hi, uint16_t* out_lo,
                       cudaStream_t st) {
  RBA_CHECK(qkv_hi && qkv_lo && (bias_table || bias_prepared) && out_hi && out_lo, "window_attn_planes: null pointer");
  RBA_CHECK(ws == WA_WS, "window_attn_planes: only window_size 12 is built (got %d)", ws);
  RBA_CHECK(heads > 0 && C == heads * WA_D, "window_attn_planes: head_dim must be 32 (C=%d heads=%d)", C, heads);
  RBA_CHECK(shift >= 0 && shift < ws, "window_attn_planes: bad shift %d", shift);
  SwinGeom g = make_swin_geom(H, W, ws, shift);
  const int64_t nwin = (int64_t)B * g.nWh * g.nWw;
  if (nwin == 0) return RBA_OK;
  RBA_CHECK(nwin < (1LL << 31), "window_attn_planes: too many windows");
  const size_t smem = (size_t)6 * WM_PLANE * 2 + 532 * 4 + 2 * WA_N * 4;
  RBA_CUDA(cudaFuncSetAttribute(window_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RBA_CHECK(nwin * heads < (1LL << 31), "window_attn_planes: grid too large");
  dim3 grid((unsigned)(nwin * heads));
  static const bool dbg_loads_only = []() { const char* e = getenv("RBA_WA_DEBUG"); return e && e[0] == '1'; }();
  window_attn_mma_kernel<<<grid, WM_THREADS, smem, st>>>(qkv_hi, qkv_lo, bias_table, bias_prepared, C, heads, g.nWh, g.nWw, shift,
                                                         (dbg_loads_only ? -1.0f : 1.0f) / sqrtf((float)WA_D), out_hi, out_lo);
  RBA_LAUNCHED();
  return RBA_OK;
}

// ------------------------------------------------------------------------------------------------
// Decoder multi-head attention core (nn.MultiheadAttention, mask2former_transformer_decoder.py:52-53,110-113)
// ------------------------------------------------------------------------------------------------
// Q = 100 queries attend to Lk keys (100 for self-attention, 2048 at 1dl, up to 32768 per image for the 3-level decoder);
// head_dim = 32.  One CTA per (b, head, key split): a thread owns one query (its q row, running max / sum and the 32 output
// accumulators live in registers) and the CTA streams its key range in tiles of 32 keys through a double-buffered
// cp.async ring, so every K / V row is fetched ONCE per (b, head) and broadcast from shared memory to all queries (the
// first version gave each (b, head, query) its own warp and re-read all keys per query: 100x the L2 traffic, 32 ms per
// forward at Lk = 32768).  Key splits write un-normalised partials (m, l, acc[32]); a second kernel merges them.
constexpr int MH_TK = 32;                  // keys per tile
constexpr int MH_THREADS = 128;            // queries per CTA
constexpr int MH_PART = 34;                // floats per partial: m, l, acc[32]

__global__ void __launch_bounds__(MH_THREADS)
mha_split_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                 const uint8_t* __restrict__ mask, int Lq, int Lk, int heads, int64_t ldq, int64_t ldk, int64_t ldv,
                 float scale, int keys_per_split, int nsplit, float* __restrict__ part) {
  __shared__ __align__(16) float sK[2][MH_TK * 32];
  __shared__ __align__(16) float sV[2][MH_TK * 32];
  const int tid = threadIdx.x;
  const int split = blockIdx.x % nsplit;
  const int bh = blockIdx.x / nsplit;
  const int head = bh % heads, b = bh / heads;
  const int qi = blockIdx.y * MH_THREADS + tid;
  const bool qok = qi < Lq;
  const int j_begin = split * keys_per_split;
  const int j_end = min(Lk, j_begin + keys_per_split);
  const int ntiles = (j_end - j_begin + MH_TK - 1) / MH_TK;

  float qr[32];
  {
    const float* qp = q + ((int64_t)b * Lq + (qok ? qi : 0)) * ldq + head * 32;
#pragma unroll
    for (int d4 = 0; d4 < 8; ++d4) {
      const float4 t = *reinterpret_cast<const float4*>(qp + d4 * 4);
      qr[d4 * 4] = t.x * scale; qr[d4 * 4 + 1] = t.y * scale; qr[d4 * 4 + 2] = t.z * scale; qr[d4 * 4 + 3] = t.w * scale;
    }
  }
  const float* kb = k + (int64_t)b * Lk * ldk + head * 32;
  const float* vb = v + (int64_t)b * Lk * ldv + head * 32;
  const uint8_t* mrow = (mask && qok) ? mask + ((int64_t)b * Lq + qi) * Lk : nullptr;

  // tile loader: 32 keys x 32 floats for K and for V = 2 x 256 16-byte chunks, two of each per thread
  auto load_tile = [&](int t, int buf) {
    const int j0 = j_begin + t * MH_TK;
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int c = tid + h2 * MH_THREADS;               // 0..255
      const int r = c >> 3, ch = c & 7;
      const int j = min(j0 + r, Lk - 1);                 // rows past the end are clamped (their scores are discarded)
      cp_async16(&sK[buf][r * 32 + ch * 4], kb + (int64_t)j * ldk + ch * 4);
      cp_async16(&sV[buf][r * 32 + ch * 4], vb + (int64_t)j * ldv + ch * 4);
    }
    asm volatile("cp.async.commit_group;" ::);
  };

  float m = -INFINITY, l = 0.f;
  float acc[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) acc[d] = 0.f;
  if (ntiles > 0) load_tile(0, 0);
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {
      load_tile(t + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::);
    } else {
      asm volatile("cp.async.wait_group 0;" ::);
    }
    __syncthreads();
    const int j0 = j_begin + t * MH_TK;
    // this query's mask bytes for the tile (two 16-byte loads when aligned, else bytewise)
    uint32_t blocked = 0;                                // bit i: key j0 + i is masked or out of range
    {
      const int nk = min(MH_TK, j_end - j0);
      if (nk < MH_TK) blocked = 0xffffffffu << nk;
      if (mrow) {
        if ((((uintptr_t)(mrow + j0)) & 15) == 0 && nk == MH_TK) {
          const uint4 m0 = *reinterpret_cast<const uint4*>(mrow + j0), m1 = *reinterpret_cast<const uint4*>(mrow + j0 + 16);
          const uint32_t w[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int bb = 0; bb < 4; ++bb)
              if ((w[i] >> (8 * bb)) & 0xffu) blocked |= 1u << (4 * i + bb);
        } else {
          for (int i = 0; i < nk; ++i)
            if (mrow[j0 + i]) blocked |= 1u << i;
        }
      }
    }
    if (blocked != 0xffffffffu) {
      float sc[MH_TK];
      float tmax = -INFINITY;
#pragma unroll
      for (int i = 0; i < MH_TK; ++i) {
        const float4* kr = reinterpret_cast<const float4*>(&sK[buf][i * 32]);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < 8; d4 += 2) {
          const float4 a4 = kr[d4], b4 = kr[d4 + 1];
          d0 = fmaf(qr[d4 * 4], a4.x, d0); d0 = fmaf(qr[d4 * 4 + 1], a4.y, d0);
          d0 = fmaf(qr[d4 * 4 + 2], a4.z, d0); d0 = fmaf(qr[d4 * 4 + 3], a4.w, d0);
          d1 = fmaf(qr[d4 * 4 + 4], b4.x, d1); d1 = fmaf(qr[d4 * 4 + 5], b4.y, d1);
          d1 = fmaf(qr[d4 * 4 + 6], b4.z, d1); d1 = fmaf(qr[d4 * 4 + 7], b4.w, d1);
        }
        sc[i] = ((blocked >> i) & 1u) ? -INFINITY : d0 + d1;
        tmax = fmaxf(tmax, sc[i]);
      }
      const float mn = fmaxf(m, tmax);
      const float corr = (m == -INFINITY) ? 0.f : expf(m - mn);
      l *= corr;
#pragma unroll
      for (int d = 0; d < 32; ++d) acc[d] *= corr;
#pragma unroll
      for (int i = 0; i < MH_TK; ++i) {
        const float pv = (sc[i] == -INFINITY) ? 0.f : expf(sc[i] - mn);
        l += pv;
        const float4* vr = reinterpret_cast<const float4*>(&sV[buf][i * 32]);
#pragma unroll
        for (int d4 = 0; d4 < 8; ++d4) {
          const float4 t4 = vr[d4];
          acc[d4 * 4] = fmaf(pv, t4.x, acc[d4 * 4]); acc[d4 * 4 + 1] = fmaf(pv, t4.y, acc[d4 * 4 + 1]);
          acc[d4 * 4 + 2] = fmaf(pv, t4.z, acc[d4 * 4 + 2]); acc[d4 * 4 + 3] = fmaf(pv, t4.w, acc[d4 * 4 + 3]);
        }
      }
      m = mn;
    }
    __syncthreads();                                     // the other buffer is overwritten by the next iteration's load
  }
  if (qok) {
    float* pp = part + ((((int64_t)b * heads + head) * Lq + qi) * nsplit + split) * MH_PART;
    pp[0] = m; pp[1] = l;
#pragma unroll
    for (int d = 0; d < 32; ++d) pp[2 + d] = acc[d];
  }
}

// merge the key splits of one (b, head, query): lane <-> output dim
__global__ void __launch_bounds__(256)
mha_combine_kernel(const float* __restrict__ part, int64_t rows /*B*heads*Lq*/, int Lq, int heads, int E, int nsplit,
                   uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* pp = part + r * nsplit * MH_PART;
  float m = -INFINITY;
  for (int s0 = 0; s0 < nsplit; ++s0) m = fmaxf(m, pp[s0 * MH_PART]);
  float l = 0.f, acc = 0.f;
  for (int s0 = 0; s0 < nsplit; ++s0) {
    const float ms = pp[s0 * MH_PART];
    if (ms == -INFINITY) continue;
    const float w = expf(ms - m);
    l = fmaf(w, pp[s0 * MH_PART + 1], l);
    acc = fmaf(w, pp[s0 * MH_PART + 2 + lane], acc);
  }
  const float o = (l > 0.f) ? acc / l : 0.f;
  const int qi = (int)(r % Lq);
  const int head = (int)((r / Lq) % heads);
  const int64_t b = r / ((int64_t)Lq * heads);
  store_split1(out_hi, out_lo, (b * Lq + qi) * E + head * 32 + lane, o);
}

static int mha_splits(int Lk, int* keys_per_split) {
  int kps = 512;
  while ((Lk + kps - 1) / kps > 32) kps *= 2;           // at most 32 splits
  *keys_per_split = kps;
  return (Lk + kps - 1) / kps;
}

int64_t mha_workspace_floats(int B, int Lq, int Lk, int heads) {
  int kps;
  const int ns = mha_splits(Lk, &kps);
  return (int64_t)B * heads * Lq * ns * MH_PART;
}

int mha(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, const uint8_t* mask, int B,
        int Lq, int Lk, int E, int heads, uint16_t* out_hi, uint16_t* out_lo, float* workspace, cudaStream_t st) {
  RBA_CHECK(q && k && v && out_hi && out_lo && workspace, "mha: null pointer");
  RBA_CHECK(heads > 0 && E == heads * 32, "mha: head_dim must be 32 (E=%d heads=%d)", E, heads);
  RBA_CHECK(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0, "mha: pitches must be multiples of 4");
  RBA_CHECK((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) & 15) == 0, "mha: q/k/v must be 16-byte aligned");
  const int64_t total = (int64_t)B * heads * Lq;
  if (total == 0 || Lk == 0) return RBA_OK;
  int kps;
  const int ns = mha_splits(Lk, &kps);
  RBA_CHECK((int64_t)B * heads * ns < (1LL << 31), "mha: grid too large");
  const dim3 grid((unsigned)(B * heads * ns), (unsigned)cdiv(Lq, MH_THREADS));
  mha_split_kernel<<<grid, MH_THREADS, 0, st>>>(q, k, v, mask, Lq, Lk, heads, ldq, ldk, ldv, 1.0f / sqrtf(32.0f), kps, ns, workspace);
  RBA_LAUNCHED();
  mha_combine_kernel<<<(unsigned)cdiv(total, 8), 256, 0, st>>>(workspace, total, Lq, heads, E, ns, out_hi, out_lo);
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

extern "C" int rba_k_window_attn_planes(const uint16_t* qkv_hi, const uint16_t* qkv_lo, const float* bias_table, int B, int H,
                                        int W, int C, int heads, int ws, int shift, uint16_t* out_hi, uint16_t* out_lo,
                                        void* stream) {
  return rba::window_attn_planes(qkv_hi, qkv_lo, bias_table, nullptr, B, H, W, C, heads, ws, shift, out_hi, out_lo, (cudaStream_t)stream);
}

extern "C" int64_t rba_k_window_attn_bias_floats(int heads) { return rba::window_attn_bias_floats(heads); }
extern "C" int rba_k_window_attn_prepare_bias(const float* bias_table, int heads, float* prepared, void* stream) {
  return rba::window_attn_prepare_bias(bias_table, heads, prepared, (cudaStream_t)stream);
}
extern "C" int64_t rba_k_mha_workspace_floats(int B, int Lq, int Lk, int heads) { return rba::mha_workspace_floats(B, Lq, Lk, heads); }

extern "C" int rba_k_mha(const float* q, const float* k, const float* v, const uint8_t* mask, int B, int Lq, int Lk, int E,
                         int heads, uint16_t* out_hi, uint16_t* out_lo, float* workspace, void* stream) {
  return rba::mha(q, E, k, E, v, E, mask, B, Lq, Lk, E, heads, out_hi, out_lo, workspace, (cudaStream_t)stream);
}

extern "C" int rba_k_attn_mask(const float* masks, int B, int Q, int h, int w, int th, int tw, uint8_t* out, void* stream) {
  return rba::attn_mask(masks, B, Q, h, w, th, tw, out, (cudaStream_t)stream);
}
