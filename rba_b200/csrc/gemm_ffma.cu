// fp32 FMA-pipe GEMM / implicit-GEMM 3x3 convolution over bf16 split-plane operands.
//
//   C[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ residual), A/W given as (hi, lo) bf16 planes, K contiguous.
//
// This is the exact-arithmetic backend (operands reconstructed as hi+lo in fp32, fp32 FMA accumulate).  It
// serves every Linear / 1x1 conv / einsum of the hot path (swin.py:139,169,36-39,335; msdeformattn.py:226,279,
// 254-260; ms_deform_attn.py:98-104,124; mask2former_transformer_decoder.py:476-479) and, with the gathered A
// loader, the FPN 3x3 output convs (msdeformattn.py:281-290).  The tcgen05 backend (gemm_tc.cu) computes the
// same contract on tensor cores; this kernel is also its on-device cross-check.
//
// Tiling: 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile per thread (two 4-row and two 4-column
// groups 64 apart so shared-memory reads are conflict-free float4s), register-prefetch double buffering.
#include "common.cuh"

namespace rba {

constexpr int GB_M = 128, GB_N = 128, GB_K = 16, GB_LD = GB_M + 4;

struct GemmParams {
  const uint16_t* a_hi; const uint16_t* a_lo; int64_t lda;
  const uint16_t* w_hi; const uint16_t* w_lo; int64_t ldw;
  int M, N, K;
  int64_t a_bs, w_bs;
  const float* bias; int bias_per_row; int64_t bias_bs;
  int act;
  const float* residual;
  float* c; int64_t ldc; int64_t c_bs;
  uint16_t* c_hi; uint16_t* c_lo; int64_t ldcp; int64_t cp_bs;
  int swin_map; SwinGeom geom;
  // conv mode
  int conv; int cH, cW, cCin;
};

__device__ __forceinline__ void unpack8(const uint4& h, const uint4& l, float* f) {
  f[0] = bf16lo(h.x) + bf16lo(l.x); f[1] = bf16hi(h.x) + bf16hi(l.x);
  f[2] = bf16lo(h.y) + bf16lo(l.y); f[3] = bf16hi(h.y) + bf16hi(l.y);
  f[4] = bf16lo(h.z) + bf16lo(l.z); f[5] = bf16hi(h.z) + bf16hi(l.z);
  f[6] = bf16lo(h.w) + bf16lo(l.w); f[7] = bf16hi(h.w) + bf16hi(l.w);
}

template <bool CONV>
__global__ void __launch_bounds__(256, 2)
gemm_ffma_kernel(GemmParams p) {
  __shared__ __align__(16) float As[2][GB_K][GB_LD];
  __shared__ __align__(16) float Bs[2][GB_K][GB_LD];
  const int tid = threadIdx.x;
  const int bz = blockIdx.z;
  const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N;
  const uint16_t* a_hi = p.a_hi + bz * p.a_bs;
  const uint16_t* a_lo = p.a_lo + bz * p.a_bs;
  const uint16_t* w_hi = p.w_hi + bz * p.w_bs;
  const uint16_t* w_lo = p.w_lo + bz * p.w_bs;

  // loader mapping: row = tid & 127, k-half = tid >> 7 (8 consecutive k each)
  const int lrow = tid & 127, lk = (tid >> 7) * 8;
  const int am = m0 + lrow, wn = n0 + lrow;
  const bool a_ok = am < p.M, w_ok = wn < p.N;
  // conv: decompose the output pixel once
  int cb = 0, ch = 0, cw = 0;
  if (CONV && a_ok) {
    cw = am % p.cW;
    int t = am / p.cW;
    ch = t % p.cH;
    cb = t / p.cH;
  }
  const int nk = p.K / GB_K;
  uint4 ra_h, ra_l, rw_h, rw_l;
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);

  auto gload = [&](int kt) {
    const int k0 = kt * GB_K + lk;
    ra_h = z4; ra_l = z4; rw_h = z4; rw_l = z4;
    if (CONV) {
      if (a_ok) {
        const int tap = k0 / p.cCin, c0 = k0 - tap * p.cCin;
        const int hs = ch + tap / 3 - 1, wsx = cw + tap % 3 - 1;
        if (hs >= 0 && hs < p.cH && wsx >= 0 && wsx < p.cW) {
          const int64_t off = (((int64_t)cb * p.cH + hs) * p.cW + wsx) * p.cCin + c0;
          ra_h = *reinterpret_cast<const uint4*>(a_hi + off);
          ra_l = *reinterpret_cast<const uint4*>(a_lo + off);
        }
      }
    } else if (a_ok) {
      const int64_t off = (int64_t)am * p.lda + k0;
      ra_h = *reinterpret_cast<const uint4*>(a_hi + off);
      ra_l = *reinterpret_cast<const uint4*>(a_lo + off);
    }
    if (w_ok) {
      const int64_t off = (int64_t)wn * p.ldw + k0;
      rw_h = *reinterpret_cast<const uint4*>(w_hi + off);
      rw_l = *reinterpret_cast<const uint4*>(w_lo + off);
    }
  };
  auto sstore = [&](int buf) {
    float fa[8], fw[8];
    unpack8(ra_h, ra_l, fa);
    unpack8(rw_h, rw_l, fw);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      As[buf][lk + i][lrow] = fa[i];
      Bs[buf][lk + i][lrow] = fw[i];
    }
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(kt + 1);
#pragma unroll
    for (int k = 0; k < GB_K; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue ----
  const float* bias = p.bias ? p.bias + bz * p.bias_bs : nullptr;
  float* c = p.c ? p.c + bz * p.c_bs : nullptr;
  const int64_t ldr = p.c ? p.ldc : p.ldcp;      // the residual shares the output's indexing
  const float* res = p.residual ? p.residual + bz * (p.c ? p.c_bs : p.cp_bs) : nullptr;
  uint16_t* c_hi = p.c_hi ? p.c_hi + bz * p.cp_bs : nullptr;
  uint16_t* c_lo = p.c_lo ? p.c_lo + bz * p.cp_bs : nullptr;
  const bool vec_c = (p.ldc & 3) == 0, vec_p = (p.ldcp & 3) == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
    int64_t orow = m;
    if (p.swin_map) {
      orow = swin_row_to_token(p.geom, m);
      if (orow < 0) continue;
    }
    const float brow = (bias && p.bias_per_row) ? bias[m] : 0.f;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      if (n >= p.N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = acc[i][jh * 4 + j];
        if (bias) t += p.bias_per_row ? brow : ((n + j < p.N) ? bias[n + j] : 0.f);
        v[j] = apply_act_rt(t, p.act);
      }
      const bool full = n + 3 < p.N;
      if (c) {
        float* cp = c + orow * p.ldc + n;
        if (full && vec_c) {
          if (res) {
            float4 r4 = *reinterpret_cast<const float4*>(res + orow * ldr + n);
            v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
          }
          *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < p.N) {
              if (res) v[j] += res[orow * ldr + n + j];
              cp[j] = v[j];
            }
        }
      } else if (res) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < p.N) v[j] += res[orow * ldr + n + j];
      }
      if (c_hi) {
        if (full && vec_p) {
          store_split4(c_hi, c_lo, orow * p.ldcp + n, v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < p.N) store_split1(c_hi, c_lo, orow * p.ldcp + n + j, v[j]);
        }
      }
    }
  }
}

int gemm_ffma_launch(const GemmParams& p, int batch, cudaStream_t st) {
  dim3 grid((unsigned)cdiv(p.N, GB_N), (unsigned)cdiv(p.M, GB_M), (unsigned)batch);
  if (p.conv) gemm_ffma_kernel<true><<<grid, 256, 0, st>>>(p);
  else gemm_ffma_kernel<false><<<grid, 256, 0, st>>>(p);
  RBA_LAUNCHED();
  return RBA_OK;
}

int gemm_tc_launch(const rba_gemm_args& a, cudaStream_t st);  // gemm_tc.cu
int conv3x3_tc_launch(const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* w_hi, const uint16_t* w_lo, int B, int H,
                      int W, int Cin, int Cout, float* y, cudaStream_t st);  // gemm_tc.cu

int gemm(const rba_gemm_args& a, cudaStream_t st) {
  RBA_CHECK(a.a_hi && a.a_lo && a.w_hi && a.w_lo, "gemm: null operand");
  RBA_CHECK(a.c || a.c_hi, "gemm: no output");
  RBA_CHECK((a.c_hi == nullptr) == (a.c_lo == nullptr), "gemm: output planes must come in pairs");
  RBA_CHECK(a.M >= 0 && a.N > 0 && a.K > 0 && a.batch >= 1, "gemm: bad shape M=%d N=%d K=%d batch=%d", a.M, a.N, a.K, a.batch);
  RBA_CHECK(a.K % 16 == 0, "gemm: K=%d must be a multiple of 16", a.K);
  RBA_CHECK(a.lda % 8 == 0 && a.ldw % 8 == 0 && a.a_bstride % 8 == 0 && a.w_bstride % 8 == 0,
            "gemm: operand pitches must be multiples of 8 elements (16 B)");
  RBA_CHECK(a.lda >= a.K && a.ldw >= a.K, "gemm: pitch smaller than K");
  RBA_CHECK(!(a.residual && a.c && a.c_hi && a.ldc != a.ldcp), "gemm: residual needs one output pitch");
  RBA_CHECK(a.act >= 0 && a.act <= 2, "gemm: bad activation %d", a.act);
  if (a.M == 0) return RBA_OK;
  if (a.backend == RBA_GEMM_TC) return gemm_tc_launch(a, st);
  RBA_CHECK(a.backend == RBA_GEMM_FFMA, "gemm: bad backend %d", a.backend);
  RBA_CHECK(a.qkv_tile_heads == 0, "gemm: the tiled q|k|v plane layout is written by the tensor-core backend only");
  GemmParams p;
  p.a_hi = a.a_hi; p.a_lo = a.a_lo; p.lda = a.lda;
  p.w_hi = a.w_hi; p.w_lo = a.w_lo; p.ldw = a.ldw;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.a_bs = a.a_bstride; p.w_bs = a.w_bstride;
  p.bias = a.bias; p.bias_per_row = a.bias_per_row; p.bias_bs = a.bias_bstride;
  p.act = a.act; p.residual = a.residual;
  p.c = a.c; p.ldc = a.ldc; p.c_bs = a.c_bstride;
  p.c_hi = a.c_hi; p.c_lo = a.c_lo; p.ldcp = a.ldcp; p.cp_bs = a.cp_bstride;
  p.swin_map = a.swin_map;
  p.geom = make_swin_geom(a.sw_H > 0 ? a.sw_H : 1, a.sw_W > 0 ? a.sw_W : 1, a.sw_ws > 0 ? a.sw_ws : 1, a.sw_shift);
  p.conv = 0; p.cH = p.cW = p.cCin = 0;
  return gemm_ffma_launch(p, a.batch, st);
}

int conv3x3(const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* w_hi, const uint16_t* w_lo, int B, int H, int W,
            int Cin, int Cout, float* y, int backend, cudaStream_t st) {
  RBA_CHECK(x_hi && x_lo && w_hi && w_lo && y, "conv3x3: null pointer");
  RBA_CHECK(Cin % 16 == 0 && Cout % 4 == 0, "conv3x3: Cin %% 16 and Cout %% 4 required");
  if ((int64_t)B * H * W == 0) return RBA_OK;
  if (backend == RBA_GEMM_TC) return conv3x3_tc_launch(x_hi, x_lo, w_hi, w_lo, B, H, W, Cin, Cout, y, st);
  GemmParams p;
  p.a_hi = x_hi; p.a_lo = x_lo; p.lda = Cin;
  p.w_hi = w_hi; p.w_lo = w_lo; p.ldw = 9 * Cin;
  p.M = B * H * W; p.N = Cout; p.K = 9 * Cin;
  p.a_bs = p.w_bs = 0;
  p.bias = nullptr; p.bias_per_row = 0; p.bias_bs = 0;
  p.act = RBA_ACT_NONE; p.residual = nullptr;
  p.c = y; p.ldc = Cout; p.c_bs = 0;
  p.c_hi = p.c_lo = nullptr; p.ldcp = 0; p.cp_bs = 0;
  p.swin_map = 0; p.geom = make_swin_geom(1, 1, 1, 0);
  p.conv = 1; p.cH = H; p.cW = W; p.cCin = Cin;
  return gemm_ffma_launch(p, 1, st);
}

}  // namespace rba

extern "C" int rba_k_gemm(const rba_gemm_args* args, void* stream) {
  using namespace rba;
  RBA_CHECK(args, "rba_k_gemm: null args");
  return gemm(*args, (cudaStream_t)stream);
}

extern "C" int rba_k_conv3x3(const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* w_hi, const uint16_t* w_lo, int B,
                             int H, int W, int Cin, int Cout, float* y, int backend, void* stream) {
  return rba::conv3x3(x_hi, x_lo, w_hi, w_lo, B, H, W, Cin, Cout, y, backend, (cudaStream_t)stream);
}
