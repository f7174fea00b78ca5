// Normalisation-family kernels: LayerNorm with Swin row gathers, GroupNorm fused with the FPN
// elementwise tail, patch embedding, fp32 -> split-plane conversion.  All HBM-bound streaming kernels:
// one pass over the input (row cached in registers), 16-byte vector accesses, one warp per row.
#include "common.cuh"

namespace rba {

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 split planes
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ x, int64_t rows, int cols, int64_t ld, uint16_t* __restrict__ hi,
             uint16_t* __restrict__ lo, int64_t ldp) {
  const int c4 = cols >> 2;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = rows * c4;
  for (; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / c4;
    int c = (int)(i - r * c4) * 4;
    float4 v = *reinterpret_cast<const float4*>(x + r * ld + c);
    store_split4(hi, lo, r * ldp + c, v.x, v.y, v.z, v.w);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (swin.py:247, :293 norm2, :334 PatchMerging.norm, :673 out norms; decoder LayerNorms)
// ------------------------------------------------------------------------------------------------
// mode 0: rows 1:1.  mode 1: Swin window gather (pad + roll(-shift) + window_partition, swin.py:250-271),
// padded rows are written as exact zeros (F.pad happens AFTER norm1, swin.py:247-254).
// mode 2: PatchMerging 2x2 gather in the order x0,x1,x2,x3 = (0,0),(1,0),(0,1),(1,1) (swin.py:327-331), LN over 4C.
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int mode,
                 int64_t out_rows, int H, int W, int C, SwinGeom geom, float eps, float* __restrict__ y,
                 uint16_t* __restrict__ y_hi, uint16_t* __restrict__ y_lo) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= out_rows) return;
  const int CO = (mode == 2) ? 4 * C : C;       // normalised width
  const int nv = CO >> 2;                        // float4s per row
  const float* base;
  bool valid = true;
  if (mode == 0) {
    base = x + r * C;
  } else if (mode == 1) {
    int64_t t = swin_row_to_token(geom, r);
    valid = t >= 0;
    base = x + (valid ? t : 0) * C;
  } else {
    const int W2 = W >> 1, H2 = H >> 1;
    int j = (int)(r % W2);
    int64_t tmp = r / W2;
    int i = (int)(tmp % H2);
    int64_t b = tmp / H2;
    base = x + ((b * H + 2 * i) * W + 2 * j) * (int64_t)C;   // segment s -> (dh, dw) = (s & 1, s >> 1)
  }
  if (!valid) {
    for (int v = lane; v < nv; v += 32) {
      if (y) *reinterpret_cast<float4*>(y + r * CO + 4 * v) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y_hi) {
        *reinterpret_cast<uint2*>(y_hi + r * CO + 4 * v) = make_uint2(0u, 0u);
        *reinterpret_cast<uint2*>(y_lo + r * CO + 4 * v) = make_uint2(0u, 0u);
      }
    }
    return;
  }
  float4 cache[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    int v = lane + 32 * k;
    if (v < nv) {
      int e = 4 * v;
      const float* p = base + e;
      if (mode == 2) {
        const int seg = e / C;
        p = base + ((int64_t)(seg & 1) * W + (seg >> 1)) * C + (e - seg * C);
      }
      cache[k] = *reinterpret_cast<const float4*>(p);
      sum += (cache[k].x + cache[k].y) + (cache[k].z + cache[k].w);
    }
  }
  const float mean = warp_sum(sum) / (float)CO;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    int v = lane + 32 * k;
    if (v < nv) {
      float a = cache[k].x - mean, b = cache[k].y - mean, c = cache[k].z - mean, d = cache[k].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)CO + eps);
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    int v = lane + 32 * k;
    if (v < nv) {
      int e = 4 * v;
      float4 g = *reinterpret_cast<const float4*>(gamma + e);
      float4 bt = *reinterpret_cast<const float4*>(beta + e);
      float4 o;
      o.x = (cache[k].x - mean) * rstd * g.x + bt.x;
      o.y = (cache[k].y - mean) * rstd * g.y + bt.y;
      o.z = (cache[k].z - mean) * rstd * g.z + bt.z;
      o.w = (cache[k].w - mean) * rstd * g.w + bt.w;
      if (y) *reinterpret_cast<float4*>(y + r * CO + e) = o;
      if (y_hi) store_split4(y_hi, y_lo, r * CO + e, o.x, o.y, o.z, o.w);
    }
  }
}

// Narrow rows (C <= 128: one float4 per lane covers the row): ROWS rows per warp, all loads issued before any reduction,
// so that each thread keeps ROWS 16-byte loads in flight -- with one row per warp the stage-0 LayerNorms (1 M tokens x
// 128 channels) reached only ~55 % of the HBM peak.  Modes 0 and 1 (mode 2 normalises 4C and uses the general kernel).
template <int ROWS>
__global__ void __launch_bounds__(256)
layernorm_narrow_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int mode,
                        int64_t out_rows, int C, SwinGeom geom, float eps, float* __restrict__ y,
                        uint16_t* __restrict__ y_hi, uint16_t* __restrict__ y_lo) {
  const int lane = threadIdx.x & 31;
  const int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ROWS;
  if (r0 >= out_rows) return;
  const int nv = C >> 2;
  const bool act = lane < nv;
  float4 v[ROWS];
  int state[ROWS];                                   // 0 past the end, 1 padded row (zeros), 2 real row
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    const int64_t r = r0 + i;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    state[i] = 0;
    if (r < out_rows) {
      int64_t t = r;
      if (mode == 1) t = swin_row_to_token(geom, r);
      state[i] = t >= 0 ? 2 : 1;
      if (t >= 0 && act) v[i] = *reinterpret_cast<const float4*>(x + t * C + 4 * lane);
    }
  }
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f), bt = g;
  if (act) {
    g = *reinterpret_cast<const float4*>(gamma + 4 * lane);
    bt = *reinterpret_cast<const float4*>(beta + 4 * lane);
  }
  float mean[ROWS], rstd[ROWS];
#pragma unroll
  for (int i = 0; i < ROWS; ++i) mean[i] = (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int i = 0; i < ROWS; ++i) mean[i] += __shfl_xor_sync(0xffffffffu, mean[i], o);
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    mean[i] /= (float)C;
    const float a = v[i].x - mean[i], b = v[i].y - mean[i], c = v[i].z - mean[i], d = v[i].w - mean[i];
    rstd[i] = act ? (a * a + b * b) + (c * c + d * d) : 0.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int i = 0; i < ROWS; ++i) rstd[i] += __shfl_xor_sync(0xffffffffu, rstd[i], o);
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    if (state[i] == 0 || !act) continue;
    const int64_t o = (r0 + i) * C + 4 * lane;
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (state[i] == 2) {
      const float rs = 1.0f / sqrtf(rstd[i] / (float)C + eps);
      out.x = (v[i].x - mean[i]) * rs * g.x + bt.x;
      out.y = (v[i].y - mean[i]) * rs * g.y + bt.y;
      out.z = (v[i].z - mean[i]) * rs * g.z + bt.z;
      out.w = (v[i].w - mean[i]) * rs * g.w + bt.w;
    }
    if (y) *reinterpret_cast<float4*>(y + o) = out;
    if (y_hi) {
      if (state[i] == 2) store_split4(y_hi, y_lo, o, out.x, out.y, out.z, out.w);
      else { *reinterpret_cast<uint2*>(y_hi + o) = make_uint2(0u, 0u); *reinterpret_cast<uint2*>(y_lo + o) = make_uint2(0u, 0u); }
    }
  }
}

int layernorm(const float* x, const float* gamma, const float* beta, int mode, int B, int H, int W, int C, int ws,
              int shift, float eps, float* y, uint16_t* y_hi, uint16_t* y_lo, cudaStream_t st) {
  RBA_CHECK(x && gamma && beta && (y || y_hi), "layernorm: null pointer");
  RBA_CHECK((y_hi == nullptr) == (y_lo == nullptr), "layernorm: planes must come in pairs");
  RBA_CHECK(C % 4 == 0 && C > 0, "layernorm: C=%d must be a multiple of 4", C);
  int64_t rows;
  SwinGeom g = make_swin_geom(H, W, ws > 0 ? ws : 1, shift);
  if (mode == 0) rows = (int64_t)B * H * W;
  else if (mode == 1) {
    RBA_CHECK(ws > 0 && shift >= 0 && shift < ws, "layernorm: bad window %d shift %d", ws, shift);
    rows = (int64_t)B * g.nWh * g.nWw * ws * ws;
  } else if (mode == 2) {
    RBA_CHECK(H % 2 == 0 && W % 2 == 0, "layernorm: PatchMerging gather needs even H,W (got %d,%d)", H, W);
    rows = (int64_t)B * (H / 2) * (W / 2);
  } else return fail(RBA_ERR_INVALID, "layernorm: bad mode %d", mode);
  if (rows == 0) return RBA_OK;
  const int CO = mode == 2 ? 4 * C : C;
  const int nv = CO / 4;
  const int warps = 8;
  dim3 grid((unsigned)cdiv(rows, warps));
#define RBA_LN(MV) layernorm_kernel<MV><<<grid, warps * 32, 0, st>>>(x, gamma, beta, mode, rows, H, W, C, g, eps, y, y_hi, y_lo)
  if (nv <= 32 && mode != 2 && rows >= 256) {
    constexpr int R = 4;
    layernorm_narrow_kernel<R><<<(unsigned)cdiv(rows, warps * R), warps * 32, 0, st>>>(x, gamma, beta, mode, rows, C, g, eps, y, y_hi, y_lo);
  } else if (nv <= 32) RBA_LN(1);
  else if (nv <= 64) RBA_LN(2);
  else if (nv <= 128) RBA_LN(4);
  else if (nv <= 256) RBA_LN(8);
  else if (nv <= 512) RBA_LN(16);
  else if (nv <= 1024) RBA_LN(32);
  else return fail(RBA_ERR_INVALID, "layernorm: width %d too large", CO);
#undef RBA_LN
  RBA_LAUNCHED();
  return RBA_OK;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm(32, C) on token-major (B, HW, C) tensors (msdeformattn.py:227,234 input_proj; :275-290 FPN norms)
// ------------------------------------------------------------------------------------------------
// Pass 1: per (b, chunk) partial sums of x and x^2 per group, accumulated in double (deterministic:
// fixed chunking, fixed summation order).  Pass 2 (fused apply): finalises mean/rstd from the partials,
// then y = GN(x) [+ bilinear_up(prev)] [relu], msdeformattn.py:356-360.
constexpr int GN_ROWS_PER_CHUNK = 256;

__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ x, int64_t x_bs, int HW, int C, int groups, int nchunks,
                double* __restrict__ part) {
  // grid: (nchunks, B); block: 256 threads. thread -> channel-quad cq = tid % (C/4), row lane rl = tid / (C/4)
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int cq_n = C >> 2;
  const int rows_par = 256 / cq_n;               // rows processed in parallel (C=256 -> 4)
  const int cq = threadIdx.x % cq_n, rl = threadIdx.x / cq_n;
  const int r0 = chunk * GN_ROWS_PER_CHUNK;
  const int r1 = min(r0 + GN_ROWS_PER_CHUNK, HW);
  const int cpg = C / groups;                    // channels per group (8)
  float s = 0.f, q = 0.f;
  if (rl < rows_par) {
    const float* xb = x + (int64_t)b * x_bs + 4 * cq;
    for (int r = r0 + rl; r < r1; r += rows_par) {
      float4 v = *reinterpret_cast<const float4*>(xb + (int64_t)r * C);
      s += (v.x + v.y) + (v.z + v.w);
      q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
  }
  __shared__ float ss[256], sq[256];
  ss[threadIdx.x] = s;
  sq[threadIdx.x] = q;
  __syncthreads();
  // one thread per group sums its quads over all row lanes, in a fixed order
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    const int q0 = g * cpg / 4, qn = cpg / 4;    // quads of this group (cpg multiple of 4)
    double ds = 0.0, dq = 0.0;
    for (int rr = 0; rr < rows_par; ++rr)
      for (int k = 0; k < qn; ++k) {
        ds += (double)ss[rr * cq_n + q0 + k];
        dq += (double)sq[rr * cq_n + q0 + k];
      }
    double* p = part + (((int64_t)b * nchunks + chunk) * groups + g) * 2;
    p[0] = ds;
    p[1] = dq;
  }
}

__global__ void __launch_bounds__(256)
gn_finalize_kernel(const double* __restrict__ part, int nchunks, int groups, int64_t count, float eps,
                   float* __restrict__ mean_rstd) {
  // grid: B; 256 threads = (256 / groups) slices x groups: slice s adds chunks s, s + S, ... in order, then thread g adds the S
  // slice sums in order -- a fixed summation tree, independent of the batch size (deterministic, batch-invariant)
  __shared__ double sh[2][256];
  const int b = blockIdx.x;
  const int S = 256 / groups;
  const int g = threadIdx.x % groups, sl = threadIdx.x / groups;
  double s = 0.0, q = 0.0;
  if (sl < S) {
    for (int c = sl; c < nchunks; c += S) {
      const double* p = part + (((int64_t)b * nchunks + c) * groups + g) * 2;
      s += p[0];
      q += p[1];
    }
  }
  sh[0][threadIdx.x] = s;
  sh[1][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x >= groups) return;
  s = 0.0; q = 0.0;
  for (int k = 0; k < S; ++k) { s += sh[0][k * groups + g]; q += sh[1][k * groups + g]; }
  double mean = s / (double)count;
  double var = q / (double)count - mean * mean;
  if (var < 0.0) var = 0.0;
  mean_rstd[((int64_t)b * groups + g) * 2 + 0] = (float)mean;
  mean_rstd[((int64_t)b * groups + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// PyTorch bilinear source index (align_corners=False, no antialias) for arbitrary sizes.
__device__ __forceinline__ void bilin_coeff(int o, int in, int out, int& i0, int& i1, float& l1) {
  float scale = (float)in / (float)out;
  float src = ((float)o + 0.5f) * scale - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = min(i0 + 1, in - 1);
  l1 = src - (float)i0;
}

// A CTA walks image ROWS (grid-stride over the B * H rows); inside a row one iteration covers 256 / (C/4) consecutive pixels x all
// channel quads, and each thread keeps UNR pixels in flight.  The image index, the row and the vertical interpolation
// coefficients are computed once per row, the horizontal scale once per kernel: the previous form decomposed a flat pixel index
// with two integer divisions and two float divisions per float4 (the one before that with three 64-bit divisions), ~120
// instructions per 16 bytes, and ran at 44 % (with the up-sampled addend) / 62 % of its HBM floor on the FPN levels.
// The arithmetic per element is unchanged.
template <bool PREV, int UNR>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ x, int64_t x_bs, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ mean_rstd, int B, int H, int W, int C,
                int groups, const float* __restrict__ prev, int64_t prev_bs, int hp, int wp, int relu,
                float* __restrict__ y, uint16_t* __restrict__ y_hi, uint16_t* __restrict__ y_lo, int64_t y_bs) {
  const int cq_n = C >> 2;
  const int ppi = 256 / cq_n;                            // pixels per CTA iteration
  const int cq = threadIdx.x % cq_n, pl = threadIdx.x / cq_n;
  if (pl >= ppi) return;
  const int c = 4 * cq;
  const int g = c / (C / groups);
  const float4 gm = *reinterpret_cast<const float4*>(gamma + c);
  const float4 bt = *reinterpret_cast<const float4*>(beta + c);
  const float xscale = PREV ? (float)wp / (float)W : 0.f;   // bilin_coeff's scale (same expression: same bits)
  const int nrows = B * H;
  for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int b = row / H, yh = row - b * H;
    const float2 mr = *reinterpret_cast<const float2*>(mean_rstd + ((int64_t)b * groups + g) * 2);
    const float* xr = x + (int64_t)b * x_bs + (int64_t)yh * W * C + c;
    const int64_t orow = (int64_t)b * y_bs + (int64_t)yh * W * C + c;
    int y0 = 0, y1 = 0; float ly = 0.f;
    const float* pb0 = nullptr; const float* pb1 = nullptr;
    if (PREV) {
      bilin_coeff(yh, hp, H, y0, y1, ly);
      pb0 = prev + (int64_t)b * prev_bs + (int64_t)y0 * wp * C + c;
      pb1 = prev + (int64_t)b * prev_bs + (int64_t)y1 * wp * C + c;
    }
    for (int x0p = pl; x0p < W; x0p += ppi * UNR) {
      float4 v[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int xw = x0p + u * ppi;
        if (xw < W) v[u] = *reinterpret_cast<const float4*>(xr + (int64_t)xw * C);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int xw = x0p + u * ppi;
        if (xw >= W) continue;
        float4 o;
        o.x = (v[u].x - mr.x) * mr.y * gm.x + bt.x;
        o.y = (v[u].y - mr.x) * mr.y * gm.y + bt.y;
        o.z = (v[u].z - mr.x) * mr.y * gm.z + bt.z;
        o.w = (v[u].w - mr.x) * mr.y * gm.w + bt.w;
        if (PREV) {
          // bilin_coeff(xw, wp, W, ...) with the scale hoisted
          float src = ((float)xw + 0.5f) * xscale - 0.5f;
          if (src < 0.f) src = 0.f;
          int x0 = (int)src;
          if (x0 > wp - 1) x0 = wp - 1;
          const int x1 = min(x0 + 1, wp - 1);
          const float lx = src - (float)x0;
          const float4 p00 = *reinterpret_cast<const float4*>(pb0 + (int64_t)x0 * C);
          const float4 p01 = *reinterpret_cast<const float4*>(pb0 + (int64_t)x1 * C);
          const float4 p10 = *reinterpret_cast<const float4*>(pb1 + (int64_t)x0 * C);
          const float4 p11 = *reinterpret_cast<const float4*>(pb1 + (int64_t)x1 * C);
          const float hy = 1.f - ly, hx = 1.f - lx;
          o.x += hy * (hx * p00.x + lx * p01.x) + ly * (hx * p10.x + lx * p11.x);
          o.y += hy * (hx * p00.y + lx * p01.y) + ly * (hx * p10.y + lx * p11.y);
          o.z += hy * (hx * p00.z + lx * p01.z) + ly * (hx * p10.z + lx * p11.z);
          o.w += hy * (hx * p00.w + lx * p01.w) + ly * (hx * p10.w + lx * p11.w);
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        const int64_t ob = orow + (int64_t)xw * C;
        if (y) *reinterpret_cast<float4*>(y + ob) = o;
        if (y_hi) store_split4(y_hi, y_lo, ob, o.x, o.y, o.z, o.w);
      }
    }
  }
}

int64_t groupnorm_ws_doubles(int B, int H, int W, int C, int groups) {
  int64_t nchunks = cdiv((int64_t)H * W, GN_ROWS_PER_CHUNK);
  return (int64_t)B * nchunks * groups * 2 + (int64_t)B * groups;  // partials + (mean,rstd) floats (2 floats = 1 double)
}

// x_bs / prev_bs / y_bs: elements between consecutive images (>= H*W*C), so a level can live inside a
// concatenated multi-level buffer.
int groupnorm(const float* x, int64_t x_bs, const float* gamma, const float* beta, int B, int H, int W, int C, int groups,
              float eps, const float* prev, int64_t prev_bs, int hp, int wp, int relu, float* y, uint16_t* y_hi,
              uint16_t* y_lo, int64_t y_bs, double* ws, cudaStream_t st) {
  RBA_CHECK(x && gamma && beta && ws && (y || y_hi), "groupnorm: null pointer");
  RBA_CHECK(C % groups == 0 && (C / groups) % 4 == 0, "groupnorm: C/groups=%d must be a multiple of 4", C / groups);
  RBA_CHECK(C <= 1024 && (256 % (C / 4)) == 0 && groups <= 256, "groupnorm: C=%d unsupported", C);
  if (B == 0) return RBA_OK;
  const int HW = H * W;
  const int nchunks = (int)cdiv(HW, GN_ROWS_PER_CHUNK);
  double* part = ws;
  float* mean_rstd = reinterpret_cast<float*>(ws + (int64_t)B * nchunks * groups * 2);
  gn_stats_kernel<<<dim3(nchunks, B), 256, 0, st>>>(x, x_bs, HW, C, groups, nchunks, part);
  RBA_LAUNCHED();
  gn_finalize_kernel<<<B, 256, 0, st>>>(part, nchunks, groups, (int64_t)HW * (C / groups), eps, mean_rstd);
  RBA_LAUNCHED();
  RBA_CHECK((int64_t)B * HW < (1LL << 31), "groupnorm: too many pixels");
  int blocks = (int)std::min<int64_t>((int64_t)B * H, 148 * 8);          // grid-stride over image rows
  if (prev)
    gn_apply_kernel<true, 2><<<blocks, 256, 0, st>>>(x, x_bs, gamma, beta, mean_rstd, B, H, W, C, groups, prev, prev_bs, hp, wp,
                                                    relu, y, y_hi, y_lo, y_bs);
  else
    gn_apply_kernel<false, 4><<<blocks, 256, 0, st>>>(x, x_bs, gamma, beta, mean_rstd, B, H, W, C, groups, prev, prev_bs, hp, wp,
                                                     relu, y, y_hi, y_lo, y_bs);
  RBA_LAUNCHED();
  return RBA_OK;
}

// ------------------------------------------------------------------------------------------------
// Patch embedding: normalise + zero pad + 4x4/4 conv + LayerNorm (maskformer_model.py:255-257, swin.py:479-495)
// ------------------------------------------------------------------------------------------------
// Persistent CTAs, one warp per GROUP of 8 horizontally adjacent tokens: the conv weights are staged once per CTA in shared
// memory transposed to [48][C]; lane l owns channels 4l..4l+3 (and 128+4l.. for C > 128), so each weight read is one
// conflict-free 16-byte load reused for the 8 tokens; the 48 (=3*4*4) normalised inputs per token are broadcast through
// shared memory; the LayerNorm is a warp reduction.
template <typename T, int G>   // G = channel groups of 128 (C <= 128*G)
__global__ void __launch_bounds__(256)
patch_embed_kernel(const T* __restrict__ img, int B, int H, int W, int Hp, int Wp, float m0, float m1, float m2, float s0,
                   float s1, float s2, const float* __restrict__ cw, const float* __restrict__ cb,
                   const float* __restrict__ gamma, const float* __restrict__ beta, int C, float* __restrict__ tokens) {
  // A warp takes PE_TOK = 8 horizontally adjacent tokens at a time (round 2: 2 -> 8): per conv tap one 16-byte weight load and two
  // broadcast 16-byte input loads feed 32 FFMA (11 instructions per 8 FFMA before), and the image rows are read 32 pixels wide.
  // The FMA order per output channel is unchanged (bias, then taps 0..47), so the result is bitwise the same.
  constexpr int PE_TOK = 8;
  extern __shared__ __align__(16) float pe_smem[];
  const int CP = 128 * G;              // padded channel count in smem
  float* sW = pe_smem;                 // [48][CP]
  float* sIn = pe_smem + 48 * CP;      // [8 warps][48 taps][PE_TOK tokens]
  for (int e = threadIdx.x; e < 48 * CP; e += blockDim.x) {
    const int k = e / CP, c = e - k * CP;
    sW[e] = (c < C) ? cw[c * 48 + k] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Th = Hp >> 2, Tw = Wp >> 2;
  const int ngx = (Tw + PE_TOK - 1) / PE_TOK;                    // the engine pads to multiples of 32 (Tw % 8 == 0); others: partial last group
  const int64_t ngrp = (int64_t)B * Th * ngx;
  float* in = sIn + warp * (48 * PE_TOK);
  for (int64_t pr = (int64_t)blockIdx.x * 8 + warp; pr < ngrp; pr += (int64_t)gridDim.x * 8) {
    const int tx = (int)(pr % ngx) * PE_TOK;
    int64_t t = pr / ngx;
    const int ty = (int)(t % Th);
    const int b = (int)(t / Th);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 12; ++r) {                               // r = ch * 4 + ky: one image row segment of 32 pixels, lane = pixel
      const int ch = r >> 2, ky = r & 3;
      const int yy = ty * 4 + ky, xx = tx * 4 + lane;
      float v = 0.f;                                             // ImageList pads the NORMALISED image with 0
      if (yy < H && xx < W) {
        const float raw = (float)img[(((int64_t)b * 3 + ch) * H + yy) * W + xx];
        const float mean = ch == 0 ? m0 : (ch == 1 ? m1 : m2);
        const float sd = ch == 0 ? s0 : (ch == 1 ? s1 : s2);
        v = (raw - mean) / sd;
      }
      in[(ch * 16 + ky * 4 + (lane & 3)) * PE_TOK + (lane >> 2)] = v;   // tap ch*16 + ky*4 + kx of token (lane >> 2)
    }
    __syncwarp();
    float acc[PE_TOK][4 * G];
#pragma unroll
    for (int gch = 0; gch < G; ++gch) {
      const int c0 = gch * 128 + 4 * lane;
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 < C) bb = *reinterpret_cast<const float4*>(cb + c0);
#pragma unroll
      for (int tk = 0; tk < PE_TOK; ++tk) { acc[tk][4 * gch] = bb.x; acc[tk][4 * gch + 1] = bb.y; acc[tk][4 * gch + 2] = bb.z; acc[tk][4 * gch + 3] = bb.w; }
    }
#pragma unroll 2
    for (int e = 0; e < 48; ++e) {
      const float4 ia = *reinterpret_cast<const float4*>(in + e * PE_TOK), ib = *reinterpret_cast<const float4*>(in + e * PE_TOK + 4);
      const float iv[PE_TOK] = {ia.x, ia.y, ia.z, ia.w, ib.x, ib.y, ib.z, ib.w};
#pragma unroll
      for (int gch = 0; gch < G; ++gch) {
        const float4 w4 = *reinterpret_cast<const float4*>(sW + e * CP + gch * 128 + 4 * lane);
#pragma unroll
        for (int tk = 0; tk < PE_TOK; ++tk) {
          acc[tk][4 * gch] = fmaf(w4.x, iv[tk], acc[tk][4 * gch]); acc[tk][4 * gch + 1] = fmaf(w4.y, iv[tk], acc[tk][4 * gch + 1]);
          acc[tk][4 * gch + 2] = fmaf(w4.z, iv[tk], acc[tk][4 * gch + 2]); acc[tk][4 * gch + 3] = fmaf(w4.w, iv[tk], acc[tk][4 * gch + 3]);
        }
      }
    }
    float4 g4[G], b4[G];
#pragma unroll
    for (int gch = 0; gch < G; ++gch) {
      const int c0 = gch * 128 + 4 * lane;
      g4[gch] = b4[gch] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 < C) { g4[gch] = *reinterpret_cast<const float4*>(gamma + c0); b4[gch] = *reinterpret_cast<const float4*>(beta + c0); }
    }
#pragma unroll
    for (int tk = 0; tk < PE_TOK; ++tk) {
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4 * G; ++j) {
        const int c = (j >> 2) * 128 + 4 * lane + (j & 3);
        if (c < C) sum += acc[tk][j];
      }
      const float mean = warp_sum(sum) / (float)C;
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < 4 * G; ++j) {
        const int c = (j >> 2) * 128 + 4 * lane + (j & 3);
        if (c < C) { const float d = acc[tk][j] - mean; sq += d * d; }
      }
      const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)C + 1e-5f);
      const int64_t tok = ((int64_t)b * Th + ty) * Tw + tx + tk;
#pragma unroll
      for (int gch = 0; gch < G; ++gch) {
        const int c0 = gch * 128 + 4 * lane;
        if (c0 < C && tx + tk < Tw) {
          float4 o;
          o.x = (acc[tk][4 * gch] - mean) * rstd * g4[gch].x + b4[gch].x;
          o.y = (acc[tk][4 * gch + 1] - mean) * rstd * g4[gch].y + b4[gch].y;
          o.z = (acc[tk][4 * gch + 2] - mean) * rstd * g4[gch].z + b4[gch].z;
          o.w = (acc[tk][4 * gch + 3] - mean) * rstd * g4[gch].w + b4[gch].w;
          *reinterpret_cast<float4*>(tokens + tok * C + c0) = o;
        }
      }
    }
  }
}

int patch_embed(const void* images, int img_dtype, int B, int H, int W, int Hp, int Wp, const float* mean,
                const float* stdv, const float* conv_w, const float* conv_b, const float* gamma, const float* beta, int C,
                float* tokens, cudaStream_t st) {
  RBA_CHECK(images && conv_w && conv_b && gamma && beta && tokens, "patch_embed: null pointer");
  RBA_CHECK(Hp % 8 == 0 && Wp % 8 == 0 && Hp >= H && Wp >= W, "patch_embed: padded size must be a multiple of 8");
  RBA_CHECK(C <= 256 && C % 4 == 0, "patch_embed: C=%d unsupported", C);
  const int64_t ngrp = (int64_t)B * (Hp / 4) * cdiv(Wp / 4, 8);          // groups of 8 tokens, one per warp iteration
  if (ngrp == 0) return RBA_OK;
  dim3 grid((unsigned)std::min<int64_t>(cdiv(ngrp, 8), 148 * 4));
  const int G = C > 128 ? 2 : 1;
  const size_t smem = (size_t)(48 * 128 * G + 8 * 48 * 8) * sizeof(float);
#define RBA_PE(T, GG)                                                                                                   \
  do {                                                                                                                  \
    RBA_CUDA(cudaFuncSetAttribute(patch_embed_kernel<T, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    patch_embed_kernel<T, GG><<<grid, 256, smem, st>>>((const T*)images, B, H, W, Hp, Wp, mean[0], mean[1], mean[2],   \
                                                       stdv[0], stdv[1], stdv[2], conv_w, conv_b, gamma, beta, C, tokens); \
  } while (0)
  if (img_dtype == RBA_IMG_U8) { if (G == 1) RBA_PE(uint8_t, 1); else RBA_PE(uint8_t, 2); }
  else if (img_dtype == RBA_IMG_F32) { if (G == 1) RBA_PE(float, 1); else RBA_PE(float, 2); }
  else return fail(RBA_ERR_INVALID, "patch_embed: bad image dtype %d", img_dtype);
#undef RBA_PE
  RBA_LAUNCHED();
  return RBA_OK;
}

// ------------------------------------------------------------------------------------------------
// Elementwise: y[r,:] = (a ? a[r,:] : 0) + (b ? b[r % period,:] : 0)  -> fp32 and/or split planes
// (with_pos_embed / level_embed adds / query broadcast: msdeformattn.py:84,123; mask2former_transformer_decoder.py:415,422-423)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ew_add_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t rows, int cols, int64_t period,
              float* __restrict__ y, uint16_t* __restrict__ y_hi, uint16_t* __restrict__ y_lo) {
  const int c4 = cols >> 2;
  const int64_t total = rows * c4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c4;
    const int c = (int)(i - r * c4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a) v = *reinterpret_cast<const float4*>(a + r * cols + c);
    if (b) {
      float4 w = *reinterpret_cast<const float4*>(b + (r % period) * cols + c);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    if (y) *reinterpret_cast<float4*>(y + r * cols + c) = v;
    if (y_hi) store_split4(y_hi, y_lo, r * cols + c, v.x, v.y, v.z, v.w);
  }
}

int ew_add(const float* a, const float* b, int64_t rows, int cols, int64_t period, float* y, uint16_t* y_hi,
           uint16_t* y_lo, cudaStream_t st) {
  RBA_CHECK((a || b) && (y || y_hi), "ew_add: null pointer");
  RBA_CHECK(cols % 4 == 0 && period > 0, "ew_add: cols must be a multiple of 4");
  if (rows == 0) return RBA_OK;
  const int64_t total = rows * (cols / 4);
  int blocks = (int)std::min<int64_t>(cdiv(total, 256), 148 * 16);
  ew_add_kernel<<<blocks, 256, 0, st>>>(a, b, rows, cols, period, y, y_hi, y_lo);
  RBA_LAUNCHED();
  return RBA_OK;
}

}  // namespace rba

// ---- C ABI ----
extern "C" int rba_k_split(const float* x, int64_t rows, int cols, int64_t ld, uint16_t* hi, uint16_t* lo, int64_t ldp,
                           void* stream) {
  using namespace rba;
  RBA_CHECK(x && hi && lo, "rba_k_split: null pointer");
  RBA_CHECK(cols % 4 == 0 && ld % 4 == 0 && ldp % 4 == 0, "rba_k_split: cols/ld/ldp must be multiples of 4");
  if (rows == 0 || cols == 0) return RBA_OK;
  int64_t total = rows * (cols / 4);
  int blocks = (int)std::min<int64_t>(cdiv(total, 256), 148 * 16);
  split_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, hi, lo, ldp);
  RBA_LAUNCHED();
  return RBA_OK;
}

extern "C" int rba_k_layernorm(const float* x, const float* gamma, const float* beta, int mode, int B, int H, int W, int C,
                               int ws, int shift, float eps, float* y, uint16_t* y_hi, uint16_t* y_lo, void* stream) {
  return rba::layernorm(x, gamma, beta, mode, B, H, W, C, ws, shift, eps, y, y_hi, y_lo, (cudaStream_t)stream);
}

extern "C" int64_t rba_k_groupnorm_ws(int B, int H, int W, int C, int groups) {
  return rba::groupnorm_ws_doubles(B, H, W, C, groups);
}

extern "C" int rba_k_groupnorm(const float* x, const float* gamma, const float* beta, int B, int H, int W, int C, int groups,
                               float eps, const float* prev, int hp, int wp, int relu, float* y, uint16_t* y_hi,
                               uint16_t* y_lo, double* workspace, void* stream) {
  const int64_t bs = (int64_t)H * W * C;
  return rba::groupnorm(x, bs, gamma, beta, B, H, W, C, groups, eps, prev, (int64_t)hp * wp * C, hp, wp, relu, y, y_hi,
                        y_lo, bs, workspace, (cudaStream_t)stream);
}

extern "C" int rba_k_patch_embed(const void* images, int img_dtype, int B, int H, int W, int Hp, int Wp, const float* mean,
                                 const float* stdv, const float* conv_w, const float* conv_b, const float* gamma,
                                 const float* beta, int C, float* tokens, void* stream) {
  return rba::patch_embed(images, img_dtype, B, H, W, Hp, Wp, mean, stdv, conv_w, conv_b, gamma, beta, C, tokens,
                          (cudaStream_t)stream);
}
