// Shared device/host helpers for the rba_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/rba_b200.h"

namespace rba {

// ---- error plumbing (thread-local message, int status through the C ABI) ----
std::string& last_error_ref();
int fail(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define RBA_CHECK(cond, ...)                                  \
  do {                                                        \
    if (!(cond)) return ::rba::fail(RBA_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define RBA_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::rba::fail(RBA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// Call right after a kernel launch: counts it and surfaces launch-configuration errors.
#define RBA_LAUNCHED()                                                                                  \
  do {                                                                                                  \
    ::rba::g_launches.fetch_add(1, std::memory_order_relaxed);                                          \
    cudaError_t _e = cudaGetLastError();                                                                \
    if (_e != cudaSuccess)                                                                              \
      return ::rba::fail(RBA_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define RBA_TRY_(expr)             \
  do {                             \
    int _rc = (expr);              \
    if (_rc != RBA_OK) return _rc; \
  } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Once-per-DEVICE guard for cudaFuncSetAttribute (function attributes are per device; a process-wide `static bool`
// leaves the second GPU of a multi-GPU process without its > 48 KB dynamic shared memory opt-in).  Usage:
//   static PerDeviceOnce once; if (once.needed()) { RBA_CUDA(cudaFuncSetAttribute(...)); once.done(); }
// Two racing threads may both set the attribute (harmless); neither launches before it is set.
struct PerDeviceOnce {
  std::atomic<uint64_t> mask{0};
  static uint64_t bit() { int d = 0; cudaGetDevice(&d); return 1ull << (d & 63); }
  bool needed() const { return !(mask.load(std::memory_order_acquire) & bit()); }
  void done() { mask.fetch_or(bit(), std::memory_order_release); }
};

// ---- bf16 split planes ----
// x ~= hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi): |x - hi - lo| <= 2^-18 |x|.
__device__ __forceinline__ void split2(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
__device__ __forceinline__ float bf16lo(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16hi(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }

// Two values -> packed bf16x2 hi word and bf16x2 lo (residual) word, element `a` in the low half.
// cvt.rn.bf16x2.f32 converts and packs a pair in one instruction; the residual x - float(hi) is ONE sm_100 mixed-precision
// FMA per element (fma.rn.f32.bf16, SASS FHFMA.BF16, reading a half of the packed register directly) instead of a shift /
// mask to rebuild float(hi) plus a subtract.  Same bits as the two-step form (both are exact).
__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  float ra, rb;
  const uint16_t m1 = 0xBF80;                              // -1.0 (bf16)
  asm("{.reg .b16 l, h; mov.b32 {l, h}, %2; fma.rn.f32.bf16 %0, l, %3, %4; fma.rn.f32.bf16 %1, h, %3, %5;}"
      : "=f"(ra), "=f"(rb)
      : "r"(hi), "h"(m1), "f"(a), "f"(b));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}

// Stores 4 consecutive values as split planes (8-byte stores, idx must be a multiple of 4).
__device__ __forceinline__ void store_split4(uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int64_t idx,
                                             float a, float b, float c, float d) {
  uint2 H, L;
  split_pack2(a, b, H.x, L.x);
  split_pack2(c, d, H.y, L.y);
  *reinterpret_cast<uint2*>(hi + idx) = H;
  *reinterpret_cast<uint2*>(lo + idx) = L;
}
__device__ __forceinline__ void store_split1(uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int64_t idx, float a) {
  __nv_bfloat16 h, l;
  split2(a, h, l);
  hi[idx] = __bfloat16_as_ushort(h);
  lo[idx] = __bfloat16_as_ushort(l);
}

// ---- warp reductions ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- activations ----
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if (ACT == RBA_ACT_RELU) return fmaxf(x, 0.0f);
  if (ACT == RBA_ACT_GELU) return gelu_erf(x);
  return x;
}
__device__ __forceinline__ float apply_act_rt(float x, int act) {
  if (act == RBA_ACT_RELU) return fmaxf(x, 0.0f);
  if (act == RBA_ACT_GELU) return gelu_erf(x);
  return x;
}

// ---- Swin window geometry (swin.py:250-271, 277-287) ----
// Windowed row r = ((b*nWh + wh)*nWw + ww)*ws*ws + i*ws + j lives at (wh*ws+i, ww*ws+j) of the SHIFTED padded
// frame; shifted = roll(x, -shift) so its source in the padded frame is ((hs+shift)%Hp, (wsft+shift)%Wp).
// Returns the token index b*H*W + h*W + w, or -1 when the source lies in the bottom/right padding.
struct SwinGeom {
  int H, W, ws, shift, nWh, nWw, Hp, Wp;
};
__host__ __device__ inline SwinGeom make_swin_geom(int H, int W, int ws, int shift) {
  SwinGeom g;
  g.H = H; g.W = W; g.ws = ws; g.shift = shift;
  g.nWh = (H + ws - 1) / ws; g.nWw = (W + ws - 1) / ws;
  g.Hp = g.nWh * ws; g.Wp = g.nWw * ws;
  return g;
}
__host__ __device__ inline int64_t swin_row_to_token(const SwinGeom& g, int64_t r) {
  const int n = g.ws * g.ws;
  int64_t win = r / n;
  int t = (int)(r - win * n);
  int i = t / g.ws, j = t - i * g.ws;
  int ww = (int)(win % g.nWw);
  int64_t tmp = win / g.nWw;
  int wh = (int)(tmp % g.nWh);
  int64_t b = tmp / g.nWh;
  int h = wh * g.ws + i + g.shift; if (h >= g.Hp) h -= g.Hp;
  int w = ww * g.ws + j + g.shift; if (w >= g.Wp) w -= g.Wp;
  if (h >= g.H || w >= g.W) return -1;
  return (b * g.H + h) * g.W + w;
}

}  // namespace rba
