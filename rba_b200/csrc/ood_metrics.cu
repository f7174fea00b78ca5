// Streaming OoD metrics on the GPU: AUROC, average precision and FPR@95%TPR over ALL pixels of a dataset,
// without moving score maps to the host and without sorting 10^8..10^9 points.
//
// Replaces OODEvaluator.evaluate_ood / calculate_ood_metrics / calculate_auroc (support.py:247-303), which run
// sklearn.metrics.roc_curve + auc + average_precision_score on the concatenation of every pixel, and the per-image
// `.cpu().numpy()` of compute_anomaly_scores (support.py:353-399).
//
// Method: scores are mapped to their order-preserving 32-bit key (sign-flipped IEEE bits); the top RBA_OOD_KEY_BITS = 24
// bits (sign, exponent, 15 mantissa bits) index a two-class histogram of 64-bit counters.  The metrics are then
// EXACTLY what sklearn computes on scores quantised to 2^-15 relative resolution: one ROC / PR point per non-empty bin,
// ties inside a bin forming one threshold (sklearn's own treatment of tied scores).
//   update   : grid-stride pass over (score, label) with warp-aggregated 64-bit atomics (__match_any_sync), so a saturated
//              score distribution (RbA piles up near -K) does not serialise on one address.  HBM-bound: 4 B + label bytes / px.
//   finalize : three small kernels over the 2 x 2^24 counters — per-chunk sums, a single-block scan of the 4096 chunk sums,
//              per-chunk sweep in descending score order accumulating trapezoid (AUROC) and step (AP) areas in fp64 and
//              locating the first threshold with TPR > 0.95 — then a fixed-order reduction (deterministic result).
// Labels follow the reference: 1 = OoD (positive), 0 = in-distribution, anything else ignored (support.py:275-279).
#include "common.cuh"

namespace rba {

constexpr int OOD_BITS = 24;
constexpr int64_t OOD_NB = (int64_t)1 << OOD_BITS;          // bins per class
constexpr int OOD_CHUNK = 4096;                              // bins per chunk
constexpr int OOD_NCHUNK = (int)(OOD_NB / OOD_CHUNK);        // 4096
constexpr int OOD_TPB = 256;
constexpr int OOD_PER_THREAD = OOD_CHUNK / OOD_TPB;          // 16

__device__ __forceinline__ uint32_t ood_key(float s) {
  const uint32_t u = __float_as_uint(s + 0.0f);               // -0.0 -> +0.0: one key for both zeros, as they compare equal
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);         // monotone: a < b  <=>  key(a) < key(b)
}

template <typename L>
__global__ void __launch_bounds__(256) ood_hist_kernel(const float* __restrict__ score, const L* __restrict__ label, int64_t n,
                                                      unsigned long long* __restrict__ hist) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_round = (n + 31) / 32 * 32;                // whole warps stay converged for __match_any_sync
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    uint32_t slot = 0xffffffffu;                             // invalid: ignored label, NaN score or tail
    if (i < n) {
      const float s = score[i];
      const long long l = (long long)label[i];
      if ((l == 0 || l == 1) && s == s) slot = ((uint32_t)l << OOD_BITS) | (ood_key(s) >> (32 - OOD_BITS));
    }
    const unsigned peers = __match_any_sync(0xffffffffu, slot);
    if (slot != 0xffffffffu && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31))
      atomicAdd(hist + slot, (unsigned long long)__popc(peers));
  }
}

// descending-score position d <-> bin NB-1-d
__global__ void __launch_bounds__(OOD_TPB) ood_chunk_sums_kernel(const unsigned long long* __restrict__ hist,
                                                                unsigned long long* __restrict__ csum) {
  __shared__ unsigned long long sp[OOD_TPB], sn[OOD_TPB];
  const int ch = blockIdx.x, t = threadIdx.x;
  unsigned long long p = 0, q = 0;
  for (int k = 0; k < OOD_PER_THREAD; ++k) {
    const int64_t d = (int64_t)ch * OOD_CHUNK + k * OOD_TPB + t;
    const int64_t b = OOD_NB - 1 - d;
    q += hist[b];
    p += hist[OOD_NB + b];
  }
  sp[t] = p; sn[t] = q;
  __syncthreads();
  for (int o = OOD_TPB / 2; o > 0; o >>= 1) {
    if (t < o) { sp[t] += sp[t + o]; sn[t] += sn[t + o]; }
    __syncthreads();
  }
  if (t == 0) { csum[ch] = sp[0]; csum[OOD_NCHUNK + ch] = sn[0]; }
}

// single block: exclusive prefix over the chunk sums (in place) + totals at [2*NCHUNK], [2*NCHUNK+1]
__global__ void __launch_bounds__(1024) ood_chunk_scan_kernel(unsigned long long* __restrict__ csum) {
  __shared__ unsigned long long sp[1024], sn[1024];
  const int t = threadIdx.x;
  constexpr int PER = OOD_NCHUNK / 1024;
  unsigned long long lp[PER], ln[PER], tp = 0, tn = 0;
  for (int k = 0; k < PER; ++k) {
    lp[k] = csum[t * PER + k]; ln[k] = csum[OOD_NCHUNK + t * PER + k];
    tp += lp[k]; tn += ln[k];
  }
  sp[t] = tp; sn[t] = tn;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {                       // Hillis-Steele inclusive scan
    unsigned long long a = 0, b = 0;
    if (t >= o) { a = sp[t - o]; b = sn[t - o]; }
    __syncthreads();
    sp[t] += a; sn[t] += b;
    __syncthreads();
  }
  unsigned long long ep = sp[t] - tp, en = sn[t] - tn;       // exclusive prefix of this thread's first chunk
  for (int k = 0; k < PER; ++k) {
    csum[t * PER + k] = ep; csum[OOD_NCHUNK + t * PER + k] = en;
    ep += lp[k]; en += ln[k];
  }
  if (t == 1023) { csum[2 * OOD_NCHUNK] = sp[t]; csum[2 * OOD_NCHUNK + 1] = sn[t]; }
}

// per chunk: sweep thresholds in descending score order
__global__ void __launch_bounds__(OOD_TPB) ood_sweep_kernel(const unsigned long long* __restrict__ hist,
                                                           const unsigned long long* __restrict__ csum,
                                                           double* __restrict__ part /* [3][NCHUNK]: auc, ap, fps@95 (or -1) */) {
  __shared__ unsigned long long sp[OOD_TPB], sn[OOD_TPB];
  __shared__ double sa[OOD_TPB], sb[OOD_TPB], sf[OOD_TPB];
  const int ch = blockIdx.x, t = threadIdx.x;
  const unsigned long long P = csum[2 * OOD_NCHUNK], N = csum[2 * OOD_NCHUNK + 1];
  // this thread owns OOD_PER_THREAD CONSECUTIVE positions
  unsigned long long hp[OOD_PER_THREAD], hn[OOD_PER_THREAD], tp = 0, tn = 0;
  const int64_t d0 = (int64_t)ch * OOD_CHUNK + (int64_t)t * OOD_PER_THREAD;
  for (int k = 0; k < OOD_PER_THREAD; ++k) {
    const int64_t b = OOD_NB - 1 - (d0 + k);
    hn[k] = hist[b]; hp[k] = hist[OOD_NB + b];
    tp += hp[k]; tn += hn[k];
  }
  sp[t] = tp; sn[t] = tn;
  __syncthreads();
  for (int o = 1; o < OOD_TPB; o <<= 1) {
    unsigned long long a = 0, b = 0;
    if (t >= o) { a = sp[t - o]; b = sn[t - o]; }
    __syncthreads();
    sp[t] += a; sn[t] += b;
    __syncthreads();
  }
  unsigned long long tps = csum[ch] + sp[t] - tp, fps = csum[OOD_NCHUNK + ch] + sn[t] - tn;   // counts above this thread's range
  double auc = 0.0, ap = 0.0, f95 = -1.0;
  const double dP = (double)P;
  for (int k = 0; k < OOD_PER_THREAD; ++k) {
    if (hp[k] | hn[k]) {                                     // a distinct threshold (sklearn: one point per distinct score)
      const unsigned long long tps1 = tps + hp[k], fps1 = fps + hn[k];
      auc += (double)(fps1 - fps) * ((double)tps1 + (double)tps) * 0.5;             // trapezoid in count units
      ap += (double)(tps1 - tps) * ((double)tps1 / ((double)tps1 + (double)fps1));  // (R_n - R_{n-1}) P_n, R in count units
      // support.py:247-257: first ROC point (descending threshold) with tpr > 0.95
      if ((double)tps1 / dP > 0.95 && !((double)tps / dP > 0.95)) f95 = (double)fps1;
      tps = tps1; fps = fps1;
    }
  }
  sa[t] = auc; sb[t] = ap; sf[t] = f95;
  __syncthreads();
  for (int o = OOD_TPB / 2; o > 0; o >>= 1) {                // fixed-order tree: deterministic
    if (t < o) { sa[t] += sa[t + o]; sb[t] += sb[t + o]; sf[t] = fmax(sf[t], sf[t + o]); }
    __syncthreads();
  }
  if (t == 0) { part[ch] = sa[0]; part[OOD_NCHUNK + ch] = sb[0]; part[2 * OOD_NCHUNK + ch] = sf[0]; }
}

__global__ void __launch_bounds__(1024) ood_reduce_kernel(const double* __restrict__ part, const unsigned long long* __restrict__ csum,
                                                         double* __restrict__ out /* auroc, aupr, fpr95, n_pos, n_neg */) {
  __shared__ double sa[1024], sb[1024], sf[1024];
  const int t = threadIdx.x;
  double a = 0, b = 0, f = -1.0;
  for (int k = t; k < OOD_NCHUNK; k += 1024) { a += part[k]; b += part[OOD_NCHUNK + k]; f = fmax(f, part[2 * OOD_NCHUNK + k]); }
  sa[t] = a; sb[t] = b; sf[t] = f;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (t < o) { sa[t] += sa[t + o]; sb[t] += sb[t + o]; sf[t] = fmax(sf[t], sf[t + o]); }
    __syncthreads();
  }
  if (t == 0) {
    const double P = (double)csum[2 * OOD_NCHUNK], N = (double)csum[2 * OOD_NCHUNK + 1];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    out[0] = (P > 0 && N > 0) ? sa[0] / (P * N) : nan;
    out[1] = (P > 0) ? sb[0] / P : nan;
    out[2] = (P > 0 && N > 0) ? (sf[0] >= 0 ? sf[0] / N : 0.0) : nan;   // never crossing 0.95 leaves fpr_best = 0 (support.py:250)
    out[3] = P; out[4] = N;
  }
}

}  // namespace rba

extern "C" int64_t rba_ood_hist_bytes(void) { return (int64_t)(2 * rba::OOD_NB * sizeof(unsigned long long)); }
extern "C" int64_t rba_ood_workspace_bytes(void) {
  return (int64_t)((2 * rba::OOD_NCHUNK + 2) * sizeof(unsigned long long) + 3 * rba::OOD_NCHUNK * sizeof(double));
}

// Adds n (score, label) pairs to the histogram.  label_dtype_bytes: 1 (uint8) or 8 (int64, what the reference's loaders yield).
extern "C" int rba_ood_hist_update(const float* score, const void* label, int label_dtype_bytes, int64_t n, void* hist, void* stream) {
  using namespace rba;
  if (n == 0) return RBA_OK;
  RBA_CHECK(score && label && hist && n > 0, "rba_ood_hist_update: bad arguments");
  RBA_CHECK(label_dtype_bytes == 1 || label_dtype_bytes == 8, "rba_ood_hist_update: labels must be uint8 or int64");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)std::min<int64_t>(cdiv(n, 256), 148 * 16);
  if (label_dtype_bytes == 1)
    ood_hist_kernel<uint8_t><<<blocks, 256, 0, st>>>(score, (const uint8_t*)label, n, (unsigned long long*)hist);
  else
    ood_hist_kernel<long long><<<blocks, 256, 0, st>>>(score, (const long long*)label, n, (unsigned long long*)hist);
  RBA_LAUNCHED();
  return RBA_OK;
}

// out (device, 5 doubles): auroc, aupr, fpr95, n_pos, n_neg.  workspace: rba_ood_workspace_bytes() device bytes.
extern "C" int rba_ood_hist_finalize(const void* hist, void* workspace, double* out, void* stream) {
  using namespace rba;
  RBA_CHECK(hist && workspace && out, "rba_ood_hist_finalize: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* csum = (unsigned long long*)workspace;
  double* part = (double*)(csum + 2 * OOD_NCHUNK + 2);
  const unsigned long long* h = (const unsigned long long*)hist;
  ood_chunk_sums_kernel<<<OOD_NCHUNK, OOD_TPB, 0, st>>>(h, csum);
  RBA_LAUNCHED();
  ood_chunk_scan_kernel<<<1, 1024, 0, st>>>(csum);
  RBA_LAUNCHED();
  ood_sweep_kernel<<<OOD_NCHUNK, OOD_TPB, 0, st>>>(h, csum, part);
  RBA_LAUNCHED();
  ood_reduce_kernel<<<1, 1024, 0, st>>>(part, csum, out);
  RBA_LAUNCHED();
  return RBA_OK;
}
