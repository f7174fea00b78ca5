// tcgen05 bf16x3 GEMM / implicit-GEMM 3x3 convolution (sm_100a).
//
//   C[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ residual)      A, W: bf16 split planes (hi, lo), K contiguous
//
// fp32-class accuracy on 5th-gen tensor cores: every K=16 step issues three tcgen05.mma (kind::f16, bf16 inputs,
// fp32 accumulate in TMEM) into the SAME accumulator: A_hi*W_hi + A_hi*W_lo + A_lo*W_hi.  The dropped terms
// (lo*lo and the 2^-18 split residuals) are ~1e-5 relative per product, random-signed; 1e-3 end-to-end parity needs
// better than TF32 and this gives ~17 mantissa bits at 2/3 of the single-pass bf16 rate with the
// operand bytes of ONE fp32 matrix (hi+lo = 4 B/element).
//
// Structure: persistent kernel, grid = min(#tiles, #SMs), 576 threads, tiles visited n-fastest so that concurrent CTAs
// share A rows and all of W through L2:
//   warp 0      TMA producer: cp.async.bulk.tensor (128B swizzle) of the 4 planes' 64-wide K blocks into a 3-stage
//               (2-stage in the store-staged configuration) mbarrier ring that runs continuously across tiles.
//   warp 1      TMEM allocator + MMA issuer: one elected thread, 12 tcgen05.mma per K block into one of TWO TMEM
//               accumulator buffers (2 x BN fp32 columns), tcgen05.commit -> `empty[s]`, last commit -> `acc_full[buf]`.
//   warps 2-17  epilogue (4 warps per TMEM lane quadrant, 32 columns each): prefetch bias / residual, wait `acc_full`,
//               tcgen05.ld 32x32b.x32, release the buffer (`acc_empty`), then bias / GELU / ReLU / residual / Swin
//               window-reverse row map -> fp32 and/or split-plane stores.  The epilogue of tile i overlaps the TMA
//               and MMA work of tile i+1.  For short K loops (K <= 512) the stores are the bottleneck: the store-staged
//               configuration routes the row-per-lane registers through a swizzled 4 KB per-warp shared-memory tile so
//               that every global store is 16 B per lane on 4 (fp32) or 8 (planes) full rows.
// The conv variant loads A through a 4-D tensor map over the NHWC planes: a tile is an 8x16 pixel patch, each of the
// 9 taps is the same box shifted by (dy-1, dx-1) with TMA out-of-bounds zero fill as the padding.
// Measured on B200 (profiles/): 350-440 TFLOP/s of useful fp32-equivalent FLOPs on the K >= 512 shapes (ceiling
// 1690/3 = 563), 4.3-5.3 TB/s on the HBM-bound K = 128 shapes.
#include "tcgen05.cuh"

namespace rba {

struct TcParams {
  int M, N, K;
  int nkb;                         // K blocks of 64 (conv: 9 * Cin/64)
  const float* bias; int bias_per_row; int64_t bias_bs;
  int act;
  const float* residual;
  float* c; int64_t ldc; int64_t c_bs;
  uint16_t* c_hi; uint16_t* c_lo; int64_t ldcp; int64_t cp_bs;
  int swin_map; SwinGeom geom;
  int a_batched, w_batched;
  // conv
  int cH, cW, cCin, tilesW, tilesH;
  // persistent tile schedule
  int tilesM, tilesN, ntiles;
  int qkv_heads, qkv_C;            // > 0: planes written in the (window, part, head) tiled layout (see rba_gemm_args.qkv_tile_heads)
  int debug;   // RBA_TC_DEBUG: 1 = skip global stores, 2 = skip the whole epilogue after the TMEM load (profiling aid)
};


// ---------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------
// Persistent: grid = min(#tiles, #SMs); every role loops over the tiles t = blockIdx.x, +gridDim.x, ... with n fastest
// (CTAs running at the same time share A rows and all of W through L2).  The smem ring runs continuously across
// tiles; the accumulator is double buffered in TMEM (2 x BN fp32 columns) so the epilogue of tile i overlaps the
// loads and MMAs of tile i+1.
constexpr int TC_EPI_WARPS = 16;                        // 4 warps per TMEM lane quadrant, each takes 32 columns
constexpr int TC_THREADS2 = (2 + TC_EPI_WARPS) * 32;    // 576

// STG = "store-staged" configuration for store-bound shapes (K <= 256): a 2-stage operand ring leaves room for a
// 4 KB per-warp staging buffer through which the epilogue turns its row-per-lane registers into coalesced 16-byte
// global stores (4 full lines per instruction instead of 32 partial ones).
// CG2 = CTA pair (`tcgen05.mma.cta_group::2`): the two CTAs of a cluster compute ONE 256 x 256 tile; each loads its own 128
// rows of A and HALF of the 256 W rows (the tensor core of either SM reads the other half from the peer's shared memory), so
// a stage is 64 KB per SM for twice the MMA work of the 128 x 128 tile: half the L2 -> shared-memory bytes per FLOP.
template <int BN, bool STG = false, bool CG2 = false>
struct TcSmem {
  static constexpr int STAGES = CG2 ? (STG ? 2 : 3) : ((BN > 128 || STG) ? 2 : 3);   // 96 KB stages for BN = 256, 64 KB (48 KB) otherwise
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;      // one plane of A per stage (16 KB)
  static constexpr int W_BYTES = (CG2 ? 128 : BN) * TC_BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
  static constexpr int NG = BN > 128 ? BN / 128 : 1;     // 128-column groups per tile (epilogue passes)
  static constexpr int BIAS_BYTES = TC_EPI_WARPS * 32 * NG * 4;
  static constexpr int OROW_BYTES = TC_EPI_WARPS * 32 * 8;
  static constexpr int STG_BYTES = STG ? TC_EPI_WARPS * 4096 : 0;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BIAS_BYTES + OROW_BYTES + STG_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

struct TileCoord {
  int m0, n0, bz;           // GEMM
  int cvb, cvh0, cvw0;      // conv
};

template <int BN, bool CONV, bool CG2 = false>
__device__ __forceinline__ TileCoord tile_coord(const TcParams& p, int t, int rank = 0) {
  TileCoord c;
  const int tn = t % p.tilesN;
  int r = t / p.tilesN;
  c.n0 = tn * BN;
  c.m0 = 0; c.bz = 0; c.cvb = 0; c.cvh0 = 0; c.cvw0 = 0;
  if (CONV) {
    c.cvw0 = (r % p.tilesW) * TC_CONV_TW;
    r /= p.tilesW;
    c.cvh0 = (r % p.tilesH) * TC_CONV_TH;
    c.cvb = r / p.tilesH;
  } else {
    c.m0 = CG2 ? (r % p.tilesM) * (2 * TC_BM) + rank * TC_BM : (r % p.tilesM) * TC_BM;   // CG2: tilesM counts 256-row tiles
    c.bz = r / p.tilesM;
  }
  return c;
}

// erf via Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7) on the FMA + XU pipes: the exact-erf GELU of swin.py:25 is
// evaluated 4C times per token, and libdevice erff (~30 instructions with branches) made the fc1 epilogue the
// bottleneck of the whole GEMM.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  const float erf_abs = fmaf(-poly, e, 1.0f);            // erf(|x|/sqrt2)
  const float erf_s = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erf_s);
}
template <int ACT>
__device__ __forceinline__ float tc_act(float x) {
  if (ACT == RBA_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == RBA_ACT_GELU) return gelu_fast(x);
  return x;
}

// ACT: RBA_ACT_*; OUTP: false = fp32 output (c), true = split-plane output (c_hi/c_lo) [both: handled by the fp32 variant
// falling back to scalar plane stores].
// ---- CTA-pair helpers ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are signalled on an mbarrier that may live in the peer CTA (shared::cluster address)
__device__ __forceinline__ void tma_load_3d_cg2(void* smem, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

template <int BN, bool CONV, int ACT, bool OUTP, bool STG, bool CG2 = false>
__global__ void __launch_bounds__(TC_THREADS2, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, const TcParams p) {
  static_assert(!CG2 || (BN == 256 && !CONV), "the CTA-pair form is the 256 x 256 GEMM tile");
  using S = TcSmem<BN, STG, CG2>;
  constexpr int TC_STAGES = S::STAGES;
  const int rank = CG2 ? (int)cluster_ctarank() : 0;     // CTA of the pair: 0 = leader (issues the MMAs)
  const int cta0 = CG2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, nctas = CG2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  // align to 1024 B (128B-swizzle atoms) with pointer arithmetic on the __shared__ array so that the compiler keeps the
  // shared address space (LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* sbias = reinterpret_cast<float*>(smem + TC_STAGES * S::STAGE_BYTES);
  int64_t* sorow = reinterpret_cast<int64_t*>(smem + TC_STAGES * S::STAGE_BYTES + S::BIAS_BYTES);
  uint8_t* sstage = smem + TC_STAGES * S::STAGE_BYTES + S::BIAS_BYTES + S::OROW_BYTES;   // [16 warps][4 KB] (STG only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sstage + S::STG_BYTES);
  uint64_t* full = bars;                          // [TC_STAGES]
  uint64_t* empty = bars + TC_STAGES;             // [TC_STAGES]
  uint64_t* acc_full = bars + 2 * TC_STAGES;      // [2]
  uint64_t* acc_empty = bars + 2 * TC_STAGES + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmW_hi); prefetch_tmap(&tmW_lo);
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], (BN >= 128 ? 4 : BN / 32) * 4 * (CG2 ? 2 : 1)); }   // one arrival per active epilogue warp (of both CTAs of a pair)
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CG2) cluster_sync_all(); else __syncthreads();      // (pair: the peer's barriers must be initialised before anything signals them)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp in uniform control flow, one elected lane issues) =====================
    uint32_t s = 0, ph = 1;                                // fresh "empty" barriers pass a wait on parity 1
    for (int t = cta0; t < p.ntiles; t += nctas) {
      const TileCoord tc = tile_coord<BN, CONV, CG2>(p, t, rank);
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(&empty[s], ph);
        if (elect_one()) {
          uint8_t* st = smem + s * S::STAGE_BYTES;
          if (CG2) {
            // both CTAs' bytes complete on the LEADER's barrier (armed by the leader for 2 x 64 KB); W: this CTA's half of the rows
            const uint32_t fb = mapa_u32(smem_u32(&full[s]), 0);
            if (rank == 0) mbar_expect_tx(&full[s], 2 * S::STAGE_BYTES);
            tma_load_3d_cg2(st, &tmA_hi, fb, kb * TC_BK, tc.m0, p.a_batched ? tc.bz : 0);
            tma_load_3d_cg2(st + S::A_BYTES, &tmA_lo, fb, kb * TC_BK, tc.m0, p.a_batched ? tc.bz : 0);
            tma_load_3d_cg2(st + 2 * S::A_BYTES, &tmW_hi, fb, kb * TC_BK, tc.n0 + rank * 128, p.w_batched ? tc.bz : 0);
            tma_load_3d_cg2(st + 2 * S::A_BYTES + S::W_BYTES, &tmW_lo, fb, kb * TC_BK, tc.n0 + rank * 128, p.w_batched ? tc.bz : 0);
          } else {
          mbar_expect_tx(&full[s], S::STAGE_BYTES);
          if (CONV) {
            const int cpb = p.cCin / TC_BK;               // channel blocks per tap
            const int tap = kb / cpb, cblk = kb - tap * cpb;
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            tma_load_4d(st, &tmA_hi, &full[s], cblk * TC_BK, tc.cvw0 + dx, tc.cvh0 + dy, tc.cvb);
            tma_load_4d(st + S::A_BYTES, &tmA_lo, &full[s], cblk * TC_BK, tc.cvw0 + dx, tc.cvh0 + dy, tc.cvb);
          } else {
            tma_load_3d(st, &tmA_hi, &full[s], kb * TC_BK, tc.m0, p.a_batched ? tc.bz : 0);
            tma_load_3d(st + S::A_BYTES, &tmA_lo, &full[s], kb * TC_BK, tc.m0, p.a_batched ? tc.bz : 0);
          }
          tma_load_3d(st + 2 * S::A_BYTES, &tmW_hi, &full[s], kb * TC_BK, tc.n0, p.w_batched ? tc.bz : 0);
          tma_load_3d(st + 2 * S::A_BYTES + S::W_BYTES, &tmW_lo, &full[s], kb * TC_BK, tc.n0, p.w_batched ? tc.bz : 0);
          }
        }
        __syncwarp();
        if (++s == TC_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The WHOLE warp walks the loop (uniform control flow, uniform counters) and one elected lane issues.  Under
    // `if (lane == 0)` ptxas wraps every tcgen05.mma in an ELECT / BRA.U.ANY loop (seen in the SASS of every instantiation of
    // the first revision; measured ~70 clk per MMA in the window-attention kernel): with 64 clk of math per 128x128x16 MMA that
    // made the BN = 128 GEMMs issue-bound at ~56 % tensor-pipe activity (profiles/r2f_ncu_kernels.md).  Descriptors are a
    // constant high word plus a low word (address / 16 | LBO) that advances by compile-time offsets.
    constexpr uint32_t idesc = make_idesc(CG2 ? 2 * TC_BM : TC_BM, BN);
    constexpr uint32_t D_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
    const uint32_t lo0 = (smem_u32(smem) >> 4) | (1u << 16);
    uint32_t lt = 0;
    uint32_t s = 0, ph = 0;
    for (int t = cta0; (!CG2 || rank == 0) && t < p.ntiles; t += nctas, ++lt) {     // pair: only the leader issues
      const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
      mbar_wait(&acc_empty[as], aph ^ 1);                // epilogue (of both CTAs) has drained this accumulator buffer
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = lo0 + s * (S::STAGE_BYTES >> 4), a_lo = a_hi + (S::A_BYTES >> 4);
          const uint32_t w_hi = a_hi + (2 * S::A_BYTES >> 4), w_lo = w_hi + (S::W_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint32_t adv = (uint32_t)(k * 32 >> 4);  // +32 B along K inside the 128 B swizzle span
            if (CG2) {
              umma_bf16_cg2(tmem_d, tc_desc(a_hi + adv, D_HI), tc_desc(w_hi + adv, D_HI), idesc, (kb | k) != 0);
              umma_bf16_cg2(tmem_d, tc_desc(a_hi + adv, D_HI), tc_desc(w_lo + adv, D_HI), idesc, 1);
              umma_bf16_cg2(tmem_d, tc_desc(a_lo + adv, D_HI), tc_desc(w_hi + adv, D_HI), idesc, 1);
            } else {
              umma_bf16(tmem_d, tc_desc(a_hi + adv, D_HI), tc_desc(w_hi + adv, D_HI), idesc, (kb | k) != 0);
              umma_bf16(tmem_d, tc_desc(a_hi + adv, D_HI), tc_desc(w_lo + adv, D_HI), idesc, 1);
              umma_bf16(tmem_d, tc_desc(a_lo + adv, D_HI), tc_desc(w_hi + adv, D_HI), idesc, 1);
            }
          }
          if (CG2) {
            umma_commit_cg2(&empty[s]);                     // frees the stage in BOTH CTAs when these MMAs retire
            if (kb == p.nkb - 1) umma_commit_cg2(&acc_full[as]);
          } else {
            umma_commit(&empty[s]);                         // frees the smem stage when these MMAs retire
            if (kb == p.nkb - 1) umma_commit(&acc_full[as]);   // accumulator complete
          }
        }
        __syncwarp();
        if (++s == TC_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: warps 2..17; TMEM lane quadrant = warp % 4, 32-column group = (warp - 2) / 4 =====
    // Each warp owns a 32-row x 32-column block of every tile.  Everything that does not depend on the accumulator
    // (output row, bias slice, residual row segment) is fetched BEFORE waiting on acc_full so that global latency
    // overlaps the MMAs of the tile; the TMEM buffer is released right after the single tcgen05.ld.
    const int ew = warp - 2;
    const int quad = warp & 3, cg = ew >> 2;
    if (cg * 32 < BN) {
      constexpr int NG = S::NG;
      float* mybias = sbias + ew * 32 * NG;
      int64_t* my_orow = sorow + ew * 32;
      uint8_t* mystage = sstage + ew * 4096;
      const int row_in_tile = quad * 32 + lane;
      uint32_t lt = 0;
      for (int t = cta0; t < p.ntiles; t += nctas, ++lt) {
        const TileCoord tc = tile_coord<BN, CONV, CG2>(p, t, rank);
        const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
        const int bz = tc.bz;
        int64_t orow = -1;                                  // output row (or -1: nothing to store)
        const int m_logical = tc.m0 + row_in_tile;
        if (CONV) {
          const int hh = tc.cvh0 + row_in_tile / TC_CONV_TW, ww = tc.cvw0 + row_in_tile % TC_CONV_TW;
          if (hh < p.cH && ww < p.cW) orow = ((int64_t)tc.cvb * p.cH + hh) * p.cW + ww;
        } else if (m_logical < p.M) {
          orow = m_logical;
          if (p.swin_map) orow = swin_row_to_token(p.geom, m_logical);
        }
        const float* bias = p.bias ? p.bias + bz * p.bias_bs : nullptr;
        float* c = p.c ? p.c + bz * p.c_bs : nullptr;
        const int64_t ldr = p.c ? p.ldc : p.ldcp;
        const float* res = p.residual ? p.residual + bz * (p.c ? p.c_bs : p.cp_bs) : nullptr;
        uint16_t* c_hi = p.c_hi ? p.c_hi + bz * p.cp_bs : nullptr;
        uint16_t* c_lo = p.c_lo ? p.c_lo + bz * p.cp_bs : nullptr;
        float brow = 0.f;
        if (bias) {
          if (p.bias_per_row) {
            if (orow >= 0 && !CONV) brow = bias[m_logical];
          } else {
            __syncwarp();
#pragma unroll
            for (int g2 = 0; g2 < NG; ++g2) {
              const int n = tc.n0 + g2 * 128 + cg * 32 + lane;
              mybias[g2 * 32 + lane] = (n < p.N) ? bias[n] : 0.f;
            }
            __syncwarp();
          }
        }
        if (STG) {
          __syncwarp();
          my_orow[lane] = orow;
          __syncwarp();
        }
#pragma unroll 1
        for (int g2 = 0; g2 < NG; ++g2) {
        const int nb = tc.n0 + g2 * 128 + cg * 32;          // first column of this warp's 32x32 block
        const bool active = orow >= 0 && nb < p.N;
        const bool full32 = nb + 31 < p.N;
        // ---- prefetch the residual row segment (independent of the accumulator) ----
        float4 r4[8];
        const bool res_vec = res && active && full32 && ((ldr & 3) == 0);
        if (res_vec) {
          const float4* rp = reinterpret_cast<const float4*>(res + orow * ldr + nb);
#pragma unroll
          for (int j = 0; j < 8; ++j) r4[j] = rp[j];
        }
        // ---- accumulator ----
        __syncwarp();
        if (g2 == 0) {
          mbar_wait(&acc_full[as], aph);
          tc_fence_after();
        }
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + g2 * 128 + cg * 32), v);
        tmem_ld_wait();
        if (g2 == NG - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {                                  // this warp is done with the TMEM buffer (pair: tell the leader)
            if (CG2) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[as]), 0)); else mbar_arrive(&acc_empty[as]);
          }
        }
        if ((p.debug & 2) || nb >= p.N) { __syncwarp(); continue; }   // warp-uniform
        // ---- bias, activation, residual in the TMEM layout (lane <-> row, register j <-> column nb + j) ----
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        if (bias) {
          if (p.bias_per_row) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] += brow;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = *reinterpret_cast<const float4*>(mybias + g2 * 32 + 4 * j);
              x[4 * j] += b4.x; x[4 * j + 1] += b4.y; x[4 * j + 2] += b4.z; x[4 * j + 3] += b4.w;
            }
          }
        }
        if (ACT != RBA_ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = tc_act<ACT>(x[j]);
        }
        if (res_vec) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { x[4 * j] += r4[j].x; x[4 * j + 1] += r4[j].y; x[4 * j + 2] += r4[j].z; x[4 * j + 3] += r4[j].w; }
        } else if (res && active) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < p.N) x[j] += res[orow * ldr + nb + j];
        }
        // ---- stores ----
        const bool staged = STG && full32 && ((OUTP ? (p.ldcp & 7) : (p.ldc & 3)) == 0) && !(!OUTP && c_hi) && !(OUTP && p.qkv_heads);   // warp-uniform
        if (p.debug & 1) {
        } else if (staged) {
          // Row-per-lane registers -> swizzled shared-memory tile -> coalesced 16-byte global stores.
          if (!OUTP) {
            // tile: 32 rows x 128 B; 16-byte chunk j of row r lives at chunk (j ^ (r & 7))
            float4* st4 = reinterpret_cast<float4*>(mystage);
#pragma unroll
            for (int j = 0; j < 8; ++j) st4[lane * 8 + (j ^ (lane & 7))] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            __syncwarp();
            const int cidx = lane & 7;
#pragma unroll
            for (int i = 0; i < 8; ++i) {                   // 4 rows x 128 B per instruction
              const int rr = i * 4 + (lane >> 3);
              const int64_t o = my_orow[rr];
              const float4 val = st4[rr * 8 + (cidx ^ (rr & 7))];
              if (o >= 0) *reinterpret_cast<float4*>(c + o * p.ldc + nb + cidx * 4) = val;
            }
          } else {
            // two tiles (hi, lo): 32 rows x 64 B; chunk j of row r lives at chunk (j ^ ((r >> 1) & 3))
            uint4* sh4 = reinterpret_cast<uint4*>(mystage);
            uint4* sl4 = sh4 + 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 Hh, Ll;
              split_pack2(x[8 * j], x[8 * j + 1], Hh.x, Ll.x);
              split_pack2(x[8 * j + 2], x[8 * j + 3], Hh.y, Ll.y);
              split_pack2(x[8 * j + 4], x[8 * j + 5], Hh.z, Ll.z);
              split_pack2(x[8 * j + 6], x[8 * j + 7], Hh.w, Ll.w);
              const int pc = j ^ ((lane >> 1) & 3);
              sh4[lane * 4 + pc] = Hh;
              sl4[lane * 4 + pc] = Ll;
            }
            __syncwarp();
            const int cidx = lane & 3;
#pragma unroll
            for (int i = 0; i < 4; ++i) {                   // 8 rows x 64 B per instruction and plane
              const int rr = i * 8 + (lane >> 2);
              const int64_t o = my_orow[rr];
              const int pc = cidx ^ ((rr >> 1) & 3);
              const uint4 hv = sh4[rr * 4 + pc], lv = sl4[rr * 4 + pc];
              if (o >= 0) {
                *reinterpret_cast<uint4*>(c_hi + o * p.ldcp + nb + cidx * 8) = hv;
                *reinterpret_cast<uint4*>(c_lo + o * p.ldcp + nb + cidx * 8) = lv;
              }
            }
          }
        } else if (active) {
          if (!OUTP) {
            float* cp = c + orow * p.ldc + nb;
            if (full32 && (p.ldc & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(cp)[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nb + j < p.N) cp[j] = x[j];
            }
            if (c_hi) {                                     // rare: both outputs requested
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nb + j < p.N) store_split1(c_hi, c_lo, orow * p.ldcp + nb + j, x[j]);
            }
          } else {
            int64_t ob = orow * p.ldcp + nb;
            int64_t cs = 8;                                 // elements between the four 8-column chunks of this row segment
            if (p.qkv_heads) {
              // tiled q | k | v planes: this warp's 32 columns are exactly one head of one part; consecutive lanes (rows) are
              // 16 bytes apart inside a chunk plane, so every store instruction writes 512 contiguous bytes
              const int part = nb / p.qkv_C, head = (nb - part * p.qkv_C) >> 5;
              const int64_t tile = orow / 144;
              const int r = (int)(orow - tile * 144);
              ob = (((tile * 3 + part) * p.qkv_heads + head) * 4) * (int64_t)(144 * 8) + r * 8;
              cs = 144 * 8;
            }
            if (full32 && (p.ldcp & 7) == 0) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {                 // 8 columns -> one 16-byte store per plane
                uint4 Hh, Ll;
                split_pack2(x[8 * j], x[8 * j + 1], Hh.x, Ll.x);
                split_pack2(x[8 * j + 2], x[8 * j + 3], Hh.y, Ll.y);
                split_pack2(x[8 * j + 4], x[8 * j + 5], Hh.z, Ll.z);
                split_pack2(x[8 * j + 6], x[8 * j + 7], Hh.w, Ll.w);
                *reinterpret_cast<uint4*>(c_hi + ob + cs * j) = Hh;
                *reinterpret_cast<uint4*>(c_lo + ob + cs * j) = Ll;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nb + j < p.N) store_split1(c_hi, c_lo, ob + j, x[j]);
            }
          }
        }
        __syncwarp();                                       // reconverge before the next .aligned TMEM op
        }  // g2
      }
    }
    tc_fence_before();
  }
  if (CG2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

static const int g_tc_stg_maxk = []() { const char* e = getenv("RBA_TC_STG_MAXK"); return e ? atoi(e) : 512; }();

// CTA-pair launch: 2-CTA clusters, grid = 2 x min(#256x256 tiles, #SMs / 2)
template <int ACT, bool OUTP, bool STG>
static int launch_cg2s(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi, const CUtensorMap& w_lo,
                       const TcParams& p, cudaStream_t st) {
  constexpr int smem = TcSmem<256, STG, true>::TOTAL;
  auto kern = gemm_tc_kernel<256, false, ACT, OUTP, STG, true>;
  static PerDeviceOnce once;
  if (once.needed()) {
    RBA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    once.done();
  }
  const int npairs = std::min<int>(p.ntiles, num_sms() / 2);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(2 * npairs));
  cfg.blockDim = dim3(TC_THREADS2);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RBA_CUDA(cudaLaunchKernelEx(&cfg, kern, a_hi, a_lo, w_hi, w_lo, p));
  RBA_LAUNCHED();
  return RBA_OK;
}
// short K loops are store-bound: 2-stage ring + staged coalesced stores (as for the single-CTA tiles); RBA_TC_CG2_STG=0 disables
template <int ACT, bool OUTP>
static int launch_cg2(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi, const CUtensorMap& w_lo,
                      const TcParams& p, cudaStream_t st) {
  static const int stg = []() { const char* e = getenv("RBA_TC_CG2_STG"); return e ? atoi(e) : 1; }();
  if (stg && p.K <= g_tc_stg_maxk) return launch_cg2s<ACT, OUTP, true>(a_hi, a_lo, w_hi, w_lo, p, st);
  return launch_cg2s<ACT, OUTP, false>(a_hi, a_lo, w_hi, w_lo, p, st);
}
static int launch_tc_cg2(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi, const CUtensorMap& w_lo,
                         const TcParams& p, cudaStream_t st) {
  const bool outp = p.c == nullptr;                       // planes only
  switch (p.act) {
    case RBA_ACT_RELU:
      return outp ? launch_cg2<RBA_ACT_RELU, true>(a_hi, a_lo, w_hi, w_lo, p, st) : launch_cg2<RBA_ACT_RELU, false>(a_hi, a_lo, w_hi, w_lo, p, st);
    case RBA_ACT_GELU:
      return outp ? launch_cg2<RBA_ACT_GELU, true>(a_hi, a_lo, w_hi, w_lo, p, st) : launch_cg2<RBA_ACT_GELU, false>(a_hi, a_lo, w_hi, w_lo, p, st);
    default:
      return outp ? launch_cg2<RBA_ACT_NONE, true>(a_hi, a_lo, w_hi, w_lo, p, st) : launch_cg2<RBA_ACT_NONE, false>(a_hi, a_lo, w_hi, w_lo, p, st);
  }
}


template <int BN, bool CONV, int ACT, bool OUTP, bool STG>
static int launch_tc3(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi, const CUtensorMap& w_lo,
                      const TcParams& p, dim3 grid, cudaStream_t st) {
  constexpr int smem = TcSmem<BN, STG>::TOTAL;
  static PerDeviceOnce once;
  if (once.needed()) {
    RBA_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, CONV, ACT, OUTP, STG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    once.done();
  }
  gemm_tc_kernel<BN, CONV, ACT, OUTP, STG><<<grid, TC_THREADS2, smem, st>>>(a_hi, a_lo, w_hi, w_lo, p);
  RBA_LAUNCHED();
  return RBA_OK;
}

// 256-wide N tiles halve the A-operand shared-memory reads per MAC.  Under sustained (board-power-capped) load they gain
// 4-7 % on the K >= 1024 shapes and lose on K = 512 / 128 (profiles/r1e_gemm_power_sustained.txt): RBA_TC_BN256 = 0 never,
// 1 always (when N % 256 == 0), 2 (default) for K >= 1024.
static const int g_tc_bn256 = []() { const char* e = getenv("RBA_TC_BN256"); return e ? atoi(e) : 2; }();

template <int BN, bool CONV, int ACT, bool OUTP>
static int launch_tc2(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi, const CUtensorMap& w_lo,
                      const TcParams& p, dim3 grid, cudaStream_t st) {
  // store-bound shapes (short K loop, BN = 128): 2-stage ring + staged coalesced stores
  if (BN == 128 && !CONV && p.K <= g_tc_stg_maxk) return launch_tc3<BN, CONV, ACT, OUTP, (BN == 128 && !CONV)>(a_hi, a_lo, w_hi, w_lo, p, grid, st);
  return launch_tc3<BN, CONV, ACT, OUTP, false>(a_hi, a_lo, w_hi, w_lo, p, grid, st);
}

template <int BN, bool CONV>
static int launch_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi, const CUtensorMap& w_lo,
                     const TcParams& p, dim3 grid, cudaStream_t st) {
  const bool outp = p.c == nullptr;                       // planes only
  if (CONV) return launch_tc2<BN, CONV, RBA_ACT_NONE, false>(a_hi, a_lo, w_hi, w_lo, p, grid, st);
  switch (p.act) {
    case RBA_ACT_RELU:
      return outp ? launch_tc2<BN, false, RBA_ACT_RELU, true>(a_hi, a_lo, w_hi, w_lo, p, grid, st)
                  : launch_tc2<BN, false, RBA_ACT_RELU, false>(a_hi, a_lo, w_hi, w_lo, p, grid, st);
    case RBA_ACT_GELU:
      return outp ? launch_tc2<BN, false, RBA_ACT_GELU, true>(a_hi, a_lo, w_hi, w_lo, p, grid, st)
                  : launch_tc2<BN, false, RBA_ACT_GELU, false>(a_hi, a_lo, w_hi, w_lo, p, grid, st);
    default:
      return outp ? launch_tc2<BN, false, RBA_ACT_NONE, true>(a_hi, a_lo, w_hi, w_lo, p, grid, st)
                  : launch_tc2<BN, false, RBA_ACT_NONE, false>(a_hi, a_lo, w_hi, w_lo, p, grid, st);
  }
}

static const int g_tc_debug = []() { const char* e = getenv("RBA_TC_DEBUG"); return e ? atoi(e) : 0; }();

static void fill_epilogue(TcParams& p, const rba_gemm_args& a) {
  p.debug = g_tc_debug;
  p.bias = a.bias; p.bias_per_row = a.bias_per_row; p.bias_bs = a.bias_bstride;
  p.act = a.act; p.residual = a.residual;
  p.c = a.c; p.ldc = a.ldc; p.c_bs = a.c_bstride;
  p.c_hi = a.c_hi; p.c_lo = a.c_lo; p.ldcp = a.ldcp; p.cp_bs = a.cp_bstride;
  p.swin_map = a.swin_map;
  p.qkv_heads = a.qkv_tile_heads; p.qkv_C = a.qkv_tile_heads * 32;
  p.geom = make_swin_geom(a.sw_H > 0 ? a.sw_H : 1, a.sw_W > 0 ? a.sw_W : 1, a.sw_ws > 0 ? a.sw_ws : 1, a.sw_shift);
}

int gemm_tc_launch(const rba_gemm_args& a, cudaStream_t st) {
  RBA_CHECK(a.K % 8 == 0, "gemm(tc): K must be a multiple of 8");
  if (a.qkv_tile_heads)
    RBA_CHECK(a.qkv_tile_heads > 0 && a.N == 3 * 32 * a.qkv_tile_heads && a.M % 144 == 0 && a.c == nullptr && a.c_hi && a.c_lo &&
                  !a.swin_map && a.batch == 1 && (a.ldcp & 7) == 0,
              "gemm(tc): the tiled q|k|v layout needs N = 96 heads, M a multiple of 144, split-plane output only");
  RBA_CHECK(((uintptr_t)a.a_hi & 15) == 0 && ((uintptr_t)a.a_lo & 15) == 0 && ((uintptr_t)a.w_hi & 15) == 0 &&
                ((uintptr_t)a.w_lo & 15) == 0, "gemm(tc): operand planes must be 16-byte aligned");
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.nkb = (a.K + TC_BK - 1) / TC_BK;
  fill_epilogue(p, a);
  p.a_batched = a.batch > 1 && a.a_bstride != 0;
  p.w_batched = a.batch > 1 && a.w_bstride != 0;
  int BN = a.N > 64 ? 128 : 64;
  if (a.N % 256 == 0 && (g_tc_bn256 == 1 || (g_tc_bn256 == 2 && a.K >= 1024))) BN = 256;   // K = 512, N >= 2048 measured slower
  // CTA pairs (256 x 256 tiles over two SMs): RBA_TC_CG2 = 0 never, 1 (default) when N % 256 == 0, K % 64 == 0, K >= 512 and
  // M >= 512 (measured, profiles/r2o_gemm_cta_pairs.txt: +4 ... +13 % on the K >= 512 shapes; the store-bound K = 128 shapes
  // lose the staged stores and a 256 x 256 tile there is two K blocks long: 2 = also those, for the record)
  static const int g_tc_cg2 = []() { const char* e = getenv("RBA_TC_CG2"); return e ? atoi(e) : 1; }();
  const bool cg2 = g_tc_cg2 != 0 && a.N % 256 == 0 && a.K % TC_BK == 0 && a.M >= 512 && (a.K >= 512 || g_tc_cg2 == 2);
  CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
  if (cg2) {
    RBA_TRY_(make_map_3d(&ta_hi, a.a_hi, a.K, a.M, a.lda, p.a_batched ? a.batch : 1, a.a_bstride, TC_BM));
    RBA_TRY_(make_map_3d(&ta_lo, a.a_lo, a.K, a.M, a.lda, p.a_batched ? a.batch : 1, a.a_bstride, TC_BM));
    RBA_TRY_(make_map_3d(&tw_hi, a.w_hi, a.K, a.N, a.ldw, p.w_batched ? a.batch : 1, a.w_bstride, 128));
    RBA_TRY_(make_map_3d(&tw_lo, a.w_lo, a.K, a.N, a.ldw, p.w_batched ? a.batch : 1, a.w_bstride, 128));
    p.tilesM = (int)cdiv(a.M, 2 * TC_BM); p.tilesN = a.N / 256;
    const int64_t nt2 = (int64_t)p.tilesM * p.tilesN * a.batch;
    RBA_CHECK(nt2 < (1LL << 31), "gemm(tc): too many tiles");
    p.ntiles = (int)nt2;
    return launch_tc_cg2(ta_hi, ta_lo, tw_hi, tw_lo, p, st);
  }
  RBA_TRY_(make_map_3d(&ta_hi, a.a_hi, a.K, a.M, a.lda, p.a_batched ? a.batch : 1, a.a_bstride, TC_BM));
  RBA_TRY_(make_map_3d(&ta_lo, a.a_lo, a.K, a.M, a.lda, p.a_batched ? a.batch : 1, a.a_bstride, TC_BM));
  RBA_TRY_(make_map_3d(&tw_hi, a.w_hi, a.K, a.N, a.ldw, p.w_batched ? a.batch : 1, a.w_bstride, BN));
  RBA_TRY_(make_map_3d(&tw_lo, a.w_lo, a.K, a.N, a.ldw, p.w_batched ? a.batch : 1, a.w_bstride, BN));
  p.tilesM = (int)cdiv(a.M, TC_BM); p.tilesN = (int)cdiv(a.N, BN);
  const int64_t nt = (int64_t)p.tilesM * p.tilesN * a.batch;
  RBA_CHECK(nt < (1LL << 31), "gemm(tc): too many tiles");
  p.ntiles = (int)nt;
  dim3 grid((unsigned)std::min<int64_t>(nt, num_sms()));
  if (BN == 256) return launch_tc<256, false>(ta_hi, ta_lo, tw_hi, tw_lo, p, grid, st);
  if (BN == 128) return launch_tc<128, false>(ta_hi, ta_lo, tw_hi, tw_lo, p, grid, st);
  return launch_tc<64, false>(ta_hi, ta_lo, tw_hi, tw_lo, p, grid, st);
}

int conv3x3_tc_launch(const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* w_hi, const uint16_t* w_lo, int B, int H,
                      int W, int Cin, int Cout, float* y, cudaStream_t st) {
  RBA_CHECK(Cin % TC_BK == 0, "conv3x3(tc): Cin must be a multiple of 64 (got %d)", Cin);
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = B * H * W; p.N = Cout; p.K = 9 * Cin;
  p.nkb = 9 * (Cin / TC_BK);
  p.c = y; p.ldc = Cout;
  p.geom = make_swin_geom(1, 1, 1, 0);
  p.cH = H; p.cW = W; p.cCin = Cin;
  p.tilesW = (int)cdiv(W, TC_CONV_TW); p.tilesH = (int)cdiv(H, TC_CONV_TH);
  CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
  RBA_TRY_(make_map_nhwc(&ta_hi, x_hi, B, H, W, Cin));
  RBA_TRY_(make_map_nhwc(&ta_lo, x_lo, B, H, W, Cin));
  const int BN = (Cout % 256 == 0 && g_tc_bn256 != 0) ? 256 : (Cout > 64 ? 128 : 64);   // K = 9 Cin >= 576: long-K shape
  RBA_TRY_(make_map_3d(&tw_hi, w_hi, 9 * Cin, Cout, 9 * Cin, 1, 0, BN));
  RBA_TRY_(make_map_3d(&tw_lo, w_lo, 9 * Cin, Cout, 9 * Cin, 1, 0, BN));
  p.tilesM = 1; p.tilesN = (int)cdiv(Cout, BN);
  const int64_t nt = (int64_t)B * p.tilesH * p.tilesW * p.tilesN;
  RBA_CHECK(nt < (1LL << 31), "conv3x3(tc): too many tiles");
  p.ntiles = (int)nt;
  dim3 grid((unsigned)std::min<int64_t>(nt, num_sms()));
  if (BN == 256) return launch_tc<256, true>(ta_hi, ta_lo, tw_hi, tw_lo, p, grid, st);
  if (BN == 128) return launch_tc<128, true>(ta_hi, ta_lo, tw_hi, tw_lo, p, grid, st);
  return launch_tc<64, true>(ta_hi, ta_lo, tw_hi, tw_lo, p, grid, st);
}

}  // namespace rba
