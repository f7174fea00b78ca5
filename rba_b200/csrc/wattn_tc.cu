// Swin window attention core on tcgen05 + TMA (swin.py:145-168).
//
//   S = (q k^T) * scale + relative_position_bias (+ shift mask);  P = softmax(S);  O = P v        per (window, head)
//
// q, k, v arrive as bf16 split planes [rows = B*nW*144, 3C] (what the QKV GEMM epilogue writes), O leaves as split planes
// [rows, C] (what the proj GEMM reads through TMA).  Both contractions run on the 5th-generation tensor cores in bf16x3
// split precision (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM), everything else in fp32.
//
// One persistent CTA per SM walks (window, head) items.  Per item:
//   TMA        six 144 x 32 bf16 boxes (q, k, v  x  hi, lo; 64-byte rows, SWIZZLE_64B) + the head's 529-entry bias table
//              (bulk copy) into one of three shared-memory stages                                     55.3 KB in flight / stage
//   S          tcgen05.mma M=128 N=144 K=32: query rows 0..127 -> TMEM buffer S1[item & 1]; query rows 128..143 (+112 rows
//              that are never read) -> S2.  A = q (K-major), B = k (K-major), 6 MMAs each.
//   softmax    thread per (row, quarter of the 144 keys): tcgen05.ld its 36 logits, bias (+ mask), max / sum exchanged with
//              the threads that own the other quarters, exp2, bf16 hi/lo split of the UN-normalised probabilities written
//              back over S with tcgen05.st (P aliases S: 144 fp32 columns = 72 + 72 packed bf16x2 columns).
//   O          tcgen05.mma M=128 N=32 K=144 with A = P read from TENSOR MEMORY and B = v straight from the TMA tile
//              (MN-major descriptor: no transpose anywhere), 27 MMAs per tile.
//   epilogue   O rows * 1 / row sum -> bf16 hi/lo -> global planes.
// The MMAs of item i+1 (S) run under the softmax of item i; TMA runs two items ahead.  Warps: 0 = TMA producer, 1 = MMA
// issuer of the 128-row tile + TMEM owner, 2 = MMA issuer of the 16-row tile (two issuers keep the two tiles' chains
// independent), 3 idle, 4..19 = softmax (quadrant w % 4, key quarter (w - 4) / 4; RBA_WT_PARTS=2 restores the 12-warp form
// with 72 keys per thread: 7.0 instead of 5.8 ms per 8 images, profiles/r2i_wattn_*): every warp serves the 128-row tile of every
// item, and the two warps of quadrant i % 4 also serve the 16-row tile of item i, whose rows the issuer places on that
// quadrant's lanes.
// TMEM columns: S1[0] 0..143 | S1[1] 144..287 | S2 288..431 | O1 432..463 | O2 464..495 (496 of 512: S2 cannot be double
// buffered; an M=64 variant that would have packed two items' 16-row tiles into one block at lane offsets 0 / 16 gave
// wrong results at offset 16 on this hardware / toolchain and was dropped).
// Measured (profiles/r2c_*): the softmax warps bound the kernel (10 warp-passes of 72 elements per item at ~8 instructions
// and one MUFU.EX2 per element; the SM sub-partition that owns the 16-row tile's lanes carries 4 of them), not the tensor
// pipe (66 MMAs, ~1.7k clk) nor HBM (74 KB per item).  RBA_WT_DEBUG bits 1/2/4/8/16 and RBA_WT_TIMELINE are the ablation /
// in-kernel clock aids those measurements came from.
#include <type_traits>

#include "tcgen05.cuh"

namespace rba {

constexpr int WT_N = 144, WT_D = 32, WT_WS = 12;
constexpr int WT_TILE_BYTES = WT_N * WT_D * 2;            // 9216 = 9 * 1024: one plane of q, k or v of one item
constexpr int WT_BIAS_FLOATS = 532;                       // == WM_BIAS_PITCH of attn.cu (prepared, log2(e)-scaled table)
constexpr int WT_BIAS_BYTES = WT_BIAS_FLOATS * 4;         // 2128 (multiple of 16: bulk-copy granularity)
constexpr int WT_STAGE_BYTES = 6 * WT_TILE_BYTES + 3072;  // 58368 = 57 * 1024
constexpr int WT_STAGES = 3;
constexpr int WT_MAXPARTS = 6;                             // softmax warps per TMEM lane quadrant (key partitions of a row): 2 or 4
constexpr int WT_THREADS_MAX = (4 + 4 * WT_MAXPARTS) * 32;
constexpr uint32_t WT_TMEM_COLS = 512;
constexpr int WT_COL_S1 = 0, WT_COL_S2 = 288, WT_COL_O1 = 432, WT_COL_O2 = 464;
constexpr int WT_EXCH_FLOATS = 2 /*parity*/ * WT_MAXPARTS /*key partition*/ * 160 /*rows (144, padded)*/;
constexpr int WT_SMEM = WT_STAGES * WT_STAGE_BYTES + 2 * WT_EXCH_FLOATS * 4 + 256 + 1024;

struct WtParams {
  const float* bias;        // [heads][532] prepared table
  uint16_t* out_hi;
  uint16_t* out_lo;
  int C, heads, nWh, nWw, shift;
  int64_t nitems;           // B * nWh * nWw * heads
  float scale_log2e;        // head_dim^-0.5 * log2(e)
  const uint16_t* t_hi;     // tiled q | k | v planes (rba_gemm_args.qkv_tile_heads layout) or null: row-major planes via tensor maps
  const uint16_t* t_lo;
  int debug;
  long long* tl;            // profiling aid (RBA_WT_TIMELINE): clock64 stamps of CTA 0, [item < 32][16 events]
};
#define WT_STAMP(item, ev)                                                                   \
  do {                                                                                       \
    if (p.tl && blockIdx.x == 0 && (item) < 32) p.tl[(item) * 16 + (ev)] = clock64();        \
  } while (0)

__device__ __forceinline__ void wt_part_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// 16 consecutive output values * inv -> bf16 hi / lo planes, two 16-byte stores each (g is a multiple of 16 elements)
__device__ __forceinline__ void wt_store16(uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo, int64_t g, const uint32_t* o, float inv) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_pack2(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv, h[i], l[i]);
  uint4* ph = reinterpret_cast<uint4*>(out_hi + g);
  uint4* pl = reinterpret_cast<uint4*>(out_lo + g);
  ph[0] = make_uint4(h[0], h[1], h[2], h[3]);
  ph[1] = make_uint4(h[4], h[5], h[6], h[7]);
  pl[0] = make_uint4(l[0], l[1], l[2], l[3]);
  pl[1] = make_uint4(l[4], l[5], l[6], l[7]);
}

__device__ __forceinline__ void wt_store8(uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo, int64_t g, const uint32_t* o, float inv) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pack2(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv, h[i], l[i]);
  *reinterpret_cast<uint4*>(out_hi + g) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(out_lo + g) = make_uint4(l[0], l[1], l[2], l[3]);
}

// barriers (shared memory, 8 bytes each)
struct WtBars {
  uint64_t full[WT_STAGES], empty[WT_STAGES];
  uint64_t s1_full[2], p1_ready[2];
  uint64_t o1_full, o1_empty;
  uint64_t s2_full, p2_ready, o2_full, o2_empty;
  uint32_t tmem_slot, pad;
};

// item -> (window index in [0, B*nWh*nWw), head); heads fastest so that consecutive items share the window's rows in L2
__device__ __forceinline__ void wt_item(const WtParams& p, int64_t it, int64_t& win, int& head) {
  win = it / p.heads;
  head = (int)(it - win * p.heads);
}

// Per-thread state of a softmax warp: quadrant = TMEM lane quadrant (= SM sub-partition) of the warp, half = which 72 of the
// 144 keys this thread owns.
struct WtThread {
  int quadrant, half, lane;      // half = key partition index (0 .. PARTS - 1)
  uint32_t lane_addr;
  int pair_bar;
};

// One softmax pass over one S tile of item `lt`: logits (72 keys of one query row) -> bias (+ mask) -> row max across the two
// halves -> exp2 -> bf16 hi / lo of the un-normalised probabilities written back over S -> "P ready".  TILE2 = the 16-row tile
// (query rows 128..143 on lanes 0..15 of this warp's quadrant; lanes 16..31 run along on rows that are never stored).
template <bool TILE2, int PARTS>
__device__ __forceinline__ void wt_pass(const WtParams& p, const WtThread& T, uint8_t* smem, WtBars* bars, float* exch_max,
                                        float* exch_sum, uint32_t tmem_base, uint32_t lt, bool lastrow, bool lastcol) {
  const int lane = T.lane, half = T.half;
  const int row = TILE2 ? 128 + lane : T.quadrant * 32 + lane;     // query index inside the window (garbage rows: >= 144)
  const int rowc = row < WT_N ? row : WT_N - 1;                     // clamped for address arithmetic only
  const int iy = rowc / WT_WS, ix = rowc - iy * WT_WS;
  // bias index of (query i, key j) = (iy - jy + 11) * 23 + (ix - jx + 11) = base - (jy * 23 + jx)
  constexpr int KP = WT_N / PARTS, PK = KP / 2;                    // keys / packed operand words of this thread (72 / 36 or 36 / 18)
  const int bias_base = (iy + WT_WS - 1) * (2 * WT_WS - 1) + ix + WT_WS - 1 - half * (KP / WT_WS) * (2 * WT_WS - 1);
  // slot of this row in the max / sum exchange arrays (160 per half: rows 0..127 of the big tile, 144..159 = rows 128..143 of
  // the small tile, 128..143 = scratch for the small tile's lanes 16..31, whose rows do not exist)
  const int er = TILE2 ? (lane < 16 ? 144 + lane : 128 + (lane - 16)) : row;
  const bool rf = iy >= WT_WS - p.shift, cf = ix >= WT_WS - p.shift;   // region flags of this query in a boundary window
  const float NEG = -100.0f * 1.4426950408889634f;
  const int s = lt % WT_STAGES, b = lt & 1;
  const float* sBias = reinterpret_cast<const float*>(smem + s * WT_STAGE_BYTES + 6 * WT_TILE_BYTES) + bias_base;
  mbar_wait(&bars->full[s], (lt / WT_STAGES) & 1);                 // the bias table of this stage has landed
  if (TILE2) mbar_wait(&bars->s2_full, lt & 1); else mbar_wait(&bars->s1_full[b], (lt >> 1) & 1);
  tc_fence_after();
  const bool stamp = !TILE2 && T.quadrant == 0 && half == 0 && lane == 0;
  if (stamp) WT_STAMP(lt, 6);
  const uint32_t saddr = tmem_base + T.lane_addr + (TILE2 ? WT_COL_S2 : WT_COL_S1 + b * WT_N);
  float x[KP];
  uint32_t* xv = reinterpret_cast<uint32_t*>(x);
  // shift mask (swin.py:416-440): -100 where query and key lie in different regions of a boundary window.  Keys of this half
  // all have jy >= 6 iff half == 1; jx >= 6 is a compile-time property of the unrolled index.
  const bool masked = lastrow || lastcol;
  const bool rowdiff = lastrow && ((half >= PARTS / 2) != rf);       // keys of this partition all have jy >= 6 iff it is in the upper half
  const float mlo = (rowdiff || (lastcol && cf)) ? NEG : 0.f, mhi = (rowdiff || (lastcol && !cf)) ? NEG : 0.f;
  float mm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};        // four independent max chains
  // x = S * scale * log2(e) + bias (+ mask) for keys [J0, J1), software-pipelined against the TMEM loads of the next chunk
  auto logits = [&](auto j0c, auto j1c) {
    constexpr int J0 = decltype(j0c)::value, J1 = decltype(j1c)::value;
    if (p.debug & 1) return;
    if (masked) {
#pragma unroll
      for (int jj = J0; jj < J1; ++jj) {
        const int off = (jj / WT_WS) * (2 * WT_WS - 1) + (jj % WT_WS);
        x[jj] = fmaf(x[jj], p.scale_log2e, sBias[-off]) + ((jj % WT_WS) < WT_WS / 2 ? mlo : mhi);
        mm[jj & 3] = fmaxf(mm[jj & 3], x[jj]);
      }
    } else {
#pragma unroll
      for (int jj = J0; jj < J1; ++jj) {
        const int off = (jj / WT_WS) * (2 * WT_WS - 1) + (jj % WT_WS);
        x[jj] = fmaf(x[jj], p.scale_log2e, sBias[-off]);
        mm[jj & 3] = fmaxf(mm[jj & 3], x[jj]);
      }
    }
  };
  if (PARTS == 2) {
    tmem_ld32(saddr + half * KP, xv);
    tmem_ld_wait_dep<32>(xv);
    tmem_ld32(saddr + half * KP + 32, xv + 32);                     // in flight under the first chunk's arithmetic
    logits(std::integral_constant<int, 0>{}, std::integral_constant<int, 32>{});
    tmem_ld_wait_dep<32>(xv + 32);
    tmem_ld8(saddr + half * KP + 64, xv + 64);
    logits(std::integral_constant<int, 32>{}, std::integral_constant<int, 64>{});
    tmem_ld_wait_dep<8>(xv + 64);
    logits(std::integral_constant<int, 64>{}, std::integral_constant<int, KP>{});
  } else if (PARTS == 4) {
    tmem_ld32(saddr + half * KP, xv);
    tmem_ld_wait_dep<32>(xv);
    tmem_ld4(saddr + half * KP + 32, xv + 32);
    logits(std::integral_constant<int, 0>{}, std::integral_constant<int, 32>{});
    tmem_ld_wait_dep<4>(xv + 32);
    logits(std::integral_constant<int, 32>{}, std::integral_constant<int, KP>{});
  } else {
    tmem_ld16(saddr + half * KP, xv);
    tmem_ld_wait_dep<16>(xv);
    tmem_ld8(saddr + half * KP + 16, xv + 16);
    logits(std::integral_constant<int, 0>{}, std::integral_constant<int, 16>{});
    tmem_ld_wait_dep<8>(xv + 16);
    logits(std::integral_constant<int, 16>{}, std::integral_constant<int, KP>{});
  }
  if (stamp) WT_STAMP(lt, 7);
  float m = (p.debug & 1) ? 0.f : fmaxf(fmaxf(mm[0], mm[1]), fmaxf(mm[2], mm[3]));
  // ---- row max across the two halves (every S column of this row has been read once both threads are here) ----
  float* emax = exch_max + (b * PARTS) * 160;
  emax[half * 160 + er] = m;
  wt_part_bar(T.pair_bar, 32 * PARTS);
#pragma unroll
  for (int o = 1; o < PARTS; ++o) m = fmaxf(m, emax[((half + o) % PARTS) * 160 + er]);
  if (stamp) WT_STAMP(lt, 8);
  // ---- un-normalised probabilities, bf16 hi / lo, written back over S: hi -> columns [0,72), lo -> [72,144) ----
  float sum = 0.f;
  uint32_t ph[PK], pl[PK];
  if (p.debug & 2) {                       // ablation: no exp / split
#pragma unroll
    for (int jj = 0; jj < PK; ++jj) { ph[jj] = __float_as_uint(x[jj]); pl[jj] = __float_as_uint(x[jj + PK]); }
    sum = 1.f;
  } else {
    float ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int jj = 0; jj < KP; jj += 2) {
      const float e0 = wt_ex2(x[jj] - m), e1 = wt_ex2(x[jj + 1] - m);
      ss[(jj >> 1) & 3] += e0 + e1;
      split_pack2(e0, e1, ph[jj >> 1], pl[jj >> 1]);
    }
    sum = (ss[0] + ss[1]) + (ss[2] + ss[3]);
  }
  if (PARTS == 2) {
    tmem_st32(saddr + half * PK, ph);
    tmem_st4(saddr + half * PK + 32, ph + 32);
    tmem_st32(saddr + 72 + half * PK, pl);
    tmem_st4(saddr + 72 + half * PK + 32, pl + 32);
  } else if (PARTS == 4) {
    tmem_st16(saddr + half * PK, ph);
    tmem_st2(saddr + half * PK + 16, ph + 16);
    tmem_st16(saddr + 72 + half * PK, pl);
    tmem_st2(saddr + 72 + half * PK + 16, pl + 16);
  } else {
    tmem_st8(saddr + half * PK, ph);
    tmem_st4(saddr + half * PK + 8, ph + 8);
    tmem_st8(saddr + 72 + half * PK, pl);
    tmem_st4(saddr + 72 + half * PK + 8, pl + 8);
  }
  if (stamp) WT_STAMP(lt, 9);
  tmem_st_wait();
  tc_fence_before();
  if (stamp) WT_STAMP(lt, 10);
  exch_sum[(b * PARTS + half) * 160 + er] = sum;
  __syncwarp();
  if (lane == 0) mbar_arrive(TILE2 ? &bars->p2_ready : &bars->p1_ready[b]);
}

// Epilogue of one tile of item `lt`: O rows * 1 / row sum -> bf16 hi / lo planes.  The partner's partial sum is visible: both
// threads of a row have passed a pair barrier (or the final one) since it was written.
template <bool TILE2, int PARTS>
__device__ __forceinline__ void wt_epilogue(const WtParams& p, const WtThread& T, WtBars* bars, const float* exch_sum,
                                            uint32_t tmem_base, uint32_t lt, int64_t row0, int head, bool release) {
  const int lane = T.lane;
  const int row = TILE2 ? 128 + lane : T.quadrant * 32 + lane;
  const int er = TILE2 ? (lane < 16 ? 144 + lane : 128 + (lane - 16)) : row;
  mbar_wait(TILE2 ? &bars->o2_full : &bars->o1_full, lt & 1);
  tc_fence_after();
  constexpr int OC = PARTS == 2 ? 16 : 8;                           // output channels of this thread; with 6 partitions only 0..3 carry any
  const bool has_out = T.half * OC < WT_D;
  const float* ps = exch_sum + ((lt & 1) * PARTS) * 160;
  float tot = 0.f;
#pragma unroll
  for (int o = 0; o < PARTS; ++o) tot += ps[o * 160 + er];
  const float inv = 1.0f / tot;
  uint32_t o[OC];
  if (PARTS == 2) tmem_ld16(tmem_base + T.lane_addr + (TILE2 ? WT_COL_O2 : WT_COL_O1) + T.half * OC, o);
  else if (has_out) tmem_ld8(tmem_base + T.lane_addr + (TILE2 ? WT_COL_O2 : WT_COL_O1) + T.half * OC, o);
  tmem_ld_wait();
  tc_fence_before();
  if (release) {
    __syncwarp();
    if (lane == 0) mbar_arrive(TILE2 ? &bars->o2_empty : &bars->o1_empty);
  }
  if (has_out && (!TILE2 || lane < 16)) {
    const int64_t g = (row0 + row) * p.C + head * WT_D + T.half * OC;
    if (PARTS == 2) wt_store16(p.out_hi, p.out_lo, g, o, inv); else wt_store8(p.out_hi, p.out_lo, g, o, inv);
  }
}

// The softmax + epilogue role of warps 4..11.  Every warp serves the 128-row tile of every item (its quadrant's 32 rows, its
// half of the keys); the 16-row tile of item i is placed on the lanes of quadrant i % 4 (the MMA issuer offsets the q rows
// accordingly) and served by that quadrant's two warps, so the extra pass rotates over the four SM sub-partitions instead of
// loading one of them with 4 of 10 heavy warp-passes (measured: that sub-partition bounded the kernel).
template <int PARTS>
__device__ __forceinline__ void wt_softmax_role(const WtParams& p, uint8_t* smem, WtBars* bars, float* exch_max, float* exch_sum,
                                                uint32_t tmem_base, int quadrant, int half, int lane) {
  WtThread T;
  T.quadrant = quadrant; T.half = half; T.lane = lane;
  T.lane_addr = (uint32_t)(quadrant * 32) << 16;
  T.pair_bar = 1 + quadrant;
  const bool tile2 = !(p.debug & 4);
  uint32_t lt = 0;
  int64_t prev_row0 = 0;
  int prev_head = 0;
  // item -> (window, head, boundary flags) without divisions in the loop: all counters advance by gridDim.x items per step
  const uint32_t nitems = (uint32_t)p.nitems, uheads = (uint32_t)p.heads, nWw = (uint32_t)p.nWw, nW = (uint32_t)(p.nWh * p.nWw);
  const uint32_t step_win = gridDim.x / uheads, step_head = gridDim.x % uheads;
  uint32_t win = blockIdx.x / uheads, head = blockIdx.x % uheads;
  uint32_t wxm = win % nWw, wim = win % nW;                       // window column / window index inside its image
  const uint32_t step_wx = step_win % nWw, step_wi = step_win % nW;
  for (uint32_t it = blockIdx.x; it < nitems; it += gridDim.x, ++lt) {
    const bool lastrow = p.shift > 0 && wim >= nW - nWw, lastcol = p.shift > 0 && wxm == nWw - 1;
    wt_pass<false, PARTS>(p, T, smem, bars, exch_max, exch_sum, tmem_base, lt, lastrow, lastcol);
    // the 16-row tile of the previous item, if this quadrant served it (its P.v was issued a whole pass ago, and the O2
    // buffer is needed again only after another quadrant's 16-row pass of this item)
    if (tile2 && lt > 0 && (int)((lt - 1) & 3) == quadrant) wt_epilogue<true, PARTS>(p, T, bars, exch_sum, tmem_base, lt - 1, prev_row0, prev_head, true);
    if (tile2 && (int)(lt & 3) == quadrant) wt_pass<true, PARTS>(p, T, smem, bars, exch_max, exch_sum, tmem_base, lt, lastrow, lastcol);
    // the 128-row tile of the previous item (its P.v has had a whole softmax to finish)
    if (lt > 0) wt_epilogue<false, PARTS>(p, T, bars, exch_sum, tmem_base, lt - 1, prev_row0, prev_head, true);
    if (!half && quadrant == 0 && lane == 0) WT_STAMP(lt, 14);
    prev_row0 = (int64_t)win * WT_N;
    prev_head = (int)head;
    head += step_head;
    uint32_t carry = 0;
    if (head >= uheads) { head -= uheads; carry = 1; }
    win += step_win + carry;
    wxm += step_wx + carry;
    if (wxm >= nWw) wxm -= nWw;
    if (wxm >= nWw) wxm -= nWw;
    wim += step_wi + carry;
    if (wim >= nW) wim -= nW;
    if (wim >= nW) wim -= nW;
  }
  if (lt > 0) {                                   // epilogues of the last item
    wt_part_bar(T.pair_bar, 32 * PARTS);           // the partners' sums of the last item are in shared memory
    if (tile2 && (int)((lt - 1) & 3) == quadrant) wt_epilogue<true, PARTS>(p, T, bars, exch_sum, tmem_base, lt - 1, prev_row0, prev_head, false);
    wt_epilogue<false, PARTS>(p, T, bars, exch_sum, tmem_base, lt - 1, prev_row0, prev_head, false);
  }
}

template <int PARTS>
__global__ void __launch_bounds__((4 + 4 * PARTS) * 32, 1)
window_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const WtParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* exch_max = reinterpret_cast<float*>(smem + WT_STAGES * WT_STAGE_BYTES);
  float* exch_sum = exch_max + WT_EXCH_FLOATS;
  WtBars* bars = reinterpret_cast<WtBars*>(smem + WT_STAGES * WT_STAGE_BYTES + 2 * WT_EXCH_FLOATS * 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_hi); prefetch_tmap(&tm_lo);
    for (int s = 0; s < WT_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 2); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bars->s1_full[b], 1); mbar_init(&bars->p1_ready[b], 4 * PARTS); }
    mbar_init(&bars->o1_full, 1); mbar_init(&bars->o1_empty, 4 * PARTS);
    mbar_init(&bars->s2_full, 1); mbar_init(&bars->p2_ready, PARTS);
    mbar_init(&bars->o2_full, 1); mbar_init(&bars->o2_empty, PARTS);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)), "r"(WT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp, one elected lane issues) =====================
    const uint32_t n = (uint32_t)((p.nitems - (int64_t)blockIdx.x + (int64_t)gridDim.x - 1) / (int64_t)gridDim.x);
    uint32_t s = 0, ph = 1;                               // fresh "empty" barriers pass a wait on parity 1
    int64_t it = blockIdx.x;
    for (uint32_t lt = 0; lt < n; ++lt, it += gridDim.x) {
      int64_t win;
      int head;
      wt_item(p, it, win, head);
      mbar_wait_sleep(&bars->empty[s], ph);
      WT_STAMP(lt, 0);
      if (elect_one()) {
        uint8_t* st = smem + s * WT_STAGE_BYTES;
        mbar_expect_tx(&bars->full[s], 6 * WT_TILE_BYTES + WT_BIAS_BYTES);
        if (p.t_hi) {                                       // tiled planes: every operand tile is 9216 contiguous bytes
#pragma unroll
          for (int part = 0; part < 3; ++part) {
            const size_t off = ((size_t)(win * 3 + part) * p.heads + head) * (WT_TILE_BYTES / 2);
            bulk_load(st + (2 * part) * WT_TILE_BYTES, p.t_hi + off, WT_TILE_BYTES, &bars->full[s]);
            bulk_load(st + (2 * part + 1) * WT_TILE_BYTES, p.t_lo + off, WT_TILE_BYTES, &bars->full[s]);
          }
        } else {
          const int r0 = (int)(win * WT_N);
#pragma unroll
          for (int part = 0; part < 3; ++part) {            // q, k, v column blocks of this head
            const int c0 = part * p.C + head * WT_D;
            tma_load_2d(st + (2 * part) * WT_TILE_BYTES, &tm_hi, &bars->full[s], c0, r0);
            tma_load_2d(st + (2 * part + 1) * WT_TILE_BYTES, &tm_lo, &bars->full[s], c0, r0);
          }
        }
        bulk_load(st + 6 * WT_TILE_BYTES, p.bias + (size_t)head * WT_BIAS_FLOATS, WT_BIAS_BYTES, &bars->full[s]);
      }
      __syncwarp();
      if (++s == WT_STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers: warp 1 = 128-row tile, warp 2 = 16-row tile =====================
    // The WHOLE warp walks the loop (uniform control flow, 32-bit uniform counters) and one elected lane issues: descriptors
    // then live in uniform registers.  (Under `if (lane == 0)` ptxas wraps every tcgen05.mma in an ELECT / R2UR.BROADCAST /
    // BRA.U.ANY loop -- measured ~70 clk per MMA, which made 66 small MMAs per item the bottleneck of the whole kernel.)
    constexpr uint32_t idS = wt_idesc(128, WT_N, false), idO = wt_idesc(128, WT_D, true);
    const bool t2 = warp == 2;
    const bool tiled = p.t_hi != nullptr;
    uint32_t n = (uint32_t)((p.nitems - (int64_t)blockIdx.x + (int64_t)gridDim.x - 1) / (int64_t)gridDim.x);
    if (t2 && (p.debug & 4)) n = 0;
    const uint32_t smem0 = smem_u32(smem);
    // first q row of this issuer's tile, in bytes.  The 16-row tile of item i is computed as q rows [128 - 32 (i & 3), + 128)
    // so that rows 128..143 land on TMEM lanes 32 (i & 3) .. + 15: the lanes of the quadrant whose softmax warps serve it
    uint32_t s_i = 0, ph_i = 0;                              // stage / phase of item i   (S issue)
    uint32_t s_j = 0;                                        // stage of item j = i - 1   (P v issue)
    const bool run = !(t2 && (p.debug & 4));
    for (uint32_t i = 0; run && i <= n; ++i) {
      const uint32_t j = i - 1;
      // ---- tile 2 frees its single S buffer first: O2(j) = P2 v ----
      if (t2 && i >= 1) {
        mbar_wait(&bars->p2_ready, j & 1);
        mbar_wait(&bars->o2_empty, (j & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sb = smem0 + s_j * WT_STAGE_BYTES;
          const uint64_t vh = tiled ? make_sdesc_ns(sb + 4 * WT_TILE_BYTES, 128, 2304) : make_sdesc64(sb + 4 * WT_TILE_BYTES);
          const uint64_t vl = tiled ? make_sdesc_ns(sb + 5 * WT_TILE_BYTES, 128, 2304) : make_sdesc64(sb + 5 * WT_TILE_BYTES);
          const uint32_t vstep = tiled ? 16u * 16u : 16u * 64u;                             // bytes per 16 keys
#pragma unroll
          for (int k = 0; k < ((p.debug & 8) ? 0 : WT_N / 16); ++k) {
            const uint64_t adv = (uint64_t)((k * vstep) >> 4);
            umma_bf16_ts(tmem_base + WT_COL_O2, tmem_base + WT_COL_S2 + k * 8, vh + adv, idO, k != 0);
            umma_bf16_ts(tmem_base + WT_COL_O2, tmem_base + WT_COL_S2 + k * 8, vl + adv, idO, 1);
            umma_bf16_ts(tmem_base + WT_COL_O2, tmem_base + WT_COL_S2 + 72 + k * 8, vh + adv, idO, 1);
          }
          umma_commit(&bars->o2_full);
          umma_commit(&bars->empty[s_j]);
        }
        __syncwarp();
        s_j = s_j + 1 == WT_STAGES ? 0 : s_j + 1;
      }
      // ---- S(i) = q k^T ----
      if (i < n) {
        mbar_wait(&bars->full[s_i], ph_i);
        tc_fence_after();
        if (!t2) WT_STAMP(i, 1);
        if (elect_one()) {
          const uint32_t sb = smem0 + s_i * WT_STAGE_BYTES;
          const uint32_t dcol = tmem_base + (t2 ? (uint32_t)WT_COL_S2 : (uint32_t)WT_COL_S1 + (i & 1) * WT_N);
          const uint32_t q_row = t2 ? 128u - 32u * (i & 3u) : 0u;
          const uint32_t q_off = q_row * (tiled ? 16u : (uint32_t)(WT_D * 2));           // bytes to the tile's first q row
          const uint64_t qh = tiled ? make_sdesc_ns(sb + q_off, 2304, 128) : make_sdesc64(sb + q_off);
          const uint64_t ql = tiled ? make_sdesc_ns(sb + WT_TILE_BYTES + q_off, 2304, 128) : make_sdesc64(sb + WT_TILE_BYTES + q_off);
          const uint64_t kh = tiled ? make_sdesc_ns(sb + 2 * WT_TILE_BYTES, 2304, 128) : make_sdesc64(sb + 2 * WT_TILE_BYTES);
          const uint64_t kl = tiled ? make_sdesc_ns(sb + 3 * WT_TILE_BYTES, 2304, 128) : make_sdesc64(sb + 3 * WT_TILE_BYTES);
          const uint32_t kstep = tiled ? 2u * 2304u : 32u;                                  // bytes per 16 channels
#pragma unroll
          for (int k = 0; k < ((p.debug & 16) ? 0 : WT_D / 16); ++k) {
            const uint64_t adv = (uint64_t)((k * kstep) >> 4);
            umma_bf16(dcol, qh + adv, kh + adv, idS, k != 0);
            umma_bf16(dcol, qh + adv, kl + adv, idS, 1);
            umma_bf16(dcol, ql + adv, kh + adv, idS, 1);
          }
          umma_commit(t2 ? &bars->s2_full : &bars->s1_full[i & 1]);
        }
        __syncwarp();
        if (!t2) WT_STAMP(i, 2);
        if (++s_i == WT_STAGES) { s_i = 0; ph_i ^= 1; }
      }
      // ---- tile 1: O1(j) = P1 v ----
      if (!t2 && i >= 1) {
        mbar_wait(&bars->p1_ready[j & 1], (j >> 1) & 1);
        WT_STAMP(j, 3);
        mbar_wait(&bars->o1_empty, (j & 1) ^ 1);
        tc_fence_after();
        WT_STAMP(j, 4);
        if (elect_one()) {
          const uint32_t sb = smem0 + s_j * WT_STAGE_BYTES;
          const uint32_t pcol = tmem_base + WT_COL_S1 + (j & 1) * WT_N;
          const uint64_t vh = tiled ? make_sdesc_ns(sb + 4 * WT_TILE_BYTES, 128, 2304) : make_sdesc64(sb + 4 * WT_TILE_BYTES);
          const uint64_t vl = tiled ? make_sdesc_ns(sb + 5 * WT_TILE_BYTES, 128, 2304) : make_sdesc64(sb + 5 * WT_TILE_BYTES);
          const uint32_t vstep = tiled ? 16u * 16u : 16u * 64u;
#pragma unroll
          for (int k = 0; k < ((p.debug & 8) ? 0 : WT_N / 16); ++k) {
            const uint64_t adv = (uint64_t)((k * vstep) >> 4);
            umma_bf16_ts(tmem_base + WT_COL_O1, pcol + k * 8, vh + adv, idO, k != 0);
            umma_bf16_ts(tmem_base + WT_COL_O1, pcol + k * 8, vl + adv, idO, 1);
            umma_bf16_ts(tmem_base + WT_COL_O1, pcol + 72 + k * 8, vh + adv, idO, 1);
          }
          umma_commit(&bars->o1_full);
          umma_commit(&bars->empty[s_j]);
          if (p.debug & 4) umma_commit(&bars->empty[s_j]);
        }
        __syncwarp();
        WT_STAMP(j, 5);
        s_j = s_j + 1 == WT_STAGES ? 0 : s_j + 1;
      }
    }
  } else if (warp >= 4) {
    wt_softmax_role<PARTS>(p, smem, bars, exch_max, exch_sum, tmem_base, warp & 3, (warp - 4) >> 2, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(WT_TMEM_COLS) : "memory");
  }
}

// bf16 [rows][cols] row-major -> 2-D map, box = (32 columns, 144 rows), SWIZZLE_64B
static int make_map_wattn(CUtensorMap* m, const uint16_t* ptr, int64_t rows, int64_t cols) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(RBA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)WT_D, (cuuint32_t)WT_N};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(RBA_ERR_CUDA, "cuTensorMapEncodeTiled(window attention) failed with %d", (int)r);
  return RBA_OK;
}

int window_attn_tc(const uint16_t* qkv_hi, const uint16_t* qkv_lo, const float* bias_prepared, int B, int H, int W, int C,
                   int heads, int ws, int shift, uint16_t* out_hi, uint16_t* out_lo, cudaStream_t st, int tiled) {
  RBA_CHECK(qkv_hi && qkv_lo && bias_prepared && out_hi && out_lo, "window_attn_tc: null pointer");
  RBA_CHECK(ws == WT_WS, "window_attn_tc: only window_size 12 is built (got %d)", ws);
  RBA_CHECK(heads > 0 && C == heads * WT_D, "window_attn_tc: head_dim must be 32 (C=%d heads=%d)", C, heads);
  RBA_CHECK(shift == 0 || shift == ws / 2, "window_attn_tc: shift must be 0 or window_size / 2 (got %d)", shift);
  RBA_CHECK((((uintptr_t)qkv_hi | (uintptr_t)qkv_lo | (uintptr_t)out_hi | (uintptr_t)out_lo | (uintptr_t)bias_prepared) & 15) == 0,
            "window_attn_tc: pointers must be 16-byte aligned");
  SwinGeom g = make_swin_geom(H, W, ws, shift);
  const int64_t nwin = (int64_t)B * g.nWh * g.nWw;
  if (nwin == 0) return RBA_OK;
  RBA_CHECK(nwin * WT_N < (1LL << 31), "window_attn_tc: too many rows");
  CUtensorMap tm_hi, tm_lo;
  memset(&tm_hi, 0, sizeof(tm_hi));
  memset(&tm_lo, 0, sizeof(tm_lo));
  if (!tiled) {
    RBA_TRY_(make_map_wattn(&tm_hi, qkv_hi, nwin * WT_N, 3 * (int64_t)C));
    RBA_TRY_(make_map_wattn(&tm_lo, qkv_lo, nwin * WT_N, 3 * (int64_t)C));
  }
  WtParams p;
  memset(&p, 0, sizeof(p));
  if (tiled) { p.t_hi = qkv_hi; p.t_lo = qkv_lo; }
  p.bias = bias_prepared; p.out_hi = out_hi; p.out_lo = out_lo;
  p.C = C; p.heads = heads; p.nWh = g.nWh; p.nWw = g.nWw; p.shift = shift;
  p.nitems = nwin * heads;
  p.scale_log2e = 1.4426950408889634f / sqrtf((float)WT_D);
  { static const int dbg = []() { const char* e = getenv("RBA_WT_DEBUG"); return e ? atoi(e) : 0; }(); p.debug = dbg; }   // profiling ablations
  // RBA_WT_PARTS: softmax warps per TMEM lane quadrant (2: 12 warps, 72 keys per thread; 4: 20 warps, 36 keys per thread)
  static const int parts = []() { const char* e = getenv("RBA_WT_PARTS"); const int v = e ? atoi(e) : 4; return v == 2 || v == 6 ? v : 4; }();
  static PerDeviceOnce once;
  if (once.needed()) {
    RBA_CUDA(cudaFuncSetAttribute(window_attn_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM));
    RBA_CUDA(cudaFuncSetAttribute(window_attn_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM));
    RBA_CUDA(cudaFuncSetAttribute(window_attn_tc_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM));
    once.done();
  }
  const unsigned grid = (unsigned)std::min<int64_t>(p.nitems, num_sms());
  static const bool timeline = getenv("RBA_WT_TIMELINE") != nullptr;       // profiling aid: prints CTA 0's clock stamps
  static long long* tl_dev = nullptr;
  if (timeline) {
    if (!tl_dev) RBA_CUDA(cudaMalloc(&tl_dev, 32 * 16 * sizeof(long long)));
    RBA_CUDA(cudaMemsetAsync(tl_dev, 0, 32 * 16 * sizeof(long long), st));
    p.tl = tl_dev;
  }
  if (parts == 2) window_attn_tc_kernel<2><<<grid, (4 + 4 * 2) * 32, WT_SMEM, st>>>(tm_hi, tm_lo, p);
  else if (parts == 6) window_attn_tc_kernel<6><<<grid, (4 + 4 * 6) * 32, WT_SMEM, st>>>(tm_hi, tm_lo, p);
  else window_attn_tc_kernel<4><<<grid, (4 + 4 * 4) * 32, WT_SMEM, st>>>(tm_hi, tm_lo, p);
  RBA_LAUNCHED();
  if (timeline) {
    long long h[32 * 16];
    RBA_CUDA(cudaStreamSynchronize(st));
    RBA_CUDA(cudaMemcpy(h, tl_dev, sizeof(h), cudaMemcpyDeviceToHost));
    const long long t0 = h[0];
    fprintf(stderr, "[wattn_tc timeline, CTA 0, clocks since first TMA issue]\n item  tma  full s1iss  p1rdy o1emp pviss | s1full  ld   bar   st   stw | epi: start o1full ldarr | end\n");
    for (int i = 0; i < 12; ++i) {
      fprintf(stderr, "%5d", i);
      for (int e = 0; e < 15; ++e) fprintf(stderr, " %6lld", h[i * 16 + e] ? h[i * 16 + e] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  return RBA_OK;
}

}  // namespace rba

extern "C" int rba_k_window_attn_tc(const uint16_t* qkv_hi, const uint16_t* qkv_lo, const float* bias_prepared, int B, int H, int W,
                                    int C, int heads, int ws, int shift, uint16_t* out_hi, uint16_t* out_lo, void* stream) {
  return rba::window_attn_tc(qkv_hi, qkv_lo, bias_prepared, B, H, W, C, heads, ws, shift, out_hi, out_lo, (cudaStream_t)stream, 0);
}
extern "C" int rba_k_window_attn_tc_tiled(const uint16_t* qkv_hi, const uint16_t* qkv_lo, const float* bias_prepared, int B, int H,
                                          int W, int C, int heads, int ws, int shift, uint16_t* out_hi, uint16_t* out_lo,
                                          void* stream) {
  return rba::window_attn_tc(qkv_hi, qkv_lo, bias_prepared, B, H, W, C, heads, ws, shift, out_hi, out_lo, (cudaStream_t)stream, 1);
}

