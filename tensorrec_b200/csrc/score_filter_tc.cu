// K2+K3 fused, filter form: ONE tcgen05 pass over the fp16 "hi" halves produces approximate scores with a proven
// error bound; per user row the kernel keeps every item that could still belong to the top-k (approximate score
// within the bound of the running k-th best).  The few survivors are re-scored exactly in fp32 and ranked by
// rescore_topk_kernel (rescore_topk.cu), which also verifies the bound and flags rows for the exact 3-pass kernel
// (score_topk_tc.cu) if it does not hold.  Same reference chain as score_topk_tc.cu: tf.matmul
// (tensorrec/prediction_graphs.py:49-50), bias_prediction_dense (tensorrec/recommendation_graphs.py:41),
// rank_predictions (:73-82) restricted to rank <= k.
//
// CTA = 256 user rows (two 128-row blocks) x a sweep over 128-item tiles: every B tile fetched from L2 feeds two
// accumulators.  The user rows live in TENSOR MEMORY (columns [0,128), written by the epilogue threads with
// tcgen05.st) and are the A operand of tcgen05.mma straight from there; the remaining 384 columns hold a ring of
// three 128-column accumulators.  Epilogue group g = warps 4+4g..7+4g drains the accumulators of user block g, one
// thread per user row.
//
// Why: the exact split-product kernel issues 3 tensor passes and its per-row sorted-list inserts serialise a warp
// (profiles/r1_v2_fused_ncu.json: tensor pipe 37 %, top stall = insert loop).  Here
//   * tensor work is 1 pass (2*U*I*d flops = the algorithmic count);
//   * items are processed in descending-bias order (host side), so within a 128-item block the biases are almost equal
//     and the admission test v_j = acc_j + bias_j / c > tau (c = user scale x GLOBAL item scale, both powers of two)
//     is bounded by max_j acc_j + blockmax / c: the hot loop is an FMNMX3 tree over the raw accumulators, one add and
//     one warp vote per 32 columns;
//   * a passing column is APPENDED raw (accumulator, position) to the row's 32-entry buffer in shared memory; when
//     some row's buffer passes half full the whole warp compacts it cooperatively: one entry per lane, raw entries
//     resolved to (approximate score, item id), a 15-step bitonic sort through shuffles, keep everything >= (k-th best
//     - 2.25m), tighten the threshold;
//   * the epilogue warps never meet at a block or group barrier.
//
// Error bound.  hi = fp16(x * 2^e) has relative error <= 2^-11 per element (absolute 2^-25 below the fp16 normal
// range), so |approx - exact| <= (2^-10 + 2^-22) |u|.|i| <= m := kMarginFactor * |u|_2 * max_j |i_j|_2 with
// kMarginFactor = 1.5 * 2^-10 (covers the fp32 accumulation of the tensor core and the flush of tiny elements).
// Every item ever excluded had approx <= theta_final, hence exact <= theta_final + m, and theta = a_k - 2.25m keeps
// theta + m strictly below the exact k-th best of the survivors (which is >= a_k - m).  rescore_topk_kernel checks
// exactly that inequality.
#include <stdlib.h>

#include "common.cuh"

namespace trk {

constexpr int kFBlockM = 128;
constexpr int kFBlockN = 128;          // item tile; TMEM: 128 columns of A operand + 3 accumulators of 128 columns
constexpr int kFAccSlots = 3;
constexpr uint32_t kFTmemAccCol = 128;   // first accumulator column (columns [0, 128): user operand, 64 per block)
constexpr int kFKBlock = 64;
constexpr int kFUmmaK = 16;
constexpr int kFThreads = 384;
constexpr uint32_t kFBTileBytes = kFBlockN * kFKBlock * 2;   // 16 KB
constexpr int kFMaxStages = 10;
constexpr uint32_t kFTmemCols = 512;
constexpr int kBufEntries = 32;      // candidate buffer per (row, epilogue group)
constexpr int kKeepMax = 16;         // entries kept by a compaction (>= k + slack); also the per-group output width
constexpr int kFilterMaxK = 12;
constexpr float kMarginFactor = 1.5f * 0.0009765625f;   // 1.5 * 2^-10
constexpr float kBiasUlps = 4.0f * 1.1920929e-7f;        // 4 ulp(1): rounding of (dot + ub) + ib
constexpr float kThetaMargins = 2.25f;                   // theta = a_k - 2.25 m  (> 2 m is what the proof needs)
constexpr int kTileEndMaxTiles = 3072;   // sweeps of up to 393216 items per split run the tile-end compaction variant
constexpr int kGiveUpOverflows = 8;   // a row whose compactions overflow this often is handed to the exact kernel

struct FilterParams {
  const __half* user_split;    // [n_users, 2 d_pad] hi | lo; the filter reads the hi half
  const float* user_scale;
  const float* user_bias;      // may be null
  const float* user_norm;      // |u|_2 per user
  const float* item_bias;      // [padded items] in PROCESSING order (see item_perm), padding = -inf
  const float* block_bias_max; // max item bias of every block of 128 processing positions (-inf for all-padding)
  const float* block_bias_min; // min item bias of every block (-inf as soon as the block holds padding); may be null
  const int32_t* item_perm;    // processing position -> local item index (items sorted by bias), or null = identity
  const float* item_stats;     // device: [0] = max_j |i_j|_2, [1] = global item scale (2^-E), [2] = max_j |bias_j|
  int64_t n_users;
  int64_t n_items;
  int32_t n_kblocks;           // d_pad / 64
  int32_t d_pad;
  int32_t n_stages;
  int32_t k;
  int32_t n_splits;
  int32_t tiles_per_split;
  int32_t n_tiles;
  int32_t n_user_pairs;        // ceil(n_users / 256)
  int32_t item_id_offset;
  int32_t tile_end_trigger;    // kTileEnd kernels: rows holding more entries than this are compacted at the END of a tile
  int32_t debug_mode;          // timing experiments only (TRK_FILTER_DEBUG): 1 = drain TMEM without filtering, 2 = no drain,
                               // 4 = nothing admitted, 6 = MMA only (no B stream, no drain), 7 = full kernel + clock readout
                               // (9, "filter half of every tile's columns", was a build-time experiment: its run-time
                               // loop bound cost 2.5 % at 1M items -- profiles/probe_r2_v18_filter_ab_full.txt)
  float* cand_score;           // [n_users, n_splits, kKeepMax] approximate scores (sentinel -inf)
  int32_t* cand_item;          // [n_users, n_splits, kKeepMax] global ids (sentinel INT32_MAX)
  float* row_theta;            // [n_users, n_splits] final admission threshold (certified by rescore_topk_kernel)
};

struct FilterLayout {
  uint32_t b_off, buf_off, bar_off, total;
};
__host__ __device__ inline FilterLayout filter_layout(int n_stages) {
  FilterLayout L;
  L.b_off = 0;
  L.buf_off = L.b_off + static_cast<uint32_t>(n_stages) * kFBTileBytes;
  L.bar_off = L.buf_off + 2u * kFBlockM * kBufEntries * 8u;       // 64 KB of candidate buffers
  L.total = L.bar_off + 512u;
  return L;
}
// The B ring is organised in TILE slots of n_kblocks k-blocks (16 KB each): one full / one empty barrier per item tile.
// barriers (uint64): [0..1] a_full (per user block) [2..4] tmem_full [5..7] tmem_empty [8 .. 8+T) b_full
// [8+T .. 8+2T) b_empty, T = n_stages / n_kblocks tile slots; TMEM base address (uint32) at byte 400 of the block.
// Accumulators: (tile it, user block b) is number q = 2 it + b and lives in TMEM slot q % 3 (its n-th use, n = q / 3,
// has barrier parity n & 1).  Epilogue group b drains the accumulators of user block b: while it works on one, the
// MMA warp can fill the next of either block.
// The item biases are NOT staged: the hot loop needs only the block maximum (one cached global load per tile,
// prefetched a tile ahead) and the rare admission path reads the few biases it needs through L2.  (A two-slot
// shared-memory ring fed by one bulk copy per tile put the copy's ~2 us latency on the critical path of every
// second tile: TRK_FILTER_DEBUG=6 showed 112 cycles per MMA step against 64 for the bare instruction stream.)

__device__ __forceinline__ void f_sts64(uint32_t addr, float s, int32_t id) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(__float_as_uint(s)), "r"(id) : "memory");
}
__device__ __forceinline__ void f_lds64(uint32_t addr, float* s, int32_t* id) {
  uint32_t a, b;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr) : "memory");
  *s = __uint_as_float(a);
  *id = static_cast<int32_t>(b);
}
// (score desc, id asc): does x come before y ?
__device__ __forceinline__ bool cand_before(float xs, int32_t xi, float ys, int32_t yi) {
  return xs > ys || (xs == ys && xi < yi);
}

// Inputs of the admission path that are the same for the whole kernel.
struct AdmitCtx {
  const float* bias;     // item biases in processing order, padded with -inf
  const int32_t* perm;   // processing position -> local item index, or null = identity
  int32_t id_offset;
  int32_t n_items;
};

// Warp-cooperative compaction of the candidate buffer of lane `src`'s row: one entry per lane, bitonic sort by
// (score desc, id asc), keep everything >= k-th best - 2.25m (at most kKeepMax), tighten that row's thresholds.
// Entries [n_res, cnt) of the buffer are RAW -- (accumulator value, processing position) exactly as the hot loop
// found them; they are turned into (approximate score, item id) here, where the bias and permutation lookups of all
// of them are independent loads issued by different lanes (one L2 latency per compaction, not one per admission).
//
// Split in two so that the lookups of the NEXT row to compact are in flight while THIS row is sorted: the rows of a warp
// overflow in bursts (their thresholds rise in step), a compaction is ~250 instructions, an L2 round trip ~700 cycles.
struct RowFetch {
  float s;        // raw accumulator (lanes >= n_res) or resolved approximate score
  int32_t id;     // processing position (raw) or item id (resolved)
  float bias;     // bias of the raw entry's position           } requested by compact_fetch,
  int32_t perm;   // original item index of that position       } first used by compact_finish
  int n, n_res;
  uint32_t addr;
};
__device__ __forceinline__ float ldg_nc_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int32_t ldg_nc_s32(const int32_t* p) {
  int32_t v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ RowFetch compact_fetch(uint32_t buf_row_addr, int lane, int src, int cnt, int n_res,
                                                  const AdmitCtx& ctx) {
  RowFetch f;
  f.n = __shfl_sync(0xffffffffu, cnt, src);
  f.n_res = __shfl_sync(0xffffffffu, n_res, src);
  f.addr = __shfl_sync(0xffffffffu, buf_row_addr, src);
  f.s = -__int_as_float(0x7f800000);
  f.id = 0x7fffffff;
  f.bias = 0.0f;
  f.perm = 0;
  if (lane < f.n) {
    f_lds64(f.addr + lane * 8, &f.s, &f.id);
    if (lane >= f.n_res) {
      f.bias = ldg_nc_f32(ctx.bias + f.id);
      // (a padded column of the last tile can be appended -- its bias is -inf, it never survives -- and has no perm entry)
      f.perm = (ctx.perm != nullptr && f.id < ctx.n_items) ? ldg_nc_s32(ctx.perm + f.id) : f.id;
    }
  }
  return f;
}
// n_ovf counts the compactions of this lane's row that found more than kKeepMax entries within the bound of the k-th
// best.  Once in a while that is harmless (the surplus is remembered in drop_max).  A row where it keeps happening holds
// massive near-ties (all-equal scores in the limit: EVERY column passes, every chunk takes the slow path and the sweep of
// 1M x 1M took 43 s instead of 0.2): after kGiveUpOverflows of them the row stops admitting (tau = +inf) and is marked
// uncertifiable (drop_max = +inf), i.e. it goes through the exact kernel -- where a tie-heavy row belongs anyway.
__device__ __forceinline__ void compact_finish(const RowFetch& f, int lane, int src, int k, int& cnt, int& n_res,
                                               float& theta, float& tau, float& drop_max, int& n_ovf, float m3,
                                               float ubias, float c, float inv_c, const AdmitCtx& ctx) {
  const float kNegInf = -__int_as_float(0x7f800000);
  const int n = f.n;
  const uint32_t addr = f.addr;
  const float m3s = __shfl_sync(0xffffffffu, m3, src);
  const float cs = __shfl_sync(0xffffffffu, c, src);
  const float ubs = __shfl_sync(0xffffffffu, ubias, src);
  float s = f.s;
  int32_t id = f.id;
  if (lane < n && lane >= f.n_res) {
    const int32_t pos = id;
    id = pos < ctx.n_items ? ctx.id_offset + f.perm : 0x7fffffff;
    s = fmaf(s, cs, ubs) + f.bias;   // approximate score: (acc * c + user bias) + item bias
  }
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, s, stride);
      const int32_t oi = __shfl_xor_sync(0xffffffffu, id, stride);
      const bool lower = (lane & stride) == 0;           // this lane holds the earlier position of the pair
      const bool descending = (lane & size) == 0;        // block direction: "before" elements first
      const bool other_first = cand_before(os, oi, s, id);
      // earlier position wants the element that comes first in a descending block (and vice versa)
      const bool take_other = (lower == descending) ? other_first : !other_first;
      if (take_other) {
        s = os;
        id = oi;
      }
    }
  }
  // lanes now hold the entries in (score desc, id asc) order, sentinels last
  const float kth = __shfl_sync(0xffffffffu, s, k - 1);
  const bool have_k = n >= k;
  const float floor_s = have_k ? kth - m3s : kNegInf;
  const unsigned keep = __ballot_sync(0xffffffffu, lane < n && s >= floor_s);
  int n_keep = __popc(keep);
  const bool ovf = n_keep > kKeepMax;
  n_keep = ovf ? kKeepMax : n_keep;
  // more than kKeepMax entries crowd within the bound of the k-th best: the surplus is dropped and the best dropped
  // score is remembered -- it only matters if it is still close to the k-th best at the END of the sweep (the
  // verification in rescore_topk_kernel compares max(theta, drop_max) + m with the exact k-th best)
  const float first_dropped = __shfl_sync(0xffffffffu, s, kKeepMax & 31);
  if (lane < n_keep) f_sts64(addr + lane * 8, s, id);
  if (lane == src) {
    cnt = n_keep;
    n_res = n_keep;
    if (ovf) {
      drop_max = fmaxf(drop_max, first_dropped);
      n_ovf += 1;
    }
    if (have_k) {
      theta = floor_s;
      // admission test runs on v = acc + bias/c; move theta there and leave a few ulps of slack (extra survivors are
      // harmless, a missed one is not)
      const float t = (theta - ubias) * inv_c;
      tau = t - 8.0f * 1.1920929e-7f * fabsf(t) - 1e-30f;
    }
    if (n_ovf >= kGiveUpOverflows) {
      tau = __int_as_float(0x7f800000);
      drop_max = __int_as_float(0x7f800000);
    }
  }
  __syncwarp();
}
// compacts every row of `rows` (bit = lane), the lookups of the next row in flight while the current one is sorted
__device__ __forceinline__ void compact_rows(unsigned rows, uint32_t buf_row_addr, int lane, int k, int& cnt, int& n_res,
                                             float& theta, float& tau, float& drop_max, int& n_ovf, float m3,
                                             float ubias, float c, float inv_c, const AdmitCtx& ctx) {
  if (rows == 0u) return;
  int src = __ffs(rows) - 1;
  rows &= rows - 1;
  RowFetch cur = compact_fetch(buf_row_addr, lane, src, cnt, n_res, ctx);
  while (true) {
    const int nxt = rows != 0u ? __ffs(rows) - 1 : -1;
    rows &= rows - 1;     // (0 stays 0)
    RowFetch ahead = cur;
    if (nxt >= 0) ahead = compact_fetch(buf_row_addr, lane, nxt, cnt, n_res, ctx);
    compact_finish(cur, lane, src, k, cnt, n_res, theta, tau, drop_max, n_ovf, m3, ubias, c, inv_c, ctx);
    if (nxt < 0) break;
    cur = ahead;
    src = nxt;
  }
}

// 16 columns of one user row per lane.  The admission test is v_j = acc_j + bias_j / c > tau.  Items are processed in
// bias-sorted order, so the biases of one 128-item block differ by ~1e-4 of their range and v_j <= max_j acc_j +
// bmax_block / c is a tight upper bound: the fast path is a pure FMNMX3 reduction of the raw accumulators plus ONE add
// (no per-score bias load, no per-score FFMA) and one vote; only when some lane's bound passes are the exact v_j
// formed.  The hitting lanes then append their survivors (approximate score + original item id) and rows whose buffer
// passed half full are compacted by the whole warp.
// (Voting once per 32 columns instead measured 7-11 % SLOWER at 1M x 1M x d128; the 16-column granularity stays.)
// maximum of 16 columns; g[q] = maximum of columns [4q, 4q + 4) (the slow path looks only into the groups that pass)
__device__ __forceinline__ float acc_max_16(const uint32_t* acc, float (&g)[4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    g[q] = fmaxf(fmaxf(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1])),
                 fmaxf(__uint_as_float(acc[4 * q + 2]), __uint_as_float(acc[4 * q + 3])));
  return fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
}
__device__ __forceinline__ float acc_max_16(const uint32_t* acc) {
  float g[4];
  return acc_max_16(acc, g);
}

// slow path of 16 columns: the lanes whose bound passed append every column that passes the same bound as a raw
// (accumulator, position) entry -- a superset of the exact test acc_j + bias_j / c > tau, since bias_j <= block max and
// rounding is monotonic; no memory is read here.  A row is compacted (by the whole warp) only when its buffer could not
// take the new entries: ~4 compactions per user at 1M items instead of 9 with a "more than half full" trigger.
// Called warp-uniformly.
__device__ __forceinline__ void admit_16(const uint32_t* acc, bool hit, int32_t pos_base, float bmax_scaled,
                                         const AdmitCtx& ctx, float c, float inv_c, float ubias, float& tau,
                                         float& theta, float& drop_max, int& n_ovf, float m3, uint32_t buf_row_addr,
                                         int& cnt, int& n_res, int lane, int k) {
  uint32_t pass = 0;
  if (hit) {
#pragma unroll
    for (int j = 0; j < 16; ++j) pass |= (__uint_as_float(acc[j]) + bmax_scaled > tau) ? (1u << j) : 0u;
  }
  __syncwarp();   // earlier appends of every lane are visible to the lanes that may now compact its row
  const unsigned need = __ballot_sync(0xffffffffu, cnt + __popc(pass) > kBufEntries);
  compact_rows(need, buf_row_addr, lane, k, cnt, n_res, theta, tau, drop_max, n_ovf, m3, ubias, c, inv_c, ctx);
  if (pass != 0 && n_ovf < kGiveUpOverflows) {   // (a row that has just given up appends nothing more)   // cnt + popc(pass) <= kBufEntries holds here (a compaction leaves at most kKeepMax = 16)
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if ((pass >> j) & 1u) {
        f_sts64(buf_row_addr + cnt * 8, __uint_as_float(acc[j]), pos_base + j);
        cnt += 1;
      }
    }
  }
}

// Bit mask of the columns of acc[0, 16) whose admission bound passes.  g[q] = maximum of columns [4q, 4q + 4) from the
// hot loop: only a group whose maximum passes is looked into (x -> x + bmax is monotonic), so the usual single hit
// costs 4 + 4 compares instead of 16; independent compares, OR'ed pairwise (a serial `mask |= ...` chain put ~80 cycles
// of dependent latency into every slow-path entry).
__device__ __forceinline__ uint32_t pass_mask_16(const uint32_t* acc, const float (&g)[4], float bmax_scaled,
                                                 float tau) {
  uint32_t mask = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (g[q] + bmax_scaled > tau) {
      const uint32_t b0 = (__uint_as_float(acc[4 * q + 0]) + bmax_scaled > tau) ? (1u << (4 * q + 0)) : 0u;
      const uint32_t b1 = (__uint_as_float(acc[4 * q + 1]) + bmax_scaled > tau) ? (1u << (4 * q + 1)) : 0u;
      const uint32_t b2 = (__uint_as_float(acc[4 * q + 2]) + bmax_scaled > tau) ? (1u << (4 * q + 2)) : 0u;
      const uint32_t b3 = (__uint_as_float(acc[4 * q + 3]) + bmax_scaled > tau) ? (1u << (4 * q + 3)) : 0u;
      mask |= (b0 | b1) | (b2 | b3);
    }
  }
  return mask;
}
// Appends the columns of `mask` (a 16-column half, its maximum `amax` known from the hot loop).  One passing column -- the
// normal case -- IS the maximum of its half (x -> x + bmax is monotonic, so the largest accumulator passes whenever any
// does): one store, no search through the registers.  Several: every slot is addressed by a prefix popcount, the
// stores are independent.
__device__ __forceinline__ void append_16(const uint32_t* acc, uint32_t mask, float amax, int32_t pos_base,
                                          uint32_t buf_row_addr, int& cnt) {
  if ((mask & (mask - 1u)) == 0u) {
    f_sts64(buf_row_addr + cnt * 8, amax, pos_base + __ffs(mask) - 1);
    cnt += 1;
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if ((mask >> j) & 1u)
        f_sts64(buf_row_addr + (cnt + __popc(mask & ((1u << j) - 1u))) * 8, __uint_as_float(acc[j]), pos_base + j);
    }
    cnt += __popc(mask);
  }
}

// 32 columns behind ONE vote (the two 16-column maxima are independent chains).  Slow path, taken when some lane's bound
// passes (~8 % of the chunks of a 1M-item sweep, ~40 % at a 125K-item shard: 32 rows share the instruction stream): the
// hitting lanes form the pass masks, ONE ballot tells whether every row's buffer can take its new entries -- the common
// case: the hitting lanes append, nobody else does anything -- and only otherwise the chunk goes through the two-step
// path (compact the rows that need it, append 16 columns at a time so that the 32-entry buffer cannot overflow between
// compactions).
__device__ __forceinline__ void filter_32(const uint32_t* acc, int32_t pos_base, float bmax_scaled, const AdmitCtx& ctx,
                                          float c, float inv_c, float ubias, float& tau, float& theta,
                                          float& drop_max, int& n_ovf, float m3, uint32_t buf_row_addr, int& cnt,
                                          int& n_res, int lane, int k) {
  float g0[4], g1[4];
  const float a0 = acc_max_16(acc, g0), a1 = acc_max_16(acc + 16, g1);
  const bool h0 = a0 + bmax_scaled > tau, h1 = a1 + bmax_scaled > tau;
  if (__any_sync(0xffffffffu, h0 || h1)) {
    uint32_t lo = 0, hi = 0;
    if (h0) lo = pass_mask_16(acc, g0, bmax_scaled, tau);
    if (h1) hi = pass_mask_16(acc + 16, g1, bmax_scaled, tau);
    if (__ballot_sync(0xffffffffu, cnt + __popc(lo) + __popc(hi) > kBufEntries) == 0u) {
      if (lo != 0u) append_16(acc, lo, a0, pos_base, buf_row_addr, cnt);
      if (hi != 0u) append_16(acc + 16, hi, a1, pos_base + 16, buf_row_addr, cnt);
    } else {
      admit_16(acc, h0, pos_base, bmax_scaled, ctx, c, inv_c, ubias, tau, theta, drop_max, n_ovf, m3, buf_row_addr, cnt,
               n_res, lane, k);
      admit_16(acc + 16, h1, pos_base + 16, bmax_scaled, ctx, c, inv_c, ubias, tau, theta, drop_max, n_ovf, m3,
               buf_row_addr, cnt, n_res, lane, k);
    }
  }
}

// ---- first tile of a work unit: a threshold to start from -----------------------------------------------------------
// With tau = -inf every column of the first tile is admitted: 128 appends and 8 warp-cooperative compactions per row,
// 32 rows of a warp one after the other (~50 us per 256-user unit; 2 % of a 1M-item sweep but 15 % of a 125K-item
// shard, which is what held the 8-GPU run at 0.74 of the tensor peak per shard).  Instead every thread first reduces
// its row of the first accumulator to 16 group maxima (8 columns each), sorts them in registers and takes the k-th
// largest, A: k DIFFERENT columns have acc >= A, their biases are >= the block minimum, so the k-th best approximate
// score of the tile is >= fma(A, c, ub) + bmin and theta may start 2.25 m below that.  The tile is then filtered as
// usual: ~1.5 k admissions per row instead of 128, no compaction.  (Measured at the 125K-item shard: -1.3 ms of 35.
// Scanning MORE tiles first -- 4 to 32 tiles reduced to a running top-16 of group maxima, then filtered in a second
// pass -- gained nothing: profiles/probe_r2_filter_shard8_scan_prologue.txt.  The cost of the admission path is
// proportional to the number of 32-column chunks in which ANY of a warp's 32 rows passes its bound, ~25 ns of SM time
// each at either size, and the early tiles are few chunks whatever they admit.)
__device__ __forceinline__ float acc_max_8(const uint32_t* acc) {
  return fmaxf(fmaxf(fmaxf(__uint_as_float(acc[0]), __uint_as_float(acc[1])),
                     fmaxf(__uint_as_float(acc[2]), __uint_as_float(acc[3]))),
               fmaxf(fmaxf(__uint_as_float(acc[4]), __uint_as_float(acc[5])),
                     fmaxf(__uint_as_float(acc[6]), __uint_as_float(acc[7]))));
}
// bitonic network on 16 registers, descending; every index is a compile-time constant after unrolling
__device__ __forceinline__ void sort16_desc(float (&g)[16]) {
#pragma unroll
  for (int size = 2; size <= 16; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int j = i ^ stride;
        if (j > i) {
          const bool desc = (i & size) == 0;
          const float hi = fmaxf(g[i], g[j]), lo = fminf(g[i], g[j]);
          g[i] = desc ? hi : lo;
          g[j] = desc ? lo : hi;
        }
      }
    }
  }
}

__device__ long long g_filter_debug_clock[2];   // {SM cycles, ns} of CTA 0, written in the timing-experiment modes only

// kNKB: k-blocks of 64 per row (d_pad / 64).  kCluster: 1, or 2 = clusters of two CTAs that work on two different
// 256-user groups over the SAME item tiles: each CTA fetches half of every tile and TMA-multicasts it into both CTAs'
// shared memory, so the L2 -> SM stream of the item operand (1 TB per launch at 1M x 1M x d128, the second largest
// consumer after the MMAs) is halved.
// kTileEnd: compile the tile-end compaction pass in (see the epilogue).  It pays for short sweeps, where the admission
// path is a large share of the time, and costs at long ones -- mostly through what its second inlined copy of the
// compaction does to the hot loop's code, so it is a template parameter and not a run-time switch
// (profiles/probe_r2_v18_filter_ab_*.txt: 125K items 32.35 -> 31.7 ms with it, 1M items 205.6 -> 209.3 ms).
template <int kNKB, int kCluster, bool kTileEnd>
__global__ void __launch_bounds__(kFThreads, 1)
score_filter_kernel(const __grid_constant__ CUtensorMap map_items, const FilterParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const FilterLayout L = filter_layout(p.n_stages);
  const int n_slots = p.n_stages / kNKB;   // B tile slots
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* a_full = bars + 0;        // [user block]
  uint64_t* tmem_full = bars + 2;     // [accumulator slot]
  uint64_t* tmem_empty = bars + 5;
  uint64_t* b_full = bars + 8;
  uint64_t* b_empty = bars + 8 + n_slots;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(smem + L.bar_off + 400);
  constexpr uint32_t kSlotBytes = kNKB * kFBTileBytes;

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  constexpr int n_kb = kNKB;
  // work unit = (group of kCluster user pairs, item split); CTA `crank` of the cluster takes pair kCluster * g + crank
  const uint32_t crank = kCluster == 2 ? cluster_ctarank() : 0u;
  const int n_groups = (p.n_user_pairs + kCluster - 1) / kCluster;
  const int64_t n_work = static_cast<int64_t>(n_groups) * p.n_splits;
  const int64_t w_first = blockIdx.x / kCluster, w_step = gridDim.x / kCluster;
  constexpr uint16_t kClusterMask = (1u << kCluster) - 1u;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&map_items);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) mbar_init(a_full + i, 4);
    for (int i = 0; i < kFAccSlots; ++i) {
      mbar_init(tmem_full + i, 1);
      mbar_init(tmem_empty + i, 4);
    }
    for (int i = 0; i < n_slots; ++i) {
      mbar_init(b_full + i, 1);
      mbar_init(b_empty + i, kCluster);   // the MMA warps of all CTAs that received the tile
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kFTmemCols>(tmem_base_smem);
  tcgen05_fence_before();
  __syncthreads();
  if (kCluster == 2) cluster_sync_all();   // the peer's barriers exist before anything is multicast to them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  long long dbg_clk = 0, dbg_ns = 0;
  if (p.debug_mode != 0 && blockIdx.x == 0 && threadIdx.x == 64) {   // timing experiments: SM clock under this load
    dbg_clk = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_ns));
  }

  if (warp == 0) {
    // ===================================== TMA producer ======================================
    {   // warp-uniform control flow, one elected lane issues (see the MMA warp)
      int ts = 0;
      uint32_t ts_phase = 0, filled = 0;
      for (int64_t w = w_first; w < n_work; w += w_step) {
        const int sp = static_cast<int>(w / n_groups);
        const int t0 = sp * p.tiles_per_split;
        const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
        for (int t = t0; t < t1; ++t) {
          if (p.debug_mode == 6 && filled >= static_cast<uint32_t>(n_slots)) continue;   // timing: no B stream
          mbar_wait(b_empty + ts, ts_phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(b_full + ts, kSlotBytes);
#pragma unroll
            for (int kb = 0; kb < kNKB; ++kb) {
              if (kCluster == 2)   // this CTA's half of the tile rows (box = 64 rows), delivered to both CTAs
                tma_load_2d_multicast(smem + L.b_off + ts * kSlotBytes + kb * kFBTileBytes + crank * (kFBTileBytes / 2),
                                      &map_items, b_full + ts, kb * kFKBlock,
                                      t * kFBlockN + static_cast<int>(crank) * (kFBlockN / 2), kClusterMask, kEvictLast);
              else
                tma_load_2d(smem + L.b_off + ts * kSlotBytes + kb * kFBTileBytes, &map_items, b_full + ts,
                            kb * kFKBlock, t * kFBlockN, kEvictLast);
            }
          }
          __syncwarp();
          ++filled;
          if (++ts == n_slots) {
            ts = 0;
            ts_phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    // The whole warp runs the (warp-uniform) control flow and polls the barriers; one elected lane issues.  Issuing
    // from inside `if (lane == 0)` makes ptxas wrap every tcgen05.mma in an ELECT / R2UR.BROADCAST loop (~17
    // instructions per MMA) because it cannot prove the operands uniform.
    //
    // The user operand is read from TENSOR MEMORY (tcgen05.mma [d], [a], b-desc): with both operands in shared memory
    // an M = 128, N = 128, K = 16 step takes 77 cycles instead of the 64 of the math (scripts/mma_probe).
    //
    // This loop has to stay LEAN: one thread issues every MMA of the SM, and 8 MMA steps are only 512 cycles of tensor
    // work.  With a runtime stage count the ring arithmetic (fill % n_stages, fill / n_stages: ~25 dependent
    // instructions each through I2F / MUFU.RCP / F2I) plus the R2UR moves put ~450 cycles of scalar latency in front of
    // every 4 steps and the pipe ran at 112 cycles per step (TRK_FILTER_DEBUG=6).  Now: ring positions advance
    // incrementally, one full/empty barrier per item tile, all 8 steps of an accumulator issued from one elected
    // block with compile-time offsets.
    {
      constexpr uint32_t idesc = umma_idesc_f16_f32(kFBlockM, kFBlockN);
      int ts = 0;
      uint32_t ts_phase = 0, witer = 0, slot = 0, slot_phase = 0, consumed = 0;
      const uint32_t b_base = smem_u32(smem + L.b_off);
      for (int64_t w = w_first; w < n_work; w += w_step) {
        const int sp = static_cast<int>(w / n_groups);
        const int t0 = sp * p.tiles_per_split;
        const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
        if (t1 <= t0) continue;
        mbar_wait(a_full + 0, witer & 1);   // both groups have written their user block into tensor memory
        mbar_wait(a_full + 1, witer & 1);
        tcgen05_fence_after();
        ++witer;
        for (int t = t0; t < t1; ++t) {
          const bool streamed = !(p.debug_mode == 6 && consumed >= static_cast<uint32_t>(n_slots));
          if (streamed) mbar_wait(b_full + ts, ts_phase);
          const uint64_t db = umma_desc_k_major_sw128(b_base + ts * kSlotBytes);
#pragma unroll
          for (int b = 0; b < 2; ++b) {     // one B tile, two user blocks, one accumulator each
            mbar_wait(tmem_empty + slot, slot_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + kFTmemAccCol + slot * kFBlockN;
            const uint32_t a_tmem = tmem_base + b * 64;
            if (elect_one()) {
#pragma unroll
              for (int kb = 0; kb < kNKB; ++kb)
#pragma unroll
                for (int ks = 0; ks < kFKBlock / kFUmmaK; ++ks)
                  umma_f16_ts(d_tmem, a_tmem + kb * (kFKBlock / 2) + ks * (kFUmmaK / 2),
                              db + static_cast<uint64_t>(kb * (kFBTileBytes >> 4) + 2 * ks), idesc,
                              static_cast<uint32_t>(kb > 0 || ks > 0));
              umma_commit(tmem_full + slot);
              if (b == 1 && streamed) {
                if (kCluster == 2)
                  umma_commit_multicast(b_empty + ts, kClusterMask);
                else
                  umma_commit(b_empty + ts);
              }
            }
            __syncwarp();
            if (++slot == kFAccSlots) {
              slot = 0;
              slot_phase ^= 1;
            }
          }
          ++consumed;
          if (++ts == n_slots) {
            ts = 0;
            ts_phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue ==========================================
    const int group = (warp - 4) / 4;
    const int quarter = warp % 4;
    const int row = quarter * 32 + lane;
    const float kNegInf = -__int_as_float(0x7f800000);
    const uint32_t buf_row_addr =   // group g owns user block g of the pair
        smem_u32(smem + L.buf_off) + static_cast<uint32_t>((group * kFBlockM + row) * kBufEntries * 8);
    const float max_item_norm = __ldg(p.item_stats + 0);
    const float item_scale = fmaxf(__ldg(p.item_stats + 1), 1e-38f);
    const float max_item_bias = __ldg(p.item_stats + 2);
    const AdmitCtx ctx = {p.item_bias, p.item_perm, p.item_id_offset, static_cast<int32_t>(p.n_items)};
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t slot = group, slot_use = 0;   // accumulator number q = 2 (tile count) + group: slot q % 3, use q / 3

    for (int64_t w = w_first; w < n_work; w += w_step) {
      const int up = static_cast<int>(w % n_groups) * kCluster + static_cast<int>(crank);
      const int sp = static_cast<int>(w / n_groups);
      const int t0 = sp * p.tiles_per_split;
      const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
      const int64_t u = (static_cast<int64_t>(up) * 2 + group) * kFBlockM + row;
      const bool u_ok = u < p.n_users;
      const float su = u_ok ? __ldg(p.user_scale + u) : 1.0f;
      const float ubias = (u_ok && p.user_bias != nullptr) ? __ldg(p.user_bias + u) : 0.0f;
      const float unorm = u_ok ? __ldg(p.user_norm + u) : 0.0f;
      const float c = su * item_scale;          // powers of two: exact
      const float inv_c = 1.0f / c;
      // error bound of one approximate score: operand rounding + the fp32 rounding of the two bias adds
      const float m3 = kThetaMargins * (kMarginFactor * unorm * max_item_norm + kBiasUlps * (fabsf(ubias) + max_item_bias));
      float tau = p.debug_mode == 4 ? -kNegInf : kNegInf, theta = kNegInf;   // 4: timing experiment, nothing admitted
      int cnt = 0, n_res = 0, n_ovf = 0;
      float drop_max = kNegInf;
      uint32_t ra[32], rb[32];

      if (t1 > t0) {
        // This row of the user operand (hi half, fp16) goes to tensor memory: lane = row, two values per column,
        // k-block kb in columns [64 group + 32 kb, +32).  Every MMA that read the previous unit's block completed
        // before this warp saw the tmem_full of that unit's last accumulator, so the columns are free.
        const uint4* src = reinterpret_cast<const uint4*>(p.user_split + u * 2 * p.d_pad);
        for (int kb = 0; kb < n_kb; ++kb) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 v = u_ok ? __ldg(src + kb * 8 + i) : make_uint4(0u, 0u, 0u, 0u);
            ra[4 * i + 0] = v.x;
            ra[4 * i + 1] = v.y;
            ra[4 * i + 2] = v.z;
            ra[4 * i + 3] = v.w;
          }
          tmem_st_32x32b_x32(tmem_lane + group * 64 + kb * (kFKBlock / 2), ra);
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full + group);
      }

      float bmax_next = t1 > t0 ? __ldg(p.block_bias_max + t0) : 0.0f;
      for (int t = t0; t < t1; ++t) {
        const float bmax_scaled = bmax_next * inv_c;
        if (t + 1 < t1) bmax_next = __ldg(p.block_bias_max + t + 1);   // in flight while this tile is filtered
        mbar_wait(tmem_full + slot, slot_use & 1);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_lane + kFTmemAccCol + slot * kFBlockN;
        const int32_t pos0 = t * kFBlockN;
        if (t == t0 && p.block_bias_min != nullptr && p.debug_mode == 0) {
          const float bmin = __ldg(p.block_bias_min + t0);   // the same for the whole CTA: warp-uniform branch
          if (bmin > kNegInf) {
            float g[16];
            tmem_ld_32x32b_x32(taddr, ra);
            tmem_ld_wait();
            tmem_ld_32x32b_x32(taddr + 32, rb);
#pragma unroll
            for (int q = 0; q < 4; ++q) g[q] = acc_max_8(ra + 8 * q);
            tmem_ld_wait();
            tmem_ld_32x32b_x32(taddr + 64, ra);
#pragma unroll
            for (int q = 0; q < 4; ++q) g[4 + q] = acc_max_8(rb + 8 * q);
            tmem_ld_wait();
            tmem_ld_32x32b_x32(taddr + 96, rb);
#pragma unroll
            for (int q = 0; q < 4; ++q) g[8 + q] = acc_max_8(ra + 8 * q);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) g[12 + q] = acc_max_8(rb + 8 * q);
            sort16_desc(g);
            float a_k = g[0];
#pragma unroll
            for (int i = 1; i < 16; ++i) a_k = (i < p.k) ? g[i] : a_k;   // g[k - 1]: the k-th largest group maximum
            const float th0 = (fmaf(a_k, c, ubias) + bmin) - m3;
            if (th0 == th0) {   // not NaN (infinite biases / margins): otherwise the sweep starts from -inf as before
              theta = th0;
              const float tt = (theta - ubias) * inv_c;
              tau = tt - 8.0f * 1.1920929e-7f * fabsf(tt) - 1e-30f;
            }
          }
        }
        if (p.debug_mode == 2 || p.debug_mode == 6) goto drained;
        if (p.debug_mode == 1) {
          float acc_dbg = 0.0f;
          for (int ch = 0; ch < kFBlockN / 32; ch += 2) {
            tmem_ld_32x32b_x32(taddr + ch * 32, ra);
            tmem_ld_32x32b_x32(taddr + (ch + 1) * 32, rb);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              acc_dbg = fmaxf(acc_dbg, fmaxf(__uint_as_float(ra[j]), __uint_as_float(rb[j])));
          }
          if (acc_dbg == 1.2345e30f) cnt = 1;
          goto drained;
        }
        tmem_ld_32x32b_x32(taddr, ra);
        tmem_ld_wait();
#pragma unroll 1
        for (int ch = 0; ch < kFBlockN / 32; ch += 2) {   // (compile-time bounds: a run-time bound here cost 2.5 %)
          tmem_ld_32x32b_x32(taddr + (ch + 1) * 32, rb);   // in flight while chunk ch is filtered
          filter_32(ra, pos0 + ch * 32, bmax_scaled, ctx, c, inv_c, ubias, tau, theta, drop_max, n_ovf, m3, buf_row_addr,
                    cnt, n_res, lane, p.k);
          tmem_ld_wait();
          if (ch + 2 < kFBlockN / 32) tmem_ld_32x32b_x32(taddr + (ch + 2) * 32, ra);
          filter_32(rb, pos0 + (ch + 1) * 32, bmax_scaled, ctx, c, inv_c, ubias, tau, theta, drop_max, n_ovf, m3,
                    buf_row_addr, cnt, n_res, lane, p.k);
          tmem_ld_wait();
        }
      drained:
        // accumulator and bias slot drained
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty + slot);
        // A compaction waits one L2 round trip for the biases / item ids of its new entries (almost all of shared memory
        // is carved out: there is no L1 to speak of) -- ~0.8 us during which, in the middle of a tile, the warp's 32 rows
        // stand still AND their accumulator slot stays occupied: the admission path cost ~10 ms of a 31 ms sweep of a
        // 125K-item shard, the same at 1.45 and at 1.9 GHz (profiles/probe_r2_v9_filter_shard8_cool.txt).  So rows whose
        // buffer is filling up are compacted HERE, after the slot has gone back to the MMA warp: the round trip overlaps
        // the MMAs of this group's next accumulator (-1.5 ms of 32 at the shard with the trigger at 26 of 32 entries,
        // profiles/probe_r2_v10_filter_shard8_tile_end_trigger.txt).  The mid-tile path remains for a row that overflows
        // inside a tile.  Compiled in only for kTileEnd (short sweeps; see the template parameter).  (Measured and not kept: releasing the slot before the last chunk is filtered, and giving the
        // MMA / TMA warps the highest warp ids -- both neutral: the epilogue warps' own time per tile is the limit.)
        if (kTileEnd && p.debug_mode == 0) {
          const unsigned early = __ballot_sync(0xffffffffu, cnt > p.tile_end_trigger);
          compact_rows(early, buf_row_addr, lane, p.k, cnt, n_res, theta, tau, drop_max, n_ovf, m3, ubias, c, inv_c, ctx);
        }
        slot += 2;                          // q += 2
        if (slot >= kFAccSlots) {
          slot -= kFAccSlots;
          ++slot_use;
        }
      }

      // end of the item range: final compaction of every row of this warp, then emit the survivors
      compact_rows(0xffffffffu, buf_row_addr, lane, p.k, cnt, n_res, theta, tau, drop_max, n_ovf, m3, ubias, c, inv_c,
                   ctx);
      if (u_ok) {
        const int64_t base = u * p.n_splits + sp;
        float* os = p.cand_score + base * kKeepMax;
        int32_t* oi = p.cand_item + base * kKeepMax;
        for (int e = 0; e < kKeepMax; ++e) {
          float s = kNegInf;
          int32_t id = 0x7fffffff;
          if (e < cnt) f_lds64(buf_row_addr + e * 8, &s, &id);
          os[e] = s;
          oi[e] = id;
        }
        // every excluded item has an approximate score <= this; a NaN (inf - inf with infinite biases) must not read
        // as "nothing was excluded": +inf makes the certificate fail and the row goes through the exact kernel
        const bool th_nan = theta != theta || drop_max != drop_max;
        p.row_theta[base] = th_nan ? __int_as_float(0x7f800000) : fmaxf(theta, drop_max);
      }
      __syncwarp();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (kCluster == 2) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it
  if (p.debug_mode != 0 && blockIdx.x == 0 && threadIdx.x == 64) {
    long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    g_filter_debug_clock[0] = clock64() - dbg_clk;
    g_filter_debug_clock[1] = ns - dbg_ns;
  }
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<kFTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------
// operand preparation: row norms + global statistics, and the globally scaled "hi" item operand
// ---------------------------------------------------------------------------------------------------------
// stats[0] = max row norm, stats[1] = max row scale (2^-e of the largest row); both via atomicMax on the float bits
// (non-negative floats order like their bit patterns) -> order independent, deterministic.  stats must be zeroed.
__global__ void operand_stats_kernel(const __half* __restrict__ split, const float* __restrict__ scale, int64_t rows,
                                     int d_pad, float* __restrict__ out_norm, float* __restrict__ stats) {
  const int lane = threadIdx.x % 32;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / 32;
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * blockDim.x / 32;
  float local_max_norm = 0.0f, local_max_scale = 0.0f;
  for (int64_t r = warp; r < rows; r += n_warps) {
    const __half* hi = split + r * 2 * d_pad;
    const __half* lo = hi + d_pad;
    float ss = 0.0f;
    for (int e = lane * 2; e < d_pad; e += 64) {
      const float2 h = __half22float2(*reinterpret_cast<const __half2*>(hi + e));
      const float2 l = __half22float2(*reinterpret_cast<const __half2*>(lo + e));
      const float x0 = h.x + l.x, x1 = h.y + l.y;
      ss = fmaf(x0, x0, ss);
      ss = fmaf(x1, x1, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float sc = scale[r];
    // the norm is used as an UPPER bound: inflate by 2^-9 to cover the 22-bit operand and the reduction rounding
    const float norm = sqrtf(ss) * sc * 1.002f;
    if (lane == 0) {
      if (out_norm != nullptr) out_norm[r] = norm;
      local_max_norm = fmaxf(local_max_norm, norm);
      if (ss > 0.0f) local_max_scale = fmaxf(local_max_scale, sc);   // all-zero rows carry the neutral scale 1
    }
  }
  if (lane == 0 && stats != nullptr) {
    atomicMax(reinterpret_cast<int*>(stats + 0), __float_as_int(local_max_norm));
    atomicMax(reinterpret_cast<int*>(stats + 1), __float_as_int(local_max_scale));
  }
}

// hi_global[r, :] = hi[r, :] * (scale_r / max_scale): an exact power-of-two rescale (values of small rows may fall
// into the fp16 subnormal range -- that loss is inside the filter's error bound).
__global__ void rescale_hi_global_kernel(const __half* __restrict__ split, const float* __restrict__ scale,
                                         const float* __restrict__ stats, const int32_t* __restrict__ perm,
                                         int64_t rows, int d_pad, __half* __restrict__ out_hi) {
  const float inv_max = 1.0f / fmaxf(stats[1], 1e-38f);
  const int64_t n_vec = rows * (d_pad / 8);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / (d_pad / 8);                     // output row = processing position
    const int e = static_cast<int>(i % (d_pad / 8)) * 8;
    const int64_t src = perm != nullptr ? perm[r] : r;     // the item placed at that position
    const float f = scale[src] * inv_max;   // power of two <= 1
    uint4 raw = *reinterpret_cast<const uint4*>(split + src * 2 * d_pad + e);
    __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 v = __half22float2(h[q]);
      h[q] = __floats2half2_rn(v.x * f, v.y * f);
    }
    *reinterpret_cast<uint4*>(out_hi + r * d_pad + e) = raw;
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn filter_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// fp16 [rows, row_elems] row-major (only the first d_pad columns are addressed), boxes 64 x box_rows, 128B swizzle
int make_hi_map(CUtensorMap* map, const void* base, int64_t rows, int row_elems, int d_pad, int box_rows) {
  EncodeTiledFn encode = filter_encode_fn();
  if (encode == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return TRK_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(d_pad), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(row_elems) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kFKBlock), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                            elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return TRK_ERR_CUDA;
  }
  return TRK_OK;
}

constexpr uint32_t kFSmemLimit = 232448;

}  // namespace

int score_filter_max_k() { return kFilterMaxK; }
int score_filter_list_width() { return kKeepMax; }

int operand_stats(const void* split, const float* scale, int64_t rows, int32_t d_pad, float* out_norm, float* stats,
                  cudaStream_t stream) {
  TRK_CHECK_ARG(split && scale && rows >= 0 && d_pad >= 64 && d_pad % 64 == 0, "operand_stats: bad arguments");
  if (rows == 0) return TRK_OK;
  const int threads = 256;
  const int64_t blocks = ceil_div(rows, threads / 32);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  operand_stats_kernel<<<static_cast<unsigned>(blocks < cap ? blocks : cap), threads, 0, stream>>>(
      static_cast<const __half*>(split), scale, rows, d_pad, out_norm, stats);
  TRK_CHECK_LAUNCH();
  return TRK_OK;
}

int rescale_hi_global(const void* split, const float* scale, const float* stats, const int32_t* perm, int64_t rows,
                      int32_t d_pad, void* out_hi, cudaStream_t stream) {
  TRK_CHECK_ARG(split && scale && stats && out_hi && rows >= 0 && d_pad >= 64 && d_pad % 64 == 0,
                "rescale_hi_global: bad arguments");
  if (rows == 0) return TRK_OK;
  const int threads = 256;
  const int64_t blocks = ceil_div(rows * (d_pad / 8), threads);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  rescale_hi_global_kernel<<<static_cast<unsigned>(blocks < cap ? blocks : cap), threads, 0, stream>>>(
      static_cast<const __half*>(split), scale, stats, perm, rows, d_pad, static_cast<__half*>(out_hi));
  TRK_CHECK_LAUNCH();
  return TRK_OK;
}

int score_filter_f16(const void* user_split, const float* user_scale, const float* user_bias,
                     const float* user_norm, const void* item_hi, const float* item_stats, const float* item_bias,
                     const float* block_bias_max, const float* block_bias_min, const int32_t* item_perm,
                     int64_t n_users, int64_t n_items, int32_t d_pad, int32_t k, int32_t n_splits,
                     int32_t item_id_offset, float* cand_score, int32_t* cand_item, float* row_theta,
                     cudaStream_t stream) {
  TRK_CHECK_ARG(user_split && user_scale && user_norm && item_hi && item_stats && item_bias && block_bias_max,
                "score_filter: null input");
  TRK_CHECK_ARG(cand_score && cand_item && row_theta, "score_filter: null output");
  TRK_CHECK_ARG(n_users >= 1 && n_items >= 1 && n_splits >= 1, "score_filter: empty shape");
  TRK_CHECK_ARG(n_users < (1ll << 31) && n_items < (1ll << 31) - 512, "score_filter: shape exceeds int32 indexing");
  if (d_pad != 64 && d_pad != 128) {
    set_error("score_filter: d_pad=%d not supported (64 or 128)", d_pad);
    return TRK_ERR_UNSUPPORTED;
  }
  if (k < 1 || k > kFilterMaxK) {
    set_error("score_filter: k=%d outside [1, %d]", k, kFilterMaxK);
    return TRK_ERR_UNSUPPORTED;
  }
  TRK_CHECK_ARG(reinterpret_cast<uintptr_t>(user_split) % 16 == 0 && reinterpret_cast<uintptr_t>(item_hi) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(item_bias) % 16 == 0,
                "score_filter: operands must be 16-byte aligned");

  FilterParams p;
  p.user_split = static_cast<const __half*>(user_split);
  p.d_pad = d_pad;
  p.user_scale = user_scale;
  p.user_bias = user_bias;
  p.user_norm = user_norm;
  p.item_bias = item_bias;
  p.block_bias_max = block_bias_max;
  p.block_bias_min = getenv("TRK_FILTER_NO_WARMSTART") != nullptr ? nullptr : block_bias_min;
  p.item_perm = item_perm;
  p.item_stats = item_stats;
  p.n_users = n_users;
  p.n_items = n_items;
  p.n_kblocks = d_pad / kFKBlock;
  p.k = k;
  p.n_tiles = static_cast<int32_t>(ceil_div(n_items, kFBlockN));
  p.n_splits = n_splits;
  p.tiles_per_split = static_cast<int32_t>(ceil_div(p.n_tiles, n_splits));
  p.n_user_pairs = static_cast<int32_t>(ceil_div(n_users, 2 * kFBlockM));
  p.item_id_offset = item_id_offset;
  // Tile-end compaction for sweeps of up to kTileEndMaxTiles item tiles per split (measured: helps at 125K items,
  // costs at 1M; the crossover interpolates to ~380K).  TRK_FILTER_TILE_END_TRIGGER (probe knob): a value below
  // kBufEntries forces it on with that trigger, kBufEntries or more forces it off.
  bool tile_end = p.tiles_per_split <= kTileEndMaxTiles;
  p.tile_end_trigger = 26;
  {
    const char* env = getenv("TRK_FILTER_TILE_END_TRIGGER");
    if (env != nullptr) {
      tile_end = atoi(env) < kBufEntries;
      if (tile_end) p.tile_end_trigger = atoi(env);
    }
    if (p.tile_end_trigger < kKeepMax + 2) p.tile_end_trigger = kKeepMax + 2;   // (a compaction leaves up to kKeepMax)
  }
  p.cand_score = cand_score;
  p.cand_item = cand_item;
  p.row_theta = row_theta;
  {
    const char* dbg = getenv("TRK_FILTER_DEBUG");
    p.debug_mode = dbg != nullptr ? atoi(dbg) : 0;
  }
  CUtensorMap map_items;
  int rc;
  p.n_stages = 0;
  for (int s = kFMaxStages; s >= 2; --s)
    if (s % p.n_kblocks == 0 && filter_layout(s).total + 1024 <= kFSmemLimit) {
      p.n_stages = s;
      break;
    }
  TRK_CHECK_ARG(p.n_stages >= 2 * p.n_kblocks, "score_filter: shared memory budget exceeded");
  const uint32_t smem_bytes = filter_layout(p.n_stages).total + 1024;

  // Launch form: clusters of two CTAs sharing every item tile through TMA multicast (default when the device can keep
  // (almost) all SMs busy with 2-CTA clusters), else independent CTAs.  TRK_FILTER_CLUSTER=1|2 forces one.
  int cluster = 2;
  {
    const char* env = getenv("TRK_FILTER_CLUSTER");
    if (env != nullptr && (atoi(env) == 1 || atoi(env) == 2)) cluster = atoi(env);
  }
  auto kernel2 = tile_end ? (p.n_kblocks == 2 ? score_filter_kernel<2, 2, true> : score_filter_kernel<1, 2, true>)
                          : (p.n_kblocks == 2 ? score_filter_kernel<2, 2, false> : score_filter_kernel<1, 2, false>);
  auto kernel1 = tile_end ? (p.n_kblocks == 2 ? score_filter_kernel<2, 1, true> : score_filter_kernel<1, 1, true>)
                          : (p.n_kblocks == 2 ? score_filter_kernel<2, 1, false> : score_filter_kernel<1, 1, false>);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  int max_clusters = 0;
  if (cluster == 2) {
    TRK_CHECK_CUDA(cudaFuncSetAttribute(kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    cfg.gridDim = dim3(2);
    cfg.blockDim = dim3(kFThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // the answer depends on (device, kernel, shared memory) only: asked once per device and kernel variant
    static int cached_clusters[64][4];
    static bool cached_valid[64][4];
    int device = 0;
    TRK_CHECK_CUDA(cudaGetDevice(&device));
    const int variant = (p.n_kblocks == 2 ? 1 : 0) + (tile_end ? 2 : 0);
    if (device >= 0 && device < 64 && cached_valid[device][variant]) {
      max_clusters = cached_clusters[device][variant];
    } else {
      if (cudaOccupancyMaxActiveClusters(&max_clusters, kernel2, &cfg) != cudaSuccess) {
        (void)cudaGetLastError();
        max_clusters = 0;
      }
      if (device >= 0 && device < 64) {
        cached_clusters[device][variant] = max_clusters;
        cached_valid[device][variant] = true;
      }
    }
    const char* env = getenv("TRK_FILTER_CLUSTER");
    if (max_clusters * 2 < sm_count() - 8 && env == nullptr) cluster = 1;   // too many SMs would sit idle
    if (max_clusters < 1) cluster = 1;
  }
  rc = make_hi_map(&map_items, item_hi, n_items, d_pad, d_pad, kFBlockN / cluster);
  if (rc != TRK_OK) return rc;
  if (cluster == 2) {
    const int64_t n_work = ceil_div(static_cast<int64_t>(p.n_user_pairs), 2) * n_splits;
    const int n_clusters = static_cast<int>(n_work < max_clusters ? n_work : max_clusters);
    cfg.gridDim = dim3(static_cast<unsigned>(2 * n_clusters));
    TRK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel2, map_items, p));
  } else {
    TRK_CHECK_CUDA(cudaFuncSetAttribute(kernel1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const int64_t n_work = static_cast<int64_t>(p.n_user_pairs) * n_splits;
    const int grid = static_cast<int>(n_work < sm_count() ? n_work : sm_count());
    kernel1<<<grid, kFThreads, smem_bytes, stream>>>(map_items, p);
  }
  TRK_CHECK_LAUNCH();
  if (p.debug_mode != 0) {   // timing experiments only: report the SM clock CTA 0 saw (synchronises)
    long long clk[2] = {0, 0};
    TRK_CHECK_CUDA(cudaStreamSynchronize(stream));
    TRK_CHECK_CUDA(cudaMemcpyFromSymbol(clk, g_filter_debug_clock, sizeof(clk)));
    fprintf(stderr, "[trk] score_filter debug=%d: %lld cycles in %.3f ms -> %.0f MHz\n", p.debug_mode, clk[0],
            clk[1] * 1e-6, clk[1] > 0 ? clk[0] * 1e3 / static_cast<double>(clk[1]) : 0.0);
  }
  return TRK_OK;
}

}  // namespace trk
