// K2 + K3 fused on tcgen05 tensor cores: user x item scores and the per-user top-k (or the dense score matrix),
// without the [n_users, n_items] matrix ever reaching HBM in the top-k form.
//
// Reference chain replaced: tf.matmul(user_repr, item_repr, transpose_b=True) (tensorrec/prediction_graphs.py:49-50,
// and :64-65 -> recommendation_graphs.py:121 for cosine, operands pre-normalised by K1), bias_prediction_dense
// (tensorrec/recommendation_graphs.py:41) and rank_predictions (:73-82) restricted to rank <= k.
//
// Arithmetic: operands are the split-fp16 rows written by K1 (hi | lo, per-row power-of-two scale).  Per 64-wide
// k-block three tcgen05.mma groups accumulate hi.hi + lo.hi + hi.lo into ONE fp32 accumulator in tensor memory,
// i.e. an fp32-grade dot product (the dropped lo.lo term is < 2^-22 relative) that is exact for integer-valued
// representations -- which is what makes rank parity with the reference testable bit for bit.
//
// CTA = 384 threads, one CTA per SM, persistent over work items (user block of 128 rows, item split):
//   warp 0      TMA producer: the A block (all k-blocks, resident for the whole sweep) then a ring of B k-block
//               tiles [128 items x 64 fp16, 128B-swizzled] over the item range;
//   warp 1      MMA issuer (one elected thread): M=128, N=128, K=16 instructions, four 128-column accumulators in
//               TMEM; tcgen05.commit releases smem stages and publishes finished accumulators;
//   warp 2      TMEM allocator;
//   warps 4-7   epilogue group 0, warps 8-11 epilogue group 1: tiles alternate between the groups, each group owns two
//               accumulators and two {scale, bias} slots fed by TMA; one thread per user row: tcgen05.ld 32 columns at a time, score = acc * scale_u *
//               scale_i + bias_u + bias_i, compare against the row's current k-th best, rare insert into the row's
//               sorted list in shared memory.  Items are visited in ascending id order and the compare is strict,
//               so equal scores keep the lower item id first -- tf.nn.top_k's order.
#include "common.cuh"

namespace trk {

constexpr int kBlockM = 128;
constexpr int kBlockN = 128;          // item tile; 4 accumulators of 128 columns fill the 512 TMEM columns
constexpr int kKBlock = 64;           // fp16 per 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kTcThreads = 384;
constexpr int kEpiThreads = 128;      // per epilogue group
constexpr uint32_t kATileBytes = kBlockM * kKBlock * 2;   // 16 KB
constexpr uint32_t kBTileBytes = kBlockN * kKBlock * 2;   // 16 KB
constexpr uint32_t kMetaBytes = kBlockN * 8;              // {item scale, item bias} per column: 1 KB per tile
constexpr int kMaxStages = 10;
constexpr uint32_t kTmemCols = 512;
constexpr int kMaxK = 32;

struct TcParams {
  const float* user_scale;
  const float* user_bias;   // may be null
  const float2* item_meta;  // [tiles*256] {scale, bias}; padding {0, -inf}
  int64_t n_users;
  int64_t n_items;
  int32_t n_kblocks;        // d_pad / 64  (hi half); the operand has 2*n_kblocks k-blocks
  int32_t n_stages;
  int32_t k;                // top-k mode
  int32_t n_splits;
  int32_t tiles_per_split;
  int32_t n_tiles;          // ceil(n_items / 256)
  int32_t n_user_blocks;
  int32_t item_id_offset;
  float* cand_score;        // top-k mode outputs
  int32_t* cand_item;
  float* dense_out;         // dense mode output
  int64_t dense_stride;
  int32_t tma_store;        // dense mode: 1 = rows are 16-byte aligned, the epilogue stores through TMA
  const int32_t* n_users_live;  // device, may be null: only the first *n_users_live user rows hold work (rows flagged
                                // by the filter's certificate, counted on the device) -- later user blocks are skipped
};

// shared-memory carve-up (offsets from a 1024-byte aligned base)
struct SmemLayout {
  uint32_t a_off, b_off, list_score_off, list_item_off, meta_off, bar_off, total;
};
constexpr uint32_t kStoreTileBytes = 32 * 32 * 4;   // one warp's 32 rows x 32 columns of fp32 scores
__host__ __device__ inline SmemLayout make_layout(int n_kblocks, int n_stages, int k, bool dense_staging = false) {
  SmemLayout L;
  L.a_off = 0;
  L.b_off = L.a_off + 2u * n_kblocks * kATileBytes;
  L.list_score_off = L.b_off + static_cast<uint32_t>(n_stages) * kBTileBytes;
  L.list_item_off = L.list_score_off + (dense_staging ? 16u * kStoreTileBytes : 2u * k * kBlockM * 4u);
  L.meta_off = L.list_item_off + (dense_staging ? 0u : 2u * k * kBlockM * 4u);
  L.bar_off = L.meta_off + 4u * kMetaBytes;      // 2 groups x 2 slots, filled by the TMA warp
  L.total = L.bar_off + 512u;
  return L;
}

// barrier block (uint64 each): [0] a_full, [1] a_empty, [2..5] tmem_full, [6..9] tmem_empty, [10..13] meta_full,
// [14..17] meta_empty, [18 .. 18+S) b_full, [18+S .. 18+2S) b_empty; then the TMEM base address (uint32) at byte 400.
// Tile `it` is drained by epilogue group g = it & 1; use = it >> 1 counts that group's tiles and
// slot = g * 2 + (use & 1) names its accumulator and meta slot: two accumulators per group, so the MMA warp fills
// one while the group drains the other.

__device__ __noinline__ float list_insert(float s, int32_t id, float* ls, int32_t* li, int k) {
  // ls/li point at this row's column of the [k][128] arrays.  Entries are sorted by (score desc, id asc) and the
  // new id is larger than every id already stored, so it goes behind all entries with score >= s.
  // Returns the row's new k-th best score (the insertion threshold).
  int j = k - 1;
  while (j > 0 && ls[(j - 1) * kBlockM] < s) {
    ls[j * kBlockM] = ls[(j - 1) * kBlockM];
    li[j * kBlockM] = li[(j - 1) * kBlockM];
    --j;
  }
  ls[j * kBlockM] = s;
  li[j * kBlockM] = id;
  return ls[(k - 1) * kBlockM];
}

// 16-byte shared-memory load through the shared window (keeps the hot loop on LDS instead of generic LD)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// Scores one 32-column chunk held in r[] (raw accumulators in, final scores out) and returns the chunk maximum.
// Branch-free: the 16 LDS.128 and 96 FP ops of a chunk pipeline freely; the rare insert path is taken per chunk.
__device__ __forceinline__ float score_chunk(uint32_t (&r)[32], uint32_t meta_addr, float su, float ubias) {
  float cmax = -__int_as_float(0x7f800000);
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const float4 m = lds128(meta_addr + j * 8);   // {scale_j, bias_j, scale_j+1, bias_j+1}
    const float s0 = fmaf(__uint_as_float(r[j]), m.x * su, ubias) + m.y;       // (acc*scales + ub) + ib
    const float s1 = fmaf(__uint_as_float(r[j + 1]), m.z * su, ubias) + m.w;
    r[j] = __float_as_uint(s0);
    r[j + 1] = __float_as_uint(s1);
    cmax = fmaxf(cmax, fmaxf(s0, s1));
  }
  return cmax;
}

// One 32-column chunk of one user row: final scores, then (top-k mode) the rare inserts, or (dense mode) the store.
template <bool kDense>
__device__ __forceinline__ void process_chunk(uint32_t (&r)[32], int c, int t, int32_t id0, uint32_t meta_base,
                                              float su, float ubias, float& thr, float* ls, int32_t* li,
                                              const TcParams& p, int64_t u, bool u_ok) {
  const float cmax = score_chunk(r, meta_base + c * 32 * 8, su, ubias);
  if constexpr (!kDense) {
    if (cmax > thr) {   // some column of this chunk enters the row's list (probability ~ 32 k / items seen)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float s = __uint_as_float(r[j]);
        if (s > thr) thr = list_insert(s, id0 + c * 32 + j, ls, li, p.k);
      }
    }
  } else {
    const int64_t i0 = static_cast<int64_t>(t) * kBlockN + c * 32;
    if (u_ok) {
      float* dst = p.dense_out + u * p.dense_stride + i0;
      if (i0 + 32 <= p.n_items && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(dst + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (i0 + j < p.n_items) dst[j] = __uint_as_float(r[j]);
      }
    }
  }
}

// Dense mode, TMA path: the warp's 32 rows x 32 columns go to a 128B-swizzled staging tile in shared memory and
// leave as ONE cp.async.bulk.tensor store (full 128-byte lines per row, rows / columns beyond the matrix clipped by the
// tensor map).  Direct stores from the row-per-thread layout write 16 bytes per lane to 32 different rows.
__device__ __forceinline__ void store_chunk_tma(const uint32_t (&r)[32], uint32_t stage_addr, int lane,
                                                const CUtensorMap* map_out, int32_t col0, int32_t row0) {
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the buffer used two chunks ago
  __syncwarp();
#pragma unroll
  for (int c16 = 0; c16 < 8; ++c16) {
    const uint32_t addr = stage_addr + lane * 128 + ((c16 ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[4 * c16]), "r"(r[4 * c16 + 1]),
                 "r"(r[4 * c16 + 2]), "r"(r[4 * c16 + 3])
                 : "memory");
  }
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map_out)),
                 "r"(stage_addr), "r"(col0), "r"(row0)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

template <bool kDense, int kNKB>   // kNKB = d_pad / 64 k-blocks per operand half
__global__ void __launch_bounds__(kTcThreads, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap map_users, const __grid_constant__ CUtensorMap map_items,
                const __grid_constant__ CUtensorMap map_out, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need a 1024-byte aligned base
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const SmemLayout L = make_layout(p.n_kblocks, p.n_stages, kDense ? 0 : p.k, kDense && p.tma_store != 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* a_full = bars + 0;
  uint64_t* a_empty = bars + 1;
  uint64_t* tmem_full = bars + 2;
  uint64_t* tmem_empty = bars + 6;
  uint64_t* meta_full = bars + 10;
  uint64_t* meta_empty = bars + 14;
  uint64_t* b_full = bars + 18;
  uint64_t* b_empty = bars + 18 + p.n_stages;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(smem + L.bar_off + 400);

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  constexpr int n_kb2 = 2 * kNKB;
  // work item w = (user block w / n_splits, item split w % n_splits): the splits of ONE user block go to consecutive
  // CTAs, so a handful of live user blocks (the device-side fallback: ~100 rows of a million) still spreads over the
  // whole machine -- with the block index minor, one live block of 8 x 148 items landed on 37 of the 148 CTAs
  const int64_t n_work = static_cast<int64_t>(p.n_user_blocks) * p.n_splits;
  // user blocks at or beyond this one hold no rows (every role skips them: the same test in all three loops)
  const int live_blocks = p.n_users_live != nullptr
                              ? static_cast<int>(min(static_cast<int64_t>(p.n_user_blocks),
                                                     ceil_div(static_cast<int64_t>(__ldg(p.n_users_live)), kBlockM)))
                              : p.n_user_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_users);
    tma_prefetch_desc(&map_items);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(tmem_full + i, 1);
      mbar_init(tmem_empty + i, kEpiThreads / 32);
      mbar_init(meta_full + i, 1);
      mbar_init(meta_empty + i, kEpiThreads / 32);
    }
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(b_full + i, 1);
      mbar_init(b_empty + i, 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_base_smem);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0) {
    // ===================================== TMA producer ======================================
    {   // warp-uniform control flow, one elected lane issues (see the MMA warp)
      int stage = 0;       // B ring position
      uint32_t stage_phase = 0;
      uint32_t witer = 0;  // non-empty work items so far
      uint32_t it = 0;     // tiles issued so far
      for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int ub = static_cast<int>(w / p.n_splits);
        const int sp = static_cast<int>(w % p.n_splits);
        const int t0 = sp * p.tiles_per_split;
        const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
        if (t1 <= t0 || ub >= live_blocks) continue;
        mbar_wait(a_empty, (witer & 1) ^ 1);  // the MMAs of the previous work item no longer read A
        if (elect_one()) {
          mbar_arrive_expect_tx(a_full, n_kb2 * kATileBytes);
          for (int kb = 0; kb < n_kb2; ++kb)
            tma_load_2d(smem + L.a_off + kb * kATileBytes, &map_users, a_full, kb * kKBlock, ub * kBlockM,
                        kEvictFirst);
        }
        __syncwarp();
        ++witer;
        for (int t = t0; t < t1; ++t, ++it) {
          // {item scale, item bias} of this tile for the epilogue group that will drain it
          const uint32_t use = it >> 1, slot = (it & 1) * 2 + (use & 1);
          mbar_wait(meta_empty + slot, ((use >> 1) & 1) ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(meta_full + slot, kMetaBytes);
            bulk_load_1d(smem + L.meta_off + slot * kMetaBytes, p.item_meta + static_cast<int64_t>(t) * kBlockN,
                         kMetaBytes, meta_full + slot);
          }
          __syncwarp();
#pragma unroll
          for (int kb = 0; kb < n_kb2; ++kb) {
            mbar_wait(b_empty + stage, stage_phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(b_full + stage, kBTileBytes);
              tma_load_2d(smem + L.b_off + stage * kBTileBytes, &map_items, b_full + stage, kb * kKBlock, t * kBlockN,
                          kEvictLast);  // the item operand is re-read by every user block: keep it in L2
            }
            __syncwarp();
            if (++stage == p.n_stages) {   // ring position advances incrementally: no div/mod on the issue path
              stage = 0;
              stage_phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    // The whole warp runs the (warp-uniform) control flow and polls the barriers; one elected lane issues.  Issuing
    // from inside `if (lane == 0)` makes ptxas wrap every tcgen05.mma in an ELECT / R2UR.BROADCAST loop (~17
    // instructions per MMA) because it cannot prove the operands uniform.
    {
      constexpr uint32_t idesc = umma_idesc_f16_f32(kBlockM, kBlockN);
      int stage = 0;
      uint32_t stage_phase = 0, witer = 0, it = 0;  // it = accumulator tiles produced so far
      const uint32_t a_base = smem_u32(smem + L.a_off);
      const uint32_t b_base = smem_u32(smem + L.b_off);
      for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int sp = static_cast<int>(w % p.n_splits);
        const int t0 = sp * p.tiles_per_split;
        const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
        if (t1 <= t0 || static_cast<int>(w / p.n_splits) >= live_blocks) continue;
        mbar_wait(a_full, witer & 1);
        ++witer;
        for (int t = t0; t < t1; ++t, ++it) {
          const uint32_t use = it >> 1, buf = (it & 1) * 2 + (use & 1);
          mbar_wait(tmem_empty + buf, ((use >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + buf * kBlockN;
          // One thread issues every MMA of the SM and a k-block is only 4-8 MMA steps: the loop is unrolled over the
          // (compile-time) k-blocks and the ring position advances incrementally -- `fill % n_stages` with a runtime
          // stage count cost ~25 dependent instructions (I2F / MUFU.RCP / F2I) in front of every k-block and kept the
          // tensor pipe at half rate (the same finding as in score_filter_tc.cu).
#pragma unroll
          for (int kb2 = 0; kb2 < n_kb2; ++kb2) {
            mbar_wait(b_full + stage, stage_phase);
            tcgen05_fence_after();
            const uint64_t db = umma_desc_k_major_sw128(b_base + stage * kBTileBytes);
            constexpr int kNKBc = kNKB;
            const bool b_is_hi = kb2 < kNKBc;
            const int kb = b_is_hi ? kb2 : kb2 - kNKBc;
            // B hi block: A_hi[kb] x B and A_lo[kb] x B ;  B lo block: A_hi[kb] x B
            if (elect_one()) {
#pragma unroll
              for (int a = 0; a < (b_is_hi ? 2 : 1); ++a) {
                const uint64_t da = umma_desc_k_major_sw128(a_base + (a == 0 ? kb : kNKBc + kb) * kATileBytes);
#pragma unroll
                for (int ks = 0; ks < kKBlock / kUmmaK; ++ks) {
                  // advancing 16 fp16 (32 bytes) inside the 128-byte swizzle atom = +2 in the address field
                  umma_f16_ss(d_tmem, da + 2u * ks, db + 2u * ks, idesc,
                              static_cast<uint32_t>(kb2 > 0 || a > 0 || ks > 0));
                }
              }
              umma_commit(b_empty + stage);  // stage reusable once these MMAs have read it
            }
            __syncwarp();
            if (++stage == p.n_stages) {
              stage = 0;
              stage_phase ^= 1;
            }
          }
          if (elect_one()) umma_commit(tmem_full + buf);  // accumulator complete
          __syncwarp();
        }
        if (elect_one()) umma_commit(a_empty);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue ==========================================
    const int group = (warp - 4) / 4;               // 0 or 1
    const int quarter = warp % 4;                   // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;            // row inside the user block == TMEM lane
    float* ls = reinterpret_cast<float*>(smem + L.list_score_off) + group * (kDense ? 0 : p.k) * kBlockM + row;
    int32_t* li = reinterpret_cast<int32_t*>(smem + L.list_item_off) + group * (kDense ? 0 : p.k) * kBlockM + row;
    const float kNegInf = -__int_as_float(0x7f800000);
    uint32_t it = 0;
    uint32_t n_stored = 0;   // dense TMA path: chunks stored by this warp (selects the staging buffer)
    const uint32_t stage_base = smem_u32(smem + L.list_score_off) + static_cast<uint32_t>(warp - 4) * 2u * kStoreTileBytes;

    for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int ub = static_cast<int>(w / p.n_splits);
      const int sp = static_cast<int>(w % p.n_splits);
      const int t0 = sp * p.tiles_per_split;
      const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
      if (ub >= live_blocks) continue;
      const int64_t u = static_cast<int64_t>(ub) * kBlockM + row;
      const bool u_ok = u < p.n_users;
      const float su = u_ok ? __ldg(p.user_scale + u) : 0.0f;
      const float ubias = (u_ok && p.user_bias != nullptr) ? __ldg(p.user_bias + u) : 0.0f;
      float thr = kNegInf;
      if constexpr (!kDense) {
        for (int j = 0; j < p.k; ++j) {
          ls[j * kBlockM] = kNegInf;
          li[j * kBlockM] = 0x7fffffff;
        }
      }

      for (int t = t0; t < t1; ++t, ++it) {
        if (static_cast<int>(it & 1) != group) continue;
        const uint32_t use = it >> 1, slot = group * 2 + (use & 1);
        mbar_wait(meta_full + slot, (use >> 1) & 1);     // this tile's {item scale, item bias}, put there by TMA
        mbar_wait(tmem_full + slot, (use >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + slot * kBlockN;
        const int32_t id0 = p.item_id_offset + t * kBlockN;
        const uint32_t meta_base = smem_u32(smem + L.meta_off) + slot * kMetaBytes;
        uint32_t ra[32], rb[32];
        tmem_ld_32x32b_x32(taddr, ra);
        tmem_ld_wait();
#pragma unroll 1
        for (int c = 0; c < kBlockN / 32; c += 2) {
          tmem_ld_32x32b_x32(taddr + (c + 1) * 32, rb);   // in flight while chunk c is scored
          if (kDense && p.tma_store) {
            score_chunk(ra, meta_base + c * 32 * 8, su, ubias);
            store_chunk_tma(ra, stage_base + (n_stored & 1) * kStoreTileBytes, lane, &map_out, t * kBlockN + c * 32,
                            ub * kBlockM + quarter * 32);
            ++n_stored;
          } else {
            process_chunk<kDense>(ra, c, t, id0, meta_base, su, ubias, thr, ls, li, p, u, u_ok);
          }
          tmem_ld_wait();
          if (c + 2 < kBlockN / 32) tmem_ld_32x32b_x32(taddr + (c + 2) * 32, ra);
          if (kDense && p.tma_store) {
            score_chunk(rb, meta_base + (c + 1) * 32 * 8, su, ubias);
            store_chunk_tma(rb, stage_base + (n_stored & 1) * kStoreTileBytes, lane, &map_out,
                            t * kBlockN + (c + 1) * 32, ub * kBlockM + quarter * 32);
            ++n_stored;
          } else {
            process_chunk<kDense>(rb, c + 1, t, id0, meta_base, su, ubias, thr, ls, li, p, u, u_ok);
          }
          tmem_ld_wait();
        }
        // accumulator drained: hand the TMEM buffer back to the MMA warp
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(tmem_empty + slot);
          mbar_arrive(meta_empty + slot);
        }
      }

      if constexpr (!kDense) {
        // both groups have finished the item range: merge the two lists of each row and emit the candidates
        named_barrier_sync(1, 2 * kEpiThreads);
        if (group == 0 && u_ok) {
          const float* l0s = ls;
          const int32_t* l0i = li;
          const float* l1s = ls + p.k * kBlockM;
          const int32_t* l1i = li + p.k * kBlockM;
          float* os = p.cand_score + (u * p.n_splits + sp) * p.k;
          int32_t* oi = p.cand_item + (u * p.n_splits + sp) * p.k;
          int a = 0, b = 0;
          for (int j = 0; j < p.k; ++j) {
            const float sa = l0s[a * kBlockM], sb = l1s[b * kBlockM];
            const int32_t ia = l0i[a * kBlockM], ib = l1i[b * kBlockM];
            const bool take_a = sa > sb || (sa == sb && ia <= ib);
            os[j] = take_a ? sa : sb;
            oi[j] = take_a ? ia : ib;
            a += take_a ? 1 : 0;
            b += take_a ? 0 : 1;
          }
        }
        named_barrier_sync(1, 2 * kEpiThreads);   // lists may be re-initialised for the next work item
      }
    }
  }

  // ---- teardown ----
  if (kDense && p.tma_store && warp >= 4 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

// [rows, 2*d_pad] fp16 row-major, boxes of 64 columns x box_rows rows, 128-byte swizzle
int make_operand_map(CUtensorMap* map, const void* base, int64_t rows, int d_pad, int box_rows) {
  EncodeTiledFn encode = get_encode_fn();
  if (encode == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return TRK_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(2 * d_pad), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(2 * d_pad) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kKBlock), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                            elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld d_pad=%d)", static_cast<int>(r),
              static_cast<long long>(rows), d_pad);
    return TRK_ERR_CUDA;
  }
  return TRK_OK;
}

constexpr uint32_t kSmemLimit = 232448;  // 227 KB opt-in limit per CTA on sm_100

int pick_stages(int n_kblocks, int k, bool dense_staging) {
  for (int s = kMaxStages; s >= 2; --s)
    if (make_layout(n_kblocks, s, k, dense_staging).total + 1024 <= kSmemLimit) return s;
  return 0;
}

}  // namespace

int score_topk_max_k(int32_t d_pad) {
  if (d_pad != 64 && d_pad != 128) return 0;
  return kMaxK;
}

template <bool kDense>
static int launch_tc(const void* user_split, const float* user_scale, const float* user_bias,
                     const void* item_split, const float* item_meta, int64_t n_users, int64_t n_items,
                     int32_t d_pad, int32_t k, int32_t n_splits, int32_t item_id_offset, float* cand_score,
                     int32_t* cand_item, float* dense_out, int64_t dense_stride, const int32_t* n_users_live,
                     cudaStream_t stream) {
  TRK_CHECK_ARG(user_split && user_scale && item_split && item_meta, "score_tc: null operand");
  TRK_CHECK_ARG(n_users >= 1 && n_items >= 1, "score_tc: empty shape");
  TRK_CHECK_ARG(n_users < (1ll << 31) && n_items < (1ll << 31) - 512, "score_tc: shape exceeds int32 indexing");
  if (d_pad != 64 && d_pad != 128) {
    set_error("score_tc: d_pad=%d not supported by the tensor-core kernel (64 or 128)", d_pad);
    return TRK_ERR_UNSUPPORTED;
  }
  TRK_CHECK_ARG(reinterpret_cast<uintptr_t>(user_split) % 16 == 0 && reinterpret_cast<uintptr_t>(item_split) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(item_meta) % 16 == 0,
                "score_tc: operands must be 16-byte aligned");
  if (!kDense) {
    TRK_CHECK_ARG(cand_score && cand_item, "score_topk: null output");
    if (k < 1 || k > kMaxK) {
      set_error("score_topk: k=%d outside [1, %d]", k, kMaxK);
      return TRK_ERR_UNSUPPORTED;
    }
  } else {
    TRK_CHECK_ARG(dense_out && dense_stride >= n_items, "score_dense: bad output");
  }
  TRK_CHECK_ARG(n_splits >= 1, "score_tc: n_splits < 1");

  TcParams p;
  p.user_scale = user_scale;
  p.user_bias = user_bias;
  p.item_meta = reinterpret_cast<const float2*>(item_meta);
  p.n_users = n_users;
  p.n_items = n_items;
  p.n_kblocks = d_pad / kKBlock;
  p.k = kDense ? 0 : k;
  p.n_tiles = static_cast<int32_t>(ceil_div(n_items, kBlockN));
  p.n_splits = n_splits;
  p.tiles_per_split = static_cast<int32_t>(ceil_div(p.n_tiles, n_splits));
  p.n_user_blocks = static_cast<int32_t>(ceil_div(n_users, kBlockM));
  p.item_id_offset = item_id_offset;
  p.cand_score = cand_score;
  p.cand_item = cand_item;
  p.dense_out = dense_out;
  p.dense_stride = dense_stride;
  p.n_users_live = n_users_live;
  p.tma_store = (kDense && dense_stride % 4 == 0 && reinterpret_cast<uintptr_t>(dense_out) % 16 == 0) ? 1 : 0;
  p.n_stages = pick_stages(p.n_kblocks, p.k, p.tma_store != 0);
  TRK_CHECK_ARG(p.n_stages >= 2, "score_tc: shared memory budget exceeded (d_pad=%d k=%d)", d_pad, k);

  CUtensorMap map_users, map_items;
  int rc = make_operand_map(&map_users, user_split, n_users, d_pad, kBlockM);
  if (rc != TRK_OK) return rc;
  rc = make_operand_map(&map_items, item_split, n_items, d_pad, kBlockN);
  if (rc != TRK_OK) return rc;

  CUtensorMap map_out = map_items;   // placeholder when the TMA store path is off
  if (p.tma_store) {
    EncodeTiledFn encode = get_encode_fn();
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(n_items), static_cast<cuuint64_t>(n_users)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(dense_stride) * 4};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t elem_strides[2] = {1, 1};
    const CUresult r = encode(&map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dense_out, dims, strides, box,
                              elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (output) failed with CUresult %d", static_cast<int>(r));
      return TRK_ERR_CUDA;
    }
  }
  const uint32_t smem_bytes = make_layout(p.n_kblocks, p.n_stages, p.k, p.tma_store != 0).total + 1024;
  auto kernel = p.n_kblocks == 2 ? score_tc_kernel<kDense, 2> : score_tc_kernel<kDense, 1>;
  TRK_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int64_t n_work = static_cast<int64_t>(p.n_user_blocks) * n_splits;
  const int grid = static_cast<int>(n_work < sm_count() ? n_work : sm_count());
  kernel<<<grid, kTcThreads, smem_bytes, stream>>>(map_users, map_items, map_out, p);
  TRK_CHECK_LAUNCH();
  return TRK_OK;
}

int score_topk_f16x3(const void* user_split, const float* user_scale, const float* user_bias,
                     const void* item_split, const float* item_meta, int64_t n_users, int64_t n_items,
                     int32_t d_pad, int32_t k, int32_t n_splits, int32_t item_id_offset, float* cand_score,
                     int32_t* cand_item, const int32_t* n_users_live, cudaStream_t stream) {
  return launch_tc<false>(user_split, user_scale, user_bias, item_split, item_meta, n_users, n_items, d_pad, k,
                          n_splits, item_id_offset, cand_score, cand_item, nullptr, 0, n_users_live, stream);
}

int score_dense_f16x3(const void* user_split, const float* user_scale, const float* user_bias,
                      const void* item_split, const float* item_meta, int64_t n_users, int64_t n_items,
                      int32_t d_pad, float* out, int64_t out_row_stride, cudaStream_t stream) {
  // split the item axis so that every SM gets work even when there are few user blocks
  const int64_t n_ub = ceil_div(n_users, kBlockM);
  const int64_t n_tiles = ceil_div(n_items, kBlockN);
  int64_t splits = ceil_div(2 * static_cast<int64_t>(sm_count()), n_ub);
  if (splits > n_tiles) splits = n_tiles;
  if (splits < 1) splits = 1;
  return launch_tc<true>(user_split, user_scale, user_bias, item_split, item_meta, n_users, n_items, d_pad, 0,
                         static_cast<int32_t>(splits), 0, nullptr, nullptr, out, out_row_stride, nullptr, stream);
}

}  // namespace trk
