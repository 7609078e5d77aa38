// Exact re-scoring and ranking of the survivors of the tensor-core filter (score_filter_tc.cu).
//
// For every user: each surviving item is scored exactly as the reference does -- fp32 dot product of the fp32
// representations (tf.matmul, tensorrec/prediction_graphs.py:49-50), then + user bias, + item bias left to right
// (tensorrec/recommendation_graphs.py:41) -- and the k best are selected in tf.nn.top_k order (score descending,
// equal scores by lower item id, recommendation_graphs.py:81).  One warp per user: lanes split the components of
// the dot product (fixed xor-tree reduction -> deterministic), the running top-k lives one entry per lane.
//
// Verification: the filter excluded only items whose approximate score was <= theta (per list), so their exact score
// is <= theta + m.  If max theta + m is not strictly below the exact k-th best found here -- or a buffer overflowed --
// the row is flagged and the host re-runs it through the exact 3-pass kernel.
#include "common.cuh"

namespace trk {

constexpr int kRescoreMaxChunks = 4;   // n_components <= 128
constexpr float kRMarginFactor = 1.5f * 0.0009765625f;
constexpr float kRBiasUlps = 4.0f * 1.1920929e-7f;

__device__ __forceinline__ bool r_before(float xs, int32_t xi, float ys, int32_t yi) {
  return xs > ys || (xs == ys && xi < yi);
}

__global__ void __launch_bounds__(256)
rescore_topk_kernel(const float* __restrict__ user_repr, const float* __restrict__ item_repr,
                    const float* __restrict__ user_bias, const float* __restrict__ item_bias,
                    const int32_t* __restrict__ cand_item, const float* __restrict__ row_theta,
                    const int32_t* __restrict__ row_flags, const float* __restrict__ user_norm,
                    const float* __restrict__ item_stats, int64_t n_users, int64_t n_items_local, int d, int n_lists,
                    int list_width, int k, int item_id_offset, float* __restrict__ out_score,
                    int32_t* __restrict__ out_item, int32_t* __restrict__ out_flag) {
  const int lane = threadIdx.x % 32;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / 32;
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * blockDim.x / 32;
  const float kNegInf = -__int_as_float(0x7f800000);
  const float max_item_norm = __ldg(item_stats + 0);
  const float max_item_bias = __ldg(item_stats + 2);
  const int n_cand = n_lists * list_width;

  for (int64_t u = warp; u < n_users; u += n_warps) {
    float uv[kRescoreMaxChunks];
#pragma unroll
    for (int j = 0; j < kRescoreMaxChunks; ++j) {
      const int e = lane + 32 * j;
      uv[j] = e < d ? __ldg(user_repr + u * d + e) : 0.0f;
    }
    const float ub = user_bias != nullptr ? __ldg(user_bias + u) : 0.0f;
    const int32_t* ci = cand_item + u * n_cand;
    // running top-k: lane j holds the j-th best so far
    float ls = kNegInf;
    int32_t li = 0x7fffffff;
    int n_real = 0;

    for (int c0 = 0; c0 < n_cand; c0 += 4) {
      int32_t ids[4];
      float iv[4][kRescoreMaxChunks];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        ids[q] = (c0 + q < n_cand) ? __ldg(ci + c0 + q) : 0x7fffffff;
        const int64_t local = static_cast<int64_t>(ids[q]) - item_id_offset;
        const bool ok = ids[q] != 0x7fffffff && local >= 0 && local < n_items_local;
        if (!ok) ids[q] = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < kRescoreMaxChunks; ++j) {
          const int e = lane + 32 * j;
          iv[q][j] = (ok && e < d) ? __ldg(item_repr + local * d + e) : 0.0f;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (ids[q] == 0x7fffffff) continue;   // warp-uniform: ids are broadcast loads
        float part = 0.0f;
#pragma unroll
        for (int j = 0; j < kRescoreMaxChunks; ++j) part = fmaf(uv[j], iv[q][j], part);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        const int64_t local = static_cast<int64_t>(ids[q]) - item_id_offset;
        float s = part;
        if (user_bias != nullptr) s = s + ub;
        if (item_bias != nullptr) s = s + __ldg(item_bias + local);
        // insert (s, id) into the lane-distributed sorted list
        const unsigned before = __ballot_sync(0xffffffffu, r_before(ls, li, s, ids[q]));
        const int pos = __popc(before);
        const float up_s = __shfl_up_sync(0xffffffffu, ls, 1);
        const int32_t up_i = __shfl_up_sync(0xffffffffu, li, 1);
        if (lane == pos) {
          ls = s;
          li = ids[q];
        } else if (lane > pos) {
          ls = up_s;
          li = up_i;
        }
        n_real += 1;
      }
    }

    // verification of the filter's bound for this user
    float theta_max = kNegInf;
    int flagged = 0;
    for (int l = lane; l < n_lists; l += 32) {
      theta_max = fmaxf(theta_max, __ldg(row_theta + u * n_lists + l));
      flagged |= __ldg(row_flags + u * n_lists + l);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      theta_max = fmaxf(theta_max, __shfl_xor_sync(0xffffffffu, theta_max, o));
      flagged |= __shfl_xor_sync(0xffffffffu, flagged, o);
    }
    const float kth = __shfl_sync(0xffffffffu, ls, k - 1);
    const float m = kRMarginFactor * __ldg(user_norm + u) * max_item_norm + kRBiasUlps * (fabsf(ub) + max_item_bias);
    bool valid = flagged == 0;
    if (theta_max > kNegInf) valid = valid && (n_real >= k) && (theta_max + m < kth);
    if (lane < k) {
      out_score[u * k + lane] = ls;
      out_item[u * k + lane] = li;
    }
    if (lane == 0) out_flag[u] = valid ? 0 : 1;
  }
}

// item biases in processing order (out[p] = bias[perm[p]]), padded to whole tiles with -inf (a padded column can
// never pass the filter); max |bias| -> stats[2]; block_max[b] = max bias of positions [128 b, 128 b + 128).
__global__ void pack_item_bias_kernel(const float* __restrict__ bias, const int32_t* __restrict__ perm, int64_t n,
                                      float* __restrict__ out, int64_t n_padded, float* __restrict__ stats,
                                      float* __restrict__ block_max) {
  // one warp per block of 128 positions
  const int lane = threadIdx.x % 32;
  const int64_t blk = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / 32;
  if (blk * 128 >= n_padded) return;
  const float kNegInf = -__int_as_float(0x7f800000);
  float vmax = kNegInf, amax = 0.0f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t pos = blk * 128 + q * 32 + lane;
    float v = kNegInf;
    if (pos < n) {
      const int64_t src = perm != nullptr ? perm[pos] : pos;
      v = bias != nullptr ? bias[src] : 0.0f;
      amax = fmaxf(amax, fabsf(v));
    }
    if (pos < n_padded) out[pos] = v;
    vmax = fmaxf(vmax, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  }
  if (lane == 0) {
    if (block_max != nullptr) block_max[blk] = vmax;
    if (stats != nullptr && amax > 0.0f) atomicMax(reinterpret_cast<int*>(stats + 2), __float_as_int(amax));
  }
}

int pack_item_bias(const float* bias, const int32_t* perm, int64_t n, float* out, int64_t n_padded, float* stats,
                   float* block_max, cudaStream_t stream) {
  TRK_CHECK_ARG(out && n >= 0 && n_padded >= n && n_padded % 128 == 0, "pack_item_bias: bad arguments");
  if (n_padded == 0) return TRK_OK;
  const int threads = 256;
  const int64_t n_blocks128 = n_padded / 128;
  pack_item_bias_kernel<<<static_cast<unsigned>(ceil_div(n_blocks128, threads / 32)), threads, 0, stream>>>(
      bias, perm, n, out, n_padded, stats, block_max);
  TRK_CHECK_LAUNCH();
  return TRK_OK;
}

int rescore_topk(const float* user_repr, const float* item_repr, const float* user_bias, const float* item_bias,
                 const int32_t* cand_item, const float* row_theta, const int32_t* row_flags, const float* user_norm,
                 const float* item_stats, int64_t n_users, int64_t n_items_local, int32_t d, int32_t n_lists,
                 int32_t list_width, int32_t k, int32_t item_id_offset, float* out_score, int32_t* out_item,
                 int32_t* out_flag, cudaStream_t stream) {
  TRK_CHECK_ARG(user_repr && item_repr && cand_item && row_theta && row_flags && user_norm && item_stats,
                "rescore_topk: null input");
  TRK_CHECK_ARG(out_score && out_item && out_flag, "rescore_topk: null output");
  TRK_CHECK_ARG(n_users >= 0 && n_items_local >= 0 && n_lists >= 1 && list_width >= 1, "rescore_topk: bad sizes");
  TRK_CHECK_ARG(d >= 1 && d <= 32 * kRescoreMaxChunks, "rescore_topk: n_components=%d outside [1, %d]", d,
                32 * kRescoreMaxChunks);
  TRK_CHECK_ARG(k >= 1 && k <= 32, "rescore_topk: k=%d outside [1, 32]", k);
  if (n_users == 0) return TRK_OK;
  const int threads = 256;
  const int64_t blocks = ceil_div(n_users, threads / 32);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  rescore_topk_kernel<<<static_cast<unsigned>(blocks < cap ? blocks : cap), threads, 0, stream>>>(
      user_repr, item_repr, user_bias, item_bias, cand_item, row_theta, row_flags, user_norm, item_stats, n_users,
      n_items_local, d, n_lists, list_width, k, item_id_offset, out_score, out_item, out_flag);
  TRK_CHECK_LAUNCH();
  return TRK_OK;
}

}  // namespace trk
