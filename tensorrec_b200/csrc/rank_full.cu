// K3 (full) -- rank_predictions (tensorrec/recommendation_graphs.py:73-82).
//
// The reference ranks by a double full sort: order = top_k(pred, k=n).indices (descending, equal values by
// lower index first), ranks = top_k(-order, k=n).indices + 1 (the inverse permutation).  In closed form
//     rank[u,i] = 1 + #{j : s[u,j] > s[u,i]} + #{j < i : s[u,j] == s[u,i]}                    (int32)
// Each score becomes the 64-bit key (descending-order bits of the float) << 32 | index; keys of a row are unique,
// ascending key order IS the reference's order, so no stable sort is needed:
//   pass 1: every chunk of kChunk keys of a row is bitonic-sorted in shared memory;
//   pass 2 (rows longer than one chunk): the rank of a key is its position in its own chunk plus, for every other
//           chunk of the row, the number of keys below it (binary search in the chunk staged to shared memory).
// Integer compares only -> exact and run-to-run deterministic.
#include "common.cuh"

namespace trk {

constexpr int kChunk = 4096;          // keys per sorted chunk: 32 KB of shared memory
constexpr int kSortThreads = 1024;
constexpr uint64_t kPadKey = ~0ull;   // sorts behind every real key

__device__ __forceinline__ uint64_t rank_key(float s, uint32_t idx) {
  s = s + 0.0f;  // -0.0 -> +0.0: the reference compares values, where the two are equal
  const uint32_t b = __float_as_uint(s);
  const uint32_t asc = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // monotone float -> uint
  return (static_cast<uint64_t>(~asc) << 32) | idx;                  // descending score, ascending index
}

// In-place ascending bitonic sort of `n` (power of two, <= kChunk) keys in shared memory by the whole block.
__device__ __forceinline__ void bitonic_sort_smem(uint64_t* keys, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int hi = lo | j;
        const uint64_t a = keys[lo], b = keys[hi];
        const bool ascending = (lo & k) == 0;
        if ((a > b) == ascending) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

// grid = (chunks per row, rows).  Single-chunk rows are ranked directly.
__global__ void __launch_bounds__(kSortThreads)
rank_chunk_sort_kernel(const float* __restrict__ scores, int32_t* __restrict__ ranks, uint64_t* __restrict__ sorted,
                       int64_t n_items, int n_chunks, int sort_n) {
  extern __shared__ uint64_t s_keys[];
  const int64_t row = blockIdx.y;
  const int chunk = blockIdx.x;
  const int64_t base = static_cast<int64_t>(chunk) * kChunk;
  const float* srow = scores + row * n_items;
  for (int t = threadIdx.x; t < sort_n; t += blockDim.x) {
    const int64_t i = base + t;
    s_keys[t] = i < n_items ? rank_key(__ldg(srow + i), static_cast<uint32_t>(i)) : kPadKey;
  }
  bitonic_sort_smem(s_keys, sort_n);
  if (n_chunks == 1) {
    int32_t* rrow = ranks + row * n_items;
    for (int t = threadIdx.x; t < sort_n; t += blockDim.x) {
      const uint64_t key = s_keys[t];
      if (key != kPadKey) rrow[static_cast<uint32_t>(key)] = t + 1;
    }
  } else {
    uint64_t* dst = sorted + (row * n_chunks + chunk) * kChunk;
    for (int t = threadIdx.x; t < kChunk; t += blockDim.x) dst[t] = s_keys[t];
  }
}

__global__ void __launch_bounds__(kSortThreads)
rank_merge_count_kernel(const uint64_t* __restrict__ sorted, int32_t* __restrict__ ranks, int64_t n_items,
                        int n_chunks) {
  __shared__ uint64_t s_other[kChunk];
  const int64_t row = blockIdx.y;
  const int chunk = blockIdx.x;
  const uint64_t* row_sorted = sorted + row * n_chunks * kChunk;
  constexpr int kPer = kChunk / kSortThreads;
  uint64_t mine[kPer];
  int count[kPer];
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    const int pos = threadIdx.x + q * kSortThreads;
    mine[q] = row_sorted[static_cast<int64_t>(chunk) * kChunk + pos];
    count[q] = pos;  // keys of the own chunk below this one
  }
  for (int other = 0; other < n_chunks; ++other) {
    if (other == chunk) continue;
    __syncthreads();
    for (int t = threadIdx.x; t < kChunk; t += kSortThreads)
      s_other[t] = row_sorted[static_cast<int64_t>(other) * kChunk + t];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      // number of keys in s_other below mine[q] (keys are unique): branch-free lower bound over 4096 = 2^12
      int lo = 0;
#pragma unroll
      for (int step = kChunk >> 1; step > 0; step >>= 1)
        if (s_other[lo + step - 1] < mine[q]) lo += step;
      if (s_other[lo] < mine[q]) lo += 1;  // covers the last slot
      count[q] += lo;
    }
  }
  int32_t* rrow = ranks + row * n_items;
#pragma unroll
  for (int q = 0; q < kPer; ++q)
    if (mine[q] != kPadKey) rrow[static_cast<uint32_t>(mine[q])] = count[q] + 1;
}

static int next_pow2(int64_t n) {
  int p = 2;
  while (p < n) p <<= 1;
  return p;
}

size_t rank_full_workspace_bytes(int64_t n_users, int64_t n_items) {
  if (n_items <= kChunk || n_users <= 0) return 0;
  const int64_t n_chunks = ceil_div(n_items, kChunk);
  return static_cast<size_t>(n_users) * static_cast<size_t>(n_chunks) * kChunk * sizeof(uint64_t);
}

int rank_full(const float* scores, int32_t* ranks, int64_t n_users, int64_t n_items, void* workspace,
              size_t workspace_bytes, cudaStream_t stream) {
  TRK_CHECK_ARG(scores && ranks, "rank_full: null pointer");
  TRK_CHECK_ARG(n_users >= 0 && n_items >= 0, "rank_full: negative size");
  TRK_CHECK_ARG(n_items < (1ll << 31), "rank_full: n_items must fit int32 (the reference returns int32 ranks)");
  if (n_users == 0 || n_items == 0) return TRK_OK;
  const int64_t n_chunks = ceil_div(n_items, kChunk);
  TRK_CHECK_ARG(n_chunks <= 65535 && n_users <= 2147483647ll, "rank_full: shape exceeds one launch");
  const size_t need = rank_full_workspace_bytes(n_users, n_items);
  TRK_CHECK_ARG(workspace_bytes >= need && (need == 0 || workspace != nullptr),
                "rank_full: workspace too small (%zu < %zu)", workspace_bytes, need);
  const int sort_n = n_chunks == 1 ? next_pow2(n_items) : kChunk;
  const int threads = sort_n / 2 < kSortThreads ? (sort_n / 2 < 32 ? 32 : sort_n / 2) : kSortThreads;
  // rows go on grid.y (<= 65535 per launch)
  for (int64_t r0 = 0; r0 < n_users; r0 += 65535) {
    const int64_t nr = n_users - r0 < 65535 ? n_users - r0 : 65535;
    const dim3 grid(static_cast<unsigned>(n_chunks), static_cast<unsigned>(nr));
    uint64_t* ws = static_cast<uint64_t*>(workspace);
    rank_chunk_sort_kernel<<<grid, threads, sort_n * sizeof(uint64_t), stream>>>(
        scores + r0 * n_items, ranks + r0 * n_items, ws ? ws + r0 * n_chunks * kChunk : nullptr, n_items,
        static_cast<int>(n_chunks), sort_n);
    TRK_CHECK_LAUNCH();
    if (n_chunks > 1) {
      rank_merge_count_kernel<<<grid, kSortThreads, 0, stream>>>(ws + r0 * n_chunks * kChunk, ranks + r0 * n_items,
                                                                 n_items, static_cast<int>(n_chunks));
      TRK_CHECK_LAUNCH();
    }
  }
  return TRK_OK;
}

}  // namespace trk
