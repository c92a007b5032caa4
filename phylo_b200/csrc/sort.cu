// Visiting order of the particles of one rank event: sorted by the (unordered) pair of child nodes, inactive
// particles last.  Keys are built by a small kernel; the sort itself is cub::DeviceRadixSort (CCCL, ships with the
// CUDA toolkit) on K (key, index) pairs -- bookkeeping next to the K*S merges, not a hot op.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "launch.h"

namespace vcsmc {
namespace {

__global__ void __launch_bounds__(256) build_keys_kernel(const int32_t* __restrict__ lsrc, const int32_t* __restrict__ rsrc,
                                                         const int32_t* __restrict__ active, int64_t K, int bits,
                                                         uint64_t* __restrict__ keys, int32_t* __restrict__ vals,
                                                         int32_t* __restrict__ count) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int on = 0;
  if (k < K) {
    on = active ? (active[k] != 0) : 1;
    const int ls = lsrc[k], rs = rsrc[k];
    // child refs are >= -256 (leaves) and < 2^31 - 256: shift to unsigned; `bits` bits per child
    const uint64_t a = (uint64_t)(min(ls, rs) + 256), b = (uint64_t)(max(ls, rs) + 256);
    keys[k] = on ? ((a << bits) | b) : ((1ull << (2 * bits)) - 1);
    vals[k] = (int32_t)k;
  }
  const int n = __syncthreads_count(on);
  if (threadIdx.x == 0 && n) atomicAdd(count, n);
}

// ---- grouping without a sort.  The merge kernels only need particles with the same (unordered) child pair to be
// ADJACENT; which group comes first is irrelevant (every particle's result is independent of its neighbours).  A hash
// table keyed by the pair hands out group slots, a warp-aggregated counter hands out ranks inside a group, one scan
// turns the counts into offsets: 4 short launches instead of the ~13 of a 40-bit radix sort.
__global__ void __launch_bounds__(256) group_insert_kernel(const int32_t* __restrict__ lsrc, const int32_t* __restrict__ rsrc,
                                                           const int32_t* __restrict__ active, int skip_leaf_pairs, int64_t K,
                                                           int log2T, unsigned long long* __restrict__ tab,
                                                           int32_t* __restrict__ cnt, int32_t* __restrict__ gslot,
                                                           int32_t* __restrict__ grank) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool on = k < K && (active ? active[k] != 0 : true);
  if (on && skip_leaf_pairs && lsrc[k] < 0 && rsrc[k] < 0) on = false;   // two leaves: scored from site patterns, not listed
  int slot = -1 - lane;  // inactive lanes never match anybody
  if (on) {
    const int ls = lsrc[k], rs = rsrc[k];
    const unsigned long long key = (((unsigned long long)(unsigned)(min(ls, rs) + 256)) << 32 | (unsigned)(max(ls, rs) + 256)) + 1ull;
    const unsigned mask = (1u << log2T) - 1u;
    unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> (64 - log2T));
    while (true) {
      const unsigned long long old = atomicCAS(tab + h, 0ull, key);
      if (old == 0ull || old == key) break;
      h = (h + 1) & mask;
    }
    slot = (int)h;
  }
  const unsigned peers = __match_any_sync(0xffffffffu, slot);
  if (on) {
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(cnt + slot, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    gslot[k] = slot;
    grank[k] = base + __popc(peers & ((1u << lane) - 1));
  } else if (k < K) {
    gslot[k] = -1;
  }
}

// group order is irrelevant, so offsets need no scan: every occupied table slot reserves its range with one atomic
__global__ void __launch_bounds__(256) group_offsets_kernel(int64_t T, const int32_t* __restrict__ cnt, int32_t* __restrict__ off,
                                                            int32_t* __restrict__ cursor) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  const int c = cnt[i];
  if (c > 0) off[i] = atomicAdd(cursor, c);
}

__global__ void __launch_bounds__(256) group_scatter_kernel(int64_t K, int64_t T, const int32_t* __restrict__ cnt,
                                                            const int32_t* __restrict__ off, const int32_t* __restrict__ gslot,
                                                            const int32_t* __restrict__ grank, int32_t* __restrict__ order,
                                                            int32_t* __restrict__ count_out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int s = gslot[k];
  if (s >= 0) order[off[s] + grank[k]] = (int32_t)k;
}

}  // namespace

int64_t group_table_entries(int64_t K) {
  int64_t T = 1024;
  while (T < 2 * K) T <<= 1;
  return T;
}

int launch_group_order(const int32_t* lsrc, const int32_t* rsrc, const int32_t* active, int skip_leaf_pairs, int64_t K, unsigned long long* tab,
                       int32_t* cnt, int32_t* off, int32_t* gslot, int32_t* grank, int32_t* order_out, int32_t* count_out,
                       void* temp, size_t temp_bytes, cudaStream_t st) {
  if (K <= 0) return VCSMC_OK;
  const int64_t T = group_table_entries(K);
  int log2T = 0;
  while (((int64_t)1 << log2T) < T) ++log2T;
  // (T depends on the K of THIS call -- a rank's share under particle sharding -- not on how the buffers were carved)
  VCSMC_CUDA(cudaMemsetAsync(tab, 0, (size_t)T * sizeof(unsigned long long), st));
  VCSMC_CUDA(cudaMemsetAsync(cnt, 0, (size_t)T * sizeof(int32_t), st));
  VCSMC_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int32_t), st));   // doubles as the range cursor
  count_launch(3);
  group_insert_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(lsrc, rsrc, active, skip_leaf_pairs, K, log2T, tab, cnt, gslot, grank);
  VCSMC_LAUNCH_CHECK("group_insert_kernel");
  group_offsets_kernel<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(T, cnt, off, count_out);
  VCSMC_LAUNCH_CHECK("group_offsets_kernel");
  group_scatter_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(K, T, cnt, off, gslot, grank, order_out, count_out);
  VCSMC_LAUNCH_CHECK("group_scatter_kernel");
  return VCSMC_OK;
}

size_t sort_temp_bytes(int64_t K) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)K, 0, 64, (cudaStream_t)0);
  return bytes;
}

size_t scan_temp_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, (cudaStream_t)0);
  return bytes;
}

int launch_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* temp, size_t temp_bytes, cudaStream_t st) {
  if (n <= 0) return VCSMC_OK;
  VCSMC_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, out, (int)n, st));
  count_launch(2);
  return VCSMC_OK;
}

int launch_sort_order(const int32_t* lsrc, const int32_t* rsrc, const int32_t* active, int64_t K, int64_t max_slot, uint64_t* keys_in,
                      uint64_t* keys_out, int32_t* vals_in, int32_t* order_out, int32_t* count_out, void* temp,
                      size_t temp_bytes, cudaStream_t st) {
  if (K <= 0) return VCSMC_OK;
  VCSMC_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int32_t), st));
  int bits = 9;  // (max_slot + 256 + 1) must fit; the all-ones key is reserved for inactive particles
  while (bits < 32 && ((int64_t)1 << bits) <= max_slot + 257) ++bits;
  build_keys_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(lsrc, rsrc, active, K, bits, keys_in, vals_in, count_out);
  VCSMC_LAUNCH_CHECK("build_keys_kernel");
  VCSMC_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, order_out, (int)K, 0, 2 * bits, st));
  count_launch(8);
  return VCSMC_OK;
}

}  // namespace vcsmc
