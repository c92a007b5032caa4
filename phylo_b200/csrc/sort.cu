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

}  // namespace

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
