// Stable LSD radix sort of (u32 key, u32 value) pairs, 8 bits per pass, hand-written for sm_100a.
// Used by the device VoxelGrid (key = PCL voxel index, value = input index: stability makes the
// in-voxel order the input order, which fixes the fp32 centroid summation order).
#include "primitives.cuh"

namespace lio {

__global__ void __launch_bounds__(kRsThreads)
rs_hist(const unsigned *__restrict__ keys, const int *__restrict__ n_dev, int shift, int *__restrict__ tile_hist) {
  __shared__ int sh[kRsBins];
  const int n = *n_dev;
  const int base = blockIdx.x * kRsTile;
  if (base >= n) return;
  sh[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kRsTile / kRsThreads; ++k) {
    int i = base + k * kRsThreads + threadIdx.x;
    if (i < n) atomicAdd(&sh[(keys[i] >> shift) & (kRsBins - 1)], 1);
  }
  __syncthreads();
  tile_hist[blockIdx.x * kRsBins + threadIdx.x] = sh[threadIdx.x];
}

// tile_hist[t][b] -> global exclusive offset of (bin b, tile t) in bin-major order.
__global__ void __launch_bounds__(kRsBins)
rs_scan(int *__restrict__ tile_hist, const int *__restrict__ n_dev) {
  __shared__ int sscan[40];
  const int n = *n_dev;
  const int ntiles = (n + kRsTile - 1) / kRsTile;
  const int b = threadIdx.x;
  int run = 0;
  for (int t = 0; t < ntiles; ++t) {
    int h = tile_hist[t * kRsBins + b];
    tile_hist[t * kRsBins + b] = run;
    run += h;
  }
  int tot;
  int bin_off = block_scan_excl(run, sscan, &tot);
  for (int t = 0; t < ntiles; ++t) tile_hist[t * kRsBins + b] += bin_off;
}

__global__ void __launch_bounds__(kRsThreads)
rs_scatter(const unsigned *__restrict__ keys, const unsigned *__restrict__ vals, unsigned *__restrict__ keys_out,
           unsigned *__restrict__ vals_out, const int *__restrict__ n_dev, int shift, const int *__restrict__ tile_off) {
  __shared__ int running[kRsBins];
  __shared__ int warpcnt[kRsThreads / 32][kRsBins];
  const int n = *n_dev;
  const int base = blockIdx.x * kRsTile;
  if (base >= n) return;
  running[threadIdx.x] = tile_off[blockIdx.x * kRsBins + threadIdx.x];
#pragma unroll
  for (int w = 0; w < kRsThreads / 32; ++w) warpcnt[w][threadIdx.x] = 0;
  __syncthreads();
  const int w = warp_id();
  for (int k = 0; k < kRsTile / kRsThreads; ++k) {
    int i = base + k * kRsThreads + threadIdx.x;
    unsigned key = 0, val = 0;
    int bin = -1;
    if (i < n) { key = keys[i]; val = vals[i]; bin = (int)((key >> shift) & (kRsBins - 1)); }
    unsigned peers = __match_any_sync(0xffffffffu, bin);
    int lrank = __popc(peers & ((1u << lane_id()) - 1u));
    if (bin >= 0 && lrank == 0) warpcnt[w][bin] = __popc(peers);
    __syncthreads();
    if (bin >= 0) {
      int pos = running[bin] + lrank;
      for (int ww = 0; ww < w; ++ww) pos += warpcnt[ww][bin];
      keys_out[pos] = key;
      vals_out[pos] = val;
    }
    __syncthreads();
    {
      int s = 0;
#pragma unroll
      for (int ww = 0; ww < kRsThreads / 32; ++ww) { s += warpcnt[ww][threadIdx.x]; warpcnt[ww][threadIdx.x] = 0; }
      running[threadIdx.x] += s;
    }
    __syncthreads();
  }
}

int radix_sort_pairs(unsigned *keys_a, unsigned *vals_a, unsigned *keys_b, unsigned *vals_b, const int *n_dev, int n_max,
                     int key_bits, RadixSortTemp &tmp, cudaStream_t st, int *launches) {
  int ntiles = (n_max + kRsTile - 1) / kRsTile;
  if (ntiles < 1) ntiles = 1;
  int passes = (key_bits + kRsBits - 1) / kRsBits;
  unsigned *ki = keys_a, *vi = vals_a, *ko = keys_b, *vo = vals_b;
  for (int p = 0; p < passes; ++p) {
    int shift = p * kRsBits;
    rs_hist<<<ntiles, kRsThreads, 0, st>>>(ki, n_dev, shift, tmp.tile_hist);
    rs_scan<<<1, kRsBins, 0, st>>>(tmp.tile_hist, n_dev);
    rs_scatter<<<ntiles, kRsThreads, 0, st>>>(ki, vi, ko, vo, n_dev, shift, tmp.tile_hist);
    if (launches) *launches += 3;
    unsigned *t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
  }
  return passes & 1;
}

}  // namespace lio
