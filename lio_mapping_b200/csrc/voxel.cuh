// Device VoxelGrid workspace (see voxel.cu).
#pragma once
#include "primitives.cuh"

namespace lio {

struct VoxelGrid {
  int cap = 0, nstatus = 0;
  unsigned *keys_a = nullptr, *vals_a = nullptr, *keys_b = nullptr, *vals_b = nullptr;
  RadixSortTemp rs;
  unsigned long long *status = nullptr;
  unsigned *bbox = nullptr;
  int *ticket = nullptr;  // [0] tile ticket, [1] sticky PCL index-overflow flag, [2] clamped input count
  int init(int cap);
  void destroy();
  // in: min(*n_dev, n_max) points (n_max <= cap).  out: centroids in ascending voxel-index order, at most out_cap of
  // them are stored; *nout_dev always receives the full count (callers compare it with out_cap after their next sync).
  // vox_key_out (optional): the voxel index of every output centroid.
  int run(const float4 *in, const int *n_dev, int n_max, float leaf, float4 *out, int out_cap, int *nout_dev,
          unsigned *vox_key_out, cudaStream_t st, int *launches);
  int *overflow_flag() const { return ticket + 1; }
};

}  // namespace lio
