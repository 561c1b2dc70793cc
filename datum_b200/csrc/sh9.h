// datum_b200 — launch interface of the SH9 kernels (internal to libdatum_ibl_cuda).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

namespace ibl
{
  struct Sh9Coefficients { float v[27]; };

  // where the 28 sums of a slab also go when several GPUs share a cube: slot [rank] of every peer's
  // [world][28] array (NVLink peer stores from the block that finishes last)
  struct Sh9Peers
  {
    double *slots[7];
    unsigned int *arrive[7];   // arrival counter in each peer's flag block, bumped once the row is stored (null: no signal)
    int count;
  };

  // fp64 solid-angle table of data/project.comp:56-60, w*h floats
  cudaError_t launch_sh9_weights(float *weights, int w, int h, cudaStream_t stream);

  int sh9_partial_blocks(int w, int h, int sm_count);

  // block_partials: blocks*28 doubles of scratch; done_counter: one zero-initialised word the kernel
  // leaves at zero; partial: 28 doubles (27 sums + weight sum).  One launch.
  // kernel: 0 = column strips (default), 1 = row segments (A/B); rows_per_item: 0 = automatic.
  // probes > 1: that many cubes probe_stride BYTES apart in one launch; block_partials then holds
  // probes x blocks x 28 doubles, done_counter `probes` words, partial probes x 28 doubles.
  cudaError_t launch_sh9_partial(void const *level0, int format, float const *weights, int w, int h, int row_begin, int row_end, double *block_partials, int blocks, unsigned int *done_counter, double *partial, Sh9Peers const &peers, int sm_count, cudaStream_t stream, int kernel = 0, int rows_per_item = 0, int probes = 1, size_t probe_stride = 0);

  cudaError_t launch_sh9_irradiance(Sh9Coefficients const &sh, int w, int h, uint32_t *words, float *f32, int sm_count, cudaStream_t stream);
}
