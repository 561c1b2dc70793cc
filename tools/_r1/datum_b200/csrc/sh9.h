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
    int count;
  }; // [k][rgb], the layout of Irradiance::L (src/renderer/envmap.h:112-115)

  // fp64 solid-angle table of data/project.comp:56-60, w*h floats
  cudaError_t launch_sh9_weights(float *weights, int w, int h, cudaStream_t stream);

  int sh9_partial_blocks(int w, int h, int sm_count);

  // block_partials: blocks*28 doubles of scratch; done_counter: one zero-initialised word the kernel
  // leaves at zero; partial: 28 doubles (27 sums + weight sum).  One launch.
  cudaError_t launch_sh9_partial(void const *level0, int format, float const *weights, int w, int h, int row_begin, int row_end, double *block_partials, int blocks, unsigned int *done_counter, double *partial, Sh9Peers const &peers, int sm_count, cudaStream_t stream);

  cudaError_t launch_sh9_irradiance(Sh9Coefficients const &sh, int w, int h, uint32_t *words, float *f32, int sm_count, cudaStream_t stream);
}
