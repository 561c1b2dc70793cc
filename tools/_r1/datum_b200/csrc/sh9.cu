// datum_b200 — SH9 irradiance projection of a cube map and its evaluation (sm_100a).
//
// Replaces the single-thread GLSL loop of data/project.comp:23-106 (reference
// paths relative to /root/reference).  The texel solid angle (project.comp:56-60)
// depends only on (x, y), not on the face, and its four-atan form cancels
// catastrophically in fp32 once faces exceed a few hundred texels, so it is
// evaluated ONCE per (x, y) in fp64 into a table that the context caches per
// face size.  The projection itself streams the slab once (16 B/texel RGBA32F):
// a CTA takes 1024-texel segments of rows, every thread issues its four 16-byte
// loads before any arithmetic, accumulates 27 sums + the weight sum in registers,
// then warp-shuffle and block-reduce in fp64.  Block partials are summed in block
// order by whichever block finishes last, so the result is run-to-run deterministic.

#include "sh9.h"
#include "ibl_math.cuh"

#include <cuda_runtime.h>
#include <cstdint>

namespace ibl
{
  // ---- solid angle table -------------------------------------------------------

  __device__ __forceinline__ double corner_angle(double x, double y) { return atan2(x * y, sqrt(x * x + y * y + 1.0)); }

  __global__ void __launch_bounds__(256) sh9_weights_kernel(float *__restrict__ weights, int w, int h)
  {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= w * h)
      return;

    int x = idx % w, y = idx / w;

    // project.comp:53, 56-60
    double u = 2 * (x + 0.5) / w - 1;
    double v = 2 * (y + 0.5) / h - 1;
    double x0 = u - 1.0 / w, x1 = u + 1.0 / w;
    double y0 = v - 1.0 / h, y1 = v + 1.0 / h;

    weights[idx] = (float)(corner_angle(x0, y0) - corner_angle(x0, y1) - corner_angle(x1, y0) + corner_angle(x1, y1));
  }

  // ---- projection ----------------------------------------------------------------

  // L2 residency: the texel stream is read once (evict first, do not allocate in L1), the solid-angle
  // table is read once per face and should survive the stream in between (evict last)
  __device__ __forceinline__ unsigned long long l2_policy_evict_first()
  {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
  }

  __device__ __forceinline__ unsigned long long l2_policy_evict_last()
  {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
  }

  __device__ __forceinline__ float4 ldg_stream(float4 const *ptr, unsigned long long policy)
  {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(policy));
    return v;
  }

  __device__ __forceinline__ float ldg_keep(float const *ptr, unsigned long long policy)
  {
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(ptr), "l"(policy));
    return v;
  }

  template<int FORMAT>
  __device__ __forceinline__ void load_texel(void const *__restrict__ level0, size_t idx, unsigned long long stream_policy, float &r, float &g, float &b)
  {
    if (FORMAT == 0)
    {
      // color.h:164-172 with the 1/511 folded into the scale (<= 1 ulp from the reference decode)
      uint32_t c = __ldg(reinterpret_cast<uint32_t const *>(level0) + idx);
      float s = u2f((((c >> 27) & 0x1Fu) + 112u) << 23) * (1.0f / 511.0f);
      r = (float)((c >> 0) & 0x1FFu) * s;
      g = (float)((c >> 9) & 0x1FFu) * s;
      b = (float)((c >> 18) & 0x1FFu) * s;
    }
    else
    {
      float4 c = ldg_stream(reinterpret_cast<float4 const *>(level0) + idx, stream_policy);
      r = c.x; g = c.y; b = c.z;
    }
  }

  // acc[3*k + c] += weight * color[c] * M_k(ray) with the MONOMIALS
  //     M = 1, y, z, x, xy, yz, z^2, zx, x^2 - y^2
  // of the basis functions of project.comp:64-92; their constant factors (and the "3 z^2 - 1" of
  // Y6) are applied once per block in sh9_basis_from_monomials: 11 multiplies less per texel.
  __device__ __forceinline__ void sh9_accumulate(float acc[28], float wr, float wg, float wb, float rx, float ry, float rz)
  {
    float mono[9];
    mono[0] = 1.0f;
    mono[1] = ry;
    mono[2] = rz;
    mono[3] = rx;
    mono[4] = rx * ry;
    mono[5] = ry * rz;
    mono[6] = rz * rz;
    mono[7] = rz * rx;
    mono[8] = fmaf(rx, rx, -ry * ry);

    acc[0] += wr; acc[1] += wg; acc[2] += wb;

    #pragma unroll
    for(int k = 1; k < 9; ++k)
    {
      acc[3*k + 0] = fmaf(wr, mono[k], acc[3*k + 0]);
      acc[3*k + 1] = fmaf(wg, mono[k], acc[3*k + 1]);
      acc[3*k + 2] = fmaf(wb, mono[k], acc[3*k + 2]);
    }
  }

  // monomial sums -> basis sums (fp64, once per block): project.comp:64-92's constants
  __device__ __forceinline__ double sh9_basis_from_monomials(int k, double const mono[28])
  {
    int band = k / 3, c = k - 3 * band;
    switch (band)
    {
      case 0: return 0.282095 * mono[k];
      case 1: case 2: case 3: return 0.488603 * mono[k];
      case 4: case 5: case 7: return 1.092548 * mono[k];
      case 6: return 0.315392 * (3.0 * mono[k] - mono[c]);
      case 8: return 0.546274 * mono[k];
      default: return mono[k];     // k == 27: the weight sum
    }
  }

  constexpr int kSh9Threads = 256;
  constexpr int kSh9Unroll = 4;                               // texels per thread and work item: four 16-byte loads in flight
  constexpr int kSh9Segment = kSh9Threads * kSh9Unroll;       // texels of one row a CTA takes at a time

  // face rays as in data/convolve.comp:85-100 (equal to project.comp:27-32's quaternions);
  // (a, b, c) = (u, v, 1) / |(u, v, 1)|
  template<int FACE>
  __device__ __forceinline__ void face_ray(float a, float b, float c, float &rx, float &ry, float &rz)
  {
    switch (FACE)
    {
      case 0: rx = c;  ry = b;  rz = a;  break;
      case 1: rx = -c; ry = b;  rz = -a; break;
      case 2: rx = a;  ry = -c; rz = -b; break;
      case 3: rx = a;  ry = c;  rz = b;  break;
      case 4: rx = a;  ry = b;  rz = -c; break;
      default: rx = -a; ry = b; rz = c;  break;
    }
  }

  // One segment of one row of one face: the CTA's threads take texels x0 + j*256 + tid.  All loads of
  // the segment are issued before the arithmetic of the first texel: memory-level parallelism is what
  // keeps this kernel near the HBM roofline (16 B per texel against ~64 instructions).
  template<int FORMAT, int FACE>
  __device__ __forceinline__ void sh9_row_segment(void const *__restrict__ level0, float const *__restrict__ weights, int w, size_t row_offset, int weight_offset, int x0, float v, float vv1, float two_inv_w, float u_bias, unsigned long long stream_policy, unsigned long long keep_policy, float acc[28])
  {
    float r[kSh9Unroll], g[kSh9Unroll], bl[kSh9Unroll], weight[kSh9Unroll];

    #pragma unroll
    for(int j = 0; j < kSh9Unroll; ++j)
    {
      int x = x0 + j * kSh9Threads + (int)threadIdx.x;
      if (x < w)
      {
        load_texel<FORMAT>(level0, row_offset + x, stream_policy, r[j], g[j], bl[j]);

        // the solid angle is symmetric in x (and y, see weight_offset): only one quadrant of the table is ever
        // touched, 17 MB at 4096^2, which the L2 keeps between faces
        int xs = x < w - 1 - x ? x : w - 1 - x;
        weight[j] = ldg_keep(weights + weight_offset + xs, keep_policy);
      }
      else
      {
        r[j] = g[j] = bl[j] = 0.0f;
        weight[j] = 0.0f;          // a texel past the row end adds exact zeros
      }
    }

    #pragma unroll
    for(int j = 0; j < kSh9Unroll; ++j)
    {
      int x = x0 + j * kSh9Threads + (int)threadIdx.x;

      // project.comp:53-54: u = 2 (x + .5) / w - 1, ray = normalize(rot * (u, v, -1))
      float u = fmaf((float)x, two_inv_w, u_bias);
      float inv = rsqrtf(fmaf(u, u, vv1));

      float rx, ry, rz;
      face_ray<FACE>(u * inv, v * inv, inv, rx, ry, rz);

      sh9_accumulate(acc, weight[j] * r[j], weight[j] * g[j], weight[j] * bl[j], rx, ry, rz);
      acc[27] += weight[j];
    }
  }

  template<int FORMAT>
  __global__ void __launch_bounds__(kSh9Threads, 4) sh9_partial_kernel(void const *__restrict__ level0, float const *__restrict__ weights, int w, int h, int row_begin, int row_end, double *__restrict__ block_partials, unsigned int *__restrict__ done_counter, double *__restrict__ partial, Sh9Peers peers)
  {
    float acc[28];
    #pragma unroll
    for(int k = 0; k < 28; ++k)
      acc[k] = 0.0f;

    const float inv_w = 1.0f / (float)w, inv_h = 1.0f / (float)h;
    const float two_inv_w = 2.0f * inv_w, u_bias = inv_w - 1.0f;
    const unsigned long long stream_policy = l2_policy_evict_first(), keep_policy = l2_policy_evict_last();

    // work items: (row of the slab, segment of that row) in row order; rows are face-major, so the
    // slab of a GPU that shares the cube with others is one contiguous range of the level.  (Tried and
    // measured slower on 4096^2 faces, 0.345 ms as is: item order with the six faces of a table stretch
    // side by side, 0.42 ms; the solid angle from a Taylor form instead of the table, 0.38 ms; five row
    // moments per channel folded once per row, 0.40 ms; the texel stream through cp.async.bulk into a
    // four-stage shared-memory ring with mbarriers, 0.41 ms — profiles/r1_summary.md 0.3.)
    const int segments = (w + kSh9Segment - 1) / kSh9Segment;
    const long long items = (long long)(row_end - row_begin) * segments;

    for(long long item = blockIdx.x; item < items; item += gridDim.x)
    {
      int row = row_begin + (int)(item / segments);
      int x0 = (int)(item % segments) * kSh9Segment;
      int face = row / h;
      int y = row - face * h;

      float v = 2.0f * ((float)y + 0.5f) * inv_h - 1.0f;
      float vv1 = fmaf(v, v, 1.0f);
      size_t row_offset = (size_t)row * w;
      int weight_offset = (y < h - 1 - y ? y : h - 1 - y) * w;

      switch (face)
      {
        case 0: sh9_row_segment<FORMAT, 0>(level0, weights, w, row_offset, weight_offset, x0, v, vv1, two_inv_w, u_bias, stream_policy, keep_policy, acc); break;
        case 1: sh9_row_segment<FORMAT, 1>(level0, weights, w, row_offset, weight_offset, x0, v, vv1, two_inv_w, u_bias, stream_policy, keep_policy, acc); break;
        case 2: sh9_row_segment<FORMAT, 2>(level0, weights, w, row_offset, weight_offset, x0, v, vv1, two_inv_w, u_bias, stream_policy, keep_policy, acc); break;
        case 3: sh9_row_segment<FORMAT, 3>(level0, weights, w, row_offset, weight_offset, x0, v, vv1, two_inv_w, u_bias, stream_policy, keep_policy, acc); break;
        case 4: sh9_row_segment<FORMAT, 4>(level0, weights, w, row_offset, weight_offset, x0, v, vv1, two_inv_w, u_bias, stream_policy, keep_policy, acc); break;
        default: sh9_row_segment<FORMAT, 5>(level0, weights, w, row_offset, weight_offset, x0, v, vv1, two_inv_w, u_bias, stream_policy, keep_policy, acc); break;
      }
    }

    // ---- warp shuffle reduction in fp64, then across the block's warps ----
    __shared__ double s_partial[kSh9Threads / 32][28];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    #pragma unroll
    for(int k = 0; k < 28; ++k)
    {
      double v = (double)acc[k];
      #pragma unroll
      for(int offset = 16; offset > 0; offset >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, offset);

      if (lane == 0)
        s_partial[warp][k] = v;
    }

    __syncthreads();

    __shared__ double s_mono[28];

    if (threadIdx.x < 28)
    {
      double v = 0;
      #pragma unroll
      for(int wi = 0; wi < kSh9Threads / 32; ++wi)
        v += s_partial[wi][threadIdx.x];

      s_mono[threadIdx.x] = v;
    }

    __syncthreads();

    if (threadIdx.x < 28)
      block_partials[(size_t)blockIdx.x * 28 + threadIdx.x] = sh9_basis_from_monomials(threadIdx.x, s_mono);

    // ---- the block that finishes last sums the block partials, in block order: deterministic, and one
    //      launch instead of two (a separate 1-CTA combine kernel cost 7.5 us, as much as a 256^2 cube) ----
    __shared__ bool s_last;

    __threadfence();
    __syncthreads();

    if (threadIdx.x == 0)
    {
      unsigned int ticket = atomicAdd(done_counter, 1u);
      s_last = ticket == gridDim.x - 1;
    }

    __syncthreads();

    if (!s_last)
      return;

    __threadfence();

    // thread (k, part): component k over the blocks part, part + kParts, ...; then the parts in order
    constexpr int kParts = kSh9Threads / 28;              // 9
    __shared__ double s_part[kParts][28];

    if (threadIdx.x < kParts * 28)
    {
      int k = threadIdx.x % 28, part = threadIdx.x / 28;
      double v = 0;

      #pragma unroll 8
      for(int i = part; i < (int)gridDim.x; i += kParts)
        v += __ldcg(block_partials + (size_t)i * 28 + k);

      s_part[part][k] = v;
    }

    __syncthreads();

    if (threadIdx.x < 28)
    {
      double v = 0;
      #pragma unroll
      for(int part = 0; part < kParts; ++part)
        v += s_part[part][threadIdx.x];

      partial[threadIdx.x] = v;

      // a cube shared by several GPUs: the slab's sums go straight into every peer's array
      for(int k = 0; k < peers.count; ++k)
        peers.slots[k][threadIdx.x] = v;
    }

    if (threadIdx.x == 0)
      *done_counter = 0;      // ready for the next launch on this stream
  }

  // ---- irradiance cube from SH9: data/lighting.inc:351-366, 371 ---------------------

  __global__ void __launch_bounds__(256) sh9_irradiance_kernel(Sh9Coefficients sh, int w, int h, uint32_t *__restrict__ words, float *__restrict__ f32)
  {
    size_t total = (size_t)6 * w * h;
    for(size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
    {
      int x = (int)(idx % w);
      int y = (int)((idx / w) % h);
      int face = (int)(idx / ((size_t)w * h));

      float u = 2.0f * ((float)x + 0.5f) / (float)w - 1.0f;
      float v = 2.0f * ((float)y + 0.5f) / (float)h - 1.0f;
      float inv = rsqrtf(fmaf(u, u, fmaf(v, v, 1.0f)));
      float a = u * inv, b = v * inv, c = inv;

      float nx, ny, nz;
      switch (face)
      {
        case 0: nx = c;  ny = b;  nz = a;  break;
        case 1: nx = -c; ny = b;  nz = -a; break;
        case 2: nx = a;  ny = -c; nz = -b; break;
        case 3: nx = a;  ny = c;  nz = b;  break;
        case 4: nx = a;  ny = b;  nz = -c; break;
        default: nx = -a; ny = b; nz = c;  break;
      }

      float L[9];
      L[0] = 3.141593f * 0.282095f;
      L[1] = 2.094395f * 0.488603f * ny;
      L[2] = 2.094395f * 0.488603f * nz;
      L[3] = 2.094395f * 0.488603f * nx;
      L[4] = 0.785398f * 1.092548f * nx * ny;
      L[5] = 0.785398f * 1.092548f * ny * nz;
      L[6] = 0.785398f * 0.315392f * (3.0f * nz * nz - 1.0f);
      L[7] = 0.785398f * 1.092548f * nz * nx;
      L[8] = 0.785398f * 0.546274f * (nx * nx - ny * ny);

      float rgb[3] = { 0.0f, 0.0f, 0.0f };
      #pragma unroll
      for(int k = 0; k < 9; ++k)
      {
        rgb[0] = fmaf(L[k], sh.v[3*k + 0], rgb[0]);
        rgb[1] = fmaf(L[k], sh.v[3*k + 1], rgb[1]);
        rgb[2] = fmaf(L[k], sh.v[3*k + 2], rgb[2]);
      }

      rgb[0] = fmaxf(rgb[0], 0.0f); rgb[1] = fmaxf(rgb[1], 0.0f); rgb[2] = fmaxf(rgb[2], 0.0f);

      if (words)
        words[idx] = rgbe_encode(rgb[0], rgb[1], rgb[2]);

      if (f32)
      {
        f32[3*idx + 0] = rgb[0]; f32[3*idx + 1] = rgb[1]; f32[3*idx + 2] = rgb[2];
      }
    }
  }

  // ---- launchers ----------------------------------------------------------------------

  cudaError_t launch_sh9_weights(float *weights, int w, int h, cudaStream_t stream)
  {
    int total = w * h;
    sh9_weights_kernel<<<(total + 255) / 256, 256, 0, stream>>>(weights, w, h);
    return cudaGetLastError();
  }

  int sh9_partial_blocks(int w, int h, int sm_count)
  {
    // upper bound for any slab of the cube: one CTA per (row, segment) item up to 4 resident CTAs per SM
    long long items = (long long)6 * h * ((w + kSh9Segment - 1) / kSh9Segment);
    long long cap = (long long)sm_count * 4;
    return (int)(items < cap ? (items < 1 ? 1 : items) : cap);
  }

  cudaError_t launch_sh9_partial(void const *level0, int format, float const *weights, int w, int h, int row_begin, int row_end, double *block_partials, int blocks, unsigned int *done_counter, double *partial, Sh9Peers const &peers, int sm_count, cudaStream_t stream)
  {
    (void)sm_count;

    if (format == 0)
      sh9_partial_kernel<0><<<blocks, kSh9Threads, 0, stream>>>(level0, weights, w, h, row_begin, row_end, block_partials, done_counter, partial, peers);
    else
      sh9_partial_kernel<1><<<blocks, kSh9Threads, 0, stream>>>(level0, weights, w, h, row_begin, row_end, block_partials, done_counter, partial, peers);

    return cudaGetLastError();
  }

  cudaError_t launch_sh9_irradiance(Sh9Coefficients const &sh, int w, int h, uint32_t *words, float *f32, int sm_count, cudaStream_t stream)
  {
    size_t total = (size_t)6 * w * h;
    size_t blocks = (total + 255) / 256;
    size_t cap = (size_t)sm_count * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
    if (grid < 1)
      grid = 1;

    sh9_irradiance_kernel<<<grid, 256, 0, stream>>>(sh, w, h, words, f32);

    return cudaGetLastError();
  }
}
