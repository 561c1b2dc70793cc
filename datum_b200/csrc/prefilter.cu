// datum_b200 — GGX prefilter of one cube-map mip level (sm_100a).
//
// Replaces the triple loop of tools/ibl.cpp:263-272 and the per-texel sample
// loop of tools/ibl.cpp:160-187 (reference paths relative to /root/reference).
//
// Work decomposition
//   tile      = 32*TPT output texels (TW x 32/TW lanes, TPT texels per lane)
//   CTA       = NW warps that all work on the SAME tile and split the level's
//               sample table round-robin; partial sums meet in shared memory.
//               (Level 1 of a 512^2 cube is only 393k texels: one thread per
//               texel could not fill 148 SMs, and eight warps walking the same
//               footprint keep the source records hot in L1.)
//   grid      = persistent: min(#tiles, SMs x resident CTAs), tiles strided.
//
// Per sample and texel the loop does: 9 FMA-pipe ops for the reflected
// direction (table entry x tangent frame), one cube-face select + reciprocal,
// a magic-add floor, ONE 16-byte gather of the quad record holding the whole
// 2x2 bilinear footprint, and the biased-mantissa accumulation of ibl_math.cuh.
// No tensor cores: nothing here is a dense contraction.

#include "prefilter.h"
#include "ibl_math.cuh"

#include <cuda_runtime.h>

namespace ibl
{
  // ---- quad records ----------------------------------------------------------
  // rec[f][j][i] = { t(i,j), t(i+1,j), t(i,j+1), t(i+1,j+1) } of the source level,
  // neighbours clamped inside the face (the clamped ones are never addressed:
  // cube_footprint keeps i <= ws-2, j <= hs-2).  One 16-byte load then fetches
  // the whole footprint of ibl.cpp:40.

  __global__ void __launch_bounds__(256) build_quad_records_kernel(uint32_t const *__restrict__ src, uint4 *__restrict__ rec, int ws, int hs)
  {
    size_t total = (size_t)6 * ws * hs;
    for(size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
    {
      int i = (int)(idx % ws);
      int j = (int)((idx / ws) % hs);
      size_t right = (i + 1 < ws) ? 1 : 0;
      size_t down = (j + 1 < hs) ? (size_t)ws : 0;

      uint4 r;
      r.x = __ldg(src + idx);
      r.y = __ldg(src + idx + right);
      r.z = __ldg(src + idx + down);
      r.w = __ldg(src + idx + down + right);
      rec[idx] = r;
    }
  }

  // ---- tile -> texel mapping ---------------------------------------------------

  template<int TW, int TPT>
  __device__ __forceinline__ bool tile_texel(PrefilterParams const &p, int tile, int lane, int k, int &x, int &row)
  {
    constexpr int TH = 32 / TW;
    if (p.tiles_x > 0)
    {
      int tx = tile % p.tiles_x;
      int ty = tile / p.tiles_x;
      x = tx * TW + (lane % TW);
      row = p.row_begin + ty * (TH * TPT) + k * TH + (lane / TW);
      return x < p.wd && row < p.row_end;
    }
    else
    {
      // levels narrower than a tile: texels of the slab taken in linear order
      int t = tile * (32 * TPT) + k * 32 + lane;
      x = t % p.wd;
      row = p.row_begin + t / p.wd;
      return row < p.row_end;
    }
  }

  // ---- the prefilter kernel ------------------------------------------------------

  template<int TW, int TPT, int NW>
  __global__ void __launch_bounds__(32 * NW) prefilter_level_kernel(PrefilterParams p)
  {
    extern __shared__ float4 smem[];
    float4 *s_table = smem;
    float *s_red = reinterpret_cast<float*>(smem + p.table_count);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    for(int i = tid; i < p.table_count; i += 32 * NW)
      s_table[i] = __ldg(p.table + i);

    __syncthreads();

    for(int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x)
    {
      Vec3f N[TPT], T[TPT], B[TPT];
      float acc[TPT][4];

      #pragma unroll
      for(int k = 0; k < TPT; ++k)
      {
        int x, row;
        bool valid = tile_texel<TW, TPT>(p, tile, lane, k, x, row);
        if (!valid) { x = 0; row = p.row_begin; }

        int face = row / p.hd;
        int y = row - face * p.hd;

        N[k] = texel_normal(p.quats[face], x, y, p.wd, p.hd);
        tangent_frame(N[k], T[k], B[k]);

        acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.0f;
      }

      #pragma unroll 2
      for(int s = warp; s < p.table_count; s += NW)
      {
        const float4 e = s_table[s];

        #pragma unroll
        for(int k = 0; k < TPT; ++k)
        {
          float Lx = fmaf(e.z, N[k].x, fmaf(e.y, B[k].x, e.x * T[k].x));
          float Ly = fmaf(e.z, N[k].y, fmaf(e.y, B[k].y, e.x * T[k].y));
          float Lz = fmaf(e.z, N[k].z, fmaf(e.y, B[k].z, e.x * T[k].z));

          float du, dv;
          uint32_t idx = cube_footprint(p.geom, Lx, Ly, Lz, du, dv);

          const uint4 rec = __ldg(p.records + idx);

          float w[4];
          footprint_weights(du, dv, e.w, e.z, w);

          accumulate_tap(rec.x, w[0], acc[k]);
          accumulate_tap(rec.y, w[1], acc[k]);
          accumulate_tap(rec.z, w[2], acc[k]);
          accumulate_tap(rec.w, w[3], acc[k]);
        }
      }

      // ---- cross-warp reduction: s_red[((warp*TPT + k)*4 + c)*32 + lane] ----
      #pragma unroll
      for(int k = 0; k < TPT; ++k)
      {
        #pragma unroll
        for(int c = 0; c < 4; ++c)
          s_red[((warp * TPT + k) * 4 + c) * 32 + lane] = acc[k][c];
      }

      __syncthreads();

      if (tid < 32 * TPT)
      {
        const int k = tid >> 5;

        float sum[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
        #pragma unroll
        for(int w = 0; w < NW; ++w)
        {
          #pragma unroll
          for(int c = 0; c < 4; ++c)
            sum[c] += s_red[((w * TPT + k) * 4 + c) * 32 + lane];
        }

        int x, row;
        if (tile_texel<TW, TPT>(p, tile, lane, k, x, row))
        {
          // sum/totalweight of ibl.cpp:186, then rgbe() of ibl.cpp:269
          float r = (sum[0] - sum[3]) * p.norm;
          float g = (sum[1] - sum[3]) * p.norm;
          float b = (sum[2] - sum[3]) * p.norm;

          size_t o = (size_t)row * p.wd + x;

          if (p.dst_words)
            p.dst_words[o] = rgbe_encode(r, g, b);

          if (p.dst_f32)
          {
            p.dst_f32[3*o + 0] = r;
            p.dst_f32[3*o + 1] = g;
            p.dst_f32[3*o + 2] = b;
          }
        }
      }

      __syncthreads();
    }
  }

  // ---- host-side launchers -----------------------------------------------------

  namespace
  {
    template<int TW, int TPT, int NW>
    cudaError_t launch_variant(PrefilterParams p, int sm_count, cudaStream_t stream, int *launched_grid)
    {
      constexpr int TH = 32 / TW;
      auto kernel = prefilter_level_kernel<TW, TPT, NW>;

      int rows = p.row_end - p.row_begin;
      if (p.wd >= TW)
      {
        p.tiles_x = (p.wd + TW - 1) / TW;
        p.tiles = p.tiles_x * ((rows + TH * TPT - 1) / (TH * TPT));
      }
      else
      {
        p.tiles_x = 0;
        p.tiles = (rows * p.wd + 32 * TPT - 1) / (32 * TPT);
      }

      size_t smem = (size_t)p.table_count * sizeof(float4) + (size_t)NW * TPT * 4 * 32 * sizeof(float);

      cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err != cudaSuccess)
        return err;

      int resident = 0;
      err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, 32 * NW, smem);
      if (err != cudaSuccess)
        return err;
      if (resident < 1)
        return cudaErrorLaunchOutOfResources;

      int grid = p.tiles < sm_count * resident ? p.tiles : sm_count * resident;
      if (grid < 1)
        grid = 1;

      kernel<<<grid, 32 * NW, smem, stream>>>(p);

      if (launched_grid)
        *launched_grid = grid;

      return cudaGetLastError();
    }
  }

  cudaError_t launch_build_quad_records(uint32_t const *src, uint4 *records, int ws, int hs, int sm_count, cudaStream_t stream)
  {
    size_t total = (size_t)6 * ws * hs;
    size_t blocks = (total + 255) / 256;
    size_t cap = (size_t)sm_count * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
    if (grid < 1)
      grid = 1;

    build_quad_records_kernel<<<grid, 256, 0, stream>>>(src, records, ws, hs);

    return cudaGetLastError();
  }

  cudaError_t launch_prefilter_level(PrefilterParams const &p, int variant, int sm_count, cudaStream_t stream, int *launched_grid)
  {
    int rows = p.row_end - p.row_begin;
    if (rows <= 0 || p.wd <= 0)
      return cudaSuccess;

    // Small slabs: fewer texels per tile and more warps per tile so the few
    // tiles there are still spread over the machine.
    size_t texels = (size_t)rows * p.wd;

    if (variant == 0)
      variant = (texels >= 64u * 148u * 4u) ? 1 : 2;

    switch (variant)
    {
      case 1: return launch_variant<8, 2, 8>(p, sm_count, stream, launched_grid);
      case 2: return launch_variant<8, 1, 16>(p, sm_count, stream, launched_grid);
      case 3: return launch_variant<16, 2, 8>(p, sm_count, stream, launched_grid);
      case 4: return launch_variant<32, 2, 8>(p, sm_count, stream, launched_grid);
      case 5: return launch_variant<8, 2, 4>(p, sm_count, stream, launched_grid);
      case 6: return launch_variant<8, 1, 8>(p, sm_count, stream, launched_grid);
      case 7: return launch_variant<16, 1, 8>(p, sm_count, stream, launched_grid);
      default: return cudaErrorInvalidValue;
    }
  }
}
