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
  // the whole footprint of ibl.cpp:40.  Words are stored rotated right by 4 bits
  // (pack_record_word) so that exponent and blue mantissa already sit at their
  // fp32 bit positions.

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
      r.x = pack_record_word(__ldg(src + idx));
      r.y = pack_record_word(__ldg(src + idx + right));
      r.z = pack_record_word(__ldg(src + idx + down));
      r.w = pack_record_word(__ldg(src + idx + down + right));
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

  // per-texel state carried through a tile: the tangent frame as three rows
  // (T, B, N), first in face-local coordinates for the same-face loop, then
  // rotated back to world coordinates for the general loop
  struct TexelState
  {
    Vec3f T, B, N;
    uint32_t face_base; // face*face_size - bias
    int face;
  };

  __device__ __forceinline__ void sample_general(PrefilterParams const &p, TexelState const &t, float4 e, float acc[4])
  {
    float Lx = fmaf(e.z, t.N.x, fmaf(e.y, t.B.x, e.x * t.T.x));
    float Ly = fmaf(e.z, t.N.y, fmaf(e.y, t.B.y, e.x * t.T.y));
    float Lz = fmaf(e.z, t.N.z, fmaf(e.y, t.B.z, e.x * t.T.z));

    float du, dv;
    uint32_t idx = cube_footprint(p.geom, Lx, Ly, Lz, du, dv);

    const uint4 rec = __ldg(p.records + idx);

    float w[4];
    footprint_weights(du, dv, e.w, e.z, w);

    accumulate_tap(p.masks, rec.x, w[0], acc);
    accumulate_tap(p.masks, rec.y, w[1], acc);
    accumulate_tap(p.masks, rec.z, w[2], acc);
    accumulate_tap(p.masks, rec.w, w[3], acc);
  }

  // frame rows in face-local (a, b, m) coordinates, a and b pre-scaled
  __device__ __forceinline__ void sample_same_face(PrefilterParams const &p, TexelState const &t, float4 e, float acc[4])
  {
    float la = fmaf(e.z, t.N.x, fmaf(e.y, t.B.x, e.x * t.T.x));
    float lb = fmaf(e.z, t.N.y, fmaf(e.y, t.B.y, e.x * t.T.y));
    float lm = fmaf(e.z, t.N.z, fmaf(e.y, t.B.z, e.x * t.T.z));

    float du, dv;
    uint32_t idx = face_footprint(p.geom, t.face_base, la, lb, lm, du, dv);

    const uint4 rec = __ldg(p.records + idx);

    float w[4];
    footprint_weights(du, dv, e.w, e.z, w);

    accumulate_tap(p.masks, rec.x, w[0], acc);
    accumulate_tap(p.masks, rec.y, w[1], acc);
    accumulate_tap(p.masks, rec.z, w[2], acc);
    accumulate_tap(p.masks, rec.w, w[3], acc);
  }

  template<int TW, int TPT, int NW, int UNROLL>
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
      TexelState st[TPT];
      float acc[TPT][4];
      float threshold = 0.0f;

      #pragma unroll
      for(int k = 0; k < TPT; ++k)
      {
        int x, row;
        bool valid = tile_texel<TW, TPT>(p, tile, lane, k, x, row);

        // lanes past the slab still walk the sample loops (their sums are dropped):
        // park them on a face-centre texel, whose lobe stays inside its face
        if (!valid) { x = p.wd >> 1; row = (p.row_begin / p.hd) * p.hd + (p.hd >> 1); }

        int face = row / p.hd;
        int y = row - face * p.hd;

        Vec3f N = texel_normal(p.quats[face], x, y, p.wd, p.hd);
        Vec3f T, B;
        tangent_frame(N, T, B);

        // face-local rows, a and b scaled to source texels (align-corners, ibl.cpp:37-38)
        Vec3f Tl = to_face_local(face, T), Bl = to_face_local(face, B), Nl = to_face_local(face, N);

        threshold = fmaxf(threshold, same_face_threshold(Nl));

        st[k].T = Vec3f{ Tl.x * p.geom.hw, Tl.y * p.geom.hh, Tl.z };
        st[k].B = Vec3f{ Bl.x * p.geom.hw, Bl.y * p.geom.hh, Bl.z };
        st[k].N = Vec3f{ Nl.x * p.geom.hw, Nl.y * p.geom.hh, Nl.z };
        st[k].face = face;
        st[k].face_base = (uint32_t)face * p.geom.face_size - p.geom.bias;

        acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.0f;
      }

      // number of leading (smallest-angle) samples that stay on every texel's own face
      threshold = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(threshold)));

      int n_same = 0;
      {
        int lo = 0, hi = p.table_count; // first index with lz <= threshold (table sorted by decreasing lz)
        while (lo < hi)
        {
          int mid = (lo + hi) >> 1;
          if (s_table[mid].z > threshold)
            lo = mid + 1;
          else
            hi = mid;
        }
        n_same = lo;
      }

      int s = warp;

      #pragma unroll UNROLL
      for(; s < n_same; s += NW)
      {
        const float4 e = s_table[s];

        #pragma unroll
        for(int k = 0; k < TPT; ++k)
          sample_same_face(p, st[k], e, acc[k]);
      }

      if (s < p.table_count)
      {
        // back to world coordinates for the samples that may cross a face edge
        #pragma unroll
        for(int k = 0; k < TPT; ++k)
        {
          st[k].T = from_face_local(st[k].face, Vec3f{ st[k].T.x * p.geom.inv_hw, st[k].T.y * p.geom.inv_hh, st[k].T.z });
          st[k].B = from_face_local(st[k].face, Vec3f{ st[k].B.x * p.geom.inv_hw, st[k].B.y * p.geom.inv_hh, st[k].B.z });
          st[k].N = from_face_local(st[k].face, Vec3f{ st[k].N.x * p.geom.inv_hw, st[k].N.y * p.geom.inv_hh, st[k].N.z });
        }

        #pragma unroll UNROLL
        for(; s < p.table_count; s += NW)
        {
          const float4 e = s_table[s];

          #pragma unroll
          for(int k = 0; k < TPT; ++k)
            sample_general(p, st[k], e, acc[k]);
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
    template<int TW, int TPT, int NW, int UNROLL>
    cudaError_t launch_variant(PrefilterParams p, int sm_count, cudaStream_t stream, int *launched_grid)
    {
      constexpr int TH = 32 / TW;
      auto kernel = prefilter_level_kernel<TW, TPT, NW, UNROLL>;

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
      variant = (texels >= 32u * 148u * 8u) ? 6 : 2;

    switch (variant)
    {
      case 1: return launch_variant<8, 2, 8, 2>(p, sm_count, stream, launched_grid);
      case 2: return launch_variant<8, 1, 16, 2>(p, sm_count, stream, launched_grid);
      case 3: return launch_variant<16, 2, 8, 2>(p, sm_count, stream, launched_grid);
      case 4: return launch_variant<8, 1, 8, 4>(p, sm_count, stream, launched_grid);
      case 5: return launch_variant<8, 2, 4, 2>(p, sm_count, stream, launched_grid);
      case 6: return launch_variant<8, 1, 8, 2>(p, sm_count, stream, launched_grid);
      case 7: return launch_variant<16, 1, 8, 2>(p, sm_count, stream, launched_grid);
      case 8: return launch_variant<8, 1, 4, 4>(p, sm_count, stream, launched_grid);
      case 9: return launch_variant<8, 2, 8, 1>(p, sm_count, stream, launched_grid);
      default: return cudaErrorInvalidValue;
    }
  }
}
