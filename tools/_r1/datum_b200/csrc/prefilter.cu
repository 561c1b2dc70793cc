// datum_b200 — GGX prefilter of one cube-map mip level, first kernel (sm_100a).
//
// Replaces the triple loop of tools/ibl.cpp:263-272 and the per-texel sample
// loop of tools/ibl.cpp:160-187 (reference paths relative to /root/reference).
// The library uses it for levels narrower than 8 texels (its tiles may be cut in
// linear texel order); wider levels run prefilter_dn.cu, which replaced this
// kernel's six-logic-op tap decode.  Variants 10..27 stay selectable for A/B timing.
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

  // One sample of one texel, split in two halves so the gather of sample s+1 can be
  // issued before the arithmetic on sample s (software pipelining: the 16-byte record
  // load is the only long-latency operation of the loop).
  struct Fetched
  {
    uint4 rec;     // the 2x2 footprint
    float du, dv;  // bilinear fractions - 0.5
    float nl, wh;  // NdotL and 0.5*NdotL of the sample
  };

  __device__ __forceinline__ Fetched fetch_general(PrefilterParams const &p, TexelState const &t, float4 e)
  {
    float Lx = fmaf(e.z, t.N.x, fmaf(e.y, t.B.x, e.x * t.T.x));
    float Ly = fmaf(e.z, t.N.y, fmaf(e.y, t.B.y, e.x * t.T.y));
    float Lz = fmaf(e.z, t.N.z, fmaf(e.y, t.B.z, e.x * t.T.z));

    Fetched f;
    uint32_t idx = cube_footprint(p.geom, Lx, Ly, Lz, f.du, f.dv);
    f.rec = __ldg(p.records + idx);
    f.nl = e.z;
    f.wh = e.w;
    return f;
  }

  // ---- packed two-wide fp32 (fma.rn.f32x2 -> SASS FFMA2/FMUL2/FADD2, new on sm_100) ----
  //
  // One FFMA2 does two FMAs for one issue slot (same lane throughput as two
  // FFMAs: measured 73 vs 72 TFLOP/s), and the hardware takes a plain fp32
  // register as a broadcast operand, so pairing costs no moves.  The loop is
  // bound by issue slots and the half-rate ALU pipe, not by FMA lanes; packing
  // the (a, b) face coordinates, the bilinear weights and the (r, g) / (b, bias)
  // accumulators removes ~18 of ~74 issue slots per sample.  Every element goes
  // through the same round-to-nearest operations as the scalar form in
  // ibl_math.cuh, so the results are bit-identical to it.
  typedef unsigned long long f32x2;

  __device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
  __device__ __forceinline__ void unpack2(f32x2 a, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); }
  __device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
  __device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
  __device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
  __device__ __forceinline__ f32x2 bcast2(float v) { return pack2(v, v); }

  // face_footprint of ibl_math.cuh with the (a, b) coordinates carried as a pair
  __device__ __forceinline__ Fetched fetch_same_face_packed(PrefilterParams const &p, TexelState const &t, float4 e)
  {
    f32x2 lab = mul2(bcast2(e.x), pack2(t.T.x, t.T.y));
    lab = fma2(bcast2(e.y), pack2(t.B.x, t.B.y), lab);
    lab = fma2(bcast2(e.z), pack2(t.N.x, t.N.y), lab);
    float lm = fmaf(e.z, t.N.z, fmaf(e.y, t.B.z, e.x * t.T.z));

    float r = rcp_fast(lm);
    f32x2 f = fma2(lab, bcast2(r), pack2(p.geom.hwm, p.geom.hhm));
    f32x2 m = add2(f, bcast2(kMagic));
    f32x2 fi = add2(m, bcast2(-kMagic));
    f32x2 d = fma2(fi, bcast2(-1.0f), f);

    float mu, mv;
    unpack2(m, mu, mv);

    Fetched out;
    unpack2(d, out.du, out.dv);
    uint32_t idx = f2u(mv) * (uint32_t)p.geom.ws + f2u(mu) + t.face_base;
    out.rec = __ldg(p.records + idx);
    out.nl = e.z;
    out.wh = e.w;
    return out;
  }

  // footprint_weights + accumulate_tap of ibl_math.cuh on (r, g) and (b, bias) accumulator pairs
  __device__ __forceinline__ void accumulate_tap_packed(DecodeMasks const &k, uint32_t word, float w, f32x2 &acc_rg, f32x2 &acc_bs)
  {
    uint32_t eb = (word & kMaskExpo) | k.bias;
    uint32_t fb = (word & kMaskExpMant) | k.bias;
    uint32_t fg = ((word << 9) & kMaskMant) | eb;
    uint32_t fr = (((word << 18) | (word >> 14)) & kMaskMant) | eb;
    f32x2 wv = bcast2(w);
    acc_rg = fma2(pack2(u2f(fr), u2f(fg)), wv, acc_rg);
    acc_bs = fma2(pack2(u2f(fb), u2f(eb)), wv, acc_bs);
  }

  __device__ __forceinline__ void consume_packed(PrefilterParams const &p, Fetched const &f, f32x2 &acc_rg, f32x2 &acc_bs)
  {
    float u0 = 0.5f - f.du, u1 = 0.5f + f.du;
    float v0 = fmaf(-f.dv, f.nl, f.wh), v1 = fmaf(f.dv, f.nl, f.wh);

    f32x2 u = pack2(u0, u1);
    float w00, w10, w01, w11;
    unpack2(mul2(u, bcast2(v0)), w00, w10);
    unpack2(mul2(u, bcast2(v1)), w01, w11);

    accumulate_tap_packed(p.masks, f.rec.x, w00, acc_rg, acc_bs);
    accumulate_tap_packed(p.masks, f.rec.y, w10, acc_rg, acc_bs);
    accumulate_tap_packed(p.masks, f.rec.z, w01, acc_rg, acc_bs);
    accumulate_tap_packed(p.masks, f.rec.w, w11, acc_rg, acc_bs);
  }

  // ---- the kernel with packed arithmetic -------------------------------------------

  template<int TW, int TPT, int NW, int UNROLL, int MINB>
  __global__ void __launch_bounds__(32 * NW, MINB) prefilter_level_packed_kernel(PrefilterParams p)
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
      f32x2 acc_rg[TPT], acc_bs[TPT];
      float threshold = 0.0f;

      #pragma unroll
      for(int k = 0; k < TPT; ++k)
      {
        int x, row;
        bool valid = tile_texel<TW, TPT>(p, tile, lane, k, x, row);

        if (!valid) { x = p.wd >> 1; row = (p.row_begin / p.hd) * p.hd + (p.hd >> 1); }

        int face = row / p.hd;
        int y = row - face * p.hd;

        Vec3f N = texel_normal(p.quats[face], x, y, p.wd, p.hd);
        Vec3f T, B;
        tangent_frame(N, T, B);

        Vec3f Tl = to_face_local(face, T), Bl = to_face_local(face, B), Nl = to_face_local(face, N);

        threshold = fmaxf(threshold, same_face_threshold(Nl));

        st[k].T = Vec3f{ Tl.x * p.geom.hw, Tl.y * p.geom.hh, Tl.z };
        st[k].B = Vec3f{ Bl.x * p.geom.hw, Bl.y * p.geom.hh, Bl.z };
        st[k].N = Vec3f{ Nl.x * p.geom.hw, Nl.y * p.geom.hh, Nl.z };
        st[k].face = face;
        st[k].face_base = (uint32_t)face * p.geom.face_size - p.geom.bias;

        acc_rg[k] = 0ull;
        acc_bs[k] = 0ull;
      }

      threshold = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(threshold)));

      int n_same = 0;
      {
        int lo = 0, hi = p.table_count;
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
        {
          Fetched f = fetch_same_face_packed(p, st[k], e);
          consume_packed(p, f, acc_rg[k], acc_bs[k]);
        }
      }

      if (s < p.table_count)
      {
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
          {
            Fetched f = fetch_general(p, st[k], e);
            consume_packed(p, f, acc_rg[k], acc_bs[k]);
          }
        }
      }

      #pragma unroll
      for(int k = 0; k < TPT; ++k)
      {
        float a[4];
        unpack2(acc_rg[k], a[0], a[1]);
        unpack2(acc_bs[k], a[2], a[3]);

        #pragma unroll
        for(int c = 0; c < 4; ++c)
          s_red[((warp * TPT + k) * 4 + c) * 32 + lane] = a[c];
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
    template<int TW, int TPT, int NW, int UNROLL, int MINB>
    cudaError_t launch_packed(PrefilterParams p, int sm_count, cudaStream_t stream, int *launched_grid)
    {
      constexpr int TH = 32 / TW;
      auto kernel = prefilter_level_packed_kernel<TW, TPT, NW, UNROLL, MINB>;

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

    // Automatic choice by slab size.  All use the packed kernel with 8x4-texel tiles, one
    // texel per lane; what changes is how many warps share a tile's samples: big slabs have
    // enough tiles to fill the machine with 4-warp CTAs (cheapest reduction, most CTAs per SM),
    // small slabs split the samples 8, 16 or 32 ways so that the few tiles still spread out.
    size_t texels = (size_t)rows * p.wd;

    if (variant == 0)
    {
      if (texels >= 32u * 148u * 8u)
        variant = 19;
      else if (texels >= 32u * 148u)
        variant = 17;
      else if (texels >= 32u * 24u)
        variant = 14;
      else
        variant = 27;
    }

    switch (variant)
    {
      case 10: return launch_packed<8, 1, 8, 2, 1>(p, sm_count, stream, launched_grid);
      case 14: return launch_packed<8, 1, 16, 2, 1>(p, sm_count, stream, launched_grid);
      case 17: return launch_packed<8, 1, 8, 2, 4>(p, sm_count, stream, launched_grid);
      case 19: return launch_packed<8, 1, 4, 2, 8>(p, sm_count, stream, launched_grid);
      case 27: return launch_packed<8, 1, 32, 1, 1>(p, sm_count, stream, launched_grid);
      default: return cudaErrorInvalidValue;
    }
  }
}
