// datum_b200 — env-BRDF and water-colour LUT kernels, FP32 peak micro-benchmark (sm_100a).
//
// image_pack_envbrdf (tools/ibl.cpp:292-308) integrates, per LUT texel, 1024 GGX
// samples (a, b: ibl.cpp:198-217) and 1024 cosine samples (c: ibl.cpp:221-234).
// One warp owns one texel; lanes take samples round-robin and meet in a
// shuffle reduction.  With the normal fixed at +z the tangent frame of
// ibl.cpp:123-125 is the constant T = (0,-1,0), B = (1,0,0).

#include "luts.h"
#include "ibl_math.cuh"

#include <cuda_runtime.h>

namespace ibl
{
  __device__ __forceinline__ float ggx_g1(float ndotx, float alpha)
  {
    float k = alpha * 0.5f; // ibl.cpp:111-115
    return ndotx / (ndotx * (1.0f - k) + k);
  }

  __device__ __forceinline__ float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }

  __device__ __forceinline__ float clamp01(float x) { return fmaxf(0.0f, fminf(x, 1.0f)); }

  __device__ __forceinline__ float warp_sum(float v)
  {
    #pragma unroll
    for(int offset = 16; offset > 0; offset >>= 1)
      v += __shfl_xor_sync(0xffffffffu, v, offset);
    return v;
  }

  __global__ void __launch_bounds__(256) envbrdf_kernel(int width, int height, int samples, uint32_t *__restrict__ words, float *__restrict__ f32)
  {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int total = width * height;

    for(int texel = blockIdx.x * warps_per_block + (threadIdx.x >> 5); texel < total; texel += gridDim.x * warps_per_block)
    {
      int x = texel % width, y = texel / width;

      // ibl.cpp:300-301, 193
      float NdotV = ((float)x + 0.5f) / (float)width;
      float roughness = ((float)y + 0.5f) / (float)height;
      float alpha = roughness * roughness;
      float a2m1 = alpha * alpha - 1.0f;
      float Vx = sqrtf(1.0f - NdotV * NdotV), Vz = NdotV;

      float a = 0.0f, b = 0.0f, c = 0.0f;

      for(int i = lane; i < samples; i += 32)
      {
        float ux = (float)i / (float)samples;
        float uy = (float)__brev((unsigned)i) * 2.3283064365386963e-10f;

        float sinphi, cosphi;
        sincosf(6.2831855f * ux, &sinphi, &cosphi);

        // ---- specular terms, ibl.cpp:200-216 ----
        {
          float costheta = sqrtf((1.0f - uy) / (1.0f + a2m1 * uy));
          float sintheta = sqrtf(1.0f - costheta * costheta);

          float Hx = sintheta * sinphi, Hy = -sintheta * cosphi, Hz = costheta;
          float VdotHraw = Vx * Hx + Vz * Hz;
          float Lz = 2.0f * VdotHraw * Hz - Vz;

          float NdotL = clamp01(Lz);
          float NdotH = clamp01(Hz);
          float VdotH = clamp01(VdotHraw);
          (void)Hy;

          if (NdotL > 0.0f)
          {
            float G = ggx_g1(NdotL, alpha) * ggx_g1(NdotV, alpha);
            float Vis = G * VdotH / (NdotH * NdotV);
            float Fc = pow5(1.0f - VdotH);

            a += (1.0f - Fc) * Vis;
            b += Fc * Vis;
          }
        }

        // ---- diffuse term, ibl.cpp:223-233, 148-158 ----
        {
          float hx = ux + 0.5f, hy = uy + 0.5f;
          float cx = hx - floorf(hx), cy = hy - floorf(hy);

          float sp, cp;
          sincosf(6.2831855f * cx, &sp, &cp);

          float costheta = sqrtf(fmaxf(0.0f, 1.0f - cy));
          float sintheta = sqrtf(cy);

          float Lx = sintheta * sp, Ly = -sintheta * cp, Lz = costheta;
          float NdotL = clamp01(Lz);

          if (NdotL > 0.0f)
          {
            float hvx = Vx + Lx, hvy = Ly, hvz = Vz + Lz;
            float inv = rsqrtf(hvx * hvx + hvy * hvy + hvz * hvz);
            float LdotH = clamp01((Lx * hvx + Ly * hvy + Lz * hvz) * inv);

            float energyfactor = (1.0f - alpha) * 1.0f + alpha * (1.0f / 1.51f);
            float f90 = 0.5f + 2.0f * LdotH * LdotH * alpha;
            float lightscatter = 1.0f + (f90 - 1.0f) * pow5(1.0f - NdotL);
            float viewscatter = 1.0f + (f90 - 1.0f) * pow5(1.0f - NdotV);

            c += lightscatter * viewscatter * energyfactor;
          }
        }
      }

      a = warp_sum(a) / (float)samples;
      b = warp_sum(b) / (float)samples;
      c = warp_sum(c) / (float)samples;

      if (lane == 0)
      {
        if (words)
          words[texel] = rgbe_encode(a, b, c);

        if (f32)
        {
          f32[3*texel + 0] = a; f32[3*texel + 1] = b; f32[3*texel + 2] = c;
        }
      }
    }
  }

  __global__ void __launch_bounds__(256) watercolor_kernel(WaterColorParams p, int width, int height, uint32_t *__restrict__ words)
  {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= width * height)
      return;

    int x = idx % width, y = idx / width;

    // ibl.cpp:320-326
    float scale = ((float)x + 0.5f) / (float)width;
    float facing = ((float)y + 0.5f) / (float)height;
    float fresnel = clamp01(p.fresnelbias + powf(facing, p.fresnelpower));
    float depth = clamp01(1.0f - exp2f(-p.depthscale * scale * 100.0f));

    float out[3];
    #pragma unroll
    for(int c = 0; c < 3; ++c)
    {
      float color = (1.0f - depth) * p.shallow[c] + depth * p.deep[c];
      out[c] = (1.0f - fresnel) * color + fresnel * p.fresnel[c];
    }

    words[idx] = rgbe_encode(out[0], out[1], out[2]);
  }

  __global__ void __launch_bounds__(256) fma_peak_kernel(float *sink, int iters)
  {
    float x[16];
    #pragma unroll
    for(int k = 0; k < 16; ++k)
      x[k] = (float)(threadIdx.x + k) * 1e-3f;

    const float m = 0.999f, a = 1e-4f;

    for(int i = 0; i < iters; ++i)
    {
      #pragma unroll
      for(int k = 0; k < 16; ++k)
        x[k] = fmaf(x[k], m, a);
    }

    float s = 0.0f;
    #pragma unroll
    for(int k = 0; k < 16; ++k)
      s += x[k];

    if (s == 123.456f)
      sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }

  // the same chains with the packed two-wide fma.rn.f32x2 (FFMA2) of sm_100: 8 chains of float2
  __global__ void __launch_bounds__(256) fma2_peak_kernel(float *sink, int iters)
  {
    unsigned long long x[8];
    #pragma unroll
    for(int k = 0; k < 8; ++k)
    {
      float2 v = make_float2((float)(threadIdx.x + k) * 1e-3f, (float)(threadIdx.x + k + 8) * 1e-3f);
      x[k] = *reinterpret_cast<unsigned long long*>(&v);
    }

    float2 mv = make_float2(0.999f, 0.999f), av = make_float2(1e-4f, 1e-4f);
    unsigned long long m = *reinterpret_cast<unsigned long long*>(&mv), a = *reinterpret_cast<unsigned long long*>(&av);

    for(int i = 0; i < iters; ++i)
    {
      #pragma unroll
      for(int k = 0; k < 8; ++k)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[k]) : "l"(m), "l"(a));
    }

    float s = 0.0f;
    #pragma unroll
    for(int k = 0; k < 8; ++k)
    {
      float2 v = *reinterpret_cast<float2*>(&x[k]);
      s += v.x + v.y;
    }

    if (s == 123.456f)
      sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }

  cudaError_t launch_fma2_peak(float *sink, int blocks, int threads, int iters, cudaStream_t stream)
  {
    fma2_peak_kernel<<<blocks, threads, 0, stream>>>(sink, iters);
    return cudaGetLastError();
  }

  cudaError_t launch_envbrdf(int width, int height, int samples, uint32_t *words, float *f32, cudaStream_t stream)
  {
    int total = width * height;
    int blocks = (total + 7) / 8;
    envbrdf_kernel<<<blocks, 256, 0, stream>>>(width, height, samples, words, f32);
    return cudaGetLastError();
  }

  cudaError_t launch_watercolor(WaterColorParams const &params, int width, int height, uint32_t *words, cudaStream_t stream)
  {
    int total = width * height;
    watercolor_kernel<<<(total + 255) / 256, 256, 0, stream>>>(params, width, height, words);
    return cudaGetLastError();
  }

  cudaError_t launch_fma_peak(float *sink, int blocks, int threads, int iters, cudaStream_t stream)
  {
    fma_peak_kernel<<<blocks, threads, 0, stream>>>(sink, iters);
    return cudaGetLastError();
  }
}
