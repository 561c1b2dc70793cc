// Micro-benchmarks of the issue/pipe rates that bound the prefilter loop on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096

// K0: FFMA, two operands loop-invariant (reuse-cache friendly): the usual "peak" loop
__global__ void k_ffma_const(float *out, float m, float a)
{
  float x[16];
  for(int k = 0; k < 16; ++k) x[k] = threadIdx.x * 0.001f + k;
  for(int i = 0; i < ITERS; ++i)
  {
    #pragma unroll
    for(int k = 0; k < 16; ++k) x[k] = fmaf(x[k], m, a);
  }
  float s = 0; for(int k = 0; k < 16; ++k) s += x[k];
  if (s == 1.2345f) out[threadIdx.x] = s;
}

// K1: FFMA with three distinct per-thread registers per instruction
__global__ void k_ffma_3reg(float *out, float m, float a)
{
  float x[16], y[16], z[16];
  for(int k = 0; k < 16; ++k) { x[k] = threadIdx.x * 0.001f + k; y[k] = m + k * 1e-6f + threadIdx.x * 1e-7f; z[k] = a + k * 1e-6f + threadIdx.x * 1e-7f; }
  for(int i = 0; i < ITERS; ++i)
  {
    #pragma unroll
    for(int k = 0; k < 16; ++k) x[k] = fmaf(x[k], y[k], z[k]);
  }
  float s = 0; for(int k = 0; k < 16; ++k) s += x[k];
  if (s == 1.2345f) out[threadIdx.x] = s;
}

// K2: LOP3 with three register operands (a & b) | c, chains
__global__ void k_lop3_3reg(uint32_t *out, uint32_t m0, uint32_t m1)
{
  uint32_t x[16], y[16], z[16];
  for(int k = 0; k < 16; ++k) { x[k] = threadIdx.x * 7u + k; y[k] = m0 + k * 3u + threadIdx.x; z[k] = m1 ^ (k * 5u + threadIdx.x); }
  for(int i = 0; i < ITERS; ++i)
  {
    #pragma unroll
    for(int k = 0; k < 16; ++k) x[k] = (x[k] & y[k]) ^ z[k];
  }
  uint32_t s = 0; for(int k = 0; k < 16; ++k) s += x[k];
  if (s == 12345u) out[threadIdx.x] = s;
}

// K3: LOP3 with one register and immediates
__global__ void k_lop3_imm(uint32_t *out, uint32_t m0)
{
  uint32_t x[16];
  for(int k = 0; k < 16; ++k) x[k] = threadIdx.x * 7u + k + m0;
  for(int i = 0; i < ITERS; ++i)
  {
    #pragma unroll
    for(int k = 0; k < 16; ++k) x[k] = (x[k] ^ 0x5bd1e995u) + 0u, x[k] = (x[k] & 0xfffffff7u) | (x[k] >> 0) ;
  }
  uint32_t s = 0; for(int k = 0; k < 16; ++k) s += x[k];
  if (s == 12345u) out[threadIdx.x] = s;
}

// K4: interleaved 1 LOP3(3 reg) : 1 FFMA(3 reg)
__global__ void k_mix(uint32_t *out, uint32_t m0, uint32_t m1, float m, float a)
{
  uint32_t x[8], y[8], z[8];
  float fx[8], fy[8], fz[8];
  for(int k = 0; k < 8; ++k) { x[k] = threadIdx.x * 7u + k; y[k] = m0 + k * 3u + threadIdx.x; z[k] = m1 ^ (k * 5u + threadIdx.x);
    fx[k] = threadIdx.x * 0.001f + k; fy[k] = m + k * 1e-6f; fz[k] = a + k * 1e-6f + threadIdx.x * 1e-7f; }
  for(int i = 0; i < ITERS; ++i)
  {
    #pragma unroll
    for(int k = 0; k < 8; ++k) { x[k] = (x[k] & y[k]) ^ z[k]; fx[k] = fmaf(fx[k], fy[k], fz[k]); }
  }
  uint32_t s = 0; float fs = 0; for(int k = 0; k < 8; ++k) { s += x[k]; fs += fx[k]; }
  if (s == 12345u || fs == 1.2345f) out[threadIdx.x] = s;
}

// K5: interleaved 2 FFMA : 1 LOP3
__global__ void k_mix21(uint32_t *out, uint32_t m0, uint32_t m1, float m, float a)
{
  uint32_t x[8], y[8], z[8];
  float fx[16], fy[16], fz[16];
  for(int k = 0; k < 8; ++k) { x[k] = threadIdx.x * 7u + k; y[k] = m0 + k * 3u + threadIdx.x; z[k] = m1 ^ (k * 5u + threadIdx.x); }
  for(int k = 0; k < 16; ++k) { fx[k] = threadIdx.x * 0.001f + k; fy[k] = m + k * 1e-6f; fz[k] = a + k * 1e-6f + threadIdx.x * 1e-7f; }
  for(int i = 0; i < ITERS; ++i)
  {
    #pragma unroll
    for(int k = 0; k < 8; ++k) { x[k] = (x[k] & y[k]) ^ z[k]; fx[2*k] = fmaf(fx[2*k], fy[2*k], fz[2*k]); fx[2*k+1] = fmaf(fx[2*k+1], fy[2*k+1], fz[2*k+1]); }
  }
  uint32_t s = 0; float fs = 0; for(int k = 0; k < 8; ++k) s += x[k]; for(int k = 0; k < 16; ++k) fs += fx[k];
  if (s == 12345u || fs == 1.2345f) out[threadIdx.x] = s;
}

// K6: SHF (funnel rotate) chains
__global__ void k_shf(uint32_t *out, uint32_t m0)
{
  uint32_t x[16];
  for(int k = 0; k < 16; ++k) x[k] = threadIdx.x * 7u + k + m0;
  for(int i = 0; i < ITERS; ++i)
  {
    #pragma unroll
    for(int k = 0; k < 16; ++k) x[k] = __funnelshift_l(x[k], x[k], 18) + 0u;
  }
  uint32_t s = 0; for(int k = 0; k < 16; ++k) s ^= x[k];
  if (s == 12345u) out[threadIdx.x] = s;
}

template<typename F> float time_ms(F launch)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  float best = 1e9f;
  for(int r = 0; r < 3; ++r) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}

int main()
{
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount, blocks = sms * 8, threads = 256;
  void *buf; cudaMalloc(&buf, 1 << 20);
  double warps = (double)blocks * threads / 32;
  auto report = [&](const char *name, float ms, double inst_per_thread) {
    double warp_inst = warps * inst_per_thread;
    double clk = 1.9e9; // nominal; ratios are what matter
    printf("%-28s %8.3f ms  %6.2f warp-inst/clk/SM (at 1.9 GHz)\n", name, ms, warp_inst / (ms * 1e-3) / clk / sms);
  };
  report("FFMA const operands", time_ms([&]{ k_ffma_const<<<blocks, threads>>>((float*)buf, 0.999f, 1e-4f); }), 16.0 * ITERS);
  report("FFMA 3 distinct regs", time_ms([&]{ k_ffma_3reg<<<blocks, threads>>>((float*)buf, 0.999f, 1e-4f); }), 16.0 * ITERS);
  report("LOP3 3 regs", time_ms([&]{ k_lop3_3reg<<<blocks, threads>>>((uint32_t*)buf, 0xfff0ff0fu, 0x01010101u); }), 16.0 * ITERS);
  report("LOP3 imm (2 ops/iter)", time_ms([&]{ k_lop3_imm<<<blocks, threads>>>((uint32_t*)buf, 3u); }), 32.0 * ITERS);
  report("mix 1 LOP3 : 1 FFMA", time_ms([&]{ k_mix<<<blocks, threads>>>((uint32_t*)buf, 0xfff0ff0fu, 0x01010101u, 0.999f, 1e-4f); }), 16.0 * ITERS);
  report("mix 1 LOP3 : 2 FFMA", time_ms([&]{ k_mix21<<<blocks, threads>>>((uint32_t*)buf, 0xfff0ff0fu, 0x01010101u, 0.999f, 1e-4f); }), 24.0 * ITERS);
  report("SHF rotate", time_ms([&]{ k_shf<<<blocks, threads>>>((uint32_t*)buf, 3u); }), 16.0 * ITERS);
  return 0;
}
