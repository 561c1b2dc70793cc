// Micro-benchmark of the footprint gather of the prefilter loop under different record layouts
// (sm_100a).  Each warp owns an 8x4 tile of output texels; for sample s every lane fetches the
// 2x2 source footprint at (2*x + ox(s), 2*y + oy(s)) — the level-1 pattern: adjacent output texels
// are two source texels apart, the per-sample offset is shared by the tile.  Offsets follow a
// GGX-like radial distribution (most within a few dozen texels, a long tail).
// Only the memory side is exercised: fetched words are xor-ed together.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define WS 512          // source face size
#define NS 1024         // samples

struct Params { const void *a; const void *b; const short2 *offs; uint32_t *out; int *counters; int tiles_x, tiles; int banded; };

// LAYOUT 0: quad records 16 B row-major            (1 LDG.128)
// LAYOUT 1: quad records 16 B in four parity planes (1 LDG.128)
// LAYOUT 2: 24 B split (16 B + 8 B arrays) row-major
// LAYOUT 3: 24 B split, parity planes
// LAYOUT 4: 8 B texels row-major, 4 LDG.64
// LAYOUT 5: 4 B texels (raw rgbe words) row-major, 4 LDG.32
// LAYOUT 6: 4 B texels, 2 LDG.64 (pairs (i,i+1) replicated: 8 B per texel position)
template<int LAYOUT>
__device__ __forceinline__ uint32_t fetch(Params const &p, int i, int j)
{
  if (LAYOUT == 0) { uint4 r = __ldg((const uint4*)p.a + (size_t)j * WS + i); return r.x ^ r.y ^ r.z ^ r.w; }
  if (LAYOUT == 1)
  {
    size_t plane = (size_t)((j & 1) * 2 + (i & 1)) * (WS / 2) * (WS / 2);
    uint4 r = __ldg((const uint4*)p.a + plane + (size_t)(j >> 1) * (WS / 2) + (i >> 1));
    return r.x ^ r.y ^ r.z ^ r.w;
  }
  if (LAYOUT == 2)
  {
    size_t idx = (size_t)j * WS + i;
    uint4 r = __ldg((const uint4*)p.a + idx); uint2 q = __ldg((const uint2*)p.b + idx);
    return r.x ^ r.y ^ r.z ^ r.w ^ q.x ^ q.y;
  }
  if (LAYOUT == 3)
  {
    size_t idx = (size_t)((j & 1) * 2 + (i & 1)) * (WS / 2) * (WS / 2) + (size_t)(j >> 1) * (WS / 2) + (i >> 1);
    uint4 r = __ldg((const uint4*)p.a + idx); uint2 q = __ldg((const uint2*)p.b + idx);
    return r.x ^ r.y ^ r.z ^ r.w ^ q.x ^ q.y;
  }
  if (LAYOUT == 4)
  {
    const uint2 *t = (const uint2*)p.a + (size_t)j * WS + i;
    uint2 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + WS), d = __ldg(t + WS + 1);
    return a.x ^ a.y ^ b.x ^ b.y ^ c.x ^ c.y ^ d.x ^ d.y;
  }
  if (LAYOUT == 5)
  {
    const uint32_t *t = (const uint32_t*)p.a + (size_t)j * WS + i;
    return __ldg(t) ^ __ldg(t + 1) ^ __ldg(t + WS) ^ __ldg(t + WS + 1);
  }
  if (LAYOUT == 7)
  {
    // column-pair records {t(i,j), t(i,j+1)} of fp16 rgb (16 B): footprint = records i and i+1 of row j
    const uint4 *t = (const uint4*)p.a + (size_t)j * WS + i;
    uint4 a = __ldg(t), c = __ldg(t + 1);
    return a.x ^ a.y ^ a.z ^ c.x ^ c.y ^ c.z;
  }
  if (LAYOUT == 8)
  {
    // row-pair records {t(i,j), t(i+1,j)} of fp16 rgb (16 B): footprint = record i of rows j and j+1
    const uint4 *t = (const uint4*)p.a + (size_t)j * WS + i;
    uint4 a = __ldg(t), c = __ldg(t + WS);
    return a.x ^ a.y ^ a.z ^ c.x ^ c.y ^ c.z;
  }
  if (LAYOUT == 6)
  {
    const uint2 *t = (const uint2*)p.a + (size_t)j * WS + i;
    uint2 a = __ldg(t), c = __ldg(t + WS);
    return a.x ^ a.y ^ c.x ^ c.y;
  }
  return 0;
}

__device__ __forceinline__ uint32_t morton_x(uint32_t m) { m &= 0x55555555u; m = (m | (m >> 1)) & 0x33333333u; m = (m | (m >> 2)) & 0x0F0F0F0Fu; m = (m | (m >> 4)) & 0x00FF00FFu; m = (m | (m >> 8)) & 0xFFFFu; return m; }

template<int LAYOUT, int SCHED>
__global__ void __launch_bounds__(128) k_gather(Params p)
{
  __shared__ short2 s_off[NS];
  for(int i = threadIdx.x; i < NS; i += blockDim.x) s_off[i] = p.offs[i];
  __syncthreads();
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t acc = 0;
  __shared__ int s_tile;
  uint32_t smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
  int nsm = gridDim.x / 9;
  int chunk = (p.tiles + nsm - 1) / nsm;
  for(int it = 0; ; ++it)
  {
    int tile;
    if (SCHED == 0) { tile = blockIdx.x + it * gridDim.x; if (tile >= p.tiles) break; }
    else
    {
      __syncthreads();
      if (threadIdx.x == 0)
      {
        int k = atomicAdd(&p.counters[smid], 1);
        s_tile = k < chunk ? (int)smid * chunk + k : -1;
      }
      __syncthreads();
      tile = s_tile;
      if (tile < 0 || tile >= p.tiles) break;
    }
    int tx, ty;
    if (SCHED == 0) { tx = tile % p.tiles_x; ty = (tile / p.tiles_x) % 64; }
    else { int f = tile >> 11, m = tile & 2047; tx = morton_x(m); ty = morton_x(m >> 1); (void)f; }
    int x = tx * 8 + (lane & 7), y = ty * 4 + (lane >> 3);
    #pragma unroll 2
    for(int q = 0; q < NS / 4; ++q)
    {
      int s = p.banded ? ((q >> 3) * 32 + warp * 8 + (q & 7)) : (warp + 4 * q);
      short2 o = s_off[s];
      int i = 2 * x + o.x, j = 2 * y + o.y;
      i = min(max(i, 0), WS - 2); j = min(max(j, 0), WS - 2);
      acc ^= fetch<LAYOUT>(p, i, j);
    }
  }
  if (acc == 0x12345u) p.out[threadIdx.x] = acc;
}

int main()
{
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double clk = khz * 1e3;

  // GGX level-1-like offsets: alpha = (1/7)^2, Hammersley, L = reflect -> angle 2*theta_h, texels = angle / (90deg/512)
  std::vector<short2> offs(NS);
  double alpha = 1.0 / 49.0;
  for(int i = 0; i < NS; ++i)
  {
    uint32_t b = i; b = (b << 16) | (b >> 16); b = ((b & 0x55555555u) << 1) | ((b & 0xAAAAAAAAu) >> 1); b = ((b & 0x33333333u) << 2) | ((b & 0xCCCCCCCCu) >> 2);
    b = ((b & 0x0F0F0F0Fu) << 4) | ((b & 0xF0F0F0F0u) >> 4); b = ((b & 0x00FF00FFu) << 8) | ((b & 0xFF00FF00u) >> 8);
    double u1 = (double)i / NS, u2 = b * 2.3283064365386963e-10;
    double ct = sqrt((1 - u2) / (1 + (alpha * alpha - 1) * u2)); double th = 2 * acos(ct);
    double r = tan(th < 1.4 ? th : 1.4) * 256.0;   // texels on the face plane
    double ph = 2 * M_PI * u1;
    offs[i] = make_short2((short)lrint(r * cos(ph)), (short)lrint(r * sin(ph)));
  }
  // the kernel's table is sorted by lobe angle: neighbouring samples have similar radius
  std::sort(offs.begin(), offs.end(), [](short2 a, short2 b){ return a.x * a.x + a.y * a.y < b.x * b.x + b.y * b.y; });

  short2 *d_off; cudaMalloc(&d_off, NS * sizeof(short2)); cudaMemcpy(d_off, offs.data(), NS * sizeof(short2), cudaMemcpyHostToDevice);
  void *a, *b; cudaMalloc(&a, (size_t)WS * WS * 16 + 4096); cudaMalloc(&b, (size_t)WS * WS * 8 + 4096);
  cudaMemset(a, 1, (size_t)WS * WS * 16); cudaMemset(b, 2, (size_t)WS * WS * 8);
  uint32_t *out; cudaMalloc(&out, 4096);
  int *counters; cudaMalloc(&counters, 4096);

  Params p; p.a = a; p.b = b; p.offs = d_off; p.out = out; p.counters = counters; p.tiles_x = 256 / 8; p.tiles = (256 / 8) * (256 / 4);
  p.tiles = p.tiles;   // one face worth of tiles: 2048 tiles -> ~1.5 per CTA at 9 CTAs/SM; repeat faces
  const int faces = 6;
  p.tiles *= faces;    // addresses repeat per face (same source) - fine for the memory pattern
  p.tiles_x = 32;

  // ring order inside bands of 32 (by angle), bands by radius
  std::vector<short2> ring = offs;
  for(int b = 0; b < NS; b += 32)
    std::sort(ring.begin() + b, ring.begin() + b + 32, [](short2 a, short2 c){ return atan2((double)a.y, (double)a.x) < atan2((double)c.y, (double)c.x); });
  short2 *d_ring; cudaMalloc(&d_ring, NS * sizeof(short2)); cudaMemcpy(d_ring, ring.data(), NS * sizeof(short2), cudaMemcpyHostToDevice);

  auto run = [&](const char *name, auto kernel) {
    int grid = sms * 9;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // tiles beyond the first face wrap in y: clamp handles it
    cudaMemset(counters, 0, 4096); kernel<<<grid, 128>>>(p); cudaDeviceSynchronize();
    float best = 1e9f;
    for(int r = 0; r < 3; ++r) { cudaMemset(counters, 0, 4096); cudaEventRecord(e0); kernel<<<grid, 128>>>(p); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double warp_samples = (double)p.tiles * NS;
    printf("%-46s %8.3f ms  %6.1f clk/warp-sample/SMSP  (%s)\n", name, best, best * 1e-3 * clk * sms * 4 / warp_samples, cudaGetErrorString(cudaGetLastError()));
  };
  for(int banded = 0; banded < 2; ++banded) {
  p.banded = banded; p.offs = banded ? d_ring : d_off;
  printf("---- sample order: %s\n", banded ? "ring-ordered bands of 32, 8 consecutive per warp" : "sorted by radius, warps interleaved");
  run("0/s0: quad 16 B row-major (current)", k_gather<0, 0>);
  run("1/s0: quad 16 B parity planes", k_gather<1, 0>);
  run("2/s0: 24 B split row-major", k_gather<2, 0>);
  run("3/s0: 24 B split parity planes", k_gather<3, 0>);
  run("6/s0: 8 B pair records, 2 x LDG.64", k_gather<6, 0>);
  run("7/s0: col-pair 16 B, 2 adjacent LDG.128", k_gather<7, 0>);
  run("8/s0: row-pair 16 B, 2 LDG.128 rows j,j+1", k_gather<8, 0>);
  run("0/s1: quad 16 B row-major (current)", k_gather<0, 1>);
  run("1/s1: quad 16 B parity planes", k_gather<1, 1>);
  run("2/s1: 24 B split row-major", k_gather<2, 1>);
  run("3/s1: 24 B split parity planes", k_gather<3, 1>);
  run("6/s1: 8 B pair records, 2 x LDG.64", k_gather<6, 1>);
  run("7/s1: col-pair 16 B, 2 adjacent LDG.128", k_gather<7, 1>);
  run("8/s1: row-pair 16 B, 2 LDG.128 rows j,j+1", k_gather<8, 1>);
  }
  return 0;
}
