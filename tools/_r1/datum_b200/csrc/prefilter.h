// datum_b200 — launch interface of the prefilter kernels (internal to libdatum_ibl_cuda).
#pragma once

#include "ibl_math.cuh"

#include <cuda_runtime.h>
#include <cstdint>

namespace ibl
{
  struct PrefilterParams
  {
    uint4 const *records;    // quad records of the SOURCE level (6*ws*hs)
    float4 const *table;     // sample table of this level: (lx, ly, lz, 0.5*lz), decreasing lz
    int table_count;
    uint32_t *dst_words;     // destination level base, rgbe words (may be null)
    float *dst_f32;          // destination level base, fp32 rgb triples before quantisation (may be null)
    int wd, hd;              // destination level size
    int row_begin, row_end;  // slab of the 6*hd face-major rows to compute
    LevelGeom geom;          // source level addressing constants
    Quatf quats[6];          // face rotations, tools/ibl.cpp:253-261
    DecodeMasks masks;       // bit masks of accumulate_tap, passed as parameters so they live in registers
    float norm;              // kAccScale / total weight
    int tiles_x, tiles;      // filled by the launcher
  };

  // ---- denormal-mantissa kernel (prefilter_dn.cu): every level at least 8 texels wide ----

  constexpr int kMaxPeers = 7;      // one probe split over at most 8 GPUs (one NVSwitch domain)
  constexpr int kSampleBand = 16;   // entries per band of the banded sample table (ibl_tables.h)

  struct PrefilterDnParams
  {
    uint4 const *records;     // quad records of the SOURCE level, words re-laid by pack_dn_word (6*ws*hs)
    float4 const *table;      // banded sample table of this level, every entry scaled by kDnTableScale
    float4 const *table_pairs; // the same entries, last band filled up, two entries interleaved per 32 bytes (ibl_tables.h)
    float const *band_min_lz; // smallest lz of each band (unscaled), decreasing
    int table_count;
    int bands;                // ceil(table_count / kSampleBand)
    uint32_t *dst_words;      // destination level base, rgbe words (may be null)
    float *dst_f32;           // destination level base, fp32 rgb triples before quantisation (may be null)
    uint32_t *peer_words[kMaxPeers]; // the same destination level in the chains of other GPUs (NVLink peer stores)
    int peers;                // how many of them: the epilogue writes every word to dst_words and to each peer
    int wd, hd;               // destination level size
    int row_begin, row_end;   // slab of the 6*hd face-major rows to compute
    LevelGeom geom;           // source level addressing constants
    Quatf quats[6];           // face rotations, tools/ibl.cpp:253-261
    float norm[3];            // per channel: sum -> radiance / total weight (dn_channel_norms)
    uint32_t exp_mul;         // 2^23 (a parameter on purpose, see scale_by_exponent)
    int *counters;            // queues+1 tile queue heads, zeroed by launch_build_dn_records
    int blocks_x, tiles;      // filled by the launcher: 4x4-blocked tile numbering
    int queues, chunk, queued;
  };

  // ---- tail levels (prefilter_dn.cu, prefilter_tail_kernel): a few hundred texels ----
  //
  // One CTA per output texel, LANES are samples: the 1024 samples of a texel are one to four steps
  // deep instead of 32, the four footprint words come straight from the source level (no record pass).
  struct PrefilterTailParams
  {
    uint32_t const *src;      // SOURCE level words (6*ws*hs), tools/ibl.cpp layout
    float4 const *table;      // banded sample table of this level scaled by kDnTableScale (any order works)
    int table_count;
    uint32_t *dst_words;
    float *dst_f32;
    uint32_t *peer_words[kMaxPeers];
    int peers;
    int wd, hd;
    int row_begin, row_end;
    LevelGeom geom;
    Quatf quats[6];
    float norm[3];
    uint32_t exp_mul;
  };

  // slabs up to this many texels go to the tail kernel
  constexpr int kTailTexels = 6144;

  cudaError_t launch_prefilter_tail(PrefilterTailParams const &p, int sm_count, cudaStream_t stream);

  // variant 0 = pick by slab size and table size; 50..58 = one sample at a time, fixed <warps per tile,
  // table in shared memory, tile queues>; 70..75 = two samples at a time (prefilter_dp_kernel)
  cudaError_t launch_prefilter_dn(PrefilterDnParams const &p, int variant, int sm_count, cudaStream_t stream, int *launched_grid);

  // Barrier between the GPUs that share one probe, on the stream: rank `rank` publishes `epoch` into
  // slot [rank] of every peer's flag array (flags[r] = rank r's array of `world` words) and waits until
  // its own array holds `epoch` in every slot.  Traps (loudly failing the context) after ~10 s.
  struct PeerFlags
  {
    uint32_t *ptr[kMaxPeers + 1];   // by rank, own array included (host array of device pointers, passed by value)
  };

  cudaError_t launch_peer_barrier(PeerFlags const &flags, int rank, int world, uint32_t epoch, cudaStream_t stream);

  // also zeroes the `ncounters` tile queue heads for the prefilter launch that follows
  cudaError_t launch_build_dn_records(uint32_t const *src, uint4 *rec, int ws, int hs, int *counters, int ncounters, int sm_count, cudaStream_t stream);

  // variant 0 = pick by slab size; 1..15 = fixed <tile width, texels per lane, warps per tile>
  cudaError_t launch_prefilter_level(PrefilterParams const &p, int variant, int sm_count, cudaStream_t stream, int *launched_grid);

  cudaError_t launch_build_quad_records(uint32_t const *src, uint4 *records, int ws, int hs, int sm_count, cudaStream_t stream);
}
