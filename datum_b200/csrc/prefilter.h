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

  // ---- half-record kernel (prefilter_f16.cu) ----

  constexpr int kSampleBand = 32;   // entries per band of the banded sample table

  struct HalfGeom
  {
    int ws, hs;            // source level size
    float hw, hh;          // 0.5*(ws-1), 0.5*(hs-1): align-corners scale of ibl.cpp:37-38
    float hwm, hhm;        // hw - 0.5, hh - 0.5
    float inv_hw, inv_hh;
    int pw;                // records per x-parity half of a row: (ws+1)/2
    int row_stride;        // 2*pw
    uint32_t face_size;    // row_stride*hs records per face
    uint32_t bias;         // what the magic-add words contribute to the raw index
  };

  HalfGeom make_half_geom(int ws, int hs);
  size_t half_record_count(int ws, int hs);

  struct PrefilterHalfParams
  {
    uint4 const *recA;       // (r,g) halves of the four taps of every footprint of the SOURCE level
    uint2 const *recB;       // b halves of the four taps
    float4 const *table;     // banded sample table of this level (ibl_tables.h): (lx, ly, lz, 0.5*lz)
    float const *band_min_lz; // smallest lz of each band, decreasing
    int table_count;
    int bands;               // ceil(table_count / kSampleBand)
    uint32_t *dst_words;     // destination level base, rgbe words (may be null)
    float *dst_f32;          // destination level base, fp32 rgb triples before quantisation (may be null)
    int wd, hd;              // destination level size
    int row_begin, row_end;  // slab of the 6*hd face-major rows to compute
    HalfGeom geom;
    Quatf quats[6];          // face rotations, tools/ibl.cpp:253-261
    float norm[3];           // per channel: sum -> radiance / total weight
    uint32_t exp_mul;        // 2^23 (a parameter on purpose, see accumulate<3>)
    int *counters;           // queues+1 tile queue heads, zeroed by launch_build_half_records
    int blocks_x, tiles;     // filled by the launcher: 4x4-blocked tile numbering
    int queues, chunk, queued;
  };

  cudaError_t launch_prefilter_half(PrefilterHalfParams const &p, int variant, int sm_count, cudaStream_t stream, int *launched_grid);

  // variants >= 50: "denormal mantissa" quad records (16 bytes, row-major), see prefilter_f16.cu
  HalfGeom make_dn_geom(int ws, int hs);
  cudaError_t launch_build_dn_records(uint32_t const *src, uint4 *rec, int ws, int hs, int *counters, int ncounters, int sm_count, cudaStream_t stream);
  constexpr float kDnTableScale = 18446744073709551616.0f;   // 2^64, applied to every entry of the sample table

  cudaError_t launch_build_half_records(uint32_t const *src, uint4 *recA, uint2 *recB, int ws, int hs, int *counters, int ncounters, int sm_count, cudaStream_t stream);

  // variant 0 = pick by slab size; 1..15 = fixed <tile width, texels per lane, warps per tile>
  cudaError_t launch_prefilter_level(PrefilterParams const &p, int variant, int sm_count, cudaStream_t stream, int *launched_grid);

  cudaError_t launch_build_quad_records(uint32_t const *src, uint4 *records, int ws, int hs, int sm_count, cudaStream_t stream);
}
