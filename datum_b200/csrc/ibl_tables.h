// datum_b200 — per-level GGX sample tables (host side).
//
// In tools/ibl.cpp:170-184 every texel recomputes, for each of the N Hammersley
// points, quantities that depend only on (i, roughness): the half vector in the
// tangent frame (ibl.cpp:119-121,127) and, because V == N (ibl.cpp:164-165), the
// reflected direction and its weight:
//     H = (sin t cos p, sin t sin p, cos t)        in (T, B, N)
//     L = 2 (V.H) H - V = (2 hz hx, 2 hz hy, 2 hz^2 - 1)
//     NdotL = 2 hz^2 - 1                            (ibl.cpp:176)
// so the accept test (ibl.cpp:178), the weight and the total weight (ibl.cpp:182)
// are per-level constants.  The table keeps only accepted samples, sorted by
// decreasing NdotL (increasing lobe angle): the kernel then knows, per tile,
// how many leading samples cannot leave the tile's own cube face.
#pragma once

#include <cstdint>
#include <vector>

namespace ibl
{
  struct SampleEntry
  {
    float lx, ly, lz; // reflected direction in the texel's (T, B, N) frame
    float wh;         // 0.5 * NdotL
  };

  struct LevelSamples
  {
    float roughness = 0;
    int accepted = 0;                  // entries with NdotL > 0
    double total_weight = 0;           // sum of NdotL over accepted samples
    std::vector<SampleEntry> entries;  // the accepted entries, sorted by decreasing lz
  };

  // level in [1, levels), samples >= 1
  LevelSamples build_level_samples(int level, int levels, int samples);

  // The same entries cut into bands of `band` consecutive entries of the lz order (the last band
  // may be short) and, inside each band, ordered by the angle atan2(ly, lx): a warp that walks a
  // band walks along a ring of the lobe, so consecutive samples fetch neighbouring footprints.
  // band_min_lz[k] = smallest lz of band k (decreasing in k): the kernel's same-face test works
  // on whole bands.
  struct BandedSamples
  {
    LevelSamples level;
    int band = 0;
    std::vector<float> band_min_lz;
  };

  // order 0 = rings (every band one ring, walked by angle); 1 = patches (rings of s*band entries cut into s
  // sectors of `band` entries each, roughly as wide as deep, bands ordered by their smallest lz)
  BandedSamples build_banded_samples(int level, int levels, int samples, int band, int order = 0);

  // The banded entries for the kernel that works on two samples at a time (prefilter_dn.cu,
  // prefilter_dp_kernel): every entry multiplied by `scale`, the last band filled up with
  // samples of no effect (direction = the normal, weight 2^-60 of a real one: it rounds away in
  // every fp32 sum), and consecutive entries (a, b) interleaved as
  //     { lx_a, lx_b, ly_a, ly_b }  { lz_a, lz_b, wh_a, wh_b }
  // so that one 16-byte load yields two register pairs for the packed fp32x2 arithmetic.
  // Returns 4 floats per entry, band * ceil(count / band) entries.  (Round 2's pair kernel read this table; it
  // is kept for the A/B variants of the tools build.  The shipped kernel reads build_sector_entries' below.)
  std::vector<float> build_paired_entries(BandedSamples const &banded, float scale);

  // The pair kernel's table with one AZIMUTH SECTOR per warp (ibl_math.cuh, sector_rho_limits): the accepted
  // samples are cut into `sectors` (4 or 8) equal azimuth sectors of the tangent plane, each sorted by
  // decreasing lz; band k holds entries [k*per, (k+1)*per) of every sector, sector w at position w*per of
  // the band (per = band / sectors: what warp w of a tile reads), ordered by azimuth inside.  Sectors that
  // run out are filled up with samples of no effect (direction = the normal, weight 2^-60 of a real one).
  // Entries are projective, (lx/lz, ly/lz, lz*scale, wh*scale), pair-interleaved like build_paired_entries.
  // rho_max[w*bands + k] = largest |(lx/lz, ly/lz)| of sector w's share of band k (rounded up; 0 for a
  // filled-up share), increasing in k: the kernel counts a warp's same-face bands against it.
  struct SectorTable
  {
    int sectors = 0, band = 0, bands = 0;
    std::vector<float> entries;    // 4 floats per entry, band * bands entries
    std::vector<float> rho_max;    // sectors * bands
  };

  SectorTable build_sector_entries(LevelSamples const &level, int sectors, int band, float scale);

  // ibl.cpp:95-104
  float radicalinverse_VdC(uint32_t bits);
}
