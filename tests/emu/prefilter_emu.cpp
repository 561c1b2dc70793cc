// TEST INFRASTRUCTURE — host build of the kernel's per-sample math.
//
// Compiles datum_b200/csrc/ibl_math.cuh and ibl_tables.cpp with g++ and runs
// the same per-texel loops the CUDA prefilter kernels run (table-driven reflected
// direction, magic-add floor, quad-record addressing; biased-mantissa accumulation
// for prefilter.cu, subnormal-mantissa accumulation over the banded table for
// prefilter_dn.cu), single-threaded and in table order.  tests/test_kernel_math_cpu.py
// compares it with the oracle so that algorithmic mistakes are caught on the
// CPU-only CI leg.  It is NOT a fallback: nothing under datum_b200/ builds,
// links or loads it.

#include "../../datum_b200/csrc/ibl_math.cuh"
#include "../../datum_b200/csrc/ibl_tables.h"

#include <cmath>
#include <cstdlib>
#include <vector>

using namespace ibl;

static double g_fast_fraction = 0;

extern "C" void emu_prefilter_level(uint32_t const *src, int ws, int hs, int level, int levels, int samples, uint32_t *words, float *f32)
{
  LevelSamples table = build_level_samples(level, levels, samples);
  LevelGeom geom = make_level_geom(ws, hs);

  const float kPi = 3.14159265358979323846f;
  float angles[6] = { -kPi/2, kPi/2, -kPi/2, kPi/2, 0.0f, kPi };
  int axes[6] = { 1, 1, 0, 0, 1, 1 };
  Quatf quats[6];
  for(int f = 0; f < 6; ++f)
  {
    float c = std::cos(angles[f]/2), s = std::sin(angles[f]/2);
    quats[f] = Quatf{ c, axes[f] == 0 ? s : 0.0f, axes[f] == 1 ? s : 0.0f, 0.0f };
  }

  int wd = ws >> 1, hd = hs >> 1;
  float norm = (float)((double)kAccScale / table.total_weight);
  DecodeMasks masks = make_decode_masks();

  // quad records exactly as build_quad_records_kernel lays them out
  struct Rec { uint32_t x, y, z, w; };
  std::vector<Rec> records((size_t)6 * ws * hs);
  for(size_t idx = 0; idx < records.size(); ++idx)
  {
    int i = (int)(idx % ws), j = (int)((idx / ws) % hs);
    size_t right = (i + 1 < ws) ? 1 : 0, down = (j + 1 < hs) ? (size_t)ws : 0;
    records[idx] = Rec{ pack_record_word(src[idx]), pack_record_word(src[idx + right]), pack_record_word(src[idx + down]), pack_record_word(src[idx + down + right]) };
  }

  long fast = 0, total = 0;

  for(int face = 0; face < 6; ++face)
  {
    for(int y = 0; y < hd; ++y)
    {
      for(int x = 0; x < wd; ++x)
      {
        Vec3f N = texel_normal(quats[face], x, y, wd, hd);
        Vec3f T, B;
        tangent_frame(N, T, B);

        Vec3f Tl = to_face_local(face, T), Bl = to_face_local(face, B), Nl = to_face_local(face, N);
        float threshold = getenv("EMU_NOFAST") ? 2.0f : same_face_threshold(Nl);

        Vec3f Ts = { Tl.x * geom.hw, Tl.y * geom.hh, Tl.z };
        Vec3f Bs = { Bl.x * geom.hw, Bl.y * geom.hh, Bl.z };
        Vec3f Ns = { Nl.x * geom.hw, Nl.y * geom.hh, Nl.z };
        uint32_t face_base = (uint32_t)face * geom.face_size - geom.bias;

        // back to world exactly as the kernel does between its two loops
        Vec3f Tw = from_face_local(face, Vec3f{ Ts.x * geom.inv_hw, Ts.y * geom.inv_hh, Ts.z });
        Vec3f Bw = from_face_local(face, Vec3f{ Bs.x * geom.inv_hw, Bs.y * geom.inv_hh, Bs.z });
        Vec3f Nw = from_face_local(face, Vec3f{ Ns.x * geom.inv_hw, Ns.y * geom.inv_hh, Ns.z });

        float acc[4] = { 0, 0, 0, 0 };

        for(int i = 0; i < table.accepted; ++i)
        {
          SampleEntry const &e = table.entries[i];

          float du, dv;
          uint32_t idx;

          if (e.lz > threshold)
          {
            float la = fmaf(e.lz, Ns.x, fmaf(e.ly, Bs.x, e.lx * Ts.x));
            float lb = fmaf(e.lz, Ns.y, fmaf(e.ly, Bs.y, e.lx * Ts.y));
            float lm = fmaf(e.lz, Ns.z, fmaf(e.ly, Bs.z, e.lx * Ts.z));
            idx = face_footprint(geom, face_base, la, lb, lm, du, dv);
            ++fast;
          }
          else
          {
            float Lx = fmaf(e.lz, Nw.x, fmaf(e.ly, Bw.x, e.lx * Tw.x));
            float Ly = fmaf(e.lz, Nw.y, fmaf(e.ly, Bw.y, e.lx * Tw.y));
            float Lz = fmaf(e.lz, Nw.z, fmaf(e.ly, Bw.z, e.lx * Tw.z));
            idx = cube_footprint(geom, Lx, Ly, Lz, du, dv);
          }
          ++total;

          if (idx >= records.size())
            __builtin_trap();

          float w[4];
          footprint_weights(du, dv, e.wh, e.lz, w);

          Rec const &rec = records[idx];
          accumulate_tap(masks, rec.x, w[0], acc);
          accumulate_tap(masks, rec.y, w[1], acc);
          accumulate_tap(masks, rec.z, w[2], acc);
          accumulate_tap(masks, rec.w, w[3], acc);
        }

        float r = (acc[0] - acc[3]) * norm;
        float g = (acc[1] - acc[3]) * norm;
        float b = (acc[2] - acc[3]) * norm;

        size_t o = ((size_t)face * hd + y) * wd + x;
        if (words)
          words[o] = rgbe_encode(r, g, b);
        if (f32)
        {
          f32[3*o + 0] = r; f32[3*o + 1] = g; f32[3*o + 2] = b;
        }
      }
    }
  }

  g_fast_fraction = total ? (double)fast / (double)total : 0.0;
}

// The same level through the math of prefilter_dn.cu: banded ring-ordered table scaled by 2^64,
// per-band same-face decision, quad records re-laid by pack_dn_word, subnormal-mantissa taps with
// the exponent folded into the weight, per-channel normalisation.
//
// pairs = true: the arithmetic of prefilter_dp_kernel instead — the pair-interleaved PROJECTIVE table with
// one azimuth sector per warp (build_sector_entries), the same-face decision from the sector's limit
// (sector_rho_limits), folded frame rows, the record index formed in the fp32
// adder and taken relative to a record pointer moved back by its bias, right-hand weights by difference
// (ibl_math.cuh "projective form"), blue summed in one partial per sample of a pair.
static void emu_dn_impl(uint32_t const *src, int ws, int hs, int level, int levels, int samples, int band, uint32_t *words, float *f32, bool pairs)
{
  BandedSamples banded = build_banded_samples(level, levels, samples, band);
  std::vector<SampleEntry> table = banded.level.entries;
  for(auto &e : table)
  {
    e.lx *= kDnTableScale; e.ly *= kDnTableScale; e.lz *= kDnTableScale; e.wh *= kDnTableScale;
  }

  SectorTable sectors;
  const int kSectors = 4;     // the kernel's shape for big levels: 4 warps per tile, one 90-degree sector each
  if (pairs)
  {
    if (!proj_usable(ws, hs))
      __builtin_trap();   // the launcher never picks the pair kernel here
    sectors = build_sector_entries(banded.level, kSectors, band, kDnTableScale);
    table.assign(sectors.entries.size() / 4, SampleEntry{});
    for(size_t i = 0; i < table.size(); i += 2)
    {
      float const *q = sectors.entries.data() + 4 * i;
      table[i] = SampleEntry{ q[0], q[2], q[4], q[6] };
      table[i + 1] = SampleEntry{ q[1], q[3], q[5], q[7] };
    }
    if (table.size() != (size_t)sectors.bands * (size_t)band)
      __builtin_trap();
  }
  LevelGeom geom = make_level_geom(ws, hs);

  const float kPi = 3.14159265358979323846f;
  float angles[6] = { -kPi/2, kPi/2, -kPi/2, kPi/2, 0.0f, kPi };
  int axes[6] = { 1, 1, 0, 0, 1, 1 };
  Quatf quats[6];
  for(int f = 0; f < 6; ++f)
  {
    float c = std::cos(angles[f]/2), s = std::sin(angles[f]/2);
    quats[f] = Quatf{ c, axes[f] == 0 ? s : 0.0f, axes[f] == 1 ? s : 0.0f, 0.0f };
  }

  int wd = ws >> 1, hd = hs >> 1;
  float norm[3];
  dn_channel_norms(banded.level.total_weight, norm);

  struct Rec { uint32_t x, y, z, w; };
  std::vector<Rec> records((size_t)6 * ws * hs);
  for(size_t idx = 0; idx < records.size(); ++idx)
  {
    int i = (int)(idx % ws), j = (int)((idx / ws) % hs);
    size_t right = (i + 1 < ws) ? 1 : 0, down = (j + 1 < hs) ? (size_t)ws : 0;
    records[idx] = Rec{ pack_dn_word(src[idx]), pack_dn_word(src[idx + right]), pack_dn_word(src[idx + down]), pack_dn_word(src[idx + down + right]) };
  }

  long fast = 0, total = 0;

  for(int face = 0; face < 6; ++face)
  {
    for(int y = 0; y < hd; ++y)
    {
      for(int x = 0; x < wd; ++x)
      {
        Vec3f N = texel_normal(quats[face], x, y, wd, hd);
        Vec3f T, B;
        tangent_frame(N, T, B);

        Vec3f Tl = to_face_local(face, T), Bl = to_face_local(face, B), Nl = to_face_local(face, N);
        float threshold = getenv("EMU_NOFAST") ? 2.0f : same_face_threshold(Nl);

        Vec3f Ts = { Tl.x * geom.hw, Tl.y * geom.hh, Tl.z };
        Vec3f Bs = { Bl.x * geom.hw, Bl.y * geom.hh, Bl.z };
        Vec3f Ns = { Nl.x * geom.hw, Nl.y * geom.hh, Nl.z };
        uint32_t face_base = (uint32_t)face * geom.face_size - geom.bias;

        Vec3f Tw = from_face_local(face, Vec3f{ Ts.x * geom.inv_hw, Ts.y * geom.inv_hh, Ts.z });
        Vec3f Bw = from_face_local(face, Vec3f{ Bs.x * geom.inv_hw, Bs.y * geom.inv_hh, Bs.z });
        Vec3f Nw = from_face_local(face, Vec3f{ Ns.x * geom.inv_hw, Ns.y * geom.inv_hh, Ns.z });

        if (pairs)
        {
          Ts = fold_face_row(geom, Tl); Bs = fold_face_row(geom, Bl); Ns = fold_face_row(geom, Nl);
          Tw = from_face_local(face, unfold_face_row(geom, Ts));
          Bw = from_face_local(face, unfold_face_row(geom, Bs));
          Nw = from_face_local(face, unfold_face_row(geom, Ns));
        }

        float acc[3] = { 0, 0, 0 };
        float blue[2] = { 0, 0 };

        // the pair kernel decides per warp (= azimuth sector) and tile; here per sector and texel
        float limits[kFrameSectors];
        sector_rho_limits(Tl, Bl, Nl, limits);

        for(size_t i = 0; i < table.size(); ++i)
        {
          SampleEntry const &e = table[i];
          bool same = !pairs && banded.band_min_lz[i / (size_t)band] > threshold;
          if (pairs && !getenv("EMU_NOFAST"))
          {
            int w = (int)((i % (size_t)band) / (size_t)(band / kSectors));
            float limit = 3.0e38f;
            for(int k = w * kFrameSectors / kSectors; k < (w + 1) * kFrameSectors / kSectors; ++k)
              limit = std::min(limit, limits[k]);
            same = sectors.rho_max[(size_t)w * sectors.bands + i / (size_t)band] <= limit;
          }

          float du, dv;
          uint32_t idx;

          if (pairs)
          {
            // e.lx, e.ly hold lx/lz, ly/lz here
            uint32_t raw, bias;
            if (same)
            {
              float la = fmaf(e.lx, Ts.x, fmaf(e.ly, Bs.x, Ns.x));
              float lb = fmaf(e.lx, Ts.y, fmaf(e.ly, Bs.y, Ns.y));
              float lm = fmaf(e.lx, Ts.z, fmaf(e.ly, Bs.z, Ns.z));
              raw = face_footprint_proj(geom, la, lb, lm, du, dv) + (uint32_t)face * geom.face_size;
              bias = kMagicBits;
              ++fast;
            }
            else
            {
              float Lx = fmaf(e.lx, Tw.x, fmaf(e.ly, Bw.x, Nw.x));
              float Ly = fmaf(e.lx, Tw.y, fmaf(e.ly, Bw.y, Nw.y));
              float Lz = fmaf(e.lx, Tw.z, fmaf(e.ly, Bw.z, Nw.z));
              uint32_t f;
              raw = cube_footprint_proj(geom, Lx, Ly, Lz, du, dv, f);
              raw += f * geom.face_size;
              bias = geom.bias_general;
            }
            // the kernel's address: unsigned 32-bit index that still carries the bias, pointer moved back by it
            long long element = (long long)raw - (long long)bias;
            if (element < 0 || element >= (long long)records.size())
              __builtin_trap();
            idx = (uint32_t)element;
          }
          else if (same)
          {
            float la = fmaf(e.lz, Ns.x, fmaf(e.ly, Bs.x, e.lx * Ts.x));
            float lb = fmaf(e.lz, Ns.y, fmaf(e.ly, Bs.y, e.lx * Ts.y));
            float lm = fmaf(e.lz, Ns.z, fmaf(e.ly, Bs.z, e.lx * Ts.z));
            idx = face_footprint(geom, face_base, la, lb, lm, du, dv);
            ++fast;
          }
          else
          {
            float Lx = fmaf(e.lz, Nw.x, fmaf(e.ly, Bw.x, e.lx * Tw.x));
            float Ly = fmaf(e.lz, Nw.y, fmaf(e.ly, Bw.y, e.lx * Tw.y));
            float Lz = fmaf(e.lz, Nw.z, fmaf(e.ly, Bw.z, e.lx * Tw.z));
            idx = cube_footprint(geom, Lx, Ly, Lz, du, dv);
          }
          ++total;

          if (idx >= records.size())
            __builtin_trap();

          float w[4];
          if (pairs)
            footprint_weights_diff(du, dv, e.wh, e.lz, w);
          else
            footprint_weights(du, dv, e.wh, e.lz, w);

          Rec const &rec = records[idx];
          if (pairs)
          {
            float one[3] = { acc[0], acc[1], blue[i & 1] };
            dn_accumulate_tap(rec.x, w[0], one);
            dn_accumulate_tap(rec.y, w[1], one);
            dn_accumulate_tap(rec.z, w[2], one);
            dn_accumulate_tap(rec.w, w[3], one);
            acc[0] = one[0]; acc[1] = one[1]; blue[i & 1] = one[2];
          }
          else
          {
            dn_accumulate_tap(rec.x, w[0], acc);
            dn_accumulate_tap(rec.y, w[1], acc);
            dn_accumulate_tap(rec.z, w[2], acc);
            dn_accumulate_tap(rec.w, w[3], acc);
          }
        }

        if (pairs)
          acc[2] = blue[0] + blue[1];

        float r = acc[0] * norm[0], g = acc[1] * norm[1], b = acc[2] * norm[2];

        size_t o = ((size_t)face * hd + y) * wd + x;
        if (words)
          words[o] = rgbe_encode(r, g, b);
        if (f32)
        {
          f32[3*o + 0] = r; f32[3*o + 1] = g; f32[3*o + 2] = b;
        }
      }
    }
  }

  g_fast_fraction = total ? (double)fast / (double)total : 0.0;
}

extern "C" void emu_prefilter_level_dn(uint32_t const *src, int ws, int hs, int level, int levels, int samples, int band, uint32_t *words, float *f32)
{
  emu_dn_impl(src, ws, hs, level, levels, samples, band, words, f32, false);
}

extern "C" void emu_prefilter_level_dp(uint32_t const *src, int ws, int hs, int level, int levels, int samples, int band, uint32_t *words, float *f32)
{
  emu_dn_impl(src, ws, hs, level, levels, samples, band, words, f32, true);
}

// pair-interleaved, filled-up table of the pair kernel; returns the entry count (a multiple of `band`)
extern "C" int emu_paired_table(int level, int levels, int samples, int band, float scale, float *out, int capacity)
{
  BandedSamples banded = build_banded_samples(level, levels, samples, band);
  std::vector<float> paired = build_paired_entries(banded, scale);
  if ((int)paired.size() > capacity)
    return -1;
  for(size_t i = 0; i < paired.size(); ++i)
    out[i] = paired[i];
  return (int)(paired.size() / 4);
}

// banded table as the dn kernel sees it (unscaled): entries, per-band minimum lz; returns the entry count
extern "C" int emu_banded_table(int level, int levels, int samples, int band, float *entries, float *band_min, int *bands)
{
  BandedSamples banded = build_banded_samples(level, levels, samples, band);
  for(size_t i = 0; i < banded.level.entries.size(); ++i)
  {
    entries[4*i + 0] = banded.level.entries[i].lx; entries[4*i + 1] = banded.level.entries[i].ly;
    entries[4*i + 2] = banded.level.entries[i].lz; entries[4*i + 3] = banded.level.entries[i].wh;
  }
  for(size_t k = 0; k < banded.band_min_lz.size(); ++k)
    band_min[k] = banded.band_min_lz[k];
  *bands = (int)banded.band_min_lz.size();
  return (int)banded.level.entries.size();
}

// the same in patch order (ibl_tables.h: order 1)
extern "C" int emu_patch_table(int level, int levels, int samples, int band, float *entries, float *band_min, int *bands)
{
  BandedSamples banded = build_banded_samples(level, levels, samples, band, 1);
  for(size_t i = 0; i < banded.level.entries.size(); ++i)
  {
    entries[4*i + 0] = banded.level.entries[i].lx; entries[4*i + 1] = banded.level.entries[i].ly;
    entries[4*i + 2] = banded.level.entries[i].lz; entries[4*i + 3] = banded.level.entries[i].wh;
  }
  for(size_t k = 0; k < banded.band_min_lz.size(); ++k)
    band_min[k] = banded.band_min_lz[k];
  *bands = (int)banded.band_min_lz.size();
  return (int)banded.level.entries.size();
}

extern "C" uint32_t emu_pack_dn_word(uint32_t w) { return pack_dn_word(w); }

// one tap through dn_accumulate_tap and the channel norms with total weight 1 and weight w: returns w * texel value
extern "C" void emu_dn_tap(uint32_t rgbe_word, float w, float *rgb)
{
  float acc[3] = { 0, 0, 0 }, norm[3];
  dn_accumulate_tap(pack_dn_word(rgbe_word), w * kDnTableScale, acc);
  dn_channel_norms(1.0, norm);
  rgb[0] = acc[0] * norm[0]; rgb[1] = acc[1] * norm[1]; rgb[2] = acc[2] * norm[2];
}

// the same tap through raw_accumulate_tap (the tail kernel: the word in the reference's own bit layout)
extern "C" void emu_raw_tap(uint32_t rgbe_word, float w, float *rgb)
{
  float acc[3] = { 0, 0, 0 }, norm[3];
  raw_accumulate_tap(rgbe_word, w * kDnTableScale, 0x00800000u, acc);
  raw_channel_norms(1.0, norm);
  for(int c = 0; c < 3; ++c)
    rgb[c] = acc[c] * norm[c];
}

extern "C" double emu_last_fast_fraction() { return g_fast_fraction; }

extern "C" uint32_t emu_rgbe_encode(float r, float g, float b) { return rgbe_encode(r, g, b); }
extern "C" void emu_rgbe_decode(uint32_t w, float *rgb) { rgbe_decode(w, rgb[0], rgb[1], rgb[2]); }
extern "C" void emu_texel_normal(int face, int x, int y, int wd, int hd, float *out)
{
  const float kPi = 3.14159265358979323846f;
  float angles[6] = { -kPi/2, kPi/2, -kPi/2, kPi/2, 0.0f, kPi };
  int axes[6] = { 1, 1, 0, 0, 1, 1 };
  float c = std::cos(angles[face]/2), s = std::sin(angles[face]/2);
  Quatf q = Quatf{ c, axes[face] == 0 ? s : 0.0f, axes[face] == 1 ? s : 0.0f, 0.0f };
  Vec3f n = texel_normal(q, x, y, wd, hd);
  out[0] = n.x; out[1] = n.y; out[2] = n.z;
}

// per-sample trace of one output texel (debug aid for the parity tests):
// out[i] = {face, i, j, du, dv, weight} for table entry i
extern "C" int emu_trace_texel(int ws, int hs, int level, int levels, int samples, int face, int x, int y, float *out, float *dirs)
{
  LevelSamples table = build_level_samples(level, levels, samples);
  LevelGeom geom = make_level_geom(ws, hs);
  const float kPi = 3.14159265358979323846f;
  float angles[6] = { -kPi/2, kPi/2, -kPi/2, kPi/2, 0.0f, kPi };
  int axes[6] = { 1, 1, 0, 0, 1, 1 };
  float c = std::cos(angles[face]/2), s = std::sin(angles[face]/2);
  Quatf q = Quatf{ c, axes[face] == 0 ? s : 0.0f, axes[face] == 1 ? s : 0.0f, 0.0f };
  int wd = ws >> 1, hd = hs >> 1;
  Vec3f N = texel_normal(q, x, y, wd, hd);
  Vec3f T, B;
  tangent_frame(N, T, B);
  for(int i = 0; i < table.accepted; ++i)
  {
    SampleEntry const &e = table.entries[i];
    float Lx = e.lx * T.x + e.ly * B.x + e.lz * N.x;
    float Ly = e.lx * T.y + e.ly * B.y + e.lz * N.y;
    float Lz = e.lx * T.z + e.ly * B.z + e.lz * N.z;
    float du, dv;
    uint32_t idx = cube_footprint(geom, Lx, Ly, Lz, du, dv);
    out[6*i + 0] = (float)(idx / geom.face_size);
    out[6*i + 1] = (float)((idx % geom.face_size) % ws);
    out[6*i + 2] = (float)((idx % geom.face_size) / ws);
    out[6*i + 3] = du + 0.5f;
    out[6*i + 4] = dv + 0.5f;
    out[6*i + 5] = e.lz;
    dirs[3*i + 0] = Lx; dirs[3*i + 1] = Ly; dirs[3*i + 2] = Lz;
  }
  return table.accepted;
}

// the level's sample table as the kernel sees it: entries[4*i..] = (lx, ly, lz, wh), *total = sum of weights
extern "C" int emu_table(int level, int levels, int samples, float *entries, double *total)
{
  LevelSamples table = build_level_samples(level, levels, samples);
  for(int i = 0; i < table.accepted; ++i)
  {
    entries[4*i + 0] = table.entries[i].lx; entries[4*i + 1] = table.entries[i].ly;
    entries[4*i + 2] = table.entries[i].lz; entries[4*i + 3] = table.entries[i].wh;
  }
  *total = table.total_weight;
  return table.accepted;
}

// ---- per-sector same-face limits (ibl_math.cuh sector_rho_limits, ibl_tables.h build_sector_entries) ----
//
// For the destination level of a ws x ws source: out[0] = share of the warp-samples the pair kernel sends
// through the cube-face selection with ONE isotropic band count per tile (round 2's rule), out[1] = the
// same with one count per warp from its sector's limit, out[2] = violations: samples a texel's own sector
// limit admits to the same-face path whose direction (evaluated in double precision) is not strictly
// inside the texel's face.  `sectors` = warps per tile (4 or 8).
extern "C" void emu_sector_study(int ws, int level, int levels, int samples, int sectors, double *out)
{
  const int band = 4 * sectors;      // the kernel's sector tables: four entries per sector and band
  LevelSamples ls = build_level_samples(level, levels, samples);
  BandedSamples banded = build_banded_samples(level, levels, samples, 16);
  SectorTable st = build_sector_entries(ls, sectors, band, 1.0f);
  const int per = band / sectors;

  const float kPi = 3.14159265358979323846f;
  float angles[6] = { -kPi/2, kPi/2, -kPi/2, kPi/2, 0.0f, kPi };
  int axes[6] = { 1, 1, 0, 0, 1, 1 };
  Quatf quats[6];
  for(int f = 0; f < 6; ++f)
  {
    float c = std::cos(angles[f]/2), s = std::sin(angles[f]/2);
    quats[f] = Quatf{ c, axes[f] == 0 ? s : 0.0f, axes[f] == 1 ? s : 0.0f, 0.0f };
  }

  int wd = ws >> 1, hd = ws >> 1;
  double old_general = 0, new_general = 0, old_total = 0, new_total = 0, violations = 0, interior_tiles = 0, tiles = 0;

  std::vector<float> limits((size_t)wd * hd * kFrameSectors);
  std::vector<float> thresholds((size_t)wd * hd);

  for(int face = 0; face < 6; ++face)
  {
    for(int y = 0; y < hd; ++y)
      for(int x = 0; x < wd; ++x)
      {
        Vec3f N = texel_normal(quats[face], x, y, wd, hd);
        Vec3f T, B;
        tangent_frame(N, T, B);
        Vec3f Tl = to_face_local(face, T), Bl = to_face_local(face, B), Nl = to_face_local(face, N);
        float *lim = limits.data() + ((size_t)y * wd + x) * kFrameSectors;
        sector_rho_limits(Tl, Bl, Nl, lim);
        thresholds[(size_t)y * wd + x] = same_face_threshold(Nl);

        // safety of the texel's own limits
        for(int w = 0; w < sectors; ++w)
        {
          float rho_lane = 3.0e38f;
          for(int k = w * kFrameSectors / sectors; k < (w + 1) * kFrameSectors / sectors; ++k)
            rho_lane = std::min(rho_lane, lim[k]);

          for(int k = 0; k < st.bands; ++k)
          {
            if (!(st.rho_max[(size_t)w * st.bands + k] <= rho_lane))
              break;
            for(int i = 0; i < per; ++i)
            {
              size_t e = (size_t)k * band + (size_t)w * per + i;
              float const *q = st.entries.data() + 4 * (e & ~(size_t)1);
              double X = q[0 + (e & 1)], Y = q[2 + (e & 1)];
              double a = X * Tl.x + Y * Bl.x + Nl.x, b = X * Tl.y + Y * Bl.y + Nl.y, m = X * Tl.z + Y * Bl.z + Nl.z;
              if (!(m > 0 && std::fabs(a) < m && std::fabs(b) < m))
                violations += 1;
            }
          }
        }
      }

    // tiles of 8x4 texels
    for(int ty = 0; ty < hd; ty += 4)
      for(int tx = 0; tx < wd; tx += 8)
      {
        float threshold = 0.0f;
        float rho_tile[8];
        for(int w = 0; w < sectors; ++w)
          rho_tile[w] = 3.0e38f;

        for(int y = ty; y < std::min(hd, ty + 4); ++y)
          for(int x = tx; x < std::min(wd, tx + 8); ++x)
          {
            threshold = std::max(threshold, thresholds[(size_t)y * wd + x]);
            float const *lim = limits.data() + ((size_t)y * wd + x) * kFrameSectors;
            for(int w = 0; w < sectors; ++w)
              for(int k = w * kFrameSectors / sectors; k < (w + 1) * kFrameSectors / sectors; ++k)
                rho_tile[w] = std::min(rho_tile[w], lim[k]);
          }

        int bands_old = (int)banded.band_min_lz.size();
        int n_same = 0;
        while (n_same < bands_old && banded.band_min_lz[n_same] > threshold)
          ++n_same;
        old_general += bands_old - n_same;
        old_total += bands_old;

        bool interior = true;
        for(int w = 0; w < sectors; ++w)
        {
          int n = 0;
          while (n < st.bands && st.rho_max[(size_t)w * st.bands + n] <= rho_tile[w])
            ++n;
          new_general += st.bands - n;
          new_total += st.bands;
          interior = interior && n == st.bands;
        }
        interior_tiles += interior ? 1 : 0;
        tiles += 1;
      }
  }

  out[0] = old_general / old_total;
  out[1] = new_general / new_total;
  out[2] = violations;
  out[3] = (double)st.bands * band / (double)(banded.band_min_lz.size() * 16);   // table growth through filling up
  out[4] = interior_tiles / tiles;   // tiles none of whose samples can leave the face
}

// Per row of 8x4 tiles of the destination level of a ws x ws source: how many (warp, band) shares the pair kernel
// sends through the cube-face selection (out_general[tile_row]) of how many in all (out_total[tile_row]).
// For studies of how evenly row slabs of a level load the GPUs that share a probe.
extern "C" void emu_row_costs(int ws, int level, int levels, int samples, int sectors, double *out_general, double *out_total)
{
  const int band = 4 * sectors;
  LevelSamples ls = build_level_samples(level, levels, samples);
  SectorTable st = build_sector_entries(ls, sectors, band, 1.0f);

  const float kPi = 3.14159265358979323846f;
  float angles[6] = { -kPi/2, kPi/2, -kPi/2, kPi/2, 0.0f, kPi };
  int axes[6] = { 1, 1, 0, 0, 1, 1 };
  Quatf quats[6];
  for(int f = 0; f < 6; ++f)
  {
    float c = std::cos(angles[f]/2), s = std::sin(angles[f]/2);
    quats[f] = Quatf{ c, axes[f] == 0 ? s : 0.0f, axes[f] == 1 ? s : 0.0f, 0.0f };
  }

  int wd = ws >> 1, hd = ws >> 1;
  std::vector<float> limits((size_t)wd * kFrameSectors);

  for(int face = 0; face < 6; ++face)
    for(int ty = 0; ty < hd; ty += 4)
    {
      int tile_row = (face * hd + ty) / 4;
      out_general[tile_row] = 0;
      out_total[tile_row] = 0;

      for(int tx = 0; tx < wd; tx += 8)
      {
        float rho_tile[8];
        for(int w = 0; w < sectors; ++w)
          rho_tile[w] = 3.0e38f;

        for(int y = ty; y < std::min(hd, ty + 4); ++y)
          for(int x = tx; x < std::min(wd, tx + 8); ++x)
          {
            Vec3f N = texel_normal(quats[face], x, y, wd, hd);
            Vec3f T, B;
            tangent_frame(N, T, B);
            float lim[kFrameSectors];
            sector_rho_limits(to_face_local(face, T), to_face_local(face, B), to_face_local(face, N), lim);
            for(int w = 0; w < sectors; ++w)
              for(int k = w * kFrameSectors / sectors; k < (w + 1) * kFrameSectors / sectors; ++k)
                rho_tile[w] = std::min(rho_tile[w], lim[k]);
          }

        for(int w = 0; w < sectors; ++w)
        {
          int n = 0;
          while (n < st.bands && st.rho_max[(size_t)w * st.bands + n] <= rho_tile[w])
            ++n;
          out_general[tile_row] += st.bands - n;
          out_total[tile_row] += st.bands;
        }
      }
    }
}

// ---- the projective footprints against the plain ones (ibl_math.cuh) ----
//
// n pseudo-random directions (and the exact cube diagonals / edge ties) through cube_footprint and
// cube_footprint_proj for a ws x hs source: out[0] = largest difference of the texel coordinates the two
// forms address (i + fraction, j + fraction), out[1] = footprints of the projective form outside
// [0, ws-2] x [0, hs-2] or on another face than the plain form's, out[2] = directions tested.
extern "C" void emu_footprint_compare(int ws, int hs, int n, unsigned seed, double *out)
{
  LevelGeom g = make_level_geom(ws, hs);
  double worst = 0, bad = 0, count = 0;
  unsigned state = seed * 2654435761u + 12345u;
  auto rnd = [&]() { state = state * 1664525u + 1013904223u; return (float)((state >> 8) & 0xFFFF) / 32768.0f - 1.0f; };

  for(int k = 0; k < n + 26; ++k)
  {
    float x, y, z;
    if (k < 26)
    {
      // axes, edge and corner diagonals: exact ties of the face selection
      int t = k + (k >= 13 ? 1 : 0);
      x = (float)(t % 3 - 1); y = (float)((t / 3) % 3 - 1); z = (float)((t / 9) % 3 - 1);
      if (x == 0 && y == 0 && z == 0)
        continue;
    }
    else
    {
      x = rnd(); y = rnd(); z = rnd();
      if (fabsf(x) + fabsf(y) + fabsf(z) < 1e-3f)
        continue;
    }

    float du0, dv0, du1, dv1;
    uint32_t idx0 = cube_footprint(g, x, y, z, du0, dv0);
    uint32_t face1;
    uint32_t raw = cube_footprint_proj(g, x, y, z, du1, dv1, face1);
    uint32_t idx1 = raw - g.bias_general + face1 * g.face_size;

    uint32_t face0 = idx0 / g.face_size;
    int i0 = (int)((idx0 % g.face_size) % (uint32_t)ws), j0 = (int)((idx0 % g.face_size) / (uint32_t)ws);
    long long local = (long long)idx1 - (long long)face1 * g.face_size;
    int i1 = (int)(local % ws), j1 = (int)(local / ws);

    count += 1;
    if (face0 != face1 || local < 0 || i1 < 0 || i1 > ws - 2 || j1 < 0 || j1 > hs - 2)
    {
      bad += 1;
      continue;
    }
    worst = std::max(worst, std::fabs(((double)i0 + du0) - ((double)i1 + du1)));
    worst = std::max(worst, std::fabs(((double)j0 + dv0) - ((double)j1 + dv1)));
  }

  out[0] = worst; out[1] = bad; out[2] = count;
}

// the pair kernel's sector table: entries (4 floats each, pair-interleaved) and rho_max; returns bands or -1
extern "C" int emu_sector_table(int level, int levels, int samples, int sectors, float scale, float *entries, int capacity, float *rho_max)
{
  SectorTable st = build_sector_entries(build_level_samples(level, levels, samples), sectors, 4 * sectors, scale);
  if ((int)st.entries.size() > capacity)
    return -1;
  for(size_t i = 0; i < st.entries.size(); ++i)
    entries[i] = st.entries[i];
  for(size_t i = 0; i < st.rho_max.size(); ++i)
    rho_max[i] = st.rho_max[i];
  return st.bands;
}
