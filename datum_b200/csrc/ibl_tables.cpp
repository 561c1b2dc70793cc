// datum_b200 — per-level GGX sample tables (host side); see ibl_tables.h.

#include "ibl_tables.h"

#include <algorithm>
#include <cmath>

namespace ibl
{
  float radicalinverse_VdC(uint32_t bits)
  {
    uint32_t r = 0;
    for(int k = 0; k < 32; ++k, bits >>= 1)
      r = (r << 1) | (bits & 1u);

    return float(r) * 2.3283064365386963e-10f; // 2^-32, rounded to fp32 like ibl.cpp:103
  }

  LevelSamples build_level_samples(int level, int levels, int samples)
  {
    LevelSamples out;

    // ibl.cpp:251, 173: roughness = level/(levels-1), alpha = roughness^2 — in fp32 like the reference
    float roughness = (float)level / (float)(levels - 1);
    float alpha = roughness * roughness;
    float a2m1 = alpha * alpha - 1;

    out.roughness = roughness;
    out.entries.reserve(samples);

    for(int i = 0; i < samples; ++i)
    {
      // ibl.cpp:106-109, 119-121 in fp32 so that (phi, theta) are the reference's values
      float ux = float(i) / float(samples);
      float uy = radicalinverse_VdC((uint32_t)i);
      float phi = 2 * 3.14159265358979323846f * ux;
      float costheta = std::sqrt((1 - uy) / (1 + a2m1 * uy));
      float sintheta = std::sqrt(1 - costheta * costheta);

      double hx = (double)sintheta * std::cos((double)phi);
      double hy = (double)sintheta * std::sin((double)phi);
      double hz = (double)costheta;

      double lz = 2 * hz * hz - 1;

      if (!(lz > 0))
        continue; // ibl.cpp:178

      SampleEntry e;
      e.lx = (float)(2 * hz * hx);
      e.ly = (float)(2 * hz * hy);
      e.lz = (float)std::min(lz, 1.0);
      e.wh = 0.5f * e.lz;

      out.total_weight += (double)e.lz;
      out.entries.push_back(e);
    }

    out.accepted = (int)out.entries.size();

    std::stable_sort(out.entries.begin(), out.entries.end(), [](SampleEntry const &a, SampleEntry const &b) { return a.lz > b.lz; });

    return out;
  }

  BandedSamples build_banded_samples(int level, int levels, int samples, int band, int order)
  {
    BandedSamples out;
    out.level = build_level_samples(level, levels, samples);
    out.band = band;

    auto &e = out.level.entries;

    auto by_angle = [](SampleEntry const &a, SampleEntry const &b) {
      return std::atan2((double)a.ly, (double)a.lx) < std::atan2((double)b.ly, (double)b.lx);
    };

    if (order == 0)
    {
      // rings: every band is one ring of the lobe-angle order, walked by angle
      for(size_t begin = 0; begin < e.size(); begin += (size_t)band)
      {
        size_t end = std::min(e.size(), begin + (size_t)band);

        out.band_min_lz.push_back(e[end - 1].lz);

        std::stable_sort(e.begin() + begin, e.begin() + end, by_angle);
      }

      return out;
    }

    // patches: a ring of s*band consecutive entries of the lobe-angle order is cut into s sectors of
    // `band` entries, s chosen so that a sector is about as wide as it is deep.  The warps of a tile
    // (they share a band) then fetch from ONE compact patch of the lobe instead of from four quarters
    // of a ring: on the wider lobes of the higher levels that is the difference between sharing the
    // footprint in L1 and not.
    auto radius = [](SampleEntry const &a) { return std::sqrt((double)a.lx * a.lx + (double)a.ly * a.ly); };

    struct Band { size_t begin; float min_lz; };
    std::vector<Band> bands;

    for(size_t begin = 0; begin < e.size(); )
    {
      size_t left = e.size() - begin;
      size_t sectors = 1;

      double best = 1e300;
      for(size_t s = 1; s <= 8 && s * (size_t)band <= left; ++s)
      {
        double r0 = radius(e[begin]), r1 = radius(e[begin + s * band - 1]);
        double depth = std::max(r1 - r0, 1e-12);
        double width = std::max(2 * 3.14159265358979323846 * 0.5 * (r0 + r1) / (double)s, 1e-12);
        double skew = std::fabs(std::log(depth / width));
        if (skew < best)
        {
          best = skew;
          sectors = s;
        }
      }

      size_t end = std::min(e.size(), begin + sectors * (size_t)band);

      std::stable_sort(e.begin() + begin, e.begin() + end, by_angle);

      for(size_t b = begin; b < end; b += (size_t)band)
      {
        size_t bend = std::min(end, b + (size_t)band);
        float min_lz = e[b].lz;
        for(size_t i = b; i < bend; ++i)
          min_lz = std::min(min_lz, e[i].lz);
        bands.push_back(Band{ b, min_lz });
      }

      begin = end;
    }

    // the kernel's same-face search needs the bands' smallest lz in decreasing order; the short band (if
    // any) stays last
    size_t full = e.size() / (size_t)band;
    std::stable_sort(bands.begin(), bands.begin() + std::min(full, bands.size()), [](Band const &a, Band const &b) { return a.min_lz > b.min_lz; });

    std::vector<SampleEntry> ordered;
    ordered.reserve(e.size());
    float floor_lz = 1e30f;
    for(auto const &b : bands)
    {
      size_t bend = std::min(e.size(), b.begin + (size_t)band);
      ordered.insert(ordered.end(), e.begin() + b.begin, e.begin() + bend);
      floor_lz = std::min(floor_lz, b.min_lz);     // monotone whatever the short band holds
      out.band_min_lz.push_back(floor_lz);
    }
    e.swap(ordered);

    return out;
  }

  std::vector<float> build_paired_entries(BandedSamples const &banded, float scale)
  {
    std::vector<SampleEntry> e = banded.level.entries;
    for(auto &s : e)
    {
      s.lx *= scale; s.ly *= scale; s.lz *= scale; s.wh *= scale;
    }

    // fill the last band: direction (0, 0, 1) in the tangent frame, i.e. the normal itself (always on
    // the texel's own face), at 2^-60 of the scale of a real entry
    size_t band = (size_t)(banded.band > 0 ? banded.band : 1);
    size_t padded = (e.size() + band - 1) / band * band;
    float tiny = std::ldexp(scale, -60);
    padded += padded & 1;   // pairs: an even count whatever the band size
    while (e.size() < padded)
      e.push_back(SampleEntry{ 0.0f, 0.0f, tiny, 0.5f * tiny });

    std::vector<float> out(4 * padded);
    for(size_t i = 0; i < padded; i += 2)
    {
      SampleEntry const &a = e[i];
      SampleEntry const &b = e[i + 1];
      float *o = out.data() + 4 * i;
      o[0] = a.lx; o[1] = b.lx; o[2] = a.ly; o[3] = b.ly;
      o[4] = a.lz; o[5] = b.lz; o[6] = a.wh; o[7] = b.wh;
    }

    return out;
  }

  SectorTable build_sector_entries(LevelSamples const &level, int sectors, int band, float scale)
  {
    SectorTable out;
    out.sectors = sectors;
    out.band = band;

    const int per = band / sectors;
    const double kPi = 3.14159265358979323846;

    struct Projective { float x, y, lz, wh; double phi; };
    std::vector<std::vector<Projective>> cut(sectors);

    for(auto const &e : level.entries)
    {
      Projective s;
      s.x = (float)((double)e.lx / (double)e.lz);     // lz > 0 for every accepted sample (ibl.cpp:178)
      s.y = (float)((double)e.ly / (double)e.lz);
      s.lz = e.lz * scale;
      s.wh = e.wh * scale;
      s.phi = std::atan2((double)s.y, (double)s.x);   // of the values the kernel multiplies with

      int k = (int)std::floor((s.phi + kPi) / (2 * kPi / sectors));
      k = std::min(std::max(k, 0), sectors - 1);
      cut[k].push_back(s);
    }

    size_t longest = 0;
    for(auto &c : cut)
    {
      std::stable_sort(c.begin(), c.end(), [](Projective const &a, Projective const &b) { return a.lz > b.lz; });
      longest = std::max(longest, c.size());
    }

    out.bands = (int)((longest + (size_t)per - 1) / (size_t)per);
    if (out.bands < 1)
      out.bands = 1;

    const float tiny = std::ldexp(scale, -60);
    std::vector<Projective> flat((size_t)out.bands * band);
    out.rho_max.assign((size_t)sectors * out.bands, 0.0f);

    for(int w = 0; w < sectors; ++w)
    {
      float running = 0.0f;    // shares are lz-sorted, so the maximum grows anyway; keep it monotone by construction

      for(int k = 0; k < out.bands; ++k)
      {
        std::vector<Projective> share;
        for(int i = 0; i < per; ++i)
        {
          size_t at = (size_t)k * per + i;
          if (at < cut[w].size())
            share.push_back(cut[w][at]);
        }

        for(auto const &s : share)
        {
          float rho = (float)std::hypot((double)s.x, (double)s.y);
          running = std::max(running, std::nextafter(rho, 3.0e38f));
        }
        out.rho_max[(size_t)w * out.bands + k] = share.empty() ? 0.0f : running;

        std::stable_sort(share.begin(), share.end(), [](Projective const &a, Projective const &b) { return a.phi < b.phi; });
        while ((int)share.size() < per)
          share.push_back(Projective{ 0.0f, 0.0f, tiny, 0.5f * tiny, 0.0 });

        for(int i = 0; i < per; ++i)
          flat[(size_t)k * band + (size_t)w * per + i] = share[i];
      }

      // a filled-up share costs nothing on either path: let it count as same-face whenever the one before does
      for(int k = 1; k < out.bands; ++k)
        if (out.rho_max[(size_t)w * out.bands + k] == 0.0f)
          out.rho_max[(size_t)w * out.bands + k] = out.rho_max[(size_t)w * out.bands + k - 1];
    }

    out.entries.resize(4 * flat.size());
    for(size_t i = 0; i < flat.size(); i += 2)
    {
      Projective const &a = flat[i];
      Projective const &b = flat[i + 1];
      float *o = out.entries.data() + 4 * i;
      o[0] = a.x; o[1] = b.x; o[2] = a.y; o[3] = b.y;
      o[4] = a.lz; o[5] = b.lz; o[6] = a.wh; o[7] = b.wh;
    }

    return out;
  }
}
