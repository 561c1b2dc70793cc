// TEST INFRASTRUCTURE — CPU oracle for the datum image-based-lighting bake.
//
// A plain restatement of the reference algorithm, written from the reference
// sources (no code copied), each function citing the file:line it follows
// (paths relative to /root/reference).  It is the parity checker for the CUDA
// path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it.  Nothing under datum_b200/ does.
//
// PARITY PINNING: the reference has no tests or golden vectors for this path
// (SURVEY.md §4), so the oracle is pinned against the reference ITSELF:
// tests/test_oracle_vs_ref.py checks that, for the reference's fixed
// 1024-sample RGBE bake, this restatement is word-for-word identical to the
// unmodified tools/ibl.cpp compiled into oracle/_ref (strict IEEE flags), and
// tests/golden/ holds vectors generated from oracle/_ref by
// tests/golden/make_golden.py.  What stays unpinned is the third-party `leap`
// math library (un-vendored, un-pinned): its lerp/normalise/fmod2 semantics are
// assumed as stated in oracle/shim/leap/lml/vector.h.
//
// Extensions over the reference (needed by BASELINE.json's configs, each
// degenerating to the reference when the parameter takes the reference's
// value): `samples` is a parameter (ibl.cpp:162 hard-codes 1024; the GLSL twin
// data/convolve.comp:8 already takes it), the fp32 pre-quantisation texels can
// be returned next to the packed words, and a level can be computed for a row
// range only (multi-GPU sharding tests).
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off, no -ffast-math, so the
// result is deterministic IEEE fp32).

#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>

#include <atomic>
#include <thread>

namespace
{
  struct V3 { float x, y, z; };
  struct C4 { float r, g, b, a; };
  struct Quat { float w, x, y, z; };

  const float kPi = 3.14159265358979323846f; // leap pi<float>()

  // ---- lml vector algebra as used on the path (term order matters for bit parity) ----

  inline float dot3(V3 const &u, V3 const &v) { float r = 0; r += u.x*v.x; r += u.y*v.y; r += u.z*v.z; return r; }
  inline V3 cross3(V3 const &u, V3 const &v) { return { u.y*v.z - u.z*v.y, u.z*v.x - u.x*v.z, u.x*v.y - u.y*v.x }; }
  inline V3 normalise3(V3 const &v) { float n = std::sqrt(dot3(v, v)); return { v.x/n, v.y/n, v.z/n }; }
  inline V3 scale3(float s, V3 const &v) { return { s*v.x, s*v.y, s*v.z }; }
  inline V3 add3(V3 const &u, V3 const &v) { return { u.x + v.x, u.y + v.y, u.z + v.z }; }
  inline V3 sub3(V3 const &u, V3 const &v) { return { u.x - v.x, u.y - v.y, u.z - v.z }; }
  inline float clampf(float v, float lo, float hi) { return std::max(lo, std::min(v, hi)); }
  inline C4 lerp4(C4 const &a, C4 const &b, float t)
  {
    float s = 1 - t;
    return { s*a.r + t*b.r, s*a.g + t*b.g, s*a.b + t*b.b, s*a.a + t*b.a };
  }
  inline float fmod2(float a, float b) { float r = std::fmod(a, b); return (r < 0) ? r + b : r; }

  // ---- quaternion / dual-quaternion rotation: src/math/transform.h:65-68, 161-178 ----

  inline Quat qmul(Quat const &a, Quat const &b)
  {
    return { a.w*b.w - a.x*b.x - a.y*b.y - a.z*b.z,
             a.w*b.x + a.x*b.w + a.y*b.z - a.z*b.y,
             a.w*b.y + a.y*b.w + a.z*b.x - a.x*b.z,
             a.w*b.z + a.z*b.w + a.x*b.y - a.y*b.x };
  }
  inline Quat qadd(Quat const &a, Quat const &b) { return { a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z }; }

  struct Xform { Quat real, dual; };

  inline Xform xrotation(V3 const &axis, float angle)
  {
    float c = std::cos(angle/2), s = std::sin(angle/2);
    return { { c, axis.x*s, axis.y*s, axis.z*s }, { 0, 0, 0, 0 } };
  }
  inline Xform xmul(Xform const &a, Xform const &b) { return { qmul(a.real, b.real), qadd(qmul(a.real, b.dual), qmul(a.dual, b.real)) }; }
  inline V3 xapply(Xform const &t, V3 const &v)
  {
    Xform p = { { 1, 0, 0, 0 }, { 0, v.x, v.y, v.z } };
    Xform c = { { t.real.w, -t.real.x, -t.real.y, -t.real.z }, { -t.dual.w, t.dual.x, t.dual.y, t.dual.z } };
    Xform r = xmul(xmul(t, p), c);
    return { r.dual.x, r.dual.y, r.dual.z };
  }

  // the six face rotations of tools/ibl.cpp:253-261 (same table in tools/hdr.cpp:335-343)
  inline void face_rotations(Xform out[6])
  {
    out[0] = xrotation({ 0, 1, 0 }, -kPi/2); // right
    out[1] = xrotation({ 0, 1, 0 }, kPi/2);  // left
    out[2] = xrotation({ 1, 0, 0 }, -kPi/2); // bottom
    out[3] = xrotation({ 1, 0, 0 }, kPi/2);  // top
    out[4] = xrotation({ 0, 1, 0 }, 0);      // front
    out[5] = xrotation({ 0, 1, 0 }, kPi);    // back
  }

  // ---- E5B9G9R9-style shared exponent codec: src/math/color.h:154-172 ----

  inline uint32_t rgbe_encode(float cr, float cg, float cb)
  {
    float r = clampf(cr, 0.0f, 65408.0f);
    float g = clampf(cg, 0.0f, 65408.0f);
    float b = clampf(cb, 0.0f, 65408.0f);
    float e = std::max(-16.0f, std::floor(std::log2(std::max(r, std::max(g, b))))) + 1;

    return ((uint32_t)(uint8_t)(e + 15) << 27)
         | ((uint32_t)(uint16_t)std::round(r / std::exp2(e) * 511) << 0)
         | ((uint32_t)(uint16_t)std::round(g / std::exp2(e) * 511) << 9)
         | ((uint32_t)(uint16_t)std::round(b / std::exp2(e) * 511) << 18);
  }

  inline C4 rgbe_decode(uint32_t c)
  {
    float r = ((c >> 0) & 0x1FF) / 511.0f;
    float g = ((c >> 9) & 0x1FF) / 511.0f;
    float b = ((c >> 18) & 0x1FF) / 511.0f;
    float e = ((c >> 27) & 0x1F) - 15.0f;

    return { r * std::exp2(e), g * std::exp2(e), b * std::exp2(e), 1.0f };
  }

  // ---- cube sampler: tools/ibl.cpp:16-93 ----

  struct Sampler
  {
    int width, height;
    uint32_t const *faces[6]; // 0 right, 1 left, 2 down, 3 up, 4 forward, 5 back (ibl.cpp:21-26)

    Sampler(int w, int h, uint32_t const *bits) : width(w), height(h)
    {
      for(int f = 0; f < 6; ++f)
        faces[f] = bits + (size_t)f * w * h;
    }

    C4 texel(int face, int i, int j) const { return rgbe_decode(faces[face][j * width + i]); }

    // ibl.cpp:34-41 — align-corners bilinear that never leaves the face
    C4 sample(int face, float tx, float ty) const
    {
      float i, j;
      float u = std::modf(fmod2(tx, 1.0f) * (width - 1), &i);
      float v = std::modf(fmod2(ty, 1.0f) * (height - 1), &j);

      return lerp4(lerp4(texel(face, (int)i, (int)j), texel(face, (int)i+1, (int)j), u),
                   lerp4(texel(face, (int)i, (int)j+1), texel(face, (int)i+1, (int)j+1), u), v);
    }

    // ibl.cpp:43-88 — strict major-axis face select.  The reference leaves the
    // result uninitialised on exact ties (|x|==|y| etc.);
    // the oracle's rule for ties: x if |x|>=max(|y|,|z|), else y if |y|>=|z|, else z.
    // Non-tie directions take exactly the reference's branch.
    C4 sample(V3 const &d) const
    {
      float mx = std::abs(d.x), my = std::abs(d.y), mz = std::abs(d.z);

      if (mx >= std::max(my, mz))
      {
        if (d.x > 0)
          return sample(0, 0.5f + 0.5f*d.z/mx, 0.5f + 0.5f*d.y/mx);
        else
          return sample(1, 0.5f - 0.5f*d.z/mx, 0.5f + 0.5f*d.y/mx);
      }
      else if (my >= mz)
      {
        if (d.y > 0)
          return sample(3, 0.5f + 0.5f*d.x/my, 0.5f + 0.5f*d.z/my);
        else
          return sample(2, 0.5f + 0.5f*d.x/my, 0.5f - 0.5f*d.z/my);
      }
      else
      {
        if (d.z > 0)
          return sample(5, 0.5f - 0.5f*d.x/mz, 0.5f + 0.5f*d.y/mz);
        else
          return sample(4, 0.5f + 0.5f*d.x/mz, 0.5f + 0.5f*d.y/mz);
      }
    }
  };

  // ---- Hammersley / GGX: tools/ibl.cpp:95-128 ----

  inline float radicalinverse_VdC(uint32_t bits)
  {
    // 32-bit reversal: byte swap, then nibbles / pairs / single bits inside each byte
    uint32_t r = __builtin_bswap32(bits);
    r = ((r >> 4) & 0x0F0F0F0Fu) | ((r & 0x0F0F0F0Fu) << 4);
    r = ((r >> 2) & 0x33333333u) | ((r & 0x33333333u) << 2);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);

    return float(r) * 2.3283064365386963e-10f; // * 2^-32 (ibl.cpp:103)
  }

  inline V3 importancesample_ggx(float ux, float uy, float alpha, V3 const &normal)
  {
    float phi = 2*kPi * ux;
    float costheta = std::sqrt((1 - uy) / (1 + (alpha*alpha - 1) * uy));
    float sintheta = std::sqrt(1 - costheta*costheta);

    V3 up = std::abs(normal.z) < 0.999f ? V3{ 0, 0, 1 } : V3{ 1, 0, 0 };
    V3 tangent = normalise3(cross3(up, normal));
    V3 bitangent = cross3(normal, tangent);

    return add3(add3(scale3(sintheta * std::cos(phi), tangent), scale3(sintheta * std::sin(phi), bitangent)), scale3(costheta, normal));
  }

  inline V3 importancesample_cosdir(float ux, float uy, V3 const &normal)
  {
    float phi = 2*kPi * ux;
    float costheta = std::sqrt(std::max(0.0f, 1 - uy));
    float sintheta = std::sqrt(uy);

    V3 up = std::abs(normal.z) < 0.999f ? V3{ 0, 0, 1 } : V3{ 1, 0, 0 };
    V3 tangent = normalise3(cross3(up, normal));
    V3 bitangent = cross3(normal, tangent);

    return add3(add3(scale3(sintheta * std::cos(phi), tangent), scale3(sintheta * std::sin(phi), bitangent)), scale3(costheta, normal));
  }

  // ---- prefilter of one direction: tools/ibl.cpp:160-187 (samples parameterised) ----

  inline void convolve(float roughness, V3 const &ray, Sampler const &envmap, int samples, float out[3])
  {
    V3 N = ray;
    V3 V = ray;

    float sum[3] = { 0, 0, 0 };
    float totalweight = 0;

    for(int i = 0; i < samples; ++i)
    {
      float ux = float(i)/float(samples);
      float uy = radicalinverse_VdC(i);
      V3 H = importancesample_ggx(ux, uy, roughness * roughness, N);
      V3 L = sub3(scale3(2 * dot3(V, H), H), V);

      float NdotL = clampf(dot3(N, L), 0.0f, 1.0f);

      if (NdotL > 0)
      {
        C4 c = envmap.sample(L);

        sum[0] += c.r * NdotL;
        sum[1] += c.g * NdotL;
        sum[2] += c.b * NdotL;

        totalweight += NdotL;
      }
    }

    out[0] = sum[0] / totalweight;
    out[1] = sum[1] / totalweight;
    out[2] = sum[2] / totalweight;
  }

  // direction of output texel (face, x, y) at a wd x hd level: tools/ibl.cpp:269
  inline V3 texel_direction(Xform const &rot, int x, int y, int wd, int hd)
  {
    return xapply(rot, normalise3({ 2 * (x + 0.5f)/wd - 1, 2 * (y + 0.5f)/hd - 1, -1.0f }));
  }

  // ---- split-sum BRDF terms: tools/ibl.cpp:111-115, 143-158, 189-237 ----

  inline float GGX(float NdotV, float alpha) { float k = alpha / 2; return NdotV / (NdotV * (1.0f - k) + k); }
  inline float fresnel_schlick(float f0, float f90, float u) { return f0 + (f90 - f0) * std::pow(1 - u, 5.0f); }
  inline float lerp1(float a, float b, float t) { return (1 - t)*a + t*b; }

  inline float diffuse_disney(float NdotV, float NdotL, float LdotH, float alpha)
  {
    float energybias = 0.5f;
    float energyfactor = lerp1(1.0f, 1.0f / 1.51f, alpha);
    float f90 = energybias + 2 * LdotH*LdotH * alpha;

    float lightscatter = fresnel_schlick(1, f90, NdotL);
    float viewscatter = fresnel_schlick(1, f90, NdotV);

    return lightscatter * viewscatter * energyfactor;
  }

  inline void integrate(float roughness, float NdotV, int samples, float out[3])
  {
    V3 V = { std::sqrt(1.0f - NdotV * NdotV), 0.0f, NdotV };
    V3 Z = { 0, 0, 1 };

    float a = 0;
    float b = 0;

    for(int i = 0; i < samples; ++i)
    {
      float ux = float(i)/float(samples);
      float uy = radicalinverse_VdC(i);
      V3 H = importancesample_ggx(ux, uy, roughness * roughness, Z);
      V3 L = sub3(scale3(2 * dot3(V, H), H), V);

      float NdotL = clampf(L.z, 0.0f, 1.0f);
      float NdotH = clampf(H.z, 0.0f, 1.0f);
      float VdotH = clampf(dot3(V, H), 0.0f, 1.0f);

      if (NdotL > 0)
      {
        float G = GGX(NdotL, roughness * roughness) * GGX(NdotV, roughness * roughness);
        float Vis = G * VdotH / (NdotH * NdotV);
        float Fc = std::pow(1 - VdotH, 5.0f);

        a += (1 - Fc) * Vis;
        b += Fc * Vis;
      }
    }

    float c = 0;

    for(int i = 0; i < samples; ++i)
    {
      float hx = float(i)/float(samples) + 0.5f;
      float hy = radicalinverse_VdC(i) + 0.5f;
      float ux = hx - std::floor(hx);
      float uy = hy - std::floor(hy);
      V3 L = importancesample_cosdir(ux, uy, Z);

      float NdotL = clampf(L.z, 0.0f, 1.0f);

      if (NdotL > 0)
      {
        float LdotH = clampf(dot3(L, normalise3(add3(V, L))), 0.0f, 1.0f);

        c += diffuse_disney(NdotV, NdotL, LdotH, roughness * roughness);
      }
    }

    out[0] = a / samples;
    out[1] = b / samples;
    out[2] = c / samples;
  }

  // rows [begin,end) handed out one at a time to `threads` std::threads (0 = all cores)
  template<typename F>
  void parallel_rows(int begin, int end, int threads, F body)
  {
    if (threads <= 0)
      threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = std::min(threads, std::max(1, end - begin));

    if (threads == 1)
    {
      for(int row = begin; row < end; ++row)
        body(row);
      return;
    }

    std::atomic<int> next(begin);
    std::vector<std::thread> pool;
    for(int t = 0; t < threads; ++t)
      pool.emplace_back([&]() { for(int row = next++; row < end; row = next++) body(row); });
    for(auto &t : pool)
      t.join();
  }

  // closed-form face directions (double) — data/convolve.comp:85-100, equal to
  // the quaternion rotations above up to 1 ulp
  inline void face_dir_d(int face, double u, double v, double d[3])
  {
    switch (face)
    {
      case 0: d[0] = 1;  d[1] = v;  d[2] = u;  break; // right
      case 1: d[0] = -1; d[1] = v;  d[2] = -u; break; // left
      case 2: d[0] = u;  d[1] = -1; d[2] = -v; break; // bottom
      case 3: d[0] = u;  d[1] = 1;  d[2] = v;  break; // top
      case 4: d[0] = u;  d[1] = v;  d[2] = -1; break; // front
      default: d[0] = -u; d[1] = v; d[2] = 1;  break; // back
    }
  }
}

extern "C"
{
  // src/math/color.h:154-162
  uint32_t oracle_rgbe_encode(float r, float g, float b) { return rgbe_encode(r, g, b); }

  // src/math/color.h:164-172
  void oracle_rgbe_decode(uint32_t word, float *rgba)
  {
    C4 c = rgbe_decode(word);
    rgba[0] = c.r; rgba[1] = c.g; rgba[2] = c.b; rgba[3] = c.a;
  }

  void oracle_rgbe_encode_array(float const *rgb, size_t count, int stride, uint32_t *words)
  {
    for(size_t i = 0; i < count; ++i)
      words[i] = rgbe_encode(rgb[i*stride + 0], rgb[i*stride + 1], rgb[i*stride + 2]);
  }

  void oracle_rgbe_decode_array(uint32_t const *words, size_t count, float *rgba)
  {
    for(size_t i = 0; i < count; ++i)
      oracle_rgbe_decode(words[i], rgba + 4*i);
  }

  // src/math/color.h:103-106, 115-118, 125-128 — ARGB32 pixel -> linear via pow 2.2
  void oracle_srgba_decode(uint32_t argb, float *rgba)
  {
    rgba[0] = std::pow((uint8_t)(argb >> 16) / 255.0f, 2.2f);
    rgba[1] = std::pow((uint8_t)(argb >> 8) / 255.0f, 2.2f);
    rgba[2] = std::pow((uint8_t)(argb >> 0) / 255.0f, 2.2f);
    rgba[3] = (uint8_t)(argb >> 24) / 255.0f;
  }

  // tools/assetbuilder.cpp:443-462: six ARGB32 images -> level 0.  Per pixel
  // image.setPixel(x, y, rgbe(srgba(image.pixel(x, y)))) (:454), then image.mirrored() (:458, a
  // vertical flip) and the memcpy behind the previous face (:460-462).
  void oracle_ingest_cube_argb32(uint32_t const *argb, int width, int height, uint32_t *level0)
  {
    size_t face_size = (size_t)width * height;

    for(int face = 0; face < 6; ++face)
    {
      for(int y = 0; y < height; ++y)
      {
        for(int x = 0; x < width; ++x)
        {
          float rgba[4];
          oracle_srgba_decode(argb[face * face_size + (size_t)y * width + x], rgba);

          level0[face * face_size + (size_t)(height - 1 - y) * width + x] = rgbe_encode(rgba[0], rgba[1], rgba[2]);
        }
      }
    }
  }

  // tools/ibl.cpp:95-104
  float oracle_radicalinverse(uint32_t bits) { return radicalinverse_VdC(bits); }

  // tools/ibl.cpp:269 + src/math/transform.h:173-178
  void oracle_texel_direction(int face, int x, int y, int wd, int hd, float *out)
  {
    Xform rot[6];
    face_rotations(rot);
    V3 d = texel_direction(rot[face], x, y, wd, hd);
    out[0] = d.x; out[1] = d.y; out[2] = d.z;
  }

  // src/math/transform.h:173-178 with the face rotations of tools/ibl.cpp:253-261
  void oracle_face_rotate(int face, float const *v, float *out)
  {
    Xform rot[6];
    face_rotations(rot);
    V3 r = xapply(rot[face], V3{ v[0], v[1], v[2] });
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
  }

  // tools/ibl.cpp:43-88 on one direction (rgba out)
  void oracle_cube_sample(uint32_t const *level, int ws, int hs, float const *dir, float *rgba)
  {
    Sampler envmap(ws, hs, level);
    C4 c = envmap.sample(V3{ dir[0], dir[1], dir[2] });
    rgba[0] = c.r; rgba[1] = c.g; rgba[2] = c.b; rgba[3] = c.a;
  }

  // tools/ibl.cpp:160-187 for one direction
  void oracle_convolve_dir(uint32_t const *level, int ws, int hs, float roughness, int samples, float const *dir, float *rgb)
  {
    Sampler envmap(ws, hs, level);
    convolve(roughness, V3{ dir[0], dir[1], dir[2] }, envmap, samples, rgb);
  }

  // per-sample reflected directions and weights of tools/ibl.cpp:170-184 for one
  // direction (debug aid for the parity tests): dirs[3*i..], ndotl[i] (<= 0 when rejected)
  void oracle_trace_samples(float roughness, int samples, float const *dir, float *dirs, float *ndotl)
  {
    V3 N = { dir[0], dir[1], dir[2] };
    for(int i = 0; i < samples; ++i)
    {
      V3 H = importancesample_ggx(float(i)/float(samples), radicalinverse_VdC(i), roughness * roughness, N);
      V3 L = sub3(scale3(2 * dot3(N, H), H), N);
      dirs[3*i + 0] = L.x; dirs[3*i + 1] = L.y; dirs[3*i + 2] = L.z;
      ndotl[i] = clampf(dot3(N, L), 0.0f, 1.0f);
    }
  }

  // Per output texel of a (wd x hd x 6) level: how many accepted samples point
  // within `eps` (relative) of a cube-face boundary, i.e. have their two largest
  // |components| closer than eps.  For those samples ibl.cpp:51-85 is decided by
  // the last bit of fp32 rounding (exact ties are undefined behaviour there, and
  // a quotient that rounds to 1.0 wraps to the opposite edge through fmod2,
  // ibl.cpp:37-38), so the reference's own strict and -ffast-math builds disagree
  // on them.  The parity tests hold texels with a non-zero count to a looser bound.
  void oracle_edge_ambiguous_counts_rows(int wd, int hd, float roughness, int samples, float eps, int row_begin, int row_end, int32_t *counts, int threads);

  void oracle_edge_ambiguous_counts(int wd, int hd, float roughness, int samples, float eps, int32_t *counts, int threads)
  {
    oracle_edge_ambiguous_counts_rows(wd, hd, roughness, samples, eps, 0, 6 * hd, counts, threads);
  }

  // the same for the face-major rows [row_begin, row_end) only; `counts` is indexed from the start of the level
  void oracle_edge_ambiguous_counts_rows(int wd, int hd, float roughness, int samples, float eps, int row_begin, int row_end, int32_t *counts, int threads)
  {
    Xform rot[6];
    face_rotations(rot);

    parallel_rows(row_begin, row_end, threads, [&](int row)
    {
      int face = row / hd;
      int y = row % hd;

      for(int x = 0; x < wd; ++x)
      {
        V3 N = texel_direction(rot[face], x, y, wd, hd);

        int count = 0;
        for(int i = 0; i < samples; ++i)
        {
          V3 H = importancesample_ggx(float(i)/float(samples), radicalinverse_VdC(i), roughness * roughness, N);
          V3 L = sub3(scale3(2 * dot3(N, H), H), N);

          if (clampf(dot3(N, L), 0.0f, 1.0f) > 0)
          {
            float m[3] = { std::abs(L.x), std::abs(L.y), std::abs(L.z) };
            std::sort(m, m + 3);
            if (m[1] >= m[2] * (1 - eps))
              ++count;
          }
        }

        counts[(size_t)row * wd + x] = count;
      }
    });
  }

  // One level of tools/ibl.cpp:247-278: `src` is a (ws x hs x 6) rgbe level, the
  // output level is (ws/2 x hs/2 x 6).  Rows are numbered face-major over
  // 6*(hs/2); [row_begin,row_end) selects a slab.  `words` / `f32` (rgb triples,
  // pre-quantisation — the value handed to rgbe() at ibl.cpp:269) are indexed
  // from the start of the level; either may be null.
  void oracle_prefilter_level(uint32_t const *src, int ws, int hs, float roughness, int samples, int row_begin, int row_end, uint32_t *words, float *f32, int threads)
  {
    Sampler envmap(ws, hs, src);

    Xform rot[6];
    face_rotations(rot);

    int wd = ws >> 1, hd = hs >> 1;

    parallel_rows(row_begin, row_end, threads, [&](int row)
    {
      int face = row / hd;
      int y = row % hd;

      for(int x = 0; x < wd; ++x)
      {
        float rgb[3];
        convolve(roughness, texel_direction(rot[face], x, y, wd, hd), envmap, samples, rgb);

        size_t idx = (size_t)row * wd + x;

        if (words)
          words[idx] = rgbe_encode(rgb[0], rgb[1], rgb[2]);

        if (f32)
        {
          f32[3*idx + 0] = rgb[0];
          f32[3*idx + 1] = rgb[1];
          f32[3*idx + 2] = rgb[2];
        }
      }
    });
  }

  // tools/ibl.cpp:242-279 — the whole chain, in place on `bits` (level 0 pre-filled).
  // `f32` (optional) receives the pre-quantisation rgb of levels >= 1, level-major,
  // starting at level 1.
  void oracle_buildmips_cube_ibl(int width, int height, int levels, int samples, uint32_t *bits, float *f32, int threads)
  {
    uint32_t *src = bits;
    uint32_t *dst = src + (size_t)width * height * 6;

    for(int level = 1; level < levels; ++level)
    {
      float roughness = (float)level / (float)(levels - 1);

      oracle_prefilter_level(src, width, height, roughness, samples, 0, 6 * (height >> 1), dst, f32, threads);

      size_t outcount = (size_t)(width >> 1) * (height >> 1) * 6;

      src += (size_t)width * height * 6;
      dst += outcount;
      if (f32)
        f32 += 3 * outcount;

      width /= 2;
      height /= 2;
    }
  }

  // SH9 projection of level 0: data/project.comp:23-106, restated with fp64
  // weights and accumulation (the shader's sequential fp32 sum over up to 1e8
  // texels is not a usable numerical reference).
  //   format 0: `level0` is rgbe words;  format 1: RGBA fp32.
  // `partial` (28 doubles: 27 unnormalised sums in [k][rgb] order, then the sum
  // of weights) covers rows [row_begin,row_end) of the 6*h face-major rows;
  // the caller finishes with oracle_sh9_finish.
  void oracle_sh9_partial(void const *level0, int format, int w, int h, int row_begin, int row_end, double *partial)
  {
    double acc[28];
    for(int k = 0; k < 28; ++k)
      acc[k] = 0;

    for(int row = row_begin; row < row_end; ++row)
    {
      int face = row / h;
      int y = row % h;

      for(int x = 0; x < w; ++x)
      {
        double u = 2 * (x + 0.5)/w - 1;
        double v = 2 * (y + 0.5)/h - 1;

        double d[3];
        face_dir_d(face, u, v, d);
        double inv = 1.0 / std::sqrt(d[0]*d[0] + d[1]*d[1] + d[2]*d[2]);
        double rx = d[0]*inv, ry = d[1]*inv, rz = d[2]*inv;

        // project.comp:56-60
        double x0 = u - 1.0/w, y0 = v - 1.0/h, x1 = u + 1.0/w, y1 = v + 1.0/h;
        double weight = std::atan2(x0*y0, std::sqrt(x0*x0 + y0*y0 + 1)) - std::atan2(x0*y1, std::sqrt(x0*x0 + y1*y1 + 1))
                      - std::atan2(x1*y0, std::sqrt(x1*x1 + y0*y0 + 1)) + std::atan2(x1*y1, std::sqrt(x1*x1 + y1*y1 + 1));

        // project.comp:62 — the level-0 texel fetched at its own centre
        size_t idx = ((size_t)face * h + y) * w + x;
        double color[3];
        if (format == 0)
        {
          C4 c = rgbe_decode(((uint32_t const *)level0)[idx]);
          color[0] = c.r; color[1] = c.g; color[2] = c.b;
        }
        else
        {
          float const *p = (float const *)level0 + 4*idx;
          color[0] = p[0]; color[1] = p[1]; color[2] = p[2];
        }

        // project.comp:64-92
        double basis[9] =
        {
          0.282095,
          0.488603 * ry,
          0.488603 * rz,
          0.488603 * rx,
          1.092548 * rx * ry,
          1.092548 * ry * rz,
          0.315392 * (3 * rz*rz - 1),
          1.092548 * rz * rx,
          0.546274 * (rx*rx - ry*ry),
        };

        for(int k = 0; k < 9; ++k)
          for(int c = 0; c < 3; ++c)
            acc[3*k + c] += weight * color[c] * basis[k];

        acc[27] += weight;
      }
    }

    for(int k = 0; k < 28; ++k)
      partial[k] = acc[k];
  }

  // project.comp:99-105
  void oracle_sh9_finish(double const *partial, double *sh)
  {
    const double pi = 3.1415926535897932384626433832795;
    for(int k = 0; k < 27; ++k)
      sh[k] = partial[k] * (4*pi / partial[27]);
  }

  void oracle_project_sh9(void const *level0, int format, int w, int h, double *sh)
  {
    double partial[28];
    oracle_sh9_partial(level0, format, w, h, 0, 6*h, partial);
    oracle_sh9_finish(partial, sh);
  }

  // Irradiance from SH9 for unit normals: data/lighting.inc:351-366 (the band
  // factors pi, 2pi/3, pi/4 as the shader's literals) with the max(.,0) of :371;
  // distance attenuation (:368-371) is a property of the placed probe, not of the bake.
  void oracle_sh9_irradiance(double const *sh, float const *normals, size_t count, float *rgb)
  {
    for(size_t i = 0; i < count; ++i)
    {
      double nx = normals[3*i + 0], ny = normals[3*i + 1], nz = normals[3*i + 2];

      double L[9] =
      {
        3.141593 * 0.282095,
        2.094395 * 0.488603 * ny,
        2.094395 * 0.488603 * nz,
        2.094395 * 0.488603 * nx,
        0.785398 * 1.092548 * nx * ny,
        0.785398 * 1.092548 * ny * nz,
        0.785398 * 0.315392 * (3 * nz*nz - 1),
        0.785398 * 1.092548 * nz * nx,
        0.785398 * 0.546274 * (nx*nx - ny*ny),
      };

      for(int c = 0; c < 3; ++c)
      {
        double e = 0;
        for(int k = 0; k < 9; ++k)
          e += L[k] * sh[3*k + c];

        rgb[3*i + c] = (float)std::max(e, 0.0);
      }
    }
  }

  // tools/ibl.cpp:292-308 (samples parameterised; reference = 1024); optional fp32 lut out
  void oracle_pack_envbrdf(int width, int height, int samples, uint32_t *bits, float *f32, int threads)
  {
    parallel_rows(0, height, threads, [&](int y)
    {
      for(int x = 0; x < width; ++x)
      {
        float NdotV = (x + 0.5f) / width;
        float roughness = (y + 0.5f) / height;

        float lut[3];
        integrate(roughness, NdotV, samples, lut);

        size_t idx = (size_t)y * width + x;

        if (bits)
          bits[idx] = rgbe_encode(lut[0], lut[1], lut[2]);

        if (f32)
        {
          f32[3*idx + 0] = lut[0];
          f32[3*idx + 1] = lut[1];
          f32[3*idx + 2] = lut[2];
        }
      }
    });
  }

  // tools/ibl.cpp:312-329
  void oracle_pack_watercolor(float const *deepcolor, float const *shallowcolor, float depthscale, float const *fresnelcolor, float fresnelbias, float fresnelpower, int width, int height, uint32_t *bits)
  {
    uint32_t *dst = bits;

    for(int y = 0; y < height; ++y)
    {
      for(int x = 0; x < width; ++x)
      {
        float scale = (x + 0.5f) / width;
        float facing = (y + 0.5f) / height;
        float fresnel = clampf(fresnelbias + std::pow(facing, fresnelpower), 0.0f, 1.0f);
        float depth = clampf(1 - std::exp2(-depthscale * scale * 100.0f), 0.0f, 1.0f);

        float out[3];
        for(int c = 0; c < 3; ++c)
        {
          float color = lerp1(shallowcolor[c], deepcolor[c], depth);
          out[c] = lerp1(color, fresnelcolor[c], fresnel);
        }

        *dst++ = rgbe_encode(out[0], out[1], out[2]);
      }
    }
  }

  // ---- equirect HDR image -> cube level 0: tools/hdr.cpp:26-74, 173-359 ----------

  namespace
  {
    struct Image
    {
      int width, height;
      C4 const *bits; // HDRImage::bits, tools/hdr.h:24

      C4 texel(int i, int j) const { return bits[(size_t)j * width + i]; }

      // hdr.cpp:33-40 — bilinear with wrap in both directions
      C4 sample(float tx, float ty) const
      {
        float i, j;
        float u = std::modf(fmod2(tx * width - 0.5f, (float)width), &i);
        float v = std::modf(fmod2(ty * height - 0.5f, (float)height), &j);

        int i0 = (int)i, j0 = (int)j;
        int i1 = (i0 + 1) % width, j1 = (j0 + 1) % height;

        return lerp4(lerp4(texel(i0, j0), texel(i1, j0), u), lerp4(texel(i0, j1), texel(i1, j1), u), v);
      }

      // hdr.cpp:44-60 — box filter of bilinear taps over `area`; the fp32 loop
      // counters are part of the behaviour (they decide the tap count)
      C4 sample_area(float tx, float ty, float ax, float ay) const
      {
        C4 sum = { 0, 0, 0, 0 };
        float totalweight = 0;

        for(float y = ty - 0.5f*ay + 0.5f/height, yend = ty + 0.5f*ay; y < yend; y += 1.0f/height)
        {
          for(float x = tx - 0.5f*ax + 0.5f/width, xend = tx + 0.5f*ax; x < xend; x += 1.0f/width)
          {
            C4 c = sample(x, y);
            sum.r += c.r; sum.g += c.g; sum.b += c.b; sum.a += c.a;

            totalweight += 1.0f;
          }
        }

        return { sum.r / totalweight, sum.g / totalweight, sum.b / totalweight, sum.a / totalweight };
      }

      // hdr.cpp:71-74 — equirect lookup (note the reference's `pi/2 +` offset on u)
      C4 sample_dir(V3 const &d, float ax, float ay) const
      {
        return sample_area(kPi/2 + std::atan2(d.x, -d.z) / (2*kPi), std::acos(d.y) / kPi, ax, ay);
      }
    };

    inline uint32_t blend3(uint32_t a, uint32_t b, uint32_t c)
    {
      C4 ca = rgbe_decode(a), cb = rgbe_decode(b), cc = rgbe_decode(c);
      return rgbe_encode(0.3f * ca.r + 0.4f * cb.r + 0.3f * cc.r, 0.3f * ca.g + 0.4f * cb.g + 0.3f * cc.g, 0.3f * ca.b + 0.4f * cb.b + 0.3f * cc.b);
    }
  }

  // tools/assetpacker.cpp:548-572 — 2x2 box mips of an rgbe image
  void oracle_image_buildmips_rgbe(int width, int height, int layers, int levels, uint32_t *bits)
  {
    uint32_t *src = bits;
    uint32_t *dst = src + (size_t)width * height * layers;

    for(int level = 1; level < levels; ++level)
    {
      for(int layer = 0; layer < layers; ++layer)
      {
        for(int y = 0; y < (height >> 1); ++y)
        {
          for(int x = 0; x < (width >> 1); ++x)
          {
            C4 t[4] = { rgbe_decode(src[(2*y)*width + 2*x]), rgbe_decode(src[(2*y)*width + 2*x + 1]), rgbe_decode(src[(2*y + 1)*width + 2*x]), rgbe_decode(src[(2*y + 1)*width + 2*x + 1]) };

            *dst++ = rgbe_encode((t[0].r + t[1].r + t[2].r + t[3].r) / 4, (t[0].g + t[1].g + t[2].g + t[3].g) / 4, (t[0].b + t[1].b + t[2].b + t[3].b) / 4);
          }
        }

        src += (size_t)width * height;
      }

      width /= 2;
      height /= 2;
    }
  }

  // tools/hdr.cpp:173-318 — 0.3/0.4/0.3 blend across the twelve cube edges.  Each
  // entry is one of the reference's twelve loops, in order: texel k of the loop
  // reads a (inner) and b (edge) on the first face and c (edge), d (inner) on the
  // second, then writes b' = blend(a,b,c) and c' = blend(b,c,d) with the OLD b, c.
  // Later loops see the texels written by earlier ones (the corners).
  void oracle_image_blend_edges(int width, int height, int levels, uint32_t *bits)
  {
    uint32_t *img = bits;

    for(int level = 0; level < levels && width > 1 && height > 1; ++level)
    {
      int w = width, h = height;
      auto tex = [=](int x, int y, int z) -> uint32_t& { return img[(size_t)z*w*h + (size_t)y*w + x]; };

      auto run = [&](int count, auto coords)
      {
        for(int k = 0; k < count; ++k)
        {
          int c[4][3];
          coords(k, c);
          uint32_t a = tex(c[0][0], c[0][1], c[0][2]), b = tex(c[1][0], c[1][1], c[1][2]);
          uint32_t cc = tex(c[2][0], c[2][1], c[2][2]), d = tex(c[3][0], c[3][1], c[3][2]);
          tex(c[1][0], c[1][1], c[1][2]) = blend3(a, b, cc);
          tex(c[2][0], c[2][1], c[2][2]) = blend3(b, cc, d);
        }
      };

      // hdr.cpp:181-223: the four vertical seams around the horizon (right edge of fa -> left edge of fb)
      int ring[4][2] = { { 4, 0 }, { 0, 5 }, { 5, 1 }, { 1, 4 } };
      for(auto &seam : ring)
      {
        int fa = seam[0], fb = seam[1];
        run(h, [=](int k, int c[4][3]) {
          int v[4][3] = { { w - 2, k, fa }, { w - 1, k, fa }, { 0, k, fb }, { 1, k, fb } };
          memcpy(c, v, sizeof(v));
        });
      }

      // hdr.cpp:225-234: bottom row of front (4) -> top row of up (3)
      run(w, [=](int k, int c[4][3]) { int v[4][3] = { { k, h - 2, 4 }, { k, h - 1, 4 }, { k, 0, 3 }, { k, 1, 3 } }; memcpy(c, v, sizeof(v)); });
      // hdr.cpp:236-245: bottom row of up (3) -> bottom row of back (5), mirrored
      run(w, [=](int k, int c[4][3]) { int v[4][3] = { { k, h - 2, 3 }, { k, h - 1, 3 }, { w - 1 - k, h - 1, 5 }, { w - 1 - k, h - 2, 5 } }; memcpy(c, v, sizeof(v)); });
      // hdr.cpp:247-256: top row of back (5) -> top row of down (2), mirrored
      run(w, [=](int k, int c[4][3]) { int v[4][3] = { { k, 1, 5 }, { k, 0, 5 }, { w - 1 - k, 0, 2 }, { w - 1 - k, 1, 2 } }; memcpy(c, v, sizeof(v)); });
      // hdr.cpp:258-267: bottom row of down (2) -> top row of front (4)
      run(w, [=](int k, int c[4][3]) { int v[4][3] = { { k, h - 2, 2 }, { k, h - 1, 2 }, { k, 0, 4 }, { k, 1, 4 } }; memcpy(c, v, sizeof(v)); });

      int m = std::min(w, h);
      // hdr.cpp:269-278: bottom row of right (0) -> right column of up (3)
      run(m, [=](int k, int c[4][3]) { int v[4][3] = { { k, h - 2, 0 }, { k, h - 1, 0 }, { w - 1, k, 3 }, { w - 2, k, 3 } }; memcpy(c, v, sizeof(v)); });
      // hdr.cpp:280-289: left column of up (3) -> bottom row of left (1), mirrored
      run(m, [=](int k, int c[4][3]) { int v[4][3] = { { 1, k, 3 }, { 0, k, 3 }, { w - 1 - k, h - 1, 1 }, { w - 1 - k, h - 2, 1 } }; memcpy(c, v, sizeof(v)); });
      // hdr.cpp:291-300: top row of left (1) -> left column of down (2)
      run(m, [=](int k, int c[4][3]) { int v[4][3] = { { k, 1, 1 }, { k, 0, 1 }, { 0, k, 2 }, { 1, k, 2 } }; memcpy(c, v, sizeof(v)); });
      // hdr.cpp:302-311: right column of down (2) -> top row of right (0), mirrored
      run(m, [=](int k, int c[4][3]) { int v[4][3] = { { w - 2, k, 2 }, { w - 1, k, 2 }, { w - 1 - k, 0, 0 }, { w - 1 - k, 1, 0 } }; memcpy(c, v, sizeof(v)); });

      img += (size_t)width * height * 6;

      width /= 2;
      height /= 2;
    }
  }

  // tools/hdr.cpp:331-359 — `pixels` is imgw*imgh RGBA fp32
  void oracle_image_pack_cube(int imgw, int imgh, float const *pixels, int width, int height, int levels, uint32_t *bits)
  {
    Image image = { imgw, imgh, reinterpret_cast<C4 const *>(pixels) };

    Xform rot[6];
    face_rotations(rot);

    float ax = 1.0f / std::min(4 * width, imgw);
    float ay = 1.0f / std::min(2 * height, imgh);

    parallel_rows(0, 6 * height, 0, [&](int row)
    {
      int face = row / height;
      int y = row % height;

      for(int x = 0; x < width; ++x)
      {
        C4 c = image.sample_dir(texel_direction(rot[face], x, y, width, height), ax, ay);

        bits[(size_t)row * width + x] = rgbe_encode(c.r, c.g, c.b);
      }
    });

    // hdr.cpp:322-327
    oracle_image_buildmips_rgbe(width, height, 6, levels, bits);
    oracle_image_blend_edges(width, height, levels, bits);
  }

  // tools/ibl.cpp:283-288
  void oracle_image_pack_cube_ibl(int imgw, int imgh, float const *pixels, int width, int height, int levels, int samples, uint32_t *bits, int threads)
  {
    oracle_image_pack_cube(imgw, imgh, pixels, width, height, 1, bits);
    oracle_buildmips_cube_ibl(width, height, levels, samples, bits, nullptr, threads);
  }

  int oracle_max_threads() { return (int)std::max(1u, std::thread::hardware_concurrency()); }
}
