// datum_b200 — device/host math shared by the IBL kernels.
//
// Everything here is `__host__ __device__` so the same expressions can be
// compiled by g++ for the CPU-side unit checks in tests/ (tests/emu/) — the
// product only ever runs them on the GPU.
//
// Reference behaviour being reproduced (paths relative to /root/reference):
//   rgbe codec ................ src/math/color.h:154-172
//   face rotations ............ tools/ibl.cpp:253-261, src/math/transform.h:65-68,173-178
//   texel direction ........... tools/ibl.cpp:269
//   GGX tangent frame ......... tools/ibl.cpp:123-125
//   cube face select + uv ..... tools/ibl.cpp:43-88
//   per-face bilinear ......... tools/ibl.cpp:34-41
#pragma once

#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define IBL_HD __host__ __device__ __forceinline__
#else
#define IBL_HD inline
#endif

namespace ibl
{
  struct Vec3f { float x, y, z; };
  struct Quatf { float w, x, y, z; };

  // ---- exactly-rounded fp32 primitives (no FMA contraction): used where the
  //      reference's own rounding decides a discrete outcome (which tangent
  //      frame, which exponent), never in the per-sample loop ----
#if defined(__CUDA_ARCH__)
  IBL_HD float mul_rn(float a, float b) { return __fmul_rn(a, b); }
  IBL_HD float add_rn(float a, float b) { return __fadd_rn(a, b); }
  IBL_HD float sub_rn(float a, float b) { return __fsub_rn(a, b); }
  IBL_HD float div_rn(float a, float b) { return __fdiv_rn(a, b); }
  IBL_HD float sqrt_rn(float a) { return __fsqrt_rn(a); }
  IBL_HD float rcp_fast(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
  IBL_HD uint32_t f2u(float f) { return __float_as_uint(f); }
  IBL_HD float u2f(uint32_t u) { return __uint_as_float(u); }
#else
  IBL_HD float mul_rn(float a, float b) { volatile float r = a * b; return r; }
  IBL_HD float add_rn(float a, float b) { volatile float r = a + b; return r; }
  IBL_HD float sub_rn(float a, float b) { volatile float r = a - b; return r; }
  IBL_HD float div_rn(float a, float b) { volatile float r = a / b; return r; }
  IBL_HD float sqrt_rn(float a) { return std::sqrt(a); }
  IBL_HD float rcp_fast(float a) { return 1.0f / a; }
  IBL_HD uint32_t f2u(float f) { uint32_t u; __builtin_memcpy(&u, &f, 4); return u; }
  IBL_HD float u2f(uint32_t u) { float f; __builtin_memcpy(&f, &u, 4); return f; }
#endif

  // ---- quaternion sandwich t * (0,v) * conj(t) with the reference's term order ----

  IBL_HD Quatf qmul_rn(Quatf a, Quatf b)
  {
    Quatf r;
    r.w = sub_rn(sub_rn(sub_rn(mul_rn(a.w, b.w), mul_rn(a.x, b.x)), mul_rn(a.y, b.y)), mul_rn(a.z, b.z));
    r.x = sub_rn(add_rn(add_rn(mul_rn(a.w, b.x), mul_rn(a.x, b.w)), mul_rn(a.y, b.z)), mul_rn(a.z, b.y));
    r.y = sub_rn(add_rn(add_rn(mul_rn(a.w, b.y), mul_rn(a.y, b.w)), mul_rn(a.z, b.x)), mul_rn(a.x, b.z));
    r.z = sub_rn(add_rn(add_rn(mul_rn(a.w, b.z), mul_rn(a.z, b.w)), mul_rn(a.x, b.y)), mul_rn(a.y, b.x));
    return r;
  }

  // Transform::rotation(axis, angle) * v for a pure rotation `q` (dual part zero):
  // the dual of (t*p)*conj(t) is (q*(0,v))*conj(q); the zero-dual terms of
  // transform.h:161-178 only ever add exact zeros.
  IBL_HD Vec3f rotate_rn(Quatf q, Vec3f v)
  {
    Quatf p = { 0.0f, v.x, v.y, v.z };
    Quatf c = { q.w, -q.x, -q.y, -q.z };
    Quatf r = qmul_rn(qmul_rn(q, p), c);
    return Vec3f{ r.x, r.y, r.z };
  }

  // normalise(Vec3(2(x+.5)/wd - 1, 2(y+.5)/hd - 1, -1)) rotated into face space, ibl.cpp:269
  IBL_HD Vec3f texel_normal(Quatf q, int x, int y, int wd, int hd)
  {
    float ex = sub_rn(div_rn(mul_rn(2.0f, add_rn((float)x, 0.5f)), (float)wd), 1.0f);
    float ey = sub_rn(div_rn(mul_rn(2.0f, add_rn((float)y, 0.5f)), (float)hd), 1.0f);
    float ez = -1.0f;
    float n = sqrt_rn(add_rn(add_rn(mul_rn(ex, ex), mul_rn(ey, ey)), mul_rn(ez, ez)));
    Vec3f v = { div_rn(ex, n), div_rn(ey, n), div_rn(ez, n) };
    return rotate_rn(q, v);
  }

  // tangent frame of importancesample_ggx, ibl.cpp:123-125
  IBL_HD void tangent_frame(Vec3f N, Vec3f &T, Vec3f &B)
  {
    Vec3f up = (fabsf(N.z) < 0.999f) ? Vec3f{ 0, 0, 1 } : Vec3f{ 1, 0, 0 };
    Vec3f c = { sub_rn(mul_rn(up.y, N.z), mul_rn(up.z, N.y)), sub_rn(mul_rn(up.z, N.x), mul_rn(up.x, N.z)), sub_rn(mul_rn(up.x, N.y), mul_rn(up.y, N.x)) };
    float n = sqrt_rn(add_rn(add_rn(mul_rn(c.x, c.x), mul_rn(c.y, c.y)), mul_rn(c.z, c.z)));
    T = Vec3f{ div_rn(c.x, n), div_rn(c.y, n), div_rn(c.z, n) };
    B = Vec3f{ sub_rn(mul_rn(N.y, T.z), mul_rn(N.z, T.y)), sub_rn(mul_rn(N.z, T.x), mul_rn(N.x, T.z)), sub_rn(mul_rn(N.x, T.y), mul_rn(N.y, T.x)) };
  }

  // ---- rgbe codec ----

  // color.h:164-172.  (m/511) * 2^e with e = E-15: both factors and the product
  // are exact in fp32 except the single rounding of m/511.
  IBL_HD void rgbe_decode(uint32_t c, float &r, float &g, float &b)
  {
    float s = u2f((((c >> 27) & 0x1Fu) + 112u) << 23); // 2^(E-15)
    r = div_rn((float)((c >> 0) & 0x1FFu), 511.0f) * s;
    g = div_rn((float)((c >> 9) & 0x1FFu), 511.0f) * s;
    b = div_rn((float)((c >> 18) & 0x1FFu), 511.0f) * s;
  }

  // color.h:154-162, written to give the same word as the reference for the
  // same fp32 inputs: log2f decides the exponent exactly as there (including
  // the roll-over just below a power of two), the scaling by 2^-e is exact.
  IBL_HD uint32_t rgbe_encode(float cr, float cg, float cb)
  {
    float r = fmaxf(0.0f, fminf(cr, 65408.0f));
    float g = fmaxf(0.0f, fminf(cg, 65408.0f));
    float b = fmaxf(0.0f, fminf(cb, 65408.0f));
    float e = fmaxf(-16.0f, floorf(log2f(fmaxf(r, fmaxf(g, b))))) + 1.0f;
    int ei = (int)e;                                  // -15 .. 16
    float inv = u2f((uint32_t)(127 - ei) << 23);      // 2^-e, exact
    uint32_t mr = (uint32_t)roundf(mul_rn(mul_rn(r, inv), 511.0f));
    uint32_t mg = (uint32_t)roundf(mul_rn(mul_rn(g, inv), 511.0f));
    uint32_t mb = (uint32_t)roundf(mul_rn(mul_rn(b, inv), 511.0f));
    return ((uint32_t)((ei + 15) & 0xFF) << 27) | ((mr & 0xFFFFu) << 0) | ((mg & 0xFFFFu) << 9) | ((mb & 0xFFFFu) << 18);
  }

  // ---- per-sample cube addressing ----

  // Magic constant for round-to-nearest-integer through the fp32 adder: adding
  // 1.5*2^23 leaves the integer in the low mantissa bits.
  constexpr float kMagic = 12582912.0f;
  constexpr float kRcpShrink = 0.99999976158142089844f; // 1 - 2^-22
  constexpr uint32_t kMagicBits = 0x4B400000u;

  struct LevelGeom
  {
    int ws, hs;          // source level size
    float hw, hh;        // 0.5*(ws-1), 0.5*(hs-1): align-corners scale of ibl.cpp:37-38
    float hwm, hhm;      // hw - 0.5, hh - 0.5
    float inv_hw, inv_hh;
    uint32_t face_size;  // ws*hs records per face
    uint32_t bias;       // kMagicBits*(ws+1) mod 2^32, removed from the raw index

    // projective form (face_footprint_proj / cube_footprint_proj below)
    float hwm_magic, hhm_magic;   // hwm + kMagic, hhm + kMagic: exact while ws and hs are even
    float neg_ws;                 // -(float)ws
    float hw_shrunk, hh_shrunk;   // hw, hh times kRcpShrink: the scale of cube_select_unshrunk's quotients
    uint32_t bias_general;        // kMagicBits - hhm*ws: what cube_footprint_proj's raw index carries
  };

  IBL_HD LevelGeom make_level_geom(int ws, int hs)
  {
    LevelGeom g;
    g.ws = ws; g.hs = hs;
    g.hw = 0.5f * (float)(ws - 1); g.hh = 0.5f * (float)(hs - 1);
    g.hwm = g.hw - 0.5f; g.hhm = g.hh - 0.5f;
    g.inv_hw = 1.0f / g.hw; g.inv_hh = 1.0f / g.hh;
    g.face_size = (uint32_t)ws * (uint32_t)hs;
    g.bias = kMagicBits * (uint32_t)(ws + 1);
    g.hwm_magic = g.hwm + kMagic; g.hhm_magic = g.hhm + kMagic;
    g.neg_ws = -(float)ws;
    g.hw_shrunk = g.hw * kRcpShrink; g.hh_shrunk = g.hh * kRcpShrink;
    g.bias_general = kMagicBits - (uint32_t)((hs - 2) / 2) * (uint32_t)ws;
    return g;
  }

  // Direction -> (record index of the 2x2 footprint's top-left texel, du, dv)
  // with du = frac_u - 0.5, dv = frac_v - 0.5 in [-0.5, 0.5].
  //
  // Face choice is ibl.cpp:51-85's strict major axis; exact ties (undefined in
  // the reference) resolve x, then y, then z like the oracle.  uv follow
  // ibl.cpp:55-83 rewritten with one reciprocal:
  //   +-x: u = .5 + .5 z/x    v = .5 + .5 y/|x|
  //   +-y: u = .5 + .5 x/|y|  v = .5 + .5 z/y
  //   +-z: u = .5 - .5 x/z    v = .5 + .5 y/|z|
  // then fu = u*(ws-1) (align-corners), i = floor(fu).  floor is taken as
  // round-to-nearest of fu-0.5 through the magic add; when fu is an exact
  // integer this may pick i-1 with frac 1, which addresses the same bilinear
  // value.  |q| <= 1 keeps i in [0, ws-2], so the footprint never leaves the face.
  // face selection and the two face coordinates in [-1, 1] (qu = 2u - 1, qv = 2v - 1)
  template<bool SHRINK>
  IBL_HD void cube_select_impl(float Lx, float Ly, float Lz, float &qu, float &qv, uint32_t &face)
  {
    float ax = fabsf(Lx), ay = fabsf(Ly), az = fabsf(Lz);
    bool px = ax >= fmaxf(ay, az);
    bool py = !px && (ay >= az);

    float major = px ? Lx : (py ? Ly : Lz);
    float un = px ? Lz : Lx;
    float vn = py ? Lz : Ly;

    // rcp.approx may round up by one ulp: on an exact tie |un| == |major| (a direction on a cube
    // edge) the quotient would then exceed 1 and the footprint would start one texel outside the
    // face.  Shrinking the reciprocal by two ulps keeps |q| <= 1; it moves a footprint by at most
    // 2.4e-7 of the face width.  (SHRINK = false: the caller multiplies by a shrunk scale instead.)
    float r = SHRINK ? rcp_fast(major) * kRcpShrink : rcp_fast(major);
    float ar = fabsf(r);
    float ru = px ? r : (py ? ar : -r);
    float rv = py ? r : ar;

    // face index: x -> 0/1, y -> 3/2, z -> 5/4 for positive/negative major
    uint32_t neg = f2u(major) >> 31;
    face = px ? neg : (py ? 3u - neg : 5u - neg);

    qu = un * ru;
    qv = vn * rv;
  }

  IBL_HD void cube_select(float Lx, float Ly, float Lz, float &qu, float &qv, uint32_t &face)
  {
    cube_select_impl<true>(Lx, Ly, Lz, qu, qv, face);
  }

  // |qu|, |qv| <= 1 + 2^-23 here: to be scaled by LevelGeom::hw_shrunk / hh_shrunk
  IBL_HD void cube_select_unshrunk(float Lx, float Ly, float Lz, float &qu, float &qv, uint32_t &face)
  {
    cube_select_impl<false>(Lx, Ly, Lz, qu, qv, face);
  }

  IBL_HD uint32_t cube_footprint(LevelGeom const &g, float Lx, float Ly, float Lz, float &du, float &dv)
  {
    float qu, qv;
    uint32_t face;
    cube_select(Lx, Ly, Lz, qu, qv, face);

    float fu = fmaf(qu, g.hw, g.hwm);
    float fv = fmaf(qv, g.hh, g.hhm);
    float mu = fu + kMagic;
    float mv = fv + kMagic;
    du = fu - (mu - kMagic);
    dv = fv - (mv - kMagic);

    return f2u(mv) * (uint32_t)g.ws + f2u(mu) + face * g.face_size - g.bias;
  }

  // ---- same-face fast path ----
  //
  // Face-local coordinates (a, b, m) of a world vector for cube face f, chosen so
  // that on face f (m > 0 major) the uv of ibl.cpp:55-83 read u = .5 + .5 a/m,
  // v = .5 + .5 b/m:
  //   0 (+x): ( z, y,  x)   1 (-x): (-z, y, -x)   2 (-y): (x, -z, -y)
  //   3 (+y): ( x, z,  y)   4 (-z): ( x, y, -z)   5 (+z): (-x, y,  z)
  IBL_HD Vec3f to_face_local(int face, Vec3f v)
  {
    switch (face)
    {
      case 0: return Vec3f{ v.z, v.y, v.x };
      case 1: return Vec3f{ -v.z, v.y, -v.x };
      case 2: return Vec3f{ v.x, -v.z, -v.y };
      case 3: return Vec3f{ v.x, v.z, v.y };
      case 4: return Vec3f{ v.x, v.y, -v.z };
      default: return Vec3f{ -v.x, v.y, v.z };
    }
  }

  IBL_HD Vec3f from_face_local(int face, Vec3f l)
  {
    switch (face)
    {
      case 0: return Vec3f{ l.z, l.y, l.x };
      case 1: return Vec3f{ -l.z, l.y, -l.x };
      case 2: return Vec3f{ l.x, -l.z, -l.y };
      case 3: return Vec3f{ l.x, l.z, l.y };
      case 4: return Vec3f{ l.x, l.y, -l.z };
      default: return Vec3f{ -l.x, l.y, l.z };
    }
  }

  // A reflected direction makes the angle acos(lz) with the normal (lz = NdotL of
  // the sample table).  It cannot leave the normal's own cube face while that
  // angle is smaller than the normal's angular distance to the nearest of the
  // four planes |a| = m, |b| = m bounding the face: sin(dist) = (m - |a|)/sqrt(2)
  // for a unit normal.  Samples with lz above the returned threshold therefore
  // need no face selection.  The margin absorbs the ~1e-7 non-orthonormality of
  // the fp32 frame.
  IBL_HD float same_face_threshold(Vec3f n_local)
  {
    float margin = n_local.z - fmaxf(fabsf(n_local.x), fabsf(n_local.y));
    margin = fmaxf(margin, 0.0f);
    return sqrtf(fmaxf(0.0f, 1.0f - 0.5f * margin * margin)) + 2e-6f;
  }

  // footprint on the texel's own face; la_s, lb_s are the a and b coordinates
  // pre-scaled by 0.5*(ws-1) and 0.5*(hs-1); face_base = face*face_size - bias
  IBL_HD uint32_t face_footprint(LevelGeom const &g, uint32_t face_base, float la_s, float lb_s, float lm, float &du, float &dv)
  {
    float r = rcp_fast(lm);
    float fu = fmaf(la_s, r, g.hwm);
    float fv = fmaf(lb_s, r, g.hhm);
    float mu = fu + kMagic;
    float mv = fv + kMagic;
    du = fu - (mu - kMagic);
    dv = fv - (mv - kMagic);

    return f2u(mv) * (uint32_t)g.ws + f2u(mu) + face_base;
  }

  // ---- projective form of both footprints (prefilter_dp_kernel) ----
  //
  // The sample table holds (X, Y) = (lx/lz, ly/lz): a direction is scale invariant (the face choice and
  // both quotients of ibl.cpp:51-85 are), so L' = X*T + Y*B + N costs two multiply-adds per component
  // instead of three.  On the texel's own face the align-corners offset goes into the frame rows,
  // a' = a*hw + hwm*m (fold_face_row), so that the texel coordinate is ONE product fu = a'/m that is
  // never rounded by itself: i comes out of fma(a', r, magic), the fraction out of fma(a', r, -i) — one
  // rounding each, three operations per coordinate instead of four.  The record index is formed in the
  // fp32 adder too: magic + i + j*ws is exact below 2^24 (ws*hs <= 2^22), its bits are the 32-bit
  // index with kMagicBits on top (the kernel moves the record pointer back by that once).
  // All of this is what face_footprint / cube_footprint compute, up to fp32 rounding.
  constexpr uint32_t kProjMaxFace = 1u << 22;

  IBL_HD bool proj_usable(int ws, int hs)
  {
    return ws >= 2 && hs >= 2 && (ws & 1) == 0 && (hs & 1) == 0 && (uint64_t)ws * (uint64_t)hs <= kProjMaxFace;
  }

  IBL_HD Vec3f fold_face_row(LevelGeom const &g, Vec3f l)
  {
    return Vec3f{ fmaf(g.hwm, l.z, l.x * g.hw), fmaf(g.hhm, l.z, l.y * g.hh), l.z };
  }

  IBL_HD Vec3f unfold_face_row(LevelGeom const &g, Vec3f s)
  {
    return Vec3f{ fmaf(-g.hwm, s.z, s.x) * g.inv_hw, fmaf(-g.hhm, s.z, s.y) * g.inv_hh, s.z };
  }

  // la, lb, lm from folded rows; returns kMagicBits + i + j*ws (face offset not included)
  IBL_HD uint32_t face_footprint_proj(LevelGeom const &g, float la, float lb, float lm, float &du, float &dv)
  {
    float r = rcp_fast(lm);
    float mu = fmaf(la, r, kMagic);
    float mv = fmaf(lb, r, kMagic);
    float niu = fmaf(mu, -1.0f, kMagic);      // -i, exact
    float niv = fmaf(mv, -1.0f, kMagic);
    du = fmaf(la, r, niu);
    dv = fmaf(lb, r, niv);
    return f2u(fmaf(niv, g.neg_ws, mu));
  }

  // any direction; returns kMagicBits + i + (j - hhm)*ws (face offset not included: see g.bias_general)
  IBL_HD uint32_t cube_footprint_proj(LevelGeom const &g, float Lx, float Ly, float Lz, float &du, float &dv, uint32_t &face)
  {
    float qu, qv;
    cube_select_unshrunk(Lx, Ly, Lz, qu, qv, face);

    float mu = fmaf(qu, g.hw_shrunk, g.hwm_magic);
    float mv = fmaf(qv, g.hh_shrunk, g.hhm_magic);
    float cu = fmaf(mu, -1.0f, g.hwm_magic);  // hwm - i, exact (hwm is an integer)
    float cv = fmaf(mv, -1.0f, g.hhm_magic);
    du = fmaf(qu, g.hw_shrunk, cu);
    dv = fmaf(qv, g.hh_shrunk, cv);
    return f2u(fmaf(cv, g.neg_ws, mu));
  }

  // footprint_weights with the right-hand column as a difference: w10 = v0 - w00 = (0.5 + du) * v0 up to
  // one rounding of the size of ulp(v0); never negative because u0 <= 1
  IBL_HD void footprint_weights_diff(float du, float dv, float wh, float nl, float w[4])
  {
    float u0 = 0.5f - du;
    float v1 = fmaf(dv, nl, wh), v0 = fmaf(-dv, nl, wh);
    w[0] = u0 * v0; w[2] = u0 * v1;
    w[1] = v0 - w[0]; w[3] = v1 - w[2];
  }

  // ---- same-face limits per azimuth sector (prefilter_dp_kernel) ----
  //
  // same_face_threshold above bounds the lobe angle whatever the azimuth.  But a sample leaves through ONE
  // edge, and only if it points towards it: with (X, Y) = (lx/lz, ly/lz) = rho * (cos phi, sin phi) — the
  // tangent-plane coordinates the projective table holds — the unnormalised direction is X*T + Y*B + N and
  //     a - m <= 0   <=>   X*(Ta - Tm) + Y*(Ba - Bm) <= Nm - Na          (edge +a; -a, +b, -b alike)
  // in face-local coordinates: rho * (p cos phi + q sin phi) <= d.  Over an azimuth sector the bracket is at
  // most |(p, q)| when (p, q) points into the sector, else its larger end value — often negative: no sample
  // of that sector ever crosses that edge.  The pair kernel gives every warp of a tile ONE sector of every
  // band (ibl_tables.h, build_sector_entries), so each warp gets its own count of same-face bands from the
  // limit of its sector: rho <= min over edges of d / max(bracket).  With (p, q) inside the sector this is
  // tan(angular distance to the edge plane), the isotropic bound; otherwise it is larger.
  // Sector k of kFrameSectors = 8 spans the azimuths [-pi + k*pi/4, -pi + (k+1)*pi/4].
  constexpr int kFrameSectors = 8;
  constexpr float kSectorMargin = 1e-5f;   // fp32 slop of the frame, the direction and the sector bounds

  IBL_HD void sector_rho_limits(Vec3f T, Vec3f B, Vec3f N, float out[kFrameSectors])
  {
    const float h = 0.70710678118654752f;
    const float ex[kFrameSectors + 1] = { -1.0f, -h, 0.0f, h, 1.0f, h, 0.0f, -h, -1.0f };
    const float ey[kFrameSectors + 1] = { 0.0f, -h, -1.0f, -h, 0.0f, h, 1.0f, h, 0.0f };

    // edges +a, -a, +b, -b of the face |a| <= m, |b| <= m
    const float p[4] = { T.x - T.z, -T.x - T.z, T.y - T.z, -T.y - T.z };
    const float q[4] = { B.x - B.z, -B.x - B.z, B.y - B.z, -B.y - B.z };
    const float d[4] = { N.z - N.x, N.z + N.x, N.z - N.y, N.z + N.y };

    for(int k = 0; k < kFrameSectors; ++k)
    {
      float limit = 3.0e38f;

      for(int e = 0; e < 4; ++e)
      {
        float c0 = ex[k] * p[e] + ey[k] * q[e];
        float c1 = ex[k + 1] * p[e] + ey[k + 1] * q[e];
        bool inside = (ex[k] * q[e] - ey[k] * p[e] >= 0.0f) && (ex[k + 1] * q[e] - ey[k + 1] * p[e] <= 0.0f);
        float most = inside ? sqrtf(p[e] * p[e] + q[e] * q[e]) : fmaxf(c0, c1);
        float room = d[e] - kSectorMargin;

        if (room <= 0.0f)
          limit = 0.0f;
        else if (most > 0.0f)
          limit = fminf(limit, room / (most * (1.0f + kSectorMargin)));
      }

      out[k] = limit * (1.0f - kSectorMargin);
    }
  }

  // ---- packed-texel accumulation ----
  //
  // A quad record holds the four rgbe words of a bilinear footprint, each rotated
  // right by 4 bits (pack_record_word) so that the exponent field E sits at bits
  // 23..27 and the blue mantissa at bits 14..22 — exactly where an fp32 keeps its
  // low exponent bits and top mantissa bits.  Each 9-bit mantissa m is turned
  // into the float
  //     bits = 0x20000000 | E<<23 | m<<14   ==   2^(E-63) * (1 + m/512)
  // with one or two logic ops and no int->float conversion; the "1 +" bias is
  // accumulated once per tap (acc[3]) and removed at the end:
  //     acc[c] - acc[3] = sum of w * 2^(E-63) * m_c/512.
  IBL_HD uint32_t pack_record_word(uint32_t rgbe) { return (rgbe >> 4) | (rgbe << 28); }

  // bit fields of a record word; the exponent offset (+64, bit 29) travels as a kernel
  // PARAMETER so that each extraction is a single three-input LOP3 of the form
  // (register & immediate) | register — with every constant immediate the compiler
  // would need two logic ops, with every constant in a register three register reads.
  constexpr uint32_t kMaskExpo = 0x0F800000u;     // E in place
  constexpr uint32_t kMaskExpMant = 0x0FFFC000u;  // E and the in-place (blue) mantissa
  constexpr uint32_t kMaskMant = 0x007FC000u;     // a mantissa moved to bits 14..22
  constexpr uint32_t kExpBias = 0x20000000u;      // exponent offset +64

  struct DecodeMasks
  {
    uint32_t bias; // kExpBias
  };

  IBL_HD DecodeMasks make_decode_masks()
  {
    DecodeMasks k;
    k.bias = kExpBias;
    return k;
  }

  IBL_HD void accumulate_tap(DecodeMasks const &k, uint32_t word, float w, float acc[4])
  {
    uint32_t eb = (word & kMaskExpo) | k.bias;
    float fb = u2f((word & kMaskExpMant) | k.bias);
    float fg = u2f(((word << 9) & kMaskMant) | eb);
    float fr = u2f((((word << 18) | (word >> 14)) & kMaskMant) | eb);
    acc[0] = fmaf(w, fr, acc[0]);
    acc[1] = fmaf(w, fg, acc[1]);
    acc[2] = fmaf(w, fb, acc[2]);
    acc[3] = fmaf(w, u2f(eb), acc[3]);
  }

  // (sum - bias) -> radiance: undo the 2^-48 exponent offset and 512 -> 511 mantissa scale
  constexpr float kAccScale = 281474976710656.0f * (512.0f / 511.0f); // 2^48 * 512/511

  // weights of the 2x2 footprint, pre-multiplied by the sample weight (wh = 0.5*NdotL)
  IBL_HD void footprint_weights(float du, float dv, float wh, float nl, float w[4])
  {
    float u1 = 0.5f + du, u0 = 0.5f - du;
    float v1 = fmaf(dv, nl, wh), v0 = fmaf(-dv, nl, wh);
    w[0] = u0 * v0; w[1] = u1 * v0; w[2] = u0 * v1; w[3] = u1 * v1;
  }

  // ---- "denormal mantissa" records (prefilter_dn.cu) ----
  //
  // E5B9G9R9 word (E 27..31, b 18..26, g 9..17, r 0..8) -> r<<23 | g<<14 | b<<5 | E.
  // A mantissa field is read as the fp32 number its bits spell under a zero exponent field:
  //     u2f(word >> 23)        = r * 2^-149
  //     u2f(word & kDnMaskG)   = g * 2^-149 * 2^14
  //     u2f(word & kDnMaskB)   = b * 2^-149 * 2^5
  // subnormals, exact, and exact again as FMA operands.  The shared exponent multiplies the
  // tap's weight instead: bits(w) + (E << 23) == bits(w * 2^E) for any normal w.
  constexpr uint32_t kDnMaskE = 0x0000001Fu;
  constexpr uint32_t kDnMaskB = 0x00003FE0u;
  constexpr uint32_t kDnMaskG = 0x007FC000u;

  // every entry of the sample table is multiplied by 2^64 (directions are scale invariant, weights
  // carry the factor) so that no weight * mantissa product falls below the normal range
  constexpr float kDnTableScale = 18446744073709551616.0f;

  IBL_HD uint32_t pack_dn_word(uint32_t w)
  {
    return ((w & 0x1FFu) << 23) | (((w >> 9) & 0x1FFu) << 14) | (((w >> 18) & 0x1FFu) << 5) | (w >> 27);
  }

  // one tap, scalar form of the kernel's accumulate(): acc[c] += w * 2^E * field_c
  IBL_HD void dn_accumulate_tap(uint32_t word, float w, float acc[3])
  {
    float ws = u2f(f2u(w) + ((word & kDnMaskE) << 23));
    acc[0] = fmaf(u2f(word >> 23), ws, acc[0]);
    acc[1] = fmaf(u2f(word & kDnMaskG), ws, acc[1]);
    acc[2] = fmaf(u2f(word & kDnMaskB), ws, acc[2]);
  }

  // The same on a word in the reference's own layout (E 27..31, b 18..26, g 9..17, r 0..8), for the
  // kernel that reads the source level directly: r and g are subnormals in place, b is brought below
  // the exponent field with one shift (bits 14..22), E comes down with one shift.  5 logic ops + 1
  // integer multiply-add per tap; re-laying the word first (pack_dn_word) cost 12.  Every product is
  // exact and the sums differ from dn_accumulate_tap's by exact powers of two per channel
  // (raw_channel_norms), so the results are bit-identical.
  IBL_HD void raw_accumulate_tap(uint32_t word, float w, uint32_t emul, float acc[3])
  {
    float ws = u2f(f2u(w) + (word >> 27) * emul);
    acc[0] = fmaf(u2f(word & 0x000001FFu), ws, acc[0]);
    acc[1] = fmaf(u2f(word & 0x0003FE00u), ws, acc[1]);
    acc[2] = fmaf(u2f((word >> 4) & 0x007FC000u), ws, acc[2]);
  }

  // sums -> radiance: radiance = (m/511) * 2^(E-15), sums hold m * 2^-149 * 2^(field position) * 2^64 * 2^E * weight
  IBL_HD void dn_channel_norms(double total_weight, float norm[3])
  {
    double base = ldexp(1.0, 149 - 64 - 15) / 511.0 / total_weight;
    norm[0] = (float)base;
    norm[1] = (float)ldexp(base, -14);
    norm[2] = (float)ldexp(base, -5);
  }

  IBL_HD void raw_channel_norms(double total_weight, float norm[3])
  {
    double base = ldexp(1.0, 149 - 64 - 15) / 511.0 / total_weight;
    norm[0] = (float)base;
    norm[1] = (float)ldexp(base, -9);
    norm[2] = (float)ldexp(base, -14);
  }
}
