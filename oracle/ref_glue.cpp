// TEST INFRASTRUCTURE — C entry points around the UNMODIFIED reference
// tools/ibl.cpp and tools/hdr.cpp, which the Makefile compiles from where they
// lie under /root/reference into oracle/_ref/libdatum_ref_ibl.so.  Only tests/,
// __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs
// may load that library; nothing under datum_b200/ does.
//
// The one function restated here is image_buildmips_rgbe
// (tools/assetpacker.cpp:548-572): tools/hdr.cpp:324 needs the symbol, but the
// rest of assetpacker.cpp depends on leap's lz4 and cannot be built here.  The
// IBL path only ever calls it with levels == 1, for which it does nothing.

#include "ibl.h"          // reference tools/ibl.h (through -I$(REF)/tools)
#include "assetpacker.h"  // reference tools/assetpacker.h
#include <cstdint>
#include <cstring>

using namespace lml;

void image_buildmips_rgbe(int width, int height, int layers, int levels, void *bits)
{
  uint32_t *src = (uint32_t*)bits;
  uint32_t *dst = src + width * height * layers;

  for(int level = 1; level < levels; ++level)
  {
    for(int layer = 0; layer < layers; ++layer)
    {
      for(int y = 0; y < (height >> 1); ++y)
      {
        for(int x = 0; x < (width >> 1); ++x)
        {
          uint32_t const *row0 = src + (2*y)*width + 2*x;
          uint32_t const *row1 = row0 + width;

          *dst++ = rgbe((rgbe(row0[0]) + rgbe(row0[1]) + rgbe(row1[0]) + rgbe(row1[1])) / 4);
        }
      }

      src += width * height;
    }

    width /= 2;
    height /= 2;
  }
}

extern "C"
{
  // tools/ibl.cpp:242
  void ref_image_buildmips_cube_ibl(int width, int height, int levels, void *bits)
  {
    image_buildmips_cube_ibl(width, height, levels, bits);
  }

  // tools/ibl.cpp:283 — `pixels` is width*height RGBA fp32 (HDRImage::bits, tools/hdr.h:24)
  void ref_image_pack_cube_ibl(int imgwidth, int imgheight, float const *pixels, int width, int height, int levels, void *bits)
  {
    HDRImage image(imgwidth, imgheight);
    memcpy(image.bits.data(), pixels, sizeof(float) * 4 * imgwidth * imgheight);
    image_pack_cube_ibl(image, width, height, levels, bits);
  }

  // tools/hdr.cpp:331
  void ref_image_pack_cube(int imgwidth, int imgheight, float const *pixels, int width, int height, int levels, void *bits)
  {
    HDRImage image(imgwidth, imgheight);
    memcpy(image.bits.data(), pixels, sizeof(float) * 4 * imgwidth * imgheight);
    image_pack_cube(image, width, height, levels, bits);
  }

  // tools/ibl.cpp:292
  void ref_image_pack_envbrdf(int width, int height, void *bits)
  {
    image_pack_envbrdf(width, height, bits);
  }

  // tools/ibl.cpp:312
  void ref_image_pack_watercolor(float const *deepcolor, float const *shallowcolor, float depthscale, float const *fresnelcolor, float fresnelbias, float fresnelpower, int width, int height, void *bits)
  {
    image_pack_watercolor(Color3(deepcolor[0], deepcolor[1], deepcolor[2]), Color3(shallowcolor[0], shallowcolor[1], shallowcolor[2]), depthscale, Color3(fresnelcolor[0], fresnelcolor[1], fresnelcolor[2]), fresnelbias, fresnelpower, width, height, bits);
  }

  // src/math/color.h:154-172, 120-128
  uint32_t ref_rgbe_encode(float r, float g, float b) { return rgbe(Color4(r, g, b, 1.0f)); }
  void ref_rgbe_decode(uint32_t word, float *rgba) { Color4 c = rgbe(word); rgba[0] = c.r; rgba[1] = c.g; rgba[2] = c.b; rgba[3] = c.a; }
  void ref_srgba_decode(uint32_t argb, float *rgba) { Color4 c = srgba(argb); rgba[0] = c.r; rgba[1] = c.g; rgba[2] = c.b; rgba[3] = c.a; }

  // src/math/transform.h:65-68, 173-178 — the six face rotations of tools/ibl.cpp:253-261
  void ref_face_rotate(int face, float const *v, float *out)
  {
    Transform transforms[] =
    {
      Transform::rotation(Vec3(0, 1, 0), -pi<float>()/2),
      Transform::rotation(Vec3(0, 1, 0), pi<float>()/2),
      Transform::rotation(Vec3(1, 0, 0), -pi<float>()/2),
      Transform::rotation(Vec3(1, 0, 0), pi<float>()/2),
      Transform::rotation(Vec3(0, 1, 0), 0),
      Transform::rotation(Vec3(0, 1, 0), pi<float>()),
    };
    Vec3 r = transforms[face] * Vec3(v[0], v[1], v[2]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
  }
}
