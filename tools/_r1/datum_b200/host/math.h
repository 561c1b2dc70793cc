// datum_b200 host shim — the few lml types that appear in the reference's
// tools/ibl.h / tools/hdr.h signatures (src/math/vec.h:22-75, src/math/color.h:20-74).
//
// Inside the reference tree this header is NOT used: tools/ibl.cpp's replacement
// includes the reference's own "datum/math.h" (see INTEGRATION.md).  It exists so
// that the shim also builds and is testable on its own, with the same member
// names and memory layout (plain consecutive floats).
#pragma once

namespace lml
{
  struct Vec2 { float x, y; };

  struct Vec3 { float x, y, z; };

  struct Color3
  {
    Color3() = default;
    constexpr Color3(float r, float g, float b) : r(r), g(g), b(b) { }

    float r, g, b;
  };

  struct Color4
  {
    Color4() = default;
    constexpr Color4(float r, float g, float b, float a = 1.0f) : r(r), g(g), b(b), a(a) { }

    float r, g, b, a;
  };
}
