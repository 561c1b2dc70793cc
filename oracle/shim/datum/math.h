// TEST INFRASTRUCTURE — replaces the reference's include/datum/math.h umbrella
// for the oracle/_ref build.  It pulls in the reference's OWN, unmodified
// src/math/{vec,color,transform}.h (found through -I$(REF)/src/math, so the
// rgbe codec and Transform algebra under test are the reference's code) on top
// of the leap stand-in under oracle/shim/leap, and skips the geometry headers
// (rect/bound/frustum/attenuation/perlin) that the IBL path never uses.
#pragma once

#include <leap/lml/vector.h>
#include <leap/lml/quaternion.h>
#include <vec.h>
#include <color.h>
#include <transform.h>

#include <string>
#include <vector>
#include <sstream>
#include <cstring>
#include <cstdlib>
#include <stdexcept>

namespace lml
{
  // tools/assetpacker.h only names Bound3 in declarations we never call
  struct Bound3 { Vec3 min, max; };
}

// leap/util.h string helpers used by tools/hdr.cpp:88-118 (load_hdr header parsing)
namespace leap
{
  using lml::pi;
  using lml::clamp;
  using lml::lerp;
  using lml::fmod2;
  using lml::frac;

  inline std::string trim(std::string const &str, const char *characters = " \t\r\n")
  {
    auto i = str.find_first_not_of(characters);
    auto j = str.find_last_not_of(characters);
    return (i == std::string::npos) ? std::string() : str.substr(i, j - i + 1);
  }

  inline std::vector<std::string> split(std::string const &str, const char *delimiters = " \t\r\n")
  {
    std::vector<std::string> result;
    size_t i = 0;
    while (i < str.size())
    {
      auto j = str.find_first_of(delimiters, i);
      if (j == std::string::npos)
        j = str.size();
      if (j > i)
        result.push_back(str.substr(i, j - i));
      i = j + 1;
    }
    return result;
  }

  inline int stricmp(std::string const &lhs, const char *rhs) { return strcasecmp(lhs.c_str(), rhs); }

  template<typename T> T ato(std::string const &str)
  {
    T value = T();
    std::istringstream(str) >> value;
    return value;
  }
}
