// datum_b200 host shim — HDRImage buffer and the Radiance .hdr loader.
//
// load_hdr follows the behaviour of the reference's tools/hdr.cpp:78-169: header
// lines until the resolution line ("-Y h +X w" only), optional EXPOSURE, then one
// new-style RLE scanline per row (2, 2, width hi, width lo; four planes of runs),
// pixels decoded as (mantissa/255) * 2^(e-128) with alpha 1.  It is host file I/O,
// not part of the GPU hot path.

#include "hdr.h"

#include <cmath>
#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <strings.h>

HDRImage::HDRImage(int width, int height, lml::Color4 const &color)
  : width(width), height(height)
{
  bits = std::vector<lml::Color4>((size_t)width * height, color);
}

namespace
{
  std::string trimmed(std::string const &s)
  {
    size_t a = s.find_first_not_of(" \t\r\n");
    size_t b = s.find_last_not_of(" \t\r\n");
    return (a == std::string::npos) ? std::string() : s.substr(a, b - a + 1);
  }

  std::vector<std::string> fields_of(std::string const &s, const char *delims)
  {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size())
    {
      size_t j = s.find_first_of(delims, i);
      if (j == std::string::npos)
        j = s.size();
      if (j > i)
        out.push_back(s.substr(i, j - i));
      i = j + 1;
    }
    return out;
  }

  bool starts_with_nocase(std::string const &s, const char *prefix, size_t n)
  {
    return s.size() >= n && strncasecmp(s.c_str(), prefix, n) == 0;
  }
}

HDRImage load_hdr(std::string const &path)
{
  HDRImage image = {};
  image.width = 0;
  image.height = 0;

  std::ifstream fin(path, std::ios_base::in | std::ios_base::binary);
  if (!fin)
    throw std::runtime_error("Unable to open file: " + path);

  std::string buffer;

  while (std::getline(fin, buffer))
  {
    std::string line = trimmed(buffer);

    if (line.empty() || line[0] == '#')
      continue;

    if (starts_with_nocase(line, "format", 6))
    {
      auto fields = fields_of(line, "=");
      if (fields.size() != 2 && fields[1] != "32-bit_rle_rgbe") // same (lenient) test as hdr.cpp:99
        throw std::runtime_error("Unsupported hdr file format");
    }

    if (starts_with_nocase(line, "exposure", 8))
    {
      auto fields = fields_of(line, "=");
      if (fields.size() > 1)
        image.exposure = std::strtof(fields[1].c_str(), nullptr);
    }

    if (line[0] == '-' || line[0] == '+')
    {
      auto fields = fields_of(line, " \t");

      if (fields.size() < 4 || fields[0] != "-Y" || fields[2] != "+X")
        throw std::runtime_error("Unsupported hdr file dimensions");

      image.width = std::atoi(fields[3].c_str());
      image.height = std::atoi(fields[1].c_str());

      break;
    }
  }

  if (image.width <= 0 || image.height <= 0 || image.width >= 32768)
    throw std::runtime_error("hdr parse error");

  image.bits.resize((size_t)image.width * image.height);

  std::vector<uint8_t> planes((size_t)4 * image.width);

  for(int y = 0; y < image.height; ++y)
  {
    uint8_t head[4];
    fin.read((char*)head, 4);

    if (!fin || head[0] != 2 || head[1] != 2 || ((head[2] << 8) | head[3]) != image.width)
      throw std::runtime_error("hdr parse error");

    for(int k = 0; k < 4; ++k)
    {
      uint8_t *plane = &planes[(size_t)k * image.width];
      int position = 0;

      while (position < image.width)
      {
        int count = fin.get();
        if (count < 0)
          throw std::runtime_error("hdr parse error");

        if (count > 128)
        {
          count -= 128;
          int value = fin.get();
          if (value < 0 || position + count > image.width)
            throw std::runtime_error("hdr parse error");
          for(int i = 0; i < count; ++i)
            plane[position + i] = (uint8_t)value;
        }
        else
        {
          if (position + count > image.width)
            throw std::runtime_error("hdr parse error");
          fin.read((char*)plane + position, count);
        }

        position += count;
      }
    }

    for(int x = 0; x < image.width; ++x)
    {
      float scale = std::exp2((float)planes[(size_t)3 * image.width + x] - 128.0f);
      float r = planes[x] / 255.0f;
      float g = planes[(size_t)image.width + x] / 255.0f;
      float b = planes[(size_t)2 * image.width + x] / 255.0f;

      image.bits[(size_t)y * image.width + x] = lml::Color4(r * scale, g * scale, b * scale, 1.0f);
    }
  }

  return image;
}
