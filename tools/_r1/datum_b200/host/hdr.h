// datum_b200 host shim — same declarations as the reference's tools/hdr.h:11-39.
#pragma once

#include "math.h"

#include <string>
#include <vector>

// tools/hdr.h:14-33.  The CPU-side sample() members of the reference are not part
// of the shim: the resample they implement runs on the GPU (datum_ibl_pack_cube).
class HDRImage
{
  public:
    HDRImage() = default;
    HDRImage(int width, int height, lml::Color4 const &color = { 0, 0, 0, 0 });

    int width;
    int height;
    float exposure = 1.0f;

    std::vector<lml::Color4> bits;
};

// tools/hdr.h:35 / tools/hdr.cpp:78-169: Radiance RLE .hdr loader (host I/O); throws std::runtime_error
HDRImage load_hdr(std::string const &path);

// tools/hdr.h:39 / tools/hdr.cpp:331-359.  Only levels == 1 is offered (the IBL path's use).
void image_pack_cube(HDRImage const &image, int width, int height, int levels, void *bits);
