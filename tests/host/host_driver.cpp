// TEST INFRASTRUCTURE — exercises the C++ host shim through the reference's own
// tools/ibl.h signatures (datum_b200/host/ibl.h), the way tools/assetbuilder.cpp does.
//
//   host_driver chain  W H LEVELS in.bin out.bin    image_buildmips_cube_ibl on a payload file
//   host_driver hdr    FILE.hdr W H LEVELS out.bin   write_skybox_asset(hdr) flow: load_hdr + image_pack_cube_ibl
//   host_driver luts   out.bin                       256x256 env-BRDF LUT followed by the water LUT of assetbuilder.cpp:503,557
//   host_driver nogpu                                expects std::runtime_error when no device is usable

#include "ibl.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <vector>

static size_t datasize(int w, int h, int layers, int levels) // tools/assetpacker.cpp:488-497
{
  size_t size = 0;
  for(int i = 0; i < levels; ++i)
    size += (size_t)(w >> i) * (h >> i) * layers * sizeof(uint32_t);
  return size;
}

int main(int argc, char **argv)
{
  try
  {
    std::string mode = argc > 1 ? argv[1] : "";

    if (mode == "chain" && argc == 7)
    {
      int w = atoi(argv[2]), h = atoi(argv[3]), levels = atoi(argv[4]);
      std::vector<char> payload(datasize(w, h, 6, levels));
      std::ifstream(argv[5], std::ios::binary).read(payload.data(), (std::streamsize)((size_t)w * h * 6 * 4));
      image_buildmips_cube_ibl(w, h, levels, payload.data());
      std::ofstream(argv[6], std::ios::binary).write(payload.data(), (std::streamsize)payload.size());
      return 0;
    }

    if (mode == "hdr" && argc == 7)
    {
      int w = atoi(argv[3]), h = atoi(argv[4]), levels = atoi(argv[5]);
      HDRImage image = load_hdr(argv[2]);
      std::vector<char> payload(datasize(w, h, 6, levels));
      image_pack_cube_ibl(image, w, h, levels, payload.data());
      std::ofstream(argv[6], std::ios::binary).write(payload.data(), (std::streamsize)payload.size());
      std::cout << image.width << " " << image.height << " " << image.exposure << std::endl;
      return 0;
    }

    if (mode == "faces" && argc == 7)
    {
      // tools/assetbuilder.cpp:416-470 after image loading: six ARGB32 images in, payload out
      int w = atoi(argv[2]), h = atoi(argv[3]), levels = atoi(argv[4]);
      std::vector<unsigned int> argb((size_t)w * h * 6);
      std::ifstream(argv[5], std::ios::binary).read((char*)argb.data(), (std::streamsize)(argb.size() * 4));
      std::vector<char> payload(datasize(w, h, 6, levels));
      image_pack_cube_faces_ibl(argb.data(), w, h, levels, payload.data());
      std::ofstream(argv[6], std::ios::binary).write(payload.data(), (std::streamsize)payload.size());
      return 0;
    }

    if (mode == "irradiance" && argc == 8)
    {
      // a skybox's diffuse side as two more IMAG payloads: SH9 (3 x 9 f32) and an iw x ih irradiance cube (rgbe)
      int w = atoi(argv[2]), h = atoi(argv[3]), iw = atoi(argv[4]), ih = atoi(argv[5]);
      std::vector<char> level0((size_t)w * h * 6 * 4);
      std::ifstream(argv[6], std::ios::binary).read(level0.data(), (std::streamsize)level0.size());
      std::vector<char> payload(datasize(3, 9, 1, 1) + datasize(iw, ih, 6, 1));
      image_pack_irradiance_sh9(w, h, level0.data(), payload.data());
      image_pack_irradiance_cube(payload.data(), iw, ih, payload.data() + datasize(3, 9, 1, 1));
      std::ofstream(argv[7], std::ios::binary).write(payload.data(), (std::streamsize)payload.size());
      return 0;
    }

    if (mode == "luts" && argc == 3)
    {
      std::vector<char> payload(2 * datasize(256, 256, 1, 1));
      image_pack_envbrdf(256, 256, payload.data());
      image_pack_watercolor(lml::Color3(0.0f, 0.007f, 0.005f), lml::Color3(0.1f, 0.6f, 0.7f), 1.0f, lml::Color3(0.0f, 0.0f, 0.0f), 0.328f, 5.0f, 256, 256, payload.data() + payload.size() / 2);
      std::ofstream(argv[2], std::ios::binary).write(payload.data(), (std::streamsize)payload.size());
      return 0;
    }

    if (mode == "nogpu")
    {
      std::vector<char> payload(datasize(8, 8, 6, 2));
      image_buildmips_cube_ibl(8, 8, 2, payload.data());
      std::cout << "unexpected success" << std::endl;
      return 3;
    }

    std::cerr << "usage: host_driver chain|hdr|luts|nogpu ..." << std::endl;
    return 2;
  }
  catch(std::exception &e)
  {
    std::cout << "Critical Error: " << e.what() << std::endl; // tools/assetbuilder.cpp:978
    return 1;
  }
}
