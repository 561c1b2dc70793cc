// datum_b200 — shared by the translation units that implement the C ABI (cabi.cu, multi.cu).
#pragma once

#include <cuda_runtime.h>
#include <string>

namespace ibl_cabi
{
  // record the calling thread's datum_ibl_last_error() text; return 1 (the ABI's failure status)
  int fail(std::string const &what);
  int fail_cuda(const char *where, cudaError_t err);
}
