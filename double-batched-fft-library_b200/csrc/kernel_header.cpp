// kernel_header.cpp -- the device header text, embedded so that NVRTC (and the offline tools)
// can compile kernel stubs without any file on disk.  kernel_header_embed.inc is generated from
// kernels/bbfft_kernels.cuh by build.py.
#include "runtime.hpp"

namespace bbfft::cuda {
namespace {
const char header_text[] =
#include "kernel_header_embed.inc"
    ;
}
char const *kernel_header_text() { return header_text; }
} // namespace bbfft::cuda
