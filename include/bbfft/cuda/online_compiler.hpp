// bbfft/cuda/online_compiler.hpp -- NVRTC / cubin module handling for the CUDA backend
// (role of the reference's include/bbfft/sycl/online_compiler.hpp:24-108 and
//  include/bbfft/ze/online_compiler.hpp: compile_to_native).
#ifndef BBFFT_CUDA_ONLINE_COMPILER_HPP
#define BBFFT_CUDA_ONLINE_COMPILER_HPP

#include "bbfft/api.hpp"

#include <cstdint>
#include <string>
#include <vector>

namespace bbfft::cuda {

// Compile CUDA C++ source (kernel stubs as written by generate_fft_kernels) to a cubin for
// `arch` (e.g. "sm_100a") with NVRTC.  Needs no GPU.
BBFFT_EXPORT auto compile_to_native(std::string const &source, std::string const &arch = "sm_100a",
                                    std::vector<std::string> const &options = {})
    -> std::vector<std::uint8_t>;

// JIT: source -> loaded module on the current device.
BBFFT_EXPORT auto build_native_module(std::string const &source, int device,
                                      std::vector<std::string> const &options = {})
    -> module_handle_t;
// Load a cubin produced by compile_to_native / nvcc -cubin.
BBFFT_EXPORT auto build_native_module(std::uint8_t const *binary, std::size_t binary_size,
                                      module_format format, int device) -> module_handle_t;
// Reference-counted ownership; the module is unloaded with the last handle.
BBFFT_EXPORT auto make_shared_handle(module_handle_t mod) -> shared_handle<module_handle_t>;
// Kernel names contained in a loaded module.
BBFFT_EXPORT auto get_kernel_names(module_handle_t mod) -> std::vector<std::string>;
// Module + its kernel names + device id, ready for aot_cache::register_module.
BBFFT_EXPORT aot_module create_aot_module(std::uint8_t const *binary, std::size_t binary_size,
                                          module_format format, int device);

} // namespace bbfft::cuda

#endif
