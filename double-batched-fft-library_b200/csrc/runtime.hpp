// runtime.hpp -- internal: thin CUDA runtime / NVRTC layer of the bbfft CUDA backend.
//
// Plays the role of the reference's `Api` policy classes (src/base/dummy_api.hpp:18-61,
// src/sycl/api.hpp:22-73, src/cl/api.hpp:24-83, src/ze/api.hpp:26-82): build a module from
// source, fetch a kernel, launch it on a queue, allocate device buffers, upload twiddle tables.
// Only the CUDA runtime API is used (linked statically), so the shared library loads on
// machines without a driver; NVRTC is dlopen'ed on first use.
#ifndef BBFFT_CUDA_RUNTIME_HPP
#define BBFFT_CUDA_RUNTIME_HPP

#include "bbfft/api.hpp"
#include "planner.hpp"

#include <cuda_runtime_api.h>

#include <cstdint>
#include <string>
#include <vector>

namespace bbfft::cuda {

struct kernel_args { // must match bbk::args
    const void *in;
    void *out;
    const void *tw;
    unsigned long long K;
    unsigned long long M;
    long long is1, is2, os1, os2;
    unsigned long long pf;
};

struct chain_kernel_args { // must match bbk::chain_args
    kernel_args step[3];
    unsigned long long *done;
    unsigned long long epoch, K, kblock;
};

class api {
  public:
    explicit api(cudaStream_t stream, int device = -1);

    int device() const { return device_; }
    cudaStream_t stream() const { return stream_; }
    device_props const &props() const { return props_; }
    device_info info() const;
    std::uint64_t device_id() const { return device_id_; }
    std::string arch() const;

    // source -> cubin -> loaded module
    shared_handle<module_handle_t> build_module(std::string const &source,
                                                std::vector<std::string> const &options = {}) const;
    // bytes of local memory per thread (register spills; the kernels have no local arrays)
    std::size_t kernel_local_bytes(cudaKernel_t k) const;
    cudaKernel_t create_kernel(module_handle_t mod, std::string const &name, std::size_t smem_bytes) const;
    void launch_kernel(cudaKernel_t k, std::uint64_t grid, int threads, std::size_t smem_bytes,
                       kernel_args const &args, cudaStream_t stream) const;
    // launch with an arbitrary by-value parameter block (chain kernels: bbk::chain_args)
    void launch_kernel_raw(cudaKernel_t k, std::uint64_t grid, int threads, std::size_t smem_bytes,
                           void *param, cudaStream_t stream) const;
    // resident CTAs per SM of `k` with this CTA shape
    int max_active_ctas_per_sm(cudaKernel_t k, int threads, std::size_t smem_bytes) const;
    void *create_device_buffer(std::size_t bytes) const;
    void release_buffer(void *ptr) const;
    // narrow the double table to `fp` bytes per real and upload it
    void *create_twiddle_table(std::vector<double> const &tw, int fp) const;

  private:
    cudaStream_t stream_;
    int device_;
    device_props props_;
    std::uint64_t device_id_;
};

// NVRTC (dlopen'ed).  Throws bbfft::cuda::error with the build log on failure.
// `refresh` bypasses (and overwrites) the persistent kernel cache entry.
std::vector<std::uint8_t> nvrtc_compile(std::string const &source, std::string const &arch,
                                        std::vector<std::string> const &extra_options, bool refresh = false);
// The device header text (kernels/bbfft_kernels.cuh), embedded at build time.
char const *kernel_header_text();

module_handle_t load_module_image(void const *image);
void unload_module(module_handle_t mod);

device_props query_device_props(int device);
std::uint64_t query_device_id(int device);

} // namespace bbfft::cuda

#endif
