// bbfft/cuda/device.hpp -- device queries for the CUDA backend
// (role of the reference's include/bbfft/sycl/device.hpp:17-32).
#ifndef BBFFT_CUDA_DEVICE_HPP
#define BBFFT_CUDA_DEVICE_HPP

#include "bbfft/api.hpp"

#include <cstdint>

namespace bbfft {

// `device` is a CUDA device ordinal.  The info uses the reference's vocabulary:
// max_work_group_size = max threads per CTA, subgroup_sizes = {32}, local_memory_size = opt-in
// shared memory per CTA.
BBFFT_EXPORT auto get_device_info(int device) -> device_info;
// Stable id used in jit_cache keys: hash of the device UUID and compute capability.
BBFFT_EXPORT auto get_device_id(int device) -> std::uint64_t;

} // namespace bbfft

#endif
