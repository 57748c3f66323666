// bbfft/cuda/error.hpp -- CUDA backend error type
// (role of the reference's include/bbfft/cl/error.hpp:20-67 and ze/error.hpp).
#ifndef BBFFT_CUDA_ERROR_HPP
#define BBFFT_CUDA_ERROR_HPP

#include "bbfft/api.hpp"

#include <exception>
#include <string>

namespace bbfft::cuda {

// Thrown for CUDA runtime / NVRTC failures; `code` is the cudaError_t (or nvrtcResult) value.
class BBFFT_EXPORT error : public std::exception {
  public:
    error(std::string what, int code) : what_(std::move(what)), code_(code) {}
    char const *what() const noexcept override { return what_.c_str(); }
    int code() const noexcept { return code_; }

  private:
    std::string what_;
    int code_;
};

BBFFT_EXPORT void throw_on_error(int cuda_error, char const *file, int line);

} // namespace bbfft::cuda

#define BBFFT_CUDA_CHECK(X) ::bbfft::cuda::throw_on_error(static_cast<int>(X), __FILE__, __LINE__)

#endif
