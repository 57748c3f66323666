// bbfft/cuda/make_plan.hpp -- plan factory of the CUDA (sm_100a) backend
// (role of the reference's include/bbfft/sycl/make_plan.hpp:16-41).
#ifndef BBFFT_CUDA_MAKE_PLAN_HPP
#define BBFFT_CUDA_MAKE_PLAN_HPP

#include "bbfft/api.hpp"

#include <cuda_runtime_api.h>
#include <memory>

namespace bbfft {
namespace cuda {

// Completion handle returned by plan::execute; the counterpart of sycl::event.  Copyable,
// reference counted; the underlying cudaEvent_t is destroyed with the last copy.
class BBFFT_EXPORT event {
  public:
    event() = default;
    explicit event(cudaStream_t stream); // records a new event on `stream`
    void wait() const;                   // host-blocking (sycl::event::wait)
    cudaEvent_t native() const { return ev_ ? *ev_ : nullptr; }
    explicit operator bool() const noexcept { return bool(ev_); }

  private:
    std::shared_ptr<cudaEvent_t> ev_;
};

} // namespace cuda

using cuda_plan = plan<cuda::event>;

// Plans launch on `stream` (of the current device unless `device` is given).  Execution is
// asynchronous and stream ordered; dependency events are honoured with cudaStreamWaitEvent.
// plan::execute is thread-safe (kernel arguments are passed by value at launch).
BBFFT_EXPORT auto make_plan(configuration const &cfg, cudaStream_t stream, jit_cache *cache = nullptr)
    -> cuda_plan;
BBFFT_EXPORT auto make_plan(configuration const &cfg, cudaStream_t stream, int device,
                            jit_cache *cache = nullptr) -> cuda_plan;

} // namespace bbfft

#endif
