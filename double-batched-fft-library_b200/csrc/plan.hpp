// plan.hpp -- internal: plan implementation classes of the CUDA backend (see plan.cpp).
#ifndef BBFFT_CUDA_PLAN_HPP
#define BBFFT_CUDA_PLAN_HPP

#include "bbfft/api.hpp"
#include "bbfft/cuda/make_plan.hpp"
#include "planner.hpp"
#include "runtime.hpp"

#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace bbfft::cuda {

problem_1d to_problem(configuration const &cfg);

// Decompose a 2d/3d configuration into its double-batched 1d passes, in execution order.
void nd_passes(configuration const &cfg, std::function<void(configuration const &)> const &visit);

class plan_base : public detail::plan_impl<event> {
  public:
    using detail::plan_impl<event>::execute;
    // stream-ordered launch without event bookkeeping (used by the C ABI and by nd plans)
    virtual void enqueue(void const *in, void *out, cudaStream_t stream) = 0;
    virtual cudaStream_t stream() const = 0;
    virtual int device() const = 0;
    virtual unsigned launches_per_execute() const = 0;
    // Slab interface for host-side pipelining / sharding: number of independent k slices, the
    // byte extent of one slice in the input/output tensor (0 = not sliceable), and a launch over
    // slices [k0, k0 + count) given the base pointers of the whole tensors.
    virtual std::uint64_t slices() const { return 0; }
    virtual std::size_t in_slice_bytes() const { return 0; }
    virtual std::size_t out_slice_bytes() const { return 0; }
    virtual void enqueue_slab(void const *, void *, std::uint64_t, std::uint64_t, cudaStream_t) {
        throw bad_configuration("plan cannot be sliced");
    }
    // A k slice is one contiguous byte range of in_slice_bytes() / out_slice_bytes() on both sides and
    // slices are laid out in k order (false e.g. for istride = {1, M*K, M}: k interleaved with n)
    virtual bool slices_contiguous() const { return false; }
    // bytes a caller's input / output buffer must span (last addressed element + 1)
    virtual std::size_t in_bytes_required() const = 0;
    virtual std::size_t out_bytes_required() const = 0;
    // identifiers of the kernels this plan launches (cache keys)
    virtual void kernel_names(std::vector<std::string> &names) const = 0;
    auto execute(void const *in, void *out, std::vector<event> const &dep_events) -> event override;
};

class fft1d_plan : public plan_base {
  public:
    fft1d_plan(configuration const &cfg, api a, jit_cache *cache, std::string const &tune = std::string());
    ~fft1d_plan() override;
    fft1d_plan(fft1d_plan const &) = delete;
    fft1d_plan &operator=(fft1d_plan const &) = delete;

    void enqueue(void const *in, void *out, cudaStream_t stream) override;
    cudaStream_t stream() const override { return api_.stream(); }
    int device() const override { return api_.device(); }
    unsigned launches_per_execute() const override { return 1; }
    kernel_plan const &kernel() const { return kp_; }
    void kernel_names(std::vector<std::string> &names) const override { names.push_back(kp_.identifier); }
    std::uint64_t slices() const override { return K_; }
    std::size_t in_slice_bytes() const override { return in_slice_bytes_; }
    std::size_t out_slice_bytes() const override { return out_slice_bytes_; }
    void enqueue_slab(void const *in, void *out, std::uint64_t k0, std::uint64_t count,
                      cudaStream_t stream) override;
    bool slices_contiguous() const override { return contiguous_; }
    std::size_t in_bytes_required() const override { return in_required_; }
    std::size_t out_bytes_required() const override { return out_required_; }

  private:
    bool contiguous_ = false;
    std::size_t in_required_ = 0, out_required_ = 0;
    api api_;
    kernel_plan kp_;
    std::uint64_t K_ = 0;
    std::size_t in_slice_bytes_ = 0, out_slice_bytes_ = 0;
    // M == 1 real transforms read/write the real tensor as aligned complex words; a real pointer at
    // an odd element offset (the reference accepts it) is served by a second kernel planned with
    // PAIR=0, created on first use
    configuration cfg_;
    jit_cache *cache_ = nullptr;
    std::string tune_;
    std::unique_ptr<fft1d_plan> unaligned_;
    std::mutex unaligned_mtx_;
    shared_handle<module_handle_t> module_;
    cudaKernel_t kernel_ = nullptr;
    void *twiddle_ = nullptr;
    std::uint64_t prefetch_ = 0; // L2 prefetch distance in CTA-batches of k (bbk::args::pf)
};

// Fused 2d c2c plan: one launch of bbk::fft2d_tile, one CTA per M x N1 x N2 tile held in shared
// memory (replaces two chained 1d passes of the reference's nd_fft when the tile fits).
class fft2d_plan : public plan_base {
  public:
    fft2d_plan(problem_2d const &prob, api a, jit_cache *cache, std::string const &tune = std::string());
    ~fft2d_plan() override;
    fft2d_plan(fft2d_plan const &) = delete;
    fft2d_plan &operator=(fft2d_plan const &) = delete;

    void enqueue(void const *in, void *out, cudaStream_t stream) override;
    cudaStream_t stream() const override { return api_.stream(); }
    int device() const override { return api_.device(); }
    unsigned launches_per_execute() const override { return 1; }
    tile_plan const &kernel() const { return tp_; }
    void kernel_names(std::vector<std::string> &names) const override { names.push_back(tp_.identifier); }
    std::uint64_t slices() const override { return K_; }
    std::size_t in_slice_bytes() const override { return in_slice_bytes_; }
    std::size_t out_slice_bytes() const override { return out_slice_bytes_; }
    void enqueue_slab(void const *in, void *out, std::uint64_t k0, std::uint64_t count,
                      cudaStream_t stream) override;
    bool slices_contiguous() const override { return true; }
    std::size_t in_bytes_required() const override { return in_slice_bytes_ * K_; }
    std::size_t out_bytes_required() const override { return out_slice_bytes_ * K_; }
    // fused real tiles with M == 1 read / write the real rows as aligned complex words
    bool pointers_ok(void const *in, void const *out) const;

  private:
    api api_;
    tile_plan tp_;
    std::uint64_t K_ = 0;
    std::size_t in_slice_bytes_ = 0, out_slice_bytes_ = 0;
    shared_handle<module_handle_t> module_;
    cudaKernel_t kernel_ = nullptr;
    void *twiddle_ = nullptr;
    std::uint64_t prefetch_ = 0; // L2 prefetch distance in tiles
    std::uint64_t resident_ctas_ = 1; // grid of the persistent tile kernel
};

// One step of a 2d/3d decomposition: either a fused tile launch over dims (1,2) or a
// double-batched 1d pass; `mult` = slices of the step per outer k.
struct nd_step {
    bool fused = false;
    problem_2d tile;
    configuration pass;
    std::uint64_t mult = 1;
};
// (for_chain: steps of one persistent chain kernel -- the fused tile must fit a single CTA)
std::vector<nd_step> nd_decompose(configuration const &cfg, device_props const &dev, bool for_chain = false,
                                  bool fuse_real = true);

class nd_plan : public plan_base {
  public:
    // fuse_real = false: real transforms as one launch per mode (the fallback of a fused real plan)
    nd_plan(configuration const &cfg, api a, jit_cache *cache, bool fuse_real = true);
    ~nd_plan() override;
    nd_plan(nd_plan const &) = delete;
    nd_plan &operator=(nd_plan const &) = delete;

    void enqueue(void const *in, void *out, cudaStream_t stream) override;
    cudaStream_t stream() const override { return api_.stream(); }
    int device() const override { return api_.device(); }
    unsigned launches_per_execute() const override;
    std::vector<std::shared_ptr<plan_base>> const &passes() const { return plans_; }
    void kernel_names(std::vector<std::string> &names) const override {
        if (chained_) names.push_back(chain_.identifier);
        for (auto const &q : plans_) q->kernel_names(names);
    }
    std::uint64_t k_block() const { return kblock_; }
    bool chained() const { return chained_; }
    std::size_t in_bytes_required() const override { return in_required_; }
    std::size_t out_bytes_required() const override { return out_required_; }

  private:
    api api_;
    unsigned dim_;
    std::vector<std::shared_ptr<plan_base>> plans_;
    std::vector<std::uint64_t> mult_; // slices of pass d per outer k
    std::uint64_t K_ = 0, kblock_ = 0; // outer batch and its L2 block (in k)
    std::size_t in_required_ = 0, out_required_ = 0;
    void *tmp_ = nullptr;
    // a fused real tile step with M == 1 needs complex-aligned real pointers; executes with other pointers go
    // through this plan (one launch per mode), built on first use
    configuration cfg_;
    jit_cache *cache_ = nullptr;
    std::shared_ptr<fft2d_plan> real_tile_;
    std::unique_ptr<nd_plan> unfused_;
    std::mutex unfused_mtx_; // (executes of one plan may come from several host threads)
    // chained execution: all steps in one persistent launch (bbk::chain)
    bool try_chain(std::vector<nd_step> const &steps, jit_cache *cache);
    bool chained_ = false;
    chain_plan_t chain_;
    std::vector<std::size_t> step_in_bytes_, step_out_bytes_; // bytes of one slab as step d sees it
    shared_handle<module_handle_t> chain_module_;
    cudaKernel_t chain_kernel_ = nullptr;
    void *chain_tw_ = nullptr, *chain_done_ = nullptr;
    std::uint64_t chain_epoch_ = 0, chain_kblock_ = 1, chain_grid_ = 0;
};

// kernel from the built-in ahead-of-time bundle, or an empty handle
shared_handle<module_handle_t> builtin_module(std::string const &kernel_name, int device);

std::shared_ptr<plan_base> select_fft_algorithm(configuration const &cfg, api a, jit_cache *cache);

} // namespace bbfft::cuda

#endif
