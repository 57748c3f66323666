// bbfft/api.hpp -- the device-independent part of the bbfft host API, B200 build.
//
// Source-compatible with the public headers of intel/double-batched-fft-library v0.5.1
// (include/bbfft/{configuration,plan,jit_cache,jit_cache_all,aot_cache,shared_handle,
// user_module,device_info,bad_configuration,module_format,generator}.hpp): same namespace,
// type names, member names, member order and semantics, so that user code written against the
// reference compiles unchanged.  The per-topic headers of the reference's names in this
// directory simply include this file.  The CUDA specific entry points live in bbfft/cuda/.
#ifndef BBFFT_API_HPP
#define BBFFT_API_HPP

#include <array>
#include <cstddef>
#include <cstdint>
#include <exception>
#include <iosfwd>
#include <limits>
#include <memory>
#include <string>
#include <string_view>
#include <type_traits>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#ifndef BBFFT_EXPORT
#define BBFFT_EXPORT __attribute__((visibility("default")))
#endif

namespace bbfft {

// ---------------------------------------------------------------------------------------
// errors (reference: include/bbfft/bad_configuration.hpp:17-40)
// ---------------------------------------------------------------------------------------
class BBFFT_EXPORT bad_configuration : public std::exception {
  public:
    bad_configuration(std::string what) : msg_(std::move(what)) {}
    bad_configuration(char const *what) : msg_(what) {}
    char const *what() const noexcept override { return msg_.c_str(); }

  private:
    std::string msg_;
};

// ---------------------------------------------------------------------------------------
// user callbacks (reference: include/bbfft/user_module.hpp:16-37).  The B200 build adds the
// cuda_c language: `data` is CUDA C++ source defining __device__ functions
//     T2 load(T2 const* in, size_t offset)            (T for the real side of r2c)
//     void store(T2* out, size_t offset, T2 value)    (T for the real side of c2r)
// which are compiled into the FFT kernel with NVRTC.  opencl_c sources are translated on the
// fly for the simple subset the reference's tests use (see src: callback translation).
// ---------------------------------------------------------------------------------------
enum class kernel_language { opencl_c, cuda_c };

struct BBFFT_EXPORT user_module {
    char const *data = nullptr;
    std::size_t length = 0;
    char const *load_function = nullptr;
    char const *store_function = nullptr;
    kernel_language language = kernel_language::opencl_c;

    explicit operator bool() const noexcept { return data != nullptr && length > 0; }
};

// ---------------------------------------------------------------------------------------
// configuration (reference: include/bbfft/configuration.hpp:24-192, src/base/configuration.cpp)
// ---------------------------------------------------------------------------------------
enum class precision : int { f32 = 4, f64 = 8 };
template <typename T> struct to_precision;
template <> struct to_precision<float> {
    static constexpr precision value = precision::f32;
};
template <> struct to_precision<double> {
    static constexpr precision value = precision::f64;
};
template <typename T> inline constexpr precision to_precision_v = to_precision<T>::value;

enum class direction : int { forward = -1, backward = 1 };
enum class transform_type : int { c2c, r2c, c2r };
BBFFT_EXPORT char const *to_string(transform_type type);

constexpr unsigned max_fft_dim = 3;
constexpr unsigned max_tensor_dim = max_fft_dim + 2;

using tensor_extent = std::array<std::size_t, max_tensor_dim>;

// Packed column-major strides of the M x N_1 x ... x N_d x K tensor; the first FFT mode of a
// real tensor is N_1 (out-of-place) or 2(N_1/2+1) (in-place), of its spectrum N_1/2+1.
BBFFT_EXPORT auto default_istride(unsigned dim, tensor_extent const &shape, transform_type type,
                                  bool inplace) -> tensor_extent;
BBFFT_EXPORT auto default_ostride(unsigned dim, tensor_extent const &shape, transform_type type,
                                  bool inplace) -> tensor_extent;

struct BBFFT_EXPORT configuration {
    unsigned dim;
    tensor_extent shape;
    precision fp;
    direction dir = direction::forward;
    transform_type type = transform_type::c2c;
    tensor_extent istride = default_istride(dim, shape, type, true);
    tensor_extent ostride = default_ostride(dim, shape, type, true);
    user_module callbacks = {};

    void set_strides_default(bool inplace);
    std::string to_string() const;
};
BBFFT_EXPORT std::ostream &operator<<(std::ostream &os, configuration const &cfg);

// ---------------------------------------------------------------------------------------
// device_info (reference: include/bbfft/device_info.hpp:18-44, src/base/device_info.cpp)
// ---------------------------------------------------------------------------------------
enum class device_type { gpu, cpu, custom };

struct BBFFT_EXPORT device_info {
    std::size_t max_work_group_size = 0;
    std::vector<std::size_t> subgroup_sizes;
    std::size_t local_memory_size = 0;
    device_type type = device_type::gpu;

    std::size_t min_subgroup_size() const;
    std::size_t max_subgroup_size() const;
    std::size_t register_space_min() const;
    std::size_t register_space_max() const;
    std::string to_string() const;
    bool operator==(device_info const &other) const;
    bool operator!=(device_info const &other) const;
};
BBFFT_EXPORT std::ostream &operator<<(std::ostream &os, device_type type);
BBFFT_EXPORT std::ostream &operator<<(std::ostream &os, device_info const &info);

enum class module_format { spirv, native };

// ---------------------------------------------------------------------------------------
// shared_handle (reference: include/bbfft/shared_handle.hpp:18-52)
// ---------------------------------------------------------------------------------------
template <typename T, typename Enable = void> class shared_handle;
template <typename T>
class shared_handle<T, std::enable_if_t<sizeof(T) <= sizeof(void *) && alignof(T) <= sizeof(void *)>> {
  public:
    shared_handle() = default;
    shared_handle(T t, void (*delete_handle)(T))
        : box_(reinterpret_cast<void *>(t),
               [delete_handle](void *p) { delete_handle(reinterpret_cast<T>(p)); }) {}
    T get() const { return reinterpret_cast<T>(box_.get()); }
    explicit operator bool() const noexcept { return static_cast<bool>(box_); }

  private:
    std::shared_ptr<void> box_;
};

// ---------------------------------------------------------------------------------------
// kernel caches (reference: include/bbfft/jit_cache.hpp:18-69, jit_cache_all.hpp:20-37,
// aot_cache.hpp:21-48).  On this backend a module handle is a cudaLibrary_t.
// ---------------------------------------------------------------------------------------
using device_handle_t = std::uintptr_t;
using module_handle_t = std::uintptr_t;

struct BBFFT_EXPORT jit_cache_key {
    std::string kernel_name = {};
    std::uint64_t device_id = std::numeric_limits<std::uint64_t>::max();
    bool operator==(jit_cache_key const &other) const {
        return device_id == other.device_id && kernel_name == other.kernel_name;
    }
};
struct jit_cache_key_hash {
    std::size_t operator()(jit_cache_key const &key) const noexcept {
        std::size_t h = std::hash<std::string>{}(key.kernel_name);
        return h ^ (std::hash<std::uint64_t>{}(key.device_id) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2));
    }
};

class BBFFT_EXPORT jit_cache {
  public:
    virtual ~jit_cache();
    virtual auto get(jit_cache_key const &key) const -> shared_handle<module_handle_t> = 0;
    virtual void store(jit_cache_key const &key, shared_handle<module_handle_t> mod) = 0;
};

class BBFFT_EXPORT jit_cache_all : public jit_cache {
  public:
    auto get(jit_cache_key const &key) const -> shared_handle<module_handle_t> override;
    void store(jit_cache_key const &key, shared_handle<module_handle_t> mod) override;
    auto kernel_names() const -> std::vector<std::string>;

  private:
    std::unordered_map<jit_cache_key, shared_handle<module_handle_t>, jit_cache_key_hash> mods_;
};

struct BBFFT_EXPORT aot_module {
    shared_handle<module_handle_t> mod;
    std::unordered_set<std::string> kernel_names;
    std::uint64_t device_id;
};

class BBFFT_EXPORT aot_cache : public jit_cache {
  public:
    auto get(jit_cache_key const &key) const -> shared_handle<module_handle_t> override;
    void store(jit_cache_key const &key, shared_handle<module_handle_t> mod) override;
    void register_module(aot_module aot_mod);

  private:
    std::vector<aot_module> aot_modules_;
};

// ---------------------------------------------------------------------------------------
// plans (reference: include/bbfft/plan.hpp:24-184, include/bbfft/detail/plan_impl.hpp:17-82)
// ---------------------------------------------------------------------------------------
namespace detail {
template <typename EventT> class plan_impl {
  public:
    using event_t = EventT;
    virtual ~plan_impl() {}
    virtual auto execute(void const *in, void *out) -> event_t {
        return execute(in, out, std::vector<event_t>{});
    }
    virtual auto execute(void const *in, void *out, event_t dep_event) -> event_t {
        return execute(in, out, std::vector<event_t>{std::move(dep_event)});
    }
    virtual auto execute(void const *in, void *out, std::vector<event_t> const &dep_events)
        -> event_t = 0;
};
template <typename EventT> class plan_unmanaged_event_impl {
  public:
    using event_t = EventT;
    virtual ~plan_unmanaged_event_impl() {}
    virtual void execute(void const *in, void *out, event_t signal_event,
                         std::uint32_t num_wait_events, event_t *wait_events) = 0;
};
} // namespace detail

template <class Impl> class base_plan {
  public:
    base_plan() : impl_(nullptr) {}
    base_plan(std::shared_ptr<Impl> impl) : impl_(std::move(impl)) {}
    inline explicit operator bool() const noexcept { return bool(impl_); }

  protected:
    std::shared_ptr<Impl> impl_;
};

template <typename EventT> class plan : public base_plan<detail::plan_impl<EventT>> {
  public:
    using base_plan<detail::plan_impl<EventT>>::base_plan;
    using event_t = EventT;

    auto execute(void const *in, void *out) -> event_t { return this->impl_->execute(in, out); }
    auto execute(void const *in, void *out, event_t dep_event) -> event_t {
        return this->impl_->execute(in, out, std::move(dep_event));
    }
    auto execute(void const *in, void *out, std::vector<event_t> const &dep_events) -> event_t {
        return this->impl_->execute(in, out, dep_events);
    }
    auto execute(void *inout) -> event_t { return this->impl_->execute(inout, inout); }
    auto execute(void *inout, event_t dep_event) -> event_t {
        return this->impl_->execute(inout, inout, std::move(dep_event));
    }
    auto execute(void *inout, std::vector<event_t> const &dep_events) -> event_t {
        return this->impl_->execute(inout, inout, dep_events);
    }
};

template <typename EventT>
class plan_unmanaged_event : public base_plan<detail::plan_unmanaged_event_impl<EventT>> {
  public:
    using base_plan<detail::plan_unmanaged_event_impl<EventT>>::base_plan;
    using event_t = EventT;

    void execute(void const *in, void *out, event_t signal_event = nullptr,
                 std::uint32_t num_wait_events = 0, event_t *wait_events = nullptr) {
        this->impl_->execute(in, out, signal_event, num_wait_events, wait_events);
    }
    void execute(void *inout, event_t signal_event = nullptr, std::uint32_t num_wait_events = 0,
                 event_t *wait_events = nullptr) {
        this->impl_->execute(inout, inout, signal_event, num_wait_events, wait_events);
    }
};

// ---------------------------------------------------------------------------------------
// offline generation (reference: include/bbfft/generator.hpp:22-36, src/base/generator.cpp)
// Writes the CUDA C++ stubs for all kernels the plans of `cfgs` need to `os` and returns the
// kernel names.  Compile the text with nvcc/NVRTC (include path: the kernels directory) to
// get a cubin for cuda::create_aot_module.
// ---------------------------------------------------------------------------------------
BBFFT_EXPORT std::vector<std::string> generate_fft_kernels(std::ostream &os,
                                                           std::vector<configuration> const &cfgs,
                                                           device_info const &info);

// ---------------------------------------------------------------------------------------
// descriptor mini-language (reference: include/bbfft/parser.hpp, src/base/parser.cpp:59-270)
// ---------------------------------------------------------------------------------------
BBFFT_EXPORT configuration parse_fft_descriptor(std::string_view desc);
BBFFT_EXPORT device_info parse_device_info(std::string_view desc);

} // namespace bbfft

#endif // BBFFT_API_HPP
