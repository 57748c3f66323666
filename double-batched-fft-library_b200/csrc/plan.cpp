// plan.cpp -- plan objects of the CUDA backend: fft1d_plan (one kernel launch per execute),
// nd_plan (2d/3d as chained double-batched 1d passes), make_plan, generate_fft_kernels.
//
// Replaces, for the CUDA backend, the reference's
//   src/common/algorithm.hpp:20-30            select_fft_algorithm
//   src/common/algorithm_1d.hpp:19-36         select_1d_fft_algorithm
//   src/common/algorithm/small_batch_fft.hpp  + factor2_slm_fft.hpp   (1d plans)
//   src/common/algorithm/nd_fft.hpp:27-152    (2d/3d)
//   src/sycl/plan.cpp:16-24                   make_plan
//   src/base/generator.cpp:16-25              generate_fft_kernels
#include "plan.hpp"

#include "bbfft/cuda/error.hpp"

#include "bbfft/cuda/online_compiler.hpp"

#include <cstdlib>
#include <dirent.h>
#include <dlfcn.h>

#include <fstream>
#include <iterator>
#include <map>
#include <mutex>
#include <ostream>
#include <sstream>

namespace bbfft::cuda {

// ------------------------------------------------------------------------------------------
// configuration -> planner problem
// ------------------------------------------------------------------------------------------
static std::string translate_callbacks(user_module const &um, precision fp) {
    std::string src(um.data, um.length);
    if (um.language == kernel_language::cuda_c) return src;
    // OpenCL-C callbacks (reference: test/callback.cpp:55-63,151-158): map the address space
    // qualifiers away and give float2/double2 the OpenCL conversions the sources rely on
    // (scalar -> vector broadcast, component-wise arithmetic).
    (void)fp;
    std::ostringstream os;
    os << "#ifdef BBFFT_EMU\n#define __device__\n#endif\n"
          "#define BBFFT_OCL_COMPAT 1\n"
          "namespace bbfft_ocl {\n"
          "template <class T> struct alignas(2 * sizeof(T)) vec2 {\n"
          "    T x, y;\n"
          "    __device__ vec2() {}\n"
          "    __device__ vec2(T a, T b) : x(a), y(b) {}\n"
          "    __device__ vec2(T a) : x(a), y(a) {}\n"
          "    __device__ vec2(int a) : x(T(a)), y(T(a)) {}\n"
          "};\n"
          "template <class T> __device__ vec2<T> operator+(vec2<T> a, vec2<T> b) { return vec2<T>(a.x + b.x, a.y + b.y); }\n"
          "template <class T> __device__ vec2<T> operator-(vec2<T> a, vec2<T> b) { return vec2<T>(a.x - b.x, a.y - b.y); }\n"
          "template <class T> __device__ vec2<T> operator*(vec2<T> a, vec2<T> b) { return vec2<T>(a.x * b.x, a.y * b.y); }\n"
          "template <class T> __device__ vec2<T> operator*(vec2<T> a, T s) { return vec2<T>(a.x * s, a.y * s); }\n"
          "template <class T> __device__ vec2<T> operator*(T s, vec2<T> a) { return vec2<T>(a.x * s, a.y * s); }\n"
          "template <class T> __device__ vec2<T> operator/(vec2<T> a, T s) { return vec2<T>(a.x / s, a.y / s); }\n"
          "}\n"
          "#define float2 bbfft_ocl::vec2<float>\n"
          "#define double2 bbfft_ocl::vec2<double>\n"
          "#define global\n#define local\n#define constant const\n#define private\n"
          "#define __global\n#define __local\n#define __constant const\n#define __private\n"
          "typedef unsigned int uint;\ntypedef unsigned long ulong;\ntypedef unsigned short ushort;\n"
          "typedef unsigned char uchar;\n";
    os << src << "\n";
    os << "#undef global\n#undef local\n#undef constant\n#undef private\n";
    return os.str();
}

problem_1d to_problem(configuration const &cfg) {
    if (cfg.dim != 1) throw bad_configuration("internal: 1d problem expected");
    if (cfg.istride[0] != 1 || cfg.ostride[0] != 1) {
        throw bad_configuration("stride[0] must be 1 for input and output tensor");
    }
    if (cfg.fp != precision::f32 && cfg.fp != precision::f64) {
        throw bad_configuration("unsupported precision");
    }
    if ((cfg.type == transform_type::r2c && cfg.dir != direction::forward) ||
        (cfg.type == transform_type::c2r && cfg.dir != direction::backward)) {
        throw bad_configuration("r2c direction must be forward and c2r direction must be backward");
    }
    if (cfg.shape[0] == 0 || cfg.shape[1] == 0) throw bad_configuration("empty FFT shape");
    problem_1d p;
    p.fp = static_cast<int>(cfg.fp);
    p.dir = static_cast<int>(cfg.dir);
    p.type = static_cast<int>(cfg.type);
    p.M = cfg.shape[0];
    p.N = cfg.shape[1];
    p.K = cfg.shape[2];
    p.is1 = std::int64_t(cfg.istride[1]);
    p.is2 = std::int64_t(cfg.istride[2]);
    p.os1 = std::int64_t(cfg.ostride[1]);
    p.os2 = std::int64_t(cfg.ostride[2]);
    if (cfg.callbacks) {
        p.cb_source = translate_callbacks(cfg.callbacks, cfg.fp);
        if (cfg.callbacks.load_function) p.cb_load = cfg.callbacks.load_function;
        if (cfg.callbacks.store_function) p.cb_store = cfg.callbacks.store_function;
    }
    return p;
}

// ------------------------------------------------------------------------------------------
// built-in ahead-of-time bundle: builtin_kernels.cubin next to this shared library, compiled by
// nvcc at build time from generate_fft_kernels output (aot.py).  One module per device.
// ------------------------------------------------------------------------------------------
namespace {
struct builtin_bundle {
    std::vector<std::vector<char>> images;
    bool tried = false;
    std::mutex mtx;
    std::map<int, std::vector<aot_module>> per_device;
};
builtin_bundle &bundle() {
    static builtin_bundle b;
    return b;
}
} // namespace

shared_handle<module_handle_t> builtin_module(std::string const &kernel_name, int device) {
    auto &b = bundle();
    std::lock_guard<std::mutex> lock(b.mtx);
    if (!b.tried) {
        b.tried = true;
        char const *off = std::getenv("BBFFT_CUDA_NO_BUILTIN");
        Dl_info info;
        if (!(off && *off == '1') && dladdr(reinterpret_cast<void *>(&builtin_module), &info) && info.dli_fname) {
            std::string dir(info.dli_fname);
            auto slash = dir.find_last_of('/');
            dir = slash == std::string::npos ? std::string(".") : dir.substr(0, slash);
            if (DIR *d = opendir(dir.c_str())) {
                while (dirent *e = readdir(d)) {
                    std::string name(e->d_name);
                    if (name.rfind("builtin_kernels", 0) != 0 || name.size() < 6 ||
                        name.substr(name.size() - 6) != ".cubin") {
                        continue;
                    }
                    std::ifstream f(dir + "/" + name, std::ios::binary);
                    if (f) {
                        b.images.emplace_back(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
                    }
                }
                closedir(d);
            }
        }
    }
    if (b.images.empty()) return {};
    auto it = b.per_device.find(device);
    if (it == b.per_device.end()) {
        std::vector<aot_module> mods;
        for (auto const &img : b.images) {
            try {
                mods.push_back(create_aot_module(reinterpret_cast<std::uint8_t const *>(img.data()), img.size(),
                                                 module_format::native, device));
            } catch (std::exception const &) {
                // unusable on this device (e.g. other architecture): those kernels are JIT-compiled
            }
        }
        it = b.per_device.emplace(device, std::move(mods)).first;
    }
    for (auto const &m : it->second) {
        if (m.kernel_names.count(kernel_name)) return m.mod;
    }
    return {};
}

// JIT policy: a capped kernel that spills must prove itself.  The planner caps registers with
// __maxnreg__ so that the planned number of CTAs is resident; when NVRTC's ptxas can only meet the cap
// by spilling, the result comes out of the spill / rematerialisation path of ptxas -- the path that
// miscompiled one kernel in round 1 (fp64 c2r M=32 N=424 r4x53: a store address built from a dead
// register, profiles/r02b_ptxas_miscompile.md).  Such a kernel is kept only if it reproduces, bit for
// bit, the output of the same source built WITHOUT the cap on a probe problem (every multiply is an
// explicit-rounding intrinsic, so two correct builds are bit-identical: the property the callback
// tests rely on).  Otherwise -- or when the probe cannot be run (user callbacks address memory the
// plan knows nothing about) -- the uncapped build is used: fewer resident CTAs, never a wrong
// result.  BBFFT_CUDA_KEEP_REGCAP=1 skips all of this (tuning runs).
struct jit_result {
    shared_handle<module_handle_t> mod;     // the build to use
    shared_handle<module_handle_t> relaxed; // the uncapped build when the capped one spills and is still to be probed
};

static jit_result jit_module(api const &a, std::string const &source, std::string const &name,
                             std::size_t smem_bytes, bool can_probe) {
    jit_result r;
    r.mod = a.build_module(source);
    char const *keep = std::getenv("BBFFT_CUDA_KEEP_REGCAP");
    if (keep && *keep == '1') return r;
    if (a.kernel_local_bytes(a.create_kernel(r.mod.get(), name, smem_bytes)) == 0) return r;
    auto relaxed = a.build_module(source, {"-DBBK_NO_REGCAP"});
    if (can_probe) {
        r.relaxed = relaxed;
    } else {
        r.mod = relaxed;
    }
    return r;
}

// Run `capped` and `relaxed` (same kernel arguments except the output) on a probe tensor and compare
// the outputs byte for byte.
static bool probe_identical(api const &a, cudaKernel_t capped, cudaKernel_t relaxed, std::uint64_t grid, int threads,
                            std::size_t smem_bytes, kernel_args args, std::size_t in_bytes, std::size_t out_bytes, int fp) {
    // input: reproducible values in (-1, 1) of the tensor's precision (any bit pattern would do for
    // bit-identity, finite values keep NaN != NaN out of the comparison)
    std::vector<unsigned char> host_in(in_bytes);
    {
        std::uint32_t s = 2463534242u;
        auto next = [&] {
            s ^= s << 13;
            s ^= s >> 17;
            s ^= s << 5;
            return double(s >> 8) / double(1u << 23) - 1.0;
        };
        if (fp == 4) {
            auto *f = reinterpret_cast<float *>(host_in.data());
            for (std::size_t i = 0; i < in_bytes / 4; ++i) f[i] = float(next());
        } else {
            auto *f = reinterpret_cast<double *>(host_in.data());
            for (std::size_t i = 0; i < in_bytes / 8; ++i) f[i] = next();
        }
    }
    void *din = a.create_device_buffer(in_bytes);
    void *d1 = a.create_device_buffer(out_bytes);
    void *d2 = a.create_device_buffer(out_bytes);
    bool same = false;
    try {
        cudaStream_t st = a.stream();
        BBFFT_CUDA_CHECK(cudaMemcpyAsync(din, host_in.data(), in_bytes, cudaMemcpyHostToDevice, st));
        BBFFT_CUDA_CHECK(cudaMemsetAsync(d1, 0x5a, out_bytes, st));
        BBFFT_CUDA_CHECK(cudaMemsetAsync(d2, 0x5a, out_bytes, st));
        args.in = din;
        args.out = d1;
        a.launch_kernel(capped, grid, threads, smem_bytes, args, st);
        args.out = d2;
        a.launch_kernel(relaxed, grid, threads, smem_bytes, args, st);
        std::vector<unsigned char> h1(out_bytes), h2(out_bytes);
        BBFFT_CUDA_CHECK(cudaMemcpyAsync(h1.data(), d1, out_bytes, cudaMemcpyDeviceToHost, st));
        BBFFT_CUDA_CHECK(cudaMemcpyAsync(h2.data(), d2, out_bytes, cudaMemcpyDeviceToHost, st));
        BBFFT_CUDA_CHECK(cudaStreamSynchronize(st));
        same = h1 == h2;
    } catch (...) {
        a.release_buffer(din);
        a.release_buffer(d1);
        a.release_buffer(d2);
        throw;
    }
    a.release_buffer(din);
    a.release_buffer(d1);
    a.release_buffer(d2);
    return same;
}

// Launches go to the plan's device whatever device is current in the calling thread (the kernel
// handle, its shared-memory attribute and the twiddle table belong to the creation device).
struct device_guard {
    int prev = -1;
    explicit device_guard(int device) {
        int cur = -1;
        BBFFT_CUDA_CHECK(cudaGetDevice(&cur));
        if (cur != device) {
            BBFFT_CUDA_CHECK(cudaSetDevice(device));
            prev = cur;
        }
    }
    ~device_guard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// L2 prefetch distance (bbk::prefetch_future_batch): the CTA that starts now prefetches the input of
// the batch `waves` waves of resident CTAs later.  OFF by default: measured on the B200 over the
// 210-kernel sweep (profiles/r02d_prefetch.txt) it loses -- 0.948 of the HBM peak without, 0.913 at
// one wave, 0.811 at 1.5, 0.72 at 2.5 (kernels that already saturate HBM pay for lines that are
// evicted before use and fetched twice) -- and the sizes below 0.8 are not waiting for memory.
// BBFFT_CUDA_PREFETCH_WAVES=<w> switches it on (some latency-bound kernels gain: c2r M=16 N=256 +11..17 %).
static std::uint64_t prefetch_distance(api const &a, cudaKernel_t k, int threads, std::size_t smem_bytes,
                                       int planned_blocks, std::uint64_t ctas_per_batch) {
    double waves = 0.0;
    if (char const *e = std::getenv("BBFFT_CUDA_PREFETCH_WAVES")) waves = std::atof(e);
    if (waves <= 0.0) return 0;
    int per_sm = a.max_active_ctas_per_sm(k, threads, smem_bytes);
    if (per_sm < 1) per_sm = std::max(1, planned_blocks);
    const double resident = double(per_sm) * double(a.props().sm_count);
    return std::max<std::uint64_t>(1, std::uint64_t(waves * resident / double(std::max<std::uint64_t>(1, ctas_per_batch)) + 0.5));
}

static std::string env_tune() {
    char const *t = std::getenv("BBFFT_CUDA_TUNE");
    return t ? std::string(t) : std::string();
}

// ------------------------------------------------------------------------------------------
// 1d
// ------------------------------------------------------------------------------------------
fft1d_plan::fft1d_plan(configuration const &cfg, api a, jit_cache *cache, std::string const &tune)
    : api_(std::move(a)), cfg_(cfg), cache_(cache), tune_(tune.empty() ? env_tune() : tune) {
    auto prob = to_problem(cfg);
    K_ = prob.K;
    {
        std::size_t real_bytes = static_cast<std::size_t>(cfg.fp);
        const std::size_t ie = (cfg.type == transform_type::r2c ? 1 : 2) * real_bytes;
        const std::size_t oe = (cfg.type == transform_type::c2r ? 1 : 2) * real_bytes;
        in_slice_bytes_ = std::size_t(prob.is2) * ie;
        out_slice_bytes_ = std::size_t(prob.os2) * oe;
        // rows stored per slice: N reals / N/2+1 spectrum bins on the real / complex side
        const std::size_t n_in = cfg.type == transform_type::c2r ? prob.N / 2 + 1 : prob.N;
        const std::size_t n_out = cfg.type == transform_type::r2c ? prob.N / 2 + 1 : prob.N;
        const std::size_t in_slice_extent = (n_in - 1) * std::size_t(prob.is1) + prob.M;
        const std::size_t out_slice_extent = (n_out - 1) * std::size_t(prob.os1) + prob.M;
        contiguous_ = std::size_t(prob.is2) >= in_slice_extent && std::size_t(prob.os2) >= out_slice_extent;
        if (prob.K > 0) {
            in_required_ = ((prob.K - 1) * std::size_t(prob.is2) + in_slice_extent) * ie;
            out_required_ = ((prob.K - 1) * std::size_t(prob.os2) + out_slice_extent) * oe;
        }
    }
    kp_ = plan_kernel_1d(prob, api_.props(), tune_);
    jit_cache_key key{kp_.identifier, api_.device_id()};
    if (cache) module_ = cache->get(key);
    if (!module_ && prob.cb_source.empty()) module_ = builtin_module(kp_.identifier, api_.device());
    jit_result jit;
    if (!module_) {
        jit = jit_module(api_, kp_.source, kp_.identifier, kp_.p.smem_bytes, prob.cb_source.empty() && prob.K > 0);
        module_ = jit.mod;
    }
    kernel_ = api_.create_kernel(module_.get(), kp_.identifier, kp_.p.smem_bytes);
    twiddle_ = api_.create_twiddle_table(kp_.twiddle, kp_.p.fp);
    if (jit.relaxed) {
        // the capped build spills: probe it against the uncapped build on a few CTAs' worth of slices
        cudaKernel_t relaxed = api_.create_kernel(jit.relaxed.get(), kp_.identifier, kp_.p.smem_bytes);
        const std::uint64_t kt = std::min<std::uint64_t>(prob.K, 3 * kp_.p.k_per_cta() + 1);
        kernel_args pa;
        pa.tw = twiddle_;
        pa.K = kt;
        pa.M = kp_.p.M;
        pa.is1 = kp_.p.is1;
        pa.is2 = kp_.p.is2;
        pa.os1 = kp_.p.os1;
        pa.os2 = kp_.p.os2;
        pa.pf = 0;
        pa.in = nullptr;
        pa.out = nullptr;
        const std::size_t ib = in_required_ - (prob.K - kt) * in_slice_bytes_;
        const std::size_t ob = out_required_ - (prob.K - kt) * out_slice_bytes_;
        if (!probe_identical(api_, kernel_, relaxed, kp_.p.grid(kt), kp_.p.threads, kp_.p.smem_bytes, pa, ib, ob, kp_.p.fp)) {
            module_ = jit.relaxed;
            kernel_ = relaxed;
        }
    }
    if (jit.mod && cache) cache->store(key, module_); // JIT-built here (not taken from a cache): publish the build in use
    prefetch_ = prefetch_distance(api_, kernel_, kp_.p.threads, kp_.p.smem_bytes, kp_.p.min_blocks,
                                  kp_.p.klanes ? 1 : (kp_.p.M + kp_.p.ML - 1) / kp_.p.ML);
}

fft1d_plan::~fft1d_plan() { api_.release_buffer(twiddle_); }

void fft1d_plan::enqueue(void const *in, void *out, cudaStream_t stream) {
    enqueue_slab(in, out, 0, K_, stream);
}

void fft1d_plan::enqueue_slab(void const *in, void *out, std::uint64_t k0, std::uint64_t count,
                              cudaStream_t stream) {
    if (in == out && kp_.inplace_unsupported) {
        throw bad_configuration("The plan does not support in-place transform on the current device.");
    }
    if (k0 + count > K_) throw bad_configuration("slab exceeds the planned batch");
    if ((kp_.p.pair_load && reinterpret_cast<std::uintptr_t>(in) % (2 * std::size_t(kp_.p.fp)) != 0) ||
        (kp_.p.pair_store && reinterpret_cast<std::uintptr_t>(out) % (2 * std::size_t(kp_.p.fp)) != 0)) {
        std::lock_guard<std::mutex> lock(unaligned_mtx_);
        if (!unaligned_) {
            unaligned_ = std::make_unique<fft1d_plan>(cfg_, api_, cache_, tune_.empty() ? "PAIR=0" : tune_ + ",PAIR=0");
        }
        unaligned_->enqueue_slab(in, out, k0, count, stream);
        return;
    }
    device_guard guard(api_.device());
    kernel_args a;
    a.in = static_cast<char const *>(in) + k0 * in_slice_bytes_;
    a.out = static_cast<char *>(out) + k0 * out_slice_bytes_;
    a.tw = twiddle_;
    a.K = count;
    a.M = kp_.p.M;
    a.is1 = kp_.p.is1;
    a.is2 = kp_.p.is2;
    a.os1 = kp_.p.os1;
    a.os2 = kp_.p.os2;
    a.pf = prefetch_;
    api_.launch_kernel(kernel_, kp_.p.grid(count), kp_.p.threads, kp_.p.smem_bytes, a, stream);
}

auto plan_base::execute(void const *in, void *out, std::vector<event> const &dep_events) -> event {
    device_guard guard(device()); // the stream, the returned event and the launch belong to the plan's device
    cudaStream_t s = stream();
    for (auto const &e : dep_events) {
        if (e) BBFFT_CUDA_CHECK(cudaStreamWaitEvent(s, e.native(), 0));
    }
    enqueue(in, out, s);
    return event(s);
}

// ------------------------------------------------------------------------------------------
// nd: chained double-batched 1d passes (pass d: M_d = M * prod_{e<d} Nc_e, K_d = prod_{e>d} Nc_e * K)
// ------------------------------------------------------------------------------------------
void nd_passes(configuration const &cfg, std::function<void(configuration const &)> const &visit) {
    const unsigned dim = cfg.dim;
    if (cfg.callbacks) {
        throw bad_configuration("User modules are unsuported for FFT dimension > 1.");
    }
    auto same = [&](tensor_extent const &x, tensor_extent const &y) {
        for (unsigned d = 0; d < cfg.dim + 2; ++d) {
            if (x[d] != y[d]) return false;
        }
        return true;
    };
    auto is_default = [&](bool inplace) {
        return same(cfg.istride, default_istride(cfg.dim, cfg.shape, cfg.type, inplace)) &&
               same(cfg.ostride, default_ostride(cfg.dim, cfg.shape, cfg.type, inplace));
    };
    bool inplace_layout = is_default(true);
    if (!inplace_layout && !is_default(false)) {
        throw bad_configuration("Only default tensor layouts are supported for the nd_fft.");
    }
    const bool real = cfg.type != transform_type::c2c;
    auto n_spectrum = [&](unsigned d) { return (d == 0 && real) ? cfg.shape[1] / 2 + 1 : cfg.shape[d + 1]; };
    auto n_signal = [&](unsigned d) {
        return (d == 0 && real && inplace_layout) ? 2 * (cfg.shape[1] / 2 + 1) : cfg.shape[d + 1];
    };
    std::size_t right = cfg.shape[dim + 1];
    for (unsigned d = 0; d < dim; ++d) right *= n_spectrum(d);
    std::size_t left = cfg.shape[0];
    std::vector<configuration> pass(dim);
    for (unsigned d = 0; d < dim; ++d) {
        right /= n_spectrum(d);
        configuration c = {};
        c.dim = 1;
        c.shape = {left, cfg.shape[d + 1], right, 0, 0};
        c.fp = cfg.fp;
        c.dir = cfg.dir;
        c.type = d == 0 ? cfg.type : transform_type::c2c;
        c.istride = {1, left, left * n_signal(d), 0, 0};
        c.ostride = {1, left, left * n_spectrum(d), 0, 0};
        if (cfg.type == transform_type::c2r) std::swap(c.istride, c.ostride);
        pass[d] = c;
        left *= n_spectrum(d);
    }
    for (unsigned d = 0; d < dim; ++d) {
        // c2r runs the modes in reverse so that the real mode comes last
        visit(cfg.type == transform_type::c2r ? pass[dim - 1 - d] : pass[d]);
    }
}

// ------------------------------------------------------------------------------------------
// fused 2d
// ------------------------------------------------------------------------------------------
fft2d_plan::fft2d_plan(problem_2d const &prob, api a, jit_cache *cache, std::string const &tune)
    : api_(std::move(a)) {
    K_ = prob.K;
    tp_ = plan_kernel_2d(prob, api_.props(), tune);
    in_slice_bytes_ = out_slice_bytes_ = std::size_t(tp_.p.tile_stride) * 2 * std::size_t(prob.fp);
    if (tp_.p.real == 1) in_slice_bytes_ = std::size_t(tp_.p.real_tile_stride) * std::size_t(prob.fp);
    if (tp_.p.real == 2) out_slice_bytes_ = std::size_t(tp_.p.real_tile_stride) * std::size_t(prob.fp);
    jit_cache_key key{tp_.identifier, api_.device_id()};
    if (cache) module_ = cache->get(key);
    if (!module_) module_ = builtin_module(tp_.identifier, api_.device());
    if (!module_) {
        // (tile kernels that spill under their cap run uncapped: no probe)
        module_ = jit_module(api_, tp_.source, tp_.identifier, tp_.p.smem_bytes, false).mod;
        if (cache) cache->store(key, module_);
    }
    kernel_ = api_.create_kernel(module_.get(), tp_.identifier, tp_.p.smem_bytes);
    twiddle_ = api_.create_twiddle_table(tp_.twiddle, tp_.p.fp);
    prefetch_ = prefetch_distance(api_, kernel_, tp_.p.threads, tp_.p.smem_bytes, tp_.p.min_blocks, 1);
    {
        int per_sm = api_.max_active_ctas_per_sm(kernel_, tp_.p.threads, tp_.p.smem_bytes);
        if (per_sm < 1) per_sm = std::max(1, tp_.p.min_blocks);
        resident_ctas_ = std::uint64_t(per_sm) * std::uint64_t(api_.props().sm_count);
    }
}

fft2d_plan::~fft2d_plan() { api_.release_buffer(twiddle_); }

void fft2d_plan::enqueue(void const *in, void *out, cudaStream_t stream) { enqueue_slab(in, out, 0, K_, stream); }

bool fft2d_plan::pointers_ok(void const *in, void const *out) const {
    if (tp_.p.real == 0 || tp_.p.M != 1) return true;
    const std::uintptr_t word = 2 * std::uintptr_t(tp_.p.fp);
    void const *r = tp_.p.real == 1 ? in : out;
    return reinterpret_cast<std::uintptr_t>(r) % word == 0;
}

void fft2d_plan::enqueue_slab(void const *in, void *out, std::uint64_t k0, std::uint64_t count,
                              cudaStream_t stream) {
    if (k0 + count > K_) throw bad_configuration("slab exceeds the planned batch");
    if (count == 0) return;
    if (!pointers_ok(in, out)) {
        throw bad_configuration("fused real 2d plan: the real tensor must be aligned to a complex number when M = 1");
    }
    device_guard guard(api_.device());
    kernel_args a = {};
    a.in = static_cast<char const *>(in) + k0 * in_slice_bytes_;
    a.out = static_cast<char *>(out) + k0 * out_slice_bytes_;
    a.tw = twiddle_;
    a.K = count;
    a.M = tp_.p.M;
    a.pf = tp_.p.persistent ? 0 : prefetch_;
    // the persistent kernel walks the tiles with a grid of resident CTAs
    if (tp_.p.persistent) {
        api_.launch_kernel(kernel_, std::min<std::uint64_t>(count, resident_ctas_), tp_.p.threads, tp_.p.smem_bytes, a, stream);
        return;
    }
    api_.launch_kernel(kernel_, count * std::uint64_t(tp_.p.cluster), tp_.p.threads, tp_.p.smem_bytes, a, stream);
}

// Steps of a 2d/3d plan.  c2c transforms in the default layout whose M x N1 x N2 tile fits into
// shared memory run modes 1 and 2 fused (BBFFT_CUDA_ND_FUSE=0 turns this off: "multi-pass").
std::vector<nd_step> nd_decompose(configuration const &cfg, device_props const &dev, bool for_chain, bool fuse_real) {
    std::vector<nd_step> steps;
    const std::uint64_t K = cfg.shape[cfg.dim + 1];
    std::vector<configuration> passes;
    nd_passes(cfg, [&](configuration const &c) { passes.push_back(c); }); // validates the layout, too
    char const *fuse_env = std::getenv("BBFFT_CUDA_ND_FUSE");
    // real transforms (even N1, default layouts): modes 1 and 2 run as one fused real tile kernel
    // (bbk::fft2d_tile_real_cta); BBFFT_CUDA_ND_FUSE_REAL=0 keeps one launch per mode.  Not part of a chain.
    const bool real = cfg.type != transform_type::c2c;
    char const *fuse_real_env = std::getenv("BBFFT_CUDA_ND_FUSE_REAL");
    bool fuse = !(fuse_env && *fuse_env == '0') && cfg.dim >= 2;
    if (real) fuse = fuse && fuse_real && !for_chain && !(fuse_real_env && *fuse_real_env == '0');
    problem_2d t;
    if (fuse) {
        t.fp = static_cast<int>(cfg.fp);
        t.dir = static_cast<int>(cfg.dir);
        t.M = cfg.shape[0];
        t.N1 = cfg.shape[1];
        t.N2 = cfg.shape[2];
        t.K = (cfg.dim == 3 ? cfg.shape[3] : 1) * K;
        t.tile_stride = t.M * (real ? t.N1 / 2 + 1 : t.N1) * t.N2;
        if (real) {
            t.real = cfg.type == transform_type::r2c ? 1 : 2;
            // (nd_passes accepted the layout: it is one of the two defaults; in place = padded real rows)
            auto const &rs = cfg.type == transform_type::r2c ? cfg.istride : cfg.ostride;
            t.inplace = rs[2] == 2 * t.M * (t.N1 / 2 + 1);
        }
        fuse = tile_fusable(t, dev, for_chain ? 1 : tile_cluster_limit());
    }
    // c2r visits the modes in reverse: the fused tile (modes 2 and 1) is the LAST step
    const bool tile_last = fuse && cfg.type == transform_type::c2r;
    auto push_tile = [&] {
        nd_step s;
        s.fused = true;
        s.tile = t;
        s.mult = K ? t.K / K : 0;
        steps.push_back(s);
    };
    if (fuse && !tile_last) push_tile();
    const std::size_t first = (fuse && !tile_last) ? 2 : 0;
    const std::size_t last = tile_last ? passes.size() - 2 : passes.size();
    for (std::size_t d = first; d < last; ++d) {
        nd_step s;
        s.pass = passes[d];
        s.mult = K ? passes[d].shape[2] / K : 0;
        steps.push_back(s);
    }
    if (tile_last) push_tile();
    return steps;
}

// All steps in one persistent launch (bbk::chain) when they can share a CTA shape:
// BBFFT_CUDA_ND_CHAIN=1 (opt-in); BBFFT_CUDA_ND_CHAIN_KBLOCK / _CTAS override the slabs per pipeline
// block and the resident CTAs per SM.  Measured on the B200 (profiles/r01f_c3c4_chain.jsonl,
// r01f_ncu_chain3d.txt) the chain does what it was built for -- 3d fp64 64^3 K=64 moves 541 MB
// through HBM instead of 1074 MB -- but it is not faster yet: 211 us against 190 us for one
// launch per step, because at the two resident CTAs per SM that the 64 KiB tile step allows the
// kernels are latency-bound, not bandwidth-bound (issue slots 26 % busy, barrier stalls on top).
// It stays a switch until the tile step hides its own load latency.
bool nd_plan::try_chain(std::vector<nd_step> const &steps, jit_cache *cache) {
    char const *env = std::getenv("BBFFT_CUDA_ND_CHAIN");
    if (!(env && *env == '1') || steps.size() < 2 || K_ == 0) return false;
    if (char const *blk = std::getenv("BBFFT_CUDA_ND_BLOCK_BYTES")) {
        if (std::strtoull(blk, nullptr, 10) > 0) return false; // host-side L2 blocking was asked for explicitly
    }
    std::vector<chain_step_problem> probs;
    std::size_t fp = 4;
    for (auto const &s : steps) {
        chain_step_problem c;
        c.tile = s.fused;
        c.mult = s.mult;
        if (s.fused) {
            c.t = s.tile;
            fp = std::size_t(s.tile.fp);
            step_in_bytes_.push_back(std::size_t(s.tile.tile_stride) * 2 * fp * s.mult);
            step_out_bytes_.push_back(step_in_bytes_.back());
        } else {
            c.p = to_problem(s.pass);
            fp = std::size_t(c.p.fp);
            step_in_bytes_.push_back(std::size_t(c.p.is2) * (c.p.type == 1 ? 1 : 2) * fp * s.mult);
            step_out_bytes_.push_back(std::size_t(c.p.os2) * (c.p.type == 2 ? 1 : 2) * fp * s.mult);
        }
        probs.push_back(c);
    }
    if (!plan_chain(probs, api_.props(), chain_)) return false;
    jit_cache_key key{chain_.identifier, api_.device_id()};
    if (cache) chain_module_ = cache->get(key);
    if (!chain_module_) chain_module_ = builtin_module(chain_.identifier, api_.device());
    if (!chain_module_) {
        chain_module_ = jit_module(api_, chain_.source, chain_.identifier, chain_.smem_bytes, false).mod;
        if (cache) cache->store(key, chain_module_);
    }
    chain_kernel_ = api_.create_kernel(chain_module_.get(), chain_.identifier, chain_.smem_bytes);
    // every CTA of the persistent grid must be resident: the runtime's occupancy figure, or (when it
    // cannot be queried) the planner's own -- __maxnreg__ and the shared-memory check guarantee it
    int per_sm = api_.max_active_ctas_per_sm(chain_kernel_, chain_.threads, chain_.smem_bytes);
    if (per_sm < 1) per_sm = chain_.min_blocks;
    if (char const *e = std::getenv("BBFFT_CUDA_ND_CHAIN_CTAS")) per_sm = std::max(1, std::min(per_sm, std::atoi(e)));
    chain_grid_ = std::uint64_t(per_sm) * std::uint64_t(api_.props().sm_count);
    // slabs per pipeline block: about two waves of the step with the fewest CTAs per slab, but at most
    // 16 MiB of tensor (three blocks are in flight: well inside the L2)
    std::uint64_t min_per_k = ~0ull;
    std::size_t slab_bytes = 1;
    for (std::size_t d = 0; d < chain_.steps.size(); ++d) {
        min_per_k = std::min(min_per_k, chain_.steps[d].per_k);
        slab_bytes = std::max(slab_bytes, std::max(step_in_bytes_[d], step_out_bytes_[d]));
    }
    chain_kblock_ = std::max<std::uint64_t>(1, (2 * chain_grid_ + min_per_k - 1) / min_per_k);
    chain_kblock_ = std::min<std::uint64_t>(chain_kblock_, std::max<std::uint64_t>(1, (16u << 20) / slab_bytes));
    if (char const *e = std::getenv("BBFFT_CUDA_ND_CHAIN_KBLOCK")) chain_kblock_ = std::max(1ull, std::strtoull(e, nullptr, 10));
    chain_tw_ = api_.create_twiddle_table(chain_.twiddle, int(fp));
    chain_done_ = api_.create_device_buffer(sizeof(std::uint64_t) * chain_.steps.size() * K_);
    BBFFT_CUDA_CHECK(cudaMemsetAsync(chain_done_, 0, sizeof(std::uint64_t) * chain_.steps.size() * K_, api_.stream()));
    BBFFT_CUDA_CHECK(cudaStreamSynchronize(api_.stream()));
    return true;
}

nd_plan::nd_plan(configuration const &cfg, api a, jit_cache *cache, bool fuse_real)
    : api_(std::move(a)), dim_(cfg.dim), cfg_(cfg), cache_(cache) {
    K_ = cfg.shape[dim_ + 1];
    char const *chain_env = std::getenv("BBFFT_CUDA_ND_CHAIN");
    const bool want_chain = chain_env && *chain_env == '1';
    auto steps = nd_decompose(cfg, api_.props(), want_chain, fuse_real);
    chained_ = try_chain(steps, cache);
    if (want_chain && !chained_) steps = nd_decompose(cfg, api_.props(), false, fuse_real);
    for (auto const &s : steps) {
        if (chained_) {
            mult_.push_back(s.mult);
            continue;
        }
        if (s.fused) {
            auto tp = std::make_shared<fft2d_plan>(s.tile, api_, cache);
            if (s.tile.real != 0) real_tile_ = tp; // (r2c: the first step reads `in`; c2r: the last step writes `out`)
            plans_.push_back(tp);
        } else {
            plans_.push_back(std::make_shared<fft1d_plan>(s.pass, api_, cache));
        }
        // step d sees mult_[d] slices per outer k
        mult_.push_back(s.mult);
    }
    std::size_t real_bytes = static_cast<std::size_t>(cfg.fp);
    std::size_t ibytes = (cfg.type == transform_type::r2c ? 1 : 2) * real_bytes;
    std::size_t obytes = (cfg.type == transform_type::c2r ? 1 : 2) * real_bytes;
    std::size_t isize = cfg.istride[dim_ + 1] * cfg.shape[dim_ + 1] * ibytes;
    std::size_t osize = cfg.ostride[dim_ + 1] * cfg.shape[dim_ + 1] * obytes;
    if (isize > osize) tmp_ = api_.create_device_buffer(isize);
    in_required_ = isize;
    out_required_ = osize;

    // Optional L2 blocking (BBFFT_CUDA_ND_BLOCK_BYTES=<bytes>, off by default): run all steps over
    // one block of outer k before moving to the next, so that later steps read what the previous
    // step just wrote from the 126 MB L2.  Measured on the B200 it LOSES (3d fp64 64^3 K=64:
    // 191 us unblocked, 242 us with 32 MiB blocks, 462 us with 8 MiB blocks -- profiles/
    // r01c_nd_block.txt): a block is only one or two waves of CTAs, and the drain at every kernel
    // boundary costs more than the L2 hits save.  Kept as a switch for other shapes.
    std::size_t block_bytes = 0;
    if (char const *e = std::getenv("BBFFT_CUDA_ND_BLOCK_BYTES")) block_bytes = std::strtoull(e, nullptr, 10);
    std::size_t per_k = std::max(cfg.istride[dim_ + 1] * ibytes, cfg.ostride[dim_ + 1] * obytes);
    kblock_ = K_;
    if (block_bytes > 0 && per_k > 0 && K_ > 1) {
        kblock_ = std::max<std::uint64_t>(1, block_bytes / per_k);
        // odd-N real transforms pair the slices (2k', 2k'+1) of their pass: keep blocks even
        if (cfg.type != transform_type::c2c && cfg.shape[1] % 2 == 1 && kblock_ % 2 == 1) ++kblock_;
        if (kblock_ >= K_) kblock_ = K_;
    }
}

nd_plan::~nd_plan() {
    api_.release_buffer(tmp_);
    api_.release_buffer(chain_tw_);
    api_.release_buffer(chain_done_);
}

unsigned nd_plan::launches_per_execute() const {
    if (chained_) return 1;
    std::uint64_t blocks = kblock_ ? (K_ + kblock_ - 1) / kblock_ : 1;
    return unsigned(plans_.size() * std::max<std::uint64_t>(1, blocks));
}

void nd_plan::enqueue(void const *in, void *out, cudaStream_t stream) {
    if (real_tile_ && !real_tile_->pointers_ok(in, out)) {
        // the user's real tensor is not aligned to a complex number: one launch per mode handles any pointer
        {
            std::lock_guard<std::mutex> lock(unfused_mtx_);
            if (!unfused_) unfused_ = std::make_unique<nd_plan>(cfg_, api_, cache_, false);
        }
        unfused_->enqueue(in, out, stream);
        return;
    }
    void *tmp = tmp_ ? tmp_ : out;
    const std::size_t n = chained_ ? chain_.steps.size() : plans_.size();
    auto src = [&](std::size_t d) { return d == 0 ? in : static_cast<void const *>(tmp); };
    auto dst = [&](std::size_t d) { return d + 1 == n ? out : tmp; };
    if (chained_) {
        // the launch number is a kernel argument and the completion counters grow monotonically: a
        // captured launch would be replayed with a stale epoch and skip its waits
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        BBFFT_CUDA_CHECK(cudaStreamIsCapturing(stream, &cap));
        if (cap != cudaStreamCaptureStatusNone) {
            throw bad_configuration("a chained nd plan (BBFFT_CUDA_ND_CHAIN=1) cannot be captured in a CUDA graph");
        }
        device_guard guard(api_.device());
        chain_kernel_args ca = {};
        const std::size_t esz = 2 * std::size_t(chain_.steps[0].tile ? chain_.steps[0].tp.p.fp : chain_.steps[0].kp.p.fp);
        for (std::size_t d = 0; d < n; ++d) {
            auto const &sp = chain_.steps[d];
            ca.step[d].in = src(d);
            ca.step[d].out = dst(d);
            ca.step[d].tw = static_cast<char const *>(chain_tw_) + std::size_t(sp.tw_offset) * esz;
            ca.step[d].K = mult_[d] * K_;
            ca.step[d].M = sp.tile ? sp.tp.p.M : sp.kp.p.M;
            ca.step[d].pf = 0;
        }
        ca.done = static_cast<unsigned long long *>(chain_done_);
        ca.epoch = ++chain_epoch_;
        ca.K = K_;
        ca.kblock = chain_kblock_;
        std::uint64_t items = 0;
        for (auto const &sp : chain_.steps) items += sp.per_k * K_;
        api_.launch_kernel_raw(chain_kernel_, std::min(chain_grid_, items), chain_.threads, chain_.smem_bytes, &ca, stream);
        return;
    }
    if (kblock_ == 0 || kblock_ >= K_) {
        for (std::size_t d = 0; d < n; ++d) plans_[d]->enqueue(src(d), dst(d), stream);
        return;
    }
    for (std::uint64_t k0 = 0; k0 < K_; k0 += kblock_) {
        const std::uint64_t cnt = std::min<std::uint64_t>(kblock_, K_ - k0);
        for (std::size_t d = 0; d < n; ++d) {
            plans_[d]->enqueue_slab(src(d), dst(d), k0 * mult_[d], cnt * mult_[d], stream);
        }
    }
}

// ------------------------------------------------------------------------------------------
// factory
// ------------------------------------------------------------------------------------------
std::shared_ptr<plan_base> select_fft_algorithm(configuration const &cfg, api a, jit_cache *cache) {
    if (cfg.dim == 0 || cfg.dim > max_fft_dim) {
        throw bad_configuration("Unsupported FFT dimension: " + std::to_string(cfg.dim));
    }
    if (cfg.dim == 1) return std::make_shared<fft1d_plan>(cfg, std::move(a), cache);
    return std::make_shared<nd_plan>(cfg, std::move(a), cache);
}

} // namespace bbfft::cuda

namespace bbfft {

auto make_plan(configuration const &cfg, cudaStream_t stream, jit_cache *cache) -> cuda_plan {
    return cuda_plan(cuda::select_fft_algorithm(cfg, cuda::api(stream), cache));
}

auto make_plan(configuration const &cfg, cudaStream_t stream, int device, jit_cache *cache) -> cuda_plan {
    BBFFT_CUDA_CHECK(cudaSetDevice(device));
    return cuda_plan(cuda::select_fft_algorithm(cfg, cuda::api(stream, device), cache));
}

// Offline generation needs no device: plan every configuration against `info` and print the
// stubs (reference: src/base/generator.cpp:16-25 runs the plan selection against dummy_api).
std::vector<std::string> generate_fft_kernels(std::ostream &os, std::vector<configuration> const &cfgs,
                                              device_info const &info) {
    cuda::device_props dev;
    if (info.max_work_group_size) dev.max_threads_per_block = int(info.max_work_group_size);
    if (info.local_memory_size) dev.max_smem_per_block = info.local_memory_size;
    std::vector<std::string> names, chain_stubs;
    auto emit = [&](configuration const &c1) {
        auto kp = cuda::plan_kernel_1d(cuda::to_problem(c1), dev, std::string());
        for (auto const &n : names) {
            if (n == kp.identifier) return;
        }
        names.push_back(kp.identifier);
        os << kp.source << "\n";
    };
    for (auto const &cfg : cfgs) {
        if (cfg.dim == 1) {
            emit(cfg);
        } else {
            // same decomposition as nd_plan, without a device
            auto steps = cuda::nd_decompose(cfg, dev);
            {
                // the persistent chain kernel nd_plan launches when the steps can share a CTA shape
                std::vector<cuda::chain_step_problem> probs;
                for (auto const &st : cuda::nd_decompose(cfg, dev, true)) {
                    cuda::chain_step_problem q;
                    q.tile = st.fused;
                    q.mult = st.mult;
                    if (st.fused) {
                        q.t = st.tile;
                    } else {
                        q.p = cuda::to_problem(st.pass);
                    }
                    probs.push_back(q);
                }
                cuda::chain_plan_t cp;
                bool seen = false;
                if (cuda::plan_chain(probs, dev, cp)) {
                    for (auto const &n : names) seen = seen || n == cp.identifier;
                    if (!seen) {
                        for (auto const &sp : cp.steps) {
                            std::string const &sid = sp.tile ? sp.tp.identifier : sp.kp.identifier;
                            bool have = false;
                            for (auto const &n : chain_stubs) have = have || n == sid;
                            if (!have) {
                                chain_stubs.push_back(sid);
                                os << (sp.tile ? sp.tp.source : sp.kp.source) << "\n";
                            }
                        }
                        names.push_back(cp.identifier);
                        os << cp.entry_source << "\n";
                    }
                }
            }
            for (auto const &st : steps) {
                if (!st.fused) {
                    emit(st.pass);
                    continue;
                }
                auto tp = cuda::plan_kernel_2d(st.tile, dev, std::string());
                bool seen = false;
                for (auto const &n : names) seen = seen || n == tp.identifier;
                if (!seen) {
                    names.push_back(tp.identifier);
                    os << tp.source << "\n";
                }
            }
        }
    }
    return names;
}

} // namespace bbfft
