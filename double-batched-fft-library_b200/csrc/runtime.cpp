// runtime.cpp -- see runtime.hpp.
#include "runtime.hpp"

#include "bbfft/cuda/device.hpp"
#include "bbfft/cuda/error.hpp"
#include "bbfft/cuda/make_plan.hpp"
#include "bbfft/cuda/online_compiler.hpp"

#include <dlfcn.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <mutex>
#include <sstream>
#include <utility>
#include <vector>

namespace bbfft::cuda {

void throw_on_error(int cuda_error, char const *file, int line) {
    if (cuda_error != 0) {
        std::ostringstream os;
        os << file << ":" << line << ": CUDA error " << cuda_error << " ("
           << cudaGetErrorName(static_cast<cudaError_t>(cuda_error))
           << "): " << cudaGetErrorString(static_cast<cudaError_t>(cuda_error));
        throw error(os.str(), cuda_error);
    }
}

// ------------------------------------------------------------------------------------------
// NVRTC, loaded lazily
// ------------------------------------------------------------------------------------------
namespace {
struct nvrtc_api {
    void *lib = nullptr;
    int (*create)(void **prog, const char *src, const char *name, int nh, const char *const *headers,
                  const char *const *include_names) = nullptr;
    int (*compile)(void *prog, int nopt, const char *const *opts) = nullptr;
    int (*destroy)(void **prog) = nullptr;
    int (*cubin_size)(void *prog, size_t *sz) = nullptr;
    int (*cubin)(void *prog, char *out) = nullptr;
    int (*log_size)(void *prog, size_t *sz) = nullptr;
    int (*log)(void *prog, char *out) = nullptr;
    const char *(*error_string)(int) = nullptr;
};

nvrtc_api &nvrtc() {
    static nvrtc_api a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                               "/usr/local/cuda/lib64/libnvrtc.so"};
        for (auto n : names) {
            a.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (a.lib) break;
        }
        if (!a.lib) return;
        auto sym = [&](const char *s) { return dlsym(a.lib, s); };
        a.create = reinterpret_cast<decltype(a.create)>(sym("nvrtcCreateProgram"));
        a.compile = reinterpret_cast<decltype(a.compile)>(sym("nvrtcCompileProgram"));
        a.destroy = reinterpret_cast<decltype(a.destroy)>(sym("nvrtcDestroyProgram"));
        a.cubin_size = reinterpret_cast<decltype(a.cubin_size)>(sym("nvrtcGetCUBINSize"));
        a.cubin = reinterpret_cast<decltype(a.cubin)>(sym("nvrtcGetCUBIN"));
        a.log_size = reinterpret_cast<decltype(a.log_size)>(sym("nvrtcGetProgramLogSize"));
        a.log = reinterpret_cast<decltype(a.log)>(sym("nvrtcGetProgramLog"));
        a.error_string = reinterpret_cast<decltype(a.error_string)>(sym("nvrtcGetErrorString"));
    });
    if (!a.lib || !a.create || !a.compile || !a.cubin) {
        throw error("bbfft-cuda: NVRTC (libnvrtc.so.12) could not be loaded; it is required to "
                    "build FFT kernels at plan creation",
                    -1);
    }
    return a;
}
} // namespace

// ------------------------------------------------------------------------------------------
// Persistent kernel cache (BBFFT_CUDA_KERNEL_CACHE=<dir>): cubins keyed by a hash of everything
// that determines them (stub source incl. pasted user callbacks, device header text,
// architecture, options).  The in-memory jit_cache of the reference (include/bbfft/jit_cache.hpp)
// lives for one process; this one carries NVRTC results across processes -- and across machines
// of the same image, which is how tools/tune_gpu.py ships pre-compiled candidates to the GPU box.
// ------------------------------------------------------------------------------------------
namespace {
std::uint64_t fnv1a(std::uint64_t h, char const *p, std::size_t n) {
    for (std::size_t i = 0; i < n; ++i) {
        h ^= static_cast<unsigned char>(p[i]);
        h *= 0x100000001b3ull;
    }
    return h;
}
std::string disk_cache_path(std::string const &source, std::string const &arch,
                            std::vector<std::string> const &options) {
    char const *dir = std::getenv("BBFFT_CUDA_KERNEL_CACHE");
    if (!dir || !*dir) return {};
    std::uint64_t h1 = 0xcbf29ce484222325ull, h2 = 0x84222325cbf29ce4ull;
    char const *hdr = kernel_header_text();
    h1 = fnv1a(h1, hdr, std::strlen(hdr));
    h1 = fnv1a(h1, source.data(), source.size());
    h2 = fnv1a(h2, source.data(), source.size());
    h2 = fnv1a(h2, arch.data(), arch.size());
    for (auto const &o : options) h2 = fnv1a(h2, o.data(), o.size() + 1);
    char const *li = std::getenv("BBFFT_CUDA_JIT_LINEINFO");
    if (li && *li == '0') h2 = fnv1a(h2, "nolineinfo", 10);
    char name[64];
    std::snprintf(name, sizeof(name), "/%016llx%016llx.cubin", (unsigned long long)h1, (unsigned long long)h2);
    return std::string(dir) + name;
}
} // namespace

static std::vector<std::uint8_t> nvrtc_compile_uncached(std::string const &source, std::string const &arch,
                                                        std::vector<std::string> const &extra_options);

std::vector<std::uint8_t> nvrtc_compile(std::string const &source, std::string const &arch,
                                        std::vector<std::string> const &extra_options, bool refresh) {
    const std::string path = disk_cache_path(source, arch, extra_options);
    if (!path.empty() && !refresh) {
        std::ifstream f(path, std::ios::binary);
        if (f) {
            std::vector<std::uint8_t> bin((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
            if (!bin.empty()) return bin;
        }
    }
    auto bin = nvrtc_compile_uncached(source, arch, extra_options);
    if (!path.empty()) {
        // write-then-rename so that concurrent processes never see a partial file; a cache that
        // cannot be written is not an error
        const std::string tmp = path + ".tmp" + std::to_string(static_cast<long long>(::getpid())) + "_" +
                                std::to_string(reinterpret_cast<std::uintptr_t>(&bin));
        std::ofstream f(tmp, std::ios::binary);
        if (f) {
            f.write(reinterpret_cast<char const *>(bin.data()), std::streamsize(bin.size()));
            f.close();
            if (!f || std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
        }
    }
    return bin;
}

static std::vector<std::uint8_t> nvrtc_compile_uncached(std::string const &source, std::string const &arch,
                                                        std::vector<std::string> const &extra_options) {
    auto &rt = nvrtc();
    void *prog = nullptr;
    const char *headers[] = {kernel_header_text()};
    const char *names[] = {"bbfft_kernels.cuh"};
    int rc = rt.create(&prog, source.c_str(), "bbfft_kernel.cu", 1, headers, names);
    if (rc != 0) {
        throw error(std::string("nvrtcCreateProgram failed: ") + rt.error_string(rc), rc);
    }
    std::vector<std::string> opts = {"--gpu-architecture=" + arch, "-std=c++17", "-default-device"};
    // line tables let ncu map SASS back to bbfft_kernels.cuh; BBFFT_CUDA_JIT_LINEINFO=0 drops them
    // (5x smaller cubins for the persistent cache, identical code)
    char const *li = std::getenv("BBFFT_CUDA_JIT_LINEINFO");
    if (!(li && *li == '0')) opts.push_back("-lineinfo");
    for (auto const &o : extra_options) opts.push_back(o);
    std::vector<const char *> copts;
    for (auto const &o : opts) copts.push_back(o.c_str());
    rc = rt.compile(prog, int(copts.size()), copts.data());
    if (rc != 0) {
        std::string log;
        size_t sz = 0;
        if (rt.log_size(prog, &sz) == 0 && sz > 1) {
            log.resize(sz);
            rt.log(prog, log.data());
        }
        rt.destroy(&prog);
        throw error(std::string("bbfft-cuda: kernel compilation failed (") + rt.error_string(rc) +
                        ")\n" + log,
                    rc);
    }
    size_t sz = 0;
    rc = rt.cubin_size(prog, &sz);
    std::vector<std::uint8_t> bin(sz);
    if (rc == 0 && sz > 0) rc = rt.cubin(prog, reinterpret_cast<char *>(bin.data()));
    rt.destroy(&prog);
    if (rc != 0 || sz == 0) {
        throw error("bbfft-cuda: nvrtcGetCUBIN failed", rc);
    }
    return bin;
}

auto compile_to_native(std::string const &source, std::string const &arch,
                       std::vector<std::string> const &options) -> std::vector<std::uint8_t> {
    return nvrtc_compile(source, arch, options);
}

// ------------------------------------------------------------------------------------------
// modules
// ------------------------------------------------------------------------------------------
module_handle_t load_module_image(void const *image) {
    cudaLibrary_t lib = nullptr;
    BBFFT_CUDA_CHECK(cudaLibraryLoadData(&lib, image, nullptr, nullptr, 0, nullptr, nullptr, 0));
    return reinterpret_cast<module_handle_t>(lib);
}

void unload_module(module_handle_t mod) {
    if (mod) cudaLibraryUnload(reinterpret_cast<cudaLibrary_t>(mod));
}

auto make_shared_handle(module_handle_t mod) -> shared_handle<module_handle_t> {
    return shared_handle<module_handle_t>(mod, &unload_module);
}

auto build_native_module(std::string const &source, int device, std::vector<std::string> const &options)
    -> module_handle_t {
    BBFFT_CUDA_CHECK(cudaSetDevice(device));
    api a(nullptr, device);
    auto bin = nvrtc_compile(source, a.arch(), options);
    return load_module_image(bin.data());
}

auto build_native_module(std::uint8_t const *binary, std::size_t, module_format format, int device)
    -> module_handle_t {
    if (format != module_format::native) {
        throw bad_configuration("the CUDA backend loads native modules (cubin/fatbin) only");
    }
    BBFFT_CUDA_CHECK(cudaSetDevice(device));
    return load_module_image(binary);
}

auto get_kernel_names(module_handle_t mod) -> std::vector<std::string> {
    auto lib = reinterpret_cast<cudaLibrary_t>(mod);
    unsigned int count = 0;
    BBFFT_CUDA_CHECK(cudaLibraryGetKernelCount(&count, lib));
    std::vector<cudaKernel_t> kernels(count);
    if (count) BBFFT_CUDA_CHECK(cudaLibraryEnumerateKernels(kernels.data(), count, lib));
    std::vector<std::string> names;
    for (auto k : kernels) {
        // the runtime has no name query for cudaKernel_t; go through the driver entry point
        using fn_t = int (*)(const char **, void *);
        static fn_t get_name = [] {
            void *f = nullptr;
            cudaDriverEntryPointQueryResult st;
            if (cudaGetDriverEntryPoint("cuKernelGetName", &f, cudaEnableDefault, &st) != cudaSuccess) f = nullptr;
            return reinterpret_cast<fn_t>(f);
        }();
        const char *nm = nullptr;
        if (get_name && get_name(&nm, k) == 0 && nm) names.emplace_back(nm);
    }
    return names;
}

aot_module create_aot_module(std::uint8_t const *binary, std::size_t binary_size, module_format format,
                             int device) {
    aot_module m;
    auto handle = build_native_module(binary, binary_size, format, device);
    m.mod = make_shared_handle(handle);
    for (auto &n : get_kernel_names(handle)) m.kernel_names.insert(n);
    m.device_id = query_device_id(device);
    return m;
}

// ------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------
device_props query_device_props(int device) {
    cudaDeviceProp p;
    BBFFT_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
    device_props d;
    d.sm_count = p.multiProcessorCount;
    d.max_threads_per_block = p.maxThreadsPerBlock;
    d.max_smem_per_block = p.sharedMemPerBlockOptin;
    d.smem_per_sm = p.sharedMemPerMultiprocessor;
    d.regs_per_sm = p.regsPerMultiprocessor;
    d.cc_major = p.major;
    d.cc_minor = p.minor;
    return d;
}

std::uint64_t query_device_id(int device) {
    cudaDeviceProp p;
    BBFFT_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
    // FNV-1a over uuid + compute capability: same physical device -> same id
    std::uint64_t h = 1469598103934665603ull;
    auto mix = [&](unsigned char c) {
        h ^= c;
        h *= 1099511628211ull;
    };
    for (unsigned char c : p.uuid.bytes) mix(c);
    mix(static_cast<unsigned char>(p.major));
    mix(static_cast<unsigned char>(p.minor));
    return h;
}

} // namespace bbfft::cuda

namespace bbfft {
auto get_device_info(int device) -> device_info {
    auto p = cuda::query_device_props(device);
    device_info info;
    info.max_work_group_size = std::size_t(p.max_threads_per_block);
    info.subgroup_sizes = {32};
    info.local_memory_size = p.max_smem_per_block;
    info.type = device_type::gpu;
    return info;
}
auto get_device_id(int device) -> std::uint64_t { return cuda::query_device_id(device); }
} // namespace bbfft

namespace bbfft::cuda {

// ------------------------------------------------------------------------------------------
// event
// ------------------------------------------------------------------------------------------
// Every execute returns an event (the reference's plans return sycl / cl events).  Creating and
// destroying a CUDA event per launch costs about as much as the launch itself, so retired events go
// back to a per-device free list (launch-bound callers: BASELINE config 1).
namespace {
struct event_pool {
    std::mutex mtx;
    std::vector<std::pair<int, cudaEvent_t>> free_list;
    cudaEvent_t take(int device) {
        {
            std::lock_guard<std::mutex> lock(mtx);
            for (std::size_t i = free_list.size(); i-- > 0;) {
                if (free_list[i].first == device) {
                    cudaEvent_t e = free_list[i].second;
                    free_list[i] = free_list.back();
                    free_list.pop_back();
                    return e;
                }
            }
        }
        cudaEvent_t e = nullptr;
        BBFFT_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        return e;
    }
    void give(int device, cudaEvent_t e) {
        std::lock_guard<std::mutex> lock(mtx);
        if (free_list.size() < 256) {
            free_list.emplace_back(device, e);
        } else {
            cudaEventDestroy(e);
        }
    }
};
event_pool &events() {
    static event_pool *p = new event_pool(); // leaked on purpose: events may be retired during exit
    return *p;
}
} // namespace

event::event(cudaStream_t stream) {
    int device = 0;
    BBFFT_CUDA_CHECK(cudaGetDevice(&device));
    cudaEvent_t e = events().take(device);
    ev_ = std::shared_ptr<cudaEvent_t>(new cudaEvent_t(e), [device](cudaEvent_t *p) {
        if (*p) events().give(device, *p);
        delete p;
    });
    BBFFT_CUDA_CHECK(cudaEventRecord(e, stream));
}

void event::wait() const {
    if (ev_) BBFFT_CUDA_CHECK(cudaEventSynchronize(*ev_));
}

// ------------------------------------------------------------------------------------------
// api
// ------------------------------------------------------------------------------------------
api::api(cudaStream_t stream, int device) : stream_(stream), device_(device) {
    if (device_ < 0) BBFFT_CUDA_CHECK(cudaGetDevice(&device_));
    props_ = query_device_props(device_);
    device_id_ = query_device_id(device_);
}

device_info api::info() const { return get_device_info(device_); }

std::string api::arch() const {
    std::ostringstream os;
    os << "sm_" << props_.cc_major << props_.cc_minor;
    if (props_.cc_major >= 9) os << "a";
    return os.str();
}

std::size_t api::kernel_local_bytes(cudaKernel_t k) const {
    cudaFuncAttributes attr;
    BBFFT_CUDA_CHECK(cudaFuncGetAttributes(&attr, reinterpret_cast<const void *>(k)));
    return attr.localSizeBytes;
}

shared_handle<module_handle_t> api::build_module(std::string const &source,
                                                 std::vector<std::string> const &options) const {
    auto bin = nvrtc_compile(source, arch(), options);
    try {
        return make_shared_handle(load_module_image(bin.data()));
    } catch (error const &) {
        // a damaged entry of the persistent kernel cache must not break plan creation: compile again
        // (overwriting the entry) and let a second failure propagate
        if (!std::getenv("BBFFT_CUDA_KERNEL_CACHE")) throw;
        bin = nvrtc_compile(source, arch(), options, true);
        return make_shared_handle(load_module_image(bin.data()));
    }
}

cudaKernel_t api::create_kernel(module_handle_t mod, std::string const &name, std::size_t smem_bytes) const {
    cudaKernel_t k = nullptr;
    BBFFT_CUDA_CHECK(cudaLibraryGetKernel(&k, reinterpret_cast<cudaLibrary_t>(mod), name.c_str()));
    if (smem_bytes > 48 * 1024) {
        BBFFT_CUDA_CHECK(cudaFuncSetAttribute(reinterpret_cast<const void *>(k),
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes)));
    }
    // The planner sizes occupancy by registers AND shared memory (mb CTAs x smem_bytes); without a stated
    // preference the driver picks the L1 / shared-memory split per launch and may keep the previous kernel's
    // split when one CTA fits, so a kernel's occupancy -- and time -- would depend on which kernel ran before
    // it (seen as +-20 % on single sizes between two runs of the same sweep, profiles/r02l_bench_ab2.txt).
    // Measured (profiles/r02m_carveout.txt): asking for the largest shared-memory carve-out costs the sweep a
    // sixth of its throughput (0.957 -> 0.79 of the HBM peak) -- the kernels live on their L1 (half-line rows
    // shared by neighbouring CTAs, twiddle tables) -- and does not move the sizes in question.  The preference
    // stays a switch, OFF by default: BBFFT_CUDA_CARVEOUT=1.
    static const bool carve = [] {
        char const *e = std::getenv("BBFFT_CUDA_CARVEOUT");
        return e && *e == '1';
    }();
    if (carve && smem_bytes > 0) {
        BBFFT_CUDA_CHECK(cudaFuncSetAttribute(reinterpret_cast<const void *>(k),
                                              cudaFuncAttributePreferredSharedMemoryCarveout,
                                              int(cudaSharedmemCarveoutMaxShared)));
    }
    return k;
}

// Programmatic dependent launch (see bbk::pdl_prologue in the device header): launches of SMALL grids
// -- at most two waves of CTAs, where the launch latency is a visible share of the kernel (measured on
// BASELINE config 1: 6.15 -> 4.45 us per back-to-back execute, profiles/r02e_tile_pdl.txt) -- carry
// the programmatic-stream-serialization attribute.  Large grids gain nothing, and inside a stream
// capture the programmatic edge costs time (3.77 -> 4.41 us per graph node), so captured launches
// and large grids are launched plainly.  BBFFT_CUDA_PDL=0 switches it off, =2 forces it everywhere.
static int pdl_mode() {
    static const int mode = [] {
        char const *e = std::getenv("BBFFT_CUDA_PDL");
        return e ? std::atoi(e) : 1;
    }();
    return mode;
}

void api::launch_kernel(cudaKernel_t k, std::uint64_t grid, int threads, std::size_t smem_bytes,
                        kernel_args const &args, cudaStream_t stream) const {
    if (grid == 0) return;
    if (grid > 0x7fffffffull) {
        throw bad_configuration("bbfft-cuda: batch too large for one launch");
    }
    kernel_args a = args;
    void *params[] = {&a};
    bool pdl = pdl_mode() == 2;
    if (pdl_mode() == 1 && std::uint64_t(threads) * grid <= 2ull * 2048ull * std::uint64_t(props_.sm_count)) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        pdl = cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone;
    }
    if (pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(unsigned(grid));
        cfg.blockDim = dim3(unsigned(threads));
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        BBFFT_CUDA_CHECK(cudaLaunchKernelExC(&cfg, reinterpret_cast<const void *>(k), params));
        return;
    }
    BBFFT_CUDA_CHECK(cudaLaunchKernel(reinterpret_cast<const void *>(k), dim3(unsigned(grid)),
                                      dim3(unsigned(threads)), params, smem_bytes, stream));
}

void api::launch_kernel_raw(cudaKernel_t k, std::uint64_t grid, int threads, std::size_t smem_bytes,
                            void *param, cudaStream_t stream) const {
    if (grid == 0) return;
    void *params[] = {param};
    BBFFT_CUDA_CHECK(cudaLaunchKernel(reinterpret_cast<const void *>(k), dim3(unsigned(grid)),
                                      dim3(unsigned(threads)), params, smem_bytes, stream));
}

int api::max_active_ctas_per_sm(cudaKernel_t k, int threads, std::size_t smem_bytes) const {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, reinterpret_cast<const void *>(k), threads, smem_bytes) !=
        cudaSuccess) {
        (void)cudaGetLastError(); // not every runtime accepts a cudaKernel_t here: the caller falls back
        return 0;
    }
    return n;
}

void *api::create_device_buffer(std::size_t bytes) const {
    void *p = nullptr;
    BBFFT_CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 1));
    return p;
}

void api::release_buffer(void *ptr) const {
    if (ptr) cudaFree(ptr);
}

void *api::create_twiddle_table(std::vector<double> const &tw, int fp) const {
    void *dev = create_device_buffer(tw.size() * std::size_t(fp));
    // The copy is issued on the plan's own stream and completed before the plan exists: a
    // cudaMemcpy from pageable memory may return before its DMA has landed and is ordered on the
    // legacy stream only, which a cudaStreamNonBlocking user stream does not wait for.
    if (fp == 4) {
        std::vector<float> f(tw.begin(), tw.end());
        BBFFT_CUDA_CHECK(cudaMemcpyAsync(dev, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice, stream_));
        BBFFT_CUDA_CHECK(cudaStreamSynchronize(stream_));
    } else {
        BBFFT_CUDA_CHECK(cudaMemcpyAsync(dev, tw.data(), tw.size() * sizeof(double), cudaMemcpyHostToDevice, stream_));
        BBFFT_CUDA_CHECK(cudaStreamSynchronize(stream_));
    }
    return dev;
}

} // namespace bbfft::cuda
