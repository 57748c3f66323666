// c_abi.cpp -- extern "C" boundary of the CUDA backend (declared in include/bbfft_cuda.h).
#pragma GCC visibility push(default)
#include "bbfft_cuda.h"
#pragma GCC visibility pop

#include "bbfft/api.hpp"
#include "bbfft/cuda/error.hpp"
#include "bbfft/cuda/make_plan.hpp"
#include "bbfft/cuda/online_compiler.hpp"
#include "plan.hpp"
#include "planner.hpp"
#include "runtime.hpp"

#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <cstring>
#include <sstream>
#include <string>

using namespace bbfft;

namespace {
thread_local std::string g_last_error;

template <class F> int guarded(F &&f) {
    try {
        f();
        return BBFFT_CUDA_OK;
    } catch (bad_configuration const &e) {
        g_last_error = e.what();
        return BBFFT_CUDA_BAD_CONFIGURATION;
    } catch (cuda::error const &e) {
        g_last_error = e.what();
        return BBFFT_CUDA_DEVICE_ERROR;
    } catch (std::exception const &e) {
        g_last_error = e.what();
        return BBFFT_CUDA_ERROR;
    } catch (...) {
        g_last_error = "unknown error";
        return BBFFT_CUDA_ERROR;
    }
}

configuration to_cpp(bbfft_cuda_config const &c) {
    configuration cfg = {};
    cfg.dim = c.dim;
    for (unsigned i = 0; i < max_tensor_dim; ++i) {
        cfg.shape[i] = c.shape[i];
        cfg.istride[i] = c.istride[i];
        cfg.ostride[i] = c.ostride[i];
    }
    cfg.fp = static_cast<precision>(c.fp);
    cfg.dir = static_cast<direction>(c.dir);
    cfg.type = static_cast<transform_type>(c.type);
    if (c.cb_source && c.cb_length) {
        cfg.callbacks.data = c.cb_source;
        cfg.callbacks.length = c.cb_length;
        cfg.callbacks.load_function = (c.cb_load && *c.cb_load) ? c.cb_load : nullptr;
        cfg.callbacks.store_function = (c.cb_store && *c.cb_store) ? c.cb_store : nullptr;
        cfg.callbacks.language = c.cb_language == 1 ? kernel_language::cuda_c : kernel_language::opencl_c;
    }
    return cfg;
}

void from_cpp(configuration const &cfg, bbfft_cuda_config &c) {
    std::memset(&c, 0, sizeof(c));
    c.dim = cfg.dim;
    for (unsigned i = 0; i < max_tensor_dim; ++i) {
        c.shape[i] = cfg.shape[i];
        c.istride[i] = cfg.istride[i];
        c.ostride[i] = cfg.ostride[i];
    }
    c.fp = static_cast<int>(cfg.fp);
    c.dir = static_cast<int>(cfg.dir);
    c.type = static_cast<int>(cfg.type);
}

char *dup_string(std::string const &s) {
    char *p = static_cast<char *>(std::malloc(s.size() + 1));
    std::memcpy(p, s.c_str(), s.size() + 1);
    return p;
}
} // namespace

struct bbfft_cuda_cache_s {
    jit_cache_all cache;
};

struct bbfft_cuda_plan_s {
    std::shared_ptr<cuda::plan_base> impl;
    std::vector<std::string> kernel_names;
    int device = 0;
};

namespace {
// Device staging for the host-buffer entry point: one pair of buffers and two copy streams per
// device, shared by all plans (grown on demand, released at process exit).
struct staging {
    std::mutex mtx;
    void *in = nullptr, *out = nullptr;
    size_t in_bytes = 0, out_bytes = 0;
    cudaStream_t streams[2] = {nullptr, nullptr};
    cudaEvent_t ev = nullptr;
    void reserve(size_t need_in, size_t need_out) {
        if (in_bytes < need_in) {
            if (in) cudaFree(in);
            in = nullptr;
            in_bytes = 0;
            BBFFT_CUDA_CHECK(cudaMalloc(&in, need_in));
            in_bytes = need_in;
        }
        if (out_bytes < need_out) {
            if (out) cudaFree(out);
            out = nullptr;
            out_bytes = 0;
            BBFFT_CUDA_CHECK(cudaMalloc(&out, need_out));
            out_bytes = need_out;
        }
    }
    void ensure_streams() {
        if (!streams[0]) {
            BBFFT_CUDA_CHECK(cudaStreamCreateWithFlags(&streams[0], cudaStreamNonBlocking));
            BBFFT_CUDA_CHECK(cudaStreamCreateWithFlags(&streams[1], cudaStreamNonBlocking));
            BBFFT_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        }
    }
};
staging &staging_for(int device) {
    static std::mutex m;
    static std::map<int, std::unique_ptr<staging>> all;
    std::lock_guard<std::mutex> lock(m);
    auto &p = all[device];
    if (!p) p = std::make_unique<staging>();
    return *p;
}
} // namespace

extern "C" {

const char *bbfft_cuda_last_error(void) { return g_last_error.c_str(); }

int bbfft_cuda_default_strides(const bbfft_cuda_config *cfg, int inplace, size_t *istride, size_t *ostride) {
    return guarded([&] {
        tensor_extent shape;
        for (unsigned i = 0; i < max_tensor_dim; ++i) shape[i] = cfg->shape[i];
        auto is = default_istride(cfg->dim, shape, static_cast<transform_type>(cfg->type), inplace != 0);
        auto os = default_ostride(cfg->dim, shape, static_cast<transform_type>(cfg->type), inplace != 0);
        for (unsigned i = 0; i < max_tensor_dim; ++i) {
            istride[i] = is[i];
            ostride[i] = os[i];
        }
    });
}

int bbfft_cuda_parse_descriptor(const char *descriptor, bbfft_cuda_config *cfg) {
    return guarded([&] { from_cpp(parse_fft_descriptor(descriptor), *cfg); });
}

int bbfft_cuda_to_descriptor(const bbfft_cuda_config *cfg, char *buffer, size_t buffer_size) {
    return guarded([&] {
        auto s = to_cpp(*cfg).to_string();
        if (s.size() + 1 > buffer_size) throw std::runtime_error("descriptor buffer too small");
        std::memcpy(buffer, s.c_str(), s.size() + 1);
    });
}

int bbfft_cuda_cache_create(bbfft_cuda_cache_t *cache) {
    return guarded([&] { *cache = new bbfft_cuda_cache_s(); });
}
int bbfft_cuda_cache_destroy(bbfft_cuda_cache_t cache) {
    return guarded([&] { delete cache; });
}
int bbfft_cuda_cache_size(bbfft_cuda_cache_t cache) {
    return cache ? int(cache->cache.kernel_names().size()) : 0;
}

int bbfft_cuda_plan_create_tuned(bbfft_cuda_plan_t *plan, const bbfft_cuda_config *cfg, void *stream,
                                 int device, bbfft_cuda_cache_t cache, const char *tune) {
    return guarded([&] {
        *plan = nullptr;
        auto c = to_cpp(*cfg);
        if (device >= 0) BBFFT_CUDA_CHECK(cudaSetDevice(device));
        cuda::api a(static_cast<cudaStream_t>(stream), device);
        auto p = std::make_unique<bbfft_cuda_plan_s>();
        p->device = a.device();
        jit_cache *jc = cache ? &cache->cache : nullptr;
        if (tune && *tune && c.dim == 2) {
            // 2d: overrides of the fused tile kernel (RA, RB, TH, PADK, MB)
            auto steps = cuda::nd_decompose(c, a.props());
            if (steps.size() != 1 || !steps[0].fused) {
                throw bad_configuration("tuned 2d plans need a configuration that runs as one fused tile kernel");
            }
            p->impl = std::make_shared<cuda::fft2d_plan>(steps[0].tile, a, jc, tune);
        } else if (tune && *tune) {
            if (c.dim != 1) throw bad_configuration("tuned plans are 1d or fused 2d only");
            p->impl = std::make_shared<cuda::fft1d_plan>(c, a, jc, tune);
        } else {
            p->impl = cuda::select_fft_algorithm(c, a, jc);
        }
        if (auto pb = std::dynamic_pointer_cast<cuda::plan_base>(p->impl)) pb->kernel_names(p->kernel_names);
        *plan = p.release();
    });
}

int bbfft_cuda_plan_create(bbfft_cuda_plan_t *plan, const bbfft_cuda_config *cfg, void *stream, int device,
                           bbfft_cuda_cache_t cache) {
    return bbfft_cuda_plan_create_tuned(plan, cfg, stream, device, cache, nullptr);
}

int bbfft_cuda_plan_execute(bbfft_cuda_plan_t plan, const void *in, void *out) {
    return guarded([&] { plan->impl->enqueue(in, out, plan->impl->stream()); });
}

int bbfft_cuda_plan_execute_on(bbfft_cuda_plan_t plan, const void *in, void *out, void *stream) {
    return guarded([&] { plan->impl->enqueue(in, out, static_cast<cudaStream_t>(stream)); });
}

int bbfft_cuda_plan_execute_host(bbfft_cuda_plan_t plan, const void *host_in, size_t in_bytes, void *host_out,
                                 size_t out_bytes) {
    return guarded([&] {
        cudaStream_t s = plan->impl->stream();
        const bool inplace = host_in == host_out;
        auto &st = staging_for(plan->device);
        std::lock_guard<std::mutex> lock(st.mtx);
        st.reserve(inplace ? std::max(in_bytes, out_bytes) : in_bytes, inplace ? 0 : out_bytes);
        void *din = st.in;
        void *dout = inplace ? st.in : st.out;
        const std::uint64_t K = plan->impl->slices();
        const std::size_t isl = plan->impl->in_slice_bytes(), osl = plan->impl->out_slice_bytes();
        // Sliceable plans are pipelined over k slabs on two streams so that the H2D copy of slab
        // c+1 overlaps the kernel and the D2H copy of slab c (PCIe is full duplex).
        std::uint64_t chunks = 1;
        if (K > 1 && isl > 0 && osl > 0 && K * isl <= in_bytes + isl && (in_bytes + out_bytes) > (32u << 20)) {
            // slabs of >= 16 MiB, at most 16: the un-overlapped head (first H2D) and tail (last D2H)
            // of the synchronous call shrink with the slab size
            chunks = std::min<std::uint64_t>(std::min<std::uint64_t>(16, K / 2), std::max<std::size_t>(2, in_bytes >> 24));
        }
        if (chunks <= 1) {
            BBFFT_CUDA_CHECK(cudaMemcpyAsync(din, host_in, in_bytes, cudaMemcpyHostToDevice, s));
            plan->impl->enqueue(din, dout, s);
            BBFFT_CUDA_CHECK(cudaMemcpyAsync(host_out, dout, out_bytes, cudaMemcpyDeviceToHost, s));
            BBFFT_CUDA_CHECK(cudaStreamSynchronize(s));
            return;
        }
        st.ensure_streams();
        BBFFT_CUDA_CHECK(cudaEventRecord(st.ev, s));
        std::uint64_t per = ((K + chunks - 1) / chunks + 1) & ~std::uint64_t(1); // even: odd-N real pairs
        for (std::uint64_t c = 0, k0 = 0; k0 < K; ++c, k0 += per) {
            std::uint64_t cnt = std::min(per, K - k0);
            cudaStream_t cs = st.streams[c % 2];
            if (c < 2) BBFFT_CUDA_CHECK(cudaStreamWaitEvent(cs, st.ev, 0));
            std::size_t ioff = k0 * isl, ooff = k0 * osl;
            std::size_t ib = std::min(cnt * isl, in_bytes > ioff ? in_bytes - ioff : 0);
            std::size_t ob = std::min(cnt * osl, out_bytes > ooff ? out_bytes - ooff : 0);
            BBFFT_CUDA_CHECK(cudaMemcpyAsync(static_cast<char *>(din) + ioff, static_cast<char const *>(host_in) + ioff,
                                             ib, cudaMemcpyHostToDevice, cs));
            plan->impl->enqueue_slab(din, dout, k0, cnt, cs);
            BBFFT_CUDA_CHECK(cudaMemcpyAsync(static_cast<char *>(host_out) + ooff, static_cast<char *>(dout) + ooff, ob,
                                             cudaMemcpyDeviceToHost, cs));
        }
        BBFFT_CUDA_CHECK(cudaStreamSynchronize(st.streams[0]));
        BBFFT_CUDA_CHECK(cudaStreamSynchronize(st.streams[1]));
    });
}

int bbfft_cuda_plan_destroy(bbfft_cuda_plan_t plan) {
    return guarded([&] {
        delete plan;
    });
}

int bbfft_cuda_plan_num_kernels(bbfft_cuda_plan_t plan) { return plan ? int(plan->kernel_names.size()) : 0; }

int bbfft_cuda_plan_launches(bbfft_cuda_plan_t plan) {
    return plan && plan->impl ? int(plan->impl->launches_per_execute()) : 0;
}

const char *bbfft_cuda_plan_kernel_name(bbfft_cuda_plan_t plan, int index) {
    if (!plan || index < 0 || index >= int(plan->kernel_names.size())) return "";
    return plan->kernel_names[index].c_str();
}

int bbfft_cuda_describe(const bbfft_cuda_config *cfg, const char *tune, bbfft_cuda_kernel_desc *desc) {
    return guarded([&] {
        std::memset(desc, 0, sizeof(*desc));
        auto c = to_cpp(*cfg);
        if (c.dim == 2) {
            // fused 2d tile kernel (the only single-kernel 2d plan)
            auto steps = cuda::nd_decompose(c, cuda::device_props{});
            if (steps.size() != 1 || !steps[0].fused) {
                throw bad_configuration("bbfft_cuda_describe: this 2d configuration is not a single fused kernel");
            }
            auto tp = cuda::plan_kernel_2d(steps[0].tile, cuda::device_props{}, tune ? tune : "");
            desc->identifier = dup_string(tp.identifier);
            desc->source = dup_string(tp.source);
            desc->twiddle_len = tp.twiddle.size();
            desc->twiddle = static_cast<double *>(std::malloc(sizeof(double) * tp.twiddle.size()));
            std::memcpy(desc->twiddle, tp.twiddle.data(), sizeof(double) * tp.twiddle.size());
            desc->grid = steps[0].tile.K;
            desc->threads = tp.p.threads;
            desc->smem_bytes = tp.p.smem_bytes;
            desc->fp = tp.p.fp;
            desc->n_stages = tp.p.a.L + tp.p.b.L;
            for (int s = 0; s < 4; ++s) desc->radix[s] = tp.p.a.radix[s];
            desc->threads_per_transform = tp.p.threads;
            desc->batch_lanes = tp.p.PADK;
            desc->batch_high = tp.p.min_blocks;
            return;
        }
        if (c.dim != 1) throw bad_configuration("bbfft_cuda_describe handles 1d and fused 2d configurations");
        auto kp = cuda::plan_kernel_1d(cuda::to_problem(c), cuda::device_props{}, tune ? tune : "");
        desc->identifier = dup_string(kp.identifier);
        desc->source = dup_string(kp.source);
        desc->twiddle_len = kp.twiddle.size();
        desc->twiddle = static_cast<double *>(std::malloc(sizeof(double) * kp.twiddle.size()));
        std::memcpy(desc->twiddle, kp.twiddle.data(), sizeof(double) * kp.twiddle.size());
        desc->grid = kp.p.grid(c.shape[2]);
        desc->threads = kp.p.threads;
        desc->smem_bytes = kp.p.smem_bytes;
        desc->inplace_unsupported = kp.inplace_unsupported;
        desc->fp = kp.p.fp;
        desc->n_stages = kp.p.L;
        for (int s = 0; s < 4; ++s) desc->radix[s] = kp.p.radix[s];
        desc->threads_per_transform = kp.p.T;
        desc->batch_lanes = kp.p.ML;
        desc->batch_high = kp.p.BH;
        desc->load_staged = kp.p.load_staged;
        desc->store_staged = kp.p.store_staged;
    });
}

int bbfft_cuda_describe_chain(const bbfft_cuda_config *cfg, bbfft_cuda_chain_desc *desc) {
    return guarded([&] {
        std::memset(desc, 0, sizeof(*desc));
        auto c = to_cpp(*cfg);
        if (c.dim < 2) throw bad_configuration("bbfft_cuda_describe_chain handles 2d and 3d configurations");
        auto steps = cuda::nd_decompose(c, cuda::device_props{});
        std::vector<cuda::chain_step_problem> probs;
        for (auto const &s : steps) {
            cuda::chain_step_problem q;
            q.tile = s.fused;
            q.mult = s.mult;
            if (s.fused) {
                q.t = s.tile;
            } else {
                q.p = cuda::to_problem(s.pass);
            }
            probs.push_back(q);
        }
        cuda::chain_plan_t cp;
        if (!cuda::plan_chain(probs, cuda::device_props{}, cp)) {
            throw bad_configuration("bbfft_cuda_describe_chain: the steps of this configuration cannot be chained");
        }
        desc->identifier = dup_string(cp.identifier);
        desc->source = dup_string(cp.source);
        desc->twiddle_len = cp.twiddle.size();
        desc->twiddle = static_cast<double *>(std::malloc(sizeof(double) * std::max<std::size_t>(1, cp.twiddle.size())));
        std::memcpy(desc->twiddle, cp.twiddle.data(), sizeof(double) * cp.twiddle.size());
        desc->threads = cp.threads;
        desc->smem_bytes = cp.smem_bytes;
        desc->min_blocks = cp.min_blocks;
        desc->fp = static_cast<int>(c.fp);
        desc->n_steps = int(cp.steps.size());
        for (std::size_t d = 0; d < cp.steps.size(); ++d) {
            desc->step_tile[d] = cp.steps[d].tile;
            desc->per_k[d] = cp.steps[d].per_k;
            desc->mult[d] = probs[d].mult;
            desc->step_M[d] = cp.steps[d].tile ? cp.steps[d].tp.p.M : cp.steps[d].kp.p.M;
            desc->tw_offset[d] = cp.steps[d].tw_offset;
        }
        desc->uses_tmp = c.type == transform_type::c2r;
    });
}

void bbfft_cuda_chain_desc_free(bbfft_cuda_chain_desc *desc) {
    if (!desc) return;
    std::free(desc->identifier);
    std::free(desc->source);
    std::free(desc->twiddle);
    std::memset(desc, 0, sizeof(*desc));
}

void bbfft_cuda_desc_free(bbfft_cuda_kernel_desc *desc) {
    if (!desc) return;
    std::free(desc->identifier);
    std::free(desc->source);
    std::free(desc->twiddle);
    std::memset(desc, 0, sizeof(*desc));
}

int bbfft_cuda_generate_kernels(const bbfft_cuda_config *cfgs, size_t n, char **source, char **names) {
    return guarded([&] {
        std::vector<configuration> v;
        for (size_t i = 0; i < n; ++i) v.push_back(to_cpp(cfgs[i]));
        std::ostringstream os;
        device_info info;
        auto ks = generate_fft_kernels(os, v, info);
        *source = dup_string(os.str());
        std::string joined;
        for (auto const &k : ks) joined += k + "\n";
        *names = dup_string(joined);
    });
}

const char *bbfft_cuda_kernel_header(void) { return cuda::kernel_header_text(); }

int bbfft_cuda_compile(const char *source, const char *arch, uint8_t **binary, size_t *binary_size) {
    return guarded([&] {
        auto bin = cuda::compile_to_native(source, arch ? arch : "sm_100a", {});
        *binary = static_cast<uint8_t *>(std::malloc(bin.size()));
        std::memcpy(*binary, bin.data(), bin.size());
        *binary_size = bin.size();
    });
}

void bbfft_cuda_free(void *ptr) { std::free(ptr); }

} // extern "C"
