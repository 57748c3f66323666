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
    void *dev_in = nullptr, *dev_out = nullptr;
    size_t dev_in_bytes = 0, dev_out_bytes = 0;
};

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
        if (tune && *tune) {
            if (c.dim != 1) throw bad_configuration("tuned plans are 1d only");
            p->impl = std::make_shared<cuda::fft1d_plan>(c, a, jc, tune);
        } else {
            p->impl = cuda::select_fft_algorithm(c, a, jc);
        }
        if (auto one = std::dynamic_pointer_cast<cuda::fft1d_plan>(p->impl)) {
            p->kernel_names.push_back(one->kernel().identifier);
        } else if (auto nd = std::dynamic_pointer_cast<cuda::nd_plan>(p->impl)) {
            for (auto const &q : nd->passes()) p->kernel_names.push_back(q->kernel().identifier);
        }
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
        size_t need_in = inplace ? std::max(in_bytes, out_bytes) : in_bytes;
        if (plan->dev_in_bytes < need_in) {
            if (plan->dev_in) cudaFree(plan->dev_in);
            BBFFT_CUDA_CHECK(cudaMalloc(&plan->dev_in, need_in));
            plan->dev_in_bytes = need_in;
        }
        if (!inplace && plan->dev_out_bytes < out_bytes) {
            if (plan->dev_out) cudaFree(plan->dev_out);
            BBFFT_CUDA_CHECK(cudaMalloc(&plan->dev_out, out_bytes));
            plan->dev_out_bytes = out_bytes;
        }
        void *din = plan->dev_in;
        void *dout = inplace ? plan->dev_in : plan->dev_out;
        BBFFT_CUDA_CHECK(cudaMemcpyAsync(din, host_in, in_bytes, cudaMemcpyHostToDevice, s));
        plan->impl->enqueue(din, dout, s);
        BBFFT_CUDA_CHECK(cudaMemcpyAsync(host_out, dout, out_bytes, cudaMemcpyDeviceToHost, s));
        BBFFT_CUDA_CHECK(cudaStreamSynchronize(s));
    });
}

int bbfft_cuda_plan_destroy(bbfft_cuda_plan_t plan) {
    return guarded([&] {
        if (!plan) return;
        if (plan->dev_in) cudaFree(plan->dev_in);
        if (plan->dev_out) cudaFree(plan->dev_out);
        delete plan;
    });
}

int bbfft_cuda_plan_num_kernels(bbfft_cuda_plan_t plan) { return plan ? int(plan->kernel_names.size()) : 0; }

const char *bbfft_cuda_plan_kernel_name(bbfft_cuda_plan_t plan, int index) {
    if (!plan || index < 0 || index >= int(plan->kernel_names.size())) return "";
    return plan->kernel_names[index].c_str();
}

int bbfft_cuda_describe(const bbfft_cuda_config *cfg, const char *tune, bbfft_cuda_kernel_desc *desc) {
    return guarded([&] {
        std::memset(desc, 0, sizeof(*desc));
        auto c = to_cpp(*cfg);
        if (c.dim != 1) throw bad_configuration("bbfft_cuda_describe handles 1d configurations");
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
