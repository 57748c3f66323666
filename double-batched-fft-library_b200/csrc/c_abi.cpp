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
// Host-buffer entry point (bbfft_cuda_plan_execute_host): a per-device ring of device slots and three
// streams.  Slab c of a plan goes  H2D (copy-in stream) -> kernel (compute stream) -> D2H (copy-out
// stream), chained by events, while slab c+1 is already on its way in and slab c-1 on its way out:
// both PCIe directions stay busy and the device footprint is SLOTS * 2 * slot bytes whatever the
// tensor size.  Slots are handed out round-robin under a mutex that only guards the bookkeeping, so
// concurrent callers (several plans, several host threads) interleave slab by slab instead of
// serialising whole transforms; nothing is freed or reallocated between plans.  Plans whose k slices
// are not contiguous byte ranges (or nd plans) take the whole-tensor path below.
struct host_ring {
    static constexpr int SLOTS = 4;
    std::mutex mtx;
    std::size_t slot_bytes = 0;
    void *in[SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    void *out[SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t loaded[SLOTS], computed[SLOTS], drained[SLOTS];
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaEvent_t entry = nullptr;
    unsigned next = 0;
    // whole-tensor path: grow-only buffers
    void *whole_in = nullptr, *whole_out = nullptr;
    std::size_t whole_in_bytes = 0, whole_out_bytes = 0;

    void init(std::size_t want_slot) {
        if (!s_in) {
            BBFFT_CUDA_CHECK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
            BBFFT_CUDA_CHECK(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking));
            BBFFT_CUDA_CHECK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
            BBFFT_CUDA_CHECK(cudaEventCreateWithFlags(&entry, cudaEventDisableTiming));
            for (int i = 0; i < SLOTS; ++i) {
                BBFFT_CUDA_CHECK(cudaEventCreateWithFlags(&loaded[i], cudaEventDisableTiming));
                BBFFT_CUDA_CHECK(cudaEventCreateWithFlags(&computed[i], cudaEventDisableTiming));
                BBFFT_CUDA_CHECK(cudaEventCreateWithFlags(&drained[i], cudaEventDisableTiming));
            }
        }
        if (slot_bytes < want_slot) {
            // first use, or a plan whose single slice exceeds the slot: drain all three stages and grow once
            BBFFT_CUDA_CHECK(cudaStreamSynchronize(s_in));
            BBFFT_CUDA_CHECK(cudaStreamSynchronize(s_k));
            BBFFT_CUDA_CHECK(cudaStreamSynchronize(s_out));
            for (int i = 0; i < SLOTS; ++i) {
                if (in[i]) cudaFree(in[i]);
                if (out[i]) cudaFree(out[i]);
                in[i] = out[i] = nullptr;
            }
            slot_bytes = 0;
            for (int i = 0; i < SLOTS; ++i) {
                BBFFT_CUDA_CHECK(cudaMalloc(&in[i], want_slot));
                BBFFT_CUDA_CHECK(cudaMalloc(&out[i], want_slot));
            }
            slot_bytes = want_slot;
        }
    }
    void reserve_whole(std::size_t need_in, std::size_t need_out) {
        if (whole_in_bytes < need_in) {
            if (whole_in) cudaFree(whole_in);
            whole_in = nullptr;
            whole_in_bytes = 0;
            BBFFT_CUDA_CHECK(cudaMalloc(&whole_in, need_in));
            whole_in_bytes = need_in;
        }
        if (whole_out_bytes < need_out) {
            if (whole_out) cudaFree(whole_out);
            whole_out = nullptr;
            whole_out_bytes = 0;
            BBFFT_CUDA_CHECK(cudaMalloc(&whole_out, need_out));
            whole_out_bytes = need_out;
        }
    }
};
host_ring &ring_for(int device) {
    static std::mutex m;
    static std::map<int, std::unique_ptr<host_ring>> all;
    std::lock_guard<std::mutex> lock(m);
    auto &p = all[device];
    if (!p) p = std::make_unique<host_ring>();
    return *p;
}
std::size_t default_slot_bytes() {
    // 16 MiB slabs: the un-overlapped head (first H2D) and tail (last D2H) of a 1 GiB transform are
    // 2 of 66 pipeline steps; BBFFT_CUDA_HOST_SLOT_MB overrides
    std::size_t mb = 16;
    if (char const *e = std::getenv("BBFFT_CUDA_HOST_SLOT_MB")) mb = std::max(1l, std::atol(e));
    return mb << 20;
}
struct device_scope {
    int prev = -1;
    explicit device_scope(int device) {
        int cur = -1;
        BBFFT_CUDA_CHECK(cudaGetDevice(&cur));
        if (cur != device) {
            BBFFT_CUDA_CHECK(cudaSetDevice(device));
            prev = cur;
        }
    }
    ~device_scope() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
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
        auto &impl = *plan->impl;
        if (in_bytes < impl.in_bytes_required() || out_bytes < impl.out_bytes_required()) {
            throw bad_configuration("bbfft_cuda_plan_execute_host: host buffer smaller than the plan's tensor (" +
                                    std::to_string(in_bytes) + " / " + std::to_string(out_bytes) + " bytes given, " +
                                    std::to_string(impl.in_bytes_required()) + " / " +
                                    std::to_string(impl.out_bytes_required()) + " needed)");
        }
        device_scope dev(plan->device);
        cudaStream_t s = impl.stream();
        const bool inplace = host_in == host_out;
        auto &ring = ring_for(plan->device);
        const std::uint64_t K = impl.slices();
        const std::size_t isl = impl.in_slice_bytes(), osl = impl.out_slice_bytes();
        const std::size_t slice_max = std::max(isl, osl);
        const bool sliceable = K > 1 && isl > 0 && osl > 0 && impl.slices_contiguous();
        if (!sliceable || K * slice_max <= (8u << 20)) {
            // one copy in, one launch, one copy out on the plan's stream
            std::lock_guard<std::mutex> lock(ring.mtx);
            ring.reserve_whole(inplace ? std::max(in_bytes, out_bytes) : in_bytes, inplace ? 0 : out_bytes);
            void *din = ring.whole_in;
            void *dout = inplace ? ring.whole_in : ring.whole_out;
            BBFFT_CUDA_CHECK(cudaMemcpyAsync(din, host_in, in_bytes, cudaMemcpyHostToDevice, s));
            impl.enqueue(din, dout, s);
            BBFFT_CUDA_CHECK(cudaMemcpyAsync(host_out, dout, out_bytes, cudaMemcpyDeviceToHost, s));
            BBFFT_CUDA_CHECK(cudaStreamSynchronize(s));
            return;
        }
        std::size_t slot = default_slot_bytes();
        if (slice_max * 2 > slot) slot = slice_max * 2; // at least one pair of slices per slab
        {
            std::lock_guard<std::mutex> lock(ring.mtx);
            ring.init(slot);
            slot = ring.slot_bytes;
            // everything the caller queued on the plan's stream precedes the first copy
            BBFFT_CUDA_CHECK(cudaEventRecord(ring.entry, s));
            BBFFT_CUDA_CHECK(cudaStreamWaitEvent(ring.s_in, ring.entry, 0));
        }
        // slices per slab: even (odd-N real transforms pair the slices 2k', 2k'+1)
        std::uint64_t per = std::max<std::uint64_t>(2, (slot / slice_max) & ~std::uint64_t(1));
        cudaEvent_t last = nullptr;
        for (std::uint64_t k0 = 0; k0 < K; k0 += per) {
            const std::uint64_t cnt = std::min(per, K - k0);
            const std::size_t ioff = k0 * isl, ooff = k0 * osl;
            const std::size_t ib = std::min(cnt * isl, in_bytes - ioff);
            const std::size_t ob = std::min(cnt * osl, out_bytes - ooff);
            std::lock_guard<std::mutex> lock(ring.mtx); // one slab is issued as a unit
            const unsigned i = ring.next++ % host_ring::SLOTS;
            void *din = ring.in[i];
            void *dout = inplace ? ring.in[i] : ring.out[i];
            // the slot is free once its previous tenant has been copied out
            BBFFT_CUDA_CHECK(cudaStreamWaitEvent(ring.s_in, ring.drained[i], 0));
            BBFFT_CUDA_CHECK(cudaMemcpyAsync(din, static_cast<char const *>(host_in) + ioff, ib, cudaMemcpyHostToDevice,
                                             ring.s_in));
            BBFFT_CUDA_CHECK(cudaEventRecord(ring.loaded[i], ring.s_in));
            BBFFT_CUDA_CHECK(cudaStreamWaitEvent(ring.s_k, ring.loaded[i], 0));
            // the slab sits at the start of its slot: hand the plan base pointers such that slice k0 lands there
            impl.enqueue_slab(static_cast<char const *>(din) - ioff, static_cast<char *>(dout) - ooff, k0, cnt, ring.s_k);
            BBFFT_CUDA_CHECK(cudaEventRecord(ring.computed[i], ring.s_k));
            BBFFT_CUDA_CHECK(cudaStreamWaitEvent(ring.s_out, ring.computed[i], 0));
            BBFFT_CUDA_CHECK(cudaMemcpyAsync(static_cast<char *>(host_out) + ooff, dout, ob, cudaMemcpyDeviceToHost,
                                             ring.s_out));
            BBFFT_CUDA_CHECK(cudaEventRecord(ring.drained[i], ring.s_out));
            last = ring.drained[i];
        }
        // copies leave in issue order on the copy-out stream: the last slab's event covers this call
        if (last) BBFFT_CUDA_CHECK(cudaStreamSynchronize(ring.s_out));
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
            desc->grid = steps[0].tile.K * std::uint64_t(tp.p.cluster);
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
        auto steps = cuda::nd_decompose(c, cuda::device_props{}, true);
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
