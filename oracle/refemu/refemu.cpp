// oracle/refemu/refemu.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// "The reference itself, run here": this file instantiates the reference's OWN plan templates
// (reference: src/common/algorithm.hpp:20-30 select_fft_algorithm, small_batch_fft.hpp,
// factor2_slm_fft.hpp, nd_fft.hpp) over an `Api` policy class (concept spelled out by
// reference src/base/dummy_api.hpp:18-61) whose "device" is the host CPU:
//
//   build_module(source)  : the OpenCL-C text the reference generator produced is rewritten
//                           textually into C++ (6 regex rules, see rewrite_opencl_c), compiled
//                           with g++ -ffp-contract=off against ocl_shim.hpp and dlopen'ed;
//   launch_kernel(...)    : every work-group is run as local_size fibers (ucontext); a
//                           barrier() yields to the group scheduler.  Work-groups are spread
//                           over host threads.
//
// Everything the reference decides on the host -- algorithm selection, factorization, work-group
// geometry, twiddle tables, nd chaining, tmp buffers, in-place guards -- is therefore the
// reference's unmodified code; only the device is emulated.  Exposed through a small C ABI so
// that tests/ and bench.py (cpu_baseline / --impl reference) can call it via ctypes.
#include "algorithm.hpp" // reference: src/common/algorithm.hpp

#include "bbfft/bad_configuration.hpp"
#include "bbfft/configuration.hpp"
#include "bbfft/device_info.hpp"
#include "bbfft/jit_cache_all.hpp"
#include "bbfft/parser.hpp"
#include "bbfft/shared_handle.hpp"

#include "ocl_shim.hpp"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <fstream>
#include <memory>
#include <mutex>
#include <regex>
#include <sstream>
#include <string>
#include <thread>
#include <ucontext.h>
#include <unistd.h>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_last_error;
std::string g_shim_dir;  // directory holding ocl_shim.hpp at run time
std::string g_work_dir;  // where rewritten kernels are compiled
int g_num_threads = 0;
std::vector<std::string> g_last_kernel_names;

// ---------------------------------------------------------------------------------------------
// OpenCL-C -> C++ textual rewrite
// ---------------------------------------------------------------------------------------------
struct kernel_sig {
    std::string name;
    std::array<std::size_t, 3> reqd_wgs;
    std::vector<std::string> param_types;
};

std::string rewrite_opencl_c(std::string src, std::vector<kernel_sig> &kernels) {
    // 1. kernel headers: remember name, work-group size and parameter list
    std::regex head(R"(kernel\s+__attribute__\(\(reqd_work_group_size\((\d+),(\d+),(\d+)\)\)\)\s+__attribute__\(\(intel_reqd_sub_group_size\((\d+)\)\)\)\s+void\s+(\w+)\(([^)]*)\))");
    std::string out;
    auto begin = std::sregex_iterator(src.begin(), src.end(), head);
    std::size_t last = 0;
    for (auto it = begin; it != std::sregex_iterator(); ++it) {
        auto const &m = *it;
        kernel_sig sig;
        sig.reqd_wgs = {std::stoul(m[1]), std::stoul(m[2]), std::stoul(m[3])};
        sig.name = m[5];
        std::string params = m[6];
        std::stringstream ps(params);
        std::string p;
        while (std::getline(ps, p, ',')) {
            // "global float2* in" -> type "float2*"
            p = std::regex_replace(p, std::regex(R"(\b(global|local|constant)\s+)"), "");
            auto pos = p.find_last_of("* ");
            std::string type = p.substr(0, pos + 1);
            type = std::regex_replace(type, std::regex(R"(^\s+|\s+$)"), "");
            sig.param_types.push_back(type);
        }
        kernels.push_back(sig);
        out += src.substr(last, m.position() - last);
        out += "static void " + sig.name + "(" + params + ")";
        last = m.position() + m.length();
    }
    out += src.substr(last);
    // 2. drop unroll hints
    out = std::regex_replace(out, std::regex(R"(__attribute__\(\(opencl_unroll_hint\(\d+\)\)\))"), "");
    // 3. local arrays -> work-group arena
    out = std::regex_replace(out, std::regex(R"(\blocal\s+(\w+)\s+(\w+)\[([^\]]+)\];)"),
                             "$1* $2 = ($1*) wg_local(sizeof($1) * ($3));");
    // 4. vector literals
    out = std::regex_replace(out, std::regex(R"(\(float2\)\s*\()"), "mk_float2(");
    out = std::regex_replace(out, std::regex(R"(\(double2\)\s*\()"), "mk_double2(");
    // 5. address-space qualifiers
    out = std::regex_replace(out, std::regex(R"(\b(global|local|constant)\s+)"), "");
    // 6. trampolines
    std::ostringstream tr;
    for (auto const &k : kernels) {
        tr << "extern \"C\" void " << k.name
           << "__entry(refemu_wi_ctx* ctx, const unsigned long* a) {\n    refemu_ctx = ctx;\n    "
           << k.name << "(";
        for (std::size_t i = 0; i < k.param_types.size(); ++i) {
            tr << (i ? ", " : "") << "(" << k.param_types[i] << ") a[" << i << "]";
        }
        tr << ");\n}\n";
    }
    return "#include \"ocl_shim.hpp\"\n" + out + "\n" + tr.str();
}

// ---------------------------------------------------------------------------------------------
// Compiled module
// ---------------------------------------------------------------------------------------------
using entry_fn = void (*)(refemu_wi_ctx *, const unsigned long *);

struct emu_module {
    void *dl = nullptr;
    std::vector<kernel_sig> kernels;
    ~emu_module() {
        if (dl) {
            dlclose(dl);
        }
    }
};

struct emu_kernel {
    entry_fn fn = nullptr;
    kernel_sig sig;
};

std::atomic<int> g_module_counter{0};

std::string tmp_dir() {
    if (!g_work_dir.empty()) {
        return g_work_dir;
    }
    char const *t = getenv("TMPDIR");
    std::string dir = std::string(t ? t : "/tmp") + "/bbfft_refemu_" + std::to_string(getpid());
    std::string cmd = "mkdir -p " + dir;
    if (system(cmd.c_str()) != 0) {
        throw std::runtime_error("refemu: cannot create " + dir);
    }
    g_work_dir = dir;
    return dir;
}

emu_module *compile_module(std::string const &cl_source) {
    auto mod = std::make_unique<emu_module>();
    std::string cpp = rewrite_opencl_c(cl_source, mod->kernels);
    if (mod->kernels.empty()) {
        throw std::runtime_error("refemu: no kernel found in generated source");
    }
    int id = g_module_counter++;
    std::string base = tmp_dir() + "/k" + std::to_string(id);
    {
        std::ofstream f(base + ".cl");
        f << cl_source;
    }
    {
        std::ofstream f(base + ".cpp");
        f << cpp;
    }
    char const *cxx = getenv("REFEMU_CXX");
    std::string cmd = std::string(cxx ? cxx : "g++") +
                      " -std=c++17 -O1 -fPIC -shared -w -ffp-contract=off -fno-fast-math -I" +
                      g_shim_dir + " -o " + base + ".so " + base + ".cpp 2> " + base + ".log";
    if (system(cmd.c_str()) != 0) {
        std::ifstream l(base + ".log");
        std::stringstream ss;
        ss << l.rdbuf();
        throw std::runtime_error("refemu: host compilation of generated kernel failed:\n" +
                                 ss.str().substr(0, 4000));
    }
    mod->dl = dlopen((base + ".so").c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!mod->dl) {
        throw std::runtime_error(std::string("refemu: dlopen failed: ") + dlerror());
    }
    return mod.release();
}

// ---------------------------------------------------------------------------------------------
// Work-group scheduler (fibers)
// ---------------------------------------------------------------------------------------------
constexpr std::size_t fiber_stack_bytes = 256 * 1024;
constexpr std::size_t local_arena_bytes = 512 * 1024;

struct fiber {
    ucontext_t uc;
    refemu_wi_ctx ctx;
    bool done = false;
    struct group_runner *owner = nullptr;
};

struct group_runner {
    ucontext_t sched;
    std::vector<fiber> fibers;
    std::vector<char> stacks;
    std::vector<char> local_mem;
    entry_fn fn = nullptr;
    const unsigned long *args = nullptr;
    fiber *current = nullptr;

    static void yield_cb(refemu_wi_ctx *c) {
        // ctx is the second member of fiber; recover the fiber
        fiber *f = reinterpret_cast<fiber *>(reinterpret_cast<char *>(c) - offsetof(fiber, ctx));
        swapcontext(&f->uc, &f->owner->sched);
    }
    static void fiber_main(unsigned lo, unsigned hi) {
        fiber *f = reinterpret_cast<fiber *>((std::uintptr_t(hi) << 32) | std::uintptr_t(lo));
        f->owner->fn(&f->ctx, f->owner->args);
        f->done = true;
        // returns to uc_link = sched
    }

    void run_group(std::array<std::size_t, 3> const &gid, std::array<std::size_t, 3> const &lws,
                   std::array<std::size_t, 3> const &ngroups) {
        std::size_t n = lws[0] * lws[1] * lws[2];
        if (fibers.size() != n) {
            fibers.assign(n, fiber{});
            stacks.resize(n * fiber_stack_bytes);
        }
        if (local_mem.empty()) {
            local_mem.resize(local_arena_bytes);
        }
        std::size_t i = 0;
        for (std::size_t l2 = 0; l2 < lws[2]; ++l2) {
            for (std::size_t l1 = 0; l1 < lws[1]; ++l1) {
                for (std::size_t l0 = 0; l0 < lws[0]; ++l0, ++i) {
                    fiber &f = fibers[i];
                    f.done = false;
                    f.owner = this;
                    f.ctx.local_id[0] = l0;
                    f.ctx.local_id[1] = l1;
                    f.ctx.local_id[2] = l2;
                    for (int d = 0; d < 3; ++d) {
                        f.ctx.group_id[d] = gid[d];
                        f.ctx.local_size[d] = lws[d];
                        f.ctx.num_groups[d] = ngroups[d];
                    }
                    f.ctx.local_base = local_mem.data();
                    f.ctx.local_used = 0;
                    f.ctx.local_cap = local_mem.size();
                    f.ctx.yield = &yield_cb;
                    getcontext(&f.uc);
                    f.uc.uc_stack.ss_sp = stacks.data() + i * fiber_stack_bytes;
                    f.uc.uc_stack.ss_size = fiber_stack_bytes;
                    f.uc.uc_link = &sched;
                    std::uintptr_t p = reinterpret_cast<std::uintptr_t>(&f);
                    makecontext(&f.uc, (void (*)())fiber_main, 2, unsigned(p & 0xffffffffu),
                                unsigned(p >> 32));
                }
            }
        }
        std::size_t remaining = n;
        while (remaining > 0) {
            std::size_t finished_this_round = 0;
            for (auto &f : fibers) {
                if (!f.done) {
                    swapcontext(&sched, &f.uc);
                    if (f.done) {
                        ++finished_this_round;
                    }
                }
            }
            remaining -= finished_this_round;
        }
    }
};

void launch(emu_kernel const &k, std::array<std::size_t, 3> gws, std::array<std::size_t, 3> lws,
            const unsigned long *args) {
    for (int d = 0; d < 3; ++d) {
        if (lws[d] != k.sig.reqd_wgs[d]) {
            throw std::runtime_error("refemu: local size does not match reqd_work_group_size");
        }
        if (gws[d] % lws[d] != 0) {
            throw std::runtime_error("refemu: global size not divisible by local size");
        }
    }
    std::array<std::size_t, 3> ng = {gws[0] / lws[0], gws[1] / lws[1], gws[2] / lws[2]};
    std::size_t total = ng[0] * ng[1] * ng[2];
    int nt = g_num_threads > 0 ? g_num_threads : int(std::thread::hardware_concurrency());
    nt = std::max(1, std::min<int>(nt, int(total)));
    std::atomic<std::size_t> next{0};
    auto worker = [&]() {
        group_runner r;
        r.fn = k.fn;
        r.args = args;
        for (;;) {
            std::size_t g = next.fetch_add(1);
            if (g >= total) {
                break;
            }
            std::array<std::size_t, 3> gid = {g % ng[0], (g / ng[0]) % ng[1], g / (ng[0] * ng[1])};
            r.run_group(gid, lws, ng);
        }
    };
    if (nt == 1) {
        worker();
    } else {
        std::vector<std::thread> ts;
        for (int t = 0; t < nt; ++t) {
            ts.emplace_back(worker);
        }
        for (auto &t : ts) {
            t.join();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The Api policy (same member list as reference src/base/dummy_api.hpp:18-61)
// ---------------------------------------------------------------------------------------------
class emu_api {
  public:
    using event_type = int;
    using plan_type = bbfft::detail::plan_impl<event_type>;
    using buffer_type = void *;
    using kernel_bundle_type = emu_module *;
    using kernel_type = emu_kernel *;

    explicit emu_api(bbfft::device_info info) : info_(std::move(info)) {}

    bbfft::device_info info() { return info_; }
    uint64_t device_id() { return 0; }

    auto build_module(std::string const &source) -> bbfft::shared_handle<bbfft::module_handle_t> {
        emu_module *m = compile_module(source);
        return bbfft::shared_handle<bbfft::module_handle_t>(
            reinterpret_cast<bbfft::module_handle_t>(m), [](bbfft::module_handle_t h) {
                delete reinterpret_cast<emu_module *>(h);
            });
    }
    auto make_kernel_bundle(bbfft::module_handle_t mod) -> kernel_bundle_type {
        return reinterpret_cast<emu_module *>(mod);
    }
    auto create_kernel(kernel_bundle_type b, std::string const &name) -> kernel_type {
        auto k = new emu_kernel;
        for (auto const &s : b->kernels) {
            if (s.name == name) {
                k->sig = s;
            }
        }
        k->fn = reinterpret_cast<entry_fn>(dlsym(b->dl, (name + "__entry").c_str()));
        if (!k->fn) {
            delete k;
            throw std::runtime_error("refemu: kernel not found: " + name);
        }
        g_last_kernel_names.push_back(name);
        return k;
    }

    struct arg_handler {
        unsigned long a[8] = {};
        template <typename T> void set_arg(unsigned i, T const &v) {
            static_assert(sizeof(T) <= sizeof(unsigned long));
            unsigned long x = 0;
            std::memcpy(&x, &v, sizeof(T));
            a[i] = x;
        }
    };

    template <typename T>
    event_type launch_kernel(kernel_type &k, std::array<std::size_t, 3> gws,
                             std::array<std::size_t, 3> lws, std::vector<event_type> const &,
                             T set_args) {
        arg_handler h;
        set_args(h);
        launch(*k, gws, lws, h.a);
        return 0;
    }

    buffer_type create_device_buffer(std::size_t bytes) { return std::malloc(bytes); }
    template <typename T> buffer_type create_device_buffer(std::size_t n) {
        return create_device_buffer(n * sizeof(T));
    }
    template <typename T> buffer_type create_twiddle_table(std::vector<T> &tw) {
        void *p = std::malloc(tw.size() * sizeof(T) + 16);
        std::memcpy(p, tw.data(), tw.size() * sizeof(T));
        return p;
    }
    static void release_event(event_type) {}
    static void release_buffer(buffer_type b) { std::free(b); }
    static void release_kernel(kernel_type k) { delete k; }

  private:
    bbfft::device_info info_;
};

struct refemu_plan {
    std::shared_ptr<emu_api::plan_type> impl;
    std::vector<std::string> kernel_names;
    // keep the user-module strings alive
    std::string cb_src, cb_load, cb_store;
};

template <typename F> int guarded(F &&f) {
    try {
        f();
        return 0;
    } catch (bbfft::bad_configuration const &e) {
        g_last_error = std::string("bad_configuration: ") + e.what();
        return 2;
    } catch (std::exception const &e) {
        g_last_error = e.what();
        return 1;
    }
}

} // namespace

extern "C" {

// POD mirror of bbfft::configuration (reference include/bbfft/configuration.hpp:151-192)
struct refemu_config {
    unsigned dim;
    unsigned long shape[5];
    int fp;   // 4 or 8
    int dir;  // -1 or +1
    int type; // 0 c2c, 1 r2c, 2 c2r
    unsigned long istride[5];
    unsigned long ostride[5];
    const char *cb_source; // OpenCL-C user module or NULL
    const char *cb_load;
    const char *cb_store;
};

const char *refemu_last_error() { return g_last_error.c_str(); }

void refemu_set_shim_dir(const char *dir) { g_shim_dir = dir; }
void refemu_set_threads(int n) { g_num_threads = n; }

// descriptor -> config using the reference's own parser (src/base/parser.cpp:59-211)
int refemu_parse_descriptor(const char *desc, refemu_config *c) {
    return guarded([&] {
        auto cfg = bbfft::parse_fft_descriptor(desc);
        c->dim = cfg.dim;
        for (int i = 0; i < 5; ++i) {
            c->shape[i] = cfg.shape[i];
            c->istride[i] = cfg.istride[i];
            c->ostride[i] = cfg.ostride[i];
        }
        c->fp = int(cfg.fp);
        c->dir = int(cfg.dir);
        c->type = int(cfg.type);
        c->cb_source = c->cb_load = c->cb_store = nullptr;
    });
}

// config -> descriptor using the reference's own printer (src/base/configuration.cpp:69-160)
int refemu_to_descriptor(const refemu_config *c, char *buf, unsigned long len) {
    return guarded([&] {
        bbfft::configuration cfg = {};
        cfg.dim = c->dim;
        for (int i = 0; i < 5; ++i) {
            cfg.shape[i] = c->shape[i];
            cfg.istride[i] = c->istride[i];
            cfg.ostride[i] = c->ostride[i];
        }
        cfg.fp = bbfft::precision(c->fp);
        cfg.dir = bbfft::direction(c->dir);
        cfg.type = bbfft::transform_type(c->type);
        auto s = cfg.to_string();
        std::snprintf(buf, len, "%s", s.c_str());
    });
}

// default strides (reference src/base/configuration.cpp:24-56)
void refemu_default_strides(const refemu_config *c, int inplace, unsigned long *is,
                            unsigned long *os) {
    std::array<std::size_t, bbfft::max_tensor_dim> shape;
    for (int i = 0; i < 5; ++i) {
        shape[i] = c->shape[i];
    }
    auto i_ = bbfft::default_istride(c->dim, shape, bbfft::transform_type(c->type), inplace);
    auto o_ = bbfft::default_ostride(c->dim, shape, bbfft::transform_type(c->type), inplace);
    for (int i = 0; i < 5; ++i) {
        is[i] = i_[i];
        os[i] = o_[i];
    }
}

int refemu_plan_create(const refemu_config *c, const char *device_info, void **plan) {
    return guarded([&] {
        if (g_shim_dir.empty()) {
            throw std::runtime_error("refemu: call refemu_set_shim_dir first");
        }
        auto p = std::make_unique<refemu_plan>();
        bbfft::configuration cfg = {};
        cfg.dim = c->dim;
        for (int i = 0; i < 5; ++i) {
            cfg.shape[i] = c->shape[i];
            cfg.istride[i] = c->istride[i];
            cfg.ostride[i] = c->ostride[i];
        }
        cfg.fp = bbfft::precision(c->fp);
        cfg.dir = bbfft::direction(c->dir);
        cfg.type = bbfft::transform_type(c->type);
        if (c->cb_source) {
            p->cb_src = c->cb_source;
            p->cb_load = c->cb_load ? c->cb_load : "";
            p->cb_store = c->cb_store ? c->cb_store : "";
            cfg.callbacks.data = p->cb_src.c_str();
            cfg.callbacks.length = p->cb_src.size();
            cfg.callbacks.load_function = c->cb_load ? p->cb_load.c_str() : nullptr;
            cfg.callbacks.store_function = c->cb_store ? p->cb_store.c_str() : nullptr;
        }
        // default device = PVC (reference tools/common/info.cpp:9)
        auto info = bbfft::parse_device_info(device_info && *device_info
                                                 ? device_info
                                                 : "{1024, {16, 32}, 131072, gpu}");
        g_last_kernel_names.clear();
        p->impl = bbfft::select_fft_algorithm<emu_api>(cfg, emu_api(info), nullptr);
        p->kernel_names = g_last_kernel_names;
        *plan = p.release();
    });
}

int refemu_plan_num_kernels(void *plan) {
    return int(static_cast<refemu_plan *>(plan)->kernel_names.size());
}
const char *refemu_plan_kernel_name(void *plan, int i) {
    return static_cast<refemu_plan *>(plan)->kernel_names[i].c_str();
}

int refemu_plan_execute(void *plan, const void *in, void *out) {
    return guarded([&] { static_cast<refemu_plan *>(plan)->impl->execute(in, out); });
}

void refemu_plan_destroy(void *plan) { delete static_cast<refemu_plan *>(plan); }

} // extern "C"
