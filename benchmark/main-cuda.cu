// benchmark/main-cuda.cu -- native benchmark driver of the CUDA backend with cuFFT timed beside it.
//
// Same command line, same workload and same CSV columns as the reference's benchmark
// executables (benchmark/main-cuda.cpp + args.cpp:11-125, test-double-batched-fft.cpp:27-127,
// test-cufft.cpp:24-117), extended with --impl and multi-device slab sharding:
//
//   bbfft-bench [-i/o] [-v] [-m <M>]... [-k <K>] [-b <B>] [--impl bbfft|cufft|both]
//               [--devices <G>] [--burst <n>] (s|d)(r|c) <N1> <N2> ...
//
// 1d FFT over the second mode of an M x N x K tensor; K defaults to a 1 GiB input tensor
// (benchmark/test.hpp:18-30).  Input is the analytic single-mode signal of benchmark/signal.hpp
// and the result is checked like benchmark/check.hpp:24-53 before timing.  Time is the reference's
// convention: minimum wall-clock of 10 synchronised executes (benchmark/common.hpp:10-21);
// bw = algorithmic bytes / time, flops = 5 N log2 N (2.5 for real) per transform
// (benchmark/adapter.hpp:21-25,51-52).  The cuFFT arm is the reference's: one strided
// cufftPlanMany executed once per m (cufft_descriptor.hpp:40-66, test-cufft.cpp:37-46).
// With --devices G the K batch is split into G contiguous slabs, one plan and stream per device,
// no communication; the reported time is the wall-clock of all slabs.
#include "bbfft/configuration.hpp"
#include "bbfft/cuda/make_plan.hpp"

#include <cuda_runtime.h>
#include <cufft.h>
#include <cufftXt.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#define CUDA_OK(x)                                                                                 \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) {                                                                   \
            std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__,     \
                         __LINE__);                                                                \
            std::exit(2);                                                                          \
        }                                                                                          \
    } while (0)
#define CUFFT_OK(x)                                                                                \
    do {                                                                                           \
        cufftResult r_ = (x);                                                                      \
        if (r_ != CUFFT_SUCCESS) {                                                                 \
            std::fprintf(stderr, "cuFFT error %d at %s:%d\n", int(r_), __FILE__, __LINE__);        \
            std::exit(2);                                                                          \
        }                                                                                          \
    } while (0)

struct options {
    bool inplace = true, inverse = false;
    char p = 's', d = 'r';
    std::vector<unsigned> MM, NN;
    unsigned long K = 0;
    std::size_t bytes = std::size_t(1) << 30;
    int burst = 0; // > 0: also time `burst` back-to-back executes between two CUDA events (launch-bound shapes)
    std::string impl = "both";
    int devices = 1;
};

static std::size_t parse_bytes(std::string b) {
    std::size_t s = 1;
    switch (b.back()) {
    case 'k': case 'K': s = std::size_t(1) << 10; break;
    case 'm': case 'M': s = std::size_t(1) << 20; break;
    case 'g': case 'G': s = std::size_t(1) << 30; break;
    default: break;
    }
    if (s != 1) b.pop_back();
    return std::stoul(b) * s;
}

static options parse(int argc, char **argv) {
    options o;
    int positional = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto need = [&]() -> std::string {
            if (i + 1 >= argc) throw std::invalid_argument("missing value for " + a);
            return argv[++i];
        };
        if (a == "-i") o.inplace = true;
        else if (a == "-o") o.inplace = false;
        else if (a == "-v") o.inverse = true;
        else if (a == "-m") o.MM.push_back(unsigned(std::stoul(need())));
        else if (a == "-k") o.K = std::stoul(need());
        else if (a == "-b") o.bytes = parse_bytes(need());
        else if (a == "--impl") o.impl = need();
        else if (a == "--devices") o.devices = std::stoi(need());
        else if (a == "--burst") o.burst = std::stoi(need());
        else if (a[0] == '-') throw std::invalid_argument("unknown option " + a);
        else if (positional++ == 0) {
            o.p = a[0];
            if (a.size() > 1) o.d = a[1];
        } else o.NN.push_back(unsigned(std::stoul(a)));
    }
    if (o.MM.empty()) o.MM.push_back(1);
    if (o.NN.empty()) throw std::invalid_argument("no FFT size given");
    return o;
}

// ---- analytic signal (benchmark/signal.hpp:23-62) ------------------------------------------
template <typename T> __device__ T periodic_delta(long z, long N) { return z % N == 0 ? T(1) : T(0); }

// element (m, n, k) of a tensor with strides (1, s1, s2); `mode` = (m + k_global) % N
template <typename T, bool Real, bool Spectrum>
__global__ void init_kernel(void *data, unsigned M, unsigned N, unsigned long K, unsigned long Ktotal,
                            unsigned long k_offset, unsigned long s1, unsigned long s2, unsigned NN) {
    const unsigned long idx = blockIdx.x * (unsigned long)blockDim.x + threadIdx.x;
    const unsigned long total = (unsigned long)M * NN * K;
    if (idx >= total) return;
    const unsigned m = idx % M;
    const unsigned n = (idx / M) % NN;
    const unsigned long k = idx / ((unsigned long)M * NN);
    const unsigned long kg = k + k_offset;
    const T scale = T(1) + T(kg) / T(Ktotal);
    const unsigned mode = (m + kg) % N;
    const unsigned long off = m + n * s1 + k * s2;
    constexpr T tau = T(6.28318530717958647693);
    if constexpr (Spectrum) {
        // the spectrum side is always complex
        T delta = periodic_delta<T>(long(n) - long(mode), N);
        if constexpr (Real) delta = (delta + periodic_delta<T>(long(n) + long(mode), N)) / T(2);
        reinterpret_cast<T *>(data)[2 * off] = scale * delta;
        reinterpret_cast<T *>(data)[2 * off + 1] = T(0);
    } else {
        const T arg = (tau / N) * mode * n;
        if constexpr (Real) {
            reinterpret_cast<T *>(data)[off] = scale * cos(arg) / T(N);
        } else {
            reinterpret_cast<T *>(data)[2 * off] = scale * cos(arg) / T(N);
            reinterpret_cast<T *>(data)[2 * off + 1] = scale * sin(arg) / T(N);
        }
    }
}

struct result {
    std::string impl, precision, domain;
    unsigned N, inplace, M;
    unsigned long K;
    bool inverse, ok;
    double time, bw, flops;
};

static void print_header() { std::printf("impl,p,N,inplace,M,K,domain,inverse,time,bw,flops,check\n"); }
static void print(result const &r) {
    std::printf("%s,%s,%u,%u,%u,%lu,%s,%d,%.6e,%.4f,%.4f,%s\n", r.impl.c_str(), r.precision.c_str(), r.N, r.inplace,
                r.M, r.K, r.domain.c_str(), int(r.inverse), r.time, r.bw, r.flops, r.ok ? "ok" : "FAILED");
    std::fflush(stdout);
}

template <typename F> static double bench(F f, int nrepeat = 10) {
    double best = std::numeric_limits<double>::max();
    for (int i = 0; i < nrepeat; ++i) {
        auto t0 = std::chrono::high_resolution_clock::now();
        f();
        auto t1 = std::chrono::high_resolution_clock::now();
        best = std::min(best, std::chrono::duration<double, std::nano>(t1 - t0).count());
    }
    return best;
}

// one device's share of the problem: k in [k0, k0 + K)
template <typename T> struct slab {
    int device = 0;
    unsigned long k0 = 0, K = 0;
    void *x = nullptr, *X = nullptr; // signal side, spectrum side (X == x in place)
    cudaStream_t stream = nullptr;
    bbfft::plan<bbfft::cuda::event> plan;
    cufftHandle cplan = 0;
};

template <typename T, bool Real> static void run_case(options const &o, unsigned M, unsigned N, std::vector<result> &out) {
    using namespace bbfft;
    const unsigned Nout = Real ? N / 2 + 1 : N;
    const std::size_t real_b = sizeof(T);
    const std::size_t sig_elem = Real ? real_b : 2 * real_b;
    unsigned long K = o.K ? o.K : std::max<unsigned long>(1, o.bytes / (std::size_t(M) * N * sig_elem));
    // strides (1, s1, s2) of the signal (x) and spectrum (X) side
    const unsigned long xs1 = M, Xs1 = M;
    const unsigned long xs2 = Real ? (o.inplace ? 2ul * Nout * M : (unsigned long)N * M) : (unsigned long)N * M;
    const unsigned long Xs2 = (unsigned long)Nout * M;
    const int G = std::max(1, o.devices);
    std::vector<slab<T>> slabs(G);
    const std::size_t spec_elem = 2 * real_b;
    for (int g = 0; g < G; ++g) {
        auto &s = slabs[g];
        s.device = g;
        unsigned long base = K / G, rem = K % G;
        s.K = base + (g < int(rem) ? 1 : 0);
        s.k0 = g * base + std::min<unsigned long>(g, rem);
        if (Real && N % 2 == 1) { // odd-N real transforms pair slices: keep slab boundaries even
            s.k0 = (g * (K / G)) / 2 * 2;
            unsigned long k1 = g + 1 == G ? K : ((g + 1) * (K / G)) / 2 * 2;
            s.K = k1 - s.k0;
        }
        CUDA_OK(cudaSetDevice(g));
        CUDA_OK(cudaStreamCreate(&s.stream));
        CUDA_OK(cudaMalloc(&s.x, std::max<std::size_t>(1, s.K * xs2 * sig_elem)));
        if (o.inplace) s.X = s.x;
        else CUDA_OK(cudaMalloc(&s.X, std::max<std::size_t>(1, s.K * Xs2 * spec_elem)));
    }
    auto initialise = [&]() {
        for (auto &s : slabs) {
            if (s.K == 0) continue;
            CUDA_OK(cudaSetDevice(s.device));
            const unsigned NN = o.inverse ? Nout : N;
            const unsigned long total = (unsigned long)M * NN * s.K;
            const unsigned threads = 256;
            const unsigned long blocks = (total + threads - 1) / threads;
            if (o.inverse) {
                init_kernel<T, Real, true><<<blocks, threads, 0, s.stream>>>(s.X, M, N, s.K, K, s.k0, Xs1, Xs2, NN);
            } else {
                init_kernel<T, Real, false><<<blocks, threads, 0, s.stream>>>(s.x, M, N, s.K, K, s.k0, xs1, xs2, NN);
            }
            CUDA_OK(cudaGetLastError());
        }
        for (auto &s : slabs) {
            CUDA_OK(cudaSetDevice(s.device));
            CUDA_OK(cudaStreamSynchronize(s.stream));
        }
    };
    // result check on a sample of slices (benchmark/check.hpp:24-53, tolerance sqrt(N) * 1e3 * eps)
    auto check = [&]() -> bool {
        const double tol = std::sqrt(double(N)) * 1.0e3 * std::numeric_limits<T>::epsilon();
        for (auto &s : slabs) {
            if (s.K == 0) continue;
            CUDA_OK(cudaSetDevice(s.device));
            const unsigned long ks[3] = {0, s.K / 2, s.K - 1};
            for (unsigned long k : ks) {
                const unsigned long kg = k + s.k0;
                const double scale = 1.0 + double(kg) / double(K);
                if (o.inverse) {
                    const bool real_out = Real;
                    std::vector<T> h(std::size_t(real_out ? 1 : 2) * xs2);
                    CUDA_OK(cudaMemcpy(h.data(), static_cast<char *>(s.x) + k * xs2 * sig_elem, h.size() * sizeof(T),
                                       cudaMemcpyDeviceToHost));
                    for (unsigned n = 0; n < N; ++n) {
                        for (unsigned m = 0; m < M; ++m) {
                            const unsigned mode = (m + kg) % N;
                            const double arg = (6.28318530717958647693 / N) * mode * n;
                            const std::size_t off = m + n * xs1;
                            double err;
                            if (real_out) {
                                err = std::abs(double(h[off]) - scale * std::cos(arg));
                            } else {
                                err = std::abs(std::complex<double>(h[2 * off], h[2 * off + 1]) -
                                               scale * std::complex<double>(std::cos(arg), std::sin(arg)));
                            }
                            if (!(err <= tol)) {
                                std::fprintf(stderr, "FFT error (%u, %u, %lu): |diff| = %g\n", m, n, kg, err);
                                return false;
                            }
                        }
                    }
                } else {
                    std::vector<T> h(2 * Xs2);
                    CUDA_OK(cudaMemcpy(h.data(), static_cast<char *>(s.X) + k * Xs2 * spec_elem, h.size() * sizeof(T),
                                       cudaMemcpyDeviceToHost));
                    for (unsigned n = 0; n < Nout; ++n) {
                        for (unsigned m = 0; m < M; ++m) {
                            const long mode = long((m + kg) % N);
                            double delta = (long(n) - mode) % long(N) == 0 ? 1.0 : 0.0;
                            if (Real) delta = (delta + ((long(n) + mode) % long(N) == 0 ? 1.0 : 0.0)) / 2.0;
                            const std::size_t off = m + n * Xs1;
                            const double err = std::abs(std::complex<double>(h[2 * off], h[2 * off + 1]) -
                                                        std::complex<double>(scale * delta, 0.0));
                            if (!(err <= tol)) {
                                std::fprintf(stderr, "FFT error (%u, %u, %lu): |diff| = %g\n", m, n, kg, err);
                                return false;
                            }
                        }
                    }
                }
            }
        }
        return true;
    };
    const double bytes = Real ? double(N * real_b + Nout * spec_elem) * M * K : 2.0 * N * spec_elem * M * K;
    const double flops = (Real ? 2.5 : 5.0) * N * std::log2(double(N)) * M * K;
    auto sync_all = [&]() {
        for (auto &s : slabs) {
            CUDA_OK(cudaSetDevice(s.device));
            CUDA_OK(cudaStreamSynchronize(s.stream));
        }
    };
    auto emit = [&](char const *impl, double ns, bool ok) {
        out.push_back(result{impl, sizeof(T) == 4 ? "f" : "d", Real ? "real" : "complex", N, unsigned(o.inplace), M, K,
                             o.inverse, ok, ns * 1e-9, bytes / ns, flops / ns});
        print(out.back());
    };

    if (o.impl == "bbfft" || o.impl == "both") {
        for (auto &s : slabs) {
            if (s.K == 0) continue;
            CUDA_OK(cudaSetDevice(s.device));
            configuration cfg = {};
            cfg.dim = 1;
            cfg.shape = {M, N, s.K, 0, 0};
            cfg.fp = sizeof(T) == 4 ? precision::f32 : precision::f64;
            cfg.dir = o.inverse ? direction::backward : direction::forward;
            cfg.type = Real ? (o.inverse ? transform_type::c2r : transform_type::r2c) : transform_type::c2c;
            cfg.istride = {1, xs1, xs2, 0, 0};
            cfg.ostride = {1, Xs1, Xs2, 0, 0};
            if (o.inverse) std::swap(cfg.istride, cfg.ostride);
            s.plan = make_plan(cfg, s.stream, s.device);
        }
        auto exec = [&]() {
            for (auto &s : slabs) {
                if (s.K == 0) continue;
                CUDA_OK(cudaSetDevice(s.device));
                if (o.inverse) s.plan.execute(s.X, s.x);
                else s.plan.execute(s.x, s.X);
            }
            sync_all();
        };
        initialise();
        exec();
        const bool ok = check();
        const double ns = bench(exec);
        emit("bbfft-cuda", ns, ok);
        if (o.burst > 0 && slabs.size() == 1) {
            // back-to-back executes on the stream, no host synchronisation in between: what a caller with
            // many small, L2-resident batches sees per transform batch (BASELINE config 1)
            auto &s0 = slabs[0];
            cudaEvent_t e0, e1;
            CUDA_OK(cudaEventCreate(&e0));
            CUDA_OK(cudaEventCreate(&e1));
            double best = 1e30;
            for (int rep = 0; rep < 10; ++rep) {
                CUDA_OK(cudaEventRecord(e0, s0.stream));
                for (int i = 0; i < o.burst; ++i) {
                    if (o.inverse) s0.plan.execute(s0.X, s0.x);
                    else s0.plan.execute(s0.x, s0.X);
                }
                CUDA_OK(cudaEventRecord(e1, s0.stream));
                CUDA_OK(cudaEventSynchronize(e1));
                float ms = 0;
                CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
                best = std::min(best, double(ms) * 1e6 / o.burst);
            }
            emit("bbfft-cuda-burst", best, ok);
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
        for (auto &s : slabs) s.plan = {};
    }
    if (o.impl == "cufft" || o.impl == "both") {
        for (auto &s : slabs) {
            if (s.K == 0) continue;
            CUDA_OK(cudaSetDevice(s.device));
            int n = int(N);
            int inembed = Real ? int(o.inplace ? 2 * Nout : N) : n, onembed = Real ? int(Nout) : n;
            int idist = inembed * int(M), odist = onembed * int(M);
            if (o.inverse) {
                std::swap(inembed, onembed);
                std::swap(idist, odist);
            }
            cufftType type;
            if (Real) type = sizeof(T) == 4 ? (o.inverse ? CUFFT_C2R : CUFFT_R2C) : (o.inverse ? CUFFT_Z2D : CUFFT_D2Z);
            else type = sizeof(T) == 4 ? CUFFT_C2C : CUFFT_Z2Z;
            CUFFT_OK(cufftPlanMany(&s.cplan, 1, &n, &inembed, int(M), idist, &onembed, int(M), odist, type, int(s.K)));
            CUFFT_OK(cufftSetStream(s.cplan, s.stream));
        }
        // cuFFT rejects a real tensor whose m-th column starts off an 8 (16) byte boundary
        // (CUFFT_INVALID_VALUE for odd m): strided real transforms with M > 1 are then reported
        // as unsupported instead of aborting the sweep
        bool cufft_supported = true;
        auto exec = [&]() {
            for (auto &s : slabs) {
                if (s.K == 0) continue;
                CUDA_OK(cudaSetDevice(s.device));
                for (unsigned i = 0; i < M && cufft_supported; ++i) {
                    char *xi = static_cast<char *>(s.x) + i * sig_elem;
                    char *Xi = static_cast<char *>(s.X) + i * spec_elem;
                    cufftResult r = o.inverse ? cufftXtExec(s.cplan, Xi, xi, CUFFT_INVERSE)
                                              : cufftXtExec(s.cplan, xi, Xi, CUFFT_FORWARD);
                    if (r == CUFFT_INVALID_VALUE && Real && M > 1) {
                        cufft_supported = false;
                    } else if (r != CUFFT_SUCCESS) {
                        CUFFT_OK(r);
                    }
                }
            }
            sync_all();
        };
        initialise();
        exec();
        if (cufft_supported) {
            const bool ok = check();
            const double ns = bench(exec);
            emit("cufft", ns, ok);
        } else {
            std::fprintf(stderr, "cufft: strided real transform with M = %u not supported (misaligned columns), skipped\n", M);
        }
        for (auto &s : slabs) {
            if (s.cplan) cufftDestroy(s.cplan);
            s.cplan = 0;
        }
    }
    for (auto &s : slabs) {
        CUDA_OK(cudaSetDevice(s.device));
        if (!o.inplace) cudaFree(s.X);
        cudaFree(s.x);
        cudaStreamDestroy(s.stream);
    }
}

int main(int argc, char **argv) {
    options o;
    try {
        o = parse(argc, argv);
    } catch (std::exception const &ex) {
        std::fprintf(stderr,
                     "Error: could not parse command line: %s\n"
                     "Usage: %s [-i/o] [-v] [-m <M>] [-k <K>] [-b <B>] [--impl bbfft|cufft|both] [--devices G] "
                     "(s|d)(r|c) <N1> <N2> ...\n",
                     ex.what(), argv[0]);
        return -1;
    }
    int ndev = 0;
    CUDA_OK(cudaGetDeviceCount(&ndev));
    if (o.devices > ndev) {
        std::fprintf(stderr, "only %d device(s) visible\n", ndev);
        return -1;
    }
    std::vector<result> results;
    print_header();
    try {
        for (unsigned M : o.MM) {
            for (unsigned N : o.NN) {
                if (o.p == 's') {
                    if (o.d == 'r') run_case<float, true>(o, M, N, results);
                    else run_case<float, false>(o, M, N, results);
                } else {
                    if (o.d == 'r') run_case<double, true>(o, M, N, results);
                    else run_case<double, false>(o, M, N, results);
                }
            }
        }
    } catch (std::exception const &ex) {
        std::fprintf(stderr, "Error: %s\n", ex.what());
        return 1;
    }
    for (auto const &r : results) {
        if (!r.ok) return 3;
    }
    return 0;
}
