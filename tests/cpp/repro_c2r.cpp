// tests/cpp/repro_c2r.cpp -- stress harness for one 1d configuration: executes a plan many times
// (fresh plan per iteration or one plan re-used; user stream blocking / non-blocking / default) and
// compares every result with a host long-double DFT, printing every differing element.
// Written to chase the round-1 nondeterministic c2r result (reference check: test/callback.cpp:18-114);
// kept as the determinism soak of tests/test_gpu_z_cpp_api.py.
//
// usage: repro_c2r <type c2c|r2c|c2r> <f32|f64> M N K iters [mode fresh|reuse] [stream blocking|nonblocking|default] [sync 0|1]
#include "bbfft/configuration.hpp"
#include "bbfft/cuda/make_plan.hpp"

#include <cuda_runtime_api.h>

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace bbfft;

#define CUDA_OK(x)                                                                                 \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) {                                                                   \
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);   \
            std::exit(2);                                                                          \
        }                                                                                          \
    } while (0)

template <typename T> int run(std::string const &type, std::size_t M, std::size_t N, std::size_t K, int iters,
                              bool fresh, cudaStream_t stream, bool sync_after_plan) {
    using cplx = std::complex<T>;
    const std::size_t ns = N / 2 + 1;
    const bool c2r = type == "c2r", r2c = type == "r2c";
    const std::size_t n_in = c2r ? ns : N, n_out = r2c ? ns : N;
    const std::size_t in_elems = M * n_in * K, out_elems = M * n_out * K;
    const std::size_t in_scalars = in_elems * (r2c ? 1 : 2), out_scalars = out_elems * (c2r ? 1 : 2);
    unsigned s = 777;
    auto rnd = [&] {
        s = s * 1664525u + 1013904223u;
        return T((s >> 8) & 0xffff) / T(65536);
    };
    std::vector<T> in(in_scalars);
    for (auto &v : in) v = rnd();
    if (c2r) {
        // Hermitian-valid spectrum: imag(X[0]) = imag(X[N/2]) = 0
        for (std::size_t k = 0; k < K; ++k)
            for (std::size_t m = 0; m < M; ++m) {
                in[2 * (m + M * (0 + ns * k)) + 1] = 0;
                if (N % 2 == 0) in[2 * (m + M * (N / 2 + ns * k)) + 1] = 0;
            }
    }
    // host reference (long double direct DFT)
    std::vector<long double> want(out_scalars);
    const long double tau = 6.283185307179586476925286766559005768L;
    const int dir = r2c ? -1 : (c2r ? 1 : -1);
    std::vector<long double> cs(N), sn(N);
    for (std::size_t i = 0; i < N; ++i) {
        cs[i] = cosl(tau * i / N);
        sn[i] = dir * sinl(tau * i / N);
    }
    for (std::size_t k = 0; k < K; ++k)
        for (std::size_t m = 0; m < M; ++m) {
            std::vector<long double> xr(N), xi(N, 0.0L);
            for (std::size_t n = 0; n < N; ++n) {
                if (r2c) {
                    xr[n] = in[m + M * (n + N * k)];
                } else if (c2r) {
                    std::size_t nn = n < ns ? n : N - n;
                    std::size_t o = 2 * (m + M * (nn + ns * k));
                    xr[n] = in[o];
                    xi[n] = n < ns ? in[o + 1] : -in[o + 1];
                    if (n == 0 || 2 * n == N) xi[n] = 0;
                } else {
                    std::size_t o = 2 * (m + M * (n + N * k));
                    xr[n] = in[o];
                    xi[n] = in[o + 1];
                }
            }
            for (std::size_t q = 0; q < n_out; ++q) {
                long double ar = 0, ai = 0;
                for (std::size_t n = 0; n < N; ++n) {
                    std::size_t idx = (q * n) % N;
                    ar += xr[n] * cs[idx] - xi[n] * sn[idx];
                    ai += xr[n] * sn[idx] + xi[n] * cs[idx];
                }
                std::size_t o = m + M * (q + n_out * k);
                if (c2r) {
                    want[o] = ar;
                } else {
                    want[2 * o] = ar;
                    want[2 * o + 1] = ai;
                }
            }
        }
    T *din = nullptr, *dout = nullptr;
    CUDA_OK(cudaMalloc(reinterpret_cast<void **>(&din), in_scalars * sizeof(T)));
    CUDA_OK(cudaMalloc(reinterpret_cast<void **>(&dout), out_scalars * sizeof(T)));
    CUDA_OK(cudaMemcpy(din, in.data(), in_scalars * sizeof(T), cudaMemcpyHostToDevice));
    configuration cfg = {1,
                         {M, N, K},
                         sizeof(T) == 4 ? precision::f32 : precision::f64,
                         dir < 0 ? direction::forward : direction::backward,
                         c2r ? transform_type::c2r : (r2c ? transform_type::r2c : transform_type::c2c),
                         {1, M, M * n_in},
                         {1, M, M * n_out}};
    const double eps = sizeof(T) == 4 ? 1e-4 : 1e-11;
    int bad_iters = 0;
    std::vector<T> got(out_scalars), first;
    cuda_plan keep;
    if (!fresh) keep = make_plan(cfg, stream);
    for (int it = 0; it < iters; ++it) {
        CUDA_OK(cudaMemset(dout, 0xff, out_scalars * sizeof(T)));
        cuda_plan p = fresh ? make_plan(cfg, stream) : keep;
        if (sync_after_plan) CUDA_OK(cudaDeviceSynchronize());
        p.execute(din, dout).wait();
        CUDA_OK(cudaMemcpy(got.data(), dout, out_scalars * sizeof(T), cudaMemcpyDeviceToHost));
        std::size_t nbad = 0;
        for (std::size_t i = 0; i < out_scalars; ++i) {
            double scale = double(N);
            if (!(std::abs(double(got[i]) - double(want[i])) <= eps * scale)) {
                if (nbad < 12 && bad_iters < 4) {
                    std::size_t e = c2r ? i : i / 2;
                    std::printf("  iter %d: scalar %zu (m=%zu n=%zu k=%zu%s) got %.9g want %.9g\n", it, i, e % M,
                                e / M % n_out, e / (M * n_out), c2r ? "" : (i % 2 ? " im" : " re"), double(got[i]),
                                double(want[i]));
                }
                ++nbad;
            }
        }
        if (nbad) {
            ++bad_iters;
            if (bad_iters == 1) {
                // compact map: which n are wrong for which m (k = 0) and how many per k
                std::printf("  wrong (m: n list) at k=0:");
                for (std::size_t m = 0; m < M; ++m) {
                    bool any = false;
                    for (std::size_t q = 0; q < n_out; ++q) {
                        std::size_t e = m + M * q;
                        bool w = c2r ? !(std::abs(double(got[e]) - double(want[e])) <= eps * double(N))
                                     : !(std::abs(double(got[2 * e]) - double(want[2 * e])) <= eps * double(N) &&
                                         std::abs(double(got[2 * e + 1]) - double(want[2 * e + 1])) <= eps * double(N));
                        if (w) {
                            if (!any) std::printf(" [m=%zu:", m);
                            any = true;
                            std::printf(" %zu", q);
                        }
                    }
                    if (any) std::printf("]");
                }
                std::printf("\n");
            }
            if (bad_iters <= 4) {
                // does the very same plan object give the right answer when it runs again?
                p.execute(din, dout).wait();
                std::vector<T> again(out_scalars);
                CUDA_OK(cudaMemcpy(again.data(), dout, out_scalars * sizeof(T), cudaMemcpyDeviceToHost));
                std::size_t nbad2 = 0;
                for (std::size_t i = 0; i < out_scalars; ++i)
                    nbad2 += !(std::abs(double(again[i]) - double(want[i])) <= eps * double(N));
                std::printf("  iter %d: %zu wrong scalars; same plan executed again: %zu wrong\n", it, nbad, nbad2);
            }
        }
        if (it == 0) first = got;
        else if (!nbad && std::memcmp(first.data(), got.data(), out_scalars * sizeof(T)) != 0 && bad_iters == 0) {
            std::printf("  iter %d: within tolerance but not bit-identical to iteration 0\n", it);
            ++bad_iters;
        }
    }
    cudaFree(din);
    cudaFree(dout);
    return bad_iters;
}

int main(int argc, char **argv) {
    if (argc < 7) {
        std::printf("usage: %s <c2c|r2c|c2r> <f32|f64> M N K iters [fresh|reuse] [blocking|nonblocking|default] [sync 0|1]\n", argv[0]);
        return 2;
    }
    std::string type = argv[1], fp = argv[2];
    std::size_t M = std::strtoull(argv[3], nullptr, 10), N = std::strtoull(argv[4], nullptr, 10),
                K = std::strtoull(argv[5], nullptr, 10);
    int iters = std::atoi(argv[6]);
    bool fresh = argc < 8 || std::string(argv[7]) == "fresh";
    std::string sk = argc < 9 ? "blocking" : argv[8];
    bool sync = argc >= 10 && std::atoi(argv[9]) != 0;
    cudaStream_t stream = nullptr;
    if (sk == "blocking") CUDA_OK(cudaStreamCreate(&stream));
    if (sk == "nonblocking") CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int bad = fp == "f32" ? run<float>(type, M, N, K, iters, fresh, stream, sync)
                          : run<double>(type, M, N, K, iters, fresh, stream, sync);
    std::printf("%s %s M=%zu N=%zu K=%zu iters=%d %s stream=%s sync=%d: %d bad iterations\n", type.c_str(), fp.c_str(), M, N,
                K, iters, fresh ? "fresh" : "reuse", sk.c_str(), int(sync), bad);
    return bad ? 1 : 0;
}
