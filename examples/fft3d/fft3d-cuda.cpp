// examples/fft3d/fft3d-cuda.cpp -- 3d r2c / c2c transform through the bbfft CUDA backend, with the
// command line, the product-of-modes check and the "<time> s, <GB/s>" report of the reference's
// fft3d example (examples/fft3d/main.cpp:12-41, common.hpp:13-62, utility.hpp:39-105), and cuFFT
// timed on the same tensor beside it (the reference's fft3d-cufft.cpp:15-73).
//
//   fft3d-cuda [-iodscrv] [-n repeats] [-x] N1 N2 N3 [K]
//     -i / -o   in-place / out-of-place          (default in-place)
//     -d / -s   double / single precision        (default single)
//     -c / -r   c2c / r2c                        (default r2c)
//     -x        also time cuFFT (cufftPlanMany)  (default off)
//     K         batch of independent 3d transforms (default 1)
//
// Build: g++ -std=c++17 -I include -I $CUDA_HOME/include fft3d-cuda.cpp -L <libdir> -lbbfft_cuda
//        -L $CUDA_HOME/lib64 -lcufft -lcudart  (tests/test_gpu_examples.py does exactly this)
#include "bbfft/configuration.hpp"
#include "bbfft/cuda/make_plan.hpp"

#include <cuda_runtime_api.h>
#include <cufft.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace {

struct options {
    bool inplace = true, dp = false, r2c = true, verbose = false, cufft = false;
    int repeats = 1;
    std::size_t n[3] = {0, 0, 0}, K = 1;
};

[[noreturn]] void usage() {
    std::fprintf(stderr, "Usage: fft3d-cuda [-iodscrvx] [-n repeats] <N1> <N2> <N3> [K]\n");
    std::exit(1);
}

options parse(int argc, char **argv) {
    options o;
    int pos = 0;
    for (int i = 1; i < argc; ++i) {
        if (argv[i][0] == '-' && argv[i][1] == 'n' && argv[i][2] == 0 && i + 1 < argc) {
            o.repeats = std::max(1, std::atoi(argv[++i]));
        } else if (argv[i][0] == '-') {
            for (char const *c = argv[i] + 1; *c; ++c) {
                switch (*c) {
                case 'i': o.inplace = true; break;
                case 'o': o.inplace = false; break;
                case 'd': o.dp = true; break;
                case 's': o.dp = false; break;
                case 'c': o.r2c = false; break;
                case 'r': o.r2c = true; break;
                case 'v': o.verbose = true; break;
                case 'x': o.cufft = true; break;
                default: usage();
                }
            }
        } else if (pos < 3) {
            o.n[pos++] = std::strtoull(argv[i], nullptr, 10);
        } else if (pos == 3) {
            o.K = std::strtoull(argv[i], nullptr, 10);
            ++pos;
        } else {
            usage();
        }
    }
    if (pos < 3 || !o.n[0] || !o.n[1] || !o.n[2] || !o.K) usage();
    return o;
}

void cuda_ok(cudaError_t e, char const *what) {
    if (e != cudaSuccess) {
        std::fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e));
        std::exit(2);
    }
}

// one mode of the separable test signal: exp(2 pi i j / N) / N (complex) or its real part; the
// spectrum of the product is a single 1 at (1,1,1), plus mirrored halves along N1 for real input
template <typename R> std::complex<R> mode(std::size_t j, std::size_t N) {
    const double arg = 6.28318530717958647693 * double(j) / double(N);
    return {R(std::cos(arg) / double(N)), R(std::sin(arg) / double(N))};
}
template <typename R> R expected(std::size_t k, std::size_t N, bool real_input) {
    const bool at1 = k % N == 1 % N, atm1 = (k + 1) % N == 0;
    if (!real_input) return at1 ? R(1) : R(0);
    return R(((at1 ? 1.0 : 0.0) + (atm1 ? 1.0 : 0.0)) / 2.0);
}

template <typename F> double best_of(int times, F &&f) {
    double best = std::numeric_limits<double>::max();
    for (int i = 0; i < times; ++i) {
        auto t0 = std::chrono::high_resolution_clock::now();
        f();
        std::chrono::duration<double> dt = std::chrono::high_resolution_clock::now() - t0;
        best = std::min(best, dt.count());
    }
    return best;
}

template <typename R> int run(options const &o) {
    using cplx = std::complex<R>;
    const std::size_t N1 = o.n[0], N2 = o.n[1], N3 = o.n[2], K = o.K;
    const std::size_t N1s = o.r2c ? N1 / 2 + 1 : N1;                      // spectrum rows
    const std::size_t N1r = o.r2c ? (o.inplace ? 2 * N1s : N1) : 0;       // stored real rows
    const std::size_t in_elems = (o.r2c ? N1r : N1) * N2 * N3 * K;        // in units of the input type
    const std::size_t out_elems = N1s * N2 * N3 * K;                      // complex
    const std::size_t in_bytes = in_elems * (o.r2c ? sizeof(R) : sizeof(cplx));
    const std::size_t out_bytes = out_elems * sizeof(cplx);
    std::printf("%zu x %zu x %zu%s\n", N1, N2, N3, K > 1 ? (" x " + std::to_string(K)).c_str() : "");

    std::vector<unsigned char> host(std::max(in_bytes, out_bytes), 0);
    for (std::size_t k = 0; k < K; ++k)
        for (std::size_t j3 = 0; j3 < N3; ++j3)
            for (std::size_t j2 = 0; j2 < N2; ++j2)
                for (std::size_t j1 = 0; j1 < N1; ++j1) {
                    if (o.r2c) {
                        const std::size_t idx = j1 + N1r * (j2 + N2 * (j3 + N3 * k));
                        reinterpret_cast<R *>(host.data())[idx] =
                            mode<R>(j1, N1).real() * mode<R>(j2, N2).real() * mode<R>(j3, N3).real();
                    } else {
                        const std::size_t idx = j1 + N1 * (j2 + N2 * (j3 + N3 * k));
                        reinterpret_cast<cplx *>(host.data())[idx] = mode<R>(j1, N1) * mode<R>(j2, N2) * mode<R>(j3, N3);
                    }
                }
    void *din = nullptr, *dout = nullptr;
    cuda_ok(cudaMalloc(&din, std::max(in_bytes, o.inplace ? out_bytes : in_bytes)), "cudaMalloc");
    if (o.inplace) {
        dout = din;
    } else {
        cuda_ok(cudaMalloc(&dout, out_bytes), "cudaMalloc");
    }
    auto upload = [&] { cuda_ok(cudaMemcpy(din, host.data(), in_bytes, cudaMemcpyHostToDevice), "H2D"); };
    std::vector<unsigned char> input_copy(host.begin(), host.begin() + in_bytes);

    cudaStream_t stream;
    cuda_ok(cudaStreamCreate(&stream), "cudaStreamCreate");
    bbfft::configuration cfg = {3,
                                {1, N1, N2, N3, K},
                                o.dp ? bbfft::precision::f64 : bbfft::precision::f32,
                                bbfft::direction::forward,
                                o.r2c ? bbfft::transform_type::r2c : bbfft::transform_type::c2c};
    cfg.set_strides_default(o.inplace);
    auto plan = bbfft::make_plan(cfg, stream);

    upload();
    plan.execute(din, dout).wait();
    cuda_ok(cudaMemcpy(host.data(), dout, out_bytes, cudaMemcpyDeviceToHost), "D2H");
    const double tol = 1.0e2 * std::numeric_limits<R>::epsilon() * std::sqrt(double(std::max({N1, N2, N3})));
    for (std::size_t k = 0; k < K; ++k)
        for (std::size_t j3 = 0; j3 < N3; ++j3)
            for (std::size_t j2 = 0; j2 < N2; ++j2)
                for (std::size_t j1 = 0; j1 < N1s; ++j1) {
                    const cplx got = reinterpret_cast<cplx *>(host.data())[j1 + N1s * (j2 + N2 * (j3 + N3 * k))];
                    const R want = expected<R>(j1, N1, o.r2c) * expected<R>(j2, N2, o.r2c) * expected<R>(j3, N3, o.r2c);
                    if (std::abs(got - cplx(want, 0)) > tol) {
                        std::fprintf(stderr, "FFT error (%zu, %zu, %zu; k=%zu): (%g,%g) != %g\n", j1, j2, j3, k,
                                     double(got.real()), double(got.imag()), double(want));
                        return 1;
                    }
                }
    // in-place transforms overwrite their input: the timing loop re-transforms whatever is in the buffer
    const int nrep = o.repeats;
    auto exec = [&] {
        for (int r = 0; r < nrep; ++r) plan.execute(din, dout);
        cuda_ok(cudaStreamSynchronize(stream), "sync");
    };
    const double t = best_of(10, exec) / nrep;
    const double bytes = 2.0 * sizeof(cplx) * double(N1s * N2 * N3 * K); // reference: 2 * sizeof(complex) * size
    std::printf("%g s, %g GB/s\n", t, bytes / t * 1e-9);

    if (o.cufft) {
        cufftHandle h;
        int nn[3] = {int(N3), int(N2), int(N1)};
        const cufftType type = o.r2c ? (o.dp ? CUFFT_D2Z : CUFFT_R2C) : (o.dp ? CUFFT_Z2Z : CUFFT_C2C);
        int inembed[3] = {int(N3), int(N2), int(o.r2c ? N1r : N1)}, onembed[3] = {int(N3), int(N2), int(N1s)};
        if (cufftPlanMany(&h, 3, nn, inembed, 1, inembed[0] * inembed[1] * inembed[2], onembed, 1,
                          onembed[0] * onembed[1] * onembed[2], type, int(K)) != CUFFT_SUCCESS) {
            std::fprintf(stderr, "cufftPlanMany failed\n");
            return 1;
        }
        cufftSetStream(h, stream);
        auto cexec = [&] {
            for (int r = 0; r < nrep; ++r) {
                if (o.r2c && o.dp) cufftExecD2Z(h, static_cast<cufftDoubleReal *>(din), static_cast<cufftDoubleComplex *>(dout));
                if (o.r2c && !o.dp) cufftExecR2C(h, static_cast<cufftReal *>(din), static_cast<cufftComplex *>(dout));
                if (!o.r2c && o.dp) cufftExecZ2Z(h, static_cast<cufftDoubleComplex *>(din), static_cast<cufftDoubleComplex *>(dout), CUFFT_FORWARD);
                if (!o.r2c && !o.dp) cufftExecC2C(h, static_cast<cufftComplex *>(din), static_cast<cufftComplex *>(dout), CUFFT_FORWARD);
            }
            cuda_ok(cudaStreamSynchronize(stream), "sync");
        };
        std::memcpy(host.data(), input_copy.data(), in_bytes);
        upload();
        const double tc = best_of(10, cexec) / nrep;
        std::printf("cuFFT: %g s, %g GB/s (bbfft speed-up %.2fx)\n", tc, bytes / tc * 1e-9, tc / t);
        cufftDestroy(h);
    }
    if (!o.inplace) cudaFree(dout);
    cudaFree(din);
    cudaStreamDestroy(stream);
    return 0;
}

} // namespace

int main(int argc, char **argv) {
    options o = parse(argc, argv);
    std::printf("%s %s %s\n", o.r2c ? "r2c" : "c2c", o.dp ? "double" : "single", o.inplace ? "in-place" : "out-of-place");
    return o.dp ? run<double>(o) : run<float>(o);
}
