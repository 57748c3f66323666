// tests/cpp/test_cuda_api.cpp -- device test of the C++ drop-in API (bbfft::configuration,
// make_plan, plan::execute, jit_cache_all, aot_cache, generate_fft_kernels, callbacks) on CUDA.
// Re-hosts the checks of the reference's device tests (test/c2c.cpp:17-127, test/r2c.cpp,
// test/callback.cpp:18-114, test/error.cpp:13-21, examples/cache/main.cpp, examples/aot/main.cpp)
// with cudaMalloc instead of sycl::malloc_device.  Built and run by tests/test_gpu_z_cpp_api.py (named to run last: a long binary).
#include "bbfft/aot_cache.hpp"
#include "bbfft/bad_configuration.hpp"
#include "bbfft/configuration.hpp"
#include "bbfft/cuda/device.hpp"
#include "bbfft/cuda/make_plan.hpp"
#include "bbfft/cuda/online_compiler.hpp"
#include "bbfft/generator.hpp"
#include "bbfft/jit_cache_all.hpp"

#include <cuda_runtime_api.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

using namespace bbfft;

static int failures = 0;
#define CHECK(cond)                                                                                \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            ++failures;                                                                            \
            std::printf("CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond);                    \
        }                                                                                          \
    } while (0)
#define CUDA_OK(x)                                                                                 \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) {                                                                   \
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);   \
            std::exit(2);                                                                          \
        }                                                                                          \
    } while (0)

template <typename T> double tol(std::size_t N) {
    return 1e2 * std::numeric_limits<T>::epsilon() * std::sqrt(double(N)); // test/fft.hpp:17-19
}

template <typename T> struct device_vec {
    T *p = nullptr;
    std::size_t n;
    explicit device_vec(std::size_t n_) : n(n_) { CUDA_OK(cudaMalloc(reinterpret_cast<void **>(&p), n * sizeof(T))); }
    ~device_vec() { cudaFree(p); }
    void upload(std::vector<T> const &h) { CUDA_OK(cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice)); }
    std::vector<T> download() const {
        std::vector<T> h(n);
        CUDA_OK(cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost));
        return h;
    }
};

// test/c2c.cpp:17-51: single Fourier mode -> scaled Kronecker delta, in-place
template <typename T> void c2c_analytic(cudaStream_t stream, std::size_t M, std::size_t N, std::size_t K, jit_cache *cache) {
    const double tau = 6.28318530717958647692;
    std::vector<std::complex<T>> x(M * N * K);
    for (std::size_t k = 0; k < K; ++k)
        for (std::size_t n = 0; n < N; ++n)
            for (std::size_t m = 0; m < M; ++m) {
                double arg = tau * double((m + k) % N) * double(n) / double(N);
                double s = (1.0 + double(k) / double(K)) / double(N);
                x[m + n * M + k * M * N] = {T(s * std::cos(arg)), T(s * std::sin(arg))};
            }
    device_vec<std::complex<T>> d(x.size());
    d.upload(x);
    configuration cfg = {1, {M, N, K}, to_precision_v<T>, direction::forward, transform_type::c2c};
    auto plan = make_plan(cfg, stream, cache);
    CHECK(bool(plan));
    plan.execute(d.p).wait();
    auto X = d.download();
    double eps = tol<T>(N);
    for (std::size_t k = 0; k < K; ++k)
        for (std::size_t n = 0; n < N; ++n)
            for (std::size_t m = 0; m < M; ++m) {
                double ref = (n == (m + k) % N) ? 1.0 + double(k) / double(K) : 0.0;
                auto v = X[m + n * M + k * M * N];
                if (std::abs(v.real() - ref) > eps || std::abs(v.imag()) > eps) {
                    ++failures;
                    std::printf("c2c mismatch M=%zu N=%zu K=%zu at (%zu,%zu,%zu): %g %g vs %g\n", M, N, K, m, n, k,
                                double(v.real()), double(v.imag()), ref);
                    return;
                }
            }
}

// test/c2c.cpp:65-82 "c2c non-packed forward": strides {1, M+1, (M+1)(N+1)}, same analytic check
template <typename T> void c2c_nonpacked(cudaStream_t stream, std::size_t M, std::size_t N, std::size_t K) {
    const double tau = 6.28318530717958647692;
    const std::size_t s1 = M + 1, s2 = (M + 1) * (N + 1);
    std::vector<std::complex<T>> x(s2 * K, std::complex<T>(T(7), T(-7))); // gaps hold a sentinel
    for (std::size_t k = 0; k < K; ++k)
        for (std::size_t n = 0; n < N; ++n)
            for (std::size_t m = 0; m < M; ++m) {
                double arg = tau * double((m + k) % N) * double(n) / double(N);
                double sc = (1.0 + double(k) / double(K)) / double(N);
                x[m + n * s1 + k * s2] = {T(sc * std::cos(arg)), T(sc * std::sin(arg))};
            }
    device_vec<std::complex<T>> d(x.size());
    d.upload(x);
    std::array<std::size_t, max_tensor_dim> stride = {1, s1, s2};
    configuration cfg = {1, {M, N, K}, to_precision_v<T>, direction::forward, transform_type::c2c, stride, stride};
    make_plan(cfg, stream).execute(d.p).wait();
    auto X = d.download();
    double eps = tol<T>(N);
    bool ok = true;
    for (std::size_t k = 0; k < K && ok; ++k)
        for (std::size_t n = 0; n <= N && ok; ++n)
            for (std::size_t m = 0; m <= M && ok; ++m) {
                auto v = X[m + n * s1 + k * s2];
                if (m == M || n == N) {
                    ok = v == std::complex<T>(T(7), T(-7)); // padding untouched
                } else {
                    double ref = (n == (m + k) % N) ? 1.0 + double(k) / double(K) : 0.0;
                    ok = std::abs(v.real() - ref) <= eps && std::abs(v.imag()) <= eps;
                }
            }
    CHECK(ok);
}

// test/c2c.cpp:84-127 "c2c identity": forward then backward in place = N * identity
template <typename T> void c2c_identity_inplace(cudaStream_t stream, std::size_t M, std::size_t N, std::size_t K) {
    std::vector<std::complex<T>> x(M * N * K);
    unsigned s = 4321;
    auto rnd = [&] {
        s = s * 1664525u + 1013904223u;
        return T((s >> 8) & 0xffff) / T(65536);
    };
    for (auto &v : x) v = {rnd(), rnd()};
    device_vec<std::complex<T>> d(x.size());
    d.upload(x);
    configuration cfg = {1, {M, N, K}, to_precision_v<T>, direction::forward};
    auto plan = make_plan(cfg, stream);
    cfg.dir = direction::backward;
    auto iplan = make_plan(cfg, stream);
    plan.execute(d.p).wait();
    iplan.execute(d.p).wait();
    auto y = d.download();
    double eps = tol<T>(N);
    bool ok = true;
    for (std::size_t i = 0; i < x.size() && ok; ++i)
        ok = std::abs(double(y[i].real()) / N - double(x[i].real())) <= eps * 2 &&
             std::abs(double(y[i].imag()) / N - double(x[i].imag())) <= eps * 2;
    CHECK(ok);
}

// test/c2c.cpp:84-127: forward o backward = N * identity, out-of-place, with event dependencies
template <typename T> void c2c_identity(cudaStream_t stream, std::size_t M, std::size_t N, std::size_t K) {
    std::vector<std::complex<T>> x(M * N * K);
    unsigned s = 12345;
    for (auto &v : x) {
        s = s * 1664525u + 1013904223u;
        T a = T((s >> 8) & 0xffff) / T(65536);
        s = s * 1664525u + 1013904223u;
        v = {a, T((s >> 8) & 0xffff) / T(65536)};
    }
    device_vec<std::complex<T>> a(x.size()), b(x.size()), c(x.size());
    a.upload(x);
    configuration cf = {1, {M, N, K}, to_precision_v<T>, direction::forward, transform_type::c2c};
    configuration cb = {1, {M, N, K}, to_precision_v<T>, direction::backward, transform_type::c2c};
    auto pf = make_plan(cf, stream);
    auto pb = make_plan(cb, stream);
    auto e1 = pf.execute(a.p, b.p);
    auto e2 = pb.execute(b.p, c.p, e1);
    pb.execute(b.p, c.p, std::vector<cuda::event>{e1, e2}).wait();
    auto y = c.download();
    double err = 0, nrm = 0;
    for (std::size_t i = 0; i < x.size(); ++i) {
        err += std::norm(std::complex<double>(y[i]) / double(N) - std::complex<double>(x[i]));
        nrm += std::norm(std::complex<double>(x[i]));
    }
    CHECK(std::sqrt(err / nrm) < (sizeof(T) == 4 ? 1e-5 : 1e-12));
}

// test/r2c.cpp: cosine -> two deltas (out-of-place), then c2r with polluted imag(X[0]) (:310-324)
template <typename T> void r2c_c2r(cudaStream_t stream, std::size_t M, std::size_t N, std::size_t K) {
    const double tau = 6.28318530717958647692;
    std::size_t Nh = N / 2 + 1;
    std::vector<T> x(M * N * K);
    for (std::size_t k = 0; k < K; ++k)
        for (std::size_t n = 0; n < N; ++n)
            for (std::size_t m = 0; m < M; ++m)
                x[m + n * M + k * M * N] = T((1.0 + double(k) / K) * std::cos(tau * double((m + k) % N) * double(n) / N) / N);
    device_vec<T> dx(x.size()), dback(x.size());
    device_vec<std::complex<T>> dX(M * Nh * K);
    dx.upload(x);
    configuration cfg = {1, {M, N, K}, to_precision_v<T>, direction::forward, transform_type::r2c};
    cfg.set_strides_default(false);
    make_plan(cfg, stream).execute(dx.p, dX.p).wait();
    auto X = dX.download();
    double eps = tol<T>(N);
    bool ok = true;
    for (std::size_t k = 0; k < K && ok; ++k)
        for (std::size_t n = 0; n < Nh && ok; ++n)
            for (std::size_t m = 0; m < M && ok; ++m) {
                long b = long((m + k) % N);
                double ref = ((long(n) - b) % long(N) == 0 ? 0.5 : 0.0) + ((long(n) + b) % long(N) == 0 ? 0.5 : 0.0);
                ref *= 1.0 + double(k) / K;
                auto v = X[m + n * M + k * M * Nh];
                ok = std::abs(v.real() - ref) <= eps && std::abs(v.imag()) <= eps;
            }
    CHECK(ok);
    for (std::size_t k = 0; k < K; ++k)
        for (std::size_t m = 0; m < M; ++m) X[m + k * M * Nh].imag(T(1.0 + m + k)); // pollute
    dX.upload(X);
    configuration cb = {1, {M, N, K}, to_precision_v<T>, direction::backward, transform_type::c2r};
    cb.set_strides_default(false);
    make_plan(cb, stream).execute(dX.p, dback.p).wait();
    auto back = dback.download();
    ok = true;
    for (std::size_t i = 0; i < x.size() && ok; ++i) ok = std::abs(double(back[i]) - double(x[i]) * N) <= eps * 2;
    CHECK(ok);
}

// test/r2c.cpp:182-252, 299-372 (2d / 3d cases): product of cosines (batch b shifts the frequency of
// the first mode) -> product of two-delta spectra, r2c forward then c2r backward = prod(N) * input,
// in-place (padded rows) and out-of-place
template <typename T, std::size_t D>
void real_nd(cudaStream_t stream, std::size_t M, std::array<std::size_t, D> N, std::size_t K, bool inplace,
             bool pollute_0 = false) {
    const double tau = 6.28318530717958647692;
    const std::size_t Nh = N[0] / 2 + 1, N0r = inplace ? 2 * Nh : N[0];
    std::size_t rest = 1, total = 1;
    for (std::size_t d = 1; d < D; ++d) rest *= N[d];
    for (std::size_t d = 0; d < D; ++d) total *= N[d];
    const std::size_t nreal = M * N0r * rest * K, nspec = M * Nh * rest * K;
    std::vector<T> x(nreal, T(0));
    auto freq = [&](std::size_t m, std::size_t k) { return (m + k) % N[0]; };
    for (std::size_t k = 0; k < K; ++k)
        for (std::size_t r = 0; r < rest; ++r)
            for (std::size_t n = 0; n < N[0]; ++n)
                for (std::size_t m = 0; m < M; ++m) {
                    double v = std::cos(tau * double(freq(m, k)) * double(n) / double(N[0])) / double(N[0]);
                    std::size_t rr = r;
                    for (std::size_t d = 1; d < D; ++d) {
                        v *= std::cos(tau * double(rr % N[d]) / double(N[d])) / double(N[d]);
                        rr /= N[d];
                    }
                    x[m + M * (n + N0r * (r + rest * k))] = T(v);
                }
    tensor_extent shape = {};
    shape[0] = M;
    for (std::size_t d = 0; d < D; ++d) shape[d + 1] = N[d];
    shape[D + 1] = K;
    configuration cf = {unsigned(D), shape, to_precision_v<T>, direction::forward, transform_type::r2c};
    cf.set_strides_default(inplace);
    configuration cb = {unsigned(D), shape, to_precision_v<T>, direction::backward, transform_type::c2r};
    cb.set_strides_default(inplace);
    device_vec<T> dreal(std::max(nreal, 2 * nspec));
    device_vec<std::complex<T>> dspec(nspec);
    {
        std::vector<T> up(dreal.n, T(0));
        std::copy(x.begin(), x.end(), up.begin());
        dreal.upload(up);
    }
    void *spec_ptr = inplace ? static_cast<void *>(dreal.p) : static_cast<void *>(dspec.p);
    make_plan(cf, stream).execute(dreal.p, spec_ptr).wait();
    std::vector<std::complex<T>> X(nspec);
    CUDA_OK(cudaMemcpy(X.data(), spec_ptr, nspec * sizeof(std::complex<T>), cudaMemcpyDeviceToHost));
    std::size_t nmax = 0;
    for (std::size_t d = 0; d < D; ++d) nmax = std::max(nmax, N[d]);
    const double eps = tol<T>(nmax);
    auto two_delta = [](long k, long f, long n) {
        return ((k - f) % n == 0 ? 0.5 : 0.0) + ((k + f) % n == 0 ? 0.5 : 0.0);
    };
    bool ok = true;
    for (std::size_t k = 0; k < K && ok; ++k)
        for (std::size_t r = 0; r < rest && ok; ++r)
            for (std::size_t n = 0; n < Nh && ok; ++n)
                for (std::size_t m = 0; m < M && ok; ++m) {
                    double ref = two_delta(long(n), long(freq(m, k)), long(N[0]));
                    std::size_t rr = r;
                    for (std::size_t d = 1; d < D; ++d) {
                        ref *= two_delta(long(rr % N[d]), 1, long(N[d]));
                        rr /= N[d];
                    }
                    auto v = X[m + M * (n + Nh * (r + rest * k))];
                    ok = std::abs(double(v.real()) - ref) <= eps && std::abs(double(v.imag())) <= eps;
                    if (!ok) std::printf("r2c %zud M=%zu K=%zu inplace=%d mismatch at (%zu,%zu,%zu,%zu): %g %g vs %g\n", D, M, K, int(inplace), m, n, r, k, double(v.real()), double(v.imag()), ref);
                }
    CHECK(ok);
    // backward from the exact spectrum we just checked
    device_vec<T> dback(std::max(nreal, 2 * nspec));
    void *back_in = inplace ? static_cast<void *>(dreal.p) : static_cast<void *>(dspec.p);
    if (pollute_0) {
        // imag(X[0]) != 0 is faulty usage that c2r must ignore (test/r2c.cpp:310-324)
        for (std::size_t k = 0; k < K; ++k)
            for (std::size_t m = 0; m < M; ++m) X[m + M * Nh * rest * k].imag(T(1.0 + m + k));
        CUDA_OK(cudaMemcpy(back_in, X.data(), nspec * sizeof(std::complex<T>), cudaMemcpyHostToDevice));
    }
    void *back_out = inplace ? static_cast<void *>(dreal.p) : static_cast<void *>(dback.p);
    make_plan(cb, stream).execute(back_in, back_out).wait();
    std::vector<T> back(nreal);
    CUDA_OK(cudaMemcpy(back.data(), back_out, nreal * sizeof(T), cudaMemcpyDeviceToHost));
    ok = true;
    for (std::size_t k = 0; k < K && ok; ++k)
        for (std::size_t r = 0; r < rest && ok; ++r)
            for (std::size_t n = 0; n < N[0] && ok; ++n)
                for (std::size_t m = 0; m < M && ok; ++m) {
                    const std::size_t i = m + M * (n + N0r * (r + rest * k));
                    ok = std::abs(double(back[i]) - double(x[i]) * double(total)) <= eps * 4;
                }
    CHECK(ok);
}

// test/callback.cpp:18-114 (load) and :116-202 (store): a c2r plan whose load callback implies a
// zero upper half of the spectrum equals the plain plan on the zero-padded spectrum, and an r2c
// plan whose store callback keeps the first N/4 bins scaled by 1/N equals the truncated, scaled
// plain result -- both compared with == (the transform arithmetic is the same code)
template <typename T> void callbacks_bit_identical(cudaStream_t stream, std::size_t M, std::size_t N) {
    const std::size_t K = 4;
    char const *real = sizeof(T) == 4 ? "float" : "double";
    unsigned s = 777;
    auto rnd = [&] {
        s = s * 1664525u + 1013904223u;
        return T((s >> 8) & 0xffff) / T(65536);
    };
    {
        const std::size_t Next = 2 * N, ns_ref = Next / 2 + 1, ns = N / 2 + 1;
        std::vector<std::complex<T>> X(M * ns * K), Xref(M * ns_ref * K, std::complex<T>(0, 0));
        for (std::size_t k = 0; k < K; ++k)
            for (std::size_t n = 0; n < ns; ++n)
                for (std::size_t m = 0; m < M; ++m) {
                    X[m + M * (n + ns * k)] = {rnd(), rnd()};
                    Xref[m + M * (n + ns_ref * k)] = X[m + M * (n + ns * k)];
                }
        std::ostringstream src;
        src << real << "2 load(global " << real << "2* in, size_t offset) {\n"
            << "    size_t m = offset % " << M << ";\n    size_t n = offset / " << M << " % " << ns_ref << ";\n"
            << "    size_t k = offset / " << (M * ns_ref) << ";\n"
            << "    if (n < " << ns << ") { return in[m + " << M << " * (n + " << ns << " * k)]; }\n"
            << "    return 0;\n}\n";
        const std::string code = src.str();
        configuration ref = {1, {M, Next, K}, to_precision_v<T>, direction::backward, transform_type::c2r,
                             {1, M, M * ns_ref}, {1, M, M * Next}};
        configuration cb = ref;
        cb.callbacks = {code.c_str(), code.size(), "load", nullptr};
        device_vec<std::complex<T>> dX(X.size()), dXref(Xref.size());
        device_vec<T> dx(M * Next * K), dxref(M * Next * K);
        dX.upload(X);
        dXref.upload(Xref);
        make_plan(ref, stream).execute(dXref.p, dxref.p).wait();
        make_plan(cb, stream).execute(dX.p, dx.p).wait();
        auto got = dx.download(), want = dxref.download();
        {
            // two plans that are wrong the same way compare equal: hold the plain plan against a
            // direct DFT as well (columns m = 0 and m = M-1 of every k; imag(X[0]) is ignored)
            const long double tau = 6.283185307179586476925286766559005768L;
            bool ok = true;
            double worst = 0;
            for (std::size_t k = 0; k < K; ++k)
                for (std::size_t m : {std::size_t(0), M - 1})
                    for (std::size_t n = 0; n < Next; ++n) {
                        long double acc = Xref[m + M * (0 + ns_ref * k)].real();
                        for (std::size_t q = 1; q < ns_ref; ++q) {
                            auto v = Xref[m + M * (q + ns_ref * k)];
                            long double a = tau * ((q * n) % Next) / Next;
                            long double w = (2 * q == Next) ? 1.0L : 2.0L;
                            acc += w * (v.real() * cosl(a) - ((2 * q == Next) ? 0.0L : v.imag() * sinl(a)));
                        }
                        double err = std::abs(double(want[m + M * (n + Next * k)]) - double(acc));
                        worst = std::max(worst, err);
                        ok = ok && err <= tol<T>(Next) * double(Next);
                    }
            if (!ok) std::printf("plain c2r plan %s M=%zu N=%zu differs from the direct DFT by %g\n", real, M, Next, worst);
            CHECK(ok);
        }
        std::size_t ndiff = 0, first = 0;
        for (std::size_t i = 0; i < got.size(); ++i) {
            if (!(got[i] == want[i])) {
                if (!ndiff) first = i;
                ++ndiff;
            }
        }
        if (ndiff) {
            std::printf("load callback %s M=%zu N=%zu: %zu of %zu elements differ, first at (m=%zu, n=%zu, k=%zu): %.9g vs %.9g\n",
                        real, M, N, ndiff, got.size(), first % M, first / M % Next, first / (M * Next), double(got[first]),
                        double(want[first]));
            // which side moved?  (diagnostic only: the check below still fails)
            make_plan(ref, stream).execute(dXref.p, dxref.p).wait();
            make_plan(cb, stream).execute(dX.p, dx.p).wait();
            auto got2 = dx.download(), want2 = dxref.download();
            std::printf("  second execution: callback plan %s, plain plan %s, callback == plain: %s\n",
                        got2 == got ? "reproduced itself" : "CHANGED", want2 == want ? "reproduced itself" : "CHANGED",
                        got2 == want2 ? "yes" : "no");
        }
        CHECK(ndiff == 0);
    }
    {
        const std::size_t N2 = 4 * N, ncut = N2 / 4, nspec = N2 / 2 + 1;
        std::vector<T> x(M * N2 * K);
        for (auto &v : x) v = rnd();
        std::ostringstream src;
        src << "void store(global " << real << "2* out, size_t offset, " << real << "2 value) {\n"
            << "    size_t m = offset % " << M << ";\n    size_t n = offset / " << M << " % " << nspec << ";\n"
            << "    size_t k = offset / " << (M * nspec) << ";\n"
            << "    if (n < " << ncut << ") { out[m + " << M << " * (n + " << ncut << " * k)] = value * ((" << real << ") "
            << std::hexfloat << (1.0 / double(N2)) << std::defaultfloat << "); }\n}\n";
        const std::string code = src.str();
        configuration ref = {1, {M, N2, K}, to_precision_v<T>, direction::forward, transform_type::r2c,
                             {1, M, M * N2}, {1, M, M * nspec}};
        configuration cb = ref;
        cb.callbacks = {code.c_str(), code.size(), nullptr, "store"};
        device_vec<T> dx(x.size());
        device_vec<std::complex<T>> dref(M * nspec * K), dcut(M * ncut * K);
        dx.upload(x);
        make_plan(ref, stream).execute(dx.p, dref.p).wait();
        make_plan(cb, stream).execute(dx.p, dcut.p).wait();
        auto full = dref.download();
        auto cut = dcut.download();
        bool ok = true;
        const T scale = T(1.0 / double(N2));
        for (std::size_t k = 0; k < K && ok; ++k)
            for (std::size_t n = 0; n < ncut && ok; ++n)
                for (std::size_t m = 0; m < M && ok; ++m) {
                    auto want = full[m + M * (n + nspec * k)] * scale;
                    ok = cut[m + M * (n + ncut * k)] == want;
                }
        CHECK(ok);
    }
}

int main() {
    int ndev = 0;
    CUDA_OK(cudaGetDeviceCount(&ndev));
    if (ndev == 0) {
        std::printf("no CUDA device\n");
        return 2;
    }
    cudaStream_t stream;
    CUDA_OK(cudaStreamCreate(&stream));
    auto info = get_device_info(0);
    std::printf("device_info %s id %llx\n", info.to_string().c_str(), (unsigned long long)get_device_id(0));
    CHECK(info.max_work_group_size == 1024);
    CHECK(info.subgroup_sizes.size() == 1 && info.subgroup_sizes[0] == 32);

    // --- c2c analytic, reference size lists (test/c2c.cpp:56-59), shared jit cache
    jit_cache_all cache;
    for (std::size_t M : {1u, 2u, 3u, 16u, 17u, 64u, 256u, 1024u})
        for (std::size_t N : {2u, 3u, 5u, 7u, 11u, 13u, 4u, 8u, 16u, 32u, 128u, 256u, 512u, 27u, 63u, 105u, 363u})
            for (std::size_t K : {1u, 32u}) {
                c2c_analytic<float>(stream, M, N, K, &cache);
                c2c_analytic<double>(stream, M, N, K, &cache);
            }
    std::size_t cached = cache.kernel_names().size();
    CHECK(cached > 0);
    // examples/cache/main.cpp:50-52: a plan that differs only in K re-uses the cached kernel
    c2c_analytic<float>(stream, 16, 128, 77, &cache);
    CHECK(cache.kernel_names().size() == cached);
    c2c_identity<float>(stream, 16, 200, 40);
    c2c_identity<double>(stream, 3, 343, 9);
    // test/c2c.cpp:65-82 (non-packed) and :84-127 (identity)
    for (std::size_t M : {1u, 32u}) {
        for (std::size_t N : {6u, 17u, 102u}) {
            c2c_nonpacked<float>(stream, M, N, 33);
            c2c_nonpacked<double>(stream, M, N, 33);
        }
        for (std::size_t N : {5u, 63u, 92u}) {
            c2c_identity_inplace<float>(stream, M, N, 16);
            c2c_identity_inplace<double>(stream, M, N, 16);
        }
    }

    // --- real transforms 1d: test/r2c.cpp:182-190 (r2c out-of-place) with the c2r mirror, :299-308 (c2r
    // out-of-place lists), :192-200 (r2c in-place), :310-324 (c2r in-place, polluted imag(X[0]))
    for (std::size_t M : {1u, 3u, 32u})
        for (std::size_t N : {2u, 4u, 5u, 8u, 27u, 16u, 32u, 128u, 105u, 256u, 512u, 102u, 220u, 10u, 26u})
            for (std::size_t K : {1u, 33u}) {
                r2c_c2r<float>(stream, M, N, K);
                r2c_c2r<double>(stream, M, N, K);
            }
    for (std::size_t M : {1u, 5u, 32u})
        for (std::size_t N : {11u, 16u, 25u, 96u, 256u, 300u, 315u}) {
            r2c_c2r<float>(stream, M, N, 33);
            r2c_c2r<double>(stream, M, N, 33);
        }
    for (std::size_t M : {1u, 3u})
        for (std::size_t N : {4u, 12u, 13u, 110u}) {
            real_nd<float, 1>(stream, M, {N}, 65, true);
            real_nd<double, 1>(stream, M, {N}, 65, true);
        }
    for (std::size_t N : {5u, 16u, 21u, 48u, 512u}) {
        real_nd<float, 1>(stream, 1, {N}, 33, true, true);
        real_nd<double, 1>(stream, 1, {N}, 33, true, true);
    }

    // --- real 2d / 3d, the reference's case lists (test/r2c.cpp:206-252 forward, :338-372 backward)
    for (std::size_t M : {1u, 3u})
        for (std::size_t K : {1u, 33u}) {
            for (auto N : {std::array<std::size_t, 2>{4, 8}, std::array<std::size_t, 2>{8, 5}}) {
                real_nd<float, 2>(stream, M, N, K, true);
                real_nd<double, 2>(stream, M, N, K, true);
            }
            for (auto N : {std::array<std::size_t, 3>{4, 8, 2}, std::array<std::size_t, 3>{8, 256, 5}}) {
                real_nd<float, 3>(stream, M, N, K, true);
                real_nd<double, 3>(stream, M, N, K, true);
            }
        }
    // c2r 2d / 3d in-place lists (test/r2c.cpp:326-346)
    for (std::size_t M : {1u, 3u})
        for (std::size_t K : {1u, 54u}) {
            for (auto N : {std::array<std::size_t, 2>{4, 2}, std::array<std::size_t, 2>{96, 96}}) {
                real_nd<float, 2>(stream, M, N, K, true);
                real_nd<double, 2>(stream, M, N, K, true);
            }
            for (auto N : {std::array<std::size_t, 3>{4, 8, 2}, std::array<std::size_t, 3>{96, 96, 80}}) {
                if (N[0] == 96 && M * K > 54) continue; // 3 x 54 x 96 x 96 x 80 reals: 1 GB per precision on the host
                real_nd<float, 3>(stream, M, N, K, true);
                real_nd<double, 3>(stream, M, N, K, true);
            }
        }
    for (std::size_t M : {1u, 7u})
        for (std::size_t K : {1u, 65u}) {
            for (auto N : {std::array<std::size_t, 2>{5, 4}, std::array<std::size_t, 2>{10, 3}}) {
                real_nd<float, 2>(stream, M, N, K, false);
                real_nd<double, 2>(stream, M, N, K, false);
            }
            for (auto N : {std::array<std::size_t, 3>{5, 4, 6}, std::array<std::size_t, 3>{10, 286, 3}}) {
                real_nd<float, 3>(stream, M, N, K, false);
                real_nd<double, 3>(stream, M, N, K, false);
            }
        }

    // --- callbacks, bit-identical to the plain plans (test/callback.cpp sizes: M in {1,32}, N in {8,64,212})
    for (std::size_t M : {1u, 32u})
        for (std::size_t N : {8u, 64u, 212u}) {
            callbacks_bit_identical<float>(stream, M, N);
            callbacks_bit_identical<double>(stream, M, N);
        }

    // --- errors (test/error.cpp:13-21, small_batch_fft.hpp:107-110)
    for (unsigned dim : {0u, 4u}) {
        bool thrown = false;
        try {
            configuration cfg = {dim, {1, 8, 1}, precision::f32};
            make_plan(cfg, stream);
        } catch (bad_configuration const &) {
            thrown = true;
        }
        CHECK(thrown);
    }
    {
        configuration cfg = {1, {64, 8, 4}, precision::f32, direction::forward, transform_type::r2c};
        auto plan = make_plan(cfg, stream); // creation succeeds, in-place execute throws
        device_vec<float> buf(64 * 10 * 4);
        bool thrown = false;
        try {
            plan.execute(buf.p);
        } catch (bad_configuration const &) {
            thrown = true;
        }
        CHECK(thrown);
    }

    // --- AOT: generate_fft_kernels -> cubin -> aot_cache (examples/aot/main.cpp:58-73)
    {
        configuration cfg = {1, {16, 48, 10}, precision::f64, direction::backward, transform_type::c2c};
        std::ostringstream src;
        auto names = generate_fft_kernels(src, {cfg}, info);
        CHECK(names.size() == 1);
        auto bin = cuda::compile_to_native(src.str());
        aot_cache aot;
        aot.register_module(cuda::create_aot_module(bin.data(), bin.size(), module_format::native, 0));
        CHECK(bool(aot.get({names[0], get_device_id(0)})));
        CHECK(!aot.get({"no_such_kernel", get_device_id(0)}));
        auto plan = make_plan(cfg, stream, &aot);
        device_vec<std::complex<double>> d(16 * 48 * 10);
        std::vector<std::complex<double>> h(d.n, {1.0, 0.0});
        d.upload(h);
        plan.execute(d.p).wait();
        auto r = d.download();
        CHECK(std::abs(r[0].real() - 48.0) < 1e-12 && std::abs(r[16].real()) < 1e-12);
    }

    // --- 3d plan + callback (OpenCL-C source as in test/callback.cpp:151-158) smoke
    {
        configuration cfg = {3, {1, 8, 4, 6, 2}, precision::f32, direction::forward, transform_type::c2c};
        auto plan = make_plan(cfg, stream);
        device_vec<std::complex<float>> d(8 * 4 * 6 * 2);
        std::vector<std::complex<float>> h(d.n, {1.0f, 0.0f});
        d.upload(h);
        plan.execute(d.p).wait();
        auto r = d.download();
        CHECK(std::abs(r[0].real() - 192.0f) < 1e-3f && std::abs(r[1].real()) < 1e-3f);
        char const src[] = "void store(global float2* out, size_t offset, float2 value) { out[offset] = value * 0.5f; }";
        configuration cc = {1, {4, 16, 3}, precision::f32, direction::forward, transform_type::c2c};
        cc.callbacks = {src, sizeof(src) - 1, nullptr, "store"};
        auto pc = make_plan(cc, stream);
        device_vec<std::complex<float>> e(4 * 16 * 3);
        std::vector<std::complex<float>> he(e.n, {1.0f, 0.0f});
        e.upload(he);
        pc.execute(e.p).wait();
        auto rc = e.download();
        CHECK(rc[0].real() == 8.0f);
        bool thrown = false;
        try {
            cfg.callbacks = cc.callbacks;
            make_plan(cfg, stream); // callbacks are 1d only (nd_fft.hpp:29-31)
        } catch (bad_configuration const &) {
            thrown = true;
        }
        CHECK(thrown);
    }
    cudaStreamDestroy(stream);
    std::printf("%s (%d failures)\n", failures ? "FAILED" : "OK", failures);
    return failures ? 1 : 0;
}
