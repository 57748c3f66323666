// examples/transpose_fft_transpose/tft-cuda.cu -- what the double-batched layout buys.
//
// An M x N x K tensor (m fastest) is to be transformed along n.  Either
//   (a) transpose to N x M x K, run a unit-stride batched FFT, transpose back -- three kernels and
//       three HBM round trips, the classic way around a library without the M mode -- or
//   (b) hand the tensor to ONE double-batched plan {M, N, K}, which walks m along the lanes.
// And one step beyond the reference's example ("fused transpose_fft_transpose"):
//   (c) when the RESULT is wanted with n fastest (N x M x K, e.g. for a following unit-stride stage of a
//       pencil decomposition), the transposing store is fused into the FFT kernel through a store callback
//       -- one kernel, one HBM round trip -- and compared with (b) followed by a separate transpose.
// Same experiment, shapes, initial data, check and report as the reference's
// examples/transpose_fft_transpose/tft.cpp (tests (M,N) = (16,16), (1120,32), (128,512), (70,16),
// fp64, ~512e6 bytes per tensor), with the transposes written as a CUDA tile kernel.
//
//   tft-cuda [K]        (K = 0 or absent: size the batch to ~512e6 bytes per tensor)
//
// Build: nvcc -std=c++17 -arch=sm_100a -I include tft-cuda.cu -L <libdir> -lbbfft_cuda
#include "bbfft/configuration.hpp"
#include "bbfft/cuda/make_plan.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

namespace {

constexpr int TILE = 32;

// out[n + m*N + k*M*N] = in[m + n*M + k*M*N]; one 32 x 32 tile per CTA through padded shared memory,
// both sides coalesced
template <typename T> __global__ void transpose_tiles(T const *__restrict__ in, T *__restrict__ out, int M, int N) {
    __shared__ T tile[TILE][TILE + 1];
    const size_t slab = size_t(blockIdx.z) * size_t(M) * size_t(N);
    const int m0 = blockIdx.x * TILE, n0 = blockIdx.y * TILE;
    for (int j = threadIdx.y; j < TILE; j += blockDim.y) {
        const int m = m0 + threadIdx.x, n = n0 + j;
        if (m < M && n < N) tile[j][threadIdx.x] = in[slab + m + size_t(n) * M];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < TILE; j += blockDim.y) {
        const int n = n0 + threadIdx.x, m = m0 + j;
        if (m < M && n < N) out[slab + n + size_t(m) * N] = tile[threadIdx.x][j];
    }
}

// std::complex<double> travels through the kernel as double2 (same size and alignment)
void transpose(cudaStream_t s, std::complex<double> const *in, std::complex<double> *out, int M, int N, size_t K) {
    dim3 grid((M + TILE - 1) / TILE, (N + TILE - 1) / TILE, unsigned(K)), block(TILE, 8);
    transpose_tiles<double2><<<grid, block, 0, s>>>(reinterpret_cast<double2 const *>(in), reinterpret_cast<double2 *>(out), M, N);
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

#define CUDA_OK(x)                                                                                 \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) {                                                                   \
            std::fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__);  \
            std::exit(2);                                                                          \
        }                                                                                          \
    } while (0)

int run(cudaStream_t stream, size_t M, size_t N, size_t K) {
    using T = double;
    using cplx = std::complex<T>;
    if (K == 0) K = std::max<size_t>(1, size_t(512) * 1000 * 1000 / (sizeof(cplx) * M * N));
    if (K > 65535) K = 65535; // grid.z of the transpose
    std::printf("%zu x %zu x %zu\n", M, N, K);
    const size_t size = M * N * K;
    std::vector<cplx> host(size), got_tft(size), got_db(size);
    for (size_t k = 0; k < K; ++k)
        for (size_t n = 0; n < N; ++n)
            for (size_t m = 0; m < M; ++m) host[m + n * M + k * M * N] = cplx(m / T(M) + n / T(N) + k / T(K), 0);
    cplx *x, *xt, *X;
    CUDA_OK(cudaMalloc(&x, size * sizeof(cplx)));
    CUDA_OK(cudaMalloc(&xt, size * sizeof(cplx)));
    CUDA_OK(cudaMalloc(&X, size * sizeof(cplx)));
    auto init = [&] { CUDA_OK(cudaMemcpy(x, host.data(), size * sizeof(cplx), cudaMemcpyHostToDevice)); };

    // (a) unit-stride plan on the transposed tensor: one row of N per (m, k)
    bbfft::configuration cfg_unit = {1, {1, N, M * K}, bbfft::precision::f64, bbfft::direction::forward,
                                     bbfft::transform_type::c2c};
    auto plan_unit = bbfft::make_plan(cfg_unit, stream);
    // (b) the double-batched plan on the tensor as it is
    bbfft::configuration cfg_db = {1, {M, N, K}, bbfft::precision::f64, bbfft::direction::forward,
                                   bbfft::transform_type::c2c};
    auto plan_db = bbfft::make_plan(cfg_db, stream);

    // (c) the same plan with a store callback that writes element (m, n, k) to n + m N + k M N:
    // FFT + transpose in one kernel (the callback is compiled into the kernel by NVRTC)
    char cb_src[512];
    std::snprintf(cb_src, sizeof(cb_src),
                  "__device__ void store_t(double2* out, size_t offset, double2 value) {\n"
                  "    size_t m = offset %% %zuull, n = offset / %zuull %% %zuull, k = offset / %zuull;\n"
                  "    out[n + m * %zuull + k * %zuull] = value;\n}\n",
                  M, M, N, M * N, N, M * N);
    bbfft::configuration cfg_fused = cfg_db;
    cfg_fused.callbacks = {cb_src, std::strlen(cb_src), nullptr, "store_t", bbfft::kernel_language::cuda_c};
    auto plan_fused = bbfft::make_plan(cfg_fused, stream);

    auto sync = [&] { CUDA_OK(cudaStreamSynchronize(stream)); };
    auto run_tft = [&] {
        transpose(stream, x, xt, int(M), int(N), K);
        plan_unit.execute(xt);
        transpose(stream, xt, X, int(N), int(M), K);
        sync();
    };
    auto run_db = [&] {
        plan_db.execute(x, xt);
        sync();
    };
    init();
    run_tft();
    CUDA_OK(cudaMemcpy(got_tft.data(), X, size * sizeof(cplx), cudaMemcpyDeviceToHost));
    run_db();
    CUDA_OK(cudaMemcpy(got_db.data(), xt, size * sizeof(cplx), cudaMemcpyDeviceToHost));
    // both routes run the same transform: they agree to rounding (different factorizations of N)
    double worst = 0, scale = 0;
    for (size_t i = 0; i < size; ++i) {
        worst = std::max(worst, double(std::abs(got_tft[i] - got_db[i])));
        scale = std::max(scale, double(std::abs(got_db[i])));
    }
    if (worst > 1e2 * std::numeric_limits<T>::epsilon() * scale) {
        std::printf("Error: routes differ by %g (scale %g)\n", worst, scale);
        return 1;
    }
    // (c) against (b) + transpose: same bits (the callback only redirects the store)
    std::vector<cplx> got_fused(size), got_bt(size);
    init();
    plan_fused.execute(x, X);
    sync();
    CUDA_OK(cudaMemcpy(got_fused.data(), X, size * sizeof(cplx), cudaMemcpyDeviceToHost));
    plan_db.execute(x, xt);
    transpose(stream, xt, X, int(M), int(N), K);
    sync();
    CUDA_OK(cudaMemcpy(got_bt.data(), X, size * sizeof(cplx), cudaMemcpyDeviceToHost));
    if (std::memcmp(got_fused.data(), got_bt.data(), size * sizeof(cplx)) != 0) {
        std::printf("Error: fused FFT+transpose differs from FFT followed by transpose\n");
        return 1;
    }
    const double tfused = best_of(10, [&] { plan_fused.execute(x, X); sync(); });
    const double tbt = best_of(10, [&] { plan_db.execute(x, xt); transpose(stream, xt, X, int(M), int(N), K); sync(); });
    const double t1 = best_of(10, [&] { transpose(stream, x, xt, int(M), int(N), K); sync(); });
    const double tu = best_of(10, [&] { plan_unit.execute(xt); sync(); });
    const double t2 = best_of(10, [&] { transpose(stream, xt, X, int(N), int(M), K); sync(); });
    const double ttft = best_of(10, run_tft);
    init();
    const double tdb = best_of(10, run_db);
    auto bw = [&](double t) { return 2.0 * sizeof(cplx) * double(size) / t * 1e-9; };
    std::printf("Transpose 1 %g s, %g GB / s\n", t1, bw(t1));
    std::printf("Unit-stride FFT %g s, %g GB / s\n", tu, bw(tu));
    std::printf("Transpose 2 %g s, %g GB / s\n", t2, bw(t2));
    std::printf("Transpose-FFT-transpose: %g s, %g GB / s\n", ttft, bw(ttft));
    std::printf("Non-unit stride FFT: %g s, %g GB / s\n", tdb, bw(tdb));
    std::printf("Speed-up: %gx\n", ttft / tdb);
    std::printf("FFT then transpose (n-fastest result): %g s, %g GB / s\n", tbt, bw(tbt));
    std::printf("FFT with the transposing store fused in: %g s, %g GB / s (%gx)\n", tfused, bw(tfused), tbt / tfused);
    cudaFree(X);
    cudaFree(xt);
    cudaFree(x);
    return 0;
}

} // namespace

int main(int argc, char **argv) {
    const size_t K = argc >= 2 ? std::strtoull(argv[1], nullptr, 10) : 0;
    cudaStream_t stream;
    CUDA_OK(cudaStreamCreate(&stream));
    int rc = 0;
    rc |= run(stream, 16, 16, K);
    rc |= run(stream, 1120, 32, K);
    rc |= run(stream, 128, 512, K);
    rc |= run(stream, 70, 16, K);
    cudaStreamDestroy(stream);
    return rc;
}
