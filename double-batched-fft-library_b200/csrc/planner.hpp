// planner.hpp -- device-independent planning for the sm_100a batched small-FFT kernels.
//
// Replaces the reference's configure_* / generate_* pair
// (reference: include/bbfft/detail/generator_impl.hpp:26-107,
//  src/base/generator/small_batch_fft.cpp:15-110, src/base/generator/factor2_slm_fft.cpp:16-116):
// given a 1d bbfft configuration it chooses the stage radices, the thread layout, the shared
// memory layout, builds the twiddle table and emits the NVRTC/nvcc "stub" that instantiates
// bbk::fft1d<C> from kernels/bbfft_kernels.cuh.
#ifndef BBFFT_CUDA_PLANNER_HPP
#define BBFFT_CUDA_PLANNER_HPP

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace bbfft::cuda {

enum kernel_mode : int { k_c2c = 0, k_r2c_half = 1, k_c2r_half = 2, k_r2c_double = 3, k_c2r_double = 4 };

struct problem_1d {
    int fp = 4;            // bytes per real: 4 | 8   (bbfft::precision)
    int dir = -1;          // -1 forward, +1 backward (bbfft::direction)
    int type = 0;          // 0 c2c, 1 r2c, 2 c2r     (bbfft::transform_type)
    std::uint64_t M = 1, N = 1, K = 1;
    std::int64_t is1 = 1, is2 = 1, os1 = 1, os2 = 1; // strides[1], strides[2]
    std::string cb_source, cb_load, cb_store;       // CUDA C callbacks (may be empty)
};

struct device_props {
    int sm_count = 148;
    int max_threads_per_block = 1024;
    std::size_t max_smem_per_block = 227 * 1024;
    std::size_t smem_per_sm = 228 * 1024;
    int regs_per_sm = 65536;
    int cc_major = 10, cc_minor = 0;
};

struct kernel_params {
    int fp = 4, dir = -1, mode = k_c2c;
    int N = 1;     // length of the complex FFT the stages run
    int nreal = 1; // user-visible N
    int L = 1;
    int radix[4] = {1, 1, 1, 1};
    int T = 1, ML = 1, BH = 1;
    bool klanes = false, load_staged = false, store_staged = false;
    bool pair_load = false, pair_store = false; // real side is read/written as aligned complex words
    bool real_fused = false; // real pre/post pass fused into the first/last stage (mirrored sub-FFT pairs)
    bool x2 = false;         // fp32: packed adds (add.f32x2 -> FADD2) in this kernel's translation unit (tune X2=1)
    bool chained = false;    // stub is a step of a chain kernel: no entry point, L2-only (ld.global.cg) loads
    std::uint64_t M = 1;
    std::int64_t is1 = 1, is2 = 1, os1 = 1, os2 = 1;
    int LL = 1, PADK = 0, ROW = 1;
    std::size_t smem_bytes = 0;
    int threads = 1;
    int min_blocks = 1; // resident CTAs per SM the register cap is sized for
    int max_regs = 255;  // __maxnreg__ of the kernel (reg_cap(threads, min_blocks))
    std::string cb_load, cb_store;

    int batch_per_cta() const { return ML * BH; }
    // number of CTAs for K slices
    std::uint64_t grid(std::uint64_t K) const;
    // k slices consumed per CTA (for odd-N real transforms a slice pair counts as one unit)
    std::uint64_t k_per_cta() const;
};

// register cap for `blocks` resident CTAs of `threads` threads (four register-file partitions)
int reg_cap(int threads, int blocks);

struct kernel_plan {
    kernel_params p;
    std::string identifier;      // cache key / kernel name
    std::string source;          // stub (includes "bbfft_kernels.cuh")
    std::vector<double> twiddle; // interleaved re,im in double; narrowed by the caller
    bool inplace_unsupported = false;
};

// Tuning overrides, "key=value,key=value" (keys: R=8x8, T, ML, BH, MB, LD, ST, RF, PADK, ROW, KL, CH).
// Used by the auto-tuner and the tests; empty string = heuristics.
kernel_plan plan_kernel_1d(problem_1d const &prob, device_props const &dev,
                           std::string const &tune = std::string());

std::string emit_stub(kernel_params const &p, std::string const &identifier,
                      std::string const &cb_source);
std::string make_identifier(kernel_params const &p);
std::vector<double> make_twiddles(kernel_params const &p);

// ---- fused 2d tile kernel (bbk::fft2d_tile): c2c, default packed layout, one CTA per tile ----
struct problem_2d {
    int fp = 4, dir = -1;
    std::uint64_t M = 1, N1 = 1, N2 = 1, K = 1; // K = number of tiles
    std::uint64_t tile_stride = 0;              // elements between consecutive tiles (0 = packed)
    // fused real tiles (bbk::fft2d_tile_real_cta): 0 = c2c, 1 = r2c, 2 = c2r; N1 is the real length (even), the
    // spectrum tile is M x (N1/2+1) x N2; `inplace` selects the padded real rows of the in-place default layout
    int real = 0;
    bool inplace = false;
};

struct tile_pass_params {
    int N = 1, S = 1, O = 1, L = 1;
    int radix[4] = {1, 1, 1, 1};
    int pitch = 0; // elements between the O blocks of the pass inside the tile (S * N unless rows are padded)
};

struct tile_params {
    int fp = 4, dir = -1;
    std::uint64_t M = 1, N1 = 1, N2 = 1, tile_stride = 0;
    tile_pass_params a, b; // axis n1, axis n2
    int threads = 256, PADK = 0, min_blocks = 1, max_regs = 255;
    bool persistent = false; // persistent grid with the asynchronous tile pipeline (bbk::fft2d_tile_persistent)
    int stage = 0;           // leading elements of the NEXT tile prefetched into a staging buffer behind the tile
                             // (bbk::fft2d_tile_staged; whole rows of pass A; implies the persistent grid)
    int stage_off = 0;       // first element of the staging buffer in shared memory
    bool bulk = false;       // staging buffer filled by cp.async.bulk + mbarrier (one thread issues) instead of cp.async
    int cluster = 1;         // CTAs per tile (thread-block cluster, bbk::fft2d_tile_cluster); 1 = one CTA owns the tile
    std::size_t smem_bytes = 0;
    bool chained = false; // see kernel_params::chained
    int real = 0;                          // problem_2d::real
    std::uint64_t real_tile_stride = 0;    // reals between consecutive real tiles
    std::uint64_t real_row = 0;            // reals between consecutive real rows (n2) of a tile
    std::uint64_t spectrum_n1 = 0;         // N1/2 + 1
};

struct tile_plan {
    tile_params p;
    std::string identifier, source;
    std::vector<double> twiddle;
};

// True when the M x N1 x N2 tile (plus padding) fits the shared memory of one CTA and is large
// enough to be worth a CTA of its own.
// largest thread-block cluster a tile may be split over (BBFFT_CUDA_TILE_CLUSTER, default 1 = off)
int tile_cluster_limit();
// (max_cluster = 1: the tile must fit one CTA)
bool tile_fusable(problem_2d const &prob, device_props const &dev, int max_cluster = 1);
// Tuning overrides: RA=8x16, RB=16x8, TH=<threads>, PADK, MB, PS, CL, SG=<rows of pass A staged ahead, -1 = as many as fit>, BK=<0|1> bulk copies.
tile_plan plan_kernel_2d(problem_2d const &prob, device_props const &dev,
                         std::string const &tune = std::string());

// ---- chained nd kernel (bbk::chain): every step of a 2d/3d decomposition in ONE persistent launch ----
struct chain_step_problem {
    bool tile = false;
    problem_2d t;           // tile == true
    problem_1d p;           // tile == false
    std::uint64_t mult = 1; // slices of the step per outer k (slab)
};
struct chain_step_plan {
    bool tile = false;
    tile_plan tp;
    kernel_plan kp;
    std::uint64_t per_k = 1; // CTAs of the step's grid per slab
    int tw_offset = 0;       // first complex element of the step's twiddles in the chain's table
};
struct chain_plan_t {
    std::vector<chain_step_plan> steps;
    std::string identifier, source;
    std::string entry_source;    // the part of `source` after the steps' stubs (chain traits + entry point)
    std::vector<double> twiddle; // all steps, interleaved re,im
    int threads = 0, min_blocks = 1, max_regs = 255;
    std::size_t smem_bytes = 0;
};
// Returns false when the steps cannot share one CTA shape (the caller then launches them one by one).
bool plan_chain(std::vector<chain_step_problem> const &steps, device_props const &dev, chain_plan_t &out);

// A radix that runs as a cooperative direct DFT from shared memory (bbk::run_stage_direct): a prime
// too large for an in-register butterfly.
bool direct_radix(int r);

// integer helpers (pinned by tests; semantics of reference src/base/prime_factorization.cpp)
std::vector<int> prime_factors(int n);
bool radix_supported(int r);

} // namespace bbfft::cuda

#endif
