/* bbfft_cuda.h -- C ABI of the B200 (sm_100a) backend of the Double-Batched FFT Library.
 *
 * This is the drop-in boundary for the batched small-FFT hot path: plain pointers and sizes,
 * no C++ or torch types.  The reference has no C ABI (its boundary is the C++ `Api` policy and
 * the virtual plan_impl::execute); each entry point below names the reference interface it
 * stands in for.  File:line citations are into intel/double-batched-fft-library v0.5.1.
 *
 * Status codes: 0 success, 1 error, 2 bad configuration (bbfft::bad_configuration),
 * 3 CUDA / NVRTC failure.  bbfft_cuda_last_error() returns the message of the last failure on
 * the calling thread.
 */
#ifndef BBFFT_CUDA_H
#define BBFFT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BBFFT_CUDA_MAX_TENSOR_DIM 5

enum { BBFFT_CUDA_OK = 0, BBFFT_CUDA_ERROR = 1, BBFFT_CUDA_BAD_CONFIGURATION = 2, BBFFT_CUDA_DEVICE_ERROR = 3 };

/* POD mirror of bbfft::configuration (include/bbfft/configuration.hpp:151-192).
 * shape = {M, N_1, ..., N_dim, K}; strides in elements of the tensor's own type; fp = 4 | 8
 * (bbfft::precision); dir = -1 | +1 (bbfft::direction); type = 0 c2c, 1 r2c, 2 c2r
 * (bbfft::transform_type).  Callback fields mirror bbfft::user_module
 * (include/bbfft/user_module.hpp:24-37); cb_language 0 = OpenCL-C subset, 1 = CUDA C++. */
typedef struct bbfft_cuda_config {
    unsigned dim;
    size_t shape[BBFFT_CUDA_MAX_TENSOR_DIM];
    int fp;
    int dir;
    int type;
    size_t istride[BBFFT_CUDA_MAX_TENSOR_DIM];
    size_t ostride[BBFFT_CUDA_MAX_TENSOR_DIM];
    const char *cb_source;
    size_t cb_length;
    const char *cb_load;
    const char *cb_store;
    int cb_language;
} bbfft_cuda_config;

typedef struct bbfft_cuda_plan_s *bbfft_cuda_plan_t;
typedef struct bbfft_cuda_cache_s *bbfft_cuda_cache_t;

const char *bbfft_cuda_last_error(void);

/* bbfft::default_istride / default_ostride (src/base/configuration.cpp:24-56). */
int bbfft_cuda_default_strides(const bbfft_cuda_config *cfg, int inplace, size_t *istride, size_t *ostride);
/* bbfft::parse_fft_descriptor (src/base/parser.cpp:59-211) and configuration::to_string
 * (src/base/configuration.cpp:69-160). */
int bbfft_cuda_parse_descriptor(const char *descriptor, bbfft_cuda_config *cfg);
int bbfft_cuda_to_descriptor(const bbfft_cuda_config *cfg, char *buffer, size_t buffer_size);

/* bbfft::jit_cache_all (include/bbfft/jit_cache_all.hpp:20-37): share compiled kernels between
 * plans.  A NULL cache is allowed everywhere. */
int bbfft_cuda_cache_create(bbfft_cuda_cache_t *cache);
int bbfft_cuda_cache_destroy(bbfft_cuda_cache_t cache);
int bbfft_cuda_cache_size(bbfft_cuda_cache_t cache);

/* bbfft::make_plan(cfg, queue, cache) (include/bbfft/sycl/make_plan.hpp:27-41, src/sycl/plan.cpp:16-24).
 * `stream` is a cudaStream_t (NULL = default stream); device < 0 = current device. */
int bbfft_cuda_plan_create(bbfft_cuda_plan_t *plan, const bbfft_cuda_config *cfg, void *stream, int device,
                           bbfft_cuda_cache_t cache);
/* Same, with planner overrides ("R=8x8,T=8,BH=2,...,X2=1" for 1d plans, "RA=8x8,RB=8x8,TH=256,MB=2,PADK=8,SG=-1,BK=1"
 * for 2d configurations -- c2c, r2c or c2r -- that run as one fused tile kernel; used by the auto-tuner and the
 * A/B tools.  X2: packed fp32 adds; SG / BK: staging buffer of the persistent tile kernel and its bulk copies). */
int bbfft_cuda_plan_create_tuned(bbfft_cuda_plan_t *plan, const bbfft_cuda_config *cfg, void *stream,
                                 int device, bbfft_cuda_cache_t cache, const char *tune);
/* plan::execute(in, out) (include/bbfft/plan.hpp:77-131): asynchronous, stream ordered;
 * in == out selects the in-place transform.  Pointers are device (or managed / pinned) memory. */
int bbfft_cuda_plan_execute(bbfft_cuda_plan_t plan, const void *in, void *out);
int bbfft_cuda_plan_execute_on(bbfft_cuda_plan_t plan, const void *in, void *out, void *stream);
/* Host buffers (the reference's plans accept host USM, docs/manual/plans.rst:113-115): copy in, transform,
 * copy out, synchronise.  k slabs travel H2D -> kernel -> D2H over a per-device ring of device slots on
 * three streams, so both PCIe directions stay busy; layouts whose k slices are not contiguous byte ranges
 * and 2d/3d plans are copied whole.  `in_bytes` / `out_bytes` must cover the plan's tensors
 * (BBFFT_CUDA_BAD_CONFIGURATION otherwise).  host_in == host_out runs in place.  Thread safe. */
int bbfft_cuda_plan_execute_host(bbfft_cuda_plan_t plan, const void *host_in, size_t in_bytes, void *host_out,
                                 size_t out_bytes);
int bbfft_cuda_plan_destroy(bbfft_cuda_plan_t plan);
/* Introspection: kernels launched per execute and their identifiers (cache keys,
 * reference: src/base/generator/small_batch_fft.cpp:60-80). */
int bbfft_cuda_plan_num_kernels(bbfft_cuda_plan_t plan);
/* kernel launches issued by one execute (2d/3d plans run their passes once per L2 block of k) */
int bbfft_cuda_plan_launches(bbfft_cuda_plan_t plan);
const char *bbfft_cuda_plan_kernel_name(bbfft_cuda_plan_t plan, int index);

/* Device-free planning and code generation (bbfft::generate_fft_kernels,
 * include/bbfft/generator.hpp:22-36, src/base/generator.cpp:16-25). */
typedef struct bbfft_cuda_kernel_desc {
    char *identifier;
    char *source;          /* CUDA C++ stub; #includes "bbfft_kernels.cuh" */
    double *twiddle;       /* interleaved re, im */
    size_t twiddle_len;    /* number of doubles */
    uint64_t grid;
    int threads;
    size_t smem_bytes;
    int inplace_unsupported;
    int fp;
    int n_stages;
    int radix[4];
    int threads_per_transform;
    int batch_lanes;
    int batch_high;
    int load_staged;
    int store_staged;
} bbfft_cuda_kernel_desc;
int bbfft_cuda_describe(const bbfft_cuda_config *cfg, const char *tune, bbfft_cuda_kernel_desc *desc);
void bbfft_cuda_desc_free(bbfft_cuda_kernel_desc *desc);

/* Device-free planning of a 2d/3d configuration as ONE persistent "chain" kernel (all steps of the
 * reference's nd_fft, src/common/algorithm/nd_fft.hpp:140-152, in one launch with L2-resident
 * intermediates).  Fails with status 2 when the steps cannot be chained. */
typedef struct bbfft_cuda_chain_desc {
    char *identifier;
    char *source;
    double *twiddle;     /* all steps, interleaved re, im */
    size_t twiddle_len;  /* number of doubles */
    int threads;
    size_t smem_bytes;
    int min_blocks;
    int fp;
    int n_steps;
    int step_tile[3];      /* 1 = fused 2d tile step, 0 = double-batched 1d pass */
    uint64_t per_k[3];     /* CTAs of the step per outer k */
    uint64_t mult[3];      /* k slices (tiles) of the step per outer k */
    uint64_t step_M[3];
    int tw_offset[3];      /* first complex element of the step's twiddles */
    int uses_tmp;          /* the plan routes intermediates through a temporary (c2r) */
} bbfft_cuda_chain_desc;
int bbfft_cuda_describe_chain(const bbfft_cuda_config *cfg, bbfft_cuda_chain_desc *desc);
void bbfft_cuda_chain_desc_free(bbfft_cuda_chain_desc *desc);
/* All kernels (1d..3d) of a list of configurations as one translation unit; *source is
 * malloc'ed, *names is a malloc'ed '\n'-separated list. */
int bbfft_cuda_generate_kernels(const bbfft_cuda_config *cfgs, size_t n, char **source, char **names);
/* Text of the device header the stubs include (for offline nvcc builds and the CPU emulator). */
const char *bbfft_cuda_kernel_header(void);
/* NVRTC: CUDA C++ -> cubin for `arch` (e.g. "sm_100a").  Needs no GPU.  *binary is malloc'ed. */
int bbfft_cuda_compile(const char *source, const char *arch, uint8_t **binary, size_t *binary_size);
void bbfft_cuda_free(void *ptr);

#ifdef __cplusplus
}
#endif

#endif /* BBFFT_CUDA_H */
