/* oracle/bbfft_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into or called by the product
 * library; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Plain-C CPU restatement of the reference's batched small-FFT path, in two independent legs:
 *
 *  (1) oracle_dft        direct O(N^2) DFT evaluated in long double over the exact bbfft tensor
 *                        layout (M x N_1..N_d x K, column-major, arbitrary strides; c2c / r2c / c2r;
 *                        reference include/bbfft/configuration.hpp:151-192, docs/manual/plans.rst:12-19).
 *                        This is the numerical ground truth the tolerances are measured against.
 *
 *  (2) oracle_bbfft      "reference-structured" restatement computed in the working precision:
 *                        algorithm selection      src/common/algorithm_1d.hpp:19-36
 *                        register model           src/base/device_info.cpp:29-57
 *                        balanced factorization   src/base/prime_factorization.cpp:27-77
 *                        f2fft configuration      src/base/generator/factor2_slm_fft.cpp:16-41
 *                        in-register DIF FFT      src/base/mixed_radix_fft.cpp:215-254
 *                        (un)scrambler            src/base/scrambler.hpp:37-49,80-93
 *                        sbfft data flow          src/base/generator/sbfft_gen.cpp:150-351
 *                        f2fft data flow          src/base/generator/f2fft_gen.cpp:95-176,203-226,253-271,349-371
 *                        table twiddles           src/common/algorithm/factor2_slm_fft.hpp:49-90
 *                        nd chaining              src/common/algorithm/nd_fft.hpp:61-99,140-152
 *
 * Pinned against: the reference's own host-side golden vectors (test/codegen.cpp:15-125: factor,
 * scrambler), the reference's analytic device tests (test/c2c.cpp, test/r2c.cpp) and the outputs of
 * the UNMODIFIED reference run under oracle/_ref (see tests/test_oracle.py, tests/golden/).
 */
#include <complex.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    unsigned dim;
    size_t shape[5];
    int fp;   /* 4 | 8 */
    int dir;  /* -1 forward | +1 backward */
    int type; /* 0 c2c | 1 r2c | 2 c2r */
    size_t istride[5];
    size_t ostride[5];
} oracle_config;

typedef struct {
    size_t max_work_group_size;
    size_t min_subgroup_size;
    size_t max_subgroup_size;
    size_t local_memory_size;
    int is_cpu;
} oracle_device;

enum { T_C2C = 0, T_R2C = 1, T_C2R = 2 };

/* ------------------------------------------------------------------------------------------- */
/* layout helpers (reference src/base/configuration.cpp:24-56)                                  */
/* ------------------------------------------------------------------------------------------- */
static void default_stride_impl(unsigned dim, const size_t *shape, int type, int inplace,
                                size_t *stride) {
    size_t shape1 = shape[1];
    if (type == T_R2C) {
        shape1 = inplace ? 2 * (shape[1] / 2 + 1) : shape[1];
    } else if (type == T_C2R) {
        shape1 = shape[1] / 2 + 1;
    }
    for (int i = 0; i < 5; ++i) {
        stride[i] = 0;
    }
    stride[0] = 1;
    stride[1] = shape[0];
    stride[2] = shape1 * shape[0];
    for (unsigned d = 1; d < dim; ++d) {
        stride[d + 2] = shape[d + 1] * stride[d + 1];
    }
}

void oracle_default_strides(const oracle_config *c, int inplace, size_t *is, size_t *os) {
    default_stride_impl(c->dim, c->shape, c->type, inplace, is);
    int otype = c->type == T_R2C ? T_C2R : (c->type == T_C2R ? T_R2C : T_C2C);
    default_stride_impl(c->dim, c->shape, otype, inplace, os);
}

/* ------------------------------------------------------------------------------------------- */
/* integer helpers (reference src/base/math.cpp, prime_factorization.cpp, scrambler.hpp)       */
/* ------------------------------------------------------------------------------------------- */
int oracle_trial_division(int n, int *factors) {
    int cnt = 0, f = 2;
    while (n > 1) {
        if (n % f == 0) {
            factors[cnt++] = f;
            n /= f;
        } else {
            ++f;
        }
    }
    return cnt;
}

static unsigned ipow_u(unsigned b, unsigned e) {
    unsigned r = 1;
    while (e--) {
        r *= b;
    }
    return r;
}
/* largest r with r^index <= n */
static unsigned iroot_u(unsigned n, unsigned index) {
    unsigned r = (unsigned)floor(pow((double)n, 1.0 / index));
    while (ipow_u(r + 1, index) <= n) {
        ++r;
    }
    while (r > 0 && ipow_u(r, index) > n) {
        --r;
    }
    return r;
}
static int is_prime_u(unsigned n) {
    if (n < 2) {
        return 0;
    }
    for (unsigned f = 2; f * f <= n; ++f) {
        if (n % f == 0) {
            return 0;
        }
    }
    return 1;
}

/* Best factorization n = f_0 ... f_{index-1} in the sense of min sum (target - f_i)^2, where the
 * leading factor of every sub-problem is searched downward from iroot(n, index) and the first
 * minimum wins -- same search order, error accumulation order and tie-breaking as the reference's
 * update_factor (src/base/prime_factorization.cpp:27-62).                                       */
static double factor_rec(unsigned n, unsigned index, double target, unsigned *factors, int *found) {
    if (index == 1) {
        double d = target - n;
        factors[0] = n;
        *found = 1;
        return d * d;
    }
    unsigned r = iroot_u(n, index);
    double best = 1.7976931348623157e308;
    unsigned sub[8];
    *found = 0;
    for (unsigned f0 = r; f0 > 0; --f0) {
        if (n % f0 == 0) {
            int fnd = 0;
            double err = factor_rec(n / f0, index - 1, target, sub, &fnd);
            if (fnd) {
                double d = target - f0;
                err += d * d;
                if (err < best) {
                    best = err;
                    factors[0] = f0;
                    memcpy(factors + 1, sub, (index - 1) * sizeof(unsigned));
                    *found = 1;
                }
            }
        }
    }
    return best;
}

static int cmp_unsigned(const void *a, const void *b) {
    unsigned x = *(const unsigned *)a, y = *(const unsigned *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

void oracle_factor(unsigned n, unsigned index, unsigned *factors) {
    if (index == 0) {
        return;
    }
    if (n == 0) {
        for (unsigned i = 0; i < index; ++i) {
            factors[i] = 0;
        }
        return;
    }
    double target = pow((double)n, 1.0 / index);
    int found = 0;
    factor_rec(n, index, target, factors, &found);
    qsort(factors, index, sizeof(unsigned), cmp_unsigned);
}

long oracle_scramble(long index, const int *factors, int L) {
    long result = 0, N = 1;
    for (int i = 0; i < L; ++i) {
        result = result * factors[i] + index % factors[i];
        index /= factors[i];
        N *= factors[i];
    }
    return result + index * N;
}
long oracle_unscramble(long index, const int *factors, int L) {
    long result = 0, N = 1;
    for (int i = L - 1; i >= 0; --i) {
        result = result * factors[i] + index % factors[i];
        index /= factors[i];
        N *= factors[i];
    }
    return result + index * N;
}

/* register-space model, reference src/base/device_info.cpp:29-57 */
static size_t register_space(const oracle_device *dev) {
    if (dev->is_cpu) {
        return 32 * 64;
    }
    size_t scale = dev->min_subgroup_size / 8u;
    if (scale < 1) {
        scale = 1;
    }
    return scale * 32u * 128u;
}

/* returns 0 = small batch (register) path, 1 = factor2 slm path; fills factorization for path 1.
 * reference src/common/algorithm_1d.hpp:19-36 + src/base/generator/factor2_slm_fft.cpp:16-41    */
int oracle_select_1d(const oracle_config *c, const oracle_device *dev, unsigned *factorization,
                     int *num_factors) {
    size_t N = c->shape[1];
    size_t sgs = dev->min_subgroup_size;
    size_t reg_space = register_space(dev);
    size_t need = 2 * (size_t)c->fp * N * sgs;
    if (c->type != T_C2C && N % 2 == 0) {
        need /= 2;
    }
    *num_factors = 0;
    if (need < reg_space / 2) {
        return 0;
    }
    size_t N_fft = N;
    if (c->type != T_C2C && N % 2 == 0) {
        N_fft /= 2;
    }
    unsigned max_in_reg = (unsigned)((reg_space / 2) / (2 * (size_t)c->fp) / sgs);
    for (unsigned index = 2; index <= 4; ++index) {
        oracle_factor((unsigned)N_fft, index, factorization);
        *num_factors = (int)index;
        unsigned fmax = 0;
        for (unsigned i = 0; i < index; ++i) {
            if (factorization[i] > fmax) {
                fmax = factorization[i];
            }
        }
        if (fmax < max_in_reg || is_prime_u(fmax)) {
            break;
        }
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------- */
/* (1) long double direct DFT                                                                  */
/* ------------------------------------------------------------------------------------------- */
typedef long double complex lcplx;

static void dft_axis(lcplx *w, size_t inner, size_t n, size_t outer, int dir) {
    /* w viewed as [outer][n][inner]; transform over the middle index */
    const long double tau = 6.283185307179586476925286766559005768L;
    lcplx *tw = (lcplx *)malloc(n * sizeof(lcplx));
    lcplx *tmp = (lcplx *)malloc(n * sizeof(lcplx));
    for (size_t j = 0; j < n; ++j) {
        long double a = dir * tau * (long double)j / (long double)n;
        tw[j] = cosl(a) + I * sinl(a);
    }
    for (size_t o = 0; o < outer; ++o) {
        for (size_t i = 0; i < inner; ++i) {
            lcplx *base = w + o * n * inner + i;
            for (size_t k = 0; k < n; ++k) {
                lcplx s = 0;
                for (size_t j = 0; j < n; ++j) {
                    s += base[j * inner] * tw[(j * k) % n];
                }
                tmp[k] = s;
            }
            for (size_t k = 0; k < n; ++k) {
                base[k * inner] = tmp[k];
            }
        }
    }
    free(tw);
    free(tmp);
}

static long double load_real(const void *p, size_t idx, int fp) {
    return fp == 4 ? (long double)((const float *)p)[idx] : (long double)((const double *)p)[idx];
}
static void store_real(void *p, size_t idx, int fp, long double v) {
    if (fp == 4) {
        ((float *)p)[idx] = (float)v;
    } else {
        ((double *)p)[idx] = (double)v;
    }
}

int oracle_dft(const oracle_config *c, const void *in, void *out) {
    unsigned d = c->dim;
    if (d < 1 || d > 3) {
        return 2;
    }
    size_t M = c->shape[0], K = c->shape[d + 1];
    size_t Nn[3] = {1, 1, 1};
    for (unsigned i = 0; i < d; ++i) {
        Nn[i] = c->shape[i + 1];
    }
    size_t N1 = Nn[0], N1h = N1 / 2 + 1;
    size_t Nin1 = c->type == T_C2R ? N1h : N1;   /* stored extent of axis 1 on input */
    size_t total = M * N1 * Nn[1] * Nn[2] * K;   /* full complex working tensor       */
    lcplx *w = (lcplx *)calloc(total, sizeof(lcplx));
    if (!w) {
        return 1;
    }
#define WIDX(m, n1, n2, n3, k) ((((k) * Nn[2] + (n3)) * Nn[1] + (n2)) * N1 + (n1)) * M + (m)
    /* gather */
    for (size_t k = 0; k < K; ++k)
        for (size_t n3 = 0; n3 < Nn[2]; ++n3)
            for (size_t n2 = 0; n2 < Nn[1]; ++n2)
                for (size_t n1 = 0; n1 < Nin1; ++n1)
                    for (size_t m = 0; m < M; ++m) {
                        size_t off = m * c->istride[0] + n1 * c->istride[1] + k * c->istride[d + 1];
                        if (d >= 2) {
                            off += n2 * c->istride[2];
                        }
                        if (d >= 3) {
                            off += n3 * c->istride[3];
                        }
                        lcplx v;
                        if (c->type == T_R2C) {
                            v = load_real(in, off, c->fp);
                        } else {
                            v = load_real(in, 2 * off, c->fp) + I * load_real(in, 2 * off + 1, c->fp);
                        }
                        w[WIDX(m, n1, n2, n3, k)] = v;
                    }
    if (c->type == T_C2R) {
        /* higher axes first (complex), then the Hermitian axis                                */
        /* working tensor currently holds n1 < N1h; transform axes 2,3 on those rows           */
        if (d >= 3) {
            dft_axis(w, M * N1 * Nn[1], Nn[2], K, c->dir);
        }
        if (d >= 2) {
            dft_axis(w, M * N1, Nn[1], Nn[2] * K, c->dir);
        }
        /* Hermitian extension along axis 1; imag of bin 0 (and Nyquist) ignored:
         * reference test/r2c.cpp:310-324, sbfft_gen.cpp:230-233                                */
        for (size_t o = 0; o < Nn[1] * Nn[2] * K; ++o)
            for (size_t m = 0; m < M; ++m) {
                lcplx *row = w + o * N1 * M + m;
                row[0] = creall(row[0]);
                if (N1 % 2 == 0) {
                    row[(N1 / 2) * M] = creall(row[(N1 / 2) * M]);
                }
                for (size_t n1 = N1h; n1 < N1; ++n1) {
                    row[n1 * M] = conjl(row[(N1 - n1) * M]);
                }
            }
        dft_axis(w, M, N1, Nn[1] * Nn[2] * K, c->dir);
    } else {
        dft_axis(w, M, N1, Nn[1] * Nn[2] * K, c->dir);
        if (d >= 2) {
            dft_axis(w, M * N1, Nn[1], Nn[2] * K, c->dir);
        }
        if (d >= 3) {
            dft_axis(w, M * N1 * Nn[1], Nn[2], K, c->dir);
        }
    }
    /* scatter */
    size_t Nout1 = c->type == T_R2C ? N1h : N1;
    for (size_t k = 0; k < K; ++k)
        for (size_t n3 = 0; n3 < Nn[2]; ++n3)
            for (size_t n2 = 0; n2 < Nn[1]; ++n2)
                for (size_t n1 = 0; n1 < Nout1; ++n1)
                    for (size_t m = 0; m < M; ++m) {
                        size_t off = m * c->ostride[0] + n1 * c->ostride[1] + k * c->ostride[d + 1];
                        if (d >= 2) {
                            off += n2 * c->ostride[2];
                        }
                        if (d >= 3) {
                            off += n3 * c->ostride[3];
                        }
                        lcplx v = w[WIDX(m, n1, n2, n3, k)];
                        if (c->type == T_C2R) {
                            store_real(out, off, c->fp, creall(v));
                        } else {
                            store_real(out, 2 * off, c->fp, creall(v));
                            store_real(out, 2 * off + 1, c->fp, cimagl(v));
                        }
                    }
#undef WIDX
    free(w);
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* (2) reference-structured restatement, instantiated for float and double                     */
/* ------------------------------------------------------------------------------------------- */
#define REAL float
#define SUF(x) x##_f32
#include "bbfft_oracle_impl.inc"
#undef REAL
#undef SUF
#define REAL double
#define SUF(x) x##_f64
#include "bbfft_oracle_impl.inc"
#undef REAL
#undef SUF

static const oracle_device pvc = {1024, 16, 32, 131072, 0}; /* reference tools/common/info.cpp:9 */

static int bbfft_1d(const oracle_config *c, const oracle_device *dev, const void *in, void *out) {
    return c->fp == 4 ? bbfft_1d_f32(c, dev, in, out) : bbfft_1d_f64(c, dev, in, out);
}

/* nd chaining, reference src/common/algorithm/nd_fft.hpp:27-112,140-152 */
int oracle_bbfft(const oracle_config *c, const oracle_device *dev_in, const void *in, void *out) {
    const oracle_device *dev = dev_in ? dev_in : &pvc;
    if (c->dim < 1 || c->dim > 3) {
        return 2;
    }
    if (c->dim == 1) {
        return bbfft_1d(c, dev, in, out);
    }
    unsigned dim = c->dim;
    size_t is_ip[5], os_ip[5], is_op[5], os_op[5];
    oracle_config tmpc = *c;
    oracle_default_strides(&tmpc, 1, is_ip, os_ip);
    oracle_default_strides(&tmpc, 0, is_op, os_op);
    int eq_ip = 1, eq_op = 1;
    for (unsigned i = 0; i < dim + 2; ++i) {
        eq_ip = eq_ip && c->istride[i] == is_ip[i] && c->ostride[i] == os_ip[i];
        eq_op = eq_op && c->istride[i] == is_op[i] && c->ostride[i] == os_op[i];
    }
    if (!eq_ip && !eq_op) {
        return 2;
    }
    int is_real = c->type != T_C2C;
    int inplace_layout = eq_ip;
    oracle_config c1[3];
    size_t Ntot = 1;
    for (unsigned d = 0; d < dim; ++d) {
        Ntot *= (d == 0 && is_real) ? c->shape[1] / 2 + 1 : c->shape[d + 1];
    }
    size_t M = c->shape[0], K = Ntot * c->shape[dim + 1];
    for (unsigned d = 0; d < dim; ++d) {
        size_t Nd = c->shape[d + 1];
        size_t Ndc = (d == 0 && is_real) ? c->shape[1] / 2 + 1 : Nd;
        size_t Ndr = (d == 0 && is_real && inplace_layout) ? 2 * (c->shape[1] / 2 + 1) : Nd;
        K /= Ndc;
        memset(&c1[d], 0, sizeof(oracle_config));
        c1[d].dim = 1;
        c1[d].shape[0] = M;
        c1[d].shape[1] = Nd;
        c1[d].shape[2] = K;
        c1[d].fp = c->fp;
        c1[d].dir = c->dir;
        c1[d].type = d == 0 ? c->type : T_C2C;
        c1[d].istride[0] = 1;
        c1[d].istride[1] = M;
        c1[d].istride[2] = M * Ndr;
        c1[d].ostride[0] = 1;
        c1[d].ostride[1] = M;
        c1[d].ostride[2] = M * Ndc;
        M *= Ndc;
    }
    oracle_config plans[3];
    if (c->type == T_C2R) {
        for (unsigned d = 0; d < dim; ++d) {
            plans[d] = c1[dim - 1 - d];
            size_t t[5];
            memcpy(t, plans[d].istride, sizeof t);
            memcpy(plans[d].istride, plans[d].ostride, sizeof t);
            memcpy(plans[d].ostride, t, sizeof t);
        }
    } else {
        for (unsigned d = 0; d < dim; ++d) {
            plans[d] = c1[d];
        }
    }
    size_t ibytes = (c->type == T_R2C ? 1 : 2) * (size_t)c->fp;
    size_t obytes = (c->type == T_C2R ? 1 : 2) * (size_t)c->fp;
    size_t isize = c->istride[dim + 1] * c->shape[dim + 1] * ibytes;
    size_t osize = c->ostride[dim + 1] * c->shape[dim + 1] * obytes;
    void *tmp = out, *tmp_alloc = NULL;
    if (isize > osize) {
        tmp_alloc = malloc(isize);
        tmp = tmp_alloc;
    }
    int rc = bbfft_1d(&plans[0], dev, in, tmp);
    for (unsigned d = 1; rc == 0 && d + 1 < dim; ++d) {
        rc = bbfft_1d(&plans[d], dev, tmp, tmp);
    }
    if (rc == 0) {
        rc = bbfft_1d(&plans[dim - 1], dev, tmp, out);
    }
    free(tmp_alloc);
    return rc;
}
