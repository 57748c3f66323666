// bbfft_kernels.cuh -- hand-written sm_100a device templates for the double-batched small FFT.
//
// This single header is the whole device side of the library.  It is
//   * compiled by NVRTC at plan creation (the host embeds this text; see runtime.cpp),
//   * compiled by nvcc for the built-in / AOT kernel bundles (same text, same stubs),
//   * compiled by g++ against tests/emu/cuda_emu.hpp (macro BBFFT_EMU) so that every index
//     map can be checked on a CPU-only box.  That emulation is test infrastructure only.
//
// A kernel is instantiated by a small generated "stub" (planner.cpp: emit_stub) that defines a
// traits struct `C` (all compile-time plan parameters) and one
//     extern "C" __global__ void <identifier>(bbk::args a) { bbk::fft1d<C>(a); }
//
// What replaces what in the reference (intel/double-batched-fft-library):
//   reg_fft / bfly        <- src/base/mixed_radix_fft.cpp:215-254 (in-register FFT), :78-126 (pair trick)
//   fft1d, L == 1         <- src/base/generator/sbfft_gen.cpp:30-134   ("small batch" kernel)
//   fft1d, L >= 2         <- src/base/generator/f2fft_gen.cpp:46-183   ("factor2 slm" kernel)
//   real pre/post passes  <- sbfft_gen.cpp:162-351, f2fft_gen.cpp:228-502
//   C::ld / C::st hooks   <- src/base/generator/tensor_accessor.cpp:34-55 (callback_accessor)
//   fft2d_tile, chain     <- src/common/algorithm/nd_fft.hpp:66-152 (2d / 3d as chained 1d launches)
//   run_stage_direct      <- prime lengths, which the reference unrolls into one work-item
// The algorithm is NOT a translation: stages are radix-2..32 Stockham-style passes that work in
// place in shared memory (digit reversal folded into the final store), threads are laid out so
// that the M batch index runs along the lanes of a warp (coalesced 128-byte rows, conflict-free
// shared memory without padding), and twiddles between stages come from a small L1-resident
// table.
//
// Tensor layout (reference include/bbfft/configuration.hpp:162-179): element (m, n, k) of the
// M x N x K tensor lives at offset m + n*s1 + k*s2 counted in elements of the tensor's own type.
#ifndef BBFFT_KERNELS_CUH
#define BBFFT_KERNELS_CUH

#ifdef BBFFT_EMU
#define BBK_DEV inline
#define BBK_HD inline
#define BBK_CE constexpr
#define BBK_RESTRICT __restrict__
#define BBK_GLOBAL
#define BBK_LAUNCH_BOUNDS(t, b)
#define BBK_KERNEL(t, r)
#define BBK_CLUSTER(n)
#else
#define BBK_DEV __device__ __forceinline__
#define BBK_HD __host__ __device__ __forceinline__
#define BBK_CE __host__ __device__ constexpr
#define BBK_GLOBAL __global__
#define BBK_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
// Kernel attribute of a stub: `t` threads per CTA, at most `r` registers per thread (the planner's
// exact cap for the planned number of resident CTAs).  -DBBK_NO_REGCAP: the host rebuilds a JIT
// kernel without the cap when the capped build needs local memory (plan.cpp: jit_module); the
// launch bound keeps the uncapped kernel launchable with `t` threads.
#ifdef BBK_NO_REGCAP
#define BBK_KERNEL(t, r) __launch_bounds__(t)
#else
#define BBK_KERNEL(t, r) __maxnreg__(r)
#endif
#define BBK_CLUSTER(n) __cluster_dims__(n, 1, 1)
#define BBK_RESTRICT __restrict__
#endif

// Shared-memory pointer type.  In product builds it is a plain pointer; the CPU emulator's race
// checker (tests/emu, -DBBFFT_EMU_RACECHECK) substitutes a proxy that records which thread read or
// wrote every word between two barriers.
#if defined(BBFFT_EMU) && defined(BBFFT_EMU_RACECHECK)
#define BBK_SPTR(E) ::bbfft_emu::checked_ptr<E>
#else
#define BBK_SPTR(E) E *
#endif

namespace bbk {

typedef unsigned long long u64;
typedef long long i64;

// transform modes (C::MODE)
enum : int {
    C2C = 0,
    R2C_HALF = 1,   // even N: half-length complex FFT + fused post-twiddle
    C2R_HALF = 2,   // even N: fused pre-twiddle + half-length complex FFT
    R2C_DOUBLE = 3, // odd N: two real rows (k, k+1) through one complex FFT
    C2R_DOUBLE = 4
};

// Kernel arguments.  One signature for every kernel; strides are only read when the stub was
// generated with run-time strides (the default is compile-time strides like the reference,
// whose identifiers carry them: src/base/generator/small_batch_fft.cpp:60-80).
struct args {
    const void *in;
    void *out;
    const void *tw; // inter-stage twiddle table (+ real post/pre table), element type cx<real_t>
    u64 K;          // number of k slices handled by this launch
    u64 M;
    i64 is1, is2, os1, os2;
    u64 pf;         // L2 prefetch distance in CTA-batches of k (tiles for the 2d kernel); 0 = off
};

template <class T> struct alignas(2 * sizeof(T)) cx {
    T x, y;
};

// Every floating-point multiply of the transform is written with an explicit-rounding intrinsic
// (never a bare `*`), so the compiler cannot contract or split multiply-adds differently from
// one instantiation to the next: a plan with user callbacks, the same plan without, the
// NVRTC-compiled and the nvcc-compiled kernel all produce bit-identical spectra
// (the reference's test/callback.cpp:102-104,186-192 compares with ==).
#ifdef BBFFT_EMU
template <class T> BBK_DEV T fmul(T a, T b) { return a * b; }
template <class T> BBK_DEV T ffma(T a, T b, T c) { return a * b + c; }
#else
BBK_DEV float fmul(float a, float b) { return __fmul_rn(a, b); }
BBK_DEV double fmul(double a, double b) { return __dmul_rn(a, b); }
BBK_DEV float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
BBK_DEV double ffma(double a, double b, double c) { return __fma_rn(a, b, c); }
#endif

template <class T> BBK_DEV cx<T> operator+(cx<T> a, cx<T> b) { return cx<T>{a.x + b.x, a.y + b.y}; }
template <class T> BBK_DEV cx<T> operator-(cx<T> a, cx<T> b) { return cx<T>{a.x - b.x, a.y - b.y}; }
#if defined(BBK_F32X2) && !defined(BBFFT_EMU)
// Packed fp32 arithmetic of sm_100 (add.f32x2 -> SASS FADD2): a complex add is ONE issue slot.  Same rounding
// as two scalar adds, so results do not change; what changes is the number of instructions the four
// schedulers of an SM have to issue, which is what bounds the fp32 tile kernels.
BBK_DEV cx<float> operator+(cx<float> a, cx<float> b) {
    cx<float> r;
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
BBK_DEV cx<float> operator-(cx<float> a, cx<float> b) {
    cx<float> r;
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tsub.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
#endif
template <class T> BBK_DEV cx<T> cmul(cx<T> a, cx<T> b) {
    return cx<T>{ffma(a.x, b.x, -fmul(a.y, b.y)), ffma(a.x, b.y, fmul(a.y, b.x))};
}
// real scalar times complex
template <class T> BBK_DEV cx<T> rmul(cx<T> a, T s) { return cx<T>{fmul(a.x, s), fmul(a.y, s)}; }
template <class T> BBK_DEV cx<T> conj(cx<T> a) { return cx<T>{a.x, -a.y}; }
// multiply by DIR*i  (= exp(DIR*i*pi/2))
template <int DIR, class T> BBK_DEV cx<T> mul_i(cx<T> a) {
    return DIR < 0 ? cx<T>{a.y, -a.x} : cx<T>{-a.y, a.x};
}

// read-only (L1-resident) load of a table entry
#ifdef BBFFT_EMU
template <class T> BBK_DEV cx<T> ldg_cx(const cx<T> *p) { return *p; }
#else
BBK_DEV cx<float> ldg_cx(const cx<float> *p) {
    float2 r = __ldg(reinterpret_cast<const float2 *>(p));
    return cx<float>{r.x, r.y};
}
BBK_DEV cx<double> ldg_cx(const cx<double> *p) {
    double2 r = __ldg(reinterpret_cast<const double2 *>(p));
    return cx<double>{r.x, r.y};
}
#endif

// L2-only load of a tensor element: for data another SM wrote during the same launch (chain)
#ifdef BBFFT_EMU
template <class T> BBK_DEV cx<T> ldcg_cx(const cx<T> *p) { return *p; }
#else
BBK_DEV cx<float> ldcg_cx(const cx<float> *p) {
    float2 r = __ldcg(reinterpret_cast<const float2 *>(p));
    return cx<float>{r.x, r.y};
}
BBK_DEV cx<double> ldcg_cx(const cx<double> *p) {
    double2 r = __ldcg(reinterpret_cast<const double2 *>(p));
    return cx<double>{r.x, r.y};
}
#endif

template <class E> BBK_DEV BBK_SPTR(E) sptr(void *p) {
#if defined(BBFFT_EMU) && defined(BBFFT_EMU_RACECHECK)
    return ::bbfft_emu::checked_ptr<E>{static_cast<E *>(p)};
#else
    return static_cast<E *>(p);
#endif
}

template <int I> struct ic {
    static constexpr int value = I;
};
template <int B, int E, class F> BBK_DEV void static_for(F &&f) {
    if constexpr (B < E) {
        f(ic<B>{});
        static_for<B + 1, E>(f);
    }
}

// ------------------------------------------------------------------------------------------
// In-register FFTs.  `W` is a generated table type: W::n entries, W::c(k) = cos(2 pi k / n),
// W::s(k) = sin(2 pi k / n), exactly rounded from long double on the host.  A transform of
// length R (R | W::n) uses every (W::n / R)-th entry.
// ------------------------------------------------------------------------------------------
template <class T, class W, int R, int K, int DIR> BBK_DEV cx<T> twc() {
    constexpr int idx = ((K % R) + R) % R * (W::n / R);
    constexpr T c = T(W::c(idx));
    constexpr T s = T(W::s(idx));
    return cx<T>{c, DIR < 0 ? -s : s};
}

// multiply by the compile-time constant w_R^K with the trivial cases folded away
template <class T, class W, int R, int K, int DIR> BBK_DEV cx<T> mul_w(cx<T> a) {
    constexpr int k = ((K % R) + R) % R;
    if constexpr (k == 0) {
        return a;
    } else if constexpr (4 * k == R) {
        return mul_i<DIR>(a);
    } else if constexpr (2 * k == R) {
        return cx<T>{-a.x, -a.y};
    } else if constexpr (4 * k == 3 * R) {
        return mul_i<-DIR>(a);
    } else if constexpr (8 * k == R || 8 * k == 3 * R || 8 * k == 5 * R || 8 * k == 7 * R) {
        // (+-1 +- i)/sqrt(2): two adds and two multiplies
        constexpr T h = T(0.70710678118654752440084436210484903928L);
        constexpr int o = 8 * k / R; // 1,3,5,7
        // w = (cr + i*ci*DIR)/sqrt2 with cr,ci in {+1,-1}
        constexpr int cr = (o == 1 || o == 7) ? 1 : -1;
        constexpr int ci0 = (o == 1 || o == 3) ? 1 : -1;
        constexpr int ci = DIR < 0 ? -ci0 : ci0;
        // (a.x + i a.y)(cr + i ci) = (cr a.x - ci a.y) + i (ci a.x + cr a.y)
        T re = (cr > 0 ? a.x : -a.x) - (ci > 0 ? a.y : -a.y);
        T im = (ci > 0 ? a.x : -a.x) + (cr > 0 ? a.y : -a.y);
        return cx<T>{fmul(re, h), fmul(im, h)};
    } else {
        return cmul(a, twc<T, W, R, k, DIR>());
    }
}

// Prime / small butterflies: natural order in, natural order out.
template <class T, class W, int R, int DIR> struct bfly {
    // generic odd prime: direct DFT with the conjugate-pair factoring
    // (same arithmetic idea as the reference's pair_optimization_esum, mixed_radix_fft.cpp:78-126)
    static BBK_DEV void run(cx<T> *v) {
        static_assert(R % 2 == 1, "generic butterfly is for odd radices");
        constexpr int H = (R - 1) / 2;
        cx<T> s[H + 1], d[H + 1];
        static_for<1, H + 1>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            s[j] = v[j] + v[R - j];
            d[j] = v[j] - v[R - j];
        });
        cx<T> x0 = v[0];
        cx<T> sum = x0;
        static_for<1, H + 1>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            sum = sum + s[j];
        });
        v[0] = sum;
        static_for<1, H + 1>([&](auto kk) {
            constexpr int k = decltype(kk)::value;
            cx<T> re = x0; // sum_j cos(jk) s_j
            cx<T> im = cx<T>{T(0), T(0)}; // sum_j sin(jk) d_j
            static_for<1, H + 1>([&](auto jj) {
                constexpr int j = decltype(jj)::value;
                constexpr int idx = (j * k) % R * (W::n / R);
                constexpr T c = T(W::c(idx));
                constexpr T sn = T(W::s(idx));
                re.x = ffma(c, s[j].x, re.x);
                re.y = ffma(c, s[j].y, re.y);
                if constexpr (j == 1) {
                    im.x = fmul(sn, d[j].x);
                    im.y = fmul(sn, d[j].y);
                } else {
                    im.x = ffma(sn, d[j].x, im.x);
                    im.y = ffma(sn, d[j].y, im.y);
                }
            });
            cx<T> rot = mul_i<DIR>(im); // DIR * i * im
            v[k] = re + rot;
            v[R - k] = re - rot;
        });
    }
};

template <class T, class W, int DIR> struct bfly<T, W, 1, DIR> {
    static BBK_DEV void run(cx<T> *) {}
};
template <class T, class W, int DIR> struct bfly<T, W, 2, DIR> {
    static BBK_DEV void run(cx<T> *v) {
        cx<T> a = v[0], b = v[1];
        v[0] = a + b;
        v[1] = a - b;
    }
};
template <class T, class W, int DIR> struct bfly<T, W, 4, DIR> {
    static BBK_DEV void run(cx<T> *v) {
        cx<T> t0 = v[0] + v[2], t1 = v[0] - v[2];
        cx<T> t2 = v[1] + v[3], t3 = mul_i<DIR>(v[1] - v[3]);
        v[0] = t0 + t2;
        v[1] = t1 + t3;
        v[2] = t0 - t2;
        v[3] = t1 - t3;
    }
};
template <class T, class W, int DIR> struct bfly<T, W, 8, DIR> {
    static BBK_DEV void run(cx<T> *v) {
        constexpr T h = T(0.70710678118654752440084436210484903928L);
        // first level: b_j = a_j + a_{j+4}, c_j = (a_j - a_{j+4}) w8^j
        cx<T> b0 = v[0] + v[4], c0 = v[0] - v[4];
        cx<T> b1 = v[1] + v[5], c1 = v[1] - v[5];
        cx<T> b2 = v[2] + v[6], c2 = mul_i<DIR>(v[2] - v[6]);
        cx<T> b3 = v[3] + v[7], c3 = v[3] - v[7];
        // w8^1 = (1 + DIR i)/sqrt2 ; w8^3 = (-1 + DIR i)/sqrt2
        cx<T> r1 = mul_i<DIR>(c1), r3 = mul_i<DIR>(c3);
        c1 = cx<T>{fmul(c1.x + r1.x, h), fmul(c1.y + r1.y, h)};
        c3 = cx<T>{fmul(r3.x - c3.x, h), fmul(r3.y - c3.y, h)};
        // even outputs: radix-4 on b, odd outputs: radix-4 on c
        cx<T> e0 = b0 + b2, e1 = b0 - b2, e2 = b1 + b3, e3 = mul_i<DIR>(b1 - b3);
        cx<T> o0 = c0 + c2, o1 = c0 - c2, o2 = c1 + c3, o3 = mul_i<DIR>(c1 - c3);
        v[0] = e0 + e2;
        v[2] = e1 + e3;
        v[4] = e0 - e2;
        v[6] = e1 - e3;
        v[1] = o0 + o2;
        v[3] = o1 + o3;
        v[5] = o0 - o2;
        v[7] = o1 - o3;
    }
};

BBK_CE int first_radix(int r) {
    if (r % 8 == 0 && r != 16) return 8;
    if (r % 4 == 0) return 4;
    if (r % 2 == 0) return 2;
    for (int p = 3; p * p <= r; p += 2) {
        if (r % p == 0) return p;
    }
    return r;
}

// Composite in-register FFT of length R on v[0..R-1] (natural order in and out).
template <class T, class W, int R, int DIR> struct reg_fft {
    static BBK_DEV void run(cx<T> *v) {
        constexpr int A = first_radix(R);
        if constexpr (A == R) {
            bfly<T, W, R, DIR>::run(v);
        } else {
            constexpr int B = R / A;
            cx<T> y[R]; // y[q*B + n]
            static_for<0, B>([&](auto nn) {
                constexpr int n = decltype(nn)::value;
                cx<T> u[A];
                static_for<0, A>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    u[j] = v[n + B * j];
                });
                bfly<T, W, A, DIR>::run(u);
                static_for<0, A>([&](auto qq) {
                    constexpr int q = decltype(qq)::value;
                    y[q * B + n] = mul_w<T, W, R, n * q, DIR>(u[q]);
                });
            });
            static_for<0, A>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                reg_fft<T, W, B, DIR>::run(y + q * B);
            });
            static_for<0, A>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                static_for<0, B>([&](auto pp) {
                    constexpr int p = decltype(pp)::value;
                    v[q + A * p] = y[q * B + p];
                });
            });
        }
    }
};

// ------------------------------------------------------------------------------------------
// Plan traits contract (generated):
//   real_t, N (complex FFT length run by the stages), NREAL (user-visible length for real
//   modes), DIR, MODE, L (stages), radix(s), T (threads per transform), ML (batch lanes, the
//   fastest thread index), BH (further batch entries per CTA), KLANES (batch lanes index k
//   instead of m; only for M == 1), M, LOAD_STAGED / STORE_STAGED, REAL_FUSED, smem layout LL / PADK /
//   ROW, twiddle block offsets tw_off(s), WR<s> table types, is1()..os2() stride accessors and
//   the ld/st hooks.
// ------------------------------------------------------------------------------------------
template <class C> struct geom {
    static constexpr int B = C::ML * C::BH;          // transforms per CTA
    static constexpr int THREADS = C::ML * C::T * C::BH;
    static BBK_CE int ns(int s) {                 // N_s: sub-problem length entering stage s
        int n = C::N;
        for (int i = 0; i < s; ++i) n /= C::radix(i);
        return n;
    }
    // padded position inside one transform's shared-memory row
    static BBK_DEV int pad(int pos) {
        if constexpr (C::PADK > 0) {
            return pos + pos / C::PADK;
        } else {
            return pos;
        }
    }
    // shared memory offset (in elements) of element `pos` of CTA-local transform b
    static BBK_DEV int soff(int b, int pos) {
        if constexpr (C::LL > 1) {
            return (b % C::LL) + C::LL * pad(pos) + C::ROW * (b / C::LL);
        } else {
            return pad(pos) + C::ROW * b;
        }
    }
};

// digit reversal: position after the last in-place stage -> output bin
template <class C> BBK_DEV int bin_of_sub(int u) {
    // u = (((q0*R1 + q1)*R2 + q2) ... + q_{L-2}); returns q0 + R0*(q1 + R1*(q2 + ...))
    int k = 0;
    int rem = u;
    // peel digits from the least significant (q_{L-2}) upwards
    int digits[4] = {0, 0, 0, 0};
    static_for<0, C::L - 1>([&](auto ii) {
        constexpr int s = C::L - 2 - decltype(ii)::value;
        digits[s] = rem % C::radix(s);
        rem /= C::radix(s);
    });
    static_for<0, C::L - 1>([&](auto ii) {
        constexpr int s = C::L - 2 - decltype(ii)::value;
        k = k * C::radix(s) + digits[s];
    });
    return k;
}
// full digit reversal of a position 0..N-1 (all L digits)
template <class C> BBK_DEV int bin_of_pos(int pos) {
    constexpr int RL = C::radix(C::L - 1);
    int q = pos % RL;
    int u = pos / RL;
    return bin_of_sub<C>(u) + (C::N / RL) * q;
}
// inverse: shared-memory position that holds output bin k after the last stage
template <class C> BBK_DEV int pos_of_bin(int k) {
    int pos = 0;
    int rem = k;
    static_for<0, C::L>([&](auto ss) {
        constexpr int s = decltype(ss)::value;
        pos = pos * C::radix(s) + rem % C::radix(s);
        rem /= C::radix(s);
    });
    return pos;
}

template <class C> struct batch_index {
    u64 m, k;
    bool ok;
};

#ifdef BBFFT_EMU
#define BBK_SYNC() ::bbfft_emu::syncthreads()
#define BBK_TID() (::bbfft_emu::thread_idx())
#define BBK_BID() (::bbfft_emu::block_idx())
#define BBK_NCTAS() (::bbfft_emu::grid_dim())
#define BBK_SMEM() (::bbfft_emu::shared_mem())
#else
#define BBK_SYNC() __syncthreads()
#define BBK_TID() (int(threadIdx.x))
#define BBK_BID() (u64(blockIdx.x))
#define BBK_NCTAS() (u64(gridDim.x))
#define BBK_SMEM() (bbk_dyn_smem)
extern __shared__ __align__(16) unsigned char bbk_dyn_smem[];
#endif

// One Stockham pass.  SRC/DST select where the elements come from / go to.
enum : int {
    IO_GLOBAL = 0,  // complex tensor in global memory
    IO_SMEM = 1,    // the CTA's shared-memory rows
    IO_G_RPAIR = 2, // real tensor, element j = (x[2j], x[2j+1])            (even-N real transforms)
    IO_G_2ROWS = 3  // real tensor, element n = (x_{2k}[n], x_{2k+1}[n])    (odd-N real transforms)
};

// first-stage element load from global memory; `k` is the k slice (IO_G_2ROWS: the slice pair)
template <class C, int SRC>
BBK_DEV cx<typename C::real_t> load_elem(args const &a, u64 m, u64 k, int pos, bool ok) {
    using T = typename C::real_t;
    cx<T> r{T(0), T(0)};
    if constexpr (SRC == IO_GLOBAL) {
        if (ok) r = C::ld(a.in, m + u64(pos) * C::is1(a) + k * C::is2(a));
    } else if constexpr (SRC == IO_G_RPAIR) {
        if (ok) {
            if constexpr (C::PAIR_LOAD) {
                // M == 1, unit stride: the pair is one aligned complex word
                r = reinterpret_cast<const cx<T> *>(a.in)[(u64(2 * pos) + k * C::is2(a)) / 2];
            } else {
                const u64 o = m + u64(2 * pos) * C::is1(a) + k * C::is2(a);
                r.x = C::ldr(a.in, o);
                r.y = C::ldr(a.in, o + C::is1(a));
            }
        }
    } else if constexpr (SRC == IO_G_2ROWS) {
        if (ok) {
            const u64 o = m + u64(pos) * C::is1(a) + (2 * k) * C::is2(a);
            r.x = C::ldr(a.in, o);
            // an unpaired last row is transformed together with zeros
            // (reference: src/base/generator/snippet.cpp:72-85)
            r.y = (2 * k + 1 < a.K) ? C::ldr(a.in, o + C::is2(a)) : T(0);
        }
    }
    return r;
}

// last-stage element store to global memory (`bin` = output index)
template <class C, int DST>
BBK_DEV void store_elem(args const &a, u64 m, u64 k, int bin, cx<typename C::real_t> v, bool ok) {
    if (!ok) return;
    using T = typename C::real_t;
    if constexpr (DST == IO_GLOBAL) {
        C::st(a.out, m + u64(bin) * C::os1(a) + k * C::os2(a), v);
    } else if constexpr (DST == IO_G_RPAIR) {
        if constexpr (C::PAIR_STORE) {
            reinterpret_cast<cx<T> *>(a.out)[(u64(2 * bin) + k * C::os2(a)) / 2] = v;
        } else {
            const u64 o = m + u64(2 * bin) * C::os1(a) + k * C::os2(a);
            C::str(a.out, o, v.x);
            C::str(a.out, o + C::os1(a), v.y);
        }
    } else if constexpr (DST == IO_G_2ROWS) {
        const u64 o = m + u64(bin) * C::os1(a) + (2 * k) * C::os2(a);
        C::str(a.out, o, v.x);
        if (2 * k + 1 < a.K) C::str(a.out, o + C::os2(a), v.y);
    }
}

template <class C, int S, int SRC, int DST>
BBK_DEV void run_stage_regs(args const &a, BBK_SPTR(cx<typename C::real_t>) sm, int t, int b, u64 m, u64 k,
                       bool ok) {
    using T = typename C::real_t;
    using G = geom<C>;
    constexpr int R = C::radix(S);
    constexpr int NS = G::ns(S);
    constexpr int NS1 = NS / R;
    constexpr int NSUB = C::N / R;
    constexpr int CNT = (NSUB + C::T - 1) / C::T;
    constexpr bool LAST = (S == C::L - 1);
    using WR = typename C::template WR<S>;
    const cx<T> *BBK_RESTRICT tw = reinterpret_cast<const cx<T> *>(a.tw) + C::tw_off(S);

    // butterflies + inter-stage twiddles of sub-FFT u held in w[0..R-1]
    auto compute = [&](cx<T> *w, int u) {
        reg_fft<T, WR, R, C::DIR>::run(w);
        if constexpr (!LAST) {
            const int n2 = u % NS1;
            static_for<1, R>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                w[q] = cmul(w[q], ldg_cx(tw + (q - 1) * NS1 + n2));
            });
        }
    };
    auto scatter = [&](cx<T> *w, int u) {
        if constexpr (DST != IO_SMEM) {
            static_assert(LAST, "only the last stage stores to global memory");
            const int k0 = bin_of_sub<C>(u);
            static_for<0, R>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                store_elem<C, DST>(a, m, k, k0 + (C::N / R) * q, w[q], ok);
            });
        } else {
            const int n2 = u % NS1, qi = u / NS1;
            const int base = qi * NS + n2;
            static_for<0, R>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                sm[G::soff(b, base + NS1 * q)] = w[q];
            });
        }
    };

    if constexpr (SRC != IO_SMEM) {
        // first stage: issue every global load of the thread before any arithmetic so that all
        // CNT*R requests are in flight together
        cx<T> v[CNT][R];
        static_for<0, CNT>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            const int u = t + C::T * i;
            if (NSUB % C::T == 0 || u < NSUB) {
                static_for<0, R>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    v[i][j] = load_elem<C, SRC>(a, m, k, u + NS1 * j, ok); // stage 0: q = 0, n2 = u
                });
            }
        });
        static_for<0, CNT>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            const int u = t + C::T * i;
            if (NSUB % C::T == 0 || u < NSUB) {
                compute(v[i], u);
                scatter(v[i], u);
            }
        });
    } else {
        // shared-memory stages: one sub-FFT at a time keeps only R elements live
        static_for<0, CNT>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            const int u = t + C::T * i;
            if (NSUB % C::T == 0 || u < NSUB) {
                cx<T> w[R];
                const int n2 = u % NS1, q = u / NS1;
                const int base = q * NS + n2;
                static_for<0, R>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    w[j] = sm[G::soff(b, base + NS1 * j)];
                });
                compute(w, u);
                scatter(w, u);
            }
        });
    }
}


// A stage whose radix is a prime too large for an in-register butterfly (> 31: the O(p^2) unrolled
// butterfly would not fit the register file; the reference unrolls such primes into one work-item
// and lets the compiler spill).  The sub-FFTs are direct DFTs done cooperatively: the stage's
// inputs sit in shared memory, every thread accumulates ceil(N/T) OUTPUTS (sum_j x[j] w_R^(jq),
// w from a table of R entries); outputs of the last stage go straight to global memory, outputs
// of an earlier stage take the inputs' places after one barrier.  O(R) multiply-adds per element: slow next to the smooth sizes, but every N works.
template <class C, int S, int SRC, int DST>
BBK_DEV void run_stage_direct(args const &a, BBK_SPTR(cx<typename C::real_t>) sm, int t, int b, u64 m, u64 k, bool ok) {
    using T = typename C::real_t;
    using G = geom<C>;
    constexpr int R = C::radix(S);
    constexpr int NS = G::ns(S);
    constexpr int NS1 = NS / R;
    constexpr int NSUB = C::N / R;
    constexpr bool LAST = (S == C::L - 1);
    constexpr int OUTS = (C::N + C::T - 1) / C::T;
    const cx<T> *BBK_RESTRICT tw = reinterpret_cast<const cx<T> *>(a.tw) + C::tw_off(S);
    const cx<T> *BBK_RESTRICT twd = reinterpret_cast<const cx<T> *>(a.tw) + C::tw_dir(S);
    if constexpr (SRC != IO_SMEM) {
        static_for<0, OUTS>([&](auto ii) {
            const int pos = t + C::T * decltype(ii)::value;
            if (C::N % C::T == 0 || pos < C::N) sm[G::soff(b, pos)] = load_elem<C, SRC>(a, m, k, pos, ok);
        });
        BBK_SYNC();
    }
    // one output: sum_j x[j] w_R^(j q) of sub-FFT u
    auto output = [&](int u, int q) {
        // consecutive threads take consecutive sub-FFTs (consecutive shared-memory positions)
        const int base = (u / NS1) * NS + u % NS1;
        cx<T> s = sm[G::soff(b, base)];
        int idx = 0;
        for (int j = 1; j < R; ++j) {
            idx += q;
            if (idx >= R) idx -= R;
            const cx<T> x = sm[G::soff(b, base + NS1 * j)];
            const cx<T> w = ldg_cx(twd + idx);
            s.x = ffma(x.x, w.x, s.x);
            s.x = ffma(-x.y, w.y, s.x);
            s.y = ffma(x.x, w.y, s.y);
            s.y = ffma(x.y, w.x, s.y);
        }
        return s;
    };
    if constexpr (DST != IO_SMEM) {
        // Results leave for global memory and nothing in shared memory is overwritten: every
        // output is stored as soon as it is complete.  A plain loop (two independent sums in
        // flight) instead of OUTS unrolled accumulators held across a barrier: round 1's unrolled
        // form made NVRTC 12.9's ptxas spill under __maxnreg__ and emit a store address from a
        // dead register (fp64 c2r M=32 N=424, profiles/r02b_ptxas_miscompile.md).
        static_assert(LAST, "only the last stage stores to global memory");
        constexpr int STEP = 2 * C::T;
#pragma unroll 1
        for (int o = t; o < C::N; o += STEP) {
            const int o2 = o + C::T;
            const int u = o % NSUB, q = o / NSUB;
            const bool two = o2 < C::N;
            const int u2 = two ? o2 % NSUB : u, q2 = two ? o2 / NSUB : q;
            const cx<T> v = output(u, q);
            const cx<T> v2 = output(u2, q2);
            store_elem<C, DST>(a, m, k, bin_of_sub<C>(u) + (C::N / R) * q, v, ok);
            if (two) store_elem<C, DST>(a, m, k, bin_of_sub<C>(u2) + (C::N / R) * q2, v2, ok);
        }
    } else {
        cx<T> acc[OUTS];
        static_for<0, OUTS>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            const int o = t + C::T * i;
            acc[i] = cx<T>{T(0), T(0)};
            if (C::N % C::T == 0 || o < C::N) acc[i] = output(o % NSUB, o / NSUB);
        });
        BBK_SYNC(); // every input is read before an output takes its place
        static_for<0, OUTS>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            const int o = t + C::T * i;
            if (C::N % C::T == 0 || o < C::N) {
                const int u = o % NSUB, q = o / NSUB;
                const int n2 = u % NS1;
                cx<T> v = acc[i];
                if constexpr (!LAST) {
                    if (q > 0) v = cmul(v, ldg_cx(tw + (q - 1) * NS1 + n2));
                }
                sm[G::soff(b, (u / NS1) * NS + n2 + NS1 * q)] = v;
            }
        });
    }
}

template <class C, int S, int SRC, int DST>
BBK_DEV void run_stage(args const &a, BBK_SPTR(cx<typename C::real_t>) sm, int t, int b, u64 m, u64 k,
                       bool ok) {
    if constexpr (C::direct(S)) {
        run_stage_direct<C, S, SRC, DST>(a, sm, t, b, m, k, ok);
    } else {
        run_stage_regs<C, S, SRC, DST>(a, sm, t, b, m, k, ok);
    }
}

// stages S..L-1 with all intermediate exchanges in shared memory
template <class C, int S, int SRC0, int DSTL> BBK_DEV void run_stages(args const &a, BBK_SPTR(cx<typename C::real_t>) sm, int t, int b, u64 m, u64 k, bool ok) {
    if constexpr (S < C::L) {
        constexpr int SRC = (S == 0) ? SRC0 : IO_SMEM;
        constexpr int DST = (S == C::L - 1) ? DSTL : IO_SMEM;
        if constexpr (S > 0) {
            BBK_SYNC();
        }
        run_stage<C, S, SRC, DST>(a, sm, t, b, m, k, ok);
        run_stages<C, S + 1, SRC0, DSTL>(a, sm, t, b, m, k, ok);
    }
}

// stages S..END-1, all of them writing to shared memory (the caller runs stage END itself)
template <class C, int S, int END, int SRC0>
BBK_DEV void run_front_stages(args const &a, BBK_SPTR(cx<typename C::real_t>) sm, int t, int b, u64 m, u64 k, bool ok) {
    if constexpr (S < END) {
        constexpr int SRC = (S == 0) ? SRC0 : IO_SMEM;
        if constexpr (S > 0) {
            BBK_SYNC();
        }
        run_stage<C, S, SRC, IO_SMEM>(a, sm, t, b, m, k, ok);
        run_front_stages<C, S + 1, END, SRC0>(a, sm, t, b, m, k, ok);
    }
}

// inverse of bin_of_sub: the last-stage sub-FFT whose first output bin is k0
template <class C> BBK_DEV int sub_of_bin(int k0) {
    int u = 0, rem = k0;
    static_for<0, C::L - 1>([&](auto ss) {
        constexpr int s = decltype(ss)::value;
        u = u * C::radix(s) + rem % C::radix(s);
        rem /= C::radix(s);
    });
    return u;
}

// ------------------------------------------------------------------------------------------
// Real transforms with the pre / post pass fused into the first / last stage.
//
// Both real schemes pair element i with element NN - i (NN = complex length).  In the LAST stage
// (radix R, NB = NN/R sub-FFTs) the sub-FFT with first bin k0 produces the bins k0 + NB q, and
// their partners NN - k0 - NB q = (NB - k0) + NB (R-1-q) all come out of the sub-FFT with first
// bin NB - k0.  In the FIRST stage (NS1 = NN/R sub-FFTs) sub-FFT u consumes the positions
// u + NS1 j, whose partners (NS1 - u) + NS1 (R-1-j) feed sub-FFT NS1 - u.  A thread that runs a
// sub-FFT together with its mirror therefore holds every pair in registers: the post-twiddle of
// r2c and the pre-twiddle of c2r need no extra trip through shared memory and no extra barrier.
// Units p = 0 .. NB/2 (resp. NS1/2); p = 0 and 2p = NB are their own mirrors.
// ------------------------------------------------------------------------------------------
template <class C, int SRC>
BBK_DEV void r2c_last_stage(args const &a, BBK_SPTR(cx<typename C::real_t>) sm, int t, int b, u64 m, u64 k, bool ok) {
    using T = typename C::real_t;
    using G = geom<C>;
    constexpr int S = C::L - 1;
    constexpr int R = C::radix(S);
    constexpr int NN = C::N;
    constexpr int NB = NN / R;
    constexpr int UNITS = NB / 2 + 1;
    constexpr int CNTU = (UNITS + C::T - 1) / C::T;
    constexpr bool HALF = (C::MODE == R2C_HALF);
    using WR = typename C::template WR<S>;
    const cx<T> *BBK_RESTRICT twr = reinterpret_cast<const cx<T> *>(a.tw) + C::TW_REAL;

    auto fetch = [&](cx<T> *v, int u) {
        static_for<0, R>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            if constexpr (SRC == IO_SMEM) {
                v[j] = sm[G::soff(b, u * R + j)];
            } else {
                v[j] = load_elem<C, SRC>(a, m, k, u * R + j, ok);
            }
        });
    };
    // yi = Y[i], yn = Y[NN - i]
    auto emit = [&](int i, cx<T> yi, cx<T> yn) {
        if constexpr (HALF) {
            const cx<T> y2 = conj(yn);
            const cx<T> w = ldg_cx(twr + i);
            const cx<T> iw = cx<T>{-w.y, w.x};
            const cx<T> aa = rmul(y2 + yi, T(0.5));
            const cx<T> bb = cmul(rmul(y2 - yi, T(0.5)), iw);
            if (ok) {
                C::st(a.out, m + u64(i) * C::os1(a) + k * C::os2(a), aa + bb);
                // X[NN] comes out of the pair (0, 0); the pair (NN/2, NN/2) has one member
                if (2 * i != NN) {
                    C::st(a.out, m + u64(NN - i) * C::os1(a) + k * C::os2(a), conj(aa - bb));
                }
            }
        } else {
            // rows 2k, 2k+1: A[idx] = (conj(Y[N-idx]) + Y[idx])/2, B[idx] = i (conj(Y[N-idx]) - Y[idx])/2
            const bool low = 2 * i <= NN;
            const int idx = low ? i : NN - i;
            const cx<T> y1 = low ? yi : yn;
            const cx<T> y2 = conj(low ? yn : yi);
            const cx<T> av = rmul(y2 + y1, T(0.5));
            const cx<T> d = rmul(y2 - y1, T(0.5));
            if (ok) {
                const u64 o = m + u64(idx) * C::os1(a) + (2 * k) * C::os2(a);
                C::st(a.out, o, av);
                if (2 * k + 1 < a.K) C::st(a.out, o + C::os2(a), cx<T>{-d.y, d.x});
            }
        }
    };

    if constexpr (SRC != IO_SMEM) {
        // Single-stage kernel (NB == 1: one unit, one thread per transform): it reads and writes
        // global memory only, and in-place rows of different m overlap, so every load of the CTA
        // precedes its first store.  The barrier sits outside any thread-dependent branch.
        static_assert(NB == 1 && C::T == 1, "r2c_last_stage from global memory is the single-stage kernel");
        cx<T> va[R];
        fetch(va, 0);
        BBK_SYNC();
        reg_fft<T, WR, R, C::DIR>::run(va);
        emit(0, va[0], va[0]);
        static_for<1, (R + 1) / 2>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            emit(q, va[q], va[R - q]);
        });
        if constexpr (R % 2 == 0) {
            emit(R / 2, va[R / 2], va[R / 2]);
        }
        return;
    }
    static_for<0, CNTU>([&](auto cc) {
        const int p = t + C::T * decltype(cc)::value;
        if (UNITS % C::T == 0 || p < UNITS) {
            const int kb = (NB - p) % NB;
            cx<T> va[R], vb[R];
            fetch(va, sub_of_bin<C>(p));
            if (kb != p) fetch(vb, sub_of_bin<C>(kb));
            reg_fft<T, WR, R, C::DIR>::run(va);
            if (kb != p) {
                reg_fft<T, WR, R, C::DIR>::run(vb);
                static_for<0, R>([&](auto qq) {
                    constexpr int q = decltype(qq)::value;
                    emit(p + NB * q, va[q], vb[R - 1 - q]);
                });
            } else if (p == 0) {
                emit(0, va[0], va[0]);
                static_for<1, (R + 1) / 2>([&](auto qq) {
                    constexpr int q = decltype(qq)::value;
                    emit(NB * q, va[q], va[R - q]);
                });
                if constexpr (R % 2 == 0) {
                    emit(NB * (R / 2), va[R / 2], va[R / 2]);
                }
            } else {
                // 2p == NB: bins p + NB q pair with p + NB (R-1-q)
                static_for<0, R / 2>([&](auto qq) {
                    constexpr int q = decltype(qq)::value;
                    emit(p + NB * q, va[q], va[R - 1 - q]);
                });
                if constexpr (R % 2 == 1) {
                    emit(p + NB * ((R - 1) / 2), va[(R - 1) / 2], va[(R - 1) / 2]);
                }
            }
        }
    });
}

template <class C, int DST>
BBK_DEV void c2r_first_stage(args const &a, BBK_SPTR(cx<typename C::real_t>) sm, int t, int b, u64 m, u64 k, bool ok) {
    using T = typename C::real_t;
    using G = geom<C>;
    constexpr int R = C::radix(0);
    constexpr int NN = C::N;
    constexpr int NS1 = NN / R;
    constexpr int UNITS = NS1 / 2 + 1;
    constexpr int CNTU = (UNITS + C::T - 1) / C::T;
    constexpr bool HALF = (C::MODE == C2R_HALF);
    constexpr bool LAST = (C::L == 1);
    using WR = typename C::template WR<0>;
    const cx<T> *BBK_RESTRICT tw = reinterpret_cast<const cx<T> *>(a.tw) + C::tw_off(0);
    const cx<T> *BBK_RESTRICT twr = reinterpret_cast<const cx<T> *>(a.tw) + C::TW_REAL;

    // spectrum element i of this thread's column (half mode), or rows A / B of the slice pair
    auto ldx = [&](int i) { return C::ld(a.in, m + u64(i) * C::is1(a) + k * C::is2(a)); };
    // half mode: xa[j] = X[p + NS1 j], xb[j] = X[ub + NS1 j] (p == 0: xb[0] = X[NN])
    // double mode: pair j of the unit is (i, NN - i) with i = p + NS1 j; xa[j] = A[min], xb[j] = B[min]
    cx<T> xa[CNTU][R], xb[CNTU][R];
    static_for<0, CNTU>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        const int p = t + C::T * c;
        static_for<0, R>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            xa[c][j] = cx<T>{T(0), T(0)};
            xb[c][j] = cx<T>{T(0), T(0)};
        });
        if ((UNITS % C::T == 0 || p < UNITS) && ok) {
            const int ub = (NS1 - p) % NS1;
            if constexpr (HALF) {
                static_for<0, R>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    xa[c][j] = ldx(p + NS1 * j);
                });
                if (ub != p) {
                    static_for<0, R>([&](auto jj) {
                        constexpr int j = decltype(jj)::value;
                        xb[c][j] = ldx(ub + NS1 * j);
                    });
                } else if (p == 0) {
                    xb[c][0] = ldx(NN);
                }
            } else {
                // NN is odd: only p == 0 is its own mirror
                static_for<0, R>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    const int i = p + NS1 * j;
                    const int idx = 2 * i <= NN ? i : NN - i;
                    // the mirror unit's pairs are the same pairs: for p == 0 only j <= R/2 is new
                    if (ub != p || j <= R / 2) {
                        const u64 o = m + u64(idx) * C::is1(a) + (2 * k) * C::is2(a);
                        xa[c][j] = C::ld(a.in, o);
                        if (2 * k + 1 < a.K) xb[c][j] = C::ld(a.in, o + C::is2(a));
                    }
                });
            }
        }
    });
    if constexpr (LAST) {
        BBK_SYNC(); // in-place rows of different m overlap: loads of the CTA before its stores
    }
    // z[i], z[NN-i] from the pair (half mode: X[i], X[NN-i]; double mode: A[idx], B[idx])
    auto pre = [&](int i, cx<T> u1, cx<T> u2, cx<T> &zi, cx<T> &zn) {
        if constexpr (HALF) {
            cx<T> x1 = u1;
            if (i == 0) x1.y = T(0);
            const cx<T> x2 = conj(u2);
            const cx<T> w = ldg_cx(twr + i);
            const cx<T> iw = cx<T>{-w.y, w.x};
            const cx<T> aa = x1 + x2;
            const cx<T> bb = cmul(x1 - x2, iw);
            zi = aa + bb;
            zn = conj(aa - bb);
        } else {
            cx<T> av = u1, bv = u2;
            if (i == 0) {
                av.y = T(0);
                bv.y = T(0);
            }
            const cx<T> ylo = cx<T>{av.x - bv.y, av.y + bv.x}; // Y[idx]
            const cx<T> yhi = cx<T>{av.x + bv.y, bv.x - av.y}; // Y[NN - idx]
            const bool low = 2 * i <= NN;
            zi = low ? ylo : yhi;
            zn = low ? yhi : ylo;
        }
    };
    auto finish = [&](cx<T> *w, int u) {
        reg_fft<T, WR, R, C::DIR>::run(w);
        if constexpr (!LAST) {
            static_for<1, R>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                w[q] = cmul(w[q], ldg_cx(tw + (q - 1) * NS1 + u));
            });
            static_for<0, R>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                sm[G::soff(b, u + NS1 * q)] = w[q];
            });
        } else {
            static_for<0, R>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                store_elem<C, DST>(a, m, k, q, w[q], ok); // single stage: NS1 == 1, u == 0
            });
        }
    };
    static_for<0, CNTU>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        const int p = t + C::T * c;
        if (UNITS % C::T == 0 || p < UNITS) {
            const int ub = (NS1 - p) % NS1;
            cx<T> za[R], zb[R];
            cx<T> dummy;
            if (ub != p) {
                static_for<0, R>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    if constexpr (HALF) {
                        pre(p + NS1 * j, xa[c][j], xb[c][R - 1 - j], za[j], zb[R - 1 - j]);
                    } else {
                        pre(p + NS1 * j, xa[c][j], xb[c][j], za[j], zb[R - 1 - j]);
                    }
                });
                finish(za, p);
                finish(zb, ub);
            } else if (p == 0) {
                if constexpr (HALF) {
                    pre(0, xa[c][0], xb[c][0], za[0], dummy);
                } else {
                    pre(0, xa[c][0], xb[c][0], za[0], dummy);
                }
                static_for<1, (R + 1) / 2>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    if constexpr (HALF) {
                        pre(NS1 * j, xa[c][j], xa[c][R - j], za[j], za[R - j]);
                    } else {
                        pre(NS1 * j, xa[c][j], xb[c][j], za[j], za[R - j]);
                    }
                });
                if constexpr (R % 2 == 0) {
                    // half mode only (NN even): i == NN/2 is its own partner
                    pre(NS1 * (R / 2), xa[c][R / 2], xa[c][R / 2], za[R / 2], dummy);
                }
                finish(za, 0);
            } else {
                // 2p == NS1 (half mode only): positions p + NS1 j pair with p + NS1 (R-1-j)
                static_for<0, R / 2>([&](auto jj) {
                    constexpr int j = decltype(jj)::value;
                    pre(p + NS1 * j, xa[c][j], xa[c][R - 1 - j], za[j], za[R - 1 - j]);
                });
                if constexpr (R % 2 == 1) {
                    pre(p + NS1 * ((R - 1) / 2), xa[c][(R - 1) / 2], xa[c][(R - 1) / 2], za[(R - 1) / 2], dummy);
                }
                finish(za, p);
            }
        }
    });
}

// Cooperative, coalesced copy of the CTA's batch of rows between global and shared memory.
// NROW = row length in elements of type E, rows are addressed (m, n, k) -> m + n*s1 + k*s2.
template <class C, class E, int NROW, bool TO_SMEM, bool REVERSE, class LD, class ST>
BBK_DEV void coop_copy(BBK_SPTR(E) sm, u64 m0, u64 k0, u64 Mtot, u64 K, i64 s1, i64 s2, int tid, LD ld, ST st) {
    using G = geom<C>;
    // iterate (ml, n, bh) with ml fastest so that consecutive threads touch consecutive addresses
    constexpr int MLC = C::KLANES ? 1 : C::ML;          // m entries per CTA
    constexpr int KB = C::KLANES ? C::ML * C::BH : C::BH; // k entries per CTA
    constexpr int TOTAL = MLC * NROW * KB;
    for (int idx = tid; idx < TOTAL; idx += G::THREADS) {
        const int ml = idx % MLC;
        const int n = (idx / MLC) % NROW;
        const int kb = idx / (MLC * NROW);
        const u64 m = m0 + ml, k = k0 + kb;
        if (m < Mtot && k < K) {
            const int b = C::KLANES ? kb : ml + C::ML * kb;
            int pos = n;
            if constexpr (REVERSE) {
                // the spectrum of an even-N real transform has one extra bin kept in slot N
                pos = (NROW > C::N && n == C::N) ? n : pos_of_bin<C>(n);
            }
            const u64 g = m + u64(n) * s1 + k * s2;
            if constexpr (TO_SMEM) {
                sm[G::soff(b, pos)] = ld(g);
            } else {
                st(g, sm[G::soff(b, pos)]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// L2 prefetch of a FUTURE batch.  A CTA is stage-synchronised: it loads, transforms, stores, and
// while it transforms it keeps no load in flight; with the two or three CTAs per SM that the
// register file allows for the long transforms, HBM idles part of the time (round 1: fp64 N=490
// at 56 % DRAM busy, warps_active 21 %).  Instead of a second shared-memory buffer per CTA (there
// is no room: 62 KiB per batch for fp64 N=490) the CTA that starts now asks the L2 -- 126 MB, far
// larger than a wave of batches -- to fetch the input of the batch that will run `pf` batches
// later (about one and a half waves), with cp.async.bulk.prefetch.L2: one instruction per 4 KiB,
// no registers, no shared memory, no completion to wait for.  The later CTA's loads then hit L2.
// Only for inputs whose rows are contiguous inside a k slice and without a user load callback
// (a callback may read anywhere).  Byte ranges are rounded inwards to 16 B and clamped to the
// slices of this launch, so nothing outside the tensor is ever named.
// ------------------------------------------------------------------------------------------
#ifdef BBFFT_EMU
BBK_DEV void l2_prefetch_range(const void *, u64, u64, int, int) {}
#else
BBK_DEV void l2_prefetch_range(const void *base, u64 lo, u64 hi, int tid, int nthreads) {
    // bytes [lo, hi) relative to base, dealt to the WARPS of the CTA in 4 KiB pieces: the instruction
    // takes its operands from uniform registers, so one lane per warp issues it (a per-lane address
    // would be serialised lane by lane)
    constexpr u64 CH = 4096;
    if ((tid & 31) != 0) return;
    const int warp = tid >> 5, nwarps = (nthreads + 31) >> 5;
    u64 p0 = reinterpret_cast<u64>(base) + lo;
    u64 p1 = reinterpret_cast<u64>(base) + hi;
    p0 = (p0 + 15) & ~u64(15);
    p1 = p1 & ~u64(15);
    for (u64 p = p0 + u64(warp) * CH; p < p1; p += u64(nwarps) * CH) {
        const unsigned n = unsigned(p1 - p < CH ? p1 - p : CH);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(n) : "memory");
    }
}
#endif

template <class C> BBK_DEV void prefetch_future_batch(args const &a, u64 bid, int tid) {
    using T = typename C::real_t;
    using G = geom<C>;
    constexpr bool IN_REAL = (C::MODE == R2C_HALF || C::MODE == R2C_DOUBLE);
    constexpr bool DOUBLE = (C::MODE == R2C_DOUBLE || C::MODE == C2R_DOUBLE);
    constexpr u64 IE = IN_REAL ? sizeof(T) : 2 * sizeof(T);
    // rows stored per k slice on the input side
    constexpr u64 NIN = C::MODE == C2C          ? u64(C::N)
                        : C::MODE == R2C_HALF   ? u64(C::NREAL)
                        : C::MODE == R2C_DOUBLE ? u64(C::N)
                        : C::MODE == C2R_HALF   ? u64(C::N) + 1
                                                : u64(C::N) / 2 + 1;
    if constexpr (!C::HAS_LOAD_CALLBACK) {
        if (a.pf == 0) return;
        if (u64(C::is1(a)) != C::M) return; // rows of a slice are not contiguous (folds at compile time)
        constexpr u64 MBLKS = C::KLANES ? 1 : (C::M + C::ML - 1) / C::ML;
        constexpr u64 KPER = C::KLANES ? u64(G::B) : u64(C::BH); // kernel-k per CTA
        const u64 part = C::KLANES ? 0 : bid % MBLKS;
        const u64 kb = (C::KLANES ? bid : bid / MBLKS) + a.pf; // the future batch of k
        const u64 sl = DOUBLE ? 2 : 1;                          // user slices per kernel-k
        const u64 uk0 = kb * KPER * sl;
        if (uk0 >= a.K) return;
        const u64 uk1 = uk0 + KPER * sl < a.K ? uk0 + KPER * sl : a.K;
        const u64 slice_bytes = NIN * C::M * IE, stride_bytes = u64(C::is2(a)) * IE;
        if (stride_bytes == slice_bytes) {
            // packed slices: one range for the whole batch, one part per m-block CTA
            const u64 lo = uk0 * stride_bytes, len = (uk1 - uk0) * stride_bytes;
            l2_prefetch_range(a.in, lo + len * part / MBLKS, lo + len * (part + 1) / MBLKS, tid, G::THREADS);
        } else {
            for (u64 uk = uk0; uk < uk1; ++uk) {
                const u64 lo = uk * stride_bytes;
                l2_prefetch_range(a.in, lo + slice_bytes * part / MBLKS, lo + slice_bytes * (part + 1) / MBLKS, tid, G::THREADS);
            }
        }
    }
}

// the work of CTA `bid` of the 1d kernel's grid
template <class C> BBK_DEV void fft1d_cta(args const &a, const u64 bid) {
    using T = typename C::real_t;
    using G = geom<C>;
    BBK_SPTR(cx<T>) sm = sptr<cx<T>>(BBK_SMEM());
    const int tid = BBK_TID();
    const int l0 = tid % C::ML;
    const int t = (tid / C::ML) % C::T;
    const int bh = tid / (C::ML * C::T);
    const int b = l0 + C::ML * bh;
    u64 m, k, m0, k0;
    if constexpr (C::KLANES) {
        m0 = 0;
        k0 = bid * G::B;
        m = 0;
        k = k0 + b;
    } else {
        constexpr u64 MBLKS = (C::M + C::ML - 1) / C::ML;
        m0 = (bid % MBLKS) * C::ML;
        k0 = (bid / MBLKS) * C::BH;
        m = m0 + l0;
        k = k0 + bh;
    }
    const bool ok = (m < C::M) && (k < a.K);
    prefetch_future_batch<C>(a, bid, tid);

    if constexpr (C::MODE == C2C) {
        constexpr int SRC0 = C::LOAD_STAGED ? IO_SMEM : IO_GLOBAL;
        constexpr int DSTL = C::STORE_STAGED ? IO_SMEM : IO_GLOBAL;
        if constexpr (C::LOAD_STAGED) {
            coop_copy<C, cx<T>, C::N, true, false>(
                sm, m0, k0, C::M, a.K, C::is1(a), C::is2(a), tid,
                [&](u64 g) { return C::ld(a.in, g); }, [&](u64, cx<T>) {});
            BBK_SYNC();
        }
        run_stages<C, 0, SRC0, DSTL>(a, sm, t, b, m, k, ok);
        if constexpr (C::STORE_STAGED) {
            BBK_SYNC();
            coop_copy<C, cx<T>, C::N, false, true>(
                sm, m0, k0, C::M, a.K, C::os1(a), C::os2(a), tid, [&](u64) { return cx<T>{}; },
                [&](u64 g, cx<T> v) { C::st(a.out, g, v); });
        }
    } else if constexpr (C::MODE == R2C_HALF) {
        // z[j] = x[2j] + i x[2j+1]; Y = FFT_h(z); X[i] = a + b, X[h-i] = conj(a - b) with
        // a = (conj(Y[h-i]) + Y[i])/2, b = (conj(Y[h-i]) - Y[i])/2 * (i w_N^i)
        // (reference: src/base/generator/sbfft_gen.cpp:180-200, f2fft_gen.cpp:253-271)
        constexpr int H = C::N;
        if constexpr (C::REAL_FUSED) {
            // post-twiddle fused into the last stage (registers only)
            if constexpr (C::L == 1) {
                r2c_last_stage<C, IO_G_RPAIR>(a, sm, t, b, m, k, ok);
            } else {
                run_front_stages<C, 0, C::L - 1, IO_G_RPAIR>(a, sm, t, b, m, k, ok);
                BBK_SYNC();
                r2c_last_stage<C, IO_SMEM>(a, sm, t, b, m, k, ok);
            }
            return;
        }
        if constexpr (C::LOAD_STAGED) {
            // M == 1: a real row is H aligned complex words
            coop_copy<C, cx<T>, H, true, false>(
                sm, m0, k0, C::M, a.K, 1, C::is2(a) / 2, tid,
                [&](u64 g) { return reinterpret_cast<const cx<T> *>(a.in)[g]; }, [&](u64, cx<T>) {});
            BBK_SYNC();
            run_stages<C, 0, IO_SMEM, IO_SMEM>(a, sm, t, b, m, k, ok);
        } else {
            run_stages<C, 0, IO_G_RPAIR, IO_SMEM>(a, sm, t, b, m, k, ok);
        }
        BBK_SYNC();
        const cx<T> *BBK_RESTRICT twr = reinterpret_cast<const cx<T> *>(a.tw) + C::TW_REAL;
        constexpr int PAIRS = H / 2 + 1;
        constexpr int PCNT = (PAIRS + C::T - 1) / C::T;
        static_for<0, PCNT>([&](auto cc) {
            const int i = t + C::T * decltype(cc)::value;
            if (i < PAIRS) {
                const int p1 = G::soff(b, pos_of_bin<C>(i));
                const int p2 = G::soff(b, pos_of_bin<C>((H - i) % H));
                const cx<T> y1 = sm[p1];
                const cx<T> y2r = sm[p2];
                const cx<T> y2 = conj(y2r);
                const cx<T> w = ldg_cx(twr + i);
                const cx<T> iw = cx<T>{-w.y, w.x};
                const cx<T> aa = rmul(y2 + y1, T(0.5));
                const cx<T> bb = cmul(rmul(y2 - y1, T(0.5)), iw);
                const cx<T> xi = aa + bb;
                const cx<T> xh = conj(aa - bb);
                if constexpr (C::STORE_STAGED) {
                    // in place: X[i] and X[h-i] take the slots of Y[i] and Y[h-i]; X[h] the extra slot
                    sm[p1] = xi;
                    if (i == 0) {
                        sm[G::soff(b, H)] = xh;
                    } else if (2 * i != H) {
                        sm[p2] = xh;
                    }
                } else if (ok) {
                    C::st(a.out, m + u64(i) * C::os1(a) + k * C::os2(a), xi);
                    if (2 * i != H) C::st(a.out, m + u64(H - i) * C::os1(a) + k * C::os2(a), xh);
                }
            }
        });
        if constexpr (C::STORE_STAGED) {
            BBK_SYNC();
            coop_copy<C, cx<T>, H + 1, false, true>(
                sm, m0, k0, C::M, a.K, C::os1(a), C::os2(a), tid, [&](u64) { return cx<T>{}; },
                [&](u64 g, cx<T> v) { C::st(a.out, g, v); });
        }
    } else if constexpr (C::MODE == C2R_HALF) {
        // z[i] = a + b, z[h-i] = conj(a - b) with x1 = X[i] (imag(X[0]) ignored), x2 = conj(X[h-i]),
        // a = x1 + x2, b = (x1 - x2) * (i w_N^i); x[2j], x[2j+1] = Re, Im of IFFT_h(z)[j]
        // (reference: src/base/generator/sbfft_gen.cpp:221-247, f2fft_gen.cpp:349-409)
        constexpr int H = C::N;
        if constexpr (C::REAL_FUSED) {
            // pre-twiddle fused into the first stage (registers only)
            c2r_first_stage<C, IO_G_RPAIR>(a, sm, t, b, m, k, ok);
            if constexpr (C::L > 1) {
                run_stages<C, 1, IO_SMEM, IO_G_RPAIR>(a, sm, t, b, m, k, ok);
            }
            return;
        }
        if constexpr (C::LOAD_STAGED) {
            coop_copy<C, cx<T>, H + 1, true, false>(
                sm, m0, k0, C::M, a.K, C::is1(a), C::is2(a), tid,
                [&](u64 g) { return C::ld(a.in, g); }, [&](u64, cx<T>) {});
            BBK_SYNC();
        }
        const cx<T> *BBK_RESTRICT twr = reinterpret_cast<const cx<T> *>(a.tw) + C::TW_REAL;
        constexpr int PAIRS = H / 2 + 1;
        constexpr int PCNT = (PAIRS + C::T - 1) / C::T;
        // all of the thread's loads are issued before the first pre-twiddle (memory-level
        // parallelism, as in the first stage of the complex transform)
        cx<T> x1[PCNT], x2[PCNT], wv[PCNT];
        static_for<0, PCNT>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            const int i = t + C::T * c;
            x1[c] = cx<T>{T(0), T(0)};
            x2[c] = cx<T>{T(0), T(0)};
            if (i < PAIRS) {
                if constexpr (C::LOAD_STAGED) {
                    x1[c] = sm[G::soff(b, i)];
                    x2[c] = sm[G::soff(b, H - i)];
                } else if (ok) {
                    x1[c] = C::ld(a.in, m + u64(i) * C::is1(a) + k * C::is2(a));
                    x2[c] = C::ld(a.in, m + u64(H - i) * C::is1(a) + k * C::is2(a));
                }
                wv[c] = ldg_cx(twr + i);
            }
        });
        static_for<0, PCNT>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            const int i = t + C::T * c;
            if (i < PAIRS) {
                cx<T> p1 = x1[c];
                if (i == 0) p1.y = T(0);
                const cx<T> p2 = conj(x2[c]);
                const cx<T> iw = cx<T>{-wv[c].y, wv[c].x};
                const cx<T> aa = p1 + p2;
                const cx<T> bb = cmul(p1 - p2, iw);
                // pair i owns the slots {i, h-i}: no other thread reads or writes them in this pass
                sm[G::soff(b, i)] = aa + bb;
                if (i != 0 && 2 * i != H) sm[G::soff(b, H - i)] = conj(aa - bb);
            }
        });
        BBK_SYNC();
        if constexpr (C::STORE_STAGED) {
            run_stages<C, 0, IO_SMEM, IO_SMEM>(a, sm, t, b, m, k, ok);
            BBK_SYNC();
            // M == 1: a real row is H aligned complex words
            coop_copy<C, cx<T>, H, false, true>(
                sm, m0, k0, C::M, a.K, 1, C::os2(a) / 2, tid, [&](u64) { return cx<T>{}; },
                [&](u64 g, cx<T> v) { reinterpret_cast<cx<T> *>(a.out)[g] = v; });
        } else {
            run_stages<C, 0, IO_SMEM, IO_G_RPAIR>(a, sm, t, b, m, k, ok);
        }
    } else if constexpr (C::MODE == R2C_DOUBLE) {
        // rows 2k and 2k+1 are the real and imaginary part of one complex FFT:
        // A[i] = (conj(Y[N-i]) + Y[i])/2, B[i] = i (conj(Y[N-i]) - Y[i])/2, i <= N/2
        // (reference: src/base/generator/sbfft_gen.cpp:274-291, f2fft_gen.cpp:314-330)
        const bool okp = (m < C::M) && (2 * k < a.K);
        if constexpr (C::REAL_FUSED) {
            if constexpr (C::L == 1) {
                r2c_last_stage<C, IO_G_2ROWS>(a, sm, t, b, m, k, okp);
            } else {
                run_front_stages<C, 0, C::L - 1, IO_G_2ROWS>(a, sm, t, b, m, k, okp);
                BBK_SYNC();
                r2c_last_stage<C, IO_SMEM>(a, sm, t, b, m, k, okp);
            }
            return;
        }
        run_stages<C, 0, IO_G_2ROWS, IO_SMEM>(a, sm, t, b, m, k, okp);
        BBK_SYNC();
        constexpr int PAIRS = C::N / 2 + 1;
        constexpr int PCNT = (PAIRS + C::T - 1) / C::T;
        static_for<0, PCNT>([&](auto cc) {
            const int i = t + C::T * decltype(cc)::value;
            if (i < PAIRS && okp) {
                const cx<T> y1 = sm[G::soff(b, pos_of_bin<C>(i))];
                const cx<T> y2r = sm[G::soff(b, pos_of_bin<C>((C::N - i) % C::N))];
                const cx<T> y2 = conj(y2r);
                const cx<T> av = rmul(y2 + y1, T(0.5));
                const cx<T> d = rmul(y2 - y1, T(0.5));
                const cx<T> bv = cx<T>{-d.y, d.x}; // i * d
                const u64 o = m + u64(i) * C::os1(a) + (2 * k) * C::os2(a);
                C::st(a.out, o, av);
                if (2 * k + 1 < a.K) C::st(a.out, o + C::os2(a), bv);
            }
        });
    } else if constexpr (C::MODE == C2R_DOUBLE) {
        // Y[i] = A[i] + i B[i], Y[N-i] = conj(A[i]) + i conj(B[i]); rows 2k, 2k+1 = Re, Im of IFFT_N(Y)
        // (reference: src/base/generator/sbfft_gen.cpp:318-351, f2fft_gen.cpp:443-502)
        const bool okp = (m < C::M) && (2 * k < a.K);
        if constexpr (C::REAL_FUSED) {
            c2r_first_stage<C, IO_G_2ROWS>(a, sm, t, b, m, k, okp);
            if constexpr (C::L > 1) {
                run_stages<C, 1, IO_SMEM, IO_G_2ROWS>(a, sm, t, b, m, k, okp);
            }
            return;
        }
        constexpr int PAIRS = C::N / 2 + 1;
        constexpr int PCNT = (PAIRS + C::T - 1) / C::T;
        cx<T> avs[PCNT], bvs[PCNT];
        static_for<0, PCNT>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            const int i = t + C::T * c;
            avs[c] = cx<T>{T(0), T(0)};
            bvs[c] = cx<T>{T(0), T(0)};
            if (i < PAIRS && okp) {
                const u64 o = m + u64(i) * C::is1(a) + (2 * k) * C::is2(a);
                avs[c] = C::ld(a.in, o);
                if (2 * k + 1 < a.K) bvs[c] = C::ld(a.in, o + C::is2(a));
            }
        });
        static_for<0, PCNT>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            const int i = t + C::T * c;
            if (i < PAIRS) {
                cx<T> av = avs[c], bv = bvs[c];
                if (i == 0) {
                    av.y = T(0);
                    bv.y = T(0);
                }
                sm[G::soff(b, i)] = cx<T>{av.x - bv.y, av.y + bv.x};
                if (i != 0) sm[G::soff(b, C::N - i)] = cx<T>{av.x + bv.y, bv.x - av.y};
            }
        });
        BBK_SYNC();
        run_stages<C, 0, IO_SMEM, IO_G_2ROWS>(a, sm, t, b, m, k, okp);
    }
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization
// attribute may be scheduled while its predecessor on the stream is still running; it must not touch
// global memory before griddepcontrol.wait (= predecessor complete and flushed).  Letting the next
// launch's CTAs become resident early takes the launch latency (~1-2 us) off the critical path of
// back-to-back executes of small, L2-resident batches (BASELINE config 1).  Without the launch
// attribute both instructions are no-ops.
BBK_DEV void pdl_prologue() {
#ifndef BBFFT_EMU
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

template <class C> BBK_DEV void fft1d(args const &a) {
    pdl_prologue();
    fft1d_cta<C>(a, BBK_BID());
}

// ------------------------------------------------------------------------------------------
// Fused 2d c2c transform ("tile kernel").  One CTA owns one M x N1 x N2 tile -- contiguous in
// global memory for the default layout, element (m, n1, n2) at m + M*n1 + M*N1*n2 -- keeps it in
// shared memory across both passes and touches HBM exactly once on the way in and once on the
// way out.  Replaces two of the chained double-batched 1d launches of the reference's nd_fft
// (src/common/algorithm/nd_fft.hpp:66-112, 140-152) when the tile fits into shared memory.
//
// Traits contract (generated): real_t, DIR, THREADS, PADK, TILE_STRIDE (elements between tiles),
// ld/st hooks and two pass descriptors PA (axis n1) and PB (axis n2), each with
//   N  transform length           S  element stride of the axis inside the tile
//   O  repeats of the S*N block   L, radix(s), tw_off(s), WR<s>  as for the 1d kernel.
// Every stage deals its S*(N/R)*O sub-FFTs round-robin to all THREADS threads with the
// unit-stride index fastest, so global accesses coalesce and shared-memory accesses are (after
// the planner's padding search) conflict-free.  Stages are in place; the last stage of pass A
// writes through the digit reversal ("sorted") so that pass B sees natural order -- every thread
// holds its sub-FFTs in registers across one extra barrier instead of a second tile buffer.
// ------------------------------------------------------------------------------------------
enum : int { T_GLOBAL = 0, T_SMEM = 1, T_SMEM_SORTED = 2, T_DSMEM = 3, T_STAGED = 4, T_GLOBAL_REAL = 5 };

// Thread-block cluster primitives (sm_90+): the barrier over all CTAs of the cluster and a load from
// another CTA's shared memory (distributed shared memory) by shared-window offset.
#ifdef BBFFT_EMU
BBK_DEV void cluster_sync() { ::bbfft_emu::fail(); } // clusters are exercised on the GPU tier only
template <class SP> BBK_DEV auto ld_cluster(SP sm, int idx, unsigned) { return sm[idx]; }
#else
BBK_DEV void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
BBK_DEV cx<float> ld_cluster(const cx<float> *sm, int idx, unsigned rank) {
    const unsigned local = static_cast<unsigned>(__cvta_generic_to_shared(sm + idx));
    unsigned remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    cx<float> v;
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(remote) : "memory");
    return v;
}
BBK_DEV cx<double> ld_cluster(const cx<double> *sm, int idx, unsigned rank) {
    const unsigned local = static_cast<unsigned>(__cvta_generic_to_shared(sm + idx));
    unsigned remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    cx<double> v;
    asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(remote) : "memory");
    return v;
}
#endif

// Asynchronous global -> shared copies (cp.async, one element per instruction: the padded tile
// layout puts every element at its own 8- / 16-byte aligned slot).  The copy engine of the SM moves
// the data while the issuing warps go on computing; cp.async.wait_group + a CTA barrier publish it.
#ifdef BBFFT_EMU
// (emulator: the destination is poisoned at the issue and the data lands at the wait, tests/emu/cuda_emu.hpp)
template <class SP, class E> BBK_DEV void async_copy_elem(SP sm, int phys, const E *src) {
    ::bbfft_emu::async_issue_raw(::bbfft_emu::raw_ptr(sm) + phys, src, sizeof(E), false);
}
template <class SP, class E> BBK_DEV void async_copy_16(SP sm, int phys, const E *src) {
    ::bbfft_emu::async_issue_raw(::bbfft_emu::raw_ptr(sm) + phys, src, 16, false);
}
BBK_DEV void async_commit() {}
BBK_DEV void async_wait_all() { ::bbfft_emu::async_wait_raw(false); }
#else
template <class E> BBK_DEV void async_copy_elem(E *sm, int phys, const E *src) {
    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(sm + phys));
    if constexpr (sizeof(E) == 8) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
    } else {
        static_assert(sizeof(E) == 16, "complex<float> or complex<double>");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
}
// 16 bytes at once (two complex<float> or one complex<double>); both addresses 16-byte aligned
template <class E> BBK_DEV void async_copy_16(E *sm, int phys, const E *src) {
    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(sm + phys));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
BBK_DEV void async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
BBK_DEV void async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#endif

// Bulk asynchronous copies (cp.async.bulk, the TMA unit without a tensor map: SASS UBLKCP) completing on an
// mbarrier in shared memory.  One thread issues them; everybody waits on the barrier's phase.  The wait is
// bounded: a kernel that would spin forever traps instead.
#ifdef BBFFT_EMU
template <class SP> BBK_DEV void mbar_init(SP, int) {}
template <class SP> BBK_DEV void mbar_expect(SP, int, unsigned) {}
template <class SP, class E> BBK_DEV void bulk_copy(SP sm, int phys, const E *src, unsigned bytes, int) {
    ::bbfft_emu::async_issue_raw(::bbfft_emu::raw_ptr(sm) + phys, src, bytes, true);
}
template <class SP> BBK_DEV void mbar_wait(SP, int, unsigned) { ::bbfft_emu::async_wait_raw(true); }
#else
template <class E> BBK_DEV void mbar_init(E *sm, int bar) {
    const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(sm + bar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
template <class E> BBK_DEV void mbar_expect(E *sm, int bar, unsigned bytes) {
    const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(sm + bar));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
template <class E> BBK_DEV void bulk_copy(E *sm, int phys, const E *src, unsigned bytes, int bar) {
    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(sm + phys));
    const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(sm + bar));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(b)
                 : "memory");
}
template <class E> BBK_DEV void mbar_wait(E *sm, int bar, unsigned phase) {
    const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(sm + bar));
    unsigned done = 0;
    u64 t0 = 0;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(b), "r"(phase)
                     : "memory");
        if (done) break;
        u64 now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        if (now - t0 > 4000000000ull) __trap(); // four seconds: the copy this phase waits for was never issued
    }
}
#endif

template <class C> BBK_DEV int tile_phys(int lin) {
    if constexpr (C::PADK > 0) {
        return lin + lin / C::PADK;
    } else {
        return lin;
    }
}

// Offsets inside the padded tile that are compile-time constants.  A stage touches `R` elements at stride A
// from a base element; tile_phys(base + A j) - tile_phys(base) does not depend on the base when
//   * the tile is unpadded, or
//   * A is a multiple of PADK (every step crosses A / PADK pads), or
//   * A divides PADK and the base sits less than A behind a multiple of PADK (SPAN and SN, the periods of the
//     base in the stage's sub-FFT index and in the outer index, are multiples of PADK): j steps cross
//     j / (PADK / A) pads.
// Then one address per sub-FFT is computed and the rest are immediates of the LDS / STS instructions
// (a third of the tile kernel's instructions were such index arithmetic).
template <class C, int A, int SPAN, int SN> struct pad_rule {
    static constexpr int PK = C::PADK;
    static constexpr bool linear = PK == 0;
    static constexpr bool r1 = PK > 0 && A % (PK > 0 ? PK : 1) == 0;
    static constexpr bool r3 = PK > 0 && !r1 && PK % A == 0 && SPAN % (PK > 0 ? PK : 1) == 0 && SN % (PK > 0 ? PK : 1) == 0;
    static constexpr bool ok = linear || r1 || r3;
    static BBK_CE int off(int j) {
        return linear ? A * j : r1 ? A * j + A * j / (PK > 0 ? PK : 1) : A * j + j / ((PK > 0 ? PK : 1) / A);
    }
};

template <class P> BBK_CE int pass_ns(int s) {
    int n = P::N;
    for (int i = 0; i < s; ++i) n /= P::radix(i);
    return n;
}

// Real rows of the fused r2c / c2r tiles (C::REAL): complex element `pos` of the half-length transform of
// batch lane m is the pair of reals (2 pos, 2 pos + 1) of the row, M reals apart.  M == 1: the pair is one
// aligned complex word (the host checks the pointer, plan.cpp), otherwise two coalesced scalar accesses.
template <class C> BBK_DEV cx<typename C::real_t> ld_real_pair(const void *in, u64 row, int m, int pos) {
    using T = typename C::real_t;
    const T *BBK_RESTRICT x = reinterpret_cast<const T *>(in) + row;
    if constexpr (C::PA::S == 1) {
        return reinterpret_cast<const cx<T> *>(x)[pos];
    } else {
        return cx<T>{x[m + C::PA::S * (2 * pos)], x[m + C::PA::S * (2 * pos + 1)]};
    }
}
template <class C> BBK_DEV void st_real_pair(void *out, u64 row, int m, int pos, cx<typename C::real_t> v) {
    using T = typename C::real_t;
    T *BBK_RESTRICT x = reinterpret_cast<T *>(out) + row;
    if constexpr (C::PA::S == 1) {
        reinterpret_cast<cx<T> *>(x)[pos] = v;
    } else {
        x[m + C::PA::S * (2 * pos)] = v.x;
        x[m + C::PA::S * (2 * pos + 1)] = v.y;
    }
}

// `after_loads` runs when every thread of the CTA holds its inputs of the stage in registers; the
// persistent tile kernel passes a hook (preceded by a barrier) that starts the asynchronous load of
// the NEXT tile into the shared memory this tile no longer needs.
struct no_hook {
    static constexpr bool active = false;
    BBK_DEV void operator()() const {}
};

template <class C, class P, int S, int SRC, int DST, class HOOK = no_hook>
BBK_DEV void tile_stage(args const &a, BBK_SPTR(cx<typename C::real_t>) sm, u64 gbase, int tid, HOOK after_loads = HOOK{},
                        unsigned rank = 0) {
    using T = typename C::real_t;
    constexpr int R = P::radix(S);
    constexpr int NS = pass_ns<P>(S);
    constexpr int NS1 = NS / R;
    constexpr int NSUB = P::N / R;
    constexpr int TOTAL = P::S * NSUB * P::O;
    constexpr int CNT = (TOTAL + C::THREADS - 1) / C::THREADS;
    constexpr bool LAST = (S == P::L - 1);
    static_assert(DST == T_SMEM || LAST, "only a pass's last stage leaves the in-place scheme");
    static_assert(SRC != T_STAGED || P::PITCH == P::S * P::N, "the staging buffer holds packed rows");
    using WR = typename P::template WR<S>;
    const cx<T> *BBK_RESTRICT tw = reinterpret_cast<const cx<T> *>(a.tw) + P::tw_off(S);

    cx<T> v[CNT][R];
    static_for<0, CNT>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        const int id = tid + C::THREADS * i;
        if (TOTAL % C::THREADS == 0 || id < TOTAL) {
            const int lo = id % P::S, r = id / P::S;
            const int u = r % NSUB, hi = r / NSUB;
            const int pos0 = (u / NS1) * NS + u % NS1; // position along the axis of input j = 0
            const int base = lo + P::S * pos0 + P::PITCH * hi;
            using LR = pad_rule<C, P::S * NS1, P::S * NS, P::PITCH>;
            [[maybe_unused]] const int pb = tile_phys<C>(base);
            static_for<0, R>([&](auto jj) {
                constexpr int j = decltype(jj)::value;
                [[maybe_unused]] const int lin = base + P::S * NS1 * j;
                [[maybe_unused]] const int phys = LR::ok ? pb + LR::off(j) : tile_phys<C>(lin);
                if constexpr (SRC == T_GLOBAL) {
                    // (P::GS: element stride of the axis in global memory; differs from the shared-memory
                    // stride P::S only for the cluster kernel's column pass)
                    v[i][j] = C::ld(a.in, gbase + u64(lo + P::GS * (pos0 + NS1 * j) + P::GPITCH * hi));
                    if constexpr (C::REAL == 2) {
                        // c2r tile, column pass: the tile holds N1/2 columns; column 0 carries the spectrum columns
                        // 0 and N1/2 -- both transform to REAL columns -- as one complex column X0 + i XH
                        if (lo < C::PA::S) {
                            cx<T> xh = C::ld(a.in, gbase + u64(lo + C::PA::S * C::PA::N + P::GS * (pos0 + NS1 * j)));
                            // the four entries (0 | N1/2, 0 | N2/2) of a real signal's spectrum are real: their
                            // imaginary parts are ignored, like the 1d c2r ignores imag X[0] (reference
                            // test/r2c.cpp:310-324: "match the behaviour of other FFT libraries")
                            if (pos0 + NS1 * j == 0 || 2 * (pos0 + NS1 * j) == P::N) {
                                v[i][j].y = T(0);
                                xh.y = T(0);
                            }
                            v[i][j] = cx<T>{v[i][j].x - xh.y, v[i][j].y + xh.x};
                        }
                    }
                } else if constexpr (SRC == T_GLOBAL_REAL) {
                    // r2c tile, first stage of the half-length pass: complex element `pos` of a row is the pair of
                    // reals (2 pos, 2 pos + 1); `gbase` and C::RROW (row pitch) count reals
                    v[i][j] = ld_real_pair<C>(a.in, gbase + u64(C::RROW) * u64(hi), lo, pos0 + NS1 * j);
                } else if constexpr (SRC == T_DSMEM) {
                    // cluster kernel, first stage of the column pass: row n2 of the tile lives in the shared
                    // memory of CTA n2 / N2L in the row-pass layout  m + M (n1 + N1 n2l);  this CTA owns the
                    // columns n1 = rank * N1L + n1l, and lo = m + M n1l
                    const int n2 = pos0 + NS1 * j;
                    const int owner = n2 / C::N2L, n2l = n2 % C::N2L;
                    v[i][j] = ld_cluster(sm, tile_phys<C>(lo + int(rank) * P::S + C::ROWLEN * n2l), unsigned(owner));
                } else if constexpr (SRC == T_STAGED) {
                    // staged kernel, first stage of pass A: the leading C::STG elements of the tile wait in
                    // the (unpadded) staging buffer behind the tile, the rest already sits in place
                    v[i][j] = lin < C::STG ? sm[C::STG_OFF + lin] : sm[phys];
                } else {
                    v[i][j] = sm[phys];
                }
            });
        }
    });
    if constexpr (DST == T_SMEM_SORTED && SRC != T_GLOBAL) {
        BBK_SYNC(); // every read of the stage happens before its out-of-place writes
    }
    if constexpr (SRC == T_DSMEM) {
        cluster_sync(); // every CTA of the cluster has gathered its columns: the row-pass layout is dead everywhere
    }
    if constexpr (HOOK::active) {
        static_assert((SRC == T_SMEM && DST == T_GLOBAL) || SRC == T_STAGED,
                      "the hook belongs to a stage that empties (a part of) shared memory");
        BBK_SYNC(); // the tile (or the staging buffer) has left shared memory
        after_loads();
    }
    static_for<0, CNT>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        const int id = tid + C::THREADS * i;
        if (TOTAL % C::THREADS == 0 || id < TOTAL) {
            const int lo = id % P::S, r = id / P::S;
            const int u = r % NSUB, hi = r / NSUB;
            reg_fft<T, WR, R, C::DIR>::run(v[i]);
            if constexpr (!LAST) {
                const int n2 = u % NS1;
                static_for<1, R>([&](auto qq) {
                    constexpr int q = decltype(qq)::value;
                    v[i][q] = cmul(v[i][q], ldg_cx(tw + (q - 1) * NS1 + n2));
                });
            }
            if constexpr (DST == T_SMEM) {
                const int base = lo + P::S * ((u / NS1) * NS + u % NS1) + P::PITCH * hi;
                using SR = pad_rule<C, P::S * NS1, P::S * NS, P::PITCH>;
                const int pb = tile_phys<C>(base);
                static_for<0, R>([&](auto qq) {
                    constexpr int q = decltype(qq)::value;
                    sm[SR::ok ? pb + SR::off(q) : tile_phys<C>(base + P::S * NS1 * q)] = v[i][q];
                });
            } else {
                const int bin0 = bin_of_sub<P>(u);
                [[maybe_unused]] const int base = lo + P::S * bin0 + P::PITCH * hi;
                // (bin0 < N / R: the base sits less than the stride S N / R behind a multiple of the row pitch)
                using SR = pad_rule<C, P::S * (P::N / R), C::PADK, P::PITCH>;
                [[maybe_unused]] const int pb = (DST == T_GLOBAL || DST == T_GLOBAL_REAL) ? 0 : tile_phys<C>(base);
                static_for<0, R>([&](auto qq) {
                    constexpr int q = decltype(qq)::value;
                    if constexpr (DST == T_GLOBAL) {
                        if constexpr (C::REAL == 1) {
                            // r2c tile: column 0 is the packed pair of the real columns 0 and N1/2; its transform is
                            // unpacked from the scratch column behind the tile (tile_r2c_unpack)
                            if (lo < C::PA::S) {
                                sm[C::SCR_OFF + lo + C::PA::S * (bin0 + (P::N / R) * q)] = v[i][q];
                            } else {
                                C::st(a.out, gbase + u64(lo + P::GS * (bin0 + (P::N / R) * q)), v[i][q]);
                            }
                        } else {
                            C::st(a.out, gbase + u64(lo + P::GS * (bin0 + (P::N / R) * q) + P::GPITCH * hi), v[i][q]);
                        }
                    } else if constexpr (DST == T_GLOBAL_REAL) {
                        st_real_pair<C>(a.out, gbase + u64(C::RROW) * u64(hi), lo, bin0 + (P::N / R) * q, v[i][q]);
                    } else {
                        sm[SR::ok ? pb + SR::off(q) : tile_phys<C>(base + P::S * (P::N / R) * q)] = v[i][q];
                    }
                });
            }
        }
    });
}

template <class C, class P, int S, int SRC0, int DSTL, class HOOK = no_hook>
BBK_DEV void tile_pass(args const &a, BBK_SPTR(cx<typename C::real_t>) sm, u64 gbase, int tid, HOOK last_hook = HOOK{},
                       unsigned rank = 0) {
    if constexpr (S < P::L) {
        constexpr int SRC = (S == 0) ? SRC0 : T_SMEM;
        constexpr int DST = (S == P::L - 1) ? DSTL : T_SMEM;
        if constexpr (S > 0) {
            BBK_SYNC();
        }
        if constexpr (S == P::L - 1 && HOOK::active) {
            tile_stage<C, P, S, SRC, DST, HOOK>(a, sm, gbase, tid, last_hook, rank);
        } else {
            tile_stage<C, P, S, SRC, DST>(a, sm, gbase, tid, no_hook{}, rank);
        }
        tile_pass<C, P, S + 1, SRC0, DSTL, HOOK>(a, sm, gbase, tid, last_hook, rank);
    }
}

template <class C> BBK_DEV void fft2d_tile_cta(args const &a, const u64 tile) {
    using T = typename C::real_t;
    BBK_SPTR(cx<T>) sm = sptr<cx<T>>(BBK_SMEM());
    const int tid = BBK_TID();
    const u64 gbase = tile * u64(C::TILE_STRIDE);
    if (a.pf != 0 && tile + a.pf < a.K) {
        // the tile that runs `pf` tiles later (same reasoning as prefetch_future_batch: one 128 KiB tile
        // per SM leaves no second CTA to overlap with; the L2 takes the role of the second buffer)
        constexpr u64 TB = u64(C::TILE_STRIDE) * 2 * sizeof(T);
        l2_prefetch_range(a.in, (tile + a.pf) * TB, (tile + a.pf + 1) * TB, tid, C::THREADS);
    }
    tile_pass<C, typename C::PA, 0, T_GLOBAL, T_SMEM_SORTED>(a, sm, gbase, tid);
    BBK_SYNC();
    tile_pass<C, typename C::PB, 0, T_SMEM, T_GLOBAL>(a, sm, gbase, tid);
}

// Persistent variant with an asynchronous tile pipeline (C::PERSIST; grid = resident CTAs).  A CTA
// walks the tiles bid, bid + grid, ...  The last stage of pass B pulls its inputs into registers;
// from that barrier on the shared-memory tile is dead, so the CTA issues the cp.async copies of its
// NEXT tile and only then computes the last butterflies and stores the finished tile: the HBM read
// of tile i+1 overlaps the arithmetic and the HBM write of tile i, inside one CTA -- which is what a
// 128 KiB tile (one CTA per SM, nobody else to overlap with) was missing in round 1 (0.65 of the
// HBM peak for 2d fp32 128 x 128).  Pass A then starts from shared memory instead of global memory.
template <class C> struct tile_loader {
    static constexpr bool active = true;
    using T = typename C::real_t;
    args const &a;
    BBK_SPTR(cx<T>) sm;
    u64 tile; // the tile to fetch (>= a.K: nothing)
    int tid;
    BBK_DEV void operator()() const {
        if (tile < a.K) {
            const cx<T> *BBK_RESTRICT src = reinterpret_cast<const cx<T> *>(a.in) + tile * u64(C::TILE_STRIDE);
            constexpr int TILE = C::PA::S * C::PA::N * C::PA::O;
            for (int lin = tid; lin < TILE; lin += C::THREADS) async_copy_elem(sm, tile_phys<C>(lin), src + lin);
        }
        async_commit();
    }
};

template <class C> BBK_DEV void fft2d_tile_persistent(args const &a) {
    using T = typename C::real_t;
    BBK_SPTR(cx<T>) sm = sptr<cx<T>>(BBK_SMEM());
    const int tid = BBK_TID();
    const u64 step = BBK_NCTAS();
    u64 tile = BBK_BID();
    tile_loader<C>{a, sm, tile, tid}();
    for (; tile < a.K; tile += step) {
        const u64 gbase = tile * u64(C::TILE_STRIDE);
        async_wait_all();
        BBK_SYNC(); // every thread's copies have landed
        tile_pass<C, typename C::PA, 0, T_SMEM, T_SMEM_SORTED>(a, sm, gbase, tid);
        BBK_SYNC();
        tile_pass<C, typename C::PB, 0, T_SMEM, T_GLOBAL, tile_loader<C>>(a, sm, gbase, tid,
                                                                            tile_loader<C>{a, sm, tile + step, tid});
    }
}

// Staged variant of the persistent kernel (C::STG > 0; grid = resident CTAs).  The pipeline above can only
// start the next tile's copies when the current tile has left shared memory, i.e. it hides them behind ONE
// stage; through the other stages of a 128 KiB tile (one CTA per SM) nothing is in flight, and those stages
// are bound by shared-memory bandwidth (six trips of the tile = 3.3 us against 5.9 us of HBM time per tile
// and SM).  Here the shared memory left beside the tile is a staging buffer for the leading C::STG elements
// (whole rows of pass A) of the NEXT tile: their copies are issued as soon as the first stage of pass A has
// pulled the current tile's share out of the buffer, and land while ALL remaining stages run; only the
// rest (a quarter of a 128 x 128 fp32 tile) waits, as before, for the last stage of pass B.  The staging
// buffer is unpadded (the first stage reads it with unit stride along the lanes), so its copies move 16 bytes.
template <class C> struct stage_loader {
    static constexpr bool active = true;
    using T = typename C::real_t;
    args const &a;
    BBK_SPTR(cx<T>) sm;
    u64 tile; // the tile to fetch (>= a.K: nothing)
    int tid;
    bool bulk; // C::BULK and the tiles are 16-byte aligned: one thread hands the copy to the TMA unit
    BBK_DEV void operator()() const {
        if (tile < a.K) {
            const cx<T> *BBK_RESTRICT src = reinterpret_cast<const cx<T> *>(a.in) + tile * u64(C::TILE_STRIDE);
            constexpr int PER = 16 / int(sizeof(cx<T>)); // elements per 16-byte copy
            if (bulk) {
                if (tid == 0) {
                    constexpr int CHUNK = 8192 / int(sizeof(cx<T>)); // elements per bulk copy
                    mbar_expect(sm, C::STG_OFF + C::STG, unsigned(C::STG * sizeof(cx<T>)));
                    for (int lin = 0; lin < C::STG; lin += CHUNK) {
                        const int cnt = C::STG - lin < CHUNK ? C::STG - lin : CHUNK;
                        bulk_copy(sm, C::STG_OFF + lin, src + lin, unsigned(cnt * sizeof(cx<T>)), C::STG_OFF + C::STG);
                    }
                }
            } else if (PER == 1 || (reinterpret_cast<u64>(src) & 15) == 0) {
                for (int lin = tid * PER; lin < C::STG; lin += C::THREADS * PER) async_copy_16(sm, C::STG_OFF + lin, src + lin);
            } else {
                for (int lin = tid; lin < C::STG; lin += C::THREADS) async_copy_elem(sm, C::STG_OFF + lin, src + lin);
            }
        }
        async_commit();
    }
};
template <class C> struct rest_loader {
    static constexpr bool active = true;
    using T = typename C::real_t;
    args const &a;
    BBK_SPTR(cx<T>) sm;
    u64 tile;
    int tid;
    BBK_DEV void operator()() const {
        if (tile < a.K) {
            const cx<T> *BBK_RESTRICT src = reinterpret_cast<const cx<T> *>(a.in) + tile * u64(C::TILE_STRIDE);
            constexpr int TILE = C::PA::S * C::PA::N * C::PA::O;
            for (int lin = C::STG + tid; lin < TILE; lin += C::THREADS) async_copy_elem(sm, tile_phys<C>(lin), src + lin);
        }
        async_commit();
    }
};

template <class C> BBK_DEV void fft2d_tile_staged(args const &a) {
    using T = typename C::real_t;
    using PA = typename C::PA;
    static_assert(C::STG % 2 == 0 && C::STG_OFF % 2 == 0, "16-byte copies into the staging buffer");
    BBK_SPTR(cx<T>) sm = sptr<cx<T>>(BBK_SMEM());
    const int tid = BBK_TID();
    const u64 step = BBK_NCTAS();
    u64 tile = BBK_BID();
    // (C::BULK: the mbarrier sits behind the staging buffer)
    const bool bulk = C::BULK && (reinterpret_cast<u64>(a.in) & 15) == 0 && (u64(C::TILE_STRIDE) * sizeof(cx<T>)) % 16 == 0;
    unsigned phase = 0;
    if (bulk) {
        if (tid == 0) mbar_init(sm, C::STG_OFF + C::STG);
        BBK_SYNC();
    }
    stage_loader<C>{a, sm, tile, tid, bulk}();
    rest_loader<C>{a, sm, tile, tid}();
    for (; tile < a.K; tile += step) {
        const u64 gbase = tile * u64(C::TILE_STRIDE);
        async_wait_all();
        if (bulk) {
            mbar_wait(sm, C::STG_OFF + C::STG, phase);
            phase ^= 1u;
        }
        BBK_SYNC(); // every thread's copies have landed
        constexpr int DST0 = PA::L == 1 ? T_SMEM_SORTED : T_SMEM;
        tile_stage<C, PA, 0, T_STAGED, DST0, stage_loader<C>>(a, sm, gbase, tid, stage_loader<C>{a, sm, tile + step, tid, bulk});
        tile_pass<C, PA, 1, T_SMEM, T_SMEM_SORTED>(a, sm, gbase, tid);
        BBK_SYNC();
        tile_pass<C, typename C::PB, 0, T_SMEM, T_GLOBAL, rest_loader<C>>(a, sm, gbase, tid,
                                                                          rest_loader<C>{a, sm, tile + step, tid});
    }
}

// Fused real 2d tiles (C::REAL = 1: r2c, 2: c2r; even N1).  The reference runs a real nd transform as one
// double-batched launch per mode (src/common/algorithm/nd_fft.hpp:66-152: r2c along n1, then c2c along n2 over
// the N1/2+1 spectrum columns): two HBM round trips.  Here one CTA owns one transform and its tile holds
// H = N1/2 complex columns:
//   r2c: the real rows are read as H complex words per row and transformed along n1 (half-length pass A); the
//        split  X[i] = A + B, X[H-i] = conj(A - B)  (the arithmetic of the 1d kernel's r2c_last_stage) runs in
//        place on the pairs (i, H - i); the columns 0 and H of the n1-spectrum are REAL, so they share tile
//        column 0 as X0 + i XH; pass B transforms the H columns along n2 and stores columns 1 .. H-1; the
//        transform of column 0 goes to a scratch column and is unpacked by conjugate symmetry,
//        F0[k] = (Z[k] + conj Z[N2-k]) / 2, FH[k] = (Z[k] - conj Z[N2-k]) / 2i, into the spectrum columns 0 and H;
//   c2r: the mirror image -- pass B loads column 0 as X0 + i XH (both are spectra of real columns), the merge
//        z[i] = A + B, z[H-i] = conj(A - B) runs on the pairs, half-length pass A stores real pairs.  Like every
//        nd c2r (cuFFT, MKL) this assumes the conjugate-even spectrum of a real signal.
// The tile is a power-of-two H x N2 block for power-of-two sizes (no ragged N1/2+1 columns), HBM traffic is one
// read of the real tile and one write of the spectrum tile (or the reverse), and in place works because every
// global load of the CTA precedes its first global store (barriers in between).
template <class C, bool FORWARD_SPLIT> BBK_DEV void tile_real_pairs(args const &a, BBK_SPTR(cx<typename C::real_t>) sm, int tid) {
    using T = typename C::real_t;
    using PA = typename C::PA;
    constexpr int H = PA::N;         // half length: complex columns of the tile
    constexpr int M = PA::S;
    constexpr int UNITS = H / 2;     // pairs (i, H - i), i = 1 .. H/2 (i = H/2 pairs with itself when H is even)
    constexpr int TOTAL = M * UNITS * PA::O;
    constexpr int CNT = (TOTAL + C::THREADS - 1) / C::THREADS;
    const cx<T> *BBK_RESTRICT twr = reinterpret_cast<const cx<T> *>(a.tw) + C::TW_REAL;
    // (a power-of-two number of units per row keeps a warp inside one row: the version that walked the H/2+1
    // units i = 0 .. H/2 of a row wrapped every warp over two rows and paid two-way bank conflicts on every
    // access, profiles/r02zb_ncu_r2c_tile.txt: 2d r2c fp32 128 x 128 297 -> 264 us.  The accesses to element i
    // still cross a pad once per half-warp (13 % excess wavefronts, r02zd_ncu_r2c_tile.txt); keeping the group
    // leaders i = 16 k out of the main loop removes that but costs a second loop and measured slower on
    // balance, profiles/r02ze_real_tiles_leaderless.log)
    static_for<0, CNT>([&](auto ii) {
        constexpr int c = decltype(ii)::value;
        const int id = tid + C::THREADS * c;
        if (TOTAL % C::THREADS == 0 || id < TOTAL) {
            const int lo = id % M, r = id / M;
            const int i = 1 + r % UNITS, hi = r / UNITS;
            const int row = lo + PA::PITCH * hi;
            const int pi = tile_phys<C>(row + M * i), pn = tile_phys<C>(row + M * (H - i));
            const cx<T> w = ldg_cx(twr + i);
            const cx<T> iw = cx<T>{-w.y, w.x};
            const cx<T> yi = sm[pi];
            const cx<T> yn = sm[pn];
            if constexpr (FORWARD_SPLIT) {
                const cx<T> y2 = conj(yn);
                const cx<T> aa = rmul(y2 + yi, T(0.5));
                const cx<T> bb = cmul(rmul(y2 - yi, T(0.5)), iw);
                sm[pi] = aa + bb;
                if (2 * i != H) sm[pn] = conj(aa - bb);
            } else {
                const cx<T> x2 = conj(yn);
                const cx<T> aa = yi + x2;
                const cx<T> bb = cmul(yi - x2, iw);
                sm[pi] = aa + bb;
                if (2 * i != H) sm[pn] = conj(aa - bb);
            }
        }
    });
    // column 0 -- r2c: y0 = (a, b) -> X[0] = a + b, X[H] = a - b (both real), kept as one complex number;
    // c2r: (X0, XH) -> z0 = (X0 + XH, X0 - XH): the same map
    for (int id = tid; id < M * PA::O; id += C::THREADS) {
        const int p0 = tile_phys<C>(id % M + PA::PITCH * (id / M));
        const cx<T> y = sm[p0];
        sm[p0] = cx<T>{y.x + y.y, y.x - y.y};
    }
}

// r2c: the transform Z of the packed column X0 + i XH sits in the scratch column (natural order); unpack it into
// the spectrum columns 0 and H of the output tile.
template <class C> BBK_DEV void tile_r2c_unpack(args const &a, BBK_SPTR(cx<typename C::real_t>) sm, u64 cbase, int tid) {
    using T = typename C::real_t;
    constexpr int M = C::PA::S, H = C::PA::N, N2 = C::PB::N;
    constexpr int UNITS = N2 / 2 + 1; // pairs (k, N2 - k)
    constexpr int TOTAL = M * UNITS;
    for (int id = tid; id < TOTAL; id += C::THREADS) {
        const int lo = id % M, k = id / M;
        const int kn = (N2 - k) % N2;
        const cx<T> zk = sm[C::SCR_OFF + lo + M * k];
        const cx<T> zr = sm[C::SCR_OFF + lo + M * kn];
        const cx<T> zn = conj(zr);
        const cx<T> f0 = rmul(zk + zn, T(0.5));
        const cx<T> d = rmul(zk - zn, T(0.5));
        const cx<T> fh = cx<T>{d.y, -d.x}; // d / i
        const u64 o = cbase + u64(lo) + u64(C::PB::GS) * u64(k);
        C::st(a.out, o, f0);
        C::st(a.out, o + u64(M * H), fh);
        if (kn != k) {
            const u64 on = cbase + u64(lo) + u64(C::PB::GS) * u64(kn);
            C::st(a.out, on, conj(f0));
            C::st(a.out, on + u64(M * H), conj(fh));
        }
    }
}

template <class C> BBK_DEV void fft2d_tile_real_cta(args const &a, const u64 tile) {
    using T = typename C::real_t;
    BBK_SPTR(cx<T>) sm = sptr<cx<T>>(BBK_SMEM());
    const int tid = BBK_TID();
    const u64 cbase = tile * u64(C::TILE_STRIDE); // spectrum tile, complex elements
    const u64 rbase = tile * u64(C::RTS);         // real tile, reals
    if constexpr (C::REAL == 1) {
        tile_pass<C, typename C::PA, 0, T_GLOBAL_REAL, T_SMEM_SORTED>(a, sm, rbase, tid);
        BBK_SYNC();
        tile_real_pairs<C, true>(a, sm, tid);
        BBK_SYNC();
        tile_pass<C, typename C::PB, 0, T_SMEM, T_GLOBAL>(a, sm, cbase, tid);
        BBK_SYNC();
        tile_r2c_unpack<C>(a, sm, cbase, tid);
    } else {
        tile_pass<C, typename C::PB, 0, T_GLOBAL, T_SMEM_SORTED>(a, sm, cbase, tid);
        BBK_SYNC();
        tile_real_pairs<C, false>(a, sm, tid);
        BBK_SYNC();
        if constexpr (C::PA::S == 1) {
            // M = 1: the last stage of pass A holds digit-reversed positions of a row, i.e. scattered 8-byte stores;
            // sort into the tile instead and copy the rows out with consecutive threads on consecutive words
            tile_pass<C, typename C::PA, 0, T_SMEM, T_SMEM_SORTED>(a, sm, rbase, tid);
            BBK_SYNC();
            using PA = typename C::PA;
            constexpr int TILE = PA::N * PA::O;
            for (int lin = tid; lin < TILE; lin += C::THREADS) {
                const int pos = lin % PA::N, hi = lin / PA::N;
                st_real_pair<C>(a.out, rbase + u64(C::RROW) * u64(hi), 0, pos, sm[tile_phys<C>(pos + PA::PITCH * hi)]);
            }
        } else {
            tile_pass<C, typename C::PA, 0, T_SMEM, T_GLOBAL_REAL>(a, sm, rbase, tid);
        }
    }
}

// Cluster variant (C::CL > 1 CTAs per tile, thread-block cluster of CL, distributed shared memory).
// A 128 x 128 fp32 tile is 128 KiB: one CTA per SM, and a stage-synchronised CTA that is alone on its SM
// overlaps nothing with its own load latency (round 1: 0.65 of the HBM peak, against 0.96 for the 32 KiB
// 64 x 64 tiles that run four to an SM).  Here CTA `rank` of the cluster holds N2 / CL ROWS of the tile
// (32 KiB for CL = 4), so CTAs of several tiles share an SM again.  Row pass: local.  Column pass: CTA
// `rank` takes the columns n1 in [rank N1/CL, (rank+1) N1/CL); the first stage pulls its inputs straight
// out of the owners' shared memory (ld.shared::cluster), one cluster barrier later every CTA reuses its
// own shared memory for the rest of the column pass, and the last stage stores N1/CL-element runs of
// every row to global memory.  HBM traffic stays one read and one write of the tile.
template <class C> BBK_DEV void fft2d_tile_cluster(args const &a) {
    using T = typename C::real_t;
    using PA = typename C::PA;
    BBK_SPTR(cx<T>) sm = sptr<cx<T>>(BBK_SMEM());
    const int tid = BBK_TID();
    const u64 tile = BBK_BID() / u64(C::CL);
    const unsigned rank = unsigned(BBK_BID() % u64(C::CL));
    const u64 gtile = tile * u64(C::TILE_STRIDE);
    // rows n2 in [rank N2L, (rank+1) N2L): a contiguous slab of the tile
    tile_pass<C, PA, 0, T_GLOBAL, T_SMEM_SORTED>(a, sm, gtile + u64(rank) * u64(PA::S * PA::N * PA::O), tid);
    cluster_sync(); // every CTA's rows are transformed and visible to the cluster
    // columns: lo = m + M n1l runs over this CTA's PB::S = M N1L fastest positions; in global memory they
    // start at rank * PB::S inside every row of the tile
    tile_pass<C, typename C::PB, 0, T_DSMEM, T_GLOBAL>(a, sm, gtile + u64(rank) * u64(C::PB::S), tid, no_hook{}, rank);
}

template <class C> BBK_DEV void fft2d_tile(args const &a) {
    pdl_prologue();
    if constexpr (C::REAL != 0) {
        const u64 tile = BBK_BID();
        if (tile >= a.K) return;
        fft2d_tile_real_cta<C>(a, tile);
    } else if constexpr (C::CL > 1) {
        fft2d_tile_cluster<C>(a);
    } else if constexpr (C::STG > 0) {
        fft2d_tile_staged<C>(a);
    } else if constexpr (C::PERSIST) {
        fft2d_tile_persistent<C>(a);
    } else {
        const u64 tile = BBK_BID();
        if (tile >= a.K) return;
        fft2d_tile_cta<C>(a, tile);
    }
}

// ------------------------------------------------------------------------------------------
// Chained nd transform: ONE persistent launch runs every step of a 2d/3d decomposition (fused
// tile kernel and/or double-batched 1d passes; reference nd_fft: one launch per mode,
// src/common/algorithm/nd_fft.hpp:140-152).  The outer batch index k is the slowest index of
// every step, so step d of slab k depends only on step d-1 of slab k.  Work items (step, k, CTA
// of that step's grid inside slab k) are ordered in pipeline slots -- slot s holds step d of
// k-block s-d -- and dealt round-robin to the resident CTAs.  A block of slabs is a few MiB, so
// what step d-1 wrote is still in the 126 MB L2 when step d reads it one slot later: HBM sees
// one read and one write of the tensor instead of one round trip per step, without the drain at
// every kernel boundary that made host-side L2 blocking lose (profiles/r01c_nd_block.txt).
//
// Dependencies: done[d*K + k] counts finished CTAs of step d, slab k (monotonic over launches:
// launch number `epoch` waits for epoch * PER_K).  A waiting item only ever waits for items that
// precede it in the item order; every CTA walks its items in that order and the whole grid is
// resident (grid <= SMs * occupancy, chosen by the host), so the earliest unfinished item is
// always running: no deadlock.  Steps d > 0 read what other SMs wrote during this launch, so
// their stubs load with ld.global.cg (L2 only; L1 is not coherent).
// ------------------------------------------------------------------------------------------
struct chain_args {
    args step[3];
    u64 *done;  // [steps][K]
    u64 epoch;  // 1, 2, 3, ... one per launch of this plan
    u64 K;      // outer batch (slabs)
    u64 kblock; // slabs per pipeline block
};

enum : int { STEP_FFT1D = 0, STEP_TILE = 1 };

#ifdef BBFFT_EMU
BBK_DEV void chain_wait(u64 const *p, u64 target) {
    if (*p < target) ::bbfft_emu::fail(); // one emulated CTA runs the items in order: never waits
}
BBK_DEV void chain_signal(u64 *p) { *p += 1; }
#else
BBK_DEV void chain_wait(u64 const *p, u64 target) {
    const volatile u64 *vp = p;
    long long t0 = 0;
    unsigned spins = 0;
    while (*vp < target) {
        __nanosleep(100);
        if ((++spins & 1023u) == 0) {
            // a dependency that does not arrive within seconds is a bug: fault instead of hanging the GPU
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (now - t0 > 8000000000ll) __trap();
        }
    }
    __threadfence();
}
BBK_DEV void chain_signal(u64 *p) {
    __threadfence();
    atomicAdd(reinterpret_cast<unsigned long long *>(p), 1ull);
}
#endif

template <class S> BBK_DEV void chain_run_step(args const &a, u64 bid) {
    if constexpr (S::KIND == STEP_TILE) {
        fft2d_tile_cta<typename S::C>(a, bid);
    } else {
        fft1d_cta<typename S::C>(a, bid);
    }
}

// S0, S1, S2: step descriptors {C, KIND, PER_K = CTAs of the step per slab}; NS of them are used.
template <int NS, class S0, class S1, class S2> BBK_DEV void chain(chain_args const &ca) {
    static_assert(NS == 2 || NS == 3, "a chain has two or three steps");
    // (a function, not an array: a dynamically indexed array would live in local memory)
    auto per_k_of = [](int d) { return d == 0 ? u64(S0::PER_K) : (d == 1 ? u64(S1::PER_K) : u64(S2::PER_K)); };
    const u64 nblk = (ca.K + ca.kblock - 1) / ca.kblock;
    const u64 nslots = nblk + NS - 1;
    auto blk_slabs = [&](u64 j) { return (j + 1) * ca.kblock <= ca.K ? ca.kblock : ca.K - j * ca.kblock; };
    auto slot_items = [&](u64 sl) {
        u64 n = 0;
        for (int d = 0; d < NS; ++d) {
            if (sl >= u64(d) && sl - d < nblk) n += blk_slabs(sl - d) * per_k_of(d);
        }
        return n;
    };
    const int tid = BBK_TID();
    u64 slot = 0, base = 0, cur = slot_items(0);
    for (u64 it = BBK_BID();; it += BBK_NCTAS()) {
        while (slot < nslots && it >= base + cur) {
            base += cur;
            ++slot;
            cur = slot < nslots ? slot_items(slot) : 0;
        }
        if (slot >= nslots) break;
        u64 r = it - base;
        int d = 0;
        u64 j = 0;
        for (; d < NS; ++d) {
            if (slot >= u64(d) && slot - d < nblk) {
                const u64 n = blk_slabs(slot - d) * per_k_of(d);
                if (r < n) {
                    j = slot - d;
                    break;
                }
                r -= n;
            }
        }
        const u64 pk = per_k_of(d);
        const u64 k = j * ca.kblock + r / pk;
        const u64 bid = k * pk + r % pk;
        if (d > 0 && tid == 0) chain_wait(ca.done + u64(d - 1) * ca.K + k, ca.epoch * per_k_of(d - 1));
        BBK_SYNC(); // dependency visible to the CTA; shared memory of the previous item is free
        if (d == 0) {
            chain_run_step<S0>(ca.step[0], bid);
        } else if (d == 1) {
            chain_run_step<S1>(ca.step[1], bid);
        } else {
            if constexpr (NS == 3) chain_run_step<S2>(ca.step[2], bid);
        }
        if (d < NS - 1) {
            BBK_SYNC(); // every store of the CTA is issued before the release below
            if (tid == 0) chain_signal(ca.done + u64(d) * ca.K + k);
        }
    }
}

} // namespace bbk

#endif // BBFFT_KERNELS_CUH
