// oracle/refemu/ocl_shim.hpp -- TEST INFRASTRUCTURE ONLY.
//
// Minimal OpenCL-C execution environment for the kernels that the UNMODIFIED reference
// generator emits (reference: src/base/generator/sbfft_gen.cpp, f2fft_gen.cpp).  A kernel,
// after the textual rewrite done in refemu.cpp, is compiled as C++ against this header and
// run one fiber per work-item; `barrier` yields to the work-group scheduler in refemu.cpp.
// Only the tiny OpenCL-C subset the generator uses is provided.
#ifndef BBFFT_ORACLE_OCL_SHIM_HPP
#define BBFFT_ORACLE_OCL_SHIM_HPP

#include <cstddef>
#include <cstdint>

typedef unsigned long ulong;
typedef unsigned int uint;
typedef unsigned short ushort;

// Work-item context; owned by the scheduler (refemu.cpp), one per fiber.
struct refemu_wi_ctx {
    std::size_t local_id[3];
    std::size_t group_id[3];
    std::size_t local_size[3];
    std::size_t num_groups[3];
    char *local_base;        // work-group local-memory arena (shared by all work-items)
    std::size_t local_used;  // per-work-item bump pointer (identical sequence in every item)
    std::size_t local_cap;
    void (*yield)(refemu_wi_ctx *); // work-group barrier: switch back to the scheduler
};

static thread_local refemu_wi_ctx *refemu_ctx = nullptr;

static inline std::size_t get_local_id(uint d) { return refemu_ctx->local_id[d]; }
static inline std::size_t get_group_id(uint d) { return refemu_ctx->group_id[d]; }
static inline std::size_t get_local_size(uint d) { return refemu_ctx->local_size[d]; }
static inline std::size_t get_num_groups(uint d) { return refemu_ctx->num_groups[d]; }
static inline std::size_t get_global_id(uint d) {
    return refemu_ctx->group_id[d] * refemu_ctx->local_size[d] + refemu_ctx->local_id[d];
}
static inline std::size_t get_global_size(uint d) {
    return refemu_ctx->num_groups[d] * refemu_ctx->local_size[d];
}

#define CLK_LOCAL_MEM_FENCE 1
#define CLK_GLOBAL_MEM_FENCE 2
static inline void barrier(int) {
    refemu_wi_ctx *c = refemu_ctx;
    c->yield(c);
    refemu_ctx = c; // another fiber of this OS thread ran in between
}
// Sub-group barriers only occur in straight-line code (sbfft_gen.cpp:289); a work-group
// barrier is a safe superset.
static inline void sub_group_barrier(int f) { barrier(f); }

static inline void *wg_local(std::size_t bytes) {
    refemu_wi_ctx *c = refemu_ctx;
    std::size_t off = (c->local_used + 15u) & ~std::size_t(15u);
    c->local_used = off + bytes;
    if (c->local_used > c->local_cap) {
        __builtin_trap();
    }
    return c->local_base + off;
}

template <typename T> struct vec2 {
    T x, y;
    vec2() = default;
    vec2(T x_, T y_) : x(x_), y(y_) {}
    vec2(int v) : x(T(v)), y(T(v)) {}
};
template <typename T> inline vec2<T> operator+(vec2<T> a, vec2<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> inline vec2<T> operator-(vec2<T> a, vec2<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> inline vec2<T> operator-(vec2<T> a) { return {-a.x, -a.y}; }
template <typename T> inline vec2<T> operator*(vec2<T> a, vec2<T> b) { return {a.x * b.x, a.y * b.y}; }
template <typename T> inline vec2<T> operator*(vec2<T> a, T s) { return {a.x * s, a.y * s}; }
template <typename T> inline vec2<T> operator*(T s, vec2<T> a) { return {s * a.x, s * a.y}; }
template <typename T> inline vec2<T> operator/(vec2<T> a, T s) { return {a.x / s, a.y / s}; }
template <typename T> inline vec2<T> &operator+=(vec2<T> &a, vec2<T> b) { a.x += b.x; a.y += b.y; return a; }
template <typename T> inline vec2<T> &operator-=(vec2<T> &a, vec2<T> b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename T> inline vec2<T> &operator*=(vec2<T> &a, T s) { a.x *= s; a.y *= s; return a; }

typedef vec2<float> float2;
typedef vec2<double> double2;
static inline float2 mk_float2(float a, float b) { return float2(a, b); }
static inline double2 mk_double2(double a, double b) { return double2(a, b); }

// OpenCL scalar select(a, b, c) = c ? b : a
template <typename T, typename U, typename C> static inline T select(T a, U b, C c) { return c ? T(b) : a; }

#endif // BBFFT_ORACLE_OCL_SHIM_HPP
