// tests/cpp/test_tensor_indexer.cpp -- host test of include/bbfft/tensor_indexer.hpp.
// The expectations are the reference's own (test/tensor.cpp:25-155): packed / strided addressing,
// may_fuse / fused, may_reshape_mode / reshaped_mode, for both storage orders.  Built and run by
// tests/test_host.py (no GPU needed).
#include "bbfft/tensor_indexer.hpp"

#include <array>
#include <cstddef>
#include <cstdio>

using namespace bbfft;

static int failures = 0;
#define EXPECT(cond)                                                                               \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            ++failures;                                                                            \
            std::printf("EXPECT failed %s:%d: %s\n", __FILE__, __LINE__, #cond);                   \
        }                                                                                          \
    } while (0)

using S = std::size_t;
template <std::size_t D> using arr = std::array<S, D>;

static void col_major() {
    {
        auto x = tensor_indexer<S, 3u, layout::col_major>({3, 4, 5});
        EXPECT(x(0, 0, 0) == 0 && x(1, 0, 0) == 1 && x(0, 1, 0) == 3 && x(0, 0, 1) == 12);
        EXPECT(x(2, 2, 2) == 2 + 2 * 3 + 2 * 12);
        EXPECT(x(2, 3, 4) == x.size() - 1 && x.size() == 60);
        EXPECT(x(arr<3>{2, 3, 4}) == 59);
    }
    {
        auto x = tensor_indexer<S, 3u, layout::col_major>({3, 4, 5}, {1, 5, 30});
        EXPECT(x(1, 0, 0) == 1 && x(0, 1, 0) == 5 && x(0, 0, 1) == 30);
        EXPECT(x(2, 2, 2) == 2 + 2 * 5 + 2 * 30 && x(2, 3, 4) == 137 && x.size() == 150);
    }
    {
        auto x1 = tensor_indexer<S, 5u, layout::col_major>({2, 3, 4, 5, 6});
        auto x2 = tensor_indexer<S, 5u, layout::col_major>({2, 3, 4, 5, 6}, {1, 3, 9, 36, 216});
        EXPECT(x1.may_fuse() && (x1.may_fuse<1, 3>()) && x1.may_fuse<4>());
        EXPECT(!x2.may_fuse() && (x2.may_fuse<1, 3>()) && (x2.may_fuse<2, 3>()) && !(x2.may_fuse<1, 4>()));
        auto f1 = x1.fused();
        auto f2 = x1.fused<1, 3>();
        EXPECT(f1.dim() == 1 && f1.shape(0) == 720 && f1.stride(0) == 1);
        EXPECT(f2.dim() == 3 && (f2.shape() == arr<3>{2, 60, 6}) && (f2.stride() == arr<3>{1, 2, 120}));
        EXPECT(f2(1, 35, 3) == x1(1, 2, 3, 2, 3));
    }
    {
        auto x1 = tensor_indexer<S, 4u, layout::col_major>({10, 8, 5, 10}, {1, 12, 120, 500});
        EXPECT(x1.may_reshape_mode(1, arr<3>{2, 2, 2}) && !x1.may_reshape_mode(1, arr<3>{2, 3, 2}));
        auto r1 = x1.reshaped_mode(0, arr<2>{2, 5});
        EXPECT(r1.dim() == 5u && (r1.shape() == arr<5>{2, 5, 8, 5, 10}) && (r1.stride() == arr<5>{1, 2, 12, 120, 500}));
        auto r2 = x1.reshaped_mode(1, arr<3>{2, 2, 2});
        EXPECT(r2.dim() == 6u && (r2.shape() == arr<6>{10, 2, 2, 2, 5, 10}) &&
               (r2.stride() == arr<6>{1, 12, 24, 48, 120, 500}));
        auto r3 = x1.reshaped_mode(3, arr<2>{5, 2});
        EXPECT(r3.dim() == 5u && (r3.shape() == arr<5>{10, 8, 5, 5, 2}) && (r3.stride() == arr<5>{1, 12, 120, 500, 2500}));
        auto r4 = x1.reshaped_mode(2, arr<1>{5});
        EXPECT(r4.dim() == 4u && r4.shape() == x1.shape() && r4.stride() == x1.stride());
    }
}

static void row_major() {
    {
        auto x = tensor_indexer<S, 3u, layout::row_major>({3, 4, 5});
        EXPECT(x(0, 0, 0) == 0 && x(1, 0, 0) == 20 && x(0, 1, 0) == 5 && x(0, 0, 1) == 1);
        EXPECT(x(2, 2, 2) == 2 * 20 + 2 * 5 + 2 && x(2, 3, 4) == x.size() - 1 && x.size() == 60);
    }
    {
        auto x = tensor_indexer<S, 3u>({3, 4, 5}, {30, 6, 1}); // row_major is the default layout
        EXPECT(x(1, 0, 0) == 30 && x(0, 1, 0) == 6 && x(0, 0, 1) == 1);
        EXPECT(x(2, 2, 2) == 2 * 30 + 2 * 6 + 2 && x(2, 3, 4) == 82 && x.size() == 90);
    }
    {
        auto x1 = tensor_indexer<S, 5u, layout::row_major>({6, 5, 4, 3, 2});
        auto x2 = tensor_indexer<S, 5u, layout::row_major>({6, 5, 4, 3, 2}, {216, 36, 9, 3, 1});
        EXPECT(x1.may_fuse() && (x1.may_fuse<1, 3>()) && x1.may_fuse<4>());
        EXPECT(!x2.may_fuse() && (x2.may_fuse<1, 3>()) && (x2.may_fuse<2, 3>()) && !(x2.may_fuse<1, 4>()));
        auto f1 = x1.fused();
        auto f2 = x1.fused<1, 3>();
        EXPECT(f1.dim() == 1 && f1.shape(0) == 720 && f1.stride(0) == 1);
        EXPECT(f2.dim() == 3 && (f2.shape() == arr<3>{6, 60, 2}) && (f2.stride() == arr<3>{120, 2, 1}));
        EXPECT(f2(1, 35, 3) == x1(1, 2, 3, 2, 3));
        // mode numbers count from the fastest mode: <0,1> merges the LAST two indices
        auto f3 = x1.fused<0, 1>();
        EXPECT((f3.shape() == arr<4>{6, 5, 4, 6}) && (f3.stride() == arr<4>{120, 24, 6, 1}));
    }
    {
        auto x1 = tensor_indexer<S, 4u, layout::row_major>({10, 8, 5, 10});
        auto r1 = x1.reshaped_mode(0, arr<2>{2, 5});
        EXPECT(r1.dim() == 5u && (r1.shape() == arr<5>{2, 5, 8, 5, 10}) && (r1.stride() == arr<5>{2000, 400, 50, 10, 1}));
        auto r2 = x1.reshaped_mode(1, arr<3>{2, 2, 2});
        EXPECT(r2.dim() == 6u && (r2.shape() == arr<6>{10, 2, 2, 2, 5, 10}) &&
               (r2.stride() == arr<6>{400, 200, 100, 50, 10, 1}));
        auto r3 = x1.reshaped_mode(3, arr<2>{5, 2});
        EXPECT(r3.dim() == 5u && (r3.shape() == arr<5>{10, 8, 5, 5, 2}) && (r3.stride() == arr<5>{400, 50, 10, 2, 1}));
    }
    {
        auto a = fit_array<5>(arr<3>{7, 8, 9}, S(1));
        EXPECT((a == arr<5>{7, 8, 9, 1, 1}));
        auto b = fit_array<2>(arr<3>{7, 8, 9});
        EXPECT((b == arr<2>{7, 8}));
    }
}

int main() {
    col_major();
    row_major();
    std::printf("%s (%d failures)\n", failures ? "FAILED" : "OK", failures);
    return failures ? 1 : 0;
}
