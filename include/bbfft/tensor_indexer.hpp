// bbfft/tensor_indexer.hpp -- address arithmetic for multi-dimensional data (host side helper of
// the public API; drop-in for the reference's include/bbfft/tensor_indexer.hpp:60-337, behaviour
// pinned by the reference's test/tensor.cpp:25-155, re-hosted in tests/cpp/test_tensor_indexer.cpp).
//
// Written from the documented behaviour, not from the reference's code: shape and strides are kept
// in the caller's own index order and every "memory order" question (which mode is d-th fastest) is
// answered by slot(d); the reference stores reversed copies for row-major tensors instead.
#ifndef BBFFT_TENSOR_INDEXER_HPP
#define BBFFT_TENSOR_INDEXER_HPP

#include <array>
#include <cstddef>
#include <type_traits>

namespace bbfft {

/// Storage order: row_major = the LAST index is fastest in memory, col_major = the FIRST one is.
enum class layout { row_major, col_major };

/// Copy `in` into an array of length Dout: truncated, or extended with `fill_value`.
template <std::size_t Dout, typename IdxT, std::size_t Din>
constexpr auto fit_array(std::array<IdxT, Din> const &in, IdxT fill_value = IdxT(0)) {
    std::array<IdxT, Dout> out{};
    for (std::size_t d = 0; d < Dout; ++d) out[d] = d < Din ? in[d] : fill_value;
    return out;
}

template <typename IdxT, unsigned int D, layout L = layout::row_major> class tensor_indexer {
    static_assert(D >= 1u, "a tensor has at least one mode");

  public:
    using multi_idx_t = std::array<IdxT, D>;

    tensor_indexer() : extent_{}, step_{} {}

    /// Packed tensor N_1 x ... x N_D: the fastest mode has stride 1, every further mode the product of
    /// the faster extents.
    tensor_indexer(multi_idx_t shape) : extent_(shape), step_{} {
        IdxT run = IdxT(1);
        for (unsigned int f = 0; f < D; ++f) {
            step_[slot(f)] = run;
            run = run * extent_[slot(f)];
        }
    }

    /// Explicit strides, given in the same order as the shape.
    tensor_indexer(multi_idx_t shape, multi_idx_t stride) : extent_(shape), step_(stride) {}

    /// Linear index of entry (i_1, ..., i_D).
    template <typename... Indices, typename = std::enable_if_t<sizeof...(Indices) == D, int>>
    IdxT operator()(Indices &&...is) const {
        const multi_idx_t idx = {static_cast<IdxT>(is)...};
        return (*this)(idx);
    }
    IdxT operator()(multi_idx_t const &idx) const {
        IdxT a = IdxT(0);
        for (unsigned int d = 0; d < D; ++d) a = a + idx[d] * step_[d];
        return a;
    }

    multi_idx_t shape() const { return extent_; }
    IdxT shape(unsigned int d) const { return extent_[d]; }
    multi_idx_t stride() const { return step_; }
    IdxT stride(unsigned int d) const { return step_[d]; }
    /// Number of elements spanned (the slowest mode's stride times its extent).
    IdxT size() const { return step_[slot(D - 1)] * extent_[slot(D - 1)]; }
    constexpr auto dim() const { return D; }

    /// Modes Dfrom..Dto -- counted from the FASTEST mode, as in the reference -- can be treated as one
    /// super-index when each is packed onto the next.
    template <unsigned int Dfrom = 0, unsigned int Dto = D - 1> bool may_fuse() const {
        static_assert(Dfrom <= Dto && Dto <= D - 1, "mode range out of bounds");
        for (unsigned int f = Dfrom; f < Dto; ++f) {
            if (step_[slot(f)] * extent_[slot(f)] != step_[slot(f + 1)]) return false;
        }
        return true;
    }

    /// Indexer with the modes Dfrom..Dto (counted from the fastest mode) merged into one.
    template <unsigned int Dfrom = 0, unsigned int Dto = D - 1> auto fused() const {
        static_assert(Dfrom <= Dto && Dto <= D - 1, "mode range out of bounds");
        constexpr unsigned int E = D - (Dto - Dfrom);
        using out_t = tensor_indexer<IdxT, E, L>;
        std::array<IdxT, E> shp{}, str{};
        auto oslot = [](unsigned int f) { return L == layout::col_major ? f : E - 1 - f; };
        for (unsigned int f = 0; f < E; ++f) {
            // f-th fastest mode of the result: untouched below Dfrom, the merged mode at Dfrom,
            // shifted above
            const unsigned int src = f <= Dfrom ? f : f + (Dto - Dfrom);
            IdxT n = extent_[slot(src)];
            if (f == Dfrom) {
                for (unsigned int g = Dfrom + 1; g <= Dto; ++g) n = n * extent_[slot(g)];
            }
            shp[oslot(f)] = n;
            str[oslot(f)] = step_[slot(src)];
        }
        return out_t(shp, str);
    }

    /// A mode of extent N can be viewed as an E-dimensional mode when the extents multiply to N.
    template <std::size_t E> bool may_reshape_mode(int mode, std::array<IdxT, E> const &mode_shape) const {
        IdxT n = IdxT(1);
        for (std::size_t i = 0; i < E; ++i) n = n * mode_shape[i];
        return n == extent_[static_cast<unsigned int>(mode)];
    }

    /// Indexer in which mode `mode` (caller's numbering) is viewed as `mode_shape` (caller's order).
    template <std::size_t E> auto reshaped_mode(int mode, std::array<IdxT, E> mode_shape) const {
        static_assert(E > 0u, "a mode is reshaped into at least one mode");
        constexpr unsigned int F = D + static_cast<unsigned int>(E) - 1u;
        using out_t = tensor_indexer<IdxT, F, L>;
        std::array<IdxT, F> shp{}, str{};
        const unsigned int md = static_cast<unsigned int>(mode);
        for (unsigned int d = 0; d < md; ++d) {
            shp[d] = extent_[d];
            str[d] = step_[d];
        }
        for (unsigned int d = md + 1; d < D; ++d) {
            shp[d + E - 1] = extent_[d];
            str[d + E - 1] = step_[d];
        }
        // the sub-modes are packed inside the old mode, fastest sub-mode on the old stride
        IdxT run = step_[md];
        for (std::size_t i = 0; i < E; ++i) {
            const std::size_t sub = L == layout::col_major ? i : E - 1 - i; // i-th fastest sub-mode
            shp[md + sub] = mode_shape[sub];
            str[md + sub] = run;
            run = run * mode_shape[sub];
        }
        return out_t(shp, str);
    }

  private:
    // position (in the caller's index order) of the f-th fastest mode
    static constexpr unsigned int slot(unsigned int f) { return L == layout::col_major ? f : D - 1 - f; }

    multi_idx_t extent_;
    multi_idx_t step_;
};

} // namespace bbfft

#endif // BBFFT_TENSOR_INDEXER_HPP
