// core.cpp -- device-independent pieces of the bbfft host API (see include/bbfft/api.hpp):
// default strides, the FFT descriptor printer and parser, device_info and the kernel caches.
// Behaviour follows the reference (src/base/configuration.cpp:24-160, src/base/parser.cpp:59-270,
// src/base/device_info.cpp:13-95, src/base/{jit_cache,jit_cache_all,aot_cache}.cpp); the code is
// written from those semantics, not copied.
#include "bbfft/api.hpp"

#include <algorithm>
#include <cctype>
#include <ostream>
#include <sstream>
#include <stdexcept>

namespace bbfft {

char const *to_string(transform_type type) {
    static char const *names[] = {"c2c", "r2c", "c2r"};
    auto i = static_cast<int>(type);
    return (i >= 0 && i < 3) ? names[i] : "unknown";
}

// ---------------------------------------------------------------------------------------
// strides
// ---------------------------------------------------------------------------------------
namespace {
// extent of the first FFT mode as stored in memory
std::size_t stored_n1(std::size_t n1, bool real_side, bool spectrum_side, bool inplace) {
    if (spectrum_side) return n1 / 2 + 1;
    if (real_side) return inplace ? 2 * (n1 / 2 + 1) : n1;
    return n1;
}
tensor_extent packed(unsigned dim, tensor_extent const &shape, std::size_t n1_stored) {
    tensor_extent s = {};
    s[0] = 1;
    s[1] = shape[0];
    s[2] = n1_stored * s[1];
    for (unsigned d = 3; d < dim + 2 && d < max_tensor_dim; ++d) {
        s[d] = s[d - 1] * shape[d - 1];
    }
    return s;
}
} // namespace

auto default_istride(unsigned dim, tensor_extent const &shape, transform_type type, bool inplace)
    -> tensor_extent {
    bool real_in = type == transform_type::r2c;
    bool spec_in = type == transform_type::c2r;
    return packed(dim, shape, stored_n1(shape[1], real_in, spec_in, inplace));
}

auto default_ostride(unsigned dim, tensor_extent const &shape, transform_type type, bool inplace)
    -> tensor_extent {
    bool real_out = type == transform_type::c2r;
    bool spec_out = type == transform_type::r2c;
    return packed(dim, shape, stored_n1(shape[1], real_out, spec_out, inplace));
}

void configuration::set_strides_default(bool inplace) {
    istride = default_istride(dim, shape, type, inplace);
    ostride = default_ostride(dim, shape, type, inplace);
}

std::string configuration::to_string() const {
    std::ostringstream os;
    os << *this;
    return os.str();
}

// ---------------------------------------------------------------------------------------
// descriptor printer:  precision domain direction placement [M.]N1[xN2[xN3]][*K] [i..] [o..]
// ---------------------------------------------------------------------------------------
std::ostream &operator<<(std::ostream &os, configuration const &cfg) {
    if (cfg.fp != precision::f32 && cfg.fp != precision::f64) {
        throw std::runtime_error("Unsupported precision");
    }
    if (cfg.type != transform_type::c2c && cfg.type != transform_type::r2c &&
        cfg.type != transform_type::c2r) {
        throw std::runtime_error("Unsupported transform type");
    }
    if (cfg.dir != direction::forward && cfg.dir != direction::backward) {
        throw std::runtime_error("Unsupported direction");
    }
    if ((cfg.type == transform_type::r2c && cfg.dir != direction::forward) ||
        (cfg.type == transform_type::c2r && cfg.dir != direction::backward)) {
        throw std::runtime_error("r2c direction must be forward and c2r direction must be backward");
    }
    os << (cfg.fp == precision::f32 ? 's' : 'd') << (cfg.type == transform_type::c2c ? 'c' : 'r')
       << (cfg.dir == direction::forward ? 'f' : 'b');

    // in-place iff input and output address the same bytes: compare strides in units of reals
    auto in_units = cfg.istride, out_units = cfg.ostride;
    for (unsigned d = 2; d < max_tensor_dim; ++d) {
        if (cfg.type == transform_type::r2c) out_units[d] *= 2;
        if (cfg.type == transform_type::c2r) in_units[d] *= 2;
    }
    bool inplace = std::equal(in_units.begin(), in_units.begin() + cfg.dim + 2, out_units.begin());
    os << (inplace ? 'i' : 'o');

    if (cfg.shape[0] != 1) os << cfg.shape[0] << '.';
    for (unsigned d = 0; d < cfg.dim; ++d) os << (d ? "x" : "") << cfg.shape[1 + d];
    if (cfg.shape[cfg.dim + 1] != 1) os << '*' << cfg.shape[cfg.dim + 1];

    auto print_strides = [&](char tag, tensor_extent const &s) {
        os << tag;
        for (unsigned d = 0; d < cfg.dim + 2; ++d) os << (d ? "," : "") << s[d];
    };
    if (cfg.istride != default_istride(cfg.dim, cfg.shape, cfg.type, inplace)) print_strides('i', cfg.istride);
    if (cfg.ostride != default_ostride(cfg.dim, cfg.shape, cfg.type, inplace)) print_strides('o', cfg.ostride);
    return os;
}

// ---------------------------------------------------------------------------------------
// descriptor parser
// ---------------------------------------------------------------------------------------
namespace {
class cursor {
  public:
    explicit cursor(std::string_view text, bool skip_blanks) : text_(text), skip_(skip_blanks) {}
    bool at_end() {
        blanks();
        return pos_ >= text_.size();
    }
    char peek() {
        blanks();
        return pos_ < text_.size() ? text_[pos_] : '\0';
    }
    char take() {
        blanks();
        return pos_ < text_.size() ? text_[pos_++] : '\0';
    }
    [[noreturn]] void fail(std::string const &msg, bool caret = true) const {
        std::ostringstream os;
        os << "==> " << text_ << " is malformed: " << msg;
        if (caret) os << "\n" << std::string(4 + pos_, ' ') << "^";
        throw std::runtime_error(os.str());
    }
    void expect(char c) {
        if (peek() != c) fail(std::string("expected '") + c + "'");
        ++pos_;
    }
    std::size_t number() {
        blanks();
        std::size_t begin = pos_, v = 0;
        while (pos_ < text_.size() && std::isdigit(static_cast<unsigned char>(text_[pos_]))) {
            v = v * 10 + std::size_t(text_[pos_++] - '0');
        }
        if (begin == pos_) fail("expected number");
        return v;
    }

  private:
    void blanks() {
        while (skip_ && pos_ < text_.size() && (text_[pos_] == ' ' || text_[pos_] == '\t')) ++pos_;
    }
    std::string_view text_;
    std::size_t pos_ = 0;
    bool skip_;
};
} // namespace

configuration parse_fft_descriptor(std::string_view desc) {
    configuration cfg = {};
    cursor c(desc, false);
    switch (c.peek()) {
    case 's': cfg.fp = precision::f32; break;
    case 'd': cfg.fp = precision::f64; break;
    default: c.fail("expected 's' (single) or 'd' (double)");
    }
    c.take();
    bool real = false;
    switch (c.peek()) {
    case 'c': real = false; break;
    case 'r': real = true; break;
    default: c.fail("expected 'c' (complex) or 'r' (real)");
    }
    c.take();
    switch (c.peek()) {
    case 'f': cfg.dir = direction::forward; break;
    case 'b': cfg.dir = direction::backward; break;
    default: c.fail("expected 'f' (forward) or 'b' (backward)");
    }
    c.take();
    cfg.type = !real ? transform_type::c2c
                     : (cfg.dir == direction::forward ? transform_type::r2c : transform_type::c2r);
    bool inplace = false;
    switch (c.peek()) {
    case 'i': inplace = true; break;
    case 'o': inplace = false; break;
    default: c.fail("expected 'i' (in-place) or 'o' (out-of-place)");
    }
    c.take();

    // shape: numbers separated by '.', 'x', '*'
    std::vector<std::size_t> nums;
    std::vector<char> seps;
    nums.push_back(c.number());
    while (c.peek() == '.' || c.peek() == 'x' || c.peek() == '*') {
        if (nums.size() >= max_tensor_dim) {
            c.fail("tensor dimension must not be larger than " + std::to_string(max_tensor_dim), false);
        }
        seps.push_back(c.take());
        nums.push_back(c.number());
    }
    bool has_m = !seps.empty() && seps.front() == '.';
    bool has_k = !seps.empty() && seps.back() == '*';
    unsigned nx = 0;
    for (std::size_t i = has_m ? 1 : 0; i + (has_k ? 1 : 0) < seps.size(); ++i) {
        if (seps[i] != 'x') {
            c.fail("'.' or '*' must only appear at the beginning or end of the tensor shape, respectively",
                   false);
        }
        ++nx;
    }
    cfg.dim = 1 + nx;
    if (cfg.dim > max_fft_dim) {
        c.fail("only " + std::to_string(max_fft_dim - 1) + " 'x' are supported", false);
    }
    cfg.shape = {};
    unsigned at = 0;
    cfg.shape[at++] = has_m ? nums.front() : 1;
    for (std::size_t i = has_m ? 1 : 0; i < nums.size() - (has_k ? 1 : 0); ++i) cfg.shape[at++] = nums[i];
    cfg.shape[at++] = has_k ? nums.back() : 1;

    bool have_i = false, have_o = false;
    auto strides = [&](tensor_extent &s) {
        for (unsigned d = 0; d < cfg.dim + 2; ++d) {
            if (d) c.expect(',');
            s[d] = c.number();
        }
    };
    while (!c.at_end()) {
        char tag = c.peek();
        if (tag == 'i') {
            c.take();
            strides(cfg.istride);
            have_i = true;
        } else if (tag == 'o') {
            c.take();
            strides(cfg.ostride);
            have_o = true;
        } else {
            c.fail("expected 'i' (istride) or 'o' (ostride)");
        }
    }
    if (!have_i) cfg.istride = default_istride(cfg.dim, cfg.shape, cfg.type, inplace);
    if (!have_o) cfg.ostride = default_ostride(cfg.dim, cfg.shape, cfg.type, inplace);
    return cfg;
}

device_info parse_device_info(std::string_view desc) {
    // "{max_work_group_size, {sgs, ...}, local_memory_size, gpu|cpu}"
    device_info info = {};
    cursor c(desc, true);
    c.expect('{');
    info.max_work_group_size = c.number();
    c.expect(',');
    c.expect('{');
    info.subgroup_sizes.push_back(c.number());
    while (c.peek() == ',') {
        c.take();
        info.subgroup_sizes.push_back(c.number());
    }
    c.expect('}');
    c.expect(',');
    info.local_memory_size = c.number();
    c.expect(',');
    char first = c.peek();
    if (first != 'g' && first != 'c') c.fail("expected gpu or cpu");
    c.take();
    c.expect('p');
    c.expect('u');
    info.type = first == 'g' ? device_type::gpu : device_type::cpu;
    c.expect('}');
    return info;
}

// ---------------------------------------------------------------------------------------
// device_info
// ---------------------------------------------------------------------------------------
std::size_t device_info::min_subgroup_size() const {
    return subgroup_sizes.empty() ? 8 : *std::min_element(subgroup_sizes.begin(), subgroup_sizes.end());
}
std::size_t device_info::max_subgroup_size() const {
    return subgroup_sizes.empty() ? 8 : *std::max_element(subgroup_sizes.begin(), subgroup_sizes.end());
}
std::size_t device_info::register_space_min() const {
    // The reference's register model (bytes of register file one sub-group can use):
    // cpu: 32 vector registers of 64 bytes; gpu: 128 registers of 32 bytes, scaled by sgs/8.
    switch (type) {
    case device_type::cpu: return 32 * 64;
    case device_type::gpu: return std::max<std::size_t>(1, min_subgroup_size() / 8) * 32 * 128;
    default: throw std::runtime_error("register_space unknown for custom device");
    }
}
std::size_t device_info::register_space_max() const { return register_space_min(); }
std::string device_info::to_string() const {
    std::ostringstream os;
    os << *this;
    return os.str();
}
bool device_info::operator==(device_info const &o) const {
    return max_work_group_size == o.max_work_group_size && subgroup_sizes == o.subgroup_sizes &&
           local_memory_size == o.local_memory_size && type == o.type;
}
bool device_info::operator!=(device_info const &o) const { return !(*this == o); }

std::ostream &operator<<(std::ostream &os, device_type type) {
    return os << (type == device_type::gpu ? "gpu" : (type == device_type::cpu ? "cpu" : "custom"));
}
std::ostream &operator<<(std::ostream &os, device_info const &info) {
    os << "{" << info.max_work_group_size << ", {";
    for (std::size_t i = 0; i < info.subgroup_sizes.size(); ++i) {
        os << (i ? ", " : "") << info.subgroup_sizes[i];
    }
    return os << "}, " << info.local_memory_size << ", " << info.type << "}";
}

// ---------------------------------------------------------------------------------------
// caches
// ---------------------------------------------------------------------------------------
jit_cache::~jit_cache() {}

auto jit_cache_all::get(jit_cache_key const &key) const -> shared_handle<module_handle_t> {
    auto it = mods_.find(key);
    return it == mods_.end() ? shared_handle<module_handle_t>{} : it->second;
}
void jit_cache_all::store(jit_cache_key const &key, shared_handle<module_handle_t> mod) {
    mods_[key] = std::move(mod);
}
auto jit_cache_all::kernel_names() const -> std::vector<std::string> {
    std::vector<std::string> names;
    names.reserve(mods_.size());
    for (auto const &kv : mods_) names.push_back(kv.first.kernel_name);
    return names;
}

auto aot_cache::get(jit_cache_key const &key) const -> shared_handle<module_handle_t> {
    for (auto const &m : aot_modules_) {
        if (m.device_id == key.device_id && m.kernel_names.count(key.kernel_name)) return m.mod;
    }
    return {};
}
void aot_cache::store(jit_cache_key const &, shared_handle<module_handle_t>) {
    // ahead-of-time cache: nothing is added at run time (reference: src/base/aot_cache.cpp:24)
}
void aot_cache::register_module(aot_module aot_mod) { aot_modules_.push_back(std::move(aot_mod)); }

} // namespace bbfft
