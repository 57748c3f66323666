// planner.cpp -- see planner.hpp.
#include "planner.hpp"

#include "bbfft/api.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>

namespace bbfft::cuda {

static void unit_root(long num, long den, int dir, double &re, double &im);
static bool packed_f32();
static void emit_w_table(std::ostringstream &os, int R);

// ------------------------------------------------------------------------------------------
// integer helpers
// ------------------------------------------------------------------------------------------
std::vector<int> prime_factors(int n) {
    // ascending trial division (reference: src/base/prime_factorization.cpp:13-25)
    std::vector<int> f;
    for (int p = 2; p * p <= n; ++p) {
        while (n % p == 0) {
            f.push_back(p);
            n /= p;
        }
    }
    if (n > 1) {
        f.push_back(n);
    }
    return f;
}

static int max_prime(int n) {
    auto f = prime_factors(n);
    return f.empty() ? 1 : f.back();
}

// A radix is computable by bbk::reg_fft when its prime factors have hand-written or generic
// butterflies; generic odd primes are O(p^2) so keep them small inside composites.
bool radix_supported(int r) {
    if (r < 2) return false;
    return true;
}

bool direct_radix(int r) { return r > 31 && max_prime(r) == r; }

static int pow2_ceil(std::uint64_t x) {
    int p = 1;
    while (std::uint64_t(p) < x && p < (1 << 30)) p <<= 1;
    return p;
}

// Largest per-thread register count with which `blocks` CTAs of `threads` threads are resident on
// one SM: the 64K-register file is four 16K partitions, a CTA's warps are dealt round-robin over
// them, and registers are allocated per warp in multiples of 8 per thread.
int reg_cap(int threads, int blocks) {
    int warps = (threads + 31) / 32;
    int per_partition = std::max(1, blocks) * ((warps + 3) / 4);
    int cap = (16384 / per_partition / 32) / 8 * 8;
    return std::max(24, std::min(cap, 255));
}

std::uint64_t kernel_params::k_per_cta() const {
    std::uint64_t per = klanes ? std::uint64_t(ML) * BH : std::uint64_t(BH);
    if (mode == k_r2c_double || mode == k_c2r_double) per *= 2;
    return per;
}

std::uint64_t kernel_params::grid(std::uint64_t K) const {
    std::uint64_t per = k_per_cta();
    std::uint64_t kblocks = (K + per - 1) / per;
    std::uint64_t mblocks = klanes ? 1 : (M + ML - 1) / ML;
    return kblocks * mblocks;
}

// ------------------------------------------------------------------------------------------
// tuning string
// ------------------------------------------------------------------------------------------
static std::map<std::string, std::string> parse_tune(std::string const &tune) {
    std::map<std::string, std::string> kv;
    std::stringstream ss(tune);
    std::string item;
    while (std::getline(ss, item, ',')) {
        auto eq = item.find('=');
        if (eq == std::string::npos) continue;
        kv[item.substr(0, eq)] = item.substr(eq + 1);
    }
    return kv;
}

static std::vector<int> parse_radices(std::string const &s) {
    std::vector<int> r;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, 'x')) {
        r.push_back(std::atoi(item.c_str()));
    }
    return r;
}

// ------------------------------------------------------------------------------------------
// factorization search
// ------------------------------------------------------------------------------------------
namespace {

struct stage_choice {
    std::vector<int> radix;
    int T = 1;
    int regs = 0;    // complex elements held per thread (max over stages)
    double cost = 0; // lower is better
};

// relative arithmetic cost of a length-r in-register FFT (per element), rough
double radix_cost(int r) {
    double c = 0;
    for (int p : prime_factors(r)) {
        c += (p == 2) ? 1.0 : (p == 3 ? 1.8 : (p == 5 ? 2.6 : (p == 7 ? 3.4 : 0.5 * p)));
    }
    return c;
}

void enum_factorizations(int n, int max_r, int max_l, std::vector<int> &cur,
                         std::vector<std::vector<int>> &out) {
    if (n == 1) {
        if (!cur.empty()) out.push_back(cur);
        return;
    }
    if (int(cur.size()) >= max_l) return;
    int lo = cur.empty() ? 2 : cur.back(); // ascending
    for (int r = lo; r <= n && r <= max_r; ++r) {
        if (n % r) continue;
        if (max_prime(r) > 31 && max_prime(r) != r) continue; // a large prime is a stage of its own
        cur.push_back(r);
        enum_factorizations(n / r, max_r, max_l, cur, out);
        cur.pop_back();
    }
}

// mirror_max > 0 (fused real transforms): the factorization must contain a radix <= mirror_max for
// the stage that holds a sub-FFT together with its mirror.
stage_choice choose_stages(int N, int fp, int max_threads_per_transform, int mirror_max = 0) {
    const int single_max = (fp == 4) ? 32 : 16; // one thread holds the whole transform
    stage_choice best;
    best.cost = 1e30;
    const bool has_direct = direct_radix(max_prime(N));
    if (N <= single_max) {
        best.radix = {N};
        best.T = 1;
        best.regs = N;
        best.cost = 0;
        return best;
    }
    double ideal = 0;
    for (int p : prime_factors(N)) ideal += radix_cost(p);
    // relax the register budget / thread limit until something fits
    // Measured on B200 (profiles/r01_tuner_probe.json): two stages with large
    // radices (18x18, 20x20, 14x25) beat three stages, so allow up to 32 (fp32) / 20 (fp64)
    // complex elements per thread before adding a stage.
    const int ebase = fp == 4 ? 32 : 20;
    for (int relax = 0; relax < 4 && best.radix.empty(); ++relax) {
        // complex elements per thread (a large prime is not held in registers: direct stage)
        const int emax = std::max(ebase << relax, has_direct ? 0 : max_prime(N));
        const int rmax_enum = std::max(emax, max_prime(N));
        const int tmax = max_threads_per_transform << relax;
        std::vector<std::vector<int>> facs;
        std::vector<int> cur;
        enum_factorizations(N, rmax_enum, 4, cur, facs);
        for (auto const &f : facs) {
            int L = int(f.size());
            if (L < 2 && N > 64 && !direct_radix(N)) continue;
            if (mirror_max > 0 && L > 1 && f.front() > mirror_max) continue; // f is ascending
            std::set<int> tcand;
            for (int r : f) {
                for (int c = 1; c <= 8; ++c) {
                    int nsub = N / r;
                    tcand.insert((nsub + c - 1) / c);
                }
            }
            if (L == 1) tcand = {1};
            if (has_direct) {
                // a direct stage deals OUTPUTS to the threads: any thread count works
                for (int c = 4; c <= 16; c *= 2) tcand.insert((N + c - 1) / c);
            }
            for (int T : tcand) {
                if (T < 1 || T > tmax) continue;
                int regs = 0, rsum = 0;
                double work = 0;
                for (int r : f) {
                    int nsub = N / r;
                    int cnt = (nsub + T - 1) / T;
                    if (direct_radix(r)) {
                        regs = std::max(regs, (N + T - 1) / T);
                        work += 0.5 * r; // O(r) multiply-adds per element
                    } else {
                        regs = std::max(regs, cnt * r);
                        work += double(cnt) * T * r / N * radix_cost(r);
                    }
                    rsum += r;
                }
                if (regs > emax) continue;
                // cost: arithmetic (counting idle lanes) + exchange passes + mild preferences
                // for ~8 elements per thread and for balanced radices
                double cost = work / ideal + 0.5 * (L - 1);
                cost += 0.02 * std::abs(std::log2(double(regs) / 8.0));
                cost += 0.001 * rsum;
                if (cost < best.cost - 1e-9) {
                    best.cost = cost;
                    best.radix = f;
                    best.T = T;
                    best.regs = regs;
                }
            }
        }
    }
    if (best.radix.empty() && mirror_max > 0) {
        return choose_stages(N, fp, max_threads_per_transform, 0);
    }
    if (best.radix.empty()) {
        throw bad_configuration("bbfft-cuda planner: no factorization found for N=" +
                                 std::to_string(N));
    }
    return best;
}

// ---- shared memory bank-conflict model -----------------------------------------------------
struct layout_eval {
    kernel_params const &p;
    int elem_bytes;
    int pad(int pos, int padk) const { return padk > 0 ? pos + pos / padk : pos; }
    int soff(int b, int pos, int padk, int row) const {
        if (p.LL > 1) return (b % p.LL) + p.LL * pad(pos, padk) + row * (b / p.LL);
        return pad(pos, padk) + row * b;
    }
    // wavefronts needed by one warp-wide access (addresses in elements; -1 = inactive lane)
    int wavefronts(std::vector<int> const &off) const {
        int lanes_per_group = std::max(1, 32 * 4 / elem_bytes); // 16 for 8 B, 8 for 16 B
        int total = 0;
        for (int g = 0; g < 32; g += lanes_per_group) {
            std::map<int, std::set<int>> bank_words;
            bool any = false;
            for (int l = g; l < g + lanes_per_group && l < 32; ++l) {
                if (off[l] < 0) continue;
                any = true;
                long byte = long(off[l]) * elem_bytes;
                for (int w = 0; w < elem_bytes / 4; ++w) {
                    int word = int(byte / 4) + w;
                    bank_words[word % 32].insert(word);
                }
            }
            if (!any) continue;
            std::size_t mx = 1;
            for (auto &kv : bank_words) mx = std::max(mx, kv.second.size());
            total += int(mx);
        }
        return total;
    }
    int ideal(std::vector<int> const &off) const {
        int lanes_per_group = std::max(1, 32 * 4 / elem_bytes);
        int total = 0;
        for (int g = 0; g < 32; g += lanes_per_group) {
            for (int l = g; l < g + lanes_per_group && l < 32; ++l) {
                if (off[l] >= 0) {
                    ++total;
                    break;
                }
            }
        }
        return total;
    }
    static int pos_of_bin(kernel_params const &p, int k) {
        int pos = 0, rem = k;
        for (int s = 0; s < p.L; ++s) {
            pos = pos * p.radix[s] + rem % p.radix[s];
            rem /= p.radix[s];
        }
        return pos;
    }
    // excess wavefronts over all shared-memory phases of the kernel for warps 0 and 1
    long score(int padk, int row) const {
        long excess = 0;
        int threads = p.threads;
        for (int warp = 0; warp < std::min(2, (threads + 31) / 32); ++warp) {
            // stage accesses
            for (int s = 0; s < p.L; ++s) {
                bool reads = (s > 0) || p.load_staged;
                bool writes = (s < p.L - 1) || p.store_staged;
                if (!reads && !writes) continue;
                int R = p.radix[s];
                int NS = p.N;
                for (int i = 0; i < s; ++i) NS /= p.radix[i];
                int NS1 = NS / R, NSUB = p.N / R;
                int CNT = (NSUB + p.T - 1) / p.T;
                for (int i = 0; i < CNT; ++i) {
                    for (int j = 0; j < R; ++j) {
                        std::vector<int> off(32, -1);
                        for (int l = 0; l < 32; ++l) {
                            int tid = warp * 32 + l;
                            if (tid >= threads) continue;
                            int l0 = tid % p.ML, t = (tid / p.ML) % p.T, bh = tid / (p.ML * p.T);
                            int b = l0 + p.ML * bh;
                            int u = t + p.T * i;
                            if (u >= NSUB) continue;
                            int n2 = u % NS1, q = u / NS1;
                            off[l] = soff(b, q * NS + n2 + NS1 * j, padk, row);
                        }
                        int e = wavefronts(off) - ideal(off);
                        excess += e * ((reads ? 1 : 0) + (writes ? 1 : 0));
                    }
                }
            }
            // cooperative copies
            int mlc = p.klanes ? 1 : p.ML;
            int kb = p.klanes ? p.ML * p.BH : p.BH;
            int total = mlc * p.N * kb;
            for (int pass = 0; pass < 2; ++pass) {
                bool active = pass == 0 ? p.load_staged : p.store_staged;
                if (!active || p.mode != k_c2c) continue;
                for (int it = 0; it < 4; ++it) {
                    std::vector<int> off(32, -1);
                    for (int l = 0; l < 32; ++l) {
                        int idx = it * threads + warp * 32 + l;
                        if (idx >= total) continue;
                        int ml = idx % mlc, n = (idx / mlc) % p.N, kk = idx / (mlc * p.N);
                        int b = p.klanes ? kk : ml + p.ML * kk;
                        int pos = pass == 0 ? n : pos_of_bin(p, n);
                        off[l] = soff(b, pos, padk, row);
                    }
                    excess += wavefronts(off) - ideal(off);
                }
            }
        }
        return excess;
    }
};

void choose_smem_layout(kernel_params &p) {
    const int elem_bytes = 2 * p.fp;
    bool uses_smem = p.L > 1 || p.load_staged || p.store_staged || p.mode != k_c2c;
    for (int s = 0; s < p.L; ++s) uses_smem = uses_smem || direct_radix(p.radix[s]);
    p.LL = (p.klanes || p.ML == 1) ? 1 : p.ML;
    int rowlen = p.N + ((p.mode == k_r2c_half || p.mode == k_c2r_half) ? 1 : 0);
    if (!uses_smem) {
        p.PADK = 0;
        p.ROW = p.LL * rowlen;
        p.smem_bytes = 0;
        return;
    }
    layout_eval ev{p, elem_bytes};
    long best_score = -1;
    int best_padk = 0, best_row = p.LL * rowlen;
    std::vector<int> padks = {0};
    if (p.LL * elem_bytes < 128) {
        for (int k : {2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 15, 16, 20, 24, 25, 27, 28, 32, 64}) {
            if (k < rowlen) padks.push_back(k);
        }
    }
    for (int padk : padks) {
        int padded = ev.pad(rowlen - 1, padk) + 1;
        int max_extra = (p.LL * elem_bytes >= 128) ? 0 : 17;
        for (int extra = 0; extra <= max_extra; ++extra) {
            int row = p.LL * padded + extra;
            long sc = ev.score(padk, row);
            // prefer less memory on ties
            long key = sc * 4096 + (row - p.LL * rowlen);
            if (best_score < 0 || key < best_score) {
                best_score = key;
                best_padk = padk;
                best_row = row;
            }
        }
    }
    p.PADK = best_padk;
    p.ROW = best_row;
    int groups = (p.batch_per_cta() + p.LL - 1) / p.LL;
    p.smem_bytes = std::size_t(groups) * p.ROW * elem_bytes;
}

} // namespace

// ------------------------------------------------------------------------------------------
// wisdom: planner overrides measured on the B200 by tools/tune_gpu.py (c2c, M = 16 sweep)
// ------------------------------------------------------------------------------------------
namespace {
struct wisdom_entry {
    int fp, n;
    char const *tune;
};
const wisdom_entry wisdom_table[] = {
#include "wisdom.inc"
    {0, 0, ""}};

char const *wisdom_lookup(int fp, int n) {
    for (auto const &w : wisdom_table) {
        if (w.fp == fp && w.n == n) return w.tune;
    }
    return nullptr;
}

// Measured overrides for single configurations outside the c2c M=16 sweep (tools/exp_real_m1.py,
// profiles/r01c_exp_m1.txt): {transform type, bytes per real, user-visible N, M, tune}.
struct wisdom_special {
    int type, fp, n;
    std::uint64_t M;
    char const *tune;
};
const wisdom_special wisdom_specials[] = {
    {0, 4, 64, 1, "R=8x8,T=8,BH=16,ST=1,MB=2"},   // BASELINE config 1 shape: 6470 GB/s (heuristic 6316)
    // config 3, round 2 search over 817 overrides per placement (tools/cases_c3.json, profiles/r02g_tune.txt):
    // r2c 5979 / 5931 GB/s out of / in place (round 1: 5702), c2r 6083 / 5913 GB/s (round 1: 5793)
    {1, 4, 256, 1, "R=16x8,T=8,BH=8,MB=6,LD=0,ST=0"},
    {2, 4, 256, 1, "R=8x16,T=8,BH=8,MB=6,LD=0,ST=0"},
// measured r2c / c2r entries of the M = 16 real sweep (tools/tune_gpu.py --type r2c|c2r); they
// keep full batch lanes, so they apply to every M that is a multiple of the lane count
#include "wisdom_real.inc"
};
char const *wisdom_special_lookup(int type, int fp, std::uint64_t n, std::uint64_t M) {
    for (auto const &w : wisdom_specials) {
        // M == 0 marks a lane-generic entry: any M that fills whole 128-byte rows
        const std::uint64_t full = 128 / (2 * fp);
        const bool m_ok = w.M == M || (w.M == 0 && M >= full && M % full == 0);
        if (w.type == type && w.fp == fp && std::uint64_t(w.n) == n && m_ok) return w.tune;
    }
    return nullptr;
}
} // namespace

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
kernel_plan plan_kernel_1d(problem_1d const &prob, device_props const &dev,
                           std::string const &tune_str) {
    if (prob.N < 1 || prob.M < 1) {
        throw bad_configuration("bbfft-cuda planner: empty shape");
    }
    if (prob.N > 8192) {
        throw bad_configuration("bbfft-cuda planner: N too large for the single-kernel path");
    }
    auto tune = parse_tune(tune_str);
    // stage order is the caller's when it names the radices itself or a measured real entry does
    bool keep_stage_order = tune.count("R") != 0;
    {
        // Measured wisdom (c2c sweep) applies when the batch rows fill the wisdom's lanes.  Real
        // transforms run the same stages on their complex length (N/2 for even N), so they take
        // the entry of that length as their starting point.  Explicit overrides win;
        // BBFFT_CUDA_NO_WISDOM=1 turns it off (used by the tuner).
        char const *off = std::getenv("BBFFT_CUDA_NO_WISDOM");
        const std::uint64_t clen = (prob.type != 0 && prob.N % 2 == 0) ? prob.N / 2 : prob.N;
        const bool use = !(off && *off == '1');
        char const *sp = use ? wisdom_special_lookup(prob.type, prob.fp, prob.N, prob.M) : nullptr;
        if (sp) {
            keep_stage_order = true;
            for (auto const &kv : parse_tune(sp)) {
                if (!tune.count(kv.first)) tune[kv.first] = kv.second;
            }
        }
        // The c2c table is for c2c only.  Real transforms have their own measured entries
        // (wisdom_real.inc, every seven-smooth N >= 27); a real size without one runs the heuristic,
        // which is what the tuner measured as its default -- a c2c entry of the same complex length
        // was tuned for a different live-register set (measured: profiles/r01e_real_sweep.jsonl,
        // r2c f32 N=441 at 0.29 of peak with the inherited 21x21 entry).
        (void)clen;
        char const *w = (use && !sp && prob.type == 0) ? wisdom_lookup(prob.fp, int(prob.N)) : nullptr;
        // BBFFT_CUDA_WISDOM_OVERRIDE="8:490:R=10x7x7,T=49,ML=8,BH=1,MB=2;4:225:..." replaces c2c table entries for
        // one run: candidates are judged inside the real sweep (bench.py), not only in the tuner's regime
        std::string env_entry;
        if (char const *ov = std::getenv("BBFFT_CUDA_WISDOM_OVERRIDE")) {
            if (use && !sp && prob.type == 0) {
                const std::string key = std::to_string(prob.fp) + ":" + std::to_string(prob.N) + ":";
                std::stringstream all(ov);
                std::string item;
                while (std::getline(all, item, ';')) {
                    if (item.rfind(key, 0) == 0) {
                        env_entry = item.substr(key.size());
                        w = env_entry.c_str();
                    }
                }
            }
        }
        if (w && *w) {
            auto wt = parse_tune(w);
            int wml = wt.count("ML") ? std::atoi(wt["ML"].c_str()) : 0;
            if (wml > 0 && prob.M >= std::uint64_t(wml) && prob.M % wml == 0) {
                for (auto const &kv : wt) {
                    if (!tune.count(kv.first)) tune[kv.first] = kv.second;
                }
            }
        }
        // Real in-place transforms need all m of a k slice in one CTA (see inplace_unsupported).
        // When the strides allow the spectrum to overlay the real rows (real k-stride = twice the
        // complex one) and M fits one warp -- the cases the reference runs in place (Mb = pow2 >= M
        // up to the sub-group size, src/base/generator/small_batch_fft.cpp:19-38) -- the lanes
        // cover M: a measured entry keeps its radices and threads per transform, the CTA keeps its
        // size by shrinking the batch, and the heuristic chooses the register cap.
        const std::int64_t rs2 = prob.type == 1 ? prob.is2 : prob.os2;
        const std::int64_t cs2 = prob.type == 1 ? prob.os2 : prob.is2;
        const bool overlay = prob.type != 0 && rs2 == 2 * cs2;
        if (overlay && prob.M > 1 && prob.M <= 32 && !parse_tune(tune_str).count("ML")) {
            const int need = pow2_ceil(prob.M);
            const int have = tune.count("ML") ? std::atoi(tune["ML"].c_str()) : need;
            if (have != need) {
                if (tune.count("BH")) {
                    tune["BH"] = std::to_string(std::max(1, std::atoi(tune["BH"].c_str()) * have / need));
                }
                if (!parse_tune(tune_str).count("MB")) tune.erase("MB");
            }
            tune["ML"] = std::to_string(need);
        }
    }
    kernel_plan plan;
    kernel_params &p = plan.p;
    p.fp = prob.fp;
    p.dir = prob.dir;
    p.M = prob.M;
    p.is1 = prob.is1;
    p.is2 = prob.is2;
    p.os1 = prob.os1;
    p.os2 = prob.os2;
    p.nreal = int(prob.N);
    p.cb_load = prob.cb_load;
    p.cb_store = prob.cb_store;
    const bool real = prob.type != 0;
    if (!real) {
        p.mode = k_c2c;
        p.N = int(prob.N);
    } else if (prob.N % 2 == 0) {
        p.mode = prob.type == 1 ? k_r2c_half : k_c2r_half;
        p.N = int(prob.N / 2);
    } else {
        p.mode = prob.type == 1 ? k_r2c_double : k_c2r_double;
        p.N = int(prob.N);
    }
    const int elem_bytes = 2 * p.fp;

    // ---- batch lanes
    const int full_lanes = 128 / elem_bytes; // lanes that make one 128-byte row
    if (prob.M == 1) {
        p.ML = 1;
    } else {
        int ml = pow2_ceil(prob.M);
        p.ML = std::min(ml, full_lanes);
        if (prob.M > std::uint64_t(full_lanes) && prob.M % full_lanes != 0) {
            // pick the lane count that wastes the fewest lanes in the last m block
            int best = full_lanes;
            std::uint64_t best_waste = ~0ull;
            for (int c : {full_lanes, full_lanes / 2, 2 * full_lanes}) {
                if (c < 2 || c > 32) continue;
                std::uint64_t waste = (prob.M + c - 1) / c * c - prob.M;
                if (waste < best_waste || (waste == best_waste && c > best)) {
                    best_waste = waste;
                    best = c;
                }
            }
            p.ML = best;
        }
    }
    if (tune.count("ML")) p.ML = std::atoi(tune["ML"].c_str());

    // ---- stages
    int max_tpt = std::max(1, 512 / p.ML);
    stage_choice sc;
    if (tune.count("R")) {
        sc.radix = parse_radices(tune["R"]);
        int prod = 1, maxr = 1;
        for (int r : sc.radix) {
            prod *= r;
            maxr = std::max(maxr, r);
        }
        if (prod != p.N) throw std::runtime_error("bbfft-cuda planner: tune R does not multiply to N");
        sc.T = sc.radix.size() == 1 ? 1 : p.N / maxr;
    } else {
        const bool fused_real = p.mode != k_c2c && prob.M > 1;
        sc = choose_stages(p.N, p.fp, max_tpt, fused_real ? (p.fp == 4 ? 16 : 8) : 0);
    }
    if (tune.count("T")) sc.T = std::atoi(tune["T"].c_str());
    if (p.mode != k_c2c && !keep_stage_order) {
        // The fused post (r2c: last stage) / pre (c2r: first stage) pass keeps a sub-FFT and its
        // mirror in registers, twice the live values of any other stage: give it the smallest radix.
        if (p.mode == k_r2c_half || p.mode == k_r2c_double) {
            std::sort(sc.radix.begin(), sc.radix.end(), std::greater<int>());
        } else {
            std::sort(sc.radix.begin(), sc.radix.end());
        }
    }
    p.L = int(sc.radix.size());
    if (p.L > 4) throw bad_configuration("bbfft-cuda planner: more than 4 stages");
    for (int s = 0; s < 4; ++s) p.radix[s] = s < p.L ? sc.radix[s] : 1;
    p.T = sc.T;
    bool has_direct = false;
    for (int s = 0; s < p.L; ++s) has_direct = has_direct || direct_radix(p.radix[s]);
    if (p.L == 1 && !has_direct) p.T = 1;
    if (has_direct && p.T < 2) p.T = std::min(64, p.N); // outputs are dealt to the threads of a transform

    // ---- M == 1, single thread per transform: let the lanes walk k
    p.klanes = false;
    if (prob.M == 1 && p.T == 1) {
        p.klanes = true;
        p.ML = 32;
    }
    if (tune.count("KL")) {
        p.klanes = std::atoi(tune["KL"].c_str()) != 0;
        if (p.klanes && prob.M != 1) throw std::runtime_error("bbfft-cuda planner: KL needs M == 1");
    }

    // ---- CTA size
    int target_threads = 256;
    int bh = std::max(1, target_threads / (p.ML * p.T));
    // keep shared memory per CTA modest so that several CTAs share an SM
    while (bh > 1 && std::size_t(p.ML) * bh * (p.N + 2) * elem_bytes > 48 * 1024) bh /= 2;
    // do not make CTAs bigger than the batch
    std::uint64_t kunits = (p.mode == k_r2c_double || p.mode == k_c2r_double) ? (prob.K + 1) / 2 : prob.K;
    if (p.klanes) {
        while (bh > 1 && std::uint64_t(p.ML) * (bh / 2) >= kunits) bh /= 2;
    } else {
        while (bh > 1 && std::uint64_t(bh / 2) >= kunits) bh /= 2;
    }
    p.BH = bh;
    if (tune.count("BH")) {
        p.BH = std::max(1, std::atoi(tune["BH"].c_str()));
        while (p.BH > 1 && !p.klanes && std::uint64_t(p.BH / 2) >= kunits) p.BH /= 2;
    }
    p.threads = p.ML * p.T * p.BH;
    if (p.threads > dev.max_threads_per_block) {
        throw bad_configuration("bbfft-cuda planner: CTA too large");
    }

    // ---- staging: go through shared memory whenever the direct access would be poorly coalesced
    const int seg_bytes = 32;
    if (p.klanes) {
        p.load_staged = true;
        p.store_staged = true;
    } else if (p.ML == 1) {
        // lanes walk t: loads touch T consecutive elements, stores are digit-reversed
        p.load_staged = p.T * elem_bytes < seg_bytes;
        p.store_staged = true;
    } else {
        bool rows_ok = std::uint64_t(p.ML) * elem_bytes >= std::uint64_t(seg_bytes) && prob.M % p.ML == 0;
        p.load_staged = !rows_ok && prob.M * elem_bytes < 64;
        p.store_staged = !rows_ok && prob.M * elem_bytes < 64;
    }
    if (p.mode != k_c2c) {
        // Real transforms: the spectrum side is always touched pair-wise (i, N-i) by consecutive
        // threads, which is coalesced; only the real side of an M == 1 tensor needs help.
        // PAIR=0: the variant for real tensors that are not aligned to a complex word (plan.cpp)
        const bool no_cb = prob.cb_load.empty() && prob.cb_store.empty() &&
                           !(tune.count("PAIR") && std::atoi(tune["PAIR"].c_str()) == 0);
        p.pair_load = p.mode == k_r2c_half && prob.M == 1 && prob.is1 == 1 && prob.is2 % 2 == 0 && no_cb;
        p.pair_store = p.mode == k_c2r_half && prob.M == 1 && prob.os1 == 1 && prob.os2 % 2 == 0 && no_cb;
        p.load_staged = false;
        p.store_staged = false;
        if (p.mode == k_r2c_half) {
            p.load_staged = p.klanes && p.pair_load;
            p.store_staged = p.klanes;
        } else if (p.mode == k_c2r_half) {
            p.load_staged = p.klanes;
            p.store_staged = prob.M == 1 && p.pair_store;
        }
    }
    // user callbacks see every element exactly once in either path, so staging stays legal.
    // (real transforms stage whole aligned complex words: only for M == 1)
    const bool stage_override_ok = p.mode == k_c2c || prob.M == 1;
    if (tune.count("LD") && stage_override_ok) p.load_staged = std::atoi(tune["LD"].c_str()) != 0;
    if (tune.count("ST") && stage_override_ok) p.store_staged = std::atoi(tune["ST"].c_str()) != 0;
    // Real pre/post pass: fused into the first/last stage (a thread runs a sub-FFT and its mirror,
    // no extra trip through shared memory) when the batch lanes walk m.  With t-lanes (M == 1) the
    // mirrored units leave half of every warp idle in the stage that issues the global loads;
    // measured on config 3 (c2r f32 N=256 M=1): 3519 GB/s fused vs 5400 GB/s as a separate pass
    // (profiles/r01e_real_sweep.jsonl vs r01d), so those keep the separate pass.
    p.real_fused = p.mode != k_c2c && !p.load_staged && !p.store_staged && p.ML > 1;
    if (tune.count("RF")) p.real_fused = p.mode != k_c2c && !p.load_staged && !p.store_staged && std::atoi(tune["RF"].c_str()) != 0;
    if (p.mode != k_c2c) {
        // the fused pass runs its stage as in-register butterflies: not for a direct (large prime) stage
        const bool r2c_mode = p.mode == k_r2c_half || p.mode == k_r2c_double;
        if (direct_radix(p.radix[r2c_mode ? p.L - 1 : 0])) p.real_fused = false;
    }

    choose_smem_layout(p);
    if (tune.count("PADK") || tune.count("ROW")) {
        if (tune.count("PADK")) p.PADK = std::atoi(tune["PADK"].c_str());
        int rowlen = p.N + ((p.mode == k_r2c_half || p.mode == k_c2r_half) ? 1 : 0);
        int padded = (p.PADK > 0 ? (rowlen - 1) + (rowlen - 1) / p.PADK : rowlen - 1) + 1;
        p.ROW = std::max(p.LL * padded, tune.count("ROW") ? std::atoi(tune["ROW"].c_str()) : 0);
        int groups = (p.batch_per_cta() + p.LL - 1) / p.LL;
        p.smem_bytes = std::size_t(groups) * p.ROW * elem_bytes;
    }
    if (p.smem_bytes > dev.max_smem_per_block) {
        // long transforms with many batch lanes: halve the lanes (shorter rows per access, same
        // kernel family) until the CTA's rows fit, unless the caller pinned the lane count
        auto given = parse_tune(tune_str);
        if (p.ML > 1 && !p.klanes && prob.M > 1 && (!given.count("ML") || given.count("MLAUTO"))) {
            // (MLAUTO marks a lane count chosen here, not by the caller: it may be narrowed again)
            std::string narrower = tune_str + (tune_str.empty() ? "" : ",") + "MLAUTO=1,ML=" + std::to_string(p.ML / 2);
            return plan_kernel_1d(prob, dev, narrower);
        }
        throw bad_configuration("bbfft-cuda planner: shared memory demand too large");
    }
    // Resident CTAs: a stage-synchronised CTA cannot overlap its own load and compute phases, so
    // ask for 2-4 CTAs per SM whenever the register cap that implies still fits the butterflies.
    // The cap is exact (reg_cap models the four register-file partitions of an SM) and is given to
    // the compiler as __maxnreg__, so the occupancy is a planner decision and not an accident of
    // the register allocator (NVRTC and nvcc differ by a few registers on the same source).
    p.min_blocks = 1;
    {
        int regs_complex = 0;
        for (int s = 0; s < p.L; ++s) {
            int nsub = p.N / p.radix[s];
            int cnt = (nsub + p.T - 1) / p.T;
            if (direct_radix(p.radix[s])) {
                regs_complex = std::max(regs_complex, (p.N + p.T - 1) / p.T + 2);
                continue;
            }
            regs_complex = std::max(regs_complex, (s == 0 ? cnt : 1) * p.radix[s]);
        }
        if (p.real_fused) {
            // fused real pre / post pass: a sub-FFT and its mirror are live together
            const bool r2c = p.mode == k_r2c_half || p.mode == k_r2c_double;
            const int rm = p.radix[r2c ? p.L - 1 : 0];
            const int units = (p.N / rm) / 2 + 1;
            const int cntu = r2c ? 1 : (units + p.T - 1) / p.T;
            regs_complex = std::max(regs_complex, 2 * cntu * rm);
        }
        // measured: unconstrained kernels use ~1.6 registers per live 32-bit word + 20
        int words = (p.fp == 4 ? 2 : 4) * regs_complex;
        int need = words + words / 4 + 32;
        for (int mb = 4; mb >= 2; --mb) {
            bool fits = reg_cap(p.threads, mb) >= need && p.threads * mb <= 2048 &&
                        (p.smem_bytes + 1024) * std::size_t(mb) <= dev.smem_per_sm;
            if (fits) {
                p.min_blocks = mb;
                break;
            }
        }
    }
    if (tune.count("MB")) p.min_blocks = std::max(1, std::atoi(tune["MB"].c_str()));
    p.max_regs = reg_cap(p.threads, p.min_blocks);

    p.chained = tune.count("CH") && std::atoi(tune["CH"].c_str()) != 0;
    // packed fp32 adds: a per-kernel choice (the stub defines BBK_F32X2 for its translation unit; bit-identical
    // results, 3-5 % faster for the add-heavy 7-point sizes, slower for others: profiles/r02v_x2.log)
    p.x2 = p.fp == 4 && !p.chained && (packed_f32() || (tune.count("X2") && std::atoi(tune["X2"].c_str()) != 0));
    if (tune.count("X2") && std::atoi(tune["X2"].c_str()) == 0) p.x2 = false;
    // real in-place transforms need one CTA to own every m of a k slice, like the reference
    // (src/base/generator/small_batch_fft.cpp:38, factor2_slm_fft.cpp:63)
    plan.inplace_unsupported = real && !p.klanes && std::uint64_t(p.ML) < prob.M;

    plan.identifier = make_identifier(p);
    plan.source = emit_stub(p, plan.identifier, prob.cb_source);
    plan.twiddle = make_twiddles(p);
    return plan;
}

// ------------------------------------------------------------------------------------------
// identifier
// ------------------------------------------------------------------------------------------
// experiment switch: packed fp32 adds (add.f32x2) in the kernels of this process (BBFFT_CUDA_F32X2=1)
static bool packed_f32() {
    char const *e = std::getenv("BBFFT_CUDA_F32X2");
    return e && *e == '1';
}

std::string make_identifier(kernel_params const &p) {
    // K is deliberately not part of the key (reference: jit_cache key = identifier + device,
    // include/bbfft/jit_cache.hpp:23-41; K is a run-time argument)
    static char const *mode_names[] = {"c2c", "r2ch", "c2rh", "r2cd", "c2rd"};
    std::ostringstream os;
    os << "bbfft_" << mode_names[p.mode] << (p.dir < 0 ? "_m1" : "_p1") << "_f" << (p.fp * 8)
       << "_M" << p.M << "_N" << p.nreal << "_r";
    for (int s = 0; s < p.L; ++s) os << (s ? "x" : "") << p.radix[s];
    os << "_T" << p.T << "_ML" << p.ML << "_BH" << p.BH << "_mb" << p.min_blocks << "_kl" << int(p.klanes) << "_ld"
       << int(p.load_staged) << "_st" << int(p.store_staged) << "_pk" << p.PADK << "_row" << p.ROW
       << "_is" << p.is1 << "_" << p.is2 << "_os" << p.os1 << "_" << p.os2;
    if (p.mode != k_c2c) os << "_pl" << int(p.pair_load) << "_ps" << int(p.pair_store) << "_rf" << int(p.real_fused);
    if (p.chained) os << "_ch";
    if (!p.cb_load.empty()) os << "_" << p.cb_load;
    if (!p.cb_store.empty()) os << "_" << p.cb_store;
    if (p.x2) os << "_x2";
    std::string s = os.str();
    for (auto &c : s) {
        if (c == '-') c = 'n';
    }
    return s;
}

// ------------------------------------------------------------------------------------------
// twiddles
// ------------------------------------------------------------------------------------------
static void unit_root(long num, long den, int dir, double &re, double &im) {
    // exp(dir * 2 pi i num/den), evaluated in long double with octant reduction
    num %= den;
    if (num < 0) num += den;
    const long double tau = 6.283185307179586476925286766559005768L;
    // reduce to first octant for symmetric, accurate values
    long n8 = num * 8;
    int oct = int(n8 / den);
    long double c, s;
    long double ang;
    switch (oct) {
    case 0:
        ang = tau * num / den;
        c = cosl(ang);
        s = sinl(ang);
        break;
    case 1:
        ang = tau * (den - 4 * num) / (4.0L * den); // pi/2 - x
        c = sinl(ang);
        s = cosl(ang);
        break;
    case 2:
        ang = tau * (4 * num - den) / (4.0L * den); // x - pi/2
        c = -sinl(ang);
        s = cosl(ang);
        break;
    case 3:
        ang = tau * (den - 2 * num) / (2.0L * den); // pi - x
        c = -cosl(ang);
        s = sinl(ang);
        break;
    case 4:
        ang = tau * (2 * num - den) / (2.0L * den); // x - pi
        c = -cosl(ang);
        s = -sinl(ang);
        break;
    case 5:
        ang = tau * (3 * den - 4 * num) / (4.0L * den); // 3pi/2 - x
        c = -sinl(ang);
        s = -cosl(ang);
        break;
    case 6:
        ang = tau * (4 * num - 3 * den) / (4.0L * den); // x - 3pi/2
        c = sinl(ang);
        s = -cosl(ang);
        break;
    default:
        ang = tau * (den - num) / (long double)den; // 2pi - x
        c = cosl(ang);
        s = -sinl(ang);
        break;
    }
    re = double(c);
    im = double(dir < 0 ? -s : s);
}

static std::vector<int> tw_offsets(kernel_params const &p, int *total) {
    std::vector<int> off(4, 0);
    int acc = 0;
    int NS = p.N;
    for (int s = 0; s < p.L; ++s) {
        off[s] = acc;
        int R = p.radix[s];
        int NS1 = NS / R;
        if (s < p.L - 1) acc += (R - 1) * NS1;
        NS = NS1;
    }
    if (total) *total = acc;
    return off;
}

// first complex element of each direct stage's root table (0 for register stages)
static std::vector<int> tw_dir_offsets(kernel_params const &p) {
    int total = 0;
    tw_offsets(p, &total);
    int acc = total + ((p.mode == k_r2c_half || p.mode == k_c2r_half) ? p.N + 1 : 0);
    std::vector<int> off(4, 0);
    for (int s = 0; s < p.L; ++s) {
        if (!direct_radix(p.radix[s])) continue;
        off[s] = acc;
        acc += p.radix[s];
    }
    return off;
}

std::vector<double> make_twiddles(kernel_params const &p) {
    std::vector<double> tw;
    int NS = p.N;
    for (int s = 0; s + 1 < p.L; ++s) {
        int R = p.radix[s];
        int NS1 = NS / R;
        for (int q = 1; q < R; ++q) {
            for (int n2 = 0; n2 < NS1; ++n2) {
                double re, im;
                unit_root(long(n2) * q, NS, p.dir, re, im);
                tw.push_back(re);
                tw.push_back(im);
            }
        }
        NS = NS1;
    }
    if (p.mode == k_r2c_half || p.mode == k_c2r_half) {
        // post/pre twiddles for the half-length trick: w_{2N}^{dir*i}, i = 0..N/2
        // (reference host table: src/common/algorithm/factor2_slm_fft.hpp:81-88)
        // (the fused first/last stages pair i with N-i for every i, so the table runs to N)
        for (int i = 0; i <= p.N; ++i) {
            double re, im;
            unit_root(i, 2L * p.N, p.dir, re, im);
            tw.push_back(re);
            tw.push_back(im);
        }
    }
    // direct stages: w_R^(dir k), k = 0..R-1 (after the stage and real tables; offsets: tw_dir_offsets)
    for (int s = 0; s < p.L; ++s) {
        if (!direct_radix(p.radix[s])) continue;
        for (int k = 0; k < p.radix[s]; ++k) {
            double re, im;
            unit_root(k, p.radix[s], p.dir, re, im);
            tw.push_back(re);
            tw.push_back(im);
        }
    }
    if (tw.empty()) {
        tw.push_back(1.0);
        tw.push_back(0.0);
    }
    return tw;
}

// ------------------------------------------------------------------------------------------
// stub
// ------------------------------------------------------------------------------------------
static void emit_w_table(std::ostringstream &os, int R) {
    os << "struct W" << R << " {\n    static constexpr int n = " << R << ";\n";
    for (int part = 0; part < 2; ++part) {
        os << "    static BBK_CE double " << (part ? "s" : "c") << "(int k) {\n        constexpr double t["
           << R << "] = {";
        for (int k = 0; k < R; ++k) {
            double re, im;
            unit_root(k, R, +1, re, im);
            char buf[64];
            std::snprintf(buf, sizeof(buf), "%.17g", part ? im : re);
            os << (k ? ", " : "") << buf;
        }
        os << "};\n        return t[k];\n    }\n";
    }
    os << "};\n";
}

std::string emit_stub(kernel_params const &p, std::string const &identifier,
                      std::string const &cb_source) {
    std::ostringstream os;
    const char *real = p.fp == 4 ? "float" : "double";
    const char *vec = p.fp == 4 ? "float2" : "double2";
    os << "// generated by bbfft-cuda planner -- do not edit\n";
    if (p.x2) os << "#define BBK_F32X2 1\n";
    os << "#include \"bbfft_kernels.cuh\"\n";
    // every stub lives in its own namespace so that several can share a translation unit
    os << "namespace stub_" << identifier << " {\n";
    if (!cb_source.empty()) {
        os << "// ---- user callbacks (reference: include/bbfft/user_module.hpp:24-37)\n";
        os << cb_source << "\n";
    }
    std::set<int> rs;
    for (int s = 0; s < p.L; ++s) rs.insert(p.radix[s]);
    for (int r : rs) {
        if (direct_radix(r)) {
            os << "struct W" << r << " {\n    static constexpr int n = " << r << ";\n};\n"; // roots come from the table
        } else {
            emit_w_table(os, r);
        }
    }
    auto doff = tw_dir_offsets(p);
    int tw_total = 0;
    auto off = tw_offsets(p, &tw_total);
    os << "struct C {\n";
    os << "    using real_t = " << real << ";\n";
    os << "    static constexpr int N = " << p.N << ", NREAL = " << p.nreal << ", DIR = " << p.dir
       << ", MODE = " << p.mode << ", L = " << p.L << ", T = " << p.T << ", ML = " << p.ML
       << ", BH = " << p.BH << ";\n";
    os << "    static constexpr bool KLANES = " << (p.klanes ? "true" : "false")
       << ", LOAD_STAGED = " << (p.load_staged ? "true" : "false")
       << ", STORE_STAGED = " << (p.store_staged ? "true" : "false")
       << ", PAIR_LOAD = " << (p.pair_load ? "true" : "false")
       << ", PAIR_STORE = " << (p.pair_store ? "true" : "false")
       << ", REAL_FUSED = " << (p.real_fused ? "true" : "false") << ";\n";
    os << "    static constexpr bbk::u64 M = " << p.M << "ull;\n";
    os << "    static constexpr int LL = " << p.LL << ", PADK = " << p.PADK << ", ROW = " << p.ROW
       << ", TW_REAL = " << tw_total << ";\n";
    os << "    static BBK_CE int radix(int s) {\n        constexpr int r[4] = {" << p.radix[0] << ", "
       << p.radix[1] << ", " << p.radix[2] << ", " << p.radix[3] << "};\n        return r[s];\n    }\n";
    os << "    static BBK_CE int tw_off(int s) {\n        constexpr int r[4] = {" << off[0] << ", "
       << off[1] << ", " << off[2] << ", " << off[3] << "};\n        return r[s];\n    }\n";
    os << "    static BBK_CE bool direct(int s) {\n        constexpr bool r[4] = {" << (direct_radix(p.radix[0]) ? "true" : "false")
       << ", " << (direct_radix(p.radix[1]) ? "true" : "false") << ", " << (direct_radix(p.radix[2]) ? "true" : "false") << ", "
       << (direct_radix(p.radix[3]) ? "true" : "false") << "};\n        return r[s];\n    }\n";
    os << "    static BBK_CE int tw_dir(int s) {\n        constexpr int r[4] = {" << doff[0] << ", " << doff[1] << ", " << doff[2]
       << ", " << doff[3] << "};\n        return r[s];\n    }\n";
    // per-stage table type
    os << "    template <int S, int Dummy = 0> struct WRsel;\n";
    for (int s = 0; s < p.L; ++s) {
        os << "    template <int Dummy> struct WRsel<" << s << ", Dummy> { using type = W" << p.radix[s]
           << "; };\n";
    }
    os << "    template <int S> using WR = typename WRsel<S>::type;\n";
    os << "    static BBK_DEV bbk::i64 is1(bbk::args const &) { return " << p.is1 << "ll; }\n";
    os << "    static BBK_DEV bbk::i64 is2(bbk::args const &) { return " << p.is2 << "ll; }\n";
    os << "    static BBK_DEV bbk::i64 os1(bbk::args const &) { return " << p.os1 << "ll; }\n";
    os << "    static BBK_DEV bbk::i64 os2(bbk::args const &) { return " << p.os2 << "ll; }\n";
    // element accessors; user callbacks are spliced in here
    // (reference: callback_accessor, src/base/generator/tensor_accessor.cpp:34-55)
    const bool in_real = p.mode == k_r2c_half || p.mode == k_r2c_double;
    const bool out_real = p.mode == k_c2r_half || p.mode == k_c2r_double;
    if (p.cb_load.empty() && p.chained) {
        os << "    static BBK_DEV bbk::cx<real_t> ld(const void *in, bbk::u64 off) {\n"
              "        return bbk::ldcg_cx(reinterpret_cast<const bbk::cx<real_t> *>(in) + off);\n    }\n";
        os << "    static BBK_DEV real_t ldr(const void *in, bbk::u64 off) {\n"
              "        return reinterpret_cast<const real_t *>(in)[off];\n    }\n";
    } else if (p.cb_load.empty()) {
        os << "    static BBK_DEV bbk::cx<real_t> ld(const void *in, bbk::u64 off) {\n"
              "        return reinterpret_cast<const bbk::cx<real_t> *>(in)[off];\n    }\n";
        os << "    static BBK_DEV real_t ldr(const void *in, bbk::u64 off) {\n"
              "        return reinterpret_cast<const real_t *>(in)[off];\n    }\n";
    } else if (in_real) {
        os << "    static BBK_DEV bbk::cx<real_t> ld(const void *, bbk::u64) { return bbk::cx<real_t>{}; }\n";
        os << "    static BBK_DEV real_t ldr(const void *in, bbk::u64 off) {\n        return " << p.cb_load
           << "(reinterpret_cast<" << real << " *>(const_cast<void *>(in)), off);\n    }\n";
    } else {
        os << "    static BBK_DEV bbk::cx<real_t> ld(const void *in, bbk::u64 off) {\n        " << vec
           << " r = " << p.cb_load << "(reinterpret_cast<" << vec
           << " *>(const_cast<void *>(in)), off);\n        return bbk::cx<real_t>{r.x, r.y};\n    }\n";
        os << "    static BBK_DEV real_t ldr(const void *, bbk::u64) { return real_t(0); }\n";
    }
    if (p.cb_store.empty()) {
        os << "    static BBK_DEV void st(void *out, bbk::u64 off, bbk::cx<real_t> v) {\n"
              "        reinterpret_cast<bbk::cx<real_t> *>(out)[off] = v;\n    }\n";
        os << "    static BBK_DEV void str(void *out, bbk::u64 off, real_t v) {\n"
              "        reinterpret_cast<real_t *>(out)[off] = v;\n    }\n";
    } else if (out_real) {
        os << "    static BBK_DEV void st(void *, bbk::u64, bbk::cx<real_t>) {}\n";
        os << "    static BBK_DEV void str(void *out, bbk::u64 off, real_t v) {\n        " << p.cb_store
           << "(reinterpret_cast<" << real << " *>(out), off, v);\n    }\n";
    } else {
        os << "    static BBK_DEV void st(void *out, bbk::u64 off, bbk::cx<real_t> v) {\n        " << vec
           << " r;\n        r.x = v.x;\n        r.y = v.y;\n        " << p.cb_store << "(reinterpret_cast<"
           << vec << " *>(out), off, r);\n    }\n";
        os << "    static BBK_DEV void str(void *, bbk::u64, real_t) {}\n";
    }
    os << "    static constexpr bool HAS_LOAD_CALLBACK = " << (p.cb_load.empty() ? "false" : "true") << ";\n";
    os << "    static constexpr bool HAS_CALLBACKS = "
       << ((p.cb_load.empty() && p.cb_store.empty()) ? "false" : "true") << ";\n";
    os << "};\n} // namespace stub_" << identifier << "\n";
    if (!p.chained) {
        os << "extern \"C\" BBK_GLOBAL void BBK_KERNEL(" << p.threads << ", " << p.max_regs << ") "
           << identifier << "(bbk::args a) {\n    bbk::fft1d<stub_" << identifier << "::C>(a);\n}\n";
    }
    os << "#ifdef BBFFT_OCL_COMPAT\n#undef float2\n#undef double2\n#undef BBFFT_OCL_COMPAT\n#endif\n";
    return os.str();
}

// ------------------------------------------------------------------------------------------
// fused 2d tile kernel
// ------------------------------------------------------------------------------------------
namespace {

// fewest stages with radices <= rmax (a prime factor above rmax is its own radix), then the
// cheapest butterflies, then the most balanced split
std::vector<int> tile_radices(int N, int rmax) {
    // measured on the B200 (tools/tune_tile.py, profiles/r01g_tile_tune.json): 64 = 4 x 16 beats the
    // balanced 8 x 8 in the tile kernel by 11 % (fp64, 5215 vs 4701 GB/s) and 13 % (fp32)
    if (N == 64 && rmax >= 16) return {4, 16};
    // 32 = 2 x 16 beats 4 x 8: +19 % fp32, +2 % fp64 (profiles/r01h_tile_tune.json)
    if (N == 32 && rmax >= 16) return {2, 16};
    std::vector<std::vector<int>> facs;
    std::vector<int> cur;
    enum_factorizations(N, std::max(rmax, max_prime(N)), 4, cur, facs);
    std::vector<int> best;
    double best_cost = 1e30;
    for (auto const &f : facs) {
        double cost = 100.0 * f.size();
        int mx = 0;
        for (int r : f) {
            cost += radix_cost(r) * 0.1;
            mx = std::max(mx, r);
        }
        cost += 0.01 * mx;
        if (cost < best_cost) {
            best_cost = cost;
            best = f;
        }
    }
    if (best.empty()) throw bad_configuration("bbfft-cuda planner: no tile factorization for N=" + std::to_string(N));
    return best;
}

int tile_regs_complex(tile_params const &p) {
    int regs = 0;
    for (tile_pass_params const *q : {&p.a, &p.b}) {
        for (int s = 0; s < q->L; ++s) {
            long total = long(q->S) * (q->N / q->radix[s]) * q->O;
            int cnt = int((total + p.threads - 1) / p.threads);
            regs = std::max(regs, cnt * q->radix[s]);
        }
    }
    return regs;
}

int tile_phys(int lin, int padk) { return padk > 0 ? lin + lin / padk : lin; }

int pass_bin_of_sub(tile_pass_params const &q, int u) {
    int digits[4] = {0, 0, 0, 0};
    int rem = u;
    for (int s = q.L - 2; s >= 0; --s) {
        digits[s] = rem % q.radix[s];
        rem /= q.radix[s];
    }
    int k = 0;
    for (int s = q.L - 2; s >= 0; --s) k = k * q.radix[s] + digits[s];
    return k;
}

// excess shared-memory wavefronts of warps 0 and 1 over every smem phase of the tile kernel
long tile_layout_score(tile_params const &p, int padk) {
    const int elem_bytes = 2 * p.fp;
    kernel_params dummy;
    layout_eval ev{dummy, elem_bytes};
    long excess = 0;
    int pass_no = 0;
    for (tile_pass_params const *q : {&p.a, &p.b}) {
        for (int s = 0; s < q->L; ++s) {
            const int R = q->radix[s];
            int NS = q->N;
            for (int i = 0; i < s; ++i) NS /= q->radix[i];
            const int NS1 = NS / R, NSUB = q->N / R;
            const long total = long(q->S) * NSUB * q->O;
            const int cnt = int((total + p.threads - 1) / p.threads);
            const bool last = s == q->L - 1;
            const bool reads = !(pass_no == 0 && s == 0);
            const bool writes_inplace = !last;
            const bool writes_sorted = last && pass_no == 0;
            for (int warp = 0; warp < std::min(2, (p.threads + 31) / 32); ++warp) {
                for (int i = 0; i < std::min(cnt, 2); ++i) {
                    for (int j = 0; j < R; ++j) {
                        std::vector<int> inpl(32, -1), sorted(32, -1);
                        for (int l = 0; l < 32; ++l) {
                            long id = long(warp) * 32 + l + long(p.threads) * i;
                            if (id >= total) continue;
                            int lo = int(id % q->S);
                            long r = id / q->S;
                            int u = int(r % NSUB), hi = int(r / NSUB);
                            const int pitch = q->pitch ? q->pitch : q->S * q->N;
                            int base = lo + q->S * ((u / NS1) * NS + u % NS1) + pitch * hi;
                            inpl[l] = tile_phys(base + q->S * NS1 * j, padk);
                            if (last) {
                                int b2 = lo + q->S * pass_bin_of_sub(*q, u) + pitch * hi;
                                sorted[l] = tile_phys(b2 + q->S * (q->N / R) * j, padk);
                            }
                        }
                        int e = ev.wavefronts(inpl) - ev.ideal(inpl);
                        excess += e * ((reads ? 1 : 0) + (writes_inplace ? 1 : 0));
                        if (writes_sorted) excess += ev.wavefronts(sorted) - ev.ideal(sorted);
                    }
                }
            }
        }
        ++pass_no;
    }
    return excess;
}

std::size_t tile_smem_bytes(std::uint64_t tile, int padk, int fp) {
    return std::size_t(tile_phys(int(tile - 1), padk) + 1) * 2 * fp;
}

} // namespace

int tile_cluster_limit() {
    if (char const *e = std::getenv("BBFFT_CUDA_TILE_CLUSTER")) {
        int want = std::atoi(e);
        if (want >= 1 && want <= 8 && (want & (want - 1)) == 0) return want;
    }
    return 1;
}

bool tile_fusable(problem_2d const &prob, device_props const &dev, int max_cluster) {
    if (prob.N1 < 2 || prob.N2 < 2) return false;
    if (prob.real != 0) {
        // one CTA per spectrum tile; the half-length trick needs an even real length
        if (prob.N1 % 2 != 0 || prob.N1 < 4) return false;
        const std::uint64_t tile = prob.M * (prob.N1 / 2) * prob.N2; // (columns 0 and N1/2 share a tile column)
        if (tile < 1024 || tile > (1u << 20)) return false;
        if (tile_smem_bytes(tile + prob.M * prob.N2, 0, prob.fp) > std::min<std::size_t>(dev.max_smem_per_block, 200 * 1024)) return false;
        return std::max(max_prime(int(prob.N1 / 2)), max_prime(int(prob.N2))) <= 31;
    }
    const std::uint64_t tile = prob.M * prob.N1 * prob.N2;
    if (tile < 1024 || tile > (1u << 20)) return false;
    // padding candidates that would not fit are skipped by the layout search; a tile that does not fit one
    // CTA is split over a thread-block cluster of up to 8 (plan_kernel_2d)
    std::uint64_t local = tile;
    for (int cl = 1; cl < max_cluster && tile_smem_bytes(local, 0, prob.fp) > std::min<std::size_t>(dev.max_smem_per_block, 200 * 1024) &&
                     prob.N1 % (2 * cl) == 0 && prob.N2 % (2 * cl) == 0;
         cl *= 2) {
        local = tile / (2 * cl);
    }
    if (tile_smem_bytes(local, 0, prob.fp) > std::min<std::size_t>(dev.max_smem_per_block, 200 * 1024)) return false;
    int big = std::max(max_prime(int(prob.N1)), max_prime(int(prob.N2)));
    return big <= 31;
}

tile_plan plan_kernel_2d(problem_2d const &prob, device_props const &dev, std::string const &tune_str) {
    auto tune = parse_tune(tune_str);
    tile_plan plan;
    tile_params &p = plan.p;
    p.fp = prob.fp;
    p.dir = prob.dir;
    p.M = prob.M;
    p.N1 = prob.N1;
    p.N2 = prob.N2;
    const bool is_real = prob.real != 0;
    const std::uint64_t NC = prob.N1 / 2 + 1; // spectrum columns of a real tile
    if (is_real && (prob.N1 % 2 != 0 || prob.N1 < 4)) throw bad_configuration("bbfft-cuda planner: fused real tiles need an even N1");
    p.real = prob.real;
    p.spectrum_n1 = NC;
    // (real tiles: pass A is the half-length complex transform; N1c = its length, the tile holds NC columns)
    const std::uint64_t N1c = is_real ? prob.N1 / 2 : prob.N1;
    const std::uint64_t tile = prob.M * N1c * prob.N2; // elements of the shared-memory tile
    p.tile_stride = prob.tile_stride ? prob.tile_stride : (is_real ? prob.M * NC * prob.N2 : tile);
    p.real_row = prob.inplace ? 2 * prob.M * NC : prob.M * prob.N1;
    p.real_tile_stride = p.real_row * prob.N2;
    const int rmax = 16;
    auto ra = tune.count("RA") ? parse_radices(tune["RA"]) : tile_radices(int(N1c), rmax);
    auto rb = tune.count("RB") ? parse_radices(tune["RB"]) : tile_radices(int(prob.N2), rmax);
    auto fill = [](tile_pass_params &q, std::vector<int> const &r, int N, int S, int O) {
        int prod = 1;
        for (int x : r) prod *= x;
        if (prod != N || r.size() > 4 || r.empty()) {
            throw std::runtime_error("bbfft-cuda planner: tile radices do not multiply to N");
        }
        q.N = N;
        q.S = S;
        q.O = O;
        q.L = int(r.size());
        for (int s = 0; s < 4; ++s) q.radix[s] = s < q.L ? r[s] : 1;
    };
    // CTAs per tile (bbk::fft2d_tile_cluster: the tile split over a thread-block cluster, rows N2 / CL per CTA
    // in the row pass, columns N1 / CL in the column pass, gathered through distributed shared memory).
    // A switch, OFF by default (CL=<n> / BBFFT_CUDA_TILE_CLUSTER=<n>): it was built to put several tiles' CTAs
    // on one SM for the 128 KiB tiles that run one CTA per SM, and measured on the B200 it loses
    // (profiles/r02f_cluster.txt: 2d fp32 128x128 at 1 GiB 524 us with one CTA per tile, 545 us with CL=4,
    // 750 / 785 us with CL=2 / 8; 3d fp64 64^3 191 -> 200 us with CL=2) -- the remote gather and the two
    // cluster barriers cost more than the extra resident CTAs buy.  With the switch on, tiles of up to
    // 8 x 200 KiB are fused instead of falling back to one launch per mode.
    {
        int cl = 1;
        const bool chained_req = tune.count("CH") && std::atoi(tune["CH"].c_str()) != 0;
        auto ok = [&](int c) {
            return prob.N1 % c == 0 && prob.N2 % c == 0 && (tile / c) * 2 * std::uint64_t(prob.fp) >= 16 * 1024 &&
                   (prob.M * prob.N1 / c) * 2 * std::uint64_t(prob.fp) >= 128; // >= 128-byte runs in the final store
        };
        const int limit = (chained_req || is_real) ? 1 : tile_cluster_limit();
        // the smallest cluster that brings the CTA's share to 32 KiB (or makes the tile fit at all)
        while (cl < limit && (tile / cl) * 2 * std::uint64_t(prob.fp) > 32 * 1024 && ok(cl * 2)) cl *= 2;
        if (tune.count("CL")) {
            int want = std::atoi(tune["CL"].c_str());
            if (want < 1 || want > 8 || (want & (want - 1)) != 0 || prob.N1 % want != 0 || prob.N2 % want != 0 ||
                ((chained_req || is_real) && want != 1)) {
                throw bad_configuration("bbfft-cuda planner: bad tile cluster size");
            }
            cl = want;
        }
        p.cluster = cl;
    }
    const std::uint64_t local_tile = tile / std::uint64_t(p.cluster);
    // row pass over the CTA's N2 / CL rows; column pass over its N1 / CL columns (S = fastest run M * N1 / CL)
    fill(p.a, ra, int(N1c), int(prob.M), int(prob.N2) / p.cluster);
    fill(p.b, rb, int(prob.N2), int(prob.M * N1c) / p.cluster, 1);

    // threads: ~16 (fp32) / ~8-16 (fp64) complex elements per thread
    if (tune.count("TH")) {
        p.threads = std::atoi(tune["TH"].c_str());
    } else {
        // fp32 64 x 64 tiles: 128 threads with 32 elements each (four resident CTAs) measured 6273 GB/s
        // against 5603 GB/s for 256 threads (profiles/r01h_tile_tune.json)
        const std::uint64_t ept = (p.fp == 4 && local_tile == 4096) ? 32 : 16;
        int th = 64;
        while (th < dev.max_threads_per_block && std::uint64_t(th) * ept < local_tile) th *= 2;
        p.threads = th;
    }
    if (p.threads < 32 || p.threads > dev.max_threads_per_block || p.threads % 32) {
        throw std::runtime_error("bbfft-cuda planner: bad tile CTA size");
    }
    // padding
    {
        long best = -1;
        int best_padk = 0;
        for (int padk : {0, 64, 32, 16, 8}) {
            if (tile_smem_bytes(local_tile, padk, p.fp) > dev.max_smem_per_block) continue;
            long sc = tile_layout_score(p, padk) * 64 + (padk ? 64 / padk : 0);
            if (best < 0 || sc < best) {
                best = sc;
                best_padk = padk;
            }
        }
        p.PADK = best_padk;
    }
    if (tune.count("PADK")) p.PADK = std::atoi(tune["PADK"].c_str());
    p.smem_bytes = tile_smem_bytes(local_tile, p.PADK, p.fp);
    int scratch_off = 0;
    if (prob.real == 1) {
        // r2c: scratch column for the transform of the packed column (bbk::tile_r2c_unpack)
        scratch_off = int(p.smem_bytes / (2 * std::size_t(p.fp)));
        p.smem_bytes += std::size_t(prob.M * prob.N2) * 2 * std::size_t(p.fp);
    }
    if (p.smem_bytes > dev.max_smem_per_block) {
        throw bad_configuration("bbfft-cuda planner: tile does not fit into shared memory");
    }
    // resident CTAs
    {
        int words = (p.fp == 4 ? 2 : 4) * tile_regs_complex(p);
        // (real tiles: 16 elements per thread of a 512-thread CTA compile to 62-64 registers without spilling,
        // and two resident CTAs measured 0.56 against 0.44 of the HBM peak for 2d r2c fp32 128 x 128)
        int need = words + words / 4 + (is_real ? 24 : 32);
        int mb = int(std::min<std::size_t>({std::size_t(4), dev.smem_per_sm / (p.smem_bytes + 1024),
                                            std::size_t(2048 / p.threads)}));
        mb = std::max(mb, 1);
        while (mb > 1 && reg_cap(p.threads, mb) < need) --mb;
        p.min_blocks = mb;
    }
    if (tune.count("MB")) p.min_blocks = std::max(1, std::atoi(tune["MB"].c_str()));
    p.max_regs = reg_cap(p.threads, p.min_blocks);
    p.chained = tune.count("CH") && std::atoi(tune["CH"].c_str()) != 0;
    if (is_real && p.chained) throw bad_configuration("bbfft-cuda planner: fused real tiles are not chained");
    // Persistent grid + asynchronous load of the next tile (bbk::fft2d_tile_persistent): a switch
    // (PS=1 / BBFFT_CUDA_TILE_ASYNC=1), off by default.  Measured on the B200 (profiles/r02e_tile_pdl.txt)
    // it loses a little -- 2d fp32 128x128 at 1 GiB: 517 us against 499 us; 3d fp64 64^3: 204 against
    // 188 us -- because the one-tile-per-CTA kernel already overlaps the store drain of tile i with the
    // loads of the CTA that replaces it, so only the last stage's arithmetic is left to hide, and the
    // extra trip through shared memory plus the barrier cost more than that.
    p.persistent = false;
    if (char const *e = std::getenv("BBFFT_CUDA_TILE_ASYNC")) p.persistent = !p.chained && *e == '1';
    if (tune.count("PS")) p.persistent = !p.chained && std::atoi(tune["PS"].c_str()) != 0;
    if (p.cluster > 1 || is_real) p.persistent = false;
    // Staged pipeline (bbk::fft2d_tile_staged): SG=<rows> / BBFFT_CUDA_TILE_STAGE=<rows> (-1: as many rows of pass A
    // as the shared memory beside the tile holds at the planned number of resident CTAs).
    {
        // Default: on, with bulk copies, for tiles that run ONE CTA per SM and more than one tile per CTA -- there
        // nothing else overlaps the loads (2d fp32 128x128 at 1 GiB: 498 -> 483 us, profiles/r02v_tile.log); off
        // for smaller tiles, whose two to four resident CTAs overlap each other better than one CTA overlaps
        // itself (fp32 64x64: 0.99 of the HBM peak plain, 0.85 staged).
        long rows = 0;
        bool bulk_default = false;
        if (p.min_blocks == 1 && prob.K >= 2 * std::uint64_t(dev.sm_count)) {
            rows = -1;
            bulk_default = true;
        }
        if (char const *e = std::getenv("BBFFT_CUDA_TILE_STAGE")) rows = std::atol(e);
        if (tune.count("SG")) rows = std::atol(tune["SG"].c_str());
        if (p.chained || p.cluster > 1 || is_real) rows = 0;
        p.stage = 0;
        p.stage_off = 0;
        if (rows != 0) {
            const std::size_t elem = 2 * std::size_t(p.fp);
            const std::size_t row = std::size_t(p.a.S) * std::size_t(p.a.N); // elements of one row of pass A
            const std::size_t off = (p.smem_bytes / elem + 1) / 2 * 2;       // 16-byte aligned for fp32, too
            const std::size_t budget = std::min<std::size_t>(dev.max_smem_per_block,
                                                             dev.smem_per_sm / std::size_t(p.min_blocks) - 1024);
            long fit = budget > off * elem ? long((budget - off * elem) / (row * elem)) : 0;
            fit = std::min<long>(fit, long(p.a.O));
            if (rows < 0 || rows > fit) rows = fit;
            if (row % 2 != 0) rows = rows / 2 * 2;
            // (16 bytes behind the staging buffer hold the mbarrier of the bulk-copy flavour)
            while (rows > 0 && (off + std::size_t(rows) * row) * elem + 16 > budget) rows -= (row % 2 != 0) ? 2 : 1;
            if (rows > 0) {
                p.stage = int(std::size_t(rows) * row);
                p.stage_off = int(off);
                p.smem_bytes = (off + std::size_t(p.stage)) * elem + 16;
                p.persistent = true;
                p.bulk = bulk_default;
                if (char const *e = std::getenv("BBFFT_CUDA_TILE_BULK")) p.bulk = *e == '1';
                if (tune.count("BK")) p.bulk = std::atoi(tune["BK"].c_str()) != 0;
            }
        }
    }

    // identifier
    {
        std::ostringstream os;
        os << (p.real == 1 ? (prob.inplace ? "bbfft_r2c2di" : "bbfft_r2c2d") : p.real == 2 ? (prob.inplace ? "bbfft_c2r2di" : "bbfft_c2r2d") : "bbfft_c2c2d")
           << (p.dir < 0 ? "_m1" : "_p1") << "_f" << (p.fp * 8) << "_M" << p.M << "_N" << p.N1
           << "x" << p.N2 << "_ra";
        for (int s = 0; s < p.a.L; ++s) os << (s ? "x" : "") << p.a.radix[s];
        os << "_rb";
        for (int s = 0; s < p.b.L; ++s) os << (s ? "x" : "") << p.b.radix[s];
        os << "_th" << p.threads << "_mb" << p.min_blocks << "_pk" << p.PADK << "_ts" << p.tile_stride;
        if (p.chained) os << "_ch";
        if (p.persistent && !p.stage) os << "_ps";
        if (p.stage) os << "_sg" << p.stage << (p.bulk ? "b" : "");
        if (p.cluster > 1) os << "_cl" << p.cluster;
        if (p.fp == 4 && packed_f32()) os << "_x2";
        plan.identifier = os.str();
    }
    // twiddles: pass A stages, then pass B stages (same construction as the 1d table)
    std::vector<int> off_a(4, 0), off_b(4, 0);
    int tw_real = 0;
    {
        auto add = [&](tile_pass_params const &q, std::vector<int> &off) {
            int NS = q.N;
            for (int s = 0; s < q.L; ++s) {
                off[s] = int(plan.twiddle.size() / 2);
                int R = q.radix[s], NS1 = NS / R;
                if (s + 1 < q.L) {
                    for (int qq = 1; qq < R; ++qq) {
                        for (int n2 = 0; n2 < NS1; ++n2) {
                            double re, im;
                            unit_root(long(n2) * qq, NS, p.dir, re, im);
                            plan.twiddle.push_back(re);
                            plan.twiddle.push_back(im);
                        }
                    }
                }
                NS = NS1;
            }
        };
        add(p.a, off_a);
        add(p.b, off_b);
        if (is_real) {
            // split / merge twiddles of the half-length trick: w_{N1}^{dir i}, i = 0 .. N1/2 (as the 1d real kernels)
            tw_real = int(plan.twiddle.size() / 2);
            for (std::uint64_t i = 0; i <= N1c; ++i) {
                double re, im;
                unit_root(long(i), long(2 * N1c), p.dir, re, im);
                plan.twiddle.push_back(re);
                plan.twiddle.push_back(im);
            }
        }
        if (plan.twiddle.empty()) {
            plan.twiddle.push_back(1.0);
            plan.twiddle.push_back(0.0);
        }
    }
    // stub
    {
        std::ostringstream os;
        const char *real = p.fp == 4 ? "float" : "double";
        os << "// generated by bbfft-cuda planner -- do not edit\n";
        if (p.fp == 4 && packed_f32()) os << "#define BBK_F32X2 1\n";
        os << "#include \"bbfft_kernels.cuh\"\n";
        os << "namespace stub_" << plan.identifier << " {\n";
        std::set<int> rs;
        for (int s = 0; s < p.a.L; ++s) rs.insert(p.a.radix[s]);
        for (int s = 0; s < p.b.L; ++s) rs.insert(p.b.radix[s]);
        for (int r : rs) emit_w_table(os, r);
        auto emit_pass = [&](char const *name, tile_pass_params const &q, std::vector<int> const &off, long gs) {
            os << "struct " << name << " {\n    static constexpr int N = " << q.N << ", S = " << q.S << ", O = " << q.O
               << ", L = " << q.L << ", GS = " << gs << ", PITCH = " << (q.pitch ? q.pitch : q.S * q.N)
               << ", GPITCH = " << gs * q.N << ";\n";
            os << "    static BBK_CE int radix(int s) {\n        constexpr int r[4] = {" << q.radix[0] << ", "
               << q.radix[1] << ", " << q.radix[2] << ", " << q.radix[3] << "};\n        return r[s];\n    }\n";
            os << "    static BBK_CE int tw_off(int s) {\n        constexpr int r[4] = {" << off[0] << ", " << off[1]
               << ", " << off[2] << ", " << off[3] << "};\n        return r[s];\n    }\n";
            os << "    template <int S_, int Dummy = 0> struct WRsel;\n";
            for (int s = 0; s < q.L; ++s) {
                os << "    template <int Dummy> struct WRsel<" << s << ", Dummy> { using type = W" << q.radix[s]
                   << "; };\n";
            }
            os << "    template <int S_> using WR = typename WRsel<S_>::type;\n};\n";
        };
        emit_pass("PassA", p.a, off_a, long(p.a.S));
        emit_pass("PassB", p.b, off_b, long(prob.M * (is_real ? NC : prob.N1))); // rows of the whole tile in global memory
        os << "struct C {\n    using real_t = " << real << ";\n    using PA = PassA;\n    using PB = PassB;\n";
        os << "    static constexpr int DIR = " << p.dir << ", THREADS = " << p.threads << ", PADK = " << p.PADK
           << ", CL = " << p.cluster << ", N2L = " << (prob.N2 / std::uint64_t(p.cluster)) << ", ROWLEN = " << (prob.M * prob.N1)
           << ", STG = " << p.stage << ", STG_OFF = " << p.stage_off << ", BULK = " << (p.bulk ? 1 : 0)
           << ", REAL = " << p.real << ", TW_REAL = " << tw_real << ", SCR_OFF = " << scratch_off
           << ";\n    static constexpr bool PERSIST = " << (p.persistent ? "true" : "false")
           << ";\n    static constexpr bbk::u64 TILE_STRIDE = " << p.tile_stride << "ull, RTS = " << p.real_tile_stride
           << "ull, RROW = " << p.real_row << "ull;\n";
        if (p.chained) {
            os << "    static BBK_DEV bbk::cx<real_t> ld(const void *in, bbk::u64 off) {\n"
                  "        return bbk::ldcg_cx(reinterpret_cast<const bbk::cx<real_t> *>(in) + off);\n    }\n";
        } else {
            os << "    static BBK_DEV bbk::cx<real_t> ld(const void *in, bbk::u64 off) {\n"
                  "        return reinterpret_cast<const bbk::cx<real_t> *>(in)[off];\n    }\n";
        }
        os << "    static BBK_DEV void st(void *out, bbk::u64 off, bbk::cx<real_t> v) {\n"
              "        reinterpret_cast<bbk::cx<real_t> *>(out)[off] = v;\n    }\n";
        os << "};\n} // namespace stub_" << plan.identifier << "\n";
        if (!p.chained) {
            os << "extern \"C\" BBK_GLOBAL void BBK_KERNEL(" << p.threads << ", " << p.max_regs << ") ";
            if (p.cluster > 1) os << "BBK_CLUSTER(" << p.cluster << ") ";
            os << plan.identifier << "(bbk::args a) {\n    bbk::fft2d_tile<stub_" << plan.identifier << "::C>(a);\n}\n";
        }
        plan.source = os.str();
    }
    return plan;
}

// ------------------------------------------------------------------------------------------
// chain: all steps of an nd decomposition in one persistent kernel (bbk::chain)
// ------------------------------------------------------------------------------------------
bool plan_chain(std::vector<chain_step_problem> const &steps, device_props const &dev, chain_plan_t &out) {
    if (steps.size() < 2 || steps.size() > 3) return false;
    out = chain_plan_t{};
    // common CTA shape: the tile kernel's thread count when there is one, else 256
    int threads = 256;
    for (auto const &s : steps) {
        if (s.tile) {
            tile_plan tp = plan_kernel_2d(s.t, dev, "CH=1");
            threads = tp.p.threads;
        }
    }
    for (auto const &s : steps) {
        chain_step_plan sp;
        sp.tile = s.tile;
        if (s.tile) {
            sp.tp = plan_kernel_2d(s.t, dev, "CH=1,TH=" + std::to_string(threads));
            if (sp.tp.p.threads != threads || s.mult == 0) return false;
            sp.per_k = s.mult; // one CTA per tile
        } else {
            // r2c reads its real input around C::ld (pair loads): fine for the first step only, whose
            // input no other SM writes during the launch
            if (s.p.type == 1 && !out.steps.empty()) return false;
            // the step's own plan gives radices, threads per transform and the preferred lanes; lanes
            // and batch are re-chosen so that the CTA has the common thread count and never
            // straddles two slabs
            kernel_plan def = plan_kernel_1d(s.p, dev, "");
            if (def.p.klanes || def.p.L < 2) return false;
            std::uint64_t kunits = s.mult; // k slices of the step per slab
            if (def.p.mode == k_r2c_double || def.p.mode == k_c2r_double) {
                if (kunits % 2) return false;
                kunits /= 2;
            }
            std::vector<int> ml_pref;
            if (s.p.M == 1) {
                ml_pref = {1};
            } else {
                for (int c = def.p.ML; c <= 32; c *= 2) ml_pref.push_back(c);
                for (int c = def.p.ML / 2; c >= 2; c /= 2) ml_pref.push_back(c);
            }
            // threads per transform: the plan's own, else more (fewer sub-FFTs per thread) up to
            // one sub-FFT of the smallest radix per thread
            int min_r = def.p.radix[0];
            for (int i = 1; i < def.p.L; ++i) min_r = std::min(min_r, def.p.radix[i]);
            int T = 0, ml = 0, bh = 0;
            for (int tc = def.p.T; tc <= def.p.N / min_r && ml == 0; ++tc) {
                if (threads % tc != 0) continue;
                const int lanes_total = threads / tc; // ML * BH
                for (int c : ml_pref) {
                    if (lanes_total % c != 0) continue;
                    const int b = lanes_total / c;
                    if (kunits % std::uint64_t(b) != 0) continue;
                    // real in-place capable layouts keep every m of a slice in one CTA
                    if (s.p.type != 0 && def.p.ML >= int(s.p.M) && c < int(s.p.M)) continue;
                    T = tc;
                    ml = c;
                    bh = b;
                    break;
                }
            }
            if (ml == 0) return false;
            std::ostringstream tn;
            tn << "CH=1,R=";
            for (int i = 0; i < def.p.L; ++i) tn << (i ? "x" : "") << def.p.radix[i];
            tn << ",T=" << T << ",ML=" << ml << ",BH=" << bh;
            sp.kp = plan_kernel_1d(s.p, dev, tn.str());
            if (sp.kp.p.threads != threads || sp.kp.p.klanes || sp.kp.p.ML != ml || sp.kp.p.BH != bh) return false;
            if (sp.kp.inplace_unsupported && !def.inplace_unsupported) return false;
            sp.per_k = (kunits / bh) * ((s.p.M + ml - 1) / ml);
        }
        out.steps.push_back(std::move(sp));
    }
    out.threads = threads;
    out.min_blocks = 8;
    for (auto const &sp : out.steps) {
        out.smem_bytes = std::max(out.smem_bytes, sp.tile ? sp.tp.p.smem_bytes : sp.kp.p.smem_bytes);
        out.min_blocks = std::min(out.min_blocks, sp.tile ? sp.tp.p.min_blocks : sp.kp.p.min_blocks);
    }
    while (out.min_blocks > 1 && ((out.smem_bytes + 1024) * std::size_t(out.min_blocks) > dev.smem_per_sm ||
                                  threads * out.min_blocks > 2048)) {
        --out.min_blocks;
    }
    if (out.smem_bytes > dev.max_smem_per_block) return false;
    out.max_regs = reg_cap(threads, out.min_blocks);
    std::ostringstream id;
    id << "bbfft_chain" << out.steps.size() << "_mb" << out.min_blocks;
    std::ostringstream stubs, os;
    stubs << "// generated by bbfft-cuda planner -- do not edit\n#include \"bbfft_kernels.cuh\"\n";
    std::uint64_t h = 1469598103934665603ull;
    std::set<std::string> emitted;
    for (auto &sp : out.steps) {
        std::string const &sid = sp.tile ? sp.tp.identifier : sp.kp.identifier;
        for (char c : sid) h = (h ^ static_cast<unsigned char>(c)) * 1099511628211ull;
        h = (h ^ sp.per_k) * 1099511628211ull;
        if (emitted.insert(sid).second) stubs << (sp.tile ? sp.tp.source : sp.kp.source);
        auto const &tw = sp.tile ? sp.tp.twiddle : sp.kp.twiddle;
        sp.tw_offset = int(out.twiddle.size() / 2);
        out.twiddle.insert(out.twiddle.end(), tw.begin(), tw.end());
    }
    {
        // the step identifiers are long: the chain's name carries a hash of them
        char buf[32];
        std::snprintf(buf, sizeof(buf), "_%016llx", static_cast<unsigned long long>(h));
        auto const &first = out.steps.front();
        id << (first.tile ? "_t" : "_p") << "_f" << ((first.tile ? first.tp.p.fp : first.kp.p.fp) * 8) << "_th"
           << threads << buf;
    }
    out.identifier = id.str();
    os << "namespace chain_" << out.identifier << " {\n";
    for (std::size_t i = 0; i < 3; ++i) {
        auto const &sp = out.steps[std::min(i, out.steps.size() - 1)];
        os << "struct S" << i << " {\n    using C = stub_" << (sp.tile ? sp.tp.identifier : sp.kp.identifier)
           << "::C;\n    static constexpr int KIND = " << (sp.tile ? 1 : 0) << ";\n    static constexpr bbk::u64 PER_K = "
           << sp.per_k << "ull;\n};\n";
    }
    os << "} // namespace chain_" << out.identifier << "\n";
    os << "extern \"C\" BBK_GLOBAL void BBK_KERNEL(" << out.threads << ", " << out.max_regs << ") " << out.identifier
       << "(bbk::chain_args ca) {\n    bbk::chain<" << out.steps.size() << ", chain_" << out.identifier << "::S0, chain_"
       << out.identifier << "::S1, chain_" << out.identifier << "::S2>(ca);\n}\n";
    out.entry_source = os.str();
    out.source = stubs.str() + out.entry_source;
    return true;
}

} // namespace bbfft::cuda
