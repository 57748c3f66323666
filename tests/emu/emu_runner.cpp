// tests/emu/emu_runner.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiled together with one generated stub (-DBBFFT_EMU -DBBFFT_EMU_KERNEL=<identifier>
// -include cuda_emu.hpp) into a shared object; emu_launch() runs the kernel grid on the host.
#include "cuda_emu.hpp"
#include "bbfft_kernels.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ucontext.h>
#include <algorithm>
#include <vector>

// chain kernels (-DBBFFT_EMU_CHAIN) take bbk::chain_args, everything else bbk::args
#ifdef BBFFT_EMU_CHAIN
typedef bbk::chain_args emu_args_t;
#else
typedef bbk::args emu_args_t;
#endif
extern "C" void BBFFT_EMU_KERNEL(emu_args_t a);

namespace bbfft_emu {
struct pending_list {
    std::vector<pending_copy> v;
};
static pending_list bulk_pending; // cp.async.bulk copies of the running CTA (one mbarrier per kernel)
static void push_copy(pending_list &l, void *dst, const void *src, unsigned bytes) {
    unsigned char *d = static_cast<unsigned char *>(dst);
    const unsigned char *s = static_cast<const unsigned char *>(src);
    for (unsigned off = 0; off < bytes; off += 16) {
        pending_copy c;
        c.dst = d + off;
        c.bytes = bytes - off < 16 ? bytes - off : 16;
        std::memcpy(c.data, s + off, c.bytes);
        l.v.push_back(c);
    }
    std::memset(dst, 0xff, bytes);
}
void async_issue_raw(void *dst, const void *src, unsigned bytes, bool bulk) {
#ifdef BBFFT_EMU_RACECHECK
    note_access(dst, bytes, true);
#endif
    push_copy(bulk ? bulk_pending : *current->pend, dst, src, bytes);
}
void async_wait_raw(bool bulk) {
    pending_list &l = bulk ? bulk_pending : *current->pend;
    for (auto const &c : l.v) {
#ifdef BBFFT_EMU_RACECHECK
        if (!bulk) note_access(c.dst, c.bytes, true);
#endif
        std::memcpy(c.dst, c.data, c.bytes);
    }
    l.v.clear();
}
thread_local thread_ctx *current = nullptr;
unsigned long long grid_size = 1;
int failed = 0;
#ifdef BBFFT_EMU_RACECHECK
word_state *shadow = nullptr;
unsigned char *shadow_base = nullptr;
long races = 0;
int report_unwritten = 0;
thread_local int epoch = 0;
void report_race(const char *kind, long word, int other_tid) {
    if (races < 8) {
        std::fprintf(stderr, "shared-memory race (%s): word %ld, threads %d and %d, barrier interval %d, CTA %llu\n", kind, word,
                     other_tid, current->tid, epoch, current->bid);
    }
    ++races;
}
#endif
}

namespace {
constexpr std::size_t stack_bytes = 256 * 1024;
struct fiber {
    ucontext_t uc;
    bbfft_emu::thread_ctx ctx;
    bool done = false;
    char *stack = nullptr;
    int epoch = 0; // barriers passed (race checker)
};
thread_local ucontext_t sched_uc;
thread_local fiber *running = nullptr;
thread_local emu_args_t *launch_args = nullptr;

void yield_to_sched(bbfft_emu::thread_ctx *) {
    fiber *f = running;
    ++f->epoch;
    swapcontext(&f->uc, &sched_uc);
}
void fiber_main() {
    fiber *f = running;
    bbfft_emu::current = &f->ctx;
    BBFFT_EMU_KERNEL(*launch_args);
    f->done = true;
    swapcontext(&f->uc, &sched_uc);
}
} // namespace

extern "C" int emu_launch(emu_args_t *a, unsigned long long grid, int threads, unsigned long smem_bytes) {
    launch_args = a;
    bbfft_emu::grid_size = grid;
    bbfft_emu::failed = 0;
    std::vector<fiber> fibers(threads);
    std::vector<bbfft_emu::pending_list> pending(threads);
    for (auto &f : fibers) {
        f.stack = static_cast<char *>(std::malloc(stack_bytes));
    }
    std::vector<unsigned char> smem(smem_bytes + 64);
#ifdef BBFFT_EMU_RACECHECK
    std::vector<bbfft_emu::word_state> shadow_words(smem.size() / 4 + 1);
    bbfft_emu::shadow = shadow_words.data();
    bbfft_emu::shadow_base = smem.data();
    bbfft_emu::races = 0;
    {
        char const *u = std::getenv("BBFFT_EMU_UNWRITTEN");
        bbfft_emu::report_unwritten = (u && *u == '1') ? 1 : 0;
    }
#endif
    // BBFFT_EMU_SMEM_FILL=<byte>: what "uninitialised" shared memory holds; BBFFT_EMU_ORDER=reverse: run
    // the threads of a barrier interval last-to-first.  A correct kernel's output depends on neither.
    int fill = 0xcd;
    if (char const *f = std::getenv("BBFFT_EMU_SMEM_FILL")) fill = std::atoi(f) & 0xff;
    char const *ord = std::getenv("BBFFT_EMU_ORDER");
    const bool reverse = ord && ord[0] == 'r';
    for (unsigned long long bb = 0; bb < grid; ++bb) {
        const unsigned long long bid = reverse ? grid - 1 - bb : bb; // CTAs run in any order, too
        std::memset(smem.data(), fill, smem.size());
        bbfft_emu::bulk_pending.v.clear();
#ifdef BBFFT_EMU_RACECHECK
        std::fill(shadow_words.begin(), shadow_words.end(), bbfft_emu::word_state{});
#endif
        for (int t = 0; t < threads; ++t) {
            fiber &f = fibers[t];
            f.done = false;
            f.epoch = 0;
            pending[t].v.clear();
            f.ctx = {t, bid, smem.data(), yield_to_sched, &pending[t]};
            getcontext(&f.uc);
            f.uc.uc_stack.ss_sp = f.stack;
            f.uc.uc_stack.ss_size = stack_bytes;
            f.uc.uc_link = nullptr;
            makecontext(&f.uc, fiber_main, 0);
        }
        for (;;) {
            int done = 0;
            for (int tt = 0; tt < threads; ++tt) {
                const int t = reverse ? threads - 1 - tt : tt;
                fiber &f = fibers[t];
                if (!f.done) {
                    running = &f;
                    bbfft_emu::current = &f.ctx;
#ifdef BBFFT_EMU_RACECHECK
                    bbfft_emu::epoch = f.epoch;
#endif
                    swapcontext(&sched_uc, &f.uc);
                }
                if (f.done) ++done;
            }
            if (done == threads) break;
            if (done != 0) {
                // some threads returned while others wait at a barrier: CUDA would hang
                for (auto &f : fibers) std::free(f.stack);
                return 2;
            }
        }
    }
    for (auto &f : fibers) std::free(f.stack);
#ifdef BBFFT_EMU_RACECHECK
    if (bbfft_emu::races) return 4;
#endif
    return bbfft_emu::failed ? 3 : 0;
}
