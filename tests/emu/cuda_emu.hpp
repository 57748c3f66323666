// tests/emu/cuda_emu.hpp -- TEST INFRASTRUCTURE ONLY (never part of the product path).
//
// Minimal CUDA execution-model shim so that the device header
// double-batched-fft-library_b200/csrc/kernels/bbfft_kernels.cuh and a generated stub can be
// compiled with g++ (-DBBFFT_EMU) and run on a CPU-only box: one fiber per CUDA thread,
// __syncthreads() yields to the CTA scheduler in emu_runner.cpp.  Used by the "not gpu" tests to
// check every index map of every planned kernel against the oracle before a GPU is involved.
#ifndef BBFFT_CUDA_EMU_HPP
#define BBFFT_CUDA_EMU_HPP
#include <stddef.h>
namespace bbfft_emu {
// An asynchronous shared-memory copy (cp.async / cp.async.bulk) in flight: the destination is undefined from the
// issue to the wait that completes it.
struct pending_copy {
    unsigned char *dst;
    unsigned bytes;
    unsigned char data[16];
};
struct pending_list; // (std::vector in emu_runner.cpp)
struct thread_ctx {
    int tid;
    unsigned long long bid;
    unsigned char *smem;
    void (*yield)(thread_ctx *);
    pending_list *pend; // this thread's cp.async copies between issue and cp.async.wait_group
};
extern unsigned long long grid_size; // CTAs of the emulated launch (persistent kernels stride by it)
extern int failed;                   // set by device code that would hang or trap on the GPU
extern thread_local thread_ctx *current;
#ifdef BBFFT_EMU_RACECHECK
extern thread_local int epoch;
#endif
inline void syncthreads() {
    thread_ctx *c = current;
    c->yield(c);
    current = c;
}
#ifdef BBFFT_EMU_RACECHECK
// Shared-memory race checker: every 4-byte word of the CTA's shared memory remembers who wrote and
// who read it in the current barrier interval ("epoch" = number of barriers the accessing thread has
// passed).  Two threads touching the same word in the same epoch, at least one of them writing, is a
// race on the GPU even though the emulator runs the threads one after the other.
struct word_state {
    int w_tid = -1, w_epoch = -1, r_tid = -1, r_epoch = -1; // r_tid == -2: several readers
};
extern word_state *shadow;
extern unsigned char *shadow_base;
extern long races;
extern int report_unwritten; // BBFFT_EMU_UNWRITTEN=1: also report reads of never-written words
extern thread_local int epoch;
void report_race(const char *kind, long word, int other_tid);
inline void note_access(const void *p, unsigned long bytes, bool write) {
    const long first = (static_cast<const unsigned char *>(p) - shadow_base) / 4;
    for (long w = first; w < first + long(bytes / 4); ++w) {
        word_state &s = shadow[w];
        const int t = current->tid;
        if (write) {
            if (s.w_epoch == epoch && s.w_tid != t) report_race("write-after-write", w, s.w_tid);
            if (s.r_epoch == epoch && s.r_tid != t) report_race("write-after-read", w, s.r_tid);
            s.w_tid = t;
            s.w_epoch = epoch;
        } else {
            if (s.w_epoch == epoch && s.w_tid != t) report_race("read-after-write", w, s.w_tid);
            if (s.w_epoch < 0 && report_unwritten) report_race("read of a word no thread of the CTA has written", w, -1);
            if (s.r_epoch != epoch) {
                s.r_tid = t;
                s.r_epoch = epoch;
            } else if (s.r_tid != t) {
                s.r_tid = -2;
            }
        }
    }
}
template <class E> struct checked_ref {
    E *p;
    operator E() const {
        note_access(p, sizeof(E), false);
        return *p;
    }
    checked_ref &operator=(E const &v) {
        note_access(p, sizeof(E), true);
        *p = v;
        return *this;
    }
    checked_ref &operator=(checked_ref const &o) { return *this = E(o); }
};
template <class E> struct checked_ptr {
    E *p;
    checked_ref<E> operator[](long i) const { return checked_ref<E>{p + i}; }
};
#endif
// Asynchronous copies, modelled pessimistically: at the issue the destination is POISONED (all-ones bytes: NaN for
// both precisions) and, under the race checker, counts as written by the issuing thread -- so a thread that still
// reads or writes the old contents in that barrier interval is a reported race; the data lands only at the wait
// (cp.async: the issuing thread's cp.async.wait_group, after which a barrier must publish it -- the landing is a
// write of that thread in the race checker; cp.async.bulk: the mbarrier wait every thread performs itself).  A
// kernel that touches the destination between issue and wait computes NaNs and fails its parity test.
void async_issue_raw(void *dst, const void *src, unsigned bytes, bool bulk);
void async_wait_raw(bool bulk);
template <class E> inline E *raw_ptr(E *p) { return p; }
#ifdef BBFFT_EMU_RACECHECK
template <class E> inline E *raw_ptr(checked_ptr<E> p) { return p.p; }
#endif
inline int thread_idx() { return current->tid; }
inline unsigned long long block_idx() { return current->bid; }
inline unsigned char *shared_mem() { return current->smem; }
inline unsigned long long grid_dim() { return grid_size; }
inline void fail() { failed = 1; }
} // namespace bbfft_emu
#endif
