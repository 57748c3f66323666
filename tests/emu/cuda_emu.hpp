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
struct thread_ctx {
    int tid;
    unsigned long long bid;
    unsigned char *smem;
    void (*yield)(thread_ctx *);
};
extern unsigned long long grid_size; // CTAs of the emulated launch (persistent kernels stride by it)
extern int failed;                   // set by device code that would hang or trap on the GPU
extern thread_local thread_ctx *current;
inline void syncthreads() {
    thread_ctx *c = current;
    c->yield(c);
    current = c;
}
inline int thread_idx() { return current->tid; }
inline unsigned long long block_idx() { return current->bid; }
inline unsigned char *shared_mem() { return current->smem; }
inline unsigned long long grid_dim() { return grid_size; }
inline void fail() { failed = 1; }
} // namespace bbfft_emu
#endif
