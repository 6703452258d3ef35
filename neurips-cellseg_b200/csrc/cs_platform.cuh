// Platform glue shared by every kernel file of libcellseg_sm100.so.
//
// The product build is nvcc for sm_100a.  Defining CS_EMU (done only by
// tests/emu/build_emu.py) swaps the CUDA headers for the test-only fiber
// emulator in tests/emu/cuda_emu.h so the GPU-less CI can exercise the kernel
// logic; the package itself never loads that build.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cstring>

#ifdef CS_EMU
#include "cuda_emu.h"
#define CS_LAUNCH(kern, grid, block, smem, stream, ...) \
    cs_emu::launch((grid), (block), (smem), [&]() { kern(__VA_ARGS__); })
#define CS_DYN_SMEM(name) unsigned char* name = cs_emu::g_dyn_smem
#else
#include <cuda_runtime.h>
#define CS_LAUNCH(kern, grid, block, smem, stream, ...)                                        \
    do {                                                                                       \
        cs::count_launch();                                                                    \
        cs::prof::Scope cs_prof_scope_(#kern, (cudaStream_t)(stream));                         \
        kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);                \
    } while (0)
#define CS_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace cs {

// Thread-local error text returned by cs_last_error(); entry points never throw.
void set_error(const char* fmt, ...);
const char* last_error();
// Number of kernels this library has launched in the process (cs_kernel_launch_count()).
void count_launch();
unsigned long long launch_count();

#ifndef CS_EMU
// Opt-in per-kernel device timing (cs_profile_enable / cs_profile_read): when enabled every launch is
// bracketed by a cudaEvent pair on its own stream; nothing is recorded, and nothing synchronises, otherwise.
namespace prof {
void enable(bool on);
bool enabled();
void begin(const char* name, cudaStream_t st, int* token);
void end(cudaStream_t st, int token);
int read(int index, char* name, size_t cap, double* total_ms, unsigned long long* count);
struct Scope {
    cudaStream_t st; int token;
    Scope(const char* name, cudaStream_t s) : st(s), token(-1) { if (enabled()) begin(name, s, &token); }
    ~Scope() { if (token >= 0) end(st, token); }
};
}  // namespace prof
#endif

enum : int {
    CS_OK = 0,
    CS_ERR_ARG = -1,       // bad argument (null pointer, negative size, unsupported connectivity …)
    CS_ERR_WORKSPACE = -2, // caller-provided workspace too small
    CS_ERR_CUDA = -3,      // a CUDA runtime call or launch failed; text in cs_last_error()
    CS_ERR_RANGE = -4,     // image too large for 32-bit pixel indices / label table
    CS_ERR_OVERFLOW = -5,  // a device-side table overflowed (retry with a larger capacity)
};

#define CS_CHECK_ARG(cond, msg)                                   \
    do { if (!(cond)) { cs::set_error("%s: %s", __func__, msg); return cs::CS_ERR_ARG; } } while (0)

#define CS_CUDA(call)                                                                  \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) {                              \
        cs::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e_));     \
        return cs::CS_ERR_CUDA; } } while (0)

#define CS_LAUNCH_CHECK()                                                              \
    do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) {                  \
        cs::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e_)); \
        return cs::CS_ERR_CUDA; } } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Bump allocator over the caller-provided workspace (256-byte granules).
struct Arena {
    unsigned char* base; size_t cap; size_t off;
    Arena(void* p, size_t n) : base((unsigned char*)p), cap(n), off(0) {}
    template <class T> T* take(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        T* r = (T*)(base ? base + off : nullptr);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= cap; }
};

}  // namespace cs
