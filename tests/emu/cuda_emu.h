// TEST INFRASTRUCTURE ONLY — never part of the product path.
//
// A single-OS-thread emulator of the CUDA execution model, just large enough to
// run the kernels under neurips-cellseg_b200/csrc/ on the GPU-less CI box so
// that `pytest -m "not gpu"` can check the kernels' *logic* (bit tricks, union
// sets, border merges, raster ranks) against the oracle.  Every CUDA thread of a
// block is a fiber; __syncthreads() and the warp collectives are fiber
// barriers; blocks run one after another; atomics are plain read-modify-write.
// It cannot find data races and its timing means nothing.  The product library
// (libcellseg_sm100.so) is always the nvcc build; nothing in the package
// imports or links this file.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static
#define __grid_constant__

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct int2 { int x, y; };
struct int4 { int x, y, z, w; } __attribute__((aligned(16)));
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; } __attribute__((aligned(16)));
struct float2 { float x, y; };
struct float4 { float x, y, z, w; } __attribute__((aligned(16)));
struct uchar4 { unsigned char x, y, z, w; };
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return {x, y}; }

namespace cs_emu {
extern dim3 g_gridDim, g_blockDim;
extern uint3 g_blockIdx, g_threadIdx;
extern unsigned char* g_dyn_smem;
void block_barrier();
void warp_barrier();
unsigned lane_id();
unsigned* warp_slots();          // 32 x 64-bit scratch of the calling fiber's warp
unsigned long long* warp_slots64();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}  // namespace cs_emu

#define gridDim cs_emu::g_gridDim
#define blockDim cs_emu::g_blockDim
#define blockIdx cs_emu::g_blockIdx
#define threadIdx cs_emu::g_threadIdx
#define warpSize 32

static inline void __syncthreads() { cs_emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cs_emu::warp_barrier(); }
namespace cs_emu { extern int g_sync_or; }
static inline int __syncthreads_or(int pred) {
    if (pred) cs_emu::g_sync_or = 1;
    cs_emu::block_barrier();
    int r = cs_emu::g_sync_or;
    cs_emu::block_barrier();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) cs_emu::g_sync_or = 0;
    cs_emu::block_barrier();
    return r;
}
namespace cs_emu { extern int g_sync_count; }
static inline int __syncthreads_count(int pred) {
    if (pred) cs_emu::g_sync_count += 1;
    cs_emu::block_barrier();
    int r = cs_emu::g_sync_count;
    cs_emu::block_barrier();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) cs_emu::g_sync_count = 0;
    cs_emu::block_barrier();
    return r;
}
static inline void __threadfence() {}
static inline void __threadfence_block() {}

// ---- bit / math intrinsics -------------------------------------------------
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    return __builtin_bswap32(v);
}
static inline unsigned long long __brevll(unsigned long long v) {
    return ((unsigned long long)__brev((unsigned)v) << 32) | (unsigned long long)__brev((unsigned)(v >> 32));
}
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline float __double2float_rn(double a) { return (float)a; }
static inline double __uint2double_rn(unsigned a) { return (double)a; }
static inline double __ull2double_rn(unsigned long long a) { return (double)a; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
using std::min;
using std::max;

// ---- atomics (single OS thread: plain RMW) ----------------------------------
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

// ---- warp collectives (full-mask use only) ----------------------------------
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned* s = cs_emu::warp_slots();
    s[cs_emu::lane_id()] = pred ? 1u : 0u;
    cs_emu::warp_barrier();
    unsigned r = 0;
    unsigned nl = std::min(32u, blockDim.x * blockDim.y * blockDim.z);
    for (unsigned i = 0; i < nl; ++i) r |= (s[i] & 1u) << i;
    cs_emu::warp_barrier();
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
template <class T> static inline T cs_emu_shfl(T v, int src) {
    static_assert(sizeof(T) <= 8, "shfl payload");
    unsigned long long* s = cs_emu::warp_slots64();
    unsigned long long raw = 0; memcpy(&raw, &v, sizeof(T));
    s[cs_emu::lane_id()] = raw;
    cs_emu::warp_barrier();
    unsigned long long got = s[src & 31];
    cs_emu::warp_barrier();
    T out; memcpy(&out, &got, sizeof(T));
    return out;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return cs_emu_shfl(v, src); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    int l = (int)cs_emu::lane_id(); return cs_emu_shfl(v, l - (int)d < 0 ? l : l - (int)d);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    int l = (int)cs_emu::lane_id(); return cs_emu_shfl(v, l + (int)d > 31 ? l : l + (int)d);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
    return cs_emu_shfl(v, (int)cs_emu::lane_id() ^ m);
}

// ---- the slice of the runtime API the host code uses ------------------------
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
