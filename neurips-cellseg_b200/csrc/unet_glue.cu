// Predictor glue, channels-last edition (the predictor is monai's UNet, baseline/predict.py:45-52).
//
// The convolutions with enough channels to be a dense contraction stay on cuDNN (out of scope by the north
// star) and run fastest on NHWC tensors (no nchw<->nhwc conversion kernels around the tensor-core convolutions).
// What cuDNN is bad at on this network are the 3-channel ends, which are HBM-bound stencils rather than
// contractions (0.26 / 0.58 ms per 100 windows against ~0.03 ms of traffic): they are written here, together with
// the norm blocks between the convolutions, so that every activation tensor is produced channels-last:
//
//   k_stem<FROM_U8>     the two stride-2 3->16 convolutions of the first ResidualUnit (unit0 and the residual branch) in one pass, from the
//                       NCHW window batch or — FROM_U8 — straight from the uint8 HWC images through the 256-entry scale table (the
//                       sliding-window gather and the /max scaling fused in): both NHWC outputs + the InstanceNorm partials of unit0
//   k_nhwc_stats        per-(sample, channel) partial statistics (chunk mean, chunk M2) of an NHWC tensor cuDNN produced
//   k_nhwc_norm_prelu   InstanceNorm2d(affine=False) -> PReLU (-> + residual + residual-conv bias) from those partials; optional
//                       second, channel-strided output into a skip connection's concatenation buffer (torch.cat never runs)
//   k_nhwc_stats_view / k_nhwc_norm_prelu_view<SHUFFLE>
//                       the same pair on channel slices of wider tensors (one merged cuDNN call for unit0 + residual branch of a
//                       down unit) and behind a transposed convolution computed as a 2x2 convolution (depth-to-space on output)
//   k_nhwc_norm_prelu_cluster   single-pass variant on thread-block clusters + distributed shared memory (workspace-free; slower)
//   k_head_convt        ConvTranspose2d(32 -> 3, k3, s2, p1, op1) reading the two 16-channel NHWC halves of the skip
//                       concatenation (no torch.cat), cp.async-staged tiles, planar output + partial statistics
//   k_head_final        InstanceNorm -> PReLU -> Conv2d(3 -> 3, k3) + bias + identity residual, planar logits out
// Convolution weights of the 3-channel kernels travel as kernel parameters (constant-bank / uniform-register operands) and the
// stem / transposed head issue packed FFMA2 (fma.rn.f32x2).
//
// Statistics are two-level and deterministic: each chunk yields (mean, M2 about its own mean); consumers merge the
// chunks of a sample with Chan's formula in double.  No atomics, no dependence on scheduling.
// The library is compiled with --fmad=false (the blend must round products and sums separately); the convolution
// loops below ask for fused multiply-adds explicitly.
#include "internal.cuh"
#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <mutex>

namespace cs {

constexpr int GL_T = 256;      // threads per CTA of every kernel in this file
constexpr int GL_VPT = 8;      // float4 per thread of an NHWC chunk → 2048 float4 (32 KB) per CTA

// ---------------------------------------------------------------------------
// per-channel CTA reduction for NHWC data
// ---------------------------------------------------------------------------
// Every thread holds a float4 = 4 consecutive channels of channel group (threadIdx.x % G), G = C/4 ∈ {1,2,4,...,64}.
// Afterwards sc.out[c] is the sum over the CTA for channel c (valid for all threads).
template <int NT>
struct ChanScratchT {
    float4 part[NT / 32][32];
    float out[256];
};
using ChanScratch = ChanScratchT<GL_T>;

template <int NT>
__device__ __forceinline__ void cta_channel_sum(float4 s, int G, ChanScratchT<NT>& sc) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int d = G; d < 32; d <<= 1) {       // lanes with equal lane % G carry the same channel group
        s.x += __shfl_xor_sync(0xffffffffu, s.x, d);
        s.y += __shfl_xor_sync(0xffffffffu, s.y, d);
        s.z += __shfl_xor_sync(0xffffffffu, s.z, d);
        s.w += __shfl_xor_sync(0xffffffffu, s.w, d);
    }
    __syncthreads();                         // previous readers of sc are done
    sc.part[wid][lane] = s;
    __syncthreads();
    const int C = 4 * G;
    for (int c = threadIdx.x; c < C; c += NT) {
        const int g = c >> 2, k = c & 3;
        float acc = 0.f;
        for (int w = 0; w < NT / 32; ++w) {              // fixed order → deterministic
            if (G > 32 && ((w * 32) % G) != (g & ~31)) continue;
            const float4 p = sc.part[w][g & 31];
            acc += k == 0 ? p.x : k == 1 ? p.y : k == 2 ? p.z : p.w;
        }
        sc.out[c] = acc;
    }
    __syncthreads();
}

__device__ __forceinline__ float prelu_f(float v, float a) { return v > 0.f ? v : a * v; }

// Traversal order of the statistics / apply pair (CELLSEG_L2_ORDER, a bit mask).  CTAs are dispatched in ascending block index, so a kernel
// that walks the samples in ascending order leaves the LAST ~100 MB it touched in the 126 MB L2 — exactly the part the next kernel, walking
// in the same direction, reaches last.  Reversing one of two consecutive kernels turns the tail of the first into the head of the second.
constexpr int ORD_STATS_REV = 1, ORD_NORM_REV = 2, ORD_STATS_KEEP = 4;     // KEEP: statistics loads are not marked evict-first
__device__ __forceinline__ int ord_x(int rev) { return rev ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x; }
__device__ __forceinline__ int ord_y(int rev) { return rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y; }

// Two fused multiply-adds per instruction (Blackwell FFMA2): acc.{x,y} += v * w.{x,y}.  The scalar is broadcast by the instruction and the
// weight pair may sit in a uniform register pair, so a convolution inner loop issues half as many FMA instructions.
#ifdef CS_EMU
__device__ __forceinline__ void fma2(float2& acc, float v, float wx, float wy) { acc.x = fmaf(v, wx, acc.x); acc.y = fmaf(v, wy, acc.y); }
#else
__device__ __forceinline__ void fma2(float2& acc, float v, float wx, float wy) {
    float2 w = make_float2(wx, wy), vv = make_float2(v, v);
    unsigned long long a = *reinterpret_cast<unsigned long long*>(&acc), ww = *reinterpret_cast<unsigned long long*>(&w);
    unsigned long long vq = *reinterpret_cast<unsigned long long*>(&vv), r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(vq), "l"(ww), "l"(a));
    acc = *reinterpret_cast<float2*>(&r);
}
#endif

// 16-byte asynchronous global → shared copy (LDGSTS, no register round trip); !ok zero-fills the destination.
#ifdef CS_EMU
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool ok) {
    *reinterpret_cast<float4*>(smem) = ok ? *reinterpret_cast<const float4*>(gmem) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void cp_async_wait_all() {}
#else
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool ok) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
#endif

// How many values chunk k of a (sample, channel) summarises.  LINEAR: consecutive chunks of `per` values over `total`
// (the last one shorter).  TILES: 2-D tiles of th x tw input pixels over an Hin x Win image, `mult` values per pixel.
struct ChunkGeom {
    int tiled, per, total, tiles_x, th, tw, Hin, Win, mult;     // LINEAR: mult = interleaved sub-lists per chunk (k → chunk k / mult)
    __host__ __device__ int count(int k) const {
        if (!tiled) { const int left = total / mult - (k / mult) * per; return left < per ? left : per; }
        const int ty = k / tiles_x, tx = k % tiles_x;
        const int rows = Hin - ty * th < th ? Hin - ty * th : th, cols = Win - tx * tw < tw ? Win - tx * tw : tw;
        return rows * cols * mult;
    }
};
__host__ __device__ static inline ChunkGeom linear_chunks(int per, int total, int mult = 1) {
    ChunkGeom g = {0, per, total * mult, 1, 1, 1, 1, 1, mult};
    return g;
}

// Merge the per-chunk (mean, M2) of every channel of one sample with all threads of the CTA: p[k*C + c], k < nstat.
// Chan's formula collapsed to two sums in double: mean = Σ n_k m_k / n,  var = Σ (M2_k + n_k m_k²) / n − mean².
// Leaves mean and 1/sqrt(var + eps) in s_mean / s_rstd (visible to the whole CTA on return).
struct MergeScratch { double s1[GL_T], s2[GL_T]; };
__device__ __forceinline__ void merge_partials_cta(const float2* __restrict__ p, int C, int nstat, const ChunkGeom& cg, float eps,
                                                   float* s_mean, float* s_rstd, MergeScratch& ms) {
    const int t = threadIdx.x, KG = GL_T / C;
    double s1 = 0.0, s2 = 0.0;
    if (t < KG * C) {
        const int c = t % C;
        for (int k = t / C; k < nstat; k += KG) {
            const double cnt = (double)cg.count(k);
            const float2 q = __ldcg(p + (size_t)k * C + c);      // L2: in the fused pair other CTAs of this launch wrote it
            s1 += cnt * (double)q.x;
            s2 += (double)q.y + cnt * (double)q.x * (double)q.x;
        }
    }
    ms.s1[t] = s1; ms.s2[t] = s2;
    __syncthreads();
    double a = 0.0, b = 0.0;
    bool writer = false;
    int c = t;
    if (C <= GL_T / 32) {
        // few channels, many groups (the 3-channel head: 85 groups): one warp per channel, lanes stride over the groups, then a
        // shuffle tree — a single thread walking 85 dependent double additions held every CTA of k_head_final for ~1 us
        const int lane = t & 31;
        c = t >> 5;
        if (c < C)
            for (int g = lane; g < KG; g += 32) { a += ms.s1[g * C + c]; b += ms.s2[g * C + c]; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); }
        writer = c < C && lane == 0;
    } else if (t < C) {
        for (int g = 0; g < KG; ++g) { a += ms.s1[g * C + t]; b += ms.s2[g * C + t]; }
        writer = true;
    }
    if (writer) {
        const double total = (double)cg.total, mean = a / total;
        double var = b / total - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[c] = (float)mean;
        s_rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// NHWC statistics + norm/activation/residual
// ---------------------------------------------------------------------------
// Two channel-group sums at once (same barriers as one): afterwards sc.out[c] / sc2[c] hold the CTA sums of a / b for channel c.
template <int NT>
__device__ __forceinline__ void cta_channel_sum2(float4 a, float4 b, int G, ChanScratchT<NT>& sc, float4 (*part2)[32], float* out2) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int d = G; d < 32; d <<= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, d); a.y += __shfl_xor_sync(0xffffffffu, a.y, d);
        a.z += __shfl_xor_sync(0xffffffffu, a.z, d); a.w += __shfl_xor_sync(0xffffffffu, a.w, d);
        b.x += __shfl_xor_sync(0xffffffffu, b.x, d); b.y += __shfl_xor_sync(0xffffffffu, b.y, d);
        b.z += __shfl_xor_sync(0xffffffffu, b.z, d); b.w += __shfl_xor_sync(0xffffffffu, b.w, d);
    }
    sc.part[wid][lane] = a;
    part2[wid][lane] = b;
    __syncthreads();
    const int C = 4 * G;
    for (int c = threadIdx.x; c < C; c += NT) {
        const int g = c >> 2, k = c & 3;
        float s1 = 0.f, s2 = 0.f;
        for (int w = 0; w < NT / 32; ++w) {              // fixed order → deterministic
            if (G > 32 && ((w * 32) % G) != (g & ~31)) continue;
            const float4 p = sc.part[w][g & 31], q = part2[w][g & 31];
            s1 += k == 0 ? p.x : k == 1 ? p.y : k == 2 ? p.z : p.w;
            s2 += k == 0 ? q.x : k == 1 ? q.y : k == 2 ? q.z : q.w;
        }
        sc.out[c] = s1; out2[c] = s2;
    }
}

// grid (chunks, N).  Chunk = P pixels (P*G <= GL_T*GL_VPT float4); partials[n][chunk][c] = (mean, M2) of the chunk.
// One sweep over the registers: sums of (x - x0) and (x - x0)² with x0 = the channel's value at the chunk's first pixel, so
// the subtraction s2 - s1²/n cancels only against the spread of the chunk, not against its mean (shifted-data variance).
struct StatsScratch { ChanScratch sc; float4 part2[GL_T / 32][32]; float out2[256]; };
__device__ __forceinline__ void nhwc_stats_body(const float* __restrict__ x, int HW, int C, int P, float2* __restrict__ partials, int chunk, int n, int nchunks,
                                                bool keep, StatsScratch& ss) {
    ChanScratch& sc = ss.sc;
    float4 (*part2)[32] = ss.part2;
    float* out2 = ss.out2;
    const int G = C >> 2, t = threadIdx.x;
    const int p0 = chunk * P, np = min(P, HW - p0), f4n = np * G;
    const float4* xp = reinterpret_cast<const float4*>(x) + ((size_t)n * HW + p0) * G;
    const float4 x0 = __ldg(xp + (t % G));
    float4 v[GL_VPT];
#pragma unroll
    for (int i = 0; i < GL_VPT; ++i) {
        const int j = t + i * GL_T;
        v[i] = x0;
        if (j < f4n) v[i] = keep ? __ldg(xp + j) : __ldcs(xp + j);
    }
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < GL_VPT; ++i) {
        const float a = v[i].x - x0.x, b = v[i].y - x0.y, c = v[i].z - x0.z, d = v[i].w - x0.w;     // 0 for the padding entries
        s.x += a; s.y += b; s.z += c; s.w += d;
        q.x += a * a; q.y += b * b; q.z += c * c; q.w += d * d;
    }
    cta_channel_sum2(s, q, G, sc, part2, out2);
    if (t < C) {
        const float inv = 1.0f / (float)np;
        const float s1 = sc.out[t], s2 = out2[t];
        const float base = reinterpret_cast<const float*>(xp)[t];        // x0 of channel t
        partials[((size_t)n * nchunks + chunk) * C + t] = make_float2(base + s1 * inv, fmaxf(s2 - s1 * s1 * inv, 0.f));
    }
}
__global__ void __launch_bounds__(GL_T) k_nhwc_stats(const float* __restrict__ x, int HW, int C, int P, float2* __restrict__ partials, int ord) {
    __shared__ StatsScratch ss;
    nhwc_stats_body(x, HW, C, P, partials, ord_x(ord & ORD_STATS_REV), ord_y(ord & ORD_STATS_REV), gridDim.x, ord & ORD_STATS_KEEP, ss);
}

// y2 (may be null): a second, channel-strided copy of the result — the slot [c2_off, c2_off + C) of an NHWC tensor with c2_total
// channels, i.e. this tensor's half of a skip-connection concatenation (torch.cat never runs).  y may be null when only the slot is wanted.
struct NormScratch { float s_mean[256], s_rstd[256]; MergeScratch ms; };
struct NormArgs {      // what the apply kernels share
    const float* res; float* y; const float* prelu_w; const float* res_bias; const float2* partials; int nstat, pstat; float eps; float* y2; int c2_total, c2_off;
};
// x and y may be the same tensor (in place): every element is read and written by the same thread, so the __restrict__ promise (which lets
// the eight loads of a thread go out before its first store) is kept in the sense that matters.  STREAM: x and the residual are read for
// the last time (evict-first).
template <bool STREAM>
__device__ __forceinline__ void nhwc_norm_body(const float* __restrict__ x, const NormArgs& a_, int HW, int C, int P, int chunk, int n, NormScratch& ns) {
    float* s_mean = ns.s_mean; float* s_rstd = ns.s_rstd; MergeScratch& ms = ns.ms;
    const float* __restrict__ res = a_.res; float* __restrict__ y = a_.y; const float* prelu_w = a_.prelu_w; const float* res_bias = a_.res_bias;
    const float2* partials = a_.partials; const int nstat = a_.nstat, pstat = a_.pstat; const float eps = a_.eps;
    float* __restrict__ y2 = a_.y2; const int c2_total = a_.c2_total, c2_off = a_.c2_off;
    const int G = C >> 2, t = threadIdx.x;
    merge_partials_cta(partials + (size_t)n * nstat * C, C, nstat, linear_chunks(pstat, HW), eps, s_mean, s_rstd, ms);
    const int p0 = chunk * P, np = min(P, HW - p0), f4n = np * G;
    const size_t base = ((size_t)n * HW + p0) * G;
    const float4* xp = reinterpret_cast<const float4*>(x) + base;
    const float4* rp = res ? reinterpret_cast<const float4*>(res) + base : nullptr;
    float4* yp = y ? reinterpret_cast<float4*>(y) + base : nullptr;
    // slot addressing: float4 j of the chunk is channel group j % G of pixel j / G; G is a power of two
    float4* y2p = y2 ? reinterpret_cast<float4*>(y2) + ((size_t)n * HW + p0) * (c2_total >> 2) + (c2_off >> 2) : nullptr;
    const int lg = 31 - __clz(G), G2 = c2_total >> 2;
    const int g4 = (t % G) * 4;
    const float4 m = make_float4(s_mean[g4], s_mean[g4 + 1], s_mean[g4 + 2], s_mean[g4 + 3]);
    const float4 r = make_float4(s_rstd[g4], s_rstd[g4 + 1], s_rstd[g4 + 2], s_rstd[g4 + 3]);
    float4 rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (res && res_bias) rb = make_float4(res_bias[g4], res_bias[g4 + 1], res_bias[g4 + 2], res_bias[g4 + 3]);
    const float a = prelu_w ? *prelu_w : 1.0f;
#pragma unroll
    for (int i = 0; i < GL_VPT; ++i) {
        const int j = t + i * GL_T;
        if (j < f4n) {
            const float4 v = STREAM ? __ldcs(xp + j) : xp[j];
            float4 o;
            o.x = prelu_f((v.x - m.x) * r.x, a); o.y = prelu_f((v.y - m.y) * r.y, a);
            o.z = prelu_f((v.z - m.z) * r.z, a); o.w = prelu_f((v.w - m.w) * r.w, a);
            if (rp) { const float4 q = STREAM ? __ldcs(rp + j) : rp[j]; o.x += q.x + rb.x; o.y += q.y + rb.y; o.z += q.z + rb.z; o.w += q.w + rb.w; }
            if (yp) yp[j] = o;
            if (y2p) y2p[(size_t)(j >> lg) * G2 + (j & (G - 1))] = o;
        }
    }
}
__global__ void __launch_bounds__(GL_T) k_nhwc_norm_prelu(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ y,
                                                          const float* __restrict__ prelu_w, const float* __restrict__ res_bias,
                                                          int HW, int C, int P, const float2* __restrict__ partials, int nstat, int pstat, float eps,
                                                          float* __restrict__ y2, int c2_total, int c2_off, int ord) {
    __shared__ NormScratch ns;
    const NormArgs a = {res, y, prelu_w, res_bias, partials, nstat, pstat, eps, y2, c2_total, c2_off};
    nhwc_norm_body<false>(x, a, HW, C, P, ord_x(ord & ORD_NORM_REV), ord_y(ord & ORD_NORM_REV), ns);
}

// ---------------------------------------------------------------------------
// the same two kernels on *views*: channel slices of wider NHWC tensors, rows with a gap, depth-to-space output
// ---------------------------------------------------------------------------
// Two cuDNN calls are saved by giving one convolution more output channels and reading slices of its result:
//  * the two stride-2 convolutions of a down ResidualUnit (unit0 and the residual branch read the same input) run as ONE
//    convolution with 2C outputs; unit0's norm reads channels [0, C), the unit's last norm reads the residual from [C, 2C);
//  * ConvTranspose2d(k3, s2, p1, output_padding 1) runs as ONE regular 2x2 convolution with 4C outputs — one group of C per
//    output parity (dy, dx) — on the input with symmetric padding 1, which cuDNN executes 2.4x faster than its transposed
//    (dgrad) kernel at 64 -> 16 channels; the result [N][H+1][W+1][4C] is used from row 1 / column 1 on, and the norm block
//    writes pixel (i, j), group (dy, dx) to output pixel (2i+dy, 2j+dx): the depth-to-space shuffle costs nothing.
struct View {          // float4 index of channel group g of pixel p (row-major over rows of `w` pixels) of sample n
    int w;             // pixels per image row
    int row;           // float4 per stored row
    int pix;           // float4 per stored pixel
    int off;           // float4 offset (first row/column/channel group of the slice)
    long long img;     // float4 per sample
    int wsh;           // log2(w) when w is a power of two, else -1
    // pixel p -> (row i, column j).  One stored "row" (row == 0: channel slices of a dense tensor) needs no split at all and a
    // power-of-two width a shift: the runtime division this replaces was ~20 instructions per 16-byte access in kernels that
    // otherwise issue ~12.
    __device__ __forceinline__ void split(int p, int& i, int& j) const {
        if (row == 0) { i = 0; j = p; }
        else if (wsh >= 0) { i = p >> wsh; j = p & (w - 1); }
        else { i = p / w; j = p - i * w; }
    }
    __device__ __forceinline__ size_t at(int n, int p, int g) const {
        int i, j;
        split(p, i, j);
        return (size_t)n * img + (size_t)i * row + (size_t)j * pix + off + g;
    }
};
static View make_view(int w, int row, int pix, int off, long long img) {
    View v = {w, row, pix, off, img, -1};
    if (w > 0 && (w & (w - 1)) == 0) { int s = 0; while ((1 << s) < w) ++s; v.wsh = s; }
    return v;
}

// k_nhwc_stats on a view with Cv channels per pixel (Cv = C, or 4C for the depth-to-space case): partials[n][chunk][Cv].
__device__ __forceinline__ void nhwc_stats_view_body(const float* __restrict__ x, const View& xv, int HW, int Cv, int P, float2* __restrict__ partials,
                                                     int chunk, int n, int nchunks, bool keep, StatsScratch& ss) {
    ChanScratch& sc = ss.sc;
    float4 (*part2)[32] = ss.part2;
    float* out2 = ss.out2;
    const int G = Cv >> 2, lg = 31 - __clz(G), t = threadIdx.x;
    const int p0 = chunk * P, np = min(P, HW - p0), f4n = np * G;
    const float4* xb = reinterpret_cast<const float4*>(x);
    const int g = t & (G - 1);
    const float4 x0 = __ldg(xb + xv.at(n, p0, g));
    float4 v[GL_VPT];
#pragma unroll
    for (int i = 0; i < GL_VPT; ++i) {
        const int j = t + i * GL_T;
        v[i] = x0;
        if (j < f4n) { const float4* q = xb + xv.at(n, p0 + (j >> lg), g); v[i] = keep ? __ldg(q) : __ldcs(q); }
    }
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < GL_VPT; ++i) {
        const float a = v[i].x - x0.x, b = v[i].y - x0.y, c = v[i].z - x0.z, d = v[i].w - x0.w;
        s.x += a; s.y += b; s.z += c; s.w += d;
        q.x += a * a; q.y += b * b; q.z += c * c; q.w += d * d;
    }
    cta_channel_sum2(s, q, G, sc, part2, out2);
    if (t < Cv) {
        const float inv = 1.0f / (float)np;
        const float s1 = sc.out[t], s2 = out2[t];
        const float base = reinterpret_cast<const float*>(xb + xv.at(n, p0, t >> 2))[t & 3];
        partials[((size_t)n * nchunks + chunk) * Cv + t] = make_float2(base + s1 * inv, fmaxf(s2 - s1 * s1 * inv, 0.f));
    }
}
__global__ void __launch_bounds__(GL_T) k_nhwc_stats_view(const float* __restrict__ x, View xv, int HW, int Cv, int P, float2* __restrict__ partials, int ord) {
    __shared__ StatsScratch ss;
    nhwc_stats_view_body(x, xv, HW, Cv, P, partials, ord_x(ord & ORD_STATS_REV), ord_y(ord & ORD_STATS_REV), gridDim.x, ord & ORD_STATS_KEEP, ss);
}

// SHUFFLE false: x (and res) are views with C channels, y dense [n][HW][C] (+ the concatenation slot y2).
// SHUFFLE true: x is a view with 4C channels per pixel, group (dy, dx) = channel / C; y dense [n][2H][2W][C]; no residual.
template <bool SHUFFLE>
__device__ __forceinline__ void nhwc_norm_view_body(const float* __restrict__ x, const View& xv, const View& rv, const NormArgs& a_, int HW, int C, int P,
                                                    int chunk, int n, NormScratch& ns) {
    float* s_mean = ns.s_mean; float* s_rstd = ns.s_rstd; MergeScratch& ms = ns.ms;
    const float* __restrict__ res = a_.res; float* __restrict__ y = a_.y; const float* prelu_w = a_.prelu_w; const float* res_bias = a_.res_bias;
    const float2* partials = a_.partials; const int nstat = a_.nstat, pstat = a_.pstat; const float eps = a_.eps;
    float* __restrict__ y2 = a_.y2; const int c2_total = a_.c2_total, c2_off = a_.c2_off;
    const int G = C >> 2, Gx = SHUFFLE ? C : G, lgx = 31 - __clz(Gx);
    const int t = threadIdx.x;
    merge_partials_cta(partials + (size_t)n * nstat * (SHUFFLE ? 4 : 1) * C, C, nstat * (SHUFFLE ? 4 : 1), linear_chunks(pstat, HW, SHUFFLE ? 4 : 1), eps,
                       s_mean, s_rstd, ms);
    const int P2 = SHUFFLE ? P / 4 : P;                        // pixels per CTA: the view holds 4x the channels per pixel
    const int p0 = chunk * P2, np = min(P2, HW - p0), f4n = np * Gx;
    const float4* xb = reinterpret_cast<const float4*>(x);
    const float4* rb4 = res ? reinterpret_cast<const float4*>(res) : nullptr;
    const int gx = t & (Gx - 1), g = gx & (G - 1), phase = SHUFFLE ? gx / G : 0;
    const int g4 = g * 4;
    const float4 m = make_float4(s_mean[g4], s_mean[g4 + 1], s_mean[g4 + 2], s_mean[g4 + 3]);
    const float4 r = make_float4(s_rstd[g4], s_rstd[g4 + 1], s_rstd[g4 + 2], s_rstd[g4 + 3]);
    float4 rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (res && res_bias) rb = make_float4(res_bias[g4], res_bias[g4 + 1], res_bias[g4 + 2], res_bias[g4 + 3]);
    const float a = prelu_w ? *prelu_w : 1.0f;
    float4* yb = y ? reinterpret_cast<float4*>(y) : nullptr;
    float4* y2b = y2 ? reinterpret_cast<float4*>(y2) : nullptr;
    const int G2 = c2_total >> 2;
#pragma unroll
    for (int i = 0; i < GL_VPT; ++i) {
        const int j = t + i * GL_T;
        if (j < f4n) {
            const int p = p0 + (j >> lgx);
            const float4 v = __ldcs(xb + xv.at(n, p, gx));
            float4 o;
            o.x = prelu_f((v.x - m.x) * r.x, a); o.y = prelu_f((v.y - m.y) * r.y, a);
            o.z = prelu_f((v.z - m.z) * r.z, a); o.w = prelu_f((v.w - m.w) * r.w, a);
            if (SHUFFLE) {
                int ii, jj;
                xv.split(p, ii, jj);
                const size_t op = (size_t)(2 * ii + (phase >> 1)) * (2 * xv.w) + 2 * jj + (phase & 1);
                yb[((size_t)n * 4 * HW + op) * G + g] = o;
            } else {
                if (rb4) { const float4 q = __ldcs(rb4 + rv.at(n, p, g)); o.x += q.x + rb.x; o.y += q.y + rb.y; o.z += q.z + rb.z; o.w += q.w + rb.w; }
                if (yb) yb[((size_t)n * HW + p) * G + g] = o;
                if (y2b) y2b[((size_t)n * HW + p) * G2 + (c2_off >> 2) + g] = o;
            }
        }
    }
}
template <bool SHUFFLE>
__global__ void __launch_bounds__(GL_T) k_nhwc_norm_prelu_view(const float* __restrict__ x, View xv, const float* __restrict__ res, View rv,
                                                               float* __restrict__ y, const float* __restrict__ prelu_w, const float* __restrict__ res_bias,
                                                               int HW, int C, int P, const float2* __restrict__ partials, int nstat, int pstat, float eps,
                                                               float* __restrict__ y2, int c2_total, int c2_off, int ord) {
    __shared__ NormScratch ns;
    const NormArgs a = {res, y, prelu_w, res_bias, partials, nstat, pstat, eps, y2, c2_total, c2_off};
    nhwc_norm_view_body<SHUFFLE>(x, xv, rv, a, HW, C, P, ord_x(ord & ORD_NORM_REV), ord_y(ord & ORD_NORM_REV), ns);
}

// ---------------------------------------------------------------------------
// statistics + apply in ONE launch: the apply CTAs of a sample run a few samples behind its statistics CTAs
// ---------------------------------------------------------------------------
// As two launches the pair reads x twice from HBM whenever the tensor is larger than L2 (400 windows x 1 MiB at the top level): the
// statistics kernel has swept the whole tensor before the first apply CTA starts.  Here both kinds of CTA live in one grid.  Work is
// handed out by ticket (atomic counter, so "earlier ticket" means "already started"), in the order
//     stats(sample 0) .. stats(sample D-1), [stats(sample s), apply(sample s-D)] for s = D .., apply of the last D samples,
// each bracket being S statistics CTAs followed by S apply CTAs.  An apply CTA waits until the S statistics CTAs of its sample have
// published their partials (a per-sample counter; they hold earlier tickets, never wait themselves, so the wait always ends) and then
// finds x in L2: D samples of x, residual and y have gone by since.  The CTAs are the unchanged bodies of the two kernels.
// sync[0] = ticket counter, sync[1 + n] = finished statistics CTAs of sample n; zeroed by a memset node in front of the launch.
#ifndef CS_EMU
struct PairArgs {
    const float* x; View xv, rv; NormArgs na; float2* partials;
    int HW, C, Cv, P, N, S, D;
    int* sync;
};
constexpr int PAIR_DENSE = 0, PAIR_SLICES = 1, PAIR_SHUFFLE = 2;

template <int KIND>
__global__ void __launch_bounds__(GL_T, 5) k_nhwc_pair(const __grid_constant__ PairArgs a) {
    __shared__ StatsScratch ss;
    __shared__ NormScratch ns;
    __shared__ int s_ticket[2];
    const int t = threadIdx.x;
    const int total = (a.N + a.D) * 2 * a.S;
    if (t == 0) s_ticket[0] = atomicAdd(a.sync, 1);
    __syncthreads();
    // persistent CTAs: the next ticket is drawn before the current item is processed, so its round trip hides behind the item
    for (int it = 0;; ++it) {
        const int cur = s_ticket[it & 1];
        if (cur >= total) break;
        int nxt = 0;
        if (t == 0) nxt = atomicAdd(a.sync, 1);
        const int slot = cur / (2 * a.S), r = cur - slot * 2 * a.S;
        if (r < a.S) {
            const int n = slot, chunk = r;
            if (n < a.N) {
                if (KIND == PAIR_DENSE) nhwc_stats_body(a.x, a.HW, a.C, a.P, a.partials, chunk, n, a.S, true, ss);
                else nhwc_stats_view_body(a.x, a.xv, a.HW, a.Cv, a.P, a.partials, chunk, n, a.S, true, ss);
                __syncthreads();                             // the partials of this CTA are written ...
                if (t == 0) { __threadfence(); atomicAdd(a.sync + 1 + n, 1); }     // ... and visible device-wide before the count says so
            }
        } else {
            const int n = slot - a.D, chunk = r - a.S;
            if (n >= 0) {
                if (t == 0) {
                    const volatile int* done = a.sync + 1 + n;
                    unsigned spins = 0;
                    while (*done < a.S) {
                        __nanosleep(64);
                        if (++spins > (1u << 24)) __trap();  // > 1 s: the counters were not zeroed — fail loudly instead of hanging
                    }
                    __threadfence();
                }
                __syncthreads();
                if (KIND == PAIR_DENSE) nhwc_norm_body<true>(a.x, a.na, a.HW, a.C, a.P, chunk, n, ns);
                else if (KIND == PAIR_SLICES) nhwc_norm_view_body<false>(a.x, a.xv, a.rv, a.na, a.HW, a.C, a.P, chunk, n, ns);
                else nhwc_norm_view_body<true>(a.x, a.xv, a.rv, a.na, a.HW, a.C, 4 * a.P, chunk, n, ns);
            }
        }
        if (t == 0) s_ticket[(it + 1) & 1] = nxt;
        __syncthreads();
    }
}
#endif

// ---------------------------------------------------------------------------
// single-pass variant: one thread-block cluster per sample
// ---------------------------------------------------------------------------
// A sample of up to 1 MiB (the largest activation of the network: 128 x 128 x 16) is held in the registers of a cluster
// of up to 8 CTAs x 512 threads x 16 float4.  Each CTA reduces its slice per channel, the cluster exchanges the per-CTA
// sums through distributed shared memory (two cluster barriers: sum, then squared deviations about the global mean —
// a true two-pass variance), and the result is written straight from registers: x is read from HBM once, where the
// statistics + apply pair reads it twice.
#ifndef CS_EMU
}  // namespace cs
#include <cooperative_groups.h>
namespace cs {
namespace cg = cooperative_groups;

constexpr int CL_T = 512;      // threads per CTA
constexpr int CL_VPT = 16;     // float4 per thread

// grid (cluster size, N), cluster dims (cluster size, 1, 1).  P = pixels per CTA (the last CTA of a sample may hold fewer).
__global__ void __launch_bounds__(CL_T, 1) k_nhwc_norm_prelu_cluster(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ y,
                                                                     const float* __restrict__ prelu_w, const float* __restrict__ res_bias,
                                                                     int HW, int C, int P, float eps) {
    __shared__ ChanScratchT<CL_T> sc;
    __shared__ float cl_sum[256], cl_m2[256], s_mean[256], s_rstd[256];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned csize = cluster.num_blocks();
    const int G = C >> 2, t = threadIdx.x, n = blockIdx.y;
    const int p0 = blockIdx.x * P, np = max(0, min(P, HW - p0)), f4n = np * G;
    const size_t base = ((size_t)n * HW + p0) * G;
    const float4* xp = reinterpret_cast<const float4*>(x) + base;
    float4 v[CL_VPT];
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CL_VPT; ++i) {
        const int j = t + i * CL_T;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < f4n) { v[i] = __ldcs(xp + j); s.x += v[i].x; s.y += v[i].y; s.z += v[i].z; s.w += v[i].w; }
    }
    cta_channel_sum(s, G, sc);
    if (t < C) cl_sum[t] = sc.out[t];
    cluster.sync();
    if (t < C) {
        float tot = 0.f;
        for (unsigned r = 0; r < csize; ++r) tot += *cluster.map_shared_rank(&cl_sum[t], r);     // rank order: same value in every CTA
        s_mean[t] = tot / (float)HW;
    }
    __syncthreads();
    const int g4 = (t % G) * 4;
    const float4 m = make_float4(s_mean[g4], s_mean[g4 + 1], s_mean[g4 + 2], s_mean[g4 + 3]);
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CL_VPT; ++i) {
        const int j = t + i * CL_T;
        if (j < f4n) {
            const float a = v[i].x - m.x, b = v[i].y - m.y, c = v[i].z - m.z, d = v[i].w - m.w;
            q.x += a * a; q.y += b * b; q.z += c * c; q.w += d * d;
        }
    }
    cta_channel_sum(q, G, sc);
    if (t < C) cl_m2[t] = sc.out[t];
    cluster.sync();
    if (t < C) {
        float tot = 0.f;
        for (unsigned r = 0; r < csize; ++r) tot += *cluster.map_shared_rank(&cl_m2[t], r);
        s_rstd[t] = rsqrtf(tot / (float)HW + eps);
    }
    __syncthreads();
    const float4 r = make_float4(s_rstd[g4], s_rstd[g4 + 1], s_rstd[g4 + 2], s_rstd[g4 + 3]);
    float4 rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (res && res_bias) rb = make_float4(res_bias[g4], res_bias[g4 + 1], res_bias[g4 + 2], res_bias[g4 + 3]);
    const float a = prelu_w ? *prelu_w : 1.0f;
    const float4* rp = res ? reinterpret_cast<const float4*>(res) + base : nullptr;
    float4* yp = reinterpret_cast<float4*>(y) + base;
#pragma unroll
    for (int i = 0; i < CL_VPT; ++i) {
        const int j = t + i * CL_T;
        if (j < f4n) {
            float4 o;
            o.x = prelu_f((v[i].x - m.x) * r.x, a); o.y = prelu_f((v[i].y - m.y) * r.y, a);
            o.z = prelu_f((v[i].z - m.z) * r.z, a); o.w = prelu_f((v[i].w - m.w) * r.w, a);
            if (rp) { const float4 w = __ldcs(rp + j); o.x += w.x + rb.x; o.y += w.y + rb.y; o.z += w.z + rb.z; o.w += w.w + rb.w; }
            yp[j] = o;
        }
    }
    cluster.sync();            // nobody leaves while a peer may still read its shared memory
}
#endif

// ---------------------------------------------------------------------------
// stem: Conv2d(3 -> 16, k3, s2, p1) twice (unit0 and the residual branch), NCHW windows in, NHWC out
// ---------------------------------------------------------------------------
// The weights travel as kernel parameters: every thread of a warp needs the same weight at the same time, and a
// parameter lives in the constant bank, so the multiply-add takes it as a direct operand (a shared-memory broadcast
// costs a load-pipe slot per four multiply-adds and was the bottleneck of the first version).
constexpr int STEM_CO = 16;
constexpr int STEM_ROW = 2 * STEM_CO + 4;    // staging row stride in floats (padded: conflict-free float4 stores)
struct StemW { float v[27][2 * STEM_CO]; };  // [ci*9 + ky*3 + kx][co]: unit0's 16 outputs, then the residual branch's

// Where the window pixels come from.  F32: the NCHW window batch (what monai's slicing + torch.cat produce).  U8: straight
// from the uint8 HWC images through the 256-entry scale table — the sliding-window gather and
// the /max scaling (a-1) fused into the first convolution, so the float window batch never exists in HBM.
struct StemSrc {
    const float* win;                        // F32: [N][3][H][W]
    const uint8_t* img; const float* lut;    // U8: [n_img][IH][IW][3], [n_img][256]
    const int* ys; const int* xs;            // window starts in padded coordinates
    int IH, IW, ny, nx, pad_y, pad_x, win_begin;
};

// grid (ceil(OH*OW / 256), N).  A/R: [N][OH*OW][16]; partials (of A): [N][chunks][16].  H, W: window size.
template <bool FROM_U8>
__global__ void __launch_bounds__(GL_T) k_stem(const StemSrc src, int H, int W, const __grid_constant__ StemW wts,
                                               float* __restrict__ A, float* __restrict__ R, float2* __restrict__ partials) {
    __shared__ __align__(16) float stage[GL_T * STEM_ROW];
    __shared__ ChanScratch sc;
    __shared__ float s_lut[256];
    const int t = threadIdx.x, n = blockIdx.y;
    const int OH = H >> 1, OW = W >> 1, OHW = OH * OW;
    const int p0 = blockIdx.x * GL_T, p = p0 + t;
    const int np = min(GL_T, OHW - p0);
    int y_base = 0, x_base = 0;
    const uint8_t* ip = nullptr;
    if (FROM_U8) {
        const int gw = src.win_begin + n, n_win = src.ny * src.nx, ni = gw / n_win, w = gw % n_win;
        y_base = src.ys[w / src.nx] - src.pad_y; x_base = src.xs[w % src.nx] - src.pad_x;
        ip = src.img + (size_t)ni * src.IH * src.IW * 3;
        s_lut[t] = src.lut[(size_t)ni * 256 + t];
        __syncthreads();
    }
    float2 acc[STEM_CO];                             // (co 2q, co 2q+1) pairs: unit0's 16 outputs, then the residual branch's
#pragma unroll
    for (int i = 0; i < STEM_CO; ++i) acc[i] = make_float2(0.f, 0.f);
    if (p < OHW) {
        const int oy = p / OW, ox = p % OW;
        const float* fp = FROM_U8 ? nullptr : src.win + (size_t)n * 3 * H * W;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = 2 * oy - 1 + ky;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = 2 * ox - 1 + kx;
                float v[3] = {0.f, 0.f, 0.f};
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
                    if (FROM_U8) {
                        const int y = y_base + iy, x = x_base + ix;
                        if (y >= 0 && y < src.IH && x >= 0 && x < src.IW) {
                            const uint8_t* q = ip + ((size_t)y * src.IW + x) * 3;
                            v[0] = s_lut[q[0]]; v[1] = s_lut[q[1]]; v[2] = s_lut[q[2]];
                        }
                    } else {
#pragma unroll
                        for (int ci = 0; ci < 3; ++ci) v[ci] = __ldg(fp + ((size_t)ci * H + iy) * W + ix);
                    }
                }
#pragma unroll
                for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                    for (int q = 0; q < STEM_CO; ++q)
                        fma2(acc[q], v[ci], wts.v[ci * 9 + ky * 3 + kx][2 * q], wts.v[ci * 9 + ky * 3 + kx][2 * q + 1]);
            }
        }
    }
    float4* st4 = reinterpret_cast<float4*>(stage + t * STEM_ROW);
#pragma unroll
    for (int q = 0; q < 2 * STEM_CO / 4; ++q) st4[q] = make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
    __syncthreads();
    // coalesced NHWC stores of the CTA's 256-pixel chunk + the chunk statistics of A
    constexpr int G = STEM_CO / 4;
    const int f4n = np * G;
    float4* Ap = reinterpret_cast<float4*>(A) + ((size_t)n * OHW + p0) * G;
    float4* Rp = reinterpret_cast<float4*>(R) + ((size_t)n * OHW + p0) * G;
    float4 v[G];
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < G; ++i) {
        const int j = t + i * GL_T;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < f4n) {
            const int px = j / G, g = j % G;
            v[i] = *reinterpret_cast<const float4*>(stage + px * STEM_ROW + g * 4);
            Ap[j] = v[i];
            Rp[j] = *reinterpret_cast<const float4*>(stage + px * STEM_ROW + STEM_CO + g * 4);
            s.x += v[i].x; s.y += v[i].y; s.z += v[i].z; s.w += v[i].w;
        }
    }
    cta_channel_sum(s, G, sc);
    const float inv = 1.0f / (float)np;
    const int g4 = (t % G) * 4;
    const float4 m = make_float4(sc.out[g4] * inv, sc.out[g4 + 1] * inv, sc.out[g4 + 2] * inv, sc.out[g4 + 3] * inv);
    const float my_mean = t < STEM_CO ? sc.out[t] * inv : 0.f;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < G; ++i) {
        const int j = t + i * GL_T;
        if (j < f4n) {
            const float a = v[i].x - m.x, b = v[i].y - m.y, c = v[i].z - m.z, d = v[i].w - m.w;
            q.x += a * a; q.y += b * b; q.z += c * c; q.w += d * d;
        }
    }
    cta_channel_sum(q, G, sc);
    if (t < STEM_CO) partials[((size_t)n * gridDim.x + blockIdx.x) * STEM_CO + t] = make_float2(my_mean, sc.out[t]);
}

#ifndef CS_EMU
// ---------------------------------------------------------------------------
// stem on the tensor cores (windows 256 wide: the engine's case)
// ---------------------------------------------------------------------------
// k_stem above issues 1376 instructions per output pixel for 864 multiply-adds (432 FFMA2, 222 constant-bank loads, the rest
// addressing, byte loads and the epilogue) and is bound by instruction issue (71 % issue utilisation at 28 % DRAM utilisation,
// profiles/r02_ncu_step_glue.txt).  As a GEMM the two convolutions are [pixels] x [27 taps → 32] x [32 outputs]: here a warp
// takes 32 pixels as two m16n8k8 tiles, `mma.sync` TF32 with the 3xTF32 split (a = a_hi + a_lo, b likewise; a_lo·b_hi + a_hi·b_lo +
// a_hi·b_hi, fp32 accumulation) so the result keeps fp32 accuracy (PyTorch's own fp32 path with TF32 off is the test's
// reference).  One CTA = 256 output pixels = two output rows of a 256-wide window: the 5 x 257 x 3 input patch is staged once in
// shared memory (through the scale table when it comes from the uint8 image), the A fragments are gathered from it, the B
// fragments (weights, split once per CTA) stay in registers.  The epilogue — NHWC stores through the staging buffer and the
// chunk statistics — is k_stem's.
constexpr int SM_PW = 260;                      // patch row stride (257 columns used)
constexpr int SM_WS = 40;                       // weight row stride: (8k + n) mod 32 is a bijection over a fragment load

__device__ __forceinline__ uint32_t tf32_of(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool FROM_U8>
__global__ void __launch_bounds__(GL_T, 4) k_stem_mma(const StemSrc src, int H, const __grid_constant__ StemW wts,
                                                      float* __restrict__ A, float* __restrict__ R, float2* __restrict__ partials) {
    constexpr int W = 256, OW = 128;
    // one buffer, two lives: input patch + split weights while the MMAs run, the pixel-major staging rows afterwards
    __shared__ __align__(16) float buf[GL_T * STEM_ROW];
    __shared__ ChanScratch sc;
    __shared__ float s_lut[256];
    float* patch = buf;                                         // [3][5][SM_PW]
    uint32_t* whi = reinterpret_cast<uint32_t*>(buf + 3 * 5 * SM_PW);      // [32][SM_WS]
    uint32_t* wlo = whi + 32 * SM_WS;
    static_assert(3 * 5 * SM_PW + 2 * 32 * SM_WS <= GL_T * STEM_ROW, "patch + weights must fit the staging buffer");
    const int t = threadIdx.x, n = blockIdx.y, lane = t & 31, warp = t >> 5;
    const int OH = H >> 1, OHW = OH * OW;
    const int p0 = blockIdx.x * GL_T;                           // 256 pixels = output rows oy0, oy0 + 1
    const int np = min(GL_T, OHW - p0);
    const int oy0 = p0 / OW;
    int y_base = 0, x_base = 0;
    const uint8_t* ip = nullptr;
    if (FROM_U8) {
        const int gw = src.win_begin + n, n_win = src.ny * src.nx, ni = gw / n_win, w = gw % n_win;
        y_base = src.ys[w / src.nx] - src.pad_y; x_base = src.xs[w % src.nx] - src.pad_x;
        ip = src.img + (size_t)ni * src.IH * src.IW * 3;
        s_lut[t] = src.lut[(size_t)ni * 256 + t];
    }
    // weights → hi / lo TF32 parts, [k][n] with k = ci*9 + ky*3 + kx padded to 32
    for (int e = t; e < 32 * 32; e += GL_T) {
        const int k = e >> 5, nn = e & 31;
        const float w = k < 27 ? wts.v[k][nn] : 0.f;
        const uint32_t h = tf32_of(w);
        whi[k * SM_WS + nn] = h;
        wlo[k * SM_WS + nn] = tf32_of(w - __uint_as_float(h));
    }
    if (FROM_U8) __syncthreads();                               // the scale table
    // input patch: rows 2*oy0 - 1 .. 2*oy0 + 3, columns -1 .. 255 (stored at column index ix + 1), zero outside the window / image.
    // All of a thread's loads are issued before the first value is used (16 independent requests in flight per thread: with one
    // load per loop iteration the staging was 46 % of the kernel's stall samples).
    constexpr int NE = 5 * 257 * 3, NIT = (NE + GL_T - 1) / GL_T;
    if (FROM_U8) {
        uint32_t raw[NIT];
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            const int e = t + i * GL_T;
            const int row = e / (257 * 3), j = e - row * (257 * 3), col = j / 3, ci = j - col * 3;
            const int iy = 2 * oy0 - 1 + row, ix = col - 1;
            const int y = y_base + iy, x = x_base + ix;
            const bool ok = e < NE && iy >= 0 && iy < H && ix >= 0 && y >= 0 && y < src.IH && x >= 0 && x < src.IW;
            raw[i] = ok ? (uint32_t)__ldg(ip + ((size_t)y * src.IW + x) * 3 + ci) : 256u;
        }
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            const int e = t + i * GL_T;
            const int row = e / (257 * 3), j = e - row * (257 * 3), col = j / 3, ci = j - col * 3;
            if (e < NE) patch[(ci * 5 + row) * SM_PW + col] = raw[i] < 256u ? s_lut[raw[i]] : 0.f;
        }
    } else {
        const float* fp = src.win + (size_t)n * 3 * H * W;
        float raw[NIT];
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            const int e = t + i * GL_T;
            const int ci = e / (5 * 257), rem = e - ci * (5 * 257), row = rem / 257, col = rem - row * 257;
            const int iy = 2 * oy0 - 1 + row, ix = col - 1;
            const bool ok = e < NE && iy >= 0 && iy < H && ix >= 0;
            raw[i] = ok ? __ldg(fp + ((size_t)ci * H + iy) * W + ix) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            const int e = t + i * GL_T;
            const int ci = e / (5 * 257), rem = e - ci * (5 * 257), row = rem / 257, col = rem - row * 257;
            if (e < NE) patch[(ci * 5 + row) * SM_PW + col] = raw[i];
        }
    }
    __syncthreads();
    const int g = lane >> 2, t4 = lane & 3;
    // B fragments: b0 = W[8ks + t4][8nt + g], b1 = W[8ks + t4 + 4][8nt + g], read from shared memory where they are used (conflict-free
    // thanks to the row stride of 40) — holding all 64 of them in registers cost two of the four CTAs an SM can hold
    const uint32_t* wh0 = whi + t4 * SM_WS + g;
    const uint32_t* wl0 = wlo + t4 * SM_WS + g;
    // patch offsets of this lane's two k columns per k-step (k = 8ks + t4, + 4); taps beyond 27 read a zero (column 257.. is never
    // written, so point them at a written zero: row 0 of a window's first patch row may be data — use a flag instead)
    int koff[4][2]; bool kval[4][2];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = 8 * ks + t4 + 4 * h;
            kval[ks][h] = k < 27;
            const int kk = kval[ks][h] ? k : 0;
            koff[ks][h] = ((kk / 9) * 5 + (kk % 9) / 3) * SM_PW + kk % 3;
        }
    float c[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) c[mt][nt][i] = 0.f;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        // rows g and g + 8 of the m-tile = pixels 32*warp + 16*mt + g (+ 8): output row (pixel / 128), column (pixel % 128)
        const int pa = 32 * warp + 16 * mt + g, pb = pa + 8;
        const int offa = (2 * (pa >> 7)) * SM_PW + 2 * (pa & 127), offb = (2 * (pb >> 7)) * SM_PW + 2 * (pb & 127);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            float av[4];
            av[0] = kval[ks][0] ? patch[koff[ks][0] + offa] : 0.f;
            av[1] = kval[ks][0] ? patch[koff[ks][0] + offb] : 0.f;
            av[2] = kval[ks][1] ? patch[koff[ks][1] + offa] : 0.f;
            av[3] = kval[ks][1] ? patch[koff[ks][1] + offb] : 0.f;
            uint32_t ah[4], al[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { ah[i] = tf32_of(av[i]); al[i] = tf32_of(av[i] - __uint_as_float(ah[i])); }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int o0 = 8 * ks * SM_WS + 8 * nt, o1 = o0 + 4 * SM_WS;
                const uint32_t h0 = wh0[o0], h1 = wh0[o1], l0 = wl0[o0], l1 = wl0[o1];
                mma_tf32(c[mt][nt], al, h0, h1);
                mma_tf32(c[mt][nt], ah, l0, l1);
                mma_tf32(c[mt][nt], ah, h0, h1);
            }
        }
    }
    __syncthreads();                                            // everybody is done with the patch and the weights
    // accumulators → pixel-major staging rows: c0, c1 = (pixel g, channels 8nt + 2*t4, + 1), c2, c3 = pixel g + 8
    float* stage = buf;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int pa = 32 * warp + 16 * mt + g;
            *reinterpret_cast<float2*>(stage + pa * STEM_ROW + 8 * nt + 2 * t4) = make_float2(c[mt][nt][0], c[mt][nt][1]);
            *reinterpret_cast<float2*>(stage + (pa + 8) * STEM_ROW + 8 * nt + 2 * t4) = make_float2(c[mt][nt][2], c[mt][nt][3]);
        }
    __syncthreads();
    // coalesced NHWC stores of the CTA's 256-pixel chunk + the chunk statistics of A (as k_stem)
    constexpr int G = STEM_CO / 4;
    const int f4n = np * G;
    float4* Ap = reinterpret_cast<float4*>(A) + ((size_t)n * OHW + p0) * G;
    float4* Rp = reinterpret_cast<float4*>(R) + ((size_t)n * OHW + p0) * G;
    float4 v[G];
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < G; ++i) {
        const int j = t + i * GL_T;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < f4n) {
            const int px = j / G, gg = j % G;
            v[i] = *reinterpret_cast<const float4*>(stage + px * STEM_ROW + gg * 4);
            Ap[j] = v[i];
            Rp[j] = *reinterpret_cast<const float4*>(stage + px * STEM_ROW + STEM_CO + gg * 4);
            s.x += v[i].x; s.y += v[i].y; s.z += v[i].z; s.w += v[i].w;
        }
    }
    cta_channel_sum(s, G, sc);
    const float inv = 1.0f / (float)np;
    const int g4 = (t % G) * 4;
    const float4 m = make_float4(sc.out[g4] * inv, sc.out[g4 + 1] * inv, sc.out[g4 + 2] * inv, sc.out[g4 + 3] * inv);
    const float my_mean = t < STEM_CO ? sc.out[t] * inv : 0.f;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < G; ++i) {
        const int j = t + i * GL_T;
        if (j < f4n) {
            const float a = v[i].x - m.x, b = v[i].y - m.y, cc = v[i].z - m.z, d = v[i].w - m.w;
            q.x += a * a; q.y += b * b; q.z += cc * cc; q.w += d * d;
        }
    }
    cta_channel_sum(q, G, sc);
    if (t < STEM_CO) partials[((size_t)n * gridDim.x + blockIdx.x) * STEM_CO + t] = make_float2(my_mean, sc.out[t]);
}
#endif  // !CS_EMU

// ---------------------------------------------------------------------------
// head, part 1: ConvTranspose2d(32 -> 3, k3, s2, p1, output_padding 1) on cat[xa, xb] (16 channels each, NHWC)
// ---------------------------------------------------------------------------
// out[2iy+dy][2ix+dx] gets:  (0,0): in[iy][ix]·w11      (0,1): in[iy][ix]·w12 + in[iy][ix+1]·w10
//   (1,0): in[iy][ix]·w21 + in[iy+1][ix]·w01      (1,1): in[iy][ix]·w22 + in[iy][ix+1]·w20 + in[iy+1][ix]·w02 + in[iy+1][ix+1]·w00
// One CTA = a tile of 8 x 32 input pixels (one per thread) staged in shared memory together with the row below and
// the column to the right: global reads are whole 64-byte channel vectors of consecutive pixels (coalesced), shared
// reads are conflict-free thanks to the 36-float pixel stride.  Weights: kernel parameters (constant-bank operands).
constexpr int HT_TH = 8, HT_TW = 32;
constexpr int HT_PIX = 36;                                   // floats per staged pixel (32 channels + pad)
struct HeadW { float v[32][28]; };                           // [ci][...]: weight pairs in the order of use (see the accumulator pairs below)

__device__ __forceinline__ void block_sum3(float& a, float& b, float& c, float* red /* [3*8 + 3] */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); c += __shfl_xor_sync(0xffffffffu, c, d);
    }
    __syncthreads();
    if (lane == 0) { red[wid * 3] = a; red[wid * 3 + 1] = b; red[wid * 3 + 2] = c; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
        for (int w = 0; w < GL_T / 32; ++w) s += red[w * 3 + threadIdx.x];
        red[24 + threadIdx.x] = s;
    }
    __syncthreads();
    a = red[24]; b = red[25]; c = red[26];
}

// grid (tiles_x, tiles_y, N).  z: [N][3][2Hin][2Win]; partials: [N][tiles_y*tiles_x][3], each over 4*(pixels of the tile) values.
__global__ void __launch_bounds__(GL_T) k_head_convt(const float* __restrict__ xa, const float* __restrict__ xb, int Hin, int Win,
                                                     const __grid_constant__ HeadW wts, float* __restrict__ z, float2* __restrict__ partials) {
    __shared__ __align__(16) float tile[(HT_TH + 1) * (HT_TW + 1) * HT_PIX];
    __shared__ float red[27];
    const int t = threadIdx.x, n = blockIdx.z;
    const int iy0 = blockIdx.y * HT_TH, ix0 = blockIdx.x * HT_TW;
    {   // staging: thread → (pixel column t/8, 16-byte piece t%8 of its 128-byte channel vector), one row per step
        const int g8 = t & 7, px = t >> 3;
        const float4* srcb = reinterpret_cast<const float4*>((g8 >> 2) ? xb : xa);
#pragma unroll
        for (int row = 0; row < HT_TH + 1; ++row) {
            const int gy = iy0 + row, gx = ix0 + px;
            const bool ok = gy < Hin && gx < Win;
            cp_async16(tile + (row * (HT_TW + 1) + px) * HT_PIX + g8 * 4, ok ? srcb + (((size_t)n * Hin + gy) * Win + gx) * 4 + (g8 & 3) : srcb, ok);
        }
        if (t < (HT_TH + 1) * 8) {           // the extra column to the right
            const int row = t >> 3, gy = iy0 + row, gx = ix0 + HT_TW;
            const bool ok = gy < Hin && gx < Win;
            cp_async16(tile + (row * (HT_TW + 1) + HT_TW) * HT_PIX + g8 * 4, ok ? srcb + (((size_t)n * Hin + gy) * Win + gx) * 4 + (g8 & 3) : srcb, ok);
        }
        cp_async_wait_all();
    }
    __syncthreads();
    const int r = t / HT_TW, c = t % HT_TW;
    const int iy = iy0 + r, ix = ix0 + c;
    const bool valid = iy < Hin && ix < Win;
    // accumulator pairs for FFMA2 (fixed pairing per source pixel, summed at the end):
    //   A[co] = (o00, o01) <- x00·(w11, w12)      B[co] = (o10, o11) <- x00·(w21, w22) + x10·(w01, w02)
    //   Cc[co] = (o01, o11) <- x01·(w10, w20)     D = o11[0..1] <- x11·w00[0..1],  d2 = o11[2] <- x11·w00[2]
    float2 A[3], B[3], Cc[3], D = make_float2(0.f, 0.f);
    float d2 = 0.f;
#pragma unroll
    for (int co = 0; co < 3; ++co) { A[co] = make_float2(0.f, 0.f); B[co] = make_float2(0.f, 0.f); Cc[co] = make_float2(0.f, 0.f); }
    const float4* p00 = reinterpret_cast<const float4*>(tile + (r * (HT_TW + 1) + c) * HT_PIX);
    const float4* p01 = p00 + HT_PIX / 4;
    const float4* p10 = p00 + (HT_TW + 1) * HT_PIX / 4;
    const float4* p11 = p10 + HT_PIX / 4;
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
        const float4 v00 = p00[c4], v01 = p01[c4], v10 = p10[c4], v11 = p11[c4];
        const float a00[4] = {v00.x, v00.y, v00.z, v00.w}, a01[4] = {v01.x, v01.y, v01.z, v01.w};
        const float a10[4] = {v10.x, v10.y, v10.z, v10.w}, a11[4] = {v11.x, v11.y, v11.z, v11.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float* w = wts.v[c4 * 4 + e];
#pragma unroll
            for (int co = 0; co < 3; ++co) {
                fma2(A[co], a00[e], w[2 * co], w[2 * co + 1]);
                fma2(B[co], a00[e], w[6 + 2 * co], w[7 + 2 * co]);
                fma2(Cc[co], a01[e], w[12 + 2 * co], w[13 + 2 * co]);
                fma2(B[co], a10[e], w[18 + 2 * co], w[19 + 2 * co]);
            }
            fma2(D, a11[e], w[24], w[25]);
            d2 = fmaf(a11[e], w[26], d2);
        }
    }
    float o00[3], o01[3], o10[3], o11[3];
#pragma unroll
    for (int co = 0; co < 3; ++co) {
        o00[co] = A[co].x; o01[co] = A[co].y + Cc[co].x; o10[co] = B[co].x;
        o11[co] = (B[co].y + Cc[co].y) + (co == 0 ? D.x : co == 1 ? D.y : d2);
    }
    if (valid) {
        const int OW = 2 * Win;
#pragma unroll
        for (int co = 0; co < 3; ++co) {
            float* zp = z + (((size_t)n * 3 + co) * (2 * Hin) + 2 * iy) * OW + 2 * ix;
            *reinterpret_cast<float2*>(zp) = make_float2(o00[co], o01[co]);
            *reinterpret_cast<float2*>(zp + OW) = make_float2(o10[co], o11[co]);
        }
    }
    const int rows = min(HT_TH, Hin - iy0), cols = min(HT_TW, Win - ix0);
    float s0 = valid ? (o00[0] + o01[0]) + (o10[0] + o11[0]) : 0.f;
    float s1 = valid ? (o00[1] + o01[1]) + (o10[1] + o11[1]) : 0.f;
    float s2 = valid ? (o00[2] + o01[2]) + (o10[2] + o11[2]) : 0.f;
    block_sum3(s0, s1, s2, red);
    const float inv = 1.0f / (4.0f * (float)(rows * cols));
    const float m0 = s0 * inv, m1 = s1 * inv, m2 = s2 * inv;
    float q0 = 0.f, q1 = 0.f, q2 = 0.f;
    if (valid) {
        float d;
        d = o00[0] - m0; q0 += d * d; d = o01[0] - m0; q0 += d * d; d = o10[0] - m0; q0 += d * d; d = o11[0] - m0; q0 += d * d;
        d = o00[1] - m1; q1 += d * d; d = o01[1] - m1; q1 += d * d; d = o10[1] - m1; q1 += d * d; d = o11[1] - m1; q1 += d * d;
        d = o00[2] - m2; q2 += d * d; d = o01[2] - m2; q2 += d * d; d = o10[2] - m2; q2 += d * d; d = o11[2] - m2; q2 += d * d;
    }
    block_sum3(q0, q1, q2, red);
    if (t < 3) {
        const float mm = t == 0 ? m0 : t == 1 ? m1 : m2, qq = t == 0 ? q0 : t == 1 ? q1 : q2;
        partials[((size_t)n * gridDim.y * gridDim.x + (size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3 + t] = make_float2(mm, qq);
    }
}

// ---------------------------------------------------------------------------
// head, part 2: InstanceNorm -> PReLU -> Conv2d(3 -> 3, k3, p1) + bias + identity residual
// ---------------------------------------------------------------------------
constexpr int HF_TH = 16, HF_TW = 128;
constexpr int HF_SW = HF_TW + 2;             // tile row incl. halo
struct FinalW { float w[3][3][9]; float b[3]; };   // [co][ci][ky*3+kx], bias

// grid (ceil(W/128), ceil(H/16), N).  z, out: [N][3][H][W].  Statistics: the tiles k_head_convt wrote for sample n.
__global__ void __launch_bounds__(GL_T) k_head_final(const float* __restrict__ z, int H, int W, const float2* __restrict__ partials, int nstat, ChunkGeom cg,
                                                     const float* __restrict__ prelu_w, const __grid_constant__ FinalW wts, float eps, float* __restrict__ out) {
    __shared__ float tile[3][HF_TH + 2][HF_SW];
    __shared__ float s_mean[4], s_rstd[4];
    __shared__ MergeScratch ms;
    const int t = threadIdx.x, n = blockIdx.z;
    const int y0 = blockIdx.y * HF_TH, x0 = blockIdx.x * HF_TW;
    merge_partials_cta(partials + (size_t)n * nstat * 3, 3, nstat, cg, eps, s_mean, s_rstd, ms);
    const float a = prelu_w ? *prelu_w : 1.0f;
    const float* zn = z + (size_t)n * 3 * H * W;
    const int lane = t & 31, wid = t >> 5;
    for (int pr = wid; pr < 3 * (HF_TH + 2); pr += GL_T / 32) {          // one (channel, tile row) per warp and step
        const int ch = pr / (HF_TH + 2), r = pr % (HF_TH + 2);
        const int gy = y0 + r - 1;
        const bool rowok = gy >= 0 && gy < H;
        const float* srow = zn + ((size_t)ch * H + (rowok ? gy : 0)) * W + (x0 - 1);
        const float mean = s_mean[ch], rstd = s_rstd[ch];
#pragma unroll
        for (int q = 0; q < (HF_SW + 31) / 32; ++q) {
            const int c = lane + 32 * q, gx = x0 + c - 1;
            if (c < HF_SW) {
                float v = 0.f;                   // zero padding applies to the activated tensor
                if (rowok && gx >= 0 && gx < W) v = prelu_f((__ldg(srow + c) - mean) * rstd, a);
                tile[ch][r][c] = v;
            }
        }
    }
    __syncthreads();
    const int r0 = 2 * wid;                  // this warp's two tile rows
    float acc[4][2][3];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int co = 0; co < 3; ++co) acc[q][r][co] = wts.b[co];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;     // tile column; the halo shifts it by one
            float v[4][3];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) v[rr][kx] = tile[ci][r0 + rr][c + kx];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int co = 0; co < 3; ++co) {
                    float s = acc[q][r][co];
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) s = fmaf(wts.w[co][ci][ky * 3 + kx], v[r + ky][kx], s);
                    if (co == ci) s += v[r + 1][1];          // identity residual of the ResidualUnit
                    acc[q][r][co] = s;
                }
            }
        }
    }
    float* on = out + (size_t)n * 3 * H * W;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int gy = y0 + r0 + r;
        if (gy >= H) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int gx = x0 + lane + 32 * q;
            if (gx >= W) continue;
#pragma unroll
            for (int co = 0; co < 3; ++co) on[((size_t)co * H + gy) * W + gx] = acc[q][r][co];
        }
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int l2_order() {
    static const int v = [] { const char* e = getenv("CELLSEG_L2_ORDER"); return e ? atoi(e) : 6; }();     // measured: 20.74 ms (0) -> 20.22 ms (6) per step
    return v;
}
// ---- fused statistics + apply launch (k_nhwc_pair) ----
// Counter sets live in the library (zero-initialised device memory, one set per stream seen, zeroed again by a memset node in front of every
// launch): callers keep the workspace contract of the two-launch pair.  CELLSEG_NORM_PAIR: 0 = never (default), 1 = when the tensor does
// not fit in L2, 2 = always.  CELLSEG_PAIR_LAG: how many items an apply item trails the statistics items of its sample by (default 2048).
// MEASURED (profiles/r02_ncu_norm_pair_fused.txt): the re-read does hit L2 (75 % of the apply pass's sectors; DRAM reads of a 100 MiB site
// 210 -> 105 MB) but the launch is no faster — 67 us against 19 + 34 us for the two kernels: every item is a chain of dependent round trips
// (ticket, flag, partials, loads) at 5 CTAs per SM, and an L2 hit is worth little when the L2 slices top out at ~1.6x the HBM rate.  The
// step is 20.9-21.3 ms with it against 20.05 ms without, so the two-launch pair with alternating traversal order stays the default.
#ifndef CS_EMU
constexpr int PAIR_SETS = 8, PAIR_MAX_N = 65535;
__device__ int g_pair_sync[PAIR_SETS][PAIR_MAX_N + 1];
static int pair_mode() { const char* e = getenv("CELLSEG_NORM_PAIR"); return e ? atoi(e) : 0; }      // read per call: tests switch it
static int pair_lag() { const char* e = getenv("CELLSEG_PAIR_LAG"); const int x = e ? atoi(e) : 2048; return x > 0 ? x : 2048; }
// the counter set of a stream (nullptr: more streams than sets → the caller uses the two-launch pair)
static int* pair_sync_for(cudaStream_t st) {
    static std::mutex mu;
    static cudaStream_t seen[PAIR_SETS];
    static int nseen = 0;
    static int* base = nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!base && cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_pair_sync) != cudaSuccess) { base = nullptr; return nullptr; }
    int k = 0;
    while (k < nseen && seen[k] != st) ++k;
    if (k == nseen) { if (nseen == PAIR_SETS) return nullptr; seen[nseen++] = st; }
    return base + (size_t)k * (PAIR_MAX_N + 1);
}
static bool pair_wanted(int N, long long sample_bytes) {
    const int m = pair_mode();
    return m >= 2 || (m == 1 && (long long)N * sample_bytes > (96ll << 20));
}
// returns CS_OK after launching, or 1 when the caller should run the two-launch pair
template <int KIND>
static int launch_pair(PairArgs a, cudaStream_t st) {
    a.sync = pair_sync_for(st);
    if (!a.sync || a.N > PAIR_MAX_N) return 1;
    a.D = std::min(a.N, std::max(1, ceil_div(pair_lag(), 2 * a.S)));
    const long long items = (long long)(a.N + a.D) * 2 * a.S;
    if (items > 0x3fffffffll) return 1;
    constexpr auto k_pair = k_nhwc_pair<KIND>;
    static int resident = 0;                               // CTAs the device holds at once (per instantiation)
    if (!resident) {
        int dev = 0, sms = 0, per_sm = 0;
        CS_CUDA(cudaGetDevice(&dev));
        CS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pair, GL_T, 0));
        resident = std::max(1, sms * per_sm);
    }
    const long long ctas = std::min<long long>(items, resident);
    CS_CUDA(cudaMemsetAsync(a.sync, 0, sizeof(int) * (size_t)(a.N + 1), st));
    if (KIND == PAIR_DENSE) { constexpr auto k_nhwc_pair_dense = k_pair; CS_LAUNCH(k_nhwc_pair_dense, dim3((unsigned)ctas), GL_T, 0, st, a); }
    else if (KIND == PAIR_SLICES) { constexpr auto k_nhwc_pair_slices = k_pair; CS_LAUNCH(k_nhwc_pair_slices, dim3((unsigned)ctas), GL_T, 0, st, a); }
    else { constexpr auto k_nhwc_pair_shuffle = k_pair; CS_LAUNCH(k_nhwc_pair_shuffle, dim3((unsigned)ctas), GL_T, 0, st, a); }
    CS_LAUNCH_CHECK();
    return CS_OK;
}
#endif

static bool nhwc_channels_ok(int C) { return C >= 4 && C <= 256 && (C & (C - 1)) == 0; }
static int nhwc_chunk_pixels(int C) { return GL_T * GL_VPT / (C / 4); }

int nhwc_stat_chunks(int HW, int C) {
    if (HW <= 0 || !nhwc_channels_ok(C)) return 0;
    return ceil_div(HW, nhwc_chunk_pixels(C));
}

int nhwc_stats(const float* x, int N, int HW, int C, float* partials, cudaStream_t st) {
    CS_CHECK_ARG(x && partials, "null pointer");
    CS_CHECK_ARG(N > 0 && N <= 65535 && HW > 0, "bad N / HW");
    CS_CHECK_ARG(nhwc_channels_ok(C), "channel count must be a power of two in [4, 256]");
    CS_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)partials & 7) == 0, "misaligned pointer");
    const int P = nhwc_chunk_pixels(C);
    CS_LAUNCH(k_nhwc_stats, dim3(ceil_div(HW, P), N), GL_T, 0, st, x, HW, C, P, reinterpret_cast<float2*>(partials), l2_order());
    CS_LAUNCH_CHECK();
    return CS_OK;
}

static int nhwc_norm_prelu2(const float* x, const float* res, float* y, const float* prelu_w, const float* res_bias, int N, int HW, int C,
                           const float* partials, int stat_chunk_pixels, float eps, float* y2, int c2_total, int c2_off, cudaStream_t st);
int nhwc_norm_prelu(const float* x, const float* res, float* y, const float* prelu_w, const float* res_bias, int N, int HW, int C,
                    const float* partials, int stat_chunk_pixels, float eps, cudaStream_t st) {
    CS_CHECK_ARG(y, "null pointer");
    return nhwc_norm_prelu2(x, res, y, prelu_w, res_bias, N, HW, C, partials, stat_chunk_pixels, eps, nullptr, 0, 0, st);
}
static int nhwc_norm_prelu2(const float* x, const float* res, float* y, const float* prelu_w, const float* res_bias, int N, int HW, int C,
                           const float* partials, int stat_chunk_pixels, float eps, float* y2, int c2_total, int c2_off, cudaStream_t st) {
    CS_CHECK_ARG(x && (y || y2) && partials, "null pointer");
    CS_CHECK_ARG(!y2 || (c2_total >= c2_off + C && c2_off >= 0 && (c2_total & 3) == 0 && (c2_off & 3) == 0 && ((uintptr_t)y2 & 15) == 0),
                 "concatenation slot must be 4-channel aligned and inside the destination");
    CS_CHECK_ARG(N > 0 && N <= 65535 && HW > 0 && stat_chunk_pixels > 0, "bad N / HW / chunk");
    CS_CHECK_ARG(nhwc_channels_ok(C), "channel count must be a power of two in [4, 256]");
    CS_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)res & 15) == 0 && ((uintptr_t)partials & 7) == 0, "misaligned pointer");
    const int P = nhwc_chunk_pixels(C);
    CS_LAUNCH(k_nhwc_norm_prelu, dim3(ceil_div(HW, P), N), GL_T, 0, st, x, res, y, prelu_w, res_bias, HW, C, P,
              reinterpret_cast<const float2*>(partials), ceil_div(HW, stat_chunk_pixels), stat_chunk_pixels, eps, y2, c2_total, c2_off, l2_order());
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// y = prelu(instance_norm(x)) (+ res + res_bias) for x [N][HW][C].  With a `partials` workspace (float2[N][nhwc_stat_chunks(HW, C)][C])
// the statistics + apply pair runs — the faster choice on B200 (measured on the 15 cuDNN-fed norm sites of a 100-window forward:
// 447 us against 726 us; many small CTAs per SM overlap their load and store phases, one 128-register CTA per SM cannot, and
// below 1 MiB per sample the second read comes from L2 anyway).  Without a workspace, samples of up to 1 MiB take the
// single-pass cluster kernel.
// y2 / c2_total / c2_off: optional concatenation slot, see k_nhwc_norm_prelu (statistics + apply pair only).
int instnorm_prelu_nhwc(const float* x, const float* res, float* y, const float* prelu_w, const float* res_bias, int N, int HW, int C, float eps,
                        float* partials, float* y2, int c2_total, int c2_off, cudaStream_t st) {
    CS_CHECK_ARG(x && (y || y2), "null pointer");
    CS_CHECK_ARG(!y2 || partials, "the concatenation slot needs the partial-statistics workspace");
    CS_CHECK_ARG(N > 0 && N <= 65535 && HW > 0, "bad N / HW");
    CS_CHECK_ARG(nhwc_channels_ok(C), "channel count must be a power of two in [4, 256]");
    CS_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)res & 15) == 0, "misaligned pointer");
#ifndef CS_EMU
    const long long f4 = (long long)HW * (C / 4);
    const long long per_cta = (long long)CL_T * CL_VPT;
    if (!partials && f4 <= 8 * per_cta) {
        int cl = 1;
        while ((long long)cl * per_cta < f4) cl <<= 1;
        const int P = ceil_div(HW, cl);                       // pixels per CTA; P * G <= per_cta because per_cta is a multiple of G
        if ((long long)P * (C / 4) <= per_cta) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cl, N); cfg.blockDim = dim3(CL_T); cfg.dynamicSmemBytes = 0; cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cs::count_launch();
            cs::prof::Scope scope("k_nhwc_norm_prelu_cluster", st);
            CS_CUDA(cudaLaunchKernelEx(&cfg, k_nhwc_norm_prelu_cluster, x, res, y, prelu_w, res_bias, HW, C, P, eps));
            return CS_OK;
        }
    }
#endif
    CS_CHECK_ARG(partials, "samples above 1 MiB need the partial-statistics workspace");
#ifndef CS_EMU
    if (pair_wanted(N, (long long)HW * C * 4)) {
        CS_CHECK_ARG(!y2 || (c2_total >= c2_off + C && c2_off >= 0 && (c2_total & 3) == 0 && (c2_off & 3) == 0 && ((uintptr_t)y2 & 15) == 0),
                     "concatenation slot must be 4-channel aligned and inside the destination");
        CS_CHECK_ARG(((uintptr_t)partials & 7) == 0, "misaligned pointer");
        PairArgs a = {};
        a.x = x; a.partials = reinterpret_cast<float2*>(partials);
        a.HW = HW; a.C = C; a.Cv = C; a.P = nhwc_chunk_pixels(C); a.N = N; a.S = ceil_div(HW, a.P);
        a.na = NormArgs{res, y, prelu_w, res_bias, a.partials, a.S, a.P, eps, y2, c2_total, c2_off};
        const int prc = launch_pair<PAIR_DENSE>(a, st);
        if (prc <= 0) return prc;
    }
#endif
    int rc = nhwc_stats(x, N, HW, C, partials, st);
    if (rc != CS_OK) return rc;
    return nhwc_norm_prelu2(x, res, y, prelu_w, res_bias, N, HW, C, partials, nhwc_chunk_pixels(C), eps, y2, c2_total, c2_off, st);
}

// The norm block on channel slices: x = channels [x_coff, x_coff + C) of an NHWC tensor with x_ctotal channels (likewise the residual);
// y dense [N][HW][C] and/or the concatenation slot y2.  All offsets and totals multiples of 4.
int instnorm_prelu_nhwc_slices(const float* x, int x_ctotal, int x_coff, const float* res, int r_ctotal, int r_coff, float* y, const float* prelu_w,
                               const float* res_bias, int N, int HW, int C, float eps, float* partials, float* y2, int c2_total, int c2_off,
                               cudaStream_t st) {
    CS_CHECK_ARG(x && (y || y2) && partials, "null pointer");
    CS_CHECK_ARG(N > 0 && N <= 65535 && HW > 0, "bad N / HW");
    CS_CHECK_ARG(nhwc_channels_ok(C), "channel count must be a power of two in [4, 256]");
    CS_CHECK_ARG(x_ctotal >= x_coff + C && x_coff >= 0 && ((x_ctotal | x_coff) & 3) == 0, "bad input channel slice");
    CS_CHECK_ARG(!res || (r_ctotal >= r_coff + C && r_coff >= 0 && ((r_ctotal | r_coff) & 3) == 0), "bad residual channel slice");
    CS_CHECK_ARG(!y2 || (c2_total >= c2_off + C && c2_off >= 0 && ((c2_total | c2_off) & 3) == 0), "bad concatenation slot");
    CS_CHECK_ARG((((uintptr_t)x | (uintptr_t)y | (uintptr_t)res | (uintptr_t)y2) & 15) == 0 && ((uintptr_t)partials & 7) == 0, "misaligned pointer");
    const int P = nhwc_chunk_pixels(C);
    const View xv = make_view(HW, 0, x_ctotal / 4, x_coff / 4, (long long)HW * (x_ctotal / 4));      // one "row" of HW pixels
    const View rv = make_view(HW, 0, r_ctotal / 4, r_coff / 4, (long long)HW * (r_ctotal / 4));
    float2* pt = reinterpret_cast<float2*>(partials);
#ifndef CS_EMU
    if (pair_wanted(N, (long long)HW * C * 4)) {
        PairArgs a = {};
        a.x = x; a.xv = xv; a.rv = rv; a.partials = pt;
        a.HW = HW; a.C = C; a.Cv = C; a.P = P; a.N = N; a.S = ceil_div(HW, P);
        a.na = NormArgs{res, y, prelu_w, res_bias, pt, a.S, P, eps, y2, c2_total, c2_off};
        const int prc = launch_pair<PAIR_SLICES>(a, st);
        if (prc <= 0) return prc;
    }
#endif
    CS_LAUNCH(k_nhwc_stats_view, dim3(ceil_div(HW, P), N), GL_T, 0, st, x, xv, HW, C, P, pt, l2_order());
    CS_LAUNCH_CHECK();
    constexpr auto k_nhwc_norm_prelu_slices = k_nhwc_norm_prelu_view<false>;
    CS_LAUNCH(k_nhwc_norm_prelu_slices, dim3(ceil_div(HW, P), N), GL_T, 0, st, x, xv, res, rv, y, prelu_w, res_bias, HW, C, P, pt, ceil_div(HW, P), P, eps,
              y2, c2_total, c2_off, l2_order());
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// The norm block behind a transposed convolution computed as a 2x2 convolution: x [N][H+1][W+1][4C] (used from row 1 / column 1
// on; channel = (dy*2 + dx)*C + c) -> y [N][2H][2W][C] = prelu(instance_norm(depth_to_space(x))).  partials: float2[N][chunks][4C]
// with chunks = nhwc_stat_chunks(H*W, 4C).
int instnorm_prelu_nhwc_shuffle(const float* x, float* y, const float* prelu_w, int N, int H, int W, int C, float eps, float* partials, cudaStream_t st) {
    CS_CHECK_ARG(x && y && partials, "null pointer");
    CS_CHECK_ARG(N > 0 && N <= 65535 && H > 0 && W > 0, "bad geometry");
    CS_CHECK_ARG(nhwc_channels_ok(C) && 4 * C <= 256, "channel count must be a power of two in [4, 64]");
    CS_CHECK_ARG((((uintptr_t)x | (uintptr_t)y) & 15) == 0 && ((uintptr_t)partials & 7) == 0, "misaligned pointer");
    const int HW = H * W, Cv = 4 * C, P = nhwc_chunk_pixels(Cv);
    const View xv = make_view(W, (W + 1) * C, C, ((W + 1) + 1) * C, (long long)(H + 1) * (W + 1) * C);      // float4 units: Cv/4 = C per pixel
    float2* pt = reinterpret_cast<float2*>(partials);
#ifndef CS_EMU
    if (pair_wanted(N, (long long)HW * Cv * 4)) {
        PairArgs a = {};
        a.x = x; a.xv = xv; a.rv = xv; a.partials = pt;
        a.HW = HW; a.C = C; a.Cv = Cv; a.P = P; a.N = N; a.S = ceil_div(HW, P);
        a.na = NormArgs{nullptr, y, prelu_w, nullptr, pt, a.S, P, eps, nullptr, 0, 0};
        const int prc = launch_pair<PAIR_SHUFFLE>(a, st);
        if (prc <= 0) return prc;
    }
#endif
    CS_LAUNCH(k_nhwc_stats_view, dim3(ceil_div(HW, P), N), GL_T, 0, st, x, xv, HW, Cv, P, pt, l2_order());
    CS_LAUNCH_CHECK();
    constexpr auto k_nhwc_norm_prelu_shuffle = k_nhwc_norm_prelu_view<true>;
    CS_LAUNCH(k_nhwc_norm_prelu_shuffle, dim3(ceil_div(HW, P), N), GL_T, 0, st, x, xv, (const float*)nullptr, xv, y, prelu_w, (const float*)nullptr,
              HW, C, 4 * P, pt, ceil_div(HW, P), P, eps, (float*)nullptr, 0, 0, l2_order());
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// w_unit0_host / w_res_host: HOST copies of the two [16][3][3][3] weights (they become kernel parameters).
static StemW stem_weights(const float* w_unit0_host, const float* w_res_host) {
    StemW w;
    for (int tap = 0; tap < 27; ++tap)
        for (int co = 0; co < STEM_CO; ++co) { w.v[tap][co] = w_unit0_host[co * 27 + tap]; w.v[tap][STEM_CO + co] = w_res_host[co * 27 + tap]; }
    return w;
}

int unet_stem(const float* win, int N, int H, int W, const float* w_unit0_host, const float* w_res_host, float* a_out, float* r_out, float* a_partials,
              cudaStream_t st) {
    CS_CHECK_ARG(win && w_unit0_host && w_res_host && a_out && r_out && a_partials, "null pointer");
    CS_CHECK_ARG(N > 0 && N <= 65535 && H > 0 && W > 0 && (H & 1) == 0 && (W & 1) == 0, "window batch must have even H and W");
    CS_CHECK_ARG(((uintptr_t)a_out & 15) == 0 && ((uintptr_t)r_out & 15) == 0 && ((uintptr_t)a_partials & 7) == 0, "misaligned pointer");
    const StemW w = stem_weights(w_unit0_host, w_res_host);
    StemSrc src = {};
    src.win = win;
#ifndef CS_EMU
    if (W == 256 && getenv("CELLSEG_STEM_MMA")) {                 // opt-in: measured 0.80 ms against 0.37 ms (FFMA2 kernel) per 400 windows
        constexpr auto k_stem_mma_f32 = k_stem_mma<false>;
        CS_LAUNCH(k_stem_mma_f32, dim3(ceil_div((H / 2) * 128, GL_T), N), GL_T, 0, st, src, H, w, a_out, r_out, reinterpret_cast<float2*>(a_partials));
        CS_LAUNCH_CHECK();
        return CS_OK;
    }
#endif
    CS_LAUNCH(k_stem<false>, dim3(ceil_div((H / 2) * (W / 2), GL_T), N), GL_T, 0, st, src, H, W, w, a_out, r_out, reinterpret_cast<float2*>(a_partials));
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// The same convolutions on windows [win_begin, win_begin + win_count) of the window grid `g`, read straight from the uint8 images.
int unet_stem_u8(const uint8_t* img, const float* lut, WinGeo g, const int* ys, const int* xs, int win_begin, int win_count,
                 const float* w_unit0_host, const float* w_res_host, float* a_out, float* r_out, float* a_partials, cudaStream_t st) {
    CS_CHECK_ARG(img && lut && ys && xs && w_unit0_host && w_res_host && a_out && r_out && a_partials, "null pointer");
    CS_CHECK_ARG(g.N > 0 && g.H > 0 && g.W > 0 && g.R > 0 && (g.R & 1) == 0 && g.ny > 0 && g.nx > 0, "bad window geometry");
    CS_CHECK_ARG(win_begin >= 0 && win_count > 0 && win_count <= 65535 && (long long)win_begin + win_count <= (long long)g.N * g.ny * g.nx, "window range out of bounds");
    CS_CHECK_ARG(((uintptr_t)a_out & 15) == 0 && ((uintptr_t)r_out & 15) == 0 && ((uintptr_t)a_partials & 7) == 0, "misaligned pointer");
    const StemW w = stem_weights(w_unit0_host, w_res_host);
    StemSrc src = {};
    src.img = img; src.lut = lut; src.ys = ys; src.xs = xs;
    src.IH = g.H; src.IW = g.W; src.ny = g.ny; src.nx = g.nx; src.pad_y = g.pad_y; src.pad_x = g.pad_x; src.win_begin = win_begin;
#ifndef CS_EMU
    if (g.R == 256 && getenv("CELLSEG_STEM_MMA")) {
        constexpr auto k_stem_mma_u8 = k_stem_mma<true>;
        CS_LAUNCH(k_stem_mma_u8, dim3(ceil_div(128 * 128, GL_T), win_count), GL_T, 0, st, src, g.R, w, a_out, r_out, reinterpret_cast<float2*>(a_partials));
        CS_LAUNCH_CHECK();
        return CS_OK;
    }
#endif
    CS_LAUNCH(k_stem<true>, dim3(ceil_div((g.R / 2) * (g.R / 2), GL_T), win_count), GL_T, 0, st, src, g.R, g.R, w, a_out, r_out,
              reinterpret_cast<float2*>(a_partials));
    CS_LAUNCH_CHECK();
    return CS_OK;
}

int unet_head_partial_chunks(int Hin, int Win) {
    if (Hin <= 0 || Win <= 0) return 0;
    return ceil_div(Hin, HT_TH) * ceil_div(Win, HT_TW);
}

// wt_host: HOST copy of the ConvTranspose2d weight [32][3][3][3] = [ci][co][ky][kx].
int unet_head_convt(const float* xa, const float* xb, int N, int Hin, int Win, const float* wt_host, float* z, float* z_partials, cudaStream_t st) {
    CS_CHECK_ARG(xa && xb && wt_host && z && z_partials, "null pointer");
    CS_CHECK_ARG(N > 0 && N <= 65535 && Hin > 0 && Win > 0 && ceil_div(Hin, HT_TH) <= 65535, "bad geometry");
    CS_CHECK_ARG(((uintptr_t)xa & 15) == 0 && ((uintptr_t)xb & 15) == 0 && ((uintptr_t)z & 7) == 0 && ((uintptr_t)z_partials & 7) == 0, "misaligned pointer");
    HeadW w;
    auto W = [&](int ci, int co, int ky, int kx) { return wt_host[((ci * 3 + co) * 3 + ky) * 3 + kx]; };
    for (int ci = 0; ci < 32; ++ci) {
        for (int co = 0; co < 3; ++co) {
            w.v[ci][2 * co] = W(ci, co, 1, 1);      w.v[ci][2 * co + 1] = W(ci, co, 1, 2);        // x00 -> (o00, o01)
            w.v[ci][6 + 2 * co] = W(ci, co, 2, 1);  w.v[ci][7 + 2 * co] = W(ci, co, 2, 2);        // x00 -> (o10, o11)
            w.v[ci][12 + 2 * co] = W(ci, co, 1, 0); w.v[ci][13 + 2 * co] = W(ci, co, 2, 0);       // x01 -> (o01, o11)
            w.v[ci][18 + 2 * co] = W(ci, co, 0, 1); w.v[ci][19 + 2 * co] = W(ci, co, 0, 2);       // x10 -> (o10, o11)
            w.v[ci][24 + co] = W(ci, co, 0, 0);                                                   // x11 -> o11
        }
        w.v[ci][27] = 0.f;
    }
    CS_LAUNCH(k_head_convt, dim3(ceil_div(Win, HT_TW), ceil_div(Hin, HT_TH), N), GL_T, 0, st, xa, xb, Hin, Win, w, z, reinterpret_cast<float2*>(z_partials));
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// w3_host [3][3][3][3], b3_host [3] (may be null): HOST copies.  z_partials: what unet_head_convt wrote for z (H = 2Hin, W = 2Win).
int unet_head_final(const float* z, const float* z_partials, int N, int H, int W, const float* prelu_w, const float* w3_host, const float* b3_host, float eps,
                    float* out, cudaStream_t st) {
    CS_CHECK_ARG(z && z_partials && w3_host && out, "null pointer");
    CS_CHECK_ARG(N > 0 && N <= 65535 && H > 0 && W > 0 && (H & 1) == 0 && (W & 1) == 0, "bad geometry");
    CS_CHECK_ARG(ceil_div(H, HF_TH) <= 65535, "image too tall");
    FinalW w;
    for (int co = 0; co < 3; ++co) {
        for (int ci = 0; ci < 3; ++ci)
            for (int k = 0; k < 9; ++k) w.w[co][ci][k] = w3_host[(co * 3 + ci) * 9 + k];
        w.b[co] = b3_host ? b3_host[co] : 0.f;
    }
    const int Hin = H / 2, Win = W / 2;
    ChunkGeom cg = {1, 0, H * W, ceil_div(Win, HT_TW), HT_TH, HT_TW, Hin, Win, 4};
    CS_LAUNCH(k_head_final, dim3(ceil_div(W, HF_TW), ceil_div(H, HF_TH), N), GL_T, 0, st, z, H, W, reinterpret_cast<const float2*>(z_partials),
              unet_head_partial_chunks(Hin, Win), cg, prelu_w, w, eps, out);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace cs
