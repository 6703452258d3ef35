// Sliding-window glue around the (out-of-scope, cuDNN) UNet forward:
//   K0  u8 HWC image → f32 CHW in [0,1]            (baseline/predict.py:107-108)
//   K1  window gather for ALL windows of a batch   (monai sliding_window_inference, predict.py:109)
//   K2  weighted blend of window logits, divide by the weight sum, class softmax, p1 > 0.5
//       → bit-packed foreground mask              (predict.py:109-113)
// Arithmetic follows MONAI's accumulation exactly: windows are visited in ascending window index,
// `acc = acc + (w * x)` with the product and the sum rounded separately (no FMA), then `acc / cnt`,
// then torch's softmax formula exp(x - max) / sum with the sum taken in class order.
#include "ccl_tile.cuh"
#include "internal.cuh"
#ifndef CS_EMU
#include <cuda.h>
#include <cstdlib>
#include <algorithm>
#endif

namespace cs {

// ---------------------------------------------------------------------------
// K0: per-image max, 256-entry scale LUT, u8 HWC → f32 CHW
// ---------------------------------------------------------------------------
// per-byte unsigned maximum of two packed words
__device__ __forceinline__ uint32_t vmaxu4(uint32_t a, uint32_t b) {
#ifdef CS_EMU
    uint32_t r = 0;
    for (int s = 0; s < 32; s += 8) r |= max((a >> s) & 0xffu, (b >> s) & 0xffu) << s;
    return r;
#else
    return __vmaxu4(a, b);
#endif
}
__global__ void k_u8_max(const uint8_t* __restrict__ img, size_t bytes_per_image, uint32_t* __restrict__ maxv) {
    const int n = blockIdx.y;
    const uint8_t* p = img + (size_t)n * bytes_per_image;
    const size_t nvec = (((uintptr_t)p & 15) == 0) ? bytes_per_image / 16 : 0;
    const uint4* v = reinterpret_cast<const uint4*>(p);
    // byte-wise maxima of four packed bytes at a time; four independent 16-byte loads in flight per thread and iteration
    uint32_t m4 = 0u;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        const uint4 a = __ldcs(v + i), b = __ldcs(v + i + stride), c = __ldcs(v + i + 2 * stride), d = __ldcs(v + i + 3 * stride);
        m4 = vmaxu4(m4, vmaxu4(vmaxu4(vmaxu4(a.x, a.y), vmaxu4(a.z, a.w)), vmaxu4(vmaxu4(b.x, b.y), vmaxu4(b.z, b.w))));
        m4 = vmaxu4(m4, vmaxu4(vmaxu4(vmaxu4(c.x, c.y), vmaxu4(c.z, c.w)), vmaxu4(vmaxu4(d.x, d.y), vmaxu4(d.z, d.w))));
    }
    for (; i < nvec; i += stride) {
        const uint4 a = v[i];
        m4 = vmaxu4(m4, vmaxu4(vmaxu4(a.x, a.y), vmaxu4(a.z, a.w)));
    }
    uint32_t m = max(max(m4 & 0xffu, (m4 >> 8) & 0xffu), max((m4 >> 16) & 0xffu, m4 >> 24));
    for (size_t j = nvec * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < bytes_per_image; j += stride)
        m = max(m, (uint32_t)p[j]);
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&maxv[n], m);
}

// lut[n][v] = (float)((double)v / (double)max_n): numpy's uint8/uint8 true-divide is float64, the
// later .type(FloatTensor) rounds to float32 (predict.py:107-108).  max == 0 gives 0/0 = NaN like numpy.
__global__ void k_scale_lut(const uint32_t* __restrict__ maxv, float* __restrict__ lut, int N) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * 256) return;
    int n = i >> 8, v = i & 255;
    double d = __ddiv_rn(__uint2double_rn((uint32_t)v), __uint2double_rn(maxv[n]));
    lut[i] = __double2float_rn(d);
}

__global__ void k_u8hwc_to_f32chw(const uint8_t* __restrict__ img, int N, int H, int W, const float* __restrict__ lut,
                                  float* __restrict__ out) {
    const size_t HW = (size_t)H * W;
    const size_t total = (size_t)N * HW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t n = i / HW, px = i % HW;
        const uint8_t* s = img + i * 3;
        const float* l = lut + n * 256;
        float* o = out + n * 3 * HW + px;
        o[0] = l[s[0]]; o[HW] = l[s[1]]; o[2 * HW] = l[s[2]];
    }
}

// ---------------------------------------------------------------------------
// K1: gather windows.  One thread per 4 consecutive window pixels of one channel-triple.
// Window w of image n is placed at batch index n*n_win + w; its top-left corner in PADDED
// coordinates is (ys[w / nx], xs[w % nx]); padded coordinate (Y,X) ↔ image (Y-pad_y, X-pad_x);
// out-of-image reads are the zero padding MONAI adds for images smaller than the ROI.
// ---------------------------------------------------------------------------
template <bool FROM_U8>
__global__ void __launch_bounds__(256)
k_gather(const void* __restrict__ src, const float* __restrict__ lut, WinGeo g, const int* __restrict__ ys,
         const int* __restrict__ xs, int win_begin, int win_count, float* __restrict__ batch) {
    const int R = g.R, R4 = R / 4;
    const size_t per_win = (size_t)R * R4;
    const size_t total = (size_t)win_count * per_win;
    const size_t HW = (size_t)g.H * g.W;
    const int n_win = g.ny * g.nx;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int wl = (int)(i / per_win);
        const int rem = (int)(i % per_win);
        const int wy = rem / R4, wx = (rem % R4) * 4;
        const int gw = win_begin + wl;
        const int n = gw / n_win, w = gw % n_win;
        const int y = ys[w / g.nx] + wy - g.pad_y;
        const int x0 = xs[w % g.nx] + wx - g.pad_x;
        float v[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) v[c][j] = 0.f;
        if (y >= 0 && y < g.H) {
            if (FROM_U8) {
                const uint8_t* s = (const uint8_t*)src + ((size_t)n * HW + (size_t)y * g.W) * 3;
                const float* l = lut + (size_t)n * 256;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int x = x0 + j;
                    if (x >= 0 && x < g.W) {
                        v[0][j] = l[s[(size_t)x * 3 + 0]]; v[1][j] = l[s[(size_t)x * 3 + 1]]; v[2][j] = l[s[(size_t)x * 3 + 2]];
                    }
                }
            } else {
                const float* s = (const float*)src + (size_t)n * 3 * HW + (size_t)y * g.W;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int x = x0 + j;
                    if (x >= 0 && x < g.W) { v[0][j] = s[x]; v[1][j] = s[HW + x]; v[2][j] = s[2 * HW + x]; }
                }
            }
        }
        float* o = batch + ((size_t)wl * 3 * R + wy) * R + wx;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            *reinterpret_cast<float4*>(o + (size_t)c * R * R) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
    }
}

// ---------------------------------------------------------------------------
// K2: blend + normalise + softmax + threshold
// ---------------------------------------------------------------------------
template <int C>
__device__ __noinline__ bool blend_pixel(const BlendArgs& a, const WinGeo& g, int n, int y, int x, float* out_p1, float* out_logits) {
    const int Y = y + g.pad_y, X = x + g.pad_x, R = g.R;
    const int cy = a.cover_y[Y], cx = a.cover_x[X];
    const int iy0 = cy >> 8, iyn = cy & 255, ix0 = cx >> 8, ixn = cx & 255;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    float cnt = 0.f;
    const size_t n_win = (size_t)g.ny * g.nx;
    for (int iy = iy0; iy < iy0 + iyn; ++iy) {
        const int wy = Y - a.ys[iy];
        for (int ix = ix0; ix < ix0 + ixn; ++ix) {
            const int wx = X - a.xs[ix];
            const float* p = a.win + (((size_t)n * n_win + (size_t)iy * g.nx + ix) * C * R + wy) * R + wx;
            const float wgt = a.imap ? a.imap[wy * R + wx] : 1.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(wgt, p[(size_t)c * R * R]));
            cnt = __fadd_rn(cnt, wgt);
        }
    }
    float o[C];
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = __fdiv_rn(acc[c], cnt);
    float m = o[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, o[c]);
    float s = 0.f, e1 = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { float e = expf(o[c] - m); s = __fadd_rn(s, e); if (c == 1) e1 = e; }
    const float p = __fdiv_rn(e1, s);
    *out_p1 = p;
    if (out_logits) {
#pragma unroll
        for (int c = 0; c < C; ++c) out_logits[c] = o[c];
    }
    return p > a.threshold;
}

// Four horizontally adjacent pixels that share one set of covering windows, all at window x-offsets that are
// multiples of 4: every window row segment is one aligned float4 per class.  Same operation order as
// blend_pixel, lane-wise.  Returns the 4 foreground bits.
// acc += wg * v (product and sum rounded separately, monai's `out += w * x; cnt += w`), lane-wise on four pixels
template <int C>
__device__ __forceinline__ void quad_add(float4 (&acc)[C], float4& cnt, const float4& wg, const float4 (&v)[C]) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
        acc[c].x = __fadd_rn(acc[c].x, __fmul_rn(wg.x, v[c].x));
        acc[c].y = __fadd_rn(acc[c].y, __fmul_rn(wg.y, v[c].y));
        acc[c].z = __fadd_rn(acc[c].z, __fmul_rn(wg.z, v[c].z));
        acc[c].w = __fadd_rn(acc[c].w, __fmul_rn(wg.w, v[c].w));
    }
    cnt.x = __fadd_rn(cnt.x, wg.x); cnt.y = __fadd_rn(cnt.y, wg.y);
    cnt.z = __fadd_rn(cnt.z, wg.z); cnt.w = __fadd_rn(cnt.w, wg.w);
}

// acc / cnt → softmax → class 1 against the threshold, for four pixels.  Returns the 4 foreground bits.
template <int C, bool MASK_ONLY>
__device__ __forceinline__ uint32_t quad_finish(const float4 (&acc)[C], const float4& cnt, float threshold, float4* out_p1, float4* out_logits) {
    uint32_t bits = 0u;
    float pr[4], ol[C][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float cj = j == 0 ? cnt.x : j == 1 ? cnt.y : j == 2 ? cnt.z : cnt.w;
        // x / 1, x / 2, x / 4 are exact scalings: the product with the exact reciprocal IS the correctly rounded quotient
        const bool pow2 = cj == 1.0f || cj == 2.0f || cj == 4.0f;
        const float rc = cj == 1.0f ? 1.0f : cj == 2.0f ? 0.5f : 0.25f;
        float o[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float aj = j == 0 ? acc[c].x : j == 1 ? acc[c].y : j == 2 ? acc[c].z : acc[c].w;
            o[c] = pow2 ? __fmul_rn(aj, rc) : __fdiv_rn(aj, cj);
        }
        if (MASK_ONLY) {
            // p1 > 0.5 without the softmax (exact): class 1 must be the strict maximum — then
            // e1 = exp(0) = 1 and fl(1/s) > 0.5 ⇔ s < 2 for the float sum s taken in class order.
            bool strict = true;
#pragma unroll
            for (int c = 0; c < C; ++c) if (c != 1) strict = strict && (o[1] > o[c]);
            if (strict) {
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) s = __fadd_rn(s, c == 1 ? 1.0f : expf(o[c] - o[1]));
                bits |= (s < 2.0f ? 1u : 0u) << j;
            }
            pr[j] = 0.f;
        } else {
            float m = o[0];
#pragma unroll
            for (int c = 1; c < C; ++c) m = fmaxf(m, o[c]);
            float s = 0.f, e1 = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) { float e = expf(o[c] - m); s = __fadd_rn(s, e); if (c == 1) e1 = e; }
            pr[j] = __fdiv_rn(e1, s);
            bits |= (pr[j] > threshold ? 1u : 0u) << j;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) ol[c][j] = o[c];
    }
    if (out_logits) {
#pragma unroll
        for (int c = 0; c < C; ++c) out_logits[c] = make_float4(ol[c][0], ol[c][1], ol[c][2], ol[c][3]);
    }
    *out_p1 = make_float4(pr[0], pr[1], pr[2], pr[3]);
    return bits;
}

template <int C, bool MASK_ONLY, bool HAS_IMAP>
__device__ __forceinline__ uint32_t blend_quad(const BlendArgs& a, const WinGeo& g, int n, int y, int x0, int cy, int cx,
                                               float4* out_p1, float4* out_logits) {
    const int Y = y + g.pad_y, X0 = x0 + g.pad_x, R = g.R;
    const int iy0 = cy >> 8, iyn = cy & 255, ix0 = cx >> 8, ixn = cx & 255;
    float4 acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 cnt = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t n_win = (size_t)g.ny * g.nx;
    const size_t plane = (size_t)R * R;
    const int nw = iyn * ixn;
    if (nw <= 4) {
        // the usual case (≤ 2 windows per axis): issue the loads of two covering windows at a time before adding, so
        // that 2 x C independent 16-byte requests are in flight per thread; accumulate in ascending window index.
#pragma unroll
        for (int kk = 0; kk < 4; kk += 2) {
            if (kk < nw) {
                float4 v[2][C], wg[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int k = kk + j;
                    if (k < nw) {
                        const int ky = ixn == 1 ? k : (ixn == 2 ? (k >> 1) : 0), kx = k - ky * ixn;
                        const int iy = iy0 + ky, ix = ix0 + kx;
                        const int wy = Y - a.ys[iy], wx = X0 - a.xs[ix];
                        const float* p = a.win + (((size_t)n * n_win + (size_t)iy * g.nx + ix) * C * R + wy) * R + wx;
                        wg[j] = HAS_IMAP ? *reinterpret_cast<const float4*>(a.imap + wy * R + wx) : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
                        for (int c = 0; c < C; ++c) v[j][c] = __ldcs(reinterpret_cast<const float4*>(p + (size_t)c * plane));   // read once: stream
                    }
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (kk + j < nw) {
                        quad_add<C>(acc, cnt, wg[j], v[j]);
                    }
                }
            }
        }
    } else {
        for (int iy = iy0; iy < iy0 + iyn; ++iy) {
            const int wy = Y - a.ys[iy];
            for (int ix = ix0; ix < ix0 + ixn; ++ix) {
                const int wx = X0 - a.xs[ix];
                const float* p = a.win + (((size_t)n * n_win + (size_t)iy * g.nx + ix) * C * R + wy) * R + wx;
                float4 wgt = make_float4(1.f, 1.f, 1.f, 1.f);
                if (HAS_IMAP) wgt = *reinterpret_cast<const float4*>(a.imap + wy * R + wx);
                float4 v[C];
#pragma unroll
                for (int c = 0; c < C; ++c) v[c] = __ldcs(reinterpret_cast<const float4*>(p + (size_t)c * plane));
                quad_add<C>(acc, cnt, wgt, v);
            }
        }
    }
    return quad_finish<C, MASK_ONLY>(acc, cnt, a.threshold, out_p1, out_logits);
}
#ifndef BLEND_MIN_CTAS
#define BLEND_MIN_CTAS 4      // measured on 64 x 1024^2: 3 CTAs per SM (80 registers) 0.389 ms, 4 (64 registers, a few spills) 0.359 ms, 5 (48) 0.367 ms
#endif
template <int C, bool MASK_ONLY, bool HAS_IMAP>
__device__ __forceinline__ void blend_tile_direct(const BlendArgs& a, const WinGeo& g, const Geo& tg, int tx, int ty, int n) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = tx * TW + lane * 4;
    const size_t HW = (size_t)g.H * g.W;
    const int X0 = x0 + g.pad_x;
    int cx = 0;
    bool quad = false;
    if (x0 < g.W) {
        cx = a.cover_x[X0];
        quad = x0 + 3 < g.W && (g.W & 3) == 0 && a.cover_x[X0 + 3] == cx;
        for (int ix = cx >> 8; quad && ix < (cx >> 8) + (cx & 255); ++ix) quad = ((X0 - a.xs[ix]) & 3) == 0;
    }
#pragma unroll 1
    for (int i = 0; i < TH / 8; ++i) {
        const int r = warp * (TH / 8) + i;
        const int y = ty * TH + r;
        if (y >= g.H) break;                      // warp-uniform
        uint32_t nib = 0u;
        if (x0 < g.W) {
            const int cy = a.cover_y[y + g.pad_y];
            if (quad) {
                const size_t o = (size_t)n * HW + (size_t)y * g.W + x0;
                if (MASK_ONLY) {
                    float4 p;
                    nib = blend_quad<C, true, HAS_IMAP>(a, g, n, y, x0, cy, cx, &p, nullptr);
                } else {
                    float4 p, lo[C];
                    nib = blend_quad<C, false, HAS_IMAP>(a, g, n, y, x0, cy, cx, &p, a.blended ? lo : nullptr);
                    if (a.p1) *reinterpret_cast<float4*>(a.p1 + o) = p;
                    if (a.blended) {
#pragma unroll
                        for (int c = 0; c < C; ++c) *reinterpret_cast<float4*>(a.blended + ((size_t)n * C + c) * HW + (size_t)y * g.W + x0) = lo[c];
                    }
                }
                if (a.mask) *reinterpret_cast<uint32_t*>(a.mask + o) = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int x = x0 + j;
                    if (x < g.W) {
                        float p, lo[C];
                        bool fg = blend_pixel<C>(a, g, n, y, x, &p, a.blended ? lo : nullptr);
                        nib |= (fg ? 1u : 0u) << j;
                        const size_t o = (size_t)n * HW + (size_t)y * g.W + x;
                        if (a.mask) a.mask[o] = fg ? 1 : 0;
                        if (a.p1) a.p1[o] = p;
                        if (a.blended) {
#pragma unroll
                            for (int c = 0; c < C; ++c) a.blended[((size_t)n * C + c) * HW + (size_t)y * g.W + x] = lo[c];
                        }
                    }
                }
            }
        }
        if (a.M) {
            uint32_t word = nib << (4 * (lane & 7));
            word |= __shfl_xor_sync(0xffffffffu, word, 1);
            word |= __shfl_xor_sync(0xffffffffu, word, 2);
            word |= __shfl_xor_sync(0xffffffffu, word, 4);
            if ((lane & 7) == 0) a.M[(((size_t)n * tg.nty + ty) * tg.ntx + tx) * TWORDS + r * WPR + (lane >> 3)] = word;
        }
    }
    // rows of the tile below the image are all background
    if (a.M) {
        for (int r = warp * (TH / 8); r < (warp + 1) * (TH / 8); ++r)
            if (ty * TH + r >= g.H && (lane & 7) == 0) a.M[(((size_t)n * tg.nty + ty) * tg.ntx + tx) * TWORDS + r * WPR + (lane >> 3)] = 0u;
    }
}

template <int C, bool MASK_ONLY, bool HAS_IMAP>
__global__ void __launch_bounds__(256, BLEND_MIN_CTAS)
k_blend(BlendArgs a, WinGeo g, Geo tg) {
    blend_tile_direct<C, MASK_ONLY, HAS_IMAP>(a, g, tg, blockIdx.x, blockIdx.y, blockIdx.z);
}

#ifndef CS_EMU
// out of line, so that the staged kernel's register allocation is its own
template <int C, bool HAS_IMAP>
__device__ __noinline__ void blend_tile_direct_call(const BlendArgs& a, const WinGeo& g, const Geo& tg, int tx, int ty, int n) {
    blend_tile_direct<C, true, HAS_IMAP>(a, g, tg, tx, ty, n);
}

// ---------------------------------------------------------------------------
// K2 with the window logits staged by the TMA
// ---------------------------------------------------------------------------
// The direct kernel above keeps every byte it has in flight in registers (2 x C float4 per thread, 80 registers, 3 CTAs per
// SM) and reaches ~50 % of the HBM peak.  Here the copy engine does the moving: the logits are one 2-D tensor
// [window·class·row][column] (a window plane is R contiguous rows of R floats), a tensor map describes it, and for every
// strip of BT_ROWS rows of the CTA's 64 x 128 output tile one warp issues one `cp.async.bulk.tensor.2d` per (covering
// window, class) — a BT_ROWS x 128 box, 4 KB — into shared memory, completion counted in bytes on an mbarrier.  One strip
// is in flight while the other is blended (2 stages x up to 4 windows x C classes x 4 KB = 96 KB, two
// CTAs per SM, up to 192 KB per SM on the wire).  Boxes are addressed in window coordinates: columns left or right of the
// window are out of the tensor's bounds and arrive as zeros, rows above or below it belong to a neighbouring plane — neither
// is ever read, because a pixel only reads the windows its cover table lists.  The arithmetic is the direct kernel's
// (quad_add / quad_finish, ascending window index).  Tiles the staging does not fit (more than BT_SLOTS windows over the
// tile, covering sets that change inside a 4-pixel group) take the direct path.
constexpr int BT_ROWS = 8, BT_STAGES = 2, BT_SLOTS = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

template <int C, bool HAS_IMAP>
__global__ void __launch_bounds__(256, 2)
k_blend_tma(const __grid_constant__ CUtensorMap map, BlendArgs a, WinGeo g, Geo tg) {
    extern __shared__ __align__(128) unsigned char bt_smem[];
    __shared__ uint64_t full[BT_STAGES];
    float* stage = reinterpret_cast<float*>(bt_smem);          // [BT_STAGES][BT_SLOTS][C][BT_ROWS][TW]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx = blockIdx.x, ty = blockIdx.y, n = blockIdx.z;
    const int x0 = tx * TW + lane * 4, X0 = x0 + g.pad_x;
    const int rows_valid = min(TH, g.H - ty * TH);
    const int Xt0 = tx * TW + g.pad_x, Xt1 = min(tx * TW + TW, g.W) - 1 + g.pad_x;
    const int Yt0 = ty * TH + g.pad_y, Yt1 = Yt0 + rows_valid - 1;
    const int ix_lo = a.cover_x[Xt0] >> 8, ix_hi = (a.cover_x[Xt1] >> 8) + (a.cover_x[Xt1] & 255) - 1;
    const int iy_lo = a.cover_y[Yt0] >> 8, iy_hi = (a.cover_y[Yt1] >> 8) + (a.cover_y[Yt1] & 255) - 1;
    const int WX = ix_hi - ix_lo + 1, WY = iy_hi - iy_lo + 1;
    int cx = 0;
    bool fits = WX * WY <= BT_SLOTS && (g.W & 3) == 0;
    if (x0 < g.W) { cx = a.cover_x[X0]; fits = fits && a.cover_x[min(X0 + 3, Xt1)] == cx; }
    if (!__syncthreads_and(fits)) {                          // block-uniform
        blend_tile_direct_call<C, HAS_IMAP>(a, g, tg, tx, ty, n);
        return;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < BT_STAGES; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nst = (rows_valid + BT_ROWS - 1) / BT_ROWS;
    const size_t n_win = (size_t)g.ny * g.nx;
    constexpr int BOX = BT_ROWS * TW;                         // floats per box
    // warp 0: the boxes of strip s into buffer `buf`
    auto issue = [&](int s, int buf) {
        const int Ys0 = Yt0 + s * BT_ROWS, Ys1 = min(Ys0 + BT_ROWS - 1, Yt1);
        const int sy_lo = a.cover_y[Ys0] >> 8, sy_hi = (a.cover_y[Ys1] >> 8) + (a.cover_y[Ys1] & 255) - 1;
        const int nbox = (sy_hi - sy_lo + 1) * WX * C;
        if (lane == 0) mbar_expect_tx(&full[buf], (uint32_t)nbox * BOX * 4u);
        __syncwarp();
        for (int b = lane; b < nbox; b += 32) {
            const int c = b % C, q = b / C;
            const int ix = ix_lo + q % WX, iy = sy_lo + q / WX;
            const int slot = (iy - iy_lo) * WX + (ix - ix_lo);
            const long long row = (((long long)n * (long long)n_win + (long long)iy * g.nx + ix) * C + c) * g.R + (Ys0 - a.ys[iy]);
            tma_load_2d(stage + ((size_t)(buf * BT_SLOTS + slot) * C + c) * BOX, &map, Xt0 - a.xs[ix], (int)row, &full[buf]);
        }
    };
    if (warp == 0) {
        issue(0, 0);
        if (nst > 1) issue(1, 1);
    }
    const size_t HW = (size_t)g.H * g.W;
    for (int s = 0; s < nst; ++s) {
        const int buf = s % BT_STAGES;
        mbar_wait(&full[buf], (uint32_t)(s / BT_STAGES) & 1u);
        const int r = s * BT_ROWS + warp;                     // one strip row per warp
        const int y = ty * TH + r;
        uint32_t nib = 0u;
        if (y < g.H && x0 < g.W) {
            const int cy = a.cover_y[y + g.pad_y];
            float4 acc[C];
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 cnt = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int iy = cy >> 8; iy < (cy >> 8) + (cy & 255); ++iy) {
                for (int ix = cx >> 8; ix < (cx >> 8) + (cx & 255); ++ix) {
                    const int slot = (iy - iy_lo) * WX + (ix - ix_lo);
                    const float* p = stage + (size_t)(buf * BT_SLOTS + slot) * C * BOX + warp * TW + lane * 4;
                    float4 v[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) v[c] = *reinterpret_cast<const float4*>(p + c * BOX);
                    float4 wg = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (HAS_IMAP) {
                        const int wy = y + g.pad_y - a.ys[iy], wx = X0 - a.xs[ix];
                        const float* ip = a.imap + wy * g.R + wx;
                        wg = (wx & 3) == 0 ? *reinterpret_cast<const float4*>(ip) : make_float4(ip[0], ip[1], ip[2], ip[3]);
                    }
                    quad_add<C>(acc, cnt, wg, v);
                }
            }
            float4 pdummy;
            nib = quad_finish<C, true>(acc, cnt, a.threshold, &pdummy, nullptr);
            if (a.mask) *reinterpret_cast<uint32_t*>(a.mask + (size_t)n * HW + (size_t)y * g.W + x0) =
                (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
        }
        if (a.M && y < g.H) {
            uint32_t word = nib << (4 * (lane & 7));
            word |= __shfl_xor_sync(0xffffffffu, word, 1);
            word |= __shfl_xor_sync(0xffffffffu, word, 2);
            word |= __shfl_xor_sync(0xffffffffu, word, 4);
            if ((lane & 7) == 0) a.M[(((size_t)n * tg.nty + ty) * tg.ntx + tx) * TWORDS + r * WPR + (lane >> 3)] = word;
        }
        __syncthreads();                                       // everybody is done with this buffer
        if (warp == 0 && s + BT_STAGES < nst) issue(s + BT_STAGES, buf);
    }
    // rows of the tile below the image are all background
    if (a.M) {
        for (int r = rows_valid + (int)threadIdx.x / WPR; r < TH; r += 256 / WPR)
            a.M[(((size_t)n * tg.nty + ty) * tg.ntx + tx) * TWORDS + r * WPR + (threadIdx.x % WPR)] = 0u;
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// → CS_OK and *launched = true when the TMA kernel was launched; *launched = false: the caller uses the direct kernel.
template <int C>
static int launch_blend_tma(const BlendArgs& a, const WinGeo& g, const Geo& tg, cudaStream_t st, bool* launched) {
    *launched = false;
    const unsigned long long rows = (unsigned long long)g.N * g.ny * g.nx * C * g.R;
    if (g.R < TW || rows >= 0x7fffffffULL || (reinterpret_cast<uintptr_t>(a.win) & 15)) return CS_OK;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return CS_OK;
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)g.R, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)g.R * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TW, (cuuint32_t)BT_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a.win), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return CS_OK;
    constexpr int smem = BT_STAGES * BT_SLOTS * C * BT_ROWS * TW * (int)sizeof(float);
    dim3 grid(tg.ntx, tg.nty, g.N);
    if (a.imap) {
        constexpr auto k_blend_tma_mask_imap = k_blend_tma<C, true>;
        static bool attr = false;
        if (!attr) { CS_CUDA(cudaFuncSetAttribute(k_blend_tma_mask_imap, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
        CS_LAUNCH(k_blend_tma_mask_imap, grid, 256, smem, st, map, a, g, tg);
    } else {
        constexpr auto k_blend_tma_mask = k_blend_tma<C, false>;
        static bool attr = false;
        if (!attr) { CS_CUDA(cudaFuncSetAttribute(k_blend_tma_mask, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
        CS_LAUNCH(k_blend_tma_mask, grid, 256, smem, st, map, a, g, tg);
    }
    CS_LAUNCH_CHECK();
    *launched = true;
    return CS_OK;
}
#endif  // !CS_EMU

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int flat_blocks(size_t work, int threads, int cap = 148 * 32) {
    size_t b = (work + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > (size_t)cap) b = cap;
    return (int)b;
}

int u8_max(const uint8_t* img, int N, size_t bytes_per_image, uint32_t* maxv, cudaStream_t st) {
    CS_CHECK_ARG(img && maxv && N > 0, "null pointer or empty batch");
    CS_CHECK_ARG(N <= 65535, "batch too large");
    CS_CUDA(cudaMemsetAsync(maxv, 0, (size_t)N * sizeof(uint32_t), st));
    // ~6 CTAs per SM over the whole batch, each long enough to amortise its start (148 CTAs per image meant 5 loads per thread at 64 images)
    dim3 grid(flat_blocks(bytes_per_image / 16 + 1, 256, std::max(2, (148 * 6 + N - 1) / N)), N);
    CS_LAUNCH(k_u8_max, grid, 256, 0, st, img, bytes_per_image, maxv);
    CS_LAUNCH_CHECK();
    return CS_OK;
}
int scale_lut_from_max(const uint32_t* maxv, int N, float* lut, cudaStream_t st) {
    CS_CHECK_ARG(maxv && lut && N > 0, "null pointer or empty batch");
    CS_LAUNCH(k_scale_lut, ceil_div(N * 256, 256), 256, 0, st, maxv, lut, N);
    CS_LAUNCH_CHECK();
    return CS_OK;
}
int scale_lut_u8(const uint8_t* img, int N, size_t bytes_per_image, uint32_t* max_scratch, float* lut, cudaStream_t st) {
    CS_CHECK_ARG(lut, "null pointer");
    int rc = u8_max(img, N, bytes_per_image, max_scratch, st);
    if (rc != CS_OK) return rc;
    return scale_lut_from_max(max_scratch, N, lut, st);
}

int u8hwc_to_f32chw(const uint8_t* img, int N, int H, int W, const float* lut, float* out, cudaStream_t st) {
    CS_CHECK_ARG(img && lut && out && N > 0 && H > 0 && W > 0, "null pointer or empty image");
    CS_LAUNCH(k_u8hwc_to_f32chw, flat_blocks((size_t)N * H * W, 256), 256, 0, st, img, N, H, W, lut, out);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

static int check_win(const WinGeo& g) {
    if (g.N <= 0 || g.H <= 0 || g.W <= 0 || g.ny <= 0 || g.nx <= 0) { set_error("empty window geometry"); return CS_ERR_ARG; }
    if (g.R <= 0 || (g.R & 3)) { set_error("roi size must be a positive multiple of 4 (got %d)", g.R); return CS_ERR_ARG; }
    if (g.pad_y < 0 || g.pad_x < 0) { set_error("negative padding"); return CS_ERR_ARG; }
    return CS_OK;
}

int gather_windows(const void* src, int from_u8, const float* lut, WinGeo g, const int* ys, const int* xs,
                   int win_begin, int win_count, float* batch, cudaStream_t st) {
    CS_CHECK_ARG(src && ys && xs && batch, "null pointer");
    CS_CHECK_ARG(!from_u8 || lut, "u8 source needs a scale LUT");
    int rc = check_win(g);
    if (rc != CS_OK) return rc;
    CS_CHECK_ARG(win_begin >= 0 && win_count >= 0 && (long long)win_begin + win_count <= (long long)g.N * g.ny * g.nx, "window range out of bounds");
    if (win_count == 0) return CS_OK;
    size_t work = (size_t)win_count * g.R * (g.R / 4);
    if (from_u8) CS_LAUNCH(k_gather<true>, flat_blocks(work, 256), 256, 0, st, src, lut, g, ys, xs, win_begin, win_count, batch);
    else         CS_LAUNCH(k_gather<false>, flat_blocks(work, 256), 256, 0, st, src, lut, g, ys, xs, win_begin, win_count, batch);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

int blend_softmax_threshold(BlendArgs a, WinGeo g, int C, cudaStream_t st) {
    CS_CHECK_ARG(a.win && a.ys && a.xs && a.cover_y && a.cover_x, "null pointer");
    CS_CHECK_ARG(a.M || a.mask || a.p1 || a.blended, "no output requested");
    int rc = check_win(g);
    if (rc != CS_OK) return rc;
    CS_CHECK_ARG(g.N <= 65535, "batch too large");
    Geo tg = make_geo(g.N, g.H, g.W);
    dim3 grid(tg.ntx, tg.nty, g.N);
    // the mask-only kernels skip the softmax division (exact for threshold 0.5)
    const bool mask_only = !a.p1 && !a.blended && a.threshold == 0.5f;
#ifndef CS_EMU
    // The TMA-staged kernel is opt-in (CELLSEG_BLEND_TMA=1): measured on B200 (64 x 1024², profiles/r02_ncu_blend_tma_vs_direct.txt)
    // it moves the logits at 3.09 TB/s against 3.39 TB/s for the direct kernel — both spend their time in the softmax
    // arithmetic (~100 instructions per pixel, 47-58 % issue utilisation at 37-42 % DRAM utilisation), and the staged kernel
    // runs it with 16 instead of 24 warps per SM.
    const char* env_tma = getenv("CELLSEG_BLEND_TMA");         // read per call: tests and A/B runs flip it in-process
    const bool use_tma = env_tma != nullptr && env_tma[0] == '1';
    if (mask_only && use_tma) {
        bool launched = false;
        rc = C == 2 ? launch_blend_tma<2>(a, g, tg, st, &launched) : C == 3 ? launch_blend_tma<3>(a, g, tg, st, &launched)
           : C == 4 ? launch_blend_tma<4>(a, g, tg, st, &launched) : CS_OK;
        if (rc != CS_OK || launched) return rc;
    }
#endif
#define CS_BLEND_CASE(CC)                                                                        \
    case CC:                                                                                     \
        if (mask_only && !a.imap) { constexpr auto k_blend_mask = k_blend<CC, true, false>; CS_LAUNCH(k_blend_mask, grid, 256, 0, st, a, g, tg); } \
        else if (mask_only) { constexpr auto k_blend_mask_imap = k_blend<CC, true, true>; CS_LAUNCH(k_blend_mask_imap, grid, 256, 0, st, a, g, tg); } \
        else if (!a.imap) { constexpr auto k_blend_full = k_blend<CC, false, false>; CS_LAUNCH(k_blend_full, grid, 256, 0, st, a, g, tg); }         \
        else { constexpr auto k_blend_full_imap = k_blend<CC, false, true>; CS_LAUNCH(k_blend_full_imap, grid, 256, 0, st, a, g, tg); }             \
        break;
    switch (C) {
        CS_BLEND_CASE(2)
        CS_BLEND_CASE(3)
        CS_BLEND_CASE(4)
        default: set_error("blend_softmax_threshold: %d classes unsupported (2..4)", C); return CS_ERR_ARG;
    }
#undef CS_BLEND_CASE
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace cs
