// Predictor glue, channels-last edition (the predictor is monai's UNet, baseline/predict.py:45-52).
//
// The convolutions with enough channels to be a dense contraction stay on cuDNN (out of scope by the north
// star) and run fastest on NHWC tensors (no nchw<->nhwc conversion kernels around the tensor-core convolutions).
// What cuDNN is bad at on this network are the 3-channel ends, which are HBM-bound stencils rather than
// contractions (0.26 / 0.58 ms per 100 windows against ~0.03 ms of traffic): they are written here, together with
// the norm blocks between the convolutions, so that every activation tensor is produced channels-last:
//
//   k_stem          the two stride-2 3->16 convolutions of the first ResidualUnit (unit0 and the residual branch) from
//                   the NCHW window batch in one pass: both NHWC outputs + the InstanceNorm partial statistics of unit0
//   k_nhwc_stats    per-(sample, channel) partial statistics (chunk mean, chunk M2) of an NHWC tensor cuDNN produced
//   k_nhwc_norm_prelu   InstanceNorm2d(affine=False) -> PReLU (-> + residual + residual-conv bias) from those partials
//   k_head_convt    ConvTranspose2d(32 -> 3, k3, s2, p1, op1) reading the two 16-channel NHWC halves of the skip
//                   concatenation (no torch.cat), planar output + partial statistics
//   k_head_final    InstanceNorm -> PReLU -> Conv2d(3 -> 3, k3) + bias + identity residual, planar logits out
//
// Statistics are two-level and deterministic: each chunk yields (mean, M2 about its own mean); consumers merge the
// chunks of a sample with Chan's formula in double.  No atomics, no dependence on scheduling.
// The library is compiled with --fmad=false (the blend must round products and sums separately); the convolution
// loops below ask for fused multiply-adds explicitly.
#include "internal.cuh"
#include <cmath>

namespace cs {

constexpr int GL_T = 256;      // threads per CTA of every kernel in this file
constexpr int GL_VPT = 8;      // float4 per thread of an NHWC chunk → 2048 float4 (32 KB) per CTA

// ---------------------------------------------------------------------------
// per-channel CTA reduction for NHWC data
// ---------------------------------------------------------------------------
// Every thread holds a float4 = 4 consecutive channels of channel group (threadIdx.x % G), G = C/4 ∈ {1,2,4,...,64}.
// Afterwards sc.out[c] is the sum over the CTA for channel c (valid for all threads).
struct ChanScratch {
    float4 part[GL_T / 32][32];
    float out[256];
};

__device__ __forceinline__ void cta_channel_sum(float4 s, int G, ChanScratch& sc) {
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
    for (int c = threadIdx.x; c < C; c += GL_T) {
        const int g = c >> 2, k = c & 3;
        float acc = 0.f;
        for (int w = 0; w < GL_T / 32; ++w) {            // fixed order → deterministic
            if (G > 32 && ((w * 32) % G) != (g & ~31)) continue;
            const float4 p = sc.part[w][g & 31];
            acc += k == 0 ? p.x : k == 1 ? p.y : k == 2 ? p.z : p.w;
        }
        sc.out[c] = acc;
    }
    __syncthreads();
}

__device__ __forceinline__ float prelu_f(float v, float a) { return v > 0.f ? v : a * v; }

// Merge the per-chunk (mean, M2) of one (sample, channel): chunk k holds cnt(k) values.  Returns mean and 1/sqrt(var+eps).
__device__ __forceinline__ void merge_partials(const float2* p, int stride, int nstat, int per_chunk, int total, float eps,
                                               float& mean_out, float& rstd_out) {
    double tot = 0.0;
    for (int k = 0; k < nstat; ++k) {
        const int cnt = min(per_chunk, total - k * per_chunk);
        tot += (double)cnt * (double)p[(size_t)k * stride].x;
    }
    const double mean = tot / (double)total;
    double m2 = 0.0;
    for (int k = 0; k < nstat; ++k) {
        const int cnt = min(per_chunk, total - k * per_chunk);
        const float2 q = p[(size_t)k * stride];
        const double d = (double)q.x - mean;
        m2 += (double)q.y + (double)cnt * d * d;
    }
    mean_out = (float)mean;
    rstd_out = (float)(1.0 / sqrt(m2 / (double)total + (double)eps));
}

// ---------------------------------------------------------------------------
// NHWC statistics + norm/activation/residual
// ---------------------------------------------------------------------------
// grid (chunks, N).  Chunk = P pixels (P*G <= GL_T*GL_VPT float4); partials[n][chunk][c] = (mean, M2) of the chunk.
__global__ void __launch_bounds__(GL_T) k_nhwc_stats(const float* __restrict__ x, int HW, int C, int P, float2* __restrict__ partials) {
    __shared__ ChanScratch sc;
    const int G = C >> 2, t = threadIdx.x, chunk = blockIdx.x, n = blockIdx.y;
    const int p0 = chunk * P, np = min(P, HW - p0), f4n = np * G;
    const float4* xp = reinterpret_cast<const float4*>(x) + ((size_t)n * HW + p0) * G;
    float4 v[GL_VPT];
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < GL_VPT; ++i) {
        const int j = t + i * GL_T;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < f4n) { v[i] = __ldg(xp + j); s.x += v[i].x; s.y += v[i].y; s.z += v[i].z; s.w += v[i].w; }
    }
    cta_channel_sum(s, G, sc);
    const float inv = 1.0f / (float)np;
    const int g4 = (t % G) * 4;
    const float4 m = make_float4(sc.out[g4] * inv, sc.out[g4 + 1] * inv, sc.out[g4 + 2] * inv, sc.out[g4 + 3] * inv);
    const float my_mean = t < C ? sc.out[t] * inv : 0.f;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < GL_VPT; ++i) {
        const int j = t + i * GL_T;
        if (j < f4n) {
            const float a = v[i].x - m.x, b = v[i].y - m.y, c = v[i].z - m.z, d = v[i].w - m.w;
            q.x += a * a; q.y += b * b; q.z += c * c; q.w += d * d;
        }
    }
    cta_channel_sum(q, G, sc);
    if (t < C) partials[((size_t)n * gridDim.x + chunk) * C + t] = make_float2(my_mean, sc.out[t]);
}

// grid (chunks of P pixels, N).  Statistics: nstat chunks of pstat pixels per sample.  y may alias x.
__global__ void __launch_bounds__(GL_T) k_nhwc_norm_prelu(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ y,
                                                          const float* __restrict__ prelu_w, const float* __restrict__ res_bias,
                                                          int HW, int C, int P, const float2* __restrict__ partials, int nstat, int pstat, float eps) {
    __shared__ float s_mean[256], s_rstd[256];
    const int G = C >> 2, t = threadIdx.x, chunk = blockIdx.x, n = blockIdx.y;
    if (t < C) merge_partials(partials + (size_t)n * nstat * C + t, C, nstat, pstat, HW, eps, s_mean[t], s_rstd[t]);
    __syncthreads();
    const int p0 = chunk * P, np = min(P, HW - p0), f4n = np * G;
    const size_t base = ((size_t)n * HW + p0) * G;
    const float4* xp = reinterpret_cast<const float4*>(x) + base;
    const float4* rp = res ? reinterpret_cast<const float4*>(res) + base : nullptr;
    float4* yp = reinterpret_cast<float4*>(y) + base;
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
            const float4 v = xp[j];
            float4 o;
            o.x = prelu_f((v.x - m.x) * r.x, a); o.y = prelu_f((v.y - m.y) * r.y, a);
            o.z = prelu_f((v.z - m.z) * r.z, a); o.w = prelu_f((v.w - m.w) * r.w, a);
            if (rp) { const float4 q = rp[j]; o.x += q.x + rb.x; o.y += q.y + rb.y; o.z += q.z + rb.z; o.w += q.w + rb.w; }
            yp[j] = o;
        }
    }
}

// ---------------------------------------------------------------------------
// stem: Conv2d(3 -> 16, k3, s2, p1) twice (unit0 and the residual branch), NCHW windows in, NHWC out
// ---------------------------------------------------------------------------
constexpr int STEM_CO = 16;
constexpr int STEM_ROW = 2 * STEM_CO + 4;    // staging row stride in floats (padded: conflict-free float4 stores)

// grid (ceil(OH*OW / 256), N).  A/R: [N][OH*OW][16]; partials (of A): [N][chunks][16].
__global__ void __launch_bounds__(GL_T) k_stem(const float* __restrict__ win, int H, int W, const float* __restrict__ w0, const float* __restrict__ wr,
                                               float* __restrict__ A, float* __restrict__ R, float2* __restrict__ partials) {
    __shared__ __align__(16) float ws[27][2 * STEM_CO];
    __shared__ __align__(16) float stage[GL_T * STEM_ROW];
    __shared__ ChanScratch sc;
    const int t = threadIdx.x, n = blockIdx.y;
    const int OH = H >> 1, OW = W >> 1, OHW = OH * OW;
    for (int i = t; i < 27 * 2 * STEM_CO; i += GL_T) {
        const int tap = i / (2 * STEM_CO), co = i % (2 * STEM_CO);          // tap = ci*9 + ky*3 + kx
        ws[tap][co] = co < STEM_CO ? w0[co * 27 + tap] : wr[(co - STEM_CO) * 27 + tap];
    }
    __syncthreads();
    const int p0 = blockIdx.x * GL_T, p = p0 + t;
    const int np = min(GL_T, OHW - p0);
    float acc[2 * STEM_CO];
#pragma unroll
    for (int i = 0; i < 2 * STEM_CO; ++i) acc[i] = 0.f;
    if (p < OHW) {
        const int oy = p / OW, ox = p % OW;
        const float* src = win + (size_t)n * 3 * H * W;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int iy = 2 * oy - 1 + ky;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int ix = 2 * ox - 1 + kx;
                    float v = 0.f;
                    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(src + ((size_t)ci * H + iy) * W + ix);
                    const float4* wp = reinterpret_cast<const float4*>(ws[ci * 9 + ky * 3 + kx]);
#pragma unroll
                    for (int q = 0; q < 2 * STEM_CO / 4; ++q) {
                        const float4 w4 = wp[q];
                        acc[4 * q + 0] = fmaf(v, w4.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
                    }
                }
            }
        }
    }
    float4* st4 = reinterpret_cast<float4*>(stage + t * STEM_ROW);
#pragma unroll
    for (int q = 0; q < 2 * STEM_CO / 4; ++q) st4[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    __syncthreads();
    // coalesced NHWC stores of the CTA's 256-pixel chunk + the chunk statistics of A
    constexpr int G = STEM_CO / 4;
    const int f4n = np * G;
    float4* Ap = reinterpret_cast<float4*>(A) + ((size_t)n * OHW + p0) * G;
    float4* Rp = reinterpret_cast<float4*>(R) + ((size_t)n * OHW + p0) * G;
    float4 v[GL_T * G / GL_T];
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

// ---------------------------------------------------------------------------
// head, part 1: ConvTranspose2d(32 -> 3, k3, s2, p1, output_padding 1) on cat[xa, xb] (16 channels each, NHWC)
// ---------------------------------------------------------------------------
// out[2iy+dy][2ix+dx] gets:  (0,0): in[iy][ix]·w11      (0,1): in[iy][ix]·w12 + in[iy][ix+1]·w10
//   (1,0): in[iy][ix]·w21 + in[iy+1][ix]·w01      (1,1): in[iy][ix]·w22 + in[iy][ix+1]·w20 + in[iy+1][ix]·w02 + in[iy+1][ix+1]·w00
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

// grid (ceil(Hin*Win / 256), N).  z: [N][3][2Hin][2Win]; partials: [N][chunks][3] over 4*(pixels of the chunk) values.
__global__ void __launch_bounds__(GL_T) k_head_convt(const float* __restrict__ xa, const float* __restrict__ xb, int Hin, int Win,
                                                     const float* __restrict__ wt, float* __restrict__ z, float2* __restrict__ partials) {
    __shared__ __align__(16) float ws[32][28];
    __shared__ float red[27];
    const int t = threadIdx.x, n = blockIdx.y;
    // per input channel the 27 weights in order of use (see the table above): [tap][co]
    for (int i = t; i < 32 * 27; i += GL_T) {
        const int ci = i / 27, k = i % 27, slot = k / 3, co = k % 3;
        const int kyx[9][2] = {{1, 1}, {1, 2}, {2, 1}, {2, 2}, {1, 0}, {2, 0}, {0, 1}, {0, 2}, {0, 0}};
        ws[ci][k] = wt[((ci * 3 + co) * 3 + kyx[slot][0]) * 3 + kyx[slot][1]];
    }
    if (t < 32) ws[t][27] = 0.f;
    __syncthreads();
    const int HWin = Hin * Win, p0 = blockIdx.x * GL_T, p = p0 + t;
    const bool valid = p < HWin;
    const int iy = valid ? p / Win : 0, ix = valid ? p % Win : 0;
    const bool has_r = valid && ix + 1 < Win, has_d = valid && iy + 1 < Hin;
    float o00[3] = {0.f, 0.f, 0.f}, o01[3] = {0.f, 0.f, 0.f}, o10[3] = {0.f, 0.f, 0.f}, o11[3] = {0.f, 0.f, 0.f};
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
#pragma unroll 2
        for (int c4 = 0; c4 < 8; ++c4) {
            const float4* src = reinterpret_cast<const float4*>(c4 < 4 ? xa : xb) + ((size_t)n * HWin + p) * 4 + (c4 & 3);
            const float4 v00 = __ldg(src);
            const float4 v01 = has_r ? __ldg(src + 4) : zero4;
            const float4 v10 = has_d ? __ldg(src + (size_t)Win * 4) : zero4;
            const float4 v11 = (has_r && has_d) ? __ldg(src + (size_t)Win * 4 + 4) : zero4;
            const float a00[4] = {v00.x, v00.y, v00.z, v00.w}, a01[4] = {v01.x, v01.y, v01.z, v01.w};
            const float a10[4] = {v10.x, v10.y, v10.z, v10.w}, a11[4] = {v11.x, v11.y, v11.z, v11.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4* wp = reinterpret_cast<const float4*>(ws[c4 * 4 + e]);
                const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4], w5 = wp[5], w6 = wp[6];
                const float b00 = a00[e], b01 = a01[e], b10 = a10[e], b11 = a11[e];
                o00[0] = fmaf(b00, w0.x, o00[0]); o00[1] = fmaf(b00, w0.y, o00[1]); o00[2] = fmaf(b00, w0.z, o00[2]);
                o01[0] = fmaf(b00, w0.w, o01[0]); o01[1] = fmaf(b00, w1.x, o01[1]); o01[2] = fmaf(b00, w1.y, o01[2]);
                o10[0] = fmaf(b00, w1.z, o10[0]); o10[1] = fmaf(b00, w1.w, o10[1]); o10[2] = fmaf(b00, w2.x, o10[2]);
                o11[0] = fmaf(b00, w2.y, o11[0]); o11[1] = fmaf(b00, w2.z, o11[1]); o11[2] = fmaf(b00, w2.w, o11[2]);
                o01[0] = fmaf(b01, w3.x, o01[0]); o01[1] = fmaf(b01, w3.y, o01[1]); o01[2] = fmaf(b01, w3.z, o01[2]);
                o11[0] = fmaf(b01, w3.w, o11[0]); o11[1] = fmaf(b01, w4.x, o11[1]); o11[2] = fmaf(b01, w4.y, o11[2]);
                o10[0] = fmaf(b10, w4.z, o10[0]); o10[1] = fmaf(b10, w4.w, o10[1]); o10[2] = fmaf(b10, w5.x, o10[2]);
                o11[0] = fmaf(b10, w5.y, o11[0]); o11[1] = fmaf(b10, w5.z, o11[1]); o11[2] = fmaf(b10, w5.w, o11[2]);
                o11[0] = fmaf(b11, w6.x, o11[0]); o11[1] = fmaf(b11, w6.y, o11[1]); o11[2] = fmaf(b11, w6.z, o11[2]);
            }
        }
        const int OW = 2 * Win;
#pragma unroll
        for (int co = 0; co < 3; ++co) {
            float* zp = z + (((size_t)n * 3 + co) * (2 * Hin) + 2 * iy) * OW + 2 * ix;
            *reinterpret_cast<float2*>(zp) = make_float2(o00[co], o01[co]);
            *reinterpret_cast<float2*>(zp + OW) = make_float2(o10[co], o11[co]);
        }
    }
    const int np = min(GL_T, HWin - p0);
    float s0 = valid ? (o00[0] + o01[0]) + (o10[0] + o11[0]) : 0.f;
    float s1 = valid ? (o00[1] + o01[1]) + (o10[1] + o11[1]) : 0.f;
    float s2 = valid ? (o00[2] + o01[2]) + (o10[2] + o11[2]) : 0.f;
    block_sum3(s0, s1, s2, red);
    const float inv = 1.0f / (4.0f * (float)np);
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
        partials[((size_t)n * gridDim.x + blockIdx.x) * 3 + t] = make_float2(mm, qq);
    }
}

// ---------------------------------------------------------------------------
// head, part 2: InstanceNorm -> PReLU -> Conv2d(3 -> 3, k3, p1) + bias + identity residual
// ---------------------------------------------------------------------------
constexpr int HF_TH = 16, HF_TW = 128;
constexpr int HF_SW = HF_TW + 2;             // tile row incl. halo

// grid (ceil(W/128), ceil(H/16), N).  z, out: [N][3][H][W].  Statistics: nstat chunks of `pstat` values per (n, channel).
__global__ void __launch_bounds__(GL_T) k_head_final(const float* __restrict__ z, int H, int W, const float2* __restrict__ partials, int nstat, int pstat,
                                                     const float* __restrict__ prelu_w, const float* __restrict__ w3, const float* __restrict__ b3,
                                                     float eps, float* __restrict__ out) {
    __shared__ float tile[3][HF_TH + 2][HF_SW];
    __shared__ float s_w[81], s_mean[3], s_rstd[3];
    const int t = threadIdx.x, n = blockIdx.z;
    const int y0 = blockIdx.y * HF_TH, x0 = blockIdx.x * HF_TW;
    if (t < 3) merge_partials(partials + (size_t)n * nstat * 3 + t, 3, nstat, pstat, H * W, eps, s_mean[t], s_rstd[t]);
    if (t >= 32 && t < 32 + 81) s_w[t - 32] = w3[t - 32];
    __syncthreads();
    const float a = prelu_w ? *prelu_w : 1.0f;
    const float* zn = z + (size_t)n * 3 * H * W;
    for (int i = t; i < 3 * (HF_TH + 2) * HF_SW; i += GL_T) {
        const int ch = i / ((HF_TH + 2) * HF_SW), rem = i % ((HF_TH + 2) * HF_SW);
        const int r = rem / HF_SW, c = rem % HF_SW;
        const int gy = y0 + r - 1, gx = x0 + c - 1;
        float v = 0.f;                       // zero padding applies to the activated tensor
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = prelu_f((__ldg(zn + ((size_t)ch * H + gy) * W + gx) - s_mean[ch]) * s_rstd[ch], a);
        tile[ch][r][c] = v;
    }
    __syncthreads();
    const int lane = t & 31, wid = t >> 5;
    const int r0 = 2 * wid;                  // this warp's two tile rows
    float acc[4][2][3];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int co = 0; co < 3; ++co) acc[q][r][co] = b3 ? b3[co] : 0.f;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
        float w[3][9];
#pragma unroll
        for (int co = 0; co < 3; ++co)
#pragma unroll
            for (int k = 0; k < 9; ++k) w[co][k] = s_w[(co * 3 + ci) * 9 + k];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;     // tile column; halo shifts by one
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
                        for (int kx = 0; kx < 3; ++kx) s = fmaf(w[co][ky * 3 + kx], v[r + ky][kx], s);
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
    CS_LAUNCH(k_nhwc_stats, dim3(ceil_div(HW, P), N), GL_T, 0, st, x, HW, C, P, reinterpret_cast<float2*>(partials));
    CS_LAUNCH_CHECK();
    return CS_OK;
}

int nhwc_norm_prelu(const float* x, const float* res, float* y, const float* prelu_w, const float* res_bias, int N, int HW, int C,
                    const float* partials, int stat_chunk_pixels, float eps, cudaStream_t st) {
    CS_CHECK_ARG(x && y && partials, "null pointer");
    CS_CHECK_ARG(N > 0 && N <= 65535 && HW > 0 && stat_chunk_pixels > 0, "bad N / HW / chunk");
    CS_CHECK_ARG(nhwc_channels_ok(C), "channel count must be a power of two in [4, 256]");
    CS_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)res & 15) == 0 && ((uintptr_t)partials & 7) == 0, "misaligned pointer");
    const int P = nhwc_chunk_pixels(C);
    CS_LAUNCH(k_nhwc_norm_prelu, dim3(ceil_div(HW, P), N), GL_T, 0, st, x, res, y, prelu_w, res_bias, HW, C, P,
              reinterpret_cast<const float2*>(partials), ceil_div(HW, stat_chunk_pixels), stat_chunk_pixels, eps);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

int unet_stem(const float* win, int N, int H, int W, const float* w_unit0, const float* w_res, float* a_out, float* r_out, float* a_partials, cudaStream_t st) {
    CS_CHECK_ARG(win && w_unit0 && w_res && a_out && r_out && a_partials, "null pointer");
    CS_CHECK_ARG(N > 0 && N <= 65535 && H > 0 && W > 0 && (H & 1) == 0 && (W & 1) == 0, "window batch must have even H and W");
    CS_CHECK_ARG(((uintptr_t)a_out & 15) == 0 && ((uintptr_t)r_out & 15) == 0 && ((uintptr_t)a_partials & 7) == 0, "misaligned pointer");
    CS_LAUNCH(k_stem, dim3(ceil_div((H / 2) * (W / 2), GL_T), N), GL_T, 0, st, win, H, W, w_unit0, w_res, a_out, r_out, reinterpret_cast<float2*>(a_partials));
    CS_LAUNCH_CHECK();
    return CS_OK;
}

int unet_head_convt(const float* xa, const float* xb, int N, int Hin, int Win, const float* wt, float* z, float* z_partials, cudaStream_t st) {
    CS_CHECK_ARG(xa && xb && wt && z && z_partials, "null pointer");
    CS_CHECK_ARG(N > 0 && N <= 65535 && Hin > 0 && Win > 0, "bad geometry");
    CS_CHECK_ARG(((uintptr_t)xa & 15) == 0 && ((uintptr_t)xb & 15) == 0 && ((uintptr_t)z & 7) == 0 && ((uintptr_t)z_partials & 7) == 0, "misaligned pointer");
    CS_LAUNCH(k_head_convt, dim3(ceil_div(Hin * Win, GL_T), N), GL_T, 0, st, xa, xb, Hin, Win, wt, z, reinterpret_cast<float2*>(z_partials));
    CS_LAUNCH_CHECK();
    return CS_OK;
}

int unet_head_final(const float* z, const float* z_partials, int N, int H, int W, const float* prelu_w, const float* w3, const float* b3, float eps,
                    float* out, cudaStream_t st) {
    CS_CHECK_ARG(z && z_partials && w3 && out, "null pointer");
    CS_CHECK_ARG(N > 0 && N <= 65535 && H > 0 && W > 0 && (H & 1) == 0 && (W & 1) == 0, "bad geometry");
    CS_CHECK_ARG(ceil_div(H, HF_TH) <= 65535, "image too tall");
    const int pix_in = (H / 2) * (W / 2);
    CS_LAUNCH(k_head_final, dim3(ceil_div(W, HF_TW), ceil_div(H, HF_TH), N), GL_T, 0, st, z, H, W, reinterpret_cast<const float2*>(z_partials),
              ceil_div(pix_in, GL_T), 4 * GL_T, prelu_w, w3, b3, eps, out);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace cs
