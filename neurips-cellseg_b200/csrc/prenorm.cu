// Pre-normalisation of the input image (baseline/predict.py:14-21, 94-104):
//   per channel: 1st / 99th percentile of the NON-ZERO values → rescale_intensity to uint8.
// The per-pixel work is two passes: a per-channel histogram (K8) and a LUT application (K9).  Between them the host
// turns the histogram into numpy's exact linearly-interpolated percentiles and a value → uint8 table in float64
// (256 or 65 536 entries per channel — every distinct input value is mapped exactly as the reference maps it).
// Channel handling follows predict.py:94-99: a 2-D image is replicated to 3 channels, more than 3 channels → first 3.
#include "internal.cuh"

namespace cs {

template <typename T> struct HistCfg;
template <> struct HistCfg<uint8_t> { static constexpr int BINS = 256; };
template <> struct HistCfg<uint16_t> { static constexpr int BINS = 65536; };

// u8: one privatised 3 x 256 histogram PER WARP in shared memory (24 KB per CTA: an eighth of the same-bin collisions of one table per
// CTA — microscopy backgrounds put most pixels into a handful of bins), flushed with one atomic per non-empty bin.  Interleaved RGB is
// read 48 bytes = 16 pixels at a time (three 16-byte loads; byte k of such a group belongs to channel k % 3), gray images 16 pixels per
// load; other layouts and the unaligned tails take the per-pixel path.
constexpr int HIST_T = 256, HIST_W = HIST_T / 32;
__device__ __forceinline__ void hist_word(uint32_t* wh, uint32_t w, int c0, int c1, int c2, int c3) {
    atomicAdd(&wh[c0 * 256 + (w & 0xffu)], 1u);
    atomicAdd(&wh[c1 * 256 + ((w >> 8) & 0xffu)], 1u);
    atomicAdd(&wh[c2 * 256 + ((w >> 16) & 0xffu)], 1u);
    atomicAdd(&wh[c3 * 256 + (w >> 24)], 1u);
}
__global__ void __launch_bounds__(HIST_T) k_hist_u8(const uint8_t* __restrict__ img, size_t npx, int C, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[HIST_W][3 * 256];
    for (int i = threadIdx.x; i < HIST_W * 3 * 256; i += HIST_T) (&sh[0][0])[i] = 0u;
    __syncthreads();
    uint32_t* wh = sh[threadIdx.x >> 5];
    const int nc = C < 3 ? C : 3;
    const size_t tid = (size_t)blockIdx.x * HIST_T + threadIdx.x, stride = (size_t)gridDim.x * HIST_T;
    size_t done = 0;                                   // pixels covered by the vector path
    if ((C == 3 || C == 1) && ((uintptr_t)img & 15) == 0) {
        const uint4* v = reinterpret_cast<const uint4*>(img);
        if (C == 3) {
            const size_t groups = npx / 16;
            for (size_t g = tid; g < groups; g += stride) {
                const uint4 a = __ldg(v + 3 * g), b = __ldg(v + 3 * g + 1), c = __ldg(v + 3 * g + 2);
                hist_word(wh, a.x, 0, 1, 2, 0); hist_word(wh, a.y, 1, 2, 0, 1); hist_word(wh, a.z, 2, 0, 1, 2); hist_word(wh, a.w, 0, 1, 2, 0);
                hist_word(wh, b.x, 1, 2, 0, 1); hist_word(wh, b.y, 2, 0, 1, 2); hist_word(wh, b.z, 0, 1, 2, 0); hist_word(wh, b.w, 1, 2, 0, 1);
                hist_word(wh, c.x, 2, 0, 1, 2); hist_word(wh, c.y, 0, 1, 2, 0); hist_word(wh, c.z, 1, 2, 0, 1); hist_word(wh, c.w, 2, 0, 1, 2);
            }
            done = groups * 16;
        } else {
            const size_t groups = npx / 16;
            for (size_t g = tid; g < groups; g += stride) {
                const uint4 a = __ldg(v + g);
                hist_word(wh, a.x, 0, 0, 0, 0); hist_word(wh, a.y, 0, 0, 0, 0); hist_word(wh, a.z, 0, 0, 0, 0); hist_word(wh, a.w, 0, 0, 0, 0);
            }
            done = groups * 16;
        }
    }
    for (size_t p = done + tid; p < npx; p += stride) {
        const uint8_t* s = img + p * C;
        for (int c = 0; c < nc; ++c) atomicAdd(&wh[c * 256 + s[c]], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nc * 256; i += HIST_T) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < HIST_W; ++w) t += sh[w][i];
        if (t) atomicAdd(&hist[i], t);
    }
}
// u16: 65 536 bins per channel do not fit in shared memory, but microscope cameras deliver 12-14 bits: values below U16_LOW are counted
// in a per-CTA shared-memory table (3 x 8192 x 4 B = 96 KB), the rest go to the global table in L2; runs of equal values inside a
// thread's pixels are merged first.  (With global atomics only the kernel ran at 70 GB/s on a 12-bit image.)
constexpr int U16_LOW = 8192;
__global__ void __launch_bounds__(256) k_hist_u16(const uint16_t* __restrict__ img, size_t npx, int C, uint32_t* __restrict__ hist) {
    CS_DYN_SMEM(smem_raw);
    uint32_t* sh = reinterpret_cast<uint32_t*>(smem_raw);           // [3][U16_LOW]
    const int nc = C < 3 ? C : 3;
    for (int i = threadIdx.x; i < nc * U16_LOW; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    constexpr int PER = 8;
    for (size_t p0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * PER; p0 < npx; p0 += (size_t)gridDim.x * blockDim.x * PER) {
        const size_t lim = npx - p0 < PER ? npx - p0 : PER;
        for (int c = 0; c < nc; ++c) {
            uint32_t cur = img[p0 * C + c], run = 1;
            for (size_t i = 1; i <= lim; ++i) {
                const uint32_t v = i < lim ? img[(p0 + i) * C + c] : 0x10000u;      // sentinel flushes the last run
                if (v == cur) { ++run; continue; }
                if (cur < (uint32_t)U16_LOW) atomicAdd(&sh[c * U16_LOW + cur], run);
                else atomicAdd(&hist[(size_t)c * 65536 + cur], run);
                cur = v; run = 1;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nc * U16_LOW; i += blockDim.x) {
        const uint32_t t = sh[i];
        if (t) atomicAdd(&hist[(size_t)(i / U16_LOW) * 65536 + (i % U16_LOW)], t);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_apply_lut(const T* __restrict__ img, size_t npx, int C, const uint8_t* __restrict__ lut, uint8_t* __restrict__ out) {
    constexpr int BINS = HistCfg<T>::BINS;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npx; p += (size_t)gridDim.x * blockDim.x) {
        const T* s = img + p * C;
        uint8_t* o = out + p * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int sc = C == 1 ? 0 : c;                    // 2-D image: the one channel feeds all three
            o[c] = sc < C ? lut[(size_t)c * BINS + s[sc]] : 0;
        }
    }
}
// Interleaved u8 RGB → u8 RGB: the three 256-entry tables in shared memory, 48 bytes = 16 pixels per thread and step in and out as
// three 16-byte accesses (byte k of a group is channel k % 3).  `groups` = npx / 16; the tail goes through k_apply_lut.
__device__ __forceinline__ uint32_t lut_word(const uint8_t* t, uint32_t w, int c0, int c1, int c2, int c3) {
    return (uint32_t)t[c0 * 256 + (w & 0xffu)] | ((uint32_t)t[c1 * 256 + ((w >> 8) & 0xffu)] << 8) |
           ((uint32_t)t[c2 * 256 + ((w >> 16) & 0xffu)] << 16) | ((uint32_t)t[c3 * 256 + (w >> 24)] << 24);
}
__global__ void __launch_bounds__(256) k_apply_lut_rgb8(const uint4* __restrict__ img, size_t groups, const uint8_t* __restrict__ lut, uint4* __restrict__ out) {
    __shared__ uint8_t t[3 * 256];
    for (int i = threadIdx.x; i < 3 * 256; i += 256) t[i] = lut[i];
    __syncthreads();
    for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < groups; g += (size_t)gridDim.x * 256) {
        const uint4 a = __ldg(img + 3 * g), b = __ldg(img + 3 * g + 1), c = __ldg(img + 3 * g + 2);
        uint4 oa, ob, oc;
        oa.x = lut_word(t, a.x, 0, 1, 2, 0); oa.y = lut_word(t, a.y, 1, 2, 0, 1); oa.z = lut_word(t, a.z, 2, 0, 1, 2); oa.w = lut_word(t, a.w, 0, 1, 2, 0);
        ob.x = lut_word(t, b.x, 1, 2, 0, 1); ob.y = lut_word(t, b.y, 2, 0, 1, 2); ob.z = lut_word(t, b.z, 0, 1, 2, 0); ob.w = lut_word(t, b.w, 1, 2, 0, 1);
        oc.x = lut_word(t, c.x, 2, 0, 1, 2); oc.y = lut_word(t, c.y, 0, 1, 2, 0); oc.z = lut_word(t, c.z, 1, 2, 0, 1); oc.w = lut_word(t, c.w, 2, 0, 1, 2);
        out[3 * g] = oa; out[3 * g + 1] = ob; out[3 * g + 2] = oc;
    }
}

// u8 gray → u8 RGB (predict.py:94-95 replicates the one channel): 16 pixels in as one 16-byte load, 48 bytes out as three.
__global__ void __launch_bounds__(256) k_apply_lut_gray8(const uint4* __restrict__ img, size_t groups, const uint8_t* __restrict__ lut, uint4* __restrict__ out) {
    __shared__ uint8_t t[3 * 256];
    for (int i = threadIdx.x; i < 3 * 256; i += 256) t[i] = lut[i];
    __syncthreads();
    for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < groups; g += (size_t)gridDim.x * 256) {
        const uint4 a = __ldg(img + g);
        const uint32_t in[4] = {a.x, a.y, a.z, a.w};
        uint32_t o[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) o[k] = 0u;
#pragma unroll
        for (int px = 0; px < 16; ++px) {
            const uint32_t v = (in[px >> 2] >> (8 * (px & 3))) & 0xffu;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int byte = 3 * px + c;
                o[byte >> 2] |= (uint32_t)t[c * 256 + v] << (8 * (byte & 3));
            }
        }
        out[3 * g] = make_uint4(o[0], o[1], o[2], o[3]);
        out[3 * g + 1] = make_uint4(o[4], o[5], o[6], o[7]);
        out[3 * g + 2] = make_uint4(o[8], o[9], o[10], o[11]);
    }
}

static int grid_for(size_t work, int threads) {
    size_t b = (work + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > 148 * 8) b = 148 * 8;
    return (int)b;
}

// hist: device uint32[3][bins] (bins = 256 for 1-byte, 65536 for 2-byte elements), zeroed here.
int prenorm_hist(const void* img, int elem_bytes, int H, int W, int C, uint32_t* hist, cudaStream_t st) {
    CS_CHECK_ARG(img && hist && H > 0 && W > 0 && C >= 1, "null pointer or empty image");
    CS_CHECK_ARG(elem_bytes == 1 || elem_bytes == 2, "element size must be 1 (uint8) or 2 (uint16)");
    const size_t npx = (size_t)H * W;
    const size_t bins = elem_bytes == 1 ? 256 : 65536;
    CS_CUDA(cudaMemsetAsync(hist, 0, 3 * bins * sizeof(uint32_t), st));
    if (elem_bytes == 1) CS_LAUNCH(k_hist_u8, grid_for(npx, HIST_T * 16), HIST_T, 0, st, (const uint8_t*)img, npx, C, hist);
    else {
        constexpr int smem = 3 * U16_LOW * (int)sizeof(uint32_t);
#ifndef CS_EMU
        CS_CUDA(cudaFuncSetAttribute(k_hist_u16, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));   // per device, so per call (microseconds)
#endif
        CS_LAUNCH(k_hist_u16, grid_for(npx, 256 * 8 * 4), 256, smem, st, (const uint16_t*)img, npx, C, hist);
    }
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// lut: device uint8[3][bins]; out: device uint8[H][W][3].
int prenorm_apply(const void* img, int elem_bytes, int H, int W, int C, const uint8_t* lut, uint8_t* out, cudaStream_t st) {
    CS_CHECK_ARG(img && lut && out && H > 0 && W > 0 && C >= 1, "null pointer or empty image");
    CS_CHECK_ARG(elem_bytes == 1 || elem_bytes == 2, "element size must be 1 (uint8) or 2 (uint16)");
    const size_t npx = (size_t)H * W;
    if (elem_bytes == 1 && C == 3 && ((((uintptr_t)img | (uintptr_t)out) & 15) == 0) && npx >= 16) {
        const size_t groups = npx / 16, tail = npx - groups * 16;
        CS_LAUNCH(k_apply_lut_rgb8, grid_for(groups, 256), 256, 0, st, (const uint4*)img, groups, lut, (uint4*)out);
        if (tail) CS_LAUNCH(k_apply_lut<uint8_t>, 1, 32, 0, st, (const uint8_t*)img + groups * 48, tail, C, lut, out + groups * 48);
    } else if (elem_bytes == 1 && C == 1 && ((((uintptr_t)img | (uintptr_t)out) & 15) == 0) && npx >= 16) {
        const size_t groups = npx / 16, tail = npx - groups * 16;
        CS_LAUNCH(k_apply_lut_gray8, grid_for(groups, 256), 256, 0, st, (const uint4*)img, groups, lut, (uint4*)out);
        if (tail) CS_LAUNCH(k_apply_lut<uint8_t>, 1, 32, 0, st, (const uint8_t*)img + groups * 16, tail, C, lut, out + groups * 48);
    } else if (elem_bytes == 1) CS_LAUNCH(k_apply_lut<uint8_t>, grid_for(npx, 256), 256, 0, st, (const uint8_t*)img, npx, C, lut, out);
    else CS_LAUNCH(k_apply_lut<uint16_t>, grid_for(npx, 256), 256, 0, st, (const uint16_t*)img, npx, C, lut, out);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace cs
