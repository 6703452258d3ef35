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

// u8: per-CTA privatised 3 x 256 histogram in shared memory, flushed with one atomic per non-empty bin.
__global__ void __launch_bounds__(256) k_hist_u8(const uint8_t* __restrict__ img, size_t npx, int C, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[3 * 256];
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const int nc = C < 3 ? C : 3;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npx; p += (size_t)gridDim.x * blockDim.x) {
        const uint8_t* s = img + p * C;
        for (int c = 0; c < nc; ++c) atomicAdd(&sh[c * 256 + s[c]], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nc * 256; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// u16: 65 536 bins per channel live in L2; runs of equal values inside a thread's pixels are merged first.
__global__ void __launch_bounds__(256) k_hist_u16(const uint16_t* __restrict__ img, size_t npx, int C, uint32_t* __restrict__ hist) {
    const int nc = C < 3 ? C : 3;
    constexpr int PER = 8;
    for (size_t p0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * PER; p0 < npx; p0 += (size_t)gridDim.x * blockDim.x * PER) {
        const size_t lim = npx - p0 < PER ? npx - p0 : PER;
        for (int c = 0; c < nc; ++c) {
            uint32_t cur = img[p0 * C + c], run = 1;
            for (size_t i = 1; i < lim; ++i) {
                const uint32_t v = img[(p0 + i) * C + c];
                if (v == cur) { ++run; continue; }
                atomicAdd(&hist[(size_t)c * 65536 + cur], run);
                cur = v; run = 1;
            }
            atomicAdd(&hist[(size_t)c * 65536 + cur], run);
        }
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
    if (elem_bytes == 1) CS_LAUNCH(k_hist_u8, grid_for(npx, 256 * 8), 256, 0, st, (const uint8_t*)img, npx, C, hist);
    else CS_LAUNCH(k_hist_u16, grid_for(npx, 256 * 8), 256, 0, st, (const uint16_t*)img, npx, C, hist);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// lut: device uint8[3][bins]; out: device uint8[H][W][3].
int prenorm_apply(const void* img, int elem_bytes, int H, int W, int C, const uint8_t* lut, uint8_t* out, cudaStream_t st) {
    CS_CHECK_ARG(img && lut && out && H > 0 && W > 0 && C >= 1, "null pointer or empty image");
    CS_CHECK_ARG(elem_bytes == 1 || elem_bytes == 2, "element size must be 1 (uint8) or 2 (uint16)");
    const size_t npx = (size_t)H * W;
    if (elem_bytes == 1) CS_LAUNCH(k_apply_lut<uint8_t>, grid_for(npx, 256), 256, 0, st, (const uint8_t*)img, npx, C, lut, out);
    else CS_LAUNCH(k_apply_lut<uint16_t>, grid_for(npx, 256), 256, 0, st, (const uint16_t*)img, npx, C, lut, out);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace cs
