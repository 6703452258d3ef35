// Glue between the cuDNN convolutions of the predictor (monai UNet, baseline/predict.py:45-52): the "ADN" block
// InstanceNorm2d(affine=False) → PReLU, optionally followed by the residual add of the enclosing ResidualUnit, as ONE
// pass over each (sample, channel) plane: the plane is read once into registers, mean and variance are reduced in
// the CTA (two-pass, biased variance, eps inside the square root — torch.nn.functional.instance_norm's formula),
// and the result is written once.  PyTorch runs the same block as 3-4 separate full-tensor kernels.
// Planes too large for the register file (the 3 x 256 x 256 output head) take the streaming variant, whose second
// read comes from L2.  The convolutions themselves stay on cuDNN (out of scope by the north star).
#include "internal.cuh"

namespace cs {

constexpr int NORM_MAX_VPT = 8;      // float4 per thread held in registers

__device__ __forceinline__ float block_sum(float v, float* red /* [33] */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    __syncthreads();                                   // red[] free again
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        float s = lane < nw ? red[lane] : 0.f;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) red[32] = s;
    }
    __syncthreads();
    return red[32];
}

__device__ __forceinline__ float prelu1(float v, float a) { return v > 0.f ? v : a * v; }

// Plane resident in registers: HW4 = HW/4 float4 per plane, blockDim.x * NORM_MAX_VPT >= HW4.
__global__ void __launch_bounds__(512) k_instnorm_prelu_reg(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ y,
                                                            const float* __restrict__ prelu_w, const float* __restrict__ res_bias, int C,
                                                            int HW4, float eps) {
    __shared__ float red[33];
    const size_t plane = (size_t)blockIdx.x * HW4;
    const float4* xp = reinterpret_cast<const float4*>(x) + plane;
    float4 v[NORM_MAX_VPT];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_VPT; ++i) {
        const int j = threadIdx.x + i * blockDim.x;
        if (j < HW4) { v[i] = xp[j]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
    }
    const float inv_n = 1.0f / (4.0f * (float)HW4);
    const float mean = block_sum(s, red) * inv_n;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_VPT; ++i) {
        const int j = threadIdx.x + i * blockDim.x;
        if (j < HW4) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float rstd = rsqrtf(block_sum(q, red) * inv_n + eps);
    const float a = prelu_w ? *prelu_w : 1.0f;
    float4* yp = reinterpret_cast<float4*>(y) + plane;
    const float4* rp = res ? reinterpret_cast<const float4*>(res) + plane : nullptr;
    const float rb = (res && res_bias) ? res_bias[blockIdx.x % C] : 0.f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_VPT; ++i) {
        const int j = threadIdx.x + i * blockDim.x;
        if (j < HW4) {
            float4 o;
            o.x = prelu1((v[i].x - mean) * rstd, a); o.y = prelu1((v[i].y - mean) * rstd, a);
            o.z = prelu1((v[i].z - mean) * rstd, a); o.w = prelu1((v[i].w - mean) * rstd, a);
            if (rp) { const float4 r = rp[j]; o.x += r.x + rb; o.y += r.y + rb; o.z += r.z + rb; o.w += r.w + rb; }
            yp[j] = o;
        }
    }
}

// Streaming variant for large planes: three sweeps (sum, squared deviations, apply); sweeps 2 and 3 hit L2.
__global__ void __launch_bounds__(1024) k_instnorm_prelu_stream(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ y,
                                                                const float* __restrict__ prelu_w, const float* __restrict__ res_bias, int C,
                                                                int HW4, float eps) {
    __shared__ float red[33];
    const size_t plane = (size_t)blockIdx.x * HW4;
    const float4* xp = reinterpret_cast<const float4*>(x) + plane;
    float s = 0.f;
    for (int j = threadIdx.x; j < HW4; j += blockDim.x) { const float4 t = xp[j]; s += (t.x + t.y) + (t.z + t.w); }
    const float inv_n = 1.0f / (4.0f * (float)HW4);
    const float mean = block_sum(s, red) * inv_n;
    float q = 0.f;
    for (int j = threadIdx.x; j < HW4; j += blockDim.x) {
        const float4 t = xp[j];
        const float a = t.x - mean, b = t.y - mean, c = t.z - mean, d = t.w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(block_sum(q, red) * inv_n + eps);
    const float a = prelu_w ? *prelu_w : 1.0f;
    float4* yp = reinterpret_cast<float4*>(y) + plane;
    const float4* rp = res ? reinterpret_cast<const float4*>(res) + plane : nullptr;
    const float rb = (res && res_bias) ? res_bias[blockIdx.x % C] : 0.f;
    for (int j = threadIdx.x; j < HW4; j += blockDim.x) {
        const float4 t = xp[j];
        float4 o;
        o.x = prelu1((t.x - mean) * rstd, a); o.y = prelu1((t.y - mean) * rstd, a);
        o.z = prelu1((t.z - mean) * rstd, a); o.w = prelu1((t.w - mean) * rstd, a);
        if (rp) { const float4 r = rp[j]; o.x += r.x + rb; o.y += r.y + rb; o.z += r.z + rb; o.w += r.w + rb; }
        yp[j] = o;
    }
}

// y[p] = prelu(instance_norm(x[p])) (+ res[p] + res_bias[p % C]) for `planes` contiguous planes of HW floats; y may alias x.
int instnorm_prelu(const float* x, const float* res, float* y, const float* prelu_w, const float* res_bias, int C, long long planes, int HW,
                   float eps, cudaStream_t st) {
    CS_CHECK_ARG(x && y, "null pointer");
    CS_CHECK_ARG(!res_bias || C > 0, "res_bias needs the channel count");
    CS_CHECK_ARG(planes > 0 && planes <= 0x7fffffffLL && HW > 0 && (HW & 3) == 0, "planes must be positive and the plane size a multiple of 4");
    CS_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)res & 15) == 0, "pointers must be 16-byte aligned");
    const int HW4 = HW / 4;
    if (HW4 <= 512 * NORM_MAX_VPT) {
        int nt = (HW4 + NORM_MAX_VPT - 1) / NORM_MAX_VPT;
        nt = ((nt + 31) / 32) * 32;
        if (nt < 64) nt = HW4 >= 64 ? 64 : 32;
        if (nt > 512) nt = 512;
        CS_LAUNCH(k_instnorm_prelu_reg, (unsigned)planes, nt, 0, st, x, res, y, prelu_w, res_bias, C > 0 ? C : 1, HW4, eps);
    } else {
        CS_LAUNCH(k_instnorm_prelu_stream, (unsigned)planes, 1024, 0, st, x, res, y, prelu_w, res_bias, C > 0 ? C : 1, HW4, eps);
    }
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace cs
