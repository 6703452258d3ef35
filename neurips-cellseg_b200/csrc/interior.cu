// Instance map -> three-class training target {0 background, 1 interior, 2 boundary}: the per-pixel part of
// data/pre_process_3class.py:31-56 (create_interior_map) of the reference:
//   boundary = binary_dilation(find_boundaries(inst, mode='inner'), disk(1));  interior = ~boundary & (inst > 0)
// then remove_small_objects(interior, 16) runs on the size-filter chain of ccl.cu and k_interior_combine writes the classes.
// Labels are compared as int16, because the reference converts them with .astype(np.int16) before the call (:98).
#include "internal.cuh"

namespace cs {

// inner boundary of skimage.segmentation.find_boundaries (connectivity 1): a foreground pixel one of whose 4-neighbours INSIDE the
// image carries another value (skimage's grey dilation/erosion reflect at the edge, so the edge itself is no boundary)
__device__ __forceinline__ bool inner_boundary(const int32_t* __restrict__ L, int H, int W, int y, int x) {
    const short c = (short)L[(size_t)y * W + x];
    if (c == 0) return false;
    bool d = false;
    if (y > 0) d = d || (short)L[(size_t)(y - 1) * W + x] != c;
    if (y + 1 < H) d = d || (short)L[(size_t)(y + 1) * W + x] != c;
    if (x > 0) d = d || (short)L[(size_t)y * W + x - 1] != c;
    if (x + 1 < W) d = d || (short)L[(size_t)y * W + x + 1] != c;
    return d;
}

__global__ void __launch_bounds__(256) k_interior_boundary(const int32_t* __restrict__ inst, int N, int H, int W, uint8_t* __restrict__ boundary,
                                                           uint8_t* __restrict__ interior) {
    const size_t HW = (size_t)H * W, total = (size_t)N * HW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t n = i / HW, p = i % HW;
        const int y = (int)(p / W), x = (int)(p % W);
        const int32_t* L = inst + n * HW;
        // binary_dilation with the cross-shaped disk(1): the pixel itself or a 4-neighbour inside the image is an inner boundary
        bool b = inner_boundary(L, H, W, y, x);
        if (!b && y > 0) b = inner_boundary(L, H, W, y - 1, x);
        if (!b && y + 1 < H) b = inner_boundary(L, H, W, y + 1, x);
        if (!b && x > 0) b = inner_boundary(L, H, W, y, x - 1);
        if (!b && x + 1 < W) b = inner_boundary(L, H, W, y, x + 1);
        boundary[i] = b ? 1 : 0;
        interior[i] = (!b && (short)L[p] > 0) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) k_interior_combine(const uint8_t* __restrict__ boundary, const uint8_t* __restrict__ interior, size_t n,
                                                          uint8_t* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = boundary[i] ? 2 : (interior[i] ? 1 : 0);
}

static int grid_for(size_t work) {
    size_t b = (work + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148 * 32) b = 148 * 32;
    return (int)b;
}

int interior_boundary(const int32_t* inst, int N, int H, int W, uint8_t* boundary, uint8_t* interior, cudaStream_t st) {
    CS_CHECK_ARG(inst && boundary && interior, "null pointer");
    CS_CHECK_ARG(N > 0 && H > 0 && W > 0, "empty image");
    CS_LAUNCH(k_interior_boundary, grid_for((size_t)N * H * W), 256, 0, st, inst, N, H, W, boundary, interior);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

int interior_combine(const uint8_t* boundary, const uint8_t* interior, size_t n, uint8_t* out, cudaStream_t st) {
    CS_CHECK_ARG(boundary && interior && out, "null pointer");
    if (n == 0) return CS_OK;
    CS_LAUNCH(k_interior_combine, grid_for(n), 256, 0, st, boundary, interior, n, out);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace cs
