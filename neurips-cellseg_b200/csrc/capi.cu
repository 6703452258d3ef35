// extern "C" surface of libcellseg_sm100.so (declared in include/cellseg_b200.h).
#include "internal.cuh"
#include "../../include/cellseg_b200.h"

namespace cs {
// RAII device buffer for the *_host entry points.
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) {
        cudaError_t e = cudaMalloc(&p, n ? n : 1);
        if (e != cudaSuccess) { p = nullptr; set_error("cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e)); return CS_ERR_CUDA; }
        return CS_OK;
    }
    template <class T> T* as() { return (T*)p; }
};
}  // namespace cs

using namespace cs;

extern "C" {

const char* cs_last_error(void) { return cs::last_error(); }
int cs_abi_version(void) { return 1; }

size_t cs_postprocess_workspace_bytes(int N, int H, int W) { return cs::postprocess_workspace_bytes(N, H, W); }

int cs_postprocess(const uint8_t* mask, int N, int H, int W, int hole_area, int min_size, int connectivity,
                   int32_t* labels, int32_t* n_out, void* workspace, size_t workspace_bytes, cs_stream_t stream) {
    return cs::postprocess(mask, N, H, W, hole_area, min_size, connectivity, labels, n_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int cs_label(const uint8_t* mask, int N, int H, int W, int connectivity, int32_t* labels, int32_t* n_out,
             void* workspace, size_t workspace_bytes, cs_stream_t stream) {
    return cs::label(mask, N, H, W, connectivity, labels, n_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int cs_remove_small(const uint8_t* mask_in, uint8_t* mask_out, int N, int H, int W, int min_size, int invert,
                    int32_t* scratch, void* workspace, size_t workspace_bytes, cs_stream_t stream) {
    return cs::remove_small(mask_in, mask_out, N, H, W, 4, min_size, invert, scratch, workspace, workspace_bytes, (cudaStream_t)stream);
}

uint32_t* cs_postprocess_packed_slot(void* workspace, size_t workspace_bytes, int N, int H, int W) {
    return cs::packed_mask_slot(workspace, workspace_bytes, N, H, W);
}
int cs_postprocess_packed(int N, int H, int W, int hole_area, int min_size, int connectivity, int32_t* labels, int32_t* n_out,
                          void* workspace, size_t workspace_bytes, cs_stream_t stream) {
    return cs::postprocess_packed(N, H, W, hole_area, min_size, connectivity, labels, n_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int cs_scale_lut_u8(const uint8_t* img, int N, size_t bytes_per_image, uint32_t* max_scratch, float* lut, cs_stream_t stream) {
    return cs::scale_lut_u8(img, N, bytes_per_image, max_scratch, lut, (cudaStream_t)stream);
}
int cs_u8hwc_to_f32chw(const uint8_t* img, int N, int H, int W, const float* lut, float* out, cs_stream_t stream) {
    return cs::u8hwc_to_f32chw(img, N, H, W, lut, out, (cudaStream_t)stream);
}
int cs_gather_windows(const float* img_chw, int N, int H, int W, int R, int ny, int nx, int pad_y, int pad_x,
                      const int* ys, const int* xs, int win_begin, int win_count, float* batch, cs_stream_t stream) {
    cs::WinGeo g{N, H, W, R, ny, nx, pad_y, pad_x};
    return cs::gather_windows(img_chw, 0, nullptr, g, ys, xs, win_begin, win_count, batch, (cudaStream_t)stream);
}
int cs_gather_windows_u8(const uint8_t* img_hwc, const float* lut, int N, int H, int W, int R, int ny, int nx, int pad_y, int pad_x,
                         const int* ys, const int* xs, int win_begin, int win_count, float* batch, cs_stream_t stream) {
    cs::WinGeo g{N, H, W, R, ny, nx, pad_y, pad_x};
    return cs::gather_windows(img_hwc, 1, lut, g, ys, xs, win_begin, win_count, batch, (cudaStream_t)stream);
}
int cs_blend_softmax_threshold(const float* win_logits, int C, int N, int H, int W, int R, int ny, int nx, int pad_y, int pad_x,
                               const int* ys, const int* xs, const int* cover_y, const int* cover_x, const float* imap,
                               float threshold, uint32_t* packed, uint8_t* mask, float* p1, float* blended, cs_stream_t stream) {
    cs::WinGeo g{N, H, W, R, ny, nx, pad_y, pad_x};
    cs::BlendArgs a{win_logits, imap, ys, xs, cover_y, cover_x, packed, mask, p1, blended, threshold};
    return cs::blend_softmax_threshold(a, g, C, (cudaStream_t)stream);
}

int cs_postprocess_host(const uint8_t* mask, int N, int H, int W, int hole_area, int min_size, int connectivity,
                        int32_t* labels, int32_t* n_out) {
    CS_CHECK_ARG(mask && labels, "null pointer");
    size_t wsb = cs::postprocess_workspace_bytes(N, H, W);
    if (!wsb) return CS_ERR_ARG;
    size_t px = (size_t)N * H * W;
    DevBuf dm, dl, dn, dw;
    int rc;
    if ((rc = dm.alloc(px)) || (rc = dl.alloc(px * 4)) || (rc = dn.alloc((size_t)N * 4)) || (rc = dw.alloc(wsb))) return rc;
    CS_CUDA(cudaMemcpyAsync(dm.p, mask, px, cudaMemcpyHostToDevice, 0));
    rc = cs::postprocess(dm.as<uint8_t>(), N, H, W, hole_area, min_size, connectivity, dl.as<int32_t>(), dn.as<int32_t>(), dw.p, wsb, 0);
    if (rc != CS_OK) return rc;
    CS_CUDA(cudaMemcpyAsync(labels, dl.p, px * 4, cudaMemcpyDeviceToHost, 0));
    if (n_out) CS_CUDA(cudaMemcpyAsync(n_out, dn.p, (size_t)N * 4, cudaMemcpyDeviceToHost, 0));
    CS_CUDA(cudaStreamSynchronize(0));
    return CS_OK;
}

int cs_label_host(const uint8_t* mask, int N, int H, int W, int connectivity, int32_t* labels, int32_t* n_out) {
    return cs_postprocess_host(mask, N, H, W, 0, 0, connectivity, labels, n_out);
}

int cs_remove_small_host(const uint8_t* mask_in, uint8_t* mask_out, int N, int H, int W, int min_size, int invert) {
    CS_CHECK_ARG(mask_in && mask_out, "null pointer");
    size_t wsb = cs::postprocess_workspace_bytes(N, H, W);
    if (!wsb) return CS_ERR_ARG;
    size_t px = (size_t)N * H * W;
    DevBuf di, dout, ds, dw;
    int rc;
    if ((rc = di.alloc(px)) || (rc = dout.alloc(px)) || (rc = ds.alloc(px * 4)) || (rc = dw.alloc(wsb))) return rc;
    CS_CUDA(cudaMemcpyAsync(di.p, mask_in, px, cudaMemcpyHostToDevice, 0));
    rc = cs::remove_small(di.as<uint8_t>(), dout.as<uint8_t>(), N, H, W, 4, min_size, invert, ds.as<int32_t>(), dw.p, wsb, 0);
    if (rc != CS_OK) return rc;
    CS_CUDA(cudaMemcpyAsync(mask_out, dout.p, px, cudaMemcpyDeviceToHost, 0));
    CS_CUDA(cudaStreamSynchronize(0));
    return CS_OK;
}

}  // extern "C"
