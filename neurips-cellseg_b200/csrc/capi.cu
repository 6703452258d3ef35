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
unsigned long long cs_kernel_launch_count(void) { return cs::launch_count(); }
#ifndef CS_EMU
int cs_profile_enable(int on) { cs::prof::enable(on != 0); return CS_OK; }
int cs_profile_read(int index, char* name, size_t name_cap, double* total_ms, unsigned long long* count) {
    return cs::prof::read(index, name, name_cap, total_ms, count);
}
#else
int cs_profile_enable(int) { return CS_OK; }
int cs_profile_read(int, char*, size_t, double*, unsigned long long*) { return 0; }
#endif

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
                    void* workspace, size_t workspace_bytes, cs_stream_t stream) {
    return cs::remove_small(mask_in, mask_out, N, H, W, 4, min_size, invert, workspace, workspace_bytes, (cudaStream_t)stream);
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
int cs_prenorm_hist(const void* img, int elem_bytes, int H, int W, int C, uint32_t* hist, cs_stream_t stream) {
    return cs::prenorm_hist(img, elem_bytes, H, W, C, hist, (cudaStream_t)stream);
}
int cs_prenorm_apply(const void* img, int elem_bytes, int H, int W, int C, const uint8_t* lut, uint8_t* out, cs_stream_t stream) {
    return cs::prenorm_apply(img, elem_bytes, H, W, C, lut, out, (cudaStream_t)stream);
}
int cs_instnorm_prelu(const float* x, const float* residual, float* y, const float* prelu_weight, const float* residual_bias, int C,
                      long long planes, int HW, float eps, cs_stream_t stream) {
    return cs::instnorm_prelu(x, residual, y, prelu_weight, residual_bias, C, planes, HW, eps, (cudaStream_t)stream);
}
int cs_nhwc_stat_chunks(int HW, int C) { return cs::nhwc_stat_chunks(HW, C); }
int cs_nhwc_stats(const float* x, int N, int HW, int C, float* partials, cs_stream_t stream) {
    return cs::nhwc_stats(x, N, HW, C, partials, (cudaStream_t)stream);
}
int cs_nhwc_norm_prelu(const float* x, const float* residual, float* y, const float* prelu_weight, const float* residual_bias, int N, int HW, int C,
                       const float* partials, int stat_chunk_pixels, float eps, cs_stream_t stream) {
    return cs::nhwc_norm_prelu(x, residual, y, prelu_weight, residual_bias, N, HW, C, partials, stat_chunk_pixels, eps, (cudaStream_t)stream);
}
int cs_instnorm_prelu_nhwc(const float* x, const float* residual, float* y, const float* prelu_weight, const float* residual_bias, int N, int HW, int C,
                           float eps, float* partials, float* y2, int c2_total, int c2_offset, cs_stream_t stream) {
    return cs::instnorm_prelu_nhwc(x, residual, y, prelu_weight, residual_bias, N, HW, C, eps, partials, y2, c2_total, c2_offset, (cudaStream_t)stream);
}
int cs_instnorm_prelu_nhwc_slices(const float* x, int x_ctotal, int x_coffset, const float* residual, int r_ctotal, int r_coffset, float* y,
                                  const float* prelu_weight, const float* residual_bias, int N, int HW, int C, float eps, float* partials, float* y2,
                                  int c2_total, int c2_offset, cs_stream_t stream) {
    return cs::instnorm_prelu_nhwc_slices(x, x_ctotal, x_coffset, residual, r_ctotal, r_coffset, y, prelu_weight, residual_bias, N, HW, C, eps, partials,
                                          y2, c2_total, c2_offset, (cudaStream_t)stream);
}
int cs_instnorm_prelu_nhwc_shuffle(const float* x, float* y, const float* prelu_weight, int N, int H, int W, int C, float eps, float* partials,
                                   cs_stream_t stream) {
    return cs::instnorm_prelu_nhwc_shuffle(x, y, prelu_weight, N, H, W, C, eps, partials, (cudaStream_t)stream);
}
int cs_unet_stem(const float* win, int N, int H, int W, const float* w_unit0, const float* w_res, float* a_out, float* r_out, float* a_partials,
                 cs_stream_t stream) {
    return cs::unet_stem(win, N, H, W, w_unit0, w_res, a_out, r_out, a_partials, (cudaStream_t)stream);
}
int cs_unet_stem_u8(const uint8_t* img_hwc, const float* lut, int N, int H, int W, int R, int ny, int nx, int pad_y, int pad_x, const int* ys,
                    const int* xs, int win_begin, int win_count, const float* w_unit0, const float* w_res, float* a_out, float* r_out,
                    float* a_partials, cs_stream_t stream) {
    cs::WinGeo g{N, H, W, R, ny, nx, pad_y, pad_x};
    return cs::unet_stem_u8(img_hwc, lut, g, ys, xs, win_begin, win_count, w_unit0, w_res, a_out, r_out, a_partials, (cudaStream_t)stream);
}
int cs_unet_head_partial_chunks(int Hin, int Win) { return cs::unet_head_partial_chunks(Hin, Win); }
int cs_unet_head_convt(const float* xa, const float* xb, int N, int Hin, int Win, const float* wt, float* z, float* z_partials, cs_stream_t stream) {
    return cs::unet_head_convt(xa, xb, N, Hin, Win, wt, z, z_partials, (cudaStream_t)stream);
}
int cs_unet_head_final(const float* z, const float* z_partials, int N, int H, int W, const float* prelu_weight, const float* w3, const float* b3,
                       float eps, float* out, cs_stream_t stream) {
    return cs::unet_head_final(z, z_partials, N, H, W, prelu_weight, w3, b3, eps, out, (cudaStream_t)stream);
}
int cs_interior_boundary(const int32_t* inst, int N, int H, int W, uint8_t* boundary, uint8_t* interior, cs_stream_t stream) {
    return cs::interior_boundary(inst, N, H, W, boundary, interior, (cudaStream_t)stream);
}
int cs_interior_combine(const uint8_t* boundary, const uint8_t* interior, size_t n, uint8_t* out, cs_stream_t stream) {
    return cs::interior_combine(boundary, interior, n, out, (cudaStream_t)stream);
}
int cs_u8_max(const uint8_t* img, int N, size_t bytes_per_image, uint32_t* maxv, cs_stream_t stream) {
    return cs::u8_max(img, N, bytes_per_image, maxv, (cudaStream_t)stream);
}
int cs_scale_lut_from_max(const uint32_t* maxv, int N, float* lut, cs_stream_t stream) {
    return cs::scale_lut_from_max(maxv, N, lut, (cudaStream_t)stream);
}
size_t cs_strip_workspace_bytes(int H, int W) { return cs::strip_workspace_bytes(H, W); }
uint32_t* cs_strip_packed_slot(void* workspace, size_t workspace_bytes, int H, int W) { return cs::strip_packed_slot(workspace, workspace_bytes, H, W); }
int cs_strip_begin(const uint8_t* mask_or_null, int H, int W, int hole_area, int min_size, int connectivity, int32_t* labels,
                   void* workspace, size_t workspace_bytes, cs_stream_t stream) {
    return cs::strip_begin(mask_or_null, H, W, hole_area, min_size, connectivity, labels, workspace, workspace_bytes, (cudaStream_t)stream);
}
int cs_strip_filter_export(int stage, int H, int W, int hole_area, int min_size, int connectivity, int32_t* labels, int32_t* rows_out,
                           void* workspace, size_t workspace_bytes, cs_stream_t stream) {
    return cs::strip_filter_export(stage, H, W, hole_area, min_size, connectivity, labels, rows_out, workspace, workspace_bytes, (cudaStream_t)stream);
}
int cs_strip_filter_resume(int stage, int H, int W, int hole_area, int min_size, int connectivity, int32_t* labels,
                           const int32_t* above_last_row, const int32_t* below_first_row, void* workspace, size_t workspace_bytes, cs_stream_t stream) {
    return cs::strip_filter_resume(stage, H, W, hole_area, min_size, connectivity, labels, above_last_row, below_first_row, workspace,
                                   workspace_bytes, (cudaStream_t)stream);
}
int cs_strip_label_export(int H, int W, int connectivity, int32_t* labels, int32_t* rows_out, int32_t* count_out, void* workspace,
                          size_t workspace_bytes, cs_stream_t stream) {
    return cs::strip_label_export(H, W, connectivity, labels, rows_out, count_out, workspace, workspace_bytes, (cudaStream_t)stream);
}
size_t cs_strip_resolve_workspace_bytes(long long total_components) { return cs::strip_resolve_workspace_bytes(total_components); }
int cs_strip_label_resolve(const int32_t* rows_all, const int32_t* offsets, int nranks, int W, int connectivity, int my_rank, int32_t first,
                           int32_t count, long long total, int32_t* lut_out, int32_t* total_out, void* workspace2, size_t workspace2_bytes,
                           cs_stream_t stream) {
    return cs::strip_label_resolve(rows_all, offsets, nranks, W, connectivity, my_rank, first, count, total, lut_out, total_out, workspace2,
                                   workspace2_bytes, (cudaStream_t)stream);
}
int cs_strip_final(int H, int W, const int32_t* lut, int32_t* labels, void* workspace, size_t workspace_bytes, cs_stream_t stream) {
    return cs::strip_final(H, W, lut, labels, workspace, workspace_bytes, (cudaStream_t)stream);
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

int cs_certify_large(const uint8_t* mask, int N, int H, int W, int hole_area, int min_size, uint8_t* large_bg, uint8_t* large_fg,
                     void* workspace, size_t workspace_bytes, cs_stream_t stream) {
    return cs::certify_masks(mask, N, H, W, hole_area, min_size, large_bg, large_fg, workspace, workspace_bytes, (cudaStream_t)stream);
}

int cs_remove_small_host(const uint8_t* mask_in, uint8_t* mask_out, int N, int H, int W, int min_size, int invert) {
    CS_CHECK_ARG(mask_in && mask_out, "null pointer");
    size_t wsb = cs::postprocess_workspace_bytes(N, H, W);
    if (!wsb) return CS_ERR_ARG;
    size_t px = (size_t)N * H * W;
    DevBuf di, dout, dw;
    int rc;
    if ((rc = di.alloc(px)) || (rc = dout.alloc(px)) || (rc = dw.alloc(wsb))) return rc;
    CS_CUDA(cudaMemcpyAsync(di.p, mask_in, px, cudaMemcpyHostToDevice, 0));
    rc = cs::remove_small(di.as<uint8_t>(), dout.as<uint8_t>(), N, H, W, 4, min_size, invert, dw.p, wsb, 0);
    if (rc != CS_OK) return rc;
    CS_CUDA(cudaMemcpyAsync(mask_out, dout.p, px, cudaMemcpyDeviceToHost, 0));
    CS_CUDA(cudaStreamSynchronize(0));
    return CS_OK;
}


int cs_label_max(const int32_t* lab, int H, int W, int ld, int32_t* out2, cs_stream_t stream) {
    return cs::label_max(lab, H, W, ld, out2, (cudaStream_t)stream);
}
size_t cs_relabel_workspace_bytes(int32_t max_label) { return cs::relabel_workspace_bytes(max_label); }
int cs_relabel(int32_t* lab, int H, int W, int ld, int border, int32_t max_label, int zero_inplace, int32_t* out, int ldo,
               int32_t* n_out, void* workspace, size_t workspace_bytes, cs_stream_t stream) {
    return cs::relabel(lab, H, W, ld, border, max_label, zero_inplace, out, ldo, n_out, workspace, workspace_bytes, (cudaStream_t)stream);
}
size_t cs_overlap_workspace_bytes(uint32_t table_capacity) { return cs::overlap_workspace_bytes(table_capacity); }
int cs_label_overlap_sparse(const int32_t* x, int ldx, const int32_t* y, int ldy, int H, int W, int32_t nt, int32_t np,
                            int32_t* pair_gt, int32_t* pair_pr, uint32_t* pair_cnt, uint32_t pair_cap, uint32_t* n_pairs,
                            uint32_t* area_gt, uint32_t* area_pr, uint32_t* status, uint32_t table_capacity,
                            void* workspace, size_t workspace_bytes, cs_stream_t stream) {
    return cs::label_overlap_sparse(x, ldx, y, ldy, H, W, nt, np, pair_gt, pair_pr, pair_cnt, pair_cap, n_pairs, area_gt, area_pr,
                                    status, table_capacity, workspace, workspace_bytes, (cudaStream_t)stream);
}
int cs_overlap_dense(const int32_t* pair_gt, const int32_t* pair_pr, const uint32_t* pair_cnt, const uint32_t* n_pairs, uint32_t pair_cap,
                     int32_t nt, int32_t np, uint64_t* ov, cs_stream_t stream) {
    return cs::dense_overlap(pair_gt, pair_pr, pair_cnt, n_pairs, pair_cap, nt, np, (unsigned long long*)ov, (cudaStream_t)stream);
}
int cs_iou_dense(const int32_t* pair_gt, const int32_t* pair_pr, const uint32_t* pair_cnt, const uint32_t* n_pairs, uint32_t pair_cap,
                 int32_t nt, int32_t np, const uint32_t* area_gt, const uint32_t* area_pr, double* iou, cs_stream_t stream) {
    return cs::dense_iou(pair_gt, pair_pr, pair_cnt, n_pairs, pair_cap, nt, np, area_gt, area_pr, iou, (cudaStream_t)stream);
}

int cs_iou_edges(const int32_t* pair_gt, const int32_t* pair_pr, const uint32_t* pair_cnt, const uint32_t* n_pairs, uint32_t pair_cap,
                 int32_t nt, int32_t np, const uint32_t* area_gt, const uint32_t* area_pr, double threshold, uint32_t* deg_gt, uint32_t* deg_pr,
                 int32_t* edge_gt, int32_t* edge_pr, uint32_t* out2, cs_stream_t stream) {
    return cs::iou_edges(pair_gt, pair_pr, pair_cnt, n_pairs, pair_cap, nt, np, area_gt, area_pr, threshold, deg_gt, deg_pr, edge_gt, edge_pr, out2,
                         (cudaStream_t)stream);
}

static int relabel_host(int32_t* mask_inout, const int32_t* mask_in, int H, int W, int border, int32_t* out, int32_t* n_out) {
    CS_CHECK_ARG((mask_inout || mask_in) && out && H > 0 && W > 0, "null pointer or empty image");
    const int32_t* src = mask_inout ? mask_inout : mask_in;
    size_t px = (size_t)H * W;
    DevBuf dl, dout, dmm, dn, dw;
    int rc;
    if ((rc = dl.alloc(px * 4)) || (rc = dout.alloc(px * 4)) || (rc = dmm.alloc(8)) || (rc = dn.alloc(4))) return rc;
    CS_CUDA(cudaMemcpyAsync(dl.p, src, px * 4, cudaMemcpyHostToDevice, 0));
    if ((rc = cs::label_max(dl.as<int32_t>(), H, W, W, dmm.as<int32_t>(), 0)) != CS_OK) return rc;
    int32_t mm[2];
    CS_CUDA(cudaMemcpyAsync(mm, dmm.p, 8, cudaMemcpyDeviceToHost, 0));
    CS_CUDA(cudaStreamSynchronize(0));
    if (mm[1] < 0) { set_error("negative labels are not supported"); return CS_ERR_ARG; }
    size_t wsb = cs::relabel_workspace_bytes(mm[0]);
    if ((rc = dw.alloc(wsb))) return rc;
    rc = cs::relabel(dl.as<int32_t>(), H, W, W, border, mm[0], mask_inout != nullptr, dout.as<int32_t>(), W, dn.as<int32_t>(), dw.p, wsb, 0);
    if (rc != CS_OK) return rc;
    CS_CUDA(cudaMemcpyAsync(out, dout.p, px * 4, cudaMemcpyDeviceToHost, 0));
    if (mask_inout) CS_CUDA(cudaMemcpyAsync(mask_inout, dl.p, px * 4, cudaMemcpyDeviceToHost, 0));
    if (n_out) CS_CUDA(cudaMemcpyAsync(n_out, dn.p, 4, cudaMemcpyDeviceToHost, 0));
    CS_CUDA(cudaStreamSynchronize(0));
    return CS_OK;
}
int cs_remove_boundary_cells_host(int32_t* mask, int H, int W, int32_t* out, int32_t* n_out) {
    return relabel_host(mask, nullptr, H, W, 2, out, n_out);
}
int cs_relabel_sequential_host(const int32_t* lab, int H, int W, int32_t* out, int32_t* n_out) {
    return relabel_host(nullptr, lab, H, W, 0, out, n_out);
}

int cs_count_nonzero3(const uint8_t* gt, const uint8_t* seg, size_t n, uint64_t* out3, cs_stream_t stream) {
    return cs::count3(gt, seg, n, (unsigned long long*)out3, (cudaStream_t)stream);
}
int cs_count_nonzero3_host(const uint8_t* gt, const uint8_t* seg, size_t n, uint64_t* out3) {
    CS_CHECK_ARG(gt && seg && out3, "null pointer");
    DevBuf da, db, dc; int rc;
    if ((rc = da.alloc(n)) || (rc = db.alloc(n)) || (rc = dc.alloc(24))) return rc;
    CS_CUDA(cudaMemcpyAsync(da.p, gt, n, cudaMemcpyHostToDevice, 0));
    CS_CUDA(cudaMemcpyAsync(db.p, seg, n, cudaMemcpyHostToDevice, 0));
    if ((rc = cs::count3(da.as<uint8_t>(), db.as<uint8_t>(), n, dc.as<unsigned long long>(), 0)) != CS_OK) return rc;
    CS_CUDA(cudaMemcpyAsync(out3, dc.p, 24, cudaMemcpyDeviceToHost, 0));
    CS_CUDA(cudaStreamSynchronize(0));
    return CS_OK;
}
int cs_label_overlap_host(const int32_t* x, const int32_t* y, int H, int W, int32_t nt, int32_t np, uint64_t* ov) {
    CS_CHECK_ARG(x && y && ov && H > 0 && W > 0 && nt >= 0 && np >= 0, "null pointer or empty image");
    size_t px = (size_t)H * W;
    DevBuf dx, dy, dpg, dpp, dpc, dmisc, dag, dap, dw, dov;
    int rc;
    if ((rc = dx.alloc(px * 4)) || (rc = dy.alloc(px * 4)) || (rc = dmisc.alloc(64)) ||
        (rc = dag.alloc(((size_t)nt + 1) * 4)) || (rc = dap.alloc(((size_t)np + 1) * 4)) ||
        (rc = dov.alloc(((size_t)nt + 1) * ((size_t)np + 1) * 8))) return rc;
    CS_CUDA(cudaMemcpyAsync(dx.p, x, px * 4, cudaMemcpyHostToDevice, 0));
    CS_CUDA(cudaMemcpyAsync(dy.p, y, px * 4, cudaMemcpyHostToDevice, 0));
    uint32_t cap = 1u << 16;
    while (cap < 8u * ((uint32_t)nt + (uint32_t)np + 2u) && cap < (1u << 30)) cap <<= 1;
    for (;;) {
        DevBuf tpg, tpp, tpc, tw;
        size_t wsb = cs::overlap_workspace_bytes(cap);
        if ((rc = tpg.alloc((size_t)cap * 4)) || (rc = tpp.alloc((size_t)cap * 4)) || (rc = tpc.alloc((size_t)cap * 4)) || (rc = tw.alloc(wsb))) return rc;
        uint32_t* n_pairs = dmisc.as<uint32_t>(); uint32_t* status = n_pairs + 1;
        rc = cs::label_overlap_sparse(dx.as<int32_t>(), W, dy.as<int32_t>(), W, H, W, nt, np, tpg.as<int32_t>(), tpp.as<int32_t>(),
                                      tpc.as<uint32_t>(), cap, n_pairs, dag.as<uint32_t>(), dap.as<uint32_t>(), status, cap, tw.p, wsb, 0);
        if (rc != CS_OK) return rc;
        uint32_t hs[2];
        CS_CUDA(cudaMemcpyAsync(hs, dmisc.p, 8, cudaMemcpyDeviceToHost, 0));
        CS_CUDA(cudaStreamSynchronize(0));
        if (hs[1] & 6u) { set_error("labels out of range (negative or above the stated maximum)"); return CS_ERR_ARG; }
        if (hs[1] & 1u) { if (cap >= (1u << 30)) { set_error("pair table overflow"); return CS_ERR_OVERFLOW; } cap <<= 2; continue; }
        rc = cs::dense_overlap(tpg.as<int32_t>(), tpp.as<int32_t>(), tpc.as<uint32_t>(), n_pairs, cap, nt, np, dov.as<unsigned long long>(), 0);
        if (rc != CS_OK) return rc;
        CS_CUDA(cudaMemcpyAsync(ov, dov.p, ((size_t)nt + 1) * ((size_t)np + 1) * 8, cudaMemcpyDeviceToHost, 0));
        CS_CUDA(cudaStreamSynchronize(0));
        return CS_OK;
    }
}

}  // extern "C"
