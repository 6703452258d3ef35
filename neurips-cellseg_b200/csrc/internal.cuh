// Host-side entry points of the kernel files, shared with capi.cu (one definition of every struct).
#pragma once
#include "cs_platform.cuh"

namespace cs {

// ---- ccl.cu ----------------------------------------------------------------
size_t postprocess_workspace_bytes(int N, int H, int W);
int postprocess(const uint8_t* mask, int N, int H, int W, int hole_area, int min_size, int connectivity,
                int32_t* labels, int32_t* n_out, void* ws, size_t ws_bytes, cudaStream_t st);
int postprocess_packed(int N, int H, int W, int hole_area, int min_size, int connectivity,
                       int32_t* labels, int32_t* n_out, void* ws, size_t ws_bytes, cudaStream_t st);
uint32_t* packed_mask_slot(void* ws, size_t ws_bytes, int N, int H, int W);
int label(const uint8_t* mask, int N, int H, int W, int connectivity, int32_t* labels, int32_t* n_out,
          void* ws, size_t ws_bytes, cudaStream_t st);
int remove_small(const uint8_t* mask_in, uint8_t* mask_out, int N, int H, int W, int connectivity, int min_size, int invert,
                 void* ws, size_t ws_bytes, cudaStream_t st);

int certify_masks(const uint8_t* mask, int N, int H, int W, int hole_area, int min_size, uint8_t* large_bg, uint8_t* large_fg,
                  void* ws, size_t ws_bytes, cudaStream_t st);

// strip-sharded phases (one row strip of a whole-slide image per GPU)
size_t strip_workspace_bytes(int H, int W);
uint32_t* strip_packed_slot(void* ws, size_t ws_bytes, int H, int W);
int strip_begin(const uint8_t* mask_or_null, int H, int W, int hole_area, int min_size, int conn, int32_t* labels, void* ws, size_t ws_bytes, cudaStream_t st);
int strip_filter_export(int stage, int H, int W, int hole_area, int min_size, int conn, int32_t* labels, int32_t* rows_out,
                        void* ws, size_t ws_bytes, cudaStream_t st);
int strip_filter_resume(int stage, int H, int W, int hole_area, int min_size, int conn, int32_t* labels, const int32_t* nb_top,
                        const int32_t* nb_bottom, void* ws, size_t ws_bytes, cudaStream_t st);
int strip_label_export(int H, int W, int conn, int32_t* labels, int32_t* rows_out, int32_t* count_out, void* ws, size_t ws_bytes, cudaStream_t st);
size_t strip_resolve_workspace_bytes(long long total);
int strip_label_resolve(const int32_t* rows_all, const int32_t* offsets, int nranks, int W, int conn, int my_rank, int32_t first, int32_t count,
                        long long total, int32_t* lut_out, int32_t* total_out, void* ws2, size_t ws2_bytes, cudaStream_t st);
int strip_final(int H, int W, const int32_t* lut, int32_t* labels, void* ws, size_t ws_bytes, cudaStream_t st);

// ---- certify.cu (certified-large pre-pass of the size filters; Geo lives in ccl_tile.cuh) ----
struct Geo;
int certify_large(const uint32_t* M0, const Geo& g, int hole_area, int min_size, uint32_t* Lbg, uint32_t* Lfg, cudaStream_t st);

// ---- window.cu -------------------------------------------------------------
struct WinGeo {
    int N, H, W;       // images
    int R;             // ROI edge
    int ny, nx;        // windows per column / per row
    int pad_y, pad_x;  // low-side zero padding
};
// cover_y[Y] / cover_x[X] (padded coordinates) = (first covering window index << 8) | count along that axis.
struct BlendArgs {
    const float* win;        // [N*n_win][C][R][R] logits
    const float* imap;       // [R][R] importance map, or nullptr for the constant map (all ones)
    const int* ys; const int* xs;
    const int* cover_y; const int* cover_x;
    uint32_t* M;             // tile-major packed mask (or nullptr)
    uint8_t* mask;           // [N][H][W] u8 0/1 (or nullptr)
    float* p1;               // [N][H][W] softmax probability of class 1 (or nullptr)
    float* blended;          // [N][C][H][W] blended logits (or nullptr)
    float threshold;
};
int scale_lut_u8(const uint8_t* img, int N, size_t bytes_per_image, uint32_t* max_scratch, float* lut, cudaStream_t st);
int u8_max(const uint8_t* img, int N, size_t bytes_per_image, uint32_t* maxv, cudaStream_t st);
int scale_lut_from_max(const uint32_t* maxv, int N, float* lut, cudaStream_t st);
int u8hwc_to_f32chw(const uint8_t* img, int N, int H, int W, const float* lut, float* out, cudaStream_t st);
int gather_windows(const void* src, int from_u8, const float* lut, WinGeo g, const int* ys, const int* xs,
                   int win_begin, int win_count, float* batch, cudaStream_t st);
int blend_softmax_threshold(BlendArgs a, WinGeo g, int C, cudaStream_t st);


// ---- prenorm.cu ------------------------------------------------------------
int prenorm_hist(const void* img, int elem_bytes, int H, int W, int C, uint32_t* hist, cudaStream_t st);
int prenorm_apply(const void* img, int elem_bytes, int H, int W, int C, const uint8_t* lut, uint8_t* out, cudaStream_t st);

// ---- norm.cu ---------------------------------------------------------------
int instnorm_prelu(const float* x, const float* res, float* y, const float* prelu_w, const float* res_bias, int C, long long planes, int HW,
                   float eps, cudaStream_t st);

// ---- unet_glue.cu (channels-last predictor glue) ------------------------------
int nhwc_stat_chunks(int HW, int C);
int nhwc_stats(const float* x, int N, int HW, int C, float* partials, cudaStream_t st);
int nhwc_norm_prelu(const float* x, const float* res, float* y, const float* prelu_w, const float* res_bias, int N, int HW, int C,
                    const float* partials, int stat_chunk_pixels, float eps, cudaStream_t st);
int instnorm_prelu_nhwc(const float* x, const float* res, float* y, const float* prelu_w, const float* res_bias, int N, int HW, int C, float eps,
                        float* partials, float* y2, int c2_total, int c2_off, cudaStream_t st);
int instnorm_prelu_nhwc_slices(const float* x, int x_ctotal, int x_coff, const float* res, int r_ctotal, int r_coff, float* y, const float* prelu_w,
                               const float* res_bias, int N, int HW, int C, float eps, float* partials, float* y2, int c2_total, int c2_off,
                               cudaStream_t st);
int instnorm_prelu_nhwc_shuffle(const float* x, float* y, const float* prelu_w, int N, int H, int W, int C, float eps, float* partials, cudaStream_t st);
int unet_stem(const float* win, int N, int H, int W, const float* w_unit0, const float* w_res, float* a_out, float* r_out, float* a_partials, cudaStream_t st);
int unet_stem_u8(const uint8_t* img, const float* lut, WinGeo g, const int* ys, const int* xs, int win_begin, int win_count,
                 const float* w_unit0_host, const float* w_res_host, float* a_out, float* r_out, float* a_partials, cudaStream_t st);
int unet_head_partial_chunks(int Hin, int Win);
int unet_head_convt(const float* xa, const float* xb, int N, int Hin, int Win, const float* wt, float* z, float* z_partials, cudaStream_t st);
int unet_head_final(const float* z, const float* z_partials, int N, int H, int W, const float* prelu_w, const float* w3, const float* b3, float eps,
                    float* out, cudaStream_t st);

// ---- interior.cu (three-class training target, data/pre_process_3class.py) ----
int interior_boundary(const int32_t* inst, int N, int H, int W, uint8_t* boundary, uint8_t* interior, cudaStream_t st);
int interior_combine(const uint8_t* boundary, const uint8_t* interior, size_t n, uint8_t* out, cudaStream_t st);

// ---- metric.cu -------------------------------------------------------------
int label_max(const int32_t* lab, int H, int W, int ld, int32_t* out2, cudaStream_t st);
size_t relabel_workspace_bytes(int32_t max_label);
int relabel(int32_t* lab, int H, int W, int ld, int border, int32_t max_label, int zero_inplace,
            int32_t* out, int ldo, int32_t* n_out, void* ws, size_t ws_bytes, cudaStream_t st);
int count3(const uint8_t* a, const uint8_t* b, size_t n, unsigned long long* out3, cudaStream_t st);
size_t overlap_workspace_bytes(uint32_t table_capacity);
int label_overlap_sparse(const int32_t* x, int ldx, const int32_t* y, int ldy, int H, int W, int32_t nt, int32_t np_,
                         int32_t* pair_gt, int32_t* pair_pr, uint32_t* pair_cnt, uint32_t pair_cap, uint32_t* n_pairs,
                         uint32_t* area_gt, uint32_t* area_pr, uint32_t* status_out, uint32_t table_capacity,
                         void* ws, size_t ws_bytes, cudaStream_t st);
int dense_overlap(const int32_t* pg, const int32_t* pp, const uint32_t* pc, const uint32_t* n_pairs, uint32_t pair_cap, int32_t nt, int32_t np_,
                  unsigned long long* ov, cudaStream_t st);
int dense_iou(const int32_t* pg, const int32_t* pp, const uint32_t* pc, const uint32_t* n_pairs, uint32_t pair_cap, int32_t nt, int32_t np_,
              const uint32_t* area_gt, const uint32_t* area_pr, double* iou, cudaStream_t st);
int iou_edges(const int32_t* pg, const int32_t* pp, const uint32_t* pc, const uint32_t* n_pairs, uint32_t pair_cap, int32_t nt, int32_t np_,
              const uint32_t* area_gt, const uint32_t* area_pr, double th, uint32_t* deg_gt, uint32_t* deg_pr, int32_t* edge_gt, int32_t* edge_pr,
              uint32_t* out, cudaStream_t st);

}  // namespace cs
