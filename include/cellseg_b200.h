/* cellseg_b200.h — C ABI of libcellseg_sm100.so
 *
 * B200 (sm_100a) implementation of the per-image inference glue and
 * post-processing path of JunMa11/NeurIPS-CellSeg.  The reference has no FFI of
 * its own: the boundary it offers is the set of third-party calls made by
 * baseline/predict.py:107-113 and the functions of baseline/compute_metric.py.
 * Each entry point below names the reference call it stands in for.
 *
 * Conventions
 *   - every function returns 0 on success, a negative cs_status otherwise; the
 *     text of the last failure on the calling thread is cs_last_error();
 *   - `*_dev` style functions (no suffix) take CALLER-OWNED DEVICE pointers and
 *     a cudaStream_t (passed as void*), enqueue work and return without
 *     synchronising; scratch memory is a caller-provided device workspace whose
 *     size comes from the matching *_workspace_bytes query;
 *   - `*_host` functions take HOST pointers, allocate/copy/synchronise
 *     internally and are what a ctypes/cffi stub in the reference would call;
 *   - images are row-major; batches are [N][H][W]; masks are uint8 (non-zero =
 *     foreground, i.e. numpy bool); instance maps are int32, 0 = background;
 *   - no entry point falls back to the CPU.
 */
#ifndef CELLSEG_B200_H
#define CELLSEG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cs_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define CS_API __attribute__((visibility("default")))
#else
#define CS_API
#endif

enum cs_status {
    CS_STATUS_OK = 0,
    CS_STATUS_ARG = -1,
    CS_STATUS_WORKSPACE = -2,
    CS_STATUS_CUDA = -3,
    CS_STATUS_RANGE = -4,
    CS_STATUS_OVERFLOW = -5
};

CS_API const char* cs_last_error(void);
CS_API int cs_abi_version(void);
/* Kernels launched by this library in the calling process so far (bench.py's gpu_launches evidence). */
CS_API unsigned long long cs_kernel_launch_count(void);
/* Opt-in per-kernel device timing.  cs_profile_enable(1) clears the accumulators and brackets every later
 * launch with a cudaEvent pair on its stream; cs_profile_read(i, ...) waits for the recorded events and returns
 * 1 with kernel i's name, summed duration (ms) and launch count, or 0 when i is past the last kernel seen.   */
CS_API int cs_profile_enable(int on);
CS_API int cs_profile_read(int index, char* name, size_t name_cap, double* total_ms, unsigned long long* count);

/* ---- a-5 / a-6 / a-7: mask -> instance map ---------------------------------
 * Replaces, at baseline/predict.py:113,
 *   measure.label(morphology.remove_small_objects(morphology.remove_small_holes(mask), 16))
 * hole_area: holes (4-connected background components) with fewer pixels are
 *   filled (skimage default 64); min_size: 4-connected objects with fewer pixels
 *   are removed (reference passes 16); connectivity of the final labelling is 8
 *   (skimage default for 2-D) or 4.  Label ids are 1..n in raster order of each
 *   component's first pixel, exactly as scipy.ndimage.label/skimage produce.
 * labels doubles as scratch until the last kernel; n_out (int32[N], may be NULL)
 * receives the number of instances per image.                                 */
CS_API size_t cs_postprocess_workspace_bytes(int N, int H, int W);
CS_API int cs_postprocess(const uint8_t* mask, int N, int H, int W, int hole_area, int min_size,
                   int connectivity, int32_t* labels, int32_t* n_out,
                   void* workspace, size_t workspace_bytes, cs_stream_t stream);
CS_API int cs_postprocess_host(const uint8_t* mask, int N, int H, int W, int hole_area, int min_size,
                        int connectivity, int32_t* labels, int32_t* n_out);

/* Same chain when the foreground mask is already bit-packed on device (what
 * cs_blend_softmax_threshold emits, so logits and masks never round-trip HBM as
 * bytes): cs_postprocess_packed_slot returns where, inside `workspace`, the packed
 * mask for an [N,H,W] batch must be written before cs_postprocess_packed runs.  */
CS_API uint32_t* cs_postprocess_packed_slot(void* workspace, size_t workspace_bytes, int N, int H, int W);
CS_API int cs_postprocess_packed(int N, int H, int W, int hole_area, int min_size, int connectivity,
                          int32_t* labels, int32_t* n_out, void* workspace, size_t workspace_bytes,
                          cs_stream_t stream);

/* ---- pre-normalisation (baseline/predict.py:14-21, 94-104) ---------------------------------------------------
 * normalize_channel: per channel, 1st/99th percentile of the non-zero values, skimage rescale_intensity to uint8.
 * cs_prenorm_hist builds the per-channel histograms (uint32[3][bins], bins = 256 for uint8 / 65536 for uint16 images,
 * img = [H][W][C] interleaved, C = 1 for a 2-D image); the host derives numpy's percentiles and the value -> uint8
 * table from them (predict.py prenorm_lut); cs_prenorm_apply maps the image through it into uint8 [H][W][3]
 * (a 2-D image is replicated, channels beyond the third are dropped, as predict.py:94-99 does).                     */
CS_API int cs_prenorm_hist(const void* img, int elem_bytes, int H, int W, int C, uint32_t* hist, cs_stream_t stream);
CS_API int cs_prenorm_apply(const void* img, int elem_bytes, int H, int W, int C, const uint8_t* lut, uint8_t* out, cs_stream_t stream);

/* ---- a-1: u8 image -> float in [0,1] ----------------------------------------
 * Replaces `pre_img_data / np.max(pre_img_data)` + `.type(torch.FloatTensor)`
 * (baseline/predict.py:107-108): lut[n][v] = (float)((double)v / (double)max_n),
 * max taken over the whole HxWx3 image n on device.  max_scratch: uint32[N].     */
CS_API int cs_scale_lut_u8(const uint8_t* img, int N, size_t bytes_per_image, uint32_t* max_scratch,
                    float* lut /* [N][256] */, cs_stream_t stream);
/* The two halves of cs_scale_lut_u8, for the strip-sharded path where the maximum is all-reduced between them. */
CS_API int cs_u8_max(const uint8_t* img, int N, size_t bytes_per_image, uint32_t* maxv /* [N] */, cs_stream_t stream);
CS_API int cs_scale_lut_from_max(const uint32_t* maxv, int N, float* lut /* [N][256] */, cs_stream_t stream);
CS_API int cs_u8hwc_to_f32chw(const uint8_t* img /* [N][H][W][3] */, int N, int H, int W, const float* lut,
                       float* out /* [N][3][H][W] */, cs_stream_t stream);

/* ---- predictor glue: InstanceNorm2d(affine=False) -> PReLU (-> + residual) of monai's UNet "ADN" blocks -----------
 * (baseline/predict.py:45-52 builds the UNet; its convolutions stay on cuDNN).  x, y: `planes` contiguous planes of HW
 * floats (NCHW: planes = N*C); y may alias x; residual (same shape), residual_bias (float[C], added per channel together
 * with the residual: the bias of the residual convolution) and prelu_weight (one float, device) may be NULL.          */
CS_API int cs_instnorm_prelu(const float* x, const float* residual, float* y, const float* prelu_weight, const float* residual_bias,
                      int C, long long planes, int HW, float eps, cs_stream_t stream);

/* ---- predictor glue, channels-last (NHWC) edition ---------------------------------------------------------------
 * The mid-network convolutions of the UNet stay on cuDNN and run on NHWC tensors; the pieces around them are ours.
 * Statistics are per-(sample, channel) partials float2[N][chunks][C] = (chunk mean, chunk M2 about that mean), merged by
 * the consumer (Chan's formula, double): deterministic, no atomics.
 *  cs_nhwc_stats        partials of x [N][HW][C] (C a power of two in [4,256]); chunks = cs_nhwc_stat_chunks(HW, C),
 *                       pixels per chunk = 2048 / (C/4)
 *  cs_nhwc_norm_prelu   y = prelu(instance_norm(x)) (+ residual + residual_bias[c]); partials made of chunks of
 *                       stat_chunk_pixels pixels (the last one shorter); y may alias x
 *  cs_unet_stem         monai UNet model.0: conv.unit0.conv and residual, both Conv2d(3 -> 16, k3, s2, p1) without bias,
 *                       from win [N][3][H][W] (H, W even) -> a_out, r_out [N][H/2*W/2][16] and the partials of a_out in
 *                       chunks of 256 pixels
 *  cs_unet_head_convt   model.2.0.conv = ConvTranspose2d(32 -> 3, k3, s2, p1, output_padding 1) without bias on the channel
 *                       concatenation of xa, xb [N][Hin*Win][16] (torch.cat of the skip connection never materialised)
 *                       -> z [N][3][2Hin][2Win] and its partials float2[N][cs_unet_head_partial_chunks(Hin, Win)][3]
 *                       (one per 8 x 32 input tile)
 *  cs_unet_head_final   out = conv3x3(y) + b3 + y with y = prelu(instance_norm(z)): model.2.0.adn then model.2.1 (a
 *                       conv-only ResidualUnit with identity residual); w3 [3][3][3][3], out [N][3][H][W]
 * The convolution weights of these three (w_unit0, w_res [16][3][3][3]; wt [32][3][3][3]; w3, b3) are read on the HOST
 * and travel as kernel parameters (constant-bank operands of the multiply-adds): pass host pointers.            */
CS_API int cs_nhwc_stat_chunks(int HW, int C);
CS_API int cs_nhwc_stats(const float* x, int N, int HW, int C, float* partials, cs_stream_t stream);
CS_API int cs_nhwc_norm_prelu(const float* x, const float* residual, float* y, const float* prelu_weight, const float* residual_bias,
                       int N, int HW, int C, const float* partials, int stat_chunk_pixels, float eps, cs_stream_t stream);
/* The norm block in one call.  With a `partials` workspace (float2[N][cs_nhwc_stat_chunks(HW, C)][C]) it runs cs_nhwc_stats +
 * cs_nhwc_norm_prelu, the faster pair on B200.  With partials == NULL samples of up to 1 MiB are normalised in a single pass by
 * one thread-block cluster per sample (registers of up to 8 CTAs hold the sample, per-CTA channel sums are exchanged through
 * distributed shared memory, x is read once, no workspace); larger samples then fail with CS_ERR_ARG.
 * y2 (may be NULL; needs partials): a second, channel-strided copy of the result into channels [c2_offset, c2_offset + C) of an
 * NHWC tensor with c2_total channels — this tensor's half of a skip connection's torch.cat([x, sub(x)], 1), which therefore
 * never runs; y may be NULL when only that slot is wanted.                                                                    */
CS_API int cs_instnorm_prelu_nhwc(const float* x, const float* residual, float* y, const float* prelu_weight, const float* residual_bias,
                           int N, int HW, int C, float eps, float* partials, float* y2, int c2_total, int c2_offset, cs_stream_t stream);
/* The norm block on channel SLICES of wider NHWC tensors: x = channels [x_coffset, x_coffset + C) of a tensor with x_ctotal channels,
 * the residual likewise.  Lets the two stride-2 convolutions of a down ResidualUnit (unit0 and the residual branch, same input) run
 * as ONE cuDNN convolution with 2C outputs.  partials: float2[N][cs_nhwc_stat_chunks(HW, C)][C].                               */
CS_API int cs_instnorm_prelu_nhwc_slices(const float* x, int x_ctotal, int x_coffset, const float* residual, int r_ctotal, int r_coffset, float* y,
                                  const float* prelu_weight, const float* residual_bias, int N, int HW, int C, float eps, float* partials,
                                  float* y2, int c2_total, int c2_offset, cs_stream_t stream);
/* The norm block behind ConvTranspose2d(k3, s2, p1, output_padding 1) computed as ONE regular 2x2 convolution with 4C outputs (one
 * group of C per output parity) on the input with symmetric padding 1: x [N][H+1][W+1][4C] is read from row 1 / column 1 on and
 * y [N][2H][2W][C] = prelu(instance_norm(depth_to_space(x))).  partials: float2[N][cs_nhwc_stat_chunks(H*W, 4C)][4C]; C <= 64. */
CS_API int cs_instnorm_prelu_nhwc_shuffle(const float* x, float* y, const float* prelu_weight, int N, int H, int W, int C, float eps,
                                   float* partials, cs_stream_t stream);
CS_API int cs_unet_stem(const float* win, int N, int H, int W, const float* w_unit0, const float* w_res, float* a_out, float* r_out,
                 float* a_partials, cs_stream_t stream);
/* cs_unet_stem on windows [win_begin, win_begin+win_count) of the window grid (arguments as cs_gather_windows_u8), read straight
 * from the uint8 HWC images through the scale table: a-1 and the gather of a-2 fused into the first convolutions.       */
CS_API int cs_unet_stem_u8(const uint8_t* img_hwc, const float* lut, int N, int H, int W, int R, int ny, int nx, int pad_y, int pad_x,
                    const int* ys, const int* xs, int win_begin, int win_count, const float* w_unit0, const float* w_res,
                    float* a_out, float* r_out, float* a_partials, cs_stream_t stream);
CS_API int cs_unet_head_partial_chunks(int Hin, int Win);
CS_API int cs_unet_head_convt(const float* xa, const float* xb, int N, int Hin, int Win, const float* wt, float* z, float* z_partials,
                       cs_stream_t stream);
CS_API int cs_unet_head_final(const float* z, const float* z_partials, int N, int H, int W, const float* prelu_weight, const float* w3,
                       const float* b3, float eps, float* out, cs_stream_t stream);

/* ---- three-class training target (data/pre_process_3class.py:31-56, create_interior_map) ------------------------
 * cs_interior_boundary: inst int32 [N][H][W] (compared as int16 like the reference's .astype(np.int16), :98) ->
 *   boundary = binary_dilation(find_boundaries(inst, mode='inner'), disk(1)) and interior = ~boundary & (inst > 0), both u8 0/1.
 * The caller then runs cs_remove_small(interior, min_size 16, 4-connected) and cs_interior_combine -> classes {0, 1 interior, 2 boundary}. */
CS_API int cs_interior_boundary(const int32_t* inst, int N, int H, int W, uint8_t* boundary, uint8_t* interior, cs_stream_t stream);
CS_API int cs_interior_combine(const uint8_t* boundary, const uint8_t* interior, size_t n, uint8_t* out, cs_stream_t stream);

/* ---- a-2: sliding-window gather ---------------------------------------------
 * Replaces the per-batch slicing + torch.cat inside monai.inferers.
 * sliding_window_inference (called at baseline/predict.py:109).  The window grid is
 * ny x nx windows of R x R pixels whose top-left corners in PADDED coordinates are
 * (ys[i], xs[j]); window index w = i*nx + j, batch index = n*ny*nx + w (MONAI's
 * order).  pad_y/pad_x is the low-side zero padding MONAI adds when the image is
 * smaller than the ROI.  Windows [win_begin, win_begin+win_count) are written to
 * batch[win_count][3][R][R].  The _u8 variant fuses a-1 (reads the u8 HWC image
 * through the scale LUT).                                                        */
CS_API int cs_gather_windows(const float* img_chw, int N, int H, int W, int R, int ny, int nx, int pad_y, int pad_x,
                      const int* ys, const int* xs, int win_begin, int win_count, float* batch,
                      cs_stream_t stream);
CS_API int cs_gather_windows_u8(const uint8_t* img_hwc, const float* lut, int N, int H, int W, int R, int ny, int nx,
                         int pad_y, int pad_x, const int* ys, const int* xs, int win_begin, int win_count,
                         float* batch, cs_stream_t stream);

/* ---- a-2 (blend) + a-4 (softmax, threshold) -----------------------------------
 * Replaces MONAI's per-window `out += w*logits; cnt += w`, the final `out / cnt`,
 * torch.nn.functional.softmax(dim=1) and `[0,1] > 0.5` (baseline/predict.py:109-113)
 * in one pass over the window logits win[N*ny*nx][C][R][R].  cover_y[Y] / cover_x[X]
 * (padded coordinates, lengths H+2*pad.. see host code) hold (first covering window
 * index << 8 | count) per axis.  imap = R*R importance map, NULL = constant (ones).
 * Any subset of outputs may be requested: packed (tile-major bit mask for
 * cs_postprocess_packed), mask u8[N][H][W], p1 f32[N][H][W], blended f32[N][C][H][W]. */
CS_API int cs_blend_softmax_threshold(const float* win_logits, int C, int N, int H, int W, int R, int ny, int nx,
                               int pad_y, int pad_x, const int* ys, const int* xs, const int* cover_y,
                               const int* cover_x, const float* imap, float threshold, uint32_t* packed,
                               uint8_t* mask, float* p1, float* blended, cs_stream_t stream);

/* measure.label(mask, connectivity) alone (baseline/predict.py:113, outermost call). */
CS_API int cs_label(const uint8_t* mask, int N, int H, int W, int connectivity, int32_t* labels,
             int32_t* n_out, void* workspace, size_t workspace_bytes, cs_stream_t stream);
CS_API int cs_label_host(const uint8_t* mask, int N, int H, int W, int connectivity, int32_t* labels,
                  int32_t* n_out);

/* morphology.remove_small_objects(mask, min_size) (invert = 0) and
 * morphology.remove_small_holes(mask, area_threshold = min_size) (invert = 1),
 * both with skimage's default connectivity=1 (4-neighbourhood)
 * (baseline/predict.py:113, inner calls).  workspace: cs_postprocess_workspace_bytes(N, H, W). */
CS_API int cs_remove_small(const uint8_t* mask_in, uint8_t* mask_out, int N, int H, int W, int min_size,
                    int invert, void* workspace, size_t workspace_bytes, cs_stream_t stream);
CS_API int cs_remove_small_host(const uint8_t* mask_in, uint8_t* mask_out, int N, int H, int W,
                         int min_size, int invert);

/* The certified-large pre-pass of the two size filters on its own (diagnostic): large_bg / large_fg (u8 0/1, [N][H][W]) mark
 * the background / foreground pixels proven — by a bit-parallel flood fill from solid blocks, without labelling — to lie in
 * 4-connected components of at least hole_area / min_size pixels.  Always a subset of those components' pixels; the chain
 * labels only the rest (remove_small_holes / remove_small_objects of baseline/predict.py:113 keep such components whatever
 * else they contain).  workspace: cs_postprocess_workspace_bytes(N, H, W). */
CS_API int cs_certify_large(const uint8_t* mask, int N, int H, int W, int hole_area, int min_size, uint8_t* large_bg,
                     uint8_t* large_fg, void* workspace, size_t workspace_bytes, cs_stream_t stream);

/* ---- strip-sharded post-processing (BASELINE config 3: one whole-slide image over several GPUs) ----------------
 * Each GPU holds one row strip [H rows x W] of the image and runs the same chain on it; pieces of components that
 * touch the strip's first / last row are reconciled with the neighbouring strips between phases.  The library never
 * communicates: the host moves the exported rows (torch.distributed all_gather over NCCL, wsi.py).  Phase order:
 *   cs_strip_begin
 *   for stage in (1 = holes, 2 = objects):  cs_strip_filter_export → [exchange] → cs_strip_filter_resume
 *   cs_strip_label_export → [exchange rows + counts] → cs_strip_label_resolve → cs_strip_final
 * Every strip but the last must be a multiple of 64 rows; a strip with neighbours on both sides needs at least
 * max(hole_area, min_size) rows.  The result is bit-identical to cs_postprocess on the whole image.
 *   rows of a filter stage:  int32[2][2][W]  (side 0 = first row / 1 = last row; {piece id or -1 bg / -2 large, size}; x)
 *   rows of the label stage: int32[2][W]     strip-local raster rank of the component under each pixel (0 = background) */
CS_API size_t cs_strip_workspace_bytes(int H, int W);
CS_API uint32_t* cs_strip_packed_slot(void* workspace, size_t workspace_bytes, int H, int W);
CS_API int cs_strip_begin(const uint8_t* mask_or_null, int H, int W, int hole_area, int min_size, int connectivity, int32_t* labels,
                   void* workspace, size_t workspace_bytes, cs_stream_t stream);
CS_API int cs_strip_filter_export(int stage, int H, int W, int hole_area, int min_size, int connectivity, int32_t* labels,
                           int32_t* rows_out, void* workspace, size_t workspace_bytes, cs_stream_t stream);
CS_API int cs_strip_filter_resume(int stage, int H, int W, int hole_area, int min_size, int connectivity, int32_t* labels,
                           const int32_t* above_last_row /* [2][W] or NULL */, const int32_t* below_first_row /* [2][W] or NULL */,
                           void* workspace, size_t workspace_bytes, cs_stream_t stream);
CS_API int cs_strip_label_export(int H, int W, int connectivity, int32_t* labels, int32_t* rows_out, int32_t* count_out,
                          void* workspace, size_t workspace_bytes, cs_stream_t stream);
CS_API size_t cs_strip_resolve_workspace_bytes(long long total_components);
CS_API int cs_strip_label_resolve(const int32_t* rows_all /* [nranks][2][W] */, const int32_t* offsets /* device [nranks+1] */, int nranks,
                           int W, int connectivity, int my_rank, int32_t first, int32_t count, long long total, int32_t* lut_out,
                           int32_t* total_out, void* workspace2, size_t workspace2_bytes, cs_stream_t stream);
CS_API int cs_strip_final(int H, int W, const int32_t* lut, int32_t* labels, void* workspace, size_t workspace_bytes, cs_stream_t stream);

/* ---- a-9: remove_boundary_cells + relabel_sequential ---------------------------
 * cs_label_max: out2[0] = max label, out2[1] = min label of an int32 image with row
 *   stride ld (device int32[2]); the caller reads it back to size the label tables.
 * cs_relabel, border = 2: compute_metric.remove_boundary_cells(mask)
 *   (baseline/compute_metric.py:124-133): every label with a pixel in the 2-pixel frame
 *   is dropped, survivors renumbered 1..n in ascending label order into `out`;
 *   zero_inplace != 0 also zeroes the dropped labels in `lab`, as the reference mutates
 *   its argument.  border = 0: skimage.segmentation.relabel_sequential(lab)
 *   (compute_metric.py:132,186-187).  n_out: device int32, number of surviving labels. */
CS_API int cs_label_max(const int32_t* lab, int H, int W, int ld, int32_t* out2, cs_stream_t stream);
CS_API size_t cs_relabel_workspace_bytes(int32_t max_label);
CS_API int cs_relabel(int32_t* lab, int H, int W, int ld, int border, int32_t max_label, int zero_inplace,
               int32_t* out, int ldo, int32_t* n_out, void* workspace, size_t workspace_bytes,
               cs_stream_t stream);
CS_API int cs_remove_boundary_cells_host(int32_t* mask, int H, int W, int32_t* out, int32_t* n_out);
CS_API int cs_relabel_sequential_host(const int32_t* lab, int H, int W, int32_t* out, int32_t* n_out);

/* ---- a-10 / a-11 / a-14: label overlap ------------------------------------------
 * compute_metric._label_overlap(x, y) (baseline/compute_metric.py:39-70) counts
 * ov[x[i], y[i]] += 1 into a dense uint64[(1+x.max()) x (1+y.max())] matrix (3 GB at
 * 20k x 20k cells with ~58k non-zeros).  cs_label_overlap_sparse returns only the
 * non-zero entries as (gt, pred, count) triples, in no particular order, plus the row
 * and column sums area_gt[nt+1] / area_pr[np+1] that _intersection_over_union (:33-35)
 * and dice (:72-81) need.  status (device uint32): bit0 = table or triple list
 * overflow (retry with a larger table_capacity / pair_cap), bit1 = negative label,
 * bit2 = label above nt / np.  table_capacity: power of two >= 1024.
 * cs_overlap_dense / cs_iou_dense materialise the reference's dense uint64 / float64
 * matrices from the triples (row-major [(nt+1)][(np+1)]).                          */
CS_API size_t cs_overlap_workspace_bytes(uint32_t table_capacity);
CS_API int cs_label_overlap_sparse(const int32_t* x, int ldx, const int32_t* y, int ldy, int H, int W, int32_t nt,
                            int32_t np, int32_t* pair_gt, int32_t* pair_pr, uint32_t* pair_cnt, uint32_t pair_cap,
                            uint32_t* n_pairs, uint32_t* area_gt, uint32_t* area_pr, uint32_t* status,
                            uint32_t table_capacity, void* workspace, size_t workspace_bytes, cs_stream_t stream);
CS_API int cs_overlap_dense(const int32_t* pair_gt, const int32_t* pair_pr, const uint32_t* pair_cnt, const uint32_t* n_pairs,
                     uint32_t pair_cap, int32_t nt, int32_t np, uint64_t* ov, cs_stream_t stream);
CS_API int cs_iou_dense(const int32_t* pair_gt, const int32_t* pair_pr, const uint32_t* pair_cnt, const uint32_t* n_pairs,
                 uint32_t pair_cap, int32_t nt, int32_t np, const uint32_t* area_gt, const uint32_t* area_pr,
                 double* iou, cs_stream_t stream);
/* The candidate graph of compute_metric._true_positive (baseline/compute_metric.py:83-105: `iou >= th` on the matrix of :21-37
 * without its background row / column): the triples with both labels > 0 and count / (area_gt + area_pr - count) >= threshold
 * (one correctly rounded float64 division of exact integers, as numpy computes it) are written to edge_gt / edge_pr
 * (int32[pair_cap]); deg_gt[nt+1] / deg_pr[np+1] (device scratch, zeroed by the call) count the edges per row / column;
 * out2[0] = number of edges, out2[1] = 1 when a row or column has more than one edge (only then the assignment is a real
 * matching problem — e.g. a ground-truth cell split into two exact halves at th = 0.5 — which the host solves on the edge list;
 * otherwise tp = out2[0]). */
CS_API int cs_iou_edges(const int32_t* pair_gt, const int32_t* pair_pr, const uint32_t* pair_cnt, const uint32_t* n_pairs,
                 uint32_t pair_cap, int32_t nt, int32_t np, const uint32_t* area_gt, const uint32_t* area_pr, double threshold,
                 uint32_t* deg_gt, uint32_t* deg_pr, int32_t* edge_gt, int32_t* edge_pr, uint32_t* out2, cs_stream_t stream);
/* dice(gt, seg) (baseline/compute_metric.py:72-81): out3 = {#gt != 0, #seg != 0, #both} over n bytes. */
CS_API int cs_count_nonzero3(const uint8_t* gt, const uint8_t* seg, size_t n, uint64_t* out3, cs_stream_t stream);
CS_API int cs_count_nonzero3_host(const uint8_t* gt, const uint8_t* seg, size_t n, uint64_t* out3);
CS_API int cs_label_overlap_host(const int32_t* x, const int32_t* y, int H, int W, int32_t nt, int32_t np, uint64_t* ov);

#ifdef __cplusplus
}
#endif
#endif /* CELLSEG_B200_H */
