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

/* measure.label(mask, connectivity) alone (baseline/predict.py:113, outermost call). */
CS_API int cs_label(const uint8_t* mask, int N, int H, int W, int connectivity, int32_t* labels,
             int32_t* n_out, void* workspace, size_t workspace_bytes, cs_stream_t stream);
CS_API int cs_label_host(const uint8_t* mask, int N, int H, int W, int connectivity, int32_t* labels,
                  int32_t* n_out);

/* morphology.remove_small_objects(mask, min_size) (invert = 0) and
 * morphology.remove_small_holes(mask, area_threshold = min_size) (invert = 1),
 * both with skimage's default connectivity=1 (4-neighbourhood)
 * (baseline/predict.py:113, inner calls).  scratch: int32[N*H*W] device buffer. */
CS_API int cs_remove_small(const uint8_t* mask_in, uint8_t* mask_out, int N, int H, int W, int min_size,
                    int invert, int32_t* scratch, void* workspace, size_t workspace_bytes,
                    cs_stream_t stream);
CS_API int cs_remove_small_host(const uint8_t* mask_in, uint8_t* mask_out, int N, int H, int W,
                         int min_size, int invert);

#ifdef __cplusplus
}
#endif
#endif /* CELLSEG_B200_H */
