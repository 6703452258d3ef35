// F1-metric kernels (baseline/compute_metric.py of the reference):
//   K7  remove_boundary_cells (:124-133) + skimage relabel_sequential (:132,186-187)
//         → border/present flags per label, prefix-count LUT, one remap pass
//   K6  _label_overlap (:39-70): ov[x[i], y[i]] += 1 — as a SPARSE pair histogram: per-CTA
//         shared-memory hash (privatised histogram), run-length compressed inserts, flushed into a
//         global open-addressing table; compaction emits (gt, pred, count) triples and the row/column
//         sums that _intersection_over_union (:21-37) and dice (:72-81) need.
//       The dense uint64 matrix of the reference's signature is materialised from the triples on request.
#include "internal.cuh"
#include <algorithm>

namespace cs {

// ---------------------------------------------------------------------------
// max label
// ---------------------------------------------------------------------------
// Row-oriented access for the label kernels below: grid (ceil(W / (4 * LAB_T)), rows), a thread owns 4 consecutive pixels of a row —
// one 16-byte load when the row pitch and the base address allow it (VEC) — and walks the rows with stride gridDim.y.  (The first
// versions were flat grid-stride loops with a 64-bit division per pixel and 4-byte loads: 0.1-1.1 TB/s.)
constexpr int LAB_T = 256;
template <bool VEC>
__device__ __forceinline__ int4 load_quad(const int32_t* row, int x, int W, int32_t fill) {
    if (VEC && x + 3 < W) return __ldg(reinterpret_cast<const int4*>(row + x));
    int4 v = make_int4(fill, fill, fill, fill);
    if (x < W) v.x = __ldg(row + x);
    if (x + 1 < W) v.y = __ldg(row + x + 1);
    if (x + 2 < W) v.z = __ldg(row + x + 2);
    if (x + 3 < W) v.w = __ldg(row + x + 3);
    return v;
}
// ~8 CTAs per SM, each walking several rows (one CTA per row meant 131 000 warps queueing on k_label_max's two atomics)
static dim3 lab_grid(int H, int W) {
    const int gx = ceil_div(ceil_div(W, 4), LAB_T);
    const int gy = std::max(1, std::min(H, (148 * 8 + gx - 1) / gx));
    return dim3((unsigned)gx, (unsigned)gy);
}
static bool lab_vec(const void* p, int ld) { return ((uintptr_t)p & 15) == 0 && (ld & 3) == 0; }

template <bool VEC>
__global__ void __launch_bounds__(LAB_T) k_label_max_t(const int32_t* __restrict__ lab, int H, int W, int ld, int32_t* __restrict__ out /* [2]: max, min */) {
    int32_t mx = INT32_MIN, mn = INT32_MAX;
    const int x = (blockIdx.x * LAB_T + threadIdx.x) * 4;
    if (x < W) {
        for (int y = blockIdx.y; y < H; y += gridDim.y) {
            const int4 v = load_quad<VEC>(lab + (size_t)y * ld, x, W, INT32_MIN);
            mx = max(max(mx, max(v.x, v.y)), max(v.z, v.w));
            const int4 u = make_int4(v.x, x + 1 < W ? v.y : v.x, x + 2 < W ? v.z : v.x, x + 3 < W ? v.w : v.x);
            mn = min(min(mn, min(u.x, u.y)), min(u.z, u.w));
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    }
    __shared__ int32_t smx[LAB_T / 32], smn[LAB_T / 32];
    if ((threadIdx.x & 31) == 0) { smx[threadIdx.x >> 5] = mx; smn[threadIdx.x >> 5] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < LAB_T / 32; ++w) { mx = max(mx, smx[w]); mn = min(mn, smn[w]); }
        if (mx >= mn) { atomicMax(&out[0], mx); atomicMin(&out[1], mn); }      // one pair of atomics per CTA
    }
}
__global__ void k_init_maxmin(int32_t* out) { out[0] = INT32_MIN; out[1] = INT32_MAX; }

// ---------------------------------------------------------------------------
// K7: flags → LUT → remap
// ---------------------------------------------------------------------------
// flags[l]: bit0 = label present, bit1 = label touches the `border`-pixel frame.
// Stores are plain (all writers write the same bits per pass) and skipped when the left neighbour
// carries the same label, which removes most of them.
// Background (label 0) is the one value millions of pixels share: it is noted per CTA and stored once, not once per run.
template <bool VEC>
__global__ void __launch_bounds__(LAB_T) k_label_flags_t(const int32_t* __restrict__ lab, int H, int W, int ld, int border, uint8_t* __restrict__ present,
                                                       uint8_t* __restrict__ onborder) {
    const int x = (blockIdx.x * LAB_T + threadIdx.x) * 4;
    bool zero_seen = false;
    if (x < W) {
        for (int y = blockIdx.y; y < H; y += gridDim.y) {
            const int32_t* row = lab + (size_t)y * ld;
            const int4 q = load_quad<VEC>(row, x, W, -1);
            const int32_t v[4] = {q.x, q.y, q.z, q.w};
            int32_t left = x > 0 ? __ldg(row + x - 1) : -1;
            const bool yfr = border > 0 && (y < border || y >= H - border);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int32_t c = v[j];
                if (c >= 0 && x + j < W) {
                    if (c == 0) zero_seen = true;
                    else if (c != left) present[c] = 1;
                    if (c != 0 && border > 0 && (yfr || x + j < border || x + j >= W - border)) onborder[c] = 1;
                }
                left = c;
            }
        }
    }
    if (__syncthreads_or(zero_seen) && threadIdx.x == 0 && present[0] == 0) present[0] = 1;      // (onborder[0] is never read)
}
// The reference drops `np.unique(mask*bd)[1:]`, i.e. it always spares the SMALLEST value of mask*bd.
// That value is 0 whenever the image has an interior or a background pixel; for a degenerate image that
// is all frame and has no background the smallest label survives.  Reproduce that corner.
__global__ void k_border_quirk(int H, int W, int border, int32_t nlab, const uint8_t* __restrict__ present, uint8_t* __restrict__ onborder) {
    if (blockIdx.x || threadIdx.x) return;
    const bool interior = H > 2 * border && W > 2 * border;
    if (interior || present[0]) return;
    for (int32_t l = 1; l < nlab; ++l) if (present[l]) { onborder[l] = 0; return; }
}

constexpr int MS_T = 256, MS_PER = 8, MS_CHUNK = MS_T * MS_PER;

__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, uint32_t* total_out) {
    __shared__ uint32_t ws[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    uint32_t inc = v;
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) ws[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < nw ? ws[lane] : 0u, si = s;
        for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, si, d); if (lane >= d) si += o; }
        ws[lane] = si - s;
        if (lane == 31) *total_out = si;
    }
    __syncthreads();
    uint32_t r = inc - v + ws[wid];
    __syncthreads();
    return r;
}
__device__ __forceinline__ bool keep_label(int32_t l, const uint8_t* present, const uint8_t* onborder) {
    return l > 0 && present[l] && !onborder[l];
}
__global__ void __launch_bounds__(MS_T) k_keep_reduce(int32_t nlab, const uint8_t* __restrict__ present, const uint8_t* __restrict__ onborder,
                                                      uint32_t* __restrict__ blocksum) {
    __shared__ uint32_t total;
    int32_t b0 = blockIdx.x * MS_CHUNK + threadIdx.x * MS_PER;
    uint32_t s = 0;
    for (int i = 0; i < MS_PER; ++i) if (b0 + i < nlab && keep_label(b0 + i, present, onborder)) ++s;
    block_scan_excl(s, &total);
    if (threadIdx.x == 0) blocksum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) k_keep_blocksums(uint32_t* __restrict__ blocksum, int nblocks, int32_t* __restrict__ n_out) {
    __shared__ uint32_t total;
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += 1024) {
        int i = b0 + threadIdx.x;
        uint32_t v = i < nblocks ? blocksum[i] : 0u;
        uint32_t ex = block_scan_excl(v, &total);
        if (i < nblocks) blocksum[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = (int32_t)carry;
}
__global__ void __launch_bounds__(MS_T) k_keep_lut(int32_t nlab, const uint8_t* __restrict__ present, const uint8_t* __restrict__ onborder,
                                                   const uint32_t* __restrict__ blocksum, int32_t offset, int32_t* __restrict__ lut) {
    __shared__ uint32_t total;
    int32_t b0 = blockIdx.x * MS_CHUNK + threadIdx.x * MS_PER;
    bool k[MS_PER]; uint32_t s = 0;
    for (int i = 0; i < MS_PER; ++i) { k[i] = b0 + i < nlab && keep_label(b0 + i, present, onborder); s += k[i] ? 1u : 0u; }
    uint32_t ex = block_scan_excl(s, &total) + blocksum[blockIdx.x];
    for (int i = 0; i < MS_PER; ++i) {
        if (b0 + i < nlab) lut[b0 + i] = k[i] ? (int32_t)(offset + ex) : 0;
        ex += k[i] ? 1u : 0u;
    }
}
// out = lut[in]; optionally zero the removed (frame-touching) labels in the input in place, as the
// reference's remove_boundary_cells mutates its argument.
template <bool VEC>
__global__ void __launch_bounds__(LAB_T) k_remap_t(int32_t* __restrict__ lab, int H, int W, int ld, const int32_t* __restrict__ lut,
                                                 const uint8_t* __restrict__ onborder, int zero_inplace, int32_t* __restrict__ out, int ldo, int vec_out) {
    const int x = (blockIdx.x * LAB_T + threadIdx.x) * 4;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        int32_t* row = lab + (size_t)y * ld;
        const int4 q = load_quad<VEC>(row, x, W, -1);
        const int32_t v[4] = {q.x, q.y, q.z, q.w};
        int32_t nv[4];
        bool drop[4];
        int32_t prev = -1, prev_nv = -1;
        bool prev_drop = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {                  // runs of one label (the rule inside a cell and in the background) look up once
            if (v[j] == prev) { nv[j] = prev_nv; drop[j] = prev_drop; }
            else {
                nv[j] = v[j] >= 0 ? __ldg(lut + v[j]) : v[j];
                drop[j] = zero_inplace && v[j] > 0 && onborder[v[j]];
                prev = v[j]; prev_nv = nv[j]; prev_drop = drop[j];
            }
        }
        int32_t* orow = out + (size_t)y * ldo;
        if (vec_out && x + 3 < W) *reinterpret_cast<int4*>(orow + x) = make_int4(nv[0], nv[1], nv[2], nv[3]);
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (x + j < W) orow[x + j] = nv[j];
        }
        if (zero_inplace) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (x + j < W && drop[j]) row[x + j] = 0;
        }
    }
}

static int flat_blocks_m(size_t work, int threads, int cap = 148 * 16) {
    size_t b = (work + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > (size_t)cap) b = cap;
    return (int)b;
}

int label_max(const int32_t* lab, int H, int W, int ld, int32_t* out2, cudaStream_t st) {
    CS_CHECK_ARG(lab && out2 && H > 0 && W > 0 && ld >= W, "bad label image");
    CS_LAUNCH(k_init_maxmin, 1, 1, 0, st, out2);
    if (lab_vec(lab, ld)) { constexpr auto k_label_max = k_label_max_t<true>; CS_LAUNCH(k_label_max, lab_grid(H, W), LAB_T, 0, st, lab, H, W, ld, out2); }
    else { constexpr auto k_label_max_scalar = k_label_max_t<false>; CS_LAUNCH(k_label_max_scalar, lab_grid(H, W), LAB_T, 0, st, lab, H, W, ld, out2); }
    CS_LAUNCH_CHECK();
    return CS_OK;
}

size_t relabel_workspace_bytes(int32_t max_label) {
    if (max_label < 0) return 0;
    Arena a(nullptr, 0);
    size_t nlab = (size_t)max_label + 1;
    a.take<uint8_t>(nlab); a.take<uint8_t>(nlab); a.take<int32_t>(nlab);
    a.take<uint32_t>(nlab / MS_CHUNK + 2);
    return a.off;
}

// border = 0: skimage.segmentation.relabel_sequential(lab) (offset 1).
// border = 2: compute_metric.remove_boundary_cells(lab): labels touching the 2-pixel frame are
//             dropped, survivors renumbered 1..n in ascending label order.
int relabel(int32_t* lab, int H, int W, int ld, int border, int32_t max_label, int zero_inplace,
            int32_t* out, int ldo, int32_t* n_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    CS_CHECK_ARG(lab && out && n_out && H > 0 && W > 0 && ld >= W && ldo >= W, "bad label image");
    CS_CHECK_ARG(border >= 0 && max_label >= 0, "negative border or max_label");
    Arena a(ws, ws_bytes);
    const size_t nlab = (size_t)max_label + 1;
    uint8_t* present = a.take<uint8_t>(nlab);
    uint8_t* onborder = a.take<uint8_t>(nlab);
    int32_t* lut = a.take<int32_t>(nlab);
    const int nblk = (int)((nlab + MS_CHUNK - 1) / MS_CHUNK);
    uint32_t* blocksum = a.take<uint32_t>((size_t)nblk + 1);
    if (!ws || !a.ok()) { set_error("relabel workspace too small: need %zu bytes, got %zu", a.off, ws_bytes); return CS_ERR_WORKSPACE; }
    CS_CUDA(cudaMemsetAsync(present, 0, (size_t)((uint8_t*)lut - present), st));
    if (lab_vec(lab, ld)) { constexpr auto k_label_flags = k_label_flags_t<true>; CS_LAUNCH(k_label_flags, lab_grid(H, W), LAB_T, 0, st, lab, H, W, ld, border, present, onborder); }
    else { constexpr auto k_label_flags_scalar = k_label_flags_t<false>; CS_LAUNCH(k_label_flags_scalar, lab_grid(H, W), LAB_T, 0, st, lab, H, W, ld, border, present, onborder); }
    CS_LAUNCH_CHECK();
    if (border > 0) { CS_LAUNCH(k_border_quirk, 1, 32, 0, st, H, W, border, (int32_t)nlab, present, onborder); CS_LAUNCH_CHECK(); }
    CS_LAUNCH(k_keep_reduce, nblk, MS_T, 0, st, (int32_t)nlab, present, onborder, blocksum);
    CS_LAUNCH_CHECK();
    CS_LAUNCH(k_keep_blocksums, 1, 1024, 0, st, blocksum, nblk, n_out);
    CS_LAUNCH_CHECK();
    CS_LAUNCH(k_keep_lut, nblk, MS_T, 0, st, (int32_t)nlab, present, onborder, blocksum, 1, lut);
    CS_LAUNCH_CHECK();
    const int vec_out = lab_vec(out, ldo) ? 1 : 0;
    if (lab_vec(lab, ld)) { constexpr auto k_remap = k_remap_t<true>; CS_LAUNCH(k_remap, lab_grid(H, W), LAB_T, 0, st, lab, H, W, ld, lut, onborder, zero_inplace, out, ldo, vec_out); }
    else { constexpr auto k_remap_scalar = k_remap_t<false>; CS_LAUNCH(k_remap_scalar, lab_grid(H, W), LAB_T, 0, st, lab, H, W, ld, lut, onborder, zero_inplace, out, ldo, vec_out); }
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// dice (compute_metric.py:72-81): counts of non-zero gt, non-zero seg and both, on u8/bool images.
// non-zero bytes of a 32-bit word: per byte (b | (b + 0x7f)) has its top bit set iff b != 0
__device__ __forceinline__ uint32_t nz_mask4(uint32_t w) { return ((w & 0x7f7f7f7fu) + 0x7f7f7f7fu | w) & 0x80808080u; }
__global__ void __launch_bounds__(256) k_count3(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n, unsigned long long* __restrict__ out3) {
    unsigned long long ca = 0, cb = 0, cab = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t done = 0;
    if ((((uintptr_t)a | (uintptr_t)b) & 15) == 0) {           // 16 pixels per load
        const size_t nv = n / 16;
        const uint4* av = reinterpret_cast<const uint4*>(a);
        const uint4* bv = reinterpret_cast<const uint4*>(b);
        uint32_t sa = 0, sb = 0, sab = 0;
        for (size_t i = tid; i < nv; i += stride) {
            const uint4 x = __ldg(av + i), y = __ldg(bv + i);
            const uint32_t xa[4] = {x.x, x.y, x.z, x.w}, ya[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t ma = nz_mask4(xa[k]), mb = nz_mask4(ya[k]);
                sa += __popc(ma); sb += __popc(mb); sab += __popc(ma & mb);
            }
        }
        ca = sa; cb = sb; cab = sab;                             // a thread sees < 2^32 pixels
        done = nv * 16;
    }
    for (size_t i = done + tid; i < n; i += stride) {
        bool x = a[i] != 0, y = b[i] != 0;
        ca += x; cb += y; cab += (x && y);
    }
    for (int d = 16; d > 0; d >>= 1) {
        ca += __shfl_xor_sync(0xffffffffu, ca, d); cb += __shfl_xor_sync(0xffffffffu, cb, d); cab += __shfl_xor_sync(0xffffffffu, cab, d);
    }
    if ((threadIdx.x & 31) == 0 && (ca | cb)) { atomicAdd(&out3[0], ca); atomicAdd(&out3[1], cb); atomicAdd(&out3[2], cab); }
}
int count3(const uint8_t* a, const uint8_t* b, size_t n, unsigned long long* out3, cudaStream_t st) {
    CS_CHECK_ARG(a && b && out3, "null pointer");
    CS_CUDA(cudaMemsetAsync(out3, 0, 24, st));
    if (n) { CS_LAUNCH(k_count3, flat_blocks_m((n + 15) / 16, 256, 148 * 8), 256, 0, st, a, b, n, out3); CS_LAUNCH_CHECK(); }
    return CS_OK;
}

// ---------------------------------------------------------------------------
// K6: sparse (gt, pred) pair histogram
// ---------------------------------------------------------------------------
constexpr unsigned long long EMPTY_KEY = 0xffffffffffffffffull;
constexpr int OV_T = 256;           // threads per CTA
constexpr int OV_PX = 16;           // consecutive pixels per thread
constexpr int OV_TILE = 64;         // tile edge: OV_TILE * OV_TILE / OV_PX = OV_T chunks
constexpr int OV_SLOTS = 1024;      // shared hash slots per CTA (keys 8 KB + counts 4 KB)

// 32-bit mix of the two labels (the 64-bit finaliser this replaces cost two 64-bit multiplications per insert)
__device__ __forceinline__ uint32_t hash64(unsigned long long k) {
    uint32_t h = (uint32_t)(k >> 32) * 0x9e3779b1u ^ (uint32_t)k * 0x85ebca77u;
    h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 13;
    return h;
}
__device__ __forceinline__ bool table_add(unsigned long long* keys, uint32_t* counts, uint32_t mask, unsigned long long key, uint32_t c,
                                          uint32_t max_probe) {
    uint32_t h = hash64(key) & mask;
    for (uint32_t p = 0; p <= max_probe; ++p) {
        unsigned long long cur = keys[h];
        if (cur != key) {
            if (cur != EMPTY_KEY) { h = (h + 1) & mask; continue; }
            cur = atomicCAS(&keys[h], EMPTY_KEY, key);
            if (cur != EMPTY_KEY && cur != key) { h = (h + 1) & mask; continue; }
        }
        atomicAdd(&counts[h], c);
        return true;
    }
    return false;
}

__device__ __noinline__ void flush_run(unsigned long long* skeys, uint32_t* scounts, unsigned long long* gkeys, uint32_t* gcounts, uint32_t gmask,
                                       unsigned long long key, uint32_t run, uint32_t* status) {
    if (!table_add(skeys, scounts, OV_SLOTS - 1, key, run, 24))
        if (!table_add(gkeys, gcounts, gmask, key, run, gmask)) atomicOr(status, 1u);
}

__global__ void __launch_bounds__(OV_T)
k_overlap_sparse(const int32_t* __restrict__ x, int ldx, const int32_t* __restrict__ y, int ldy, int H, int W,
                 unsigned long long* __restrict__ gkeys, uint32_t* __restrict__ gcounts, uint32_t gmask, uint32_t* __restrict__ status) {
    __shared__ unsigned long long skeys[OV_SLOTS];
    __shared__ uint32_t scounts[OV_SLOTS];
    for (int i = threadIdx.x; i < OV_SLOTS; i += OV_T) { skeys[i] = EMPTY_KEY; scounts[i] = 0u; }
    __syncthreads();
    // Work is dealt out in 64 x 64-pixel tiles (256 chunks of 16 pixels: one per thread), a contiguous run of tiles per CTA: a cell then
    // lands in the shared tables of one or two CTAs instead of one per image row it crosses, and the flush into the global table — the
    // expensive part: random L2 atomics — shrinks accordingly (with row-major chunks every CTA met ~500 different cells).
    // tile = th rows x tw columns with th * tw = 4096 pixels; th = 64 unless the map is flatter (the reference-named functions pass
    // ravelled 1-D arrays: one long row)
    int th = OV_TILE;
    while (th > 1 && th > H) th >>= 1;
    const int cpr = OV_T / th, tw = cpr * OV_PX;        // chunks per tile row, tile width
    const int tiles_x = (W + tw - 1) / tw, tiles_y = (H + th - 1) / th;
    const long long ntiles = (long long)tiles_x * tiles_y;
    const long long per_cta = (ntiles + gridDim.x - 1) / gridDim.x;
    const long long t_begin = (long long)blockIdx.x * per_cta, t_end = t_begin + per_cta < ntiles ? t_begin + per_cta : ntiles;
    const bool vec = ((((uintptr_t)x | (uintptr_t)y) & 15) == 0) && ((ldx | ldy) & 3) == 0;
    uint32_t zeros = 0;                                // (0, 0) pixels of this thread: < 2^32 by the chunk count per thread
    const int trow = threadIdx.x / cpr, tcol = (threadIdx.x % cpr) * OV_PX;
    for (long long tile = t_begin; tile < t_end; ++tile) {
        const int ty = (int)(tile / tiles_x), tx = (int)(tile - (long long)ty * tiles_x);
        const int row = ty * th + trow, c0 = tx * tw + tcol;
        if (row >= H || c0 >= W) continue;
        const int lim = min(OV_PX, W - c0);
        const int32_t* xr = x + (size_t)row * ldx + c0;
        const int32_t* yr = y + (size_t)row * ldy + c0;
        unsigned long long cur = EMPTY_KEY; uint32_t run = 0;
        // one 16-byte load per map and 4 pixels; the loop over the groups stays rolled (fully unrolled, sixteen inlined copies of the
        // two hash insertions blew the instruction cache: 0.17 -> 1.04 ms)
#pragma unroll 1
        for (int q = 0; q < OV_PX / 4; ++q) {
            const int i0 = 4 * q;
            if (i0 >= lim) break;
            int32_t av[4], bv[4];
            if (vec && i0 + 3 < lim) {
                const int4 u = __ldg(reinterpret_cast<const int4*>(xr) + q), v = __ldg(reinterpret_cast<const int4*>(yr) + q);
                av[0] = u.x; av[1] = u.y; av[2] = u.z; av[3] = u.w;
                bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { av[j] = i0 + j < lim ? xr[i0 + j] : -1; bv[j] = i0 + j < lim ? yr[i0 + j] : -1; }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (i0 + j >= lim) break;
                const int32_t a = av[j], b = bv[j];
                if (a < 0 || b < 0) { atomicOr(status, 2u); continue; }
                const unsigned long long key = ((unsigned long long)(uint32_t)a << 32) | (uint32_t)b;
                if (key == 0ull) { ++zeros; continue; }       // background over background: most pixels; counted in a register
                if (key == cur) { ++run; continue; }
                if (run) flush_run(skeys, scounts, gkeys, gcounts, gmask, cur, run, status);
                cur = key; run = 1;
            }
        }
        if (run) flush_run(skeys, scounts, gkeys, gcounts, gmask, cur, run, status);
    }
    for (int d = 16; d > 0; d >>= 1) zeros += __shfl_xor_sync(0xffffffffu, zeros, d);
    if ((threadIdx.x & 31) == 0 && zeros && !table_add(skeys, scounts, OV_SLOTS - 1, 0ull, zeros, 24))
        if (!table_add(gkeys, gcounts, gmask, 0ull, zeros, gmask)) atomicOr(status, 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < OV_SLOTS; i += OV_T) {
        unsigned long long k = skeys[i];
        if (k != EMPTY_KEY && !table_add(gkeys, gcounts, gmask, k, scounts[i], gmask)) atomicOr(status, 1u);
    }
}

__global__ void k_fill_keys(unsigned long long* keys, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) keys[i] = EMPTY_KEY;
}

// Compact the table into triples and accumulate row/column sums (n_pixels_true / n_pixels_pred).
__global__ void k_overlap_compact(const unsigned long long* __restrict__ gkeys, const uint32_t* __restrict__ gcounts, size_t cap,
                                  int32_t nt, int32_t np_, int32_t* __restrict__ pair_gt, int32_t* __restrict__ pair_pr,
                                  uint32_t* __restrict__ pair_cnt, uint32_t pair_cap, uint32_t* __restrict__ n_pairs,
                                  uint32_t* __restrict__ area_gt, uint32_t* __restrict__ area_pr, uint32_t* __restrict__ status) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long k = gkeys[i];
        if (k == EMPTY_KEY) continue;
        int32_t a = (int32_t)(k >> 32), b = (int32_t)(k & 0xffffffffu);
        uint32_t c = gcounts[i];
        if (a > nt || b > np_) { atomicOr(status, 4u); continue; }
        atomicAdd(&area_gt[a], c); atomicAdd(&area_pr[b], c);
        uint32_t slot = atomicAdd(n_pairs, 1u);
        if (slot < pair_cap) { pair_gt[slot] = a; pair_pr[slot] = b; pair_cnt[slot] = c; }
        else atomicOr(status, 1u);
    }
}

size_t overlap_workspace_bytes(uint32_t table_capacity) {
    Arena a(nullptr, 0);
    a.take<unsigned long long>(table_capacity); a.take<uint32_t>(table_capacity); a.take<uint32_t>(64);
    return a.off;
}

// x = ground truth labels (0..nt), y = predicted labels (0..np).  table_capacity must be a power of two;
// status_out (device uint32): bit0 table/pair list overflow, bit1 negative label seen, bit2 label above nt/np.
int label_overlap_sparse(const int32_t* x, int ldx, const int32_t* y, int ldy, int H, int W, int32_t nt, int32_t np_,
                         int32_t* pair_gt, int32_t* pair_pr, uint32_t* pair_cnt, uint32_t pair_cap, uint32_t* n_pairs,
                         uint32_t* area_gt, uint32_t* area_pr, uint32_t* status_out, uint32_t table_capacity,
                         void* ws, size_t ws_bytes, cudaStream_t st) {
    CS_CHECK_ARG(x && y && pair_gt && pair_pr && pair_cnt && n_pairs && area_gt && area_pr && status_out, "null pointer");
    CS_CHECK_ARG(H > 0 && W > 0 && ldx >= W && ldy >= W && nt >= 0 && np_ >= 0, "bad geometry");
    CS_CHECK_ARG(table_capacity >= 1024 && (table_capacity & (table_capacity - 1)) == 0, "table_capacity must be a power of two >= 1024");
    Arena a(ws, ws_bytes);
    unsigned long long* gkeys = a.take<unsigned long long>(table_capacity);
    uint32_t* gcounts = a.take<uint32_t>(table_capacity);
    if (!ws || !a.ok()) { set_error("overlap workspace too small: need %zu bytes, got %zu", a.off, ws_bytes); return CS_ERR_WORKSPACE; }
    CS_LAUNCH(k_fill_keys, flat_blocks_m(table_capacity, 256), 256, 0, st, gkeys, (size_t)table_capacity);
    CS_CUDA(cudaMemsetAsync(gcounts, 0, (size_t)table_capacity * 4, st));
    CS_CUDA(cudaMemsetAsync(n_pairs, 0, 4, st));
    CS_CUDA(cudaMemsetAsync(status_out, 0, 4, st));
    CS_CUDA(cudaMemsetAsync(area_gt, 0, ((size_t)nt + 1) * 4, st));
    CS_CUDA(cudaMemsetAsync(area_pr, 0, ((size_t)np_ + 1) * 4, st));
    static_assert(OV_TILE * OV_TILE / OV_PX == OV_T, "one chunk per thread and tile");
    int th = OV_TILE;
    while (th > 1 && th > H) th >>= 1;                 // as in the kernel
    const int tw = OV_T / th * OV_PX;
    const size_t ntiles = (size_t)((H + th - 1) / th) * ((W + tw - 1) / tw);
    CS_LAUNCH(k_overlap_sparse, flat_blocks_m(ntiles, 1, 148 * 8), OV_T, 0, st, x, ldx, y, ldy, H, W, gkeys, gcounts,
              table_capacity - 1, status_out);
    CS_LAUNCH_CHECK();
    CS_LAUNCH(k_overlap_compact, flat_blocks_m(table_capacity, 256), 256, 0, st, gkeys, gcounts, (size_t)table_capacity, nt, np_,
              pair_gt, pair_pr, pair_cnt, pair_cap, n_pairs, area_gt, area_pr, status_out);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// Dense forms of the reference's signatures, materialised from the triples.
__global__ void k_dense_overlap(const int32_t* __restrict__ pg, const int32_t* __restrict__ pp, const uint32_t* __restrict__ pc,
                                const uint32_t* __restrict__ n_pairs, int32_t np_, unsigned long long* __restrict__ ov) {
    const uint32_t n = *n_pairs;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        ov[(size_t)pg[i] * ((size_t)np_ + 1) + pp[i]] = pc[i];
}
__global__ void k_dense_iou(const int32_t* __restrict__ pg, const int32_t* __restrict__ pp, const uint32_t* __restrict__ pc,
                            const uint32_t* __restrict__ n_pairs, int32_t np_, const uint32_t* __restrict__ area_gt,
                            const uint32_t* __restrict__ area_pr, double* __restrict__ iou) {
    const uint32_t n = *n_pairs;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned long long c = pc[i];
        unsigned long long den = (unsigned long long)area_gt[pg[i]] + area_pr[pp[i]] - c;
        iou[(size_t)pg[i] * ((size_t)np_ + 1) + pp[i]] = __ddiv_rn(__ull2double_rn(c), __ull2double_rn(den));
    }
}

int dense_overlap(const int32_t* pg, const int32_t* pp, const uint32_t* pc, const uint32_t* n_pairs, uint32_t pair_cap, int32_t nt, int32_t np_,
                  unsigned long long* ov, cudaStream_t st) {
    CS_CHECK_ARG(pg && pp && pc && n_pairs && ov && nt >= 0 && np_ >= 0, "null pointer");
    CS_CUDA(cudaMemsetAsync(ov, 0, ((size_t)nt + 1) * ((size_t)np_ + 1) * 8, st));
    CS_LAUNCH(k_dense_overlap, flat_blocks_m(pair_cap, 256), 256, 0, st, pg, pp, pc, n_pairs, np_, ov);
    CS_LAUNCH_CHECK();
    return CS_OK;
}
int dense_iou(const int32_t* pg, const int32_t* pp, const uint32_t* pc, const uint32_t* n_pairs, uint32_t pair_cap, int32_t nt, int32_t np_,
              const uint32_t* area_gt, const uint32_t* area_pr, double* iou, cudaStream_t st) {
    CS_CHECK_ARG(pg && pp && pc && n_pairs && iou && area_gt && area_pr && nt >= 0 && np_ >= 0, "null pointer");
    CS_CUDA(cudaMemsetAsync(iou, 0, ((size_t)nt + 1) * ((size_t)np_ + 1) * 8, st));   // 0/0 → NaN → 0 in the reference (:36)
    CS_LAUNCH(k_dense_iou, flat_blocks_m(pair_cap, 256), 256, 0, st, pg, pp, pc, n_pairs, np_, area_gt, area_pr, iou);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// _true_positive's candidate graph on the device (compute_metric.py:83-105 after :21-37): the pairs with both labels > 0 and
// IoU = n / (area_gt + area_pr - n) >= th — the division in float64 on exact integers, as numpy does it — are written to
// (edge_gt, edge_pr) and counted; deg_gt / deg_pr count the edges per row / column.  out[0] = number of edges, out[1] = 1 when some
// row or column carries more than one edge (then the host solves the small matching; otherwise every edge is matched and tp = out[0]).
__global__ void k_iou_edges(const int32_t* __restrict__ pg, const int32_t* __restrict__ pp, const uint32_t* __restrict__ pc,
                            const uint32_t* __restrict__ n_pairs, const uint32_t* __restrict__ area_gt, const uint32_t* __restrict__ area_pr,
                            double th, uint32_t* __restrict__ deg_gt, uint32_t* __restrict__ deg_pr, int32_t* __restrict__ edge_gt,
                            int32_t* __restrict__ edge_pr, uint32_t* __restrict__ out) {
    const uint32_t n = *n_pairs;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int32_t a = pg[i], b = pp[i];
        if (a <= 0 || b <= 0) continue;
        const unsigned long long c = pc[i];
        const unsigned long long den = (unsigned long long)area_gt[a] + area_pr[b] - c;
        const double iou = __ddiv_rn(__ull2double_rn(c), __ull2double_rn(den));
        if (!(iou >= th)) continue;
        const uint32_t slot = atomicAdd(&out[0], 1u);
        edge_gt[slot] = a; edge_pr[slot] = b;
        const uint32_t da = atomicAdd(&deg_gt[a], 1u), db = atomicAdd(&deg_pr[b], 1u);
        if (da | db) out[1] = 1u;
    }
}
// deg_gt: uint32[nt + 1], deg_pr: uint32[np + 1] (zeroed here); edge_gt / edge_pr: int32[pair_cap]; out: uint32[2].
int iou_edges(const int32_t* pg, const int32_t* pp, const uint32_t* pc, const uint32_t* n_pairs, uint32_t pair_cap, int32_t nt, int32_t np_,
              const uint32_t* area_gt, const uint32_t* area_pr, double th, uint32_t* deg_gt, uint32_t* deg_pr, int32_t* edge_gt, int32_t* edge_pr,
              uint32_t* out, cudaStream_t st) {
    CS_CHECK_ARG(pg && pp && pc && n_pairs && area_gt && area_pr && deg_gt && deg_pr && edge_gt && edge_pr && out && nt >= 0 && np_ >= 0, "null pointer");
    CS_CUDA(cudaMemsetAsync(deg_gt, 0, ((size_t)nt + 1) * 4, st));
    CS_CUDA(cudaMemsetAsync(deg_pr, 0, ((size_t)np_ + 1) * 4, st));
    CS_CUDA(cudaMemsetAsync(out, 0, 8, st));
    CS_LAUNCH(k_iou_edges, flat_blocks_m(pair_cap, 256, 148 * 4), 256, 0, st, pg, pp, pc, n_pairs, area_gt, area_pr, th, deg_gt, deg_pr, edge_gt, edge_pr, out);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace cs
