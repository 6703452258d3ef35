// Mask → instance-map post-processing on device:
//   remove_small_holes(area 64, 4-conn) → remove_small_objects(16, 4-conn) →
//   measure.label(8-conn, ids in raster order of first pixel)
// as called at baseline/predict.py:113 of the reference.
//
// Data layout in HBM
//   M      packed mask, tile-major: tile (n,ty,tx) is 256 contiguous words (row r, word k at r*4+k)
//   R      root bitmap, row-major, WSR words per image row (bit set ⇔ pixel is the first pixel of a component)
//   base   exclusive prefix of popc(R) over the whole batch (+1 total entry)
//   Brow   [n][ty][top|bottom][Wp]  image-linear root id of each tile-border pixel (-1 background, -2 "large")
//   Bcol   [n][tx][left|right][Hp]
//   P      union-find parents over image-linear pixel ids, touched only at roots of border-touching
//          components; it lives in the caller's int32 label buffer until the final pass overwrites it
//   S      component sizes for the two size filters, same sparse indexing (workspace)
// Every tile kernel recomputes its tile-local components from the 1-bit mask instead of
// round-tripping per-pixel provisional labels through HBM; the only 4-byte-per-pixel traffic is the
// final label store.
#include "ccl_tile.cuh"
#include "internal.cuh"

namespace cs {

constexpr int32_t B_BG = -1;
constexpr int32_t B_LARGE = -2;
constexpr uint32_t AUX_UNDECIDED = 0x80000000u;
constexpr uint32_t AUX_KEEP = 0x40000000u;
constexpr uint32_t AUX_SIZE = 0x3fffffffu;

// ---------------------------------------------------------------------------
// pack / unpack
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TWORDS) k_pack_u8(const uint8_t* __restrict__ mask, uint32_t* __restrict__ M, Geo g) {
    const int tx = blockIdx.x, ty = blockIdx.y, n = blockIdx.z;
    const int t = threadIdx.x, r = t / WPR, k = t % WPR;
    const int y = ty * TH + r, x0 = tx * TW + k * 32;
    uint32_t w = 0;
    if (y < g.H && x0 < g.W) {
        const uint8_t* src = mask + (size_t)n * g.HW + (size_t)y * g.W + x0;
        if (x0 + 32 <= g.W && (((uintptr_t)src) & 15) == 0) {
            const uint4* v = reinterpret_cast<const uint4*>(src);
            uint4 a = v[0], b = v[1];
            uint32_t q[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t u = q[i];
                uint32_t nib = ((u & 0xffu) ? 1u : 0u) | ((u & 0xff00u) ? 2u : 0u) |
                               ((u & 0xff0000u) ? 4u : 0u) | ((u & 0xff000000u) ? 8u : 0u);
                w |= nib << (4 * i);
            }
        } else {
            int lim = min(32, g.W - x0);
            for (int i = 0; i < lim; ++i) w |= (src[i] ? 1u : 0u) << i;
        }
    }
    size_t tile = ((size_t)n * g.nty + ty) * g.ntx + tx;
    M[tile * TWORDS + t] = w;
}

__global__ void __launch_bounds__(TWORDS) k_unpack_u8(const uint32_t* __restrict__ M, uint8_t* __restrict__ mask, Geo g) {
    const int tx = blockIdx.x, ty = blockIdx.y, n = blockIdx.z;
    const int t = threadIdx.x, r = t / WPR, k = t % WPR;
    const int y = ty * TH + r, x0 = tx * TW + k * 32;
    if (y >= g.H || x0 >= g.W) return;
    size_t tile = ((size_t)n * g.nty + ty) * g.ntx + tx;
    uint32_t w = M[tile * TWORDS + t];
    uint8_t* dst = mask + (size_t)n * g.HW + (size_t)y * g.W + x0;
    int lim = min(32, g.W - x0);
    for (int i = 0; i < lim; ++i) dst[i] = (uint8_t)((w >> i) & 1u);
}

// ---------------------------------------------------------------------------
// global (cross-tile) union-find on image-linear ids
// ---------------------------------------------------------------------------
__device__ __forceinline__ int32_t guf_find(const volatile int32_t* P, int32_t a) {
    int32_t p = P[a];
    while (p != a) { a = p; p = P[a]; }
    return a;
}
__device__ __forceinline__ void guf_unite(int32_t* P, int32_t a, int32_t b) {
    for (;;) {
        a = guf_find(P, a); b = guf_find(P, b);
        if (a == b) return;
        if (a < b) { int32_t t = a; a = b; b = t; }
        if (atomicCAS(&P[a], a, b) == a) return;
    }
}

// ---------------------------------------------------------------------------
// size filters (remove_small_holes / remove_small_objects), 4-connectivity
// ---------------------------------------------------------------------------
// Pass 1: tile-local components and sizes.  A component that does not touch
// the tile border is complete and is decided here.  A border-touching
// component with local size >= thr is large whatever happens outside.  Only
// border-touching pieces smaller than thr are left undecided: they are kept
// for now, published (P, S, B, list) and resolved by k_filter_merge/accum/final.
template <bool INVERT>
__global__ void __launch_bounds__(TWORDS)
k_filter_local(const uint32_t* __restrict__ Min, uint32_t* __restrict__ Mout, Geo g, uint32_t thr,
               int32_t* __restrict__ Brow, int32_t* __restrict__ Bcol, int32_t* __restrict__ P,
               int32_t* __restrict__ S, int2* __restrict__ list, uint32_t* __restrict__ list_count,
               uint32_t* __restrict__ tileflag) {
    CS_DYN_SMEM(smem);
    uint32_t* parent = reinterpret_cast<uint32_t*>(smem);
    uint32_t* aux = parent + TPX;
    uint32_t* sb = aux + TPX;
    const int tx = blockIdx.x, ty = blockIdx.y, n = blockIdx.z;
    const int t = threadIdx.x;
    const uint32_t base = (uint32_t)t * 32u;
    const size_t tile = ((size_t)n * g.nty + ty) * g.ntx + tx;

    const uint32_t w = tile_load_bits(Min, g, n, ty, tx, INVERT, sb);
    const uint32_t orig = Min[tile * TWORDS + t];
    for (uint32_t st = w & ~(w << 1); st; st &= st - 1) aux[base + __ffs((int)st) - 1] = 0u;
    tile_ccl<4>(sb, parent);
    for (uint32_t st = w & ~(w << 1); st; st &= st - 1) {
        int s = __ffs((int)st) - 1;
        atomicAdd(&aux[parent[base + s]], (uint32_t)__popc(seg_mask(w, s)));
    }
    __syncthreads();
    int32_t* Pn = P + (size_t)n * g.HW;
    int32_t* Sn = S + (size_t)n * g.HW;
    for (int e = t; e < TPERIM; e += TWORDS) {
        int r, c; border_px(e, r, c);
        uint32_t bw = sb[(r + 1) * SBW + (c >> 5) + 1];
        int32_t val = B_BG;
        if ((bw >> (c & 31)) & 1u) {
            uint32_t root = parent[(uint32_t)(r * WPR + (c >> 5)) * 32u + seg_start(bw, c & 31)];
            uint32_t a = aux[root];
            if ((a & AUX_SIZE) >= thr) val = B_LARGE;
            else {
                int32_t gid = tile_to_gid(root, g, ty, tx);
                val = gid;
                uint32_t old = atomicOr(&aux[root], AUX_UNDECIDED);
                if (!(old & AUX_UNDECIDED)) {
                    Pn[gid] = gid; Sn[gid] = (int32_t)(old & AUX_SIZE);
                    uint32_t slot = atomicAdd(list_count, 1u);
                    list[slot] = make_int2(n, gid);
                    tileflag[tile] = 1u;
                }
            }
        }
        *border_slot(e, g, n, ty, tx, Brow, Bcol) = val;
    }
    __syncthreads();
    uint32_t removed = 0u;
    for (uint32_t st = w & ~(w << 1); st; st &= st - 1) {
        int s = __ffs((int)st) - 1;
        uint32_t a = aux[parent[base + s]];
        if (!(a & AUX_UNDECIDED) && (a & AUX_SIZE) < thr) removed |= seg_mask(w, s);
    }
    Mout[tile * TWORDS + t] = INVERT ? (orig | removed) : (w & ~removed);
}

// Pass 2: unions between undecided pieces facing each other across a tile edge
// (4-connectivity: straight neighbours only); a piece facing a large component
// is itself part of a large component.
__global__ void k_filter_merge(Geo g, uint32_t thr, const int32_t* __restrict__ Brow, const int32_t* __restrict__ Bcol,
                               int32_t* __restrict__ P, int32_t* __restrict__ S) {
    const size_t nh = (size_t)g.N * (g.nty - 1) * g.W;
    const size_t nv = (size_t)g.N * (g.ntx - 1) * g.H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nh + nv; i += (size_t)gridDim.x * blockDim.x) {
        int32_t a, b; int n;
        if (i < nh) {
            int x = (int)(i % g.W); size_t q = i / g.W; int ty = (int)(q % (g.nty - 1)); n = (int)(q / (g.nty - 1));
            a = Brow[(((size_t)n * g.nty + ty) * 2 + 1) * g.Wp + x];
            b = Brow[(((size_t)n * g.nty + ty + 1) * 2 + 0) * g.Wp + x];
        } else {
            size_t j = i - nh;
            int y = (int)(j % g.H); size_t q = j / g.H; int tx = (int)(q % (g.ntx - 1)); n = (int)(q / (g.ntx - 1));
            a = Bcol[(((size_t)n * g.ntx + tx) * 2 + 1) * g.Hp + y];
            b = Bcol[(((size_t)n * g.ntx + tx + 1) * 2 + 0) * g.Hp + y];
        }
        int32_t* Pn = P + (size_t)n * g.HW;
        int32_t* Sn = S + (size_t)n * g.HW;
        if (a >= 0 && b >= 0) guf_unite(Pn, a, b);
        else if (a >= 0 && b == B_LARGE) atomicMax(&Sn[a], (int32_t)thr);
        else if (b >= 0 && a == B_LARGE) atomicMax(&Sn[b], (int32_t)thr);
    }
}

// Pass 3: every undecided piece adds its size to its global root.
__global__ void k_filter_accum(Geo g, const int2* __restrict__ list, const uint32_t* __restrict__ list_count,
                               const int32_t* __restrict__ P, int32_t* __restrict__ S) {
    const uint32_t cnt = *list_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        int2 e = list[i];
        const int32_t* Pn = P + (size_t)e.x * g.HW;
        int32_t* Sn = S + (size_t)e.x * g.HW;
        int32_t r = guf_find(Pn, e.y);
        if (r != e.y) atomicAdd(&Sn[r], Sn[e.y]);
    }
}

// Pass 4: tiles that published undecided pieces redo their tile-local
// components and drop the pieces whose global size stayed below thr.
template <bool INVERT>
__global__ void __launch_bounds__(TWORDS)
k_filter_final(const uint32_t* __restrict__ Min, uint32_t* __restrict__ Mout, Geo g, uint32_t thr,
               const int32_t* __restrict__ P, const int32_t* __restrict__ S, const uint32_t* __restrict__ tileflag) {
    CS_DYN_SMEM(smem);
    uint32_t* parent = reinterpret_cast<uint32_t*>(smem);
    uint32_t* aux = parent + TPX;
    uint32_t* sb = aux + TPX;
    const int tx = blockIdx.x, ty = blockIdx.y, n = blockIdx.z;
    const int t = threadIdx.x;
    const uint32_t base = (uint32_t)t * 32u;
    const size_t tile = ((size_t)n * g.nty + ty) * g.ntx + tx;
    if (!tileflag[tile]) return;

    const uint32_t w = tile_load_bits(Min, g, n, ty, tx, INVERT, sb);
    const uint32_t orig = Min[tile * TWORDS + t];
    for (uint32_t st = w & ~(w << 1); st; st &= st - 1) aux[base + __ffs((int)st) - 1] = 0u;
    tile_ccl<4>(sb, parent);
    for (uint32_t st = w & ~(w << 1); st; st &= st - 1) {
        int s = __ffs((int)st) - 1;
        atomicAdd(&aux[parent[base + s]], (uint32_t)__popc(seg_mask(w, s)));
    }
    __syncthreads();
    const int32_t* Pn = P + (size_t)n * g.HW;
    const int32_t* Sn = S + (size_t)n * g.HW;
    for (int e = t; e < TPERIM; e += TWORDS) {
        int r, c; border_px(e, r, c);
        uint32_t bw = sb[(r + 1) * SBW + (c >> 5) + 1];
        if ((bw >> (c & 31)) & 1u) {
            uint32_t root = parent[(uint32_t)(r * WPR + (c >> 5)) * 32u + seg_start(bw, c & 31)];
            uint32_t a = aux[root];
            if ((a & AUX_SIZE) >= thr || (a & AUX_KEEP)) continue;
            int32_t gid = tile_to_gid(root, g, ty, tx);
            if (Sn[guf_find(Pn, gid)] >= (int32_t)thr) atomicOr(&aux[root], AUX_KEEP);
        }
    }
    __syncthreads();
    uint32_t removed = 0u;
    for (uint32_t st = w & ~(w << 1); st; st &= st - 1) {
        int s = __ffs((int)st) - 1;
        uint32_t a = aux[parent[base + s]];
        if (!(a & AUX_KEEP) && (a & AUX_SIZE) < thr) removed |= seg_mask(w, s);
    }
    Mout[tile * TWORDS + t] = INVERT ? (orig | removed) : (w & ~removed);
}

// ---------------------------------------------------------------------------
// labelling with raster-order ids
// ---------------------------------------------------------------------------
// Pass 1: tile-local components; publish the root bitmap and, for components
// touching the tile border, their image-linear root id (B) and a singleton in P.
template <int CONN>
__global__ void __launch_bounds__(TWORDS)
k_label_local(const uint32_t* __restrict__ M, Geo g, uint32_t* __restrict__ R,
              int32_t* __restrict__ Brow, int32_t* __restrict__ Bcol, int32_t* __restrict__ P) {
    CS_DYN_SMEM(smem);
    uint32_t* parent = reinterpret_cast<uint32_t*>(smem);
    uint32_t* sb = parent + 2 * TPX;
    const int tx = blockIdx.x, ty = blockIdx.y, n = blockIdx.z;
    const int t = threadIdx.x, r = t / WPR, k = t % WPR;
    const uint32_t base = (uint32_t)t * 32u;

    const uint32_t w = tile_load_bits(M, g, n, ty, tx, false, sb);
    tile_ccl<CONN>(sb, parent);
    uint32_t roots = 0u;
    for (uint32_t st = w & ~(w << 1); st; st &= st - 1) {
        int s = __ffs((int)st) - 1;
        if (parent[base + s] == base + s) roots |= 1u << s;
    }
    const int y = ty * TH + r;
    if (y < g.H) R[((size_t)n * g.H + y) * g.WSR + tx * WPR + k] = roots;
    int32_t* Pn = P + (size_t)n * g.HW;
    for (int e = t; e < TPERIM; e += TWORDS) {
        int br, bc; border_px(e, br, bc);
        uint32_t bw = sb[(br + 1) * SBW + (bc >> 5) + 1];
        int32_t val = B_BG;
        if ((bw >> (bc & 31)) & 1u) {
            uint32_t root = parent[(uint32_t)(br * WPR + (bc >> 5)) * 32u + seg_start(bw, bc & 31)];
            val = tile_to_gid(root, g, ty, tx);
            Pn[val] = val;
        }
        *border_slot(e, g, n, ty, tx, Brow, Bcol) = val;
    }
}

// Pass 2: unions across tile edges (straight, plus diagonal for 8-connectivity;
// Brow/Bcol span whole image rows/columns so corner-crossing diagonals are covered).
template <int CONN>
__global__ void k_label_merge(Geo g, const int32_t* __restrict__ Brow, const int32_t* __restrict__ Bcol, int32_t* __restrict__ P) {
    const size_t nh = (size_t)g.N * (g.nty - 1) * g.W;
    const size_t nv = (size_t)g.N * (g.ntx - 1) * g.H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nh + nv; i += (size_t)gridDim.x * blockDim.x) {
        if (i < nh) {
            int x = (int)(i % g.W); size_t q = i / g.W; int ty = (int)(q % (g.nty - 1)); int n = (int)(q / (g.nty - 1));
            int32_t lo = Brow[(((size_t)n * g.nty + ty + 1) * 2 + 0) * g.Wp + x];
            if (lo < 0) continue;
            const int32_t* uprow = Brow + (((size_t)n * g.nty + ty) * 2 + 1) * g.Wp;
            int32_t* Pn = P + (size_t)n * g.HW;
            const int d = CONN == 8 ? 1 : 0;
            for (int dx = -d; dx <= d; ++dx) {
                int xx = x + dx;
                if (xx < 0 || xx >= g.W) continue;
                int32_t u = uprow[xx];
                if (u >= 0) guf_unite(Pn, lo, u);
            }
        } else {
            size_t j = i - nh;
            int y = (int)(j % g.H); size_t q = j / g.H; int tx = (int)(q % (g.ntx - 1)); int n = (int)(q / (g.ntx - 1));
            int32_t ri = Bcol[(((size_t)n * g.ntx + tx + 1) * 2 + 0) * g.Hp + y];
            if (ri < 0) continue;
            const int32_t* lcol = Bcol + (((size_t)n * g.ntx + tx) * 2 + 1) * g.Hp;
            int32_t* Pn = P + (size_t)n * g.HW;
            const int d = CONN == 8 ? 1 : 0;
            for (int dy = -d; dy <= d; ++dy) {
                int yy = y + dy;
                if (yy < 0 || yy >= g.H) continue;
                int32_t l = lcol[yy];
                if (l >= 0) guf_unite(Pn, ri, l);
            }
        }
    }
}

// Pass 3: replace every border entry by its global root and clear the root bit
// of tile-local roots that were hooked under another tile's component.
__global__ void k_label_flatten(Geo g, int32_t* __restrict__ Brow, int32_t* __restrict__ Bcol,
                                const int32_t* __restrict__ P, uint32_t* __restrict__ R) {
    const size_t nr = (size_t)g.N * g.nty * 2 * g.Wp;
    const size_t nc = (size_t)g.N * g.ntx * 2 * g.Hp;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nr + nc; i += (size_t)gridDim.x * blockDim.x) {
        int32_t* slot; int n;
        if (i < nr) { slot = Brow + i; n = (int)(i / ((size_t)g.nty * 2 * g.Wp)); }
        else { slot = Bcol + (i - nr); n = (int)((i - nr) / ((size_t)g.ntx * 2 * g.Hp)); }
        int32_t b = *slot;
        if (b < 0) continue;
        int32_t r = guf_find(P + (size_t)n * g.HW, b);
        if (r != b) {
            int y = b / g.W, x = b % g.W;
            atomicAnd(&R[((size_t)n * g.H + y) * g.WSR + (x >> 5)], ~(1u << (x & 31)));
            *slot = r;
        }
    }
}

// Pass 4: exclusive prefix count of root bits (three small kernels).
constexpr int SCAN_T = 256, SCAN_PER = 8, SCAN_CHUNK = SCAN_T * SCAN_PER;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total_out) {
    __shared__ uint32_t warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    uint32_t inc = v;
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < nw ? warp_sums[lane] : 0u;
        uint32_t si = s;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, si, d);
            if (lane >= d) si += o;
        }
        warp_sums[lane] = si - s;  // exclusive warp offsets
        if (lane == 31) *total_out = si;
    }
    __syncthreads();
    uint32_t r = inc - v + warp_sums[wid];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_reduce(const uint32_t* __restrict__ R, size_t nwords, uint32_t* __restrict__ blocksum) {
    __shared__ uint32_t total;
    size_t b0 = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_PER;
    uint32_t s = 0;
    for (int i = 0; i < SCAN_PER; ++i) if (b0 + i < nwords) s += (uint32_t)__popc(R[b0 + i]);
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) blocksum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) k_scan_blocksums(uint32_t* __restrict__ blocksum, int nblocks, uint32_t* __restrict__ grand_total) {
    __shared__ uint32_t total;
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += 1024) {
        int i = b0 + threadIdx.x;
        uint32_t v = i < nblocks ? blocksum[i] : 0u;
        uint32_t ex = block_exclusive_scan(v, &total);
        if (i < nblocks) blocksum[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_apply(const uint32_t* __restrict__ R, size_t nwords, const uint32_t* __restrict__ blocksum,
                                                       uint32_t* __restrict__ base) {
    __shared__ uint32_t total;
    size_t b0 = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_PER;
    uint32_t c[SCAN_PER]; uint32_t s = 0;
    for (int i = 0; i < SCAN_PER; ++i) { c[i] = b0 + i < nwords ? (uint32_t)__popc(R[b0 + i]) : 0u; s += c[i]; }
    uint32_t ex = block_exclusive_scan(s, &total) + blocksum[blockIdx.x];
    for (int i = 0; i < SCAN_PER; ++i) { if (b0 + i < nwords) base[b0 + i] = ex; ex += c[i]; }
}
__global__ void k_counts(Geo g, const uint32_t* __restrict__ base, int32_t* __restrict__ n_out) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= g.N) return;
    size_t per = (size_t)g.H * g.WSR;
    n_out[n] = (int32_t)(base[(size_t)(n + 1) * per] - base[(size_t)n * per]);
}

// Pass 5: tile-local components again, translate each root to the global
// root's raster rank and store the int32 labels (the one 4 B/pixel stream).
template <int CONN>
__global__ void __launch_bounds__(TWORDS)
k_label_final(const uint32_t* __restrict__ M, Geo g, const uint32_t* __restrict__ R, const uint32_t* __restrict__ base,
              const int32_t* __restrict__ Brow, const int32_t* __restrict__ Bcol, int32_t* __restrict__ labels) {
    CS_DYN_SMEM(smem);
    uint32_t* parent = reinterpret_cast<uint32_t*>(smem);
    uint32_t* aux = parent + TPX;
    uint32_t* sb = aux + TPX;
    const int tx = blockIdx.x, ty = blockIdx.y, n = blockIdx.z;
    const int t = threadIdx.x;
    const uint32_t wbase = (uint32_t)t * 32u;

    const uint32_t w = tile_load_bits(M, g, n, ty, tx, false, sb);
    tile_ccl<CONN>(sb, parent);
    for (uint32_t st = w & ~(w << 1); st; st &= st - 1) {
        uint32_t s = wbase + __ffs((int)st) - 1;
        if (parent[s] == s) aux[s] = (uint32_t)tile_to_gid(s, g, ty, tx);
    }
    __syncthreads();
    for (int e = t; e < TPERIM; e += TWORDS) {
        int br, bc; border_px(e, br, bc);
        uint32_t bw = sb[(br + 1) * SBW + (bc >> 5) + 1];
        if ((bw >> (bc & 31)) & 1u) {
            uint32_t root = parent[(uint32_t)(br * WPR + (bc >> 5)) * 32u + seg_start(bw, bc & 31)];
            aux[root] = (uint32_t)*border_slot(e, g, n, ty, tx, const_cast<int32_t*>(Brow), const_cast<int32_t*>(Bcol));
        }
    }
    __syncthreads();
    const size_t img_words = (size_t)g.H * g.WSR;
    const uint32_t img_base = base[(size_t)n * img_words];
    for (uint32_t st = w & ~(w << 1); st; st &= st - 1) {
        uint32_t s = wbase + __ffs((int)st) - 1;
        if (parent[s] == s) {
            int32_t gid = (int32_t)aux[s];
            int y = gid / g.W, x = gid % g.W;
            size_t wi = (size_t)n * img_words + (size_t)y * g.WSR + (x >> 5);
            aux[s] = base[wi] - img_base + (uint32_t)__popc(R[wi] & ((1u << (x & 31)) - 1u)) + 1u;
        }
    }
    __syncthreads();
    for (uint32_t st = w & ~(w << 1); st; st &= st - 1) {
        uint32_t s = wbase + __ffs((int)st) - 1;
        parent[s] = aux[parent[s]];
    }
    __syncthreads();
    // store: a warp covers one 128-pixel tile row per iteration, 4 pixels per lane
    const int lane = t & 31, warp = t >> 5;
    const int k = lane >> 3, b0 = (lane & 7) * 4;
    const int x = tx * TW + lane * 4;
    int32_t* out = labels + (size_t)n * g.HW;
    const bool vec = (g.W & 3) == 0;
    for (int r = warp; r < TH; r += TWORDS / 32) {
        const int y = ty * TH + r;
        if (y >= g.H) break;
        uint32_t rw = sb[(r + 1) * SBW + k + 1];
        uint32_t sbase = (uint32_t)(r * WPR + k) * 32u;
        int32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int b = b0 + j;
            v[j] = ((rw >> b) & 1u) ? (int32_t)parent[sbase + seg_start(rw, b)] : 0;
        }
        int32_t* dst = out + (size_t)y * g.W + x;
        if (vec) { if (x < g.W) *reinterpret_cast<int4*>(dst) = make_int4(v[0], v[1], v[2], v[3]); }
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (x + j < g.W) dst[j] = v[j];
        }
    }
}

// ---------------------------------------------------------------------------
// host-side orchestration
// ---------------------------------------------------------------------------
struct PostWs {
    uint32_t *M0, *M1, *R, *base, *blocksum, *list_count, *tileflag, *grand;
    int32_t *Brow, *Bcol, *S;
    int2* list;
    size_t bytes;
    int nscan;
};
static PostWs carve(void* ws, size_t cap, const Geo& g, bool need_sizes) {
    Arena a(ws, cap);
    PostWs p;
    size_t ntiles = (size_t)g.N * g.nty * g.ntx;
    size_t nwords = (size_t)g.N * g.H * g.WSR;
    p.nscan = (int)((nwords + SCAN_CHUNK - 1) / SCAN_CHUNK);
    p.M0 = a.take<uint32_t>(ntiles * TWORDS);
    p.M1 = a.take<uint32_t>(ntiles * TWORDS);
    p.R = a.take<uint32_t>(nwords);
    p.base = a.take<uint32_t>(nwords + 1);
    p.blocksum = a.take<uint32_t>((size_t)p.nscan + 1);
    p.list_count = a.take<uint32_t>(64);
    p.grand = p.list_count + 1;
    p.tileflag = a.take<uint32_t>(ntiles);
    p.Brow = a.take<int32_t>((size_t)g.N * g.nty * 2 * g.Wp);
    p.Bcol = a.take<int32_t>((size_t)g.N * g.ntx * 2 * g.Hp);
    p.list = a.take<int2>(ntiles * TPERIM);
    p.S = need_sizes ? a.take<int32_t>((size_t)g.N * g.HW) : nullptr;
    p.bytes = a.off;
    return p;
}

static int check_geo(int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0) { set_error("image batch must have positive N, H, W (got %d, %d, %d)", N, H, W); return CS_ERR_ARG; }
    if ((long long)H * W > 0x7fffffffLL) { set_error("H*W = %lld exceeds the 32-bit pixel-id range", (long long)H * W); return CS_ERR_RANGE; }
    if (ceil_div(H, TH) > 65535 || N > 65535) { set_error("grid dimension out of range"); return CS_ERR_RANGE; }
    return CS_OK;
}

size_t postprocess_workspace_bytes(int N, int H, int W) {
    if (check_geo(N, H, W) != CS_OK) return 0;
    Geo g = make_geo(N, H, W);
    return carve(nullptr, 0, g, true).bytes;
}

static bool g_attr_done = false;
static int ensure_attrs() {
    if (g_attr_done) return CS_OK;
    CS_CUDA(cudaFuncSetAttribute(k_filter_local<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM));
    CS_CUDA(cudaFuncSetAttribute(k_filter_local<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM));
    CS_CUDA(cudaFuncSetAttribute(k_filter_final<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM));
    CS_CUDA(cudaFuncSetAttribute(k_filter_final<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM));
    CS_CUDA(cudaFuncSetAttribute(k_label_local<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM));
    CS_CUDA(cudaFuncSetAttribute(k_label_local<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM));
    CS_CUDA(cudaFuncSetAttribute(k_label_final<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM));
    CS_CUDA(cudaFuncSetAttribute(k_label_final<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM));
    g_attr_done = true;
    return CS_OK;
}

static int flat_grid(size_t work, int threads) {
    size_t b = (work + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > 148 * 16) b = 148 * 16;
    return (int)b;
}

// One size filter: Min → Mout (packed).  invert=true fills background components (holes).
static int run_filter(const Geo& g, const PostWs& p, const uint32_t* Min, uint32_t* Mout, bool invert, int thr,
                      int32_t* P, cudaStream_t st) {
    dim3 grid(g.ntx, g.nty, g.N);
    size_t ntiles = (size_t)g.N * g.nty * g.ntx;
    CS_CUDA(cudaMemsetAsync(p.list_count, 0, sizeof(uint32_t), st));
    CS_CUDA(cudaMemsetAsync(p.tileflag, 0, ntiles * sizeof(uint32_t), st));
    if (invert) CS_LAUNCH(k_filter_local<true>, grid, TWORDS, TILE_SMEM, st, Min, Mout, g, (uint32_t)thr, p.Brow, p.Bcol, P, p.S, p.list, p.list_count, p.tileflag);
    else        CS_LAUNCH(k_filter_local<false>, grid, TWORDS, TILE_SMEM, st, Min, Mout, g, (uint32_t)thr, p.Brow, p.Bcol, P, p.S, p.list, p.list_count, p.tileflag);
    CS_LAUNCH_CHECK();
    size_t edges = (size_t)g.N * (g.nty - 1) * g.W + (size_t)g.N * (g.ntx - 1) * g.H;
    if (edges) {
        CS_LAUNCH(k_filter_merge, flat_grid(edges, 256), 256, 0, st, g, (uint32_t)thr, p.Brow, p.Bcol, P, p.S);
        CS_LAUNCH_CHECK();
    }
    CS_LAUNCH(k_filter_accum, flat_grid(ntiles * 8, 256), 256, 0, st, g, p.list, p.list_count, P, p.S);
    CS_LAUNCH_CHECK();
    if (invert) CS_LAUNCH(k_filter_final<true>, grid, TWORDS, TILE_SMEM, st, Min, Mout, g, (uint32_t)thr, P, p.S, p.tileflag);
    else        CS_LAUNCH(k_filter_final<false>, grid, TWORDS, TILE_SMEM, st, Min, Mout, g, (uint32_t)thr, P, p.S, p.tileflag);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

static int run_label(const Geo& g, const PostWs& p, const uint32_t* M, int conn, int32_t* labels, int32_t* n_out, cudaStream_t st) {
    dim3 grid(g.ntx, g.nty, g.N);
    size_t nwords = (size_t)g.N * g.H * g.WSR;
    if (conn == 8) CS_LAUNCH(k_label_local<8>, grid, TWORDS, TILE_SMEM, st, M, g, p.R, p.Brow, p.Bcol, labels);
    else           CS_LAUNCH(k_label_local<4>, grid, TWORDS, TILE_SMEM, st, M, g, p.R, p.Brow, p.Bcol, labels);
    CS_LAUNCH_CHECK();
    size_t edges = (size_t)g.N * (g.nty - 1) * g.W + (size_t)g.N * (g.ntx - 1) * g.H;
    if (edges) {
        if (conn == 8) CS_LAUNCH(k_label_merge<8>, flat_grid(edges, 256), 256, 0, st, g, p.Brow, p.Bcol, labels);
        else           CS_LAUNCH(k_label_merge<4>, flat_grid(edges, 256), 256, 0, st, g, p.Brow, p.Bcol, labels);
        CS_LAUNCH_CHECK();
        size_t nb = (size_t)g.N * g.nty * 2 * g.Wp + (size_t)g.N * g.ntx * 2 * g.Hp;
        CS_LAUNCH(k_label_flatten, flat_grid(nb, 256), 256, 0, st, g, p.Brow, p.Bcol, labels, p.R);
        CS_LAUNCH_CHECK();
    }
    CS_LAUNCH(k_scan_reduce, p.nscan, SCAN_T, 0, st, p.R, nwords, p.blocksum);
    CS_LAUNCH_CHECK();
    CS_LAUNCH(k_scan_blocksums, 1, 1024, 0, st, p.blocksum, p.nscan, p.base + nwords);
    CS_LAUNCH_CHECK();
    CS_LAUNCH(k_scan_apply, p.nscan, SCAN_T, 0, st, p.R, nwords, p.blocksum, p.base);
    CS_LAUNCH_CHECK();
    if (n_out) {
        CS_LAUNCH(k_counts, ceil_div(g.N, 128), 128, 0, st, g, p.base, n_out);
        CS_LAUNCH_CHECK();
    }
    if (conn == 8) CS_LAUNCH(k_label_final<8>, grid, TWORDS, TILE_SMEM, st, M, g, p.R, p.base, p.Brow, p.Bcol, labels);
    else           CS_LAUNCH(k_label_final<4>, grid, TWORDS, TILE_SMEM, st, M, g, p.R, p.base, p.Brow, p.Bcol, labels);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

static int prep(int N, int H, int W, void* ws, size_t ws_bytes, bool need_sizes, Geo& g, PostWs& p) {
    int rc = check_geo(N, H, W);
    if (rc != CS_OK) return rc;
    g = make_geo(N, H, W);
    p = carve(ws, ws_bytes, g, need_sizes);
    if (!ws || p.bytes > ws_bytes) {
        set_error("workspace too small: need %zu bytes, got %zu", p.bytes, ws_bytes);
        return CS_ERR_WORKSPACE;
    }
    return ensure_attrs();
}

int pack_mask(const uint8_t* mask, const Geo& g, uint32_t* M, cudaStream_t st) {
    CS_LAUNCH(k_pack_u8, dim3(g.ntx, g.nty, g.N), TWORDS, 0, st, mask, M, g);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// mask u8[N,H,W] (non-zero = foreground) → labels int32[N,H,W], n_out int32[N]
int postprocess(const uint8_t* mask, int N, int H, int W, int hole_area, int min_size, int connectivity,
                int32_t* labels, int32_t* n_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    CS_CHECK_ARG(mask && labels, "null mask or labels pointer");
    CS_CHECK_ARG(connectivity == 4 || connectivity == 8, "connectivity must be 4 or 8");
    CS_CHECK_ARG(hole_area >= 0 && min_size >= 0, "negative size threshold");
    Geo g; PostWs p;
    int rc = prep(N, H, W, ws, ws_bytes, true, g, p);
    if (rc != CS_OK) return rc;
    if ((rc = pack_mask(mask, g, p.M0, st)) != CS_OK) return rc;
    uint32_t* cur = p.M0; uint32_t* nxt = p.M1;
    if (hole_area > 1) { if ((rc = run_filter(g, p, cur, nxt, true, hole_area, labels, st)) != CS_OK) return rc; uint32_t* t = cur; cur = nxt; nxt = t; }
    if (min_size > 1)  { if ((rc = run_filter(g, p, cur, nxt, false, min_size, labels, st)) != CS_OK) return rc; uint32_t* t = cur; cur = nxt; nxt = t; }
    return run_label(g, p, cur, connectivity, labels, n_out, st);
}

// Same chain starting from an already packed (tile-major) mask living in the workspace's M0.
int postprocess_packed(int N, int H, int W, int hole_area, int min_size, int connectivity,
                       int32_t* labels, int32_t* n_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    CS_CHECK_ARG(labels, "null labels pointer");
    CS_CHECK_ARG(connectivity == 4 || connectivity == 8, "connectivity must be 4 or 8");
    Geo g; PostWs p;
    int rc = prep(N, H, W, ws, ws_bytes, true, g, p);
    if (rc != CS_OK) return rc;
    uint32_t* cur = p.M0; uint32_t* nxt = p.M1;
    if (hole_area > 1) { if ((rc = run_filter(g, p, cur, nxt, true, hole_area, labels, st)) != CS_OK) return rc; uint32_t* t = cur; cur = nxt; nxt = t; }
    if (min_size > 1)  { if ((rc = run_filter(g, p, cur, nxt, false, min_size, labels, st)) != CS_OK) return rc; uint32_t* t = cur; cur = nxt; nxt = t; }
    return run_label(g, p, cur, connectivity, labels, n_out, st);
}

uint32_t* packed_mask_slot(void* ws, size_t ws_bytes, int N, int H, int W) {
    if (check_geo(N, H, W) != CS_OK) return nullptr;
    Geo g = make_geo(N, H, W);
    PostWs p = carve(ws, ws_bytes, g, true);
    return (ws && p.bytes <= ws_bytes) ? p.M0 : nullptr;
}

// measure.label on its own.
int label(const uint8_t* mask, int N, int H, int W, int connectivity, int32_t* labels, int32_t* n_out,
          void* ws, size_t ws_bytes, cudaStream_t st) {
    return postprocess(mask, N, H, W, 0, 0, connectivity, labels, n_out, ws, ws_bytes, st);
}

// remove_small_objects (invert=0) / remove_small_holes (invert=1) on a u8 mask; `scratch` is an
// int32[N*H*W] buffer for the sparse union-find parents.
int remove_small(const uint8_t* mask_in, uint8_t* mask_out, int N, int H, int W, int connectivity, int min_size, int invert,
                 int32_t* scratch, void* ws, size_t ws_bytes, cudaStream_t st) {
    CS_CHECK_ARG(mask_in && mask_out && scratch, "null pointer");
    CS_CHECK_ARG(connectivity == 4, "size filters implement the reference's default 4-connectivity only");
    CS_CHECK_ARG(min_size >= 0, "negative size threshold");
    Geo g; PostWs p;
    int rc = prep(N, H, W, ws, ws_bytes, true, g, p);
    if (rc != CS_OK) return rc;
    if ((rc = pack_mask(mask_in, g, p.M0, st)) != CS_OK) return rc;
    uint32_t* res = p.M0;
    if (min_size > 1) { if ((rc = run_filter(g, p, p.M0, p.M1, invert != 0, min_size, scratch, st)) != CS_OK) return rc; res = p.M1; }
    CS_LAUNCH(k_unpack_u8, dim3(g.ntx, g.nty, g.N), TWORDS, 0, st, res, mask_out, g);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace cs
