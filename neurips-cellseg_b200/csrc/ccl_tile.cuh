// Tile-local connected components on a bit-packed mask, in shared memory.
//
// One CTA owns a TH x TW pixel tile (64 x 128 → 256 32-bit words, one word per
// thread).  Work is done per *segment* (a maximal run of set bits inside one
// word) rather than per pixel: a segment is named by the tile-linear index of
// its first pixel (13 bits), `P[]` is a union-find forest over those names (packed
// 32-bit entries, see below) with "smaller index wins" hooking, so a component's root is its
// first pixel in raster order — which is what makes scipy/skimage-compatible
// label ids a prefix count later (reference: skimage.measure.label called at
// baseline/predict.py:113; semantics in SURVEY.md §8 a-7).
#pragma once
#include "cs_platform.cuh"

namespace cs {

constexpr int TH = 64;            // tile rows
constexpr int TW = 128;           // tile cols
constexpr int WPR = TW / 32;      // words per tile row
constexpr int TWORDS = TH * WPR;  // 256 words = threads per CTA
constexpr int TPX = TH * TW;      // 8192 pixels
constexpr int SBW = WPR + 2;      // padded bit-row stride (zero word either side)
constexpr int TPERIM = 2 * TW + 2 * TH;  // border entries per tile: top,bottom rows; left,right cols
constexpr int SEGMAX = TPX / 2;   // segments per tile (≤ 16 per word)
constexpr int NODES_PER_TILE = 256;   // undecided border pieces per tile (≤ 190: ring pixels of two pieces are never adjacent)

// Union-find entry of a segment, in shared memory, keyed by (first pixel >> 1) — two segment starts are never horizontally
// adjacent, so 4096 32-bit entries cover the 8192 pixels and the hook is a native 32-bit compare-and-swap (a 16-bit parent
// array needed a CAS emulation loop that was 18 % of the tile kernel's instructions, profiles/r01_ncu_lines_k_tile.txt).
//   bits  0..12  parent: tile-linear pixel index of another segment start (itself at a root)
//   bit   13     UF_LARGE: the component touches pixels already certified as part of a large component
//   bit   14     UF_EXT:   the component has a pixel on the tile border whose neighbour across the border may belong to it
//   bit   15     UF_UNDEC: root of an undecided piece (then bits 16.. hold its node ordinal instead of its size)
//   bits 16..29  size of the component in pixels (accumulated at the root during the flatten pass)
// The flag / size bits are only ever written after the union phase, so hooks compare against the bare index.
constexpr uint32_t UF_PAR = 0x1fffu;
constexpr uint32_t UF_LARGE = 1u << 13;
constexpr uint32_t UF_EXT = 1u << 14;
constexpr uint32_t UF_UNDEC = 1u << 15;
constexpr int UF_SIZE_SHIFT = 16;
constexpr uint32_t UF_SIZE = 0x3fffu;

struct Geo {
    int N, H, W;        // images, rows, cols
    int nty, ntx;       // tiles per column / row
    int WSR;            // words per row of the row-major root bitmap (= ntx*WPR)
    int Wp, Hp;         // padded extents ntx*TW, nty*TH
    size_t HW;          // H*W
};

static inline Geo make_geo(int N, int H, int W) {
    Geo g; g.N = N; g.H = H; g.W = W;
    g.nty = ceil_div(H, TH); g.ntx = ceil_div(W, TW);
    g.WSR = g.ntx * WPR; g.Wp = g.ntx * TW; g.Hp = g.nty * TH;
    g.HW = (size_t)H * W;
    return g;
}

// First bit of the run of ones that contains bit b of w (b must be set).
__device__ __forceinline__ int seg_start(uint32_t w, int b) {
    uint32_t zeros_below = ~w & ((1u << b) - 1u);
    return 32 - __clz((int)zeros_below);
}
// Mask of the run of ones of w that starts at bit s.
__device__ __forceinline__ uint32_t seg_mask(uint32_t w, int s) {
    uint32_t t = w + (1u << s);
    return (w ^ t) & w;
}
// Ordinal (within the word) of the segment that contains set bit b: segments are counted by their start bits.
__device__ __forceinline__ int seg_ordinal(uint32_t w, int b) {
    uint32_t starts = w & ~(w << 1);
    return __popc(starts & ((2u << b) - 1u)) - 1;
}

// Every run of ones of w that holds a bit of s (s inside w), whole: carry propagation upwards, the same on the reversed word
// downwards.
__device__ __forceinline__ uint32_t fill_runs32(uint32_t s, uint32_t w) {
    const uint32_t rw = __brev(w), rs = __brev(s);
    return ((w ^ (w + s)) & w) | s | __brev((rw ^ (rw + rs)) & rw);
}
// bit i set iff bits i .. i+b-1 of w are all set (1 <= b <= 32)
__device__ __forceinline__ uint32_t runs_of32(uint32_t w, int b) {
    uint32_t h = w; int p = 1;
    while (2 * p <= b) { h &= h >> p; p *= 2; }
    if (b > p) h &= h >> (b - p);
    return h;
}

// Union phase only (entries are bare indices): find with path halving, hook the larger root under the smaller.
__device__ __forceinline__ uint32_t uf_find(volatile uint32_t* P, uint32_t a) {
    uint32_t p = P[a >> 1];
    while (p != a) {
        uint32_t g = P[p >> 1];
        if (g != p) P[a >> 1] = g;  // path halving: only ever replaces a parent by an ancestor
        a = p; p = g;
    }
    return a;
}
// Any phase: the flag / size bits are masked off.
__device__ __forceinline__ uint32_t uf_find_ro(const volatile uint32_t* P, uint32_t a) {
    uint32_t p = P[a >> 1] & UF_PAR;
    while (p != a) { a = p; p = P[a >> 1] & UF_PAR; }
    return a;
}
__device__ __forceinline__ void uf_unite(uint32_t* P, uint32_t a, uint32_t b) {
    for (;;) {
        a = uf_find(P, a); b = uf_find(P, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }
        if (atomicCAS(&P[a >> 1], a, b) == a) return;  // hook the larger root under the smaller
    }
}

// Unions of this thread's word with the row above.  One union per run of vertically touching pixels (a run
// that merely continues from the left word is that word's job); diagonal unions only where neither the pixel
// above nor the horizontal neighbour already carries the connection.  DIAG_ONLY adds just the diagonal cases
// (turning a finished 4-connected forest into the 8-connected one).
template <int CONN, bool DIAG_ONLY>
__device__ __forceinline__ void unite_with_row_above(uint32_t* P, uint32_t base, uint32_t w, uint32_t wl, uint32_t wr,
                                                     uint32_t up, uint32_t ul, uint32_t ur) {
    const uint32_t upbase = base - (uint32_t)TW;
    if (!DIAG_ONLY) {
        uint32_t c = w & up;
        uint32_t sv = c & ~(c << 1) & ~((wl & ul) >> 31);
        while (sv) {
            int b = __ffs((int)sv) - 1; sv &= sv - 1;
            uf_unite(P, base + seg_start(w, b), upbase + seg_start(up, b));
        }
    }
    if (CONN == 8) {
        uint32_t upl = (up << 1) | (ul >> 31), curl = (w << 1) | (wl >> 31);
        uint32_t dl = w & upl & ~up & ~curl;   // pixel b, neighbour (row-1, b-1)
        while (dl) {
            int b = __ffs((int)dl) - 1; dl &= dl - 1;
            uint32_t u = b > 0 ? upbase + seg_start(up, b - 1) : upbase - 32u + seg_start(ul, 31);
            uf_unite(P, base + b, u);
        }
        uint32_t uprr = (up >> 1) | (ur << 31), curr = (w >> 1) | (wr << 31);
        uint32_t dr = w & uprr & ~up & ~curr;  // pixel b, neighbour (row-1, b+1)
        while (dr) {
            int b = __ffs((int)dr) - 1; dr &= dr - 1;
            uf_unite(P, base + seg_start(w, b), upbase + b + 1);
        }
    }
}

// Union-find over the tile's segments; the tile's bits are in the zero-padded array `sb`
// (visible to all threads on entry).  All TWORDS threads must call.  Two levels keep the pointer chains
// short: a warp owns a band of 8 rows and first resolves it on its own (__syncwarp only), then the band
// tops are united across bands.  On return every thread has flattened its own segment starts
// (P[s >> 1] & UF_PAR = root = first pixel of the component, tile-linear); the caller must __syncthreads() before
// reading other threads' entries.  CONN is 4 or 8.
// two_level (block-uniform): worth it when components are large (long vertical chains); a tile full of small
// pieces has short chains anyway and is better off with the single pass.
// SIZES: the flatten pass adds every segment's length to its root's size field and ORs `large` / `ext` (masks of
// this word's pixels that touch certified-large pixels / have a possible continuation across the tile border) into the
// root's flags for the segments they hit, so the size filters need no loops of their own.
template <int CONN, bool SIZES>
__device__ __forceinline__ void tile_ccl(const uint32_t* sb, uint32_t* P, bool two_level, uint32_t large = 0u, uint32_t ext = 0u) {
    const int t = threadIdx.x;
    const int r = t / WPR, k = t % WPR;
    const uint32_t* row = sb + (r + 1) * SBW + (k + 1);
    const uint32_t* upr_ = row - SBW;
    const uint32_t w = row[0], wl = row[-1], wr = row[1];
    const uint32_t up = upr_[0], ul = upr_[-1], ur = upr_[1];
    const uint32_t base = (uint32_t)t * 32u;
    const uint32_t starts = w & ~(w << 1);
    const bool band_top = (r & (32 / WPR - 1)) == 0;       // first row of this warp's band

    // 1. every segment starts as its own set; a run that continues across the
    //    word boundary points at the segment holding the left word's bit 31.
    for (uint32_t st = starts; st; st &= st - 1) {
        int s = __ffs((int)st) - 1;
        uint32_t p = base + s;
        if (s == 0 && (wl >> 31)) p = base - 32u + (uint32_t)seg_start(wl, 31);
        P[(base + s) >> 1] = p;
    }
    if (two_level) {
        __syncwarp();
        // 2. inside the band
        if (w && !band_top) unite_with_row_above<CONN, false>(P, base, w, wl, wr, up, ul, ur);
        __syncwarp();
        for (uint32_t st = starts; st; st &= st - 1) {
            int s = __ffs((int)st) - 1;
            P[(base + s) >> 1] = uf_find_ro(P, base + s);
        }
        __syncthreads();
        // 3. across bands
        if (w && band_top && r > 0) unite_with_row_above<CONN, false>(P, base, w, wl, wr, up, ul, ur);
    } else {
        __syncthreads();
        if (w) unite_with_row_above<CONN, false>(P, base, w, wl, wr, up, ul, ur);
    }
    __syncthreads();
    // 4. flatten own entries (concurrent readers only ever see a parent replaced by an ancestor; a root's entry is never
    //    stored to here, only atomically added to)
    for (uint32_t st = starts; st; st &= st - 1) {
        int s = __ffs((int)st) - 1;
        const uint32_t root = uf_find_ro(P, base + s);
        if (root != base + s) P[(base + s) >> 1] = root;
        if (SIZES) {
            const uint32_t m = seg_mask(w, s);
            atomicAdd(&P[root >> 1], (uint32_t)__popc(m) << UF_SIZE_SHIFT);
            const uint32_t fl = ((m & large) ? UF_LARGE : 0u) | ((m & ext) ? UF_EXT : 0u);
            if (fl) atomicOr(&P[root >> 1], fl);
        }
    }
}

// A finished, flattened 4-connected forest (bare indices) of the bits in `sb` → 8-connected: add the diagonal unions, flatten.
// Same contract as tile_ccl.
__device__ __forceinline__ void tile_ccl_add_diagonals(const uint32_t* sb, uint32_t* P) {
    const int t = threadIdx.x;
    const int r = t / WPR, k = t % WPR;
    const uint32_t* row = sb + (r + 1) * SBW + (k + 1);
    const uint32_t* upr_ = row - SBW;
    const uint32_t w = row[0], wl = row[-1], wr = row[1];
    const uint32_t base = (uint32_t)t * 32u;
    if (w) unite_with_row_above<8, true>(P, base, w, wl, wr, upr_[0], upr_[-1], upr_[1]);
    __syncthreads();
    for (uint32_t st = w & ~(w << 1); st; st &= st - 1) {
        int s = __ffs((int)st) - 1;
        P[(base + s) >> 1] = uf_find_ro(P, base + s);
    }
}

// Barrier that also tells whether the tile is "blobby": few words hold three or more segments.
__device__ __forceinline__ bool sync_and_pick_two_level(uint32_t w) {
    const uint32_t starts = w & ~(w << 1);
    return __syncthreads_count(__popc(starts) >= 3) < TWORDS / 8;
}

// Exclusive prefix sum of one small value per thread over the 256-thread CTA (scratch: 8 words of shared memory).
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t v, uint32_t* scratch8) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    __syncthreads();                       // previous users of scratch8 are done
    if (lane == 31) scratch8[wid] = inc;
    __syncthreads();
    uint32_t off = 0;
#pragma unroll
    for (int i = 0; i < TWORDS / 32; ++i) off += (i < wid) ? scratch8[i] : 0u;
    return off + inc - v;
}

// The tile's TPERIM border pixels spread over the 256 threads in two rounds: round 0 — thread t takes pixel t & 127 of the
// top (t < 128) or bottom row; round 1 — threads below 128 take row t & 63 of the left (t < 64) or right column.
// `off` is the element offset in the Brow (round 0) / Bcol (round 1) array.
struct BorderBases { size_t row0, col0; int Wp, Hp; };
__device__ __forceinline__ BorderBases border_bases(const Geo& g, int n, int ty, int tx) {
    BorderBases b;
    b.row0 = ((size_t)n * g.nty + ty) * 2 * g.Wp + (size_t)tx * TW;
    b.col0 = ((size_t)n * g.ntx + tx) * 2 * g.Hp + (size_t)ty * TH;
    b.Wp = g.Wp; b.Hp = g.Hp;
    return b;
}
__device__ __forceinline__ bool border_entry(int round, const BorderBases& bb, int& br, int& bc, size_t& off) {
    const int t = threadIdx.x;
    if (round == 0) {
        const int side = t >> 7, c = t & (TW - 1);
        br = side ? TH - 1 : 0; bc = c;
        off = bb.row0 + (size_t)side * bb.Wp + c;
        return true;
    }
    if (t >= 2 * TH) return false;
    const int side = t >> 6, r = t & (TH - 1);
    br = r; bc = side ? TW - 1 : 0;
    off = bb.col0 + (size_t)side * bb.Hp + r;
    return true;
}

// Border entry e in [0, TPERIM): 0..TW-1 top row, TW..2TW-1 bottom row,
// then TH left-column, TH right-column entries.
__device__ __forceinline__ void border_px(int e, int& r, int& c) {
    if (e < TW) { r = 0; c = e; }
    else if (e < 2 * TW) { r = TH - 1; c = e - TW; }
    else if (e < 2 * TW + TH) { r = e - 2 * TW; c = 0; }
    else { r = e - 2 * TW - TH; c = TW - 1; }
}
// Where border entry e of tile (n,ty,tx) lives in Brow / Bcol.
__device__ __forceinline__ int32_t* border_slot(int e, const Geo& g, int n, int ty, int tx,
                                                int32_t* Brow, int32_t* Bcol) {
    if (e < 2 * TW) {
        int side = e / TW, c = e % TW;
        return Brow + (((size_t)n * g.nty + ty) * 2 + side) * g.Wp + (size_t)tx * TW + c;
    }
    int q = e - 2 * TW;
    int side = q / TH, r = q % TH;
    return Bcol + (((size_t)n * g.ntx + tx) * 2 + side) * g.Hp + (size_t)ty * TH + r;
}
// Image-linear pixel index (y*W+x) of a tile-linear pixel index.
__device__ __forceinline__ int32_t tile_to_gid(uint32_t tl, const Geo& g, int ty, int tx) {
    int r = (int)(tl / TW), c = (int)(tl % TW);
    return (int32_t)((ty * TH + r) * g.W + tx * TW + c);
}

}  // namespace cs
