// Certified-large pre-pass of the size filters (remove_small_holes / remove_small_objects, baseline/predict.py:113).
//
// Almost every pixel a size filter looks at belongs to a component that is obviously large: the background of a microscopy
// field, the body of a cell.  Labelling those components only to learn that they stay is what the tile kernel spent most of
// its time on (profiles/r02_phase_cycles_before.txt: the background labelling alone was 35-53 % of it).  This kernel finds them
// without labelling anything: a bit-parallel flood fill.
//
//   unit      FU = 16 rows x 128 columns of one tile (4 units per tile), one THREAD per unit, the 16 rows of 128 bits in
//             registers — no shared memory, no barriers, no divergence inside the fill
//   seeds     top-left pixels of solid b x b blocks, b = ceil(sqrt(threshold)) (8 for the 64-pixel hole filter, 4 for the
//             16-pixel object filter): b*b connected pixels, so the component is large whatever lies outside the unit.
//             A unit without such a block (salt-and-pepper masks) starts from one pixel instead and keeps the fill only if
//             it reaches `threshold` pixels inside the unit.
//   fill      raster sweeps (down, up, down, up): row i takes the filled pixels of the previous row that it shares and
//             extends them along its runs of ones.  The extension is carry propagation: for seeds s inside runs w,
//             ((w ^ (w + s)) & w) | s fills every run from its lowest seed upwards; the same on the bit-reversed row fills
//             downwards.  A sweep therefore crosses the whole unit however the component meanders.
//
// The result L (one bit per pixel, tile-major like the mask) is a SUBSET of the pixels whose component has at least
// `threshold` pixels — not necessarily closed under connectivity (two sweep pairs do not always converge, floods stop at
// unit borders).  The tile kernel labels only what is left (w & ~L) and treats any contact with L as "large" (ccl.cu), so
// the result never depends on how far the flood got; it only decides how little is left to label.
// Both polarities are certified on the INPUT mask: L_bg for hole filling, L_fg for the object filter that runs after it —
// hole filling only adds foreground, so a foreground component that is large before it is part of a large one after it.
#include "ccl_tile.cuh"
#include "internal.cuh"

namespace cs {

constexpr int FU = 16;                 // rows per flood unit
constexpr int FBANDS = TH / FU;        // units per tile
constexpr int FLOOD_PAIRS = 2;         // (down, up) sweep pairs

struct R128 { unsigned long long lo, hi; };   // bit i of the 128-pixel row: i < 64 in lo (pixel 0 = bit 0)

__device__ __forceinline__ R128 operator&(R128 a, R128 b) { return {a.lo & b.lo, a.hi & b.hi}; }
__device__ __forceinline__ R128 operator|(R128 a, R128 b) { return {a.lo | b.lo, a.hi | b.hi}; }
__device__ __forceinline__ R128 operator^(R128 a, R128 b) { return {a.lo ^ b.lo, a.hi ^ b.hi}; }
__device__ __forceinline__ R128 operator~(R128 a) { return {~a.lo, ~a.hi}; }
__device__ __forceinline__ bool nz(R128 a) { return (a.lo | a.hi) != 0ull; }
__device__ __forceinline__ R128 add128(R128 a, R128 b) {
    R128 r; r.lo = a.lo + b.lo; r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull); return r;
}
__device__ __forceinline__ R128 brev128(R128 a) { return {__brevll(a.hi), __brevll(a.lo)}; }
// bit i of the result = bit i + n of a (1 <= n <= 63)
__device__ __forceinline__ R128 shr128(R128 a, int n) { return {(a.lo >> n) | (a.hi << (64 - n)), a.hi >> n}; }
// every run of ones of w that holds a bit of s (s inside w): from its lowest such bit upwards
__device__ __forceinline__ R128 fill_up(R128 s, R128 w) { return ((w ^ add128(w, s)) & w) | s; }
// … and in both directions
__device__ __forceinline__ R128 fill_runs(R128 s, R128 w) {
    return fill_up(s, w) | brev128(fill_up(brev128(s), brev128(w)));
}
// bit i set iff bits i .. i+b-1 of w are all set (2 <= b <= 64)
__device__ __forceinline__ R128 runs_of(R128 w, int b) {
    R128 h = w; int p = 1;
    while (2 * p <= b) { h = h & shr128(h, p); p *= 2; }
    if (b > p) h = h & shr128(h, b - p);
    return h;
}
__device__ __forceinline__ R128 lowest_bit(R128 w) {
    if (w.lo) return {w.lo & (0ull - w.lo), 0ull};
    return {0ull, w.hi & (0ull - w.hi)};
}

// Certify one polarity of one unit.  w[i]: the unit's rows of that polarity.  b: block edge (0 = no block seeds).
__device__ __forceinline__ void certify_unit(const R128 (&w)[FU], R128 (&F)[FU], int b, uint32_t thr) {
    bool any = false;
    if (b >= 2 && b <= FU) {
#pragma unroll
        for (int i = 0; i < FU; ++i) F[i] = runs_of(w[i], b);
        // vertical AND over b consecutive rows, by doubling; rows past the unit count as empty
        int p = 1;
#pragma unroll
        for (int pp = 1; pp < FU; pp *= 2) {
            if (2 * pp <= b) {
#pragma unroll
                for (int i = 0; i < FU; ++i) F[i] = (i + pp < FU) ? (F[i] & F[i + pp]) : R128{0ull, 0ull};
                p = 2 * pp;
            }
        }
        const int rest = b - p;
#pragma unroll
        for (int o = 1; o < FU; ++o) {
            if (o == rest) {
#pragma unroll
                for (int i = 0; i < FU; ++i) F[i] = (i + o < FU) ? (F[i] & F[i + o]) : R128{0ull, 0ull};
            }
        }
#pragma unroll
        for (int i = 0; i < FU; ++i) any |= nz(F[i]);
    } else {
#pragma unroll
        for (int i = 0; i < FU; ++i) F[i] = R128{0ull, 0ull};
    }
    const bool single = !any;
    if (single) {
        // one pixel, verified by counting what it reaches
        constexpr int order[5] = {FU / 2, FU / 4, (3 * FU) / 4, 0, FU - 1};
#pragma unroll
        for (int j = 0; j < 5; ++j)
            if (!any && nz(w[order[j]])) { F[order[j]] = lowest_bit(w[order[j]]); any = true; }
        if (!any) return;                                    // nothing of this polarity in the chosen rows: F stays empty
    }
#pragma unroll
    for (int pass = 0; pass < FLOOD_PAIRS; ++pass) {
#pragma unroll
        for (int i = 0; i < FU; ++i) {
            R128 s = F[i];
            if (i > 0) s = s | (F[i - 1] & w[i]);
            F[i] = fill_runs(s, w[i]);
        }
#pragma unroll
        for (int i = FU - 2; i >= 0; --i) F[i] = fill_runs(F[i] | (F[i + 1] & w[i]), w[i]);
    }
    if (single) {
        uint32_t cnt = 0;
#pragma unroll
        for (int i = 0; i < FU; ++i) cnt += (uint32_t)(__popcll(F[i].lo) + __popcll(F[i].hi));
        if (cnt < thr) {
#pragma unroll
            for (int i = 0; i < FU; ++i) F[i] = R128{0ull, 0ull};
        }
    }
}

__global__ void __launch_bounds__(128) k_certify(const uint32_t* __restrict__ M0, Geo g, uint32_t thr_bg, int b_bg, uint32_t thr_fg, int b_fg,
                                                 uint32_t* __restrict__ Lbg, uint32_t* __restrict__ Lfg, size_t nunits) {
    const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nunits) return;
    const size_t tile = u / FBANDS;
    const int band = (int)(u % FBANDS);
    const int tx = (int)(tile % g.ntx), ty = (int)((tile / g.ntx) % g.nty);
    const int ncols = min(TW, g.W - tx * TW);
    const int nrows = min(FU, g.H - (ty * TH + band * FU));          // valid rows of the unit (may be <= 0)
    R128 cv;                                                         // valid columns
    cv.lo = ncols >= 64 ? ~0ull : ((1ull << ncols) - 1ull);
    cv.hi = ncols >= 128 ? ~0ull : (ncols > 64 ? ((1ull << (ncols - 64)) - 1ull) : 0ull);
    const uint4* src = reinterpret_cast<const uint4*>(M0 + tile * TWORDS + (size_t)band * FU * WPR);
    uint4* dbg = reinterpret_cast<uint4*>(Lbg + tile * TWORDS + (size_t)band * FU * WPR);
    uint4* dfg = reinterpret_cast<uint4*>(Lfg + tile * TWORDS + (size_t)band * FU * WPR);
    R128 m[FU], w[FU], F[FU];
#pragma unroll
    for (int i = 0; i < FU; ++i) {
        const uint4 v = src[i];
        m[i].lo = (unsigned long long)v.x | ((unsigned long long)v.y << 32);
        m[i].hi = (unsigned long long)v.z | ((unsigned long long)v.w << 32);
        m[i] = (i < nrows) ? (m[i] & cv) : R128{0ull, 0ull};
    }
    // background (hole filter)
    if (thr_bg > 1u) {
#pragma unroll
        for (int i = 0; i < FU; ++i) w[i] = (i < nrows) ? (~m[i] & cv) : R128{0ull, 0ull};
        certify_unit(w, F, b_bg, thr_bg);
#pragma unroll
        for (int i = 0; i < FU; ++i)
            dbg[i] = make_uint4((uint32_t)F[i].lo, (uint32_t)(F[i].lo >> 32), (uint32_t)F[i].hi, (uint32_t)(F[i].hi >> 32));
    }
    // foreground (object filter)
    if (thr_fg > 1u) {
        certify_unit(m, F, b_fg, thr_fg);
#pragma unroll
        for (int i = 0; i < FU; ++i)
            dfg[i] = make_uint4((uint32_t)F[i].lo, (uint32_t)(F[i].lo >> 32), (uint32_t)F[i].hi, (uint32_t)(F[i].hi >> 32));
    }
}

static int block_edge(int thr) {
    if (thr <= 1) return 0;
    int b = 1;
    while (b * b < thr) ++b;
    return b <= FU ? b : 0;          // a block taller than a unit cannot be seen inside one
}

// M0: packed tile-major mask of g.  Lbg / Lfg: certified-large background / foreground pixels (same layout); an array of a
// disabled filter (threshold <= 1) is not written.
int certify_large(const uint32_t* M0, const Geo& g, int hole_area, int min_size, uint32_t* Lbg, uint32_t* Lfg, cudaStream_t st) {
    if (hole_area <= 1 && min_size <= 1) return CS_OK;
    const size_t nunits = (size_t)g.N * g.nty * g.ntx * FBANDS;
    const unsigned grid = (unsigned)((nunits + 127) / 128);
    CS_LAUNCH(k_certify, grid, 128, 0, st, M0, g, (uint32_t)max(hole_area, 0), block_edge(hole_area), (uint32_t)max(min_size, 0),
              block_edge(min_size), Lbg, Lfg, nunits);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace cs
