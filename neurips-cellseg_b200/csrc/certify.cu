// Certified-large pre-pass of the size filters (remove_small_holes / remove_small_objects, baseline/predict.py:113).
//
// Almost every pixel a size filter looks at belongs to a component that is obviously large: the background of a microscopy
// field, the body of a cell.  Labelling those components only to learn that they stay is what the tile kernel spent most of
// its time on (profiles/r02_phase_cycles_before.txt: the background labelling alone was 35-53 % of it).  This kernel finds them
// without labelling anything: a bit-parallel flood fill.
//
//   unit      FU = 8 rows x 128 columns of one tile (8 units per tile), one THREAD per unit and polarity; a row is one
//             128-bit value — no barriers, no divergence inside the fill
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

constexpr int FU = 8;                  // rows per flood unit
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

// One thread certifies one polarity of one unit: the unit is one 128-byte line of the tile-major mask, read once with
// eight 16-byte loads; rows and flood stay in registers (8 + 8 rows of 128 bits) and the sweeps are unrolled over the rows.
// History (profiles/r02_ncu_certify_*.txt): the kernel's duration is the latency of ONE thread times the number of waves, not
// its instruction count.  16-row units with everything in registers and unrolled: 4096 straight-line instructions walked by
// ~7 warps per SM, stalled on instruction fetch 5 cycles out of 6.  Rows + flood in shared memory, rolled loops: 512 bytes per
// thread, all of an SM's shared memory for 14 warps and a second wave for the last 13 % of a 64-image batch.  Flood in shared
// memory, rows re-read through L1 with look-ahead seeds: ~100 uncoalesced 16-byte loads per thread, LSU-bound.
constexpr int CERT_THREADS = 128;

__device__ __forceinline__ R128 ld128(const uint4* p) {
    const uint4 v = *p;
    return {(unsigned long long)v.x | ((unsigned long long)v.y << 32), (unsigned long long)v.z | ((unsigned long long)v.w << 32)};
}
__device__ __forceinline__ uint4 as_uint4(R128 a) {
    return make_uint4((uint32_t)a.lo, (uint32_t)(a.lo >> 32), (uint32_t)a.hi, (uint32_t)(a.hi >> 32));
}

__global__ void __launch_bounds__(CERT_THREADS) k_certify(const uint32_t* __restrict__ M0, Geo g, uint32_t thr_bg, int b_bg, uint32_t thr_fg, int b_fg,
                                                          uint32_t* __restrict__ Lbg, uint32_t* __restrict__ Lfg, size_t nunits) {
    const size_t gi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t u = gi >> 1;
    if (u >= nunits) return;
    const bool fgpol = gi & 1;
    const uint32_t thr = fgpol ? thr_fg : thr_bg;
    const int b = fgpol ? b_fg : b_bg;
    if (thr <= 1u) return;
    const size_t tile = u / FBANDS;
    const int band = (int)(u % FBANDS);
    const int tx = (int)(tile % g.ntx), ty = (int)((tile / g.ntx) % g.nty);
    const int ncols = min(TW, g.W - tx * TW);
    const int nrows = g.H - (ty * TH + band * FU);                  // valid rows of the unit (may be <= 0 or > FU)
    R128 cv;                                                         // valid columns
    cv.lo = ncols >= 64 ? ~0ull : ((1ull << ncols) - 1ull);
    cv.hi = ncols >= 128 ? ~0ull : (ncols > 64 ? ((1ull << (ncols - 64)) - 1ull) : 0ull);
    const size_t woff = tile * TWORDS + (size_t)band * FU * WPR;
    const uint4* src = reinterpret_cast<const uint4*>(M0 + woff);
    uint4* dst = reinterpret_cast<uint4*>((fgpol ? Lfg : Lbg) + woff);

    R128 w[FU], F[FU];
    R128 all_or{0ull, 0ull}, all_and{~0ull, ~0ull};
#pragma unroll
    for (int i = 0; i < FU; ++i) {
        R128 x = ld128(src + i);
        x = fgpol ? x : ~x;
        w[i] = (i < nrows) ? (x & cv) : R128{0ull, 0ull};
        all_or = all_or | w[i]; all_and = all_and & w[i];
    }
    // Trivial units: nothing of this polarity (the foreground of an empty field) or all of it (its background).
    if (!nz(all_or) || (!nz(~all_and) && (uint32_t)(FU * TW) >= thr)) {
        const uint4 v = nz(all_or) ? make_uint4(~0u, ~0u, ~0u, ~0u) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int i = 0; i < FU; ++i) dst[i] = v;
        return;
    }
    // Seeds: top-left pixels of solid b x b blocks on rows that are multiples of b inside the unit (b = 8: the unit's eight
    // rows; b = 4: rows 0-3 and 4-7) — a solid region of at least 2b - 1 rows holds one.
    bool any = false;
#pragma unroll
    for (int i = 0; i < FU; ++i) F[i] = R128{0ull, 0ull};
    if (b >= 2 && b <= FU) {
#pragma unroll
        for (int i0 = 0; i0 < FU; ++i0) {
            if (i0 % b != 0 || i0 + b > FU) continue;
            R128 v = w[i0];
#pragma unroll
            for (int d = 1; d < FU; ++d)
                if (d < b && i0 + d < FU) v = v & w[i0 + d];
            if (nz(v)) { F[i0] = runs_of(v, b); any |= nz(F[i0]); }
        }
    }
    const bool single = !any;
    if (single) {
        // no aligned solid block: one pixel, verified below by counting what it reaches inside the unit
#pragma unroll
        for (int j = 0; j < FU; ++j) {
            const int i = (FU / 2 + j) % FU;
            if (!any && nz(w[i])) { F[i] = lowest_bit(w[i]); any = true; }
        }
    }
    // A row whose seeds are what it already holds is run-closed from the sweep before: nothing to extend.  (The first sweep
    // has to close the seeds' own runs, so it extends every row.)
    uint32_t cnt = 0;
    bool first = true;
    for (int pass = 0; pass < FLOOD_PAIRS; ++pass) {
        bool changed = false;                                            // by the upward sweep
        R128 prev{0ull, 0ull};
#pragma unroll
        for (int i = 0; i < FU; ++i) {                                   // down
            const R128 s = F[i] | (prev & w[i]);
            if (first || nz(s ^ F[i])) F[i] = fill_runs(s, w[i]);
            prev = F[i];
        }
        first = false;
        cnt = (uint32_t)(__popcll(prev.lo) + __popcll(prev.hi));
#pragma unroll
        for (int i = FU - 2; i >= 0; --i) {                              // up
            const R128 s = F[i] | (prev & w[i]);
            if (nz(s ^ F[i])) { F[i] = fill_runs(s, w[i]); changed = true; }
            prev = F[i];
            cnt += (uint32_t)(__popcll(prev.lo) + __popcll(prev.hi));
        }
        if (!changed) break;             // stable under both sweeps: converged
    }
    const bool keep = !single || cnt >= thr;
#pragma unroll
    for (int i = 0; i < FU; ++i) dst[i] = keep ? as_uint4(F[i]) : make_uint4(0u, 0u, 0u, 0u);
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
    const unsigned grid = (unsigned)((2 * nunits + CERT_THREADS - 1) / CERT_THREADS);      // one thread per (unit, polarity)
    CS_LAUNCH(k_certify, grid, CERT_THREADS, 0, st, M0, g, (uint32_t)max(hole_area, 0), block_edge(hole_area), (uint32_t)max(min_size, 0),
              block_edge(min_size), Lbg, Lfg, nunits);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace cs
