// Mask → instance-map post-processing on device:
//   remove_small_holes(area 64, 4-conn) → remove_small_objects(16, 4-conn) →
//   measure.label(8-conn, ids in raster order of first pixel)
// as called at baseline/predict.py:113 of the reference.
//
// k_certify (certify.cu) first marks, without labelling, the pixels that certainly belong to components of at least `thr`
// pixels (bit-parallel flood fill from solid blocks); everything here works on what is left.
// One tile kernel (k_tile) carries a 64x128-pixel tile through all three stages in a single shared-memory
// residency: per size filter — extend the certified set inside the tile (closure rounds), decide lone runs by bit
// tricks, label the rest (union-find over word segments, sizes at the roots), fill / drop what is small — then the
// 8-connected components of the final foreground → publish.  A component that does not touch the tile border is
// complete and is decided on the spot; a border piece is large if it is `thr` pixels itself or touches a certified pixel
// (inside the tile or across the border, read from the neighbouring tile's certified set), complete if every border
// pixel of it faces a pixel that certainly is of the other polarity.  Only the remaining border pieces are
// "undecided": the tile registers each as a node (tile*256 + k) of a small global union-find (P_s, S_s), records per
// word-segment which node it belongs to (cache_s), appends itself to the stage's stop list and stops.  After the
// cross-tile merge + size accumulation (both driven by the stop list) k_tile_resume processes the listed tiles; it
// applies the decisions from the cache (no second labelling of that stage) and carries on with the remaining stages.
//
// Data layout in HBM
//   M0/M1/M2  packed masks, tile-major: tile (n,ty,tx) is 256 contiguous words (row r, word k at r*4+k);
//             M0 input, M1 the mask of a stopped tile as its stage left it, M2 final foreground
//   Lbg/Lfg   certified-large background / foreground pixels, same layout (k_certify; a stopped tile stores its
//             extended set back so that the resume pass sees the same segmentation)
//   stoplist  tile indices that stopped at stage 1 / 2, appended with an atomic counter
//   Brow_s    [n][ty][top|bottom][Wp], Bcol_s [n][tx][left|right][Hp]: per tile-border pixel of stage s:
//             -1 background, -2 part of a large piece, else node id (s = 1,2) / image-linear root id (s = 3)
//   P_s, S_s  union-find parent / size per node, s = 1 (holes), 2 (objects)
//   cache_s   one byte per word-segment of a stopped tile: 0 keep, 1 remove, 2+k undecided node k
//   cache3    uint16 per word-segment: tile-linear root pixel of its 8-connected component (read by k_final,
//             which therefore never labels again)
//   R         root bitmap, row-major, WSR words per image row (bit set ⇔ pixel is the first pixel of a component)
//   base      exclusive prefix of popc(R) over the whole batch (+1 total entry)
//   P3        union-find parents over image-linear pixel ids for the labelling stage, touched only at roots of
//             border-touching components; it lives in the caller's int32 label buffer until k_final overwrites it
// The only 4-byte-per-pixel traffic is the final label store.
#include "ccl_tile.cuh"
#include "internal.cuh"

namespace cs {

// One- or two-level union-find (tile_ccl): by default the heuristic of sync_and_pick_two_level decides (few words with three or
// more segments → components are large → two levels).  Measured after the move to packed 32-bit entries: forcing the single pass
// everywhere makes the labelling phase cheaper on sparse cell fields (10.9 % against 21.3 % of the tile kernel's cycles) and
// specks (5.2 % / 8.1 %) but the whole chain 4 % slower on dense blobs (0.383 against 0.367 ms), so the heuristic stays.
// CS_LABEL_FORCE / CS_FILTER_FORCE = 0 | 1 in a tools/phase_timing.py build force either variant for comparison.
#ifndef CS_LABEL_FORCE
#define CS_LABEL_FORCE -1
#endif
#ifndef CS_FILTER_FORCE
#define CS_FILTER_FORCE -1
#endif
#ifndef TILE_CTAS_PER_SM
#define TILE_CTAS_PER_SM 8
#endif
constexpr int CLOSURE_ROUNDS = 24;   // upper bound on the in-tile extension rounds of a certified set (stage_filter)
constexpr int32_t B_BG = -1;
constexpr int32_t B_LARGE = -2;

// Phase timers of the tile kernel (tools/phase_timing.py): a -DCS_PHASE_TIMING build accumulates, per phase, the clock
// cycles thread 0 of every CTA spends between two marks (barrier waits included).  Compiled out of the product build.
#ifdef CS_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[128];
struct PhaseTimer {
    long long t0;
    __device__ PhaseTimer() { t0 = clock64(); }
    __device__ void mark(int i) { if (threadIdx.x == 0) { long long t1 = clock64(); atomicAdd(&g_phase_cycles[i], (unsigned long long)(t1 - t0)); t0 = t1; } }
};
#define CS_PHASE(i) cs_phase_timer.mark(i)
#define CS_PHASE_DECL PhaseTimer cs_phase_timer
#define CS_PHASE_ARG , PhaseTimer& cs_phase_timer, int cs_phase_base
#define CS_PHASE_PASS(b) , cs_phase_timer, (b)
#define CS_PHASE_REL(i) cs_phase_timer.mark(cs_phase_base + (i))
#define CS_PHASE_COUNT(i) do { if (threadIdx.x == 0) atomicAdd(&g_phase_cycles[(cs_phase_base / 40) * 40 + (i)], 1ull); } while (0)
#else
#define CS_PHASE(i)
#define CS_PHASE_DECL
#define CS_PHASE_ARG
#define CS_PHASE_PASS(b)
#define CS_PHASE_REL(i)
#define CS_PHASE_COUNT(i)
#endif

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
// global (cross-tile) union-find: smaller index wins
// ---------------------------------------------------------------------------
__device__ __forceinline__ int32_t guf_find(const volatile int32_t* P, int32_t a) {
    int32_t p = P[a];
    while (p != a) { a = p; p = P[a]; }
    return a;
}
// … with path halving (a parent is only ever replaced by an ancestor, so concurrent walkers stay on a valid chain)
__device__ __forceinline__ int32_t guf_find_halve(volatile int32_t* P, int32_t a) {
    int32_t p = P[a];
    while (p != a) {
        const int32_t gp = P[p];
        if (gp != p) P[a] = gp;
        a = p; p = gp;
    }
    return a;
}
__device__ __forceinline__ void guf_unite(int32_t* P, int32_t a, int32_t b) {
    for (;;) {
        a = guf_find_halve(P, a); b = guf_find_halve(P, b);
        if (a == b) return;
        if (a < b) { int32_t t = a; a = b; b = t; }
        if (atomicCAS(&P[a], a, b) == a) return;
    }
}

// ---------------------------------------------------------------------------
// the tile kernel
// ---------------------------------------------------------------------------
struct FilterStage {
    int32_t *Brow, *Bcol;     // border values of this stage
    int32_t *P, *S;           // node union-find
    uint32_t* list;           // node ids of all undecided pieces
    uint32_t* list_n;
    uint8_t* cache;           // [tile][SEGMAX]
    uint32_t* cnt;            // nodes per tile
    uint32_t* L;              // certified-large pixels of this stage's polarity (certify.cu), tile-major; never null when thr > 1.
                              // A tile that stops at this stage stores its extended set back (the resume pass must see the same
                              // segmentation); a neighbour that reads the old or the new word reads a certified set either way.
    uint32_t thr;             // pieces with fewer pixels are removed; ≤ 1 disables the stage
};
struct ChainArgs {
    Geo g;
    const uint32_t* M0; uint32_t* M1; uint32_t* M2;
    FilterStage f1, f2;
    int32_t *Brow3, *Bcol3, *P3;
    uint32_t* R;
    uint16_t* cache3;
    uint32_t* stoplist[2];    // tiles that stopped at stage 1 / 2 …
    uint32_t* stop_n;         // … and how many (two counters)
    int open_rows;            // strip mode: rows beyond the first / last row of g belong to a neighbouring strip (never "outside")
};

// What lies across the tile border, per border pixel of the tile (both filters are 4-connected, so
// only the straight neighbour matters).  Taken from the INPUT mask and the certified sets of the neighbouring tiles, which
// are final before the tile kernel starts:
//   hrow[side][q][k]   word k of the neighbour's facing row (side 0: above, 1: below); q = 0 mask, 1 L_bg, 2 L_fg
//   hcol[side][r]      bits 0..2 = the same three values of the pixel left (side 0) / right (side 1) of tile row r
//   hside[4]           neighbour tile exists: above, below, left, right
struct Halo {
    uint32_t (*hrow)[3][WPR];
    uint8_t (*hcol)[TH];
    uint32_t* hside;
};
struct TileSmem {
    uint32_t* P;        // [TPX/2] union-find entries (ccl_tile.cuh)
    uint32_t* sb;       // [(TH+2)*SBW] zero-padded bits being labelled
    uint32_t* sl;       // [(TH+2)*SBW] zero-padded certified-large bits of the same polarity
    uint32_t* sp;       // [(TH+2)*SBW] zero-padded bits of the whole polarity as the stage found it
    uint32_t* scan;     // [8]
    uint32_t* count;    // [1]
    uint8_t* dec;       // [NODES_PER_TILE]
    Halo halo;
};
constexpr int TILE_SMEM = (TPX / 2) * 4 + 3 * (TH + 2) * SBW * 4 + 16 * 4 + NODES_PER_TILE + 2 * 3 * WPR * 4 + 2 * TH + 4 * 4;

__device__ __forceinline__ TileSmem carve_smem(unsigned char* smem) {
    TileSmem s;
    s.P = reinterpret_cast<uint32_t*>(smem);
    s.sb = s.P + TPX / 2;
    s.sl = s.sb + (TH + 2) * SBW;
    s.sp = s.sl + (TH + 2) * SBW;
    s.scan = s.sp + (TH + 2) * SBW;
    s.count = s.scan + 8;
    s.halo.hrow = reinterpret_cast<uint32_t(*)[3][WPR]>(s.scan + 16);
    s.halo.hside = s.scan + 16 + 2 * 3 * WPR;
    s.dec = reinterpret_cast<uint8_t*>(s.halo.hside + 4);
    s.halo.hcol = reinterpret_cast<uint8_t(*)[TH]>(s.dec + NODES_PER_TILE);
    return s;
}

// Load the halo (all threads call; the caller's next barrier publishes it).
__device__ __forceinline__ void load_halo(const ChainArgs& a, const TileSmem& sm, int tx, int ty, size_t tile) {
    const Geo& g = a.g;
    const int t = threadIdx.x;
    const uint32_t* Lb = a.f1.thr > 1u ? a.f1.L : nullptr;     // (read as they are at this moment, see FilterStage::L)
    const uint32_t* Lf = a.f2.thr > 1u ? a.f2.L : nullptr;
    if (t < 2 * WPR) {                                    // facing rows
        const int side = t / WPR, k = t % WPR;
        const bool ex = side ? (ty + 1 < g.nty) : (ty > 0);
        const size_t w = side ? (tile + g.ntx) * TWORDS + k : (tile - g.ntx) * TWORDS + (TH - 1) * WPR + k;
        sm.halo.hrow[side][0][k] = ex ? a.M0[w] : 0u;
        sm.halo.hrow[side][1][k] = (ex && Lb) ? Lb[w] : 0u;
        sm.halo.hrow[side][2][k] = (ex && Lf) ? Lf[w] : 0u;
        if (k == 0) sm.halo.hside[side] = ex ? 1u : 0u;
    } else if (t >= 64 && t < 64 + 2 * TH) {              // facing columns
        const int side = (t - 64) / TH, r = (t - 64) % TH;
        const bool ex = side ? (tx + 1 < g.ntx) : (tx > 0);
        const size_t w = side ? (tile + 1) * TWORDS + r * WPR : (tile - 1) * TWORDS + r * WPR + (WPR - 1);
        const int sh = side ? 0 : 31;
        uint32_t v = 0u;
        if (ex) v = ((a.M0[w] >> sh) & 1u) | (Lb ? ((Lb[w] >> sh) & 1u) << 1 : 0u) | (Lf ? ((Lf[w] >> sh) & 1u) << 2 : 0u);
        sm.halo.hcol[side][r] = (uint8_t)v;
        if (r == 0) sm.halo.hside[2 + side] = ex ? 1u : 0u;
    }
}

// Links of this thread's word `w` (pixels of the stage's polarity) across the tile border:
//   xl  the neighbour is certified large in the same polarity                     → the piece is large
//   xu  the neighbour may be of the same polarity but is not certified            → the piece may continue there
// Everything else on the border faces a pixel that certainly is of the other polarity (or the outside of the image): if
// all border pixels of a piece are like that the piece is complete.  INVERT (holes, background polarity): the neighbour's
// polarity is exact (the input mask).  Objects (after hole filling): a background neighbour stays background for certain
// only if its component is certified large; anything else may have been filled.
template <bool INVERT>
__device__ __forceinline__ void border_links(uint32_t w, const ChainArgs& a, const TileSmem& sm, int ty, uint32_t& xl, uint32_t& xu) {
    const int t = threadIdx.x, r = t / WPR, k = t % WPR;
    xl = 0u; xu = 0u;
    if (!w) return;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        if (r != (side ? TH - 1 : 0)) continue;
        if (sm.halo.hside[side]) {
            const uint32_t m = sm.halo.hrow[side][0][k], lb = sm.halo.hrow[side][1][k], lf = sm.halo.hrow[side][2][k];
            const uint32_t cert = INVERT ? lb : lf;
            const uint32_t maybe = INVERT ? ~m : ~lb;
            xl |= w & cert;
            xu |= w & maybe & ~cert;
        } else if (a.open_rows && (side ? ty == a.g.nty - 1 : ty == 0)) {
            xu |= w;                                       // strip boundary: the neighbouring strip is reconciled by the host
        }
    }
    if (k == 0 && (w & 1u) && sm.halo.hside[2]) {
        const uint32_t v = sm.halo.hcol[0][r];
        const uint32_t cert = INVERT ? (v >> 1) & 1u : (v >> 2) & 1u;
        const uint32_t maybe = INVERT ? (~v & 1u) : (~(v >> 1) & 1u);
        xl |= cert; xu |= maybe & ~cert;
    }
    if (k == WPR - 1 && (w >> 31) && sm.halo.hside[3]) {
        const uint32_t v = sm.halo.hcol[1][r];
        const uint32_t cert = INVERT ? (v >> 1) & 1u : (v >> 2) & 1u;
        const uint32_t maybe = INVERT ? (~v & 1u) : (~(v >> 1) & 1u);
        xl |= cert << 31; xu |= (maybe & ~cert) << 31;
    }
}

// One size filter on the tile.  `cur` is the thread's word of the current foreground; INVERT works on the background
// (hole filling); `Lw` is the thread's word of certified-large pixels of that polarity.  Only w & ~Lw is labelled; a piece
// of it that touches Lw (inside the tile or across the border) is large.  Returns true (block-uniform) when the tile has
// undecided pieces and must stop; otherwise `cur` is updated.  Border values are published either way.
// `clean` is set when the labelled set was the whole polarity (nothing certified in this tile), i.e. the forest covers it.
template <bool INVERT>
__device__ __forceinline__ bool stage_filter(uint32_t& cur, uint32_t valid, uint32_t Lw, const FilterStage& f, const ChainArgs& a, const TileSmem& sm,
                                             int n, int ty, int tx, size_t tile, bool& clean CS_PHASE_ARG) {
    const Geo& g = a.g;
    const int t = threadIdx.x, r = t / WPR, k = t % WPR;
    const uint32_t base = (uint32_t)t * 32u;
    uint32_t pol = INVERT ? (~cur & valid) : cur;
    uint32_t L = Lw & pol;
    uint32_t xl, xu;                                       // links across the tile border
    border_links<INVERT>(pol, a, sm, ty, xl, xu);
    sm.sl[(r + 1) * SBW + k + 1] = L;
    sm.sp[(r + 1) * SBW + k + 1] = pol;
    if (t == 0) *sm.count = 0u;
    if (L | xl) sm.count[1] = 1u;                          // "something is certified in or next to this tile" (cleared below)
    const bool any_open = __syncthreads_or((pol & ~L) != 0u) != 0;
    if (any_open) {
        // Lone runs: a run of this polarity inside one word with nothing of the polarity above or below it and no possible
        // continuation across a word or tile border is a whole component — decided here by its length.  The masks of a
        // predictor that has not learnt anything are mostly such specks (~1000 per tile); they never reach the union-find.
        const uint32_t* prow = sm.sp + (r + 1) * SBW + (k + 1);
        const uint32_t seeds = (pol & (prow[-SBW] | prow[SBW])) | xl | xu | (pol & 1u & (prow[-1] >> 31)) | (pol & 0x80000000u & (prow[1] << 31));
        uint32_t lone = pol & ~fill_runs32(seeds, pol) & ~L;
        if (lone && f.thr <= 32u) lone &= ~fill_runs32(runs_of32(pol, (int)f.thr), pol);      // … unless it is `thr` pixels long
        pol &= ~lone;
        cur = INVERT ? (cur | lone) : (cur & ~lone);
    }
    // Closure: the flood works on 8-row units and stops after two sweep pairs, so slivers of certified components remain
    // next to them (and next to the certified pixels of the neighbouring tiles).  A few rounds of "take over the straight
    // neighbours' certified pixels, extend along the runs" absorb them (each round: four shared-memory loads, ~25 integer
    // instructions, one barrier) — much cheaper than labelling them.  Reading a neighbour's word while it is being replaced
    // is harmless: both values are certified sets.
    if (any_open && sm.count[1] != 0u) {
        const uint32_t rpol = __brev(pol);
        uint32_t* lrow = sm.sl + (r + 1) * SBW + (k + 1);
        for (int it = 0; it < CLOSURE_ROUNDS; ++it) {
            uint32_t s = pol & (lrow[-SBW] | lrow[SBW] | (L << 1) | (lrow[-1] >> 31) | (L >> 1) | (lrow[1] << 31) | xl) & ~L;
            bool grew = false;
            if (s) {
                s |= L;
                const uint32_t rs = __brev(s);
                L = ((pol ^ (pol + s)) & pol) | s | __brev((rpol ^ (rpol + rs)) & rpol);
                lrow[0] = L;
                grew = true;
            }
            if (__syncthreads_or(grew) == 0) break;      // (letting each warp first iterate inside its own 8-row band was measured: no gain)
        }
    }
    const uint32_t w = pol & ~L;                           // what is left to label
    const uint32_t starts = w & ~(w << 1);
    sm.sb[(r + 1) * SBW + k + 1] = w;
    const BorderBases bb = border_bases(g, n, ty, tx);
    const int nrest = __syncthreads_count(w != 0u);
    clean = false;
    CS_PHASE_REL(0);
    CS_PHASE_COUNT(INVERT ? (nrest ? 21 : 20) : (nrest ? 23 : 22));
    if (nrest == 0) {
        if (t == 0) sm.count[1] = 0u;
        // everything of this polarity is certified: nothing to label, nothing to remove
#pragma unroll
        for (int round = 0; round < 2; ++round) {
            int br, bc; size_t off;
            if (!border_entry(round, bb, br, bc, off)) continue;
            const uint32_t lw = sm.sl[(br + 1) * SBW + (bc >> 5) + 1];
            (round == 0 ? f.Brow : f.Bcol)[off] = ((lw >> (bc & 31)) & 1u) ? B_LARGE : B_BG;
        }
        __syncthreads();
        CS_PHASE_REL(1);
        return false;
    }
    // contact with certified pixels: straight neighbours inside the tile …
    uint32_t large, ext;
    {
        const uint32_t* lrow = sm.sl + (r + 1) * SBW + (k + 1);
        const uint32_t l0 = lrow[0];
        large = w & (lrow[-SBW] | lrow[SBW] | (l0 << 1) | (lrow[-1] >> 31) | (l0 >> 1) | (lrow[1] << 31));
        large |= w & xl; ext = w & xu;                     // … and across the border
    }
    const bool had_cert = sm.count[1] != 0u;
    const bool two_level = CS_FILTER_FORCE >= 0 ? CS_FILTER_FORCE != 0 : __syncthreads_count(__popc(starts) >= 3) < TWORDS / 8;
    tile_ccl<4, true>(sm.sb, sm.P, two_level, large, ext);     // components, sizes and flags at the roots
    __syncthreads();
    if (t == 0) sm.count[1] = 0u;                          // (every thread has read it, barriers ago)
    CS_PHASE_REL(2);
    // roots: small pieces that may continue across the border become nodes of the global union-find
    bool found = false;
    for (uint32_t st = starts; st; st &= st - 1) {
        const uint32_t s = base + (uint32_t)(__ffs((int)st) - 1);
        const uint32_t e = sm.P[s >> 1];
        if ((e & UF_PAR) != s) continue;
        const uint32_t size = (e >> UF_SIZE_SHIFT) & UF_SIZE;
        if (!(e & UF_LARGE) && size < f.thr && (e & UF_EXT)) {
            found = true;
            const uint32_t kk = atomicAdd(sm.count, 1u);
            sm.P[s >> 1] = s | UF_EXT | UF_UNDEC | (kk << UF_SIZE_SHIFT);
            const int32_t node = (int32_t)(tile * NODES_PER_TILE + kk);
            f.P[node] = node; f.S[node] = (int32_t)size;
            f.list[atomicAdd(f.list_n, 1u)] = (uint32_t)node;
        }
    }
    const bool stop = __syncthreads_or(found) != 0;
    // 0 keep (large), 1 remove (complete and small), 2+k undecided node k
    auto verdict = [&](uint32_t root) -> uint32_t {
        const uint32_t e = sm.P[root >> 1];
        if (e & UF_UNDEC) return 2u + ((e >> UF_SIZE_SHIFT) & 0xffu);
        return (!(e & UF_LARGE) && ((e >> UF_SIZE_SHIFT) & UF_SIZE) < f.thr) ? 1u : 0u;
    };
#pragma unroll
    for (int round = 0; round < 2; ++round) {
        int br, bc; size_t off;
        if (!border_entry(round, bb, br, bc, off)) continue;
        const int wi = (br + 1) * SBW + (bc >> 5) + 1, bit = bc & 31;
        int32_t val = B_BG;
        if ((sm.sl[wi] >> bit) & 1u) val = B_LARGE;
        else {
            const uint32_t bw = sm.sb[wi];
            if ((bw >> bit) & 1u) {
                const uint32_t v = verdict(sm.P[((uint32_t)(br * WPR + (bc >> 5)) * 32u + seg_start(bw, bit)) >> 1] & UF_PAR);
                val = v == 0u ? B_LARGE : (v == 1u ? B_BG : (int32_t)(tile * NODES_PER_TILE + (v - 2u)));
            }
        }
        (round == 0 ? f.Brow : f.Bcol)[off] = val;
    }
    CS_PHASE_REL(3);
    if (stop) CS_PHASE_COUNT(INVERT ? 24 : 25);
    if (!stop) {
        uint32_t removed = 0u;
        for (uint32_t st = starts; st; st &= st - 1) {
            int s = __ffs((int)st) - 1;
            if (verdict(sm.P[(base + s) >> 1] & UF_PAR) == 1u) removed |= seg_mask(w, s);
        }
        cur = INVERT ? (cur | removed) : (cur & ~removed);
        __syncthreads();      // everybody is done reading the entries
        if (!had_cert) {
            // the forest spans the whole polarity: strip the flag / size bits so the labelling stage can extend it
            for (uint32_t st = starts; st; st &= st - 1) { const uint32_t s = base + (uint32_t)(__ffs((int)st) - 1); sm.P[s >> 1] &= UF_PAR; }
            clean = true;
        }
        __syncthreads();      // … and with sb / sl before the next stage rewrites them
        CS_PHASE_REL(4);
        return false;
    }
    // stopped: remember, per word-segment, what the resume pass has to do with it
    f.L[tile * TWORDS + t] = L;
    const uint32_t segbase = cta_exclusive_scan((uint32_t)__popc(starts), sm.scan);
    uint8_t* cache = f.cache + tile * SEGMAX + segbase;
    int j = 0;
    for (uint32_t st = starts; st; st &= st - 1, ++j) {
        int s = __ffs((int)st) - 1;
        cache[j] = (uint8_t)verdict(sm.P[(base + s) >> 1] & UF_PAR);
    }
    if (t == 0) f.cnt[tile] = *sm.count;
    CS_PHASE_REL(5);
    return true;
}

// Resume a stopped stage: the global sizes are known now; apply them through the segment cache.
template <bool INVERT>
__device__ __forceinline__ uint32_t stage_resume(uint32_t cur, uint32_t valid, uint32_t Lw, const FilterStage& f, const TileSmem& sm, size_t tile) {
    const int t = threadIdx.x;
    const uint32_t pol = INVERT ? (~cur & valid) : cur;
    const uint32_t w = pol & ~Lw;
    const uint32_t starts = w & ~(w << 1);
    const uint32_t nn = f.cnt[tile];
    if ((uint32_t)t < nn) sm.dec[t] = f.S[guf_find(f.P, (int32_t)(tile * NODES_PER_TILE + t))] >= (int32_t)f.thr ? 1 : 0;
    const uint32_t segbase = cta_exclusive_scan((uint32_t)__popc(starts), sm.scan);   // its barriers also publish dec[]
    const uint8_t* cache = f.cache + tile * SEGMAX + segbase;
    uint32_t removed = 0u;
    int j = 0;
    for (uint32_t st = starts; st; st &= st - 1, ++j) {
        int s = __ffs((int)st) - 1;
        const uint32_t code = cache[j];
        if (code == 1u || (code >= 2u && !sm.dec[code - 2u])) removed |= seg_mask(w, s);
    }
    __syncthreads();
    return INVERT ? (cur | removed) : (cur & ~removed);
}

// Labelling stage: tile-local components of the final foreground; publish the root bitmap, the per-segment
// roots for k_final and, for components touching the tile border, their image-linear root id and a singleton in P3.
// have_forest4 (block-uniform): `P` still holds the flattened 4-connected forest (bare indices) of a superset of w whose
// extra components were removed whole (the object filter just ran over the whole foreground) — only the diagonal unions
// are missing.
template <int CONN>
__device__ __forceinline__ void stage_label(uint32_t w, bool have_forest4, const ChainArgs& a, const TileSmem& sm, int n, int ty, int tx, size_t tile CS_PHASE_ARG) {
    const Geo& g = a.g;
    const int t = threadIdx.x, r = t / WPR, k = t % WPR;
    const uint32_t base = (uint32_t)t * 32u;
    const uint32_t starts = w & ~(w << 1);
    if (__syncthreads_or(w != 0u) == 0) {
        // nothing left in this tile (the usual outcome on the salt-like masks of an untrained predictor: the object filter
        // removed every speck): no roots, background on the whole border
        const int y0 = ty * TH + r;
        if (y0 < g.H) a.R[((size_t)n * g.H + y0) * g.WSR + tx * WPR + k] = 0u;
        const BorderBases bb0 = border_bases(g, n, ty, tx);
#pragma unroll
        for (int round = 0; round < 2; ++round) {
            int br, bc; size_t off;
            if (!border_entry(round, bb0, br, bc, off)) continue;
            (round == 0 ? a.Brow3 : a.Bcol3)[off] = B_BG;
        }
        CS_PHASE_REL(0);
        return;
    }
    sm.sb[(r + 1) * SBW + k + 1] = w;
    const bool two_level = CS_LABEL_FORCE >= 0 ? (__syncthreads(), CS_LABEL_FORCE != 0) : sync_and_pick_two_level(w);
    if (!have_forest4) tile_ccl<CONN, false>(sm.sb, sm.P, two_level);
    else if (CONN == 8) tile_ccl_add_diagonals(sm.sb, sm.P);
    CS_PHASE_REL(0);
    uint32_t roots = 0u;
    for (uint32_t st = starts; st; st &= st - 1) {
        int s = __ffs((int)st) - 1;
        if (sm.P[(base + s) >> 1] == base + s) roots |= 1u << s;
    }
    const int y = ty * TH + r;
    if (y < g.H) a.R[((size_t)n * g.H + y) * g.WSR + tx * WPR + k] = roots;
    const uint32_t segbase = cta_exclusive_scan((uint32_t)__popc(starts), sm.scan);   // barrier: flattened parents visible
    uint16_t* cache = a.cache3 + tile * SEGMAX + segbase;
    int j = 0;
    for (uint32_t st = starts; st; st &= st - 1, ++j) cache[j] = (uint16_t)sm.P[(base + __ffs((int)st) - 1) >> 1];
    CS_PHASE_REL(1);
    int32_t* Pn = a.P3 + (size_t)n * g.HW;
    const BorderBases bb = border_bases(g, n, ty, tx);
#pragma unroll
    for (int round = 0; round < 2; ++round) {
        int br, bc; size_t off;
        if (!border_entry(round, bb, br, bc, off)) continue;
        const uint32_t bw = sm.sb[(br + 1) * SBW + (bc >> 5) + 1];
        int32_t val = B_BG;
        if ((bw >> (bc & 31)) & 1u) {
            const uint32_t root = sm.P[((uint32_t)(br * WPR + (bc >> 5)) * 32u + seg_start(bw, bc & 31)) >> 1];
            val = tile_to_gid(root, g, ty, tx);
            Pn[val] = val;
        }
        (round == 0 ? a.Brow3 : a.Bcol3)[off] = val;
    }
    CS_PHASE_REL(2);
}

// START 0: all tiles from the input mask.  START 1 / 2: only tiles stopped at that stage, resumed.
// CONN 0: no labelling stage (size filters only).
template <int START, int CONN>
__device__ __forceinline__ void tile_body(const ChainArgs& a, const TileSmem& sm, int tx, int ty, int n) {
    const Geo& g = a.g;
    const int t = threadIdx.x, r = t / WPR, k = t % WPR;
    const size_t tile = ((size_t)n * g.nty + ty) * g.ntx + tx;
    CS_PHASE_DECL;

    for (int i = t; i < (TH + 2) * SBW; i += TWORDS) { sm.sb[i] = 0u; sm.sl[i] = 0u; sm.sp[i] = 0u; }      // the zero frames are written once
    if (t == 0) sm.count[1] = 0u;
    const int y = ty * TH + r, x0 = tx * TW + k * 32;
    uint32_t valid = 0u;
    if (y < g.H && x0 < g.W) valid = (g.W - x0 >= 32) ? 0xffffffffu : ((1u << (g.W - x0)) - 1u);
    uint32_t cur;
    if (START != 0) cur = a.M1[tile * TWORDS + t];         // as the stage that stopped left it
    else cur = a.M0[tile * TWORDS + t] & valid;
    const uint32_t L1w = (START <= 1 && a.f1.thr > 1u) ? a.f1.L[tile * TWORDS + t] : 0u;
    const uint32_t L2w = a.f2.thr > 1u ? a.f2.L[tile * TWORDS + t] : 0u;
    if (START <= 1) load_halo(a, sm, tx, ty, tile);
    __syncthreads();

    if (START == 0 && __syncthreads_or(cur != 0u) == 0) {
        // Empty tile (no foreground at all — most tiles of a sparse microscopy field): its background is one component of the
        // tile's valid pixels, nothing can be filled, removed or labelled.  When that component is at least `hole_area` large
        // (every tile except slivers at the image edge) the outputs of all three stages are known without labelling anything:
        // stage-1 border values B_LARGE on valid pixels, stage-2 / stage-3 border values B_BG, empty masks, no roots.
        const int vrows = min(TH, g.H - ty * TH), vcols = min(TW, g.W - tx * TW);
        if ((a.f1.thr <= 1u || (uint32_t)(vrows * vcols) >= a.f1.thr)) {
            const BorderBases bb = border_bases(g, n, ty, tx);
#pragma unroll
            for (int round = 0; round < 2; ++round) {
                int br, bc; size_t off;
                if (!border_entry(round, bb, br, bc, off)) continue;
                const bool v = br < vrows && bc < vcols;
                if (a.f1.thr > 1u) (round == 0 ? a.f1.Brow : a.f1.Bcol)[off] = v ? B_LARGE : B_BG;
                if (a.f2.thr > 1u) (round == 0 ? a.f2.Brow : a.f2.Bcol)[off] = B_BG;
                if (CONN != 0) (round == 0 ? a.Brow3 : a.Bcol3)[off] = B_BG;
            }
            a.M2[tile * TWORDS + t] = 0u;
            if (CONN != 0 && y < g.H) a.R[((size_t)n * g.H + y) * g.WSR + tx * WPR + k] = 0u;
            CS_PHASE(START * 40 + 1);
            return;
        }
    }
    CS_PHASE(START * 40 + 0);
    if (START == 1) cur = stage_resume<true>(cur, valid, L1w, a.f1, sm, tile);
    if (START == 2) cur = stage_resume<false>(cur, valid, L2w, a.f2, sm, tile);
    CS_PHASE(START * 40 + 2);
    bool clean = false;
    if (START == 0 && a.f1.thr > 1u) {
        if (stage_filter<true>(cur, valid, L1w, a.f1, a, sm, n, ty, tx, tile, clean CS_PHASE_PASS(START * 40 + 3))) {
            a.M1[tile * TWORDS + t] = cur;                   // (lone runs already filled)
            if (t == 0) a.stoplist[0][atomicAdd(a.stop_n + 0, 1u)] = (uint32_t)tile;
            return;
        }
    }
    bool have_forest4 = false;
    if (START <= 1 && a.f2.thr > 1u) {
        if (stage_filter<false>(cur, valid, L2w, a.f2, a, sm, n, ty, tx, tile, clean CS_PHASE_PASS(START * 40 + 10))) {
            a.M1[tile * TWORDS + t] = cur;
            if (t == 0) a.stoplist[1][atomicAdd(a.stop_n + 1, 1u)] = (uint32_t)tile;
            return;
        }
        have_forest4 = clean;
    }
    a.M2[tile * TWORDS + t] = cur;
    if (CONN != 0) stage_label<(CONN == 0 ? 8 : CONN)>(cur, have_forest4, a, sm, n, ty, tx, tile CS_PHASE_PASS(START * 40 + 17));
}

// START 0: every tile of the batch, one CTA each.
template <int CONN>
__global__ void __launch_bounds__(TWORDS, TILE_CTAS_PER_SM) k_tile(const ChainArgs a) {
    CS_DYN_SMEM(smem);
    tile_body<0, CONN>(a, carve_smem(smem), blockIdx.x, blockIdx.y, blockIdx.z);
}
// START 1 / 2: the tiles that stopped at that stage (a.stoplist[START - 1], filled by the launches before), dealt out over a
// fixed grid — a field of cells leaves a handful of them, and a launch of one CTA per tile of the batch just to find that out
// cost 10-16 us per stage.
template <int START, int CONN>
__global__ void __launch_bounds__(TWORDS, TILE_CTAS_PER_SM) k_tile_resume(const ChainArgs a) {
    CS_DYN_SMEM(smem);
    const TileSmem sm = carve_smem(smem);
    const Geo& g = a.g;
    const uint32_t cnt = a.stop_n[START - 1];
    for (uint32_t i = blockIdx.x; i < cnt; i += gridDim.x) {
        const uint32_t tile = a.stoplist[START - 1][i];
        const int tx = (int)(tile % (uint32_t)g.ntx), ty = (int)((tile / (uint32_t)g.ntx) % (uint32_t)g.nty);
        const int n = (int)(tile / ((uint32_t)g.ntx * (uint32_t)g.nty));
        tile_body<START, CONN>(a, sm, tx, ty, n);
        __syncthreads();                                   // the next tile reuses the shared memory
    }
}

// ---------------------------------------------------------------------------
// cross-tile passes of a size filter
// ---------------------------------------------------------------------------
// Unions between undecided pieces facing each other across a tile edge (4-connectivity: straight
// neighbours only); a piece facing a large component is itself part of a large component.
__device__ __forceinline__ void filter_edge(int32_t a, int32_t b, uint32_t thr, int32_t* __restrict__ P, int32_t* __restrict__ S) {
    if (a >= 0 && b >= 0) guf_unite(P, a, b);
    else if (a >= 0 && b == B_LARGE) atomicMax(&S[a], (int32_t)thr);
    else if (b >= 0 && a == B_LARGE) atomicMax(&S[b], (int32_t)thr);
}

// Undecided pieces only exist in tiles that stopped, so the stop list drives the pass: one CTA per listed tile walks the
// tile's TPERIM border entries; an entry that holds a node looks at the entry facing it in the neighbouring tile.  (A pair of
// nodes is seen from both sides; the second union finds them united.)  The first version streamed the border arrays of the
// whole batch — 12 MB per pass to find a few hundred nodes on a field of cells.
__global__ void __launch_bounds__(128) k_filter_merge(Geo g, uint32_t thr, const uint32_t* __restrict__ list, const uint32_t* __restrict__ list_n,
                                                      const int32_t* __restrict__ Brow, const int32_t* __restrict__ Bcol,
                                                      int32_t* __restrict__ P, int32_t* __restrict__ S) {
    const uint32_t cnt = *list_n;
    for (uint32_t i = blockIdx.x; i < cnt; i += gridDim.x) {
        const uint32_t tile = list[i];
        const int tx = (int)(tile % (uint32_t)g.ntx), ty = (int)((tile / (uint32_t)g.ntx) % (uint32_t)g.nty);
        const int n = (int)(tile / ((uint32_t)g.ntx * (uint32_t)g.nty));
        for (int e = threadIdx.x; e < TPERIM; e += blockDim.x) {
            int32_t me, other;
            if (e < 2 * TW) {
                const int side = e / TW, c = e % TW;
                const int x = tx * TW + c;
                const int nty_ = side ? ty + 1 : ty - 1;
                if (x >= g.W || nty_ < 0 || nty_ >= g.nty) continue;
                me = Brow[(((size_t)n * g.nty + ty) * 2 + side) * g.Wp + x];
                if (me < 0) continue;
                other = Brow[(((size_t)n * g.nty + nty_) * 2 + (1 - side)) * g.Wp + x];
            } else {
                const int q = e - 2 * TW, side = q / TH, rr = q % TH;
                const int y = ty * TH + rr;
                const int ntx_ = side ? tx + 1 : tx - 1;
                if (y >= g.H || ntx_ < 0 || ntx_ >= g.ntx) continue;
                me = Bcol[(((size_t)n * g.ntx + tx) * 2 + side) * g.Hp + y];
                if (me < 0) continue;
                other = Bcol[(((size_t)n * g.ntx + ntx_) * 2 + (1 - side)) * g.Hp + y];
            }
            filter_edge(me, other, thr, P, S);
        }
    }
}

// Every undecided piece adds its size to its global root.
__global__ void k_filter_accum(const uint32_t* __restrict__ list, const uint32_t* __restrict__ list_n,
                               const int32_t* __restrict__ P, int32_t* __restrict__ S) {
    const uint32_t cnt = *list_n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const int32_t e = (int32_t)list[i];
        const int32_t r = guf_find(P, e);
        if (r != e) atomicAdd(&S[r], S[e]);
    }
}

// ---------------------------------------------------------------------------
// cross-tile passes of the labelling
// ---------------------------------------------------------------------------
// Unions across tile edges (straight, plus diagonal for 8-connectivity; Brow/Bcol span whole image
// rows/columns so corner-crossing diagonals are covered).  A pair whose two ids repeat the previous
// position's pair was already united there.
// Grid: x = chunk of 256 positions along the edge, y = edge inside the image (the nty - 1 horizontal ones, then the ntx - 1
// vertical ones), z = image.
template <int CONN>
__global__ void __launch_bounds__(256) k_label_merge(Geo g, const int32_t* __restrict__ Brow, const int32_t* __restrict__ Bcol, int32_t* __restrict__ P,
                                                     uint32_t* __restrict__ R) {
    const int d = CONN == 8 ? 1 : 0;
    const int n = blockIdx.z, e = blockIdx.y;
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t *mine, *other; int len;
    if (e < g.nty - 1) {
        mine = Brow + (((size_t)n * g.nty + e + 1) * 2 + 0) * g.Wp;
        other = Brow + (((size_t)n * g.nty + e) * 2 + 1) * g.Wp;
        len = g.W;
    } else {
        const int tx = e - (g.nty - 1);
        mine = Bcol + (((size_t)n * g.ntx + tx + 1) * 2 + 0) * g.Hp;
        other = Bcol + (((size_t)n * g.ntx + tx) * 2 + 1) * g.Hp;
        len = g.H;
    }
    if (pos >= len) return;
    const int32_t me = mine[pos];
    if (me < 0) return;
    int32_t* Pn = P + (size_t)n * g.HW;
    uint32_t* Rn = R + (size_t)n * g.H * g.WSR;
    const int32_t prev_me = pos > 0 ? mine[pos - 1] : B_BG;
    for (int dd = -d; dd <= d; ++dd) {
        const int q = pos + dd;
        if (q < 0 || q >= len) continue;
        const int32_t u = other[q];
        if (u < 0) continue;
        bool dup = false;     // (me, u) already united by the previous position's thread?
        if (prev_me == me)
            for (int ee = -d; ee <= d; ++ee) { const int qq = pos - 1 + ee; if (qq >= 0 && qq < len && other[qq] == u) dup = true; }
        if (dup) continue;
        // unite; the root that gets hooked stops being the first pixel of a component: clear its bit in the root bitmap
        int32_t x = me, y = u;
        for (;;) {
            x = guf_find_halve(Pn, x); y = guf_find_halve(Pn, y);
            if (x == y) break;
            if (x < y) { int32_t tmp = x; x = y; y = tmp; }
            if (atomicCAS(&Pn[x], x, y) == x) {
                const int yy = x / g.W, xx = x % g.W;
                atomicAnd(&Rn[(size_t)yy * g.WSR + (xx >> 5)], ~(1u << (xx & 31)));
                break;
            }
        }
    }
}

// Replace every border entry by its global root.  Grid: x = chunk of 1024 entries of one border row / column, y = that row
// or column inside the image (2 nty rows, then 2 ntx columns), z = image.
__global__ void __launch_bounds__(256) k_label_flatten(Geo g, int32_t* __restrict__ Brow, int32_t* __restrict__ Bcol, const int32_t* __restrict__ P) {
    const int n = blockIdx.z, e = blockIdx.y;
    const int pos = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    int32_t* slot; int len;
    if (e < 2 * g.nty) { slot = Brow + ((size_t)n * g.nty * 2 + e) * g.Wp; len = g.Wp; }
    else { slot = Bcol + ((size_t)n * g.ntx * 2 + (e - 2 * g.nty)) * g.Hp; len = g.Hp; }
    if (pos >= len) return;
    slot += pos;
    const int4 v = *reinterpret_cast<const int4*>(slot);
    const int32_t ent[4] = {v.x, v.y, v.z, v.w};
    const int32_t* Pn = P + (size_t)n * g.HW;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int32_t b = ent[j];
        if (b < 0) continue;
        const int32_t rt = guf_find(Pn, b);
        if (rt != b) slot[j] = rt;
    }
}

// Exclusive prefix count of root bits.
constexpr int SCAN_T = 256, SCAN_PER = 32, SCAN_CHUNK = SCAN_T * SCAN_PER;   // 8192 words per CTA: the look-back frontier moves 32 chunks per round, so few large chunks beat many small ones (1024 chunks of 2048 words: 22 us for 2M words)

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

// Exclusive prefix of popc(R) in one pass (decoupled look-back): a CTA takes the next chunk by ticket, publishes the chunk's
// count, then walks back over its predecessors' published counts until it meets one that already knows its prefix.
// state[c]: bits 63..62 = 1 "count of chunk c", 2 "count of chunks 0..c"; low 32 bits the value; state[nchunks] is the ticket.
// The caller zeroes state[0 .. nchunks] before the launch.  Replaces three kernels (reduce, scan of the block sums, apply).
__global__ void __launch_bounds__(SCAN_T) k_scan_lookback(const uint32_t* __restrict__ R, size_t nwords, uint32_t* __restrict__ base,
                                                          unsigned long long* __restrict__ state, int nchunks) {
    __shared__ uint32_t total, s_chunk, s_prefix;
    if (threadIdx.x == 0) s_chunk = (uint32_t)atomicAdd(&state[nchunks], 1ull);
    __syncthreads();
    const uint32_t chunk = s_chunk;
    const size_t b0 = (size_t)chunk * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_PER;
    uint32_t c[SCAN_PER]; uint32_t s = 0;
    const bool whole = b0 + SCAN_PER <= nwords && (reinterpret_cast<uintptr_t>(R) & 15) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0;
    if (whole) {                                             // a thread's 32 words are one 128-byte line: eight 16-byte loads
#pragma unroll
        for (int i = 0; i < SCAN_PER; i += 4) {
            const uint4 v = *reinterpret_cast<const uint4*>(R + b0 + i);
            c[i] = (uint32_t)__popc(v.x); c[i + 1] = (uint32_t)__popc(v.y); c[i + 2] = (uint32_t)__popc(v.z); c[i + 3] = (uint32_t)__popc(v.w);
            s += c[i] + c[i + 1] + c[i + 2] + c[i + 3];
        }
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_PER; ++i) { c[i] = b0 + i < nwords ? (uint32_t)__popc(R[b0 + i]) : 0u; s += c[i]; }
    }
    uint32_t ex = block_exclusive_scan(s, &total);
    if (threadIdx.x < 32) {
        // the first warp looks back 32 predecessors at a time: lane l reads chunk - 1 - l (- 32, - 64, …), the nearest
        // predecessor that already knows its inclusive prefix ends the walk
        volatile unsigned long long* vs = state;
        const int lane = threadIdx.x;
        const uint32_t tot = total;
        if (lane == 0 && chunk > 0) { vs[chunk] = (1ull << 62) | tot; __threadfence(); }
        uint32_t prefix = 0;
        for (int j0 = (int)chunk - 1; j0 >= 0; j0 -= 32) {
            const int j = j0 - lane;
            unsigned long long v = 2ull << 62;             // chunks before 0: an inclusive prefix of 0
            if (j >= 0) do { v = vs[j]; } while ((v >> 62) == 0ull);
            const uint32_t incl = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
            const int stop = __ffs((int)incl) - 1;         // nearest lane holding an inclusive prefix (one always exists at j < 0)
            uint32_t part = (stop < 0 || lane <= stop) ? (uint32_t)v : 0u;
            for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
            prefix += part;
            if (stop >= 0) break;
        }
        if (lane == 0) {
            vs[chunk] = (2ull << 62) | (unsigned long long)(prefix + tot);
            __threadfence();
            s_prefix = prefix;
            if ((int)chunk == nchunks - 1) base[nwords] = prefix + tot;
        }
    }
    __syncthreads();
    ex += s_prefix;
    if (whole) {
#pragma unroll
        for (int i = 0; i < SCAN_PER; i += 4) {
            uint4 o;
            o.x = ex; ex += c[i]; o.y = ex; ex += c[i + 1]; o.z = ex; ex += c[i + 2]; o.w = ex; ex += c[i + 3];
            *reinterpret_cast<uint4*>(base + b0 + i) = o;
        }
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_PER; ++i) { if (b0 + i < nwords) base[b0 + i] = ex; ex += c[i]; }
    }
}
__global__ void k_counts(Geo g, const uint32_t* __restrict__ base, int32_t* __restrict__ n_out) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= g.N) return;
    size_t per = (size_t)g.H * g.WSR;
    n_out[n] = (int32_t)(base[(size_t)(n + 1) * per] - base[(size_t)n * per]);
}

// ---------------------------------------------------------------------------
// final pass: segment roots from the cache → raster rank of the global root → int32 labels
// ---------------------------------------------------------------------------
struct FinalSmem {
    uint32_t* aux;       // [TPX/2] keyed by root pixel >> 1: gid, then label
    uint16_t* segroot;   // [SEGMAX] tile-linear root pixel per word-segment
    uint32_t* sb;        // [(TH+2)*SBW]
    uint16_t* segbase;   // [TWORDS]
    uint32_t* scan;      // [8]
};
constexpr int FINAL_SMEM = (TPX / 2) * 4 + SEGMAX * 2 + (TH + 2) * SBW * 4 + TWORDS * 2 + 8 * 4;

// label_lut (may be null): final label = label_lut[label_offset + local rank]; used by the strip-sharded path,
// where ranks are strip-local and the table holds the global ids (wsi.py).
__global__ void __launch_bounds__(TWORDS, 8)
k_final(const uint32_t* __restrict__ M, Geo g, const uint32_t* __restrict__ R, const uint32_t* __restrict__ base,
        const int32_t* __restrict__ Brow, const int32_t* __restrict__ Bcol, const uint16_t* __restrict__ cache3,
        const int32_t* __restrict__ label_lut, int32_t* __restrict__ labels, int32_t* __restrict__ n_out) {
    CS_DYN_SMEM(smem);
    FinalSmem sm;
    sm.aux = reinterpret_cast<uint32_t*>(smem);
    sm.sb = sm.aux + TPX / 2;
    sm.scan = sm.sb + (TH + 2) * SBW;
    sm.segroot = reinterpret_cast<uint16_t*>(sm.scan + 8);
    sm.segbase = sm.segroot + SEGMAX;
    const int tx = blockIdx.x, ty = blockIdx.y, n = blockIdx.z;
    const int t = threadIdx.x, r = t / WPR, k = t % WPR;
    const uint32_t wbase = (uint32_t)t * 32u;
    const size_t tile = ((size_t)n * g.nty + ty) * g.ntx + tx;

    const uint32_t w = M[tile * TWORDS + t];
    const uint32_t starts = w & ~(w << 1);
    const int lane = t & 31, warp = t >> 5;
    const int x = tx * TW + lane * 4;
    int32_t* out = labels + (size_t)n * g.HW;
    const bool vec = (g.W & 3) == 0;
    if (n_out && tx == 0 && ty == 0 && t == 0) {             // instances of image n
        const size_t iw = (size_t)g.H * g.WSR;
        n_out[n] = (int32_t)(base[(size_t)(n + 1) * iw] - base[(size_t)n * iw]);
    }
    if (__syncthreads_or(w != 0u) == 0) {
        // a tile without foreground (most tiles of a field of cells): zeros, straight away
        for (int rr = warp; rr < TH; rr += TWORDS / 32) {
            const int y = ty * TH + rr;
            if (y >= g.H) break;
            int32_t* dst = out + (size_t)y * g.W + x;
            if (vec) { if (x < g.W) __stcs(reinterpret_cast<int4*>(dst), make_int4(0, 0, 0, 0)); }
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (x + j < g.W) dst[j] = 0;
            }
        }
        return;
    }
    const uint32_t segbase = cta_exclusive_scan((uint32_t)__popc(starts), sm.scan);
    sm.sb[(r + 1) * SBW + k + 1] = w;
    sm.segbase[t] = (uint16_t)segbase;
    {
        const uint16_t* cache = cache3 + tile * SEGMAX + segbase;
        int j = 0;
        for (uint32_t st = starts; st; st &= st - 1, ++j) {
            const uint32_t s = wbase + __ffs((int)st) - 1;
            const uint16_t root = cache[j];
            sm.segroot[segbase + j] = root;
            if (root == s) sm.aux[s >> 1] = (uint32_t)tile_to_gid(s, g, ty, tx);
        }
    }
    __syncthreads();
    // components touching the tile border take the global root recorded there
    const BorderBases bb = border_bases(g, n, ty, tx);
#pragma unroll
    for (int round = 0; round < 2; ++round) {
        int br, bc; size_t off;
        if (!border_entry(round, bb, br, bc, off)) continue;
        const uint32_t bw = sm.sb[(br + 1) * SBW + (bc >> 5) + 1];
        if ((bw >> (bc & 31)) & 1u) {
            const int wt = br * WPR + (bc >> 5);
            const uint32_t root = sm.segroot[sm.segbase[wt] + seg_ordinal(bw, bc & 31)];
            sm.aux[root >> 1] = (uint32_t)(round == 0 ? Brow : Bcol)[off];
        }
    }
    __syncthreads();
    const size_t img_words = (size_t)g.H * g.WSR;
    const uint32_t img_base = base[(size_t)n * img_words];
    {
        int j = 0;
        for (uint32_t st = starts; st; st &= st - 1, ++j) {
            const uint32_t s = wbase + __ffs((int)st) - 1;
            if (sm.segroot[segbase + j] == s) {
                const int32_t gid = (int32_t)sm.aux[s >> 1];
                const int y = gid / g.W, x = gid % g.W;
                const size_t wi = (size_t)n * img_words + (size_t)y * g.WSR + (x >> 5);
                uint32_t lab = base[wi] - img_base + (uint32_t)__popc(R[wi] & ((1u << (x & 31)) - 1u)) + 1u;
                if (label_lut) lab = (uint32_t)label_lut[lab];
                sm.aux[s >> 1] = lab;
            }
        }
    }
    __syncthreads();
    // store: a warp covers one 128-pixel tile row per iteration, 4 pixels per lane
    const int kk = lane >> 3, b0 = (lane & 7) * 4;
    for (int rr = warp; rr < TH; rr += TWORDS / 32) {
        const int y = ty * TH + rr;
        if (y >= g.H) break;
        const uint32_t rw = sm.sb[(rr + 1) * SBW + kk + 1];
        const uint32_t nib = (rw >> b0) & 0xfu;
        int32_t v[4] = {0, 0, 0, 0};
        if (nib) {
            const uint32_t rstarts = rw & ~(rw << 1);
            const uint32_t sb0 = sm.segbase[rr * WPR + kk];
            const int first = __ffs((int)nib) - 1;
            const uint32_t shifted = nib >> first;
            if (((shifted + 1u) & shifted) == 0u) {
                // the quad's foreground pixels are one run (the rule on cell-like masks): one look-up serves them all
                const int b = b0 + first;
                const int32_t lab = (int32_t)sm.aux[sm.segroot[sb0 + __popc(rstarts & ((2u << b) - 1u)) - 1] >> 1];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = ((nib >> j) & 1u) ? lab : 0;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int b = b0 + j;
                    if ((rw >> b) & 1u) v[j] = (int32_t)sm.aux[sm.segroot[sb0 + __popc(rstarts & ((2u << b) - 1u)) - 1] >> 1];
                }
            }
        }
        int32_t* dst = out + (size_t)y * g.W + x;
        if (vec) { if (x < g.W) __stcs(reinterpret_cast<int4*>(dst), make_int4(v[0], v[1], v[2], v[3])); }
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (x + j < g.W) dst[j] = v[j];
        }
    }
}

// ---------------------------------------------------------------------------
// host-side orchestration
// ---------------------------------------------------------------------------
struct StageWs { int32_t *Brow, *Bcol, *P, *S; uint32_t *list, *list_n, *cnt; uint8_t* cache; int32_t *canon, *myroot; };
struct PostWs {
    uint32_t *M0, *M1, *M2, *Lbg, *Lfg, *R, *base, *counters, *stoplist[2];
    unsigned long long* scanstate;
    StageWs s1, s2;
    int32_t *Brow3, *Bcol3;
    uint16_t* cache3;
    size_t bytes;
    int nscan;
};
// strip = true adds what the strip-sharded (multi-GPU) path needs: room for the neighbours' boundary pieces in the
// node union-find (2*Wp foreign nodes), canonical-x scratch and this strip's boundary-row roots.
static PostWs carve(void* ws, size_t cap, const Geo& g, bool strip = false) {
    Arena a(ws, cap);
    PostWs p;
    const size_t ntiles = (size_t)g.N * g.nty * g.ntx;
    const size_t nwords = (size_t)g.N * g.H * g.WSR;
    const size_t nbrow = (size_t)g.N * g.nty * 2 * g.Wp, nbcol = (size_t)g.N * g.ntx * 2 * g.Hp;
    const size_t nnodes = ntiles * NODES_PER_TILE;
    p.nscan = (int)((nwords + SCAN_CHUNK - 1) / SCAN_CHUNK);
    p.M0 = a.take<uint32_t>(ntiles * TWORDS);
    p.M1 = a.take<uint32_t>(ntiles * TWORDS);
    p.M2 = a.take<uint32_t>(ntiles * TWORDS);
    p.Lbg = a.take<uint32_t>(ntiles * TWORDS);
    p.Lfg = a.take<uint32_t>(ntiles * TWORDS);
    p.R = a.take<uint32_t>(nwords);
    p.base = a.take<uint32_t>(nwords + 1);
    p.scanstate = a.take<unsigned long long>((size_t)p.nscan + 1);
    p.counters = a.take<uint32_t>(64);
    p.stoplist[0] = a.take<uint32_t>(ntiles);
    p.stoplist[1] = a.take<uint32_t>(ntiles);
    StageWs* st[2] = {&p.s1, &p.s2};
    for (int i = 0; i < 2; ++i) {
        st[i]->Brow = a.take<int32_t>(nbrow);
        st[i]->Bcol = a.take<int32_t>(nbcol);
        st[i]->P = a.take<int32_t>(nnodes + (strip ? 2 * (size_t)g.Wp : 0));
        st[i]->S = a.take<int32_t>(nnodes + (strip ? 2 * (size_t)g.Wp : 0));
        st[i]->canon = strip ? a.take<int32_t>(2 * nnodes) : nullptr;
        st[i]->myroot = strip ? a.take<int32_t>(2 * (size_t)g.Wp) : nullptr;
        st[i]->list = a.take<uint32_t>(nnodes);
        st[i]->cnt = a.take<uint32_t>(ntiles);
        st[i]->cache = a.take<uint8_t>(ntiles * SEGMAX);
        st[i]->list_n = p.counters + i;
    }
    p.Brow3 = a.take<int32_t>(nbrow);
    p.Bcol3 = a.take<int32_t>(nbcol);
    p.cache3 = a.take<uint16_t>(ntiles * SEGMAX);
    p.bytes = a.off;
    return p;
}

static int check_geo(int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0) { set_error("image batch must have positive N, H, W (got %d, %d, %d)", N, H, W); return CS_ERR_ARG; }
    if ((long long)H * W > 0x7fffffffLL) { set_error("H*W = %lld exceeds the 32-bit pixel-id range", (long long)H * W); return CS_ERR_RANGE; }
    if (ceil_div(H, TH) + ceil_div(W, TW) > 32767 || N > 65535) { set_error("grid dimension out of range"); return CS_ERR_RANGE; }
    if ((long long)N * ceil_div(H, TH) * ceil_div(W, TW) > (0x7fffffffLL / NODES_PER_TILE)) { set_error("too many tiles for 32-bit node ids"); return CS_ERR_RANGE; }
    return CS_OK;
}

size_t postprocess_workspace_bytes(int N, int H, int W) {
    if (check_geo(N, H, W) != CS_OK) return 0;
    Geo g = make_geo(N, H, W);
    return carve(nullptr, 0, g).bytes;
}

template <int START, int CONN>
static int set_attr() {
    if (START == 0) CS_CUDA(cudaFuncSetAttribute(k_tile<CONN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM));
    else CS_CUDA(cudaFuncSetAttribute(k_tile_resume<(START == 0 ? 1 : START), CONN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM));
    return CS_OK;
}
static bool g_attr_done = false;
static int ensure_attrs() {
    if (g_attr_done) return CS_OK;
    int rc;
    if ((rc = set_attr<0, 0>()) || (rc = set_attr<0, 4>()) || (rc = set_attr<0, 8>()) ||
        (rc = set_attr<1, 0>()) || (rc = set_attr<1, 4>()) || (rc = set_attr<1, 8>()) ||
        (rc = set_attr<2, 0>()) || (rc = set_attr<2, 4>()) || (rc = set_attr<2, 8>())) return rc;
    CS_CUDA(cudaFuncSetAttribute(k_final, cudaFuncAttributeMaxDynamicSharedMemorySize, FINAL_SMEM));
    g_attr_done = true;
    return CS_OK;
}

static int flat_grid(size_t work, int threads) {
    size_t b = (work + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > 148 * 16) b = 148 * 16;
    return (int)b;
}

static int resume_grid(const Geo& g) {
    const size_t ntiles = (size_t)g.N * g.nty * g.ntx;
    return (int)(ntiles < (size_t)148 * TILE_CTAS_PER_SM ? ntiles : (size_t)148 * TILE_CTAS_PER_SM);
}
#define CS_TILE_LAUNCH(CONN, NAME)                                                   \
    {                                                                                 \
        constexpr auto NAME = k_tile<CONN>;                                           \
        CS_LAUNCH(NAME, grid, TWORDS, TILE_SMEM, st, a);                              \
    }
#define CS_RESUME_LAUNCH(START, CONN, NAME)                                          \
    {                                                                                 \
        constexpr auto NAME = k_tile_resume<START, CONN>;                             \
        CS_LAUNCH(NAME, resume_grid(a.g), TWORDS, TILE_SMEM, st, a);                  \
    }
static int launch_tile(int start, const ChainArgs& a, int conn, cudaStream_t st) {
    dim3 grid(a.g.ntx, a.g.nty, a.g.N);
    const int key = start * 10 + conn;
    switch (key) {
        case 0:  CS_TILE_LAUNCH(0, k_tile_fresh_filters) break;
        case 4:  CS_TILE_LAUNCH(4, k_tile_fresh_label4) break;
        case 8:  CS_TILE_LAUNCH(8, k_tile_fresh_label8) break;
        case 10: CS_RESUME_LAUNCH(1, 0, k_tile_resume1_filters) break;
        case 14: CS_RESUME_LAUNCH(1, 4, k_tile_resume1_label4) break;
        case 18: CS_RESUME_LAUNCH(1, 8, k_tile_resume1_label8) break;
        case 20: CS_RESUME_LAUNCH(2, 0, k_tile_resume2_filters) break;
        case 24: CS_RESUME_LAUNCH(2, 4, k_tile_resume2_label4) break;
        case 28: CS_RESUME_LAUNCH(2, 8, k_tile_resume2_label8) break;
        default: set_error("launch_tile: bad stage/connectivity %d/%d", start, conn); return CS_ERR_ARG;
    }
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// Size filters (thr ≤ 1 disables one) and, for conn = 4 | 8, the labelling passes up to the prefix scan.
// The final foreground ends up in p.M2.  P3 is the caller's int32 label buffer (conn != 0).
static ChainArgs build_args(const Geo& g, const PostWs& p, const uint32_t* M0, int hole_area, int min_size, int32_t* P3, bool strip = false) {
    ChainArgs a;
    a.g = g; a.M0 = M0; a.M1 = p.M1; a.M2 = p.M2;
    a.f1.L = p.Lbg; a.f2.L = p.Lfg; a.open_rows = strip ? 1 : 0;
    const StageWs* sw[2] = {&p.s1, &p.s2};
    FilterStage* fs[2] = {&a.f1, &a.f2};
    const int thr[2] = {hole_area, min_size};
    for (int i = 0; i < 2; ++i) {
        fs[i]->Brow = sw[i]->Brow; fs[i]->Bcol = sw[i]->Bcol; fs[i]->P = sw[i]->P; fs[i]->S = sw[i]->S;
        fs[i]->list = sw[i]->list; fs[i]->list_n = sw[i]->list_n; fs[i]->cache = sw[i]->cache; fs[i]->cnt = sw[i]->cnt;
        fs[i]->thr = thr[i] > 1 ? (uint32_t)thr[i] : 0u;
    }
    a.Brow3 = p.Brow3; a.Bcol3 = p.Bcol3; a.P3 = P3; a.R = p.R; a.cache3 = p.cache3;
    a.stoplist[0] = p.stoplist[0]; a.stoplist[1] = p.stoplist[1]; a.stop_n = p.counters + 2;
    return a;
}

// Cross-tile unions and size accumulation of one filter stage (inside one image / strip).
static int run_filter_merge(const Geo& g, const FilterStage& f, const uint32_t* stoplist, const uint32_t* stop_n, cudaStream_t st) {
    const size_t ntiles = (size_t)g.N * g.nty * g.ntx;
    const int grid = (int)(ntiles < 148 * 16 ? ntiles : 148 * 16);
    CS_LAUNCH(k_filter_merge, grid, 128, 0, st, g, f.thr, stoplist, stop_n, f.Brow, f.Bcol, f.P, f.S);
    CS_LAUNCH_CHECK();
    CS_LAUNCH(k_filter_accum, flat_grid(ntiles * 4, 256), 256, 0, st, f.list, f.list_n, f.P, f.S);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

static int run_tiles(const Geo& g, const PostWs& p, const uint32_t* M0, int hole_area, int min_size, int conn, int32_t* P3, cudaStream_t st) {
    ChainArgs a = build_args(g, p, M0, hole_area, min_size, P3);
    CS_CUDA(cudaMemsetAsync(p.counters, 0, 64 * sizeof(uint32_t), st));
    int rc;
    if ((rc = certify_large(M0, g, hole_area, min_size, p.Lbg, p.Lfg, st)) != CS_OK) return rc;
    if ((rc = launch_tile(0, a, conn, st)) != CS_OK) return rc;
    const FilterStage* fs[2] = {&a.f1, &a.f2};
    for (int i = 0; i < 2; ++i) {
        if (!fs[i]->thr) continue;
        if ((rc = run_filter_merge(g, *fs[i], a.stoplist[i], a.stop_n + i, st)) != CS_OK) return rc;
        if ((rc = launch_tile(i + 1, a, conn, st)) != CS_OK) return rc;
    }
    return CS_OK;
}

// Exclusive prefix count of the set bits of a bitmap: base[i] = popc(R[0..i)), base[nwords] = total.
int scan_bitmap(const uint32_t* R, size_t nwords, uint32_t* base, unsigned long long* state, cudaStream_t st) {
    const int nscan = (int)((nwords + SCAN_CHUNK - 1) / SCAN_CHUNK);
    CS_CUDA(cudaMemsetAsync(state, 0, ((size_t)nscan + 1) * sizeof(unsigned long long), st));
    CS_LAUNCH(k_scan_lookback, nscan, SCAN_T, 0, st, R, nwords, base, state, nscan);
    CS_LAUNCH_CHECK();
    return CS_OK;
}
size_t scan_blocks(size_t nwords) { return (nwords + SCAN_CHUNK - 1) / SCAN_CHUNK + 1; }

// Cross-tile label unions, root bitmap clean-up and the raster-rank prefix scan.
static int run_label_merge_scan(const Geo& g, const PostWs& p, int conn, int32_t* labels, int32_t* n_out, cudaStream_t st) {
    const size_t nwords = (size_t)g.N * g.H * g.WSR;
    const size_t edges = (size_t)g.N * (g.nty - 1) * g.W + (size_t)g.N * (g.ntx - 1) * g.H;
    if (edges) {
        const int longest = g.W > g.H ? g.W : g.H;
        const dim3 mg(ceil_div(longest, 256), (g.nty - 1) + (g.ntx - 1), g.N);
        if (conn == 8) CS_LAUNCH(k_label_merge<8>, mg, 256, 0, st, g, p.Brow3, p.Bcol3, labels, p.R);
        else           CS_LAUNCH(k_label_merge<4>, mg, 256, 0, st, g, p.Brow3, p.Bcol3, labels, p.R);
        CS_LAUNCH_CHECK();
        const int longestp = g.Wp > g.Hp ? g.Wp : g.Hp;
        const dim3 fg(ceil_div(longestp, 1024), 2 * g.nty + 2 * g.ntx, g.N);
        CS_LAUNCH(k_label_flatten, fg, 256, 0, st, g, p.Brow3, p.Bcol3, labels);
        CS_LAUNCH_CHECK();
    }
    int rc = scan_bitmap(p.R, nwords, p.base, p.scanstate, st);
    if (rc != CS_OK) return rc;
    if (n_out) {
        CS_LAUNCH(k_counts, ceil_div(g.N, 128), 128, 0, st, g, p.base, n_out);
        CS_LAUNCH_CHECK();
    }
    return CS_OK;
}

static int run_final(const Geo& g, const PostWs& p, int32_t* labels, const int32_t* label_lut, int32_t* n_out, cudaStream_t st) {
    dim3 grid(g.ntx, g.nty, g.N);
    CS_LAUNCH(k_final, grid, TWORDS, FINAL_SMEM, st, p.M2, g, p.R, p.base, p.Brow3, p.Bcol3, p.cache3, label_lut, labels, n_out);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

static int run_label_global(const Geo& g, const PostWs& p, int conn, int32_t* labels, int32_t* n_out, const int32_t* label_lut, cudaStream_t st) {
    int rc = run_label_merge_scan(g, p, conn, labels, nullptr, st);      // the per-image counts come out of k_final
    if (rc != CS_OK) return rc;
    return run_final(g, p, labels, label_lut, n_out, st);
}

static int prep(int N, int H, int W, void* ws, size_t ws_bytes, Geo& g, PostWs& p, bool strip = false) {
    int rc = check_geo(N, H, W);
    if (rc != CS_OK) return rc;
    g = make_geo(N, H, W);
    p = carve(ws, ws_bytes, g, strip);
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
    int rc = prep(N, H, W, ws, ws_bytes, g, p);
    if (rc != CS_OK) return rc;
    if ((rc = pack_mask(mask, g, p.M0, st)) != CS_OK) return rc;
    if ((rc = run_tiles(g, p, p.M0, hole_area, min_size, connectivity, labels, st)) != CS_OK) return rc;
    return run_label_global(g, p, connectivity, labels, n_out, nullptr, st);
}

// Same chain starting from an already packed (tile-major) mask living in the workspace's M0.
int postprocess_packed(int N, int H, int W, int hole_area, int min_size, int connectivity,
                       int32_t* labels, int32_t* n_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    CS_CHECK_ARG(labels, "null labels pointer");
    CS_CHECK_ARG(connectivity == 4 || connectivity == 8, "connectivity must be 4 or 8");
    CS_CHECK_ARG(hole_area >= 0 && min_size >= 0, "negative size threshold");
    Geo g; PostWs p;
    int rc = prep(N, H, W, ws, ws_bytes, g, p);
    if (rc != CS_OK) return rc;
    if ((rc = run_tiles(g, p, p.M0, hole_area, min_size, connectivity, labels, st)) != CS_OK) return rc;
    return run_label_global(g, p, connectivity, labels, n_out, nullptr, st);
}

uint32_t* packed_mask_slot(void* ws, size_t ws_bytes, int N, int H, int W) {
    if (check_geo(N, H, W) != CS_OK) return nullptr;
    Geo g = make_geo(N, H, W);
    PostWs p = carve(ws, ws_bytes, g);
    return (ws && p.bytes <= ws_bytes) ? p.M0 : nullptr;
}

// The certified-large pre-pass on its own (diagnostic / test entry): u8 mask → u8 0/1 maps of the background and foreground
// pixels certify.cu proves to lie in components of at least hole_area / min_size pixels.
int certify_masks(const uint8_t* mask, int N, int H, int W, int hole_area, int min_size, uint8_t* large_bg, uint8_t* large_fg,
                  void* ws, size_t ws_bytes, cudaStream_t st) {
    CS_CHECK_ARG(mask && large_bg && large_fg, "null pointer");
    CS_CHECK_ARG(hole_area >= 0 && min_size >= 0, "negative size threshold");
    Geo g; PostWs p;
    int rc = prep(N, H, W, ws, ws_bytes, g, p);
    if (rc != CS_OK) return rc;
    if ((rc = pack_mask(mask, g, p.M0, st)) != CS_OK) return rc;
    const size_t bytes = (size_t)N * g.nty * g.ntx * TWORDS * sizeof(uint32_t);
    CS_CUDA(cudaMemsetAsync(p.Lbg, 0, bytes, st));
    CS_CUDA(cudaMemsetAsync(p.Lfg, 0, bytes, st));
    if ((rc = certify_large(p.M0, g, hole_area, min_size, p.Lbg, p.Lfg, st)) != CS_OK) return rc;
    CS_LAUNCH(k_unpack_u8, dim3(g.ntx, g.nty, g.N), TWORDS, 0, st, p.Lbg, large_bg, g);
    CS_LAUNCH_CHECK();
    CS_LAUNCH(k_unpack_u8, dim3(g.ntx, g.nty, g.N), TWORDS, 0, st, p.Lfg, large_fg, g);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// measure.label on its own.
int label(const uint8_t* mask, int N, int H, int W, int connectivity, int32_t* labels, int32_t* n_out,
          void* ws, size_t ws_bytes, cudaStream_t st) {
    return postprocess(mask, N, H, W, 0, 0, connectivity, labels, n_out, ws, ws_bytes, st);
}

// remove_small_objects (invert=0) / remove_small_holes (invert=1) on a u8 mask.
int remove_small(const uint8_t* mask_in, uint8_t* mask_out, int N, int H, int W, int connectivity, int min_size, int invert,
                 void* ws, size_t ws_bytes, cudaStream_t st) {
    CS_CHECK_ARG(mask_in && mask_out, "null pointer");
    CS_CHECK_ARG(connectivity == 4, "size filters implement the reference's default 4-connectivity only");
    CS_CHECK_ARG(min_size >= 0, "negative size threshold");
    Geo g; PostWs p;
    int rc = prep(N, H, W, ws, ws_bytes, g, p);
    if (rc != CS_OK) return rc;
    if ((rc = pack_mask(mask_in, g, p.M0, st)) != CS_OK) return rc;
    if ((rc = run_tiles(g, p, p.M0, invert ? min_size : 0, invert ? 0 : min_size, 0, nullptr, st)) != CS_OK) return rc;
    CS_LAUNCH(k_unpack_u8, dim3(g.ntx, g.nty, g.N), TWORDS, 0, st, p.M2, mask_out, g);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// ---------------------------------------------------------------------------
// strip-sharded variant: one row strip of a whole-slide image per GPU (wsi.py)
// ---------------------------------------------------------------------------
// A strip is post-processed like an image of its own; the pieces of components that touch its first / last row
// are reconciled with the neighbouring strips between the phases below.  The host (torch.distributed over NCCL)
// moves the exported rows; every kernel here works on one strip.
//   begin                     k_tile fresh
//   for stage s = 1, 2:       filter_export(s) → [all-gather rows] → filter_resume(s)
//   label_export              → [all-gather rows + counts] → label_resolve → final(lut)
// Strips other than the last must be a multiple of TH rows (their last row is then a tile border), and a strip
// with neighbours on both sides must have at least max(hole_area, min_size) rows, so that a piece smaller than
// the threshold never reaches across a whole strip.

// Row `side` (0 top, 1 bottom) of this strip's stage border values.
__device__ __forceinline__ const int32_t* strip_row(const Geo& g, const int32_t* Brow, int side) {
    return Brow + ((size_t)(side ? g.nty - 1 : 0) * 2 + side) * g.Wp;
}

// Export the undecided pieces on the strip's first and last rows: out[side][0][x] = canonical id of the piece
// (smallest x at which it appears in that row), -1 background, -2 large; out[side][1][x] = its strip-level size.
// One CTA per side.
__global__ void __launch_bounds__(1024) k_strip_filter_export(Geo g, FilterStage f, int32_t* __restrict__ canon, int32_t* __restrict__ myroot,
                                                               int32_t* __restrict__ out, size_t nnodes) {
    const int side = blockIdx.x;
    const int32_t* row = strip_row(g, f.Brow, side);
    int32_t* cn = canon + (size_t)side * nnodes;
    int32_t* mr = myroot + (size_t)side * g.Wp;
    int32_t* oid = out + (size_t)side * 2 * g.W;
    int32_t* osz = oid + g.W;
    for (int x = threadIdx.x; x < g.W; x += blockDim.x) {
        const int32_t b = row[x];
        int32_t r = b;
        if (b >= 0) {
            r = guf_find(f.P, b);
            if (f.S[r] >= (int32_t)f.thr) r = B_LARGE; else cn[r] = 0x7fffffff;
        }
        mr[x] = r;
    }
    __syncthreads();
    for (int x = threadIdx.x; x < g.W; x += blockDim.x) { const int32_t r = mr[x]; if (r >= 0) atomicMin(&cn[r], x); }
    __syncthreads();
    for (int x = threadIdx.x; x < g.W; x += blockDim.x) {
        const int32_t r = mr[x];
        oid[x] = r >= 0 ? cn[r] : r;
        osz[x] = r >= 0 ? f.S[r] : 0;
    }
}

// Import the facing row of a neighbouring strip (nb = [2][W]: ids, sizes) on `side` of this strip: the
// neighbour's pieces become nodes nnodes + side*Wp + id of this strip's union-find; straight (4-connectivity)
// contacts across the cut are united, sizes re-accumulated at the new roots, contact with a large piece marks
// the root large.  One CTA per call.
__global__ void __launch_bounds__(1024) k_strip_filter_import(Geo g, FilterStage f, const int32_t* __restrict__ myroot, const int32_t* __restrict__ nb,
                                                               int side, size_t nnodes) {
    const int32_t* mr = myroot + (size_t)side * g.Wp;
    const int32_t* nid = nb;
    const int32_t* nsz = nb + g.W;
    const int32_t fbase = (int32_t)(nnodes + (size_t)side * g.Wp);
    for (int x = threadIdx.x; x < g.W; x += blockDim.x) {
        const int32_t id = nid[x];
        if (id >= 0) { f.P[fbase + id] = fbase + id; f.S[fbase + id] = nsz[x]; }
    }
    __syncthreads();
    for (int x = threadIdx.x; x < g.W; x += blockDim.x) {
        const int32_t m = mr[x], id = nid[x];
        if (m >= 0 && id >= 0) guf_unite(f.P, m, fbase + id);
    }
    __syncthreads();
    for (int x = threadIdx.x; x < g.W; x += blockDim.x) {
        const int32_t m = mr[x], id = nid[x];
        if (m >= 0) {
            const int32_t rr = guf_find(f.P, m);
            if (rr != m) { const int32_t v = atomicExch(&f.S[m], 0); if (v) atomicAdd(&f.S[rr], v); }
        }
        if (id >= 0) {
            const int32_t F = fbase + id, rr = guf_find(f.P, F);
            if (rr != F) { const int32_t v = atomicExch(&f.S[F], 0); if (v) atomicAdd(&f.S[rr], v); }
        }
    }
    __syncthreads();
    for (int x = threadIdx.x; x < g.W; x += blockDim.x) {
        const int32_t m = mr[x], id = nid[x];
        if (m >= 0 && id == B_LARGE) atomicMax(&f.S[guf_find(f.P, m)], (int32_t)f.thr);
        // a neighbour's small piece that leans on one of this strip's large pieces is large — and so is everything
        // of this strip that hangs on it
        if (id >= 0 && m == B_LARGE) atomicMax(&f.S[guf_find(f.P, fbase + id)], (int32_t)f.thr);
    }
}

// Strip-local raster rank (1..n, 0 background) of the component under every pixel of the strip's first / last row.
__global__ void k_strip_label_export(Geo g, const int32_t* __restrict__ Brow3, const uint32_t* __restrict__ R,
                                     const uint32_t* __restrict__ base, int32_t* __restrict__ out /* [2][W] */) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * g.W; i += gridDim.x * blockDim.x) {
        const int side = i / g.W, x = i % g.W;
        const int32_t b = strip_row(g, Brow3, side)[x];
        int32_t lab = 0;
        if (b >= 0) {
            const int yy = b / g.W, xx = b % g.W;
            const size_t wi = (size_t)yy * g.WSR + (xx >> 5);
            lab = (int32_t)(base[wi] + (uint32_t)__popc(R[wi] & ((1u << (xx & 31)) - 1u)) + 1u);
        }
        out[i] = lab;
    }
}

// Global ids: nodes are all strip-local components, node = offsets[strip] + rank - 1 (raster order is preserved:
// strips are ordered, ranks are raster-ordered inside a strip).
__global__ void k_iota(int32_t* P, int32_t n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) P[i] = i;
}
template <int CONN>
__global__ void k_strip_label_unite(const int32_t* __restrict__ rows /* [nranks][2][W] */, const int32_t* __restrict__ offsets, int nranks, int W,
                                    int32_t* __restrict__ P) {
    const int d = CONN == 8 ? 1 : 0;
    const int total = (nranks - 1) * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int j = i / W, x = i % W;                      // boundary between strip j and j + 1
        const int32_t lo = rows[((size_t)(j + 1) * 2 + 0) * W + x];
        if (!lo) continue;
        const int32_t* up = rows + ((size_t)j * 2 + 1) * W;
        for (int dd = -d; dd <= d; ++dd) {
            const int q = x + dd;
            if (q < 0 || q >= W) continue;
            const int32_t u = up[q];
            if (u) guf_unite(P, offsets[j + 1] + lo - 1, offsets[j] + u - 1);
        }
    }
}
__global__ void k_root_bitmap(const int32_t* __restrict__ P, int32_t n, uint32_t* __restrict__ Rg, int nw) {
    for (int wI = blockIdx.x * blockDim.x + threadIdx.x; wI < nw; wI += gridDim.x * blockDim.x) {
        uint32_t bits = 0u;
        for (int b = 0; b < 32; ++b) { const int i = wI * 32 + b; if (i < n && P[i] == i) bits |= 1u << b; }
        Rg[wI] = bits;
    }
}
__global__ void k_strip_lut(const int32_t* __restrict__ P, const uint32_t* __restrict__ Rg, const uint32_t* __restrict__ baseg,
                            int32_t first, int32_t count, int32_t* __restrict__ lut) {
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l <= count; l += gridDim.x * blockDim.x) {
        int32_t v = 0;
        if (l > 0) {
            const int32_t r = guf_find(P, first + l - 1);
            v = (int32_t)(baseg[r >> 5] + (uint32_t)__popc(Rg[r >> 5] & ((1u << (r & 31)) - 1u)) + 1u);
        }
        lut[l] = v;
    }
}

size_t strip_workspace_bytes(int H, int W) {
    if (check_geo(1, H, W) != CS_OK) return 0;
    Geo g = make_geo(1, H, W);
    return carve(nullptr, 0, g, true).bytes;
}
uint32_t* strip_packed_slot(void* ws, size_t ws_bytes, int H, int W) {
    if (check_geo(1, H, W) != CS_OK) return nullptr;
    Geo g = make_geo(1, H, W);
    PostWs p = carve(ws, ws_bytes, g, true);
    return (ws && p.bytes <= ws_bytes) ? p.M0 : nullptr;
}

static int strip_check(int hole_area, int min_size, int conn) {
    if (conn != 4 && conn != 8) { set_error("connectivity must be 4 or 8"); return CS_ERR_ARG; }
    if (hole_area < 0 || min_size < 0) { set_error("negative size threshold"); return CS_ERR_ARG; }
    return CS_OK;
}

// mask_or_null: u8 strip mask to pack; NULL = the packed mask is already in the workspace slot.
int strip_begin(const uint8_t* mask_or_null, int H, int W, int hole_area, int min_size, int conn, int32_t* labels, void* ws, size_t ws_bytes, cudaStream_t st) {
    CS_CHECK_ARG(labels, "null labels pointer");
    int rc = strip_check(hole_area, min_size, conn);
    if (rc != CS_OK) return rc;
    Geo g; PostWs p;
    if ((rc = prep(1, H, W, ws, ws_bytes, g, p, true)) != CS_OK) return rc;
    if (mask_or_null && (rc = pack_mask(mask_or_null, g, p.M0, st)) != CS_OK) return rc;
    ChainArgs a = build_args(g, p, p.M0, hole_area, min_size, labels, true);
    CS_CUDA(cudaMemsetAsync(p.counters, 0, 64 * sizeof(uint32_t), st));
    if ((rc = certify_large(p.M0, g, hole_area, min_size, p.Lbg, p.Lfg, st)) != CS_OK) return rc;
    return launch_tile(0, a, conn, st);
}

// stage 1 = holes, 2 = objects.  rows_out: device int32[2][2][W] (side, {id, size}, x).
int strip_filter_export(int stage, int H, int W, int hole_area, int min_size, int conn, int32_t* labels, int32_t* rows_out,
                        void* ws, size_t ws_bytes, cudaStream_t st) {
    CS_CHECK_ARG(stage == 1 || stage == 2, "stage must be 1 or 2");
    CS_CHECK_ARG(rows_out, "null rows_out");
    Geo g; PostWs p;
    int rc = prep(1, H, W, ws, ws_bytes, g, p, true);
    if (rc != CS_OK) return rc;
    ChainArgs a = build_args(g, p, p.M0, hole_area, min_size, labels, true);
    const FilterStage& f = stage == 1 ? a.f1 : a.f2;
    const StageWs& sw = stage == 1 ? p.s1 : p.s2;
    if (!f.thr) { CS_CUDA(cudaMemsetAsync(rows_out, 0xff, (size_t)4 * W * sizeof(int32_t), st)); return CS_OK; }
    if ((rc = run_filter_merge(g, f, a.stoplist[stage - 1], a.stop_n + (stage - 1), st)) != CS_OK) return rc;
    const size_t nnodes = (size_t)g.nty * g.ntx * NODES_PER_TILE;
    CS_LAUNCH(k_strip_filter_export, 2, 1024, 0, st, g, f, sw.canon, sw.myroot, rows_out, nnodes);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// nb_top: the last-row export [2][W] of the strip above (its side 1), nb_bottom: the first-row export of the strip
// below (its side 0); NULL at the image border.
int strip_filter_resume(int stage, int H, int W, int hole_area, int min_size, int conn, int32_t* labels, const int32_t* nb_top,
                        const int32_t* nb_bottom, void* ws, size_t ws_bytes, cudaStream_t st) {
    CS_CHECK_ARG(stage == 1 || stage == 2, "stage must be 1 or 2");
    Geo g; PostWs p;
    int rc = prep(1, H, W, ws, ws_bytes, g, p, true);
    if (rc != CS_OK) return rc;
    ChainArgs a = build_args(g, p, p.M0, hole_area, min_size, labels, true);
    const FilterStage& f = stage == 1 ? a.f1 : a.f2;
    const StageWs& sw = stage == 1 ? p.s1 : p.s2;
    if (!f.thr) return CS_OK;
    CS_CHECK_ARG(!nb_bottom || H % TH == 0, "a strip with a neighbour below must be a multiple of 64 rows");
    CS_CHECK_ARG(!(nb_top && nb_bottom) || (uint32_t)H >= f.thr, "a strip between two neighbours must have at least `threshold` rows");
    const size_t nnodes = (size_t)g.nty * g.ntx * NODES_PER_TILE;
    if (nb_top) { CS_LAUNCH(k_strip_filter_import, 1, 1024, 0, st, g, f, sw.myroot, nb_top, 0, nnodes); CS_LAUNCH_CHECK(); }
    if (nb_bottom) { CS_LAUNCH(k_strip_filter_import, 1, 1024, 0, st, g, f, sw.myroot, nb_bottom, 1, nnodes); CS_LAUNCH_CHECK(); }
    return launch_tile(stage, a, conn, st);
}

// Cross-tile label merge + rank scan inside the strip; rows_out: device int32[2][W] strip-local ranks on the first /
// last row; count_out: device int32, components in the strip.
int strip_label_export(int H, int W, int conn, int32_t* labels, int32_t* rows_out, int32_t* count_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    CS_CHECK_ARG(labels && rows_out && count_out, "null pointer");
    Geo g; PostWs p;
    int rc = prep(1, H, W, ws, ws_bytes, g, p, true);
    if (rc != CS_OK) return rc;
    if ((rc = run_label_merge_scan(g, p, conn, labels, count_out, st)) != CS_OK) return rc;
    CS_LAUNCH(k_strip_label_export, flat_grid((size_t)2 * W, 256), 256, 0, st, g, p.Brow3, p.R, p.base, rows_out);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

size_t strip_resolve_workspace_bytes(long long total) {
    if (total < 0 || total > 0x7fffffffLL) return 0;
    Arena a(nullptr, 0);
    const size_t nw = ((size_t)total + 31) / 32 + 1;
    a.take<int32_t>((size_t)total + 1); a.take<uint32_t>(nw); a.take<uint32_t>(nw + 1); a.take<unsigned long long>(scan_blocks(nw));
    return a.off;
}

// rows_all: device int32[nranks][2][W] (every strip's label_export rows), offsets: device int32[nranks+1] exclusive
// prefix of the strips' component counts.  lut_out: device int32[count_me + 1], lut[strip-local rank] = global id
// (ids 1..n in raster order of the whole image); total_out: device int32, components in the whole image.
int strip_label_resolve(const int32_t* rows_all, const int32_t* offsets, int nranks, int W, int conn, int my_rank, int32_t first, int32_t count,
                        long long total, int32_t* lut_out, int32_t* total_out, void* ws2, size_t ws2_bytes, cudaStream_t st) {
    CS_CHECK_ARG(rows_all && offsets && lut_out && nranks >= 1 && W > 0 && my_rank >= 0 && my_rank < nranks, "bad arguments");
    CS_CHECK_ARG(conn == 4 || conn == 8, "connectivity must be 4 or 8");
    CS_CHECK_ARG(total >= 0 && total <= 0x7fffffffLL && first >= 0 && count >= 0 && (long long)first + count <= total, "bad component counts");
    Arena a(ws2, ws2_bytes);
    const size_t nw = ((size_t)total + 31) / 32 + 1;
    int32_t* P = a.take<int32_t>((size_t)total + 1);
    uint32_t* Rg = a.take<uint32_t>(nw);
    uint32_t* baseg = a.take<uint32_t>(nw + 1);
    unsigned long long* scanstate = a.take<unsigned long long>(scan_blocks(nw));
    if (!ws2 || !a.ok()) { set_error("resolve workspace too small: need %zu bytes, got %zu", a.off, ws2_bytes); return CS_ERR_WORKSPACE; }
    CS_LAUNCH(k_iota, flat_grid((size_t)total + 1, 256), 256, 0, st, P, (int32_t)total);
    CS_LAUNCH_CHECK();
    if (nranks > 1) {
        if (conn == 8) CS_LAUNCH(k_strip_label_unite<8>, flat_grid((size_t)(nranks - 1) * W, 256), 256, 0, st, rows_all, offsets, nranks, W, P);
        else           CS_LAUNCH(k_strip_label_unite<4>, flat_grid((size_t)(nranks - 1) * W, 256), 256, 0, st, rows_all, offsets, nranks, W, P);
        CS_LAUNCH_CHECK();
    }
    CS_LAUNCH(k_root_bitmap, flat_grid(nw, 256), 256, 0, st, P, (int32_t)total, Rg, (int)nw);
    CS_LAUNCH_CHECK();
    int rc = scan_bitmap(Rg, nw, baseg, scanstate, st);
    if (rc != CS_OK) return rc;
    CS_LAUNCH(k_strip_lut, flat_grid((size_t)count + 1, 256), 256, 0, st, P, Rg, baseg, first, count, lut_out);
    CS_LAUNCH_CHECK();
    if (total_out) CS_CUDA(cudaMemcpyAsync(total_out, baseg + nw, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    return CS_OK;
}

// Store the strip's int32 labels; lut (device, from strip_label_resolve) maps strip-local ranks to global ids.
int strip_final(int H, int W, const int32_t* lut, int32_t* labels, void* ws, size_t ws_bytes, cudaStream_t st) {
    CS_CHECK_ARG(labels, "null labels pointer");
    Geo g; PostWs p;
    int rc = prep(1, H, W, ws, ws_bytes, g, p, true);
    if (rc != CS_OK) return rc;
    return run_final(g, p, labels, lut, nullptr, st);
}

}  // namespace cs

#ifdef CS_PHASE_TIMING
// debug build only (tools/phase_timing.py); not part of the C ABI
extern "C" __attribute__((visibility("default"))) int cs_debug_phase_cycles(unsigned long long* out64, int reset) {
    if (cudaMemcpyFromSymbol(out64, cs::g_phase_cycles, sizeof(unsigned long long) * 128) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[128] = {0}; if (cudaMemcpyToSymbol(cs::g_phase_cycles, z, sizeof(z)) != cudaSuccess) return -1; }
    return 0;
}
#endif
