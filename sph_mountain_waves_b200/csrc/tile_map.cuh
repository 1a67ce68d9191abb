// Neighbourhood tiles of the pair passes: which runs of the cell-sorted particle arrays a block
// of TM_BLOCK consecutive particles can reach through _apply_binary! (src/core.jl:94-112), so
// that a pair kernel stages exactly those runs in shared memory with bulk (TMA) copies and
// gathers its neighbours from there instead of from L1/L2.
//
// Geometry.  In the x-chunked physical cell order (Grid, sphmw_internal.h) the cells of one
// "chunk row" — fixed chunk ci and fixed row rest = j + Ly*k — are contiguous in memory, and so
// are their particles.  A block's particles sit in the cells [key[first], key[last]] of the
// physical order: one PIECE per chunk row they touch (normally one, two when the block straddles
// the end of a chunk row).  A home cell (i, rest) reaches the cells i-1..i+1 of the NR = 3 (2D)
// or 9 (3D) rows rest + nb_drest[r]; for a piece with columns [i0, i1] that is, per row r,
//   MAIN   columns max(i0-1, chunk_lo) .. min(i1+1, chunk_hi) of the same chunk   (contiguous)
//   LEFT   the single column i0-1 when it lies in the previous chunk — or, for i0 = 0, the cell
//          the reference's unchecked linear key arithmetic wraps to (core.jl:98: column Lx-1 of
//          row rest'-1)
//   RIGHT  the single column i1+1 in the next chunk, or the wrapped cell (0, rest'+1)
// — exactly the cells neighbour_pkey() (sphmw_internal.h) yields for the block's particles.
// Segment id: sid = (piece * 3 + part) * NR + r.  Every non-empty segment gets a run of tile
// SLOTS; a neighbour at global position q in segment sid has slot q + delta[sid].  Starts and
// lengths are rounded to TM_ALIGN particles so that every field (8-byte doubles, 4-byte mirror
// words) can be fetched with 16-byte-aligned bulk copies; the padding slots are never addressed.
//
// The map is rebuilt by every cell-list build (k_tile_map, one warp per block) and read by all
// tiled passes of that cell-list generation: they agree on the slots, so the pair list of the
// recording pass stores 16-bit slots instead of 32-bit positions.
#pragma once
#include "sphmw_internal.h"

#define TM_BLOCK 128       // particles (threads) per tile; == NL_BLOCK
#define TM_ALIGN 4         // segment starts/lengths are multiples of this many particles
#define TM_NSID 108        // segment ids per block: pieces * 3 parts * NR rows
#define TM_HDR 4           // header words
#define TM_WORDS (TM_HDR + 3 * TM_NSID)  // u32 words of one block's record
// record layout (u32 words):
//   [0] crow_lo   chunk row (pkey >> cx_shift) of the block's first particle
//   [1] npieces   chunk rows the block touches; > max pieces => the block is not tiled
//   [2] total     slots of the tile (sum of the aligned segment lengths)
//   [3] nstage    non-empty segments, listed compactly in stage_gs/stage_cb
//   [4 .. 4+NSID)           delta[sid]     (int32) slot = position + delta; TM_NO_SEG if empty
//   [4+NSID .. 4+2NSID)     stage_gs[k]    aligned global start of the k-th non-empty segment
//   [4+2NSID .. 4+3NSID)    stage_cb[k]    (aligned length << 16) | first slot
#define TM_NO_SEG 0x7FFFFFFF
#define TM_MAX_SLOTS 65536  // slots are stored in 16 bits

NL_HD int tm_rows(const Grid &g) { return g.ndiff / 3; }               // NR
NL_HD int tm_max_pieces(const Grid &g) { return TM_NSID / g.ndiff; }   // 4 in 3D, 12 in 2D

struct TilePiece {
    int ci, rest;             // chunk and row of the piece's home cells
    int i0, i1;               // their column range
    int chunk_lo, chunk_hi;   // columns of that chunk inside the grid
};

// piece P of a block whose first/last particles sit in cells (key_first, col_first) / (key_last, col_last)
NL_HD TilePiece tm_piece(const Grid &g, uint32_t key_first, uint32_t col_first, uint32_t key_last,
                         uint32_t col_last, int P) {
    const uint32_t crow_lo = key_first >> g.cx_shift, crow_hi = key_last >> g.cx_shift;
    const uint32_t crow = crow_lo + (uint32_t)P;
    TilePiece t;
    t.ci = (int)(crow / (uint32_t)g.rows);
    t.rest = (int)(crow - (uint32_t)t.ci * (uint32_t)g.rows);
    t.chunk_lo = t.ci << g.cx_shift;
    const int hi = t.chunk_lo + (1 << g.cx_shift) - 1, lx1 = (int)g.lim[0] - 1;
    t.chunk_hi = hi < lx1 ? hi : lx1;
    t.i0 = crow == crow_lo ? (int)col_first : t.chunk_lo;
    t.i1 = crow == crow_hi ? (int)col_last : t.chunk_hi;
    return t;
}

// the particle run [gs, ge) of segment (piece, part, r); false if the segment is empty
// (part: 0 MAIN, 1 LEFT, 2 RIGHT)
NL_HD bool tm_segment(const Grid &g, const TilePiece &t, int part, int r, const uint32_t *cell_start,
                      uint32_t &gs, uint32_t &ge) {
    const int rows = (int)g.rows, lx = (int)g.lim[0];
    int rest = t.rest + g.nb_drest[r];  // nb_drest[r], r < NR: the di = -1 group lists every row once
    int a, b;
    if (part == 0) {
        a = t.i0 - 1 > t.chunk_lo ? t.i0 - 1 : t.chunk_lo;
        b = t.i1 + 1 < t.chunk_hi ? t.i1 + 1 : t.chunk_hi;
    } else if (part == 1) {
        if (t.i0 - 1 >= t.chunk_lo) return false;
        a = t.i0 - 1;
        if (a < 0) {  // core.jl:98 has no per-axis check: the linear key wraps into the row before
            a += lx;
            rest -= 1;
        }
        b = a;
    } else {
        if (t.i1 + 1 <= t.chunk_hi) return false;
        a = t.i1 + 1;
        if (a >= lx) {
            a -= lx;
            rest += 1;
        }
        b = a;
    }
    if (rest < 0 || rest >= rows || a < 0 || b >= lx || a > b) return false;
    gs = cell_start[pkey_of(g, a, rest)];
    ge = cell_start[pkey_of(g, b, rest) + 1];
    return ge > gs;
}

// which part the neighbour column i + di of a home cell in this chunk belongs to
NL_HD int tm_part(int i_nb, int chunk_lo, int chunk_hi) { return i_nb < chunk_lo ? 1 : (i_nb > chunk_hi ? 2 : 0); }
NL_HD int tm_sid(int NR, int piece, int part, int r) { return (piece * 3 + part) * NR + r; }

// One block's record, computed serially (the emulation harness and the host-side property tests
// call this; k_tile_map below computes the same record with one warp).
NL_HD void tm_build_record(const Grid &g, const uint32_t *key, const uint32_t *cellx,
                           const uint32_t *cell_start, int64_t n, int64_t block, uint32_t *rec) {
    const int64_t first = block * TM_BLOCK;
    const int64_t last = (first + TM_BLOCK < n ? first + TM_BLOCK : n) - 1;
    const int NR = tm_rows(g);
    const uint32_t kf = key[first], kl = key[last], cf = cellx[first], cl = cellx[last];
    const uint32_t npieces = (kl >> g.cx_shift) - (kf >> g.cx_shift) + 1u;
    rec[0] = kf >> g.cx_shift;
    rec[1] = npieces;
    uint32_t total = 0, nstage = 0;
    for (int s = 0; s < TM_NSID; ++s) rec[TM_HDR + s] = (uint32_t)TM_NO_SEG;
    if (npieces <= (uint32_t)tm_max_pieces(g)) {
        for (int P = 0; P < (int)npieces; ++P) {
            const TilePiece t = tm_piece(g, kf, cf, kl, cl, P);
            for (int part = 0; part < 3; ++part)
                for (int r = 0; r < NR; ++r) {
                    uint32_t gs, ge;
                    if (!tm_segment(g, t, part, r, cell_start, gs, ge)) continue;
                    const uint32_t gs_al = gs & ~(uint32_t)(TM_ALIGN - 1);
                    const uint32_t len = (ge - gs_al + TM_ALIGN - 1) & ~(uint32_t)(TM_ALIGN - 1);
                    rec[TM_HDR + tm_sid(NR, P, part, r)] = (uint32_t)((int32_t)total - (int32_t)gs_al);
                    if (total + len <= TM_MAX_SLOTS && len < 65536u) {
                        rec[TM_HDR + TM_NSID + nstage] = gs_al;
                        rec[TM_HDR + 2 * TM_NSID + nstage] = (len << 16) | total;
                    }
                    ++nstage;
                    total += len;
                }
        }
    }
    rec[2] = total;
    rec[3] = nstage;
}

#if defined(__CUDACC__)
// one warp per block of TM_BLOCK particles; lanes stride over the segment ids
__global__ void __launch_bounds__(128)
k_tile_map(Grid g, const uint32_t *__restrict__ key, const uint32_t *__restrict__ cellx,
           const uint32_t *__restrict__ cell_start, int64_t n, int64_t nblocks, uint32_t *__restrict__ tab) {
    const int64_t block = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    if (block >= nblocks) return;
    uint32_t *rec = tab + (size_t)block * TM_WORDS;
    const int64_t first = block * TM_BLOCK;
    const int64_t last = (first + TM_BLOCK < n ? first + TM_BLOCK : n) - 1;
    const int NR = tm_rows(g);
    const uint32_t kf = key[first], kl = key[last], cf = cellx[first], cl = cellx[last];
    const uint32_t npieces = (kl >> g.cx_shift) - (kf >> g.cx_shift) + 1u;
    const bool tiled = npieces <= (uint32_t)tm_max_pieces(g);
    const int nsid = tiled ? (int)npieces * 3 * NR : 0;
    uint32_t run_total = 0, run_stage = 0;
    for (int s0 = 0; s0 < TM_NSID; s0 += 32) {
        const int s = s0 + (int)lane;
        uint32_t gs_al = 0, len = 0;
        bool have = false;
        if (s < nsid) {
            const int P = s / (3 * NR), part = (s / NR) % 3, r = s % NR;
            const TilePiece t = tm_piece(g, kf, cf, kl, cl, P);
            uint32_t gs, ge;
            if (tm_segment(g, t, part, r, cell_start, gs, ge)) {
                have = true;
                gs_al = gs & ~(uint32_t)(TM_ALIGN - 1);
                len = (ge - gs_al + TM_ALIGN - 1) & ~(uint32_t)(TM_ALIGN - 1);
            }
        }
        // exclusive prefix of the lengths and of the non-empty flags over the lanes
        uint32_t incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += v;
        }
        const unsigned havem = __ballot_sync(0xffffffffu, have);
        const uint32_t base = run_total + incl - len;
        const uint32_t k = run_stage + (uint32_t)__popc(havem & ((1u << lane) - 1u));
        if (s < TM_NSID) rec[TM_HDR + s] = have ? (uint32_t)((int32_t)base - (int32_t)gs_al) : (uint32_t)TM_NO_SEG;
        if (have && base + len <= TM_MAX_SLOTS && len < 65536u) {
            rec[TM_HDR + TM_NSID + k] = gs_al;
            rec[TM_HDR + 2 * TM_NSID + k] = (len << 16) | base;
        }
        run_total += __shfl_sync(0xffffffffu, incl, 31);
        run_stage += (uint32_t)__popc(havem);
    }
    if (lane == 0) {
        rec[0] = kf >> g.cx_shift;
        rec[1] = npieces;
        rec[2] = run_total;
        rec[3] = run_stage;
    }
}
#endif
