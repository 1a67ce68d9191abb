// Tiled pair kernels: _apply_binary! (src/core.jl:94-112) with the neighbourhood of a block staged
// in shared memory by bulk (TMA) copies.
//
// The thread-per-particle list kernels (pair_list.cuh) are bound by the L1 data pipe: every
// accepted pair gathers 4 (density) or 11 (force) neighbour fields, and one gather instruction of a
// warp touches ~11.5 different 128-byte lines (profiles/r01b_pair_list.md).  Here a block of
// TM_BLOCK consecutive particles first copies the particle runs its cells can reach (tile_map.cuh:
// 9 row segments in 3D, 3 in 2D, each contiguous in the cell-sorted arrays) into shared memory —
// `cp.async.bulk` global->shared, completion on an mbarrier, no LSU instructions and no registers —
// and then runs the same per-thread loops with the neighbour fields read from the tile:
//
//   k_tile_build  recording pass (the fused density pass): walks the 9/27 neighbour cells in
//                 key_diff order with the integer pre-test on the staged 10-bit mirror, queues the
//                 survivors per thread, then runs the exact test `r > sys.h` (core.jl:104-105)
//                 and the closure over the queue and writes the ACCEPTED neighbours, as 16-bit
//                 tile slots, to the pair list.
//   k_tile_list   replaying pass (the fused force pass): closure body per list entry; positions
//                 have not changed since the recording pass, so the accepted set is the list
//                 itself and no test is repeated.
//
// STATUS: opt-in (SPHMW_FLAG_TILES).  Correct — CPU emulation and GPU tests — but 2.5x slower than the
// packed-record kernels of pair_list.cuh on B200: a 114 KB tile per 4 warps leaves 8 warps per SM, and a
// thread's closure is one dependent chain that two warps per scheduler cannot hide
// (profiles/r02_pair_kernels.md §3).
//
// Visiting order is that of k_binary (pair_ops.cu) — key_diff order, cell entries front to back —
// so every FP64 sum keeps its bits.  Blocks whose tile does not fit (crowded cells) or that
// straddle too many chunk rows (narrow or sparse grids), and particles whose candidates overflow
// the queue, fall back to the cell walk on global memory inside the same kernels
// (cnt == NL_NONE).
#pragma once
#include "pair_list.cuh"
#include "tile_map.cuh"

// slots per tile (compile-time: field arrays sit at immediate offsets from a slot's address).
// 3D: 9 staged doubles + one global index per slot = 76 B -> 114.3 KB + 1.3 KB of tables: two
// blocks per SM.  A block of 128 particles reaches ~1310 particles on the 1.8 dr cubic lattice.
template <int DIM>
struct TileGeom {
    static constexpr int CAP = DIM == 3 ? 1504 : 640;
};
#define TS_MBAR 0
#define TS_TAB 16
#define TS_DATA (TS_TAB + 4 * TM_WORDS)  // 1328: multiple of 16

// ---- shared-window primitives ---------------------------------------------------------------
#ifdef SPHMW_EMU
// host build for the CPU tests: the "shared window" is one array, addresses are offsets into it;
// a kernel is run once per sync point for all threads of a block (tests/emu/), returning at the
// sync point it has reached — everything before it is idempotent
extern unsigned char emu_tile_smem[];
extern int emu_phase, emu_sync_seen;
extern bool emu_block_or;
inline unsigned char *ts_window() { return emu_tile_smem; }
inline uint32_t ts_addr(const void *p) { return (uint32_t)((const unsigned char *)p - emu_tile_smem); }
inline double ts_ld_f64(uint32_t a) { double v; memcpy(&v, emu_tile_smem + a, 8); return v; }
inline uint32_t ts_ld_u32(uint32_t a) { uint32_t v; memcpy(&v, emu_tile_smem + a, 4); return v; }
inline void ts_st_u32(uint32_t a, uint32_t v) { memcpy(emu_tile_smem + a, &v, 4); }
// the per-thread queue of the recording pass holds 16-bit slots
inline void ts_push(unsigned &top, uint32_t v, bool pass) {
    if (pass) {
        const uint16_t h = (uint16_t)v;
        memcpy(emu_tile_smem + top, &h, 2);
        top += TM_BLOCK * 2u;
    }
}
inline uint32_t ts_ld_u16(uint32_t a) { uint16_t v; memcpy(&v, emu_tile_smem + a, 2); return v; }
inline double2 ts_ldg_f64x2(const double *p) { return double2{p[0], p[1]}; }
inline void ts_st_f64x2(uint32_t a, double2 v) { memcpy(emu_tile_smem + a, &v, 16); }
struct emu_uint4 { uint32_t x, y, z, w; };
inline emu_uint4 ts_ldg_u32x4(const uint32_t *p) { return emu_uint4{p[0], p[1], p[2], p[3]}; }
inline void ts_st_u32x4(uint32_t a, emu_uint4 v) { memcpy(emu_tile_smem + a, &v, 16); }
inline void ts_mbar_init(uint32_t, uint32_t) {}
inline void ts_mbar_expect_tx(uint32_t, uint32_t) {}
inline void ts_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t) { memcpy(emu_tile_smem + dst, src, bytes); }
inline void ts_mbar_wait(uint32_t, uint32_t) {}
#define TS_SYNC()                                  \
    do {                                           \
        if (emu_sync_seen++ == emu_phase) return;  \
    } while (0)
#define TS_BLOCK_ANY(out, pred)                    \
    do {                                           \
        if (emu_phase == 0) emu_block_or = emu_block_or || (pred); \
        if (emu_sync_seen++ == emu_phase) return;  \
        out = emu_block_or;                        \
    } while (0)
#else
__device__ __forceinline__ uint32_t ts_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double ts_ld_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t ts_ld_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void ts_st_u32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// the per-thread queue of the recording pass holds 16-bit slots: one predicated store and one
// predicated add per candidate
__device__ __forceinline__ void ts_push(unsigned &top, uint32_t v, bool pass) {
    asm volatile(
        "{\n\t"
        ".reg .pred ps;\n\t"
        ".reg .b16 hv;\n\t"
        "setp.ne.u32 ps, %2, 0;\n\t"
        "cvt.u16.u32 hv, %1;\n\t"
        "@ps st.shared.u16 [%0], hv;\n\t"
        "@ps add.u32 %0, %0, %3;\n\t"
        "}"
        : "+r"(top)
        : "r"(v), "r"((unsigned)pass), "n"(TM_BLOCK * 2));
}
__device__ __forceinline__ uint32_t ts_ld_u16(uint32_t a) {
    uint32_t v;
    asm volatile("{\n\t.reg .b16 hv;\n\tld.shared.u16 hv, [%1];\n\tcvt.u32.u16 %0, hv;\n\t}" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 ts_ldg_f64x2(const double *p) { return __ldg((const double2 *)p); }
__device__ __forceinline__ void ts_st_f64x2(uint32_t a, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ uint4 ts_ldg_u32x4(const uint32_t *p) { return __ldg((const uint4 *)p); }
__device__ __forceinline__ void ts_st_u32x4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// mbarrier + 1-D bulk copy (TMA engine; SASS: SYNCS.*, UBLKCP)
__device__ __forceinline__ void ts_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void ts_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ts_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void ts_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred pw;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 pw, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, pw;\n\t"
            "}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
#define TS_SYNC() __syncthreads()
#define TS_BLOCK_ANY(out, pred) out = __syncthreads_or(pred)
#endif

// neighbour accessor of the tiled kernels: staged fields from the tile (array k of the operator's
// tile_index map sits k * CAP * 8 bytes behind array 0), anything else from global memory through
// the slot's global position
template <int DIM, class Op>
struct TileQ {
    uint32_t a;   // shared-window address of the slot in array 0
    uint32_t ga;  // shared-window address of the slot's global position (if the tile carries them)
    const Fields &f;
    template <int SLOT>
    __device__ double get() const {
        constexpr int k = Op::template tile_index<DIM>(SLOT);
        if constexpr (k >= 0) return ts_ld_f64(a + (uint32_t)k * (TileGeom<DIM>::CAP * 8u));
        else return f.s[SLOT][ts_ld_u32(ga)];
    }
    __device__ double rho_floored(const Params &c) const { return jl_max(get<S_RHO>(), c.rho_floor); }
};

// does the operator read a neighbour field that is not staged?  (then the tile carries the
// global position of every slot)
template <int DIM, class Op>
__host__ __device__ constexpr bool tile_needs_gidx() {
    return Op::REC_KIND == 2 && (Op::template tile_index<DIM>(S_RHO) < 0 || Op::template tile_index<DIM>(S_CS) < 0);
}
template <int DIM, class Op>
__host__ __device__ constexpr int tile_bytes_list() {
    return TS_DATA + TileGeom<DIM>::CAP * (8 * Op::template tile_fields<DIM>() + 4);
}
template <int DIM, class Op>
__host__ __device__ constexpr int tile_bytes_build(int stride) {
    return TS_DATA + TileGeom<DIM>::CAP * (4 + 8 * Op::template tile_fields<DIM>() + 4) + stride * TM_BLOCK * 2;
}

// ---- staging --------------------------------------------------------------------------------
// The block's record goes to shared memory, the global position of every slot is written out
// (gidx: the copy loops and the non-staged fields address global memory through it), then the
// field arrays follow — either
//   TILE_STAGE_TMA 1  one bulk copy (cp.async.bulk, TMA engine; SASS UBLKCP.S.G) per (segment, array),
//                     completion on an mbarrier (SYNCS.*): no LSU instructions, no registers
//   TILE_STAGE_TMA 0  coalesced 16-byte loads and stores by all threads (default)
// Both were measured on B200 and run the passes at the same speed (profiles/r02_pair_kernels.md):
// the tiled kernels are not limited by the staging but by their 8 resident warps per SM.
#ifndef TILE_STAGE_TMA
#define TILE_STAGE_TMA 0
#endif

template <int DIM, class Op>
__device__ __forceinline__ void tile_fill_gidx(unsigned char *sm, const uint32_t *stab, int off_gidx) {
    const uint32_t nstage = stab[3];
    const uint32_t ga = ts_addr(sm + off_gidx);
    for (uint32_t k = 0; k < nstage; ++k) {
        const uint32_t gs = stab[TM_HDR + TM_NSID + k], cb = stab[TM_HDR + 2 * TM_NSID + k];
        const uint32_t len = cb >> 16, base = cb & 0xFFFFu;
        for (uint32_t j = threadIdx.x; j < len; j += TM_BLOCK) ts_st_u32(ga + (base + j) * 4u, gs + j);
    }
}

#if TILE_STAGE_TMA
// off_xq < 0: no mirror.  Every thread of the block calls this.
template <int DIM, class Op>
__device__ __forceinline__ void tile_copy(unsigned char *sm, const uint32_t *stab, const Fields &f,
                                          const uint32_t *__restrict__ xq, int off_xq, int off_d, int) {
    constexpr int CAP = TileGeom<DIM>::CAP;
    constexpr int NF = Op::template tile_fields<DIM>();
    const uint32_t nstage = stab[3];
    const uint32_t bar = ts_addr(sm + TS_MBAR);
    const int narr = NF + (off_xq >= 0 ? 1 : 0);
    // global array of each staged field, in tile order
    const double *src[NF];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const int k = Op::template tile_index<DIM>(s);
        if (k >= 0) src[k] = f.s[s];
    }
    for (uint32_t w = threadIdx.x; w < nstage * (uint32_t)narr; w += TM_BLOCK) {
        const uint32_t k = w / (uint32_t)narr, a = w - k * (uint32_t)narr;
        const uint32_t gs = stab[TM_HDR + TM_NSID + k], cb = stab[TM_HDR + 2 * TM_NSID + k];
        const uint32_t len = cb >> 16, base = cb & 0xFFFFu;
        if (a < (uint32_t)NF) {
            const double *g = nullptr;
#pragma unroll
            for (int j = 0; j < NF; ++j)
                if (a == (uint32_t)j) g = src[j];
            ts_bulk_g2s(ts_addr(sm + off_d) + (a * CAP + base) * 8u, g + gs, len * 8u, bar);
        } else {
            ts_bulk_g2s(ts_addr(sm + off_xq) + base * 4u, xq + gs, len * 4u, bar);
        }
    }
}
#else
// Slots come in aligned groups of TM_ALIGN = 4 that are contiguous in global memory: a thread
// moves two slots of every array (16 bytes) per step, all its loads in flight before the stores.
template <int DIM, class Op>
__device__ __forceinline__ void tile_copy(unsigned char *sm, const uint32_t *stab, const Fields &f,
                                          const uint32_t *__restrict__ xq, int off_xq, int off_d, int off_gidx) {
    constexpr int CAP = TileGeom<DIM>::CAP;
    constexpr int NF = Op::template tile_fields<DIM>();
    const uint32_t total = stab[2];
    const uint32_t ga = ts_addr(sm + off_gidx), da = ts_addr(sm + off_d);
    const double *src[NF];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const int k = Op::template tile_index<DIM>(s);
        if (k >= 0) src[k] = f.s[s];
    }
    for (uint32_t s = 2u * threadIdx.x; s < total; s += 2u * TM_BLOCK) {
        const uint32_t gp = ts_ld_u32(ga + s * 4u);
        double2 v[NF];
#pragma unroll
        for (int k = 0; k < NF; ++k) v[k] = ts_ldg_f64x2(src[k] + gp);
#pragma unroll
        for (int k = 0; k < NF; ++k) ts_st_f64x2(da + ((uint32_t)k * CAP + s) * 8u, v[k]);
    }
    if (off_xq >= 0) {
        const uint32_t xa = ts_addr(sm + off_xq);
        for (uint32_t s = 4u * threadIdx.x; s < total; s += 4u * TM_BLOCK)
            ts_st_u32x4(xa + s * 4u, ts_ldg_u32x4(xq + ts_ld_u32(ga + s * 4u)));
    }
}
#endif

// common prologue of both kernels; `tiled` tells whether the block has a staged tile.
// (A macro because TS_SYNC returns from the kernel in the emulation build.)
#define TILE_PROLOGUE(OFF_XQ, OFF_D, OFF_GIDX)                                                          \
    unsigned char *sm = ts_window();                                                                    \
    uint32_t *stab = (uint32_t *)(sm + TS_TAB);                                                         \
    const int64_t p = blockIdx.x * (int64_t)TM_BLOCK + threadIdx.x;                                     \
    const bool live = p < n;                                                                            \
    CellCoord home{0, 0};                                                                               \
    bool sel = false;                                                                                   \
    if (live) {                                                                                         \
        home = cell_of(g, key[p], cellx[p]);                                                            \
        sel = !cf.on || col_selected(cf, home.i);                                                       \
    }                                                                                                   \
    bool any;                                                                                           \
    TS_BLOCK_ANY(any, sel);                                                                             \
    if (!any) {                                                                                         \
        if (live) {                                                                                     \
            if (cf.copy) Op::template skip<DIM>(f, out, p);                                             \
            if (RECORDING) pl.cnt[p] = NL_NONE;                                                         \
        }                                                                                               \
        return;                                                                                         \
    }                                                                                                   \
    {                                                                                                   \
        const uint32_t *rec = tile_tab + (size_t)blockIdx.x * TM_WORDS;                                 \
        for (int w = threadIdx.x; w < TM_WORDS; w += TM_BLOCK) stab[w] = rec[w];                        \
        if (TILE_STAGE_TMA && threadIdx.x == 0) {                                                       \
            /* the barrier expects the tile's bytes before any copy is issued */                        \
            ts_mbar_init(ts_addr(sm + TS_MBAR), 1);                                                     \
            if (rec[1] <= (uint32_t)tm_max_pieces(g) && rec[2] <= (uint32_t)TileGeom<DIM>::CAP &&       \
                rec[3] <= (uint32_t)TM_NSID)                                                            \
                ts_mbar_expect_tx(ts_addr(sm + TS_MBAR),                                                \
                                  rec[2] * (uint32_t)(8 * Op::template tile_fields<DIM>() + ((OFF_XQ) >= 0 ? 4 : 0))); \
        }                                                                                               \
    }                                                                                                   \
    TS_SYNC();                                                                                          \
    const bool tiled = stab[1] <= (uint32_t)tm_max_pieces(g) && stab[2] <= (uint32_t)TileGeom<DIM>::CAP && \
                       stab[3] <= (uint32_t)TM_NSID;                                                    \
    if (tiled) tile_fill_gidx<DIM, Op>(sm, stab, OFF_GIDX);                                             \
    TS_SYNC();                                                                                          \
    if (tiled) tile_copy<DIM, Op>(sm, stab, f, pl.xq, OFF_XQ, OFF_D, OFF_GIDX);                         \
    TS_SYNC();                                                                                          \
    if (TILE_STAGE_TMA && tiled) ts_mbar_wait(ts_addr(sm + TS_MBAR), 0);                                \
    if (!live) return;                                                                                  \
    if (!sel) {                                                                                         \
        if (cf.copy) Op::template skip<DIM>(f, out, p);                                                 \
        if (RECORDING) pl.cnt[p] = NL_NONE;                                                             \
        return;                                                                                         \
    }

#ifndef SPHMW_EMU
__device__ __forceinline__ unsigned char *ts_window() {
    extern __shared__ __align__(16) unsigned char tile_smem_dyn[];
    return tile_smem_dyn;
}
#endif

// pair list of the tiled kernels: 16-bit slots, two per word; word kk of particle p is
// list16[((p >> 5) * (stride / 2) + kk) * 32 + (p & 31)]
template <int DIM, class Op>
__global__ void __launch_bounds__(TM_BLOCK)
k_tile_build(Fields f, Fields out, Params prm, Grid g, const uint32_t *__restrict__ key,
             const uint32_t *__restrict__ cellx, const uint32_t *__restrict__ cell_start, int64_t n, int self,
             unsigned long long *pair_counter, ColFilter cf, PairList pl, const uint32_t *__restrict__ tile_tab) {
    constexpr bool RECORDING = true;
    constexpr int CAP = TileGeom<DIM>::CAP;
    constexpr int NF = Op::template tile_fields<DIM>();
    constexpr int OFF_XQ = TS_DATA, OFF_D = TS_DATA + CAP * 4, OFF_G = OFF_D + NF * CAP * 8, OFF_Q = OFF_G + CAP * 4;
    TILE_PROLOGUE(OFF_XQ, OFF_D, OFF_G)
    const double px = f.s[S_X0][p], py = f.s[S_X1][p], pz = DIM == 3 ? f.s[S_X2][p] : 0.0;
    Op op;
    op.template init<DIM>(f, prm, p);
    unsigned accepted = 0;
    bool fits = tiled;
    const uint32_t stride = (uint32_t)pl.stride;
    const unsigned qbase = ts_addr(sm + OFF_Q) + threadIdx.x * 2u;
    const unsigned qend = qbase + stride * (TM_BLOCK * 2u);
    unsigned qtop = qbase;
    const int NR = tm_rows(g);
    uint32_t own_slot = 0;
    if (tiled) {
        // ---- phase 1: candidates -> queue (as k_binary_build, on tile slots) -------------------
        const int P = (int)((key[p] >> g.cx_shift) - stab[0]);
        const int chunk_lo = (home.i >> g.cx_shift) << g.cx_shift;
        const int chunk_hi_raw = chunk_lo + (1 << g.cx_shift) - 1;
        const int chunk_hi = chunk_hi_raw < (int)g.lim[0] - 1 ? chunk_hi_raw : (int)g.lim[0] - 1;
        own_slot = (uint32_t)((int32_t)p + (int32_t)stab[TM_HDR + tm_sid(NR, P, 0, NR / 2)]);
        const uint32_t xs = ts_addr(sm + OFF_XQ), ds = ts_addr(sm + OFF_D);
        const uint32_t ow = ts_ld_u32(xs + own_slot * 4u);
        const int qx = (int)(ow & 1023u), qy = (int)((ow >> 10) & 1023u), qz = (int)(ow >> 20);
        const int ly = (int)g.lim[1];
        const int hj = home.rest % ly, hk = home.rest / ly;
        for (int d = 0; d < g.ndiff && fits; ++d) {
            unsigned nk;
            if (!neighbour_pkey(g, home, d, nk)) continue;
            const uint32_t b = cell_start[nk], e = cell_start[nk + 1];
            if (e == b) continue;
            if (!((e - b) < (1u << 20) && qtop + (e - b) * (TM_BLOCK * 2u) <= qend)) {  // room for the whole run
                fits = false;
                break;
            }
            const int di = g.nb_di[d], dj = g.nb_dj[d], dk = g.nb_dk[d];
            const int sid = tm_sid(NR, P, tm_part(home.i + di, chunk_lo, chunk_hi), d % NR);
            const uint32_t sb = (uint32_t)((int32_t)b + (int32_t)stab[TM_HDR + sid]), se = sb + (e - b);
            const bool regular = (unsigned)(home.i + di) < (unsigned)g.lim[0] && (unsigned)(hj + dj) < (unsigned)ly &&
                                 (unsigned)(hk + dk) < (unsigned)g.lim[2];
            if (regular) {
                int ox = qx - NL_Q10_ONE * di, oy = qy - NL_Q10_ONE * dj, oz = qz - NL_Q10_ONE * dk;
#ifndef SPHMW_EMU
                asm volatile("" : "+r"(ox), "+r"(oy), "+r"(oz));  // keep them out of the inner loop
#endif
                const uint32_t last = se - 1;
                for (uint32_t s = sb; s < se; s += 4) {
                    uint32_t w[4];  // slots past the run (still inside the tile arrays) are masked
#pragma unroll
                    for (int i = 0; i < 4; ++i) w[i] = ts_ld_u32(xs + (s + i) * 4u);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int dx = ox - (int)(w[i] & 1023u);
                        const int dy = oy - (int)((w[i] >> 10) & 1023u);
                        int s2 = dx * dx + dy * dy;
                        if (DIM == 3) {
                            const int dz = oz - (int)(w[i] >> 20);
                            s2 += dz * dz;
                        }
                        const uint32_t si = s + i;
                        ts_push(qtop, si, !((s2 > NL_Q10_R2MAX) || (i > 0 && si > last)));
                    }
                }
            } else {
                // a cell reached through the reference's row wrap (core.jl:98): exact test
                for (uint32_t s = sb; s < se; ++s) {
                    const uint32_t sa = ds + s * 8u;
                    double dx = px - ts_ld_f64(sa);
                    double dy = py - ts_ld_f64(sa + CAP * 8u);
                    double r2 = dx * dx + dy * dy;
                    if (DIM == 3) {
                        double dz = pz - ts_ld_f64(sa + 2u * CAP * 8u);
                        r2 = r2 + dz * dz;
                    }
                    ts_push(qtop, s, !(r2 > g.r2_max));
                }
            }
        }
    }
    if (fits) {
        // ---- phase 2: exact test + closure over the queue; accepted slots -> list -------------
        const uint32_t ds = ts_addr(sm + OFF_D);
        uint32_t *row = pl.list16 + ((size_t)(p >> 5) * (stride >> 1)) * 32 + (size_t)(p & 31);
        uint32_t pend = 0;
        for (unsigned a = qbase; a < qtop; a += TM_BLOCK * 2u) {
            const uint32_t s = ts_ld_u16(a);
            const uint32_t sa = ds + s * 8u;
            // dist(p,q) — core.jl:8-10, algebra.jl:49-60: left-to-right, no FMA
            double dx = px - ts_ld_f64(sa);
            double dy = py - ts_ld_f64(sa + CAP * 8u);
            double dz = 0.0;
            double r2 = dx * dx + dy * dy;
            if (DIM == 3) {
                dz = pz - ts_ld_f64(sa + 2u * CAP * 8u);
                r2 = r2 + dz * dz;
            }
            if ((r2 > g.r2_max) || (s == own_slot)) continue;  // core.jl:105, decided on r2 (Grid::r2_max)
            double r = sqrt(r2);
            op.template pair_q<DIM>(prm, TileQ<DIM, Op>{sa, 0u, f}, dx, dy, dz, r);
            if (accepted & 1u) __stcs(row + (size_t)(accepted >> 1) * 32, pend | (s << 16));
            else pend = s;
            ++accepted;
        }
        if (accepted & 1u) __stcs(row + (size_t)(accepted >> 1) * 32, pend);
        pl.cnt[p] = accepted;
    } else {
        pl.cnt[p] = NL_NONE;
        if (tiled) atomicAdd(pl.overflow, 1ull);
        nl_walk<DIM>(op, f, prm, g, home, p, px, py, pz, cell_start, accepted);
    }
    if (self) op.template pair<DIM>(f, prm, p, p, 0.0, 0.0, 0.0, 0.0);  // core.jl:155-157
    op.template finish<DIM>(f, out, prm, p);
    nl_count_pairs(pair_counter, accepted);
}

template <int DIM, class Op>
__global__ void __launch_bounds__(TM_BLOCK)
k_tile_list(Fields f, Fields out, Params prm, Grid g, const uint32_t *__restrict__ key,
            const uint32_t *__restrict__ cellx, const uint32_t *__restrict__ cell_start, int64_t n, int self,
            unsigned long long *pair_counter, ColFilter cf, PairList pl, const uint32_t *__restrict__ tile_tab) {
    constexpr bool RECORDING = false;
    constexpr int CAP = TileGeom<DIM>::CAP;
    constexpr int NF = Op::template tile_fields<DIM>();
    constexpr int OFF_D = TS_DATA;
    constexpr int OFF_G = OFF_D + NF * CAP * 8;
    TILE_PROLOGUE(-1, OFF_D, OFF_G)
    const uint32_t cnt = pl.cnt[p];
    const double px = f.s[S_X0][p], py = f.s[S_X1][p], pz = DIM == 3 ? f.s[S_X2][p] : 0.0;
    Op op;
    op.template init<DIM>(f, prm, p);
    unsigned accepted = 0;
    if (tiled && cnt != NL_NONE) {
        const uint32_t ds = ts_addr(sm + OFF_D);
        const uint32_t gs = ts_addr(sm + OFF_G);
        const uint32_t *row = pl.list16 + ((size_t)(p >> 5) * ((uint32_t)pl.stride >> 1)) * 32 + (size_t)(p & 31);
        uint32_t wn = cnt ? __ldcs(row) : 0u;
        for (uint32_t k = 0; k < cnt; k += 2) {
            const uint32_t w = wn;
            if (k + 2 < cnt) wn = __ldcs(row + (size_t)((k >> 1) + 1) * 32);  // one word ahead
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (half == 1 && k + 1 >= cnt) break;
                const uint32_t s = half ? (w >> 16) : (w & 0xFFFFu);
                const uint32_t sa = ds + s * 8u;
                double dx = px - ts_ld_f64(sa);
                double dy = py - ts_ld_f64(sa + CAP * 8u);
                double dz = 0.0;
                double r2 = dx * dx + dy * dy;
                if (DIM == 3) {
                    dz = pz - ts_ld_f64(sa + 2u * CAP * 8u);
                    r2 = r2 + dz * dz;
                }
                double r = sqrt(r2);
                op.template pair_q<DIM>(prm, TileQ<DIM, Op>{sa, gs + s * 4u, f}, dx, dy, dz, r);
            }
        }
        accepted = cnt;
    } else {
        nl_walk<DIM>(op, f, prm, g, home, p, px, py, pz, cell_start, accepted);
    }
    if (self) op.template pair<DIM>(f, prm, p, p, 0.0, 0.0, 0.0, 0.0);
    op.template finish<DIM>(f, out, prm, p);
    nl_count_pairs(pair_counter, accepted);
}
