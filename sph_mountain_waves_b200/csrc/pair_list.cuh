// Pair-list kernels: _apply_binary! (src/core.jl:94-112) split into a recording pass and
// replaying passes.  See PairList in sphmw_internal.h for the layout and the invariants.
//
//   k_binary_build  walks the 9/27 neighbour cells in key_diff order (structs.jl:73-81) with a
//                   cheap cut-off test (integers on a 10-bit mirror, or the exact FP64 one), queues
//                   the survivors per thread in shared memory, then — with the lanes of a warp
//                   compacted onto ~26 survivors instead of ~157 candidates — runs the exact test
//                   `r > sys.h` (core.jl:104-105) and the closure body, and streams the queue
//                   to the list.
//   k_binary_list   replays a recorded list: exact test + closure body per entry.
//
// Both visit accepted neighbours in exactly the order of k_binary (pair_ops.cu), so every FP64
// sum is bit-identical to the cell walk.  A particle whose survivors do not fit the list
// stride, or that was outside the recording pass's column filter, has cnt == NL_NONE and
// walks the cells with the original loop.
#pragma once
#include "q_access.cuh"
#include "sphmw_internal.h"

// the loop of k_binary, for particles without a list
template <int DIM, class Op>
__device__ __forceinline__ void nl_walk(Op &op, const Fields &f, const Params &prm, const Grid &g,
                                        const CellCoord &home, int64_t p, double px, double py,
                                        double pz, const uint32_t *__restrict__ cell_start,
                                        unsigned &accepted) {
    for (int d = 0; d < g.ndiff; ++d) {
        unsigned nk;
        if (!neighbour_pkey(g, home, d, nk)) continue;  // core.jl:98
        uint32_t b = cell_start[nk], e = cell_start[nk + 1];
        for (uint32_t q = b; q < e; ++q) {
            double dx = px - f.s[S_X0][q];
            double dy = py - f.s[S_X1][q];
            double dz = 0.0;
            double r2 = dx * dx + dy * dy;
            if (DIM == 3) {
                dz = pz - f.s[S_X2][q];
                r2 = r2 + dz * dz;
            }
            if ((r2 > g.r2_max) || (q == p)) continue;
            double r = sqrt(r2);
            op.template pair<DIM>(f, prm, p, q, dx, dy, dz, r);
            ++accepted;
        }
    }
}

// exact test + closure body for one recorded candidate (the particle itself is among them)
template <int DIM, class Op>
__device__ __forceinline__ void nl_entry(Op &op, const Fields &f, const Params &prm, const Grid &g,
                                         int64_t p, uint32_t q, double px, double py, double pz,
                                         unsigned &accepted) {
    // dist(p,q) — core.jl:8-10, algebra.jl:49-60: left-to-right, no FMA
    double dx = px - f.s[S_X0][q];
    double dy = py - f.s[S_X1][q];
    double dz = 0.0;
    double r2 = dx * dx + dy * dy;
    if (DIM == 3) {
        dz = pz - f.s[S_X2][q];
        r2 = r2 + dz * dz;
    }
    if ((r2 > g.r2_max) || (q == (uint32_t)p)) return;  // core.jl:105, decided on r2 (Grid::r2_max)
    double r = sqrt(r2);
    op.template pair<DIM>(f, prm, p, q, dx, dy, dz, r);
    ++accepted;
}

// ---- packed-record variants (SPHMW_FLAG_PACKED_RECORDS; NbRec in sphmw_internal.h) ----------
// density closure: position and mass of q from record A
template <int DIM, class Op>
__device__ __forceinline__ void nl_entry_rec_density(Op &op, const Params &prm, const Grid &g, int64_t p,
                                                     uint32_t q, double px, double py, double pz,
                                                     const NbRec *__restrict__ recA, unsigned &accepted) {
    const NbRec A = nb_load(recA + q);
    double dx = px - A.a;
    double dy = py - A.b;
    double dz = 0.0;
    double r2 = dx * dx + dy * dy;
    if (DIM == 3) {
        dz = pz - A.c;
        r2 = r2 + dz * dz;
    }
    if ((r2 > g.r2_max) || (q == (uint32_t)p)) return;
    double r = sqrt(r2);
    op.template pair_q<DIM>(prm, RecAQ{A.d}, dx, dy, dz, r);
    ++accepted;
}
// force closure: all three records are requested up front (the exact test rejects few entries)
template <int DIM, class Op>
__device__ __forceinline__ void nl_entry_rec_force(Op &op, const Params &prm, const Grid &g, int64_t p,
                                                   uint32_t q, double px, double py, double pz,
                                                   const PairList &pl, unsigned &accepted) {
    const NbRec A = nb_load(pl.recA + q);
    const NbRec B = nb_load(pl.recB + q);
    const NbRec C = nb_load(pl.recC + q);
    double dx = px - A.a;
    double dy = py - A.b;
    double dz = 0.0;
    double r2 = dx * dx + dy * dy;
    if (DIM == 3) {
        dz = pz - A.c;
        r2 = r2 + dz * dz;
    }
    if ((r2 > g.r2_max) || (q == (uint32_t)p)) return;
    double r = sqrt(r2);
    op.template pair_q<DIM>(prm, RecQ{A.d, B, C}, dx, dy, dz, r);
    ++accepted;
}

// per-thread queue column in shared memory, addressed with 32-bit shared-window addresses so
// that a push is one predicated st.shared and one predicated add (the queue is touched only
// through these volatile statements, which keep their order).  Room for a whole cell run is
// checked before the run starts (nl_room), not per candidate.
__device__ __forceinline__ bool nl_room(unsigned top, unsigned end, uint32_t run) {
    // run < 2^20 keeps the product inside 32 bits; longer runs never fit a stride <= 96 anyway
    return run < (1u << 20) && top + run * (NL_BLOCK * 4u) <= end;
}
#ifdef SPHMW_EMU  // host build for the CPU tests: the "shared window" is the array itself
inline void nl_push(unsigned &top, uint32_t q, bool pass) {
    if (pass) {
        *(uint32_t *)((char *)nl_queue_emu() + top) = q;
        top += NL_BLOCK * 4u;
    }
}
inline uint32_t nl_peek(unsigned addr) { return *(const uint32_t *)((const char *)nl_queue_emu() + addr); }
#else
__device__ __forceinline__ void nl_push(unsigned &top, uint32_t q, bool pass) {
    asm volatile(
        "{\n\t"
        ".reg .pred ps;\n\t"
        "setp.ne.u32 ps, %2, 0;\n\t"
        "@ps st.shared.u32 [%0], %1;\n\t"
        "@ps add.u32 %0, %0, %3;\n\t"
        "}"
        : "+r"(top)
        : "r"(q), "r"((unsigned)pass), "n"(NL_BLOCK * 4));
}
__device__ __forceinline__ uint32_t nl_peek(unsigned addr) {
    uint32_t q;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(q) : "r"(addr));
    return q;
}
#endif

__device__ __forceinline__ void nl_count_pairs(unsigned long long *pair_counter, unsigned accepted) {
    if (pair_counter) {
        unsigned m = __activemask();
        unsigned tot = __reduce_add_sync(m, accepted);
        if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(pair_counter, (unsigned long long)tot);
    }
}

// FILTER: how phase 1 tests a candidate — the exact FP64 test (three 8-byte loads), or integers
// on the 10-bit cell-relative mirror (one 4-byte load; survivors get the exact test in phase 2).
// An FP32 mirror of the absolute positions (float4, 16-byte loads) was measured slower:
// profiles/r01b_pair_list.md.
#define NL_FILTER_F64 0
#define NL_FILTER_Q10 2
// REC (fused density pass only): phase 2 reads q from record A, and the particle's own records B
// and C are written once its density, smoothing length and pressure are final
template <int DIM, class Op, int FILTER, bool REC = false>
__global__ void __launch_bounds__(NL_BLOCK)
k_binary_build(Fields f, Fields out, Params prm, Grid g, const uint32_t *__restrict__ key,
               const uint32_t *__restrict__ cellx, const uint32_t *__restrict__ cell_start, int64_t n,
               int self, unsigned long long *pair_counter, ColFilter cf, PairList pl) {
    extern __shared__ uint32_t nl_queue[];  // [stride][NL_BLOCK]: a private column per thread
    const int64_t p = blockIdx.x * (int64_t)NL_BLOCK + threadIdx.x;
    if (p >= n) return;
    const CellCoord home = cell_of(g, key[p], cellx[p]);
    if (cf.on && !col_selected(cf, home.i)) {
        if (cf.copy) Op::template skip<DIM>(f, out, p);
        pl.cnt[p] = NL_NONE;
        return;
    }
    const uint32_t stride = (uint32_t)pl.stride;
    const unsigned qbase = (unsigned)__cvta_generic_to_shared(nl_queue + threadIdx.x);
    const unsigned qend = qbase + stride * (NL_BLOCK * 4u);
    unsigned qtop = qbase;
    const double px = f.s[S_X0][p], py = f.s[S_X1][p], pz = DIM == 3 ? f.s[S_X2][p] : 0.0;
    bool fits = true;
    // ---- phase 1: candidates -> queue -------------------------------------------------
    if (FILTER == NL_FILTER_Q10) {
        // own position in h/1024 inside the home cell, and the home cell's row coordinates: a
        // neighbour cell reached without any wrap of the linear key arithmetic (core.jl:98 has no
        // per-axis check) lies exactly (di, dj, dk) cells away; the few wrapped ones (cells on
        // the faces of the grid) take the exact FP64 test instead
        const uint32_t ow = pl.xq[p];
        const int qx = (int)(ow & 1023u), qy = (int)((ow >> 10) & 1023u), qz = (int)(ow >> 20);
        const int ly = (int)g.lim[1];
        const int hj = home.rest % ly, hk = home.rest / ly;
        for (int d = 0; d < g.ndiff; ++d) {
            unsigned nk;
            if (!neighbour_pkey(g, home, d, nk)) continue;
            const uint32_t b = cell_start[nk], e = cell_start[nk + 1];
            if (!nl_room(qtop, qend, e - b)) {
                fits = false;
                break;
            }
            const int di = g.nb_di[d], dj = g.nb_dj[d], dk = g.nb_dk[d];
            const bool regular = (unsigned)(home.i + di) < (unsigned)g.lim[0] && (unsigned)(hj + dj) < (unsigned)ly &&
                                 (unsigned)(hk + dk) < (unsigned)g.lim[2];
            if (regular) {
                int ox = qx - NL_Q10_ONE * di, oy = qy - NL_Q10_ONE * dj, oz = qz - NL_Q10_ONE * dk;
                asm volatile("" : "+r"(ox), "+r"(oy), "+r"(oz));  // keep them out of the inner loop
                const uint32_t last = e - 1;
                for (uint32_t q = b; q < e; q += 4) {
                    const uint32_t *__restrict__ cp = pl.xq + q;  // padded: slots past the run are masked
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) w[i] = cp[i];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int dx = ox - (int)(w[i] & 1023u);
                        const int dy = oy - (int)((w[i] >> 10) & 1023u);
                        int s2 = dx * dx + dy * dy;
                        if (DIM == 3) {
                            const int dz = oz - (int)(w[i] >> 20);
                            s2 += dz * dz;
                        }
                        const uint32_t qi = q + i;
                        nl_push(qtop, qi, !((s2 > NL_Q10_R2MAX) || (i > 0 && qi > last)));
                    }
                }
            } else {
                for (uint32_t q = b; q < e; ++q) {
                    double dx = px - f.s[S_X0][q];
                    double dy = py - f.s[S_X1][q];
                    double r2 = dx * dx + dy * dy;
                    if (DIM == 3) {
                        double dz = pz - f.s[S_X2][q];
                        r2 = r2 + dz * dz;
                    }
                    nl_push(qtop, q, !(r2 > g.r2_max));
                }
            }
        }
    } else {
        for (int d = 0; d < g.ndiff; ++d) {
            unsigned nk;
            if (!neighbour_pkey(g, home, d, nk)) continue;
            const uint32_t b = cell_start[nk], e = cell_start[nk + 1];
            if (!nl_room(qtop, qend, e - b)) {
                fits = false;
                break;
            }
            for (uint32_t q = b; q < e; ++q) {
                double dx = px - f.s[S_X0][q];
                double dy = py - f.s[S_X1][q];
                double r2 = dx * dx + dy * dy;
                if (DIM == 3) {
                    double dz = pz - f.s[S_X2][q];
                    r2 = r2 + dz * dz;
                }
                nl_push(qtop, q, !(r2 > g.r2_max));
            }
        }
    }
    // ---- phase 2: exact test + closure over the queue; queue -> list --------------------
    Op op;
    op.template init<DIM>(f, prm, p);
    unsigned accepted = 0;
    if (fits) {
        uint32_t *row = pl.list + ((size_t)(p >> 5) * stride) * 32 + (size_t)(p & 31);
        for (unsigned a = qbase; a < qtop; a += NL_BLOCK * 4u, row += 32) {
            const uint32_t q = nl_peek(a);
            __stcs(row, q);
            if constexpr (REC && Op::REC_KIND == 1)
                nl_entry_rec_density<DIM>(op, prm, g, p, q, px, py, pz, pl.recA, accepted);
            else nl_entry<DIM>(op, f, prm, g, p, q, px, py, pz, accepted);
        }
        pl.cnt[p] = (qtop - qbase) / (NL_BLOCK * 4u);
    } else {
        pl.cnt[p] = NL_NONE;
        atomicAdd(pl.overflow, 1ull);
        nl_walk<DIM>(op, f, prm, g, home, p, px, py, pz, cell_start, accepted);
    }
    if (self) op.template pair<DIM>(f, prm, p, p, 0.0, 0.0, 0.0, 0.0);  // core.jl:155-157
    op.template finish<DIM>(f, out, prm, p);
    if constexpr (REC && Op::REC_KIND == 1) {
        // what finish() just stored for p (same thread: reads see its own writes), as records
        nb_store(pl.recB + p, f.s[S_V0][p], f.s[S_V1][p], DIM == 3 ? f.s[S_V2][p] : 0.0, f.s[S_H][p]);
        const double rho = f.s[S_RHO][p];
        const double rfl = (rho != rho) ? rho : (rho < prm.rho_floor ? prm.rho_floor : rho);  // jl_max
        nb_store(pl.recC + p, f.s[S_PR2][p], rfl, f.s[S_CS][p], 0.0);
    }
    nl_count_pairs(pair_counter, accepted);
}

// REC (fused force pass only): entries are evaluated from the packed records
template <int DIM, class Op, bool REC = false>
__global__ void __launch_bounds__(NL_BLOCK)
k_binary_list(Fields f, Fields out, Params prm, Grid g, const uint32_t *__restrict__ key,
              const uint32_t *__restrict__ cellx, const uint32_t *__restrict__ cell_start, int64_t n,
              int self, unsigned long long *pair_counter, ColFilter cf, PairList pl) {
    const int64_t p = blockIdx.x * (int64_t)NL_BLOCK + threadIdx.x;
    if (p >= n) return;
    const CellCoord home = cell_of(g, key[p], cellx[p]);
    if (cf.on && !col_selected(cf, home.i)) {
        if (cf.copy) Op::template skip<DIM>(f, out, p);
        return;
    }
    const uint32_t cnt = pl.cnt[p];
    const double px = f.s[S_X0][p], py = f.s[S_X1][p], pz = DIM == 3 ? f.s[S_X2][p] : 0.0;
    Op op;
    op.template init<DIM>(f, prm, p);
    unsigned accepted = 0;
    if (cnt != NL_NONE) {
        const uint32_t *row = pl.list + ((size_t)(p >> 5) * (uint32_t)pl.stride) * 32 + (size_t)(p & 31);
        // (loading the next entry's position one iteration ahead as well was measured slower:
        // profiles/r01b_pair_list.md)
        uint32_t qn = cnt ? __ldcs(row) : 0u;
        for (uint32_t k = 0; k < cnt; ++k) {
            const uint32_t q = qn;
            if (k + 1 < cnt) qn = __ldcs(row + (size_t)(k + 1) * 32);  // one entry ahead
            if constexpr (REC && Op::REC_KIND == 2)
                nl_entry_rec_force<DIM>(op, prm, g, p, q, px, py, pz, pl, accepted);
            else nl_entry<DIM>(op, f, prm, g, p, q, px, py, pz, accepted);
        }
    } else {
        nl_walk<DIM>(op, f, prm, g, home, p, px, py, pz, cell_start, accepted);
    }
    if (self) op.template pair<DIM>(f, prm, p, p, 0.0, 0.0, 0.0, 0.0);
    op.template finish<DIM>(f, out, prm, p);
    nl_count_pairs(pair_counter, accepted);
}
