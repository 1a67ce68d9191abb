// Cell-centric, pair-parallel neighbour pass for the fused WCSPH operators.
//
// One warp owns one home cell at a time (grid-stride over cells).  All home particles of
// a cell share the same 9/27 neighbour cells, so
//   * the candidates (the neighbour cells' runs, in the reference's key_diff order) are
//     loaded ONCE per cell, 32 at a time, one candidate per lane (contiguous runs ->
//     coalesced), instead of once per particle;
//   * every lane tests its candidate against each home particle (exact r2 <= r2_max test,
//     bit-identical to `r > sys.h` of core.jl:105); the accepted (p, q) pairs are compacted
//     into a per-warp queue with ballot/popc, so the expensive closure body — FP64 divisions,
//     square roots, the artificial-viscosity branch — runs with (nearly) all 32 lanes busy
//     instead of at the 16-35 % acceptance rate a thread-per-particle loop suffers;
//   * the per-pair contributions are then added to each home particle by its own lane,
//     strictly in the reference's order (key_diff order, then the cell's stored order,
//     core.jl:96-110), so every FP64 sum is BIT-IDENTICAL to the generic thread-per-particle
//     kernel (k_binary) and to the oracle — tests/test_gpu_parity.py compares them bitwise.
//
// No tensor cores: this is not a dense contraction.  The pass is bound by the FP64 pipe
// and the L1/LSU path, not by HBM ("Rooflines").
#pragma once

#include "kernels_sph.cuh"
#include "sphmw_internal.h"

#define CP_WARPS 8          // warps per block
#define CP_HP 16            // home particles handled together (cells hold ~3-8)
#define CP_CMAX 256         // candidate positions staged per batch
#define CP_QCAP (CP_HP * 32)

__device__ __forceinline__ double cp_jl_max(double a, double b) {
    if (a != a || b != b) return a + b;
    return a < b ? b : a;
}

// ---- density + finalize + smoothing + pressure (wcsph_perturbed_witch.jl:316-323) --------
struct CP_Density {
    static constexpr int NPD = 4;  // x y z h
    static constexpr int NC = 1;
    double rho, hp;
    template <int DIM>
    __device__ void load(const Fields &f, const Params &, int64_t p, double *pd) {
        rho = 0.0;  // reset_density!
        hp = f.s[S_H][p];
        pd[0] = f.s[S_X0][p];
        pd[1] = f.s[S_X1][p];
        pd[2] = DIM == 3 ? f.s[S_X2][p] : 0.0;
        pd[3] = hp;
    }
    // closure body for one accepted pair -> contribution (no accumulation here)
    template <int DIM>
    static __device__ void eval(const Fields &f, const Params &, const double *pd, int64_t q,
                                double, double, double, double r, double *c) {
        c[0] = f.s[S_M][q] * sph_W<DIM>(pd[3], r);  // compute_density! :226-228
    }
    __device__ void accumulate(const double *c) { rho += c[0]; }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &c, int64_t p) {
        double y = f.s[S_X1][p];
        double rbg = c.rho0 * exp(-y * c.g / (c.R_mass * c.T_bg));  // finalize_density!
        double rho_p = rho - rbg;
        double rfl = cp_jl_max(rho, c.rho_floor);  // update_smoothing!
        double m = f.s[S_M][p];
        double hn = DIM == 2 ? c.eta * sqrt(m / rfl) : c.eta * cbrt(m / rfl);
        double pbg = c.R_mass * c.T_bg * rbg;  // compute_pressure!
        double pp = sph_pow2(c.c) * rho_p;
        double P = pbg + pp;
        f.s[S_RHO][p] = rho;
        f.s[S_RHO_BG][p] = rbg;
        f.s[S_RHO_P][p] = rho_p;
        f.s[S_H][p] = hn;
        f.s[S_P_BG][p] = pbg;
        f.s[S_P_P][p] = pp;
        f.s[S_P][p] = P;
        f.s[S_PR2][p] = pp / sph_pow2(rfl);
        f.s[S_CS][p] = sqrt(c.gamma * P / rfl);
    }
    template <int DIM>
    static __device__ void skip(const Fields &, const Fields &, int64_t) {}
};

// ---- balance_of_momentum! + accelerate! (wcsph_perturbed_witch.jl:261-286, :298-303) -----
struct CP_Momentum {
    static constexpr int NPD = 10;  // x y z vx vy vz h prho pr2 cs
    static constexpr int NC = 7;    // conservative (3), viscous (3), viscous flag
    double dv0, dv1, dv2, v0, v1, v2;
    template <int DIM>
    __device__ void load(const Fields &f, const Params &c, int64_t p, double *pd) {
        dv0 = dv1 = dv2 = 0.0;
        v0 = f.s[S_V0][p];
        v1 = f.s[S_V1][p];
        v2 = DIM == 3 ? f.s[S_V2][p] : 0.0;
        pd[0] = f.s[S_X0][p];
        pd[1] = f.s[S_X1][p];
        pd[2] = DIM == 3 ? f.s[S_X2][p] : 0.0;
        pd[3] = v0;
        pd[4] = v1;
        pd[5] = v2;
        pd[6] = f.s[S_H][p];
        pd[7] = cp_jl_max(f.s[S_RHO][p], c.rho_floor);
        pd[8] = f.s[S_PR2][p];
        pd[9] = f.s[S_CS][p];
    }
    template <int DIM>
    static __device__ void eval(const Fields &f, const Params &c, const double *pd, int64_t q,
                                double dx, double dy, double dz, double r, double *o) {
        double vx = pd[3] - f.s[S_V0][q], vy = pd[4] - f.s[S_V1][q];
        double dot_product = dx * vx + dy * vy;
        if (DIM == 3) {
            double vz = pd[5] - f.s[S_V2][q];
            dot_product = dot_product + dz * vz;
        }
        double h_ij = 0.5 * (pd[6] + f.s[S_H][q]);
        double ker = sph_rDW<DIM>(h_ij, r);
        double qm = f.s[S_M][q];
        double fc = -qm * (pd[8] + f.s[S_PR2][q]) * ker;
        o[0] = fc * dx;
        o[1] = fc * dy;
        o[2] = DIM == 3 ? fc * dz : 0.0;
        o[6] = 0.0;
        if (dot_product < 0.0) {
            double qrho = cp_jl_max(f.s[S_RHO][q], c.rho_floor);
            double c_ij = 0.5 * (pd[9] + f.s[S_CS][q]);
            double rho_ij = 0.5 * (pd[7] + qrho);
            double mu_ij = (h_ij * dot_product) / (r * r + c.eps * h_ij * h_ij);
            double pi_ij = (-c.alpha * c_ij * mu_ij + c.beta * mu_ij * mu_ij) / rho_ij;
            double fv = -qm * pi_ij * ker;
            o[3] = fv * dx;
            o[4] = fv * dy;
            o[5] = DIM == 3 ? fv * dz : 0.0;
            o[6] = 1.0;
        }
    }
    __device__ void accumulate(const double *c) {
        dv0 += c[0];
        dv1 += c[1];
        dv2 += c[2];
        if (c[6] != 0.0) {
            dv0 += c[3];
            dv1 += c[4];
            dv2 += c[5];
        }
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &out, const Params &c, int64_t p) {
        double n0 = v0, n1 = v1, n2 = v2;
        if (f.s[S_TYPE][p] == c.fluid) {  // accelerate!
            const double rho_p = f.s[S_RHO_P][p], rho = f.s[S_RHO][p];
            const bool sponge = f.s[S_X1][p] >= c.sponge_z0;
            const double hdt = 0.5 * c.dt;
            n0 = v0 + hdt * (dv0 + -c.g * 0.0 * rho_p / rho + (sponge ? c.sponge_y * 0.0 : 0.0));
            n1 = v1 + hdt * (dv1 + -c.g * 1.0 * rho_p / rho + (sponge ? c.sponge_y * 1.0 : 0.0));
            if (DIM == 3)
                n2 = v2 + hdt * (dv2 + -c.g * 0.0 * rho_p / rho + (sponge ? c.sponge_y * 0.0 : 0.0));
        }
        out.s[S_V0][p] = n0;
        out.s[S_V1][p] = n1;
        if (DIM == 3) out.s[S_V2][p] = n2;
    }
    template <int DIM>
    static __device__ void skip(const Fields &f, const Fields &out, int64_t p) {
        out.s[S_V0][p] = f.s[S_V0][p];
        out.s[S_V1][p] = f.s[S_V1][p];
        if (DIM == 3) out.s[S_V2][p] = f.s[S_V2][p];
    }
};

template <class OP>
struct CPWarpShared {
    uint32_t cand[CP_CMAX];
    unsigned short queue[CP_QCAP];
    double pdata[CP_HP][OP::NPD];
    double contrib[32][OP::NC];
};

template <int DIM, class OP>
__global__ void __launch_bounds__(CP_WARPS * 32)
k_cell_pairs(Fields f, Fields out, Params prm, Grid g, const uint32_t *__restrict__ cell_start,
             const uint32_t *__restrict__ cellx, int col_lo, int col_hi,
             unsigned long long *pair_counter) {
    __shared__ CPWarpShared<OP> shared[CP_WARPS];
    const unsigned lane = threadIdx.x & 31;
    const unsigned ltmask = (1u << lane) - 1u;
    CPWarpShared<OP> &ws = shared[threadIdx.x >> 5];
    const long long nwarps = (long long)gridDim.x * CP_WARPS;
    unsigned long long npairs = 0;

    for (long long cell = (long long)blockIdx.x * CP_WARPS + (threadIdx.x >> 5); cell < g.pkey_max;
         cell += nwarps) {
        const uint32_t hb = cell_start[cell], he = cell_start[cell + 1];
        if (hb == he) continue;
        const CellCoord home = cell_of(g, (unsigned)cell, cellx[hb]);
        if (col_lo > 0) {  // slab mode: ghost columns outside the pass are carried over
            if (home.i < col_lo || home.i > col_hi) {
                for (uint32_t p = hb + lane; p < he; p += 32) OP::template skip<DIM>(f, out, p);
                continue;
            }
        }
        // the neighbour cells' runs, lane d <-> key_diff[d]  (structs.jl:73-81 order)
        uint32_t rb = 0, len = 0;
        if ((int)lane < g.ndiff) {
            unsigned nk;
            if (neighbour_pkey(g, home, (int)lane, nk)) {  // core.jl:98 — no per-axis wrap check
                rb = cell_start[nk];
                len = cell_start[nk + 1] - rb;
            }
        }
        uint32_t incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += t;
        }
        const uint32_t off = incl - len;
        const uint32_t T = __shfl_sync(0xffffffffu, incl, 31);

        for (uint32_t hp0 = hb; hp0 < he; hp0 += CP_HP) {
            const int np = (int)min((uint32_t)CP_HP, he - hp0);
            OP st;
            __syncwarp();
            if ((int)lane < np) st.template load<DIM>(f, prm, hp0 + lane, ws.pdata[lane]);
            uint32_t my_lo = 0, my_cnt = 0;

            for (uint32_t c0 = 0; c0 < T; c0 += CP_CMAX) {
                __syncwarp();
                for (uint32_t i = 0; i < len; ++i) {
                    uint32_t t = off + i;
                    if (t >= c0 && t < c0 + CP_CMAX) ws.cand[t - c0] = rb + i;
                }
                __syncwarp();
                const uint32_t nb = min((uint32_t)CP_CMAX, T - c0);
                for (uint32_t k = 0; k < nb; k += 32) {
                    const bool valid = k + lane < nb;
                    const uint32_t q = valid ? ws.cand[k + lane] : 0u;
                    double qx = 0.0, qy = 0.0, qz = 0.0;
                    if (valid) {
                        qx = f.s[S_X0][q];
                        qy = f.s[S_X1][q];
                        if (DIM == 3) qz = f.s[S_X2][q];
                    }
                    // ---- filter: this lane's candidate against every home particle
                    uint32_t base = 0;
                    for (int j = 0; j < np; ++j) {
                        // dist(p,q) — core.jl:8-10, algebra.jl:49-60: left to right, no FMA
                        double dx = ws.pdata[j][0] - qx;
                        double dy = ws.pdata[j][1] - qy;
                        double r2 = dx * dx + dy * dy;
                        if (DIM == 3) {
                            double dz = ws.pdata[j][2] - qz;
                            r2 = r2 + dz * dz;
                        }
                        // core.jl:105 `r > sys.h || p == q` decided on r2 (Grid::r2_max)
                        const bool acc = valid && !(r2 > g.r2_max) && (q != hp0 + j);
                        const unsigned m = __ballot_sync(0xffffffffu, acc);
                        if (acc) ws.queue[base + __popc(m & ltmask)] = (unsigned short)(j | (lane << 8));
                        if ((int)lane == j) {
                            my_lo = base;
                            my_cnt = __popc(m);
                        }
                        base += __popc(m);
                    }
                    __syncwarp();
                    npairs += base;
                    // ---- evaluate the queued pairs 32 at a time, then add them in order
                    for (uint32_t e0 = 0; e0 < base; e0 += 32) {
                        const uint32_t e = e0 + lane;
                        if (e < base) {
                            const unsigned ent = ws.queue[e];
                            const int j = ent & 0xff;
                            const uint32_t qq = ws.cand[k + (ent >> 8)];
                            const double *pd = ws.pdata[j];
                            double dx = pd[0] - f.s[S_X0][qq];
                            double dy = pd[1] - f.s[S_X1][qq];
                            double dz = 0.0;
                            double r2 = dx * dx + dy * dy;
                            if (DIM == 3) {
                                dz = pd[2] - f.s[S_X2][qq];
                                r2 = r2 + dz * dz;
                            }
                            double r = sqrt(r2);
                            OP::template eval<DIM>(f, prm, pd, qq, dx, dy, dz, r, ws.contrib[lane]);
                        }
                        __syncwarp();
                        if ((int)lane < np) {
                            uint32_t lo = max(my_lo, e0), hi = min(my_lo + my_cnt, e0 + 32);
                            for (uint32_t t = lo; t < hi; ++t) st.accumulate(ws.contrib[t - e0]);
                        }
                        __syncwarp();
                    }
                }
            }
            if ((int)lane < np) st.template finish<DIM>(f, out, prm, hp0 + lane);
        }
    }
    if (pair_counter) {
        // every lane counted the same totals: one lane per warp reports
        if (lane == 0 && npairs) atomicAdd(pair_counter, npairs);
    }
}
