// Every closure of the operator menu as a device functor (src/current/*.jl,
// sph_jl/examples/collapse_dry.jl, sph_jl/tests/test_collision_2d.jl, src/utils/new_packing.jl,
// src/legacy/isothermal_flow_witch.jl), plus the menu itself.  The fused closures of the "wcsph"
// step live in wcsph_ops.cuh.  In a header so that the CPU emulation harness (tests/emu/) can
// compile the same functors for the host and compare them with the oracle.
//
// Arithmetic: compiled with -fmad=false; every product/sum is written in the reference's
// evaluation order (Julia's n-ary * and + fold left).
#pragma once
#include <math.h>

#include "kernels_sph.cuh"
#include "sphmw_internal.h"
#include "wcsph_ops.cuh"

// ===========================================================================
// Unary operators: struct with static void apply<DIM>(Fields&, Params&, pos)
// ===========================================================================
#define FLD(slot) f.s[slot][p]

// reset_density!  :220-223
struct U_wcsph_reset_density {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &, int64_t p) {
        FLD(S_RHO) = 0.0;
        FLD(S_RHO_P) = 0.0;
    }
};
// finalize_density!  :230-233
struct U_wcsph_finalize_density {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        double rbg = background_density(c, FLD(S_X1));
        FLD(S_RHO_BG) = rbg;
        FLD(S_RHO_P) = FLD(S_RHO) - rbg;
    }
};
// update_smoothing!  :235-238  (3D extrusion: cube root)
struct U_wcsph_update_smoothing {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        double rho = jl_max(FLD(S_RHO), c.rho_floor);
        FLD(S_H) = DIM == 2 ? c.eta * sqrt(FLD(S_M) / rho) : c.eta * cbrt(FLD(S_M) / rho);
    }
};
// compute_pressure!  :195-199
struct U_wcsph_compute_pressure {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        double pbg = background_pressure(c, FLD(S_X1));
        double pp = sph_pow2(c.c) * FLD(S_RHO_P);
        FLD(S_P_BG) = pbg;
        FLD(S_P_P) = pp;
        FLD(S_P) = pbg + pp;
    }
};
// find_temperature!  :205-208
struct U_wcsph_find_temperature {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        double T = FLD(S_P) / (c.R_mass * FLD(S_RHO));
        FLD(S_T) = T;
        FLD(S_T_P) = T - FLD(S_T_BG);
    }
};
// find_pot_temp!  :210-214
struct U_wcsph_find_pot_temp {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        double th = FLD(S_T) * pow((c.T_bg * c.R_gas * c.rho0) / FLD(S_P), 2.0 / 7.0);
        double thbg = background_pot_temperature(c, FLD(S_X1));
        FLD(S_TH) = th;
        FLD(S_TH_BG) = thbg;
        FLD(S_TH_P) = th - thbg;
    }
};

// hopkins_perturbed_witch.jl:200-203
struct U_hopkins_reset_pressure {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &, int64_t p) {
        FLD(S_P) = 0.0;
        FLD(S_P_P) = 0.0;
    }
};
// hopkins_perturbed_witch.jl:210-214
struct U_hopkins_finalize_pressure {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        double P = pow(FLD(S_P), c.gamma);
        double pbg = background_pressure(c, FLD(S_X1));
        FLD(S_P) = P;
        FLD(S_P_BG) = pbg;
        FLD(S_P_P) = P - pbg;
    }
};
// hopkins_total_witch.jl:170-172
struct U_ht_reset_pressure {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &, int64_t p) { FLD(S_P) = 0.0; }
};
// hopkins_total_witch.jl:179-181
struct U_ht_finalize_pressure {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        FLD(S_P) = pow(FLD(S_P), c.gamma);
    }
};
// hopkins_total_witch.jl:187-189
struct U_ht_find_temperature {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        FLD(S_T) = FLD(S_P) / (c.R_mass * FLD(S_RHO));
    }
};
// hopkins_total_witch.jl:191-193
struct U_ht_find_pot_temp {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        FLD(S_TH) = FLD(S_T) * pow((c.T_bg * c.R_gas * c.rho0) / FLD(S_P), 2.0 / 7.0);
    }
};
// hopkins_total_witch.jl:203-205
struct U_ht_reset_density {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &, int64_t p) { FLD(S_RHO) = 0.0; }
};
// hopkins_total_witch.jl:270-272 — not type-gated
struct U_ht_move {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        FLD(S_X0) += c.dt * FLD(S_V0);
        FLD(S_X1) += c.dt * FLD(S_V1);
        if (DIM == 3) FLD(S_X2) += c.dt * FLD(S_V2);
    }
};
// hopkins_total_witch.jl:274-277, gravity :225-228
struct U_ht_accelerate {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        const bool sponge = FLD(S_X1) >= c.sponge_z0;
        const double hdt = 0.5 * c.dt;
        FLD(S_V0) += hdt * (FLD(S_DV0) + -c.g * 0.0 + (sponge ? c.sponge_y * 0.0 : 0.0));
        FLD(S_V1) += hdt * (FLD(S_DV1) + -c.g * 1.0 + (sponge ? c.sponge_y * 1.0 : 0.0));
        if (DIM == 3)
            FLD(S_V2) += hdt * (FLD(S_DV2) + -c.g * 0.0 + (sponge ? c.sponge_y * 0.0 : 0.0));
        FLD(S_DV0) = 0.0;
        FLD(S_DV1) = 0.0;
        if (DIM == 3) FLD(S_DV2) = 0.0;
    }
};

// collapse_dry.jl:123-127
struct U_dam_find_pressure {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        double rho = FLD(S_RHO) + FLD(S_DRHO) * c.dt;
        FLD(S_RHO) = rho;
        FLD(S_DRHO) = 0.0;
        FLD(S_P) = sph_pow2(c.c) * (rho - c.rho0);
    }
};
// collapse_dry.jl:148-153
struct U_dam_move {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        FLD(S_DV0) = 0.0;
        FLD(S_DV1) = 0.0;
        if (DIM == 3) FLD(S_DV2) = 0.0;
        if (FLD(S_TYPE) == c.fluid) {
            FLD(S_X0) += 0.5 * c.dt * FLD(S_V0);
            FLD(S_X1) += 0.5 * c.dt * FLD(S_V1);
            if (DIM == 3) FLD(S_X2) += 0.5 * c.dt * FLD(S_V2);
        }
    }
};
// collapse_dry.jl:155-159
struct U_dam_accelerate {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        if (FLD(S_TYPE) == c.fluid) {
            FLD(S_V0) += 0.5 * c.dt * (FLD(S_DV0) + c.gx);
            FLD(S_V1) += 0.5 * c.dt * (FLD(S_DV1) + c.gy);
            if (DIM == 3) FLD(S_V2) += 0.5 * c.dt * (FLD(S_DV2) + c.gz);
        }
    }
};
// test_collision_2d.jl:74-76
struct U_col_find_pressure {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        FLD(S_P) = sph_pow2(c.c) * (FLD(S_RHO) - FLD(S_RHO0));
    }
};
// test_collision_2d.jl:83-85
struct U_col_reset_a {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &, int64_t p) {
        FLD(S_DV0) = 0.0;
        FLD(S_DV1) = 0.0;
        if (DIM == 3) FLD(S_DV2) = 0.0;
    }
};
// test_collision_2d.jl:87-89
struct U_col_reset_rho {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &, int64_t p) { FLD(S_RHO) = 0.0; }
};
// test_collision_2d.jl:91-93
struct U_col_move {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        FLD(S_X0) += c.dt * FLD(S_V0);
        FLD(S_X1) += c.dt * FLD(S_V1);
        if (DIM == 3) FLD(S_X2) += c.dt * FLD(S_V2);
    }
};
// test_collision_2d.jl:95-97
struct U_col_accelerate {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        FLD(S_V0) += 0.5 * c.dt * FLD(S_DV0);
        FLD(S_V1) += 0.5 * c.dt * FLD(S_DV1);
        if (DIM == 3) FLD(S_V2) += 0.5 * c.dt * FLD(S_DV2);
    }
};
// new_packing.jl:5-9
struct U_pack_reset_rho {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        if (FLD(S_TYPE) == c.fluid) FLD(S_RHO) = 0.0;
    }
};
// new_packing.jl:49-57
struct U_pack_accelerate {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        if (FLD(S_TYPE) == c.fluid) {
            double den = 1.0 + c.zeta_pack * c.dt_pack;
            FLD(S_V0) = (FLD(S_V0) + c.dt_pack * FLD(S_DV0)) / den;
            FLD(S_V1) = (FLD(S_V1) + c.dt_pack * FLD(S_DV1)) / den;
            if (DIM == 3) FLD(S_V2) = (FLD(S_V2) + c.dt_pack * FLD(S_DV2)) / den;
        }
        FLD(S_DV0) = 0.0;
        FLD(S_DV1) = 0.0;
        if (DIM == 3) FLD(S_DV2) = 0.0;
    }
};
// new_packing.jl:59-63
struct U_pack_move {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        if (FLD(S_TYPE) == c.fluid) {
            FLD(S_X0) += c.dt_pack * FLD(S_V0);
            FLD(S_X1) += c.dt_pack * FLD(S_V1);
            if (DIM == 3) FLD(S_X2) += c.dt_pack * FLD(S_V2);
        }
    }
};
// ---- src/legacy/isothermal_flow_witch.jl (u stored in v, Du in Dv, T = T_bg, h = kh) ----
// find_pressure!  :156-160
struct U_flow_find_pressure {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        double rho = FLD(S_RHO) + FLD(S_DRHO) * c.dt;
        FLD(S_RHO) = rho;
        FLD(S_DRHO) = 0.0;
        FLD(S_P) = rho * c.R_mass * c.T_bg;
    }
};
// set_density!  :162-164
struct U_flow_set_density {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        FLD(S_RHO) = c.rho0 * exp(-FLD(S_X1) * c.g / (c.R_mass * c.T_bg));
    }
};
// find_pot_temp!  :167-169
struct U_flow_find_pot_temp {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        FLD(S_TH) = c.T_bg * pow((c.T_bg * c.R_gas * c.rho0) / FLD(S_P), c.R_gas / c.cp);
    }
};
// move!  :204-209
struct U_flow_move {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        FLD(S_DV0) = 0.0;
        FLD(S_DV1) = 0.0;
        if (DIM == 3) FLD(S_DV2) = 0.0;
        double t = FLD(S_TYPE);
        if (t == c.fluid || t == c.inflow) {
            FLD(S_X0) += c.dt * FLD(S_V0);
            FLD(S_X1) += c.dt * FLD(S_V1);
            if (DIM == 3) FLD(S_X2) += c.dt * FLD(S_V2);
        }
    }
};
// accelerate!  :211-215, damping_structure :192-198 (positive scalar; sponge_y = -that)
struct U_flow_accelerate {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        if (FLD(S_TYPE) == c.fluid) {
            double damp = FLD(S_X1) >= c.sponge_z0 ? -c.sponge_y : 0.0;
            FLD(S_V0) += 0.5 * c.dt * (FLD(S_DV0) - c.g * 0.0 - damp * 0.0);
            FLD(S_V1) += 0.5 * c.dt * (FLD(S_DV1) - c.g * 1.0 - damp * 1.0);
            if (DIM == 3) FLD(S_V2) += 0.5 * c.dt * (FLD(S_DV2) - c.g * 0.0 - damp * 0.0);
        }
    }
};
// ---- src/legacy/adiabatic_flow_witch.jl (u in v, Du in Dv, T0 = T_bg, h = kh, cv = cp - R_mass) ----
// accelerate! (:225-229) is U_flow_accelerate, internal_force! (:146-153) is B_flow_force.
// move!  :217-223
struct U_aflow_move {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        FLD(S_DV0) = 0.0;
        FLD(S_DV1) = 0.0;
        if (DIM == 3) FLD(S_DV2) = 0.0;
        if (FLD(S_TYPE) == c.fluid) {
            FLD(S_X0) += c.dt * FLD(S_V0);
            FLD(S_X1) += c.dt * FLD(S_V1);
            if (DIM == 3) FLD(S_X2) += c.dt * FLD(S_V2);
            FLD(S_RHO) = 0.0;
        }
    }
};
// find_s!  :165-169
struct U_aflow_find_s {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        if (FLD(S_TYPE) == c.fluid) FLD(S_ENT_D) = FLD(S_ENT) * FLD(S_RHO) / FLD(S_M);
    }
};
// find_pressure!  :171-176
struct U_aflow_find_pressure {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        if (FLD(S_TYPE) == c.fluid) {
            const double cv = c.cp - c.R_mass;
            const double rho = FLD(S_RHO);
            const double T = (pow(rho, c.gamma - 1.0)) * exp(FLD(S_ENT_D) / (rho * cv)) / (cv * (c.gamma - 1.0));
            FLD(S_T) = T;
            FLD(S_P) = c.R_mass * rho * T;
        }
    }
};
// find_pot_temp!  :178-182
struct U_aflow_find_pot_temp {
    template <int DIM>
    static __device__ void apply(const Fields &f, const Params &c, int64_t p) {
        if (FLD(S_TYPE) == c.fluid) {
            const double b = (c.T_bg * c.R_gas * c.rho0) / FLD(S_P);
            FLD(S_TH) = FLD(S_T) * pow(b * b, 1.0 / 7.0);
        }
    }
};
#undef FLD

// ===========================================================================
// Binary operators.  A functor keeps the fields of p it reads/writes in
// registers: init() loads them, pair() is the closure body for one accepted
// neighbour q, finish() stores what the closure wrote to p.
// ===========================================================================
#define PF(slot) f.s[slot][p]
#define QF(slot) f.s[slot][q]

// compute_density!  wcsph_perturbed_witch.jl:226-228
struct B_wcsph_density : PairOpBase {
    double rho, hp;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &, int64_t p) {
        rho = PF(S_RHO);
        hp = PF(S_H);
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &, int64_t, int64_t q, double, double,
                         double, double r) {
        rho += QF(S_M) * sph_W<DIM>(hp, r);
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &, int64_t p) {
        PF(S_RHO) = rho;
    }
};

// shared body of balance_of_momentum!  wcsph_perturbed_witch.jl:261-286
struct MomentumState {
    double dv0, dv1, dv2;
    double v0, v1, v2, hp, rho, prho, Pp, P;
};

struct B_wcsph_momentum : PairOpBase {
    MomentumState s;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &c, int64_t p) {
        s.dv0 = PF(S_DV0);
        s.dv1 = PF(S_DV1);
        s.dv2 = DIM == 3 ? PF(S_DV2) : 0.0;
        s.v0 = PF(S_V0);
        s.v1 = PF(S_V1);
        s.v2 = DIM == 3 ? PF(S_V2) : 0.0;
        s.hp = PF(S_H);
        s.rho = PF(S_RHO);
        s.prho = jl_max(s.rho, c.rho_floor);
        s.Pp = PF(S_P_P);
        s.P = PF(S_P);
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &c, int64_t, int64_t q, double dx, double dy,
                         double dz, double r) {
        double vx = s.v0 - QF(S_V0), vy = s.v1 - QF(S_V1);
        double dot_product = dx * vx + dy * vy;
        if (DIM == 3) {
            double vz = s.v2 - QF(S_V2);
            dot_product = dot_product + dz * vz;
        }
        double h_ij = 0.5 * (s.hp + QF(S_H));
        double ker = sph_rDW<DIM>(h_ij, r);
        double qrho = jl_max(QF(S_RHO), c.rho_floor);
        double qm = QF(S_M);
        // -q.m * (p.P'/prho^2 + q.P'/qrho^2) * ker * x_pq  (left fold, vector last)
        double fc = -qm * (s.Pp / sph_pow2(s.prho) + QF(S_P_P) / sph_pow2(qrho)) * ker;
        s.dv0 += fc * dx;
        s.dv1 += fc * dy;
        if (DIM == 3) s.dv2 += fc * dz;
        if (dot_product < 0.0) {
            double c_i = sqrt(c.gamma * s.P / s.prho);
            double c_j = sqrt(c.gamma * QF(S_P) / qrho);
            double c_ij = 0.5 * (c_i + c_j);
            double rho_ij = 0.5 * (s.prho + qrho);
            double mu_ij = (h_ij * dot_product) / (r * r + c.eps * h_ij * h_ij);
            double pi_ij = (-c.alpha * c_ij * mu_ij + c.beta * mu_ij * mu_ij) / rho_ij;
            double fv = -qm * pi_ij * ker;
            s.dv0 += fv * dx;
            s.dv1 += fv * dy;
            if (DIM == 3) s.dv2 += fv * dz;
        }
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &, int64_t p) {
        PF(S_DV0) = s.dv0;
        PF(S_DV1) = s.dv1;
        if (DIM == 3) PF(S_DV2) = s.dv2;
    }
};

// compute_pressure! (binary)  hopkins_perturbed_witch.jl:205-208
struct B_hopkins_pressure : PairOpBase {
    double P, hp;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &, int64_t p) {
        P = PF(S_P);
        hp = PF(S_H);
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &c, int64_t, int64_t q, double, double,
                         double, double r) {
        double ker = sph_W<DIM>(0.5 * (hp + QF(S_H)), r);
        P += QF(S_M) * pow(QF(S_A), 1 / c.gamma) * ker;
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &, int64_t p) {
        PF(S_P) = P;
    }
};

// balance_of_momentum!  hopkins_total_witch.jl:233-264
struct B_ht_momentum : PairOpBase {
    MomentumState s;
    double A;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &c, int64_t p) {
        s.dv0 = PF(S_DV0);
        s.dv1 = PF(S_DV1);
        s.dv2 = DIM == 3 ? PF(S_DV2) : 0.0;
        s.v0 = PF(S_V0);
        s.v1 = PF(S_V1);
        s.v2 = DIM == 3 ? PF(S_V2) : 0.0;
        s.hp = PF(S_H);
        s.rho = PF(S_RHO);
        s.prho = jl_max(s.rho, c.rho_floor);
        s.P = PF(S_P);
        A = PF(S_A);
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &c, int64_t, int64_t q, double dx, double dy,
                         double dz, double r) {
        double vx = s.v0 - QF(S_V0), vy = s.v1 - QF(S_V1);
        double dot_product = dx * vx + dy * vy;
        if (DIM == 3) {
            double vz = s.v2 - QF(S_V2);
            dot_product = dot_product + dz * vz;
        }
        double qm = QF(S_M), qh = QF(S_H), qPraw = QF(S_P);
        double prefac = qm * pow(A * QF(S_A), 1 / c.gamma);
        double expfac = 1.0 - 2.0 / c.gamma;
        double ker_i = sph_rDW<DIM>(s.hp, r);
        double ker_j = sph_rDW<DIM>(qh, r);
        double pP = jl_max(c.P_floor, s.P);
        double qP = jl_max(c.P_floor, qPraw);
        double fc = -prefac * (pow(pP, expfac) * ker_i + pow(qP, expfac) * ker_j);
        s.dv0 += fc * dx;
        s.dv1 += fc * dy;
        if (DIM == 3) s.dv2 += fc * dz;
        if (dot_product < 0.0) {
            double h_ij = 0.5 * (s.hp + qh);
            double ker_ij = sph_rDW<DIM>(h_ij, r);
            double qrho = jl_max(QF(S_RHO), c.rho_floor);
            double c_i = sqrt(c.gamma * s.P / s.prho);
            double c_j = sqrt(c.gamma * qPraw / qrho);
            double c_ij = 0.5 * (c_i + c_j);
            double rho_ij = 0.5 * (s.prho + qrho);
            double mu_ij = (h_ij * dot_product) / (r * r + c.eps * h_ij * h_ij);
            double pi_ij = (-c.alpha * c_ij * mu_ij + c.beta * mu_ij * mu_ij) / rho_ij;
            double fv = -qm * pi_ij * ker_ij;
            s.dv0 += fv * dx;
            s.dv1 += fv * dy;
            if (DIM == 3) s.dv2 += fv * dz;
        }
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &, int64_t p) {
        PF(S_DV0) = s.dv0;
        PF(S_DV1) = s.dv1;
        if (DIM == 3) PF(S_DV2) = s.dv2;
    }
};

// balance_of_momentum!  full_hopkins_perturbed_witch.jl:284-326
struct B_hf_momentum : PairOpBase {
    MomentumState s;
    double A, A_bg, P_bg;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &c, int64_t p) {
        s.dv0 = PF(S_DV0);
        s.dv1 = PF(S_DV1);
        s.dv2 = DIM == 3 ? PF(S_DV2) : 0.0;
        s.v0 = PF(S_V0);
        s.v1 = PF(S_V1);
        s.v2 = DIM == 3 ? PF(S_V2) : 0.0;
        s.hp = PF(S_H);
        s.rho = PF(S_RHO);
        s.prho = jl_max(s.rho, c.rho_floor);
        s.P = PF(S_P);
        A = PF(S_A);
        A_bg = PF(S_A_BG);
        P_bg = PF(S_P_BG);
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &c, int64_t, int64_t q, double dx, double dy,
                         double dz, double r) {
        double vx = s.v0 - QF(S_V0), vy = s.v1 - QF(S_V1);
        double dot_product = dx * vx + dy * vy;
        if (DIM == 3) {
            double vz = s.v2 - QF(S_V2);
            dot_product = dot_product + dz * vz;
        }
        double qm = QF(S_M), qh = QF(S_H), qPraw = QF(S_P);
        double prefac = qm * pow(A * QF(S_A), 1 / c.gamma);
        double expfac = 1.0 - 2.0 / c.gamma;
        double ker_i = sph_rDW<DIM>(s.hp, r);
        double ker_j = sph_rDW<DIM>(qh, r);
        double pP = jl_max(c.P_floor, s.P);
        double qP = jl_max(c.P_floor, qPraw);
        double f_tot = -prefac * (pow(pP, expfac) * ker_i + pow(qP, expfac) * ker_j);
        double prefac_bg = qm * pow(A_bg * QF(S_A_BG), 1 / c.gamma);
        double pP_bg = jl_max(c.P_floor, P_bg);
        double qP_bg = jl_max(c.P_floor, QF(S_P_BG));
        double f_bg = -prefac_bg * (pow(pP_bg, expfac) * ker_i + pow(qP_bg, expfac) * ker_j);
        s.dv0 += f_tot * dx - f_bg * dx;  // p.Dv += a_tot - a_bg
        s.dv1 += f_tot * dy - f_bg * dy;
        if (DIM == 3) s.dv2 += f_tot * dz - f_bg * dz;
        if (dot_product < 0.0) {
            double h_ij = 0.5 * (s.hp + qh);
            double ker_ij = sph_rDW<DIM>(h_ij, r);
            double qrho = jl_max(QF(S_RHO), c.rho_floor);
            double c_i = sqrt(c.gamma * s.P / s.prho);
            double c_j = sqrt(c.gamma * qPraw / qrho);
            double c_ij = 0.5 * (c_i + c_j);
            double rho_ij = 0.5 * (s.prho + qrho);
            double mu_ij = (h_ij * dot_product) / (r * r + c.eps * h_ij * h_ij);
            double pi_ij = (-c.alpha * c_ij * mu_ij + c.beta * mu_ij * mu_ij) / rho_ij;
            double fv = -qm * pi_ij * ker_ij;
            s.dv0 += fv * dx;
            s.dv1 += fv * dy;
            if (DIM == 3) s.dv2 += fv * dz;
        }
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &, int64_t p) {
        PF(S_DV0) = s.dv0;
        PF(S_DV1) = s.dv1;
        if (DIM == 3) PF(S_DV2) = s.dv2;
    }
};

// balance_of_mass!  collapse_dry.jl:112-115 (fixed h = kh, fixed mass m)
struct B_dam_mass : PairOpBase {
    double drho, v0, v1, v2, rho;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &, int64_t p) {
        drho = PF(S_DRHO);
        v0 = PF(S_V0);
        v1 = PF(S_V1);
        v2 = DIM == 3 ? PF(S_V2) : 0.0;
        rho = PF(S_RHO);
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &c, int64_t, int64_t q, double dx, double dy,
                         double dz, double r) {
        double ker = c.m * rDwendland2(c.kh, r);
        double d = dx * (v0 - QF(S_V0)) + dy * (v1 - QF(S_V1));
        if (DIM == 3) d = d + dz * (v2 - QF(S_V2));
        drho += ker * (d + 2 * c.nu * (rho - QF(S_RHO)));
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &, int64_t p) {
        PF(S_DRHO) = drho;
    }
};
// internal_force!  collapse_dry.jl:135-141
struct B_dam_force : PairOpBase {
    double dv0, dv1, dv2, v0, v1, v2, P, rho;
    bool fluid;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &c, int64_t p) {
        dv0 = PF(S_DV0);
        dv1 = PF(S_DV1);
        dv2 = DIM == 3 ? PF(S_DV2) : 0.0;
        v0 = PF(S_V0);
        v1 = PF(S_V1);
        v2 = DIM == 3 ? PF(S_V2) : 0.0;
        P = PF(S_P);
        rho = PF(S_RHO);
        fluid = PF(S_TYPE) == c.fluid;
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &c, int64_t, int64_t q, double dx, double dy,
                         double dz, double r) {
        if (!fluid) return;
        double ker = c.m * rDwendland2(c.kh, r);
        double qrho = QF(S_RHO);
        double a1 = -ker * (P / sph_pow2(rho) + QF(S_P) / sph_pow2(qrho));
        dv0 += a1 * dx;
        dv1 += a1 * dy;
        if (DIM == 3) dv2 += a1 * dz;
        double a2 = +2 * ker * c.mu / sph_pow2(c.rho0);
        dv0 += a2 * (v0 - QF(S_V0));
        dv1 += a2 * (v1 - QF(S_V1));
        if (DIM == 3) dv2 += a2 * (v2 - QF(S_V2));
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &, int64_t p) {
        PF(S_DV0) = dv0;
        PF(S_DV1) = dv1;
        if (DIM == 3) PF(S_DV2) = dv2;
    }
};
// balance_of_mass!  isothermal_flow_witch.jl:140-143
struct B_flow_mass : PairOpBase {
    double drho, v0, v1, v2;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &, int64_t p) {
        drho = PF(S_DRHO);
        v0 = PF(S_V0);
        v1 = PF(S_V1);
        v2 = DIM == 3 ? PF(S_V2) : 0.0;
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &c, int64_t, int64_t q, double dx, double dy,
                         double dz, double r) {
        double ker = QF(S_M) * rDwendland2(c.kh, r);
        double d = dx * (v0 - QF(S_V0)) + dy * (v1 - QF(S_V1));
        if (DIM == 3) d = d + dz * (v2 - QF(S_V2));
        drho += ker * d;
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &, int64_t p) {
        PF(S_DRHO) = drho;
    }
};
// find_density!  adiabatic_flow_witch.jl:159-163 (the driver applies it with self = true)
struct B_aflow_density : PairOpBase {
    double rho;
    bool fluid;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &c, int64_t p) {
        rho = PF(S_RHO);
        fluid = PF(S_TYPE) == c.fluid;
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &c, int64_t, int64_t q, double, double, double, double r) {
        if (fluid && QF(S_TYPE) == c.fluid) rho += QF(S_M) * wendland2(c.kh, r);
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &, int64_t p) {
        PF(S_RHO) = rho;
    }
};
// entropy_production!  adiabatic_flow_witch.jl:184-191
struct B_aflow_entropy : PairOpBase {
    double S, v0, v1, v2, m, T, rho;
    bool fluid;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &c, int64_t p) {
        S = PF(S_ENT);
        v0 = PF(S_V0);
        v1 = PF(S_V1);
        v2 = DIM == 3 ? PF(S_V2) : 0.0;
        m = PF(S_M);
        T = PF(S_T);
        rho = PF(S_RHO);
        fluid = PF(S_TYPE) == c.fluid;
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &c, int64_t, int64_t q, double dx, double dy, double dz,
                         double r) {
        if (fluid && QF(S_TYPE) == c.fluid) {
            double ker = rDwendland2(c.kh, r);
            double d = (v0 - QF(S_V0)) * dx + (v1 - QF(S_V1)) * dy;  // dot(u_pq, x_pq)
            if (DIM == 3) d = d + (v2 - QF(S_V2)) * dz;
            S += -4.0 * m * QF(S_M) * ker * c.mu / (T * rho * QF(S_RHO)) * (d * d) / (r * r + 0.01 * c.kh * c.kh) * c.dt;
        }
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &, int64_t p) {
        PF(S_ENT) = S;
    }
};
// internal_force!  isothermal_flow_witch.jl:145-150
struct B_flow_force : PairOpBase {
    double dv0, dv1, dv2, v0, v1, v2, P, rho;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &, int64_t p) {
        dv0 = PF(S_DV0);
        dv1 = PF(S_DV1);
        dv2 = DIM == 3 ? PF(S_DV2) : 0.0;
        v0 = PF(S_V0);
        v1 = PF(S_V1);
        v2 = DIM == 3 ? PF(S_V2) : 0.0;
        P = PF(S_P);
        rho = PF(S_RHO);
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &c, int64_t, int64_t q, double dx, double dy,
                         double dz, double r) {
        double ker = QF(S_M) * rDwendland2(c.kh, r);
        double qrho = QF(S_RHO);
        double a1 = -ker * (P / sph_pow2(rho) + QF(S_P) / sph_pow2(qrho));
        dv0 += a1 * dx;
        dv1 += a1 * dy;
        if (DIM == 3) dv2 += a1 * dz;
        double d = (v0 - QF(S_V0)) * dx + (v1 - QF(S_V1)) * dy;  // dot(p.u - q.u, x_pq)
        if (DIM == 3) d = d + (v2 - QF(S_V2)) * dz;
        double a2 = 8.0 * ker * c.mu / (rho * qrho) * d / (r * r + 0.01 * c.kh * c.kh);
        dv0 += a2 * dx;
        dv1 += a2 * dy;
        if (DIM == 3) dv2 += a2 * dz;
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &, int64_t p) {
        PF(S_DV0) = dv0;
        PF(S_DV1) = dv1;
        if (DIM == 3) PF(S_DV2) = dv2;
    }
};
// find_rho! / find_rho0!  test_collision_2d.jl:66-72
template <int SLOT>
struct B_col_rho : PairOpBase {
    double acc;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &, int64_t p) { acc = PF(SLOT); }
    template <int DIM>
    __device__ void pair(const Fields &, const Params &c, int64_t, int64_t, double, double, double,
                         double r) {
        acc += c.m * wendland2(c.kh, r);
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &, int64_t p) {
        PF(SLOT) = acc;
    }
};
// internal_force!  test_collision_2d.jl:78-81
struct B_col_force : PairOpBase {
    double dv0, dv1, dv2, P;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &, int64_t p) {
        dv0 = PF(S_DV0);
        dv1 = PF(S_DV1);
        dv2 = DIM == 3 ? PF(S_DV2) : 0.0;
        P = PF(S_P);
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &c, int64_t, int64_t q, double dx, double dy,
                         double dz, double r) {
        double ker = c.m * rDwendland2(c.kh, r);
        double a = -ker * (P / sph_pow2(c.rho0) + QF(S_P) / sph_pow2(c.rho0));
        dv0 += a * dx;
        dv1 += a * dy;
        if (DIM == 3) dv2 += a * dz;
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &, int64_t p) {
        PF(S_DV0) = dv0;
        PF(S_DV1) = dv1;
        if (DIM == 3) PF(S_DV2) = dv2;
    }
};
// accumulate_rho_pack!  new_packing.jl:11-15
struct B_pack_rho : PairOpBase {
    double rho, hp;
    bool fluid;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &c, int64_t p) {
        rho = PF(S_RHO);
        hp = PF(S_H);
        fluid = PF(S_TYPE) == c.fluid;
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &, int64_t, int64_t q, double, double, double,
                         double r) {
        if (fluid) rho += QF(S_M) * sph_W<DIM>(hp, r);
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &, int64_t p) {
        PF(S_RHO) = rho;
    }
};
// balance_of_momentum_pack!  new_packing.jl:23-46
struct B_pack_momentum : PairOpBase {
    double dv0, dv1, dv2, hp, rho_i, Pi, y;
    bool fluid;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &c, int64_t p) {
        dv0 = PF(S_DV0);
        dv1 = PF(S_DV1);
        dv2 = DIM == 3 ? PF(S_DV2) : 0.0;
        hp = PF(S_H);
        y = PF(S_X1);
        rho_i = jl_max(PF(S_RHO), c.rho_floor);
        Pi = sph_pow2(c.c_pack) * (rho_i - background_density(c, y));
        fluid = PF(S_TYPE) == c.fluid;
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &c, int64_t, int64_t q, double, double dy,
                         double, double r) {
        if (!(fluid && QF(S_TYPE) == c.fluid)) return;
        double rho_j = jl_max(QF(S_RHO), c.rho_floor);
        double Pj = sph_pow2(c.c_pack) * (rho_j - background_density(c, QF(S_X1)));
        double ker = sph_rDW<DIM>(0.5 * (hp + QF(S_H)), r);
        double f1 = -QF(S_M) * (Pi / sph_pow2(rho_i) + Pj / sph_pow2(rho_j)) * ker * dy;
        dv0 += f1 * 0.0;
        dv1 += f1 * 1.0;
        if (DIM == 3) dv2 += f1 * 0.0;
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &, int64_t p) {
        PF(S_DV0) = dv0;
        PF(S_DV1) = dv1;
        if (DIM == 3) PF(S_DV2) = dv2;
    }
};

// ---- fused passes of the Hopkins drivers (sphmw_step, schemes "hopkins", "hopkins_full") -------
// hopkins_perturbed_witch.jl:324-349 runs three binary operators on one cell list; fused:
//   pass 1 (records the pair list)  reset_density! + compute_density! + finalize_density! +
//                                   update_smoothing!                                  (:331-334)
//   pass 2 (replays it)             reset_pressure! + compute_pressure! + finalize_pressure!
//                                                                                      (:337-339)
//   pass 3 (replays it)             balance_of_momentum! + accelerate!                 (:346-347)
// find_temperature! / find_pot_temp! (:342-343) are diagnostics, rebuilt on demand.  Each fused
// operator performs its unfused parts' arithmetic in the same order: same bits.
struct B_hopkins_density_fused : PairOpBase {
    double rho, hp;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &, int64_t p) {
        rho = 0.0;  // reset_density!
        hp = PF(S_H);
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &, int64_t, int64_t q, double, double, double, double r) {
        rho += QF(S_M) * sph_W<DIM>(hp, r);
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &c, int64_t p) {
        double rbg = background_density(c, PF(S_X1));  // finalize_density!
        PF(S_RHO) = rho;
        PF(S_RHO_BG) = rbg;
        PF(S_RHO_P) = rho - rbg;
        double rfl = jl_max(rho, c.rho_floor);         // update_smoothing!
        // (in place: the density closure reads only the particle's OWN h, never a neighbour's)
        double m = PF(S_M);
        PF(S_H) = DIM == 2 ? c.eta * sqrt(m / rfl) : c.eta * cbrt(m / rfl);
    }
};
struct B_hopkins_pressure_fused : PairOpBase {
    double P, hp;
    template <int DIM>
    __device__ void init(const Fields &f, const Params &, int64_t p) {
        P = 0.0;  // reset_pressure!
        hp = PF(S_H);
    }
    template <int DIM>
    __device__ void pair(const Fields &f, const Params &c, int64_t, int64_t q, double, double, double, double r) {
        double ker = sph_W<DIM>(0.5 * (hp + QF(S_H)), r);
        P += QF(S_M) * pow(QF(S_A), 1 / c.gamma) * ker;
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &, const Params &c, int64_t p) {
        double Pf = pow(P, c.gamma);  // finalize_pressure!
        double pbg = background_pressure(c, PF(S_X1));
        PF(S_P) = Pf;
        PF(S_P_BG) = pbg;
        PF(S_P_P) = Pf - pbg;
    }
};
// a force operator + the trailing accelerate! of the step: Dv starts at 0 and is never stored, the
// new velocity goes to `out` (neighbours still read the old one)
template <class Force>
struct B_force_kick_fused : Force {
    template <int DIM>
    static __device__ void skip(const Fields &f, const Fields &out, int64_t p) {
        out.s[S_V0][p] = f.s[S_V0][p];
        out.s[S_V1][p] = f.s[S_V1][p];
        if (DIM == 3) out.s[S_V2][p] = f.s[S_V2][p];
    }
    template <int DIM>
    __device__ void finish(const Fields &f, const Fields &out, const Params &c, int64_t p) {
        wcsph_force_finish<DIM>(f, out, c, p, this->s.v0, this->s.v1, this->s.v2, this->s.dv0, this->s.dv1, this->s.dv2);
    }
};

#undef PF
#undef QF

// The operator menu: name, device functor, fields read, fields written, host bookkeeping after the
// launch.  One list for the dispatch table of pair_ops.cu and for the CPU emulation harness
// (tests/emu/), which checks every functor against the oracle under the same name.
#define SPHMW_OPERATOR_MENU(U, B) \
    U("wcsph.accelerate", U_wcsph_accelerate<true>, SL(S_TYPE, S_RHO_P, S_RHO, S_X0, S_DV0), SL(S_V0, S_DV0), c->dv_zero = true) \
    U("wcsph.move", U_wcsph_move, SL(S_TYPE, S_V0), SL(S_X0), c->cell_list_valid = false) \
    U("wcsph.reset_density", U_wcsph_reset_density, SL(S_TYPE), SL(S_RHO, S_RHO_P),) \
    B("wcsph.compute_density", B_wcsph_density, SL(S_X0, S_M, S_H), SL(S_RHO),) \
    U("wcsph.finalize_density", U_wcsph_finalize_density, SL(S_X0, S_RHO), SL(S_RHO_BG, S_RHO_P),) \
    U("wcsph.update_smoothing", U_wcsph_update_smoothing, SL(S_M, S_RHO), SL(S_H),) \
    U("wcsph.compute_pressure", U_wcsph_compute_pressure, SL(S_X0, S_RHO_P), SL(S_P_BG, S_P_P, S_P),) \
    U("wcsph.find_temperature", U_wcsph_find_temperature, SL(S_P, S_RHO, S_T_BG), SL(S_T, S_T_P),) \
    U("wcsph.find_pot_temp", U_wcsph_find_pot_temp, SL(S_T, S_P, S_X0), SL(S_TH, S_TH_BG, S_TH_P),) \
    B("wcsph.balance_of_momentum", B_wcsph_momentum, SL(S_X0, S_V0, S_H, S_M, S_RHO, S_P_P, S_P), SL(S_DV0), c->dv_zero = false) \
    U("hopkins.reset_pressure", U_hopkins_reset_pressure, SL(S_TYPE), SL(S_P, S_P_P),) \
    B("hopkins.compute_pressure", B_hopkins_pressure, SL(S_X0, S_M, S_H, S_A), SL(S_P),) \
    U("hopkins.finalize_pressure", U_hopkins_finalize_pressure, SL(S_X0), SL(S_P, S_P_BG, S_P_P),) \
    U("hopkins_total.reset_pressure", U_ht_reset_pressure, SL(S_TYPE), SL(S_P),) \
    U("hopkins_total.finalize_pressure", U_ht_finalize_pressure, SL(S_TYPE), SL(S_P),) \
    U("hopkins_total.find_temperature", U_ht_find_temperature, SL(S_P, S_RHO), SL(S_T),) \
    U("hopkins_total.find_pot_temp", U_ht_find_pot_temp, SL(S_T, S_P), SL(S_TH),) \
    U("hopkins_total.reset_density", U_ht_reset_density, SL(S_TYPE), SL(S_RHO),) \
    B("hopkins_total.balance_of_momentum", B_ht_momentum, SL(S_X0, S_V0, S_H, S_M, S_RHO, S_P, S_A), SL(S_DV0), c->dv_zero = false) \
    B("hopkins_full.balance_of_momentum", B_hf_momentum, SL(S_X0, S_V0, S_H, S_M, S_RHO, S_P, S_P_BG, S_A, S_A_BG), SL(S_DV0), c->dv_zero = false) \
    U("hopkins_total.move", U_ht_move, SL(S_V0), SL(S_X0), c->cell_list_valid = false) \
    U("hopkins_total.accelerate", U_ht_accelerate, SL(S_X0, S_DV0), SL(S_V0, S_DV0), c->dv_zero = true) \
    B("dambreak.balance_of_mass", B_dam_mass, SL(S_X0, S_V0, S_RHO), SL(S_DRHO),) \
    U("dambreak.find_pressure", U_dam_find_pressure, SL(S_DRHO), SL(S_RHO, S_DRHO, S_P),) \
    B("dambreak.internal_force", B_dam_force, SL(S_X0, S_V0, S_P, S_RHO, S_TYPE), SL(S_DV0), c->dv_zero = false) \
    U("dambreak.move", U_dam_move, SL(S_TYPE, S_V0), SL(S_X0, S_DV0), (c->cell_list_valid = false, c->dv_zero = true)) \
    U("dambreak.accelerate", U_dam_accelerate, SL(S_TYPE, S_DV0), SL(S_V0),) \
    B("collision.find_rho", B_col_rho<S_RHO>, SL(S_X0), SL(S_RHO),) \
    B("collision.find_rho0", B_col_rho<S_RHO0>, SL(S_X0), SL(S_RHO0),) \
    U("collision.find_pressure", U_col_find_pressure, SL(S_RHO, S_RHO0), SL(S_P),) \
    B("collision.internal_force", B_col_force, SL(S_X0, S_P), SL(S_DV0), c->dv_zero = false) \
    U("collision.reset_a", U_col_reset_a, SL(S_X0), SL(S_DV0), c->dv_zero = true) \
    U("collision.reset_rho", U_col_reset_rho, SL(S_X0), SL(S_RHO),) \
    U("collision.move", U_col_move, SL(S_V0), SL(S_X0), c->cell_list_valid = false) \
    U("collision.accelerate", U_col_accelerate, SL(S_DV0), SL(S_V0),) \
    B("flow.balance_of_mass", B_flow_mass, SL(S_X0, S_V0, S_M), SL(S_DRHO),) \
    B("flow.internal_force", B_flow_force, SL(S_X0, S_V0, S_M, S_P, S_RHO), SL(S_DV0), c->dv_zero = false) \
    U("flow.find_pressure", U_flow_find_pressure, SL(S_DRHO), SL(S_RHO, S_DRHO, S_P),) \
    U("flow.set_density", U_flow_set_density, SL(S_X0), SL(S_RHO),) \
    U("flow.find_pot_temp", U_flow_find_pot_temp, SL(S_P), SL(S_TH),) \
    U("flow.move", U_flow_move, SL(S_TYPE, S_V0), SL(S_X0, S_DV0), (c->cell_list_valid = false, c->dv_zero = true)) \
    U("flow.accelerate", U_flow_accelerate, SL(S_TYPE, S_X0, S_DV0), SL(S_V0),) \
    B("aflow.find_density", B_aflow_density, SL(S_X0, S_M, S_TYPE, S_RHO), SL(S_RHO),) \
    B("aflow.entropy_production", B_aflow_entropy, SL(S_X0, S_V0, S_M, S_TYPE, S_T, S_RHO, S_ENT), SL(S_ENT),) \
    U("aflow.find_s", U_aflow_find_s, SL(S_TYPE, S_ENT, S_RHO, S_M), SL(S_ENT_D),) \
    U("aflow.find_pressure", U_aflow_find_pressure, SL(S_TYPE, S_RHO, S_ENT_D), SL(S_T, S_P),) \
    U("aflow.find_pot_temp", U_aflow_find_pot_temp, SL(S_TYPE, S_P, S_T), SL(S_TH),) \
    U("aflow.move", U_aflow_move, SL(S_TYPE, S_V0), SL(S_X0, S_DV0, S_RHO), (c->cell_list_valid = false, c->dv_zero = true)) \
    U("packing.reset_rho", U_pack_reset_rho, SL(S_TYPE), SL(S_RHO),) \
    B("packing.accumulate_rho", B_pack_rho, SL(S_X0, S_M, S_H, S_TYPE), SL(S_RHO),) \
    B("packing.balance_of_momentum", B_pack_momentum, SL(S_X0, S_M, S_H, S_TYPE, S_RHO), SL(S_DV0), c->dv_zero = false) \
    U("packing.accelerate", U_pack_accelerate, SL(S_TYPE, S_DV0), SL(S_V0, S_DV0), c->dv_zero = true) \
    U("packing.move", U_pack_move, SL(S_TYPE, S_V0), SL(S_X0), c->cell_list_valid = false) \
    /* end of menu */
