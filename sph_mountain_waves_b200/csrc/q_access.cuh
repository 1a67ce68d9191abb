// Where a pair closure finds the fields of its neighbour q.  The fused closure bodies
// (wcsph_ops.cuh) are written once, against an accessor Q with get<SLOT>() — the field value, bit
// for bit the SoA entry — and rho_floored(c) = max(rho_q, rho_floor):
//   GlobalQ      the SoA arrays in global memory (cell walk, pair list)
//   RecAQ, RecQ  the packed 32-byte records (SPHMW_FLAG_PACKED_RECORDS, pair_list.cuh)
//   TileQ        the block's shared-memory tile (pair_tile.cuh)
#pragma once
#include "sphmw_internal.h"

// Julia's max(a,b) propagates NaN (Base.max); fmax does not.
__device__ __forceinline__ double jl_max(double a, double b) {
    if (a != a || b != b) return a + b;
    return a < b ? b : a;
}

#define QG(slot) q.template get<slot>()
struct GlobalQ {
    const Fields &f;
    int64_t q;
    template <int SLOT>
    __device__ double get() const { return f.s[SLOT][q]; }
    __device__ double rho_floored(const Params &c) const { return jl_max(f.s[S_RHO][q], c.rho_floor); }
};
// A = {x, y, z, m} only (density closure)
struct RecAQ {
    double qm;
    template <int SLOT>
    __device__ double get() const {
        static_assert(SLOT == S_M, "record A carries the mass only");
        return qm;
    }
};
// A.d = m, B = {vx, vy, vz, h}, C = {P'/rho^2, max(rho, rho_floor), c_s}
struct RecQ {
    double qm;
    const NbRec &B, &C;
    template <int SLOT>
    __device__ double get() const {
        if constexpr (SLOT == S_M) return qm;
        else if constexpr (SLOT == S_V0) return B.a;
        else if constexpr (SLOT == S_V1) return B.b;
        else if constexpr (SLOT == S_V2) return B.c;
        else if constexpr (SLOT == S_H) return B.d;
        else if constexpr (SLOT == S_PR2) return C.a;
        else {
            static_assert(SLOT == S_CS, "field not in the packed records");
            return C.c;
        }
    }
    __device__ double rho_floored(const Params &) const { return C.b; }
};

