// Device smoothing kernels — src/kernels.jl (all normalised to support radius h).
// The reference marks them @fastmath; integer powers lower to repeated squaring.
// Written with plain operators: this translation unit is compiled with
// -fmad=false, so no multiply-add is contracted and the values are bit-identical
// to an IEEE left-to-right evaluation.
#pragma once

__device__ __forceinline__ double sph_pow2(double a) { return a * a; }
__device__ __forceinline__ double sph_pow3(double a) { return a * a * a; }
__device__ __forceinline__ double sph_pow4(double a) { double b = a * a; return b * b; }
__device__ __forceinline__ double sph_pow5(double a) { double b = a * a; return b * b * a; }
__device__ __forceinline__ double sph_pos(double x) { return x > 0.0 ? x : 0.0; }  // kernels.jl:3-5

// kernels.jl:108-115
__device__ __forceinline__ double wendland2(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return 2.228169203286535 * sph_pow4(1.0 - x) * (1.0 + 4.0 * x) / sph_pow2(h);
}
// kernels.jl:124-131
__device__ __forceinline__ double Dwendland2(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -44.563384065730695 * x * sph_pow3(1.0 - x) / sph_pow3(h);
}
// kernels.jl:140-147
__device__ __forceinline__ double rDwendland2(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -44.563384065730695 * sph_pow3(1.0 - x) / sph_pow4(h);
}
// kernels.jl:156-163
__device__ __forceinline__ double wendland3(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return 3.3422538049298023 * sph_pow4(1.0 - x) * (1.0 + 4.0 * x) / sph_pow3(h);
}
// kernels.jl:172-179
__device__ __forceinline__ double Dwendland3(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -66.84507609859604 * x * sph_pow3(1.0 - x) / sph_pow4(h);
}
// kernels.jl:188-195
__device__ __forceinline__ double rDwendland3(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -66.84507609859604 * sph_pow3(1.0 - x) / sph_pow5(h);
}
// kernels.jl:197-204
__device__ __forceinline__ double DDwendland3(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -66.84507609859604 * ((1.0 - 4.0 * x) * sph_pow2(1.0 - x)) / sph_pow5(h);
}
// kernels.jl:206-212
__device__ __forceinline__ double wendland1(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return 1.5 * sph_pow4(1.0 - x) * (1.0 + 4.0 * x) / h;
}
// kernels.jl:214-220
__device__ __forceinline__ double Dwendland1(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -30.0 * x * sph_pow3(1.0 - x) / sph_pow2(h);
}
// kernels.jl:222-228
__device__ __forceinline__ double rDwendland1(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -30.0 * sph_pow3(1.0 - x) / sph_pow3(h);
}
// kernels.jl:14-25
__device__ __forceinline__ double spline23(double h, double r) {
    double x = r / h;
    if (x < 0.5)
        return 1.8189136353359467 * (1.0 - 6.0 * sph_pow2(x) + 6.0 * sph_pow3(x)) / sph_pow2(h);
    else if (x < 1.0)
        return 3.6378272706718935 * sph_pow3(1.0 - x) / sph_pow2(h);
    return 0.0;
}
// kernels.jl:34-43
__device__ __forceinline__ double Dspline23(double h, double r) {
    double x = r / h;
    if (x < 0.5) return -10.91348181201568 * (2.0 * x - 3.0 * sph_pow2(x)) / sph_pow3(h);
    else if (x < 1.0) return -10.91348181201568 * sph_pow2(1.0 - x) / sph_pow3(h);
    return 0.0;
}
// kernels.jl:52-61
__device__ __forceinline__ double rDspline23(double h, double r) {
    double x = r / h;
    if (x < 0.5) return -10.91348181201568 * (2.0 - 3.0 * x) / sph_pow4(h);
    else if (x < 1.0) return -10.91348181201568 * sph_pow2(1.0 - x) / (x * sph_pow4(h));
    return 0.0;
}
// kernels.jl:70-73
__device__ __forceinline__ double spline24(double h, double r) {
    double x = r / h;
    return 6.222175110452539 *
           (sph_pow4(sph_pos(1.0 - x)) - 5 * sph_pow4(sph_pos(0.6 - x)) +
            10 * sph_pow4(sph_pos(0.2 - x))) /
           sph_pow2(h);
}
// kernels.jl:82-85
__device__ __forceinline__ double Dspline24(double h, double r) {
    double x = r / h;
    return -24.888700441810155 *
           (sph_pow3(sph_pos(1.0 - x)) - 5 * sph_pow3(sph_pos(0.6 - x)) +
            10 * sph_pow3(sph_pos(0.2 - x))) /
           sph_pow3(h);
}
// kernels.jl:94-100
__device__ __forceinline__ double rDspline24(double h, double r) {
    double x = r / h;
    if (x > 0.2)
        return -24.888700441810155 *
               (sph_pow3(sph_pos(1.0 - x)) - 5 * sph_pow3(sph_pos(0.6 - x))) / (x * sph_pow4(h));
    return -24.888700441810155 * (1.2 - 6.0 * sph_pow2(x)) / sph_pow4(h);
}

__device__ __forceinline__ double sph_kernel_by_id(int which, double h, double r) {
    switch (which) {
        case 0: return wendland1(h, r);
        case 1: return Dwendland1(h, r);
        case 2: return rDwendland1(h, r);
        case 3: return wendland2(h, r);
        case 4: return Dwendland2(h, r);
        case 5: return rDwendland2(h, r);
        case 6: return wendland3(h, r);
        case 7: return Dwendland3(h, r);
        case 8: return rDwendland3(h, r);
        case 9: return DDwendland3(h, r);
        case 10: return spline23(h, r);
        case 11: return Dspline23(h, r);
        case 12: return rDspline23(h, r);
        case 13: return spline24(h, r);
        case 14: return Dspline24(h, r);
        default: return rDspline24(h, r);
    }
}

// dimension-dispatched kernels of the mountain-wave drivers: the 2D drivers call
// wendland2/rDwendland2 (wcsph_perturbed_witch.jl:227,267); the 3D extrusion
//  uses wendland3/rDwendland3.
template <int DIM>
__device__ __forceinline__ double sph_W(double h, double r) {
    return DIM == 2 ? wendland2(h, r) : wendland3(h, r);
}
template <int DIM>
__device__ __forceinline__ double sph_rDW(double h, double r) {
    return DIM == 2 ? rDwendland2(h, r) : rDwendland3(h, r);
}
