// Lattice/CSG input generation on the device for the mountain-wave drivers:
// what make_system() of src/current/wcsph_perturbed_witch.jl:152-170 does on the host with
//   grid     = Grid(dr, :hexagonal | :square | :cubic)          src/grids.jl:50-93,176-196
//   domain   = Rectangle / Box                                  src/geometry.jl:15-43
//   fence    = BoundaryLayer(domain, grid, bc_width)            src/geometry.jl:196-232
//   mountain = Specification(domain, x -> x[2] <= profile(x))   src/geometry.jl:176-187
//   generate_particles!(sys, grid, domain - mountain, FLUID); (fence, WALL); (mountain, ...)
// Every lattice site of the fence's bounding box is classified by one thread with the same
// floating-point expressions as the host generators (grids.py / geometry.py), three exclusive
// scans give each site its index inside its group, and the particles are written in the
// reference's order: group by group (fluid, walls, mountain — SURVEY quirk 12), lattice index i
// outermost inside a group (quirk 13).  The particle constructor (:103-145) is evaluated per
// particle for the fields the step carries (h, x, m, v, rho, rho', type).
#include <math.h>

#include <vector>

#include "sphmw_internal.h"

struct LatticeJob {
    int grid;       // 0 square, 1 hexagonal, 2 cubic
    int mountain;   // 0 none, 1 Witch of Agnesi (2D), 2 bell hill (3D)
    double sx, sy, sz;  // lattice steps (dr, or the hexagonal a, b)
    long long i0, j0, k0, ni, nj, nk;
    double dmin[3], dmax[3];
    double width, h_m, a, U;
    double t_fluid, t_wall, t_mountain;
    double h0, dr;
    int rng;            // ceil(width / dr): largest lattice offset of the boundary layer
    int n_off;          // hexagonal: number of offsets in the ball
    const double *off;  // hexagonal: offsets (x, y) pairs on the device
    // slab clip (multi-GPU): only sites whose cell column lies in [col_lo, col_hi)
    int clip;
    long long col_lo, col_hi, phase0;
    double cell_h;
};

__device__ __forceinline__ bool in_domain(const LatticeJob &J, double x, double y, double z) {
    // geometry.jl:24-30
    return J.dmin[0] <= x && x <= J.dmax[0] && J.dmin[1] <= y && y <= J.dmax[1] && J.dmin[2] <= z &&
           z <= J.dmax[2];
}

__device__ __forceinline__ void site_position(const LatticeJob &J, long long s, double &x, double &y, double &z) {
    long long k = s % J.nk;
    long long j = (s / J.nk) % J.nj;
    long long i = s / (J.nk * J.nj);
    i += J.i0;
    j += J.j0;
    k += J.k0;
    if (J.grid == 1) {
        // grids.jl:85-86 — (i + (j % 2)/2) * a with Julia's truncating remainder
        double shift = (double)(j % 2) / 2;
        x = ((double)i + shift) * J.sx;
        y = (double)j * J.sy;
        z = 0.0;
    } else {
        x = (double)i * J.sx;  // grids.jl:62,190
        y = (double)j * J.sy;
        z = J.grid == 2 ? (double)k * J.sz : 0.0;
    }
}

// smallest |n| (n = 0, -1, 1, -2, 2, ...) with lo <= v + n*dr <= hi, or -1
__device__ __forceinline__ int min_steps(double v, double lo, double hi, double dr, int rng) {
    for (int m = 0; m <= rng + 1; ++m) {
        double a = v + (double)(-m) * dr, b = v + (double)m * dr;
        if ((lo <= a && a <= hi) || (lo <= b && b <= hi)) return m;
    }
    return -1;
}

// geometry.jl:207-217 for an axis-aligned box
__device__ bool in_fence(const LatticeJob &J, double x, double y, double z) {
    if (in_domain(J, x, y, z)) return false;
    if (J.grid == 1) {
        for (int t = 0; t < J.n_off; ++t)
            if (in_domain(J, x + J.off[2 * t], y + J.off[2 * t + 1], z + 0.0)) return true;
        return false;
    }
    // square / cubic: is_inside(x + dx, Box) is separable per axis and the ball test is
    // monotone in each |n| (geometry.py BoundaryLayer._box_fast)
    int n0 = min_steps(x, J.dmin[0], J.dmax[0], J.dr, J.rng);
    int n1 = min_steps(y, J.dmin[1], J.dmax[1], J.dr, J.rng);
    int n2 = J.grid == 2 ? min_steps(z, J.dmin[2], J.dmax[2], J.dr, J.rng) : 0;
    if (n0 < 0 || n1 < 0 || n2 < 0 || n0 > J.rng || n1 > J.rng || n2 > J.rng) return false;
    double d0 = (double)n0 * J.dr, d1 = (double)n1 * J.dr, d2 = (double)n2 * J.dr;
    // Ball: (x-0)^2 + (y-0)^2 + (z-0)^2 <= r^2   (geometry.jl:252-254)
    return (d0 - 0.0) * (d0 - 0.0) + (d1 - 0.0) * (d1 - 0.0) + (d2 - 0.0) * (d2 - 0.0) <= J.width * J.width;
}

__device__ __forceinline__ bool in_mountain(const LatticeJob &J, double x, double y, double z) {
    if (J.mountain == 1) return y <= (J.h_m * (J.a * J.a)) / (x * x + J.a * J.a);  // :158
    if (J.mountain == 2) {
        double t = 1 + (x * x + z * z) / (J.a * J.a);
        return y <= J.h_m / (t * sqrt(t));
    }
    return false;
}

// group of a site: 0 fluid (domain - mountain), 1 wall (fence), 2 mountain, 3 none
__device__ int site_group(const LatticeJob &J, double x, double y, double z) {
    if (J.clip) {
        long long col = (long long)floor(x / J.cell_h) - J.phase0;
        if (col < J.col_lo || col >= J.col_hi) return 3;
    }
    if (in_domain(J, x, y, z)) return in_mountain(J, x, y, z) ? 2 : 0;
    return in_fence(J, x, y, z) ? 1 : 3;
}

__global__ void k_lattice_classify(LatticeJob J, long long nsites, uint32_t *f0, uint32_t *f1, uint32_t *f2) {
    long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= nsites) return;
    double x, y, z;
    site_position(J, s, x, y, z);
    int g = site_group(J, x, y, z);
    f0[s] = g == 0;
    f1[s] = g == 1;
    f2[s] = g == 2;
}

template <int DIM>
__global__ void k_lattice_emit(LatticeJob J, Params c, long long nsites, const uint32_t *r0,
                               const uint32_t *r1, const uint32_t *r2, long long base1, long long base2,
                               long long first, Fields f) {
    long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= nsites) return;
    double x, y, z;
    site_position(J, s, x, y, z);
    int g = site_group(J, x, y, z);
    if (g == 3) return;
    long long p = first + (g == 0 ? r0[s] : g == 1 ? base1 + r1[s] : base2 + r2[s]);
    // Particle(x, v, type), wcsph_perturbed_witch.jl:103-145
    double type = g == 0 ? J.t_fluid : g == 1 ? J.t_wall : J.t_mountain;
    bool wind = g == 0 || (g == 2 && J.t_mountain == J.t_fluid);
    double rho = 0.0 + c.rho0 * exp(-y * c.g / (c.R_mass * c.T_bg));  // :129,139
    double m = rho * J.dr * J.dr;                                      // :143
    if (DIM == 3) m = m * J.dr;
    f.s[S_H][p] = J.h0;
    f.s[S_X0][p] = x;
    f.s[S_X1][p] = y;
    if (DIM == 3) f.s[S_X2][p] = z;
    f.s[S_V0][p] = wind ? J.U : 0.0;
    f.s[S_V1][p] = 0.0;
    if (DIM == 3) f.s[S_V2][p] = 0.0;
    f.s[S_M][p] = m;
    f.s[S_RHO][p] = rho;
    f.s[S_RHO_P][p] = 0.0;
    f.s[S_TYPE][p] = type;
}

static void axis_range(double lo, double hi, double step, long long &i0, long long &n, long long extra_lo) {
    i0 = (long long)floor(lo / step) - extra_lo;  // grids.jl:57-60,80-83
    long long i1 = (long long)ceil(hi / step);
    n = i1 - i0 + 1;
}

extern "C" int sphmw_generate_mountain_wave(sphmw_ctx *c, const sphmw_lattice_setup *su, int64_t *n_out,
                                            int64_t group_counts[3]) {
    if (!c || !su) { sphmw_set_error("null argument"); return SPHMW_E_INVALID; }
    CUDA_TRY(cudaSetDevice(c->device));
    const int dim = c->grid.dim;
    if ((su->grid == 2) != (dim == 3)) { sphmw_set_error("generate: the cubic lattice needs a 3D system and vice versa"); return SPHMW_E_INVALID; }
    if (su->grid < 0 || su->grid > 2 || !(su->dr > 0)) { sphmw_set_error("generate: bad lattice"); return SPHMW_E_INVALID; }
    LatticeJob J{};
    J.grid = su->grid;
    J.mountain = su->mountain;
    J.dr = su->dr;
    if (su->grid == 1) {
        J.sx = pow(4.0 / 3.0, 1.0 / 4.0) * su->dr;  // grids.jl:74
        J.sy = pow(3.0 / 4.0, 1.0 / 4.0) * su->dr;
        J.sz = 1.0;
    } else {
        J.sx = J.sy = J.sz = su->dr;
    }
    for (int a = 0; a < 3; ++a) { J.dmin[a] = su->dom_min[a]; J.dmax[a] = su->dom_max[a]; }
    J.width = su->bc_width;
    J.h_m = su->h_m;
    J.a = su->a;
    J.U = su->U;
    J.t_fluid = su->type_fluid;
    J.t_wall = su->type_wall;
    J.t_mountain = su->type_mountain;
    J.h0 = su->h0;
    J.rng = (int)ceil(su->bc_width / su->dr);
    // bounding box of the fence (geometry.jl:219-232)
    const double w = su->bc_width;
    axis_range(J.dmin[0] - w, J.dmax[0] + w, J.sx, J.i0, J.ni, su->grid == 1 ? 1 : 0);
    axis_range(J.dmin[1] - w, J.dmax[1] + w, J.sy, J.j0, J.nj, 0);
    if (dim == 3) axis_range(J.dmin[2] - w, J.dmax[2] + w, J.sz, J.k0, J.nk, 0);
    else { J.k0 = 0; J.nk = 1; }
    J.clip = c->slab_lo >= 0;
    if (J.clip) {
        // global column range owned by this rank; Grid::phase[0] was shifted by the slab
        J.col_lo = c->slab_lo;
        J.col_hi = c->slab_hi;
        J.phase0 = c->grid.phase[0] - (c->slab_lo - c->grid.ghost);
        J.cell_h = c->grid.h;
        // restrict the enumeration to the lattice planes that can fall into the slab
        double xa = (double)(J.phase0 + J.col_lo) * J.cell_h - 2 * J.sx;
        double xb = (double)(J.phase0 + J.col_hi) * J.cell_h + 2 * J.sx;
        long long ia = (long long)floor(xa / J.sx) - 1, ib = (long long)ceil(xb / J.sx) + 1;
        long long i1 = J.i0 + J.ni - 1;
        long long n0 = std::max(J.i0, ia), n1 = std::min(i1, ib);
        J.i0 = n0;
        J.ni = std::max<long long>(0, n1 - n0 + 1);
    }
    const long long nsites = J.ni * J.nj * J.nk;
    if (nsites <= 0) { if (n_out) *n_out = c->n; return SPHMW_OK; }
    if (nsites >= (long long)0x7FFFFFF0) { sphmw_set_error("generate: too many lattice sites"); return SPHMW_E_CAPACITY; }

    // hexagonal boundary layer: the offsets covering(grid, Ball(0,0,0,width)) (geometry.jl:202)
    std::vector<double> off;
    if (su->grid == 1) {
        long long oi0, oni, oj0, onj;
        axis_range(-w, w, J.sx, oi0, oni, 1);
        axis_range(-w, w, J.sy, oj0, onj, 0);
        for (long long i = oi0; i < oi0 + oni; ++i)
            for (long long j = oj0; j < oj0 + onj; ++j) {
                double x = ((double)i + (double)(j % 2) / 2) * J.sx, y = (double)j * J.sy;
                if ((x - 0.0) * (x - 0.0) + (y - 0.0) * (y - 0.0) + (0.0 - 0.0) * (0.0 - 0.0) <= w * w) {
                    off.push_back(x);
                    off.push_back(y);
                }
            }
        J.n_off = (int)(off.size() / 2);
    }
    double *d_off = nullptr;
    uint32_t *flags = nullptr;
    int rc = [&]() -> int {
        if (!off.empty()) {
            CUDA_TRY(cudaMalloc(&d_off, sizeof(double) * off.size()));
            CUDA_TRY(cudaMemcpyAsync(d_off, off.data(), sizeof(double) * off.size(), cudaMemcpyHostToDevice, c->stream));
        }
        J.off = d_off;
        CUDA_TRY(cudaMalloc(&flags, sizeof(uint32_t) * 3 * (size_t)nsites));
        uint32_t *f0 = flags, *f1 = flags + nsites, *f2 = flags + 2 * nsites;
        {
            TIMED(c, "lattice_classify");
            k_lattice_classify<<<grid_for(nsites, 256), 256, 0, c->stream>>>(J, nsites, f0, f1, f2);
        }
        uint32_t lastf[3], lastr[3];
        for (int g = 0; g < 3; ++g)
            CUDA_TRY(cudaMemcpyAsync(&lastf[g], flags + (size_t)g * nsites + nsites - 1, sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, c->stream));
        for (int g = 0; g < 3; ++g) TRY(sphmw_exclusive_scan_u32(c, flags + (size_t)g * nsites, nsites));
        for (int g = 0; g < 3; ++g)
            CUDA_TRY(cudaMemcpyAsync(&lastr[g], flags + (size_t)g * nsites + nsites - 1, sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        long long cnt[3];
        for (int g = 0; g < 3; ++g) cnt[g] = (long long)lastf[g] + lastr[g];
        const long long total = cnt[0] + cnt[1] + cnt[2];
        if (group_counts)
            for (int g = 0; g < 3; ++g) group_counts[g] = cnt[g];
        const int64_t first = c->n;
        TRY(sphmw_resize(c, first + total));  // iota indices, zeroed fields
        for (int s : {S_H, S_X0, S_X1, S_M, S_V0, S_V1, S_RHO, S_RHO_P, S_TYPE}) TRY(sphmw_ensure_slot(c, s));
        if (dim == 3) { TRY(sphmw_ensure_slot(c, S_X2)); TRY(sphmw_ensure_slot(c, S_V2)); }
        if (total > 0) {
            TIMED(c, "lattice_emit");
            if (dim == 2)
                k_lattice_emit<2><<<grid_for(nsites, 256), 256, 0, c->stream>>>(J, c->prm, nsites, f0, f1, f2, cnt[0],
                                                                               cnt[0] + cnt[1], first, c->cur);
            else
                k_lattice_emit<3><<<grid_for(nsites, 256), 256, 0, c->stream>>>(J, c->prm, nsites, f0, f1, f2, cnt[0],
                                                                               cnt[0] + cnt[1], first, c->cur);
        }
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (n_out) *n_out = c->n;
        return SPHMW_OK;
    }();
    cudaFree(d_off);
    cudaFree(flags);
    return rc;
}
