// Neighbour-grid setup of a ParticleSystem — src/structs.jl:63-82 (bounding box, key_phase,
// key_lim, key_max, key_diff with di outermost) plus what this library adds to it: the slab
// window, the x-chunked physical cell order and the exact r^2 threshold of the
// cut-off test.  Host-only and free of CUDA calls: sphmw_create uses it, and so does the
// emulation harness of the CPU test suite (tests/emu/).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "sphmw_internal.h"

// derived driver constants (same names as sphmw_set_param)
void sphmw_derive_params(Params &p) {
    // damping_structure, wcsph_perturbed_witch.jl:245-251: a constant vector.
    // Evaluated once on the host in the written order.
    p.sponge_z0 = p.z_t - p.z_b;
    if (p.z_b != 0.0) {
        double sn = sin(M_PI / 2 * (1 - (p.z_t - p.z_b) / p.z_b));
        p.sponge_y = -p.gamma_r * (sn * sn);
    } else {
        p.sponge_y = 0.0;
    }
}

// Physical cell order (sphmw_internal.h, Grid::zrun).  zrun: x outermost, the innermost axis of the
// reference's key_diff loops fastest — the default.  Otherwise x-chunked rows: a pass needs three
// x-y planes of cells at a time; chunks are sized so that three chunk-planes (~6 particles x ~100 B
// per cell) stay near 40 MB, well inside the 126 MB L2: about 22000 / Ly columns (measured on the
// 64 M case, profiles/r01_tuning.md: 256 columns best of 32..2048).  Grids up to 1.5x that wide, and
// 2D grids, stay one chunk (plain x-fastest rows); wider ones are cut into equal-looking
// power-of-two chunks.
void sphmw_grid_set_order(Grid &g, bool zrun) {
    g.zrun = zrun ? 1 : 0;
    g.rows = g.lim[1] * g.lim[2];
    g.cx_shift = 0;
    if (zrun) {
        g.pkey_max = g.lim[0] * g.rows;
        return;
    }
    const long long lx = g.lim[0];
    const long long want = g.dim == 3 ? std::max<long long>(32, 22000 / std::max<long long>(1, g.lim[1])) : lx;
    long long cols = lx;
    if (2 * lx > 3 * want) {
        const long long nchunks = (lx + want - 1) / want;
        cols = (lx + nchunks - 1) / nchunks;
    }
    while ((1LL << g.cx_shift) < cols) ++g.cx_shift;
    if (getenv("SPHMW_CX_SHIFT")) g.cx_shift = atoi(getenv("SPHMW_CX_SHIFT"));
    const long long cx = 1LL << g.cx_shift;
    const long long nchunks = (g.lim[0] + cx - 1) / cx;
    g.pkey_max = nchunks * cx * g.rows;
}

int sphmw_grid_setup(Grid &g, const double box_min[3], const double box_max[3], double h, int64_t slab_lo,
                     int64_t slab_hi, int64_t *global_cols, int ghost) {
    g.h = h;
    g.ghost = slab_lo >= 0 ? ghost : 0;
    for (int a = 0; a < 3; ++a) {
        g.box[a] = box_min[a];
        g.box[3 + a] = box_max[a];
    }
    // structs.jl:66-68
    g.key_max = 1;
    for (int a = 0; a < 3; ++a) {
        g.phase[a] = (long long)floor(g.box[a] / g.h);
        g.lim[a] = (long long)floor(g.box[3 + a] / g.h) - g.phase[a] + 1;
        if (g.lim[a] <= 0) {
            sphmw_set_error("empty bounding box along axis %d", a);
            return SPHMW_E_INVALID;
        }
        g.key_max *= g.lim[a];
    }
    if (global_cols) *global_cols = 0;
    if (slab_lo >= 0) {
        // local grid = owned columns + `ghost` ghost columns each side; keys are local
        if (!(slab_hi > slab_lo) || slab_hi > g.lim[0]) {
            sphmw_set_error("invalid slab [%lld,%lld) for %lld columns", (long long)slab_lo, (long long)slab_hi,
                            g.lim[0]);
            return SPHMW_E_INVALID;
        }
        if (slab_hi - slab_lo < 2 * ghost) {
            sphmw_set_error("a slab must own at least %d cell columns", 2 * ghost);
            return SPHMW_E_INVALID;
        }
        if (global_cols) *global_cols = g.lim[0];
        long long width = (slab_hi - slab_lo) + 2 * ghost;
        g.phase[0] += slab_lo - ghost;
        g.key_max = g.key_max / g.lim[0] * width;
        g.lim[0] = width;
    }
    if (g.key_max >= (long long)0xFFFFFFF0u) {
        sphmw_set_error("too many cells (%lld)", g.key_max);
        return SPHMW_E_INVALID;
    }
    // structs.jl:70-82 — di outermost
    g.ndiff = 0;
    if (g.lim[2] == 1) {
        g.dim = 2;
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj) {
                g.nb_di[g.ndiff] = di;
                g.nb_drest[g.ndiff] = dj;
                g.nb_dj[g.ndiff] = dj;
                g.nb_dk[g.ndiff] = 0;
                g.key_diff[g.ndiff++] = (int)(di + g.lim[0] * dj);
            }
    } else {
        g.dim = 3;
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj)
                for (int dk = -1; dk <= 1; ++dk) {
                    g.nb_di[g.ndiff] = di;
                    g.nb_dj[g.ndiff] = dj;
                    g.nb_dk[g.ndiff] = dk;
                    g.nb_drest[g.ndiff] = (int)(dj + g.lim[1] * dk);
                    g.key_diff[g.ndiff++] = (int)(di + g.lim[0] * (dj + g.lim[1] * dk));
                }
    }
    sphmw_grid_set_order(g, !(getenv("SPHMW_CELL_ORDER") && !strcmp(getenv("SPHMW_CELL_ORDER"), "xchunk")));
    if (g.pkey_max >= (long long)0x7FFFFFF0) {
        sphmw_set_error("too many cells (%lld)", g.pkey_max);
        return SPHMW_E_INVALID;
    }
    {
        // exact threshold for the cut-off test (sqrt is monotone and correctly rounded on
        // host and device alike)
        double t = g.h * g.h;
        while (sqrt(t) > g.h) t = nextafter(t, 0.0);
        while (sqrt(nextafter(t, INFINITY)) <= g.h) t = nextafter(t, INFINITY);
        g.r2_max = t;
    }
    return SPHMW_OK;
}
