// CPU emulation harness of the pair kernels of libsphmw — TEST INFRASTRUCTURE, never shipped.
//
// Compiles the device headers (csrc/pair_list.cuh, wcsph_ops.cuh, kernels_sph.cuh, cell_gather.cuh) with g++
// against tests/emu/cuda_runtime.h and runs the kernels one "thread" at a time:
//   walk        the cell walk of _apply_binary! (src/core.jl:94-112), as k_binary does it
//   list        k_binary_build (integer pre-test) + k_binary_list
//   list_f64    the same with the exact FP64 test in the recording pass
//   records     k_binary_build / k_binary_list with the packed neighbour records
// for the two fused passes of verlet_step! (wcsph_perturbed_witch.jl:316-331) on the particle
// state it is given (positions already advanced).  All variants must agree bit for bit; the
// result of the first one is written out so that the Python test can compare it with the oracle
// (-ffp-contract=off and the same libm: strict arithmetic is expected to match exactly).
//
// With nsteps > 0 in the input header it instead runs whole verlet_step!s (:309-332) — accelerate!,
// move!, cell list, density pass, force pass + kick — cycling through the four variants, and
// writes x, v, rho, h after the last step.
//
// usage: emu_pairs <input.bin> <output.bin>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "cuda_runtime.h"
#include "cell_gather.cuh"
#include "pair_list.cuh"
#include "pair_tile.cuh"
#include "sphmw_internal.h"
#include "wcsph_ops.cuh"

uint3 threadIdx, blockIdx, blockDim, gridDim;
uint32_t nl_queue[(96 + NL_QUEUE_SLACK) * NL_BLOCK];
// the tiled kernels' shared window (pair_tile.cuh) and the sync-point bookkeeping of their emulation
alignas(16) unsigned char emu_tile_smem[256 * 1024];
int emu_phase = 0, emu_sync_seen = 0;
bool emu_block_or = false;

void sphmw_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
}

template <class F>
static void launch(int64_t n, F &&thread_body) {
    blockDim = uint3{NL_BLOCK, 1, 1};
    for (int64_t b = 0; b * NL_BLOCK < n; ++b) {
        blockIdx = uint3{(unsigned)b, 0, 0};
        for (unsigned t = 0; t < NL_BLOCK; ++t) {
            threadIdx = uint3{t, 0, 0};
            thread_body();
        }
    }
}

// a fixed grid whose threads stride over their work themselves (k_binary_list_cols)
template <class F>
static void launch_grid(unsigned nblocks, F &&thread_body) {
    blockDim = uint3{NL_BLOCK, 1, 1};
    gridDim = uint3{nblocks, 1, 1};
    for (unsigned b = 0; b < nblocks; ++b) {
        blockIdx = uint3{b, 0, 0};
        for (unsigned t = 0; t < NL_BLOCK; ++t) {
            threadIdx = uint3{t, 0, 0};
            thread_body();
        }
    }
}

// A tiled kernel has four block-wide sync points (selection vote, table loaded, slot positions
// written, fields copied):
// every "thread" runs up to sync point k in pass k — what precedes a sync point is idempotent —
// and to the end in the last pass.
template <class F>
static void launch_tiled(int64_t n, F &&thread_body) {
    blockDim = uint3{TM_BLOCK, 1, 1};
    for (int64_t b = 0; b * TM_BLOCK < n; ++b) {
        blockIdx = uint3{(unsigned)b, 0, 0};
        emu_block_or = false;
        memset(emu_tile_smem, 0xA5, sizeof(emu_tile_smem));  // stale bytes must never be read
        for (emu_phase = 0; emu_phase < 5; ++emu_phase)
            for (unsigned t = 0; t < TM_BLOCK; ++t) {
                threadIdx = uint3{t, 0, 0};
                emu_sync_seen = 0;
                thread_body();
            }
    }
}

// what k_binary (pair_ops.cu) does for one particle
template <int DIM, class Op>
static void walk_thread(Fields f, Fields out, Params prm, Grid g, const uint32_t *key, const uint32_t *cellx,
                        const uint32_t *cell_start, int64_t n, unsigned long long *pc) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const CellCoord home = cell_of(g, key[p], cellx[p]);
    Op op;
    op.template init<DIM>(f, prm, p);
    const double px = f.s[S_X0][p], py = f.s[S_X1][p], pz = DIM == 3 ? f.s[S_X2][p] : 0.0;
    unsigned acc = 0;
    nl_walk<DIM>(op, f, prm, g, home, p, px, py, pz, cell_start, acc);
    op.template finish<DIM>(f, out, prm, p);
    *pc += acc;
}

struct State {
    int64_t n = 0;
    std::vector<double> cur[NSLOT], alt[NSLOT];
    std::vector<uint32_t> key, cellx, cell_start, idx, xq, list, cnt, list16, tile_tab;
    std::vector<NbRec> rec[3];
    Fields fcur{}, falt{};
    void bind() {
        for (int s = 0; s < NSLOT; ++s) {
            fcur.s[s] = cur[s].empty() ? nullptr : cur[s].data();
            falt.s[s] = alt[s].empty() ? nullptr : alt[s].data();
        }
    }
};

static const int WRITTEN[] = {S_RHO, S_RHO_BG, S_RHO_P, S_H, S_P_BG, S_P_P, S_P, S_PR2, S_CS};

struct Result {
    std::vector<double> fields[9], vnew[3], vold[3];
    std::vector<uint32_t> column;  // cell column of every particle (reference index order)
    unsigned long long pairs_density = 0, pairs_force = 0, overflow = 0;
    long long tiled_blocks = 0, blocks = 0;  // tiles variants: blocks with a staged tile
    unsigned long long list_entries = 0;     // list variants: entries recorded (particles with a list)
    unsigned long long listed_pairs = 0;     // ... and the accepted pairs of those particles
    bool same_as(const Result &o) const {
        for (int k = 0; k < 9; ++k)
            if (memcmp(fields[k].data(), o.fields[k].data(), sizeof(double) * fields[k].size())) return false;
        for (int k = 0; k < 3; ++k)
            if (memcmp(vnew[k].data(), o.vnew[k].data(), sizeof(double) * vnew[k].size())) return false;
        return pairs_density == o.pairs_density && pairs_force == o.pairs_force;
    }
};

enum Variant { WALK, LIST_Q6, LIST_F64, RECORDS, TILES, SLAB_LIST, SLAB_RECORDS, SLAB_TILES };
// the shared-memory tiles stage x-rows of cells and run on the x-chunked cell order (with its 10-bit
// mirror); everything else runs on the zrun order (6-bit mirror)
static bool needs_xchunk(Variant v) { return v == TILES || v == SLAB_TILES; }

template <int DIM, class DensityOp, class ForceOp>
static Result run_variant(State st, const Grid &g, const Params &prm, Variant v, int stride) {
    st.bind();
    const int64_t n = st.n;
    Result r;
    unsigned long long counters[4] = {0, 0, 0, 0};
    ColFilter cf{0, 0, (int)g.lim[0] - 1, 1, 0, 1};
    PairList pl{};
    pl.list = st.list.data();
    pl.cnt = st.cnt.data();
    pl.xq = st.xq.data();
    pl.stride = stride;
    pl.qrows = stride + NL_QUEUE_SLACK;
    pl.overflow = &counters[2];
    pl.recA = st.rec[0].data();
    pl.recB = st.rec[1].data();
    pl.recC = st.rec[2].data();
    pl.list16 = st.list16.data();
    const uint32_t *tab = st.tile_tab.data();
    const uint32_t *key = st.key.data(), *cellx = st.cellx.data(), *cs = st.cell_start.data();
    // density pass (in place), then force pass (new velocity into alt)
    if (v == WALK) {
        launch(n, [&] { walk_thread<DIM, DensityOp>(st.fcur, st.fcur, prm, g, key, cellx, cs, n, &counters[0]); });
        launch(n, [&] { walk_thread<DIM, ForceOp>(st.fcur, st.falt, prm, g, key, cellx, cs, n, &counters[1]); });
    } else if (v == LIST_Q6) {
        launch(n, [&] { k_binary_build<DIM, DensityOp, NL_FILTER_Q6>(st.fcur, st.fcur, prm, g, key, cellx, cs, n, 0, &counters[0], cf, pl); });
        launch(n, [&] { k_binary_list<DIM, ForceOp>(st.fcur, st.falt, prm, g, key, cellx, cs, n, 0, &counters[1], cf, pl); });
    } else if (v == LIST_F64) {
        launch(n, [&] { k_binary_build<DIM, DensityOp, NL_FILTER_F64>(st.fcur, st.fcur, prm, g, key, cellx, cs, n, 0, &counters[0], cf, pl); });
        launch(n, [&] { k_binary_list<DIM, ForceOp>(st.fcur, st.falt, prm, g, key, cellx, cs, n, 0, &counters[1], cf, pl); });
    } else if (v == RECORDS) {
        launch(n, [&] { k_binary_build<DIM, DensityOp, NL_FILTER_Q6, true>(st.fcur, st.fcur, prm, g, key, cellx, cs, n, 0, &counters[0], cf, pl); });
        launch(n, [&] { k_binary_list<DIM, ForceOp, true>(st.fcur, st.falt, prm, g, key, cellx, cs, n, 0, &counters[1], cf, pl); });
    } else if (v == TILES) {
        for (size_t b = 0; b * TM_WORDS < st.tile_tab.size(); ++b) {
            const uint32_t *rec = tab + b * TM_WORDS;
            r.blocks += 1;
            r.tiled_blocks += rec[1] <= (uint32_t)tm_max_pieces(g) && rec[2] <= (uint32_t)TileGeom<DIM>::CAP;
        }
        launch_tiled(n, [&] { k_tile_build<DIM, DensityOp>(st.fcur, st.fcur, prm, g, key, cellx, cs, n, 0, &counters[0], cf, pl, tab); });
        launch_tiled(n, [&] { k_tile_list<DIM, ForceOp>(st.fcur, st.falt, prm, g, key, cellx, cs, n, 0, &counters[1], cf, pl, tab); });
    } else {
        // the column filters of a slab context (pair_ops.cu filter_for_depth): the density pass
        // covers all but the outermost column of each side, the force pass all but the outermost
        // two (split into two launches like the overlapped schedule); skipped particles carry
        // their velocity over.  Compared with the cell walk on the columns that were evaluated.
        const int W = (int)g.lim[0];
        const ColFilter cd{1, 1, W - 2, 1, 0, 1};
        const ColFilter cedge{1, 2, 3, W - 4, W - 3, 1}, cint{1, 4, W - 5, 1, 0, 0};
        // list variants on the zrun order: the interior launch goes first and carries the velocity of
        // everything it skips, then the edge columns run over their particle ranges (three blocks
        // striding, as k_binary_list_cols is launched by the overlapped slab step)
        const ColFilter cint_copy{1, 4, W - 5, 1, 0, 1}, cedge_sparse{1, 2, 3, W - 4, W - 3, 0, 1};
        if (v == SLAB_LIST) {
            launch(n, [&] { k_binary_build<DIM, DensityOp, NL_FILTER_Q6>(st.fcur, st.fcur, prm, g, key, cellx, cs, n, 0, &counters[0], cd, pl); });
            launch(n, [&] { k_binary_list<DIM, ForceOp>(st.fcur, st.falt, prm, g, key, cellx, cs, n, 0, &counters[3], cint_copy, pl); });
            launch_grid(3, [&] { k_binary_list_cols<DIM, ForceOp>(st.fcur, st.falt, prm, g, key, cellx, cs, n, 0, &counters[1], cedge_sparse, pl); });
        } else if (v == SLAB_TILES) {
            launch_tiled(n, [&] { k_tile_build<DIM, DensityOp>(st.fcur, st.fcur, prm, g, key, cellx, cs, n, 0, &counters[0], cd, pl, tab); });
            launch_tiled(n, [&] { k_tile_list<DIM, ForceOp>(st.fcur, st.falt, prm, g, key, cellx, cs, n, 0, &counters[1], cedge, pl, tab); });
            launch_tiled(n, [&] { k_tile_list<DIM, ForceOp>(st.fcur, st.falt, prm, g, key, cellx, cs, n, 0, &counters[3], cint, pl, tab); });
        } else {
            launch(n, [&] { k_binary_build<DIM, DensityOp, NL_FILTER_Q6, true>(st.fcur, st.fcur, prm, g, key, cellx, cs, n, 0, &counters[0], cd, pl); });
            launch(n, [&] { k_binary_list<DIM, ForceOp, true>(st.fcur, st.falt, prm, g, key, cellx, cs, n, 0, &counters[3], cint_copy, pl); });
            launch_grid(3, [&] { k_binary_list_cols<DIM, ForceOp, true>(st.fcur, st.falt, prm, g, key, cellx, cs, n, 0, &counters[1], cedge_sparse, pl); });
        }
    }
    r.pairs_density = counters[0];
    r.pairs_force = counters[1];
    if (v == LIST_Q6 || v == LIST_F64 || v == RECORDS)
        for (int64_t p = 0; p < n; ++p)
            if (st.cnt[p] != NL_NONE) r.list_entries += st.cnt[p];
    r.overflow = counters[2];
    // back to reference index order
    for (int k = 0; k < 9; ++k) {
        r.fields[k].resize(n);
        for (int64_t p = 0; p < n; ++p) r.fields[k][st.idx[p]] = st.cur[WRITTEN[k]][p];
    }
    for (int k = 0; k < 3; ++k) {
        r.vnew[k].assign(n, 0.0);
        r.vold[k].assign(n, 0.0);
        if (k < DIM)
            for (int64_t p = 0; p < n; ++p) {
                r.vnew[k][st.idx[p]] = st.alt[S_V0 + k][p];
                r.vold[k][st.idx[p]] = st.cur[S_V0 + k][p];
            }
    }
    r.column.resize(n);
    for (int64_t p = 0; p < n; ++p) r.column[st.idx[p]] = st.cellx[p];
    return r;
}

static void read_exact(FILE *fp, void *dst, size_t bytes) {
    if (bytes && fread(dst, 1, bytes, fp) != bytes) {
        fprintf(stderr, "emu_pairs: truncated input\n");
        exit(2);
    }
}

int main(int argc, char **argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: emu_pairs <input.bin> <output.bin>\n");
        return 2;
    }
    FILE *fp = fopen(argv[1], "rb");
    if (!fp) {
        perror(argv[1]);
        return 2;
    }
    int32_t head[6];  // fast, stride, cx_shift (-1: library default), nparams, nsteps, reserved
    int64_t n;
    double box[7];  // min[3], max[3], h
    read_exact(fp, head, sizeof(head));
    read_exact(fp, &n, sizeof(n));
    read_exact(fp, box, sizeof(box));
    const int fast = head[0], stride = head[1], nparams = head[3], nsteps = head[4];
    if (head[2] >= 0) setenv("SPHMW_CX_SHIFT", std::to_string(head[2]).c_str(), 1);  // chunk width of the x-chunked order
    Params prm;
    memset(&prm, 0, sizeof(prm));
    struct Named {
        const char *name;
        double Params::*field;
    };
    static const Named TABLE[] = {{"dt", &Params::dt}, {"g", &Params::g}, {"c", &Params::c}, {"gamma", &Params::gamma},
                                  {"alpha", &Params::alpha}, {"beta", &Params::beta}, {"eps", &Params::eps},
                                  {"eta", &Params::eta}, {"rho0", &Params::rho0}, {"R_mass", &Params::R_mass},
                                  {"R_gas", &Params::R_gas}, {"T_bg", &Params::T_bg}, {"rho_floor", &Params::rho_floor},
                                  {"P_floor", &Params::P_floor}, {"z_t", &Params::z_t}, {"z_b", &Params::z_b},
                                  {"gamma_r", &Params::gamma_r}, {"fluid", &Params::fluid}};
    for (int k = 0; k < nparams; ++k) {
        char name[16];
        double value;
        read_exact(fp, name, 16);
        read_exact(fp, &value, sizeof(value));
        name[15] = 0;
        for (const Named &t : TABLE)
            if (!strcmp(t.name, name)) prm.*(t.field) = value;
    }
    sphmw_derive_params(prm);
    Grid g;  // zrun cell order (the library default); gx: the same grid in the x-chunked order
    memset(&g, 0, sizeof(g));
    int64_t global_cols = 0;
    if (sphmw_grid_setup(g, box, box + 3, box[6], -1, -1, &global_cols) != SPHMW_OK) return 3;
    sphmw_grid_set_order(g, true);
    Grid gx = g;
    sphmw_grid_set_order(gx, false);
    const int dim = g.dim;

    // fields in reference index order, component-major
    const int IN_SLOTS[] = {S_X0, S_X1, S_X2, S_V0, S_V1, S_V2, S_M, S_H, S_RHO, S_RHO_P, S_TYPE};
    std::vector<double> in[11];
    for (int k = 0; k < 11; ++k) {
        in[k].resize(n);
        read_exact(fp, in[k].data(), sizeof(double) * n);
    }
    fclose(fp);

    // ---- cell list: keys (structs.jl:97-106), order (cell ascending, index descending) --------
    auto build_state = [&](State &st, const Grid &g) -> int {
    std::vector<uint32_t> pkey(n), col(n), order(n);
    for (int64_t i = 0; i < n; ++i) {
        const long long ci = (long long)floor(in[0][i] / g.h) - g.phase[0];
        const long long cj = (long long)floor(in[1][i] / g.h) - g.phase[1];
        const long long ck = dim == 3 ? (long long)floor(in[2][i] / g.h) - g.phase[2] : 0;
        if (ci < 0 || ci >= g.lim[0] || cj < 0 || cj >= g.lim[1] || ck < 0 || ck >= g.lim[2]) {
            fprintf(stderr, "emu_pairs: particle %lld is outside the box (not supported here)\n", (long long)i);
            return 3;
        }
        pkey[i] = pkey_ijk(g, (int)ci, (int)cj, (int)ck);
        col[i] = (uint32_t)ci;
    }
    std::iota(order.begin(), order.end(), 0u);
    std::sort(order.begin(), order.end(),
              [&](uint32_t a, uint32_t b) { return pkey[a] != pkey[b] ? pkey[a] < pkey[b] : a > b; });
    st.n = n;
    st.key.resize(n);
    st.cellx.resize(n);
    st.idx.resize(n);
    st.xq.assign(n + 4, 0u);
    st.cell_start.assign(g.pkey_max + 2, 0u);
    for (int s = 0; s < NSLOT; ++s) {
        st.cur[s].assign(n + 4, 0.0);  // + 4: bulk copies fetch whole groups of 4 particles
        st.alt[s].assign(n + 4, 0.0);
    }
    for (int k = 0; k < 3; ++k) st.rec[k].assign(n, NbRec{0, 0, 0, 0});
    // the permutation itself is the library's k_gather (csrc/cell_gather.cuh): fields, indices,
    // keys, the pre-test mirror and neighbour record A
    for (int64_t i = 0; i < n; ++i) st.cell_start[pkey[i] + 1] += 1;
    {
        std::vector<uint32_t> ident(n), tag_in(n, 0u), tag_out(n), pos_of_idx(n);
        std::iota(ident.begin(), ident.end(), 0u);
        GatherList gl;
        memset(&gl, 0, sizeof(gl));
        gl.count = 0;
        for (int a = 0; a < 3; ++a) gl.xpos[a] = -1;
        gl.mpos = -1;
        for (int k = 0; k < 11; ++k) {
            const int slot = IN_SLOTS[k];
            if (dim == 2 && (slot == S_X2 || slot == S_V2)) continue;  // 2D systems keep no third component
            gl.from[gl.count] = in[k].data();
            gl.to[gl.count] = st.cur[slot].data();
            if (slot >= S_X0 && slot <= S_X2) gl.xpos[slot - S_X0] = gl.count;
            if (slot == S_M) gl.mpos = gl.count;
            ++gl.count;
        }
        gl.h = g.h;
        gl.q6 = g.zrun;
        gl.dim = dim;
        gl.run_phase = g.phase[dim == 3 ? 2 : 1];
        gl.xq = st.xq.data();
        gl.recA = st.rec[0].data();
        launch(n, [&] {
            k_gather(gl, order.data(), ident.data(), st.idx.data(), pos_of_idx.data(), pkey.data(), st.key.data(),
                     tag_in.data(), tag_out.data(), col.data(), st.cellx.data(), n);
        });
        for (int64_t p = 0; p < n; ++p)
            if (pos_of_idx[st.idx[p]] != (uint32_t)p) {
                fprintf(stderr, "emu_pairs: k_gather left an inconsistent inverse map\n");
                return 3;
            }
    }
    for (long long c = 0; c <= g.pkey_max; ++c) st.cell_start[c + 1] += st.cell_start[c];
    const size_t warps = (size_t)((n + 31) / 32);
    st.list.assign(warps * (size_t)stride * 32, 0u);
    st.cnt.assign(n, 0u);
    st.list16.assign(warps * (size_t)((stride + 1) / 2) * 32, 0u);
    // the tile map of this cell list (csrc/tile_map.cuh), one record per block
    const int64_t nblocks = (n + TM_BLOCK - 1) / TM_BLOCK;
    st.tile_tab.assign((size_t)nblocks * TM_WORDS, 0u);
    for (int64_t b = 0; b < nblocks && !g.zrun; ++b)  // tiles: x-chunked order only
        tm_build_record(g, st.key.data(), st.cellx.data(), st.cell_start.data(), n, b, st.tile_tab.data() + (size_t)b * TM_WORDS);
    return 0;
    };
    State st, stx;
    if (build_state(st, g) || build_state(stx, gx)) return 3;

    // ---- the variants --------------------------------------------------------------------------
    auto run_on = [&](Variant v, const State &s, const Grid &gg) -> Result {
        if (dim == 2)
            return fast ? run_variant<2, B_wcsph_density_fast, B_wcsph_momentum_fast>(s, gg, prm, v, stride)
                        : run_variant<2, B_wcsph_density_fused, B_wcsph_momentum_fused>(s, gg, prm, v, stride);
        return fast ? run_variant<3, B_wcsph_density_fast, B_wcsph_momentum_fast>(s, gg, prm, v, stride)
                    : run_variant<3, B_wcsph_density_fused, B_wcsph_momentum_fused>(s, gg, prm, v, stride);
    };
    auto run = [&](Variant v) -> Result { return needs_xchunk(v) ? run_on(v, stx, gx) : run_on(v, st, g); };
    if (nsteps > 0) {
        // master copy stays in reference index order; kick and drift are per-particle
        Fields master{};
        for (int k = 0; k < 11; ++k) master.s[IN_SLOTS[k]] = in[k].data();
        unsigned long long pairs = 0;
        for (int step = 0; step < nsteps; ++step) {
            launch(n, [&] {
                const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
                if (p >= n) return;
                if (dim == 2) {
                    U_wcsph_accelerate<false>::apply<2>(master, prm, p);
                    U_wcsph_move::apply<2>(master, prm, p);
                } else {
                    U_wcsph_accelerate<false>::apply<3>(master, prm, p);
                    U_wcsph_move::apply<3>(master, prm, p);
                }
            });
            st = State();
            stx = State();
            if (build_state(st, g) || build_state(stx, gx)) return 3;
            const Result r = run((Variant)(step % 5));
            pairs = r.pairs_force;
            for (int64_t i = 0; i < n; ++i) {
                in[8][i] = r.fields[0][i];   // rho
                in[9][i] = r.fields[2][i];   // rho'
                in[7][i] = r.fields[3][i];   // h
                for (int a = 0; a < 3; ++a) in[3 + a][i] = r.vnew[a][i];
            }
        }
        FILE *out = fopen(argv[2], "wb");
        if (!out) {
            perror(argv[2]);
            return 2;
        }
        const int64_t meta[6] = {n, dim, (int64_t)pairs, (int64_t)pairs, 0, gx.cx_shift};
        fwrite(meta, sizeof(meta), 1, out);
        for (int k : {8, 9, 9, 7, 9, 9, 9}) fwrite(in[k].data(), sizeof(double), n, out);  // rho, -, -, h, -, -, -
        for (int k = 0; k < 3; ++k) fwrite(in[3 + k].data(), sizeof(double), n, out);       // v
        for (int k = 0; k < 3; ++k) fwrite(in[k].data(), sizeof(double), n, out);           // x
        fclose(out);
        printf("emu_pairs: %d steps, n=%lld dim=%d pairs=%llu\n", nsteps, (long long)n, dim, pairs);
        return 0;
    }
    const Result base = run(WALK);
    if (!run_on(WALK, stx, gx).same_as(base)) {
        fprintf(stderr, "emu_pairs: the cell walk depends on the physical cell order\n");
        return 1;
    }
    {
        // the x-chunked order records with the FP64 test (no 6-bit mirror there)
        const Result r = run_on(LIST_F64, stx, gx);
        if (!r.same_as(base)) {
            fprintf(stderr, "emu_pairs: list_f64 on the x-chunked order differs from the cell walk\n");
            return 1;
        }
    }
    static const char *NAMES[] = {"walk", "list", "list_f64", "records", "tiles", "slab_list", "slab_records", "slab_tiles"};
    unsigned long long overflow = 0;
    long long tiled_blocks = 0, blocks = 0;
    for (Variant v : {LIST_Q6, LIST_F64, RECORDS, TILES}) {
        const Result r = run(v);
        if (v == TILES) tiled_blocks = r.tiled_blocks, blocks = r.blocks;
        if (!r.same_as(base)) {
            fprintf(stderr, "emu_pairs: variant '%s' differs from the cell walk (pairs %llu/%llu vs %llu/%llu)\n",
                    NAMES[v], r.pairs_density, r.pairs_force, base.pairs_density, base.pairs_force);
            return 1;
        }
        overflow = std::max(overflow, r.overflow);
        // the list holds the accepted neighbours and nothing else (no self entry, no false accepts)
        if (v != TILES && r.overflow == 0 && r.list_entries != base.pairs_density) {
            fprintf(stderr, "emu_pairs: variant '%s' recorded %llu list entries for %llu accepted pairs\n", NAMES[v],
                    r.list_entries, base.pairs_density);
            return 1;
        }
    }
    if (g.lim[0] >= 10) {
        const int W = (int)g.lim[0];
        for (Variant v : {SLAB_LIST, SLAB_RECORDS, SLAB_TILES}) {
            const Result r = run(v);
            for (int64_t i = 0; i < n; ++i) {
                const int c = (int)r.column[i];
                bool ok = true;
                if (c >= 1 && c <= W - 2)
                    for (int k = 0; k < 9; ++k) ok = ok && !memcmp(&r.fields[k][i], &base.fields[k][i], sizeof(double));
                for (int k = 0; k < 3; ++k) {
                    const double want = (c >= 2 && c <= W - 3) ? base.vnew[k][i] : r.vold[k][i];
                    // interior launch has copy = 0: columns 0, 1, W-2, W-1 are copied by the edge launch
                    ok = ok && !memcmp(&r.vnew[k][i], &want, sizeof(double));
                }
                if (!ok) {
                    fprintf(stderr, "emu_pairs: slab-filtered variant %d differs at particle %lld (column %d of %d)\n",
                            (int)v, (long long)i, c, W);
                    return 1;
                }
            }
        }
    }
    FILE *out = fopen(argv[2], "wb");
    if (!out) {
        perror(argv[2]);
        return 2;
    }
    const int64_t meta[6] = {n, dim, (int64_t)base.pairs_density, (int64_t)base.pairs_force, (int64_t)overflow,
                             gx.cx_shift};
    fwrite(meta, sizeof(meta), 1, out);
    for (int k = 0; k < 7; ++k) fwrite(base.fields[k].data(), sizeof(double), n, out);  // rho .. P
    for (int k = 0; k < 3; ++k) fwrite(base.vnew[k].data(), sizeof(double), n, out);
    fclose(out);
    printf("emu_pairs: n=%lld dim=%d pairs=%llu overflow=%llu cx_shift=%d tiled=%lld/%lld: walk == list == list_f64 =="
           " records == tiles (== slab-filtered launches on their columns)\n",
           (long long)n, dim, base.pairs_force, overflow, gx.cx_shift, tiled_blocks, blocks);
    return 0;
}
