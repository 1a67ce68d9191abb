// CPU emulation of the whole operator menu of libsphmw — TEST INFRASTRUCTURE, never shipped.
//
// Compiles csrc/ops_menu.cuh (every device functor + SPHMW_OPERATOR_MENU, the very list the
// dispatch table of pair_ops.cu is built from), pair_list.cuh and cell_gather.cuh with g++ behind
// tests/emu/cuda_runtime.h and applies a sequence of operators BY NAME, exactly as
// sphmw_apply / sphmw_create_cell_list would: unary operators per particle, binary operators
// through the cell walk, the recording kernel or the replaying kernel (cycling, so that every
// kernel sees every closure).  The Python test runs the same sequence in the oracle and expects
// the same bits (-ffp-contract=off, same libm).
//
// usage: emu_ops <input.bin> <output.bin>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "cuda_runtime.h"
#include "cell_gather.cuh"
#include "ops_menu.cuh"
#include "pair_list.cuh"
#include "sphmw_internal.h"

uint3 threadIdx, blockIdx, blockDim, gridDim;
uint32_t nl_queue[(96 + NL_QUEUE_SLACK) * NL_BLOCK];

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

static void read_exact(FILE *fp, void *dst, size_t bytes) {
    if (bytes && fread(dst, 1, bytes, fp) != bytes) {
        fprintf(stderr, "emu_ops: truncated input\n");
        exit(2);
    }
}

struct System {
    int64_t n = 0;
    int dim = 3, stride = 40;
    Grid g;
    Params prm;
    std::vector<double> master[NSLOT];  // reference index order
    std::vector<double> cur[NSLOT];     // physical (cell-sorted) order
    std::vector<uint32_t> key, cellx, cell_start, idx, xq, list, cnt;
    std::vector<NbRec> recA;
    Fields f{};
    bool sorted = false;
    int generation = 0;       // cell lists built so far
    bool list_built = false;  // a pair list exists for this generation
    unsigned long long pairs = 0, overflow = 0;

    void unsort() {
        if (!sorted) return;
        for (int s = 0; s < NSLOT; ++s)
            for (int64_t p = 0; p < n; ++p) master[s][idx[p]] = cur[s][p];
    }

    int create_cell_list() {
        unsort();
        std::vector<uint32_t> pkey(n), col(n), order(n);
        for (int64_t i = 0; i < n; ++i) {
            const double x = master[S_X0][i], y = master[S_X1][i], z = dim == 3 ? master[S_X2][i] : 0.0;
            const long long ci = (long long)floor(x / g.h) - g.phase[0], cj = (long long)floor(y / g.h) - g.phase[1];
            const long long ck = dim == 3 ? (long long)floor(z / g.h) - g.phase[2] : 0;
            if (ci < 0 || ci >= g.lim[0] || cj < 0 || cj >= g.lim[1] || ck < 0 || ck >= g.lim[2]) {
                fprintf(stderr, "emu_ops: particle %lld is outside the box (removal is not emulated)\n", (long long)i);
                return 3;
            }
            pkey[i] = pkey_ijk(g, (int)ci, (int)cj, (int)ck);
            col[i] = (uint32_t)ci;
        }
        std::iota(order.begin(), order.end(), 0u);
        std::sort(order.begin(), order.end(),
                  [&](uint32_t a, uint32_t b) { return pkey[a] != pkey[b] ? pkey[a] < pkey[b] : a > b; });
        key.assign(n, 0u);
        cellx.assign(n, 0u);
        idx.assign(n, 0u);
        xq.assign(n + 4, 0u);
        recA.assign(n, NbRec{0, 0, 0, 0});
        cell_start.assign(g.pkey_max + 2, 0u);
        for (int64_t i = 0; i < n; ++i) cell_start[pkey[i] + 1] += 1;
        for (long long c = 0; c <= g.pkey_max; ++c) cell_start[c + 1] += cell_start[c];
        std::vector<uint32_t> ident(n), tag_in(n, 0u), tag_out(n), pos_of_idx(n);
        std::iota(ident.begin(), ident.end(), 0u);
        GatherList gl;
        memset(&gl, 0, sizeof(gl));
        for (int a = 0; a < 3; ++a) gl.xpos[a] = -1;
        gl.mpos = -1;
        for (int s = 0; s < NSLOT; ++s) {
            cur[s].assign(n, 0.0);
            gl.from[gl.count] = master[s].data();
            gl.to[gl.count] = cur[s].data();
            if (s >= S_X0 && s < S_X0 + dim) gl.xpos[s - S_X0] = gl.count;
            if (s == S_M) gl.mpos = gl.count;
            ++gl.count;
        }
        gl.h = g.h;
        gl.q6 = g.zrun;
        gl.dim = dim;
        gl.run_phase = g.phase[dim == 3 ? 2 : 1];
        gl.xq = xq.data();
        gl.recA = recA.data();
        launch(n, [&] {
            k_gather(gl, order.data(), ident.data(), idx.data(), pos_of_idx.data(), pkey.data(), key.data(),
                     tag_in.data(), tag_out.data(), col.data(), cellx.data(), n);
        });
        for (int s = 0; s < NSLOT; ++s) f.s[s] = cur[s].data();
        list.assign((size_t)((n + 31) / 32) * (size_t)stride * 32, 0u);
        cnt.assign(n, 0u);
        sorted = true;
        generation += 1;
        list_built = false;
        return 0;
    }

    template <class Op>
    void unary() {
        launch(n, [&] {
            const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
            if (p >= n) return;
            if (dim == 2) Op::template apply<2>(f, prm, p);
            else Op::template apply<3>(f, prm, p);
        });
    }

    template <int DIM, class Op>
    void binary_dim(int self) {
        unsigned long long counters[4] = {0, 0, 0, 0};
        const ColFilter cf{0, 0, (int)g.lim[0] - 1, 1, 0, 1};
        PairList pl{};
        pl.list = list.data();
        pl.cnt = cnt.data();
        pl.xq = xq.data();
        pl.stride = stride;
        pl.qrows = stride + NL_QUEUE_SLACK;
        pl.overflow = &counters[2];
        const uint32_t *k = key.data(), *cx = cellx.data(), *cs = cell_start.data();
        if (list_built) {  // replay
            launch(n, [&] { k_binary_list<DIM, Op>(f, f, prm, g, k, cx, cs, n, self, &counters[0], cf, pl); });
        } else if (generation % 3 != 0) {  // record (integer or FP64 pre-test)
            if (generation % 3 == 1 && g.zrun)  // the 6-bit pre-test needs the zrun cell order
                launch(n, [&] { k_binary_build<DIM, Op, NL_FILTER_Q6>(f, f, prm, g, k, cx, cs, n, self, &counters[0], cf, pl); });
            else
                launch(n, [&] { k_binary_build<DIM, Op, NL_FILTER_F64>(f, f, prm, g, k, cx, cs, n, self, &counters[0], cf, pl); });
            list_built = true;
        } else {  // walk, as k_binary does
            launch(n, [&] {
                const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
                if (p >= n) return;
                const CellCoord home = cell_of(g, k[p], cx[p]);
                Op op;
                op.template init<DIM>(f, prm, p);
                const double px = f.s[S_X0][p], py = f.s[S_X1][p], pz = DIM == 3 ? f.s[S_X2][p] : 0.0;
                unsigned acc = 0;
                nl_walk<DIM>(op, f, prm, g, home, p, px, py, pz, cs, acc);
                if (self) op.template pair<DIM>(f, prm, p, p, 0.0, 0.0, 0.0, 0.0);
                op.template finish<DIM>(f, f, prm, p);
                counters[0] += acc;
            });
        }
        pairs = counters[0];
        overflow += counters[2];
    }
    template <class Op>
    void binary(int self) {
        if (dim == 2) binary_dim<2, Op>(self);
        else binary_dim<3, Op>(self);
    }

    // sphmw_apply by name, through the menu pair_ops.cu dispatches from
    bool apply(const char *name, int self) {
#define EMU_U(NAME, OP, READS, WRITES, EXTRA) \
    if (!strcmp(name, NAME)) {                \
        unary<OP>();                          \
        return true;                          \
    }
#define EMU_B(NAME, OP, READS, WRITES, EXTRA) \
    if (!strcmp(name, NAME)) {                \
        binary<OP>(self);                     \
        return true;                          \
    }
        SPHMW_OPERATOR_MENU(EMU_U, EMU_B)
#undef EMU_U
#undef EMU_B
        return false;
    }
};

int main(int argc, char **argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: emu_ops <input.bin> <output.bin>\n");
        return 2;
    }
    FILE *fp = fopen(argv[1], "rb");
    if (!fp) {
        perror(argv[1]);
        return 2;
    }
    int32_t head[4];  // stride, cx_shift (-1: library default), nparams, nfields
    int64_t n;
    double box[7];
    read_exact(fp, head, sizeof(head));
    read_exact(fp, &n, sizeof(n));
    read_exact(fp, box, sizeof(box));
    System sys;
    sys.n = n;
    sys.stride = head[0];
    if (head[1] >= 0) {  // x-chunked cell order with this chunk width; default: the zrun order
        setenv("SPHMW_CX_SHIFT", std::to_string(head[1]).c_str(), 1);
        setenv("SPHMW_CELL_ORDER", "xchunk", 1);
    }
    memset(&sys.prm, 0, sizeof(Params));
    struct Named {
        const char *name;
        double Params::*field;
    };
#define PRM(x) {#x, &Params::x}
    static const Named TABLE[] = {PRM(dt), PRM(g), PRM(c), PRM(gamma), PRM(alpha), PRM(beta), PRM(eps), PRM(eta),
                                  PRM(rho0), PRM(R_mass), PRM(R_gas), PRM(T_bg), PRM(rho_floor), PRM(P_floor),
                                  PRM(z_t), PRM(z_b), PRM(gamma_r), PRM(fluid), PRM(m), PRM(nu), PRM(mu), PRM(gx),
                                  PRM(gy), PRM(gz), PRM(kh), PRM(dt_pack), PRM(c_pack), PRM(zeta_pack), PRM(U_max),
                                  PRM(cp), PRM(bc_width), PRM(x_inflow), PRM(dr), PRM(inflow)};
#undef PRM
    for (int k = 0; k < head[2]; ++k) {
        char name[16];
        double value;
        read_exact(fp, name, 16);
        read_exact(fp, &value, sizeof(value));
        name[15] = 0;
        bool known = false;
        for (const Named &t : TABLE)
            if (!strcmp(t.name, name)) {
                sys.prm.*(t.field) = value;
                known = true;
            }
        if (!known) {
            fprintf(stderr, "emu_ops: unknown parameter '%s'\n", name);
            return 2;
        }
    }
    sphmw_derive_params(sys.prm);
    memset(&sys.g, 0, sizeof(Grid));
    int64_t global_cols = 0;
    if (sphmw_grid_setup(sys.g, box, box + 3, box[6], -1, -1, &global_cols) != SPHMW_OK) return 3;
    sys.dim = sys.g.dim;
    for (int s = 0; s < NSLOT; ++s) sys.master[s].assign(n, 0.0);  // unset fields read as the constructor's zero
    for (int k = 0; k < head[3]; ++k) {
        int32_t slot;
        read_exact(fp, &slot, sizeof(slot));
        if (slot < 0 || slot >= NSLOT) return 2;
        read_exact(fp, sys.master[slot].data(), sizeof(double) * n);
    }
    int32_t nops;
    read_exact(fp, &nops, sizeof(nops));
    std::vector<unsigned long long> pair_counts;
    for (int k = 0; k < nops; ++k) {
        char name[48];
        int32_t self;
        read_exact(fp, name, 48);
        read_exact(fp, &self, sizeof(self));
        name[47] = 0;
        if (!strcmp(name, "create_cell_list")) {
            if (sys.create_cell_list()) return 3;
            continue;
        }
        if (!sys.sorted) {
            fprintf(stderr, "emu_ops: '%s' before the first create_cell_list\n", name);
            return 2;
        }
        if (!sys.apply(name, self)) {
            fprintf(stderr, "emu_ops: operator '%s' is not in the menu\n", name);
            return 2;
        }
    }
    fclose(fp);
    sys.unsort();
    FILE *out = fopen(argv[2], "wb");
    if (!out) {
        perror(argv[2]);
        return 2;
    }
    const int64_t meta[4] = {n, sys.dim, (int64_t)sys.pairs, (int64_t)sys.overflow};
    fwrite(meta, sizeof(meta), 1, out);
    for (int s = 0; s < NSLOT; ++s) fwrite(sys.master[s].data(), sizeof(double), n, out);
    fclose(out);
    printf("emu_ops: n=%lld dim=%d ops=%d cell lists=%d last pairs=%llu overflow=%llu\n", (long long)n, sys.dim, nops,
           sys.generation, sys.pairs, sys.overflow);
    return 0;
}
