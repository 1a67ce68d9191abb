/* Plain-C user of libsphmw (include/sphmw.h) — test infrastructure.
 *
 * What a Julia `ccall` shim, or any other host, sees: no Python, no torch, only the C ABI.
 *   test_capi layout          sizeof/offsetof of the two structs that cross the ABI, as JSON
 *                             (tests/test_capi_layout.py compares them with the ctypes and the
 *                             Julia declarations; needs no GPU)
 *   test_capi run             one GPU: create -> upload -> create_cell_list -> step("wcsph", 3)
 *                             -> download; prints checksums (the GPU test recomputes them through
 *                             the ctypes binding and against the oracle)
 *   test_capi slabs W         W GPUs, one host thread per GPU, x-slabs with the halo transport
 *                             inside the library (sphmw_comm_init): sphmw_create_cell_list and
 *                             sphmw_step are the SAME calls as on one GPU, and the gathered result
 *                             must equal the one-GPU result bit for bit
 *                             (≙ verlet_step!, src/current/wcsph_perturbed_witch.jl:309-332;
 *                             the reference's caller is single-threaded, src/core.jl:125-142)
 */
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sphmw.h"

#define CHECK(expr)                                                                     \
    do {                                                                                \
        int rc_ = (expr);                                                               \
        if (rc_ < 0) {                                                                  \
            fprintf(stderr, "%s -> %d: %s\n", #expr, rc_, sphmw_last_error());          \
            exit(1);                                                                    \
        }                                                                               \
    } while (0)

/* the driver constants of wcsph_perturbed_witch.jl:25-75 for dr = 26 km / 12 */
static const double DOM_H = 26e3, NY = 12.0, ETA = 1.8, RHO0 = 1.393, G = 9.81, RM = 287.05, TBG = 250.0;
static double dr_(void) { return DOM_H / NY; }
static double h0_(void) { return ETA * dr_(); }
static double c_(void) { return sqrt(65e3 * (7.0 / 5.0) / RHO0); }

static void set_params(sphmw_ctx *ctx) {
    const double cp = 7 * RM / 2, cv = cp - RM;
    const char *names[] = {"dt", "g", "c", "gamma", "alpha", "beta", "eps", "eta", "rho0", "R_mass", "R_gas",
                           "T_bg", "rho_floor", "P_floor", "z_t", "z_b", "gamma_r", "fluid"};
    const double vals[] = {0.01 * h0_() / c_(), G, c_(), cp / cv, 0.1, 0.2, 0.01, ETA, RHO0, RM, 8.314,
                           TBG, 1e-6, 1e-10, DOM_H, 12e3, 10 * sqrt(0.0196), 0.0};
    for (size_t k = 0; k < sizeof(vals) / sizeof(vals[0]); ++k) CHECK(sphmw_set_param(ctx, names[k], vals[k]));
}

typedef struct {
    int64_t n;
    double *x, *v, *m, *h, *rho, *rho_p, *type; /* x, v: component-major (3 x n) */
} Cloud;

/* cubic lattice nx x ny x nz, density falling with height, a sheared wind so that the pair force works;
 * only + - * / so that a numpy restatement of the inputs (tests/test_gpu_capi_c.py) has the same bits */
static Cloud make_cloud(int nx, int ny, int nz) {
    Cloud c;
    c.n = (int64_t)nx * ny * nz;
    c.x = malloc(sizeof(double) * 3 * c.n);
    c.v = malloc(sizeof(double) * 3 * c.n);
    c.m = malloc(sizeof(double) * c.n);
    c.h = malloc(sizeof(double) * c.n);
    c.rho = malloc(sizeof(double) * c.n);
    c.rho_p = calloc(c.n, sizeof(double));
    c.type = calloc(c.n, sizeof(double));
    const double dr = dr_();
    int64_t p = 0;
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            for (int k = 0; k < nz; ++k, ++p) {
                const double x = (i + 0.5) * dr, y = (j + 0.5) * dr, z = (k + 0.5) * dr;
                c.x[p] = x, c.x[c.n + p] = y, c.x[2 * c.n + p] = z;
                c.v[p] = 20.0 + 5.0 * ((7 * j + 3 * k) % 11) / 11.0, c.v[c.n + p] = 0.5 * ((9 * i) % 7) / 7.0 - 0.25, c.v[2 * c.n + p] = 0.0;
                c.rho[p] = RHO0 / (1.0 + y * G / (RM * TBG));
                c.m[p] = c.rho[p] * dr * dr * dr;
                c.h[p] = h0_();
            }
    return c;
}

static sphmw_config box_config(int nx, int ny, int nz, int64_t cap, int device) {
    sphmw_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    const double dr = dr_();
    cfg.box_max[0] = nx * dr, cfg.box_max[1] = ny * dr, cfg.box_max[2] = nz * dr;
    cfg.h = h0_();
    cfg.capacity = cap;
    cfg.device = device;
    cfg.flags = SPHMW_FLAG_NONE;
    cfg.slab_lo = cfg.slab_hi = -1;
    return cfg;
}

static void upload_all(sphmw_ctx *ctx, const Cloud *c, const int64_t *sel, int64_t m) {
    /* sel == NULL: everything; else the m selected particles, in that order */
    const int64_t n = sel ? m : c->n;
    double *buf = malloc(sizeof(double) * 3 * n);
    const struct { const char *name; const double *src; int nc; } F[] = {
        {"x", c->x, 3}, {"v", c->v, 3}, {"m", c->m, 1}, {"h", c->h, 1}, {"rho", c->rho, 1}, {"rho_p", c->rho_p, 1},
        {"type", c->type, 1}};
    CHECK(sphmw_resize(ctx, n));
    for (size_t f = 0; f < sizeof(F) / sizeof(F[0]); ++f) {
        for (int k = 0; k < F[f].nc; ++k)
            for (int64_t i = 0; i < n; ++i) buf[k * n + i] = F[f].src[k * c->n + (sel ? sel[i] : i)];
        CHECK(sphmw_upload(ctx, F[f].name, buf, n, F[f].nc));
    }
    free(buf);
}

static uint64_t bits_sum(const double *a, int64_t n) {
    uint64_t s = 0;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t b;
        memcpy(&b, a + i, 8);
        s = s * 1099511628211ull + b; /* order-dependent */
    }
    return s;
}

#define NX 40
#define NY_ 10
#define NZ 8
#define NSTEPS 3

static void run_single(int device, double **x_out, double **v_out, double **rho_out, int64_t *n_out, int quiet) {
    Cloud c = make_cloud(NX, NY_, NZ);
    sphmw_config cfg = box_config(NX, NY_, NZ, c.n + 1024, device);
    sphmw_ctx *ctx = NULL;
    CHECK(sphmw_create(&cfg, &ctx));
    set_params(ctx);
    upload_all(ctx, &c, NULL, 0);
    int64_t alive = 0;
    CHECK(sphmw_create_cell_list(ctx, &alive));
    CHECK(sphmw_step(ctx, "wcsph", NSTEPS));
    int64_t n = 0;
    CHECK(sphmw_count(ctx, &n));
    double *x = malloc(sizeof(double) * 3 * n), *v = malloc(sizeof(double) * 3 * n), *rho = malloc(sizeof(double) * n);
    CHECK(sphmw_download(ctx, "x", x, n, 3));
    CHECK(sphmw_download(ctx, "v", v, n, 3));
    CHECK(sphmw_download(ctx, "rho", rho, n, 1));
    int64_t launches = 0;
    CHECK(sphmw_launch_count(ctx, &launches));
    if (!quiet)
        printf("{\"mode\": \"run\", \"n0\": %lld, \"alive\": %lld, \"n\": %lld, \"launches\": %lld, \"x_bits\": \"%016llx\", "
               "\"v_bits\": \"%016llx\", \"rho_bits\": \"%016llx\", \"rho_first\": %.17g, \"vx_last\": %.17g}\n",
               (long long)c.n, (long long)alive, (long long)n, (long long)launches, (unsigned long long)bits_sum(x, 3 * n),
               (unsigned long long)bits_sum(v, 3 * n), (unsigned long long)bits_sum(rho, n), rho[0], v[n - 1]);
    CHECK(sphmw_destroy(ctx));
    *x_out = x, *v_out = v, *rho_out = rho, *n_out = n;
}

typedef struct {
    int rank, world;
    unsigned char id[128];
    const Cloud *cloud;
    int64_t lim0, phase0;
    /* results, by global index */
    double *x, *v, *rho;
    int64_t owned;
    int failed;
} RankArg;

static void *rank_main(void *argp) {
    RankArg *a = (RankArg *)argp;
    const Cloud *c = a->cloud;
    const int64_t lo = a->lim0 * a->rank / a->world, hi = a->lim0 * (a->rank + 1) / a->world;
    int64_t *sel = malloc(sizeof(int64_t) * c->n), m = 0;
    for (int64_t p = 0; p < c->n; ++p) {
        const int64_t col = (int64_t)floor(c->x[p] / h0_()) - a->phase0;
        if (col >= lo && col < hi) sel[m++] = p;
    }
    const int64_t halo_cap = c->n / a->world + 4096;
    sphmw_config cfg = box_config(NX, NY_, NZ, m + 2 * halo_cap + 1024, a->rank);
    cfg.slab_lo = lo, cfg.slab_hi = hi;
    sphmw_ctx *ctx = NULL;
    CHECK(sphmw_create(&cfg, &ctx));
    set_params(ctx);
    upload_all(ctx, c, sel, m);
    CHECK(sphmw_set_index(ctx, sel, m));                                   /* global particle indices */
    CHECK(sphmw_comm_init(ctx, a->rank, a->world, a->id, halo_cap));       /* collective */
    CHECK(sphmw_create_cell_list(ctx, NULL));                              /* halo exchange + sort */
    CHECK(sphmw_step(ctx, "wcsph", 1));                                    /* plain schedule */
    CHECK(sphmw_step(ctx, "wcsph", NSTEPS - 1));                           /* overlapped schedule */
    int64_t nres = 0, nown = 0;
    CHECK(sphmw_slab_counts(ctx, &nres, &nown));
    int64_t *gidx = malloc(sizeof(int64_t) * nres);
    int32_t *tag = malloc(sizeof(int32_t) * nres);
    double *x = malloc(sizeof(double) * 3 * nres), *v = malloc(sizeof(double) * 3 * nres), *rho = malloc(sizeof(double) * nres);
    CHECK(sphmw_download_index(ctx, gidx, tag, nres));
    CHECK(sphmw_download_raw(ctx, "x", x, nres, 3));
    CHECK(sphmw_download_raw(ctx, "v", v, nres, 3));
    CHECK(sphmw_download_raw(ctx, "rho", rho, nres, 1));
    for (int64_t p = 0; p < nres; ++p) {
        if (tag[p] != 0) continue; /* ghosts */
        const int64_t gi = gidx[p];
        for (int k = 0; k < 3; ++k) a->x[k * c->n + gi] = x[k * nres + p], a->v[k * c->n + gi] = v[k * nres + p];
        a->rho[gi] = rho[p];
        a->owned += 1;
    }
    if (a->owned != nown) a->failed = 1;
    CHECK(sphmw_destroy(ctx));
    free(sel), free(gidx), free(tag), free(x), free(v), free(rho);
    return NULL;
}

static int run_slabs(int world) {
    double *x1, *v1, *r1;
    int64_t n1;
    run_single(0, &x1, &v1, &r1, &n1, 1);
    Cloud c = make_cloud(NX, NY_, NZ);
    /* key tables of the whole domain (structs.jl:66-68) */
    sphmw_config cfg = box_config(NX, NY_, NZ, 1024, 0);
    sphmw_ctx *probe = NULL;
    int64_t phase[3], lim[3], key_max;
    int32_t dim;
    CHECK(sphmw_create(&cfg, &probe));
    CHECK(sphmw_key_tables(probe, phase, lim, &key_max, &dim));
    CHECK(sphmw_destroy(probe));
    RankArg *args = calloc(world, sizeof(RankArg));
    pthread_t *th = calloc(world, sizeof(pthread_t));
    double *x = calloc(3 * c.n, sizeof(double)), *v = calloc(3 * c.n, sizeof(double)), *rho = calloc(c.n, sizeof(double));
    unsigned char id[128];
    CHECK(sphmw_comm_unique_id(id));
    for (int r = 0; r < world; ++r) {
        args[r].rank = r, args[r].world = world, args[r].cloud = &c, args[r].lim0 = lim[0], args[r].phase0 = phase[0];
        memcpy(args[r].id, id, 128);
        args[r].x = x, args[r].v = v, args[r].rho = rho;
        pthread_create(&th[r], NULL, rank_main, &args[r]);
    }
    int64_t owned = 0;
    int failed = 0;
    for (int r = 0; r < world; ++r) {
        pthread_join(th[r], NULL);
        owned += args[r].owned;
        failed |= args[r].failed;
    }
    const int same = owned == n1 && n1 == c.n && !memcmp(x, x1, sizeof(double) * 3 * n1) &&
                     !memcmp(v, v1, sizeof(double) * 3 * n1) && !memcmp(rho, r1, sizeof(double) * n1);
    printf("{\"mode\": \"slabs\", \"world\": %d, \"n\": %lld, \"owned\": %lld, \"bitwise_equal_to_one_gpu\": %s}\n", world,
           (long long)n1, (long long)owned, same && !failed ? "true" : "false");
    return same && !failed ? 0 : 1;
}

int main(int argc, char **argv) {
    const char *mode = argc > 1 ? argv[1] : "layout";
    if (!strcmp(mode, "layout")) {
        printf("{\"sizeof_config\": %zu, \"config\": {\"box_min\": %zu, \"box_max\": %zu, \"h\": %zu, \"capacity\": %zu, "
               "\"device\": %zu, \"flags\": %zu, \"slab_lo\": %zu, \"slab_hi\": %zu}, \"sizeof_lattice_setup\": %zu, "
               "\"lattice_setup\": {\"grid\": %zu, \"mountain\": %zu, \"dr\": %zu, \"dom_min\": %zu, \"dom_max\": %zu, "
               "\"bc_width\": %zu, \"h_m\": %zu, \"a\": %zu, \"U\": %zu, \"type_fluid\": %zu, \"type_wall\": %zu, "
               "\"type_mountain\": %zu, \"h0\": %zu}, \"halo_record_doubles\": %d}\n",
               sizeof(sphmw_config), offsetof(sphmw_config, box_min), offsetof(sphmw_config, box_max),
               offsetof(sphmw_config, h), offsetof(sphmw_config, capacity), offsetof(sphmw_config, device),
               offsetof(sphmw_config, flags), offsetof(sphmw_config, slab_lo), offsetof(sphmw_config, slab_hi),
               sizeof(sphmw_lattice_setup), offsetof(sphmw_lattice_setup, grid), offsetof(sphmw_lattice_setup, mountain),
               offsetof(sphmw_lattice_setup, dr), offsetof(sphmw_lattice_setup, dom_min),
               offsetof(sphmw_lattice_setup, dom_max), offsetof(sphmw_lattice_setup, bc_width),
               offsetof(sphmw_lattice_setup, h_m), offsetof(sphmw_lattice_setup, a), offsetof(sphmw_lattice_setup, U),
               offsetof(sphmw_lattice_setup, type_fluid), offsetof(sphmw_lattice_setup, type_wall),
               offsetof(sphmw_lattice_setup, type_mountain), offsetof(sphmw_lattice_setup, h0),
               sphmw_halo_record_doubles());
        return 0;
    }
    if (!strcmp(mode, "run")) {
        double *x, *v, *rho;
        int64_t n;
        run_single(0, &x, &v, &rho, &n, 0);
        return 0;
    }
    if (!strcmp(mode, "slabs")) return run_slabs(argc > 2 ? atoi(argv[2]) : 2);
    fprintf(stderr, "usage: test_capi layout | run | slabs W\n");
    return 2;
}
