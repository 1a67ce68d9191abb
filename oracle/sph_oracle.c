/*
 * oracle/sph_oracle.c — CPU restatement of the reference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (sph_mountain_waves_b200/,
 * libsphmw.so) links, imports or calls this file.  It may be used by tests/,
 * by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * leg, always as the checker or the timed CPU baseline, never as the product.
 *
 * PARITY UNPINNED at the bit level: the reference (moschehaus/sph-mountain-waves,
 * a SmoothedParticles.jl v0.2.0 fork) is 100 % Julia, `julia` is not installed in
 * the build image and the reference ships no golden arrays for cell keys,
 * neighbour lists, density or velocity.  What the reference's own
 * tests do assert — kernel normalisation/derivative properties
 * (sph_jl/tests/test_kernels.jl:19-43) and the two-disc collision invariants
 * (sph_jl/tests/test_collision_2d.jl:141-147) — is checked against this file in
 * tests/test_oracle_*.py.
 *
 * Every function cites the reference file:line it restates.  Build with
 *   gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC
 * so that no multiply-add is contracted into an FMA (Julia/LLVM never contracts
 * without @fastmath) and sqrt/div are IEEE.  libm's exp/pow/sin/cbrt stand in for
 * Julia's (<= 1 ulp apart; covered by the 1e-10 tolerance of the north star).
 *
 * Indices are 1-based inside cells (0 = vacant slot, as in the reference) and
 * 0-based across the C API.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* Particle: superset of the driver structs.                                 */
/*   wcsph_perturbed_witch.jl:83-102 (18 fields)                             */
/*   hopkins_perturbed_witch.jl:103 (+A), full_hopkins_perturbed_witch.jl:104 */
/*   (+A_bg), collapse_dry.jl:74-82 (rho, Drho), test_collision_2d.jl:37-44  */
/*   (a == Dv here, rho0)                                                    */
/* Each particle is a separately allocated object and the system holds a     */
/* vector of pointers, like Julia's Vector{mutable struct} (structs.jl:53).  */
/* ------------------------------------------------------------------------- */
typedef struct particle {
    double h;
    double x[3];
    double m;
    double v[3];
    double Dv[3];
    double rho_bg, rho_p, rho;
    double P_bg, P_p, P;
    double th_bg, th_p, th;
    double T_bg, T_p, T;
    double type;
    double A, A_bg;
    double Drho, rho0;
    double S, s; /* entropy, entropy density: src/legacy/adiabatic_flow_witch.jl:75-76 */
} particle;

typedef struct {
    const char *name;
    size_t off;
    int ncomp;
} field_desc;

#define OFF(f) offsetof(particle, f)
static const field_desc FIELDS[] = {
    {"h", OFF(h), 1},         {"x", OFF(x), 3},       {"m", OFF(m), 1},
    {"v", OFF(v), 3},         {"Dv", OFF(Dv), 3},     {"rho_bg", OFF(rho_bg), 1},
    {"rho_p", OFF(rho_p), 1}, {"rho", OFF(rho), 1},   {"P_bg", OFF(P_bg), 1},
    {"P_p", OFF(P_p), 1},     {"P", OFF(P), 1},       {"theta_bg", OFF(th_bg), 1},
    {"theta_p", OFF(th_p), 1},{"theta", OFF(th), 1},  {"T_bg", OFF(T_bg), 1},
    {"T_p", OFF(T_p), 1},     {"T", OFF(T), 1},       {"type", OFF(type), 1},
    {"A", OFF(A), 1},         {"A_bg", OFF(A_bg), 1}, {"Drho", OFF(Drho), 1},
    {"rho0", OFF(rho0), 1},   {"S", OFF(S), 1},       {"s", OFF(s), 1},
    {NULL, 0, 0}};

/* driver constants: wcsph_perturbed_witch.jl:25-75, collapse_dry.jl:30-66,
 * test_collision_2d.jl:14-35 */
typedef struct {
    double dt, g, c, gamma, alpha, beta, eps, eta, rho0, R_mass, R_gas, T_bg;
    double rho_floor, P_floor, z_t, z_b, gamma_r, fluid;
    double m, nu, mu, gx, gy, gz, kh; /* fixed-mass / fixed-h examples */
    double dt_pack, c_pack, zeta_pack; /* utils/new_packing.jl:1-3 */
    /* legacy flow drivers: src/legacy/isothermal_flow_witch.jl:24-60 */
    double U_max, cp, bc_width, x_inflow, dr, inflow;
} params;

typedef struct {
    const char *name;
    size_t off;
} param_desc;
#define POFF(f) offsetof(params, f)
static const param_desc PARAMS[] = {
    {"dt", POFF(dt)},           {"g", POFF(g)},         {"c", POFF(c)},
    {"gamma", POFF(gamma)},     {"alpha", POFF(alpha)}, {"beta", POFF(beta)},
    {"eps", POFF(eps)},         {"eta", POFF(eta)},     {"rho0", POFF(rho0)},
    {"R_mass", POFF(R_mass)},   {"R_gas", POFF(R_gas)}, {"T_bg", POFF(T_bg)},
    {"rho_floor", POFF(rho_floor)}, {"P_floor", POFF(P_floor)},
    {"z_t", POFF(z_t)},         {"z_b", POFF(z_b)},     {"gamma_r", POFF(gamma_r)},
    {"fluid", POFF(fluid)},     {"m", POFF(m)},         {"nu", POFF(nu)},
    {"mu", POFF(mu)},           {"gx", POFF(gx)},       {"gy", POFF(gy)},
    {"gz", POFF(gz)},           {"kh", POFF(kh)},       {"dt_pack", POFF(dt_pack)},
    {"c_pack", POFF(c_pack)},   {"zeta_pack", POFF(zeta_pack)}, {"U_max", POFF(U_max)},
    {"cp", POFF(cp)},           {"bc_width", POFF(bc_width)},   {"x_inflow", POFF(x_inflow)},
    {"dr", POFF(dr)},           {"inflow", POFF(inflow)},       {NULL, 0}};

/* structs.jl:22-26 — cell = growable index vector + lock */
typedef struct {
    int64_t *e;
    int64_t len;
#ifdef _OPENMP
    omp_lock_t lock;
#endif
} cell;

/* structs.jl:43-56 */
typedef struct orc_system {
    double h;
    double box[6]; /* x1_min x2_min x3_min x1_max x2_max x3_max (geometry.jl:15-22) */
    int64_t key_phase[3], key_lim[3], key_max;
    int64_t key_diff[27];
    int ndiff;
    int dim;
    particle **p;
    int64_t n, cap;
    cell *cells;
    cell removal;
    params prm;
    int64_t pair_count; /* accepted pairs of the last binary pass */
} orc_system;

/* ------------------------------------------------------------------------- */
/* algebra.jl:49-60 — dot and norm, left-to-right, no fastmath                */
/* ------------------------------------------------------------------------- */
static inline double dot3(const double *a, const double *b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static inline double norm3(const double *a) { return sqrt(dot3(a, a)); }

/* core.jl:8-10 */
static inline double dist(const particle *p, const particle *q) {
    double d[3] = {p->x[0] - q->x[0], p->x[1] - q->x[1], p->x[2] - q->x[2]};
    return norm3(d);
}

/* ------------------------------------------------------------------------- */
/* kernels.jl — smoothing kernels.  @fastmath in the reference: integer powers */
/* lower to llvm.powi with a constant exponent (repeated squaring); the        */
/* evaluation order written here is the obvious left-to-right one. */
/* ------------------------------------------------------------------------- */
static inline double pow2(double a) { return a * a; }
static inline double pow3(double a) { return a * a * a; }
static inline double pow4(double a) { double b = a * a; return b * b; }
static inline double pow5(double a) { double b = a * a; return b * b * a; }

/* kernels.jl:108-115 */
double orc_wendland2(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return 2.228169203286535 * pow4(1.0 - x) * (1.0 + 4.0 * x) / pow2(h);
}
/* kernels.jl:124-131 */
double orc_Dwendland2(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -44.563384065730695 * x * pow3(1.0 - x) / pow3(h);
}
/* kernels.jl:140-147 */
double orc_rDwendland2(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -44.563384065730695 * pow3(1.0 - x) / pow4(h);
}
/* kernels.jl:156-163 */
double orc_wendland3(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return 3.3422538049298023 * pow4(1.0 - x) * (1.0 + 4.0 * x) / pow3(h);
}
/* kernels.jl:172-179 */
double orc_Dwendland3(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -66.84507609859604 * x * pow3(1.0 - x) / pow4(h);
}
/* kernels.jl:188-195 */
double orc_rDwendland3(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -66.84507609859604 * pow3(1.0 - x) / pow5(h);
}
/* kernels.jl:197-204 */
double orc_DDwendland3(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -66.84507609859604 * ((1.0 - 4.0 * x) * pow2(1.0 - x)) / pow5(h);
}
/* kernels.jl:206-212 */
double orc_wendland1(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return 1.5 * pow4(1.0 - x) * (1.0 + 4.0 * x) / h;
}
/* kernels.jl:214-220 */
double orc_Dwendland1(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -30.0 * x * pow3(1.0 - x) / pow2(h);
}
/* kernels.jl:222-228 */
double orc_rDwendland1(double h, double r) {
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -30.0 * pow3(1.0 - x) / pow3(h);
}
/* kernels.jl:3-5 */
static inline double pos(double x) { return x > 0.0 ? x : 0.0; }
/* kernels.jl:14-25 */
double orc_spline23(double h, double r) {
    double x = r / h;
    if (x < 0.5) return 1.8189136353359467 * (1.0 - 6.0 * pow2(x) + 6.0 * pow3(x)) / pow2(h);
    else if (x < 1.0) return 3.6378272706718935 * pow3(1.0 - x) / pow2(h);
    return 0.0;
}
/* kernels.jl:34-43 */
double orc_Dspline23(double h, double r) {
    double x = r / h;
    if (x < 0.5) return -10.91348181201568 * (2.0 * x - 3.0 * pow2(x)) / pow3(h);
    else if (x < 1.0) return -10.91348181201568 * pow2(1.0 - x) / pow3(h);
    return 0.0;
}
/* kernels.jl:52-61 */
double orc_rDspline23(double h, double r) {
    double x = r / h;
    if (x < 0.5) return -10.91348181201568 * (2.0 - 3.0 * x) / pow4(h);
    else if (x < 1.0) return -10.91348181201568 * pow2(1.0 - x) / (x * pow4(h));
    return 0.0;
}
/* kernels.jl:70-73 */
double orc_spline24(double h, double r) {
    double x = r / h;
    return 6.222175110452539 *
           (pow4(pos(1.0 - x)) - 5 * pow4(pos(0.6 - x)) + 10 * pow4(pos(0.2 - x))) / pow2(h);
}
/* kernels.jl:82-85 */
double orc_Dspline24(double h, double r) {
    double x = r / h;
    return -24.888700441810155 *
           (pow3(pos(1.0 - x)) - 5 * pow3(pos(0.6 - x)) + 10 * pow3(pos(0.2 - x))) / pow3(h);
}
/* kernels.jl:94-100 */
double orc_rDspline24(double h, double r) {
    double x = r / h;
    if (x > 0.2)
        return -24.888700441810155 * (pow3(pos(1.0 - x)) - 5 * pow3(pos(0.6 - x))) / (x * pow4(h));
    return -24.888700441810155 * (1.2 - 6.0 * pow2(x)) / pow4(h);
}

/* ------------------------------------------------------------------------- */
/* geometry.jl:24-30 — closed-interval box test; NaN fails every comparison   */
/* ------------------------------------------------------------------------- */
static inline int is_inside_box(const double *x, const double *b) {
    return b[0] <= x[0] && x[0] <= b[3] && b[1] <= x[1] && x[1] <= b[4] &&
           b[2] <= x[2] && x[2] <= b[5];
}

/* structs.jl:97-106 — IEEE division then floor, 1-based key */
static inline int64_t find_key(const orc_system *s, const double *x) {
    int64_t i = 1 + (int64_t)floor(x[0] / s->h) - s->key_phase[0];
    int64_t j = 1 + (int64_t)floor(x[1] / s->h) - s->key_phase[1];
    int64_t k = 1 + (int64_t)floor(x[2] / s->h) - s->key_phase[2];
    return i + s->key_lim[0] * (j - 1) + s->key_lim[0] * s->key_lim[1] * (k - 1);
}

static void cell_init(cell *c) {
    c->e = NULL;
    c->len = 0;
#ifdef _OPENMP
    omp_init_lock(&c->lock);
#endif
}
static void cell_free(cell *c) {
    free(c->e);
#ifdef _OPENMP
    omp_destroy_lock(&c->lock);
#endif
}

/* core.jl:13-24 — first vacant (0) slot, grow by one when full */
static int64_t find_vacation(cell *c) {
    for (int64_t i = 0; i < c->len; ++i)
        if (c->e[i] == 0) return i;
    int64_t i = c->len;
    c->e = (int64_t *)realloc(c->e, (size_t)(i + 1) * sizeof(int64_t));
    c->len = i + 1;
    return i;
}

/* core.jl:26-41 — insert under the cell's lock, keep entries DESCENDING */
static void add_index(cell *c, int64_t idx1) {
#ifdef _OPENMP
    omp_set_lock(&c->lock);
#endif
    int64_t ind = find_vacation(c);
    c->e[ind] = idx1;
    while (ind > 0 && c->e[ind - 1] < c->e[ind]) {
        int64_t t = c->e[ind];
        c->e[ind] = c->e[ind - 1];
        c->e[ind - 1] = t;
        --ind;
    }
#ifdef _OPENMP
    omp_unset_lock(&c->lock);
#endif
}

/* structs.jl:57-91 — constructor: key tables from the bounding box */
orc_system *orc_create(const double *box, double h) {
    if (!(h > 0.0)) return NULL; /* @assert h > 0 (structs.jl:59) */
    orc_system *s = (orc_system *)calloc(1, sizeof(orc_system));
    s->h = h;
    memcpy(s->box, box, 6 * sizeof(double));
    s->key_max = 1;
    for (int a = 0; a < 3; ++a) {
        s->key_phase[a] = (int64_t)floor(box[a] / h);
        s->key_lim[a] = (int64_t)floor(box[3 + a] / h) - s->key_phase[a] + 1;
        s->key_max *= s->key_lim[a];
    }
    s->ndiff = 0;
    if (s->key_lim[2] == 1) { /* structs.jl:70-75: 2D, di outer / dj inner */
        s->dim = 2;
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj)
                s->key_diff[s->ndiff++] = di + s->key_lim[0] * dj;
    } else { /* structs.jl:76-82 */
        s->dim = 3;
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj)
                for (int dk = -1; dk <= 1; ++dk)
                    s->key_diff[s->ndiff++] =
                        di + s->key_lim[0] * (dj + s->key_lim[1] * dk);
    }
    s->cells = (cell *)malloc((size_t)s->key_max * sizeof(cell));
    for (int64_t k = 0; k < s->key_max; ++k) cell_init(&s->cells[k]);
    cell_init(&s->removal);
    s->prm.fluid = 0.0;
    return s;
}

void orc_destroy(orc_system *s) {
    if (!s) return;
    for (int64_t i = 0; i < s->n; ++i) free(s->p[i]);
    free(s->p);
    for (int64_t k = 0; k < s->key_max; ++k) cell_free(&s->cells[k]);
    free(s->cells);
    cell_free(&s->removal);
    free(s);
}

int orc_dim(const orc_system *s) { return s->dim; }
int64_t orc_n(const orc_system *s) { return s->n; }
int64_t orc_key_max(const orc_system *s) { return s->key_max; }
void orc_key_tables(const orc_system *s, int64_t *phase, int64_t *lim) {
    for (int a = 0; a < 3; ++a) { phase[a] = s->key_phase[a]; lim[a] = s->key_lim[a]; }
}
int64_t orc_pair_count(const orc_system *s) { return s->pair_count; }

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_set_param(orc_system *s, const char *name, double v) {
    for (const param_desc *d = PARAMS; d->name; ++d)
        if (!strcmp(d->name, name)) {
            *(double *)((char *)&s->prm + d->off) = v;
            return 0;
        }
    return -1;
}

/* grids.jl:305-310 — push!(sys.particles, ctor(x)); all fields start at 0 */
int orc_append(orc_system *s, int64_t n_new) {
    if (s->n + n_new > s->cap) {
        s->cap = (s->n + n_new) * 2;
        s->p = (particle **)realloc(s->p, (size_t)s->cap * sizeof(particle *));
    }
    for (int64_t i = 0; i < n_new; ++i)
        s->p[s->n + i] = (particle *)calloc(1, sizeof(particle));
    s->n += n_new;
    return 0;
}

static const field_desc *find_field(const char *name) {
    for (const field_desc *d = FIELDS; d->name; ++d)
        if (!strcmp(d->name, name)) return d;
    return NULL;
}

/* SoA host buffers, component-major: buf[c*n + i] */
int orc_set_field(orc_system *s, const char *name, const double *buf, int64_t first,
                  int64_t n) {
    const field_desc *d = find_field(name);
    if (!d || first < 0 || first + n > s->n) return -1;
    for (int c = 0; c < d->ncomp; ++c)
        for (int64_t i = 0; i < n; ++i)
            ((double *)((char *)s->p[first + i] + d->off))[c] = buf[(int64_t)c * n + i];
    return 0;
}
int orc_get_field(const orc_system *s, const char *name, double *buf) {
    const field_desc *d = find_field(name);
    if (!d) return -1;
    for (int c = 0; c < d->ncomp; ++c)
        for (int64_t i = 0; i < s->n; ++i)
            buf[(int64_t)c * s->n + i] = ((const double *)((const char *)s->p[i] + d->off))[c];
    return 0;
}

/* ------------------------------------------------------------------------- */
/* core.jl:51-90 — create_cell_list!                                          */
/* ------------------------------------------------------------------------- */
void orc_create_cell_list(orc_system *s) {
    /* :54-58 declare all entries null */
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < s->key_max; ++k) {
        cell *c = &s->cells[k];
        for (int64_t e = 0; e < c->len; ++e) c->e[e] = 0;
    }
    /* :59-61 */
    for (int64_t e = 0; e < s->removal.len; ++e) s->removal.e[e] = 0;

    /* :64-69 identify particles outside the (bounding-box) domain */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < s->n; ++i)
        if (!is_inside_box(s->p[i]->x, s->box)) add_index(&s->removal, i + 1);

    /* :72-81 serial swap-from-end in DESCENDING removed-index order.  The Julia
     * vector only drops the reference; the object is garbage.  We must free
     * exactly the removed objects, so remember them first. */
    {
        int64_t i = 1;
        int64_t nrem = 0;
        while (nrem < s->removal.len && s->removal.e[nrem] != 0) ++nrem;
        particle **dead = NULL;
        if (nrem) {
            dead = (particle **)malloc((size_t)nrem * sizeof(particle *));
            for (int64_t r = 0; r < nrem; ++r) dead[r] = s->p[s->removal.e[r] - 1];
        }
        while (i <= s->removal.len && s->removal.e[i - 1] != 0) {
            s->p[s->removal.e[i - 1] - 1] = s->p[s->n - i]; /* particles[end+1-i] */
            ++i;
        }
        if (i > 1) s->n = s->n + 1 - i;
        for (int64_t r = 0; r < nrem; ++r) free(dead[r]);
        free(dead);
    }

    /* :84-89 fill the cell list */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < s->n; ++i) {
        int64_t key = find_key(s, s->p[i]->x);
        add_index(&s->cells[key - 1], i + 1);
    }
}

/* test hooks ---------------------------------------------------------------- */
void orc_cell_keys(const orc_system *s, int64_t *keys) { /* 0-based keys */
    for (int64_t i = 0; i < s->n; ++i) keys[i] = find_key(s, s->p[i]->x) - 1;
}
/* entries of one cell in stored (descending) order, 0-based particle indices */
int64_t orc_cell_entries(const orc_system *s, int64_t key0, int64_t *out, int64_t cap) {
    const cell *c = &s->cells[key0];
    int64_t n = 0;
    for (int64_t e = 0; e < c->len && c->e[e] != 0; ++e) {
        if (n < cap) out[n] = c->e[e] - 1;
        ++n;
    }
    return n;
}

typedef void (*unary_fn)(particle *, const orc_system *);
typedef void (*binary_fn)(particle *, const particle *, double, const orc_system *);

/* core.jl:94-112 — neighbour traversal for one particle */
static inline int64_t apply_binary_one(const orc_system *s, binary_fn f, particle *p) {
    int64_t cnt = 0;
    int64_t key = find_key(s, p->x);
    for (int d = 0; d < s->ndiff; ++d) {
        int64_t nk = key + s->key_diff[d];
        if (1 <= nk && nk <= s->key_max) { /* :98 — no per-axis wrap check */
            const cell *c = &s->cells[nk - 1];
            for (int64_t e = 0; e < c->len; ++e) {
                int64_t j = c->e[e];
                if (j == 0) break; /* :100-102 */
                particle *q = s->p[j - 1];
                double r = dist(p, q);
                if ((r > s->h) || (p == q)) continue; /* :105 */
                f(p, q, r, s);
                ++cnt;
            }
        }
    }
    return cnt;
}

/* core.jl:125-129 */
static void apply_binary(orc_system *s, binary_fn f) {
    int64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int64_t i = 0; i < s->n; ++i) total += apply_binary_one(s, f, s->p[i]);
    s->pair_count = total;
}
/* core.jl:138-142 */
static void apply_unary(orc_system *s, unary_fn f) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < s->n; ++i) f(s->p[i], s);
}

/* accepted (p,q) pairs in traversal order; returns the total, fills <= cap */
int64_t orc_pairs(const orc_system *s, int64_t *pi, int64_t *pj, int64_t cap) {
    int64_t n = 0;
    for (int64_t i = 0; i < s->n; ++i) {
        const particle *p = s->p[i];
        int64_t key = find_key(s, p->x);
        for (int d = 0; d < s->ndiff; ++d) {
            int64_t nk = key + s->key_diff[d];
            if (1 <= nk && nk <= s->key_max) {
                const cell *c = &s->cells[nk - 1];
                for (int64_t e = 0; e < c->len; ++e) {
                    int64_t j = c->e[e];
                    if (j == 0) break;
                    const particle *q = s->p[j - 1];
                    double r = dist(p, q);
                    if ((r > s->h) || (p == q)) continue;
                    if (n < cap) { pi[n] = i; pj[n] = j - 1; }
                    ++n;
                }
            }
        }
    }
    return n;
}

/* ------------------------------------------------------------------------- */
/* Operators of src/current/wcsph_perturbed_witch.jl                          */
/* In 3D (our extrusion) wendland2 -> wendland3 and            */
/* sqrt(m/rho) -> cbrt(m/rho); everything else is dimension-agnostic.         */
/* ------------------------------------------------------------------------- */
static inline double W(const orc_system *s, double h, double r) {
    return s->dim == 2 ? orc_wendland2(h, r) : orc_wendland3(h, r);
}
static inline double rDW(const orc_system *s, double h, double r) {
    return s->dim == 2 ? orc_rDwendland2(h, r) : orc_rDwendland3(h, r);
}

/* :177-179 */
static inline double background_density(const params *c, double y) {
    return c->rho0 * exp(-y * c->g / (c->R_mass * c->T_bg));
}
/* :181-184 */
static inline double background_pressure(const params *c, double y) {
    double rho_bg = background_density(c, y);
    return c->R_mass * c->T_bg * rho_bg;
}
/* :186-189 */
static inline double background_pot_temperature(const params *c, double y) {
    double P_bg = background_pressure(c, y);
    return c->T_bg * pow((c->T_bg * c->R_gas * c->rho0) / P_bg, 2.0 / 7.0);
}

/* :195-199 */
static void w_compute_pressure(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    p->P_bg = background_pressure(c, p->x[1]);
    p->P_p = pow2(c->c) * p->rho_p;
    p->P = p->P_bg + p->P_p;
}
/* :205-208 */
static void w_find_temperature(particle *p, const orc_system *s) {
    p->T = p->P / (s->prm.R_mass * p->rho);
    p->T_p = p->T - p->T_bg;
}
/* :210-214 */
static void w_find_pot_temp(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    p->th = p->T * pow((c->T_bg * c->R_gas * c->rho0) / p->P, 2.0 / 7.0);
    p->th_bg = background_pot_temperature(c, p->x[1]);
    p->th_p = p->th - p->th_bg;
}
/* :220-223 */
static void w_reset_density(particle *p, const orc_system *s) {
    (void)s;
    p->rho = 0.0;
    p->rho_p = 0.0;
}
/* :226-228 */
static void w_compute_density(particle *p, const particle *q, double r, const orc_system *s) {
    p->rho += q->m * W(s, p->h, r);
}
/* :230-233 */
static void w_finalize_density(particle *p, const orc_system *s) {
    p->rho_bg = background_density(&s->prm, p->x[1]);
    p->rho_p = p->rho - p->rho_bg;
}
/* :235-238 */
static void w_update_smoothing(particle *p, const orc_system *s) {
    double rho = fmax(p->rho, s->prm.rho_floor);
    p->h = s->dim == 2 ? s->prm.eta * sqrt(p->m / rho) : s->prm.eta * cbrt(p->m / rho);
}
/* :245-251 — constant vector above z_t - z_b */
static inline double damping_y(const params *c, double z, int *active) {
    if (z >= (c->z_t - c->z_b)) {
        double sn = sin(M_PI / 2 * (1 - (c->z_t - c->z_b) / c->z_b));
        *active = 1;
        return -c->gamma_r * (sn * sn);
    }
    *active = 0;
    return 0.0;
}
/* :261-286 */
static void w_balance_of_momentum(particle *p, const particle *q, double r,
                                  const orc_system *s) {
    const params *c = &s->prm;
    double x_pq[3], v_pq[3];
    for (int a = 0; a < 3; ++a) { x_pq[a] = p->x[a] - q->x[a]; v_pq[a] = p->v[a] - q->v[a]; }
    double dot_product = dot3(x_pq, v_pq);
    double h_ij = 0.5 * (p->h + q->h);
    double ker = rDW(s, h_ij, r);
    double prho = fmax(p->rho, c->rho_floor);
    double qrho = fmax(q->rho, c->rho_floor);
    /* -q.m * (..) * ker * x_pq : left fold, scalar part first */
    double f = -q->m * (p->P_p / pow2(prho) + q->P_p / pow2(qrho)) * ker;
    for (int a = 0; a < 3; ++a) p->Dv[a] += f * x_pq[a];
    if (dot_product < 0.0) {
        double c_i = sqrt(c->gamma * p->P / prho);
        double c_j = sqrt(c->gamma * q->P / qrho);
        double c_ij = 0.5 * (c_i + c_j);
        double rho_ij = 0.5 * (prho + qrho);
        double mu_ij = (h_ij * dot_product) / (r * r + c->eps * h_ij * h_ij);
        double pi_ij = (-c->alpha * c_ij * mu_ij + c->beta * mu_ij * mu_ij) / rho_ij;
        double fv = -q->m * pi_ij * ker;
        for (int a = 0; a < 3; ++a) p->Dv[a] += fv * x_pq[a];
    }
}
/* :292-296 */
static void w_move(particle *p, const orc_system *s) {
    if (p->type == s->prm.fluid)
        for (int a = 0; a < 3; ++a) p->x[a] += s->prm.dt * p->v[a];
}
/* :298-303 with buyoancy_force :253-256.  VECY arithmetic is done per
 * component exactly as StaticArrays would: (-g*e_a)*rho_p/rho. */
static void w_accelerate(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    if (p->type == c->fluid) {
        static const double ey[3] = {0.0, 1.0, 0.0};
        int act;
        double dy = damping_y(c, p->x[1], &act);
        for (int a = 0; a < 3; ++a) {
            double buoy = -c->g * ey[a] * p->rho_p / p->rho;
            double damp = act ? dy * ey[a] : 0.0;
            p->v[a] += 0.5 * c->dt * (p->Dv[a] + buoy + damp);
        }
    }
    p->Dv[0] = p->Dv[1] = p->Dv[2] = 0.0;
}

/* ------------------------------------------------------------------------- */
/* Hopkins pressure-entropy variants (src/current/hopkins_*.jl)               */
/* ------------------------------------------------------------------------- */
/* hopkins_perturbed_witch.jl:200-203 */
static void h_reset_pressure(particle *p, const orc_system *s) {
    (void)s;
    p->P = 0.0;
    p->P_p = 0.0;
}
/* hopkins_perturbed_witch.jl:205-208 */
static void h_compute_pressure(particle *p, const particle *q, double r, const orc_system *s) {
    double ker = W(s, 0.5 * (p->h + q->h), r);
    p->P += q->m * pow(q->A, 1 / s->prm.gamma) * ker;
}
/* hopkins_perturbed_witch.jl:210-214 */
static void h_finalize_pressure(particle *p, const orc_system *s) {
    p->P = pow(p->P, s->prm.gamma);
    p->P_bg = background_pressure(&s->prm, p->x[1]);
    p->P_p = p->P - p->P_bg;
}
/* hopkins_total_witch.jl:170-172, :179-181 (11-field particle: no P_p, P_bg) */
static void ht_reset_pressure(particle *p, const orc_system *s) {
    (void)s;
    p->P = 0.0;
}
static void ht_finalize_pressure(particle *p, const orc_system *s) {
    p->P = pow(p->P, s->prm.gamma);
}
/* hopkins_total_witch.jl:187-193 */
static void ht_find_temperature(particle *p, const orc_system *s) {
    p->T = p->P / (s->prm.R_mass * p->rho);
}
static void ht_find_pot_temp(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    p->th = p->T * pow((c->T_bg * c->R_gas * c->rho0) / p->P, 2.0 / 7.0);
}
/* hopkins_total_witch.jl:203-205 */
static void ht_reset_density(particle *p, const orc_system *s) {
    (void)s;
    p->rho = 0.0;
}
/* shared AV block, hopkins_total_witch.jl:250-263 */
static inline void monaghan_av(particle *p, const particle *q, double r, const orc_system *s,
                               const double *x_pq, double dot_product) {
    const params *c = &s->prm;
    double h_ij = 0.5 * (p->h + q->h);
    double ker_ij = rDW(s, h_ij, r);
    double prho = fmax(p->rho, c->rho_floor);
    double qrho = fmax(q->rho, c->rho_floor);
    double c_i = sqrt(c->gamma * p->P / prho);
    double c_j = sqrt(c->gamma * q->P / qrho);
    double c_ij = 0.5 * (c_i + c_j);
    double rho_ij = 0.5 * (prho + qrho);
    double mu_ij = (h_ij * dot_product) / (r * r + c->eps * h_ij * h_ij);
    double pi_ij = (-c->alpha * c_ij * mu_ij + c->beta * mu_ij * mu_ij) / rho_ij;
    double fv = -q->m * pi_ij * ker_ij;
    for (int a = 0; a < 3; ++a) p->Dv[a] += fv * x_pq[a];
}
/* hopkins_total_witch.jl:233-264 */
static void ht_balance_of_momentum(particle *p, const particle *q, double r,
                                   const orc_system *s) {
    const params *c = &s->prm;
    double x_pq[3], v_pq[3];
    for (int a = 0; a < 3; ++a) { x_pq[a] = p->x[a] - q->x[a]; v_pq[a] = p->v[a] - q->v[a]; }
    double dot_product = dot3(x_pq, v_pq);
    double prefac = q->m * pow(p->A * q->A, 1 / c->gamma);
    double expfac = 1.0 - 2.0 / c->gamma;
    double ker_i = rDW(s, p->h, r);
    double ker_j = rDW(s, q->h, r);
    double pP = fmax(c->P_floor, p->P);
    double qP = fmax(c->P_floor, q->P);
    double f = -prefac * (pow(pP, expfac) * ker_i + pow(qP, expfac) * ker_j);
    for (int a = 0; a < 3; ++a) p->Dv[a] += f * x_pq[a];
    if (dot_product < 0.0) monaghan_av(p, q, r, s, x_pq, dot_product);
}
/* full_hopkins_perturbed_witch.jl:284-326 — total minus background pressure gradient */
static void hf_balance_of_momentum(particle *p, const particle *q, double r,
                                   const orc_system *s) {
    const params *c = &s->prm;
    double x_pq[3], v_pq[3];
    for (int a = 0; a < 3; ++a) { x_pq[a] = p->x[a] - q->x[a]; v_pq[a] = p->v[a] - q->v[a]; }
    double dot_product = dot3(x_pq, v_pq);
    double prefac = q->m * pow(p->A * q->A, 1 / c->gamma);
    double expfac = 1.0 - 2.0 / c->gamma;
    double ker_i = rDW(s, p->h, r);
    double ker_j = rDW(s, q->h, r);
    double pP = fmax(c->P_floor, p->P);
    double qP = fmax(c->P_floor, q->P);
    double f_tot = -prefac * (pow(pP, expfac) * ker_i + pow(qP, expfac) * ker_j);
    double prefac_bg = q->m * pow(p->A_bg * q->A_bg, 1 / c->gamma);
    double pP_bg = fmax(c->P_floor, p->P_bg);
    double qP_bg = fmax(c->P_floor, q->P_bg);
    double f_bg = -prefac_bg * (pow(pP_bg, expfac) * ker_i + pow(qP_bg, expfac) * ker_j);
    /* p.Dv += a_tot - a_bg  (vectors) */
    for (int a = 0; a < 3; ++a) p->Dv[a] += f_tot * x_pq[a] - f_bg * x_pq[a];
    if (dot_product < 0.0) monaghan_av(p, q, r, s, x_pq, dot_product);
}
/* hopkins_total_witch.jl:270-277 — NOT type-gated; gravity :225-228 */
static void ht_move(particle *p, const orc_system *s) {
    for (int a = 0; a < 3; ++a) p->x[a] += s->prm.dt * p->v[a];
}
static void ht_accelerate(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    static const double ey[3] = {0.0, 1.0, 0.0};
    int act;
    double dy = damping_y(c, p->x[1], &act);
    for (int a = 0; a < 3; ++a) {
        double grav = -c->g * ey[a];
        double damp = act ? dy * ey[a] : 0.0;
        p->v[a] += 0.5 * c->dt * (p->Dv[a] + grav + damp);
    }
    p->Dv[0] = p->Dv[1] = p->Dv[2] = 0.0;
}

/* ------------------------------------------------------------------------- */
/* sph_jl/examples/collapse_dry.jl (BASELINE config 1)                        */
/* ------------------------------------------------------------------------- */
/* :112-115 */
static void d_balance_of_mass(particle *p, const particle *q, double r, const orc_system *s) {
    const params *c = &s->prm;
    double ker = c->m * orc_rDwendland2(c->kh, r);
    double x_pq[3], v_pq[3];
    for (int a = 0; a < 3; ++a) { x_pq[a] = p->x[a] - q->x[a]; v_pq[a] = p->v[a] - q->v[a]; }
    p->Drho += ker * (dot3(x_pq, v_pq) + 2 * c->nu * (p->rho - q->rho));
}
/* :123-127 */
static void d_find_pressure(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    p->rho += p->Drho * c->dt;
    p->Drho = 0.0;
    p->P = pow2(c->c) * (p->rho - c->rho0);
}
/* :135-141 */
static void d_internal_force(particle *p, const particle *q, double r, const orc_system *s) {
    const params *c = &s->prm;
    if (p->type == c->fluid) {
        double ker = c->m * orc_rDwendland2(c->kh, r);
        double a1 = -ker * (p->P / pow2(p->rho) + q->P / pow2(q->rho));
        for (int a = 0; a < 3; ++a) p->Dv[a] += a1 * (p->x[a] - q->x[a]);
        double a2 = +2 * ker * c->mu / pow2(c->rho0);
        for (int a = 0; a < 3; ++a) p->Dv[a] += a2 * (p->v[a] - q->v[a]);
    }
}
/* :148-153 */
static void d_move(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    p->Dv[0] = p->Dv[1] = p->Dv[2] = 0.0;
    if (p->type == c->fluid)
        for (int a = 0; a < 3; ++a) p->x[a] += 0.5 * c->dt * p->v[a];
}
/* :155-159 */
static void d_accelerate(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    const double g[3] = {c->gx, c->gy, c->gz};
    if (p->type == c->fluid)
        for (int a = 0; a < 3; ++a) p->v[a] += 0.5 * c->dt * (p->Dv[a] + g[a]);
}

/* ------------------------------------------------------------------------- */
/* sph_jl/tests/test_collision_2d.jl (the reference's only integration test)  */
/* field `a` of that Particle is stored in Dv here.                           */
/* ------------------------------------------------------------------------- */
/* :66-68 */
static void c_find_rho(particle *p, const particle *q, double r, const orc_system *s) {
    (void)q;
    p->rho += s->prm.m * orc_wendland2(s->prm.kh, r);
}
/* :70-72 */
static void c_find_rho0(particle *p, const particle *q, double r, const orc_system *s) {
    (void)q;
    p->rho0 += s->prm.m * orc_wendland2(s->prm.kh, r);
}
/* :74-76 */
static void c_find_pressure(particle *p, const orc_system *s) {
    p->P = pow2(s->prm.c) * (p->rho - p->rho0);
}
/* :78-81 */
static void c_internal_force(particle *p, const particle *q, double r, const orc_system *s) {
    const params *c = &s->prm;
    double ker = c->m * orc_rDwendland2(c->kh, r);
    double f = -ker * (p->P / pow2(c->rho0) + q->P / pow2(c->rho0));
    for (int a = 0; a < 3; ++a) p->Dv[a] += f * (p->x[a] - q->x[a]);
}
/* :83-89 */
static void c_reset_a(particle *p, const orc_system *s) {
    (void)s;
    p->Dv[0] = p->Dv[1] = p->Dv[2] = 0.0;
}
static void c_reset_rho(particle *p, const orc_system *s) {
    (void)s;
    p->rho = 0.0;
}
/* :91-97 */
static void c_move(particle *p, const orc_system *s) {
    for (int a = 0; a < 3; ++a) p->x[a] += s->prm.dt * p->v[a];
}
static void c_accelerate(particle *p, const orc_system *s) {
    for (int a = 0; a < 3; ++a) p->v[a] += 0.5 * s->prm.dt * p->Dv[a];
}

/* ------------------------------------------------------------------------- */
/* src/utils/new_packing.jl:5-60                                              */
/* ------------------------------------------------------------------------- */
static void p_reset_rho(particle *p, const orc_system *s) {
    if (p->type == s->prm.fluid) p->rho = 0.0;
}
static void p_accumulate_rho(particle *p, const particle *q, double r, const orc_system *s) {
    if (p->type == s->prm.fluid) p->rho += q->m * W(s, p->h, r);
}
static void p_balance_of_momentum(particle *p, const particle *q, double r,
                                  const orc_system *s) {
    const params *c = &s->prm;
    if (p->type == c->fluid && q->type == c->fluid) {
        double x_pq1 = p->x[1] - q->x[1];
        double rho_i = fmax(p->rho, c->rho_floor);
        double rho_j = fmax(q->rho, c->rho_floor);
        double rho_ti = background_density(c, p->x[1]);
        double rho_tj = background_density(c, q->x[1]);
        double Pi = pow2(c->c_pack) * (rho_i - rho_ti);
        double Pj = pow2(c->c_pack) * (rho_j - rho_tj);
        double ker = rDW(s, 0.5 * (p->h + q->h), r);
        double f1 = -q->m * (Pi / pow2(rho_i) + Pj / pow2(rho_j)) * ker * x_pq1;
        /* p.Dv += f[2]*VECY : x and z receive f[2]*0.0 */
        p->Dv[0] += f1 * 0.0;
        p->Dv[1] += f1 * 1.0;
        p->Dv[2] += f1 * 0.0;
    }
}
static void p_accelerate(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    if (p->type == c->fluid)
        for (int a = 0; a < 3; ++a)
            p->v[a] = (p->v[a] + c->dt_pack * p->Dv[a]) / (1.0 + c->zeta_pack * c->dt_pack);
    p->Dv[0] = p->Dv[1] = p->Dv[2] = 0.0;
}
static void p_move(particle *p, const orc_system *s) {
    if (p->type == s->prm.fluid)
        for (int a = 0; a < 3; ++a) p->x[a] += s->prm.dt_pack * p->v[a];
}

/* ------------------------------------------------------------------------- */
/* src/legacy/isothermal_flow_witch.jl — constant-U flow over the mountain    */
/* with inflow re-seeding.  Field u is stored in v, Du in Dv.  */
/* T is the driver's constant temperature (params.T_bg), h its fixed kernel   */
/* radius (params.kh).                                                        */
/* ------------------------------------------------------------------------- */
/* :140-143 */
static void f_balance_of_mass(particle *p, const particle *q, double r, const orc_system *s) {
    double ker = q->m * orc_rDwendland2(s->prm.kh, r);
    double x_pq[3], u_pq[3];
    for (int a = 0; a < 3; ++a) { x_pq[a] = p->x[a] - q->x[a]; u_pq[a] = p->v[a] - q->v[a]; }
    p->Drho += ker * dot3(x_pq, u_pq);
}
/* :145-150 */
static void f_internal_force(particle *p, const particle *q, double r, const orc_system *s) {
    const params *c = &s->prm;
    double ker = q->m * orc_rDwendland2(c->kh, r);
    double x_pq[3], u_pq[3];
    for (int a = 0; a < 3; ++a) { x_pq[a] = p->x[a] - q->x[a]; u_pq[a] = p->v[a] - q->v[a]; }
    double a1 = -ker * (p->P / pow2(p->rho) + q->P / pow2(q->rho));
    for (int a = 0; a < 3; ++a) p->Dv[a] += a1 * x_pq[a];
    double a2 = 8.0 * ker * c->mu / (p->rho * q->rho) * dot3(u_pq, x_pq) / (r * r + 0.01 * c->kh * c->kh);
    for (int a = 0; a < 3; ++a) p->Dv[a] += a2 * x_pq[a];
}
/* :156-160 */
static void f_find_pressure(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    p->rho += p->Drho * c->dt;
    p->Drho = 0.0;
    p->P = p->rho * c->R_mass * c->T_bg;
}
/* :162-164 */
static void f_set_density(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    p->rho = c->rho0 * exp(-p->x[1] * c->g / (c->R_mass * c->T_bg));
}
/* :167-169 */
static void f_find_pot_temp(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    p->th = c->T_bg * pow((c->T_bg * c->R_gas * c->rho0) / p->P, c->R_gas / c->cp);
}
/* :204-209 */
static void f_move(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    p->Dv[0] = p->Dv[1] = p->Dv[2] = 0.0;
    if (p->type == c->fluid || p->type == c->inflow)
        for (int a = 0; a < 3; ++a) p->x[a] += c->dt * p->v[a];
}
/* :211-215 with damping_structure :192-198 (a positive scalar here) */
static void f_accelerate(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    if (p->type == c->fluid) {
        static const double ey[3] = {0.0, 1.0, 0.0};
        double damp = 0.0;
        if (p->x[1] >= (c->z_t - c->z_b)) {
            double sn = sin(M_PI / 2 * (1 - (c->z_t - c->z_b) / c->z_b));
            damp = c->gamma_r * (sn * sn);
        }
        for (int a = 0; a < 3; ++a) p->v[a] += 0.5 * c->dt * (p->Dv[a] - c->g * ey[a] - damp * ey[a]);
    }
}
/* :175-186 add_new_particles! with the Particle constructor :72-82 */
int64_t orc_flow_add_new_particles(orc_system *s) {
    const params *c = &s->prm;
    int64_t n0 = s->n, added = 0;
    for (int64_t i = 0; i < n0; ++i) {
        particle *p = s->p[i];
        if (p->type == c->inflow && p->x[0] >= c->x_inflow) {
            p->type = c->fluid;
            orc_append(s, 1);
            particle *q = s->p[s->n - 1];
            static const double ex[3] = {1.0, 0.0, 0.0};
            for (int a = 0; a < 3; ++a) {
                q->x[a] = p->x[a] - c->bc_width * ex[a];
                q->v[a] = c->U_max * ex[a];
            }
            q->type = c->inflow;
            q->rho = c->rho0 * exp(-q->x[1] * c->g / (c->R_mass * c->T_bg));
            q->m = q->rho * pow2(c->dr);
            q->P = q->rho * c->T_bg * c->R_mass;
            q->th = c->T_bg * pow((c->T_bg * c->R_gas * c->rho0) / q->P, c->R_gas / c->cp);
            ++added;
        }
    }
    return added;
}

/* ------------------------------------------------------------------------- */
/* src/legacy/adiabatic_flow_witch.jl — the same flow with adiabatic           */
/* thermodynamics: entropy S is carried per particle, T and P follow from rho  */
/* and the entropy density s.  u is stored in v, Du in Dv, T0 = params.T_bg,   */
/* h = params.kh, cv = cp - R_mass (:50), gamma = cp/cv (:51).                 */
/* accelerate! (:225-229) and internal_force! (:146-153) are the isothermal    */
/* driver's, character for character: f_accelerate, f_internal_force.          */
/* ------------------------------------------------------------------------- */
/* :159-163 (applied with self = true, :238) */
static void af_find_density(particle *p, const particle *q, double r, const orc_system *s) {
    const params *c = &s->prm;
    if (p->type == c->fluid && q->type == c->fluid) p->rho += q->m * orc_wendland2(c->kh, r);
}
/* :165-169 */
static void af_find_s(particle *p, const orc_system *s) {
    if (p->type == s->prm.fluid) p->s = p->S * p->rho / p->m;
}
/* :171-176 */
static void af_find_pressure(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    if (p->type == c->fluid) {
        const double cv = c->cp - c->R_mass;
        p->T = (pow(p->rho, c->gamma - 1.0)) * exp(p->s / (p->rho * cv)) / (cv * (c->gamma - 1.0));
        p->P = c->R_mass * p->rho * p->T;
    }
}
/* :178-182 */
static void af_find_pot_temp(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    if (p->type == c->fluid) {
        double b = (c->T_bg * c->R_gas * c->rho0) / p->P;
        p->th = p->T * pow(b * b, 1.0 / 7.0);
    }
}
/* :184-191 */
static void af_entropy_production(particle *p, const particle *q, double r, const orc_system *s) {
    const params *c = &s->prm;
    if (p->type == c->fluid && q->type == c->fluid) {
        double ker = orc_rDwendland2(c->kh, r);
        double x_pq[3], u_pq[3];
        for (int a = 0; a < 3; ++a) { x_pq[a] = p->x[a] - q->x[a]; u_pq[a] = p->v[a] - q->v[a]; }
        double d = dot3(u_pq, x_pq);
        p->S += -4.0 * p->m * q->m * ker * c->mu / (p->T * p->rho * q->rho) * (d * d) /
                (r * r + 0.01 * c->kh * c->kh) * c->dt;
    }
}
/* :217-223 */
static void af_move(particle *p, const orc_system *s) {
    const params *c = &s->prm;
    p->Dv[0] = p->Dv[1] = p->Dv[2] = 0.0;
    if (p->type == c->fluid) {
        for (int a = 0; a < 3; ++a) p->x[a] += c->dt * p->v[a];
        p->rho = 0.0;
    }
}
/* the Particle constructor :82-91 */
static void af_construct(particle *q, const params *c) {
    const double cv = c->cp - c->R_mass;
    q->T = c->T_bg;
    q->rho = c->rho0 * exp(-q->x[1] * c->g / (c->R_mass * q->T));
    q->m = q->rho * pow2(c->dr);
    q->P = c->R_mass * q->T * q->rho;
    double b = (c->T_bg * c->R_gas * c->rho0) / q->P;
    q->th = q->T * pow(b * b, 1.0 / 7.0);
    q->S = q->m * cv * log((cv * q->T * (c->gamma - 1.0)) / (c->gamma * pow(q->rho, c->gamma - 1.0)));
}
/* :197-208 add_new_particles! */
int64_t orc_aflow_add_new_particles(orc_system *s) {
    const params *c = &s->prm;
    int64_t n0 = s->n, added = 0;
    for (int64_t i = 0; i < n0; ++i) {
        particle *p = s->p[i];
        if (p->type == c->inflow && p->x[0] >= c->x_inflow) {
            p->type = c->fluid;
            orc_append(s, 1);
            particle *q = s->p[s->n - 1];
            static const double ex[3] = {1.0, 0.0, 0.0};
            for (int a = 0; a < 3; ++a) {
                q->x[a] = p->x[a] - c->bc_width * ex[a];
                q->v[a] = c->U_max * ex[a];
            }
            q->type = c->inflow;
            af_construct(q, c);
            ++added;
        }
    }
    return added;
}
/* the constructor applied to every particle of a freshly generated system (make_system, :106-110) */
void orc_aflow_construct_all(orc_system *s) {
    for (int64_t i = 0; i < s->n; ++i) af_construct(s->p[i], &s->prm);
}

/* ------------------------------------------------------------------------- */
/* operator menu (names shared with the product's enum so tests read alike)   */
/* ------------------------------------------------------------------------- */
typedef struct {
    const char *name;
    unary_fn u;
    binary_fn b;
} op_desc;
static const op_desc OPS[] = {
    {"wcsph.accelerate", w_accelerate, NULL},
    {"wcsph.move", w_move, NULL},
    {"wcsph.reset_density", w_reset_density, NULL},
    {"wcsph.compute_density", NULL, w_compute_density},
    {"wcsph.finalize_density", w_finalize_density, NULL},
    {"wcsph.update_smoothing", w_update_smoothing, NULL},
    {"wcsph.compute_pressure", w_compute_pressure, NULL},
    {"wcsph.find_temperature", w_find_temperature, NULL},
    {"wcsph.find_pot_temp", w_find_pot_temp, NULL},
    {"wcsph.balance_of_momentum", NULL, w_balance_of_momentum},
    {"hopkins.reset_pressure", h_reset_pressure, NULL},
    {"hopkins.compute_pressure", NULL, h_compute_pressure},
    {"hopkins.finalize_pressure", h_finalize_pressure, NULL},
    {"hopkins_total.reset_pressure", ht_reset_pressure, NULL},
    {"hopkins_total.finalize_pressure", ht_finalize_pressure, NULL},
    {"hopkins_total.find_temperature", ht_find_temperature, NULL},
    {"hopkins_total.find_pot_temp", ht_find_pot_temp, NULL},
    {"hopkins_total.reset_density", ht_reset_density, NULL},
    {"hopkins_total.balance_of_momentum", NULL, ht_balance_of_momentum},
    {"hopkins_full.balance_of_momentum", NULL, hf_balance_of_momentum},
    {"hopkins_total.move", ht_move, NULL},
    {"hopkins_total.accelerate", ht_accelerate, NULL},
    {"dambreak.balance_of_mass", NULL, d_balance_of_mass},
    {"dambreak.find_pressure", d_find_pressure, NULL},
    {"dambreak.internal_force", NULL, d_internal_force},
    {"dambreak.move", d_move, NULL},
    {"dambreak.accelerate", d_accelerate, NULL},
    {"collision.find_rho", NULL, c_find_rho},
    {"collision.find_rho0", NULL, c_find_rho0},
    {"collision.find_pressure", c_find_pressure, NULL},
    {"collision.internal_force", NULL, c_internal_force},
    {"collision.reset_a", c_reset_a, NULL},
    {"collision.reset_rho", c_reset_rho, NULL},
    {"collision.move", c_move, NULL},
    {"collision.accelerate", c_accelerate, NULL},
    {"flow.balance_of_mass", NULL, f_balance_of_mass},
    {"flow.internal_force", NULL, f_internal_force},
    {"flow.find_pressure", f_find_pressure, NULL},
    {"flow.set_density", f_set_density, NULL},
    {"flow.find_pot_temp", f_find_pot_temp, NULL},
    {"flow.move", f_move, NULL},
    {"flow.accelerate", f_accelerate, NULL},
    {"aflow.find_density", NULL, af_find_density},
    {"aflow.find_s", af_find_s, NULL},
    {"aflow.find_pressure", af_find_pressure, NULL},
    {"aflow.find_pot_temp", af_find_pot_temp, NULL},
    {"aflow.entropy_production", NULL, af_entropy_production},
    {"aflow.move", af_move, NULL},
    {"packing.reset_rho", p_reset_rho, NULL},
    {"packing.accumulate_rho", NULL, p_accumulate_rho},
    {"packing.balance_of_momentum", NULL, p_balance_of_momentum},
    {"packing.accelerate", p_accelerate, NULL},
    {"packing.move", p_move, NULL},
    {NULL, NULL, NULL}};

typedef struct {
    orc_system *s;
    binary_fn b;
} self_ctx;

/* core.jl:151-161 — apply!: binary if f has a (T,T,Float64) method; `self`
 * adds f(p,p,0.0) as a second, unary sweep. */
int orc_apply(orc_system *s, const char *name, int self) {
    for (const op_desc *d = OPS; d->name; ++d)
        if (!strcmp(d->name, name)) {
            if (d->b) {
                apply_binary(s, d->b);
                if (self) {
                    binary_fn b = d->b;
#pragma omp parallel for schedule(static)
                    for (int64_t i = 0; i < s->n; ++i) b(s->p[i], s->p[i], 0.0, s);
                }
            } else {
                apply_unary(s, d->u);
            }
            return 0;
        }
    return -1;
}

/* wcsph_perturbed_witch.jl:309-332 */
static void verlet_wcsph(orc_system *s) {
    apply_unary(s, w_accelerate);
    apply_unary(s, w_move);
    orc_create_cell_list(s);
    apply_unary(s, w_reset_density);
    apply_binary(s, w_compute_density);
    apply_unary(s, w_finalize_density);
    apply_unary(s, w_update_smoothing);
    orc_create_cell_list(s);
    apply_unary(s, w_compute_pressure);
    apply_unary(s, w_find_temperature);
    apply_unary(s, w_find_pot_temp);
    apply_binary(s, w_balance_of_momentum);
    apply_unary(s, w_accelerate);
}
/* hopkins_perturbed_witch.jl:325-349 */
static void verlet_hopkins(orc_system *s) {
    apply_unary(s, w_accelerate);
    apply_unary(s, w_move);
    orc_create_cell_list(s);
    apply_unary(s, w_reset_density);
    apply_binary(s, w_compute_density);
    apply_unary(s, w_finalize_density);
    apply_unary(s, w_update_smoothing);
    orc_create_cell_list(s);
    apply_unary(s, h_reset_pressure);
    apply_binary(s, h_compute_pressure);
    apply_unary(s, h_finalize_pressure);
    apply_unary(s, w_find_temperature);
    apply_unary(s, w_find_pot_temp);
    apply_binary(s, w_balance_of_momentum);
    apply_unary(s, w_accelerate);
}
/* full_hopkins_perturbed_witch.jl:350-374 (same sequence, its own momentum closure) */
static void verlet_hopkins_full(orc_system *s) {
    apply_unary(s, w_accelerate);
    apply_unary(s, w_move);
    orc_create_cell_list(s);
    apply_unary(s, w_reset_density);
    apply_binary(s, w_compute_density);
    apply_unary(s, w_finalize_density);
    apply_unary(s, w_update_smoothing);
    orc_create_cell_list(s);
    apply_unary(s, h_reset_pressure);
    apply_binary(s, h_compute_pressure);
    apply_unary(s, h_finalize_pressure);
    apply_unary(s, w_find_temperature);
    apply_unary(s, w_find_pot_temp);
    apply_binary(s, hf_balance_of_momentum);
    apply_unary(s, w_accelerate);
}
/* hopkins_total_witch.jl:283-308 */
static void verlet_hopkins_total(orc_system *s) {
    apply_unary(s, ht_accelerate);
    apply_unary(s, ht_move);
    orc_create_cell_list(s);
    apply_unary(s, ht_reset_density);
    apply_binary(s, w_compute_density);
    apply_unary(s, w_update_smoothing);
    orc_create_cell_list(s);
    apply_unary(s, ht_reset_pressure);
    apply_binary(s, h_compute_pressure);
    apply_unary(s, ht_finalize_pressure);
    apply_unary(s, ht_find_temperature);
    apply_unary(s, ht_find_pot_temp);
    apply_binary(s, ht_balance_of_momentum);
    apply_unary(s, ht_accelerate);
}
/* collapse_dry.jl:203-211 */
static void step_dambreak(orc_system *s) {
    apply_unary(s, d_accelerate);
    apply_unary(s, d_move);
    orc_create_cell_list(s);
    apply_binary(s, d_balance_of_mass);
    apply_unary(s, d_find_pressure);
    apply_unary(s, d_move);
    orc_create_cell_list(s);
    apply_binary(s, d_internal_force);
    apply_unary(s, d_accelerate);
}
/* isothermal_flow_witch.jl:221-232 */
static void step_flow(orc_system *s) {
    apply_unary(s, f_accelerate);
    apply_unary(s, f_move);
    orc_flow_add_new_particles(s);
    orc_create_cell_list(s);
    apply_binary(s, f_balance_of_mass);
    apply_unary(s, f_find_pressure);
    apply_unary(s, f_find_pot_temp);
    apply_binary(s, f_internal_force);
    apply_unary(s, f_accelerate);
}
/* adiabatic_flow_witch.jl:231-243 */
static void step_aflow(orc_system *s) {
    apply_unary(s, f_accelerate);
    apply_unary(s, af_move);
    orc_aflow_add_new_particles(s);
    orc_create_cell_list(s);
    orc_apply(s, "aflow.find_density", 1);
    apply_unary(s, af_find_s);
    apply_unary(s, af_find_pressure);
    apply_binary(s, af_entropy_production);
    apply_binary(s, f_internal_force);
    apply_unary(s, f_accelerate);
}
/* test_collision_2d.jl:106-116 */
static void step_collision(orc_system *s) {
    apply_unary(s, c_accelerate);
    apply_unary(s, c_move);
    orc_create_cell_list(s);
    apply_unary(s, c_reset_rho);
    orc_apply(s, "collision.find_rho", 1);
    apply_unary(s, c_find_pressure);
    apply_unary(s, c_reset_a);
    apply_binary(s, c_internal_force);
    apply_unary(s, c_accelerate);
}

int orc_step(orc_system *s, const char *scheme, int nsteps) {
    void (*f)(orc_system *) = NULL;
    if (!strcmp(scheme, "wcsph")) f = verlet_wcsph;
    else if (!strcmp(scheme, "hopkins")) f = verlet_hopkins;
    else if (!strcmp(scheme, "hopkins_total")) f = verlet_hopkins_total;
    else if (!strcmp(scheme, "hopkins_full")) f = verlet_hopkins_full;
    else if (!strcmp(scheme, "dambreak")) f = step_dambreak;
    else if (!strcmp(scheme, "collision")) f = step_collision;
    else if (!strcmp(scheme, "flow")) f = step_flow;
    else if (!strcmp(scheme, "aflow")) f = step_aflow;
    if (!f) return -1;
    for (int k = 0; k < nsteps; ++k) f(s);
    return 0;
}
