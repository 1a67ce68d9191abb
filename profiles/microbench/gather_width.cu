// Micro-benchmark for DESIGN.md §7: how should the replay kernels (csrc/pair_list.cuh
// k_binary_list) fetch the fields of a neighbour?  They are bound by the L1 data pipe: one
// 128-byte line per cycle per SM, and a warp-wide gather of ONE FP64 field touches ~11.5 lines
// (profiles/r01b_pair_list.md).  The force closure needs 11 fields per accepted pair.
//
// This program builds the real access pattern — a cubic lattice, h = 1.8 dr, particles sorted by
// cell (x fastest), per-particle neighbour lists in the reference's traversal order, the 32
// particles of a warp interleaved exactly like PairList — and times four ways of reading 11 (or
// 12) doubles per entry:
//   soa64    11 separate arrays, LDG.64 each                     (what k_binary_list does today)
//   pair128  6 arrays of double2, LDG.128 each
//   rec128   3 arrays of 32-byte records, two LDG.128 per record
//   rec256   3 arrays of 32-byte records, one 256-bit load per record (sm_100: LDG.E.256)
// Every variant adds up what it loads (the sum is checked across variants), so nothing is
// optimised away and the FP64 work is negligible.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_width gather_width.cu
//   ./gather_width [nx ny nz]          (default 160 120 100 = 1.92 M particles)
//   ncu --set full -k regex:k_ ./gather_width 96 64 64      (wavefronts per variant)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>
#include <vector>

#define CHECK(x)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (x);                                                             \
        if (e_ != cudaSuccess) {                                                          \
            fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                      \
        }                                                                                 \
    } while (0)

static const int STRIDE = 32;  // list entries per particle (26 on the undisturbed lattice)
static const int NF = 12;      // doubles per particle (11 used by the force closure + 1 pad)

struct __align__(32) Rec {
    double a, b, c, d;
};

__device__ __forceinline__ Rec ld256(const Rec *p) {
    Rec r;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
    return r;
}

struct Soa {
    const double *f[NF];
};

// list[((p >> 5) * STRIDE + k) * 32 + (p & 31)], cnt[p]
#define LIST_LOOP_BEGIN                                                                  \
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;                    \
    if (p >= n) return;                                                                  \
    const uint32_t c = cnt[p];                                                           \
    const uint32_t *row = list + ((size_t)(p >> 5) * STRIDE) * 32 + (size_t)(p & 31);    \
    double acc = 0.0;                                                                    \
    uint32_t qn = c ? row[0] : 0u;                                                       \
    for (uint32_t k = 0; k < c; ++k) {                                                   \
        const uint32_t q = qn;                                                           \
        if (k + 1 < c) qn = row[(size_t)(k + 1) * 32];
#define LIST_LOOP_END \
    }                 \
    out[p] = acc;

__global__ void __launch_bounds__(128) k_soa64(Soa s, const uint32_t *__restrict__ list,
                                               const uint32_t *__restrict__ cnt, double *out, int64_t n) {
    LIST_LOOP_BEGIN
#pragma unroll
    for (int f = 0; f < 11; ++f) acc += s.f[f][q];
    LIST_LOOP_END
}

__global__ void __launch_bounds__(128) k_pair128(const double2 *const *__restrict__ arr_unused, const double2 *a0,
                                                 const double2 *a1, const double2 *a2, const double2 *a3,
                                                 const double2 *a4, const double2 *a5,
                                                 const uint32_t *__restrict__ list,
                                                 const uint32_t *__restrict__ cnt, double *out, int64_t n) {
    LIST_LOOP_BEGIN
    const double2 v0 = a0[q], v1 = a1[q], v2 = a2[q], v3 = a3[q], v4 = a4[q], v5 = a5[q];
    acc += v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y + v4.x + v4.y + v5.x;  // 11 of 12
    LIST_LOOP_END
}

__global__ void __launch_bounds__(128) k_rec128(const Rec *ra, const Rec *rb, const Rec *rc,
                                                const uint32_t *__restrict__ list,
                                                const uint32_t *__restrict__ cnt, double *out, int64_t n) {
    LIST_LOOP_BEGIN
    const double2 *pa = (const double2 *)(ra + q), *pb = (const double2 *)(rb + q), *pc = (const double2 *)(rc + q);
    const double2 a0 = pa[0], a1 = pa[1], b0 = pb[0], b1 = pb[1], c0 = pc[0], c1 = pc[1];
    acc += a0.x + a0.y + a1.x + a1.y + b0.x + b0.y + b1.x + b1.y + c0.x + c0.y + c1.x;
    LIST_LOOP_END
}

__global__ void __launch_bounds__(128) k_rec256(const Rec *ra, const Rec *rb, const Rec *rc,
                                                const uint32_t *__restrict__ list,
                                                const uint32_t *__restrict__ cnt, double *out, int64_t n) {
    LIST_LOOP_BEGIN
    const Rec a = ld256(ra + q), b = ld256(rb + q), cc = ld256(rc + q);
    acc += a.a + a.b + a.c + a.d + b.a + b.b + b.c + b.d + cc.a + cc.b + cc.c;
    LIST_LOOP_END
}

int main(int argc, char **argv) {
    int nx = 160, ny = 120, nz = 100;
    if (argc == 4) {
        nx = atoi(argv[1]);
        ny = atoi(argv[2]);
        nz = atoi(argv[3]);
    }
    const double dr = 1.0, h = 1.8 * dr;
    const int64_t n = (int64_t)nx * ny * nz;
    // ---- lattice, cell keys (x fastest), cell-sorted order ---------------------------------
    const int lx = (int)floor((nx - 1) * dr / h) + 1, ly = (int)floor((ny - 1) * dr / h) + 1,
              lz = (int)floor((nz - 1) * dr / h) + 1;
    std::vector<double> x(n), y(n), z(n);
    std::vector<uint32_t> key(n), order(n);
    for (int64_t i = 0; i < n; ++i) {
        const int ix = (int)(i % nx), iy = (int)((i / nx) % ny), iz = (int)(i / ((int64_t)nx * ny));
        x[i] = ix * dr;
        y[i] = iy * dr;
        z[i] = iz * dr;
        key[i] = (uint32_t)((int)floor(x[i] / h) + lx * ((int)floor(y[i] / h) + ly * (int)floor(z[i] / h)));
    }
    std::iota(order.begin(), order.end(), 0u);
    // (cell ascending, index descending) — the physical order of the library
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return key[a] != key[b] ? key[a] < key[b] : a > b;
    });
    const int64_t ncell = (int64_t)lx * ly * lz;
    std::vector<uint32_t> cell_start(ncell + 1, 0);
    for (int64_t i = 0; i < n; ++i) cell_start[key[i] + 1]++;
    for (int64_t c = 0; c < ncell; ++c) cell_start[c + 1] += cell_start[c];
    std::vector<double> sx(n), sy(n), sz(n);
    std::vector<uint32_t> skey(n);
    for (int64_t pos = 0; pos < n; ++pos) {
        sx[pos] = x[order[pos]];
        sy[pos] = y[order[pos]];
        sz[pos] = z[order[pos]];
        skey[pos] = key[order[pos]];
    }
    // ---- neighbour lists in traversal order (di outermost, structs.jl:73-81) -----------------
    const int64_t nwarps = (n + 31) / 32;
    std::vector<uint32_t> list((size_t)nwarps * STRIDE * 32, 0u), cnt(n, 0u);
    int64_t total = 0;
    uint32_t cmax = 0;
    for (int64_t p = 0; p < n; ++p) {
        const int ci = (int)(skey[p] % lx), cj = (int)((skey[p] / lx) % ly), ck = (int)(skey[p] / ((int64_t)lx * ly));
        uint32_t c = 0;
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj)
                for (int dk = -1; dk <= 1; ++dk) {
                    const int i = ci + di, j = cj + dj, k = ck + dk;
                    if (i < 0 || i >= lx || j < 0 || j >= ly || k < 0 || k >= lz) continue;
                    const int64_t cell = i + (int64_t)lx * (j + (int64_t)ly * k);
                    for (uint32_t q = cell_start[cell]; q < cell_start[cell + 1]; ++q) {
                        const double dx = sx[p] - sx[q], dy = sy[p] - sy[q], dz = sz[p] - sz[q];
                        if (dx * dx + dy * dy + dz * dz > h * h || q == p) continue;
                        if (c < (uint32_t)STRIDE) list[((size_t)(p >> 5) * STRIDE + c) * 32 + (p & 31)] = q;
                        ++c;
                    }
                }
        cnt[p] = std::min<uint32_t>(c, STRIDE);
        cmax = std::max(cmax, c);
        total += cnt[p];
    }
    printf("particles %lld  cells %dx%dx%d  entries/particle %.2f (max %u)\n", (long long)n, lx, ly, lz,
           (double)total / n, cmax);
    // ---- fields -----------------------------------------------------------------------------
    std::vector<double> soa((size_t)NF * n);
    for (int f = 0; f < NF; ++f)
        for (int64_t p = 0; p < n; ++p) soa[(size_t)f * n + p] = (double)((p * 7 + f * 13) % 1000) * 1e-3;
    std::vector<double> pairs((size_t)NF * n), recs((size_t)NF * n);
    for (int64_t p = 0; p < n; ++p)
        for (int f = 0; f < NF; ++f) {
            pairs[(size_t)(f / 2) * 2 * n + 2 * p + (f % 2)] = soa[(size_t)f * n + p];  // 6 arrays of double2
            recs[(size_t)(f / 4) * 4 * n + 4 * p + (f % 4)] = soa[(size_t)f * n + p];   // 3 arrays of Rec
        }
    double *d_soa, *d_pairs, *d_recs, *d_out;
    uint32_t *d_list, *d_cnt;
    CHECK(cudaMalloc(&d_soa, sizeof(double) * NF * n));
    CHECK(cudaMalloc(&d_pairs, sizeof(double) * NF * n));
    CHECK(cudaMalloc(&d_recs, sizeof(double) * NF * n));
    CHECK(cudaMalloc(&d_out, sizeof(double) * n));
    CHECK(cudaMalloc(&d_list, sizeof(uint32_t) * list.size()));
    CHECK(cudaMalloc(&d_cnt, sizeof(uint32_t) * n));
    CHECK(cudaMemcpy(d_soa, soa.data(), sizeof(double) * NF * n, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d_pairs, pairs.data(), sizeof(double) * NF * n, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d_recs, recs.data(), sizeof(double) * NF * n, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d_list, list.data(), sizeof(uint32_t) * list.size(), cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d_cnt, cnt.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice));
    Soa s;
    for (int f = 0; f < NF; ++f) s.f[f] = d_soa + (size_t)f * n;
    const double2 *pa[6];
    for (int f = 0; f < 6; ++f) pa[f] = (const double2 *)(d_pairs + (size_t)f * 2 * n);
    const Rec *ra = (const Rec *)d_recs, *rb = (const Rec *)(d_recs + (size_t)4 * n), *rc = (const Rec *)(d_recs + (size_t)8 * n);

    const unsigned blocks = (unsigned)((n + 127) / 128);
    std::vector<double> out(n);
    auto checksum = [&]() {
        CHECK(cudaMemcpy(out.data(), d_out, sizeof(double) * n, cudaMemcpyDeviceToHost));
        double sum = 0.0;
        for (double v : out) sum += v;
        return sum;
    };
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    const int reps = 10;
    double ref = 0.0;
    for (int variant = 0; variant < 4; ++variant) {
        auto launch = [&]() {
            switch (variant) {
                case 0: k_soa64<<<blocks, 128>>>(s, d_list, d_cnt, d_out, n); break;
                case 1: k_pair128<<<blocks, 128>>>(nullptr, pa[0], pa[1], pa[2], pa[3], pa[4], pa[5], d_list, d_cnt, d_out, n); break;
                case 2: k_rec128<<<blocks, 128>>>(ra, rb, rc, d_list, d_cnt, d_out, n); break;
                case 3: k_rec256<<<blocks, 128>>>(ra, rb, rc, d_list, d_cnt, d_out, n); break;
            }
        };
        for (int w = 0; w < 3; ++w) launch();
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) launch();
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        CHECK(cudaGetLastError());
        float ms = 0.f;
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double sum = checksum();
        if (variant == 0) ref = sum;
        static const char *names[] = {"soa64   (11 x LDG.64)", "pair128 ( 6 x LDG.128)", "rec128  ( 6 x LDG.128 on 3 records)",
                                      "rec256  ( 3 x LDG.256 on 3 records)"};
        printf("%-38s %8.3f ms/launch  %7.2f G entries/s  checksum %s\n", names[variant], ms / reps,
               (double)total / (ms / reps * 1e-3) / 1e9, fabs(sum - ref) <= 1e-9 * fabs(ref) ? "ok" : "MISMATCH");
    }
    return 0;
}
