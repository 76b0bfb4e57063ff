/* Call latency of the single-point / small-batch path through the C ABI:
 * dind_eval_points with host coordinates and a host result, n = 1 .. 4096.
 *   gcc -O2 -Iinclude scripts/latency_bench.c -o scripts/latency_bench \
 *       -Ldatainterpolationsnd.jl_b200 -ldind -Wl,-rpath,$PWD/datainterpolationsnd.jl_b200 -lm
 * Prints one line per configuration: median and 95th-percentile microseconds per call. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "dind.h"

#define CHECK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, dind_last_error()); exit(1); } } while (0)

static double now_us(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
static int cmp(const void* a, const void* b) { double x = *(const double*)a, y = *(const double*)b; return (x > y) - (x < y); }
static double frand(void) { return rand() / (RAND_MAX + 1.0); }

static void run(const char* label, int n_in, int kind, int degree, int n_knots, int n_comp) {
    dind_handle_t h;
    CHECK(dind_create(&h, n_in, DIND_F64, 0));
    int64_t dims[9];
    int64_t total = 1;
    double* t = malloc(sizeof(double) * n_knots);
    t[0] = 0;
    for (int i = 1; i < n_knots; ++i) t[i] = t[i - 1] + 0.5 + frand();
    for (int d = 0; d < n_in; ++d) {
        CHECK(dind_set_dim(h, d, kind, t, n_knots, degree, NULL, 1));
        dims[d] = kind == DIND_BSPLINE ? n_knots + degree - 1 : n_knots;
        total *= dims[d];
    }
    int nd = n_in;
    if (n_comp > 1) { dims[nd++] = n_comp; total *= n_comp; }
    double* u = malloc(sizeof(double) * total);
    for (int64_t i = 0; i < total; ++i) u[i] = frand();
    CHECK(dind_set_u(h, u, dims, nd, 0));
    const int sizes[] = {1, 16, 256, 2048, 4096};
    for (int s = 0; s < 5; ++s) {
        const int n = sizes[s];
        double* pts[8];
        const void* ptrs[8];
        for (int d = 0; d < n_in; ++d) {
            pts[d] = malloc(sizeof(double) * n);
            for (int i = 0; i < n; ++i) pts[d][i] = t[0] + frand() * (t[n_knots - 1] - t[0]);
            ptrs[d] = pts[d];
        }
        double* out = malloc(sizeof(double) * n * n_comp);
        const int reps = 2000;
        double* us = malloc(sizeof(double) * reps);
        for (int r = 0; r < 200; ++r) CHECK(dind_eval_points(h, out, ptrs, n, NULL, 0));
        for (int r = 0; r < reps; ++r) {
            pts[0][0] = t[0] + frand() * (t[n_knots - 1] - t[0]);
            const double a = now_us();
            CHECK(dind_eval_points(h, out, ptrs, n, NULL, 0));
            us[r] = now_us() - a;
        }
        qsort(us, reps, sizeof(double), cmp);
        printf("%-34s n=%5d  median %7.2f us  p95 %7.2f us  (%.3f Mpoints/s)  kernel=%s\n", label, n, us[reps / 2],
               us[(int)(reps * 0.95)], n / us[reps / 2], dind_last_kernel(h));
        for (int d = 0; d < n_in; ++d) free(pts[d]);
        free(out); free(us);
    }
    free(u); free(t);
    CHECK(dind_destroy(h));
}

int main(void) {
    srand(1);
    run("2-D linear F64, 2 components (C1)", 2, DIND_LINEAR, 0, 64, 2);
    run("3-D cubic B-spline F64, 3 comps (C3)", 3, DIND_BSPLINE, 3, 126, 3);
    run("1-D linear F64 scalar", 1, DIND_LINEAR, 0, 1000, 1);
    return 0;
}
