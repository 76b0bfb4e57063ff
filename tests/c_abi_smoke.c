/* Plain-C client of the drop-in boundary (include/dind.h): no Python, no torch — what a Julia
 * `ccall` or any other FFI sees.  Built and run by tests/test_c_client.py.
 * Exit code: 0 = evaluated on a GPU and matched the closed form; 77 = no CUDA device (the library
 * refused with DIND_ERR_NO_DEVICE: there is no CPU fallback); anything else = failure. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "dind.h"

#define CHECK(call)                                                          \
    do {                                                                     \
        int rc_ = (call);                                                    \
        if (rc_ != DIND_OK) {                                                \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, dind_last_error()); \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main(void) {
    dind_handle_t h = NULL;
    int rc = dind_create(&h, 2, DIND_F64, 0);
    if (rc == DIND_ERR_NO_DEVICE) {
        printf("no device: %s\n", dind_last_error());
        return 77;
    }
    if (rc != DIND_OK) { fprintf(stderr, "dind_create -> %d: %s\n", rc, dind_last_error()); return 1; }

    /* 2-D linear interpolation of the plane f = 3 + 2.3 t1 - 4.7 t2 (test/test_interpolations.jl:70-82) */
    enum { N1 = 6, N2 = 5, NE = 4 };
    double t1[N1], t2[N2], u[N1 * N2], x1[NE] = {0.25, 1.5, 2.75, 4.9}, x2[NE] = {0.1, 3.2, 1.0, 2.5};
    for (int i = 0; i < N1; ++i) t1[i] = i;
    for (int j = 0; j < N2; ++j) t2[j] = 0.8 * j;
    for (int j = 0; j < N2; ++j)
        for (int i = 0; i < N1; ++i) u[i + N1 * j] = 3.0 + 2.3 * t1[i] - 4.7 * t2[j];   /* column-major like Julia */
    CHECK(dind_set_dim(h, 0, DIND_LINEAR, t1, N1, 0, NULL, 1));
    CHECK(dind_set_dim(h, 1, DIND_LINEAR, t2, N2, 0, NULL, 1));
    CHECK(dind_set_t_eval(h, 0, x1, NE, 0, 0));
    CHECK(dind_set_t_eval(h, 1, x2, NE, 0, 0));
    int64_t dims[2] = {N1, N2};
    CHECK(dind_set_u(h, u, dims, 2, 0));

    double out[NE], grid[NE * NE], d1[NE];
    int orders[2] = {1, 0};
    CHECK(dind_eval_unstructured(h, out, NULL, 0));
    CHECK(dind_eval_unstructured(h, d1, orders, 0));
    CHECK(dind_eval_grid(h, grid, NULL, 0));
    for (int k = 0; k < NE; ++k) {
        if (fabs(out[k] - (3.0 + 2.3 * x1[k] - 4.7 * x2[k])) > 1e-12) { fprintf(stderr, "value mismatch at %d\n", k); return 2; }
        if (fabs(d1[k] - 2.3) > 1e-12) { fprintf(stderr, "derivative mismatch at %d\n", k); return 3; }
        for (int l = 0; l < NE; ++l)
            if (fabs(grid[k + NE * l] - (3.0 + 2.3 * x1[k] - 4.7 * x2[l])) > 1e-12) { fprintf(stderr, "grid mismatch\n"); return 4; }
    }
    int64_t idx[NE];
    CHECK(dind_get_idx_eval(h, 0, idx));
    if (idx[0] != 1 || idx[1] != 2 || idx[2] != 3 || idx[3] != 5) { fprintf(stderr, "idx mismatch\n"); return 5; }

    /* validation errors come back as codes, not exceptions */
    double bad[3] = {0.0, 0.0, 1.0};
    if (dind_set_dim(h, 0, DIND_LINEAR, bad, 3, 0, NULL, 1) != DIND_ERR_INVALID_ARGUMENT) return 6;
    CHECK(dind_destroy(h));
    printf("C client ok: kernel ran on the GPU, values/derivatives/grid/idx match\n");
    return 0;
}
