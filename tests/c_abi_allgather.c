/* Plain-C client of the multi-GPU part of the boundary (include/dind.h: dind_comm_* / dind_allgather_out): one
 * process per GPU, no Python, no torch, no MPI — the NCCL unique id travels through a file.  Built and started
 * (once per rank) by tests/test_c_client.py:
 *     c_abi_allgather <rank> <world> <id file>
 * Every rank evaluates its dind_shard_range slice of a point list with its own handle, the ranks all-gather the
 * outputs, and every rank checks the global result against the closed form.  Device arrays come from the CUDA
 * runtime (what CUDA.jl's CuArray would provide on the Julia side).
 * Exit code: 0 = ok, 77 = fewer devices than ranks, anything else = failure. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "dind.h"

#define CHECK(call)                                                                       \
    do {                                                                                  \
        int rc_ = (call);                                                                 \
        if (rc_ != DIND_OK) {                                                             \
            fprintf(stderr, "[rank %d] %s -> %d: %s\n", rank, #call, rc_, dind_last_error()); \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) { fprintf(stderr, "[rank %d] %s: %s\n", rank, #call, cudaGetErrorString(e_)); return 1; } \
    } while (0)

static double f0(double a, double b) { return 3.0 + 2.3 * a - 4.7 * b; }
static double f1(double a, double b) { return -1.0 + 0.5 * a + 0.25 * b; }

static int run(int rank, int world, dind_handle_t h, int64_t n_total) {
    /* 2-D linear interpolation of two planes (u has 2 components): exact for linear interpolation */
    enum { N1 = 6, N2 = 5 };
    double t1[N1], t2[N2], u[N1 * N2 * 2];
    for (int i = 0; i < N1; ++i) t1[i] = i;
    for (int j = 0; j < N2; ++j) t2[j] = 0.8 * j;
    for (int j = 0; j < N2; ++j)
        for (int i = 0; i < N1; ++i) {
            u[i + N1 * j] = f0(t1[i], t2[j]);
            u[i + N1 * j + N1 * N2] = f1(t1[i], t2[j]);
        }
    int64_t lo, hi;
    CHECK(dind_shard_range(n_total, rank, world, &lo, &hi));
    const int64_t n_local = hi - lo;
    double* x1 = (double*)malloc(sizeof(double) * (size_t)(n_local + 1));
    double* x2 = (double*)malloc(sizeof(double) * (size_t)(n_local + 1));
    for (int64_t k = 0; k < n_local; ++k) {   /* the GLOBAL point k + lo */
        x1[k] = 5.0 * (double)((k + lo) * 7919 % 1000) / 1000.0;
        x2[k] = 3.2 * (double)((k + lo) * 104729 % 1000) / 1000.0;
    }
    CHECK(dind_set_dim(h, 0, DIND_LINEAR, t1, N1, 0, NULL, 1));
    CHECK(dind_set_dim(h, 1, DIND_LINEAR, t2, N2, 0, NULL, 1));
    CHECK(dind_set_t_eval(h, 0, x1, n_local, 0, 0));
    CHECK(dind_set_t_eval(h, 1, x2, n_local, 0, 0));
    int64_t dims[3] = {N1, N2, 2};
    CHECK(dind_set_u(h, u, dims, 3, 0));
    double *d_local = NULL, *d_global = NULL;
    CU(cudaMalloc((void**)&d_local, sizeof(double) * (size_t)(2 * n_local + 1)));
    CU(cudaMalloc((void**)&d_global, sizeof(double) * (size_t)(2 * n_total)));
    CU(cudaMemset(d_global, 0xff, sizeof(double) * (size_t)(2 * n_total)));
    CHECK(dind_eval_unstructured(h, d_local, NULL, DIND_FLAG_ASYNC));          /* (n_local, 2), stays on the stream */
    CHECK(dind_allgather_out(h, d_local, d_global, n_total, 1, 2, 0));         /* (n_total, 2) on every rank */
    double* g = (double*)malloc(sizeof(double) * (size_t)(2 * n_total));
    CU(cudaMemcpy(g, d_global, sizeof(double) * (size_t)(2 * n_total), cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int64_t k = 0; k < n_total && !bad; ++k) {
        const double a = 5.0 * (double)(k * 7919 % 1000) / 1000.0, b = 3.2 * (double)(k * 104729 % 1000) / 1000.0;
        if (!(fabs(g[k] - f0(a, b)) <= 1e-12 * 30) || !(fabs(g[k + n_total] - f1(a, b)) <= 1e-12 * 30)) {
            fprintf(stderr, "[rank %d] n_total=%lld: mismatch at point %lld: %g %g vs %g %g\n", rank, (long long)n_total,
                    (long long)k, g[k], g[k + n_total], f0(a, b), f1(a, b));
            bad = 1;
        }
    }
    /* the same result from ONE collective call: the shard is evaluated straight into its rows of the global array and
     * the rows of the other ranks arrive chunk by chunk while the next chunk is computed (3 chunks here) */
    double* d_fused = NULL;
    CU(cudaMalloc((void**)&d_fused, sizeof(double) * (size_t)(2 * n_total + 1)));
    CU(cudaMemset(d_fused, 0xff, sizeof(double) * (size_t)(2 * n_total + 1)));
    CHECK(dind_eval_unstructured_allgather(h, d_fused, n_total, NULL, 3, 0));
    double* g2 = (double*)malloc(sizeof(double) * (size_t)(2 * n_total + 1));
    CU(cudaMemcpy(g2, d_fused, sizeof(double) * (size_t)(2 * n_total), cudaMemcpyDeviceToHost));
    if (!bad && memcmp(g, g2, sizeof(double) * (size_t)(2 * n_total))) {
        fprintf(stderr, "[rank %d] n_total=%lld: dind_eval_unstructured_allgather differs from eval + allgather\n", rank, (long long)n_total);
        bad = 1;
    }
    free(g2);
    cudaFree(d_fused);
    free(g); free(x1); free(x2);
    cudaFree(d_local); cudaFree(d_global);
    return bad ? 2 : 0;
}

int main(int argc, char** argv) {
    if (argc != 4) { fprintf(stderr, "usage: %s rank world idfile\n", argv[0]); return 64; }
    const int rank = atoi(argv[1]), world = atoi(argv[2]);
    const char* idfile = argv[3];
    if (dind_device_count() < world) { printf("need %d devices, have %d\n", world, dind_device_count()); return 77; }
    CU(cudaSetDevice(rank));
    dind_handle_t h = NULL;
    CHECK(dind_create(&h, 2, DIND_F64, rank));

    unsigned char id[DIND_UNIQUE_ID_BYTES];
    if (rank == 0) {
        CHECK(dind_comm_get_unique_id(id));
        char tmp[4096];
        snprintf(tmp, sizeof tmp, "%s.tmp", idfile);
        FILE* f = fopen(tmp, "wb");
        if (!f || fwrite(id, 1, sizeof id, f) != sizeof id) return 65;
        fclose(f);
        if (rename(tmp, idfile)) return 66;
    } else {
        FILE* f = NULL;
        for (int tries = 0; tries < 600 && !(f = fopen(idfile, "rb")); ++tries) usleep(100000);
        if (!f || fread(id, 1, sizeof id, f) != sizeof id) { fprintf(stderr, "[rank %d] no unique id\n", rank); return 67; }
        fclose(f);
    }
    CHECK(dind_comm_init(h, id, rank, world));
    int r = -1, w = -1, ver = 0;
    CHECK(dind_comm_info(h, &r, &w, &ver));
    if (r != rank || w != world || ver < 20000) { fprintf(stderr, "[rank %d] comm_info %d %d %d\n", rank, r, w, ver); return 3; }

    int rc = run(rank, world, h, 4096);            /* even shards: ncclAllGather per component column */
    if (!rc) rc = run(rank, world, h, 4099);       /* uneven shards: ncclBroadcast per rank and column */
    if (!rc) rc = run(rank, world, h, world - 1);  /* some rank owns nothing */
    CHECK(dind_comm_destroy(h));
    CHECK(dind_destroy(h));
    if (!rc) printf("[rank %d] all-gather ok (NCCL %d)\n", rank, ver);
    return rc;
}
