/* examples/lid2d_driver.c -- a plain-C driver for the 2-D lid-driven cavity that binds libmglc.so through include/mglc.h only.
 *
 * It plays the role of the reference's C program (MPI/Lid_driven_cavity/c/lid_driven_cavity.c: main() :50-80, output_binary()
 * :428-457) with every per-step subroutine replaced by the library: the same 200 x 200 lattice, Re = 1000, u_zero = 0.1, the
 * residual check every 2000 iterations, and at the end the same `flow_binary` file (x, y, rho, u, v as double[NX][NY]).  With
 * MGLC_ARITH_STRICT that file is byte-identical to the one the reference program writes after the same number of iterations
 * (tests/test_examples_gpu.py compares its SHA-256 with the committed hash of the reference's own file).
 *
 *   gcc -std=c99 -O2 -Iinclude examples/lid2d_driver.c -Lmglc_b200 -lmglc -Wl,-rpath,$PWD/mglc_b200 -o lid2d_driver
 *   ./lid2d_driver [max_iterations = 2000] [strict = 1] [output = flow_binary] [model = 2]
 * model follows the program's own switch (c:13-14): 2 = MRT (its shipped setting), 1 = SRT / BGK.
 */
#include <stdio.h>
#include <stdlib.h>

#include "mglc.h"

#define CHECK(call)                                                                     \
    do {                                                                                \
        int rc_ = (call);                                                               \
        if (rc_ != MGLC_OK) {                                                           \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, mglc_last_error());     \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

int main(int argc, char **argv) {
    const int itc_max = argc > 1 ? atoi(argv[1]) : 2000;
    const int strict = argc > 2 ? atoi(argv[2]) : 1;
    const char *out = argc > 3 ? argv[3] : "flow_binary";
    const int model = argc > 4 ? atoi(argv[4]) : 2;        /* c:13-14: const int SRT = 1, MRT = 2; const int model = 2; */
    if (model != 1 && model != 2) {
        fprintf(stderr, "lid2d_driver: model must be 1 (SRT) or 2 (MRT)\n");
        return 2;
    }
    const double eps = 1e-6;

    mglc_l2d_desc d;
    CHECK(mglc_l2d_desc_init(&d, model == 1 ? MGLC_L2D_C_SRT : MGLC_L2D_C));   /* 200 x 200, Re = 1000, u_zero = 0.1, rho_zero = 1 */
    d.arith = strict ? MGLC_ARITH_STRICT : MGLC_ARITH_FAST;
    const int NX = d.total_nx, NY = d.total_ny;
    const double height = (double)NX;
    const int zero2[2] = {0, 0};
    mglc_l2d *h = NULL;
    CHECK(mglc_l2d_create(&h, &d, zero2, 1, 0, 0, NULL));
    CHECK(mglc_l2d_initial(h));                            /* initial() */

    double error = 1.0;
    int itc = 0;
    while (error >= eps && itc < itc_max) {
        const int n = itc_max - itc < 2000 ? itc_max - itc : 2000;
        CHECK(mglc_l2d_step(h, n));                        /* n x (collision, streaming, boundary, macro) */
        itc += n;
        if (itc % 2000 == 0) {
            CHECK(mglc_l2d_check(h, &error));              /* check() */
            printf("%d %.15e\n", itc, error);
        }
    }

    /* the library hands rho,u,v back as (nx, ny) with i fastest; the C program's arrays are [NX][NY] with j fastest */
    const size_t n = (size_t)NX * NY;
    double *col = malloc(3 * n * sizeof(double)), *x = malloc(n * sizeof(double)), *y = malloc(n * sizeof(double));
    double *row = malloc(3 * n * sizeof(double));
    if (!col || !row || !x || !y) return 2;
    CHECK(mglc_l2d_download(h, 0, NULL, NULL, col, col + n, col + 2 * n));
    for (int q = 0; q < 3; ++q)
        for (int i = 0; i < NX; ++i)
            for (int j = 0; j < NY; ++j) row[q * n + (size_t)i * NY + j] = col[q * n + (size_t)j * NX + i];
    const double delta_x = height / (NX - 1), delta_y = height / (NY - 1);
    for (int i = 0; i < NX; ++i)
        for (int j = 0; j < NY; ++j) { x[i * NX + j] = i * delta_x; y[i * NX + j] = j * delta_y; }
    FILE *fp = fopen(out, "wb+");
    if (!fp) return 3;
    fwrite(x, sizeof(double), n, fp); fwrite(y, sizeof(double), n, fp);
    fwrite(row, sizeof(double), n, fp); fwrite(row + n, sizeof(double), n, fp); fwrite(row + 2 * n, sizeof(double), n, fp);
    fclose(fp);
    free(col); free(row); free(x); free(y);
    CHECK(mglc_l2d_destroy(h));
    return 0;
}
