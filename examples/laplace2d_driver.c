/* examples/laplace2d_driver.c -- a plain-C driver for the 2-D Laplace / Jacobi problem that binds libmglc.so through
 * include/mglc.h only.
 *
 * It plays the role of the reference's C program MPI/Laplace/c/laplace2d.c (main() :28-66): the same 4320 x 4320 array whose
 * outer ring is the boundary (top row = 1, the rest 0, :41-50), at most 1000 Jacobi iterations, the max-norm difference of two
 * successive iterates evaluated after EVERY iteration for the stopping test (tolerance 1e-5, :55) and printed every 100 (:60)
 * in the program's own format.  jacobi() :71-86 and swap() :88-97 are replaced by the library:
 *   - the library's subdomain is the interior (nx-2) x (ny-2) with the boundary ring as its ghost layer; mglc_jacobi_init()
 *     sets exactly the program's boundary values (top ghost row 1, corners included);
 *   - the program's x is the library's first (fastest) index, so the four neighbours are added in the program's order
 *     A[x-1][y] + A[x+1][y] + A[x][y-1] + A[x][y+1] and every iterate is bit-identical to the program's;
 *   - mglc_jacobi_check_diff() returns max |A - A_p| and then sets A_p = A, so called once per iteration it is the program's
 *     `error`.
 * The printed lines equal the reference program's stdout; with a 4th argument the final array is written as the program holds
 * it (double[nx][ny], y fastest) for a bitwise comparison.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/laplace2d_driver.c -Lmglc_b200 -lmglc -Wl,-rpath,$PWD/mglc_b200 -o laplace2d_driver
 *   ./laplace2d_driver [nx = 4320] [ny = 4320] [itc_max = 1000] [dump_file]
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
    const int nx = argc > 1 ? atoi(argv[1]) : 4320;
    const int ny = argc > 2 ? atoi(argv[2]) : 4320;
    const int itc_max = argc > 3 ? atoi(argv[3]) : 1000;
    const char *dump = argc > 4 ? argv[4] : NULL;
    const double tolerance = 1e-5;
    if (nx < 3 || ny < 3) {
        fprintf(stderr, "laplace2d_driver: the array needs an interior (nx, ny >= 3)\n");
        return 2;
    }

    const int gn[3] = {nx - 2, ny - 2, 1}, zero3[3] = {0, 0, 0};
    mglc_jacobi *h = NULL;
    CHECK(mglc_jacobi_create(&h, 2, gn, zero3, 1, 0, 0, NULL));
    CHECK(mglc_jacobi_init(h));                                 /* memset + "set top boundary = 1.0", :41-50 */

    int itc = 0;
    double error = 1.0;
    while (error > tolerance && itc++ < itc_max) {              /* :55 */
        CHECK(mglc_jacobi_step(h, 1));                          /* jacobi() + swap() */
        CHECK(mglc_jacobi_check_diff(h, &error));               /* the value jacobi() returns */
        if (itc % 100 == 0) printf("%5d, %0.6f\n", itc, error); /* :60 */
    }

    if (dump) {
        const size_t n = (size_t)nx * ny;
        double *lib = malloc(n * sizeof(double)), *prog = malloc(n * sizeof(double));
        if (!lib || !prog) return 2;
        CHECK(mglc_jacobi_download(h, 0, lib, NULL));           /* (0:nx-1, 0:ny-1) of the program, x fastest */
        for (int x = 0; x < nx; ++x)
            for (int y = 0; y < ny; ++y) prog[(size_t)x * ny + y] = lib[(size_t)x + (size_t)nx * y];
        FILE *fp = fopen(dump, "wb");
        if (!fp || fwrite(prog, sizeof(double), n, fp) != n) {
            fprintf(stderr, "laplace2d_driver: cannot write %s\n", dump);
            return 3;
        }
        fclose(fp);
        free(lib);
        free(prog);
    }
    CHECK(mglc_jacobi_destroy(h));
    return 0;
}
