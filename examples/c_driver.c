/* Minimal C host for the C-ABI of include/mohid_adt.h (what the Fortran shim does through ISO_C_BINDING):
 * a 16 x 12 x 4 closed box of still water, one property, one step; a constant field must stay constant.
 *
 *   gcc -std=c99 -Iinclude examples/c_driver.c -Lmohid_b200 -lmohid_adt -Wl,-rpath,$PWD/mohid_b200 -lm -o c_driver
 *
 * Needs a CUDA device to run (the library has no CPU fallback); tests/test_capi_cpu.py only compiles and links it. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "mohid_adt.h"

int main(void) {
    const int I = 16, J = 12, K = 4, ld = I + 2, nj = J + 2, nk = K + 2;
    const size_t n2 = (size_t)ld * nj, n3 = n2 * nk;
    mohid_adt_size3d size = {0, I + 1, 0, J + 1, 0, K + 1}, work = {1, I, 1, J, 1, K};
    mohid_adt_options opt = {0};
    opt.Docycle_method = 1; opt.device = -1;
    int handle = 0, rc;
    char msg[512];
    int msglen = (int)sizeof msg;

    double *d2[4], *d3[11], *prop = calloc(n3, sizeof(double));
    int *kfloor = calloc(n2, sizeof(int)), *bnd = calloc(n2, sizeof(int)), *m3[6];
    for (int a = 0; a < 4; ++a) { d2[a] = malloc(n2 * sizeof(double)); for (size_t q = 0; q < n2; ++q) d2[a][q] = 500.0; }
    for (int a = 0; a < 11; ++a) d3[a] = calloc(n3, sizeof(double));
    for (int a = 0; a < 6; ++a) m3[a] = calloc(n3, sizeof(int));
    for (size_t q = 0; q < n2; ++q) kfloor[q] = 1;
    for (int k = 0; k < nk; ++k)
        for (int j = 0; j < nj; ++j)
            for (int i = 0; i < ld; ++i) {
                const size_t q = i + (size_t)ld * (j + (size_t)nj * k);
                const int in = i >= 1 && i <= I && j >= 1 && j <= J && k >= 1 && k <= K;
                d3[3][q] = d3[4][q] = 500.0 * 500.0 * 5.0;      /* VolumeZOld, VolumeZ */
                d3[5][q] = 5.0; d3[6][q] = 1e-3;               /* Visc_H, Diff_V */
                d3[7][q] = d3[8][q] = 5.0;                     /* DWZ, DZZ */
                d3[9][q] = d3[10][q] = 500.0 * 5.0;            /* AreaU, AreaV */
                m3[0][q] = m3[2][q] = in;                      /* OpenPoints3D, WaterPoints3D */
                m3[1][q] = 0;                                  /* LandPoints3D */
                m3[3][q] = in && j >= 2;                       /* ComputeFacesU3D: faces between two water cells */
                m3[4][q] = in && i >= 2;                       /* ComputeFacesV3D */
                m3[5][q] = in && k >= 2;                       /* ComputeFacesW3D */
                prop[q] = in ? 17.5 : 0.0;
            }
    mohid_adt_params p = {0};
    p.Schmidt_H = 1.0; p.SchmidtCoef_V = 1.0; p.SchmidtBackground_V = 1e-8;
    p.AdvMethodH = MOHID_P2_TVD; p.TVDLimitationH = MOHID_SuperBee; p.AdvMethodV = MOHID_P2_TVD; p.TVDLimitationV = MOHID_SuperBee;
    p.Upwind2H = 1; p.Upwind2V = 1; p.VolumeRelMax = 1.5; p.DTProp = 30.0; p.ImpExp_AdvV = 1.0; p.ImpExp_DifV = 1.0;
    p.BoundaryCondition = MOHID_BC_None;

    rc = mohid_adt_create(&handle, &size, &work, &ld, &opt);
    if (!rc) rc = mohid_adt_set_grid2d(&handle, d2[0], d2[1], d2[2], d2[3], kfloor, bnd);
    if (!rc) rc = mohid_adt_set_step(&handle, d3[0], d3[1], d3[2], d3[3], d3[4], d3[5], d3[6], d3[7], d3[8], d3[9], d3[10],
                                     m3[0], m3[1], m3[2], m3[3], m3[4], m3[5], NULL);
    double *props[1] = {prop};
    const int one = 1;
    if (!rc) rc = mohid_adt_advect_batch(&handle, &one, props, NULL, &p);
    if (rc) {
        mohid_adt_last_error(&handle, msg, &msglen);
        fprintf(stderr, "mohid_adt error %d: %s\n", rc, msg);
        return 1;
    }
    double worst = 0.0;
    for (int k = 1; k <= K; ++k)
        for (int j = 1; j <= J; ++j)
            for (int i = 1; i <= I; ++i) {
                const double e = fabs(prop[i + (size_t)ld * (j + (size_t)nj * k)] - 17.5);
                if (e > worst) worst = e;
            }
    printf("max |C - 17.5| after one step: %.3e\n", worst);
    mohid_adt_destroy(&handle);
    return worst < 1e-12 ? 0 : 2;
}
