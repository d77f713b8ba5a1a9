/* Two ranks, two GPUs, one process: the column-slab decomposition and the NCCL halo exchange of include/mohid_adt.h from
 * plain C (what an MPI host does per rank; here two threads stand in for the ranks and a shared buffer for MPI_Bcast).
 *
 * A 16 x 24 x 4 closed box of still water with a tracer that varies along j: three diffusive steps on the two slabs
 * (12 owned columns + 2 ghost columns each, halos exchanged after every step) must reproduce, bit for bit, the run of
 * the undivided box on one GPU.
 *
 *   gcc -std=c99 -pthread -Iinclude examples/c_driver_2rank.c -Lmohid_b200 -lmohid_adt -Wl,-rpath,$PWD/mohid_b200 -lm -o c2
 *
 * Needs two CUDA devices to do anything (prints "skipped" and returns 0 otherwise); tests/test_capi_cpu.py compiles and
 * links it, tests/test_gpu_multi.py runs it. */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mohid_adt.h"

enum { I = 16, JG = 24, K = 4, GHOST = 2, STEPS = 3 };

typedef struct {
    int J, j_offset;            /* local work columns 1..J are the global columns j_offset+1 .. j_offset+J */
    int ld, nj, nk;
    size_t n2, n3;
    double *d2[4], *d3[11], *prop;
    int *kfloor, *bnd, *m3[6];
} Slab;

static void build(Slab *s, int j_first, int j_last) {
    s->J = j_last - j_first + 1; s->j_offset = j_first - 1;
    s->ld = I + 2; s->nj = s->J + 2; s->nk = K + 2;
    s->n2 = (size_t)s->ld * s->nj; s->n3 = s->n2 * s->nk;
    for (int a = 0; a < 4; ++a) { s->d2[a] = malloc(s->n2 * sizeof(double)); for (size_t q = 0; q < s->n2; ++q) s->d2[a][q] = 500.0; }
    for (int a = 0; a < 11; ++a) s->d3[a] = calloc(s->n3, sizeof(double));
    for (int a = 0; a < 6; ++a) s->m3[a] = calloc(s->n3, sizeof(int));
    s->prop = calloc(s->n3, sizeof(double));
    s->kfloor = calloc(s->n2, sizeof(int)); s->bnd = calloc(s->n2, sizeof(int));
    for (size_t q = 0; q < s->n2; ++q) s->kfloor[q] = 1;
    for (int k = 0; k < s->nk; ++k)
        for (int j = 0; j < s->nj; ++j)
            for (int i = 0; i < s->ld; ++i) {
                const size_t q = i + (size_t)s->ld * (j + (size_t)s->nj * k);
                const int gj = j + s->j_offset;                                   /* global column */
                const int in = i >= 1 && i <= I && gj >= 1 && gj <= JG && k >= 1 && k <= K;
                s->d3[3][q] = s->d3[4][q] = 500.0 * 500.0 * 5.0;                  /* VolumeZOld, VolumeZ */
                s->d3[5][q] = 5.0; s->d3[6][q] = 1e-3;                            /* Visc_H, Diff_V */
                s->d3[7][q] = s->d3[8][q] = 5.0;                                  /* DWZ, DZZ */
                s->d3[9][q] = s->d3[10][q] = 500.0 * 5.0;                         /* AreaU, AreaV */
                s->m3[0][q] = s->m3[2][q] = in;                                   /* OpenPoints3D, WaterPoints3D */
                s->m3[3][q] = in && gj >= 2;                                      /* ComputeFacesU3D */
                s->m3[4][q] = in && i >= 2;                                       /* ComputeFacesV3D */
                s->m3[5][q] = in && k >= 2;                                       /* ComputeFacesW3D */
                s->prop[q] = in ? 10.0 + 0.25 * gj + 0.01 * i + 0.1 * k : 0.0;
            }
}

static mohid_adt_params params(void) {
    mohid_adt_params p;
    memset(&p, 0, sizeof p);
    p.Schmidt_H = 1.0; p.SchmidtCoef_V = 1.0; p.SchmidtBackground_V = 1e-8;
    p.AdvMethodH = MOHID_P2_TVD; p.TVDLimitationH = MOHID_SuperBee; p.AdvMethodV = MOHID_P2_TVD; p.TVDLimitationV = MOHID_SuperBee;
    p.Upwind2H = 1; p.Upwind2V = 1; p.VolumeRelMax = 1.5; p.DTProp = 3600.0; p.ImpExp_AdvV = 1.0; p.ImpExp_DifV = 1.0;
    p.BoundaryCondition = MOHID_BC_None;
    return p;
}

static int check(int rc, const int *handle, const char *what) {
    if (!rc) return 0;
    char msg[512];
    int n = (int)sizeof msg;
    mohid_adt_last_error(handle, msg, &n);
    fprintf(stderr, "%s: mohid_adt error %d: %s\n", what, rc, msg);
    return rc;
}

/* create + inputs + upload on `device`; returns the handle (0 on failure) */
static int start(Slab *s, int device) {
    mohid_adt_size3d size = {0, I + 1, 0, s->J + 1, 0, K + 1}, work = {1, I, 1, s->J, 1, K};
    mohid_adt_options opt;
    memset(&opt, 0, sizeof opt);
    opt.Docycle_method = 1; opt.device = device;
    int h = 0, one = 1;
    const double *pp[1] = {s->prop};
    if (check(mohid_adt_create(&h, &size, &work, &s->ld, &opt), &h, "create")) return 0;
    if (check(mohid_adt_set_grid2d(&h, s->d2[0], s->d2[1], s->d2[2], s->d2[3], s->kfloor, s->bnd), &h, "set_grid2d")) return 0;
    if (check(mohid_adt_set_step(&h, s->d3[0], s->d3[1], s->d3[2], s->d3[3], s->d3[4], s->d3[5], s->d3[6], s->d3[7], s->d3[8],
                                 s->d3[9], s->d3[10], s->m3[0], s->m3[1], s->m3[2], s->m3[3], s->m3[4], s->m3[5], NULL),
              &h, "set_step")) return 0;
    if (check(mohid_adt_upload_props(&h, &one, pp, NULL), &h, "upload_props")) return 0;
    return h;
}

static pthread_barrier_t bar;
static char unique_id[128];
static Slab slab[2];
static int failed[2];

static void *rank_main(void *arg) {
    const int rank = (int)(size_t)arg, nranks = 2, ghost = GHOST, overlap = 1, one = 1, nid = 128;
    Slab *s = &slab[rank];
    /* rank 0 owns the global columns 1..12 (+ 13, 14 as ghosts), rank 1 owns 13..24 (+ 11, 12) */
    const int j_begin = rank == 0 ? 1 : 1 + GHOST, j_count = JG / 2;
    mohid_adt_params p = params();
    double *pp[1] = {s->prop};
    int h = start(s, rank);
    failed[rank] = h == 0;
    if (rank == 0 && h) failed[0] = check(mohid_adt_comm_get_unique_id(unique_id, &nid), NULL, "comm_get_unique_id") != 0;
    pthread_barrier_wait(&bar);                       /* "MPI_Bcast" of the 128-byte id */
    if (failed[0] || failed[1]) return NULL;
    int rc = check(mohid_adt_set_active_columns(&h, &j_begin, &j_count), &h, "set_active_columns");
    if (!rc) rc = check(mohid_adt_comm_init(&h, &nranks, &rank, unique_id, &ghost, &overlap), &h, "comm_init");
    for (int t = 0; t < STEPS && !rc; ++t) {
        rc = check(mohid_adt_advect_device(&h, &one, &p, &one), &h, "advect_device");
        if (!rc) rc = check(mohid_adt_exchange_halos(&h, &one), &h, "exchange_halos");
    }
    if (!rc) rc = check(mohid_adt_download_props(&h, &one, pp), &h, "download_props");
    if (!rc) rc = check(mohid_adt_comm_destroy(&h), &h, "comm_destroy");
    mohid_adt_destroy(&h);
    failed[rank] = rc != 0;
    return NULL;
}

int main(void) {
    /* the undivided box on device 0 */
    Slab whole;
    build(&whole, 1, JG);
    int h = start(&whole, 0), one = 1, steps = STEPS;
    if (!h) return 1;
    mohid_adt_params p = params();
    double *pp[1] = {whole.prop};
    if (check(mohid_adt_advect_device(&h, &one, &p, &steps), &h, "advect_device")) return 1;
    if (check(mohid_adt_download_props(&h, &one, pp), &h, "download_props")) return 1;
    mohid_adt_destroy(&h);

    /* is there a second device?  (create on device 1 fails with a message when there is not) */
    {
        mohid_adt_size3d size = {0, 3, 0, 3, 0, 3}, work = {1, 2, 1, 2, 1, 2};
        mohid_adt_options opt;
        memset(&opt, 0, sizeof opt);
        opt.device = 1;
        int ld = 4, probe = 0;
        if (mohid_adt_create(&probe, &size, &work, &ld, &opt)) { printf("skipped: one CUDA device only\n"); return 0; }
        mohid_adt_destroy(&probe);
    }
    build(&slab[0], 1, JG / 2 + GHOST);
    build(&slab[1], JG / 2 + 1 - GHOST, JG);
    pthread_barrier_init(&bar, NULL, 2);
    pthread_t th[2];
    for (size_t r = 0; r < 2; ++r) pthread_create(&th[r], NULL, rank_main, (void *)r);
    for (int r = 0; r < 2; ++r) pthread_join(th[r], NULL);
    if (failed[0] || failed[1]) return 1;

    int bad = 0;
    double moved = 0.0;
    for (int r = 0; r < 2; ++r) {
        const Slab *s = &slab[r];
        for (int k = 1; k <= K; ++k)
            for (int j = 1; j <= s->J; ++j)
                for (int i = 1; i <= I; ++i) {
                    const double a = s->prop[i + (size_t)s->ld * (j + (size_t)s->nj * k)];
                    const double b = whole.prop[i + (size_t)whole.ld * (j + s->j_offset + (size_t)whole.nj * k)];
                    if (memcmp(&a, &b, sizeof a)) ++bad;               /* owned AND exchanged ghost columns */
                    moved = fmax(moved, fabs(b - (10.0 + 0.25 * (j + s->j_offset) + 0.01 * i + 0.1 * k)));
                }
    }
    printf("two slabs vs the undivided box after %d steps: %d cells differ (field moved by up to %.3e)\n", STEPS, bad, moved);
    return (bad == 0 && moved > 1e-6) ? 0 : 2;
}
