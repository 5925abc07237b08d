/* phi_cabi.c -- gen.phi(gen.genealogy(path)) through the C ABI of libgenlib_cuda.so, from plain C.
 *
 * The same sequence of calls as the Julia shim (genlib.jl_b200/julia/GenLibCUDA.jl) and the Python
 * ctypes mirror: loader -> pro -> ranks -> plan -> engine -> run -> fetch.  No torch, no Python.
 *
 *   gcc -O2 -Iinclude examples/phi_cabi.c -o phi_cabi -Lgenlib.jl_b200 -lgenlib_cuda \
 *       -Wl,-rpath,$PWD/genlib.jl_b200
 *   ./phi_cabi tests/data/geneaJi.csv            # prints the kinship matrix of the probands
 *   ./phi_cabi tests/data/geneaJi.csv sparse     # the values gen.sparse_phi keeps
 *   ./phi_cabi tests/data/genea140.csv devices=0,1   # one process, two GPUs (genlib_phi_multi)
 *
 * Exit status: 0, or the GENLIB_E* code of the failing call (4 = no CUDA device: there is no CPU
 * fallback). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "genlib_cuda.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ != GENLIB_OK) {                                                  \
            fprintf(stderr, "%s: status %d: %s\n", #call, rc_, genlib_last_error()); \
            return rc_;                                                          \
        }                                                                        \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s pedigree.csv [sparse | devices=0,1,...]\n", argv[0]); return GENLIB_EINVAL; }
    const int schedule = argc > 2 && strcmp(argv[2], "sparse") == 0 ? GENLIB_SCHEDULE_SPARSE_PHI : GENLIB_SCHEDULE_PHI;
    int32_t devices[16], n_dev = 0;
    if (argc > 2 && strncmp(argv[2], "devices=", 8) == 0)
        for (char *tok = strtok(argv[2] + 8, ","); tok && n_dev < 16; tok = strtok(NULL, ",")) devices[n_dev++] = atoi(tok);
    genlib_pedigree *ped = NULL;
    CHECK(genlib_genealogy_csv(argv[1], 1, &ped));                 /* gen.genealogy(path), src/create.jl:131-189 */
    const int64_t n = genlib_pedigree_n(ped);
    int32_t *father = malloc((size_t)n * sizeof *father), *mother = malloc((size_t)n * sizeof *mother);
    int64_t *pro = malloc((size_t)n * sizeof *pro), *ids = malloc((size_t)n * sizeof *ids);
    CHECK(genlib_pedigree_arrays(ped, ids, father, mother, NULL));
    const int64_t n_pro = genlib_pedigree_pro(ped, pro);           /* gen.pro(ped), src/identify.jl:35-39 */
    int32_t *ranks = malloc((size_t)n_pro * sizeof *ranks);
    CHECK(genlib_pedigree_ranks(ped, n_pro, pro, ranks));
    if (n_dev > 0) {                                               /* the one-call form on several devices */
        int32_t *seen = calloc((size_t)n, sizeof *seen), nu = 0;
        for (int64_t t = 0; t < n_pro; t++) if (!seen[ranks[t]]++) nu++;
        float *phi = malloc((size_t)nu * nu * sizeof *phi);
        CHECK(genlib_phi_multi((int32_t)n, father, mother, (int32_t)n_pro, ranks, phi, GENLIB_F32, GENLIB_NUMERICS_REFERENCE,
                               n_dev, devices, NULL));
        for (int32_t a = 0; a < nu; a++) {
            for (int32_t b = 0; b < nu; b++) printf("%s%.9g", b ? " " : "", (double)phi[(size_t)a * nu + b]);
            printf("\n");
        }
        genlib_pedigree_destroy(ped);
        free(father); free(mother); free(pro); free(ids); free(ranks); free(phi); free(seen);
        return 0;
    }
    genlib_plan *plan = NULL;
    /* the IDs order the founders in sparse_phi's queue (src/identify.jl:15-19); phi ignores them */
    CHECK(genlib_plan_create_ex((int32_t)n, father, mother, ids, (int32_t)n_pro, ranks, 1, schedule, &plan));
    const int32_t nu = genlib_plan_n_unique(plan);
    float *phi = malloc((size_t)nu * nu * sizeof *phi);
    genlib_engine *eng = NULL;
    CHECK(genlib_engine_create(plan, GENLIB_NUMERICS_REFERENCE, -1, &eng));   /* GENLIB_ECUDA without a GPU */
    CHECK(genlib_engine_run(eng, 0));
    CHECK(genlib_engine_fetch(eng, phi, GENLIB_F32));
    for (int32_t a = 0; a < nu; a++) {
        for (int32_t b = 0; b < nu; b++) printf("%s%.9g", b ? " " : "", (double)phi[(size_t)a * nu + b]);
        printf("\n");
    }
    genlib_engine_destroy(eng);
    genlib_plan_destroy(plan);
    genlib_pedigree_destroy(ped);
    free(father); free(mother); free(pro); free(ids); free(ranks); free(phi);
    return 0;
}
