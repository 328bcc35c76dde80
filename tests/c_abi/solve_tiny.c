/* solve_tiny.c — a plain C (C99) caller of libezpz_b200.so: proves include/ezpz_b200.h is a C header (no C++ leaks
 * through the boundary) and that the C ABI alone — no Python, no C++ — solves a sketch.
 *
 * Builds the records of test_cases/tiny by hand (ids p = (0,1), q = (2,3); rows Fixed(p.x=0), Fixed(p.y=0),
 * Fixed(q.y=0), Vertical(p,q): SURVEY.md §8c worked example), checks the sparsity pattern against the worked CSC/CSR
 * arrays, solves through ezpz_b200_solve (what the Rust shim's `solve` binds, lib.rs:80-87) and through the batched
 * entry with host buffers, and prints "C ABI ok".  With --no-device it stops after the host-only part (pattern).
 *
 *   gcc -std=c99 -pedantic -Wall -Wextra -Werror -Iinclude tests/c_abi/solve_tiny.c -Lezpz_b200/_lib -lezpz_b200
 */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "ezpz_b200.h"

static ezpz_constraint_t rec(uint32_t kind, const uint32_t* ids, int n_ids, double p0) {
    ezpz_constraint_t c;
    int k;
    memset(&c, 0, sizeof c);
    c.kind = kind;
    for (k = 0; k < n_ids; ++k) c.ids[k] = ids[k];
    c.p0 = p0;
    c.weight = 1.0;
    return c;
}

#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            fprintf(stderr, "%s:%d: check failed: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main(int argc, char** argv) {
    const int no_device = argc > 1 && strcmp(argv[1], "--no-device") == 0;
    ezpz_constraint_t cons[4];
    const uint32_t px[1] = {0}, py[1] = {1}, qy[1] = {3}, line[4] = {0, 1, 2, 3};
    const double guesses[4] = {0.1, 0.2, 0.3, 4.0};
    const uint32_t want_col_ptr[5] = {0, 2, 3, 4, 5}, want_row_idx[5] = {0, 3, 1, 3, 2};
    const uint32_t want_row_ptr[5] = {0, 1, 2, 3, 5}, want_col_idx[5] = {0, 1, 3, 0, 2};
    const uint32_t *cp, *ri, *rp, *ci;
    ezpz_structure_t* st = NULL;
    ezpz_context_t* ctx = NULL;
    ezpz_error_detail_t det;
    ezpz_config_t cfg;
    ezpz_outcome_t out;
    double finals[4];
    uint64_t unsat[4];
    uint32_t m = 0, n = 0;
    uint64_t nnz = 0;
    int32_t rc;
    int k;

    CHECK(sizeof(ezpz_constraint_t) == 64);
    CHECK(ezpz_b200_abi_version() == EZPZ_B200_ABI_VERSION);
    cons[0] = rec(EZPZ_K_FIXED, px, 1, 0.0);
    cons[1] = rec(EZPZ_K_FIXED, py, 1, 0.0);
    cons[2] = rec(EZPZ_K_FIXED, qy, 1, 0.0);
    cons[3] = rec(EZPZ_K_VERTICAL, line, 4, 0.0);

    rc = ezpz_b200_structure_create(cons, 4, NULL, 4, &st, &det);
    CHECK(rc == EZPZ_OK);
    CHECK(ezpz_b200_structure_dims(st, &m, &n, &nnz, NULL, NULL, NULL) == EZPZ_OK);
    CHECK(m == 4 && n == 4 && nnz == 5);
    CHECK(ezpz_b200_structure_pattern(st, &cp, &ri, &rp, &ci) == EZPZ_OK);
    for (k = 0; k < 5; ++k) CHECK(cp[k] == want_col_ptr[k] && rp[k] == want_row_ptr[k]);
    for (k = 0; k < 5; ++k) CHECK(ri[k] == want_row_idx[k] && ci[k] == want_col_idx[k]);
    printf("Problem size: %u rows, %u vars\n", m, n);
    if (no_device) {
        ezpz_b200_structure_destroy(st);
        printf("C ABI ok (host only)\n");
        return 0;
    }

    rc = ezpz_b200_context_create(0, &ctx, &det);
    if (rc != EZPZ_OK) {
        fprintf(stderr, "context: %s (%s)\n", ezpz_b200_status_name(rc), det.message);
        return 1;
    }
    ezpz_b200_config_default(&cfg);
    CHECK(cfg.max_iterations == 35);

    /* ezpz::solve */
    memset(&out, 0, sizeof out);
    out.final_values = finals;
    out.unsatisfied = unsat;
    rc = ezpz_b200_solve(ctx, cons, NULL, NULL, 4, NULL, guesses, 4, &cfg, 0, &out, &det);
    CHECK(rc == EZPZ_OK);
    CHECK(out.converged == 1 && out.n_unsatisfied == 0 && out.num_eqs == 4 && out.num_vars == 4);
    CHECK(fabs(finals[0]) < 1e-6 && fabs(finals[1]) < 1e-6 && fabs(finals[2]) < 1e-6 && fabs(finals[3]) < 1e-6);
    printf("Iterations needed: %llu\n", (unsigned long long)out.iterations);

    /* the batched entry, host buffers: three problems of this topology */
    {
        const double g3[12] = {0.1, 0.2, 0.3, 4.0, -1.0, 2.0, 5.0, 0.5, 9.0, 9.0, 9.0, 9.0};
        double f3[12];
        uint32_t it3[3], mask3[3];
        uint8_t st3[3];
        ezpz_batch_io_t io;
        memset(&io, 0, sizeof io);
        io.guesses = g3;
        io.final_values = f3;
        io.iterations = it3;
        io.status = st3;
        io.unsat_mask = mask3;
        rc = ezpz_b200_solve_batch(ctx, st, &cfg, 3, &io, &det);
        CHECK(rc == EZPZ_OK);
        for (k = 0; k < 3; ++k) {
            CHECK((st3[k] & EZPZ_ST_CONVERGED) && !(st3[k] & EZPZ_ST_UNSATISFIED) && mask3[k] == 0);
            CHECK(fabs(f3[4 * k]) < 1e-6 && fabs(f3[4 * k + 1]) < 1e-6 && fabs(f3[4 * k + 2]) < 1e-6 && fabs(f3[4 * k + 3]) < 1e-6);
        }
        CHECK(it3[0] == out.iterations);
    }
    CHECK(ezpz_b200_context_launches(ctx) >= 2);
    ezpz_b200_context_destroy(ctx);
    ezpz_b200_structure_destroy(st);
    printf("C ABI ok\n");
    return 0;
}
