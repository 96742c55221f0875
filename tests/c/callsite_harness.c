/*
 * callsite_harness.c -- imitates the reference handlers' use of the L1 API, in C, the way
 * src/compare_handler.c:98-160,403-416, src/post_handler.c:244,333, src/put_handler.c:209,247,
 * src/delete_handler.c:96 and src/main.c:349,398 use it: includes "vector_database.h" /
 * "kdtree.h", reaches into db->size / db->kdtree / vec->uuid, passes Vector by value.
 *
 * The SAME object file is linked once against oracle/_ref/libsvdb_ref.so (the reference) and
 * once against libsvdb_b200.so (the drop-in); tests/test_gpu_dropin_c.py compares the two
 * outputs byte for byte.
 *
 * usage: harness <input.bin> <scratch.db>
 * input: u64 n, D, K, nq, npairs, nops ; f64 rows[n*D] ; f64 queries[nq*D] ;
 *        u64 pairs[npairs*2] ; then nops records {u64 code(1 update,2 delete), u64 index, f64 row[D]}
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kdtree.h"
#include "vector_database.h"

static void *must(void *p) {
    if (!p) { fprintf(stderr, "harness: allocation or read failed\n"); exit(2); }
    return p;
}

static Vector make_vec(const double *src, size_t D, const char *tag, size_t i) {
    Vector v;
    memset(&v, 0, sizeof v);
    snprintf(v.uuid, UUID_SIZE, "%s-%zu", tag, i);
    v.dimension = D;
    v.data = (double *)must(malloc(D * sizeof(double)));   /* ownership passes to the store */
    memcpy(v.data, src, D * sizeof(double));
    return v;
}

static void answer_queries(VectorDatabase *db, const double *Q, size_t nq, size_t D, const char *phase) {
    for (size_t i = 0; i < nq; i++) {
        size_t idx = kdtree_nearest(db->kdtree, Q + i * D);             /* compare_handler.c:403 */
        if (idx == (size_t)-1) { printf("%s q%zu none\n", phase, i); continue; }
        Vector *v = vector_db_read(db, idx);                            /* :411 */
        if (v) printf("%s q%zu idx=%zu uuid=%s v0=%a\n", phase, i, idx, v->uuid, v->data[0]);
        else   printf("%s q%zu idx=%zu not-found\n", phase, i, idx);    /* :417-419 */
    }
}

int main(int argc, char **argv) {
    if (argc < 3) return 64;
    FILE *f = (FILE *)must(fopen(argv[1], "rb"));
    size_t hdr[6];
    if (fread(hdr, sizeof(size_t), 6, f) != 6) return 65;
    const size_t n = hdr[0], D = hdr[1], K = hdr[2], nq = hdr[3], np = hdr[4], nops = hdr[5];
    double *rows = (double *)must(malloc(n * D * sizeof(double)));
    double *Q = (double *)must(malloc(nq * D * sizeof(double)));
    size_t *pairs = (size_t *)must(malloc(np * 2 * sizeof(size_t)));
    if (fread(rows, sizeof(double), n * D, f) != n * D || fread(Q, sizeof(double), nq * D, f) != nq * D ||
        fread(pairs, sizeof(size_t), np * 2, f) != np * 2) return 66;

    VectorDatabase *db = (VectorDatabase *)must(vector_db_init(0, K));   /* main.c:351 */
    for (size_t i = 0; i < n; i++) {
        size_t got = vector_db_insert(db, make_vec(rows + i * D, D, "row", i));   /* post_handler.c:333 */
        if (got != i) { printf("insert %zu -> %zu\n", i, got); return 1; }
    }
    printf("size=%zu kdtree=%s dim=%zu\n", db->size, db->kdtree ? "yes" : "no", db->kdtree->dimension);
    answer_queries(db, Q, nq, D, "A");

    for (size_t i = 0; i < np; i++) {                                    /* compare_handler.c:98-160 */
        size_t i1 = pairs[2 * i], i2 = pairs[2 * i + 1];
        if (i1 >= db->size || i2 >= db->size) { printf("pair %zu out-of-bounds\n", i); continue; }
        Vector *v1 = vector_db_read(db, i1), *v2 = vector_db_read(db, i2);
        double c = cosine_similarity(*v1, *v2), e = euclidean_distance(*v1, *v2), d = dot_product(*v1, *v2);
        printf("pair %zu cos=%a euc=%a dot=%a\n", i, c, e, d);
    }

    double *row = (double *)must(malloc(D * sizeof(double)));
    for (size_t i = 0; i < nops; i++) {
        size_t op[2];
        if (fread(op, sizeof(size_t), 2, f) != 2 || fread(row, sizeof(double), D, f) != D) return 67;
        if (op[0] == 1) {
            if (op[1] < db->size) vector_db_update(db, op[1], make_vec(row, D, "upd", i));   /* put_handler.c:155,247 */
        } else {
            if (op[1] < db->size) vector_db_delete(db, op[1]);                                 /* delete_handler.c:84,96 */
        }
    }
    printf("size=%zu after %zu ops\n", db->size, nops);
    answer_queries(db, Q, nq, D, "B");
    Vector *u = vector_db_read_by_uuid(db, "row-3");                     /* get_handler.c:113 */
    printf("by-uuid row-3: %s\n", u ? u->uuid : "none");

    vector_db_save(db, argv[2]);                                         /* main.c:398 */
    VectorDatabase *db2 = vector_db_load(argv[2], K);                    /* main.c:349 */
    printf("reloaded size=%zu\n", db2 ? db2->size : (size_t)0);
    if (db2) {
        answer_queries(db2, Q, nq < 8 ? nq : 8, D, "C");
        vector_db_free(db2);
    }
    vector_db_free(db);
    free(rows); free(Q); free(pairs); free(row);
    fclose(f);
    return 0;
}
