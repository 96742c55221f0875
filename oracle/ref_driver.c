/*
 * ref_driver.c -- thin batch driver over the REFERENCE's own C API.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/svdb_oracle.c header).  This file is
 * ours; it is compiled together with the reference's untouched
 * src/kdtree.c + src/vector_database.c (read in place from /root/reference by
 * oracle/Makefile) into oracle/_ref/libsvdb_ref.so.  It calls nothing but the
 * public functions declared in the reference's include/vector_database.h and
 * include/kdtree.h, exactly as the handlers do:
 *   vector_db_init / vector_db_insert    (main.c:351, post_handler.c:333)
 *   kdtree_nearest(db->kdtree, q)        (compare_handler.c:403, no lock held)
 *   vector_db_read + the three metrics   (compare_handler.c:113-114,153-159)
 * and exports the same cpu_* names as the port so bench.py binds either.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "vector_database.h"   /* the reference's header, via -I/root/reference/include */

void *cpu_build(const double *rows, size_t n, size_t D, size_t K) {
    VectorDatabase *db = vector_db_init(n ? n : 1, K);
    if (!db) return NULL;
    for (size_t i = 0; i < n; i++) {
        Vector v;
        memset(&v, 0, sizeof v);
        v.dimension = D;
        v.data = (double *)malloc(D * sizeof(double));   /* the store takes ownership */
        memcpy(v.data, rows + i * D, D * sizeof(double));
        if (vector_db_insert(db, v) == (size_t)-1) { vector_db_free(db); return NULL; }
    }
    return db;
}

void cpu_free(void *h) { vector_db_free((VectorDatabase *)h); }

typedef struct { VectorDatabase *db; const double *Q; size_t stride, lo, hi; size_t *out; } nn_job;
static void *nn_worker(void *p) {
    nn_job *j = (nn_job *)p;
    for (size_t i = j->lo; i < j->hi; i++) j->out[i] = kdtree_nearest(j->db->kdtree, j->Q + i * j->stride);
    return NULL;
}

void cpu_nearest_batch(void *h, const double *Q, size_t nq, size_t stride, size_t nthreads, size_t *out) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nq) nthreads = nq ? nq : 1;
    pthread_t *th = (pthread_t *)malloc(nthreads * sizeof(pthread_t));
    nn_job *jobs = (nn_job *)malloc(nthreads * sizeof(nn_job));
    for (size_t t = 0; t < nthreads; t++) {
        jobs[t] = (nn_job){ (VectorDatabase *)h, Q, stride, nq * t / nthreads, nq * (t + 1) / nthreads, out };
        if (nthreads == 1) nn_worker(&jobs[t]); else pthread_create(&th[t], NULL, nn_worker, &jobs[t]);
    }
    if (nthreads > 1) for (size_t t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

typedef struct { VectorDatabase *db; int metric; const size_t *i1, *i2; size_t lo, hi; float *out; } cmp_job;
static void *cmp_worker(void *p) {
    cmp_job *j = (cmp_job *)p;
    for (size_t i = j->lo; i < j->hi; i++) {
        /* rows are read without the mutex here: the batch is read-only */
        const Vector a = j->db->vectors[j->i1[i]], b = j->db->vectors[j->i2[i]];
        j->out[i] = j->metric == 0 ? cosine_similarity(a, b)
                  : j->metric == 1 ? euclidean_distance(a, b) : dot_product(a, b);
    }
    return NULL;
}

void cpu_compare_batch(void *h, int metric, const size_t *i1, const size_t *i2, size_t n, size_t nthreads, float *out) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n ? n : 1;
    pthread_t *th = (pthread_t *)malloc(nthreads * sizeof(pthread_t));
    cmp_job *jobs = (cmp_job *)malloc(nthreads * sizeof(cmp_job));
    for (size_t t = 0; t < nthreads; t++) {
        jobs[t] = (cmp_job){ (VectorDatabase *)h, metric, i1, i2, n * t / nthreads, n * (t + 1) / nthreads, out };
        if (nthreads == 1) cmp_worker(&jobs[t]); else pthread_create(&th[t], NULL, cmp_worker, &jobs[t]);
    }
    if (nthreads > 1) for (size_t t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

const char *cpu_kind(void) { return "reference"; }
