/*
 * svdb_dropin.h -- the reference's L1 C API, served by the B200 engine.
 *
 * libsvdb_b200.so exports the same sixteen functions, with the same signatures,
 * sentinels and ownership rules, as the reference declares in
 *     include/vector_database.h:39-135   (vector_db_*, cosine_similarity,
 *                                         euclidean_distance, dot_product)
 *     include/kdtree.h:32-57             (kdtree_create/insert/free/nearest)
 * so the reference's handlers (src/compare_handler.c:113-114,153-159,403,
 * src/post_handler.c:333, src/put_handler.c:247, src/delete_handler.c:96,
 * src/get_handler.c:85-118, src/main.c:349-402) link against it unchanged.
 *
 * Callers reach INTO these structs (db->size, db->vectors[i].uuid/.data,
 * db->kdtree, vec->dimension), so the leading members below have
 * the reference's order and types.  The GPU state hangs off the allocation
 * behind them; never allocate these structs yourself.
 *
 * What differs, on purpose:
 *   - rows and kd-points live in HBM; the host keeps the caller-visible copy of
 *     each row (the reference's own malloc'ed vec.data, whose ownership
 *     vector_db_insert/update take exactly as before);
 *   - KDTree.root is non-NULL once the log is non-empty but is not a walkable
 *     tree.  The reference's non-static helpers that no header declares and nothing
 *     outside their own file calls -- kdtree_create_node, kdtree_insert_rec,
 *     kdtree_free_rec, kdtree_nearest_rec (src/kdtree.c:15,47,98,131) and the unused
 *     qsort comparator `compare` (src/vector_database.c:361) -- are deliberately NOT
 *     exported: they take or return KDTreeNode pointers into a heap tree that does
 *     not exist here;
 *   - no progress chatter on stdout (the reference prints per insert level);
 *   - kd_dim > vector dimension is rejected ((size_t)-1) instead of reading past
 *     the row (src/kdtree.c:26-28);
 *   - a CUDA failure surfaces as the call's usual sentinel plus a line on stderr.
 */
#ifndef SVDB_DROPIN_H
#define SVDB_DROPIN_H

#include <pthread.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UUID_SIZE 37

typedef struct KDTreeNode {
    double *point;
    size_t index;
    struct KDTreeNode *left, *right;
} KDTreeNode;

typedef struct KDTree {
    KDTreeNode *root;
    size_t dimension;
} KDTree;

typedef struct Vector {
    char uuid[UUID_SIZE];
    size_t dimension;
    double *data;
} Vector;

typedef struct VectorDatabase {
    Vector *vectors;
    size_t size;
    size_t capacity;
    KDTree *kdtree;
    pthread_mutex_t mutex;
} VectorDatabase;

/* -- include/kdtree.h:32-57 -- */
KDTree *kdtree_create(size_t dimension);
void kdtree_insert(KDTree *tree, const double *point, size_t index);
void kdtree_free(KDTree *tree);
size_t kdtree_nearest(KDTree *tree, const double *point);
/* One documented difference in kdtree_nearest: between DISTINCT kd-points at exactly equal minimal distance the reference
 * returns whichever its insertion-order tree reaches first (kdtree.c:139, :147-159), and this library reproduces that by
 * keeping a tree of the same shape on the GPU -- up to a depth of 8192 (option "tree.max_depth").  Input inserted in
 * monotone order (points along a line) makes the reference's tree a list; past that depth the shape is dropped
 * (svdb_stats.tree_dropped = 1) and such exact ties resolve to the earliest insert instead.  Unique minima and copies of
 * one point are unaffected. */

/* -- include/vector_database.h:39-135 -- */
VectorDatabase *vector_db_init(size_t initial_capacity, size_t dimension);
void vector_db_free(VectorDatabase *db);
size_t vector_db_insert(VectorDatabase *db, Vector vec);
Vector *vector_db_read(VectorDatabase *db, size_t index);
Vector *vector_db_read_by_uuid(VectorDatabase *db, const char *uuid);
void vector_db_update(VectorDatabase *db, size_t index, Vector vec);
void vector_db_delete(VectorDatabase *db, size_t index);
void vector_db_save(VectorDatabase *db, const char *filename);
VectorDatabase *vector_db_load(const char *filename, size_t dimension);
float cosine_similarity(Vector vec1, Vector vec2);
float euclidean_distance(Vector vec1, Vector vec2);
float dot_product(Vector vec1, Vector vec2);

/* -- batched extensions behind the same semantics (last row) -- */
#define SVDB_DROPIN_MAX_K 24     /* largest k of kdtree_nearest_batch (= SVDB_MAX_K) */
/* k nearest log entries for each of nq queries (ldq doubles apart); index_out is nq x k. */
int kdtree_nearest_batch(KDTree *tree, const double *queries, size_t nq, size_t ldq, size_t k,
                         size_t *index_out, double *dist_out);
/* metric: 0 cosine, 1 euclidean, 2 dot; out[i] for rows (index1[i], index2[i]); -1.0f if out of range. */
int vector_db_compare_batch(VectorDatabase *db, int metric, const size_t *index1, const size_t *index2,
                            size_t n, float *out);

#ifdef __cplusplus
}
#endif
#endif /* SVDB_DROPIN_H */
