/*
 * svdb_oracle.c -- CPU restatement of the simple-vector-db similarity hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (simple-vector-db_b200/,
 * include/) may link, import or execute this file.  It exists so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * have an independent CPU checker for the CUDA path.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 *   (a) against the reference's own L1 code compiled unmodified into
 *       oracle/_ref/libsvdb_ref.so (see oracle/Makefile), and
 *   (b) against tests/golden/*.npz, which were produced by that library
 *       (tests/golden/make_golden.py).
 * The reference ships no golden vectors or tests of its own.
 *
 * What is restated (reference file:line, relative to /root/reference):
 *   - append of a (kd-point, index) entry and its BST placement
 *       src/kdtree.c:15-35 (node = first K coords + caller's index)
 *       src/kdtree.c:47-62 (descent: cd = depth % K; strictly-less goes left,
 *                           equal goes right)
 *   - exact 1-NN with near-side-first traversal and hyper-plane pruning
 *       src/kdtree.c:131-162, entry src/kdtree.c:171-178
 *   - the three /compare metrics with their float accumulators
 *       src/vector_database.c:301-313, 322-333, 342-352
 *   - store semantics that decide WHICH points are searchable
 *       src/vector_database.c:81-119 (insert appends entry with index=size)
 *       src/vector_database.c:169-177 (update re-appends, stale entry stays)
 *       src/vector_database.c:185-195 (delete shifts rows, log untouched)
 *
 * The data structure is deliberately not the reference's (no per-node malloc,
 * children are sequence numbers in flat arrays, traversal uses an explicit
 * stack) -- only the algorithm and its arithmetic order are the same.
 * Build with -ffp-contract=off: the reference's Makefile uses no -O/-march,
 * so every a*b+c is a rounded multiply followed by a rounded add.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NONE ((int64_t)-1)

typedef struct orc_log {
    size_t K;        /* kd_dim: coordinates that take part in the distance   */
    size_t n, cap;   /* entries appended so far / allocated                   */
    double *pts;     /* n x K, entry s at pts + s*K                           */
    size_t *idx;     /* index the entry was appended with                     */
    int64_t *lo;     /* child on the "strictly less" side, ORC_NONE if none   */
    int64_t *hi;     /* child on the "greater or equal" side                  */
} orc_log;

orc_log *orc_log_create(size_t K) {
    orc_log *L = (orc_log *)calloc(1, sizeof *L);
    if (!L) return NULL;
    L->K = K;
    return L;
}

void orc_log_free(orc_log *L) {
    if (!L) return;
    free(L->pts); free(L->idx); free(L->lo); free(L->hi); free(L);
}

size_t orc_log_size(const orc_log *L) { return L ? L->n : 0; }

static int orc_log_grow(orc_log *L) {
    size_t nc = L->cap ? L->cap * 2 : 1024;
    double *p = (double *)realloc(L->pts, nc * L->K * sizeof(double));
    if (!p) return -1;
    L->pts = p;
    size_t *ix = (size_t *)realloc(L->idx, nc * sizeof(size_t));
    if (!ix) return -1;
    L->idx = ix;
    int64_t *lo = (int64_t *)realloc(L->lo, nc * sizeof(int64_t));
    if (!lo) return -1;
    L->lo = lo;
    int64_t *hi = (int64_t *)realloc(L->hi, nc * sizeof(int64_t));
    if (!hi) return -1;
    L->hi = hi;
    L->cap = nc;
    return 0;
}

/* kdtree.c:87-91 -> :47-62 -> :15-35.  Returns the entry's sequence number. */
int64_t orc_log_append(orc_log *L, const double *point, size_t index) {
    if (!L) return -1;
    if (L->n == L->cap && orc_log_grow(L) != 0) return -1;
    const size_t K = L->K, s = L->n;
    memcpy(L->pts + s * K, point, K * sizeof(double));
    L->idx[s] = index;
    L->lo[s] = L->hi[s] = ORC_NONE;
    if (s > 0) {
        int64_t cur = 0;
        size_t depth = 0;
        for (;;) {
            const size_t cd = depth % K;
            int64_t *slot = (point[cd] < L->pts[(size_t)cur * K + cd]) ? &L->lo[cur] : &L->hi[cur];
            if (*slot == ORC_NONE) { *slot = (int64_t)s; break; }
            cur = *slot;
            ++depth;
        }
    }
    L->n = s + 1;
    return (int64_t)s;
}

/* kdtree.c:134-137: sequential sum, each term (n-q)*(n-q) rounded before the add. */
double orc_sqdist(const double *node, const double *q, size_t K) {
    double d = 0;
    for (size_t i = 0; i < K; i++) {
        const double t = node[i] - q[i];
        d += t * t;
    }
    return d;
}

/*
 * kdtree.c:131-162 with the recursion unrolled onto an explicit stack.
 * A frame is pushed for the far child BEFORE descending to the near child and
 * carries the plane distance; the prune test (:157, strict <) is evaluated
 * when the frame is popped, i.e. after the whole near subtree has been
 * searched -- the same moment the recursive code evaluates it.
 * Returns the winning sequence number (ORC_NONE when the log is empty).
 */
int64_t orc_tree_nearest_seq(const orc_log *L, const double *q, double *best_out, size_t *visited_out) {
    if (!L || L->n == 0) return ORC_NONE;
    const size_t K = L->K;
    typedef struct { int64_t node; size_t depth; double plane; } frame;
    size_t scap = 256, sp = 0, visited = 0;
    frame *st = (frame *)malloc(scap * sizeof(frame));
    double best = INFINITY;
    int64_t best_s = ORC_NONE;
    int64_t cur = 0;
    size_t depth = 0;
    for (;;) {
        while (cur != ORC_NONE) {
            const double *p = L->pts + (size_t)cur * K;
            const double d = orc_sqdist(p, q, K);
            ++visited;
            if (d < best) { best = d; best_s = cur; }
            const size_t cd = depth % K;
            int64_t near_c, far_c;
            if (q[cd] < p[cd]) { near_c = L->lo[cur]; far_c = L->hi[cur]; }
            else               { near_c = L->hi[cur]; far_c = L->lo[cur]; }
            if (sp == scap) { scap *= 2; st = (frame *)realloc(st, scap * sizeof(frame)); }
            st[sp].node = far_c;
            st[sp].depth = depth + 1;
            st[sp].plane = (q[cd] - p[cd]) * (q[cd] - p[cd]);
            ++sp;
            cur = near_c;
            ++depth;
        }
        int found = 0;
        while (sp > 0) {
            --sp;
            if (st[sp].plane < best && st[sp].node != ORC_NONE) {
                cur = st[sp].node; depth = st[sp].depth; found = 1;
                break;
            }
        }
        if (!found) break;
    }
    free(st);
    if (best_out) *best_out = best;
    if (visited_out) *visited_out = visited;
    return best_s;
}

/* kdtree.c:171-178: (size_t)-1 for an empty tree, else the carried index. */
size_t orc_tree_nearest(const orc_log *L, const double *q) {
    const int64_t s = orc_tree_nearest_seq(L, q, NULL, NULL);
    return s == ORC_NONE ? (size_t)-1 : L->idx[s];
}

/*
 * Flat restatement of the same search: the k smallest entries under the total
 * order (d, seq), d exactly as kdtree.c:134-137 computes it.  For k = 1 this
 * names the same entry as the tree wherever the minimum is unique or tied only
 * between identical kd-points (SURVEY.md s8a tie rule); it is what the GPU scan
 * is compared with at sizes where replaying the tree is too slow.
 * Outputs are filled up to min(k, n); returns that count.
 */
size_t orc_flat_topk(const orc_log *L, const double *q, size_t k,
                     int64_t *out_seq, size_t *out_idx, double *out_d) {
    if (!L || L->n == 0 || k == 0) return 0;
    const size_t K = L->K;
    size_t m = 0;
    for (size_t s = 0; s < L->n; s++) {
        const double d = orc_sqdist(L->pts + s * K, q, K);
        if (m == k && !(d < out_d[m - 1])) continue;   /* later seq loses ties */
        size_t pos = m < k ? m : k - 1;
        while (pos > 0 && d < out_d[pos - 1]) {
            out_d[pos] = out_d[pos - 1]; out_seq[pos] = out_seq[pos - 1]; out_idx[pos] = out_idx[pos - 1];
            --pos;
        }
        out_d[pos] = d; out_seq[pos] = (int64_t)s; out_idx[pos] = L->idx[s];
        if (m < k) ++m;
    }
    return m;
}

/* ---- /compare metrics, vector_database.c:301-352 ------------------------- */

/* The float accumulator of "acc += a[i]*b[i]" (:308-310, :349): the product and
 * the add happen in double, the result is rounded to float every step. */
float orc_dot_f(const double *a, const double *b, size_t D) {
    float acc = 0.0f;
    for (size_t i = 0; i < D; i++) acc = (float)((double)acc + a[i] * b[i]);
    return acc;
}

float orc_dot_product(const double *a, const double *b, size_t D) { return orc_dot_f(a, b, D); }

float orc_cosine_similarity(const double *a, const double *b, size_t D) {
    const float dot = orc_dot_f(a, b, D), na = orc_dot_f(a, a, D), nb = orc_dot_f(b, b, D);
    return (float)((double)dot / (sqrt((double)na) * sqrt((double)nb)));   /* :312 */
}

float orc_euclidean_distance(const double *a, const double *b, size_t D) {
    float sum = 0.0f;
    for (size_t i = 0; i < D; i++) {
        const float diff = (float)(a[i] - b[i]);   /* :329 */
        sum = sum + diff * diff;                    /* :330, float multiply and add */
    }
    return (float)sqrt((double)sum);                /* :332 */
}

/* ---- store model: which (point, index) pairs are searchable ------------- */

typedef struct orc_db {
    size_t D, size, cap;
    double **rows;   /* rows[i] = current D doubles of index i (owned) */
    orc_log *log;
} orc_db;

orc_db *orc_db_create(size_t D, size_t K) {
    orc_db *db = (orc_db *)calloc(1, sizeof *db);
    if (!db) return NULL;
    db->D = D;
    db->log = orc_log_create(K);
    return db;
}

void orc_db_free(orc_db *db) {
    if (!db) return;
    for (size_t i = 0; i < db->size; i++) free(db->rows[i]);
    free(db->rows);
    orc_log_free(db->log);
    free(db);
}

size_t orc_db_size(const orc_db *db) { return db->size; }
orc_log *orc_db_log(orc_db *db) { return db->log; }
const double *orc_db_row(const orc_db *db, size_t i) { return i < db->size ? db->rows[i] : NULL; }

static double *orc_dup(const double *v, size_t D) {
    double *r = (double *)malloc(D * sizeof(double));
    if (r) memcpy(r, v, D * sizeof(double));
    return r;
}

/* vector_database.c:81-119: entry carries index = size before the increment. */
size_t orc_db_insert(orc_db *db, const double *v) {
    if (db->size == db->cap) {
        size_t nc = db->cap ? db->cap * 2 : 10;
        double **r = (double **)realloc(db->rows, nc * sizeof(double *));
        if (!r) return (size_t)-1;
        db->rows = r; db->cap = nc;
    }
    db->rows[db->size] = orc_dup(v, db->D);
    orc_log_append(db->log, v, db->size);
    return db->size++;
}

/* :169-177: row replaced, a NEW entry appended with the same index; the old
 * entry stays searchable. Out of range: silent no-op. */
void orc_db_update(orc_db *db, size_t index, const double *v) {
    if (index >= db->size) return;
    free(db->rows[index]);
    db->rows[index] = orc_dup(v, db->D);
    orc_log_append(db->log, v, index);
}

/* :185-195: rows above shift down by one; the log is not touched. */
void orc_db_delete(orc_db *db, size_t index) {
    if (index >= db->size) return;
    free(db->rows[index]);
    memmove(db->rows + index, db->rows + index + 1, (db->size - 1 - index) * sizeof(double *));
    db->size--;
}

size_t orc_db_nearest(const orc_db *db, const double *q) { return orc_tree_nearest(db->log, q); }

/* metric: 0 cosine, 1 euclidean, 2 dot */
float orc_db_compare(const orc_db *db, int metric, size_t i1, size_t i2) {
    const double *a = db->rows[i1], *b = db->rows[i2];
    switch (metric) {
        case 0: return orc_cosine_similarity(a, b, db->D);
        case 1: return orc_euclidean_distance(a, b, db->D);
        default: return orc_dot_product(a, b, db->D);
    }
}

/* ---- uniform "cpu_*" driver, same names as oracle/ref_driver.c ----------- */
/* Lets bench.py time either library through one binding. */

void *cpu_build(const double *rows, size_t n, size_t D, size_t K) {
    orc_db *db = orc_db_create(D, K);
    if (!db) return NULL;
    for (size_t i = 0; i < n; i++) orc_db_insert(db, rows + i * D);
    return db;
}

void cpu_free(void *h) { orc_db_free((orc_db *)h); }

typedef struct { const orc_db *db; const double *Q; size_t stride, lo, hi; size_t *out; } nn_job;
static void *nn_worker(void *p) {
    nn_job *j = (nn_job *)p;
    for (size_t i = j->lo; i < j->hi; i++) j->out[i] = orc_db_nearest(j->db, j->Q + i * j->stride);
    return NULL;
}

/* Independent queries on a read-only tree, one contiguous slice per thread
 * (the reference handler holds no lock there: compare_handler.c:403). */
void cpu_nearest_batch(void *h, const double *Q, size_t nq, size_t stride, size_t nthreads, size_t *out) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nq) nthreads = nq ? nq : 1;
    pthread_t *th = (pthread_t *)malloc(nthreads * sizeof(pthread_t));
    nn_job *jobs = (nn_job *)malloc(nthreads * sizeof(nn_job));
    for (size_t t = 0; t < nthreads; t++) {
        jobs[t] = (nn_job){ (const orc_db *)h, Q, stride, nq * t / nthreads, nq * (t + 1) / nthreads, out };
        if (nthreads == 1) nn_worker(&jobs[t]); else pthread_create(&th[t], NULL, nn_worker, &jobs[t]);
    }
    if (nthreads > 1) for (size_t t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

typedef struct { const orc_db *db; int metric; const size_t *i1, *i2; size_t lo, hi; float *out; } cmp_job;
static void *cmp_worker(void *p) {
    cmp_job *j = (cmp_job *)p;
    for (size_t i = j->lo; i < j->hi; i++) j->out[i] = orc_db_compare(j->db, j->metric, j->i1[i], j->i2[i]);
    return NULL;
}

void cpu_compare_batch(void *h, int metric, const size_t *i1, const size_t *i2, size_t n, size_t nthreads, float *out) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n ? n : 1;
    pthread_t *th = (pthread_t *)malloc(nthreads * sizeof(pthread_t));
    cmp_job *jobs = (cmp_job *)malloc(nthreads * sizeof(cmp_job));
    for (size_t t = 0; t < nthreads; t++) {
        jobs[t] = (cmp_job){ (const orc_db *)h, metric, i1, i2, n * t / nthreads, n * (t + 1) / nthreads, out };
        if (nthreads == 1) cmp_worker(&jobs[t]); else pthread_create(&th[t], NULL, cmp_worker, &jobs[t]);
    }
    if (nthreads > 1) for (size_t t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

const char *cpu_kind(void) { return "port"; }
