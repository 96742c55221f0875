// dropin.cu -- the reference's L1 C API (include/svdb_dropin.h) on top of the engine.
//
// Mirrors, function by function, src/vector_database.c and src/kdtree.c of the reference:
// same signatures, sentinels, ownership and (bug-compatible) store semantics, with the
// data the hot path reads held in HBM:
//   KDTree          -> a LOG_ONLY engine: the append-only (kd-point, index) log
//   VectorDatabase  -> host array of Vector (callers index into it) + a NO_LOG engine
//                      holding the rows for the /compare kernels
// When the store's tree measures whole rows (kd_dim == row dimension, i.e. the server was started
// with -d D) the two share ONE engine whose log aliases its rows: no second copy in HBM.
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/svdb_b200.h"
#include "../../include/svdb_dropin.h"

namespace {

int env_device() {
    const char *v = getenv("SVDB_DEVICE");
    return v ? atoi(v) : 0;
}

struct TreeImpl {
    KDTree pub;              // must be first: callers hold KDTree*
    svdb_engine *eng;        // created on first insert
    bool shared;             // eng belongs to the VectorDatabase this tree was created for
    bool attached;           // created by vector_db_init (may still become shared)
    KDTreeNode root_node;    // what pub.root points to once the log is non-empty
    std::vector<double> root_point;
};

struct DbImpl {
    VectorDatabase pub;      // must be first
    svdb_engine *rows;       // created on first insert (needs the row dimension): NO_LOG, or rows + log when unified
    bool unified;            // rows engine also carries the tree's log (kd_dim == row dimension)
    size_t row_dim;
    std::vector<unsigned char> dim_ok;   // per index: row has the engine's dimension
};

void complain(const char *where) { fprintf(stderr, "svdb_b200: %s: %s\n", where, svdb_last_error()); }

bool tree_engine(TreeImpl *t) {
    if (t->eng) return true;
    svdb_config c;
    memset(&c, 0, sizeof c);
    c.dimension = c.kd_dim = t->pub.dimension;
    c.device = env_device();
    c.flags = SVDB_FLAG_LOG_ONLY;
    if (svdb_engine_create(&c, &t->eng) != SVDB_OK) {
        complain("kdtree: engine");
        t->eng = nullptr;
        return false;
    }
    return true;
}

bool db_engine(DbImpl *d, size_t dim) {
    if (d->rows) return true;
    TreeImpl *t = reinterpret_cast<TreeImpl *>(d->pub.kdtree);
    // one engine for rows and log when the tree measures whole rows and nothing went into it yet
    const bool unify = t && t->attached && !t->eng && t->pub.dimension == dim;
    svdb_config c;
    memset(&c, 0, sizeof c);
    c.dimension = dim;
    c.kd_dim = unify ? dim : 1;
    c.device = env_device();
    c.flags = unify ? 0u : SVDB_FLAG_NO_LOG;
    if (svdb_engine_create(&c, &d->rows) != SVDB_OK) {
        complain("vector_db: engine");
        d->rows = nullptr;
        return false;
    }
    d->row_dim = dim;
    d->unified = unify;
    if (unify) {
        t->eng = d->rows;
        t->shared = true;
    }
    return true;
}

void tree_note_root(TreeImpl *t, const double *point, size_t index) {
    if (t->pub.root) return;
    t->root_point.assign(point, point + t->pub.dimension);
    t->root_node.point = t->root_point.data();
    t->root_node.index = index;
    t->root_node.left = t->root_node.right = NULL;
    t->pub.root = &t->root_node;
}

// Double the slot array of the host-visible Vector table; false (and a line on stderr) when
// the byte count would overflow or realloc fails.
bool grow_slots(VectorDatabase *db) {
    const size_t limit = SIZE_MAX / sizeof(Vector) / 2;
    Vector *nv = db->capacity >= 1 && db->capacity <= limit
                     ? (Vector *)realloc(db->vectors, 2 * db->capacity * sizeof(Vector))
                     : NULL;
    if (!nv) {
        fprintf(stderr, "svdb_b200: cannot grow the vector table beyond %zu slots\n", db->capacity);
        return false;
    }
    db->vectors = nv;
    db->capacity *= 2;
    return true;
}

float metric_call(int metric, const Vector &a, const Vector &b) {
    if (a.dimension != b.dimension) {
        fprintf(stderr, "Vectors have different dimensions\n");   // vector_database.c:303
        return -1.0f;
    }
    float out = -1.0f;
    if (a.dimension == 0) {   // the loops of :307/:328/:348 never run
        if (metric == SVDB_COSINE) return 0.0f / (0.0f * 0.0f);
        return 0.0f;
    }
    // by-value vectors need not belong to any store (compare_handler.c:153-159 passes copies):
    // upload both, K4 + K3 on the device
    if (svdb_compare_vectors(env_device(), metric, a.data, b.data, a.dimension, &out) != SVDB_OK) {
        complain("compare");
        return -1.0f;
    }
    return out;
}

}  // namespace

extern "C" {

// ---- kdtree.c:70-78 ----
KDTree *kdtree_create(size_t dimension) {
    TreeImpl *t = new (std::nothrow) TreeImpl();
    if (!t) return NULL;
    t->pub.root = NULL;
    t->pub.dimension = dimension;
    t->eng = nullptr;
    t->shared = t->attached = false;
    return &t->pub;
}

// ---- kdtree.c:87-91 (+ :15-35 copy of the first `dimension` coordinates) ----
void kdtree_insert(KDTree *tree, const double *point, size_t index) {
    if (tree == NULL) return;
    TreeImpl *t = reinterpret_cast<TreeImpl *>(tree);
    if (tree->dimension == 0 || !point || !tree_engine(t)) return;
    if (svdb_append_kdpoints(t->eng, point, &index, 1, tree->dimension) != SVDB_OK) {
        complain("kdtree_insert");
        return;
    }
    tree_note_root(t, point, index);
}

// ---- kdtree.c:112-118 ----
void kdtree_free(KDTree *tree) {
    if (!tree) return;
    TreeImpl *t = reinterpret_cast<TreeImpl *>(tree);
    if (t->eng && !t->shared) svdb_engine_destroy(t->eng);
    tree->root = NULL;
    delete t;
}

// ---- kdtree.c:171-178 ----
size_t kdtree_nearest(KDTree *tree, const double *point) {
    if (tree == NULL || tree->root == NULL) return (size_t)-1;
    TreeImpl *t = reinterpret_cast<TreeImpl *>(tree);
    size_t idx = (size_t)-1;
    if (svdb_nearest_batch(t->eng, point, 1, tree->dimension, 1, &idx, NULL, NULL) != SVDB_OK) {
        complain("kdtree_nearest");
        return (size_t)-1;
    }
    return idx;
}

int kdtree_nearest_batch(KDTree *tree, const double *queries, size_t nq, size_t ldq, size_t k, size_t *index_out,
                         double *dist_out) {
    if (!tree || !queries || !index_out || k < 1) return SVDB_ERR_ARG;
    TreeImpl *t = reinterpret_cast<TreeImpl *>(tree);
    if (!tree->root) {
        for (size_t i = 0; i < nq * k; i++) {
            index_out[i] = (size_t)-1;
            if (dist_out) dist_out[i] = __builtin_inf();
        }
        return SVDB_OK;
    }
    int rc = svdb_nearest_batch(t->eng, queries, nq, ldq, k, index_out, dist_out, NULL);
    if (rc != SVDB_OK) complain("kdtree_nearest_batch");
    return rc;
}

// ---- vector_database.c:17-54 ----
VectorDatabase *vector_db_init(size_t initial_capacity, size_t dimension) {
    DbImpl *d = new (std::nothrow) DbImpl();
    if (!d) {
        fprintf(stderr, "Failed to allocate memory for database\n");
        return NULL;
    }
    d->rows = nullptr;
    d->unified = false;
    d->row_dim = 0;
    d->pub.size = 0;
    d->pub.capacity = initial_capacity > 0 ? initial_capacity : 10;
    d->pub.vectors = (Vector *)malloc(d->pub.capacity * sizeof(Vector));
    if (!d->pub.vectors) {
        fprintf(stderr, "Failed to allocate memory for vectors\n");
        delete d;
        return NULL;
    }
    d->pub.kdtree = kdtree_create(dimension);
    if (!d->pub.kdtree) {
        fprintf(stderr, "Failed to create KDTree\n");
        free(d->pub.vectors);
        delete d;
        return NULL;
    }
    reinterpret_cast<TreeImpl *>(d->pub.kdtree)->attached = true;
    if (pthread_mutex_init(&d->pub.mutex, NULL) != 0) {
        fprintf(stderr, "Failed to initialize mutex\n");
        kdtree_free(d->pub.kdtree);
        free(d->pub.vectors);
        delete d;
        return NULL;
    }
    return &d->pub;
}

// ---- vector_database.c:61-73 ----
void vector_db_free(VectorDatabase *db) {
    if (!db) return;
    DbImpl *d = reinterpret_cast<DbImpl *>(db);
    for (size_t i = 0; i < db->size; ++i) {
        free(db->vectors[i].data);
    }
    kdtree_free(db->kdtree);
    free(db->vectors);
    if (d->rows) svdb_engine_destroy(d->rows);
    pthread_mutex_destroy(&db->mutex);
    delete d;
}

// ---- vector_database.c:81-119 ----
size_t vector_db_insert(VectorDatabase *db, Vector vec) {
    DbImpl *d = reinterpret_cast<DbImpl *>(db);
    pthread_mutex_lock(&db->mutex);
    if (db->size == db->capacity && !grow_slots(db)) {   // :85-103, doubling
        pthread_mutex_unlock(&db->mutex);
        return (size_t)-1;
    }
    if (!db->kdtree || !vec.data || vec.dimension < db->kdtree->dimension) {
        fprintf(stderr, "svdb_b200: vector_db_insert: no tree, or vector (dimension %zu) shorter than kd_dim\n",
                vec.dimension);
        pthread_mutex_unlock(&db->mutex);
        return (size_t)-1;
    }
    if (!db_engine(d, vec.dimension)) {
        pthread_mutex_unlock(&db->mutex);
        return (size_t)-1;
    }
    // HBM copy of the row for /compare.  A row of another dimension keeps its slot: unified stores take
    // its first row_dim (= kd_dim) coordinates -- exactly what the tree measures -- others a zero row;
    // either way /compare on it answers the mismatch sentinel (dim_ok).
    const bool same = vec.dimension == d->row_dim;
    int rc;
    if (same || d->unified) {
        rc = svdb_insert_batch(d->rows, vec.data, 1, vec.dimension, NULL);
    } else {
        std::vector<double> z(d->row_dim, 0.0);
        rc = svdb_insert_batch(d->rows, z.data(), 1, d->row_dim, NULL);
    }
    if (rc != SVDB_OK) {
        complain("vector_db_insert");
        pthread_mutex_unlock(&db->mutex);
        return (size_t)-1;
    }
    d->dim_ok.push_back(same ? 1 : 0);
    vec.uuid[UUID_SIZE - 1] = '\0';
    db->vectors[db->size] = vec;                      // takes ownership of vec.data (:113)
    if (d->unified) tree_note_root(reinterpret_cast<TreeImpl *>(db->kdtree), vec.data, db->size);   // the insert above logged it
    else kdtree_insert(db->kdtree, vec.data, db->size);                                              // :114
    const size_t index = db->size++;
    pthread_mutex_unlock(&db->mutex);
    return index;
}

// ---- vector_database.c:128-136 ----
Vector *vector_db_read(VectorDatabase *db, size_t index) {
    pthread_mutex_lock(&db->mutex);
    Vector *vec = NULL;
    if (index < db->size) vec = &db->vectors[index];
    pthread_mutex_unlock(&db->mutex);
    return vec;
}

// ---- vector_database.c:145-159 ----
Vector *vector_db_read_by_uuid(VectorDatabase *db, const char *uuid) {
    pthread_mutex_lock(&db->mutex);
    Vector *vec = NULL;
    for (size_t i = 0; i < db->size; ++i) {
        if (strncmp(db->vectors[i].uuid, uuid, UUID_SIZE) == 0) {
            vec = &db->vectors[i];
            break;
        }
    }
    pthread_mutex_unlock(&db->mutex);
    return vec;
}

// ---- vector_database.c:169-177: the old kd-point stays searchable ----
void vector_db_update(VectorDatabase *db, size_t index, Vector vec) {
    DbImpl *d = reinterpret_cast<DbImpl *>(db);
    pthread_mutex_lock(&db->mutex);
    if (index < db->size) {
        if (!vec.data || !db->kdtree || vec.dimension < db->kdtree->dimension) {
            fprintf(stderr, "svdb_b200: vector_db_update: vector shorter than kd_dim, ignored\n");
            pthread_mutex_unlock(&db->mutex);
            return;
        }
        // the HBM side first: the host table changes only once the engine has taken the update, so a failure (device out
        // of memory, a CUDA error) leaves host table, kd log and rows engine consistent -- the index keeps its old row
        const bool same = d->rows && vec.dimension == d->row_dim;
        if (d->rows) {
            std::vector<double> z;
            const double *src = vec.data;
            if (!same && !d->unified) {
                z.assign(d->row_dim, 0.0);
                src = z.data();
            }
            if (svdb_update_batch(d->rows, &index, src, 1, d->row_dim) != SVDB_OK) {
                complain("vector_db_update");
                free(vec.data);                       // ours since the call (vector_database.c:172 frees the OLD row; we keep it)
                pthread_mutex_unlock(&db->mutex);
                return;
            }
            d->dim_ok[index] = same ? 1 : 0;
        }
        if (!d->unified) kdtree_insert(db->kdtree, vec.data, index);   // :174; unified: the update above re-appended
        free(db->vectors[index].data);
        db->vectors[index] = vec;
    }
    pthread_mutex_unlock(&db->mutex);
}

// ---- vector_database.c:185-195: rows shift down, the kd log is left alone ----
void vector_db_delete(VectorDatabase *db, size_t index) {
    DbImpl *d = reinterpret_cast<DbImpl *>(db);
    pthread_mutex_lock(&db->mutex);
    if (index < db->size) {
        free(db->vectors[index].data);
        memmove(&db->vectors[index], &db->vectors[index + 1], (db->size - 1 - index) * sizeof(Vector));
        db->size--;
        if (d->rows) {
            if (svdb_delete_batch(d->rows, &index, 1) != SVDB_OK) complain("vector_db_delete");
            d->dim_ok.erase(d->dim_ok.begin() + index);
        }
    }
    pthread_mutex_unlock(&db->mutex);
}

// ---- vector_database.c:203-229: u64 count, then {char[37] uuid, u64 dim, f64[dim]} per row ----
void vector_db_save(VectorDatabase *db, const char *filename) {
    pthread_mutex_lock(&db->mutex);
    FILE *file = fopen(filename, "wb");
    if (!file) {
        perror("Failed to open file for writing");
        pthread_mutex_unlock(&db->mutex);
        return;
    }
    fwrite(&db->size, sizeof(size_t), 1, file);
    for (size_t i = 0; i < db->size; ++i) {
        const Vector *v = &db->vectors[i];
        if (v->dimension == 0 || v->data == NULL) {
            fprintf(stderr, "Invalid vector at index %zu, skipping\n", i);   // :215-218 (count is not corrected)
            continue;
        }
        fwrite(v->uuid, sizeof(char), UUID_SIZE, file);
        fwrite(&v->dimension, sizeof(size_t), 1, file);
        fwrite(v->data, sizeof(double), v->dimension, file);
    }
    fclose(file);
    pthread_mutex_unlock(&db->mutex);
}

// ---- vector_database.c:237-292: rows are re-inserted in index order ----
VectorDatabase *vector_db_load(const char *filename, size_t dimension) {
    FILE *file = fopen(filename, "rb");
    if (!file) {
        perror("Failed to open file for reading");
        return NULL;
    }
    size_t count = 0;
    if (fread(&count, sizeof(size_t), 1, file) != 1) count = 0;
    VectorDatabase *db = vector_db_init(count > 0 ? count : 10, dimension);
    if (!db) {
        fclose(file);
        return NULL;
    }
    for (size_t i = 0; i < count; ++i) {
        Vector v;
        memset(&v, 0, sizeof v);
        if (fread(v.uuid, sizeof(char), UUID_SIZE, file) != UUID_SIZE || fread(&v.dimension, sizeof(size_t), 1, file) != 1) {
            fprintf(stderr, "svdb_b200: vector_db_load: file ends after %zu of %zu rows\n", i, count);
            break;
        }
        v.data = (double *)malloc((v.dimension ? v.dimension : 1) * sizeof(double));
        if (!v.data || fread(v.data, sizeof(double), v.dimension, file) != v.dimension) {
            fprintf(stderr, "svdb_b200: vector_db_load: short row %zu\n", i);
            free(v.data);
            break;
        }
        if (db->kdtree && v.dimension < db->kdtree->dimension) {
            // a row shorter than kd_dim: the reference's tree would read past it (kdtree.c:26-28).  Keep the row's slot --
            // later rows keep the index the file gives them -- with the missing coordinates as zeros.
            fprintf(stderr, "svdb_b200: vector_db_load: row %zu has %zu < kd_dim values, zero-padded to kd_dim\n", i, v.dimension);
            double *p = (double *)calloc(db->kdtree->dimension, sizeof(double));
            if (p) {
                memcpy(p, v.data, v.dimension * sizeof(double));
                free(v.data);
                v.data = p;
                v.dimension = db->kdtree->dimension;
            }
        }
        if (vector_db_insert(db, v) == (size_t)-1) {
            free(v.data);
            vector_db_free(db);
            fclose(file);
            return NULL;
        }
    }
    fclose(file);
    return db;
}

// ---- vector_database.c:301-352 ----
float cosine_similarity(Vector vec1, Vector vec2) { return metric_call(SVDB_COSINE, vec1, vec2); }
float euclidean_distance(Vector vec1, Vector vec2) { return metric_call(SVDB_EUCLIDEAN, vec1, vec2); }
float dot_product(Vector vec1, Vector vec2) { return metric_call(SVDB_DOT, vec1, vec2); }

int vector_db_compare_batch(VectorDatabase *db, int metric, const size_t *index1, const size_t *index2, size_t n,
                            float *out) {
    if (!db || !out || (!index1 && n) || (!index2 && n)) return SVDB_ERR_ARG;
    DbImpl *d = reinterpret_cast<DbImpl *>(db);
    pthread_mutex_lock(&db->mutex);
    int rc = SVDB_OK;
    if (!d->rows) {
        for (size_t i = 0; i < n; i++) out[i] = -1.0f;
    } else {
        rc = svdb_compare_batch(d->rows, metric, index1, index2, n, out);
        if (rc == SVDB_OK) {
            for (size_t i = 0; i < n; i++)   // rows of a foreign dimension: mismatch sentinel
                if (index1[i] < db->size && index2[i] < db->size && !(d->dim_ok[index1[i]] && d->dim_ok[index2[i]]))
                    out[i] = -1.0f;
        } else {
            complain("vector_db_compare_batch");
        }
    }
    pthread_mutex_unlock(&db->mutex);
    return rc;
}

}  // extern "C"
