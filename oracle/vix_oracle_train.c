/*
 * vix_oracle_train.c -- CPU ORACLE, trainers (TEST INFRASTRUCTURE ONLY; see vix_oracle.h).
 *
 * Restates the reference's stochastic training code paths with the exact RNG consumption order:
 *   - RNGState LCG                         Utilities/RNG.swift:33-104
 *   - kmeansPlusPlusSeed                   Kernels/KMeansSeeding.swift:167-409
 *   - kmeans_minibatch_f32 (lloydMiniBatch, AoS, subsampleN = 0)
 *                                          Kernels/KMeansMiniBatchKernel.swift:401-724
 *   - Xoroshiro128 / randperm / selection sampling   Kernels/PQTrain.swift:712-795
 *   - pq_train_f32 (Lloyd + mini-batch)    Kernels/PQTrain.swift:83-388, 856-1442
 *   - pq_train_streaming_f32               Kernels/PQTrain.swift:391-706, 1444-1646
 * Pins: PQTrainTests.swift:813-816 (bit patterns) and .bench/post-phase3/ivf_search.json:54.
 */
#include "vix_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* LCG (Utilities/RNG.swift)                                                                  */
/* ------------------------------------------------------------------------------------------ */

void vo_lcg_init(vo_lcg* r, uint64_t seed, uint64_t stream) {
    uint64_t base = (seed == 0) ? 1 : seed;     /* RNG.swift:47-52 */
    r->s = base ^ (stream << 32);
}
uint64_t vo_lcg_next(vo_lcg* r) {               /* RNG.swift:61-65 */
    r->s = 2862933555777941757ULL * r->s + 3037000493ULL;
    return r->s;
}
static double lcg_next_double(vo_lcg* r) {      /* RNG.swift:88-91 */
    uint64_t u = vo_lcg_next(r) >> 11;
    return (double)u / 9007199254740992.0;
}
static int64_t lcg_next_int(vo_lcg* r, int64_t bound) { /* RNG.swift:100-103 */
    return (int64_t)(vo_lcg_next(r) % (uint64_t)bound);
}

/* ------------------------------------------------------------------------------------------ */
/* k-means++ (Kernels/KMeansSeeding.swift:167-409)                                            */
/* ------------------------------------------------------------------------------------------ */

static void km11_update(const float* data, int64_t n, int d, const float* c, float* d2) {
#pragma omp parallel for schedule(static) if (n >= 4096)
    for (int64_t i = 0; i < n; ++i) {
        float ds = vo_km11_l2sq(data + i * (int64_t)d, c, d);
        float safe = (isfinite(ds) && ds >= 0.0f) ? ds : 0.0f;       /* :357-359 */
        d2[i] = (safe < d2[i]) ? safe : d2[i];                        /* Swift min(a,b): b<a ? b : a */
    }
}

static int64_t km11_sample(const float* w, int64_t n, vo_lcg* rng) { /* :368-409 */
    double total = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double v = (double)w[i];
        if (isfinite(v) && v >= 0.0) total += v;
    }
    if (total <= 0.0) return lcg_next_int(rng, n);
    double thr = lcg_next_double(rng) * total;
    double cum = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double v = (double)w[i];
        if (isfinite(v) && v >= 0.0) cum += v;
        if (cum >= thr) return i;
    }
    return n - 1;
}

int vo_kmeanspp_seed(const float* data, int64_t n, int d, int k, uint64_t seed, uint64_t stream,
                     float* centroids_out, int64_t* chosen_out) {
    if (d < 1 || n < 1 || k < 1 || k > n) return -1;
    vo_lcg rng;
    vo_lcg_init(&rng, seed, stream);
    float* d2 = (float*)malloc(sizeof(float) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) d2[i] = INFINITY;
    int64_t first = lcg_next_int(&rng, n);                            /* :223 */
    if (chosen_out) chosen_out[0] = first;
    memcpy(centroids_out, data + first * (int64_t)d, sizeof(float) * (size_t)d);
    km11_update(data, n, d, data + first * (int64_t)d, d2);
    for (int t = 1; t < k; ++t) {
        int64_t sel = km11_sample(d2, n, &rng);
        if (chosen_out) chosen_out[t] = sel;
        memcpy(centroids_out + (size_t)t * d, data + sel * (int64_t)d, sizeof(float) * (size_t)d);
        km11_update(data, n, d, data + sel * (int64_t)d, d2);
    }
    free(d2);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Mini-batch k-means (Kernels/KMeansMiniBatchKernel.swift:401-724), lloydMiniBatch/AoS        */
/* ------------------------------------------------------------------------------------------ */

static int km12_assign_one(const float* xv, const float* c, int kc, int d) {   /* :341-359 */
    int best = 0;
    float bd = vo_km12_l2sq(xv, c, d);
    for (int ci = 1; ci < kc; ++ci) {
        float dist = vo_km12_l2sq(xv, c + (size_t)ci * d, d);
        if (dist < bd || (dist == bd && ci < best)) { bd = dist; best = ci; }
    }
    return best;
}

/* Returns 0 (success) or the negative KMeansMBStatus codes (:131-138).
 * init_centroids NULL => k-means++ with (seed, stream) (:430-447).
 * empties_per_batch (optional, capacity empties_cap) records the number of "empty" centroids
 * repaired in each processed batch (used by the E3 pin). */
int vo_kmeans_minibatch(const float* x, int64_t n, int d, int kc, const float* init_centroids,
                        int batch_size, int epochs, float tol, uint64_t seed, uint64_t stream,
                        float* centroids_out, int32_t* assign_out, int64_t* empties_per_batch,
                        int empties_cap, int* epochs_done, int64_t* batches_done) {
    if (d < 1 || d > 32768) return -1;
    if (kc < 1 || (int64_t)kc > n) return -2;
    if (!centroids_out) return -3;
    if (init_centroids) memcpy(centroids_out, init_centroids, sizeof(float) * (size_t)kc * d);
    else if (vo_kmeanspp_seed(x, n, d, kc, seed, stream, centroids_out, NULL) != 0) return -2;

    vo_lcg rng;
    vo_lcg_init(&rng, seed, stream);                                   /* :474 (fresh stream) */
    int max_touched = batch_size < kc ? batch_size : kc;
    double* sums = (double*)calloc((size_t)max_touched * d, sizeof(double));
    int* touched_list = (int*)calloc((size_t)max_touched, sizeof(int));
    int* sum_index = (int*)malloc(sizeof(int) * (size_t)kc);
    uint32_t* batch_tag = (uint32_t*)calloc((size_t)kc, sizeof(uint32_t));
    int* batch_counts = (int*)calloc((size_t)kc, sizeof(int));
    int64_t* bidx = (int64_t*)malloc(sizeof(int64_t) * (size_t)(batch_size > 0 ? batch_size : 1));
    int* bassign = (int*)malloc(sizeof(int) * (size_t)(batch_size > 0 ? batch_size : 1));
    for (int c = 0; c < kc; ++c) sum_index[c] = -1;
    uint32_t current_tag = 1;
    double prev_inertia = INFINITY;
    int edone = 0;
    int64_t nbatches = 0;
    int nep = epochs > 1 ? epochs : 1;

    for (int epoch = 0; epoch < nep; ++epoch) {
        edone = epoch + 1;
        int64_t processed = 0;
        while (processed < n) {
            int64_t remaining = n - processed;
            int bc = (int)(batch_size < remaining ? batch_size : remaining);
            for (int bi = 0; bi < bc; ++bi) bidx[bi] = (int64_t)(vo_lcg_next(&rng) % (uint64_t)n); /* :524-528 */
            current_tag += 1;
            int touched = 0;
            /* assignments use the centroids as of batch start (centroids only written below) */
#pragma omp parallel for schedule(static) if (bc >= 64)
            for (int bi = 0; bi < bc; ++bi)
                bassign[bi] = km12_assign_one(x + bidx[bi] * (int64_t)d, centroids_out, kc, d);
            for (int bi = 0; bi < bc; ++bi) {
                int cb = bassign[bi];
                if (batch_tag[cb] != current_tag) {
                    batch_tag[cb] = current_tag;
                    sum_index[cb] = touched;
                    touched_list[touched] = cb;
                    double* z = sums + (size_t)touched * d;
                    for (int j = 0; j < d; ++j) z[j] = 0.0;
                    touched += 1;
                }
                double* s = sums + (size_t)sum_index[cb] * d;
                const float* v = x + bidx[bi] * (int64_t)d;
                for (int j = 0; j < d; ++j) s[j] += (double)v[j];    /* :582 */
                batch_counts[cb] += 1;
            }
            for (int t = 0; t < touched; ++t) {                       /* :595-607 */
                int c = touched_list[t];
                int nc = batch_counts[c];
                if (nc > 0) {
                    double inv = 1.0 / (double)nc;
                    const double* s = sums + (size_t)t * d;
                    float* dst = centroids_out + (size_t)c * d;
                    for (int j = 0; j < d; ++j) dst[j] = (float)(s[j] * inv);
                    batch_counts[c] = 0;
                }
            }
            /* :609-627 + :290-331: empties = untouched this batch; counts are all zero by now, so
             * cMax = 0; farthest batch point from the (already updated) centroid 0, first max. */
            int64_t nempty = 0;
            for (int c = 0; c < kc; ++c) if (batch_tag[c] != current_tag) ++nempty;
            if (nempty > 0) {
                int far = 0;
                float fard = -INFINITY;
                for (int bi = 0; bi < bc; ++bi) {
                    float dist = vo_km12_l2sq(x + bidx[bi] * (int64_t)d, centroids_out, d);
                    if (dist > fard) { fard = dist; far = bi; }
                }
                const float* v = x + bidx[far] * (int64_t)d;
                for (int c = 0; c < kc; ++c)
                    if (batch_tag[c] != current_tag)
                        memcpy(centroids_out + (size_t)c * d, v, sizeof(float) * (size_t)d);
            }
            if (empties_per_batch && nbatches < empties_cap) empties_per_batch[nbatches] = nempty;
            processed += bc;
            nbatches += 1;
        }
        /* :635-682 inertia on a reservoir sample of min(n, 10000) (consumes n - m rng draws) */
        int64_t sm = n < 10000 ? n : 10000;
        int64_t* res = (int64_t*)malloc(sizeof(int64_t) * (size_t)sm);
        for (int64_t i = 0; i < sm; ++i) res[i] = i;
        for (int64_t i = sm; i < n; ++i) {
            int64_t j = (int64_t)(vo_lcg_next(&rng) % (uint64_t)(i + 1));
            if (j < sm) res[j] = i;
        }
        float* best = (float*)malloc(sizeof(float) * (size_t)sm);
#pragma omp parallel for schedule(static) if (sm >= 64)
        for (int64_t t = 0; t < sm; ++t) {
            const float* v = x + res[t] * (int64_t)d;
            float b = vo_km12_l2sq(v, centroids_out, d);
            for (int c = 1; c < kc; ++c) {
                float dist = vo_km12_l2sq(v, centroids_out + (size_t)c * d, d);
                if (dist < b) b = dist;
            }
            best[t] = b;
        }
        double inertia = 0.0;
        for (int64_t t = 0; t < sm; ++t) inertia += (double)best[t];
        free(best); free(res);
        if (epoch > 0) {
            double denom = prev_inertia > 4.9406564584124654e-324 ? prev_inertia : 4.9406564584124654e-324;
            double improvement = (prev_inertia - inertia) / denom;
            if (improvement < (double)tol) break;
        }
        prev_inertia = inertia;
    }
    if (assign_out) vo_assign(x, n, centroids_out, kc, d, assign_out, NULL);   /* :689-706 */
    if (epochs_done) *epochs_done = edone;
    if (batches_done) *batches_done = nbatches;
    free(sums); free(touched_list); free(sum_index); free(batch_tag); free(batch_counts);
    free(bidx); free(bassign);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Xoroshiro128** and samplers (Kernels/PQTrain.swift:712-795)                                */
/* ------------------------------------------------------------------------------------------ */

typedef struct { uint64_t s0, s1; } xoro;
static inline uint64_t rotl64(uint64_t x, unsigned k) { return (x << k) | (x >> (64 - k)); }
static uint64_t splitmix_next(uint64_t* st) {
    *st += 0x9E3779B97F4A7C15ULL;
    uint64_t z = *st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static void xoro_split(xoro* r, uint64_t seed, uint64_t stream, uint64_t task) {   /* :732-745 */
    uint64_t s = seed ^ (stream * 0xD1B54A32D192ED03ULL) ^ (task * 0x94D049BB133111EBULL);
    uint64_t a = splitmix_next(&s), b = splitmix_next(&s);
    if (a != 0 || b != 0) { r->s0 = a; r->s1 = b; }
    else { r->s0 = 0x9E3779B97F4A7C15ULL; r->s1 = 0xD1B54A32D192ED03ULL; }
}
static uint64_t xoro_u64(xoro* r) {                                                /* :747-753 */
    uint64_t res = rotl64(r->s0 * 5, 7) * 9;
    uint64_t t = r->s0 ^ r->s1;
    r->s0 = rotl64(r->s0, 24) ^ t ^ (t << 16);
    r->s1 = rotl64(t, 37);
    return res;
}
static uint32_t xoro_u32(xoro* r) { return (uint32_t)(xoro_u64(r) >> 32); }
static double xoro_f64(xoro* r) { return (double)(xoro_u64(r) >> 11) * (1.0 / 9007199254740992.0); }

static void randperm(uint32_t* a, int64_t count, xoro* r) {                        /* :761-768 */
    if (count <= 1) return;
    for (int64_t i = count - 1; i >= 1; --i) {
        uint32_t rr = xoro_u32(r);
        int64_t j = (int64_t)(((uint64_t)rr * (uint64_t)(i + 1)) >> 32);
        uint32_t t = a[i]; a[i] = a[j]; a[j] = t;
    }
}
/* :770-782 selection sampling; returns count written (ascending indices) */
static int64_t sample_wo_repl(uint32_t n, uint32_t k, xoro* r, uint32_t* out) {
    uint32_t t = 0, m = 0;
    while (m < k && t < n) {
        double u = xoro_f64(r);
        if ((double)(n - t) * u >= (double)(k - m)) { t += 1; }
        else { out[m] = t; t += 1; m += 1; }
    }
    return m;
}

/* PQTrain.swift:797-813 / :833-852: l2Sq and its residual form ((x - g) - c). */
static inline float pq_l2(const float* a, const float* b, int len) { return vo_pqtrain_l2sq(a, b, len); }
static inline float pq_l2_res(const float* x, const float* c, int dsub, const float* g) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int l8 = dsub & ~7;
    for (int i = 0; i < l8; i += 8)
        for (int l = 0; l < 8; ++l) {
            float r = (x[i + l] - g[i + l]) - c[i + l];
            acc[l] = acc[l] + r * r;
        }
    float v[4];
    for (int l = 0; l < 4; ++l) v[l] = acc[l] + acc[4 + l];
    float s = ((v[0] + v[1]) + v[2]) + v[3];
    for (int i = l8; i < dsub; ++i) { float r = (x[i] - g[i]) - c[i]; s = s + r * r; }
    return s;
}
static inline float pq_dist(const float* xs, const float* c, int dsub, const float* gs) {
    return gs ? pq_l2_res(xs, c, dsub, gs) : pq_l2(xs, c, dsub);
}

void vo_pq_train_cfg_default(vo_pq_train_cfg* c) {                                  /* :20-43 */
    memset(c, 0, sizeof(*c));
    c->ks = 256; c->m = 1; c->algorithm = 0; c->max_iters = 25; c->batch_size = 1024;
    c->empty_policy = 0; c->sample_n = 0; c->seed = 42; c->stream_id = 0; c->tol = 1e-4f;
    c->precompute_x_norm2 = 0; c->compute_centroid_norms = 1;
}

/* stable "descending by mins" order (Swift sorted { mins[$0] > mins[$1] }; Swift 5's sort is a
 * stable merge sort in practice, ties keep ascending index) */
typedef struct { float v; int64_t i; } ord_t;
static int ord_cmp(const void* a, const void* b) {
    const ord_t* x = (const ord_t*)a;
    const ord_t* y = (const ord_t*)b;
    if (x->v > y->v) return -1;
    if (x->v < y->v) return 1;
    return (x->i < y->i) ? -1 : (x->i > y->i);
}

/* :970-1019 kmeansppSeedSubspaceDense */
static void seed_dense(const float* xd, int64_t n, int dsub, int ks, xoro* rng, float* C) {
    int64_t i0 = (int64_t)(xoro_f64(rng) * (double)n);
    if (i0 < 0) i0 = 0;
    if (i0 > n - 1) i0 = n - 1;
    memcpy(C, xd + i0 * dsub, sizeof(float) * (size_t)dsub);
    float* dmin = (float*)malloc(sizeof(float) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) dmin[i] = pq_l2(xd + i * dsub, C, dsub);
    for (int k = 1; k < ks; ++k) {
        double sum = 0.0;
        for (int64_t i = 0; i < n; ++i) sum += (double)dmin[i];
        int64_t pick;
        if (!(sum > 0)) {
            pick = (int64_t)(xoro_f64(rng) * (double)n);
            if (pick < 0) pick = 0;
            if (pick > n - 1) pick = n - 1;
        } else {
            double r = xoro_f64(rng) * sum;
            pick = n - 1;
            for (int64_t i = 0; i < n; ++i) {
                r -= (double)dmin[i];
                if (r <= 0) { pick = i; break; }
            }
        }
        memcpy(C + (size_t)k * dsub, xd + pick * dsub, sizeof(float) * (size_t)dsub);
        for (int64_t i = 0; i < n; ++i) {
            float di = pq_l2(xd + i * dsub, C + (size_t)k * dsub, dsub);
            if (di < dmin[i]) dmin[i] = di;
        }
    }
    free(dmin);
}

/* :856-968 kmeansppSeedSubspace (strided x, optional residual) */
static void seed_strided(const float* x, int64_t n, int d, int j, int dsub, int ks,
                         const float* coarse, const int32_t* assign, xoro* rng, float* C) {
    int64_t i0 = (int64_t)(xoro_f64(rng) * (double)n);
    if (i0 < 0) i0 = 0;
    if (i0 >= n) i0 = n - 1;
    const float* x0 = x + i0 * (int64_t)d + (size_t)j * dsub;
    if (coarse) {
        const float* g0 = coarse + (int64_t)assign[i0] * d + (size_t)j * dsub;
        for (int u = 0; u < dsub; ++u) C[u] = x0[u] - g0[u];
    } else memcpy(C, x0, sizeof(float) * (size_t)dsub);
    float* dmin = (float*)malloc(sizeof(float) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const float* xs = x + i * (int64_t)d + (size_t)j * dsub;
        const float* gs = coarse ? coarse + (int64_t)assign[i] * d + (size_t)j * dsub : NULL;
        dmin[i] = pq_dist(xs, C, dsub, gs);
    }
    for (int k = 1; k < ks; ++k) {
        double sum = 0.0;
        for (int64_t i = 0; i < n; ++i) sum += (double)dmin[i];
        int64_t pick;
        if (!(sum > 0)) {
            pick = (int64_t)(xoro_f64(rng) * (double)n);
            if (pick < 0) pick = 0;
            if (pick > n - 1) pick = n - 1;
        } else {
            double r = xoro_f64(rng) * sum;
            pick = n - 1;
            for (int64_t i = 0; i < n; ++i) {
                r -= (double)dmin[i];
                if (r <= 0) { pick = i; break; }
            }
        }
        const float* xp = x + pick * (int64_t)d + (size_t)j * dsub;
        float* ck = C + (size_t)k * dsub;
        if (coarse) {
            const float* gp = coarse + (int64_t)assign[pick] * d + (size_t)j * dsub;
            for (int u = 0; u < dsub; ++u) ck[u] = xp[u] - gp[u];
        } else memcpy(ck, xp, sizeof(float) * (size_t)dsub);
        for (int64_t i = 0; i < n; ++i) {
            const float* xs = x + i * (int64_t)d + (size_t)j * dsub;
            const float* gs = coarse ? coarse + (int64_t)assign[i] * d + (size_t)j * dsub : NULL;
            float di = pq_dist(xs, ck, dsub, gs);
            if (di < dmin[i]) dmin[i] = di;
        }
    }
    free(dmin);
}

/* :1023-1202 lloydKMeansSubspace */
static void lloyd_subspace(const float* x, int64_t n, int d, int j, int dsub, int ks,
                           const float* coarse, const int32_t* assign, const vo_pq_train_cfg* cfg,
                           float* C, double* out_dist, int* out_iters) {
    double prev = INFINITY;
    int it = 0;
    int use_dot = cfg->precompute_x_norm2 && coarse == NULL;
    float* qn = NULL;
    if (use_dot) {
        qn = (float*)malloc(sizeof(float) * (size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            const float* xs = x + i * (int64_t)d + (size_t)j * dsub;
            float s = 0.0f;
            for (int u = 0; u < dsub; ++u) s += xs[u] * xs[u];
            qn[i] = s;
        }
    }
    double* sums = (double*)malloc(sizeof(double) * (size_t)ks * dsub);
    int64_t* counts = (int64_t*)malloc(sizeof(int64_t) * (size_t)ks);
    float* cn = (float*)calloc((size_t)ks, sizeof(float));
    int32_t* best_k = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    float* best_d = (float*)malloc(sizeof(float) * (size_t)n);
    int max_iters = cfg->max_iters > 1 ? cfg->max_iters : 1;
    for (int iter = 0; iter < max_iters; ++iter) {
        if (use_dot)
            for (int k = 0; k < ks; ++k) {
                float s = 0.0f;
                for (int u = 0; u < dsub; ++u) { float v = C[(size_t)k * dsub + u]; s += v * v; }
                cn[k] = s;
            }
        memset(sums, 0, sizeof(double) * (size_t)ks * dsub);
        memset(counts, 0, sizeof(int64_t) * (size_t)ks);
        /* assignment (parallel over rows; per-row arithmetic unchanged) */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            const float* xs = x + i * (int64_t)d + (size_t)j * dsub;
            const float* gs = coarse ? coarse + (int64_t)assign[i] * d + (size_t)j * dsub : NULL;
            int bk = 0;
            float bd;
            if (use_dot) {
                float dot = 0.0f;
                for (int u = 0; u < dsub; ++u) dot += xs[u] * C[u];
                bd = (qn[i] + cn[0]) - 2.0f * dot;
                for (int k = 1; k < ks; ++k) {
                    const float* c = C + (size_t)k * dsub;
                    float dt = 0.0f;
                    for (int u = 0; u < dsub; ++u) dt += xs[u] * c[u];
                    float dk = (qn[i] + cn[k]) - 2.0f * dt;
                    if (dk < bd || (dk == bd && k < bk)) { bd = dk; bk = k; }
                }
            } else {
                bd = pq_dist(xs, C, dsub, gs);
                for (int k = 1; k < ks; ++k) {
                    float dk = pq_dist(xs, C + (size_t)k * dsub, dsub, gs);
                    if (dk < bd || (dk == bd && k < bk)) { bd = dk; bk = k; }
                }
            }
            best_k[i] = bk;
            best_d[i] = bd;
        }
        /* f64 accumulation strictly in row order (:1111, :1131) */
        double distortion = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            const float* xs = x + i * (int64_t)d + (size_t)j * dsub;
            double* s = sums + (size_t)best_k[i] * dsub;
            if (coarse) {
                const float* gs = coarse + (int64_t)assign[i] * d + (size_t)j * dsub;
                for (int u = 0; u < dsub; ++u) s[u] += (double)(xs[u] - gs[u]);
            } else {
                for (int u = 0; u < dsub; ++u) s[u] += (double)xs[u];
            }
            float bd = best_d[i];
            if (bd < 0) bd = 0;
            counts[best_k[i]] += 1;
            distortion += (double)bd;
        }
        int empties = 0;
        for (int k = 0; k < ks; ++k) {
            if (counts[k] > 0) {
                double inv = 1.0 / (double)counts[k];
                for (int u = 0; u < dsub; ++u)
                    C[(size_t)k * dsub + u] = (float)(sums[(size_t)k * dsub + u] * inv);
            } else empties += 1;
        }
        if (empties > 0) {
            if (cfg->empty_policy == 1) {                                        /* .reseed :1126-1136 */
                uint64_t seed = cfg->seed ^ ((uint64_t)j * 0x9E3779B97F4A7C15ULL) ^
                                ((uint64_t)iter * 0xD1B54A32D192ED03ULL);
                for (int k = 0; k < ks; ++k)
                    if (counts[k] == 0) {
                        seed = 2862933555777941757ULL * seed + 3037000493ULL;
                        int64_t pick = (int64_t)(seed % (uint64_t)n);
                        memcpy(C + (size_t)k * dsub, x + pick * (int64_t)d + (size_t)j * dsub,
                               sizeof(float) * (size_t)dsub);
                    }
            } else if (cfg->empty_policy == 0) {                                 /* .split :1137-1188 */
                int64_t sample = n < (128 > n / 4 ? 128 : n / 4) ? n : (128 > n / 4 ? 128 : n / 4);
                int64_t stride = n / sample > 1 ? n / sample : 1;
                int64_t cnt = 0;
                for (int64_t idx = 0; idx < n; idx += stride) ++cnt;
                ord_t* o = (ord_t*)malloc(sizeof(ord_t) * (size_t)cnt);
                int64_t t = 0;
                for (int64_t idx = 0; idx < n; idx += stride, ++t) {
                    const float* xs = x + idx * (int64_t)d + (size_t)j * dsub;
                    float md = pq_l2(xs, C, dsub);       /* NB: repair ignores the residual (:1155) */
                    for (int kk = 1; kk < ks; ++kk) {
                        float di = pq_l2(xs, C + (size_t)kk * dsub, dsub);
                        if (di < md) md = di;
                    }
                    o[t].v = md; o[t].i = t;
                }
                qsort(o, (size_t)cnt, sizeof(ord_t), ord_cmp);
                int r = 0;
                for (int k = 0; k < ks && r < cnt; ++k)
                    if (counts[k] == 0) {
                        int64_t pick = o[r].i * stride;
                        memcpy(C + (size_t)k * dsub, x + pick * (int64_t)d + (size_t)j * dsub,
                               sizeof(float) * (size_t)dsub);
                        ++r;
                    }
                free(o);
            }
        }
        double improve = (prev - distortion) / (prev == 0 ? 1 : prev);
        prev = distortion;
        it = iter + 1;
        if (cfg->tol > 0 && iter > 0 && improve >= 0 && improve < (double)cfg->tol) break;
    }
    double denom = (double)(n > 1 ? n : 1);
    if (!isfinite(prev) || prev < 0) *out_dist = (prev > 0 ? prev : 0) / denom;
    else *out_dist = prev / denom;
    *out_iters = it;
    free(qn); free(sums); free(counts); free(cn); free(best_k); free(best_d);
}

/* running-mean blend shared by :1297-1316 and :1552-1569 */
static void blend_update(float* C, const double* sums, const int64_t* counts, int64_t* global_counts,
                         int64_t* pass_counts_after, int ks, int dsub) {
    for (int k = 0; k < ks; ++k) {
        int64_t ck = counts[k];
        if (ck > 0) {
            int64_t old_n = global_counts[k];
            int64_t new_n = old_n + ck;
            global_counts[k] = new_n;
            double old_w = (double)old_n / (double)new_n;
            double new_w = (double)ck / (double)new_n;
            for (int u = 0; u < dsub; ++u) {
                double old_val = (double)C[(size_t)k * dsub + u];
                double batch_mean = sums[(size_t)k * dsub + u] / (double)ck;
                float v = (float)(old_w * old_val + new_w * batch_mean);
                C[(size_t)k * dsub + u] = isfinite(v) ? v : 0.0f;
            }
            if (pass_counts_after) pass_counts_after[k] += ck;
        }
    }
}

static int assign_sub(const float* xs, const float* gs, const float* C, int ks, int dsub) {
    int bk = 0;
    float bd = pq_dist(xs, C, dsub, gs);
    for (int k = 1; k < ks; ++k) {
        float dk = pq_dist(xs, C + (size_t)k * dsub, dsub, gs);
        if (dk < bd || (dk == bd && k < bk)) { bd = dk; bk = k; }
    }
    return bk;
}

/* :1206-1442 minibatchKMeansSubspace (no warm start) */
static void minibatch_subspace(const float* x, int64_t n, int d, int j, int dsub, int ks,
                               const float* coarse, const int32_t* assign,
                               const vo_pq_train_cfg* cfg, int64_t sample_n_eff, int dist_eval_n,
                               xoro* rng, float* C, double* out_dist, int* out_iters) {
    if (n == 0) return;
    uint32_t* idx = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) idx[i] = (uint32_t)i;
    int B = cfg->batch_size > 1 ? cfg->batch_size : 1;
    int iters = 0;
    int64_t* gcounts = (int64_t*)calloc((size_t)ks, sizeof(int64_t));
    double* sums = (double*)malloc(sizeof(double) * (size_t)ks * dsub);
    int64_t* counts = (int64_t*)malloc(sizeof(int64_t) * (size_t)ks);
    int passes = cfg->max_iters > 1 ? cfg->max_iters : 1;
    for (int p = 0; p < passes; ++p) {
        randperm(idx, n, rng);
        int64_t processed = 0;
        int64_t limit = (sample_n_eff > 0) ? (n < sample_n_eff ? n : sample_n_eff) : n;
        while (processed < limit) {
            int64_t s = processed;
            int64_t e = (s + B < limit) ? s + B : limit;
            memset(sums, 0, sizeof(double) * (size_t)ks * dsub);
            memset(counts, 0, sizeof(int64_t) * (size_t)ks);
            for (int64_t t = s; t < e; ++t) {
                int64_t i = idx[t];
                const float* xs = x + i * (int64_t)d + (size_t)j * dsub;
                const float* gs = coarse ? coarse + (int64_t)assign[i] * d + (size_t)j * dsub : NULL;
                int bk = assign_sub(xs, gs, C, ks, dsub);
                double* sk = sums + (size_t)bk * dsub;
                if (gs) for (int u = 0; u < dsub; ++u) sk[u] += (double)(xs[u] - gs[u]);
                else    for (int u = 0; u < dsub; ++u) sk[u] += (double)xs[u];
                counts[bk] += 1;
            }
            blend_update(C, sums, counts, gcounts, NULL, ks, dsub);
            iters += 1;
            processed = e;
        }
        /* pass-level repair for clusters that never received anything (:1325-1395) */
        int nempty = 0;
        for (int k = 0; k < ks; ++k) if (gcounts[k] == 0) ++nempty;
        if (nempty > 0) {
            int64_t eval_lim = sample_n_eff > 0 ? sample_n_eff : dist_eval_n;
            if (eval_lim > n) eval_lim = n;
            if (eval_lim > 0) {
                ord_t* o = (ord_t*)malloc(sizeof(ord_t) * (size_t)eval_lim);
                int64_t* inds = (int64_t*)malloc(sizeof(int64_t) * (size_t)eval_lim);
                for (int64_t t = 0; t < eval_lim; ++t) {
                    int64_t i = idx[t % n];
                    inds[t] = i;
                    const float* xs = x + i * (int64_t)d + (size_t)j * dsub;
                    const float* gs = coarse ? coarse + (int64_t)assign[i] * d + (size_t)j * dsub : NULL;
                    float md = pq_dist(xs, C, dsub, gs);
                    for (int kk = 1; kk < ks; ++kk) {
                        float di = pq_dist(xs, C + (size_t)kk * dsub, dsub, gs);
                        if (di < md) md = di;
                    }
                    o[t].v = md; o[t].i = t;
                }
                qsort(o, (size_t)eval_lim, sizeof(ord_t), ord_cmp);
                int64_t rank = 0;
                for (int k = 0; k < ks && rank < eval_lim; ++k)
                    if (gcounts[k] == 0) {
                        int64_t i = inds[o[rank].i];
                        const float* xs = x + i * (int64_t)d + (size_t)j * dsub;
                        if (coarse) {
                            const float* gs = coarse + (int64_t)assign[i] * d + (size_t)j * dsub;
                            for (int u = 0; u < dsub; ++u) C[(size_t)k * dsub + u] = xs[u] - gs[u];
                        } else memcpy(C + (size_t)k * dsub, xs, sizeof(float) * (size_t)dsub);
                        gcounts[k] = 1;
                        ++rank;
                    }
                free(o); free(inds);
            }
        }
    }
    for (int t = 0; t < ks * dsub; ++t) if (!isfinite(C[t])) C[t] = 0.0f;
    double total = 0.0;
    int64_t used = 0;
    int64_t eval_lim = sample_n_eff > 0 ? sample_n_eff : dist_eval_n;
    if (eval_lim > n) eval_lim = n;
    for (int64_t t = 0; t < eval_lim; ++t) {
        int64_t i = idx[t % n];
        const float* xs = x + i * (int64_t)d + (size_t)j * dsub;
        float best = pq_l2(xs, C, dsub);
        for (int k = 1; k < ks; ++k) {
            float dk = pq_l2(xs, C + (size_t)k * dsub, dsub);
            if (dk < best) best = dk;
        }
        if (best < 0) best = 0;
        if (isfinite(best)) { total += (double)best; used += 1; }
    }
    *out_dist = (used > 0 && isfinite(total)) ? total / (double)used : 1.0;
    *out_iters = iters;
    free(idx); free(gcounts); free(sums); free(counts);
}

/* Kernels/PQTrain.swift:83-388 (pq_train_f32).  Returns 0 or a negative error:
 * -1 invalidDimension, -2 invalidParameter, -3 contractViolation, -4 emptyInput (:96-135). */
int vo_pq_train(const float* x, int64_t n, int d, const vo_pq_train_cfg* in_cfg,
                const float* coarse, const int32_t* assign,
                float* codebooks_out, float* norms_out, double* distortion_out) {
    vo_pq_train_cfg cfg = *in_cfg;
    int m = cfg.m, ks = cfg.ks;
    if (d <= 0 || m <= 0 || n < 0) return -1;
    if (d % m != 0) return -1;
    if (ks < 1 || ks > 65536) return -2;
    if ((coarse == NULL) != (assign == NULL)) return -3;
    int64_t need = cfg.sample_n > 0 ? cfg.sample_n : n;
    if (need < ks) return -4;
    int dsub = d / m;
    const int dist_eval_n = 2000;
    if (cfg.algorithm == 1) {                                                     /* :144-149 */
        if (cfg.sample_n <= 0 && n > dist_eval_n) cfg.sample_n = dist_eval_n;
        if (cfg.batch_size <= 0) cfg.batch_size = 512;
        cfg.empty_policy = 1;
    }
    if (cfg.max_iters <= 0) cfg.max_iters = 25;
    if (cfg.tol <= 0) cfg.tol = 1e-4f;
    double total_dist = 0.0;
    for (int j = 0; j < m; ++j) {
        xoro rng;
        xoro_split(&rng, cfg.seed, cfg.stream_id, (uint64_t)j);
        /* buildSampleIndex :784-795 */
        int64_t ns;
        uint32_t* idx = NULL;
        if (cfg.sample_n <= 0 || cfg.sample_n >= n) { ns = n; }
        else {
            idx = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)cfg.sample_n);
            ns = sample_wo_repl((uint32_t)n, (uint32_t)cfg.sample_n, &rng, idx);
            ns = cfg.sample_n;
        }
        float* Cj = codebooks_out + (size_t)j * ks * dsub;
        memset(Cj, 0, sizeof(float) * (size_t)ks * dsub);
        int64_t seeding_cap = 4LL * ks;
        int use_subset = ns > seeding_cap;
        int64_t ns_seed = use_subset ? seeding_cap : ns;
        if (ns == n && !use_subset) {
            seed_strided(x, n, d, j, dsub, ks, coarse, assign, &rng, Cj);
        } else {
            float* tmp = (float*)malloc(sizeof(float) * (size_t)ns_seed * dsub);
            uint32_t* pos = NULL;
            if (use_subset) {
                pos = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)ns_seed);
                sample_wo_repl((uint32_t)ns, (uint32_t)ns_seed, &rng, pos);
            }
            for (int64_t t = 0; t < ns_seed; ++t) {
                int64_t pool = use_subset ? pos[t] : t;
                int64_t i = (ns == n) ? pool : idx[pool];
                const float* xs = x + i * (int64_t)d + (size_t)j * dsub;
                if (coarse) {
                    const float* gs = coarse + (int64_t)assign[i] * d + (size_t)j * dsub;
                    for (int u = 0; u < dsub; ++u) tmp[t * dsub + u] = xs[u] - gs[u];
                } else memcpy(tmp + t * dsub, xs, sizeof(float) * (size_t)dsub);
            }
            seed_dense(tmp, ns_seed, dsub, ks, &rng, Cj);
            free(tmp); free(pos);
        }
        double dist = 0.0;
        int iters = 0;
        if (cfg.algorithm == 1)
            minibatch_subspace(x, n, d, j, dsub, ks, coarse, assign, &cfg, cfg.sample_n, dist_eval_n,
                               &rng, Cj, &dist, &iters);
        else
            lloyd_subspace(x, n, d, j, dsub, ks, coarse, assign, &cfg, Cj, &dist, &iters);
        if (norms_out)                                                            /* :299-307 */
            for (int k = 0; k < ks; ++k) {
                float s = 0.0f;
                for (int u = 0; u < dsub; ++u) { float v = Cj[(size_t)k * dsub + u]; s += v * v; }
                norms_out[(size_t)j * ks + k] = s;
            }
        total_dist += dist;
        free(idx);
    }
    if (distortion_out) *distortion_out = total_dist;
    return 0;
}

/* :1577-1646 streamingKMeansppSeed (no residual) */
static void streaming_seed(const float* const* chunks, const int64_t* cn, int nchunks, int d, int j,
                           int dsub, int ks, xoro* rng, float* C) {
    int64_t total = 0;
    for (int c = 0; c < nchunks; ++c) total += cn[c];
    int64_t pick = (int64_t)(xoro_f64(rng) * (double)total);
    if (pick < 0) pick = 0;
    if (pick >= total) pick = total - 1;
    int ci = 0;
    int64_t off = pick;
    while (off >= cn[ci]) { off -= cn[ci]; ci += 1; }
    memcpy(C, chunks[ci] + off * (int64_t)d + (size_t)j * dsub, sizeof(float) * (size_t)dsub);
    float** dmin = (float**)malloc(sizeof(float*) * (size_t)nchunks);
    for (int c = 0; c < nchunks; ++c) {
        dmin[c] = (float*)malloc(sizeof(float) * (size_t)(cn[c] > 0 ? cn[c] : 1));
        for (int64_t i = 0; i < cn[c]; ++i)
            dmin[c][i] = pq_l2(chunks[c] + i * (int64_t)d + (size_t)j * dsub, C, dsub);
    }
    for (int k = 1; k < ks; ++k) {
        double sum = 0.0;
        for (int c = 0; c < nchunks; ++c)
            for (int64_t i = 0; i < cn[c]; ++i) sum += (double)dmin[c][i];
        float* ck = C + (size_t)k * dsub;
        if (!(sum > 0)) {
            memcpy(ck, chunks[0] + (size_t)j * dsub, sizeof(float) * (size_t)dsub);
        } else {
            double r = xoro_f64(rng) * sum;
            int chosen = 0;
            for (int c = 0; c < nchunks && !chosen; ++c)
                for (int64_t i = 0; i < cn[c]; ++i) {
                    r -= (double)dmin[c][i];
                    if (r <= 0) {
                        memcpy(ck, chunks[c] + i * (int64_t)d + (size_t)j * dsub,
                               sizeof(float) * (size_t)dsub);
                        chosen = 1;
                        break;
                    }
                }
            if (!chosen) memcpy(ck, C, sizeof(float) * (size_t)dsub);
        }
        for (int c = 0; c < nchunks; ++c)
            for (int64_t i = 0; i < cn[c]; ++i) {
                float di = pq_l2(chunks[c] + i * (int64_t)d + (size_t)j * dsub, ck, dsub);
                if (di < dmin[c][i]) dmin[c][i] = di;
            }
    }
    for (int c = 0; c < nchunks; ++c) free(dmin[c]);
    free(dmin);
}

/* Kernels/PQTrain.swift:391-706 (pq_train_streaming_f32), no residual.  The pinned golden vector
 * (PQTrainTests.swift:724-817) exercises: streamingKMeansppSeed (totalN <= 4*ks), ten passes of
 * minibatchKMeansSubspaceChunk with sampleProb = 1, and no empty repair. */
int vo_pq_train_streaming(const float* const* chunks, const int64_t* chunk_n, int nchunks, int d,
                          const vo_pq_train_cfg* in_cfg, float* codebooks_out, float* norms_out) {
    vo_pq_train_cfg cfg = *in_cfg;
    int m = cfg.m, ks = cfg.ks;
    if (d <= 0 || m <= 0 || d % m != 0) return -1;
    if (ks < 1 || ks > 65536) return -2;
    cfg.algorithm = 1;
    if (cfg.max_iters <= 0) cfg.max_iters = 15;
    if (cfg.batch_size <= 0) cfg.batch_size = 8192;
    int64_t total = 0;
    for (int c = 0; c < nchunks; ++c) total += chunk_n[c];
    if (cfg.sample_n <= 0 && total > 2000) cfg.sample_n = 2000;
    int dsub = d / m;
    const int streaming_repair_eval_n = 512;
    for (int j = 0; j < m; ++j) {
        xoro rng;
        xoro_split(&rng, cfg.seed, cfg.stream_id, (uint64_t)j);
        float* Cj = codebooks_out + (size_t)j * ks * dsub;
        memset(Cj, 0, sizeof(float) * (size_t)ks * dsub);
        int64_t cap = 4LL * ks;
        if (total > cap) {
            int64_t sn = total < cap ? total : cap;
            uint32_t* picks = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)sn);
            int64_t got = sample_wo_repl((uint32_t)total, (uint32_t)sn, &rng, picks);
            float* tmp = (float*)calloc((size_t)sn * dsub, sizeof(float));
            for (int64_t t = 0; t < got; ++t) {
                int64_t g = picks[t];
                int c = 0;
                int64_t pre = 0;
                while (c < nchunks - 1 && g >= pre + chunk_n[c]) { pre += chunk_n[c]; c += 1; }
                int64_t i = g - pre;
                memcpy(tmp + t * dsub, chunks[c] + i * (int64_t)d + (size_t)j * dsub,
                       sizeof(float) * (size_t)dsub);
            }
            seed_dense(tmp, sn, dsub, ks, &rng, Cj);
            free(tmp); free(picks);
        } else {
            streaming_seed(chunks, chunk_n, nchunks, d, j, dsub, ks, &rng, Cj);
        }
        int64_t* gcounts = (int64_t*)calloc((size_t)ks, sizeof(int64_t));
        int64_t* pcounts = (int64_t*)calloc((size_t)ks, sizeof(int64_t));
        double* sums = (double*)malloc(sizeof(double) * (size_t)ks * dsub);
        int64_t* counts = (int64_t*)malloc(sizeof(int64_t) * (size_t)ks);
        int B = cfg.batch_size > 1 ? cfg.batch_size : 1;
        for (int pass = 0; pass < cfg.max_iters; ++pass) {
            memset(pcounts, 0, sizeof(int64_t) * (size_t)ks);
            int64_t limit = cfg.sample_n > 0 ? (total < cfg.sample_n ? total : cfg.sample_n) : total;
            double prob = (double)limit / (double)(total > 1 ? total : 1);
            if (prob < 0.0) prob = 0.0;
            if (prob > 1.0) prob = 1.0;
            for (int c = 0; c < nchunks; ++c) {
                int64_t nc = chunk_n[c];
                if (nc <= 0) continue;
                /* :1444-1575 minibatchKMeansSubspaceChunk */
                uint32_t* idx = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)nc);
                for (int64_t i = 0; i < nc; ++i) idx[i] = (uint32_t)i;
                randperm(idx, nc, &rng);
                for (int64_t s = 0; s < nc;) {
                    int64_t e = (s + B < nc) ? s + B : nc;
                    memset(sums, 0, sizeof(double) * (size_t)ks * dsub);
                    memset(counts, 0, sizeof(int64_t) * (size_t)ks);
                    for (int64_t t = s; t < e; ++t) {
                        if (prob < 1.0) {
                            double u = xoro_f64(&rng);
                            if (u > prob) continue;
                        }
                        int64_t i = idx[t];
                        const float* xs = chunks[c] + i * (int64_t)d + (size_t)j * dsub;
                        int bk = assign_sub(xs, NULL, Cj, ks, dsub);
                        double* sk = sums + (size_t)bk * dsub;
                        for (int u = 0; u < dsub; ++u) sk[u] += (double)xs[u];
                        counts[bk] += 1;
                    }
                    blend_update(Cj, sums, counts, gcounts, pcounts, ks, dsub);
                    s = e;
                }
                free(idx);
            }
            /* pass-level repair (:543-640), no-residual branch */
            int nempty = 0;
            for (int k = 0; k < ks; ++k) if (gcounts[k] == 0) ++nempty;
            if (nempty > 0) {
                int64_t eval_n = total < streaming_repair_eval_n ? total : streaming_repair_eval_n;
                if (eval_n > 0) {
                    ord_t* o = (ord_t*)malloc(sizeof(ord_t) * (size_t)eval_n);
                    int* pc = (int*)malloc(sizeof(int) * (size_t)eval_n);
                    int64_t* pi = (int64_t*)malloc(sizeof(int64_t) * (size_t)eval_n);
                    for (int64_t t = 0; t < eval_n; ++t) {
                        int64_t g = (int64_t)(xoro_f64(&rng) * (double)total);
                        int c = 0;
                        int64_t pre = 0;
                        while (c < nchunks - 1 && g >= pre + chunk_n[c]) { pre += chunk_n[c]; c += 1; }
                        int64_t i = g - pre;
                        pc[t] = c; pi[t] = i;
                        const float* xs = chunks[c] + i * (int64_t)d + (size_t)j * dsub;
                        float md = pq_l2(xs, Cj, dsub);
                        for (int kk = 1; kk < ks; ++kk) {
                            float di = pq_l2(xs, Cj + (size_t)kk * dsub, dsub);
                            if (di < md) md = di;
                        }
                        o[t].v = md; o[t].i = t;
                    }
                    qsort(o, (size_t)eval_n, sizeof(ord_t), ord_cmp);
                    int64_t rank = 0;
                    for (int k = 0; k < ks && rank < eval_n; ++k)
                        if (gcounts[k] == 0) {
                            int64_t t = o[rank].i;
                            memcpy(Cj + (size_t)k * dsub,
                                   chunks[pc[t]] + pi[t] * (int64_t)d + (size_t)j * dsub,
                                   sizeof(float) * (size_t)dsub);
                            gcounts[k] = 1;
                            ++rank;
                        }
                    free(o); free(pc); free(pi);
                }
            }
        }
        if (norms_out)
            for (int k = 0; k < ks; ++k) {
                float s = 0.0f;
                for (int u = 0; u < dsub; ++u) { float v = Cj[(size_t)k * dsub + u]; s += v * v; }
                norms_out[(size_t)j * ks + k] = s;
            }
        free(gcounts); free(pcounts); free(sums); free(counts);
    }
    return 0;
}
