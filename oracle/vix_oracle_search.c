/*
 * vix_oracle_search.c -- CPU ORACLE, search path (TEST INFRASTRUCTURE ONLY; see vix_oracle.h).
 *
 * Restates, operation for operation, the reference's scoring / selection / coarse-quantiser /
 * PQ-encode / LUT / ADC arithmetic.  "SIMD4 accumulators" of the Swift code are written out as
 * explicit per-lane scalar accumulators; every multiply and add is a separate fp32 operation
 * (compile with -ffp-contract=off).  File:line citations are relative to
 * /root/reference/Sources/VectorIndex unless another root is given.
 */
#include "vix_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Host threads of the OpenMP loops below (over independent rows / queries only).  n > 0 sets the count (the
 * launcher of a multi-process job exports OMP_NUM_THREADS=1, which would otherwise silently serialise the
 * CPU baseline); returns the count the next parallel region will use. */
int vo_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

static inline float hsum4(const float* v) { return ((v[0] + v[1]) + v[2]) + v[3]; }

/* ------------------------------------------------------------------------------------------ */
/* Pair kernels                                                                               */
/* ------------------------------------------------------------------------------------------ */

/* Operations/Scoring/L2SqrKernel.swift:192-238 (_l2sqr_single_direct, kahan=false):
 * four SIMD4 accumulators over 16-element strides, lane-wise ((a0+a1)+a2)+a3, hsum, scalar tail. */
float vo_l2sqr_direct(const float* q, const float* x, int d) {
    float acc[16];
    for (int l = 0; l < 16; ++l) acc[l] = 0.0f;
    int j = 0;
    while (j + 15 < d) {
        for (int l = 0; l < 16; ++l) {
            float df = q[j + l] - x[j + l];
            acc[l] = acc[l] + df * df;
        }
        j += 16;
    }
    float v[4];
    for (int l = 0; l < 4; ++l) v[l] = ((acc[l] + acc[4 + l]) + acc[8 + l]) + acc[12 + l];
    float sum = hsum4(v);
    for (int t = j; t < d; ++t) {
        float df = q[t] - x[t];
        sum = sum + df * df;
    }
    return sum;
}

/* Operations/Support/Norms.swift:105-130 (l2NormSquared): 16-stride four accumulators ->
 * hsum4(a0+a1+a2+a3) (lane-wise left-to-right), then 4-groups sum += hsum4(v*v), then scalar tail. */
float vo_norm_l2sq(const float* x, int d) {
    if (d == 0) return 0.0f;
    float acc[16];
    for (int l = 0; l < 16; ++l) acc[l] = 0.0f;
    int d16 = d & ~15;
    int j = 0;
    while (j < d16) {
        for (int l = 0; l < 16; ++l) acc[l] = acc[l] + x[j + l] * x[j + l];
        j += 16;
    }
    float v[4];
    for (int l = 0; l < 4; ++l) v[l] = ((acc[l] + acc[4 + l]) + acc[8 + l]) + acc[12 + l];
    float sum = hsum4(v);
    int d4 = d & ~3;
    while (j < d4) {
        float p[4];
        for (int l = 0; l < 4; ++l) p[l] = x[j + l] * x[j + l];
        sum = sum + hsum4(p);
        j += 4;
    }
    while (j < d) { sum = sum + x[j] * x[j]; ++j; }
    return sum;
}

/* Operations/Scoring/L2SqrKernel.swift:411-448 (_l2sqr_block_dot_fused_serial), one row:
 * 16-stride four-accumulator dot over dBlocked=(d/4)*4, scalar tail from j, then
 * dist = qNorm + xn - 2*dot, clamped at 0.  x_norm NaN => computed by l2NormSquared. */
float vo_l2sqr_dot_fused(const float* q, const float* row, int d, float q_norm, float x_norm) {
    float acc[16];
    for (int l = 0; l < 16; ++l) acc[l] = 0.0f;
    int d_blocked = (d / 4) * 4;
    int j = 0;
    while (j + 15 < d_blocked) {
        for (int l = 0; l < 16; ++l) acc[l] = acc[l] + q[j + l] * row[j + l];
        j += 16;
    }
    float v[4];
    for (int l = 0; l < 4; ++l) v[l] = ((acc[l] + acc[4 + l]) + acc[8 + l]) + acc[12 + l];
    float dot = hsum4(v);
    for (int t = j; t < d; ++t) dot = dot + q[t] * row[t];
    float xn = isnan(x_norm) ? vo_norm_l2sq(row, d) : x_norm;
    float dist = (q_norm + xn) - 2.0f * dot;
    if (dist < 0.0f) dist = 0.0f;
    return dist;
}

/* Operations/Scoring/InnerProduct.swift:115-150 (generic) and :153-184 (ip_r1_D; r4/r8 use the
 * same per-row order): ONE SIMD4 accumulator fed by every 4-group in order, hsum, <=3 scalar tail. */
float vo_ip(const float* q, const float* x, int d) {
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    int d4 = d & ~3;
    for (int j = 0; j < d4; j += 4)
        for (int l = 0; l < 4; ++l) acc[l] = acc[l] + q[j + l] * x[j + l];
    float sum = hsum4(acc);
    for (int j = d4; j < d; ++j) sum = sum + q[j] * x[j];
    return sum;
}

/* Kernels/KMeansMiniBatchKernel.swift:198-225 (_vi_km12_l2sq_aos): two SIMD4 accumulators over
 * 8-strides, result hsum(acc0) + hsum(acc1), scalar tail.  BIT-EXACT contract (IVF assignment). */
float vo_km12_l2sq(const float* a, const float* b, int d) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int d8 = (d / 8) * 8;
    for (int j = 0; j < d8; j += 8)
        for (int l = 0; l < 8; ++l) {
            float df = a[j + l] - b[j + l];
            acc[l] = acc[l] + df * df;
        }
    float sum = hsum4(acc) + hsum4(acc + 4);
    for (int j = d8; j < d; ++j) {
        float df = a[j] - b[j];
        sum = sum + df * df;
    }
    return sum;
}

/* Kernels/KMeansSeeding.swift:302-361 (_vi_km11_updateSquaredDistances inner distance): same lanes
 * as km12 but the two accumulators are added lane-wise first: (acc0+acc1).sum(), then tail. */
float vo_km11_l2sq(const float* a, const float* b, int d) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int d8 = (d / 8) * 8;
    for (int j = 0; j < d8; j += 8)
        for (int l = 0; l < 8; ++l) {
            float df = a[j + l] - b[j + l];
            acc[l] = acc[l] + df * df;
        }
    float v[4];
    for (int l = 0; l < 4; ++l) v[l] = acc[l] + acc[4 + l];
    float sum = hsum4(v);
    for (int j = d8; j < d; ++j) {
        float df = a[j] - b[j];
        sum = sum + df * df;
    }
    return sum;
}

/* Kernels/PQTrain.swift:797-813 (l2Sq): identical lane structure to km11 ((acc0+acc1) then hsum). */
float vo_pqtrain_l2sq(const float* a, const float* b, int d) { return vo_km11_l2sq(a, b, d); }

/* Operations/Quantization/PQLUT.swift:69-103 (_simd_l2sqr): two SIMD4 accumulators per 8-stride;
 * sum = hsum(acc0); sum += hsum(acc1); scalar tail. */
float vo_lut_l2sqr(const float* a, const float* b, int len) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int len8 = len & ~7;
    for (int i = 0; i < len8; i += 8)
        for (int l = 0; l < 8; ++l) {
            float df = a[i + l] - b[i + l];
            acc[l] = acc[l] + df * df;
        }
    float sum = hsum4(acc);
    sum = sum + hsum4(acc + 4);
    for (int i = len8; i < len; ++i) {
        float df = a[i] - b[i];
        sum = sum + df * df;
    }
    return sum;
}

/* Operations/Quantization/PQLUT.swift:105-140 (_simd_dot). */
float vo_lut_dot(const float* a, const float* b, int len) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int len8 = len & ~7;
    for (int i = 0; i < len8; i += 8)
        for (int l = 0; l < 8; ++l) acc[l] = acc[l] + a[i + l] * b[i + l];
    float sum = hsum4(acc);
    sum = sum + hsum4(acc + 4);
    for (int i = len8; i < len; ++i) sum = sum + a[i] * b[i];
    return sum;
}

/* Operations/Quantization/PQEncode.swift:637-648 (sqnorm): per 4-group acc += sum4(v*v); tail. */
float vo_pq_sqnorm(const float* a, int d) {
    float acc = 0.0f;
    int dv = d & ~3;
    int j = 0;
    while (j < dv) {
        float p[4];
        for (int l = 0; l < 4; ++l) p[l] = a[j + l] * a[j + l];
        acc = acc + hsum4(p);
        j += 4;
    }
    while (j < d) { acc = acc + a[j] * a[j]; ++j; }
    return acc;
}

/* ------------------------------------------------------------------------------------------ */
/* Block scoring                                                                              */
/* ------------------------------------------------------------------------------------------ */

/* Operations/Scoring/L2SqrKernel.swift:78-138 (l2sqr_f32_block, algo .auto, default opts):
 * dot-trick iff norms are available (xb_norm != NULL or q_norm not NaN) or d >= 256 (:95-106);
 * otherwise the direct kernel.  Row partitioning over threads does not change per-row results. */
void vo_l2sqr_block(const float* q, const float* xb, int64_t n, int d, float* out,
                    const float* xb_norm, float q_norm) {
    if (n <= 0 || d <= 0) return;
    int can_dot = (xb_norm != NULL) || !isnan(q_norm);
    int use_dot = can_dot || (d >= 256);
    if (use_dot) {
        float qn = isnan(q_norm) ? vo_norm_l2sq(q, d) : q_norm;
#pragma omp parallel for schedule(static) if (n >= 4096)
        for (int64_t i = 0; i < n; ++i)
            out[i] = vo_l2sqr_dot_fused(q, xb + i * (int64_t)d, d, qn, xb_norm ? xb_norm[i] : NAN);
    } else {
#pragma omp parallel for schedule(static) if (n >= 4096)
        for (int64_t i = 0; i < n; ++i) out[i] = vo_l2sqr_direct(q, xb + i * (int64_t)d, d);
    }
}

/* Operations/Scoring/InnerProduct.swift:8-101 (run). */
void vo_ip_block(const float* q, const float* xb, int64_t n, int d, float* out) {
    if (n <= 0) return;
    if (d == 0) { for (int64_t i = 0; i < n; ++i) out[i] = 0.0f; return; }
#pragma omp parallel for schedule(static) if (n >= 4096)
    for (int64_t i = 0; i < n; ++i) out[i] = vo_ip(q, xb + i * (int64_t)d, d);
}

/* Operations/Scoring/ScoreBlock.swift:24-70: euclidean => L2^2 (no sqrt), dotProduct => raw dot. */
void vo_score_block(const float* q, const float* xb, int64_t n, int d, int metric, float* out) {
    if (metric == VO_METRIC_L2) vo_l2sqr_block(q, xb, n, d, out, NULL, NAN);
    else vo_ip_block(q, xb, n, d, out);
}

/* ------------------------------------------------------------------------------------------ */
/* Selection (Operations/Selection/TopK.swift, TopKMerge.swift)                               */
/* ------------------------------------------------------------------------------------------ */

/* TopK.swift:8-31 */
static inline int is_better(int ord, float as, int32_t ai, float bs, int32_t bi) {
    if (ord == VO_ORDER_MIN) return (as < bs) || (as == bs && ai < bi);
    return (as > bs) || (as == bs && ai < bi);
}
static inline int is_worse(int ord, float as, int32_t ai, float bs, int32_t bi) {
    if (ord == VO_ORDER_MIN) return (as > bs) || (as == bs && ai > bi);
    return (as < bs) || (as == bs && ai > bi);
}

typedef struct { float* s; int32_t* id; int count; int cap; int ord; } vo_heap;

/* TopK.swift:95-103 (_siftDown): root holds the WORST retained element. */
static void heap_sift_down(vo_heap* h, int idx) {
    int i = idx;
    float s = h->s[i];
    int32_t id = h->id[i];
    for (;;) {
        int left = (i << 1) + 1;
        if (left >= h->count) break;
        int right = left + 1;
        int w = left;
        float ws = h->s[left];
        int32_t wid = h->id[left];
        if (right < h->count) {
            float rs = h->s[right];
            int32_t rid = h->id[right];
            if (is_worse(h->ord, rs, rid, ws, wid)) { w = right; ws = rs; wid = rid; }
        }
        if (is_worse(h->ord, ws, wid, s, id)) { h->s[i] = ws; h->id[i] = wid; i = w; }
        else break;
    }
    h->s[i] = s;
    h->id[i] = id;
}

typedef struct { float s; int32_t id; int ord; } vo_pair;
static int pair_cmp_min(const void* a, const void* b) {
    const vo_pair* x = (const vo_pair*)a;
    const vo_pair* y = (const vo_pair*)b;
    if (is_better(x->ord, x->s, x->id, y->s, y->id)) return -1;
    if (is_better(x->ord, y->s, y->id, x->s, x->id)) return 1;
    return 0;
}

/* TopK.swift:127-164 (selectTopK, streaming algorithm; the hybrid quickselect returns the same set
 * because (score, id) is a total order) + :77-82 (extractSorted best->worst).
 * ids == NULL => ids are 0..n-1.  Returns kEff = min(k, n) written entries. */
int vo_select_topk(const float* scores, const int32_t* ids, int64_t n, int k, int ordering,
                   float* out_scores, int32_t* out_ids) {
    int64_t keff64 = k < n ? k : n;
    int keff = keff64 < 0 ? 0 : (int)keff64;
    if (keff <= 0) return 0;
    vo_heap h;
    h.s = (float*)malloc(sizeof(float) * (size_t)keff);
    h.id = (int32_t*)malloc(sizeof(int32_t) * (size_t)keff);
    h.cap = keff;
    h.ord = ordering;
    h.count = keff;
    for (int i = 0; i < keff; ++i) { h.s[i] = scores[i]; h.id[i] = ids ? ids[i] : (int32_t)i; }
    for (int i = (h.count / 2) - 1; i >= 0; --i) heap_sift_down(&h, i);   /* heapify :83 */
    for (int64_t i = keff; i < n; ++i) {
        int32_t id = ids ? ids[i] : (int32_t)i;
        float s = scores[i];
        float rs = h.s[0];
        int32_t rid = h.id[0];
        int repl = (ordering == VO_ORDER_MIN) ? ((s < rs) || (s == rs && id < rid))
                                              : ((s > rs) || (s == rs && id < rid));
        if (repl) { h.s[0] = s; h.id[0] = id; heap_sift_down(&h, 0); }
    }
    vo_pair* p = (vo_pair*)malloc(sizeof(vo_pair) * (size_t)keff);
    for (int i = 0; i < keff; ++i) { p[i].s = h.s[i]; p[i].id = h.id[i]; p[i].ord = ordering; }
    qsort(p, (size_t)keff, sizeof(vo_pair), pair_cmp_min);
    for (int i = 0; i < keff; ++i) { out_scores[i] = p[i].s; out_ids[i] = p[i].id; }
    free(p); free(h.s); free(h.id);
    return keff;
}

/* TopKMerge.swift:11-61 + :66-71: k-way merge of best->worst lists; order (score, id), exact
 * duplicates from different lists broken by the smaller list index.  Lists are stored as rows of
 * a [nlists x list_stride] matrix with lens[l] valid entries.  Returns number written. */
int vo_merge_topk(const float* scores, const int32_t* ids, const int32_t* lens, int nlists,
                  int list_stride, int k, int ordering, float* out_scores, int32_t* out_ids) {
    if (k <= 0 || nlists <= 0) return 0;
    int* head = (int*)calloc((size_t)nlists, sizeof(int));
    int outc = 0;
    while (outc < k) {
        int best = -1;
        for (int l = 0; l < nlists; ++l) {
            if (head[l] >= lens[l]) continue;
            if (best < 0) { best = l; continue; }
            float as = scores[(size_t)l * list_stride + head[l]];
            int32_t ai = ids[(size_t)l * list_stride + head[l]];
            float bs = scores[(size_t)best * list_stride + head[best]];
            int32_t bi = ids[(size_t)best * list_stride + head[best]];
            /* strictly better wins; on exact duplicates the smaller list index (= earlier l) stays */
            if (is_better(ordering, as, ai, bs, bi)) best = l;
        }
        if (best < 0) break;
        out_scores[outc] = scores[(size_t)best * list_stride + head[best]];
        out_ids[outc] = ids[(size_t)best * list_stride + head[best]];
        ++outc;
        ++head[best];
    }
    free(head);
    return outc;
}

/* ------------------------------------------------------------------------------------------ */
/* Coarse quantiser                                                                           */
/* ------------------------------------------------------------------------------------------ */

/* IVFIndex.swift:470-485 (rebuildCentroidCache): ||c||^2 by Norms.l2NormSquared. */
void vo_centroid_norms(const float* c, int kc, int d, float* out) {
    for (int i = 0; i < kc; ++i) out[i] = vo_norm_l2sq(c + (size_t)i * d, d);
}

/* Kernels/CentroidBatchScore.swift:39-88: out = alpha * Q * C^T (cblas_sgemm, alpha=-2 L2 / -1 IP)
 * then L2 adds ||c||^2 per column (||q||^2 omitted).  Accelerate's summation order is unknown and
 * closed-source; this oracle uses the netlib reference order (sequential over the contraction
 * index).  alpha is a power of two, so alpha*(sequential sum) == netlib's sum of (alpha*b)*a bit
 * for bit.  Parity against Accelerate itself is by tolerance only. */
void vo_centroid_batch_score(const float* queries, int64_t q, const float* centroids, int kc, int d,
                             int metric, const float* centroid_norms, float* out) {
    float alpha = (metric == VO_METRIC_L2) ? -2.0f : -1.0f;
#pragma omp parallel for schedule(static) if (q * (int64_t)kc >= 4096)
    for (int64_t qi = 0; qi < q; ++qi) {
        const float* qp = queries + qi * (int64_t)d;
        float* row = out + qi * (int64_t)kc;
        for (int ci = 0; ci < kc; ++ci) {
            const float* cp = centroids + (size_t)ci * d;
            float acc = 0.0f;
            for (int t = 0; t < d; ++t) acc = acc + qp[t] * cp[t];
            float v = alpha * acc;
            if (metric == VO_METRIC_L2) v = v + centroid_norms[ci];
            row[ci] = v;
        }
        if (metric == VO_METRIC_COSINE) {
            /* CentroidBatchScore.swift:70-84 (queriesAreNormalized == false): the row holds -<q, c>;
             * 1 - dot qInv cInv = 1 + row qInv cInv, except where the near-zero-norm guard of the single-query path
             * (IVFIndex.swift:558-561) forces the largest distance, 1.  centroid_norms = ||c||^2 by
             * Norms.l2NormSquared, cInv = 1 / (sqrt(||c||^2) + 1e-12) (rebuildCentroidCache, IVFIndex.swift:470-485). */
            const float qn = vo_norm_l2sq(qp, d);
            const float qinv = 1.0f / (sqrtf(qn) + 1e-12f);
            for (int ci = 0; ci < kc; ++ci) {
                const float cinv = 1.0f / (sqrtf(centroid_norms[ci]) + 1e-12f);
                const float denom = sqrtf(qn * centroid_norms[ci]);
                row[ci] = denom > 1.1920929e-07f ? (1.0f + (row[ci] * qinv) * cinv) : 1.0f;
            }
        }
    }
}

typedef struct { float s; int32_t i; } vo_probe;
static int probe_cmp(const void* a, const void* b) {
    const vo_probe* x = (const vo_probe*)a;
    const vo_probe* y = (const vo_probe*)b;
    /* IVFIndex.swift:593-595 (probeIsOrderedBefore): score ascending, then index ascending */
    if (x->s != y->s) return (x->s < y->s) ? -1 : 1;
    return (x->i < y->i) ? -1 : (x->i > y->i);
}

/* IVFIndex.swift:920-927: sort all (ci, score) by probeIsOrderedBefore, prefix(nprobeEff);
 * nprobeEff = min(nprobe, kc) (:897); entries beyond nprobeEff are padded with -1 / NaN in the
 * style of Kernels/IVFSelect.swift:399-413. */
void vo_probe_select(const float* scores, int kc, int nprobe, int32_t* out_idx, float* out_scores) {
    vo_probe* p = (vo_probe*)malloc(sizeof(vo_probe) * (size_t)kc);
    for (int i = 0; i < kc; ++i) { p[i].s = scores[i]; p[i].i = i; }
    qsort(p, (size_t)kc, sizeof(vo_probe), probe_cmp);
    int eff = nprobe < kc ? nprobe : kc;
    for (int i = 0; i < eff; ++i) { out_idx[i] = p[i].i; if (out_scores) out_scores[i] = p[i].s; }
    for (int i = eff; i < nprobe; ++i) { out_idx[i] = -1; if (out_scores) out_scores[i] = NAN; }
    free(p);
}

/* IVFIndex.swift:865-931 (batchSearch probe stage): CentroidBatchScore + per-row ordered prefix. */
void vo_probe_select_batch(const float* queries, int64_t q, const float* centroids, int kc, int d,
                           int metric, const float* centroid_norms, int nprobe,
                           int32_t* out_idx, float* out_scores) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t qi = 0; qi < q; ++qi) {
        float* row = (float*)malloc(sizeof(float) * (size_t)kc);
        vo_centroid_batch_score(queries + qi * (int64_t)d, 1, centroids, kc, d, metric,
                                centroid_norms, row);
        vo_probe_select(row, kc, nprobe, out_idx + qi * (int64_t)nprobe,
                        out_scores ? out_scores + qi * (int64_t)nprobe : NULL);
        free(row);
    }
}

/* Kernels/KMeansMiniBatchKernel.swift:341-359 (_vi_km12_assignAOS) applied to every row
 * (:689-706): argmin over c of km12 L2^2, tie -> lower c.  BIT-EXACT contract. */
void vo_assign(const float* x, int64_t n, const float* c, int kc, int d, int32_t* assign_out,
               float* dist_out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float* xv = x + i * (int64_t)d;
        int best = 0;
        float bd = vo_km12_l2sq(xv, c, d);
        for (int ci = 1; ci < kc; ++ci) {
            float dist = vo_km12_l2sq(xv, c + (size_t)ci * d, d);
            if (dist < bd || (dist == bd && ci < best)) { bd = dist; best = ci; }
        }
        assign_out[i] = best;
        if (dist_out) dist_out[i] = bd;
    }
}

/* IVFIndex.swift:376-435 (dot/cosine list build): tiles of rows -> CentroidBatchScore ->
 * first-min argmin (strict <, ascending ci).  For L2 this is the CentroidBatchScore variant. */
void vo_assign_metric(const float* x, int64_t n, const float* c, int kc, int d, int metric,
                      const float* centroid_norms, int32_t* assign_out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float* row = (float*)malloc(sizeof(float) * (size_t)kc);
        vo_centroid_batch_score(x + i * (int64_t)d, 1, c, kc, d, metric, centroid_norms, row);
        int best = -1;
        float bs = INFINITY;
        for (int ci = 0; ci < kc; ++ci)
            if (row[ci] < bs) { bs = row[ci]; best = ci; }
        assign_out[i] = best;
        free(row);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* PQ encode: restatement of /root/reference/Sources/CPQEncode/pq_encode.c (x86 scalar path).   */
/* Validated bit-for-bit against the compiled reference in oracle/_ref by tests.               */
/* ------------------------------------------------------------------------------------------ */

/* pq_encode.c:74-80 (pq_argmin_update) */
static inline void argmin_update(float dist, int k, float* bd, int* bk) {
    if (dist < *bd || (dist == *bd && k < *bk)) { *bd = dist; *bk = k; }
}
/* pq_encode.c:188-195 (dot_only, scalar) */
static inline float enc_dot(const float* x, const float* c, int dsub) {
    float dot = 0.0f;
    for (int i = 0; i < dsub; ++i) dot += x[i] * c[i];
    return dot;
}
/* pq_encode.c:83-90 (l2_sq_scalar) */
static inline float enc_l2(const float* a, const float* b, int dsub) {
    float acc = 0.0f;
    for (int i = 0; i < dsub; ++i) { float df = a[i] - b[i]; acc += df * df; }
    return acc;
}
/* pq_encode.c:126-134 (dist_dp_scalar): interleaved dot / c2 accumulation, x2 + c2 - 2*dot */
static inline float enc_dist_dp(const float* x, const float* c, int dsub, float x2) {
    float dot = 0.0f, c2 = 0.0f;
    for (int i = 0; i < dsub; ++i) { float ci = c[i]; dot += x[i] * ci; c2 += ci * ci; }
    return x2 + c2 - 2.0f * dot;
}
/* pq_encode.c:199-207 (l2_sq_residual_scalar) */
static inline float enc_l2_res(const float* x, const float* g, const float* c, int dsub) {
    float acc = 0.0f;
    for (int i = 0; i < dsub; ++i) { float r = (x[i] - g[i]) - c[i]; acc += r * r; }
    return acc;
}
/* pq_encode.c:246-257 (dist_dp_residual_scalar) */
static inline float enc_dist_dp_res(const float* x, const float* g, const float* c, int dsub,
                                    float r2) {
    float dot = 0.0f, c2 = 0.0f;
    for (int i = 0; i < dsub; ++i) {
        float ri = x[i] - g[i];
        float ci = c[i];
        dot += ri * ci;
        c2 += ci * ci;
    }
    return r2 + c2 - 2.0f * dot;
}

/* One (vector, subspace) code.  mode selection mirrors the public entry points:
 *   csq != NULL, g == NULL : encode_subspace_u8_dot_with_csq        (pq_encode.c:332-366)
 *   csq != NULL, g != NULL : encode_subspace_u8_residual_with_csq   (:368-410)
 *   csq == NULL, g == NULL : use_dot ? encode_subspace_u8_dot (:296-330) : _direct (:279-294)
 *   csq == NULL, g != NULL : encode_subspace_u8_residual(use_dot)   (:412-447)
 * The k-tiling of the reference only affects prefetch hints, not the visiting order 0..ks-1. */
static int encode_one(const float* xs, const float* gs, const float* cb, const float* csq, int ks,
                      int dsub, int use_dot) {
    int bk = 0;
    float bd;
    if (csq && !gs) {
        float x2 = 0.0f;
        for (int i = 0; i < dsub; ++i) x2 += xs[i] * xs[i];
        bd = x2 + csq[0] - 2.0f * enc_dot(xs, cb, dsub);
        for (int k = 1; k < ks; ++k) {
            float dd = x2 + csq[k] - 2.0f * enc_dot(xs, cb + (size_t)k * dsub, dsub);
            argmin_update(dd, k, &bd, &bk);
        }
    } else if (csq && gs) {
        float r2 = 0.0f;
        for (int i = 0; i < dsub; ++i) { float ri = xs[i] - gs[i]; r2 += ri * ri; }
        float dot0 = enc_dot(xs, cb, dsub) - enc_dot(gs, cb, dsub);
        bd = r2 + csq[0] - 2.0f * dot0;
        for (int k = 1; k < ks; ++k) {
            const float* ck = cb + (size_t)k * dsub;
            float dotk = enc_dot(xs, ck, dsub) - enc_dot(gs, ck, dsub);
            float dd = r2 + csq[k] - 2.0f * dotk;
            argmin_update(dd, k, &bd, &bk);
        }
    } else if (!gs) {
        if (use_dot) {
            float x2 = 0.0f;
            for (int i = 0; i < dsub; ++i) x2 += xs[i] * xs[i];
            bd = enc_dist_dp(xs, cb, dsub, x2);
            for (int k = 1; k < ks; ++k)
                argmin_update(enc_dist_dp(xs, cb + (size_t)k * dsub, dsub, x2), k, &bd, &bk);
        } else {
            bd = enc_l2(xs, cb, dsub);
            for (int k = 1; k < ks; ++k)
                argmin_update(enc_l2(xs, cb + (size_t)k * dsub, dsub), k, &bd, &bk);
        }
    } else {
        if (use_dot) {
            float r2 = 0.0f;
            for (int i = 0; i < dsub; ++i) { float ri = xs[i] - gs[i]; r2 += ri * ri; }
            bd = enc_dist_dp_res(xs, gs, cb, dsub, r2);
            for (int k = 1; k < ks; ++k)
                argmin_update(enc_dist_dp_res(xs, gs, cb + (size_t)k * dsub, dsub, r2), k, &bd, &bk);
        } else {
            bd = enc_l2_res(xs, gs, cb, dsub);
            for (int k = 1; k < ks; ++k)
                argmin_update(enc_l2_res(xs, gs, cb + (size_t)k * dsub, dsub), k, &bd, &bk);
        }
    }
    return bk;
}

/* pq_encode.c:479-739: cpq_encode_u8_f32 / _with_csq / residual variants, AoS codes[i*m+j].
 * centroid_sq NULL => the no-csq entry points (use_dot selects dot-trick vs direct);
 * coarse/assign NULL => non-residual. */
void vo_pq_encode_u8(const float* x, int64_t n, int d, int m, int ks, const float* codebooks,
                     const float* centroid_sq, const float* coarse, const int32_t* assign,
                     int use_dot, uint8_t* codes) {
    int dsub = d / m;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
        const float* xi = x + i * (int64_t)d;
        const float* gc = coarse ? coarse + (int64_t)assign[i] * d : NULL;
        for (int j = 0; j < m; ++j) {
            const float* cb = codebooks + ((size_t)j * ks) * dsub;
            const float* csq = centroid_sq ? centroid_sq + (size_t)j * ks : NULL;
            codes[i * (int64_t)m + j] = (uint8_t)encode_one(
                xi + (size_t)j * dsub, gc ? gc + (size_t)j * dsub : NULL, cb, csq, ks, dsub, use_dot);
        }
    }
}

/* pq_encode.c:558-599, 692-739: u4 (ks=16) always uses the direct L2 (residual: fused direct),
 * two codes packed per byte: low nibble = even subspace. */
void vo_pq_encode_u4(const float* x, int64_t n, int d, int m, int ks, const float* codebooks,
                     const float* coarse, const int32_t* assign, uint8_t* codes) {
    int dsub = d / m;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
        const float* xi = x + i * (int64_t)d;
        const float* gc = coarse ? coarse + (int64_t)assign[i] * d : NULL;
        uint8_t* out = codes + i * (int64_t)(m >> 1);
        for (int j = 0; j < m; j += 2) {
            int c0 = encode_one(xi + (size_t)j * dsub, gc ? gc + (size_t)j * dsub : NULL,
                                codebooks + ((size_t)j * ks) * dsub, NULL, ks, dsub, 0);
            int c1 = encode_one(xi + (size_t)(j + 1) * dsub, gc ? gc + (size_t)(j + 1) * dsub : NULL,
                                codebooks + ((size_t)(j + 1) * ks) * dsub, NULL, ks, dsub, 0);
            out[j >> 1] = (uint8_t)((c0 & 0x0F) | ((c1 & 0x0F) << 4));
        }
    }
}

/* Operations/Quantization/PQEncode.swift:543-565 (ensureCentroidSqNorms): Swift sqnorm per centroid. */
void vo_pq_centroid_sq_swift(const float* codebooks, int m, int ks, int dsub, float* out) {
    for (int j = 0; j < m; ++j)
        for (int k = 0; k < ks; ++k)
            out[(size_t)j * ks + k] = vo_pq_sqnorm(codebooks + ((size_t)j * ks + k) * dsub, dsub);
}
/* Tests/VectorIndexTests/PQEncodeParity_AoS_C_vs_Swift_Tests.swift:19-31 (computeCentroidSq) and
 * Kernels/PQTrain.swift:299-307: strictly sequential s += v*v. */
void vo_pq_centroid_sq_seq(const float* codebooks, int m, int ks, int dsub, float* out) {
    for (int j = 0; j < m; ++j)
        for (int k = 0; k < ks; ++k) {
            const float* c = codebooks + ((size_t)j * ks + k) * dsub;
            float s = 0.0f;
            for (int i = 0; i < dsub; ++i) s += c[i] * c[i];
            out[(size_t)j * ks + k] = s;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* PQ LUT (Operations/Quantization/PQLUT.swift) -- parity unpinned by reference tests          */
/* ------------------------------------------------------------------------------------------ */

static inline float scalar_dot(const float* a, const float* b, int len) {
    float s = 0.0f;
    for (int i = 0; i < len; ++i) s += a[i] * b[i];
    return s;
}
static inline float scalar_l2(const float* a, const float* b, int len) {
    float s = 0.0f;
    for (int i = 0; i < len; ++i) { float df = a[i] - b[i]; s += df * df; }
    return s;
}

/* PQLUT.swift:191-261 (pq_lut_l2_f32).  use_dot: -1 auto (norms given && ks >= 64), 0, 1. */
void vo_pq_lut_l2(const float* q, int d, int m, int ks, const float* codebooks, float* lut,
                  const float* centroid_norms, const float* q_sub_norms, int use_dot,
                  int include_q, int strict_fp) {
    int dsub = d / m;
    int dot = (use_dot < 0) ? ((centroid_norms != NULL) && (ks >= 64)) : use_dot;
    for (int j = 0; j < m; ++j) {
        const float* qj = q + (size_t)j * dsub;
        const float* cbj = codebooks + ((size_t)j * ks) * dsub;
        float* lutj = lut + (size_t)j * ks;
        float qn = 0.0f;
        if (include_q)
            qn = q_sub_norms ? q_sub_norms[j]
                             : (strict_fp ? scalar_dot(qj, qj, dsub) : vo_lut_dot(qj, qj, dsub));
        if (dot) {
            const float* cn = centroid_norms + (size_t)j * ks;
            for (int k = 0; k < ks; ++k) {
                const float* c = cbj + (size_t)k * dsub;
                float dp = strict_fp ? scalar_dot(qj, c, dsub) : vo_lut_dot(qj, c, dsub);
                lutj[k] = ((include_q ? qn : 0.0f) + cn[k]) - 2.0f * dp;
            }
        } else {
            for (int k = 0; k < ks; ++k) {
                const float* c = cbj + (size_t)k * dsub;
                lutj[k] = strict_fp ? scalar_l2(qj, c, dsub) : vo_lut_l2sqr(qj, c, dsub);
            }
        }
    }
}

/* PQLUT.swift:266-386 (pq_lut_residual_l2_f32). */
void vo_pq_lut_residual_l2(const float* q, const float* coarse, int d, int m, int ks,
                           const float* codebooks, float* lut, const float* centroid_norms,
                           int use_dot, int include_q, int strict_fp) {
    int dsub = d / m;
    int dot = (use_dot < 0) ? ((centroid_norms != NULL) && (ks >= 64)) : use_dot;
    int len8 = dsub & ~7;
    for (int j = 0; j < m; ++j) {
        const float* qj = q + (size_t)j * dsub;
        const float* cj = coarse + (size_t)j * dsub;
        const float* cbj = codebooks + ((size_t)j * ks) * dsub;
        float* lutj = lut + (size_t)j * ks;
        if (dot) {
            const float* cn = centroid_norms + (size_t)j * ks;
            float rn = 0.0f;
            if (include_q) rn = strict_fp ? scalar_l2(qj, cj, dsub) : vo_lut_l2sqr(qj, cj, dsub);
            for (int k = 0; k < ks; ++k) {
                const float* c = cbj + (size_t)k * dsub;
                float dp;
                if (strict_fp) {
                    dp = 0.0f;
                    for (int i = 0; i < dsub; ++i) dp += (qj[i] - cj[i]) * c[i];
                } else {
                    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    for (int i = 0; i < len8; i += 8)
                        for (int l = 0; l < 8; ++l) {
                            float rc = qj[i + l] - cj[i + l];
                            acc[l] = acc[l] + rc * c[i + l];
                        }
                    dp = hsum4(acc);
                    dp = dp + hsum4(acc + 4);
                    for (int i = len8; i < dsub; ++i) dp = dp + (qj[i] - cj[i]) * c[i];
                }
                lutj[k] = (rn + cn[k]) - 2.0f * dp;
            }
        } else {
            for (int k = 0; k < ks; ++k) {
                const float* c = cbj + (size_t)k * dsub;
                float s;
                if (strict_fp) {
                    s = 0.0f;
                    for (int i = 0; i < dsub; ++i) {
                        float df = (qj[i] - cj[i]) - c[i];
                        s += df * df;
                    }
                } else {
                    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    for (int i = 0; i < len8; i += 8)
                        for (int l = 0; l < 8; ++l) {
                            float r = (qj[i + l] - cj[i + l]) - c[i + l];
                            acc[l] = acc[l] + r * r;
                        }
                    s = hsum4(acc);
                    s = s + hsum4(acc + 4);
                    for (int i = len8; i < dsub; ++i) {
                        float df = (qj[i] - cj[i]) - c[i];
                        s = s + df * df;
                    }
                }
                lutj[k] = s;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* ADC scan (Operations/Quantization/ADCScan.swift) -- parity unpinned by reference tests      */
/* ------------------------------------------------------------------------------------------ */

/* ADCScan.swift:190-283 (scanU8AoS): default 4 accumulators s[j mod 4] over 4-groups, leftover
 * (m mod 4) entries all go to s0, out = (((s0+s1)+s2)+s3) + bias; strictFP && m >= 64 => Kahan. */
void vo_adc_scan_u8(const uint8_t* codes, int64_t n, int m, int ks, const float* lut, float* out,
                    int stride, float bias, int strict_fp) {
    int st = stride > 0 ? stride : m;
    int kahan = strict_fp && m >= 64;
#pragma omp parallel for schedule(static) if (n >= 4096)
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t* row = codes + i * (int64_t)st;
        if (kahan) {
            float sum = 0.0f, c = 0.0f;
            for (int j = 0; j < m; ++j) {
                float value = lut[(size_t)j * ks + row[j]];
                float y = value - c;
                float t = sum + y;
                c = (t - sum) - y;
                sum = t;
            }
            out[i] = sum + bias;
        } else {
            float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            int j = 0;
            while (j + 3 < m) {
                s0 += lut[(size_t)(j + 0) * ks + row[j + 0]];
                s1 += lut[(size_t)(j + 1) * ks + row[j + 1]];
                s2 += lut[(size_t)(j + 2) * ks + row[j + 2]];
                s3 += lut[(size_t)(j + 3) * ks + row[j + 3]];
                j += 4;
            }
            while (j < m) { s0 += lut[(size_t)j * ks + row[j]]; ++j; }
            out[i] = (((s0 + s1) + s2) + s3) + bias;
        }
    }
}

/* ADCScan.swift:384-456 (scanU4AoS): packed nibbles, low nibble = even subspace; strictFP &&
 * m >= 64 => Kahan (:419-437), otherwise a single sequential accumulator (:438-447). */
void vo_adc_scan_u4(const uint8_t* codes, int64_t n, int m, int ks, const float* lut, float* out,
                    int stride, float bias, int strict_fp) {
    int mb = m / 2;
    int st = stride > 0 ? stride : mb;
    int kahan = strict_fp && m >= 64;
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t* row = codes + i * (int64_t)st;
        if (kahan) {
            float sum = 0.0f, c = 0.0f;
            for (int b = 0; b < mb; ++b) {
                uint8_t byte = row[b];
                float vals[2] = { lut[(size_t)(2 * b) * ks + (byte & 0x0F)],
                                  lut[(size_t)(2 * b + 1) * ks + (byte >> 4)] };
                for (int t = 0; t < 2; ++t) {
                    float y = vals[t] - c;
                    float tt = sum + y;
                    c = (tt - sum) - y;
                    sum = tt;
                }
            }
            out[i] = sum + bias;
        } else {
            /* :438-447: ONE sequential accumulator over j = 0..m-1 (unlike the u8 path) */
            float sum = 0.0f;
            for (int b = 0; b < mb; ++b) {
                uint8_t byte = row[b];
                sum += lut[(size_t)(2 * b) * ks + (byte & 0x0F)];
                sum += lut[(size_t)(2 * b + 1) * ks + ((byte >> 4) & 0x0F)];
            }
            out[i] = sum + bias;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Composed searches                                                                          */
/* ------------------------------------------------------------------------------------------ */

/* FlatIndexOptimized.swift:390-477 (fastSearchWithMicrokernels): ScoreBlock.run -> selectTopK with
 * ids 0..n-1 (.min for L2, .max for IP) -> extractSorted -> API distance (L2: sqrt, IP: -dot).
 * out_raw (optional) receives the raw kernel scores of the winners.  Unused slots: id -1, NaN. */
/* computeQueryInvNorm_impl / the on-the-fly row norm of Cosine.run (Cosine.swift:113-114, 186-190), epsilon = 1e-12 */
float vo_cosine_inv_norm(const float* x, int d) {
    if (d == 0) return 1.0f / 1e-12f;
    return 1.0f / (sqrtf(vo_norm_l2sq(x, d)) + 1e-12f);
}

void vo_flat_search(const float* queries, int64_t nq, const float* xb, int64_t n, int d,
                    int metric, int k, float* out_dist, int64_t* out_ids, float* out_raw) {
    if (k <= 0) return;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t qi = 0; qi < nq; ++qi) {
        float* scores = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        float* ts = (float*)malloc(sizeof(float) * (size_t)k);
        int32_t* ti = (int32_t*)malloc(sizeof(int32_t) * (size_t)k);
        const float* q = queries + qi * (int64_t)d;
        if (metric == VO_METRIC_L2) {
            if (d >= 256) {
                float qn = vo_norm_l2sq(q, d);
                for (int64_t i = 0; i < n; ++i)
                    scores[i] = vo_l2sqr_dot_fused(q, xb + i * (int64_t)d, d, qn, NAN);
            } else {
                for (int64_t i = 0; i < n; ++i) scores[i] = vo_l2sqr_direct(q, xb + i * (int64_t)d, d);
            }
        } else if (metric == VO_METRIC_COSINE) {
            /* Cosine.run without cached norms (Operations/Scoring/Cosine.swift:38-131, the branch ScoreBlock.run takes
             * for a FlatIndexOptimized without a cosineNormsHandle, ScoreBlock.swift:42-58): InnerProduct.run, then per
             * row v = (dot * qInv) * inv with inv = 1 / (sqrt(l2NormSquared(row)) + 1e-12), clamped to [-1, 1] */
            const float qinv = vo_cosine_inv_norm(q, d);
            for (int64_t i = 0; i < n; ++i) {
                const float* row = xb + i * (int64_t)d;
                const float v = (vo_ip(q, row, d) * qinv) * vo_cosine_inv_norm(row, d);
                scores[i] = v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v);
            }
        } else {
            for (int64_t i = 0; i < n; ++i) scores[i] = vo_ip(q, xb + i * (int64_t)d, d);
        }
        int got = vo_select_topk(scores, NULL, n, k, metric == VO_METRIC_L2 ? VO_ORDER_MIN : VO_ORDER_MAX,
                                 ts, ti);
        for (int t = 0; t < k; ++t) {
            int64_t o = qi * (int64_t)k + t;
            if (t < got) {
                out_ids[o] = ti[t];
                /* FlatIndexOptimized.swift:457-474: L2 => sqrt, dot => -dot, cosine => 1 - similarity */
                out_dist[o] = (metric == VO_METRIC_L2) ? sqrtf(ts[t]) : (metric == VO_METRIC_COSINE ? 1.0f - ts[t] : -ts[t]);
                if (out_raw) out_raw[o] = ts[t];
            } else {
                out_ids[o] = -1;
                out_dist[o] = NAN;
                if (out_raw) out_raw[o] = NAN;
            }
        }
        free(scores); free(ts); free(ti);
    }
}

/* IVF-PQ query = the composition written down in
 * /root/reference/docs/kernel-specs/DONE_22_adc_scan.md:831-881:
 *   probe (batch form: CentroidBatchScore + ordered prefix, IVFIndex.swift:905-927)
 *   -> per probed list: pq_lut_residual_l2_f32 (defaults; dot-trick iff cb_norms given && ks>=64)
 *   -> adc_scan_u8 (defaults) -> selectTopK(k, .min) over local positions
 *   -> ids via list ids -> mergeTopK(.min).
 * Lists are CSR: list l occupies rows [list_offsets[l], list_offsets[l+1]) of codes/ids.
 * metric IP (no reference arithmetic exists, SURVEY 0.7 -- PARITY UNPINNED): probe by -<q,c>,
 * LUT[j][k] = <q_j, cb_jk> (vo_lut_dot order), ADC with bias = <q, c_list> (sequential dot),
 * selection .max, API distance = -score.
 * ids must fit int32 for the reference's TopK (TopK.swift:59); the oracle asserts nothing and
 * truncates exactly like the reference would. */
void vo_ivfpq_search(const float* queries, int64_t nq, int d, const float* coarse, int kc,
                     const float* coarse_norms, int m, int ks, const float* codebooks,
                     const float* cb_norms, const int64_t* list_offsets, const uint8_t* codes,
                     const int64_t* ids, int nprobe, int k, int metric,
                     float* out_dist, int64_t* out_ids, int32_t* out_probes) {
    if (k <= 0) return;
    int npe = nprobe < kc ? nprobe : kc;
    int ord = (metric == VO_METRIC_L2) ? VO_ORDER_MIN : VO_ORDER_MAX;
    int dsub = d / m;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t qi = 0; qi < nq; ++qi) {
        const float* q = queries + qi * (int64_t)d;
        int32_t* probes = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nprobe > 0 ? nprobe : 1));
        float* crow = (float*)malloc(sizeof(float) * (size_t)kc);
        vo_centroid_batch_score(q, 1, coarse, kc, d, metric, coarse_norms, crow);
        vo_probe_select(crow, kc, nprobe, probes, NULL);
        if (out_probes)
            for (int p = 0; p < nprobe; ++p) out_probes[qi * (int64_t)nprobe + p] = probes[p];
        float* lut = (float*)malloc(sizeof(float) * (size_t)m * ks);
        float* ls = (float*)malloc(sizeof(float) * (size_t)npe * k);
        int32_t* li = (int32_t*)malloc(sizeof(int32_t) * (size_t)npe * k);
        int32_t* lens = (int32_t*)calloc((size_t)npe, sizeof(int32_t));
        for (int p = 0; p < npe; ++p) {
            int l = probes[p];
            int64_t b = list_offsets[l], e = list_offsets[l + 1];
            int64_t len = e - b;
            if (len <= 0) { lens[p] = 0; continue; }
            float bias = 0.0f;
            if (metric == VO_METRIC_L2) {
                vo_pq_lut_residual_l2(q, coarse + (size_t)l * d, d, m, ks, codebooks, lut, cb_norms,
                                      -1, 1, 0);
            } else {
                for (int j = 0; j < m; ++j)
                    for (int kk = 0; kk < ks; ++kk)
                        lut[(size_t)j * ks + kk] =
                            vo_lut_dot(q + (size_t)j * dsub, codebooks + ((size_t)j * ks + kk) * dsub, dsub);
                const float* cl = coarse + (size_t)l * d;
                for (int t = 0; t < d; ++t) bias = bias + q[t] * cl[t];
            }
            float* dist = (float*)malloc(sizeof(float) * (size_t)len);
            /* ks = 16: packed nibbles, m / 2 bytes per row (adc_scan_u4, ADCScan.swift:384-456) */
            if (ks == 16) vo_adc_scan_u4(codes + b * (int64_t)(m / 2), len, m, ks, lut, dist, 0, bias, 0);
            else vo_adc_scan_u8(codes + b * (int64_t)m, len, m, ks, lut, dist, 0, bias, 0);
            float* ts = ls + (size_t)p * k;
            int32_t* ti = li + (size_t)p * k;
            int got = vo_select_topk(dist, NULL, len, k, ord, ts, ti);
            for (int t = 0; t < got; ++t) ti[t] = (int32_t)ids[b + ti[t]];
            /* local-position tie order == id order when list ids ascend (true for our builders);
             * re-sort by (score, id) so the merge precondition (best->worst lists) holds regardless */
            vo_pair* pr = (vo_pair*)malloc(sizeof(vo_pair) * (size_t)(got > 0 ? got : 1));
            for (int t = 0; t < got; ++t) { pr[t].s = ts[t]; pr[t].id = ti[t]; pr[t].ord = ord; }
            qsort(pr, (size_t)got, sizeof(vo_pair), pair_cmp_min);
            for (int t = 0; t < got; ++t) { ts[t] = pr[t].s; ti[t] = pr[t].id; }
            free(pr);
            lens[p] = got;
            free(dist);
        }
        float* ms = (float*)malloc(sizeof(float) * (size_t)k);
        int32_t* mi = (int32_t*)malloc(sizeof(int32_t) * (size_t)k);
        int got = vo_merge_topk(ls, li, lens, npe, k, k, ord, ms, mi);
        for (int t = 0; t < k; ++t) {
            int64_t o = qi * (int64_t)k + t;
            if (t < got) {
                out_ids[o] = mi[t];
                out_dist[o] = (metric == VO_METRIC_L2) ? ms[t] : -ms[t];
            } else { out_ids[o] = -1; out_dist[o] = NAN; }
        }
        free(ms); free(mi); free(lut); free(ls); free(li); free(lens); free(crow); free(probes);
    }
}

/* IVFIndex.swift:865-1042 (batchSearch, no filter): probes as above; candidates = union of probed
 * lists; exact per-candidate distance (VectorCore distanceSquared is un-vendored: restated with the
 * in-tree L2Sqr direct / InnerProduct kernels, which the reference's own tests hold to 1e-4 of it,
 * MicrokernelIntegrationTests.swift:5-46); ascending sort, truncate to k.  The reference's final
 * sort is unstable on ties (IVFIndex.swift:1038); the oracle uses (distance, id). */
void vo_ivfflat_search(const float* queries, int64_t nq, int d, const float* coarse, int kc,
                       const float* coarse_norms, const int64_t* list_offsets, const float* vecs,
                       const int64_t* ids, int nprobe, int k, int metric,
                       float* out_dist, int64_t* out_ids) {
    if (k <= 0) return;
    int npe = nprobe < kc ? nprobe : kc;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t qi = 0; qi < nq; ++qi) {
        const float* q = queries + qi * (int64_t)d;
        int32_t* probes = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nprobe > 0 ? nprobe : 1));
        float* crow = (float*)malloc(sizeof(float) * (size_t)kc);
        vo_centroid_batch_score(q, 1, coarse, kc, d, metric, coarse_norms, crow);
        vo_probe_select(crow, kc, nprobe, probes, NULL);
        int64_t total = 0;
        for (int p = 0; p < npe; ++p) total += list_offsets[probes[p] + 1] - list_offsets[probes[p]];
        vo_pair* pr = (vo_pair*)malloc(sizeof(vo_pair) * (size_t)(total > 0 ? total : 1));
        int64_t c = 0;
        for (int p = 0; p < npe; ++p) {
            int l = probes[p];
            for (int64_t r = list_offsets[l]; r < list_offsets[l + 1]; ++r) {
                const float* v = vecs + r * (int64_t)d;
                float dist;
                if (metric == VO_METRIC_L2) dist = sqrtf(vo_l2sqr_direct(q, v, d));
                else if (metric == VO_METRIC_IP) dist = -vo_ip(q, v, d);
                else {
                    /* DistanceUtils.swift:22-38: 1 - clamp(dot / sqrt(|a|^2 |b|^2)), 1 under the near-zero guard.
                     * dot / sumOfSquares are VectorCore's (not in the tree): stand-ins, tolerance parity only. */
                    float a2 = 0.0f, b2 = 0.0f;
                    for (int t = 0; t < d; ++t) { a2 = a2 + q[t] * q[t]; b2 = b2 + v[t] * v[t]; }
                    const float denom = sqrtf(a2 * b2);
                    if (!(denom > 1.1920929e-07f)) dist = 1.0f;
                    else {
                        float sim = vo_ip(q, v, d) / denom;
                        sim = sim > 1.0f ? 1.0f : (sim < -1.0f ? -1.0f : sim);
                        dist = 1.0f - sim;
                    }
                }
                pr[c].s = dist; pr[c].id = (int32_t)ids[r]; pr[c].ord = VO_ORDER_MIN;
                ++c;
            }
        }
        qsort(pr, (size_t)c, sizeof(vo_pair), pair_cmp_min);
        for (int t = 0; t < k; ++t) {
            int64_t o = qi * (int64_t)k + t;
            if (t < c) { out_ids[o] = pr[t].id; out_dist[o] = pr[t].s; }
            else { out_ids[o] = -1; out_dist[o] = NAN; }
        }
        free(pr); free(crow); free(probes);
    }
}
