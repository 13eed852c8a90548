/*
 * vix_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
 *
 * A plain-C restatement of the reference's (gifton/VectorIndex) CPU arithmetic for the batched
 * search hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  Every function cites the reference file:line it follows
 * (paths relative to /root/reference/Sources/VectorIndex unless stated).
 *
 * Build: gcc -O2 -ffp-contract=off [-fopenmp]  (no -march=native, no -ffast-math): all products and
 * sums are separate IEEE fp32 operations, exactly as Swift emits them (no FMA contraction).
 *
 * Parity status (see DESIGN.md "Oracle pins"):
 *   - PQ encode: pinned against the reference's own C encoder compiled unmodified (oracle/_ref).
 *   - PQ streaming train: pinned against the bit-level golden vector PQTrainTests.swift:813-816.
 *   - k-means++ / mini-batch k-means / IVF search: pinned against the published recall
 *     0.9565000000000008 (.bench/post-phase3/ivf_search.json:54).
 *   - TopK / probe tie-breaks: pinned against TelemetryRecorderTests.swift:229-241,
 *     IVFSelectTests.swift:305-347.
 *   - PQ LUT + ADC scan: PARITY UNPINNED by any reference test (the reference has no live test or
 *     caller for adc_scan_u8 / pq_lut_*); authority is the source text only.
 */
#ifndef VIX_ORACLE_H
#define VIX_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { VO_METRIC_L2 = 0, VO_METRIC_IP = 1, VO_METRIC_COSINE = 2 };
enum { VO_ORDER_MIN = 0, VO_ORDER_MAX = 1 };

/* ---- scalar pair kernels (Appendix A of SURVEY.md) ---- */
float vo_l2sqr_direct(const float* q, const float* x, int d);
float vo_norm_l2sq(const float* x, int d);
float vo_l2sqr_dot_fused(const float* q, const float* row, int d, float q_norm, float x_norm);
float vo_ip(const float* q, const float* x, int d);
float vo_km12_l2sq(const float* a, const float* b, int d);
float vo_km11_l2sq(const float* a, const float* b, int d);
float vo_pqtrain_l2sq(const float* a, const float* b, int d);
float vo_lut_l2sqr(const float* a, const float* b, int len);
float vo_lut_dot(const float* a, const float* b, int len);
float vo_pq_sqnorm(const float* a, int d);

/* ---- block scoring ---- */
/* OpenMP threads of the oracle's row / query loops: n > 0 sets, returns the count in effect */
int vo_set_threads(int n);

void vo_l2sqr_block(const float* q, const float* xb, int64_t n, int d, float* out,
                    const float* xb_norm, float q_norm);
void vo_ip_block(const float* q, const float* xb, int64_t n, int d, float* out);
void vo_score_block(const float* q, const float* xb, int64_t n, int d, int metric, float* out);

/* ---- selection ---- */
int vo_select_topk(const float* scores, const int32_t* ids, int64_t n, int k, int ordering,
                   float* out_scores, int32_t* out_ids);
int vo_merge_topk(const float* scores, const int32_t* ids, const int32_t* lens, int nlists,
                  int list_stride, int k, int ordering, float* out_scores, int32_t* out_ids);

/* ---- coarse quantiser ---- */
void vo_centroid_norms(const float* c, int kc, int d, float* out);
void vo_centroid_batch_score(const float* queries, int64_t q, const float* centroids, int kc, int d,
                             int metric, const float* centroid_norms, float* out);
void vo_probe_select(const float* scores, int kc, int nprobe, int32_t* out_idx, float* out_scores);
void vo_probe_select_batch(const float* queries, int64_t q, const float* centroids, int kc, int d,
                           int metric, const float* centroid_norms, int nprobe,
                           int32_t* out_idx, float* out_scores);
void vo_assign(const float* x, int64_t n, const float* c, int kc, int d, int32_t* assign_out,
               float* dist_out);
void vo_assign_metric(const float* x, int64_t n, const float* c, int kc, int d, int metric,
                      const float* centroid_norms, int32_t* assign_out);

/* ---- PQ encode (restatement of Sources/CPQEncode/pq_encode.c, x86 scalar path) ---- */
void vo_pq_encode_u8(const float* x, int64_t n, int d, int m, int ks, const float* codebooks,
                     const float* centroid_sq, const float* coarse, const int32_t* assign,
                     int use_dot, uint8_t* codes);
void vo_pq_encode_u4(const float* x, int64_t n, int d, int m, int ks, const float* codebooks,
                     const float* coarse, const int32_t* assign, uint8_t* codes);
void vo_pq_centroid_sq_swift(const float* codebooks, int m, int ks, int dsub, float* out);
void vo_pq_centroid_sq_seq(const float* codebooks, int m, int ks, int dsub, float* out);

/* ---- PQ LUT + ADC ---- */
void vo_pq_lut_l2(const float* q, int d, int m, int ks, const float* codebooks, float* lut,
                  const float* centroid_norms, const float* q_sub_norms, int use_dot,
                  int include_q, int strict_fp);
void vo_pq_lut_residual_l2(const float* q, const float* coarse, int d, int m, int ks,
                           const float* codebooks, float* lut, const float* centroid_norms,
                           int use_dot, int include_q, int strict_fp);
void vo_adc_scan_u8(const uint8_t* codes, int64_t n, int m, int ks, const float* lut, float* out,
                    int stride, float bias, int strict_fp);
void vo_adc_scan_u4(const uint8_t* codes, int64_t n, int m, int ks, const float* lut, float* out,
                    int stride, float bias, int strict_fp);

/* ---- composed searches ---- */
void vo_flat_search(const float* queries, int64_t nq, const float* xb, int64_t n, int d,
                    int metric, int k, float* out_dist, int64_t* out_ids, float* out_raw);
void vo_ivfpq_search(const float* queries, int64_t nq, int d, const float* coarse, int kc,
                     const float* coarse_norms, int m, int ks, const float* codebooks,
                     const float* cb_norms, const int64_t* list_offsets, const uint8_t* codes,
                     const int64_t* ids, int nprobe, int k, int metric,
                     float* out_dist, int64_t* out_ids, int32_t* out_probes);
void vo_ivfflat_search(const float* queries, int64_t nq, int d, const float* coarse, int kc,
                       const float* coarse_norms, const int64_t* list_offsets, const float* vecs,
                       const int64_t* ids, int nprobe, int k, int metric,
                       float* out_dist, int64_t* out_ids);

/* ---- RNGs + trainers ---- */
typedef struct { uint64_t s; } vo_lcg;
void     vo_lcg_init(vo_lcg* r, uint64_t seed, uint64_t stream);
uint64_t vo_lcg_next(vo_lcg* r);

int vo_kmeanspp_seed(const float* data, int64_t n, int d, int k, uint64_t seed, uint64_t stream,
                     float* centroids_out, int64_t* chosen_out);
int vo_kmeans_minibatch(const float* x, int64_t n, int d, int kc, const float* init_centroids,
                        int batch_size, int epochs, float tol, uint64_t seed, uint64_t stream,
                        float* centroids_out, int32_t* assign_out, int64_t* empties_per_batch,
                        int empties_cap, int* epochs_done, int64_t* batches_done);

typedef struct {
    int      ks;
    int      m;
    int      algorithm;      /* 0 = lloyd, 1 = minibatch */
    int      max_iters;
    int      batch_size;
    int      empty_policy;   /* 0 = split, 1 = reseed, 2 = ignore */
    int64_t  sample_n;
    uint64_t seed;
    uint64_t stream_id;
    float    tol;
    int      precompute_x_norm2;
    int      compute_centroid_norms;
} vo_pq_train_cfg;
void vo_pq_train_cfg_default(vo_pq_train_cfg* c);
int  vo_pq_train(const float* x, int64_t n, int d, const vo_pq_train_cfg* cfg,
                 const float* coarse, const int32_t* assign,
                 float* codebooks_out, float* norms_out, double* distortion_out);
int  vo_pq_train_streaming(const float* const* chunks, const int64_t* chunk_n, int nchunks, int d,
                           const vo_pq_train_cfg* cfg, float* codebooks_out, float* norms_out);

#ifdef __cplusplus
}
#endif
#endif
