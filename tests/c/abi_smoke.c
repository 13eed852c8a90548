/* A plain C11 consumer of the C ABI (no torch, no Python, no C++): what a cgo / SwiftPM system-library target sees.
 * Built and run by tests/test_c_consumer.py.  Without a CUDA device every compute entry point must fail loudly with
 * VIX_ERR_NO_DEVICE; with one, a small exact flat search and a PQ encode through the reference's cpq_* symbols run. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cpq_encode.h"
#include "vindex_cuda.h"

static int fail(const char* what, int rc) {
    fprintf(stderr, "FAIL %s: status %d (%s)\n", what, rc, vix_last_error());
    return 1;
}

int main(void) {
    if (vix_version() < 100) return fail("vix_version", vix_version());
    if (sizeof(PQEncodeOpts) != 24) return fail("sizeof(PQEncodeOpts)", (int)sizeof(PQEncodeOpts));
    enum { N = 64, D = 8, NQ = 3, K = 4 };
    float xb[N * D], q[NQ * D], dist[NQ * K];
    int64_t ids[NQ * K];
    for (int i = 0; i < N * D; ++i) xb[i] = (float)((i * 37) % 101) * 0.01f;
    for (int r = 0; r < NQ; ++r) memcpy(q + r * D, xb + (5 + 7 * r) * D, sizeof(float) * D);   /* queries = rows 5, 12, 19 */
    const int rc = vix_flat_search_f32(q, NQ, xb, N, D, VIX_METRIC_L2, K, dist, ids);
    if (vix_device_count() < 1) {
        if (rc != VIX_ERR_NO_DEVICE) return fail("flat search without a device must report VIX_ERR_NO_DEVICE", rc);
        if (strlen(vix_last_error()) == 0) return fail("no error text", rc);
        vix_index_params p;
        vix_index_params_default(&p);
        vix_index_t* h = NULL;
        if (vix_index_create(&p, &h) != VIX_ERR_NO_DEVICE || h != NULL) return fail("index_create without a device", 0);
        printf("ok (no device: every entry point refuses, no CPU fallback)\n");
        return 0;
    }
    if (rc != VIX_OK) return fail("vix_flat_search_f32", rc);
    for (int r = 0; r < NQ; ++r) {
        if (ids[r * K] != 5 + 7 * r || dist[r * K] != 0.0f) return fail("nearest neighbour of a stored row is itself", (int)ids[r * K]);
        for (int t = 1; t < K; ++t)
            if (!(dist[r * K + t] >= dist[r * K + t - 1])) return fail("distances ascend", t);
    }
    /* the reference's own encoder symbol (include/cpq_encode.h), host pointers in and out */
    enum { M = 2, KS = 256, DSUB = D / M };
    float* cb = (float*)malloc(sizeof(float) * M * KS * DSUB);
    uint8_t codes[N * M];
    for (int i = 0; i < M * KS * DSUB; ++i) cb[i] = (float)((i * 53) % 97) * 0.011f;
    cpq_encode_u8_f32(xb, N, D, M, KS, cb, codes, NULL);
    for (int i = 0; i < N; ++i) {                                        /* brute-force argmin in plain C, first minimum wins */
        for (int j = 0; j < M; ++j) {
            int best = 0; float bd = INFINITY;
            for (int c = 0; c < KS; ++c) {
                float s = 0.0f;
                for (int e = 0; e < DSUB; ++e) { const float df = xb[i * D + j * DSUB + e] - cb[(j * KS + c) * DSUB + e]; s += df * df; }
                if (s < bd) { bd = s; best = c; }
            }
            float sc = 0.0f;                                             /* accept a different code only at an exact tie in distance */
            for (int e = 0; e < DSUB; ++e) { const float df = xb[i * D + j * DSUB + e] - cb[(j * KS + codes[i * M + j]) * DSUB + e]; sc += df * df; }
            if (codes[i * M + j] != best && fabsf(sc - bd) > 1e-6f * (1.0f + bd)) return fail("cpq_encode_u8_f32 code", i);
        }
    }
    free(cb);
    printf("ok (device: flat search + cpq_encode_u8_f32 through the C ABI)\n");
    return 0;
}
