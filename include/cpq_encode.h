/*
 * cpq_encode.h -- drop-in C ABI of the reference's CPQEncode module, implemented on B200.
 *
 * libvindex_b200.so exports exactly the eight symbols that
 * /root/reference/Sources/CPQEncode/include/cpq_encode.h:41-122 declares, with identical prototypes
 * and an identical PQEncodeOpts layout (LP64: size 24, offsets 0/4/5/8/12/16/20, checked by
 * /root/reference/tools/pq_align_check.c:22-28 and by tests/test_abi.py).  A SwiftPM C target whose
 * module map points at this header can therefore replace the CPQEncode target unchanged
 * (see INTEGRATION.md).
 *
 * Semantics kept from the reference (pq_encode.c):
 *   - codes are bit-exact with the reference's x86 scalar path (sequential dot, unfused multiply-add,
 *     `x2 + csq[k] - 2*dot` evaluated left to right, argmin tie -> smaller k, pq_encode.c:74-80);
 *   - opts == NULL => AoS layout, dot-trick iff ks >= 64 (pq_encode.c:468-476);
 *   - all three output layouts of idx_layout_u8 (pq_encode.c:260-276);
 *   - buffers are caller-owned; on this implementation each pointer may be a host pointer OR a CUDA
 *     device pointer (detected per call), so pipelines can keep data resident in HBM;
 *   - void return, like the reference.  The reference assert()s on bad arguments; this library
 *     never aborts: it records the failure (vix_last_error() in vindex_cuda.h) and leaves `codes`
 *     untouched.  There is NO CPU fallback: without a usable CUDA device every call fails loudly.
 */
#ifndef VIX_CPQ_ENCODE_H
#define VIX_CPQ_ENCODE_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef CPQ_RESTRICT
#if defined(__GNUC__) || defined(__clang__)
#define CPQ_RESTRICT __restrict
#else
#define CPQ_RESTRICT restrict
#endif
#endif

/* code layouts, same numbering as the reference enum */
typedef enum {
    PQ_LAYOUT_AOS = 0,               /* codes[i*m + j]                                   */
    PQ_LAYOUT_SOA_BLOCKED = 1,       /* codes[j*ceil(n/B)*B + (i/B)*B + i%B]   (padded)  */
    PQ_LAYOUT_INTERLEAVED_BLOCK = 2  /* codes[(i/g)*m*g + j*g + i%g]           (padded)  */
} PQLayout;

typedef struct {
    PQLayout layout;             /* default PQ_LAYOUT_AOS                       */
    bool     use_dot_trick;      /* default (ks >= 64)                          */
    bool     precompute_x_norm2; /* accepted, no effect on results              */
    int      prefetch_distance;  /* accepted, ignored (CPU cache hint)          */
    int      num_threads;        /* accepted, ignored (the GPU grid is sized by n) */
    int      soa_block_B;        /* B for SOA_BLOCKED (<=0 => 64)               */
    int      interleave_g;       /* g for INTERLEAVED_BLOCK (<=0 => 8)          */
} PQEncodeOpts;

/* u8 codes, ks must be 256.  x[n*d], codebooks[m*ks*dsub] (j,k,component), codes[n*m]. */
void cpq_encode_u8_f32(const float* CPQ_RESTRICT x, int64_t n, int d, int m, int ks,
                       const float* CPQ_RESTRICT codebooks, uint8_t* CPQ_RESTRICT codes,
                       const PQEncodeOpts* opts);

/* as above with precomputed centroid squared norms centroid_sq[m*ks] (always the dot-trick path) */
void cpq_encode_u8_f32_with_csq(const float* CPQ_RESTRICT x, int64_t n, int d, int m, int ks,
                                const float* CPQ_RESTRICT codebooks,
                                const float* CPQ_RESTRICT centroid_sq, uint8_t* CPQ_RESTRICT codes,
                                const PQEncodeOpts* opts);

/* u4 codes, ks must be 16, m even; two codes per byte (low nibble = even subspace): codes[n*(m/2)] */
void cpq_encode_u4_f32(const float* CPQ_RESTRICT x, int64_t n, int d, int m, int ks,
                       const float* CPQ_RESTRICT codebooks, uint8_t* CPQ_RESTRICT codes,
                       const PQEncodeOpts* opts);

/* residual (IVF-PQ) variants: encode x[i] - coarse_centroids[assignments[i]] without materialising it */
void cpq_encode_residual_u8_f32(const float* CPQ_RESTRICT x, int64_t n, int d, int m, int ks,
                                const float* CPQ_RESTRICT codebooks,
                                const float* CPQ_RESTRICT coarse_centroids,
                                const int32_t* CPQ_RESTRICT assignments,
                                uint8_t* CPQ_RESTRICT codes, const PQEncodeOpts* opts);

void cpq_encode_residual_u8_f32_with_csq(const float* CPQ_RESTRICT x, int64_t n, int d, int m, int ks,
                                         const float* CPQ_RESTRICT codebooks,
                                         const float* CPQ_RESTRICT centroid_sq,
                                         const float* CPQ_RESTRICT coarse_centroids,
                                         const int32_t* CPQ_RESTRICT assignments,
                                         uint8_t* CPQ_RESTRICT codes, const PQEncodeOpts* opts);

void cpq_encode_residual_u4_f32(const float* CPQ_RESTRICT x, int64_t n, int d, int m, int ks,
                                const float* CPQ_RESTRICT codebooks,
                                const float* CPQ_RESTRICT coarse_centroids,
                                const int32_t* CPQ_RESTRICT assignments,
                                uint8_t* CPQ_RESTRICT codes, const PQEncodeOpts* opts);

/* nibble pack / unpack of ONE vector's m codes (host-side helpers, m even) */
void cpq_pack_u4_bulk(const uint8_t* CPQ_RESTRICT codes, int m, uint8_t* CPQ_RESTRICT packed);
void cpq_unpack_u4_bulk(const uint8_t* CPQ_RESTRICT packed, int m, uint8_t* CPQ_RESTRICT codes);

#ifdef __cplusplus
}
#endif
#endif /* VIX_CPQ_ENCODE_H */
