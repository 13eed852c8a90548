// vix_pq_encode.cuh -- the reference's PQ-encode arithmetic (pq_encode.c, x86 scalar order) as device functions shared
// by the CUDA-core encoder (vix_pq_encode.cu) and the tensor-core shortlist encoder (vix_pq_tc.cu).
#pragma once

#include "vix_common.cuh"

namespace vix {

enum EncMode { ENC_CSQ = 0, ENC_CSQ_RES = 1, ENC_DOT = 2, ENC_DIRECT = 3, ENC_DOT_RES = 4, ENC_DIRECT_RES = 5 };

// pq_encode.c:260-276 (idx_layout_u8)
__device__ __forceinline__ size_t code_index(int64_t i, int j, int64_t n, int m, int layout, int B, int g) {
    if (layout == PQ_LAYOUT_SOA_BLOCKED) {
        int64_t blocks = (n + B - 1) / B;
        return (size_t)((int64_t)j * blocks * B + (i / B) * B + (i % B));
    }
    if (layout == PQ_LAYOUT_INTERLEAVED_BLOCK) return (size_t)((i / g) * (int64_t)m * g + (int64_t)j * g + (i % g));
    return (size_t)(i * (int64_t)m + j);
}

// One (vector, subspace) argmin in the reference order.  xs/gs: this thread's sub-vector (registers
// when DSUB > 0, else strided shared memory); cb/csq: shared memory.
template <int MODE, int DSUB, typename XF, typename GF>
__device__ __forceinline__ void encode_chunk(XF xf, GF gf, const float* __restrict__ cb,
                                             const float* __restrict__ csq, int k0, int k1, int dsub_rt,
                                             float base2, float& bd, int& bk) {
    const int dsub = DSUB > 0 ? DSUB : dsub_rt;
    for (int k = k0; k < k1; ++k) {
        const float* c = cb + (size_t)(k - k0) * dsub;
        float dist;
        if (MODE == ENC_CSQ) {
            // encode_subspace_u8_dot_with_csq (pq_encode.c:332-366): x2 + csq[k] - 2*dot
            float dot = 0.0f;
#pragma unroll
            for (int i = 0; i < dsub; ++i) dot = fadd(dot, fmul(xf(i), c[i]));
            dist = fsub(fadd(base2, csq[k - k0]), fmul(2.0f, dot));
        } else if (MODE == ENC_CSQ_RES) {
            // encode_subspace_u8_residual_with_csq (:368-410): dot(x,c) - dot(g,c), sequential each
            float dx = 0.0f, dg = 0.0f;
#pragma unroll
            for (int i = 0; i < dsub; ++i) dx = fadd(dx, fmul(xf(i), c[i]));
#pragma unroll
            for (int i = 0; i < dsub; ++i) dg = fadd(dg, fmul(gf(i), c[i]));
            float dot = fsub(dx, dg);
            dist = fsub(fadd(base2, csq[k - k0]), fmul(2.0f, dot));
        } else if (MODE == ENC_DOT) {
            // dist_dp_scalar (:126-134): interleaved dot / c2, x2 + c2 - 2*dot
            float dot = 0.0f, c2 = 0.0f;
#pragma unroll
            for (int i = 0; i < dsub; ++i) {
                float ci = c[i];
                dot = fadd(dot, fmul(xf(i), ci));
                c2 = fadd(c2, fmul(ci, ci));
            }
            dist = fsub(fadd(base2, c2), fmul(2.0f, dot));
        } else if (MODE == ENC_DIRECT) {
            // l2_sq_scalar (:83-90)
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < dsub; ++i) {
                float df = fsub(xf(i), c[i]);
                acc = fadd(acc, fmul(df, df));
            }
            dist = acc;
        } else if (MODE == ENC_DOT_RES) {
            // dist_dp_residual_scalar (:246-257)
            float dot = 0.0f, c2 = 0.0f;
#pragma unroll
            for (int i = 0; i < dsub; ++i) {
                float ri = fsub(xf(i), gf(i));
                float ci = c[i];
                dot = fadd(dot, fmul(ri, ci));
                c2 = fadd(c2, fmul(ci, ci));
            }
            dist = fsub(fadd(base2, c2), fmul(2.0f, dot));
        } else {
            // l2_sq_residual_scalar (:199-207): ((x - g) - c)^2
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < dsub; ++i) {
                float r = fsub(fsub(xf(i), gf(i)), c[i]);
                acc = fadd(acc, fmul(r, r));
            }
            dist = acc;
        }
        // pq_argmin_update (:74-80); k ascends, so "dist == bd && k < bk" can never fire
        if (k == 0 || dist < bd) { bd = dist; bk = k; }
    }
}

}  // namespace vix
