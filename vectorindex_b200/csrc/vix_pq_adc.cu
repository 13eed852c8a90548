// vix_pq_adc.cu -- kernel-level PQ look-up tables and ADC scans with the reference's arithmetic:
//
//   a13  vix_pq_lut_batch_l2_f32 / vix_pq_lut_residual_l2_f32    (Operations/Quantization/PQLUT.swift)
//   a14  vix_adc_scan_u8 / vix_adc_scan_u4                        (Operations/Quantization/ADCScan.swift)
//
// These are the drop-in, materialising forms (a LUT in, a distance per code row out) and follow the
// reference's summation orders exactly (8-lane LUT sums; 4-accumulator / Kahan / sequential scan
// sums), so they compare bit-for-bit with the oracle.  The search path proper never materialises
// either array: see the fused kernel in vix_index.cu.
#include "vix_exact.cuh"

namespace vix {

// One thread per LUT entry (pair, j, k).  coarse == nullptr: pq_lut_l2_f32 (PQLUT.swift:191-261);
// otherwise pq_lut_residual_l2_f32 (:266-386) with coarse row coarse_ids[pair].
__global__ void pq_lut_kernel(const float* __restrict__ queries, const int32_t* __restrict__ coarse_ids,
                              const float* __restrict__ coarse, int64_t nq, int d, int m, int ks, int dsub,
                              const float* __restrict__ codebooks, const float* __restrict__ cnorms, int use_dot,
                              int include_q, int strict_fp, float* __restrict__ luts) {
    const int64_t total = nq * (int64_t)m * ks;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(e % ks);
        const int j = (int)((e / ks) % m);
        const int64_t qi = e / ((int64_t)ks * m);
        const float* qj = queries + qi * (int64_t)d + (size_t)j * dsub;
        const float* c = codebooks + ((size_t)j * ks + k) * dsub;
        const float* gj = coarse ? coarse + (int64_t)coarse_ids[qi] * d + (size_t)j * dsub : nullptr;
        const int len8 = dsub & ~7;
        float val;
        if (use_dot) {
            float qn = 0.0f;
            float dp;
            if (!gj) {
                if (include_q) qn = strict_fp ? exact_pair<SpecSeqDot>(qj, qj, dsub) : exact_pair<SpecLut8Dot>(qj, qj, dsub);
                dp = strict_fp ? exact_pair<SpecSeqDot>(qj, c, dsub) : exact_pair<SpecLut8Dot>(qj, c, dsub);
            } else {
                // rNorm = ||q_j - coarse_j||^2 ; dp = sum (q - coarse) * c   (PQLUT.swift:293-341)
                if (include_q) qn = strict_fp ? exact_pair<SpecSeqL2>(qj, gj, dsub) : exact_pair<SpecLut8L2>(qj, gj, dsub);
                if (strict_fp) {
                    dp = 0.0f;
                    for (int i = 0; i < dsub; ++i) dp = fadd(dp, fmul(fsub(qj[i], gj[i]), c[i]));
                } else {
                    float outer = 0.0f;
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        float inner = 0.0f;
#pragma unroll
                        for (int l = 0; l < 4; ++l) {
                            float acc = 0.0f;
                            for (int i = 4 * o + l; i < len8; i += 8) acc = fadd(acc, fmul(fsub(qj[i], gj[i]), c[i]));
                            inner = (l == 0) ? acc : fadd(inner, acc);
                        }
                        outer = (o == 0) ? inner : fadd(outer, inner);
                    }
                    dp = outer;
                    for (int i = len8; i < dsub; ++i) dp = fadd(dp, fmul(fsub(qj[i], gj[i]), c[i]));
                }
            }
            val = fsub(fadd(include_q ? qn : 0.0f, cnorms[(size_t)j * ks + k]), fmul(2.0f, dp));
        } else {
            if (!gj) {
                val = strict_fp ? exact_pair<SpecSeqL2>(qj, c, dsub) : exact_pair<SpecLut8L2>(qj, c, dsub);
            } else if (strict_fp) {
                float s = 0.0f;
                for (int i = 0; i < dsub; ++i) {
                    float df = fsub(fsub(qj[i], gj[i]), c[i]);
                    s = fadd(s, fmul(df, df));
                }
                val = s;
            } else {
                float outer = 0.0f;
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    float inner = 0.0f;
#pragma unroll
                    for (int l = 0; l < 4; ++l) {
                        float acc = 0.0f;
                        for (int i = 4 * o + l; i < len8; i += 8) {
                            float r = fsub(fsub(qj[i], gj[i]), c[i]);
                            acc = fadd(acc, fmul(r, r));
                        }
                        inner = (l == 0) ? acc : fadd(inner, acc);
                    }
                    outer = (o == 0) ? inner : fadd(outer, inner);
                }
                float s = outer;
                for (int i = len8; i < dsub; ++i) {
                    float df = fsub(fsub(qj[i], gj[i]), c[i]);
                    s = fadd(s, fmul(df, df));
                }
                val = s;
            }
        }
        luts[e] = val;
    }
}

int pq_lut_device(const float* queries, const int32_t* coarse_ids, const float* coarse, int64_t nq, int d, int m,
                  int ks, const float* codebooks, const float* cnorms, int use_dot, int include_q, int strict_fp,
                  float* luts) {
    const int64_t total = nq * (int64_t)m * ks;
    if (total == 0) return VIX_OK;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    pq_lut_kernel<<<(unsigned)blocks, 256, 0, ctx().stream>>>(queries, coarse_ids, coarse, nq, d, m, ks, d / m,
                                                             codebooks, cnorms, use_dot, include_q, strict_fp, luts);
    VIX_LAUNCH_CHECK();
    return VIX_OK;
}

// ------------------------------------------------------------------------------------------------
// ADC scan, reference order.  The LUT is staged in shared memory when it fits; one thread per row.
// u8: ADCScan.swift:190-283 (4 accumulators s[j mod 4], leftovers to s0, Kahan when strictFP && m>=64)
// u4: ADCScan.swift:384-456 (packed nibbles, ONE sequential accumulator, Kahan as above)
// interleavedBlock (layout 1): code(i, j) = codes[(i/g)*m*g + j*g + i%g]   (ADCScan.swift:288-379)
// ------------------------------------------------------------------------------------------------
template <bool U4>
__global__ void __launch_bounds__(256)
adc_scan_kernel(const uint8_t* __restrict__ codes, int64_t n, int m, int ks, const float* __restrict__ lut_g,
                int lut_in_smem, float* __restrict__ out, int layout, int g, int stride, float bias, int kahan) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const float* lut = lut_g;
    if (lut_in_smem) {
        float* s = reinterpret_cast<float*>(smem_raw);
        for (int e = threadIdx.x; e < m * ks; e += blockDim.x) s[e] = lut_g[e];
        __syncthreads();
        lut = s;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        auto code_at = [&](int j) -> int {
            if (U4) {
                // u4 is AoS only: byte j/2, low nibble = even subspace
                uint8_t byte = codes[i * (int64_t)stride + (j >> 1)];
                return (j & 1) ? ((byte >> 4) & 0x0F) : (byte & 0x0F);
            }
            if (layout == 1) return codes[(i / g) * (int64_t)m * g + (int64_t)j * g + (i % g)];
            return codes[i * (int64_t)stride + j];
        };
        float result;
        if (kahan) {
            float sum = 0.0f, c = 0.0f;
            for (int j = 0; j < m; ++j) {
                float value = lut[(size_t)j * ks + code_at(j)];
                float y = fsub(value, c);
                float t = fadd(sum, y);
                c = fsub(fsub(t, sum), y);
                sum = t;
            }
            result = fadd(sum, bias);
        } else if (U4 || layout == 1) {
            // u4 and interleaved u8: ONE sequential accumulator (ADCScan.swift:438-447, 364-378)
            float sum = 0.0f;
            for (int j = 0; j < m; ++j) sum = fadd(sum, lut[(size_t)j * ks + code_at(j)]);
            result = fadd(sum, bias);
        } else {
            float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
            int j = 0;
            for (; j + 3 < m; j += 4) {
                s0 = fadd(s0, lut[(size_t)(j + 0) * ks + code_at(j + 0)]);
                s1 = fadd(s1, lut[(size_t)(j + 1) * ks + code_at(j + 1)]);
                s2 = fadd(s2, lut[(size_t)(j + 2) * ks + code_at(j + 2)]);
                s3 = fadd(s3, lut[(size_t)(j + 3) * ks + code_at(j + 3)]);
            }
            for (; j < m; ++j) s0 = fadd(s0, lut[(size_t)j * ks + code_at(j)]);
            result = fadd(fadd(fadd(fadd(s0, s1), s2), s3), bias);
        }
        out[i] = result;
    }
}

template <bool U4>
static int adc_scan_entry(const char* fn, const uint8_t* codes, int64_t n, int m, int ks, const float* lut,
                          float* out, const vix_adc_scan_opts* opts) {
    VIX_TRY(ensure_device());
    VIX_REQUIRE(codes && lut && out, VIX_ERR_NULL_PTR, "%s: null pointer", fn);
    VIX_REQUIRE(n >= 0 && m > 0, VIX_ERR_INVALID_DIM, "%s: bad n/m", fn);
    VIX_REQUIRE(ks == (U4 ? 16 : 256), VIX_ERR_INVALID_K, "%s: ks must be %d", fn, U4 ? 16 : 256);
    VIX_REQUIRE(!U4 || (m & 1) == 0, VIX_ERR_INVALID_DIM, "%s: m must be even", fn);
    if (n == 0) return VIX_OK;
    int layout = 0, g = 8, stride = 0, strict = 0;
    float bias = 0.0f;
    if (opts) {
        layout = opts->layout; g = opts->group_size; stride = opts->stride; bias = opts->add_bias;
        strict = opts->strict_fp ? 1 : 0;
    }
    VIX_REQUIRE(layout == 0 || (layout == 1 && !U4 && g > 0), VIX_ERR_INVALID_LAYOUT,
                "%s: unsupported layout/group size", fn);   // ADCScan.swift:297 (groupSize must be > 0)
    const int row_bytes = U4 ? m / 2 : m;
    if (stride <= 0) stride = row_bytes;
    VIX_REQUIRE(stride >= row_bytes, VIX_ERR_INVALID_PARAM, "%s: stride < row bytes", fn);
    size_t code_bytes = (layout == 1) ? (size_t)((n + g - 1) / g) * m * g : (size_t)(n - 1) * stride + row_bytes;
    In<uint8_t> dc;
    In<float> dl;
    Out<float> dout;
    VIX_TRY(dc.stage(codes, code_bytes));
    VIX_TRY(dl.stage(lut, (size_t)m * ks));
    VIX_TRY(dout.stage(out, (size_t)n));
    const size_t lut_bytes = (size_t)m * ks * 4;
    const int in_smem = lut_bytes <= 200 * 1024;
    auto kern = adc_scan_kernel<U4>;
    if (in_smem) VIX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lut_bytes));
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, 256, in_smem ? lut_bytes : 0, ctx().stream>>>(dc.dev, n, m, ks, dl.dev, in_smem, dout.dev,
                                                                          layout, g, stride, bias,
                                                                          strict && m >= 64);
    VIX_LAUNCH_CHECK();
    VIX_TRY(dout.commit());
    return finish(dout.is_host());
}

static int lut_entry(const char* fn, const float* queries, const int32_t* coarse_ids, const float* coarse,
                     bool residual, int64_t nq, int d, int m, int ks, const float* codebooks, float* luts,
                     const float* cnorms, const vix_pq_lut_opts* opts) {
    VIX_TRY(ensure_device());
    VIX_REQUIRE(queries && codebooks && luts, VIX_ERR_NULL_PTR, "%s: null pointer", fn);
    VIX_REQUIRE(!residual || (coarse && coarse_ids), VIX_ERR_NULL_PTR, "%s: coarse centroids / ids null", fn);
    VIX_REQUIRE(d > 0 && m > 0 && d % m == 0 && nq >= 0, VIX_ERR_INVALID_DIM, "%s: need d %% m == 0", fn);
    VIX_REQUIRE(ks > 0, VIX_ERR_INVALID_K, "%s: ks must be > 0", fn);
    if (nq == 0) return VIX_OK;
    int use_dot = -1, include_q = 1, strict = 0;
    if (opts) { use_dot = opts->use_dot_trick; include_q = opts->include_q_norm ? 1 : 0; strict = opts->strict_fp ? 1 : 0; }
    if (use_dot < 0) use_dot = (cnorms != nullptr && ks >= 64) ? 1 : 0;   // PQLUT.swift:207-210
    VIX_REQUIRE(!use_dot || cnorms, VIX_ERR_NULL_PTR, "%s: dot-trick needs centroid_norms", fn);
    In<float> dq, dcb, dcn, dco;
    In<int32_t> dids;
    Out<float> dl;
    VIX_TRY(dq.stage(queries, (size_t)nq * d));
    VIX_TRY(dcb.stage(codebooks, (size_t)m * ks * (d / m)));
    VIX_TRY(dcn.stage(cnorms, cnorms ? (size_t)m * ks : 0));
    if (residual) {
        VIX_TRY(dids.stage(coarse_ids, (size_t)nq));
        int64_t rows = 0;
        if (!is_device_ptr(coarse)) {
            VIX_REQUIRE(!is_device_ptr(coarse_ids), VIX_ERR_INVALID_PARAM, "%s: host coarse with device ids", fn);
            for (int64_t i = 0; i < nq; ++i) {
                VIX_REQUIRE(coarse_ids[i] >= 0, VIX_ERR_INVALID_PARAM, "%s: negative coarse id", fn);
                if (coarse_ids[i] + 1 > rows) rows = coarse_ids[i] + 1;
            }
        }
        VIX_TRY(dco.stage(coarse, (size_t)rows * d));
    }
    VIX_TRY(dl.stage(luts, (size_t)nq * m * ks));
    VIX_TRY(pq_lut_device(dq.dev, residual ? dids.dev : nullptr, residual ? dco.dev : nullptr, nq, d, m, ks, dcb.dev,
                          dcn.dev, use_dot, include_q, strict, dl.dev));
    VIX_TRY(dl.commit());
    return finish(dl.is_host());
}

}  // namespace vix

using namespace vix;

namespace vix {
// pq_query_subnorms_f32 (PQLUT.swift:174-187): out[q][j] = _simd_dot(q_j, q_j, dsub), the LUT kernels' own reduction
__global__ void query_subnorms_kernel(const float* __restrict__ q, int64_t total, int m, int dsub, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (query, sub-space)
    if (i >= total) return;
    const float* qj = q + i * dsub;                                        // rows are contiguous: q[(query * m + j) * dsub]
    out[i] = exact_pair<SpecLut8Dot>(qj, qj, dsub);
    (void)m;
}
}  // namespace vix

extern "C" {

int vix_pq_query_subnorms_f32(const float* queries, int64_t nq, int d, int m, float* out) {
    VIX_TRY(ensure_device());
    VIX_REQUIRE(queries && out, VIX_ERR_NULL_PTR, "vix_pq_query_subnorms_f32: null pointer");
    VIX_REQUIRE(d > 0 && m > 0 && d % m == 0, VIX_ERR_INVALID_DIM, "vix_pq_query_subnorms_f32: d must be divisible by m");
    if (nq <= 0) return VIX_OK;
    In<float> dq;
    Out<float> dout;
    VIX_TRY(dq.stage(queries, (size_t)nq * d));
    VIX_TRY(dout.stage(out, (size_t)nq * m));
    const int64_t total = nq * (int64_t)m;
    query_subnorms_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx().stream>>>(dq.dev, total, m, d / m, dout.dev);
    VIX_LAUNCH_CHECK();
    VIX_TRY(dout.commit());
    return finish(dout.is_host());
}

int vix_pq_lut_batch_l2_f32(const float* queries, int64_t nq, int d, int m, int ks, const float* codebooks,
                            float* luts, const float* centroid_norms, const vix_pq_lut_opts* opts) {
    return lut_entry("vix_pq_lut_batch_l2_f32", queries, nullptr, nullptr, false, nq, d, m, ks, codebooks, luts,
                     centroid_norms, opts);
}

int vix_pq_lut_residual_l2_f32(const float* queries, const int32_t* coarse_ids, int64_t nq, int d,
                               const float* coarse_centroids, int m, int ks, const float* codebooks, float* luts,
                               const float* centroid_norms, const vix_pq_lut_opts* opts) {
    return lut_entry("vix_pq_lut_residual_l2_f32", queries, coarse_ids, coarse_centroids, true, nq, d, m, ks,
                     codebooks, luts, centroid_norms, opts);
}

int vix_adc_scan_u8(const uint8_t* codes, int64_t n, int m, int ks, const float* lut, float* out,
                    const vix_adc_scan_opts* opts) {
    return adc_scan_entry<false>("vix_adc_scan_u8", codes, n, m, ks, lut, out, opts);
}

int vix_adc_scan_u4(const uint8_t* codes, int64_t n, int m, int ks, const float* lut, float* out,
                    const vix_adc_scan_opts* opts) {
    return adc_scan_entry<true>("vix_adc_scan_u4", codes, n, m, ks, lut, out, opts);
}

}  // extern "C"
