// vix_pq_encode.cu -- PQ encoding on B200: the eight cpq_* symbols of the reference's CPQEncode C
// module (/root/reference/Sources/CPQEncode/include/cpq_encode.h:41-122), bit-exact with the
// reference's x86 scalar arithmetic (/root/reference/Sources/CPQEncode/pq_encode.c).
//
// Kernel: one thread per vector; the CTA walks the m subspaces in order.  For each subspace the
// sub-codebook [ks x dsub] (+ optional ||c||^2) is staged in shared memory (every lane of a warp reads
// the same centroid component -> pure broadcast, no bank conflicts) and each thread runs the
// reference's k = 0..ks-1 loop in the reference's exact operation order with unfused multiply / add
// (strict '<' argmin == "tie -> smaller k", pq_encode.c:74-80).  AoS codes are staged in shared
// memory and written back as whole rows, so both the x reads (a [256 x d] slab per CTA) and the code
// writes (a contiguous [256 x m] byte slab) are full-sector HBM transactions.
//
// Roofline: 2*n*ks*d flop on CUDA cores against 4*n*d + n*m bytes of HBM; at dsub = 8
// the kernel is fp32-issue bound (mul and add are separate instructions by contract).
#include "vix_common.cuh"
#include "vix_pq_encode.cuh"

#include <assert.h>

namespace vix {

constexpr int kEncThreads = 256;

// MODE: arithmetic variant; DSUB > 0: sub-vector in registers, DSUB == 0: generic (shared memory).
// THREADS = rows per CTA: 256, or 64 when the generic path stages long sub-vectors (d / m = 128 with residuals, the shape
// of the reference's ResidualKernelTests.swift:126-200, needs 2 x 128 x THREADS floats).
template <int MODE, int DSUB, int THREADS = kEncThreads>
__global__ void __launch_bounds__(THREADS)
pq_encode_kernel(const float* __restrict__ x, int64_t n, int d, int m, int ks, int dsub_rt,
                 const float* __restrict__ codebooks, const float* __restrict__ centroid_sq,
                 const float* __restrict__ coarse, const int32_t* __restrict__ assign,
                 uint8_t* __restrict__ codes, int layout, int B, int g, int u4, int kchunk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int dsub = DSUB > 0 ? DSUB : dsub_rt;
    constexpr bool kRes = (MODE == ENC_CSQ_RES || MODE == ENC_DOT_RES || MODE == ENC_DIRECT_RES);
    constexpr bool kCsq = (MODE == ENC_CSQ || MODE == ENC_CSQ_RES);

    uint8_t* s_codes = smem_raw;                                     // [T x m] bytes (AoS staging)
    const size_t codes_bytes = ((size_t)THREADS * m + 15) & ~(size_t)15;
    float* s_cb = reinterpret_cast<float*>(smem_raw + codes_bytes);  // [kchunk x dsub]
    float* s_csq = s_cb + (size_t)kchunk * dsub;                     // [kchunk]
    float* s_x = s_csq + kchunk;                                     // generic: [dsub x T] (transposed)
    float* s_g = s_x + (DSUB > 0 ? 0 : (size_t)dsub * THREADS);  // generic residual: [dsub x T]
    (void)s_g;

    const int t = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * THREADS;
    const int64_t i = i0 + t;
    const bool live = i < n;
    const float* xi = x + (live ? i : 0) * (int64_t)d;
    const float* gi = nullptr;
    if (kRes) gi = coarse + (int64_t)(live ? assign[i] : 0) * d;

    for (int j = 0; j < m; ++j) {
        float xr[DSUB > 0 ? DSUB : 1];
        float gr[DSUB > 0 ? DSUB : 1];
        if (DSUB > 0) {
#pragma unroll
            for (int e = 0; e < DSUB; ++e) xr[e] = live ? xi[(size_t)j * DSUB + e] : 0.0f;
            if (kRes) {
#pragma unroll
                for (int e = 0; e < DSUB; ++e) gr[e] = live ? gi[(size_t)j * DSUB + e] : 0.0f;
            }
        } else {
            for (int e = 0; e < dsub; ++e) s_x[(size_t)e * THREADS + t] = live ? xi[(size_t)j * dsub + e] : 0.0f;
            if (kRes)
                for (int e = 0; e < dsub; ++e) s_g[(size_t)e * THREADS + t] = live ? gi[(size_t)j * dsub + e] : 0.0f;
        }
        auto xf = [&](int e) -> float { return DSUB > 0 ? xr[e] : s_x[(size_t)e * THREADS + t]; };
        auto gf = [&](int e) -> float { return DSUB > 0 ? gr[e] : s_g[(size_t)e * THREADS + t]; };

        // base2: x2 (or r2), sequential sum, computed once per (vector, subspace)
        float base2 = 0.0f;
        if (MODE == ENC_CSQ || MODE == ENC_DOT) {
#pragma unroll
            for (int e = 0; e < dsub; ++e) base2 = fadd(base2, fmul(xf(e), xf(e)));
        } else if (MODE == ENC_CSQ_RES || MODE == ENC_DOT_RES) {
#pragma unroll
            for (int e = 0; e < dsub; ++e) {
                float ri = fsub(xf(e), gf(e));
                base2 = fadd(base2, fmul(ri, ri));
            }
        }

        float bd = 0.0f;
        int bk = 0;
        const float* cbj = codebooks + (size_t)j * ks * dsub;
        for (int k0 = 0; k0 < ks; k0 += kchunk) {
            int k1 = min(ks, k0 + kchunk);
            __syncthreads();   // previous chunk fully consumed
            const int cnt = (k1 - k0) * dsub;
            for (int e = t; e < cnt; e += THREADS) s_cb[e] = cbj[(size_t)k0 * dsub + e];
            if (kCsq)
                for (int e = t; e < k1 - k0; e += THREADS) s_csq[e] = centroid_sq[(size_t)j * ks + k0 + e];
            __syncthreads();
            encode_chunk<MODE, DSUB>(xf, gf, s_cb, s_csq, k0, k1, dsub, base2, bd, bk);
        }
        if (layout == PQ_LAYOUT_AOS) {
            s_codes[(size_t)t * m + j] = (uint8_t)bk;
        } else if (live) {
            codes[code_index(i, j, n, m, layout, B, g)] = (uint8_t)bk;
        }
    }

    if (layout != PQ_LAYOUT_AOS) return;
    __syncthreads();
    const int64_t rows = (n - i0 < THREADS) ? (n - i0) : (int64_t)THREADS;
    if (!u4) {
        // contiguous [rows x m] byte slab
        uint8_t* dst = codes + i0 * (int64_t)m;
        const int64_t total = rows * m;
        const bool al = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
        if (al) {
            const int64_t nv = total >> 4;
            const uint4* src4 = reinterpret_cast<const uint4*>(s_codes);
            uint4* dst4 = reinterpret_cast<uint4*>(dst);
            for (int64_t e = t; e < nv; e += THREADS) dst4[e] = src4[e];
            for (int64_t e = (nv << 4) + t; e < total; e += THREADS) dst[e] = s_codes[e];
        } else {
            for (int64_t e = t; e < total; e += THREADS) dst[e] = s_codes[e];
        }
    } else {
        // u4: two codes per byte, low nibble = even subspace (pq_encode.c:594-596)
        const int mb = m >> 1;
        uint8_t* dst = codes + i0 * (int64_t)mb;
        const int64_t total = rows * mb;
        for (int64_t e = t; e < total; e += THREADS) {
            int64_t r = e / mb;
            int b = (int)(e - r * mb);
            uint8_t c0 = s_codes[(size_t)r * m + 2 * b];
            uint8_t c1 = s_codes[(size_t)r * m + 2 * b + 1];
            dst[e] = (uint8_t)((c0 & 0x0F) | ((c1 & 0x0F) << 4));
        }
    }
}

template <int MODE, int DSUB, int THREADS>
static int launch_encode_rows(const float* x, int64_t n, int d, int m, int ks, int dsub, const float* cb,
                              const float* csq, const float* coarse, const int32_t* assign, uint8_t* codes,
                              int layout, int B, int g, int u4, bool& fits) {
    constexpr bool kRes = (MODE == ENC_CSQ_RES || MODE == ENC_DOT_RES || MODE == ENC_DIRECT_RES);
    const size_t fixed = (((size_t)THREADS * m + 15) & ~(size_t)15)                // code staging
                       + (DSUB > 0 ? 0 : (size_t)dsub * THREADS * 4 * (kRes ? 2 : 1));
    const size_t budget = 200 * 1024;
    fits = fixed + (size_t)(dsub + 1) * 4 <= budget;
    if (!fits) return VIX_OK;
    size_t kfit = (budget - fixed) / ((size_t)(dsub + 1) * 4);
    int kchunk = (int)(kfit < (size_t)ks ? kfit : (size_t)ks);
    if (kchunk < 1) kchunk = 1;
    size_t smem = (size_t)kchunk * (dsub + 1) * 4 + fixed + 16;
    auto kern = pq_encode_kernel<MODE, DSUB, THREADS>;
    VIX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = (n + THREADS - 1) / THREADS;
    kern<<<(unsigned)grid, THREADS, smem, ctx().stream>>>(x, n, d, m, ks, dsub, cb, csq, coarse, assign, codes,
                                                          layout, B, g, u4, kchunk);
    VIX_LAUNCH_CHECK();
    return VIX_OK;
}

template <int MODE, int DSUB>
static int launch_encode(const float* x, int64_t n, int d, int m, int ks, int dsub, const float* cb,
                         const float* csq, const float* coarse, const int32_t* assign, uint8_t* codes,
                         int layout, int B, int g, int u4) {
    bool fits = false;
    VIX_TRY((launch_encode_rows<MODE, DSUB, kEncThreads>(x, n, d, m, ks, dsub, cb, csq, coarse, assign, codes, layout, B, g,
                                                         u4, fits)));
    if (fits) return VIX_OK;
    if constexpr (DSUB == 0) {
        // long sub-vectors: a quarter of the rows per CTA (the staged [dsub x rows] tiles are what does not fit)
        VIX_TRY((launch_encode_rows<MODE, 0, 64>(x, n, d, m, ks, dsub, cb, csq, coarse, assign, codes, layout, B, g, u4,
                                                 fits)));
        if (fits) return VIX_OK;
    }
    set_error("pq_encode: d/m = %d with m = %d does not fit the shared-memory staging", dsub, m);
    return VIX_ERR_UNSUPPORTED;
}

template <int MODE>
static int dispatch_dsub(const float* x, int64_t n, int d, int m, int ks, int dsub, const float* cb,
                         const float* csq, const float* coarse, const int32_t* assign, uint8_t* codes,
                         int layout, int B, int g, int u4) {
#define VIX_ENC_CASE(D) \
    case D: return launch_encode<MODE, D>(x, n, d, m, ks, dsub, cb, csq, coarse, assign, codes, layout, B, g, u4)
    switch (dsub) {
        VIX_ENC_CASE(1);
        VIX_ENC_CASE(2);
        VIX_ENC_CASE(4);
        VIX_ENC_CASE(8);
        VIX_ENC_CASE(12);
        VIX_ENC_CASE(16);
        VIX_ENC_CASE(32);
        default: return launch_encode<MODE, 0>(x, n, d, m, ks, dsub, cb, csq, coarse, assign, codes, layout, B, g, u4);
    }
#undef VIX_ENC_CASE
}

bool pq_encode_tc_supported(const float* x, int64_t n, int d, int m, int ks, int mode, int layout, int u4);
int pq_encode_tc_device(const float* x, int64_t n, int d, int m, const float* cb, const float* csq, const float* coarse,
                        const int32_t* assign, uint8_t* codes, int mode);

// Device-pointer core shared by the cpq_* entry points and the index "add" path.
int pq_encode_device(const float* x, int64_t n, int d, int m, int ks, const float* cb, const float* csq,
                     const float* coarse, const int32_t* assign, uint8_t* codes, int use_dot, int layout,
                     int B, int g, int u4) {
    if (n == 0) return VIX_OK;
    const int dsub = d / m;
    const bool res = coarse != nullptr;
    int mode;
    if (csq) mode = res ? ENC_CSQ_RES : ENC_CSQ;
    else if (use_dot) mode = res ? ENC_DOT_RES : ENC_DOT;
    else mode = res ? ENC_DIRECT_RES : ENC_DIRECT;
    // tensor-core shortlist + exact finalists (vix_pq_tc.cu): same codes, bit for bit
    if (pq_encode_tc_supported(x, n, d, m, ks, mode, layout, u4))
        return pq_encode_tc_device(x, n, d, m, cb, csq, coarse, assign, codes, mode);
    switch (mode) {
        case ENC_CSQ: return dispatch_dsub<ENC_CSQ>(x, n, d, m, ks, dsub, cb, csq, coarse, assign, codes, layout, B, g, u4);
        case ENC_CSQ_RES: return dispatch_dsub<ENC_CSQ_RES>(x, n, d, m, ks, dsub, cb, csq, coarse, assign, codes, layout, B, g, u4);
        case ENC_DOT: return dispatch_dsub<ENC_DOT>(x, n, d, m, ks, dsub, cb, csq, coarse, assign, codes, layout, B, g, u4);
        case ENC_DIRECT: return dispatch_dsub<ENC_DIRECT>(x, n, d, m, ks, dsub, cb, csq, coarse, assign, codes, layout, B, g, u4);
        case ENC_DOT_RES: return dispatch_dsub<ENC_DOT_RES>(x, n, d, m, ks, dsub, cb, csq, coarse, assign, codes, layout, B, g, u4);
        default: return dispatch_dsub<ENC_DIRECT_RES>(x, n, d, m, ks, dsub, cb, csq, coarse, assign, codes, layout, B, g, u4);
    }
}

// Host-or-device pointer front end: validation in the reference's terms (its assert()s become a
// recorded error and an untouched output buffer), staging, launch, copy back.
static int encode_entry(const char* fn, const float* x, int64_t n, int d, int m, int ks, int ks_required,
                        const float* codebooks, const float* csq, bool need_csq, const float* coarse,
                        const int32_t* assign, bool residual, uint8_t* codes, const PQEncodeOpts* opts_in,
                        int u4) {
    VIX_TRY(ensure_device());
    VIX_REQUIRE(x && codebooks && codes, VIX_ERR_NULL_PTR, "%s: null pointer argument", fn);
    VIX_REQUIRE(!need_csq || csq, VIX_ERR_NULL_PTR, "%s: centroid_sq is null", fn);
    VIX_REQUIRE(!residual || (coarse && assign), VIX_ERR_NULL_PTR, "%s: coarse_centroids/assignments is null", fn);
    VIX_REQUIRE(n >= 0 && d > 0 && m > 0 && (d % m) == 0, VIX_ERR_INVALID_DIM,
                "%s: need n >= 0, d > 0, m > 0, d %% m == 0 (n=%lld d=%d m=%d)", fn, (long long)n, d, m);
    VIX_REQUIRE(ks == ks_required, VIX_ERR_INVALID_K, "%s: ks must be %d (got %d)", fn, ks_required, ks);
    VIX_REQUIRE(!u4 || (m & 1) == 0, VIX_ERR_INVALID_DIM, "%s: m must be even for u4 codes", fn);
    if (n == 0) return VIX_OK;

    // pq_opts_default (pq_encode.c:468-476)
    int layout = PQ_LAYOUT_AOS, B = 64, g = 8;
    int use_dot = (ks >= 64);
    if (opts_in) {
        layout = (int)opts_in->layout;
        use_dot = opts_in->use_dot_trick ? 1 : 0;
        if (opts_in->soa_block_B > 0) B = opts_in->soa_block_B;
        if (opts_in->interleave_g > 0) g = opts_in->interleave_g;
    }
    if (layout != PQ_LAYOUT_SOA_BLOCKED && layout != PQ_LAYOUT_INTERLEAVED_BLOCK) layout = PQ_LAYOUT_AOS;
    if (u4) { layout = PQ_LAYOUT_AOS; use_dot = 0; }   // u4: AoS enforced, direct L2 (pq_encode.c:571,590-591)

    const int dsub = d / m;
    size_t out_count;
    if (u4) out_count = (size_t)n * (m >> 1);
    else if (layout == PQ_LAYOUT_SOA_BLOCKED) out_count = (size_t)m * (size_t)((n + B - 1) / B) * B;
    else if (layout == PQ_LAYOUT_INTERLEAVED_BLOCK) out_count = (size_t)((n + g - 1) / g) * m * g;
    else out_count = (size_t)n * m;

    In<float> dx, dcb, dcsq, dco;
    In<int32_t> das;
    Out<uint8_t> dcodes;
    VIX_TRY(dx.stage(x, (size_t)n * d));
    VIX_TRY(dcb.stage(codebooks, (size_t)m * ks * dsub));
    if (need_csq) VIX_TRY(dcsq.stage(csq, (size_t)m * ks));
    int64_t kc_needed = 0;
    if (residual) {
        VIX_TRY(das.stage(assign, (size_t)n));
        if (!is_device_ptr(coarse)) {
            // host coarse table: its row count is implied by max(assignments)
            if (is_device_ptr(assign)) {
                set_error("%s: host coarse_centroids with device assignments is not supported", fn);
                return VIX_ERR_INVALID_PARAM;
            }
            int32_t mx = 0;
            for (int64_t i = 0; i < n; ++i) {
                VIX_REQUIRE(assign[i] >= 0, VIX_ERR_INVALID_PARAM, "%s: negative assignment at row %lld", fn, (long long)i);
                if (assign[i] > mx) mx = assign[i];
            }
            kc_needed = (int64_t)mx + 1;
        }
        VIX_TRY(dco.stage(coarse, (size_t)kc_needed * d));
    }
    VIX_TRY(dcodes.stage(codes, out_count));
    if (dcodes.is_host() && layout != PQ_LAYOUT_AOS)   // padded layouts: keep the caller's padding bytes
        VIX_CUDA(cudaMemcpyAsync(dcodes.dev, codes, out_count, cudaMemcpyHostToDevice, ctx().stream));

    VIX_TRY(pq_encode_device(dx.dev, n, d, m, ks, dcb.dev, need_csq ? dcsq.dev : nullptr,
                             residual ? dco.dev : nullptr, residual ? das.dev : nullptr, dcodes.dev, use_dot,
                             layout, B, g, u4));
    VIX_TRY(dcodes.commit());
    return finish(dcodes.is_host());
}

}  // namespace vix

using namespace vix;

extern "C" {

// /root/reference/Sources/CPQEncode/include/cpq_encode.h:41-50 ; pq_encode.c:479-520
void cpq_encode_u8_f32(const float* x, int64_t n, int d, int m, int ks, const float* codebooks,
                       uint8_t* codes, const PQEncodeOpts* opts) {
    (void)encode_entry("cpq_encode_u8_f32", x, n, d, m, ks, 256, codebooks, nullptr, false, nullptr, nullptr, false,
                       codes, opts, 0);
}

// cpq_encode.h:52-62 ; pq_encode.c:522-556
void cpq_encode_u8_f32_with_csq(const float* x, int64_t n, int d, int m, int ks, const float* codebooks,
                                const float* centroid_sq, uint8_t* codes, const PQEncodeOpts* opts) {
    (void)encode_entry("cpq_encode_u8_f32_with_csq", x, n, d, m, ks, 256, codebooks, centroid_sq, true, nullptr,
                       nullptr, false, codes, opts, 0);
}

// cpq_encode.h:64-72 ; pq_encode.c:558-599
void cpq_encode_u4_f32(const float* x, int64_t n, int d, int m, int ks, const float* codebooks,
                       uint8_t* codes, const PQEncodeOpts* opts) {
    (void)encode_entry("cpq_encode_u4_f32", x, n, d, m, ks, 16, codebooks, nullptr, false, nullptr, nullptr, false,
                       codes, opts, 1);
}

// cpq_encode.h:74-86 ; pq_encode.c:601-649
void cpq_encode_residual_u8_f32(const float* x, int64_t n, int d, int m, int ks, const float* codebooks,
                                const float* coarse_centroids, const int32_t* assignments, uint8_t* codes,
                                const PQEncodeOpts* opts) {
    (void)encode_entry("cpq_encode_residual_u8_f32", x, n, d, m, ks, 256, codebooks, nullptr, false,
                       coarse_centroids, assignments, true, codes, opts, 0);
}

// cpq_encode.h:88-100 ; pq_encode.c:651-690
void cpq_encode_residual_u8_f32_with_csq(const float* x, int64_t n, int d, int m, int ks,
                                         const float* codebooks, const float* centroid_sq,
                                         const float* coarse_centroids, const int32_t* assignments,
                                         uint8_t* codes, const PQEncodeOpts* opts) {
    (void)encode_entry("cpq_encode_residual_u8_f32_with_csq", x, n, d, m, ks, 256, codebooks, centroid_sq, true,
                       coarse_centroids, assignments, true, codes, opts, 0);
}

// cpq_encode.h:102-112 ; pq_encode.c:692-739
void cpq_encode_residual_u4_f32(const float* x, int64_t n, int d, int m, int ks, const float* codebooks,
                                const float* coarse_centroids, const int32_t* assignments, uint8_t* codes,
                                const PQEncodeOpts* opts) {
    (void)encode_entry("cpq_encode_residual_u4_f32", x, n, d, m, ks, 16, codebooks, nullptr, false,
                       coarse_centroids, assignments, true, codes, opts, 1);
}

// cpq_encode.h:114-122 ; pq_encode.c:441-466: ONE vector's m codes, host side
void cpq_pack_u4_bulk(const uint8_t* codes, int m, uint8_t* packed) {
    if (!codes || !packed || (m & 1)) { set_error("cpq_pack_u4_bulk: null pointer or odd m"); return; }
    for (int j = 0, o = 0; j < m; j += 2, ++o)
        packed[o] = (uint8_t)((codes[j] & 0x0F) | ((codes[j + 1] & 0x0F) << 4));
}

void cpq_unpack_u4_bulk(const uint8_t* packed, int m, uint8_t* codes) {
    if (!codes || !packed || (m & 1)) { set_error("cpq_unpack_u4_bulk: null pointer or odd m"); return; }
    for (int j = 0, o = 0; j < m; j += 2, ++o) {
        uint8_t b = packed[o];
        codes[j] = (uint8_t)(b & 0x0F);
        codes[j + 1] = (uint8_t)((b >> 4) & 0x0F);
    }
}

}  // extern "C"
