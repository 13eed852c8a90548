// vix_ivfpq_scan.cu -- the fused IVF-PQ list scan (query-only LUT -> ADC over the probed lists -> top-k).
//
// Reference composition: pq_lut_residual_l2_f32 -> adc_scan_u8 -> selectTopK -> mergeTopK
// (/root/reference/docs/kernel-specs/DONE_22_adc_scan.md:831-881; PQLUT.swift:266-386;
// ADCScan.swift:190-283; TopK.swift:54-176).  One persistent CTA serves one query at a time:
//
//   ||q - x^||^2 = ||q - c_l||^2 + (||r^||^2 + 2<c_l, r^>) - 2<q, r^>,   x^ = c_l + r^
//                  bias (per probe)  t_x (per stored vector)     sum_j T[j][code_j],  T = -2<q_j, cb_j[.]>
//
// so ONE table per query serves every probed list (the reference builds one residual LUT per
// (query, list)); the result agrees with the reference's sum to fp32 rounding (1e-5, tested).
//
// The scan is bound by shared-memory look-ups (one per code byte; an SM retires 32 per clock), so the
// kernel is organised to spend exactly one conflict-free LDS and two other instructions per byte:
//
//   * "lane = sub-quantiser": a warp owns 32 stored vectors at a time and lane l looks up
//     sub-quantiser 32 f + l of every one of them, accumulating 32 running sums (one per vector) in
//     registers.  The table is stored code-major, T[code][64 slots] (256 B per code), and lane l only
//     ever touches slot l (or l + 32): its bank is its lane id, so EVERY warp-wide look-up is a single
//     wavefront whatever the codes are.
//   * the byte offset code * 256 + 4 * lane is ONE byte-permute (PRMT) of the packed code word and a
//     per-lane constant; the slot half / table index is the LDS immediate.
//   * the 32 x 32 partial sums are transposed with a butterfly of shuffles (31 SHFL per 1024 look-ups)
//     so that lane v ends up with the distance of vector v.
//   * codes are stored per list in blocks of 32 vectors, transposed to [sub-quantiser][vector] (32 B
//     per sub-quantiser), so a lane's 32 codes are two 128-bit loads and a warp reads 1 KB
//     contiguously; the next pass is prefetched into registers while the current one is looked up.
//   * m = 32 F + R sub-quantisers (R in {0, 8, 16}): the R left-over sub-quantisers are handled by
//     splitting the warp into 32 / R lane groups that work on 32 / R different vector blocks at once
//     (their table slots are replicated so the bank == lane rule still holds).
//
// Top-k: per-warp shared-memory queues keyed (score, id) with a CTA-wide acceptance threshold, merged at
// the end of the query; no distance array ever reaches HBM.  Queries are handed out by an atomic
// counter in an order sorted by first probed list, so CTAs running concurrently scan neighbouring
// lists and share them through L2.
#include "vix_common.cuh"
#include "vix_topk.cuh"
#include "vix_scan.cuh"

namespace vix {

constexpr int kScanWarps = 8;
constexpr int kScanThreads = kScanWarps * 32;

template <int F_, int R_>
struct ScanShape {
    static constexpr int F = F_, R = R_;
    static constexpr int M = 32 * F + R;
    static constexpr int NG = (R == 0) ? 1 : 32 / R;             // vector blocks per chunk
    static constexpr int CH = 32 * NG;                           // slots per chunk
    static constexpr int NPASS = NG * F + (R ? 1 : 0);
    static constexpr int SLOTS = 32 * (F + (R ? 1 : 0));
    static constexpr int NTAB = (SLOTS + 63) / 64;               // 64 KB tables
    static_assert(R == 0 || R == 8 || R == 16, "m = 32 F + R with R in {0, 8, 16}");
    static_assert(R != 8 || F == 0, "R = 8 only for m = 8");
};

// byte offset of table slot s for the LDS immediate
__host__ __device__ constexpr int slot_imm(int s) { return (s >> 6) * 65536 + (s & 63) * 4; }

// acc[v] += T[code of vector v][slot column IMM] for the 32 vectors of one block
template <int IMM>
__device__ __forceinline__ void lookup32(const uint4& w0, const uint4& w1, const char* __restrict__ lut_b,
                                         uint32_t laneconst, float (&acc)[32]) {
    const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int v = 0; v < 32; ++v) {
        // bytes: [0] = 4 * lane, [1] = code, [2] = [3] = 0   ->   code * 256 + 4 * lane
        const uint32_t off = __byte_perm(w[v >> 2], laneconst, 0x6504 | ((v & 3) << 4));
        acc[v] += *reinterpret_cast<const float*>(lut_b + off + IMM);
    }
}

// Butterfly transpose-reduce over groups of GROUP lanes: afterwards acc[0 .. 32/GROUP) of lane l hold the
// group totals of vectors (32/GROUP) * (l % GROUP) + i.
template <int N, int O>
struct Butterfly {
    __device__ __forceinline__ static void run(float (&acc)[32], int lane) {
        const bool upper = (lane & O) != 0;
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            const float send = upper ? acc[i] : acc[i + N / 2];
            const float keep = upper ? acc[i + N / 2] : acc[i];
            acc[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, O);
        }
        Butterfly<N / 2, O / 2>::run(acc, lane);
    }
};
template <int N>
struct Butterfly<N, 0> {
    __device__ __forceinline__ static void run(float (&)[32], int) {}
};

struct ChunkPos {
    int p;            // probe index
    int within;       // first slot of the chunk inside the list
    int64_t g0;       // first slot of the chunk (global)
};

template <typename S>
__device__ __forceinline__ const uint4* pass_ptr(const uint8_t* __restrict__ codes, int64_t g0, int ps, int lane) {
    // pass ps < NG*F: block ps / F, sub-quantiser row 32 (ps % F) + lane; last pass: the R left-overs
    int blk, row;
    if (S::F > 0 && ps < S::NG * S::F) { blk = ps / (S::F > 0 ? S::F : 1); row = 32 * (ps % (S::F > 0 ? S::F : 1)) + lane; }
    else { blk = lane / (S::R ? S::R : 32); row = 32 * S::F + lane % (S::R ? S::R : 32); }
    return reinterpret_cast<const uint4*>(codes + ((g0 >> 5) + blk) * (int64_t)(32 * S::M) + (int64_t)row * 32);
}

template <int F, int R>
__global__ void __launch_bounds__(kScanThreads, (ScanShape<F, R>::NTAB == 1) ? 2 : 1)
ivfpq_scan_kernel(ScanArgs a) {
    using S = ScanShape<F, R>;
    constexpr int m = S::M;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_lut = reinterpret_cast<float*>(smem_raw);                    // NTAB x [256][64]
    float* s_q = s_lut + (size_t)S::NTAB * 16384;                         // [d]
    float* s_bias = s_q + a.d;                                            // [nprobe]
    int* s_start = reinterpret_cast<int*>(s_bias + a.nprobe);             // [nprobe]  first slot / 32
    int* s_len = s_start + a.nprobe;                                      // [nprobe]
    int* s_pref = s_len + a.nprobe;                                       // [nprobe + 1] chunk prefix
    int* s_misc = s_pref + a.nprobe + 1;                                  // [0] work item, [1] CTA threshold
    u64* s_wq = reinterpret_cast<u64*>((reinterpret_cast<uintptr_t>(s_misc + 2) + 15) & ~(uintptr_t)15);
    u64* s_merge = s_wq + (size_t)kScanWarps * a.Pw;                      // [P2]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int order_max = (a.metric == VIX_METRIC_IP);
    const float lut_scale = order_max ? 1.0f : -2.0f;
    u64* wq = s_wq + (size_t)warp * a.Pw;
    const char* lut_b = reinterpret_cast<const char*>(s_lut);
    const uint32_t laneconst = 4u * lane;
    unsigned long long scanned_local = 0;
    volatile uint32_t* cta_thr = reinterpret_cast<volatile uint32_t*>(s_misc + 1);

    for (;;) {
        __syncthreads();                                   // previous query fully drained
        if (tid == 0) s_misc[0] = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const int item = s_misc[0];
        if (item >= a.nq) break;
        const int64_t qi = a.order ? a.order[item] : item;

        // ---- prologue: query, probe table, bias, LUT ----
        for (int e = tid; e < a.d; e += kScanThreads) s_q[e] = a.queries[qi * (int64_t)a.d + e];
        if (tid < a.nprobe) {
            const int l = a.probes[qi * (int64_t)a.nprobe + tid];
            s_start[tid] = l >= 0 ? (int)(a.list_off[l] >> 5) : 0;
            s_len[tid] = l >= 0 ? a.list_len[l] : 0;
        }
        if (tid == 0) *cta_thr = 0xFFFFFFFFu;
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int p = 0; p < a.nprobe; ++p) { s_pref[p] = acc; acc += (s_len[p] + S::CH - 1) / S::CH; }
            s_pref[a.nprobe] = acc;
        }
        for (int p = warp; p < a.nprobe; p += kScanWarps) {
            const int l = a.probes[qi * (int64_t)a.nprobe + p];
            float part = 0.0f;
            if (l >= 0) {
                const float* c = a.coarse + (int64_t)l * a.d;
                if (order_max) for (int e = lane; e < a.d; e += 32) part = fmaf(s_q[e], c[e], part);
                else for (int e = lane; e < a.d; e += 32) { float df = s_q[e] - c[e]; part = fmaf(df, df, part); }
            }
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
            if (lane == 0) s_bias[p] = part;
        }
        {
            // T[c][slot(j)] = scale * <q_j, cb_j[c]>; codebooks_t is [256][m][dsub] so that consecutive
            // threads read consecutive memory and write consecutive banks
            const int dsub = a.dsub;
            for (int e = tid; e < m * 256; e += kScanThreads) {
                const int c = e / m, j = e - c * m;
                const float* cw = a.codebooks_t + (size_t)e * dsub;
                const float* qj = s_q + j * dsub;
                float dot = 0.0f;
                for (int t = 0; t < dsub; ++t) dot = fmaf(qj[t], __ldg(cw + t), dot);
                const float v = lut_scale * dot;
                if (j < 32 * F) {
                    s_lut[(j >> 6) * 16384 + c * 64 + (j & 63)] = v;
                } else {
#pragma unroll
                    for (int t = 0; t < S::NG; ++t) {
                        const int s = j + t * R;
                        s_lut[(s >> 6) * 16384 + c * 64 + (s & 63)] = v;
                    }
                }
            }
        }
        for (int i = lane; i < a.Pw; i += 32) wq[i] = kEmptyKey;
        __syncthreads();

        // ---- scan: warp-strided over chunks of CH slots ----
        const int nchunks = s_pref[a.nprobe];
        int cnt = 0;
        uint32_t thr_u = 0xFFFFFFFFu;
        int p = 0;
        int ch = warp;
        ChunkPos cur{0, 0, 0};
        uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
        if (ch < nchunks) {
            while (ch >= s_pref[p + 1]) ++p;
            cur.p = p; cur.within = (ch - s_pref[p]) * S::CH; cur.g0 = ((int64_t)s_start[p] << 5) + cur.within;
            const uint4* src = pass_ptr<S>(a.slot_codes, cur.g0, R > 0 ? S::NPASS - 1 : 0, lane);
            c0 = __ldg(src); c1 = __ldg(src + 1);
        }
        while (ch < nchunks) {
            // position of this warp's next chunk (for the prefetch of its first pass)
            const int chn = ch + kScanWarps;
            ChunkPos nxt = cur;
            if (chn < nchunks) {
                while (chn >= s_pref[p + 1]) ++p;
                nxt.p = p; nxt.within = (chn - s_pref[p]) * S::CH; nxt.g0 = ((int64_t)s_start[p] << 5) + nxt.within;
            }
            float res[S::NG];
            float rem[S::NG];
            float acc[32];
            int ps = 0;
            // pass order: the R left-overs first (all blocks of the chunk at once), then F passes per block
            if (R > 0) {
                // prefetch: the next pass of this chunk, or the first pass of the next chunk
                uint4 n0 = c0, n1 = c1;
                if (F > 0) { const uint4* s2 = pass_ptr<S>(a.slot_codes, cur.g0, 0, lane); n0 = __ldg(s2); n1 = __ldg(s2 + 1); }
                else if (chn < nchunks) { const uint4* s2 = pass_ptr<S>(a.slot_codes, nxt.g0, S::NPASS - 1, lane); n0 = __ldg(s2); n1 = __ldg(s2 + 1); }
#pragma unroll
                for (int v = 0; v < 32; ++v) acc[v] = 0.0f;
                lookup32<slot_imm(32 * F)>(c0, c1, lut_b, laneconst, acc);
                Butterfly<32, R / 2>::run(acc, lane);
#pragma unroll
                for (int i = 0; i < S::NG; ++i) rem[i] = acc[i];
                c0 = n0; c1 = n1;
            }
            if (F > 0) {
#pragma unroll
                for (int blk = 0; blk < S::NG; ++blk) {
#pragma unroll
                    for (int v = 0; v < 32; ++v) acc[v] = 0.0f;
#pragma unroll
                    for (int f = 0; f < F; ++f) {
                        ps = blk * F + f;
                        uint4 n0 = c0, n1 = c1;
                        if (ps + 1 < S::NG * F) {
                            const uint4* s2 = pass_ptr<S>(a.slot_codes, cur.g0, ps + 1, lane);
                            n0 = __ldg(s2); n1 = __ldg(s2 + 1);
                        } else if (chn < nchunks) {
                            const uint4* s2 = pass_ptr<S>(a.slot_codes, nxt.g0, R > 0 ? S::NPASS - 1 : 0, lane);
                            n0 = __ldg(s2); n1 = __ldg(s2 + 1);
                        }
                        // compile-time slot immediate per f
                        if (f == 0) lookup32<slot_imm(0)>(c0, c1, lut_b, laneconst, acc);
                        else if (f == 1) lookup32<slot_imm(32)>(c0, c1, lut_b, laneconst, acc);
                        else if (f == 2) lookup32<slot_imm(64)>(c0, c1, lut_b, laneconst, acc);
                        else lookup32<slot_imm(96)>(c0, c1, lut_b, laneconst, acc);
                        c0 = n0; c1 = n1;
                    }
                    Butterfly<32, 16>::run(acc, lane);
                    res[blk] = acc[0];
                }
                if (R > 0) {
                    // left-over sums live in lane blk * R + v / NG, entry v % NG
#pragma unroll
                    for (int blk = 0; blk < S::NG; ++blk) {
                        const int srcl = blk * R + lane / S::NG;
                        float pick = 0.0f;
#pragma unroll
                        for (int i = 0; i < S::NG; ++i) {
                            const float t = __shfl_sync(0xFFFFFFFFu, rem[i], srcl);
                            if ((lane % S::NG) == i) pick = t;
                        }
                        res[blk] += pick;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < S::NG; ++i) res[i] = rem[i];
            }

            // ---- candidates: result i of lane l is  (F > 0) block i, vector l
            //                                          (F == 0) block l / R, vector NG * (l % R) + i
            const float bias = s_bias[cur.p];
            const int len = s_len[cur.p];
            {
                const uint32_t t = *cta_thr;
                if (t < thr_u) thr_u = t;
            }
#pragma unroll
            for (int i = 0; i < S::NG; ++i) {
                const int vec = (F > 0) ? (32 * i + lane) : (32 * (lane / (R ? R : 32)) + S::NG * (lane % (R ? R : 32)) + i);
                const int within = cur.within + vec;
                const bool valid = within < len;
                const int64_t g = cur.g0 + vec;
                const float tx = valid ? __ldg(a.slot_tx + g) : 0.0f;
                const float sum = (bias + tx) + res[i];
                if (valid) ++scanned_local;
                const u64 key = make_key(sum, 0u, order_max);
                const bool pass = valid && ((uint32_t)(key >> 32) <= thr_u);
                const unsigned ball = __ballot_sync(0xFFFFFFFFu, pass);
                if (ball) {
                    if (pass) {
                        const uint32_t id = (uint32_t)a.slot_ids[g];
                        wq[a.k + cnt + __popc(ball & ((1u << lane) - 1u))] = key | (u64)id;
                    }
                    cnt += __popc(ball);
                    __syncwarp();
                    if (cnt + 32 > a.Pw - a.k) {
                        for (int t = a.k + cnt + lane; t < a.Pw; t += 32) wq[t] = kEmptyKey;
                        __syncwarp();
                        bitonic_sort_keys<true>(wq, a.Pw, lane, 32);
                        cnt = 0;
                        const u64 t = wq[a.k - 1];
                        if (t != kEmptyKey) {
                            const uint32_t tu = (uint32_t)(t >> 32);
                            if (tu < thr_u) thr_u = tu;
                            if (lane == 0) atomicMin(reinterpret_cast<unsigned int*>(s_misc + 1), tu);
                        }
                    }
                }
            }
            cur = nxt;
            ch = chn;
        }
        // ---- epilogue: flush warp queues, merge, write ----
        if (cnt > 0) {
            for (int i = a.k + cnt + lane; i < a.Pw; i += 32) wq[i] = kEmptyKey;
            __syncwarp();
            bitonic_sort_keys<true>(wq, a.Pw, lane, 32);
        }
        __syncwarp();
        for (int i = lane; i < a.k; i += 32) s_merge[warp * a.k + i] = wq[i];
        for (int i = kScanWarps * a.k + tid; i < a.P2; i += kScanThreads) s_merge[i] = kEmptyKey;
        __syncthreads();
        bitonic_sort_keys<false>(s_merge, a.P2, tid, kScanThreads);
        for (int i = tid; i < a.k; i += kScanThreads) {
            const u64 key = s_merge[i];
            const size_t o = (size_t)qi * a.k + i;
            if (key == kEmptyKey) { a.out_dist[o] = __int_as_float(0x7fc00000); a.out_ids[o] = -1; }
            else {
                const float sc = key_score(key, order_max);
                a.out_dist[o] = order_max ? -sc : sc;     // IP: API distance = -score (DistanceUtils.swift:40-46)
                a.out_ids[o] = (int64_t)key_id(key);
            }
        }
    }
    if (a.scanned) {
        for (int o = 16; o > 0; o >>= 1) scanned_local += __shfl_xor_sync(0xFFFFFFFFu, scanned_local, o);
        if (lane == 0 && scanned_local) atomicAdd(a.scanned, scanned_local);
    }
}

// ------------------------------------------------------------------------------------------------
// generic m (not 32 F + {0, 8, 16}): plain AoS codes, [m][256] table, one stored vector per thread
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads)
ivfpq_scan_generic_kernel(ScanArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = a.m;
    float* s_lut = reinterpret_cast<float*>(smem_raw);          // [m][256]
    float* s_q = s_lut + (size_t)m * 256;
    float* s_bias = s_q + a.d;
    int* s_start = reinterpret_cast<int*>(s_bias + a.nprobe);
    int* s_len = s_start + a.nprobe;
    int* s_pref = s_len + a.nprobe;
    u64* s_wq = reinterpret_cast<u64*>((reinterpret_cast<uintptr_t>(s_pref + a.nprobe + 1) + 15) & ~(uintptr_t)15);
    u64* s_merge = s_wq + (size_t)kScanWarps * a.Pw;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int order_max = (a.metric == VIX_METRIC_IP);
    const float lut_scale = order_max ? 1.0f : -2.0f;
    u64* wq = s_wq + (size_t)warp * a.Pw;
    unsigned long long scanned_local = 0;

    for (int64_t qi = blockIdx.x; qi < a.nq; qi += gridDim.x) {
        __syncthreads();
        for (int e = tid; e < a.d; e += kScanThreads) s_q[e] = a.queries[qi * (int64_t)a.d + e];
        if (tid < a.nprobe) {
            const int l = a.probes[qi * (int64_t)a.nprobe + tid];
            s_start[tid] = l >= 0 ? (int)(a.list_off[l] >> 5) : 0;
            s_len[tid] = l >= 0 ? a.list_len[l] : 0;
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int p = 0; p < a.nprobe; ++p) { s_pref[p] = acc; acc += (s_len[p] + 31) >> 5; }
            s_pref[a.nprobe] = acc;
        }
        for (int p = warp; p < a.nprobe; p += kScanWarps) {
            const int l = a.probes[qi * (int64_t)a.nprobe + p];
            float part = 0.0f;
            if (l >= 0) {
                const float* c = a.coarse + (int64_t)l * a.d;
                if (order_max) for (int e = lane; e < a.d; e += 32) part = fmaf(s_q[e], c[e], part);
                else for (int e = lane; e < a.d; e += 32) { float df = s_q[e] - c[e]; part = fmaf(df, df, part); }
            }
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
            if (lane == 0) s_bias[p] = part;
        }
        for (int e = tid; e < m * 256; e += kScanThreads) {
            const int j = e >> 8, c = e & 255;
            const float* cw = a.codebooks + (size_t)e * a.dsub;
            const float* qj = s_q + j * a.dsub;
            float dot = 0.0f;
            for (int t = 0; t < a.dsub; ++t) dot = fmaf(qj[t], __ldg(cw + t), dot);
            s_lut[e] = lut_scale * dot;
        }
        for (int i = lane; i < a.Pw; i += 32) wq[i] = kEmptyKey;
        __syncthreads();

        const int nchunks = s_pref[a.nprobe];
        int cnt = 0;
        float thr_s = order_max ? -INFINITY : INFINITY;
        int p = 0;
        for (int ch = warp; ch < nchunks; ch += kScanWarps) {
            while (ch >= s_pref[p + 1]) ++p;
            const int within = (ch - s_pref[p]) * 32 + lane;
            const bool valid = within < s_len[p];
            const int64_t g = ((int64_t)s_start[p] << 5) + within;
            const uint8_t* src = a.slot_codes + g * (int64_t)m;
            float s0 = 0.f;
            if (valid) for (int j = 0; j < m; ++j) s0 += s_lut[(size_t)j * 256 + src[j]];
            const float tx = valid ? a.slot_tx[g] : 0.0f;
            const float sum = (s_bias[p] + tx) + s0;
            if (valid) ++scanned_local;
            const bool pass = valid && (order_max ? !(sum < thr_s) : !(sum > thr_s));
            const unsigned ball = __ballot_sync(0xFFFFFFFFu, pass);
            if (ball) {
                if (pass) {
                    const uint32_t id = (uint32_t)a.slot_ids[g];
                    wq[a.k + cnt + __popc(ball & ((1u << lane) - 1u))] = make_key(sum, id, order_max);
                }
                cnt += __popc(ball);
                __syncwarp();
                if (cnt + 32 > a.Pw - a.k) {
                    for (int i = a.k + cnt + lane; i < a.Pw; i += 32) wq[i] = kEmptyKey;
                    __syncwarp();
                    bitonic_sort_keys<true>(wq, a.Pw, lane, 32);
                    cnt = 0;
                    const u64 t = wq[a.k - 1];
                    if (t != kEmptyKey) thr_s = key_score(t, order_max);
                }
            }
        }
        if (cnt > 0) {
            for (int i = a.k + cnt + lane; i < a.Pw; i += 32) wq[i] = kEmptyKey;
            __syncwarp();
            bitonic_sort_keys<true>(wq, a.Pw, lane, 32);
        }
        __syncwarp();
        for (int i = lane; i < a.k; i += 32) s_merge[warp * a.k + i] = wq[i];
        for (int i = kScanWarps * a.k + tid; i < a.P2; i += kScanThreads) s_merge[i] = kEmptyKey;
        __syncthreads();
        bitonic_sort_keys<false>(s_merge, a.P2, tid, kScanThreads);
        for (int i = tid; i < a.k; i += kScanThreads) {
            const u64 key = s_merge[i];
            const size_t o = (size_t)qi * a.k + i;
            if (key == kEmptyKey) { a.out_dist[o] = __int_as_float(0x7fc00000); a.out_ids[o] = -1; }
            else {
                const float sc = key_score(key, order_max);
                a.out_dist[o] = order_max ? -sc : sc;
                a.out_ids[o] = (int64_t)key_id(key);
            }
        }
    }
    if (a.scanned) {
        for (int o = 16; o > 0; o >>= 1) scanned_local += __shfl_xor_sync(0xFFFFFFFFu, scanned_local, o);
        if (lane == 0 && scanned_local) atomicAdd(a.scanned, scanned_local);
    }
}

// ------------------------------------------------------------------------------------------------
// layout + launch
// ------------------------------------------------------------------------------------------------
ScanLayout scan_layout(int m) {
    ScanLayout L;
    const int F = m / 32, R = m % 32;
    const bool fast = m > 0 && m <= 128 && (R == 0 || R == 16 || (R == 8 && F == 0));
    L.fast = fast;
    L.ng = fast ? (R == 0 ? 1 : 32 / R) : 1;
    L.align = 32 * L.ng;
    return L;
}

static size_t scan_smem_bytes(const ScanArgs& a, const ScanLayout& L) {
    size_t s;
    if (L.fast) {
        const int slots = 32 * (a.m / 32 + ((a.m % 32) ? 1 : 0));
        s = (size_t)((slots + 63) / 64) * 65536;
    } else {
        s = (size_t)a.m * 256 * 4;
    }
    s += (size_t)a.d * 4 + (size_t)a.nprobe * 12 + (size_t)(a.nprobe + 1) * 4 + 8 + 16;
    s += (size_t)kScanWarps * a.Pw * 8 + (size_t)a.P2 * 8;
    return s;
}

template <int F, int R>
static int launch_fast(ScanArgs& a, size_t smem) {
    auto kern = ivfpq_scan_kernel<F, R>;
    VIX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    VIX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kScanThreads, smem));
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)num_sms() * occ;
    if (grid > a.nq) grid = a.nq;
    kern<<<(unsigned)grid, kScanThreads, smem, ctx().stream>>>(a);
    VIX_LAUNCH_CHECK();
    return VIX_OK;
}

int launch_ivfpq_scan(ScanArgs& a) {
    a.Pw = next_pow2(a.k + 32);
    a.P2 = next_pow2(kScanWarps * a.k);
    const ScanLayout L = scan_layout(a.m);
    const size_t smem = scan_smem_bytes(a, L);
    VIX_REQUIRE(smem <= 227 * 1024, VIX_ERR_UNSUPPORTED,
                "ivfpq scan: m = %d, k = %d, nprobe = %d need %zu bytes of shared memory", a.m, a.k, a.nprobe, smem);
    VIX_REQUIRE(a.nprobe <= kScanThreads, VIX_ERR_INVALID_K, "ivfpq scan: nprobe > %d", kScanThreads);
    if (!L.fast) {
        VIX_CUDA(cudaFuncSetAttribute(ivfpq_scan_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 1;
        VIX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ivfpq_scan_generic_kernel, kScanThreads, smem));
        int64_t grid = (int64_t)num_sms() * (occ < 1 ? 1 : occ);
        if (grid > a.nq) grid = a.nq;
        ivfpq_scan_generic_kernel<<<(unsigned)grid, kScanThreads, smem, ctx().stream>>>(a);
        VIX_LAUNCH_CHECK();
        return VIX_OK;
    }
    VIX_REQUIRE(a.work_counter != nullptr, VIX_ERR_NULL_PTR, "ivfpq scan: work counter missing");
    VIX_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(int), ctx().stream));
    switch (a.m) {
        case 8:   return launch_fast<0, 8>(a, smem);
        case 16:  return launch_fast<0, 16>(a, smem);
        case 32:  return launch_fast<1, 0>(a, smem);
        case 48:  return launch_fast<1, 16>(a, smem);
        case 64:  return launch_fast<2, 0>(a, smem);
        case 80:  return launch_fast<2, 16>(a, smem);
        case 96:  return launch_fast<3, 0>(a, smem);
        case 112: return launch_fast<3, 16>(a, smem);
        case 128: return launch_fast<4, 0>(a, smem);
    }
    set_error("ivfpq scan: unsupported m = %d", a.m);
    return VIX_ERR_UNSUPPORTED;
}

}  // namespace vix
