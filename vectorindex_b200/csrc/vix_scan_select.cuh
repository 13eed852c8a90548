// vix_scan_select.cuh -- device pieces shared by the two IVF-PQ scan paths (vix_ivfpq_scan.cu: look-up tables on the
// CUDA cores; vix_ivfpq_tc.cu: tensor-core shortlist + the same arithmetic for the finalists).
#pragma once

#include "vix_common.cuh"
#include "vix_topk.cuh"

#ifndef VIX_SCAN_FN
#define VIX_SCAN_FN __noinline__
#endif

namespace vix {

// one entry of the query-only table at dsub = 2: scale * <q_j, cb_j[c]> with the query pre-scaled (q0s = q[2j] * scale,
// q1s = q[2j + 1] * scale).  build_lut and the finalist evaluation of the tensor-core path both call THIS function, so the
// two paths produce the same bits.
__device__ __forceinline__ float lut_entry2(float q0s, float q1s, float2 v) { return fmaf(q1s, v.y, q0s * v.x); }

// k-th smallest (0-based rank kth) of one 32-bit key per lane: bitonic network over the warp
static __device__ VIX_SCAN_FN uint32_t warp_kth_smallest(uint32_t v, int kth, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, v, stride);
            const bool up = (lane & size) == 0 || size == 32;
            const bool lower = (lane & stride) == 0;
            v = (lower == up) ? min(v, other) : max(v, other);
        }
    }
    return __shfl_sync(0xFFFFFFFFu, v, kth);
}

// ---- per-query prologue / epilogue pieces, kept out of line so that the scan loop owns the registers ----

// warp 0: the k best of the n published candidates.  Keys are unique, so "the smallest key greater than the
// last one taken" walks them in order.  Up to 512 candidates live in registers (16 per lane); each of the k
// rounds is a lane-local minimum over the registers plus two warp REDUX steps.
static __device__ __forceinline__ void write_result(u64 mn, int order_max, size_t o, float* __restrict__ out_dist,
                                             int64_t* __restrict__ out_ids) {
    if (mn == kEmptyKey) { out_dist[o] = __int_as_float(0x7fc00000); out_ids[o] = -1; }
    else {
        const float sc = key_score(mn, order_max);
        out_dist[o] = order_max ? -sc : sc;   // IP: API distance = -score (DistanceUtils.swift:40-46)
        out_ids[o] = (int64_t)key_id(mn);
    }
}

static __device__ VIX_SCAN_FN void select_and_write(const u64* __restrict__ s_cand, int n, int k, int order_max, int64_t qi,
                                             float* __restrict__ out_dist, int64_t* __restrict__ out_ids) {
    const int lane = threadIdx.x & 31;
    if (n <= 512) {
        constexpr int R = 16;
        uint32_t kh[R], kl[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int t = lane + 32 * r;
            const u64 key = (t < n) ? s_cand[t] : kEmptyKey;
            kh[r] = (uint32_t)(key >> 32); kl[r] = (uint32_t)key;
        }
        for (int i = 0; i < k; ++i) {
            uint32_t mh = 0xFFFFFFFFu, ml = 0xFFFFFFFFu;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool less = kh[r] < mh || (kh[r] == mh && kl[r] < ml);
                mh = less ? kh[r] : mh; ml = less ? kl[r] : ml;
            }
            const uint32_t gh = __reduce_min_sync(0xFFFFFFFFu, mh);
            const uint32_t gl = __reduce_min_sync(0xFFFFFFFFu, mh == gh ? ml : 0xFFFFFFFFu);
            // ONE owner retires ONE copy of the key: the same (score, id) pair may have been stored more than once
            // (an id added twice), and every copy is a result of its own
            int mine = -1;
#pragma unroll
            for (int r = R - 1; r >= 0; --r)
                if (kh[r] == gh && kl[r] == gl) mine = r;
            const unsigned owners = __ballot_sync(0xFFFFFFFFu, mine >= 0 && (gh & gl) != 0xFFFFFFFFu);
            if (owners && lane == __ffs(owners) - 1) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (r == mine) { kh[r] = 0xFFFFFFFFu; kl[r] = 0xFFFFFFFFu; }
            }
            if (lane == 0) write_result(((u64)gh << 32) | gl, order_max, (size_t)qi * k + i, out_dist, out_ids);
        }
        return;
    }
    // more candidates than the registers hold (rare): "the smallest key greater than the last one taken" walks the
    // keys in order; a key stored several times (an id added twice) is taken once per copy
    uint32_t last_hi = 0, last_lo = 0;
    int taken = 0;                                          // copies of `last` written so far (0: nothing taken yet)
    for (int i = 0; i < k; ++i) {
        uint32_t mh = 0xFFFFFFFFu, ml = 0xFFFFFFFFu;
        int same = 0;
        for (int t = lane; t < n; t += 32) {
            const u64 key = s_cand[t];
            const uint32_t kh = (uint32_t)(key >> 32), kl = (uint32_t)key;
            same += (taken > 0 && kh == last_hi && kl == last_lo) ? 1 : 0;
            const bool after = taken == 0 || kh > last_hi || (kh == last_hi && kl > last_lo);
            if (after && (kh < mh || (kh == mh && kl < ml))) { mh = kh; ml = kl; }
        }
        same = __reduce_add_sync(0xFFFFFFFFu, same);
        uint32_t gh, gl;
        if (same > taken) { gh = last_hi; gl = last_lo; taken += 1; }
        else {
            gh = __reduce_min_sync(0xFFFFFFFFu, mh);
            gl = __reduce_min_sync(0xFFFFFFFFu, mh == gh ? ml : 0xFFFFFFFFu);
            last_hi = gh; last_lo = gl; taken = 1;
        }
        if (lane == 0) write_result(((u64)gh << 32) | gl, order_max, (size_t)qi * k + i, out_dist, out_ids);
    }
}

}  // namespace vix
