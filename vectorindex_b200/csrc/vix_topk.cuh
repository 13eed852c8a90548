// vix_topk.cuh -- fused top-k selection in shared memory (device side).
//
// A selection queue is a shared-memory array of P 64-bit keys (P a power of two):
//   keys[0, k)      the current k best candidates, ascending (kEmptyKey padded)
//   keys[k, k+cnt)  unsorted candidates accepted since the last flush
// A candidate is accepted only if key < thr (thr == keys[k-1] after the last flush), so once the
// threshold has tightened almost nothing is written.  A flush pads the tail with kEmptyKey and
// bitonic-sorts all P keys; (score, id) is a total order (Operations/Selection/TopK.swift:8-31 of the
// reference), so the result equals the reference's heap selection whatever the arrival order.
#pragma once

#include "vix_common.cuh"

namespace vix {

__host__ __device__ inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Bitonic sort of P keys in shared memory by `nthreads` cooperating threads (thread index `t`).
// WARP == true: the cooperating threads are one warp (sync with __syncwarp);
// otherwise they are the whole CTA (sync with __syncthreads) and every thread of the CTA must call.
template <bool WARP>
__device__ __forceinline__ void bitonic_sort_keys(u64* s, int P, int t, int nthreads) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = t; i < (P >> 1); i += nthreads) {
                int lo = 2 * i - (i & (stride - 1));
                int hi = lo + stride;
                bool asc = ((lo & size) == 0);
                u64 a = s[lo], b = s[hi];
                if ((a > b) == asc) { s[lo] = b; s[hi] = a; }
            }
            if (WARP) __syncwarp(); else __syncthreads();
        }
    }
}

// Warp-owned queue (one query row per queue, used by the tiled pair kernel).
struct WarpQueue {
    u64* keys;     // [P]
    int* cnt;      // candidates since last flush
    u64* thr;      // acceptance threshold
    int k, P;

    __device__ __forceinline__ void init(int lane) {
        for (int i = lane; i < P; i += 32) keys[i] = kEmptyKey;
        if (lane == 0) { *cnt = 0; *thr = kEmptyKey; }
    }
    // any thread of the CTA may push (capacity is guaranteed by the caller's flush policy)
    __device__ __forceinline__ void push(u64 key) {
        if (key < *thr) {
            int pos = atomicAdd(cnt, 1);
            keys[k + pos] = key;
        }
    }
    __device__ __forceinline__ void flush(int lane) {
        int c = *cnt;
        __syncwarp();
        for (int i = k + c + lane; i < P; i += 32) keys[i] = kEmptyKey;
        __syncwarp();
        bitonic_sort_keys<true>(keys, P, lane, 32);
        if (lane == 0) { *cnt = 0; *thr = keys[k - 1]; }
        __syncwarp();
    }
};

// CTA-owned queue (one query per CTA: IVF-PQ scan, row selection, merges).
struct BlockQueue {
    u64* keys;
    int* cnt;
    u64* thr;
    int k, P;

    __device__ __forceinline__ void init() {
        for (int i = threadIdx.x; i < P; i += blockDim.x) keys[i] = kEmptyKey;
        if (threadIdx.x == 0) { *cnt = 0; *thr = kEmptyKey; }
        __syncthreads();
    }
    __device__ __forceinline__ void push(u64 key) {
        if (key < *thr) {
            int pos = atomicAdd(cnt, 1);
            keys[k + pos] = key;
        }
    }
    // every thread of the CTA must call; `need` = number of pushes the next round may add
    __device__ __forceinline__ void flush_if_needed(int need) {
        __syncthreads();
        if (*cnt + need > P - k) flush();
    }
    __device__ __forceinline__ void flush() {
        __syncthreads();
        int c = *cnt;
        __syncthreads();
        for (int i = k + c + threadIdx.x; i < P; i += blockDim.x) keys[i] = kEmptyKey;
        __syncthreads();
        bitonic_sort_keys<false>(keys, P, threadIdx.x, blockDim.x);
        if (threadIdx.x == 0) { *cnt = 0; *thr = keys[k - 1]; }
        __syncthreads();
    }
};

}  // namespace vix
