// vix_exact.cuh -- the reference's fp32 reduction orders as compile-time "specs", and two engines
// that evaluate them bit-exactly on the GPU:
//
//   exact_pair<Spec>(a, b, d)            one (a, b) pair by one thread (pointers to global/shared)
//   PairTile<Spec, TV, TC>               a register-tiled [16*TV x 16*TC] block of pairs per CTA
//
// Every reduction the reference uses on this path has the same shape: element
// t < dmain = (d / L) * L feeds SIMD lane (t mod L); the L lane sums are combined by a two-level,
// left-associated tree (NOUT outer terms, each the left-associated sum of NIN lanes); the scalar tail
// t >= dmain is then added sequentially.  A lane sum is itself a sequential chain, so the engines run
// ONE lane at a time (a single accumulator per pair), in tree order -- the same operations in the
// same order as the Swift SIMD4 code, with unfused multiply/add.
//
//   spec        L  NOUT NIN lane(o,i)  reference function
//   Km12        8   2   4   4o+i      _vi_km12_l2sq_aos      KMeansMiniBatchKernel.swift:198-225
//   Lut8        8   2   4   4o+i      _simd_l2sqr/_simd_dot  PQLUT.swift:69-140  (same tree as Km12)
//   Km11        8   4   2   o+4i      km11 update / PQTrain l2Sq   KMeansSeeding.swift:302-361, PQTrain.swift:797-813
//   Direct16   16   4   4   o+4i      _l2sqr_single_direct   L2SqrKernel.swift:192-238
//   Dot16      16   4   4   o+4i      _l2sqr_block_dot_fused_serial dot   L2SqrKernel.swift:411-448
//   Ip4         4   4   1   o         InnerProduct.generic / ip_r1_D   InnerProduct.swift:115-184
//   Seq         1   1   1   0         sequential sum (pq_encode.c dot_only; netlib-order sgemm of the oracle)
#pragma once

#include "vix_common.cuh"

namespace vix {

enum PairOp { OP_DIFFSQ = 0, OP_PROD = 1 };

template <int OP_, int L_, int NOUT_, int NIN_, bool STRIDED_>
struct ReduceSpec {
    static constexpr int OP = OP_;
    static constexpr int L = L_;
    static constexpr int NOUT = NOUT_;
    static constexpr int NIN = NIN_;
    static constexpr bool STRIDED = STRIDED_;
    static_assert(NOUT_ * NIN_ == L_, "lanes = NOUT * NIN");
    __host__ __device__ static constexpr int lane(int o, int i) { return STRIDED_ ? o + i * NOUT_ : o * NIN_ + i; }
};

using SpecKm12L2 = ReduceSpec<OP_DIFFSQ, 8, 2, 4, false>;
using SpecLut8L2 = ReduceSpec<OP_DIFFSQ, 8, 2, 4, false>;
using SpecLut8Dot = ReduceSpec<OP_PROD, 8, 2, 4, false>;
using SpecKm11L2 = ReduceSpec<OP_DIFFSQ, 8, 4, 2, true>;
using SpecDirect16L2 = ReduceSpec<OP_DIFFSQ, 16, 4, 4, true>;
using SpecDot16 = ReduceSpec<OP_PROD, 16, 4, 4, true>;
using SpecIp4 = ReduceSpec<OP_PROD, 4, 4, 1, false>;
using SpecSeqDot = ReduceSpec<OP_PROD, 1, 1, 1, false>;
using SpecSeqL2 = ReduceSpec<OP_DIFFSQ, 1, 1, 1, false>;

template <int OP>
__device__ __forceinline__ float pair_term(float a, float b) {
    if (OP == OP_DIFFSQ) {
        float df = fsub(a, b);
        return fmul(df, df);
    }
    return fmul(a, b);
}

// One pair, one thread.  a/b may point to global or shared memory; stride_a/stride_b are element
// strides (1 for contiguous rows).
template <typename Spec>
__device__ __forceinline__ float exact_pair(const float* __restrict__ a, const float* __restrict__ b, int d,
                                            int stride_a = 1, int stride_b = 1) {
    const int nstr = d / Spec::L;
    const int dmain = nstr * Spec::L;
    float outer = 0.0f;
#pragma unroll
    for (int o = 0; o < Spec::NOUT; ++o) {
        float inner = 0.0f;
#pragma unroll
        for (int i = 0; i < Spec::NIN; ++i) {
            const int lane = Spec::lane(o, i);
            float acc = 0.0f;
            for (int s = 0; s < nstr; ++s) {
                const int t = s * Spec::L + lane;
                acc = fadd(acc, pair_term<Spec::OP>(a[(size_t)t * stride_a], b[(size_t)t * stride_b]));
            }
            inner = (i == 0) ? acc : fadd(inner, acc);
        }
        outer = (o == 0) ? inner : fadd(outer, inner);
    }
    for (int t = dmain; t < d; ++t)
        outer = fadd(outer, pair_term<Spec::OP>(a[(size_t)t * stride_a], b[(size_t)t * stride_b]));
    return outer;
}

// Norms.l2NormSquared (Operations/Support/Norms.swift:105-130): 16-stride four accumulators ->
// hsum(((a0+a1)+a2)+a3) ; then 4-groups sum += hsum(v*v) ; then scalar tail.
__device__ __forceinline__ float exact_norm_l2sq(const float* __restrict__ x, int d) {
    if (d == 0) return 0.0f;
    const int d16 = d & ~15;
    float outer = 0.0f;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        float inner = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int lane = o + 4 * i;
            float acc = 0.0f;
            for (int t = lane; t < d16; t += 16) acc = fadd(acc, fmul(x[t], x[t]));
            inner = (i == 0) ? acc : fadd(inner, acc);
        }
        outer = (o == 0) ? inner : fadd(outer, inner);
    }
    float sum = outer;
    const int d4 = d & ~3;
    int j = d16;
    for (; j < d4; j += 4) {
        float p = hsum4(fmul(x[j], x[j]), fmul(x[j + 1], x[j + 1]), fmul(x[j + 2], x[j + 2]), fmul(x[j + 3], x[j + 3]));
        sum = fadd(sum, p);
    }
    for (; j < d; ++j) sum = fadd(sum, fmul(x[j], x[j]));
    return sum;
}

// Register-tiled block of exact pairs.  The CTA has 256 threads laid out 16 (A rows) x 16 (B rows);
// thread (tx, ty) owns A rows {tx + 16 v} and B rows {ty + 16 c}.  As/Bs are row-major shared-memory
// tiles with an ODD pitch (d | 1) so that the 16 different rows a warp touches fall in 16 different
// banks (and the two B rows of a warp are broadcasts).
template <typename Spec, int TV, int TC>
struct PairTile {
    static constexpr int TA = 16 * TV;
    static constexpr int TB = 16 * TC;

    __device__ __forceinline__ static void compute(const float* __restrict__ As, const float* __restrict__ Bs,
                                                   int pitch, int d, int tx, int ty, float (&out)[TV][TC]) {
        const int nstr = d / Spec::L;
        const int dmain = nstr * Spec::L;
        const float* ap[TV];
        const float* bp[TC];
#pragma unroll
        for (int v = 0; v < TV; ++v) ap[v] = As + (size_t)(tx + 16 * v) * pitch;
#pragma unroll
        for (int c = 0; c < TC; ++c) bp[c] = Bs + (size_t)(ty + 16 * c) * pitch;

#pragma unroll
        for (int o = 0; o < Spec::NOUT; ++o) {
            float inner[TV][TC];
#pragma unroll
            for (int i = 0; i < Spec::NIN; ++i) {
                const int lane = Spec::lane(o, i);
                float acc[TV][TC];
#pragma unroll
                for (int v = 0; v < TV; ++v)
#pragma unroll
                    for (int c = 0; c < TC; ++c) acc[v][c] = 0.0f;
#pragma unroll 2
                for (int s = 0; s < nstr; ++s) {
                    const int t = s * Spec::L + lane;
                    float a[TV], b[TC];
#pragma unroll
                    for (int v = 0; v < TV; ++v) a[v] = ap[v][t];
#pragma unroll
                    for (int c = 0; c < TC; ++c) b[c] = bp[c][t];
#pragma unroll
                    for (int v = 0; v < TV; ++v)
#pragma unroll
                        for (int c = 0; c < TC; ++c) acc[v][c] = fadd(acc[v][c], pair_term<Spec::OP>(a[v], b[c]));
                }
#pragma unroll
                for (int v = 0; v < TV; ++v)
#pragma unroll
                    for (int c = 0; c < TC; ++c) inner[v][c] = (i == 0) ? acc[v][c] : fadd(inner[v][c], acc[v][c]);
            }
#pragma unroll
            for (int v = 0; v < TV; ++v)
#pragma unroll
                for (int c = 0; c < TC; ++c) out[v][c] = (o == 0) ? inner[v][c] : fadd(out[v][c], inner[v][c]);
        }
        for (int t = dmain; t < d; ++t) {
            float a[TV], b[TC];
#pragma unroll
            for (int v = 0; v < TV; ++v) a[v] = ap[v][t];
#pragma unroll
            for (int c = 0; c < TC; ++c) b[c] = bp[c][t];
#pragma unroll
            for (int v = 0; v < TV; ++v)
#pragma unroll
                for (int c = 0; c < TC; ++c) out[v][c] = fadd(out[v][c], pair_term<Spec::OP>(a[v], b[c]));
        }
    }

    // cooperative load of `rows` rows (row-major, leading dimension d) into a pitched tile; rows
    // beyond `valid` are zero-filled.
    __device__ __forceinline__ static void load_rows(float* __restrict__ S, int pitch, const float* __restrict__ G,
                                                     int64_t row0, int64_t nrows_total, int rows, int d) {
        const int tid = threadIdx.x;
        const int nthr = blockDim.x;
        const bool vec = ((d & 3) == 0) && ((reinterpret_cast<uintptr_t>(G) & 15) == 0);
        if (vec) {
            const int d4 = d >> 2;
            for (int e = tid; e < rows * d4; e += nthr) {
                int r = e / d4, c4 = e - r * d4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row0 + r < nrows_total) v = reinterpret_cast<const float4*>(G + (row0 + r) * (int64_t)d)[c4];
                float* dst = S + (size_t)r * pitch + 4 * c4;
                dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
            }
        } else {
            for (int e = tid; e < rows * d; e += nthr) {
                int r = e / d, c = e - r * d;
                S[(size_t)r * pitch + c] = (row0 + r < nrows_total) ? G[(row0 + r) * (int64_t)d + c] : 0.0f;
            }
        }
    }
};

}  // namespace vix
