// vix_gemm.cu -- tensor-core (tcgen05) contraction path for the GEMM-shaped stages (coarse probe
// selection, flat scan).  Round-1 state: the entry point below forwards to the exact CUDA-core
// kernels of vix_scoring.cu; the tcgen05 + TMA shortlist kernel replaces the body, with the exact
// kernels kept as the rescoring stage (see DESIGN.md "Tensor-core plan").
#include "vix_common.cuh"

namespace vix {

int probe_select_device(const float* q, int64_t nq, const float* c, int kc, int d, int metric, int nprobe,
                        const float* cnorm, const uint64_t* disabled, int32_t* out_idx, float* out_scores);

int probe_select_fast_device(const float* q, int64_t nq, const float* c, int kc, int d, int metric, int nprobe,
                             const float* cnorm, int32_t* out_idx, float* out_scores) {
    return probe_select_device(q, nq, c, kc, d, metric, nprobe, cnorm, nullptr, out_idx, out_scores);
}

}  // namespace vix
