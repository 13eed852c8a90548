// vix_train_parity.cu -- reference-parity trainers (mode 0): the reference's control flow replayed on
// the host (RNG streams, batch composition, repairs) with the distance passes on the GPU.
#include "vix_common.cuh"

namespace vix {

int kmeans_parity_device(const float* x, int64_t n, int d, int kc, const float* init, const vix_kmeans_cfg* cfg,
                         float* centroids_out, int32_t* assign_out) {
    set_error("kmeans_minibatch (reference-parity mode) is not implemented yet; use cfg.mode = 1");
    return VIX_ERR_UNSUPPORTED;
}
int kmeanspp_parity_device(const float* x, int64_t n, int d, int k, uint64_t seed, uint64_t stream, float* centroids_out,
                           int64_t* chosen_out) {
    set_error("kmeanspp_seed (reference-parity mode) is not implemented yet");
    return VIX_ERR_UNSUPPORTED;
}
int pq_train_parity_device(const float* x, int64_t n, int d, int m, int ks, const float* coarse, const int32_t* assign,
                           const vix_pq_train_cfg* cfg, float* codebooks_out, float* norms_out) {
    set_error("pq_train (reference-parity mode) is not implemented yet; use cfg.mode = 1");
    return VIX_ERR_UNSUPPORTED;
}

}  // namespace vix
