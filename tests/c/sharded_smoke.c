/* The multi-GPU step from a plain C11 host, no Python, no torch, no MPI: the parent forks one process per GPU, hands the
 * NCCL id of rank 0 to the others through pipes (a Swift host would use whatever starts its workers), and every rank runs
 *   vix_comm_create -> vix_index_set_coarse / set_codebooks -> vix_sharded_add -> vix_sharded_search
 * on synthetic data.  Rank 0 also holds a single-GPU index with ALL rows and requires the sharded result to equal its
 * result: same ids, distances to fp32 rounding (the order in which a vector's table entries are summed follows its slot in
 * its list, and a shard receives its rows in another order than the single index).  Usage: sharded_smoke [world]  (needs `world` GPUs; built and run by
 * tests/test_sharded_gpu.py).  VIX_NO_P2P=1 exercises the NCCL all-gather fallback. */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include "vindex_cuda.h"

enum { D = 32, M = 16, KC = 24, N = 6000, NQ = 203, K = 7, NPROBE = 5 };

static float frand(uint64_t* s) {
    *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (float)((*s >> 40) & 0xFFFFFF) / 16777216.0f * 2.0f - 1.0f;
}

#define CHECK(call)                                                                              \
    do {                                                                                         \
        int rc_ = (call);                                                                        \
        if (rc_ != VIX_OK) {                                                                     \
            fprintf(stderr, "rank %d: %s -> %d (%s)\n", rank, #call, rc_, vix_last_error());     \
            return 10 + rank;                                                                    \
        }                                                                                        \
    } while (0)

static int run_rank(int rank, int world, const unsigned char* id) {
    CHECK(vix_set_device(rank));
    vix_comm_t* comm = NULL;
    CHECK(vix_comm_create(id, 128, rank, world, &comm));
    /* identical parameters and data on every rank (same seed) */
    uint64_t seed = 12345;
    float* coarse = malloc(sizeof(float) * KC * D);
    float* cb = malloc(sizeof(float) * M * 256 * (D / M));
    float* xb = malloc(sizeof(float) * N * D);
    float* q = malloc(sizeof(float) * NQ * D);
    for (int i = 0; i < KC * D; ++i) coarse[i] = frand(&seed);
    for (int i = 0; i < M * 256 * (D / M); ++i) cb[i] = 0.3f * frand(&seed);
    for (int i = 0; i < N; ++i) {
        const int c = i % KC;
        for (int e = 0; e < D; ++e) xb[i * D + e] = coarse[c * D + e] + 0.25f * frand(&seed);
    }
    for (int i = 0; i < NQ * D; ++i) q[i] = frand(&seed);
    vix_index_params p;
    vix_index_params_default(&p);
    p.kind = VIX_INDEX_IVF_PQ; p.d = D; p.metric = VIX_METRIC_L2; p.nlist = KC; p.nprobe = NPROBE; p.m = M; p.ks = 256;
    vix_index_t* shard = NULL;
    CHECK(vix_index_create(&p, &shard));
    CHECK(vix_index_set_coarse(shard, coarse, KC));
    CHECK(vix_index_set_codebooks(shard, cb, NULL));
    /* rank r contributes rows r, r + world, ...: ragged, and rank world - 1 contributes nothing in the second call */
    int64_t mine = 0;
    float* xr = malloc(sizeof(float) * N * D);
    int64_t* ir = malloc(sizeof(int64_t) * N);
    for (int i = rank; i < N; i += world) { memcpy(xr + mine * D, xb + (size_t)i * D, sizeof(float) * D); ir[mine++] = 1000 + i; }
    const int64_t half = mine / 2;
    CHECK(vix_sharded_add(shard, comm, NULL, xr, ir, half));
    const int64_t rest = rank == world - 1 ? 0 : mine - half;       /* the last rank keeps its second half for call three */
    CHECK(vix_sharded_add(shard, comm, NULL, xr + half * D, ir + half, rest));
    CHECK(vix_sharded_add(shard, comm, NULL, xr + half * D, ir + half, rank == world - 1 ? mine - half : 0));
    float* sd = malloc(sizeof(float) * NQ * K);
    int64_t* si = malloc(sizeof(int64_t) * NQ * K);
    CHECK(vix_sharded_search(shard, comm, q, NQ, K, 0, sd, si));
    int64_t total = vix_index_count(shard);
    printf("rank %d: %lld rows in its lists, peer memory %d\n", rank, (long long)total, vix_comm_uses_peer_memory(comm));
    int bad = 0;
    if (rank == 0) {
        vix_index_t* full = NULL;
        CHECK(vix_index_create(&p, &full));
        CHECK(vix_index_set_coarse(full, coarse, KC));
        CHECK(vix_index_set_codebooks(full, cb, NULL));
        int64_t* ids = malloc(sizeof(int64_t) * N);
        for (int i = 0; i < N; ++i) ids[i] = 1000 + i;
        CHECK(vix_index_add(full, xb, ids, N));
        float* fd = malloc(sizeof(float) * NQ * K);
        int64_t* fi = malloc(sizeof(int64_t) * NQ * K);
        CHECK(vix_index_search(full, q, NQ, K, 0, fd, fi));
        for (int i = 0; i < NQ * K; ++i)
            if (fi[i] != si[i] || fabsf(fd[i] - sd[i]) > 2e-6f * fabsf(fd[i])) {
                if (bad < 5) fprintf(stderr, "mismatch at %d: sharded (%lld, %g) single (%lld, %g)\n", i, (long long)si[i], sd[i], (long long)fi[i], fd[i]);
                ++bad;
            }
        printf("rank 0: sharded == single-GPU on %d results: %s\n", NQ * K, bad ? "NO" : "yes");
        vix_index_destroy(full);
    }
    /* a second, smaller batch (regions are reused) and an empty one */
    CHECK(vix_sharded_search(shard, comm, q, 9, K, 3, sd, si));
    CHECK(vix_sharded_search(shard, comm, q, 0, K, 3, sd, si));
    vix_index_destroy(shard);
    vix_comm_destroy(comm);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    const int world = argc > 1 ? atoi(argv[1]) : 2;
    /* no CUDA call before the forks (a forked child cannot use a CUDA context its parent initialised): the caller makes
     * sure `world` GPUs are visible */
    int pipes[64][2];
    pid_t pids[64];
    if (world < 1 || world > 64) return 2;
    for (int r = 1; r < world; ++r) {
        if (pipe(pipes[r]) != 0) return 3;
        pids[r] = fork();
        if (pids[r] == 0) {                                       /* child = rank r: the id arrives through its pipe */
            unsigned char id[128];
            close(pipes[r][1]);
            size_t got = 0;
            while (got < sizeof(id)) { ssize_t n = read(pipes[r][0], id + got, sizeof(id) - got); if (n <= 0) return 4; got += (size_t)n; }
            return run_rank(r, world, id);
        }
        close(pipes[r][0]);
    }
    unsigned char id[128];
    if (vix_comm_unique_id(id, sizeof(id)) != VIX_OK) { fprintf(stderr, "vix_comm_unique_id: %s\n", vix_last_error()); return 5; }
    for (int r = 1; r < world; ++r) { if (write(pipes[r][1], id, sizeof(id)) != (ssize_t)sizeof(id)) return 6; close(pipes[r][1]); }
    int rc = run_rank(0, world, id);
    for (int r = 1; r < world; ++r) { int st = 0; waitpid(pids[r], &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = rc ? rc : 20 + r; }
    printf(rc ? "FAILED (%d)\n" : "ok%.0d\n", rc);
    return rc;
}
