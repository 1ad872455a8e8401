/*
 * c_client.c — the C ABI of libssd_b200.so used from plain C (C99, no C++ / torch / Python): what a binding in any
 * compiled host language does.  Two handles with the same (seed, first_env_id) step a small cleanup batch with the
 * CleanupContract fused in; observations, rewards and dones must agree byte for byte (counter-based randomness), and
 * the pipelined host-buffer step must deliver the same rewards as the device-resident one.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/c_client.c -o c_client \
 *       -L contracts_b200 -lssd_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/contracts_b200:/usr/local/cuda/lib64
 *
 * Exit codes: 0 ok; 3 ssd_create failed (e.g. no CUDA device: there is no CPU fallback); 1 any other failure.
 * Reference interface replaced: env_creator('CleanupNew', ...) + ContractWrapperSubgame, reset() / step()
 * (utils/env_creator_functions.py:12-34, environments/two_stage_train.py:62-121,159-187).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "ssd_b200.h"

#define E 1024
#define N 4
#define STEPS 40

static const char* MAP[6] = { "@@@@@@", "@PPPP@", "@PPPP@", "@HBBR@", "@PPPP@", "@@@@@@" };

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define SSD(h, x) do { int rc_ = (x); if (rc_ != SSD_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, ssd_last_error(h)); return 1; } } while (0)

typedef struct { ssd_handle* h; uint8_t* obs; double* rew; uint8_t* done; uint8_t* info; uint8_t* act; } batch;

static int make(batch* b, const char* flat)
{
    ssd_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = SSD_ABI_VERSION; cfg.struct_size = (int32_t)sizeof cfg;
    cfg.env_kind = SSD_ENV_CLEANUP; cfg.num_envs = E; cfg.num_agents = N;
    cfg.map_h = 6; cfg.map_w = 6; cfg.ascii_map = flat; cfg.horizon = 25;
    cfg.contract_kind = SSD_CONTRACT_CLEANUP; cfg.theta_low = 0.0; cfg.theta_high = (double)0.2f;
    cfg.seed = 73907u; cfg.first_env_id = 500u; cfg.device = 0;
    int rc = ssd_create(&cfg, &b->h);
    if (rc != SSD_OK) { fprintf(stderr, "ssd_create failed (%d): %s\n", rc, ssd_last_error(NULL)); return 3; }
    CU(cudaMalloc((void**)&b->obs, (size_t)E * N * SSD_OBS_BYTES_PER_AGENT));
    CU(cudaMalloc((void**)&b->rew, (size_t)E * N * sizeof(double)));
    CU(cudaMalloc((void**)&b->done, E));
    CU(cudaMalloc((void**)&b->info, (size_t)E * N * 4));
    CU(cudaMalloc((void**)&b->act, (size_t)E * N));
    return 0;
}

static uint64_t fnv(const void* p, size_t n, uint64_t h)
{
    const uint8_t* b = (const uint8_t*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

int main(void)
{
    char flat[37];
    for (int r = 0; r < 6; r++) memcpy(flat + 6 * r, MAP[r], 6);
    flat[36] = 0;
    if (ssd_abi_version() != SSD_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }

    batch a, b;
    memset(&a, 0, sizeof a); memset(&b, 0, sizeof b);
    int rc = make(&a, flat);
    if (rc) return rc;
    rc = make(&b, flat);
    if (rc) return rc;

    const size_t obs_bytes = (size_t)E * N * SSD_OBS_BYTES_PER_AGENT, rew_bytes = (size_t)E * N * sizeof(double);
    uint8_t* obs_a = (uint8_t*)malloc(obs_bytes); uint8_t* obs_b = (uint8_t*)malloc(obs_bytes);
    double* rew_a = (double*)malloc(rew_bytes); double* rew_b = (double*)malloc(rew_bytes); double* rew_x = (double*)malloc(rew_bytes);
    uint8_t done_a[E], done_b[E];

    /* handle b goes through the pipelined host-buffer step: pinned actions in, one compact result block out */
    ssd_host_layout lay;
    SSD(b.h, ssd_host_result_layout(b.h, &lay));
    void* block; uint8_t* act_host;
    CU(cudaMallocHost(&block, (size_t)lay.total_bytes));
    CU(cudaMallocHost((void**)&act_host, (size_t)E * N));

    SSD(a.h, ssd_reset(a.h, NULL, a.obs, 0, NULL));
    SSD(b.h, ssd_reset(b.h, NULL, b.obs, 0, NULL));

    uint64_t h_obs = 1469598103934665603ull, h_rew = h_obs;
    double total = 0.0; long finished = 0, records = 0;
    for (int t = 0; t < STEPS; t++) {
        ssd_step_io io;
        memset(&io, 0, sizeof io);
        SSD(a.h, ssd_random_actions(a.h, (uint32_t)t, 9, a.act, NULL));      /* stands in for the policy */
        io.actions_dev = a.act; io.obs_dev = a.obs; io.rew_dev = a.rew; io.done_dev = a.done; io.info_dev = a.info;
        io.auto_reset = 1;                                                   /* finished envs restart behind the step */
        SSD(a.h, ssd_step(a.h, &io, NULL));
        CU(cudaMemcpy(act_host, a.act, (size_t)E * N, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(obs_a, a.obs, obs_bytes, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(rew_a, a.rew, rew_bytes, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(done_a, a.done, E, cudaMemcpyDeviceToHost));

        ssd_step_io jo;
        memset(&jo, 0, sizeof jo);
        jo.obs_dev = b.obs; jo.rew_dev = b.rew; jo.done_dev = b.done; jo.info_dev = b.info; jo.auto_reset = 1;
        int64_t ticket = -1;
        SSD(b.h, ssd_step_host_async(b.h, &jo, act_host, block, &ticket, NULL));
        SSD(b.h, ssd_step_host_wait(b.h, ticket));
        SSD(b.h, ssd_host_result_expand(b.h, block, rew_x));
        CU(cudaMemcpy(obs_b, b.obs, obs_bytes, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(rew_b, b.rew, rew_bytes, cudaMemcpyDeviceToHost));
        memcpy(done_b, (const uint8_t*)block + lay.done_offset, E);
        records += *(const uint32_t*)((const uint8_t*)block + lay.count_offset);

        if (memcmp(obs_a, obs_b, obs_bytes) || memcmp(rew_a, rew_b, rew_bytes) || memcmp(done_a, done_b, E)) {
            fprintf(stderr, "step %d: the two handles diverged\n", t); return 1;
        }
        if (memcmp(rew_a, rew_x, rew_bytes)) { fprintf(stderr, "step %d: result block != dense rewards\n", t); return 1; }
        h_obs = fnv(obs_a, obs_bytes, h_obs); h_rew = fnv(rew_a, rew_bytes, h_rew);
        for (int i = 0; i < E * N; i++) total += rew_a[i];
        for (int i = 0; i < E; i++) finished += done_a[i];
    }
    if (finished != E) { fprintf(stderr, "expected every env to finish once at t = 25, got %ld\n", finished); return 1; }
    printf("c_client ok: %d envs x %d agents x %d steps, obs fnv %016llx, rew fnv %016llx, sum of rewards %.17g, "
           "%ld exact records, %lld kernel launches\n", E, N, STEPS, (unsigned long long)h_obs, (unsigned long long)h_rew,
           total, records, (long long)(ssd_kernel_launches(a.h) + ssd_kernel_launches(b.h)));
    ssd_destroy(a.h); ssd_destroy(b.h);
    return 0;
}
