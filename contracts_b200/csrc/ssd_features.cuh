// ssd_features.cuh — sm_100a kernels for the hand-designed-feature envs CleanupFeatures / HarvestFeatures
// ('Cleanup' / 'Harvest' tags) with the subgame contract wrapper fused in.
//
// Reference behaviour restated here (paths relative to the reference root):
//   environments/cleanup_features.py  step :156-254, reset :256-284, initialize_players :103-109,
//       spawn_apples_and_waste :111-125, compute_closest_* :127-154, compute_probabilities :286-303
//   environments/harvest_features.py  step :173-287, reset :289-336, spawn_apples :139-151,
//       count_apples_in_radius :128-137
//   contract/contract_list.py :22-27, :45-54 ; environments/two_stage_train.py :62-121, :159-187
//
// Mapping: these envs are list manipulations on <= 155 apple / 119 waste points with data-dependent sequential scans
// (HarvestFeatures' regrowth reads the list it is appending to), and the benchmark runs ~1M of them: one THREAD per
// env.  The current apple / waste lists are a presence bitmask (shared memory while stepping) plus a birth stamp per
// point (global memory, struct of arrays) — list order = stamp order, which decides np.argmin ties in
// compute_closest_*.  Feature rows (12+n or 10+2n doubles per agent) are transposed through a warp-private shared
// tile so that every global store instruction writes one contiguous row.
#pragma once
#include "ssd_common.cuh"

#define FEAT_THREADS 128
#define FEAT_MASK_WORDS 8               // up to 256 apple and 256 waste points
#define FEAT_MAXF 26                    // 10 + 2 * 8

struct FeatParams {
    int E, n, kind, H, W, F, horizon, contract;
    int n_apple, n_waste, n_spawn, potential;
    uint32_t seed, first_env_id;
    double theta_low, theta_high, null_prob;
    uint32_t thr_harvest[4], thr_waste;
    // static tables (global memory, read-only)
    const uint8_t* wall;        // [H*W]
    const int16_t* apple_idx;   // [H*W] index into the apple point list or -1
    const int16_t* waste_idx;   // [H*W]
    const uint16_t* apple_rc;   // [n_apple] row << 8 | col
    const uint16_t* waste_rc;   // [n_waste]
    const int16_t* apple_nbr;   // [n_apple][8] apple indices of the 3x3 neighbours or -1 (harvest regrowth)
    const uint16_t* spawn_rc;   // [n_spawn]
    const uint8_t* waste_start; // [n_waste] 1 if the point starts as waste ('H')
    const uint32_t* thr_apple;  // [potential + 1] apple spawn threshold by #waste (cleanup)
    const uint8_t* waste_on;    // [potential + 1]
    // state, struct of arrays
    uint32_t* agents;           // [n][E] row | col << 8 | ori << 16
    uint32_t* apple_mask;       // [FEAT_MASK_WORDS][E]
    uint32_t* waste_mask;       // [FEAT_MASK_WORDS][E]
    uint16_t* apple_stamp;      // [n_apple][E]
    uint16_t* waste_stamp;      // [n_waste][E]
    uint32_t* counters;         // [4][E] next apple stamp, next waste stamp, t, episode | initialised << 31
    double* theta;              // [E]
    double* metrics;            // [8][E] dirt, raw, transfers, apples, low_density
    uint32_t* sum_raw;          // [n][E]
    unsigned long long* tsum_raw; // [n][E]
    double* sum_tr;             // [n][E]
    double* tsum_tr;            // [n][E]
};

struct FeatIO {
    const uint8_t* actions;     // [E][n]
    double* obs;                // [E][n][F]
    double* rew; double* base_rew; double* transfers;   // [E][n]
    uint8_t* info;              // [E][n][4]: cleanup (cleaned_squares,0,0,0); harvest (eaten_apples, eaten_close_apples,0,0)
    uint8_t* done;              // [E]
    int auto_reset;             // next-step auto-reset (see feat_step_kernel)
};

struct FeatDraws {              // k-th random.random() of a step: Philox block cached
    uint32_t seed, env_id, episode, t, blk;
    Philox4 q;
    __device__ __forceinline__ uint32_t get(uint32_t k)
    {
        if ((k >> 2) != blk) { blk = k >> 2; q = philox4x32_10(blk, SITE_FEAT_SPAWN, t, episode, seed, env_id); }
        return pick(q, k & 3u);
    }
};

// per-thread views of the presence masks in shared memory: word w of this thread = m[w * FEAT_THREADS]
__device__ __forceinline__ bool mask_test(const uint32_t* m, int idx) { return (m[(idx >> 5) * FEAT_THREADS] >> (idx & 31)) & 1u; }
__device__ __forceinline__ void mask_set(uint32_t* m, int idx) { m[(idx >> 5) * FEAT_THREADS] |= 1u << (idx & 31); }
__device__ __forceinline__ void mask_clear(uint32_t* m, int idx) { m[(idx >> 5) * FEAT_THREADS] &= ~(1u << (idx & 31)); }

// count_apples_in_radius(radius, loc): j*j + k*k <= radius (sic) over the live list
__device__ __forceinline__ int feat_count_radius5(const FeatParams& p, const uint32_t* am, int r, int c)
{
    int cnt = 0;
    for (int j = -2; j <= 2; j++)
        for (int k = -2; k <= 2; k++) {
            if (j * j + k * k > 5) continue;
            const int rr = r + j, cc = c + k;
            if (rr < 0 || rr >= p.H || cc < 0 || cc >= p.W) continue;
            const int i = __ldg(p.apple_idx + rr * p.W + cc);
            if (i >= 0 && mask_test(am, i)) cnt++;
        }
    return cnt;
}

// spawn_apples_and_waste (cleanup_features.py:111-125) / spawn_apples (harvest_features.py:139-151)
__device__ __forceinline__ void feat_spawn(const FeatParams& p, int env, uint32_t* am, uint32_t* wm, const uint32_t* pos, FeatDraws& dr,
                                           uint32_t& next_apple, uint32_t& next_waste, int& n_cur_apple, int& n_cur_waste)
{
    const int n = p.n;
    // apple points under an agent are not eligible
    uint32_t occ[FEAT_MASK_WORDS];
#pragma unroll
    for (int w = 0; w < FEAT_MASK_WORDS; w++) occ[w] = 0u;
#pragma unroll
    for (int a = 0; a < SSD_MAXN; a++) {
        if (a >= n) continue;
        const int i = __ldg(p.apple_idx + (int)(pos[a] & 255u) * p.W + (int)((pos[a] >> 8) & 255u));
#pragma unroll
        for (int w = 0; w < FEAT_MASK_WORDS; w++) if (i >= 0 && (i >> 5) == w) occ[w] |= 1u << (i & 31);
    }
    const int aw = (p.n_apple + 31) >> 5;
    uint32_t k = 0;
    if (p.kind == SSD_ENV_CLEANUP_FEATURES) {
        const uint32_t thrA = __ldg(p.thr_apple + n_cur_waste);
        const bool waste_on = __ldg(p.waste_on + n_cur_waste) != 0;
#pragma unroll
        for (int w = 0; w < FEAT_MASK_WORDS; w++) {
            if (w >= aw) continue;
            const uint32_t valid = (w == aw - 1 && (p.n_apple & 31)) ? ((1u << (p.n_apple & 31)) - 1u) : 0xffffffffu;
            uint32_t elig = ~am[w * FEAT_THREADS] & ~occ[w] & valid;
            if (thrA == 0u) { k += (uint32_t)__popc(elig); continue; }      // r < 0 never holds; the draws are still consumed
            while (elig) {
                const int b = __ffs(elig) - 1; elig &= elig - 1;
                if (dr.get(k++) < thrA) {
                    const int i = w * 32 + b;
                    mask_set(am, i); p.apple_stamp[(size_t)i * p.E + env] = (uint16_t)next_apple++; n_cur_apple++;
                }
            }
        }
        if (waste_on) {
            const int ww = (p.n_waste + 31) >> 5;
            bool spawned = false;
            for (int w = 0; w < ww && !spawned; w++) {
                const uint32_t valid = (w == ww - 1 && (p.n_waste & 31)) ? ((1u << (p.n_waste & 31)) - 1u) : 0xffffffffu;
                uint32_t cand = ~wm[w * FEAT_THREADS] & valid;
                while (cand) {
                    const int b = __ffs(cand) - 1; cand &= cand - 1;
                    if (dr.get(k++) < p.thr_waste) {
                        const int i = w * 32 + b;
                        mask_set(wm, i); p.waste_stamp[(size_t)i * p.E + env] = (uint16_t)next_waste++; n_cur_waste++;
                        spawned = true;
                        break;
                    }
                }
            }
        }
    } else {
#pragma unroll
        for (int w = 0; w < FEAT_MASK_WORDS; w++) {
            if (w >= aw) continue;
            const uint32_t valid = (w == aw - 1 && (p.n_apple & 31)) ? ((1u << (p.n_apple & 31)) - 1u) : 0xffffffffu;
            uint32_t elig = ~am[w * FEAT_THREADS] & ~occ[w] & valid;
            while (elig) {
                const int b = __ffs(elig) - 1; elig &= elig - 1;
                const int i = w * 32 + b;
                int num = 0;                                                // live list: sees this loop's earlier spawns
                const int4 n0 = __ldg(reinterpret_cast<const int4*>(p.apple_nbr + i * 8));
                const int nb[8] = { (short)(n0.x & 0xFFFF), (short)(n0.x >> 16), (short)(n0.y & 0xFFFF), (short)(n0.y >> 16),
                                    (short)(n0.z & 0xFFFF), (short)(n0.z >> 16), (short)(n0.w & 0xFFFF), (short)(n0.w >> 16) };
#pragma unroll
                for (int q = 0; q < 8; q++) if (nb[q] >= 0 && mask_test(am, nb[q])) num++;
                const uint32_t kk = k++;
                if (num > 0 && dr.get(kk) < p.thr_harvest[num < 3 ? num : 3]) {
                    mask_set(am, i); p.apple_stamp[(size_t)i * p.E + env] = (uint16_t)next_apple++; n_cur_apple++;
                }
            }
        }
    }
}

// closest point of a list to each agent: smallest (L1 distance, birth stamp)  (np.argmin over the list in birth order).
// One key per (point, agent): distance << 24 | birth stamp << 8 | point index, so the search is a running minimum —
// no data-dependent stamp loads on ties (the stamp of every live point is read once, coalesced across the warp's
// envs: stamp[i][env]).  The L1 distances of a point to four agents at a time come from two VABSDIFF4.U8 (rows, columns:
// agents packed one per byte) and one add; distances are < 256, stamps 16 bits, indices < 256 (FEAT_MASK_WORDS * 32).
__device__ __forceinline__ void feat_closest(int n, int E, int env, const uint32_t* mask, int npts, const uint16_t* rc,
                                             const uint16_t* stamp, const uint32_t* pos, uint32_t* out_rc /* [MAXN] */)
{
    uint32_t best[SSD_MAXN];
    uint32_t ar[2] = { 0x7F7F7F7Fu, 0x7F7F7F7Fu }, ac[2] = { 0x7F7F7F7Fu, 0x7F7F7F7Fu };     // absent agents: (127, 127), no byte carry
#pragma unroll
    for (int a = 0; a < SSD_MAXN; a++) {
        best[a] = 0xFFFFFFFFu;
        if (a < n) {
            const uint32_t sh = 8u * (a & 3);
            ar[a >> 2] = (ar[a >> 2] & ~(0xFFu << sh)) | ((pos[a] & 255u) << sh);
            ac[a >> 2] = (ac[a >> 2] & ~(0xFFu << sh)) | (((pos[a] >> 8) & 255u) << sh);
        }
    }
    // Software-pipelined over the live points: the position word and the birth stamp of the NEXT point are requested before
    // the current one is evaluated, so the kernel's longest-latency loads (20 % of its stall samples when they sat in the
    // chain) overlap the key updates.
    const int nw = (npts + 31) >> 5;
    int w = 0;
    uint32_t m = nw > 0 ? mask[0] : 0u;
    int i_nx = -1; uint32_t prc_nx = 0, st_nx = 0;
    auto fetch_next = [&]() {
        while (m == 0u && ++w < nw) m = mask[w * FEAT_THREADS];
        if (m) {
            const int b = __ffs(m) - 1; m &= m - 1;
            i_nx = w * 32 + b;
            prc_nx = __ldg(rc + i_nx);
            st_nx = stamp[(size_t)i_nx * E + env];
        } else i_nx = -1;
    };
    fetch_next();
    while (i_nx >= 0) {
        const int i = i_nx; const uint32_t prc = prc_nx;
        const uint32_t base = (st_nx << 8) | (uint32_t)i;
        fetch_next();
        const uint32_t pr4 = __byte_perm(prc, 0u, 0x1111), pc4 = __byte_perm(prc, 0u, 0x0000);   // row / col in every byte
        const uint32_t d[2] = { __vabsdiffu4(pr4, ar[0]) + __vabsdiffu4(pc4, ac[0]), __vabsdiffu4(pr4, ar[1]) + __vabsdiffu4(pc4, ac[1]) };
#pragma unroll
        for (int a = 0; a < SSD_MAXN; a++) {
            if (a >= n) continue;
            const uint32_t key = __byte_perm(d[a >> 2], 0u, 0x0444u | ((uint32_t)(a & 3) << 12)) | base;   // distance -> byte 3
            best[a] = min(best[a], key);
        }
    }
#pragma unroll
    for (int a = 0; a < SSD_MAXN; a++) out_rc[a] = (a < n && best[a] != 0xFFFFFFFFu) ? (uint32_t)__ldg(rc + (best[a] & 255u)) : 0u;   // sentinel [0, 0]
}

// feature rows of all agents -> global memory.  Every thread stores its env's n rows itself (F float64 each, 16-byte
// stores): a row is one contiguous 160 / 208-byte run, an env's n rows are contiguous, and the L2 merges the sectors a warp's
// stores touch.  (Round 1 transposed the rows through a warp-private shared tile so that every store instruction wrote one
// contiguous row: 32 x n predicated store iterations per warp — 16 % of the kernel's instructions.)
__device__ __forceinline__ void feat_write_obs(const FeatParams& p, bool mine, int env, double* obs, const uint32_t* pos,
                                               const uint32_t* ca, const uint32_t* cw, const int* close5, const int* cleaned,
                                               int n_cur_apple, int n_cur_waste)
{
    if (!mine) return;
    const int n = p.n, F = p.F;
    const int cp0 = n > 1 ? 1 : 0;
    const bool cleanup = p.kind == SSD_ENV_CLEANUP_FEATURES;
    const bool even = (F & 1) == 0;                                   // rows are 16-byte aligned iff F is even
#pragma unroll
    for (int a = 0; a < SSD_MAXN; a++) {
        if (a >= n) continue;
        double v[FEAT_MAXF];                                          // fully unrolled below: stays in registers
        const uint32_t me = pos[a], other = pos[a == 0 ? cp0 : 0];    // compute_closest_pos quirk
        v[0] = (double)(me & 255u); v[1] = (double)((me >> 8) & 255u); v[2] = (double)((me >> 16) & 3u);
        v[3] = (double)(other & 255u); v[4] = (double)((other >> 8) & 255u); v[5] = (double)((other >> 16) & 3u);
        v[6] = (double)(ca[a] >> 8); v[7] = (double)(ca[a] & 255u);
        if (cleanup) {
            v[8] = (double)(cw[a] >> 8); v[9] = (double)(cw[a] & 255u); v[10] = (double)n_cur_apple; v[11] = (double)n_cur_waste;
#pragma unroll
            for (int i = 0; i < SSD_MAXN; i++) v[12 + i] = i < n ? (double)cleaned[i] : 0.0;
#pragma unroll
            for (int i = 20; i < FEAT_MAXF; i++) v[i] = 0.0;
        } else {
            v[8] = (double)close5[a]; v[9] = (double)n_cur_apple;
#pragma unroll
            for (int i = 10; i < FEAT_MAXF; i++) v[i] = 0.0;
        }
        double* o = obs + ((size_t)env * n + a) * F;
        if (even) {
#pragma unroll
            for (int k = 0; k < FEAT_MAXF; k += 2) if (k < F) reinterpret_cast<double2*>(o)[k >> 1] = make_double2(v[k], v[k + 1]);
        } else {
#pragma unroll
            for (int k = 0; k < FEAT_MAXF; k++) if (k < F) o[k] = v[k];
        }
    }
}

struct FeatState {      // per-thread scalars
    uint32_t next_apple, next_waste, episode;
    int t, n_cur_apple, n_cur_waste;
};

__device__ __forceinline__ int feat_popcount_mask(const uint32_t* m, int npts)
{
    int c = 0;
    for (int w = 0; w < ((npts + 31) >> 5); w++) c += __popc(m[w * FEAT_THREADS]);
    return c;
}

// reset of one env by its thread (cleanup_features.py:256-284 / harvest_features.py:289-336 + two_stage_train.py:159-187)
__device__ __forceinline__ void feat_reset_env(const FeatParams& p, int env, uint32_t* am, uint32_t* wm, uint32_t (&pos)[SSD_MAXN],
                                               uint32_t (&ca)[SSD_MAXN], uint32_t (&cw)[SSD_MAXN], int (&close5)[SSD_MAXN],
                                               int& n_cur_apple, int& n_cur_waste)
{
    const int n = p.n;
    {
        const uint32_t c3 = p.counters[(size_t)3 * p.E + env];
        const uint32_t episode = (c3 & 0x80000000u) ? (c3 & 0x7fffffffu) + 1u : 0u;
        const uint32_t env_id = p.first_env_id + (uint32_t)env;
        // initialize_arrays
        uint32_t next_apple = 1u, next_waste = 1u;
#pragma unroll
        for (int w = 0; w < FEAT_MASK_WORDS; w++) { am[w * FEAT_THREADS] = 0u; wm[w * FEAT_THREADS] = 0u; }
        if (p.kind == SSD_ENV_CLEANUP_FEATURES) {
            for (int i = 0; i < p.n_waste; i++)
                if (__ldg(p.waste_start + i)) { mask_set(wm, i); p.waste_stamp[(size_t)i * p.E + env] = (uint16_t)next_waste++; n_cur_waste++; }
        } else {
            for (int i = 0; i < p.n_apple; i++) { mask_set(am, i); p.apple_stamp[(size_t)i * p.E + env] = (uint16_t)next_apple++; n_cur_apple++; }
        }
        // initialize_players: agent a takes the spawn point with the (a+1)-th smallest (key, index).  Every key is drawn
        // once (one Philox block per four points) and inserted into a sorted list of the n smallest.
        {
            uint32_t bk[SSD_MAXN]; int bj[SSD_MAXN];
#pragma unroll
            for (int q = 0; q < SSD_MAXN; q++) { bk[q] = 0xffffffffu; bj[q] = 0x7fffffff; }
            Philox4 blk = { 0, 0, 0, 0 };
            for (int j = 0; j < p.n_spawn; j++) {
                if ((j & 3) == 0) blk = draw_block(p.seed, env_id, episode, 0u, SITE_FEAT_ORDER, 0u, (uint32_t)(j >> 2));
                uint32_t kk = pick(blk, (uint32_t)j & 3u); int jj = j;
#pragma unroll
                for (int q = 0; q < SSD_MAXN; q++) {                 // insertion: (kk, jj) sinks to its place, the rest shift down
                    const bool less = kk < bk[q] || (kk == bk[q] && jj < bj[q]);
                    const uint32_t tk = bk[q]; const int tj = bj[q];
                    if (less) { bk[q] = kk; bj[q] = jj; kk = tk; jj = tj; }
                }
            }
#pragma unroll
            for (int a = 0; a < SSD_MAXN; a++) {
                if (a >= n) continue;
                const uint32_t rc = __ldg(p.spawn_rc + bj[a]);
                const uint32_t o = draw_u32(p.seed, env_id, episode, 0u, SITE_FEAT_ROT, (uint32_t)a, 0u) >> 30;
                pos[a] = (rc >> 8) | ((rc & 255u) << 8) | (o << 16);
            }
        }
        FeatDraws dr = { p.seed, env_id, episode, 0u, 0xffffffffu, { 0, 0, 0, 0 } };
        feat_spawn(p, env, am, wm, pos, dr, next_apple, next_waste, n_cur_apple, n_cur_waste);
        feat_closest(n, p.E, env, am, p.n_apple, p.apple_rc, p.apple_stamp, pos, ca);
        if (p.kind == SSD_ENV_CLEANUP_FEATURES) feat_closest(n, p.E, env, wm, p.n_waste, p.waste_rc, p.waste_stamp, pos, cw);
        else {
#pragma unroll
            for (int a = 0; a < SSD_MAXN; a++) if (a < n) close5[a] = feat_count_radius5(p, am, (int)(pos[a] & 255u), (int)((pos[a] >> 8) & 255u));
        }
        double theta = 0.0;
        if (p.contract != SSD_CONTRACT_NONE) {
            const double u0 = __dmul_rn((double)draw_u32(p.seed, env_id, episode, 0u, SITE_CONTRACT, 0u, 0u), 1.0 / 4294967296.0);
            const double u1 = __dmul_rn((double)draw_u32(p.seed, env_id, episode, 0u, SITE_CONTRACT, 0u, 1u), 1.0 / 4294967296.0);
            theta = (u0 > p.null_prob) ? __dadd_rn(p.theta_low, __dmul_rn(__dsub_rn(p.theta_high, p.theta_low), u1)) : p.theta_low;
        }
        // state out
#pragma unroll
        for (int w = 0; w < FEAT_MASK_WORDS; w++) { p.apple_mask[(size_t)w * p.E + env] = am[w * FEAT_THREADS]; p.waste_mask[(size_t)w * p.E + env] = wm[w * FEAT_THREADS]; }
#pragma unroll
        for (int a = 0; a < SSD_MAXN; a++) {
            if (a >= n) continue;
            p.agents[(size_t)a * p.E + env] = pos[a];
            p.sum_raw[(size_t)a * p.E + env] = 0u; p.tsum_raw[(size_t)a * p.E + env] = 0ull;
            p.sum_tr[(size_t)a * p.E + env] = 0.0; p.tsum_tr[(size_t)a * p.E + env] = 0.0;
        }
        p.counters[(size_t)0 * p.E + env] = next_apple; p.counters[(size_t)1 * p.E + env] = next_waste;
        p.counters[(size_t)2 * p.E + env] = 0u; p.counters[(size_t)3 * p.E + env] = episode | 0x80000000u;
        p.theta[env] = theta;
        for (int q = 0; q < 8; q++) p.metrics[(size_t)q * p.E + env] = 0.0;
    }
}

__global__ void __launch_bounds__(FEAT_THREADS) feat_reset_kernel(const FeatParams p, const uint8_t* mask, double* obs)
{
    __shared__ uint32_t s_am[FEAT_MASK_WORDS * FEAT_THREADS], s_wm[FEAT_MASK_WORDS * FEAT_THREADS];
    const int env = blockIdx.x * FEAT_THREADS + threadIdx.x;
    const bool mine = env < p.E && (!mask || mask[env]);
    uint32_t* am = s_am + threadIdx.x; uint32_t* wm = s_wm + threadIdx.x;
    uint32_t pos[SSD_MAXN], ca[SSD_MAXN], cw[SSD_MAXN];
    int close5[SSD_MAXN], cleaned[SSD_MAXN];
    int n_cur_apple = 0, n_cur_waste = 0;
#pragma unroll
    for (int a = 0; a < SSD_MAXN; a++) { pos[a] = 0u; ca[a] = cw[a] = 0u; close5[a] = 0; cleaned[a] = 0; }
    if (mine) feat_reset_env(p, env, am, wm, pos, ca, cw, close5, n_cur_apple, n_cur_waste);
    if (obs) feat_write_obs(p, mine, env, obs, pos, ca, cw, close5, cleaned,
                            n_cur_apple, n_cur_waste);
}

#ifndef FEAT_MIN_BLOCKS
#define FEAT_MIN_BLOCKS 4          // 128 registers: 16 warps per SM (measured best of 3..6, tools/sweep_feat.sh)
#endif
__global__ void __launch_bounds__(FEAT_THREADS, FEAT_MIN_BLOCKS) feat_step_kernel(const FeatParams p, const FeatIO io)
{
    __shared__ uint32_t s_am[FEAT_MASK_WORDS * FEAT_THREADS], s_wm[FEAT_MASK_WORDS * FEAT_THREADS];
    const int env = blockIdx.x * FEAT_THREADS + threadIdx.x;
    const int n = p.n, W = p.W;
    const bool mine = env < p.E;
    uint32_t* am = s_am + threadIdx.x; uint32_t* wm = s_wm + threadIdx.x;
    uint32_t pos[SSD_MAXN], ca[SSD_MAXN], cw[SSD_MAXN];
    int close5[SSD_MAXN], cleaned[SSD_MAXN];
    int n_cur_apple = 0, n_cur_waste = 0;
#pragma unroll
    for (int a = 0; a < SSD_MAXN; a++) { pos[a] = 0u; ca[a] = cw[a] = 0u; close5[a] = 0; cleaned[a] = 0; }
    // next-step auto-reset (ssd_feat_io.auto_reset): an env that reached its horizon in the previous step starts its next
    // episode in this one — reset observation, zero rewards, done cleared, the actions of this step ignored
    const bool restart = mine && io.auto_reset && (int)p.counters[(size_t)2 * p.E + env] == p.horizon;
    if (restart) {
        feat_reset_env(p, env, am, wm, pos, ca, cw, close5, n_cur_apple, n_cur_waste);
        for (int a = 0; a < n; a++) {
            const size_t o = (size_t)env * n + a;
            io.rew[o] = 0.0;
            if (io.base_rew) io.base_rew[o] = 0.0;
            if (io.transfers) io.transfers[o] = 0.0;
            if (io.info) reinterpret_cast<uint32_t*>(io.info)[o] = 0u;
        }
        if (io.done) io.done[env] = 0;
    } else if (mine) {
        const uint32_t env_id = p.first_env_id + (uint32_t)env;
        uint32_t next_apple = p.counters[(size_t)0 * p.E + env], next_waste = p.counters[(size_t)1 * p.E + env];
        int t = (int)p.counters[(size_t)2 * p.E + env];
        const uint32_t episode = p.counters[(size_t)3 * p.E + env] & 0x7fffffffu;
        const double theta = p.theta[env];
#pragma unroll
        for (int w = 0; w < FEAT_MASK_WORDS; w++) { am[w * FEAT_THREADS] = p.apple_mask[(size_t)w * p.E + env]; wm[w * FEAT_THREADS] = p.waste_mask[(size_t)w * p.E + env]; }
        n_cur_apple = feat_popcount_mask(am, p.n_apple);
        n_cur_waste = feat_popcount_mask(wm, p.n_waste);
        int act[SSD_MAXN]; uint32_t claim[SSD_MAXN]; bool has[SSD_MAXN];
        int reward[SSD_MAXN], eaten[SSD_MAXN], eaten_close[SSD_MAXN];
        uint2 apk = make_uint2(0x04040404u, 0x04040404u);
        const bool packed_actions = n == 8 && (reinterpret_cast<uintptr_t>(io.actions) & 7u) == 0;
        if (packed_actions) apk = *reinterpret_cast<const uint2*>(io.actions + (size_t)env * 8);      // one 8-byte load
#pragma unroll
        for (int a = 0; a < SSD_MAXN; a++) {
            has[a] = false; claim[a] = 0u; reward[a] = eaten[a] = eaten_close[a] = 0; act[a] = 4;
            if (a < n) {
                pos[a] = p.agents[(size_t)a * p.E + env];
                act[a] = packed_actions ? (int)(((a < 4 ? apk.x : apk.y) >> (8 * (a & 3))) & 255u) : (int)io.actions[(size_t)env * n + a];
            }
        }
        const bool cleanup = p.kind == SSD_ENV_CLEANUP_FEATURES;
        // stay first: highest priority (cleanup: act == 4; harvest: every act > 3)
#pragma unroll
        for (int a = 0; a < SSD_MAXN; a++) if (a < n && (cleanup ? act[a] == 4 : act[a] > 3)) { claim[a] = pos[a] & 0xFFFFu; has[a] = true; }
        // movers, in agent order: blocked by walls and by squares already claimed
#pragma unroll
        for (int a = 0; a < SSD_MAXN; a++) {
            if (a >= n || act[a] > 3) continue;
            int r = (int)(pos[a] & 255u), c = (int)((pos[a] >> 8) & 255u);
            const int tr = r + (act[a] == 2 ? -1 : (act[a] == 3 ? 1 : 0)), tc = c + (act[a] == 0 ? -1 : (act[a] == 1 ? 1 : 0));
            const uint32_t tgt = (uint32_t)tr | ((uint32_t)tc << 8);
            bool blocked = tr < 0 || tr >= p.H || tc < 0 || tc >= W ? false : __ldg(p.wall + tr * W + tc) != 0;
#pragma unroll
            for (int b = 0; b < SSD_MAXN; b++) if (b < n && has[b] && claim[b] == tgt) blocked = true;
            claim[a] = blocked ? (pos[a] & 0xFFFFu) : tgt;
            has[a] = true;
        }
#pragma unroll
        for (int a = 0; a < SSD_MAXN; a++) if (a < n && has[a]) pos[a] = (pos[a] & 0xFFFF0000u) | claim[a];
        // consume in move_squares insertion order: stays (agent order), then movers (agent order)
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
#pragma unroll
            for (int a = 0; a < SSD_MAXN; a++) {
                if (a >= n || !has[a]) continue;
                const bool mover = act[a] <= 3;
                if ((pass == 1) != mover) continue;
                const int r = (int)(pos[a] & 255u), c = (int)((pos[a] >> 8) & 255u);
                const int i = __ldg(p.apple_idx + r * W + c);
                if (i < 0 || !mask_test(am, i)) continue;
                reward[a] += 1;
                if (!cleanup) {
                    eaten[a] += 1;
                    if (feat_count_radius5(p, am, r, c) < 4) eaten_close[a] += 1;
                }
                mask_clear(am, i); n_cur_apple--;
            }
        }
        // rotations
#pragma unroll
        for (int a = 0; a < SSD_MAXN; a++) {
            if (a >= n) continue;
            uint32_t o = (pos[a] >> 16) & 3u;
            if (act[a] == 5) o = (o + 1u) & 3u;
            if (act[a] == 6) o = (o + 3u) & 3u;
            pos[a] = (pos[a] & 0xFFFFu) | (o << 16);
        }
        // cleaning beams (cleanup_features.py:196-219): 3 rays x 6 cells incl. the agent's own, stopped by walls only
        int dirt = 0;
        if (cleanup) {
#pragma unroll
            for (int a = 0; a < SSD_MAXN; a++) {
                if (a >= n || (act[a] != 7 && act[a] != 8)) continue;
                const int o = (int)((pos[a] >> 16) & 3u), o1 = (o + 1) & 3;
                const int dr = o == 0 ? -1 : (o == 2 ? 1 : 0), dc = o == 1 ? 1 : (o == 3 ? -1 : 0);
                const int sr = o1 == 0 ? -1 : (o1 == 2 ? 1 : 0), sc = o1 == 1 ? 1 : (o1 == 3 ? -1 : 0);
                const int r0 = (int)(pos[a] & 255u), c0 = (int)((pos[a] >> 8) & 255u);
                for (int b = 0; b < 3; b++) {
                    const int br = r0 + (b == 1 ? sr : (b == 2 ? -sr : 0)), bc = c0 + (b == 1 ? sc : (b == 2 ? -sc : 0));
                    for (int j = 0; j < 6; j++) {
                        const int r = br + j * dr, c = bc + j * dc;
                        if (r < 0 || r >= p.H || c < 0 || c >= W) continue;
                        if (__ldg(p.wall + r * W + c)) break;
                        if (act[a] == 7) {
                            const int wi = __ldg(p.waste_idx + r * W + c);
                            if (wi >= 0 && mask_test(wm, wi)) { mask_clear(wm, wi); n_cur_waste--; cleaned[a]++; dirt++; }
                        }
                    }
                }
            }
        }
        FeatDraws dr = { p.seed, env_id, episode, (uint32_t)t + 1u, 0xffffffffu, { 0, 0, 0, 0 } };
        feat_spawn(p, env, am, wm, pos, dr, next_apple, next_waste, n_cur_apple, n_cur_waste);
        feat_closest(n, p.E, env, am, p.n_apple, p.apple_rc, p.apple_stamp, pos, ca);
        if (cleanup) feat_closest(n, p.E, env, wm, p.n_waste, p.waste_rc, p.waste_stamp, pos, cw);
        else {
#pragma unroll
            for (int a = 0; a < SSD_MAXN; a++) if (a < n) close5[a] = feat_count_radius5(p, am, (int)(pos[a] & 255u), (int)((pos[a] >> 8) & 255u));
        }
        t += 1;
        // rewards, contract transfers (contract_list.py:22-27,45-54), redistribution (two_stage_train.py:71-92)
        double r[SSD_MAXN], tr[SSD_MAXN], total = 0.0, raw = 0.0;
        int n_eaten = 0, n_close = 0;
#pragma unroll
        for (int a = 0; a < SSD_MAXN; a++) {
            r[a] = (double)reward[a]; tr[a] = 0.0;
            if (a >= n) continue;
            raw = __dadd_rn(raw, r[a]);
            n_eaten += eaten[a]; n_close += eaten_close[a];
            if (p.contract == SSD_CONTRACT_CLEANUP) tr[a] = __dmul_rn(-theta, (double)cleaned[a]);
            else if (p.contract == SSD_CONTRACT_HARVEST_LOCAL) tr[a] = (close5[a] < 4 && eaten_close[a] > 0) ? theta : 0.0;
        }
        if (p.contract != SSD_CONTRACT_NONE) {
#pragma unroll
            for (int i = 0; i < SSD_MAXN; i++) {
                if (i >= n) continue;
                r[i] = __dsub_rn(r[i], tr[i]); total = __dadd_rn(total, tr[i]);
                const double share = __ddiv_rn(tr[i], (double)(n - 1));
#pragma unroll
                for (int j = 0; j < SSD_MAXN; j++) if (j < n && j != i) r[j] = __dadd_rn(r[j], share);
            }
        }
        // outputs + state
#pragma unroll
        for (int a = 0; a < SSD_MAXN; a++) {
            if (a >= n) continue;
            const size_t o = (size_t)env * n + a, so = (size_t)a * p.E + env;
            io.rew[o] = r[a];
            if (io.base_rew) io.base_rew[o] = (double)reward[a];
            if (io.transfers) io.transfers[o] = tr[a];
            if (io.info) reinterpret_cast<uint32_t*>(io.info)[o] = cleanup ? (uint32_t)cleaned[a] : ((uint32_t)eaten[a] | ((uint32_t)eaten_close[a] << 8));
            p.agents[so] = pos[a];
            // episode accumulators as fire-and-forget reductions (one add per address and step: a float64 RED rounds
            // exactly like `x = x + y`), so no load round trip sits in the thread's dependency chain
            if (reward[a]) {
                red_add(p.sum_raw + so, (uint32_t)reward[a]);
                red_add(reinterpret_cast<long long*>(p.tsum_raw + so), (long long)((unsigned long long)(t - 1) * (unsigned long long)reward[a]));
            }
            if (r[a] != 0.0) {
                red_add(p.sum_tr + so, r[a]);
                red_add(p.tsum_tr + so, __dmul_rn((double)(t - 1), r[a]));
            }
        }
#pragma unroll
        for (int w = 0; w < FEAT_MASK_WORDS; w++) { p.apple_mask[(size_t)w * p.E + env] = am[w * FEAT_THREADS]; p.waste_mask[(size_t)w * p.E + env] = wm[w * FEAT_THREADS]; }
        p.counters[(size_t)0 * p.E + env] = next_apple; p.counters[(size_t)1 * p.E + env] = next_waste;
        p.counters[(size_t)2 * p.E + env] = (uint32_t)t;
        if (dirt) red_add(p.metrics + (size_t)0 * p.E + env, (double)dirt);
        if (raw != 0.0) red_add(p.metrics + (size_t)1 * p.E + env, raw);
        if (p.contract != SSD_CONTRACT_NONE && total != 0.0) red_add(p.metrics + (size_t)2 * p.E + env, total);
        if (n_eaten) red_add(p.metrics + (size_t)3 * p.E + env, (double)n_eaten);
        if (n_close) red_add(p.metrics + (size_t)4 * p.E + env, (double)n_close);
        // birth stamps are 16 bits (list order = stamp order decides np.argmin ties): an episode with more than 65535
        // spawns of one kind would wrap them — flag it instead of picking a wrong closest point silently
        if ((next_apple | next_waste) > 0xFFFFu) p.metrics[(size_t)5 * p.E + env] = 1.0;
        if (io.done) io.done[env] = t == p.horizon ? 1 : 0;
    }
    feat_write_obs(p, mine, env, io.obs, pos, ca, cw, close5, cleaned,
                   n_cur_apple, n_cur_waste);
}

// metrics [E][40]: dirt, raw, transfers, apples, low_density, 0, 0, 0, sum_raw[8], tsum_raw[8], sum_tr[8], tsum_tr[8]
__global__ void feat_get_metrics_kernel(const FeatParams p, double* out)
{
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.E) return;
    double* o = out + (size_t)env * 40;
    for (int q = 0; q < 8; q++) o[q] = q < 6 ? p.metrics[(size_t)q * p.E + env] : 0.0;        // [5]: err_flags
    for (int a = 0; a < SSD_MAXN; a++) {
        const bool v = a < p.n;
        const size_t so = (size_t)a * p.E + env;
        o[8 + a] = v ? (double)p.sum_raw[so] : 0.0; o[16 + a] = v ? (double)p.tsum_raw[so] : 0.0;
        o[24 + a] = v ? p.sum_tr[so] : 0.0; o[32 + a] = v ? p.tsum_tr[so] : 0.0;
    }
}
// pos int32 [E][n][2], ori int32 [E][n], cells u8 [E][H][W] (1 apple, 2 waste), theta f64 [E], t int32 [E]
__global__ void feat_get_state_kernel(const FeatParams p, int32_t* pos, int32_t* ori, uint8_t* cells, double* theta, int32_t* t)
{
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.E) return;
    for (int a = 0; a < p.n; a++) {
        const uint32_t v = p.agents[(size_t)a * p.E + env];
        if (pos) { pos[((size_t)env * p.n + a) * 2] = (int)(v & 255u); pos[((size_t)env * p.n + a) * 2 + 1] = (int)((v >> 8) & 255u); }
        if (ori) ori[(size_t)env * p.n + a] = (int)((v >> 16) & 3u);
    }
    if (cells) {
        uint8_t* c = cells + (size_t)env * p.H * p.W;
        for (int i = 0; i < p.H * p.W; i++) c[i] = 0;
        for (int i = 0; i < p.n_apple; i++)
            if ((p.apple_mask[(size_t)(i >> 5) * p.E + env] >> (i & 31)) & 1u) { const uint32_t rc = p.apple_rc[i]; c[(rc >> 8) * p.W + (rc & 255u)] = 1; }
        for (int i = 0; i < p.n_waste; i++)
            if ((p.waste_mask[(size_t)(i >> 5) * p.E + env] >> (i & 31)) & 1u) { const uint32_t rc = p.waste_rc[i]; c[(rc >> 8) * p.W + (rc & 255u)] = 2; }
    }
    if (theta) theta[env] = p.theta[env];
    if (t) t[env] = (int)p.counters[(size_t)2 * p.E + env];
}
__global__ void feat_set_theta_kernel(const FeatParams p, const double* theta)
{
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env < p.E) p.theta[env] = theta[env];
}
__global__ void feat_random_actions_kernel(const FeatParams p, uint32_t step_index, uint32_t* counter, int num_actions, uint8_t* actions)
{
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (counter) step_index = *reinterpret_cast<volatile uint32_t*>(counter);
    if (env < p.E) {
        const uint32_t env_id = p.first_env_id + (uint32_t)env;
        for (int b = 0; b * 4 < p.n; b++) {
            Philox4 q = philox4x32_10((uint32_t)b, SITE_ACTIONS, step_index, 0u, p.seed, env_id);
            uint32_t w[4] = { q.x, q.y, q.z, q.w };
            for (int j = 0; j < 4 && b * 4 + j < p.n; j++)
                actions[(size_t)env * p.n + b * 4 + j] = (uint8_t)(((uint64_t)w[j] * (uint32_t)num_actions) >> 32);
        }
    }
    if (counter) counter_finish(counter, step_index);
}
