// ssd_selfdrive.cuh — sm_100a kernels for SelfAcceleratingCarEnv (+ SelfdriveContractDistprop subgame wrapper).
//
// Reference behaviour restated here (paths relative to the reference root):
//   environments/self_driving_car_accelerate.py  reset :49-79, step :151-250, update_rel_rank :110-125,
//       update_infos :127-149, make_new_pos_consistent :92-108 (collision_on=False)
//   contract/contract_list.py :66-102 ; environments/two_stage_train.py :62-121, :159-187
//
// Mapping: the kinematics are <= 8 cars of serial float64 arithmetic per env, the output is a 2n+5-double
// observation row per car that is a pure function of the 2n positions / velocities.  So: one THREAD per env for the
// step logic (state is struct-of-arrays: coalesced), positions / velocities are parked in shared memory, and each warp
// then expands the observation rows of its 32 envs with lanes = row elements, which makes the dominant HBM write
// (n * (2n+5) * 8 B per env) contiguous.  Everything is float64 and rounds like the reference (-fmad=false).
#pragma once
#include "ssd_common.cuh"

#define CAR_THREADS 128
#define CAR_STRIDE 129                  // doubles per car row in shared memory (odd: conflict-free column walks)

struct CarParams {
    int E, n, D, contract;
    double low_bound, high_bound, start_vel, start_vel_amb, theta_low, theta_high, null_prob;
    uint32_t seed, first_env_id;
    // state, struct of arrays
    double* pos;        // [n][E]
    double* vel;        // [n][E]
    double* theta;      // [E]
    double* m_transfers;// [E]
    double* dist_front; // [E]  dist_to_front['a{n-1}'] (:144)
    uint32_t* crossed;  // [E]  crossed_agents, 4 bits per entry
    uint32_t* meta;     // [E]  n_crossed | done mask << 8 | all_done << 16 | initialised << 31
    int32_t* t;         // [E]
    uint32_t* episode;  // [E]
};

struct CarIO {
    const float* actions;   // [E][n]
    double* obs;            // [E][n][D]
    double* rew; double* base_rew; double* transfers;   // [E][n]
    double* info;           // [E][n][4]: just_passed, active, ambulance_rank, ambulance_dist_to_front (row of the first acting agent)
    uint8_t* done;          // [E][n+1]
    int auto_reset;         // next-step auto-reset (see car_step_kernel)
};

__device__ __forceinline__ double py_min(double x, double y) { return y < x ? y : x; }   // Python min([x, y])
__device__ __forceinline__ double py_max(double x, double y) { return y > x ? y : x; }
__device__ __forceinline__ double car_u01(const CarParams& p, uint32_t env_id, uint32_t episode, uint32_t site, uint32_t idx)
{
    return __dmul_rn((double)draw_u32(p.seed, env_id, episode, 0u, site, 0u, idx), 1.0 / 4294967296.0);
}

// each warp writes the observation rows of its 32 envs (:244-249).  The n * D doubles of an env and the envs of a warp
// are contiguous in `obs`, so the warp walks the 32 * n * D block with flat, fully coalesced 256-byte stores; a lane
// tracks (env, car, element) of its flat index incrementally.
__device__ __forceinline__ void car_write_obs(const CarParams& p, const double* s_pos, const double* s_vel, double* obs,
                                              int env0, int lane, unsigned valid_mask)
{
    const int n = p.n, D = p.D, total = 32 * n * D;
    double* base = obs + (size_t)env0 * n * D;
    const int col0 = threadIdx.x & ~31;
    int el = 0, k = 0, j = lane;                        // flat index f = (el * n + k) * D + j
    while (j >= D) { j -= D; k++; }
    while (k >= n) { k -= n; el++; }
    for (int f = lane; f < total; f += 32) {
        if ((valid_mask >> el) & 1u) {
            const int col = col0 + el;
            const double pk = s_pos[k * CAR_STRIDE + col];
            double v;
            if (j < 2) v = j == 0 ? pk : s_vel[k * CAR_STRIDE + col];
            else if (j < 2 + 2 * n) {
                const int q = j - 2;
                const double x = (q < n ? s_pos : s_vel)[(q < n ? q : q - n) * CAR_STRIDE + col];
                v = q < n ? __dsub_rn(x, pk) : x;
            } else if (j == 2 + 2 * n) v = s_pos[col] > 0 ? 1.0 : 0.0;
            else if (j == 3 + 2 * n) v = pk > 0 ? 1.0 : 0.0;
            else v = 0.0;
            base[f] = v;
        }
        j += 32;
        while (j >= D) { j -= D; k++; }
        while (k >= n) { k -= n; el++; }
    }
}

// reset of one env by its thread (self_driving_car_accelerate.py:49-79 + two_stage_train.py:159-187); leaves the new
// positions / velocities in the thread's shared-memory columns for car_write_obs
__device__ __forceinline__ void car_reset_env(const CarParams& p, int env, double* s_pos, double* s_vel)
{
    const int n = p.n;
    {
        const uint32_t meta = p.meta[env];
        const uint32_t episode = (meta & 0x80000000u) ? p.episode[env] + 1u : 0u;
        const uint32_t env_id = p.first_env_id + (uint32_t)env;
        for (int k = 0; k < n; k++) {
            const double u = car_u01(p, env_id, episode, SITE_SELFDRIVE_RESET, (uint32_t)k);
            double x, v;
            if (k == 0) {        // random.random() * low / 2 + low / 2  (:53)
                x = __dadd_rn(__ddiv_rn(__dmul_rn(u, p.low_bound), 2.0), __ddiv_rn(p.low_bound, 2.0));
                v = p.start_vel_amb;
            } else {             // random.random() * low / 16 + low * 3 / 16  (:57)
                x = __dadd_rn(__ddiv_rn(__dmul_rn(u, p.low_bound), 16.0), __ddiv_rn(__dmul_rn(p.low_bound, 3.0), 16.0));
                v = p.start_vel;
            }
            p.pos[(size_t)k * p.E + env] = x; p.vel[(size_t)k * p.E + env] = v;
            s_pos[k * CAR_STRIDE + threadIdx.x] = x; s_vel[k * CAR_STRIDE + threadIdx.x] = v;
        }
        double theta = 0.0;
        if (p.contract != SSD_CONTRACT_NONE) {       // two_stage_train.py:163-166
            const double u0 = car_u01(p, env_id, episode, SITE_CONTRACT, 0u), u1 = car_u01(p, env_id, episode, SITE_CONTRACT, 1u);
            theta = (u0 > p.null_prob) ? __dadd_rn(p.theta_low, __dmul_rn(__dsub_rn(p.theta_high, p.theta_low), u1)) : p.theta_low;
        }
        p.theta[env] = theta; p.m_transfers[env] = 0.0; p.dist_front[env] = -1.0;
        p.crossed[env] = 0u; p.meta[env] = 0x80000000u; p.t[env] = 0; p.episode[env] = episode;
    }
}

__global__ void __launch_bounds__(CAR_THREADS) car_reset_kernel(const CarParams p, const uint8_t* mask, double* obs)
{
    __shared__ double s_pos[SSD_MAXN * CAR_STRIDE], s_vel[SSD_MAXN * CAR_STRIDE];
    const int env = blockIdx.x * CAR_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool mine = env < p.E && (!mask || mask[env]);
    if (mine) car_reset_env(p, env, s_pos, s_vel);
    const unsigned valid = __ballot_sync(0xffffffffu, mine);
    __syncwarp();
    if (obs) car_write_obs(p, s_pos, s_vel, obs, env - lane, lane, valid);
}

__global__ void __launch_bounds__(CAR_THREADS) car_step_kernel(const CarParams p, const CarIO io)
{
    __shared__ double s_pos[SSD_MAXN * CAR_STRIDE], s_vel[SSD_MAXN * CAR_STRIDE], s_new[SSD_MAXN * CAR_STRIDE];
    const int env = blockIdx.x * CAR_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, tx = threadIdx.x;
    const int n = p.n;
    const bool mine = env < p.E;
    // next-step auto-reset (ssd_selfdrive_io.auto_reset): an env whose episode ended in the previous step starts its next
    // episode in this one — reset observation, zero rewards, dones cleared, the actions of this step ignored
    const bool restart = mine && io.auto_reset && ((p.meta[env] >> 16) & 1u);
    if (restart) {
        car_reset_env(p, env, s_pos, s_vel);
        for (int k = 0; k < n; k++) {
            const size_t o = (size_t)env * n + k;
            io.rew[o] = 0.0;
            if (io.base_rew) io.base_rew[o] = 0.0;
            if (io.transfers) io.transfers[o] = 0.0;
            if (io.info) reinterpret_cast<double4*>(io.info)[o] = make_double4(0.0, 0.0, 0.0, 0.0);
            if (io.done) io.done[(size_t)env * (n + 1) + k] = 0;
        }
        if (io.done) io.done[(size_t)env * (n + 1) + n] = 0;
    } else if (mine) {
        uint32_t meta = p.meta[env];
        int n_crossed = (int)(meta & 0xFFu);
        uint32_t done_mask = (meta >> 8) & 0xFFu;
        bool all_done = (meta >> 16) & 1u;
        uint32_t crossed = p.crossed[env];
        const double theta = p.theta[env];
        for (int k = 0; k < n; k++) { s_pos[k * CAR_STRIDE + tx] = p.pos[(size_t)k * p.E + env]; s_vel[k * CAR_STRIDE + tx] = p.vel[(size_t)k * p.E + env]; }
        const uint32_t active = all_done ? 0u : (~done_mask & ((1u << n) - 1u));   // done agents stop acting (RLlib)
        double rew[SSD_MAXN], tr[SSD_MAXN];
        double info2 = 0.0, info3 = 0.0;
        uint32_t just = 0u;
#pragma unroll
        for (int k = 0; k < SSD_MAXN; k++) { rew[k] = 0.0; tr[k] = 0.0; }
        const int first = active ? __ffs(active) - 1 : -1;
        if (active) {
            p.t[env] += 1;
            for (int k = 0; k < n; k++) {
                double x = s_pos[k * CAR_STRIDE + tx];
                if ((active >> k) & 1u) {
                    const double a = (double)io.actions[(size_t)env * n + k];
                    const double v = py_max(py_min(__dadd_rn(py_max(py_min(a, 0.1), -0.1), s_vel[k * CAR_STRIDE + tx]),
                                                   k == 0 ? 1.0 : 0.25), 0.0);          // :172-174
                    s_vel[k * CAR_STRIDE + tx] = v;
                    x = __dadd_rn(v, x);                                                 // :180
                }
                s_new[k * CAR_STRIDE + tx] = x;
            }
            // infos defaults (:183-189)
            int ci = -1;
            for (int i = 0; i < n_crossed; i++) if (((crossed >> (4 * i)) & 15u) == 0u) ci = i;
            info2 = ci >= 0 ? (double)(ci + 1) : (double)n;
            { const double d0 = n == 1 ? p.dist_front[env] : -1.0; info3 = d0 > -1.0 ? d0 : __dsub_rn(p.high_bound, p.low_bound); }
            // update_rel_rank (:110-125): crossers appended in index order
            for (int k = 0; k < n; k++) {
                if (((active >> k) & 1u) && s_pos[k * CAR_STRIDE + tx] < 0.0 && s_new[k * CAR_STRIDE + tx] > 0.0) {
                    just |= 1u << k;
                    crossed |= (uint32_t)k << (4 * n_crossed);
                    n_crossed++;
                }
            }
            // update_infos (:127-149)
            for (int k = 0; k < n; k++) {
                if (!((just >> k) & 1u)) continue;
                const double nk = s_new[k * CAR_STRIDE + tx];
                double d = 0.0;
                for (int i = 0; i < n; i++) {
                    if (i == k) continue;
                    const bool act_i = (active >> i) & 1u;
                    const double ni = s_new[i * CAR_STRIDE + tx];
                    if (!act_i) { if (__dsub_rn(p.high_bound, nk) > d) d = __dsub_rn(__dadd_rn(p.high_bound, 1.0), nk); }
                    else if (ni > nk) { if (__dsub_rn(ni, nk) > d) d = __dsub_rn(ni, ni); }      // sic (:143)
                }
                p.dist_front[env] = d;
                if (k == 0) {
                    int c0 = 0;
                    for (int i = 0; i < n_crossed; i++) if (((crossed >> (4 * i)) & 15u) == 0u) c0 = i;
                    info2 = (double)(c0 + 1); info3 = d;
                }
            }
            // make_new_pos_consistent (:92-108)
            uint32_t pre = 0u;
            for (int i = 0; i + 1 < n_crossed; i++) {
                const int f = (int)((crossed >> (4 * i)) & 15u), b = (int)((crossed >> (4 * i + 4)) & 15u);
                const bool af = (active >> f) & 1u, ab = (active >> b) & 1u;
                const double pf = af ? s_new[f * CAR_STRIDE + tx] : s_pos[f * CAR_STRIDE + tx];
                const double pb = ab ? s_new[b * CAR_STRIDE + tx] : s_pos[b * CAR_STRIDE + tx];
                if (pf < pb && af && ab) {
                    const double nb = __dsub_rn(pf, 0.01);
                    s_new[b * CAR_STRIDE + tx] = nb;
                    if (nb < 0) pre |= 1u << b;
                }
            }
            if (pre) {
                uint32_t kept = 0u; int m = 0;
                for (int i = 0; i < n_crossed; i++) {
                    const uint32_t a = (crossed >> (4 * i)) & 15u;
                    if (!((pre >> a) & 1u)) { kept |= a << (4 * m); m++; }
                }
                crossed = kept; n_crossed = m;
            }
            // new positions, rewards, dones (:216-238)
#pragma unroll
            for (int k = 0; k < SSD_MAXN; k++) {
                if (k >= n) continue;
                double x = ((active >> k) & 1u) ? s_new[k * CAR_STRIDE + tx] : s_pos[k * CAR_STRIDE + tx];
                if ((active >> k) & 1u) rew[k] = k == 0 ? __dsub_rn(-1.0, 99.0) : -1.0;
                if (x > p.high_bound) { x = __dadd_rn(p.high_bound, 1.0); done_mask |= 1u << k; }
                s_pos[k * CAR_STRIDE + tx] = x;
                p.pos[(size_t)k * p.E + env] = x; p.vel[(size_t)k * p.E + env] = s_vel[k * CAR_STRIDE + tx];
            }
            all_done = (active & ~done_mask) == 0u;
        }
        double base[SSD_MAXN];
#pragma unroll
        for (int k = 0; k < SSD_MAXN; k++) base[k] = rew[k];
        // SelfdriveContractDistprop (contract_list.py:66-102) + redistribution (two_stage_train.py:71-92)
        if (p.contract == SSD_CONTRACT_SELFDRIVE_DISTPROP && (active & 1u) && (just & 1u)) {
            const double x0 = s_pos[tx];
            uint32_t behind = 0u;
            double sum = 0.0;
            for (int i = 1; i < n; i++) {
                const double rel = __dsub_rn(s_pos[i * CAR_STRIDE + tx], x0);
                if (rel < 0) { behind |= 1u << i; sum = __dadd_rn(sum, -rel); }
            }
            double total = 0.0;
            if (behind) {
                const double v0 = __dmul_rn(theta, sum);
                tr[0] = v0;
                rew[0] = __dsub_rn(rew[0], v0); total = __dadd_rn(total, v0);
#pragma unroll
                for (int j = 1; j < SSD_MAXN; j++)
                    if (j < n && ((behind >> j) & 1u) && ((active >> j) & 1u)) {
                        const double dj = -__dsub_rn(s_pos[j * CAR_STRIDE + tx], x0);
                        rew[j] = __dadd_rn(rew[j], __dmul_rn(v0, __ddiv_rn(dj, sum)));
                    }
            }
#pragma unroll
            for (int i = 1; i < SSD_MAXN; i++) {
                if (i >= n || !((active >> i) & 1u) || ((behind >> i) & 1u)) continue;
                const double vi = __dmul_rn(theta, __dsub_rn(s_pos[i * CAR_STRIDE + tx], x0));
                tr[i] = vi;
                rew[i] = __dsub_rn(rew[i], vi); total = __dadd_rn(total, vi);
                rew[0] = __dadd_rn(rew[0], vi);
            }
            p.m_transfers[env] = __dadd_rn(p.m_transfers[env], total);
        }
        p.crossed[env] = crossed;
        p.meta[env] = 0x80000000u | (uint32_t)n_crossed | (done_mask << 8) | ((all_done ? 1u : 0u) << 16);
#pragma unroll
        for (int k = 0; k < SSD_MAXN; k++) {
            if (k >= n) continue;
            const size_t o = (size_t)env * n + k;
            io.rew[o] = rew[k];
            if (io.base_rew) io.base_rew[o] = base[k];
            if (io.transfers) io.transfers[o] = tr[k];
            if (io.info) {
                double4 r = make_double4(((just >> k) & 1u) ? 1.0 : 0.0, ((active >> k) & 1u) ? 1.0 : 0.0,
                                         k == first ? info2 : 0.0, k == first ? info3 : 0.0);
                reinterpret_cast<double4*>(io.info)[o] = r;
            }
            if (io.done) io.done[(size_t)env * (n + 1) + k] = (uint8_t)((done_mask >> k) & 1u);
        }
        if (io.done) io.done[(size_t)env * (n + 1) + n] = all_done ? 1 : 0;
    }
    const unsigned valid = __ballot_sync(0xffffffffu, mine);
    __syncwarp();
    car_write_obs(p, s_pos, s_vel, io.obs, env - lane, lane, valid);
}

__global__ void car_random_actions_kernel(const CarParams p, uint32_t step_index, uint32_t* counter, float lo, float hi,
                                          float* actions)
{
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (counter) step_index = *reinterpret_cast<volatile uint32_t*>(counter);
    if (env < p.E) {
        const uint32_t env_id = p.first_env_id + (uint32_t)env;
        for (int b = 0; b * 4 < p.n; b++) {
            Philox4 q = philox4x32_10((uint32_t)b, SITE_ACTIONS, step_index, 0u, p.seed, env_id);
            uint32_t w[4] = { q.x, q.y, q.z, q.w };
            for (int j = 0; j < 4 && b * 4 + j < p.n; j++)
                actions[(size_t)env * p.n + b * 4 + j] = lo + (hi - lo) * ((float)(w[j] >> 8) * (1.0f / 16777216.0f));
        }
    }
    if (counter) counter_finish(counter, step_index);
}

__global__ void car_get_state_kernel(const CarParams p, double* pos, double* vel, double* theta, double* m_transfers, int32_t* t)
{
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.E) return;
    for (int k = 0; k < p.n; k++) {
        if (pos) pos[(size_t)env * p.n + k] = p.pos[(size_t)k * p.E + env];
        if (vel) vel[(size_t)env * p.n + k] = p.vel[(size_t)k * p.E + env];
    }
    if (theta) theta[env] = p.theta[env];
    if (m_transfers) m_transfers[env] = p.m_transfers[env];
    if (t) t[env] = p.t[env];
}
__global__ void car_set_theta_kernel(const CarParams p, const double* theta)
{
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env < p.E) p.theta[env] = theta[env];
}
