// ssd_grid3.cuh — the gridworld decision logic with EIGHT LANES PER ENV (an "octet"; lane = agent), four envs per warp.
//
// Same function as grid_logic_kernel (ssd_grid2.cuh: update_moves map_env.py:483-676, consume :244-247, beams :678-814, and — cleanup
// — SeparateContractEnv.step two_stage_train.py:62-121), same inputs and outputs (the env's 128-byte hot line, the per-agent
// result words, rewards / infos / accumulators).  Why a second mapping: the thread-per-env kernel is bound by the latency of one
// thread's chain — E / 148 = 886 threads per SM is all the parallelism it has, 18 of 32 lanes are active on average, and with
// ~3500 SASS instructions it misses the instruction cache; removing a quarter of its instructions moved its time by 1 %
// (DESIGN.md §4.1).  With a lane per agent the per-agent work (decode, rotation, target cell, wall test, apple under the agent,
// ray occupancy, reward, transfers) is parallel and 8 x as many warps hide each other's latency.  It executes ~3 x the
// warp-instructions per env though (the sequential beams run four envs per warp-iteration instead of 32), so it wins only while
// the batch does not fill the GPU: 13.6 vs 15.8 us at 16384 envs, 75 vs 45 us at 131072 — ssd_create picks it for E <= 24576.
// How it works:
//   * contested moves: one match.any over the targets + one shuffle per agent decide whether the env needs the literal
//     reference ordering; the few % that do are resolved one env at a time by the whole warp (resolve_moves_slow, unchanged);
//   * consume: of co-located agents the lowest index eats — one match.any over the agents' cells;
//   * beams stay sequential per env (shuffled order, a ray stops at the waste an earlier beam left), but a shooter's beam is
//     evaluated by its octet: the static part is a table row (GridParams::beam_tab), the 15 waste bits are tested two per lane,
//     every lane tests its own agent against the rays (one PRMT), two octet OR-reductions give the masks;
//   * rewards: lane = agent; the float64 redistribution runs in the reference's order through shuffles.
#pragma once
#include "ssd_grid2.cuh"

#define L8_WARPS 8
#define L8_THREADS (L8_WARPS * 32)
#define L8_ENVS_PER_CTA (L8_WARPS * 4)
#define L8_OCT_WORDS 32                 // per octet: the hot line
#ifndef L8_MIN_BLOCKS
#define L8_MIN_BLOCKS 4
#endif

__device__ __forceinline__ uint32_t o8_or(uint32_t v)
{
    v |= __shfl_xor_sync(FULL, v, 1); v |= __shfl_xor_sync(FULL, v, 2); v |= __shfl_xor_sync(FULL, v, 4);
    return v;
}
__device__ __forceinline__ int o8_sum(int v)
{
    v += __shfl_xor_sync(FULL, v, 1); v += __shfl_xor_sync(FULL, v, 2); v += __shfl_xor_sync(FULL, v, 4);
    return v;
}
__device__ __forceinline__ uint32_t o8_min(uint32_t v)
{
    v = min(v, __shfl_xor_sync(FULL, v, 1)); v = min(v, __shfl_xor_sync(FULL, v, 2)); v = min(v, __shfl_xor_sync(FULL, v, 4));
    return v;
}
__device__ __forceinline__ double o8_shfl_f64(double v, int src)
{
    const int lo = __shfl_sync(FULL, __double2loint(v), src, 8), hi = __shfl_sync(FULL, __double2hiint(v), src, 8);
    return __hiloint2double(hi, lo);
}
// bit `idx` of a mask kept as consecutive words
__device__ __forceinline__ uint32_t hot_bit(const uint32_t* words, uint32_t idx) { return (words[idx >> 5] >> (idx & 31u)) & 1u; }

// count_apples_in_radius(5, loc) (explicit bounds): harvest_new.py:326-336, on the octet's apple mask
__device__ __noinline__ int g3_count_r5(const uint16_t* ci, const uint32_t* am, int row, int col, int H, int W, int Wp)
{
    int cnt = 0;
#pragma unroll 1
    for (int dr = -2; dr <= 2; dr++)
#pragma unroll
        for (int dc = -2; dc <= 2; dc++)
            if (dr * dr + dc * dc <= 5) {
                const int r = row + dr, c = col + dc;
                if ((unsigned)r < (unsigned)H && (unsigned)c < (unsigned)W) {
                    const uint32_t cw = ci[r * Wp + c];
                    cnt += (int)((cw >> 14) & hot_bit(am, cw & CI_IDX));
                }
            }
    return cnt;
}

template <int KIND>
__global__ void __launch_bounds__(L8_THREADS, L8_MIN_BLOCKS) grid_logic8_kernel(const GridParams p, const StepIO io, uint32_t* __restrict__ res_g)
{
    extern __shared__ __align__(16) uint8_t dsm[];
    pdl_launch_dependents();                                      // the observe kernel's CTAs may take the SM slots this grid frees
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, a = lane & 7, obase = lane & 24;
    const int n = p.n, H = p.H, W = p.W, Wp = p.Wp;
    const int ci_bytes = (H * Wp * 2 + 15) & ~15;
    const uint16_t* ci = reinterpret_cast<const uint16_t*>(dsm);
    for (int i = threadIdx.x; i < (ci_bytes >> 2); i += L8_THREADS)
        reinterpret_cast<uint32_t*>(dsm)[i] = __ldg(reinterpret_cast<const uint32_t*>(p.cell_info) + i);
    pdl_wait();                                                   // the actions (and the records) come from the preceding kernels
    __syncthreads();
    uint32_t* s_hot = reinterpret_cast<uint32_t*>(dsm + ci_bytes) + (warp * 4 + (lane >> 3)) * L8_OCT_WORDS;
    uint32_t* s_am = s_hot + RO_AMASK / 4;
    uint32_t* s_wm = s_hot + RO_WMASK / 4;
    const uint32_t raymask = (31u << RAY_L) | (31u << RAY_C) | (31u << RAY_R);

    const int ngroups = (p.E + L8_ENVS_PER_CTA - 1) / L8_ENVS_PER_CTA;
#pragma unroll 1
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int env_raw = (grp * L8_WARPS + warp) * 4 + (lane >> 3);
        const bool valid = env_raw < p.E;
        const int env = valid ? env_raw : p.E - 1;
        uint8_t* hdr = p.state + (size_t)env * p.rec_stride;       // this env's record
        {
            uint4 h4 = make_uint4(0u, 0u, 0u, 0u);
            if (valid) h4 = *reinterpret_cast<const uint4*>(hdr + 16 * a);
            *reinterpret_cast<uint4*>(s_hot + 4 * a) = h4;
        }
        const bool on = valid && a < n;                             // this lane is an agent of a real env
        int act = 4;
        if (on) act = io.actions[(size_t)env * n + a];
        __syncwarp();
        const int t = (int)s_hot[RO_T / 4] + 1;                     // map_env.py:230
        const uint32_t episode = s_hot[RO_EPISODE / 4];
        const double theta = __hiloint2double((int)s_hot[RO_THETA / 4 + 1], (int)s_hot[RO_THETA / 4]);
        const uint32_t flags = s_hot[RO_FLAGS / 4];
        int hcount = (int)s_hot[RO_HCOUNT / 4];
        const uint32_t v = s_hot[a];
        uint32_t err = 0;

        // ---- decode, rotations (map_env.py:514-516), candidate cells (Agent.py:8-16,161-162,198-199)
        const int row = (int)(v & 255u), col = (int)((v >> 8) & 255u);
        int ori = (int)((v >> 16) & 3u);
        bool mover = false, fire = false, clean = false, want = false;
        uint32_t cand = 0;
        if (on) {
            if (act <= 4) {
                mover = true;
                if (act < 4) {
                    // egocentric -> world direction: LEFT ori+3, RIGHT ori+1, UP ori, DOWN ori+2 (rotate_action :844-853)
                    const int d = (ori + ((0x2013 >> (4 * act)) & 3)) & 3;
                    const int nr = row + ori_dr(d), nc = col + ori_dc(d);
                    if ((unsigned)nr < (unsigned)H && (unsigned)nc < (unsigned)W) { want = true; cand = (uint32_t)nr | ((uint32_t)nc << 8); }
                }
            } else if (act == 5) ori = (ori + 1) & 3;                                // TURN_CLOCKWISE
            else if (act == 6) ori = (ori + 3) & 3;                                  // TURN_COUNTERCLOCKWISE
            else if (KIND == SSD_ENV_HARVEST) { if (act == 7) fire = true; else { err = 8; mover = true; } }
            else if (act == 7) clean = true;
            else if (act == 8) fire = true;
            else { err = 8; mover = true; }
        }
        const uint32_t cur = v & 0xFFFFu;
        uint32_t pos = cur | ((uint32_t)ori << 16);
        // return_valid_pos (Agent.py:111-119): walls are static
        const uint32_t tg = (want && !(ci[rc_off(cand, Wp)] & CI_WALL)) ? cand : cur;

        // ---- fast path test: no two movers share a target and no real move targets an occupied cell
        const bool real = mover && tg != cur;
        // (the match runs in every lane: not under a short-circuit)
        const unsigned same_tg = __match_any_sync(FULL, mover ? (((uint32_t)obase << 16) | tg) : (0x80000000u | (uint32_t)lane));
        bool slow = mover && __popc(same_tg) > 1;
#pragma unroll 1
        for (int b = 0; b < n; b++) {
            const uint32_t cb = __shfl_sync(FULL, cur, b, 8);
            if (real && b != a && cb == tg) slow = true;
        }
        const uint32_t slowm = __ballot_sync(FULL, slow);
        const uint32_t moverm = __ballot_sync(FULL, mover);
        if (!((slowm >> obase) & 0xFFu) && mover) pos = (pos & 0xFFFF0000u) | tg;
        // ---- contested moves: the literal reference ordering, one env at a time, lane = agent (lanes 0 .. n-1)
        if (slowm) {
#pragma unroll 1
            for (int r = 0; r < 4; r++) {
                if (!((slowm >> (8 * r)) & 0xFFu)) continue;
                const int src = 8 * r + a;
                const uint32_t cur_r = __shfl_sync(FULL, cur, src), tg_r = __shfl_sync(FULL, tg, src);
                const uint32_t movers_r = (moverm >> (8 * r)) & 0xFFu;
                const uint32_t ep_r = __shfl_sync(FULL, episode, 8 * r), t_r = __shfl_sync(FULL, (uint32_t)t, 8 * r);
                const uint32_t env_r = __shfl_sync(FULL, (uint32_t)env, 8 * r);
                int ao = 0, tgt = 0; bool has_move = false;
                if (lane < n) {
                    ao = (int)rc_lex(cur_r);                           // lexicographic (row, col) cell id
                    has_move = (movers_r >> lane) & 1u;
                    tgt = has_move ? (int)rc_lex(tg_r) : ao;
                }
                const uint32_t mval = has_move ? (uint32_t)tgt : (0xFFFF0000u | (uint32_t)lane);
                const unsigned same = __match_any_sync(FULL, mval);
                const bool contested = has_move && (__popc(same) > 1);
                const uint32_t rr = resolve_moves_slow(lane, n, p.seed, p.first_env_id + env_r, ep_r, t_r, ao, has_move, tgt, movers_r, contested);
                const uint32_t e2 = __reduce_or_sync(FULL, rr >> 16);
                const uint32_t back = __shfl_sync(FULL, rr, a);        // lane l of octet r takes the result of lane l & 7
                if ((lane >> 3) == r) {
                    if (a < n) pos = (pos & 0xFFFF0000u) | rc_lex(back & 0xFFFFu);
                    err |= e2;
                }
            }
        }

        // ---- cells under the agents: stale-list infos on the start-of-step map (cleanup_new.py:220-223,
        //      harvest_new.py:190-199), then consume (map_env.py:244-247): of co-located agents the lowest index eats
        const uint32_t cw = on ? (uint32_t)ci[rc_off(pos, Wp)] : 0u;
        const bool on_apple = on && (((cw >> 14) & hot_bit(s_am, cw & CI_IDX)) != 0u);
        const bool first = (__ffs(__match_any_sync(FULL, on ? (((uint32_t)obase << 16) | (pos & 0xFFFFu)) : (0x80000000u | (uint32_t)lane))) - 1) == lane;
        uint32_t res = 0;
        if (__any_sync(FULL, on_apple)) {
            if (on_apple) {
                if (!(flags & RF_STALE_EMPTY)) {
                    res = RS_EATEN;
                    if (KIND == SSD_ENV_HARVEST &&
                        g3_count_r5(ci, s_am, (int)(pos & 255u), (int)((pos >> 8) & 255u), H, W, Wp) < 4) res |= RS_EATEN_CLOSE;
                }
                if (first) res += 1u << RS_REWARD_SHIFT;
            }
            __syncwarp();                                           // the counts read the start-of-step mask
            if (on_apple && first) atomicAnd(&s_am[(cw & CI_IDX) >> 5], ~(1u << (cw & 31u)));
            __syncwarp();
        }

        // ---- beams in shuffled agent order (map_env.py:678-693); keys only matter when >= 2 agents of an env fire
        const bool shooter = fire || clean;
        const uint32_t shootm = __ballot_sync(FULL, shooter);
        int ncleaned = 0;
        if (shootm) {
            uint32_t rem = (shootm >> obase) & 0xFFu;
            const uint32_t cleanm = (__ballot_sync(FULL, clean) >> obase) & 0xFFu;
            const bool multi = (rem & (rem - 1u)) != 0u;
            uint32_t key = 0u;
            if (__any_sync(FULL, multi)) {
                if (multi && shooter) {
                    const uint4 q = draw_block_ool(p.seed, p.first_env_id + (uint32_t)env, episode, (uint32_t)t, SITE_BEAM_ORDER, (uint32_t)(a >> 2));
                    key = (a & 3) == 0 ? q.x : ((a & 3) == 1 ? q.y : ((a & 3) == 2 ? q.z : q.w));
                }
            }
#pragma unroll 1
            for (int it = 0; it < SSD_MAXN && __any_sync(FULL, rem != 0u); it++) {      // (an octet has at most 8 shooters)
                // the next shooter of the env: smallest key, lowest index on equal keys
                const bool mine_rem = (rem >> a) & 1u;
                const uint32_t kk = mine_rem ? (multi ? key : 0u) : 0xFFFFFFFFu;
                const uint32_t kmin = o8_min(kk);
                const uint32_t cm = (__ballot_sync(FULL, mine_rem && kk == kmin) >> obase) & 0xFFu;
                const int s = cm ? __ffs(cm) - 1 : -1;
                const bool go = s >= 0;
                if (go) rem &= ~(1u << s);
                const uint32_t sh = __shfl_sync(FULL, pos, go ? s : 0, 8);
                const bool cl = go && ((cleanm >> s) & 1u);
                // the static part of the beam: which ray cells are walls or outside the map, which are waste points
                const int srow = (int)(sh & 255u), scol = (int)((sh >> 8) & 255u), sori = (int)((sh >> 16) & 3u);
                const uint4* trow = p.beam_tab + ((size_t)(srow * Wp + scol) * 4 + sori) * 2;
                uint32_t wall = raymask, w0 = 0xFFFFFFFFu, w1 = 0xFFFFFFFFu, w2 = 0xFFFFFFFFu, w3 = 0xFFFFFFFFu;
                if (go) { const uint4 t0 = __ldg(trow); wall = t0.x; w0 = t0.y; w1 = t0.z; w2 = t0.w; }
                uint32_t wbits = 0u;
                if (cl) {
                    w3 = __ldg(trow + 1).x;
                    // 15 ray cells, lane a tests cells a and a + 8
                    const uint32_t ia = ((a < 4 ? w0 : w1) >> (8 * (a & 3))) & 255u;
                    const uint32_t ib = ((a < 4 ? w2 : w3) >> (8 * (a & 3))) & 255u;
                    const int pa = (a < 5 ? RAY_L : RAY_C - 5) + a;                       // cells 0..7: left 0-4, centre 0-2
                    const int kb = a + 8;
                    const int pb = (kb < 10 ? RAY_C - 5 : RAY_R - 10) + kb;               // cells 8..14: centre 3-4, right 0-4
                    if (ia != 255u) wbits |= hot_bit(s_wm, ia) << pa;
                    if (kb < 15 && ib != 255u) wbits |= hot_bit(s_wm, ib) << pb;
                }
                // agents on the rays (see g2_fire): every lane tests its own agent
                const uint32_t sel = sori == ORI_UP ? 0x3214u : (sori == ORI_DOWN ? 0x3250u : (sori == ORI_RIGHT ? 0x3201u : 0x3245u));
                uint32_t obit = 0u;
                if (go && on) {
                    const uint32_t x = (pos & 0xFFFFu) + (0x4040u - (sh & 0xFFFFu));
                    const uint32_t z = __byte_perm(x, 0x8080u - x, sel) - 0x3F40u;
                    if ((z & 0xFFFFFCF8u) == 0u) obit = (1u << ((z & 7u) | ((z >> 5) & 0x18u))) & raymask;
                }
                const uint32_t waste = o8_or(wbits), occ = o8_or(obit);
                const uint32_t stop = wall | waste | occ;
                int nup = 0;
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    const int shb = b == 0 ? RAY_L : (b == 1 ? RAY_C : RAY_R);
                    const uint32_t s5 = (stop >> shb) & 31u;
                    const uint32_t f = (s5 & (0u - s5)) << shb;                           // the ray's first stopping cell (0: none)
                    const bool eff = go && (f & ~wall) != 0u;
                    if (eff && (f & waste)) {                                              // CLEAN: H -> R (cleanup_new.py:285-290)
                        const int k = 5 * b + __ffs(f) - 1 - shb;
                        const uint32_t wk = k < 8 ? (k < 4 ? w0 : w1) : (k < 12 ? w2 : w3);
                        const uint32_t c = (wk >> (8 * (k & 3))) & 255u;
                        if (a == 0) s_wm[c >> 5] &= ~(1u << (c & 31u));
                        nup++;
                    }
                    // Agent.hit (Agent.py:224-226): the highest index wins duplicates (agent_by_pos)
                    const uint32_t hm = (__ballot_sync(FULL, eff && !cl && (f & occ) && obit == f) >> obase) & 0xFFu;
                    if (hm && a == 31 - __clz(hm)) res -= 50u << RS_REWARD_SHIFT;
                    if (p.beam && go && a == 0) {                                          // firing_points -> beam_pos (map_env.py:789,812)
                        const int dr = ori_dr(sori), dc = ori_dc(sori);
                        const int rr = ori_dr((sori + 1) & 3), rcl = ori_dc((sori + 1) & 3);
                        const int step = dr * Wp + dc, side = rr * Wp + rcl, base = srow * Wp + scol;
                        const int cnt = f ? __ffs(f) - 1 - shb + ((f & ~wall) ? 1 : 0) : 5; // cells up to and including a non-wall stop
                        uint8_t* beam = p.beam + (size_t)env * p.map_bytes;
                        for (int i = 0; i < cnt; i++)
                            beam[base + (b == 0 ? -side : (b == 1 ? step : side)) + i * step] = cl ? (uint8_t)'C' : (uint8_t)'F';
                    }
                }
                if (go && a == s) {
                    if (cl) res |= (uint32_t)nup;
                    else res -= 1u << RS_REWARD_SHIFT;                                     // fire cost (Agent.py:217-219)
                }
                if (cl) ncleaned += nup;
                __syncwarp();                                                              // the next beam sees this one's cleaning
            }
        }
        if (KIND == SSD_ENV_CLEANUP) hcount -= ncleaned;
        const uint32_t errall = o8_or(err);

        // ---- hot line back (the observe kernel reads agents, t, episode, flags, hcount, masks from it)
        __syncwarp();
        if (valid) {
            if (a < n) s_hot[a] = pos;
            if (a == 0) {
                s_hot[RO_T / 4] = (uint32_t)t;
                s_hot[RO_FLAGS / 4] = (flags & ~RF_STALE_EMPTY) | (errall ? (errall << RF_ERR_SHIFT) : 0u);
                s_hot[RO_HCOUNT / 4] = (uint32_t)hcount;
            }
        }
        __syncwarp();
        if (valid) {
            *reinterpret_cast<uint4*>(hdr + 16 * a) = *reinterpret_cast<const uint4*>(s_hot + 4 * a);
            res_g[(size_t)env * SSD_MAXN + a] = a < n ? res : 0u;
            if (KIND == SSD_ENV_CLEANUP && ncleaned && a == 0) red_add(reinterpret_cast<uint32_t*>(hdr + RO_DIRT), (uint32_t)ncleaned);
        }

        // ---- cleanup: nothing below depends on the spawn, so rewards / outputs are finished here (lane = agent).
        //      contract_list.py:22-27; two_stage_train.py:72-99; cleanup_new.py:213-253
        if (KIND == SSD_ENV_CLEANUP) {
            const size_t o = (size_t)env * n + a;
            const int reward = (int)res >> RS_REWARD_SHIFT;
            const uint32_t cleaned = res & RS_CLEANED_MASK, eaten = (res >> 2) & 1u;
            double rj = (double)reward;
            if (p.reward_mode) {                               // shaped env rewards (map_env.py:289-301) and their f64 episode sums
                const int coll = o8_sum(on ? reward : 0);
                const bool collective = p.reward_mode & RM_COLLECTIVE;
                const int ra = collective ? coll : reward;
                int sp = 0, sn = 0;
#pragma unroll 1
                for (int j = 0; j < n; j++) {
                    const int d = (collective ? coll : __shfl_sync(FULL, reward, j, 8)) - ra;
                    if (d > 0) sp += d; else sn += d;
                }
                rj = (p.reward_mode & RM_INEQUITY) ? inequity_term(ra, sp, sn, p.alpha, p.beta, n) : (double)ra;
                double raw_step = 0.0;                         // raw_rewards = ((0 + r0) + r1) + ... (cleanup_new.py:228-232)
#pragma unroll 1
                for (int j = 0; j < n; j++) raw_step = __dadd_rn(raw_step, o8_shfl_f64(rj, j));
                if (on) {
                    double* xs = reinterpret_cast<double*>(hdr + RO_XSUM) + a;
                    double* xt = reinterpret_cast<double*>(hdr + RO_XTSUM) + a;
                    *xs = __dadd_rn(*xs, rj);
                    *xt = __dadd_rn(*xt, __dmul_rn((double)(t - 1), rj));
                    if (a == 0) { double* xr = reinterpret_cast<double*>(hdr + RO_XRAW); *xr = __dadd_rn(*xr, raw_step); }
                }
            }
            if (on && io.base_rew) io.base_rew[o] = rj;
            double tr = 0.0, total_tr = 0.0;
            if (on && p.contract == SSD_CONTRACT_CLEANUP) tr = __dmul_rn(-theta, (double)cleaned);
            if (on && io.transfers) io.transfers[o] = tr;
            // a zero transfer leaves every reward bit-identical (rewards are never -0.0): the loop runs only if somebody pays
            if (p.contract != SSD_CONTRACT_NONE && __any_sync(FULL, tr != 0.0)) {
                const double share = __ddiv_rn(tr, (double)(n - 1));
#pragma unroll 1
                for (int i = 0; i < n; i++) {
                    const double tri = o8_shfl_f64(tr, i), shi = o8_shfl_f64(share, i);
                    if (tri != 0.0) {
                        rj = __dadd_rn(rj, i == a ? -tri : shi);
                        total_tr = __dadd_rn(total_tr, tri);
                    }
                }
            }
            const uint32_t n_eaten = (uint32_t)o8_sum((int)eaten);
            if (on) {
                if (io.rew) io.rew[o] = rj;
                if (io.info) reinterpret_cast<uint32_t*>(io.info)[o] = eaten | (cleaned << 8);
                // episode accumulators: fire-and-forget reductions (one add per address and step: a float64 RED rounds exactly
                // like `*p = *p + x`; zero terms are skipped, sums are never -0.0)
                if (reward != 0 && !p.reward_mode) {
                    red_add(reinterpret_cast<int*>(hdr + RO_SUM_RAW) + a, reward);
                    red_add(reinterpret_cast<long long*>(hdr + RO_TSUM_RAW) + a, (long long)(t - 1) * reward);
                }
                if (p.contract != SSD_CONTRACT_NONE && rj != 0.0) {
                    red_add(reinterpret_cast<double*>(hdr + RO_SUM_TR) + a, rj);
                    red_add(reinterpret_cast<double*>(hdr + RO_TSUM_TR) + a, __dmul_rn((double)(t - 1), rj));
                }
                if (cleaned) red_add(reinterpret_cast<uint32_t*>(hdr + RO_AGENT_A) + a, cleaned);
            }
            if (valid && a == 0) {
                if (n_eaten) red_add(reinterpret_cast<uint32_t*>(hdr + RO_APPLES), n_eaten);
                if (total_tr != 0.0) red_add(reinterpret_cast<double*>(hdr + RO_TRANSFERS), total_tr);
                if (io.done) io.done[env] = t == p.horizon ? 1 : 0;
            }
            if (io.c_rew8) {                                   // compact result block (ssd_step_host_async)
                int vi = 0;
                const bool fits = reward_fits_i8(rj, vi);
                if (on) io.c_rew8[o] = (int8_t)(fits ? vi : -128);
                if (valid && a == 0) io.c_done[env] = t == p.horizon ? 1 : 0;
                const uint32_t badm = __ballot_sync(FULL, on && !fits);
                if (badm) {                                    // one record per env with such a reward, one atomic per warp
                    const bool need = ((badm >> obase) & 0xFFu) != 0u;
                    const uint32_t needm = __ballot_sync(FULL, need && a == 0);
                    const int leader = __ffs(needm) - 1;
                    uint32_t base = 0u;
                    if (lane == leader) base = atomicAdd(io.c_count, (uint32_t)__popc(needm));
                    base = __shfl_sync(FULL, base, leader);
                    if (need) {
                        uint8_t* rec = io.c_rec + (size_t)(base + (uint32_t)__popc(needm & ((1u << obase) - 1u))) * (size_t)(8 + 8 * n);
                        if (a == 0) *reinterpret_cast<int2*>(rec) = make_int2(env, 0);
                        if (a < n) reinterpret_cast<double*>(rec + 8)[a] = rj;
                    }
                }
            }
        }
        __syncwarp();                                           // the octet's hot line is rewritten by the next round
    }
}
