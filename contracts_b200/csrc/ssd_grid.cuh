// ssd_grid.cuh — sm_100a kernels for the gridworld envs (cleanup_new / harvest_new).
//
// Mapping: one WARP per environment, one LANE per agent for the move /
// rotate / conflict logic (n <= 8, decisions are serial per env, so warp ballots + shuffles
// give the reference's ordering without block barriers); all 32 lanes for the data-parallel
// phases (tile expand, spawn scans, observation gather).  The map lives as a padded uint8
// tile in shared memory (7 cells of C_OUTSIDE border = the view radius), agents are
// painted into the high nibble.  All bulk HBM traffic goes through the bulk-copy (TMA) engine:
// each warp prefetches the NEXT env's whole record (map + header) into a double-buffered
// shared slot with cp.async.bulk + an mbarrier while it works on the current one, writes the
// updated record back with one bulk store, and the rotated 15x15x3 windows are gathered
// row-wise (4 output rows = 180 B per lane, conflict-free word stores) into a per-warp staging
// buffer that leaves through one more bulk store per env.
//
// Reference behaviour restated here (paths relative to the reference root):
//   environments/map_env.py   step :216-304, update_moves :483-676, update_custom_moves :678-693,
//                             update_map_fire :721-814, color_view :397-411, reset :306-342,
//                             spawn_point :816-827, spawn_rotation :829-832
//   environments/cleanup_new.py  step :211-267, custom_action :269-292, spawn_apples_and_waste :322-349,
//                             compute_probabilities :351-368, custom_reset :171-189
//   environments/harvest_new.py  step :181-239, spawn_apples :284-317, count_apples_in_radius :326-336
//   contract/contract_list.py :22-27, :45-54 ; environments/two_stage_train.py :62-121, :159-187
#pragma once
#include "ssd_common.cuh"

#ifndef GRID_WARPS
#define GRID_WARPS 8
#endif
#define GRID_THREADS (GRID_WARPS * 32)
#define GRID_MIN_BLOCKS 3               // 24 warps / SM: registers <= 80, ~9 KB shared per warp
#define SCRATCH_DRAWS 256               // u32 draws per warp scratch
#define SCRATCH_KEYS 256                // u32 shuffle keys per warp scratch
#define MAX_POINT_ROUNDS 8              // point lists are scanned 32 per round => <= 256 points
#define FULL 0xffffffffu

// record layout after the map bytes (all offsets relative to rec + map_bytes)
#define RO_AGENTS 0        // u32[8]: row | col << 8 | ori << 16
#define RO_T 32            // i32
#define RO_EPISODE 36      // u32
#define RO_THETA 40        // f64
#define RO_FLAGS 48        // u32
#define RO_HCOUNT 52       // u32 (#waste cells, cleanup)
#define RO_APPLES 56       // u32 total_apples_eaten
#define RO_LOWDENS 60      // u32 low_density_apples_eaten
#define RO_DIRT 64         // u32 dirt_cleaned
#define RO_TRANSFERS 72    // f64 metrics['transfers']
#define RO_SUM_TR 80       // f64[8]
#define RO_TSUM_TR 144     // f64[8]
#define RO_TSUM_RAW 208    // i64[8]
#define RO_SUM_RAW 272     // i32[8]
#define RO_AGENT_A 304     // u32[8] waste_cleaned | apples_consumed
#define RO_AGENT_B 336     // u32[8] close_apples_consumed
#define RO_SIZE 368
// extension, present only when rewards are shaped (use_collective_reward / inequity_averse_reward,
// map_env.py:289-301): the env rewards are float64 then, so their episode sums are too
#define RO_XSUM 368        // f64[8] sum_t r
#define RO_XTSUM 432       // f64[8] sum_t t * r
#define RO_XRAW 496        // f64 metrics['raw_env_rewards']
#define RO_XSIZE 512
#define RM_COLLECTIVE 1
#define RM_INEQUITY 2

struct GridParams {
    int E, n, H, W, Wp, S, TH;
    int wpw;                 // words per map row (Wp / 4)
    uint32_t wpw_magic;      // ceil(65536 / wpw): q = (w * magic) >> 16 for w < 65536 / wpw
    int map_bytes;           // H * Wp rounded up to 16
    int rec_stride;          // map_bytes + hdr_bytes (multiple of 16: one bulk copy moves a record)
    int hdr_bytes;           // RO_SIZE, or RO_XSIZE when rewards are shaped
    int reward_mode;         // RM_* bits
    double alpha, beta;      // inequity aversion weights (map_env.py:71-72)
    // shared memory layout.  CTA tables first, then one region per warp:
    //   [tile | rec slot 0 | rec slot 1 | stage (obs staging, aliased by the spawn scratch) | misc]
    int tile_r16, stage_r16, warp_bytes, off_rec, off_stage, off_misc;
    int sm_thr, sm_won, sm_apple, sm_waste, sm_apple_rc, sm_waste_rc, sm_warp0, smem_bytes;
    int obs_items;           // ceil(15 n / 4): 4-row (180 B) work items of the observation gather
    // observe kernel (ssd_grid2.cuh), per warp: [tile | stage | misc]
    int g2_stage, g2_misc, g2_warp_bytes, g2_smem_bytes;
    int S2, hpw, tile2_off;  // transposed tile: row stride 8 + Hp + 8, words per transposed row (Hp / 4), byte offset from T
    int kind, contract, horizon;
    int n_apple, n_waste, n_spawn, n_waste_start, F;
    uint32_t seed, first_env_id;
    double theta_low, theta_high, null_prob;
    uint32_t thr_harvest[4]; // SPAWN_PROB thresholds (harvest_new.py:34)
    uint32_t thr_waste;      // wasteSpawnProbability = 0.5
    const uint32_t* pal;     // [16] packed RGB per tile byte value (cell codes 0..5, agents 6..13, outside 15)
    uint32_t s_magic;        // ceil(2^32 / S): row = umulhi(offset, s_magic)
    const uint16_t* apple_pts;
    const uint16_t* waste_pts;
    const uint16_t* spawn_pts;
    const uint16_t* apple_rc;    // row << 8 | col of each apple point (feature_obs only)
    const uint16_t* waste_rc;
    const uint32_t* thr_apple;   // [n_waste + 1] apple spawn threshold by #waste
    const uint8_t* waste_on;     // [n_waste + 1]
    const uint8_t* reset_map;    // [map_bytes] initial codes (walls + custom_reset)
    uint8_t* state;
    uint8_t* beam;               // optional [E][map_bytes]: beam_pos of the last step as a char overlay (render only)
    double* stats;               // optional [8]: ssd_set_episode_stats accumulator (added to by the reset kernel)
};

// use_collective_reward / inequity_averse_reward (map_env.py:289-301): the shaped reward of agent a from the
// integer env rewards.  Differences and their partial sums are exact integers; alpha * sum, beta * sum, their
// sum, the division by (num_agents - 1) and the subtraction are float64, in that order.  (Scalar arguments and
// no local arrays: a GridParams reference into a non-inlined function would copy the kernel parameters to the stack.)
__device__ __forceinline__ double inequity_term(int ra, int sp, int sn, double alpha, double beta, int n)
{
    const double dis = __dmul_rn(alpha, (double)sp), adv = __dmul_rn(beta, (double)sn);
    return __dsub_rn((double)ra, __ddiv_rn(__dadd_rn(dis, adv), (double)(n - 1)));
}
// thread per env: the int16 rewards sit in the RS_* words rsp[j * rstride] (bits 16-31)
__device__ __noinline__ double shaped_reward(int mode, double alpha, double beta, int n, const uint32_t* rsp, int rstride, int a)
{
    int coll = 0;
    for (int j = 0; j < n; j++) coll += (int)rsp[j * rstride] >> 16;
    const bool collective = mode & RM_COLLECTIVE;
    const int ra = collective ? coll : ((int)rsp[a * rstride] >> 16);
    if (!(mode & RM_INEQUITY)) return (double)ra;
    int sp = 0, sn = 0;
    for (int j = 0; j < n; j++) {
        const int d = (collective ? coll : ((int)rsp[j * rstride] >> 16)) - ra;
        if (d > 0) sp += d; else sn += d;
    }
    return inequity_term(ra, sp, sn, alpha, beta, n);
}
// warp per env, lane = agent (every lane calls)
__device__ __noinline__ double shaped_reward_warp(int mode, double alpha, double beta, int n, int reward)
{
    int coll = 0;
    for (int j = 0; j < n; j++) coll += __shfl_sync(FULL, reward, j);
    const bool collective = mode & RM_COLLECTIVE;
    const int ra = collective ? coll : reward;
    int sp = 0, sn = 0;
    for (int j = 0; j < n; j++) {
        const int d = (collective ? coll : __shfl_sync(FULL, reward, j)) - ra;
        if (d > 0) sp += d; else sn += d;
    }
    return (mode & RM_INEQUITY) ? inequity_term(ra, sp, sn, alpha, beta, n) : (double)ra;
}

// per-warp misc area
#define MISC_VDESC 0        // int4[8]: per-agent view descriptor {offset of out[0][0], pixel step, row step, 0}
#define MISC_MBAR 128       // u64[2]: record-prefetch mbarriers
#define MISC_BYTES 144

struct StepIO {
    const uint8_t* actions;
    uint8_t* obs; long long obs_stride;
    double* rew; double* base_rew; double* transfers;
    uint8_t* info; double* feat; uint8_t* done;
    // compact result block of ssd_step_host_async (all null otherwise): int8 rewards, dones, and records
    // { int32 env; int32 0; double rew[n] } for the envs whose rewards are not all integers in [-127, 127]
    int8_t* c_rew8; uint8_t* c_done; uint32_t* c_count; uint8_t* c_rec;
};

// is `r` exactly an integer in [-127, 127]?  (bit pattern compared: -0.0 is not)
__device__ __forceinline__ bool reward_fits_i8(double r, int& v)
{
    v = __double2int_rn(r);
    return v >= -127 && v <= 127 && __double_as_longlong((double)v) == __double_as_longlong(r);
}
// slot of this thread's record in the compact block: one atomic per group of converged threads
__device__ __forceinline__ uint32_t compact_slot(uint32_t* count)
{
    const unsigned am = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(am) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(count, (uint32_t)__popc(am));
    base = __shfl_sync(am, base, leader);
    return base + (uint32_t)__popc(am & ((1u << lane) - 1u));
}

struct EnvRng { uint32_t seed, env_id, episode, t; };

// per-CTA copies of the static tables in shared memory
struct SharedTables {
    const uint32_t* pal; const uint32_t* thr_apple; const uint8_t* waste_on;
    const uint16_t* apple; const uint16_t* waste; const uint16_t* apple_rc; const uint16_t* waste_rc;
};

__device__ __forceinline__ uint32_t lanemask_lt(int lane) { return (1u << lane) - 1u; }
__device__ __forceinline__ int dir_delta(int ori, int S)
{
    return ori == ORI_UP ? -S : (ori == ORI_RIGHT ? 1 : (ori == ORI_DOWN ? S : -1));
}

// ---------------------------------------------------------------------------------------------
// tile <-> HBM record
__device__ __forceinline__ void tile_load(const GridParams& p, const uint8_t* rec, uint8_t* tile, int lane)
{
    const uint4* src = reinterpret_cast<const uint4*>(rec);
    uint32_t* tw = reinterpret_cast<uint32_t*>(tile);
    const int S4 = p.S >> 2, nvec = p.map_bytes >> 4, nwords = p.H * p.wpw;
    for (int v = lane; v < nvec; v += 32) {
        uint4 q = __ldg(src + v);
        uint32_t w4[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int w = v * 4 + j;
            if (w < nwords) {
                int row = (int)(((uint32_t)w * p.wpw_magic) >> 16);
                int cw = w - row * p.wpw;
                tw[(row + SSD_VIEW) * S4 + 2 + cw] = w4[j];
            }
        }
    }
}
__device__ __forceinline__ void tile_store(const GridParams& p, uint8_t* rec, const uint8_t* tile, int lane)
{
    uint4* dst = reinterpret_cast<uint4*>(rec);
    const uint32_t* tw = reinterpret_cast<const uint32_t*>(tile);
    const int S4 = p.S >> 2, nvec = p.map_bytes >> 4, nwords = p.H * p.wpw;
    for (int v = lane; v < nvec; v += 32) {
        uint32_t w4[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int w = v * 4 + j;
            uint32_t x = TILE_FILL4;
            if (w < nwords) {
                int row = (int)(((uint32_t)w * p.wpw_magic) >> 16);
                int cw = w - row * p.wpw;
                x = tw[(row + SSD_VIEW) * S4 + 2 + cw] & CODE_MASK4;   // strip agent paint / occupancy
            }
            w4[j] = x;
        }
        dst[v] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
}

// shared-memory record slot <-> padded tile (step kernel; the record arrives / leaves by bulk copy).
// Lane l moves map words l, l + 32, ...; the tile word of each is loop-invariant across envs, so the
// first MAP_REG_WORDS of them are computed once per warp and kept in registers.
#define MAP_REG_WORDS 6
struct MapWords { int tw[MAP_REG_WORDS]; };
__device__ __forceinline__ int map_tile_word(const GridParams& p, int w)
{
    int row = (int)(((uint32_t)w * p.wpw_magic) >> 16);
    return (row + SSD_VIEW) * (p.S >> 2) + 2 + (w - row * p.wpw);
}
__device__ __forceinline__ MapWords map_words_init(const GridParams& p, int lane)
{
    MapWords m;
#pragma unroll
    for (int k = 0; k < MAP_REG_WORDS; k++) m.tw[k] = map_tile_word(p, lane + 32 * k);
    return m;
}
__device__ __forceinline__ void tile_expand(const GridParams& p, const MapWords& m, const uint8_t* recbuf, uint8_t* tile, int lane)
{
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(recbuf);
    uint32_t* tw = reinterpret_cast<uint32_t*>(tile);
    const int nwords = p.H * p.wpw;
#pragma unroll
    for (int k = 0; k < MAP_REG_WORDS; k++)
        if (lane + 32 * k < nwords) tw[m.tw[k]] = rw[lane + 32 * k];
    for (int w = lane + 32 * MAP_REG_WORDS; w < nwords; w += 32) tw[map_tile_word(p, w)] = rw[w];
}
__device__ __forceinline__ void tile_compress(const GridParams& p, const MapWords& m, uint8_t* recbuf, const uint8_t* tile, int lane)
{
    uint32_t* rw = reinterpret_cast<uint32_t*>(recbuf);
    const uint32_t* tw = reinterpret_cast<const uint32_t*>(tile);
    const int nwords = p.H * p.wpw;
#pragma unroll
    for (int k = 0; k < MAP_REG_WORDS; k++)
        if (lane + 32 * k < nwords) rw[lane + 32 * k] = tw[m.tw[k]] & CODE_MASK4;   // strip agent paint / occupancy
    for (int w = lane + 32 * MAP_REG_WORDS; w < nwords; w += 32) rw[w] = tw[map_tile_word(p, w)] & CODE_MASK4;
}

// ---------------------------------------------------------------------------------------------
// bulk-copy (TMA) + mbarrier primitives.  Bulk async-groups are per thread: lane 0 issues and waits.
__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, completion signalled on `bar` (expect_tx armed by the same lane)
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, void* bar)
{
    const uint32_t b = smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
// shared -> global, one bulk group per call
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int PENDING>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(PENDING) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One out-of-line copy of the 10 Philox rounds for the step kernel (the inlined form is ~80
// instructions per call site; the hot loop has to fit the instruction cache).
__device__ __noinline__ uint4 draw_block_ool(uint32_t seed, uint32_t env_id, uint32_t episode, uint32_t t,
                                             uint32_t site_call, uint32_t block)
{
    Philox4 q = philox4x32_10(block, site_call, t, episode, seed, env_id);
    return make_uint4(q.x, q.y, q.z, q.w);
}
__device__ __forceinline__ uint32_t draw_u32_ool(const EnvRng& g, uint32_t t, uint32_t site, uint32_t idx)
{
    uint4 q = draw_block_ool(g.seed, g.env_id, g.episode, t, site, idx >> 2);
    const uint32_t w = idx & 3u;
    return w == 0 ? q.x : (w == 1 ? q.y : (w == 2 ? q.z : q.w));
}

// Two Philox blocks of the same (env, episode, t) with interleaved rounds: two independent dependency chains instead of
// two calls back to back (the spawn is latency-bound).  Results go to shared memory; a null pointer skips the store.
__device__ __noinline__ void draw_blocks2_ool(uint32_t seed, uint32_t env_id, uint32_t episode, uint32_t t,
                                              uint32_t site_a, uint32_t block_a, uint4* dst_a,
                                              uint32_t site_b, uint32_t block_b, uint4* dst_b)
{
    uint32_t a0 = block_a, a1 = site_a, a2 = t, a3 = episode, b0 = block_b, b1 = site_b, b2 = t, b3 = episode;
    uint32_t k0 = seed, k1 = env_id;
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint32_t ah0 = __umulhi(0xD2511F53u, a0), al0 = 0xD2511F53u * a0, ah1 = __umulhi(0xCD9E8D57u, a2), al1 = 0xCD9E8D57u * a2;
        const uint32_t bh0 = __umulhi(0xD2511F53u, b0), bl0 = 0xD2511F53u * b0, bh1 = __umulhi(0xCD9E8D57u, b2), bl1 = 0xCD9E8D57u * b2;
        const uint32_t na0 = ah1 ^ a1 ^ k0, na2 = ah0 ^ a3 ^ k1, nb0 = bh1 ^ b1 ^ k0, nb2 = bh0 ^ b3 ^ k1;
        a0 = na0; a1 = al1; a2 = na2; a3 = al0;
        b0 = nb0; b1 = bl1; b2 = nb2; b3 = bl0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    if (dst_a) *dst_a = make_uint4(a0, a1, a2, a3);
    if (dst_b) *dst_b = make_uint4(b0, b1, b2, b3);
}

// ---------------------------------------------------------------------------------------------
// The literal ordering of the reference (rare: some movers share a target or a real move targets an
// occupied cell).  Out of line: it is cold and would otherwise bloat the hot loop's instruction
// footprint.  Returns err << 16 | new tile offset.
__device__ __noinline__ uint32_t resolve_moves_slow(int lane, int n, uint32_t seed, uint32_t env_id, uint32_t episode,
                                                    uint32_t t, int ao, bool has_move, int tgt, unsigned movers,
                                                    bool contested)
{
    const bool act_lane = lane < n;
    uint32_t err = 0;
    // shuffled priority (map_env.py:545-547): key of list position m = rank among movers
    uint32_t key = 0xFFFFFFFFu;
    if (has_move) key = draw_u32(seed, env_id, episode, t, SITE_MOVE_ORDER, 0,
                                 (uint32_t)__popc(movers & lanemask_lt(lane)));
    int cur_mv = tgt;                               // agent_moves[agent]
    // contested cells in lexicographic (row, col) order == increasing tile offset (np.unique, :548)
    unsigned remaining = __ballot_sync(FULL, contested);
    while (remaining) {
        uint32_t v = ((remaining >> lane) & 1u) ? (uint32_t)tgt : 0xFFFFFFFFu;
        int cmin = (int)__reduce_min_sync(FULL, v);
        bool in_group = has_move && tgt == cmin;
        unsigned group = __ballot_sync(FULL, in_group);
        remaining &= ~group;
        unsigned occm = __ballot_sync(FULL, act_lane && ao == cmin);   // live positions (:574)
        bool cell_free = true;
        if (occm) {
            int occ = 31 - __clz(occm);             // agent_by_pos: the highest index wins duplicates
            bool occ_has = (movers >> occ) & 1u;
            int occ_mv = __shfl_sync(FULL, cur_mv, occ);
            unsigned swapm = __ballot_sync(FULL, in_group && ao == occ_mv);
            if (!occ_has || occ_mv == cmin) cell_free = false;      // conditions (1),(2) :585-595
            else if (swapm) cell_free = false;                      // condition (3) :599-605
        }
        if (cell_free) {
            uint32_t kmin = __reduce_min_sync(FULL, in_group ? key : 0xFFFFFFFFu);
            unsigned cand = __ballot_sync(FULL, in_group && key == kmin);
            if (lane == __ffs(cand) - 1) ao = cmin;                 // first contestant of the shuffled list (:610)
        }
        if (in_group) cur_mv = ao;                                   // :620-621
    }
    // remaining moves, multi-pass (:624-676)
    unsigned live = movers;
    while (live) {
        const int snap = ao;                        // agent_by_pos snapshot (:625)
        const unsigned copy = live;
        const int num = __popc(live);
        for (int i = 0; i < n; i++) {
            if (!((copy >> i) & 1u) || !((live >> i) & 1u)) continue;
            int mv_i = __shfl_sync(FULL, cur_mv, i);
            int pos_i = __shfl_sync(FULL, ao, i);
            unsigned occ_live = __ballot_sync(FULL, act_lane && ao == mv_i);
            if (occ_live) {
                unsigned snap_b = __ballot_sync(FULL, act_lane && snap == mv_i);
                if (!snap_b) { err |= 2; live &= ~(1u << i); continue; }      // KeyError in the reference
                int occ = 31 - __clz(snap_b);
                int pos_occ = __shfl_sync(FULL, ao, occ);
                int mv_occ_raw = __shfl_sync(FULL, cur_mv, occ);
                int mv_occ = ((live >> occ) & 1u) ? mv_occ_raw : pos_occ;      // agent_moves.get(occ, pos)
                if (occ == i) live &= ~(1u << i);
                else if (!((copy >> occ) & 1u) || pos_occ == mv_occ) live &= ~(1u << i);
                else if (mv_occ == pos_i && mv_i == pos_occ) live &= ~((1u << i) | (1u << occ));
            } else {
                if (lane == i) ao = mv_i;
                live &= ~(1u << i);
            }
        }
        if (__popc(live) == num) {                  // no progress: move everyone left (:673-676)
            if ((live >> lane) & 1u) ao = cur_mv;
            break;
        }
    }
    return (err << 16) | (uint32_t)ao;
}

// update_moves (map_env.py:483-676).  Lane i < n holds agent i: `ao` tile offset of its cell.
// Returns error bits.  On the fast path (no two movers share a target and no real move targets an
// occupied cell) every mover simply moves, which is what the reference's first while-pass does.
__device__ __forceinline__ uint32_t resolve_moves(int lane, int n, uint8_t* tile, const EnvRng& g,
                                                  int& ao, bool has_move, int tgt)
{
    const bool act_lane = lane < n;
    const unsigned movers = __ballot_sync(FULL, has_move);
    if (movers == 0) return 0;
    uint32_t mval = has_move ? (uint32_t)tgt : (0xFFFF0000u | (uint32_t)lane);
    unsigned same = __match_any_sync(FULL, mval);
    bool contested = has_move && (__popc(same) > 1);
    if (act_lane) tile[ao] |= OCC_BIT;             // co-located lanes write the same value
    __syncwarp();
    bool real = has_move && tgt != ao;
    bool occupied = real && (tile[tgt] & OCC_BIT);
    __syncwarp();
    if (act_lane) tile[ao] &= 0x7F;
    __syncwarp();
    if (__ballot_sync(FULL, contested || occupied) == 0) {
        if (real) ao = tgt;
        return 0;
    }
    const uint32_t r = resolve_moves_slow(lane, n, g.seed, g.env_id, g.episode, g.t, ao, has_move, tgt, movers, contested);
    ao = (int)(r & 0xFFFFu);
    return r >> 16;
}

// ---------------------------------------------------------------------------------------------
// update_custom_moves + update_map_fire (map_env.py:678-693, 721-814).  Occupancy bits must be set.
// cls: 0 none, 1 FIRE, 2 CLEAN.  Returns #waste cells cleaned by all agents this step.
__device__ __forceinline__ int fire_beams(int lane, int n, int S, uint8_t* tile, const EnvRng& g,
                                          int ao, int ori, int cls, int& reward, int& cleaned)
{
    const bool act_lane = lane < n;
    unsigned rem = __ballot_sync(FULL, act_lane && cls != 0);
    if (!rem) return 0;
    int total_cleaned = 0;
    // shuffled firing order (map_env.py:684-685); irrelevant (and not drawn) when a single agent fires
    uint32_t bkey = (uint32_t)lane;
    if (rem & (rem - 1))
        bkey = act_lane ? draw_u32_ool(g, g.t, SITE_BEAM_ORDER, (uint32_t)lane) : 0xFFFFFFFFu;
    const bool ray_lane = lane < 15;
    const int b = lane / 5, k = lane - 5 * b;
    while (rem) {
        bool mine = (rem >> lane) & 1u;
        uint32_t kmin = __reduce_min_sync(FULL, mine ? bkey : 0xFFFFFFFFu);
        unsigned cand = __ballot_sync(FULL, mine && bkey == kmin);
        const int s = __ffs(cand) - 1;
        rem &= ~(1u << s);
        const int so = __shfl_sync(FULL, ao, s), sori = __shfl_sync(FULL, ori, s);
        const bool sclean = __shfl_sync(FULL, cls, s) == 2;
        const int d = dir_delta(sori, S), rt = dir_delta((sori + 1) & 3, S);
        int cell = so + (b == 1 ? rt - d : (b == 2 ? -rt - d : 0)) + (k + 1) * d;
        uint32_t code = ray_lane ? tile[cell] : (uint32_t)C_WALL;
        uint32_t c = code & CODE_MASK;
        bool pass = c != C_WALL && c != C_OUTSIDE;
        bool agent_here = (code & OCC_BIT) != 0;
        bool isH = c == C_WASTE;
        unsigned stopm = __ballot_sync(FULL, ray_lane && (!pass || agent_here || (sclean && isH)));
        unsigned raybits = (stopm >> (5 * b)) & 31u;
        int first = raybits ? __ffs(raybits) - 1 : 5;
        bool covered = ray_lane && pass && k <= first;
        bool upd = covered && sclean && isH;
        unsigned updm = __ballot_sync(FULL, upd);
        unsigned hitm = __ballot_sync(FULL, covered && agent_here && !sclean);
        if (upd) tile[cell] = (uint8_t)((code & OCC_BIT) | C_RIVER);
        while (hitm) {                               // Agent.hit(b"F"): -50 (Agent.py:178-180,224-226)
            int hl = __ffs(hitm) - 1; hitm &= hitm - 1;
            int hc = __shfl_sync(FULL, cell, hl);
            unsigned victims = __ballot_sync(FULL, act_lane && ao == hc);
            if (lane == 31 - __clz(victims)) reward -= 50;
        }
        int nclean = __popc(updm);
        if (lane == s) { if (sclean) cleaned = nclean; else reward -= 1; }
        total_cleaned += nclean;
        __syncwarp();
    }
    return total_cleaned;
}

// ---------------------------------------------------------------------------------------------
// cleanup spawn (cleanup_new.py:322-349).  Occupancy bits must be set.  `t` = 0 at reset.
// Updates hcount (#waste cells); returns whether the map changed.
__device__ __forceinline__ bool cleanup_spawn_active(const SharedTables& tb, int hcount)
{
    return tb.thr_apple[hcount] != 0 || tb.waste_on[hcount] != 0;
}
template <int ROUNDS>
__device__ __forceinline__ bool cleanup_spawn(const GridParams& p, const SharedTables& tb, int lane, uint8_t* tile,
                                              uint32_t* scratch, const EnvRng& g, uint32_t t, int& hcount)
{
    const uint16_t* sm_apple = tb.apple; const uint16_t* sm_waste = tb.waste;
    const uint32_t thrA = tb.thr_apple[hcount];
    const bool waste_on = tb.waste_on[hcount] != 0;
    if (thrA == 0 && !waste_on) return false;
    bool spawned = false;
    uint32_t* draws = scratch;
    uint32_t* keys = scratch + SCRATCH_DRAWS;
    // apples: eligible = no agent there and not already 'A'; draw index = rank among eligible (:328-335)
    unsigned eligm[ROUNDS];
    int M = 0;
#pragma unroll
    for (int q = 0; q < ROUNDS; q++) {
        int j = lane + 32 * q;
        bool e = false;
        if (q * 32 < p.n_apple) {
            if (j < p.n_apple) { uint32_t code = tile[sm_apple[j]]; e = !(code & OCC_BIT) && (code & CODE_MASK) != C_APPLE; }
            eligm[q] = __ballot_sync(FULL, e);
            M += __popc(eligm[q]);
        } else eligm[q] = 0;
    }
    // waste candidates = non-'H' cells of waste_points (apple points and waste points are disjoint, so the apple
    // spawns below cannot change them)
    unsigned candm[ROUNDS] = {};
    int C = 0;
    if (waste_on) {
#pragma unroll
        for (int q = 0; q < ROUNDS; q++) {
            int j = lane + 32 * q;
            bool cnd = false;
            if (q * 32 < p.n_waste) {
                if (j < p.n_waste) cnd = (tile[sm_waste[j]] & CODE_MASK) != C_WASTE;
                candm[q] = __ballot_sync(FULL, cnd);
                C += __popc(candm[q]);
            } else candm[q] = 0;
        }
    }
    // one Philox pass for everything this spawn can need: the apple draws [0, M), the first waste draws [M, M + wfirst)
    // (the waste scan stops at its first success, p = 0.5) and the shuffle keys of the waste points
    const int wfirst = C > 0 ? min(min(C, 8), SCRATCH_DRAWS - M) : 0;
    const int nblk = (thrA || wfirst) ? (M + wfirst + 3) >> 2 : 0;
    const int nkblk = C > 0 ? (p.n_waste + 3) >> 2 : 0;
    for (int bl = lane; __any_sync(FULL, bl < nblk || bl < nkblk); bl += 32)
        draw_blocks2_ool(g.seed, g.env_id, g.episode, t,
                         SITE_SPAWN_DRAWS, (uint32_t)bl, bl < nblk ? reinterpret_cast<uint4*>(draws) + bl : nullptr,
                         SITE_WASTE_ORDER, (uint32_t)bl, bl < nkblk ? reinterpret_cast<uint4*>(keys) + bl : nullptr);
    __syncwarp();
    if (thrA) {
        int base = 0;
#pragma unroll
        for (int q = 0; q < ROUNDS; q++) {
            if (q * 32 < p.n_apple) {
                if ((eligm[q] >> lane) & 1u) {
                    int r = base + __popc(eligm[q] & lanemask_lt(lane));
                    if (draws[r] < thrA) { int cell = sm_apple[lane + 32 * q]; tile[cell] = (uint8_t)((tile[cell] & OCC_BIT) | C_APPLE); spawned = true; }
                }
                base += __popc(eligm[q]);
            }
        }
        __syncwarp();
    }
    const bool apples = __any_sync(FULL, spawned);
    if (!waste_on || C == 0) return apples;
    // waste: shuffle waste_points (stateless: key per canonical index), scan non-'H' cells in that
    // order, draw continues at rank M; first success spawns and breaks (:338-348)
    // number of failed draws before the first success
    int kstar = -1;
    {
        const unsigned succ = __ballot_sync(FULL, lane < wfirst && draws[M + lane] < p.thr_waste);
        if (succ) kstar = __ffs(succ) - 1;
    }
    for (int k0 = wfirst; k0 < C && kstar < 0; k0 += 32) {
        uint32_t dr = draw_u32_ool(g, t, SITE_SPAWN_DRAWS, (uint32_t)(M + k0 + lane));
        unsigned succ = __ballot_sync(FULL, (k0 + lane) < C && dr < p.thr_waste);
        if (succ) kstar = k0 + __ffs(succ) - 1;
    }
    if (kstar < 0) return apples;
    // the (kstar+1)-th smallest (key, index) among the candidates
    int chosen = -1;
    for (int it = 0; it <= kstar; it++) {
        uint32_t bk = 0xFFFFFFFFu; int bj = 0x7FFFFFFF;
#pragma unroll
        for (int q = 0; q < ROUNDS; q++) {
            if ((candm[q] >> lane) & 1u) {
                int j = lane + 32 * q; uint32_t kk = keys[j];
                if (kk < bk || (kk == bk && j < bj)) { bk = kk; bj = j; }
            }
        }
        uint32_t kmin = __reduce_min_sync(FULL, bk);
        int jmin = (int)__reduce_min_sync(FULL, (bk == kmin) ? (uint32_t)bj : 0x7FFFFFFFu);
        chosen = jmin;
        // remove it from the candidate set (uniform update of the owning lane's bit)
        int ql = jmin >> 5, ll = jmin & 31;
#pragma unroll
        for (int q = 0; q < ROUNDS; q++) if (q == ql) candm[q] &= ~(1u << ll);
    }
    if (lane == 0) { int cell = sm_waste[chosen]; tile[cell] = (uint8_t)((tile[cell] & OCC_BIT) | C_WASTE); }
    __syncwarp();
    hcount += 1;
    return true;
}

// harvest spawn (harvest_new.py:284-317): neighbour counts read the pre-spawn map.
// Every eligible point consumes one draw (by rank), but only a draw below the LARGEST spawn probability (0.05) can spawn
// whatever the neighbour count is.  So the 3x3 neighbour counts (8 tile reads) are evaluated only for those ~5 % of the
// eligible points, compacted into a list: one pass of the warp instead of one per 32 points.
template <int ROUNDS>
__device__ __forceinline__ bool harvest_spawn(const GridParams& p, int lane, uint8_t* tile, uint32_t* scratch,
                                              const uint16_t* sm_apple, const EnvRng& g, uint32_t t)
{
    const int S = p.S;
    unsigned eligm[ROUNDS];
    int M = 0;
#pragma unroll
    for (int q = 0; q < ROUNDS; q++) {
        eligm[q] = 0;
        if (q * 32 < p.n_apple) {
            const int j = lane + 32 * q;
            bool e = false;
            if (j < p.n_apple) { const uint32_t code = tile[sm_apple[j]]; e = !(code & OCC_BIT) && (code & CODE_MASK) != C_APPLE; }
            eligm[q] = __ballot_sync(FULL, e);
            M += __popc(eligm[q]);
        }
    }
    if (M == 0) return false;
    {   // draws [0, M): Philox blocks lane and lane + 32 as two interleaved chains (M > 128 needs more than 32 blocks)
        const int nblk = (M + 3) >> 2;
        uint4* d4 = reinterpret_cast<uint4*>(scratch);
        if (nblk <= 32) {
            if (lane < nblk) d4[lane] = draw_block_ool(g.seed, g.env_id, g.episode, t, SITE_SPAWN_DRAWS, (uint32_t)lane);
        } else {
            for (int bl = lane; __any_sync(FULL, bl < nblk); bl += 64)
                draw_blocks2_ool(g.seed, g.env_id, g.episode, t, SITE_SPAWN_DRAWS, (uint32_t)bl, bl < nblk ? d4 + bl : nullptr,
                                 SITE_SPAWN_DRAWS, (uint32_t)(bl + 32), bl + 32 < nblk ? d4 + bl + 32 : nullptr);
        }
        __syncwarp();                                                  // also orders the tile reads above
    }
    // candidates: point index | draw rank << 16, compacted behind the draws (M <= 256 draws + M <= 256 entries)
    uint32_t* list = scratch + SCRATCH_DRAWS;
    const uint32_t thr_max = p.thr_harvest[3];                         // SPAWN_PROB is increasing (harvest_new.py:34)
    int base = 0, ncand = 0;
#pragma unroll
    for (int q = 0; q < ROUNDS; q++) {
        if (q * 32 < p.n_apple) {
            const bool el = (eligm[q] >> lane) & 1u;
            const int r = base + __popc(eligm[q] & lanemask_lt(lane));
            const bool c = el && scratch[r] < thr_max;
            const unsigned cm = __ballot_sync(FULL, c);
            if (c) list[ncand + __popc(cm & lanemask_lt(lane))] = (uint32_t)(lane + 32 * q) | ((uint32_t)r << 16);
            ncand += __popc(cm);
            base += __popc(eligm[q]);
        }
    }
    if (ncand == 0) return false;
    __syncwarp();
    // neighbour counts of the candidates on the PRE-spawn map (j*j + k*k <= APPLE_RADIUS(=2): the 3x3 block; the own cell
    // is not 'A'), decisions first, writes after all of them
    for (int k = lane; k < ncand; k += 32) {
        const uint32_t ent = list[k];
        const int cell = sm_apple[ent & 0xFFFFu];
        int cnt = 0;
#pragma unroll
        for (int dr = -1; dr <= 1; dr++)
#pragma unroll
            for (int dc = -1; dc <= 1; dc++)
                if (dr | dc) cnt += ((tile[cell + dr * S + dc] & CODE_MASK) == C_APPLE);
        list[k] = scratch[ent >> 16] < p.thr_harvest[cnt < 3 ? cnt : 3] ? (uint32_t)cell : 0xFFFFFFFFu;
    }
    __syncwarp();
    bool spawned = false;
    for (int k = lane; k < ncand; k += 32) {
        const uint32_t cell = list[k];
        if (cell != 0xFFFFFFFFu) { tile[cell] = (uint8_t)((tile[cell] & OCC_BIT) | C_APPLE); spawned = true; }
    }
    __syncwarp();
    return __any_sync(FULL, spawned);
}

// count_apples_in_radius(5, loc) (harvest_new.py:326-336): the 21 cells with j*j + k*k <= 5
__device__ __forceinline__ int count_apples_r5(const uint8_t* tile, int o, int S)
{
    int cnt = 0;
#pragma unroll
    for (int dr = -2; dr <= 2; dr++)
#pragma unroll
        for (int dc = -2; dc <= 2; dc++)
            if (dr * dr + dc * dc <= 5) cnt += ((tile[o + dr * S + dc] & CODE_MASK) == C_APPLE);
    return cnt;
}

// ---------------------------------------------------------------------------------------------
// infos['feature_obs'] (cleanup_new.py:235-251, harvest_new.py:208-222).  Lists are the row-major
// scans of the post-step map (compute_current_apples / _wastes); np.argmin takes the first minimum,
// i.e. the smallest (L1 distance, list index).  Returns (dist << 16 | index) or 0xFFFFFFFF.
__device__ __forceinline__ uint32_t closest_point(int lane, const uint8_t* tile, const uint16_t* pts,
                                                  const uint16_t* rc, int npts, uint32_t code, int ar, int ac)
{
    uint32_t best = 0xFFFFFFFFu;
    for (int j = lane; j < npts; j += 32) {
        if ((tile[pts[j]] & CODE_MASK) == code) {
            int r = rc[j] >> 8, c = rc[j] & 255;
            uint32_t key = ((uint32_t)(abs(r - ar) + abs(c - ac)) << 16) | (uint32_t)j;
            best = key < best ? key : best;
        }
    }
    return __reduce_min_sync(FULL, best);
}
__device__ __forceinline__ int count_points(int lane, const uint8_t* tile, const uint16_t* pts, int npts, uint32_t code)
{
    int cnt = 0;
    for (int j0 = 0; j0 < npts; j0 += 32) {
        int j = j0 + lane;
        cnt += __popc(__ballot_sync(FULL, j < npts && (tile[pts[j]] & CODE_MASK) == code));
    }
    return cnt;
}
template <int KIND>
__device__ __forceinline__ void write_features(const GridParams& p, const SharedTables& tb, int lane, const uint8_t* tile,
                                               int ao, int ori, int cleaned, int total_close, int hcount, double* out)
{
    const int n = p.n, S = p.S;
    const int my_r = ao / S - SSD_VIEW, my_c = ao % S - 8;
    const int n_apples = count_points(lane, tile, tb.apple, p.n_apple, C_APPLE);
    int ca_r = 0, ca_c = 0, cw_r = 0, cw_c = 0;        // sentinel [0, 0] when the list is empty
    for (int a = 0; a < n; a++) {
        int ar = __shfl_sync(FULL, my_r, a), ac = __shfl_sync(FULL, my_c, a);
        uint32_t ka = closest_point(lane, tile, tb.apple, tb.apple_rc, p.n_apple, C_APPLE, ar, ac);
        uint32_t kw = 0xFFFFFFFFu;
        if (KIND == SSD_ENV_CLEANUP) kw = closest_point(lane, tile, tb.waste, tb.waste_rc, p.n_waste, C_WASTE, ar, ac);
        if (lane == a) {
            if (ka != 0xFFFFFFFFu) { ca_r = tb.apple_rc[ka & 0xFFFFu] >> 8; ca_c = tb.apple_rc[ka & 0xFFFFu] & 255; }
            if (kw != 0xFFFFFFFFu) { cw_r = tb.waste_rc[kw & 0xFFFFu] >> 8; cw_c = tb.waste_rc[kw & 0xFFFFu] & 255; }
        }
    }
    // compute_closest_pos (cleanup_new.py:405-412): distances are 0 for every other agent, so the
    // argmin is agent 0 (agent 1 for agent 0; itself when alone)
    const int cp = lane == 0 ? (n > 1 ? 1 : 0) : 0;
    const int cp_r = __shfl_sync(FULL, my_r, cp), cp_c = __shfl_sync(FULL, my_c, cp), cp_o = __shfl_sync(FULL, ori, cp);
    int cl[SSD_MAXN];
#pragma unroll
    for (int j = 0; j < SSD_MAXN; j++) cl[j] = __shfl_sync(FULL, cleaned, j);
    if (lane < n) {
        double* f = out + (size_t)lane * p.F;
        f[0] = my_r; f[1] = my_c; f[2] = ori; f[3] = cp_r; f[4] = cp_c; f[5] = cp_o; f[6] = ca_r; f[7] = ca_c;
        if (KIND == SSD_ENV_CLEANUP) {
            f[8] = cw_r; f[9] = cw_c; f[10] = n_apples; f[11] = hcount;
#pragma unroll
            for (int j = 0; j < SSD_MAXN; j++) if (j < n) f[12 + j] = cl[j];
        } else {
            f[8] = total_close; f[9] = n_apples;
            for (int j = 0; j < 2 * n; j++) f[10 + j] = 0.0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// observations (color_view, map_env.py:397-411): out[i][j] of agent a is a strided walk over the
// painted tile: V[i][j] (UP), V[j][14-i] (LEFT), V[14-i][14-j] (DOWN), V[14-j][i] (RIGHT), with
// V[a][b] = tile[origin + a*S + b].  So output row i starts at start0 + i*rs and advances by `step`
// per pixel; {start0, step, rs} per agent live in vdesc.  The env's observation stream is 15 n rows
// of 45 B; a work item is 4 consecutive rows = 60 pixels = 45 words (word aligned whatever the
// agents), one item per lane, staged with conflict-free word stores (lane stride 45 words) and
// shipped with one bulk async store (head / tail up to the 16-B boundaries by plain stores).
__device__ __forceinline__ void pack4(uint32_t* dst, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3)
{
    dst[0] = __byte_perm(p0, p1, 0x4210);
    dst[1] = __byte_perm(p1, p2, 0x5421);
    dst[2] = __byte_perm(p2, p3, 0x6542);
}
// the staged observation stream of one env (L bytes at stage + shift, congruent to gdst modulo 16) -> HBM:
// one bulk async store for the 16-byte aligned middle, plain stores for the head / tail
__device__ __forceinline__ void obs_stream_store(int lane, uint8_t* stage, int shift, bool word_ok, int L, uint8_t* gdst)
{
    uint32_t* sw = reinterpret_cast<uint32_t*>(stage + shift);
    __syncwarp();
    if (word_ok) {
        const int a0 = (16 - shift) & 15;               // head bytes up to the first 16-B boundary
        const int mid = (L - a0) & ~15;
        const int tail0 = a0 + mid;
        // head / tail: whole words first, then bytes
        if (lane < (a0 >> 2)) reinterpret_cast<uint32_t*>(gdst)[lane] = sw[lane];
        {
            int tw = (L - tail0) >> 2;
            if (lane < tw) reinterpret_cast<uint32_t*>(gdst + tail0)[lane] = sw[(tail0 >> 2) + lane];
            int tb0 = tail0 + tw * 4;
            if (lane < L - tb0) gdst[tb0 + lane] = (stage + shift)[tb0 + lane];
        }
        if (mid > 0) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bulk_store(gdst + a0, stage + shift + a0, (uint32_t)mid);
        }
    } else {
        for (int i = lane; i < L; i += 32) gdst[i] = stage[i];
    }
}


// PENDING: bulk groups of this lane-0 that may stay in flight while `stage` is rewritten
template <int PENDING>
__device__ __forceinline__ void gather_obs(const GridParams& p, int lane, const uint8_t* tile, uint8_t* stage,
                                           const uint32_t* sm_pal, int4* vdesc, int ao, int ori, uint8_t* gdst)
{
    const int n = p.n, S = p.S;
    const int L = n * SSD_OBS_BYTES;
    const uint32_t gaddr = (uint32_t)(reinterpret_cast<uintptr_t>(gdst) & 15u);
    const bool word_ok = (gaddr & 3u) == 0;
    const int shift = word_ok ? (int)gaddr : 0;          // staged stream is congruent to gdst modulo 16
    uint32_t* sw = reinterpret_cast<uint32_t*>(stage + shift);
    if (lane < n) {
        const int origin = ao - SSD_VIEW * S - SSD_VIEW;     // V[0][0] of this lane's agent
        const int last = SSD_OBSW - 1;
        int start0 = origin, step = 1, rs = S;               // UP
        if (ori == ORI_RIGHT) { start0 = origin + last * S; step = -S; rs = 1; }
        else if (ori == ORI_DOWN) { start0 = origin + last * S + last; step = -1; rs = -S; }
        else if (ori == ORI_LEFT) { start0 = origin + last; step = S; rs = -1; }
        vdesc[lane] = make_int4(start0, step, rs, 0);
    }
    if (lane == 0) bulk_wait_read<PENDING>();                // the previous env's store has drained `stage`
    __syncwarp();
    if (lane < p.obs_items) {
        uint32_t col[60];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int r = 4 * lane + q;
            int a = (int)(((uint32_t)r * 0x8889u) >> 19);    // r / 15
            int i = r - 15 * a;
            if (a >= n) { a = n - 1; i = SSD_OBSW - 1; }     // rows past the stream end: harmless in-bounds reads
            const int4 d = vdesc[a];
            const uint8_t* src = tile + (d.x + i * d.z);
#pragma unroll
            for (int j = 0; j < SSD_OBSW; j++)      // a tile byte IS the byte offset of its palette entry
                col[15 * q + j] = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(sm_pal) + src[j * d.y]);
        }
        uint32_t* dst = sw + 45 * lane;
#pragma unroll
        for (int g4 = 0; g4 < 15; g4++) pack4(dst + 3 * g4, col[4 * g4], col[4 * g4 + 1], col[4 * g4 + 2], col[4 * g4 + 3]);
    }
    obs_stream_store(lane, stage, shift, word_ok, L, gdst);
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ SharedTables load_shared_tables(const GridParams& p, uint8_t* smem)
{
    uint32_t* pal = reinterpret_cast<uint32_t*>(smem);          // [16]: 16 distinct banks, no replication needed
    uint32_t* thr = reinterpret_cast<uint32_t*>(smem + p.sm_thr);
    uint8_t* won = smem + p.sm_won;
    uint16_t* apple = reinterpret_cast<uint16_t*>(smem + p.sm_apple);
    uint16_t* waste = reinterpret_cast<uint16_t*>(smem + p.sm_waste);
    uint16_t* apple_rc = reinterpret_cast<uint16_t*>(smem + p.sm_apple_rc);
    uint16_t* waste_rc = reinterpret_cast<uint16_t*>(smem + p.sm_waste_rc);
    for (int i = threadIdx.x; i < 16; i += blockDim.x) pal[i] = __ldg(p.pal + i);
    for (int i = threadIdx.x; i <= p.n_waste; i += blockDim.x) { thr[i] = __ldg(p.thr_apple + i); won[i] = __ldg(p.waste_on + i); }
    for (int i = threadIdx.x; i < p.n_apple; i += blockDim.x) apple[i] = __ldg(p.apple_pts + i);
    for (int i = threadIdx.x; i < p.n_waste; i += blockDim.x) waste[i] = __ldg(p.waste_pts + i);
    for (int i = threadIdx.x; i < p.n_apple; i += blockDim.x) apple_rc[i] = __ldg(p.apple_rc + i);
    for (int i = threadIdx.x; i < p.n_waste; i += blockDim.x) waste_rc[i] = __ldg(p.waste_rc + i);
    SharedTables t = { pal, thr, won, apple, waste, apple_rc, waste_rc };
    return t;
}

// =============================================================================================
// STEP
// ROUNDS: 32-point rounds the apple / waste point lists need (4 covers the stock cleanup map);
// FEAT: the variant that also writes infos['feature_obs'] (kept out of the hot variant's code).
template <int KIND, int ROUNDS, bool FEAT>
__global__ void __launch_bounds__(GRID_THREADS, GRID_MIN_BLOCKS) grid_step_kernel(const GridParams p, const StepIO io)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const SharedTables tb = load_shared_tables(p, smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* tile = smem + p.sm_warp0 + warp * p.warp_bytes;
    uint8_t* recs = tile + p.off_rec;                 // two record slots (double buffer)
    uint8_t* stage = tile + p.off_stage;
    // the spawn scratch (draws + shuffle keys) aliases the observation staging buffer: both are
    // dead outside their phase once the previous env's bulk store has been read out
    uint32_t* scratch = reinterpret_cast<uint32_t*>(stage);
    int4* vdesc = reinterpret_cast<int4*>(tile + p.off_misc + MISC_VDESC);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(tile + p.off_misc + MISC_MBAR);
    for (int i = lane; i < (p.tile_r16 >> 2); i += 32) reinterpret_cast<uint32_t*>(tile)[i] = TILE_FILL4;
    if (lane == 0) { mbar_init(mbar, 1); mbar_init(mbar + 1, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();                                  // tables + barriers visible
    const int n = p.n, S = p.S;
    const bool act_lane = lane < n;
    const int env_stride = gridDim.x * GRID_WARPS;
    const uint32_t rec_bytes = (uint32_t)p.rec_stride;
    const MapWords mw = map_words_init(p, lane);

    int env = blockIdx.x * GRID_WARPS + warp;
    // running global pointers of the current env (advanced by one grid stride per iteration)
    uint8_t* g_rec = p.state + (size_t)env * p.rec_stride;
    const size_t g_rec_step = (size_t)env_stride * p.rec_stride;
    const uint8_t* g_act = io.actions + (size_t)env * n + (act_lane ? lane : 0);
    const size_t g_act_step = (size_t)env_stride * n;
    uint8_t* g_obs = io.obs + (size_t)env * (size_t)io.obs_stride;
    const size_t g_obs_step = (size_t)env_stride * (size_t)io.obs_stride;
    int act_next = 4;
    if (env < p.E) {
        if (lane == 0) bulk_load(recs, g_rec, rec_bytes, mbar);
        if (act_lane) act_next = *g_act;
    }
    for (uint32_t it = 0; env < p.E; env += env_stride, it++) {
        const uint32_t slot = it & 1u;
        uint8_t* rec = recs + slot * p.rec_stride;    // this env's record, in shared memory
        uint8_t* hdr = rec + p.map_bytes;
        // ---- prefetch the next env's record into the other slot (its previous contents left
        //      through the bulk store of the previous iteration: wait until that has read it)
        const int env_nx = env + env_stride;
        int act = act_next;
        if (env_nx < p.E) {
            if (lane == 0) {
                bulk_wait_read<1>();                  // all but the newest group (an observation store)
                bulk_load(recs + (slot ^ 1u) * p.rec_stride, g_rec + g_rec_step, rec_bytes, mbar + (slot ^ 1u));
            }
            if (act_lane) act_next = g_act[g_act_step];
        }
        mbar_wait(mbar + slot, (it >> 1) & 1u);
        tile_expand(p, mw, rec, tile, lane);
        // ---- per-env scalars + agent registers
        int t = *reinterpret_cast<const int*>(hdr + RO_T) + 1;                 // map_env.py:230
        const uint32_t episode = *reinterpret_cast<const uint32_t*>(hdr + RO_EPISODE);
        const double theta = *reinterpret_cast<const double*>(hdr + RO_THETA);
        uint32_t flags = *reinterpret_cast<const uint32_t*>(hdr + RO_FLAGS);
        int hcount = *reinterpret_cast<const int*>(hdr + RO_HCOUNT);
        EnvRng g = { p.seed, p.first_env_id + (uint32_t)env, episode, (uint32_t)t };
        int ao = 0, ori = 0;
        if (act_lane) {
            uint32_t a = reinterpret_cast<const uint32_t*>(hdr + RO_AGENTS)[lane];
            ao = ((int)(a & 255u) + SSD_VIEW) * S + 8 + (int)((a >> 8) & 255u);
            ori = (int)((a >> 16) & 3u);
        } else act = 4;
        __syncwarp();

        // ---- action decode (Agent.py:8-16,161-162,198-199) + rotations (map_env.py:514-516)
        int cls = 0;                          // 0 none, 1 FIRE, 2 CLEAN
        bool has_move = false;
        int tgt = ao;
        uint32_t err = 0;
        if (act_lane) {
            if (act <= 4) {
                has_move = true;
                if (act < 4) {
                    // egocentric -> world direction: LEFT ori+3, RIGHT ori+1, UP ori, DOWN ori+2 (rotate_action :844-853)
                    int rel = (0x2013 >> (4 * act)) & 3;
                    int cand = ao + dir_delta((ori + rel) & 3, S);
                    uint32_t c = tile[cand] & CODE_MASK;
                    if (c != C_WALL && c != C_OUTSIDE) tgt = cand;           // return_valid_pos (Agent.py:111-119)
                }
            } else if (act == 5) ori = (ori + 1) & 3;                          // TURN_CLOCKWISE
            else if (act == 6) ori = (ori + 3) & 3;                            // TURN_COUNTERCLOCKWISE
            else if (KIND == SSD_ENV_HARVEST) { if (act == 7) cls = 1; else { err |= 8; has_move = true; } }
            else if (act == 7) cls = 2;
            else if (act == 8) cls = 1;
            else { err |= 8; has_move = true; }
        }
        err |= resolve_moves(lane, n, tile, g, ao, has_move, tgt);

        // ---- stale-list infos (cleanup_new.py:220-223, harvest_new.py:190-199) on the start-of-step map
        int eaten = 0, eaten_close = 0, total_close = 0;
        int reward = 0, cleaned = 0;
        // co-located agents (possible after unresolved conflicts) share one cell
        const unsigned grp = __match_any_sync(FULL, act_lane ? (uint32_t)ao : (0x40000000u | (uint32_t)lane));
        if (act_lane) {
            bool on_apple = (tile[ao] & CODE_MASK) == C_APPLE;
            if (on_apple && !(flags & RF_STALE_EMPTY)) {
                eaten = 1;
                if (KIND == SSD_ENV_HARVEST) eaten_close = count_apples_r5(tile, ao, S) < 4;
            }
            // consume in agent order: the lowest-index co-located agent eats (map_env.py:244-247)
            if (on_apple && lane == __ffs(grp) - 1) reward += 1;
        }
        __syncwarp();
        if (act_lane) {
            uint32_t c = tile[ao];
            if ((c & CODE_MASK) == C_APPLE) c = C_EMPTY;
            tile[ao] = (uint8_t)(c | OCC_BIT);      // occupancy for beams + spawn eligibility
        }
        __syncwarp();

        // ---- beams, spawning
        int ncleaned = fire_beams(lane, n, S, tile, g, ao, ori, cls, reward, cleaned);
        if (KIND == SSD_ENV_CLEANUP) {
            hcount -= ncleaned;
            if (cleanup_spawn_active(tb, hcount)) {
                if (lane == 0) bulk_wait_read<0>();  // previous env's observation store has drained `stage` (= scratch)
                __syncwarp();
                cleanup_spawn<ROUNDS>(p, tb, lane, tile, scratch, g, (uint32_t)t, hcount);
            }
        } else {
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            harvest_spawn<ROUNDS>(p, lane, tile, scratch, tb.apple, g, (uint32_t)t);
        }
        if (KIND == SSD_ENV_HARVEST && act_lane) total_close = count_apples_r5(tile, ao, S);
        if (FEAT) write_features<KIND>(p, tb, lane, tile, ao, ori, cleaned, total_close, hcount,
                                       io.feat + (size_t)env * n * p.F);

        // ---- map back into the record slot (paint stripped)
        tile_compress(p, mw, rec, tile, lane);

        // ---- rewards + contract transfers (contract_list.py, two_stage_train.py:69-92), lane j = agent j
        double base = (double)reward;
        if (p.reward_mode)                             // shaped env rewards (map_env.py:289-301)
            base = shaped_reward_warp(p.reward_mode, p.alpha, p.beta, n, act_lane ? reward : 0);
        double tr = 0.0;
        if (p.contract == SSD_CONTRACT_CLEANUP) tr = __dmul_rn(-theta, (double)cleaned);
        else if (p.contract == SSD_CONTRACT_HARVEST_LOCAL) tr = (total_close < 4 && eaten_close > 0) ? theta : 0.0;
        double r = base, total_tr = 0.0;
        // Redistribution in the reference's order (two_stage_train.py:72-90).  When every transfer is
        // +-0.0 the loop is the identity on r (and adds +0.0 to the totals), so it is skipped.
        if (p.contract != SSD_CONTRACT_NONE && __ballot_sync(FULL, act_lane && tr != 0.0) != 0u) {
            const double share = __ddiv_rn(tr, (double)(n - 1));      // transfers[i] / (len(acts) - 1)
            for (int i = 0; i < n; i++) {
                double ti = __shfl_sync(FULL, tr, i), qi = __shfl_sync(FULL, share, i);
                r = (i == lane) ? __dsub_rn(r, ti) : __dadd_rn(r, qi);
                total_tr = __dadd_rn(total_tr, ti);
            }
        }
        const bool done = t == p.horizon;
        double raw_step = 0.0;
        if (p.reward_mode)
            for (int i = 0; i < n; i++) raw_step = __dadd_rn(raw_step, __shfl_sync(FULL, base, i));
        if (act_lane) {
            size_t o = (size_t)env * n + lane;
            if (io.rew) io.rew[o] = r;
            if (io.base_rew) io.base_rew[o] = base;
            if (io.transfers) io.transfers[o] = tr;
            if (io.info) reinterpret_cast<uint32_t*>(io.info)[o] =
                (uint32_t)eaten | ((uint32_t)(KIND == SSD_ENV_CLEANUP ? cleaned : eaten_close) << 8) | ((uint32_t)total_close << 16);
            // agent record + accumulators (in the shared record slot)
            uint32_t trow = __umulhi((uint32_t)ao, p.s_magic);
            uint32_t row = trow - SSD_VIEW, col = (uint32_t)ao - trow * (uint32_t)S - 8u;
            reinterpret_cast<uint32_t*>(hdr + RO_AGENTS)[lane] = row | (col << 8) | ((uint32_t)ori << 16);
            if (p.reward_mode) {
                double* xs = reinterpret_cast<double*>(hdr + RO_XSUM) + lane;
                double* xt = reinterpret_cast<double*>(hdr + RO_XTSUM) + lane;
                *xs = __dadd_rn(*xs, base);
                *xt = __dadd_rn(*xt, __dmul_rn((double)(t - 1), base));
            } else if (reward != 0) {
                reinterpret_cast<int*>(hdr + RO_SUM_RAW)[lane] += reward;
                reinterpret_cast<long long*>(hdr + RO_TSUM_RAW)[lane] += (long long)(t - 1) * reward;
            }
            if (p.contract != SSD_CONTRACT_NONE) {
                const double tm1 = (double)(t - 1);
                double* st = reinterpret_cast<double*>(hdr + RO_SUM_TR) + lane;
                double* tt = reinterpret_cast<double*>(hdr + RO_TSUM_TR) + lane;
                *st = __dadd_rn(*st, r);
                *tt = __dadd_rn(*tt, __dmul_rn(tm1, r));
            }
            if (KIND == SSD_ENV_CLEANUP) { if (cleaned) reinterpret_cast<uint32_t*>(hdr + RO_AGENT_A)[lane] += (uint32_t)cleaned; }
            else {
                reinterpret_cast<uint32_t*>(hdr + RO_AGENT_A)[lane] += (uint32_t)eaten;
                reinterpret_cast<uint32_t*>(hdr + RO_AGENT_B)[lane] += (uint32_t)eaten_close;
            }
        }
        unsigned eatm = __ballot_sync(FULL, eaten != 0), closem = __ballot_sync(FULL, eaten_close != 0);
        unsigned errm = __ballot_sync(FULL, err != 0);
        uint32_t errbits = __reduce_or_sync(FULL, err);
        if (lane == 0) {
            *reinterpret_cast<int*>(hdr + RO_T) = t;
            *reinterpret_cast<uint32_t*>(hdr + RO_FLAGS) = (flags & ~RF_STALE_EMPTY) | (errm ? (errbits << RF_ERR_SHIFT) : 0u);
            *reinterpret_cast<int*>(hdr + RO_HCOUNT) = hcount;
            *reinterpret_cast<uint32_t*>(hdr + RO_APPLES) += (uint32_t)__popc(eatm);
            if (KIND == SSD_ENV_HARVEST) *reinterpret_cast<uint32_t*>(hdr + RO_LOWDENS) += (uint32_t)__popc(closem);
            else *reinterpret_cast<uint32_t*>(hdr + RO_DIRT) += (uint32_t)ncleaned;
            if (p.contract != SSD_CONTRACT_NONE) {
                double* mt = reinterpret_cast<double*>(hdr + RO_TRANSFERS);
                *mt = __dadd_rn(*mt, total_tr);
            }
            if (p.reward_mode) {                       // raw_rewards = ((0 + r0) + r1) + ... (cleanup_new.py:228-232)
                double* xr = reinterpret_cast<double*>(hdr + RO_XRAW);
                *xr = __dadd_rn(*xr, raw_step);
            }
            if (io.done) io.done[env] = done ? 1 : 0;
            if (io.c_done) io.c_done[env] = done ? 1 : 0;
        }
        if (io.c_rew8) {                               // compact result block (ssd_step_host_async), lane = agent
            int v = 0;
            const bool fits = reward_fits_i8(r, v);
            const bool sparse = __ballot_sync(FULL, act_lane && !fits) != 0u;
            if (act_lane) io.c_rew8[(size_t)env * n + lane] = (int8_t)(fits ? v : -128);
            if (sparse) {
                uint32_t slot = 0;
                if (lane == 0) slot = atomicAdd(io.c_count, 1u);
                slot = __shfl_sync(FULL, slot, 0);
                uint8_t* rec_o = io.c_rec + (size_t)slot * (8 + 8 * n);
                if (lane == 0) *reinterpret_cast<int2*>(rec_o) = make_int2(env, 0);
                if (act_lane) reinterpret_cast<double*>(rec_o + 8)[lane] = r;
            }
        }
        // ---- the record leaves with one bulk store
        fence_async_smem();
        __syncwarp();
        if (lane == 0) bulk_store(g_rec, rec, rec_bytes);

        // ---- paint agents in agent order: the highest index wins a shared cell (map_env.py:257-261).
        // Palette index 6 + i = agent i; the interior of the tile is rewritten by the next tile_expand.
        if (act_lane && lane == 31 - __clz(grp)) tile[ao] = (uint8_t)PAINT_CODE(lane);
        // gather_obs: waits (lane 0) until at most the record store above is still in flight, syncs the warp
        gather_obs<1>(p, lane, tile, stage, tb.pal, vdesc, ao, ori, g_obs);
        g_rec += g_rec_step; g_act += g_act_step; g_obs += g_obs_step;
        __syncwarp();
    }
    if (lane == 0) bulk_wait_read<0>();      // smem must outlive the async bulk reads
}

// =============================================================================================
// RESET: setup_agents + reset_map + custom_reset + reset-time spawn + contract sample + reset obs
template <int KIND>
__global__ void __launch_bounds__(GRID_THREADS) grid_reset_kernel(const GridParams p, const uint8_t* mask,
                                                                  uint8_t* obs, long long obs_stride)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const SharedTables tb = load_shared_tables(p, smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* tile = smem + p.sm_warp0 + warp * p.warp_bytes;
    uint8_t* stage = tile + p.off_stage;
    uint32_t* scratch = reinterpret_cast<uint32_t*>(stage);
    int4* vdesc = reinterpret_cast<int4*>(tile + p.off_misc + MISC_VDESC);
    for (int i = lane; i < (p.tile_r16 >> 2); i += 32) reinterpret_cast<uint32_t*>(tile)[i] = TILE_FILL4;
    __syncthreads();
    const int n = p.n, S = p.S;
    const bool act_lane = lane < n;

    // The warp's envs are env0 + i * estride.  The mask bytes of 32 of them are read at once (one round trip, one
    // ballot), so a sparse mask — a vectorised sampler in steady state resets ~E / horizon envs per step — costs a
    // few loads per warp instead of one dependent load per env.
    const int env0 = blockIdx.x * GRID_WARPS + warp, estride = gridDim.x * GRID_WARPS;
    for (int i0 = 0; env0 + (long long)i0 * estride < p.E; i0 += 32) {
      const long long el = env0 + (long long)(i0 + lane) * estride;
      unsigned todo = __ballot_sync(FULL, el < p.E && (!mask || mask[el]));
      while (todo) {
        const int env = env0 + (i0 + __ffs(todo) - 1) * estride;
        todo &= todo - 1;
        uint8_t* rec = p.state + (size_t)env * p.rec_stride;
        uint8_t* hdr = rec + p.map_bytes;
        uint32_t flags = *reinterpret_cast<const uint32_t*>(hdr + RO_FLAGS);
        uint32_t episode = *reinterpret_cast<const uint32_t*>(hdr + RO_EPISODE);
        if (p.stats && (flags & 0x80000000u)) {       // a finished episode is replaced: hand its accumulators over
            double tr = 0.0, raw = 0.0;
            if (act_lane) {
                tr = reinterpret_cast<const double*>(hdr + RO_SUM_TR)[lane];
                raw = p.reward_mode ? reinterpret_cast<const double*>(hdr + RO_XSUM)[lane]
                                    : (double)reinterpret_cast<const int*>(hdr + RO_SUM_RAW)[lane];
            }
            for (int o = 16; o; o >>= 1) { tr += __shfl_xor_sync(FULL, tr, o); raw += __shfl_xor_sync(FULL, raw, o); }
            if (lane == 0) {
                atomicAdd(p.stats + 0, (double)*reinterpret_cast<const uint32_t*>(hdr + RO_APPLES));
                atomicAdd(p.stats + 1, p.reward_mode ? *reinterpret_cast<const double*>(hdr + RO_XRAW) : raw);
                atomicAdd(p.stats + 2, *reinterpret_cast<const double*>(hdr + RO_TRANSFERS));
                atomicAdd(p.stats + 3, (double)*reinterpret_cast<const uint32_t*>(hdr + RO_DIRT));
                atomicAdd(p.stats + 4, tr);
                atomicAdd(p.stats + 5, raw);
                atomicAdd(p.stats + 6, 1.0);
                // non-negative doubles order like their bit patterns
                atomicMax(reinterpret_cast<unsigned long long*>(p.stats + 7),
                          (unsigned long long)__double_as_longlong((double)((flags >> RF_ERR_SHIFT) & 0xFFFFu)));
            }
        }
        // the first reset of a record (flags bit 31 clear) is episode 0; later resets increment
        episode = (flags & 0x80000000u) ? episode + 1u : 0u;
        EnvRng g = { p.seed, p.first_env_id + (uint32_t)env, episode, 0u };

        // ---- setup_agents (cleanup_new.py:302-320 / harvest_new.py:132-141; map_env.py:816-832)
        int ao = 0, ori = 0;
        unsigned taken[4] = { 0, 0, 0, 0 };          // bit per canonical spawn entry (<= 128), uniform
        for (int i = 0; i < n; i++) {
            // shuffled list = canonical entries ordered by (key, idx); the LAST free entry wins (:822-825)
            uint32_t bk = 0; int bj = -1;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int j = lane + 32 * q;
                if (j < p.n_spawn && !((taken[q] >> lane) & 1u)) {
                    uint32_t kk = draw_u32(g.seed, g.env_id, g.episode, 0, SITE_SPAWN_POINT, (uint32_t)i, (uint32_t)j);
                    if (bj < 0 || kk > bk || (kk == bk && j > bj)) { bk = kk; bj = j; }
                }
            }
            uint32_t kmax = __reduce_max_sync(FULL, bj >= 0 ? bk : 0u);
            int jmax = (int)__reduce_max_sync(FULL, (bj >= 0 && bk == kmax) ? (uint32_t)(bj + 1) : 0u) - 1;
            int cell = (int)__ldg(p.spawn_pts + jmax);
            // every canonical entry on that cell becomes occupied (cleanup lists each point twice)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int j = lane + 32 * q;
                bool same = j < p.n_spawn && (int)__ldg(p.spawn_pts + j) == cell;
                taken[q] |= __ballot_sync(FULL, same);
            }
            uint32_t rdraw = draw_u32(g.seed, g.env_id, g.episode, 0, SITE_SPAWN_ROT, (uint32_t)i, 0u) >> 30;
            // list(ORIENTATIONS.keys()) = LEFT, RIGHT, UP, DOWN (map_env.py:22, :829-832)
            int o = rdraw == 0 ? ORI_LEFT : (rdraw == 1 ? ORI_RIGHT : (rdraw == 2 ? ORI_UP : ORI_DOWN));
            if (lane == i) { ao = cell; ori = o; }
        }
        // ---- reset_map + custom_reset: copy the initial map, mark occupancy, reset-time spawn (map_env.py:319-320)
        tile_load(p, p.reset_map, tile, lane);
        if (lane == 0) bulk_wait_read<0>();          // `stage` doubles as the spawn scratch
        __syncwarp();
        if (act_lane) tile[ao] |= OCC_BIT;
        __syncwarp();
        int hcount = p.n_waste_start;
        if (KIND == SSD_ENV_CLEANUP) cleanup_spawn<MAX_POINT_ROUNDS>(p, tb, lane, tile, scratch, g, 0u, hcount);
        else harvest_spawn<MAX_POINT_ROUNDS>(p, lane, tile, scratch, tb.apple, g, 0u);
        tile_store(p, rec, tile, lane);
        __syncwarp();
        if (act_lane) tile[ao] &= CODE_MASK;              // MapEnv.reset never paints agents into the colour grid
        __syncwarp();
        if (obs) gather_obs<0>(p, lane, tile, stage, tb.pal, vdesc, ao, ori, obs + (size_t)env * (size_t)obs_stride);

        // ---- SeparateContractSubgameStage.reset (two_stage_train.py:163-168)
        double theta = 0.0;
        if (p.contract != SSD_CONTRACT_NONE) {
            Philox4 q = draw_block(g.seed, g.env_id, g.episode, 0, SITE_CONTRACT, 0, 0);
            double u0 = __dmul_rn((double)q.x, 1.0 / 4294967296.0), u1 = __dmul_rn((double)q.y, 1.0 / 4294967296.0);
            theta = (u0 > p.null_prob) ? __dadd_rn(p.theta_low, __dmul_rn(__dsub_rn(p.theta_high, p.theta_low), u1))
                                       : p.theta_low;
        }
        // ---- record header
        if (act_lane) {
            uint32_t row = (uint32_t)(ao / S) - SSD_VIEW, col = (uint32_t)(ao % S) - 8u;
            reinterpret_cast<uint32_t*>(hdr + RO_AGENTS)[lane] = row | (col << 8) | ((uint32_t)ori << 16);
        }
        for (int i = lane; i < (p.hdr_bytes - RO_T) / 4; i += 32) reinterpret_cast<uint32_t*>(hdr + RO_T)[i] = 0u;
        __syncwarp();
        if (lane == 0) {
            *reinterpret_cast<uint32_t*>(hdr + RO_EPISODE) = episode;
            *reinterpret_cast<double*>(hdr + RO_THETA) = theta;
            *reinterpret_cast<uint32_t*>(hdr + RO_FLAGS) = 0x80000000u | (KIND == SSD_ENV_CLEANUP ? RF_STALE_EMPTY : 0u);
            *reinterpret_cast<int*>(hdr + RO_HCOUNT) = hcount;
        }
        __syncwarp();
      }
    }
    if (lane == 0) bulk_wait_read<0>();
}
