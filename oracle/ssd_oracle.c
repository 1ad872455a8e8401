/*
 * ssd_oracle.c — CPU restatement of the reference's gridworld hot path.
 *
 * TEST INFRASTRUCTURE (oracle/).  Not part of the product: only tests/, bench.py's
 * cpu_baseline / `--impl reference` leg and __graft_entry__.smoke() may load it, and only
 * as the checker.  Parity is PINNED: tests/test_oracle_golden.py checks this file against
 * golden vectors produced by the unmodified reference (`oracle/make_golden.py`, which runs
 * /root/reference under the RNG injection of oracle/ref_harness.py).
 *
 * It follows the reference's own data structures literally (python lists / dicts become
 * small arrays in insertion order) so that every quirk is reproduced, and cites the
 * reference file:line each function restates (paths relative to /root/reference):
 *   environments/map_env.py    MapEnv.step :216-304, reset :306-342, update_moves :483-676,
 *                              update_custom_moves :678-693, update_map_fire :721-814,
 *                              spawn_point :816-827, spawn_rotation :829-832, color_view :397-411
 *   environments/Agent.py      action tables :8-16,161-162,198-199, consume/hit/fire_beam :178-234
 *   environments/cleanup_new.py  custom_reset :171-189, step :211-267, custom_action :269-292,
 *                              spawn_apples_and_waste :322-349, compute_probabilities :351-368,
 *                              feature caches :378-420
 *   environments/harvest_new.py  custom_reset :143-156, step :181-239, spawn_apples :284-317,
 *                              count_apples_in_radius :326-336
 *   contract/contract_list.py  CleanupContract :22-27, HarvestFeaturemodLocalContract :45-54
 *   environments/two_stage_train.py  SeparateContractEnv.step :62-121,
 *                              SeparateContractSubgameStage.reset :159-187
 * Randomness: every RNG call site is replaced by the counter-based Philox stream of
 * oracle/philox.py (same addressing), exactly as oracle/ref_harness.py injects it.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 8
#define MAXH 48
#define MAXW 64
#define VIEW 7
#define OBSW 15
#define MAXPTS (MAXH * MAXW)

enum { KIND_CLEANUP = 0, KIND_HARVEST = 1 };
enum { CONTRACT_NONE = 0, CONTRACT_CLEANUP = 1, CONTRACT_HARVEST_LOCAL = 2 };
enum { ORI_UP = 0, ORI_RIGHT = 1, ORI_DOWN = 2, ORI_LEFT = 3 };   /* Agent.py:18-23 */
enum {
    SITE_MOVE_ORDER = 1, SITE_BEAM_ORDER = 2, SITE_SPAWN_DRAWS = 3, SITE_WASTE_ORDER = 4,
    SITE_SPAWN_ROT = 5, SITE_SPAWN_POINT = 6, SITE_CONTRACT = 7, SITE_NEGOTIATE = 8
};

typedef struct { int16_t r, c; } pt;

/* tables derived from the ascii map; identical for every env of a batch */
typedef struct {
    char base_map[MAXH][MAXW];
    pt spawn_points[2 * 64]; int n_spawn;          /* canonical (sorted) order */
    pt apple_points[MAXPTS]; int n_apple;
    pt waste_points[MAXPTS]; int n_waste;          /* H or R cells, canonical (row-major) */
    pt waste_start[MAXPTS]; int n_waste_start;
    pt river[MAXPTS]; int n_river;
    pt stream[MAXPTS]; int n_stream;
    pt wall[MAXPTS]; int n_wall;
    int potential_waste_area;
} static_t;

typedef struct {
    /* config */
    int kind, n, H, W, horizon, contract;
    double theta_low, theta_high, null_prob;
    int reward_mode;                                /* 1 use_collective_reward | 2 inequity_averse_reward (map_env.py:289-301) */
    double alpha, beta;
    uint32_t seed, env_id;
    const static_t* s;
    /* dynamic state */
    char world_map[MAXH][MAXW];
    uint8_t color[MAXH + 2 * VIEW][MAXW + 2 * VIEW][3];   /* world_map_color, map_env.py:111 */
    pt pos[MAXN]; int ori[MAXN]; int reward_acc[MAXN]; int cleaned[MAXN];
    int timesteps; uint32_t episode;
    double p_apple, p_waste;
    pt cur_apples[MAXPTS]; int n_cur_apples;       /* current_apple_points (stale by one step) */
    pt cur_wastes[MAXPTS]; int n_cur_wastes;
    double theta;
    char beam[MAXH][MAXW];                          /* beam_pos of the last step as an overlay (0 = none; later entries win) */
    /* metrics accumulators (cleanup_new.py:186-189, harvest_new.py:152-156, two_stage_train.py:92-99) */
    double m_apples, m_low_density, m_raw, m_transfers, m_dirt;
    double m_agent_a[MAXN], m_agent_b[MAXN];       /* waste_cleaned | apples_consumed, close_apples_consumed */
    double sum_raw[MAXN], tsum_raw[MAXN], sum_tr[MAXN], tsum_tr[MAXN];
    int err;
} env_t;

/* ------------------------------------------------------------------ Philox4x32-10 */
static void philox4x32_10(const uint32_t c_in[4], const uint32_t k_in[2], uint32_t out[4])
{
    uint32_t c0 = c_in[0], c1 = c_in[1], c2 = c_in[2], c3 = c_in[3], k0 = k_in[0], k1 = k_in[1];
    for (int i = 0; i < 10; i++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void oracle_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }

static uint32_t draw_u32(const env_t* e, uint32_t t, int site, uint32_t call, uint32_t idx)
{
    uint32_t ctr[4] = { idx >> 2, (uint32_t)site | (call << 8), t, e->episode };
    uint32_t key[2] = { e->seed, e->env_id }, out[4];
    philox4x32_10(ctr, key, out);
    return out[idx & 3];
}
static double draw_f64(const env_t* e, uint32_t t, int site, uint32_t call, uint32_t idx)
{
    return (double)draw_u32(e, t, site, call, idx) * (1.0 / 4294967296.0);
}
/* stateless shuffle: order[j] = canonical index of the element at shuffled position j
 * (stable argsort of one u32 key per canonical element). */
static void shuffle_order(const env_t* e, uint32_t t, int site, uint32_t call, int n, int* order)
{
    uint32_t keys[MAXPTS];
    for (int i = 0; i < n; i++) { keys[i] = draw_u32(e, t, site, call, (uint32_t)i); order[i] = i; }
    for (int i = 1; i < n; i++) {           /* insertion sort = stable */
        int o = order[i]; int j = i - 1;
        while (j >= 0 && keys[order[j]] > keys[o]) { order[j + 1] = order[j]; j--; }
        order[j + 1] = o;
    }
}

/* ------------------------------------------------------------------ colours (map_env.py:24-42, cleanup_new.py:42-47) */
static void color_of(char ch, uint8_t rgb[3])
{
    int r = 0, g = 0, b = 0;
    switch (ch) {
    case ' ': case '0': break;
    case '@': r = 180; g = 180; b = 180; break;
    case 'A': g = 255; break;
    case 'F': r = 255; g = 255; break;
    case 'P': r = 159; g = 67; b = 255; break;
    case '1': b = 255; break;
    case '2': r = 2; g = 81; b = 154; break;
    case '3': r = 204; b = 204; break;
    case '4': r = 216; g = 30; b = 54; break;
    case '5': r = 254; g = 151; break;
    case '6': r = 100; g = 255; b = 255; break;
    case '7': r = 99; g = 99; b = 255; break;
    case '8': r = 250; g = 204; b = 255; break;
    case '9': r = 238; g = 223; b = 16; break;
    case 'C': r = 100; g = 255; b = 255; break;
    case 'S': case 'R': r = 113; g = 75; b = 24; break;
    case 'H': r = 99; g = 156; b = 194; break;
    default: break;
    }
    rgb[0] = (uint8_t)r; rgb[1] = (uint8_t)g; rgb[2] = (uint8_t)b;
}
static void single_update_world_color_map(env_t* e, int row, int col, char ch)   /* map_env.py:705-708 */
{
    color_of(ch, e->color[row + VIEW][col + VIEW]);
}
static void single_update_map(env_t* e, int row, int col, char ch)               /* map_env.py:701-703 */
{
    e->world_map[row][col] = ch;
    color_of(ch, e->color[row + VIEW][col + VIEW]);
}

/* ------------------------------------------------------------------ helpers */
static int in_agent_pos(const env_t* e, int r, int c)       /* `[r, c] in self.agent_pos` */
{
    for (int i = 0; i < e->n; i++) if (e->pos[i].r == r && e->pos[i].c == c) return 1;
    return 0;
}
/* dict {tuple(pos): id} built in agent order: a later agent overwrites an earlier one */
static int agent_by_pos_lookup(const pt* snap, int n, int r, int c)
{
    int found = -1;
    for (int i = 0; i < n; i++) if (snap[i].r == r && snap[i].c == c) found = i;
    return found;
}
static int is_tile_walkable(const env_t* e, int r, int c)   /* Agent.py:141-147; walls are static */
{
    return r >= 0 && r < e->H && c >= 0 && c < e->W && e->s->base_map[r][c] != '@';
}
static void update_agent_pos(env_t* e, int i, int r, int c) /* Agent.py:121-139 */
{
    if (is_tile_walkable(e, r, c)) { e->pos[i].r = r; e->pos[i].c = c; }
}
/* rotate_action (map_env.py:844-859): numpy dot with the TURN matrices
 * TURN_CLOCKWISE = [[0,1],[-1,0]]  TURN_COUNTERCLOCKWISE = [[0,-1],[1,0]] (map_env.py:17-18) */
static pt rotate_left(pt v) { pt o = { 0 * v.r + -1 * v.c, 1 * v.r + 0 * v.c }; return o; }
static pt rotate_right(pt v) { pt o = { 0 * v.r + 1 * v.c, -1 * v.r + 0 * v.c }; return o; }
static pt rotate_action(pt v, int ori)
{
    if (ori == ORI_UP) return v;
    if (ori == ORI_LEFT) return rotate_left(v);
    if (ori == ORI_RIGHT) return rotate_right(v);
    return rotate_left(rotate_left(v));
}
static int update_rotation(int ccw, int ori)                /* map_env.py:862-880 */
{
    if (ccw) {
        if (ori == ORI_LEFT) return ORI_DOWN;
        if (ori == ORI_DOWN) return ORI_RIGHT;
        if (ori == ORI_RIGHT) return ORI_UP;
        return ORI_LEFT;
    }
    if (ori == ORI_LEFT) return ORI_UP;
    if (ori == ORI_UP) return ORI_RIGHT;
    if (ori == ORI_RIGHT) return ORI_DOWN;
    return ORI_LEFT;
}
static pt orientation_vec(int ori)                          /* ORIENTATIONS, map_env.py:22 */
{
    pt v = { 0, 0 };
    if (ori == ORI_LEFT) v.c = -1; else if (ori == ORI_RIGHT) v.c = 1;
    else if (ori == ORI_UP) v.r = -1; else v.r = 1;
    return v;
}

/* action classes (Agent.py:8-16,161-162,198-199) */
enum { A_MOVE, A_TURN_CW, A_TURN_CCW, A_FIRE, A_CLEAN, A_BAD };
static int action_class(const env_t* e, int a, pt* vec)
{
    static const pt mv[5] = { {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {0, 0} };   /* map_env.py:11-16 */
    if (a >= 0 && a <= 4) { *vec = mv[a]; return A_MOVE; }
    if (a == 5) return A_TURN_CW;
    if (a == 6) return A_TURN_CCW;
    if (e->kind == KIND_HARVEST) return a == 7 ? A_FIRE : A_BAD;
    if (a == 7) return A_CLEAN;
    if (a == 8) return A_FIRE;
    return A_BAD;
}

/* ------------------------------------------------------------------ update_moves (map_env.py:483-676) */
static void update_moves(env_t* e, const int* acls, const pt* avec)
{
    int n = e->n;
    /* reserved_slots / agent_moves: dict in agent order (insertion order) */
    int has_move[MAXN]; pt mv[MAXN];
    int n_movers = 0;
    for (int i = 0; i < n; i++) {
        has_move[i] = 0;
        if (acls[i] == A_MOVE) {
            pt rot = rotate_action(avec[i], e->ori[i]);
            int nr = e->pos[i].r + rot.r, nc = e->pos[i].c + rot.c;
            if (!is_tile_walkable(e, nr, nc)) { nr = e->pos[i].r; nc = e->pos[i].c; }   /* return_valid_pos */
            has_move[i] = 1; mv[i].r = nr; mv[i].c = nc; n_movers++;
        } else if (acls[i] == A_TURN_CW || acls[i] == A_TURN_CCW) {
            e->ori[i] = update_rotation(acls[i] == A_TURN_CCW, e->ori[i]);
        }
    }
    if (n_movers == 0) return;

    /* shuffle_list = list(zip(agent_to_slot, move_slots)); np.random.shuffle (map_env.py:545-547) */
    int list_agent[MAXN]; pt list_slot[MAXN]; int m = 0;
    for (int i = 0; i < n; i++) if (has_move[i]) { list_agent[m] = i; list_slot[m] = mv[i]; m++; }
    int order[MAXN];
    shuffle_order(e, (uint32_t)e->timesteps, SITE_MOVE_ORDER, 0, m, order);
    int agent_to_slot[MAXN]; pt move_slots[MAXN];
    for (int j = 0; j < m; j++) { agent_to_slot[j] = list_agent[order[j]]; move_slots[j] = list_slot[order[j]]; }

    /* np.unique(move_slots, axis=0, return_index, return_counts): lexicographic rows */
    pt uniq[MAXN]; int uidx[MAXN], ucnt[MAXN]; int nu = 0;
    for (int j = 0; j < m; j++) {
        int k;
        for (k = 0; k < nu; k++) if (uniq[k].r == move_slots[j].r && uniq[k].c == move_slots[j].c) break;
        if (k == nu) { uniq[nu] = move_slots[j]; uidx[nu] = j; ucnt[nu] = 1; nu++; }
        else ucnt[k]++;
    }
    for (int a = 1; a < nu; a++) {          /* sort unique rows lexicographically */
        pt u = uniq[a]; int ui = uidx[a], uc = ucnt[a]; int b = a - 1;
        while (b >= 0 && (uniq[b].r > u.r || (uniq[b].r == u.r && uniq[b].c > u.c))) {
            uniq[b + 1] = uniq[b]; uidx[b + 1] = uidx[b]; ucnt[b + 1] = ucnt[b]; b--;
        }
        uniq[b + 1] = u; uidx[b + 1] = ui; ucnt[b + 1] = uc;
    }

    /* agent_by_pos (map_env.py:521), refreshed after each winner moves (:611) */
    pt abp[MAXN];
    memcpy(abp, e->pos, sizeof(abp));

    for (int k = 0; k < nu; k++) {
        if (ucnt[k] <= 1) continue;
        pt move = uniq[k];
        int conflict_cell_free = 1;
        for (int j = 0; j < m; j++) {
            if (!(move_slots[j].r == move.r && move_slots[j].c == move.c)) continue;
            int agent_id = agent_to_slot[j];
            /* moves_copy = agent_moves.copy(): membership == has_move (never deleted in this phase) */
            if (in_agent_pos(e, move.r, move.c)) {
                int conflicting = agent_by_pos_lookup(abp, n, move.r, move.c);
                if (conflicting < 0) { e->err |= 1; continue; }
                pt curr_pos = e->pos[agent_id];
                pt curr_conflict_pos = e->pos[conflicting];
                pt conflict_move = has_move[conflicting] ? mv[conflicting] : curr_conflict_pos;
                if (agent_id == conflicting) {
                    conflict_cell_free = 0;
                } else if (!has_move[conflicting] ||
                           (curr_conflict_pos.r == conflict_move.r && curr_conflict_pos.c == conflict_move.c)) {
                    conflict_cell_free = 0;
                } else if (has_move[conflicting]) {
                    if (mv[conflicting].r == curr_pos.r && mv[conflicting].c == curr_pos.c &&
                        move.r == e->pos[conflicting].r && move.c == e->pos[conflicting].c)
                        conflict_cell_free = 0;
                }
            }
        }
        if (conflict_cell_free) {
            update_agent_pos(e, agent_to_slot[uidx[k]], move.r, move.c);
            memcpy(abp, e->pos, sizeof(abp));
        }
        for (int j = 0; j < m; j++)
            if (move_slots[j].r == move.r && move_slots[j].c == move.c)
                mv[agent_to_slot[j]] = e->pos[agent_to_slot[j]];
    }

    /* make the remaining un-conflicted moves (map_env.py:624-676) */
    int live[MAXN]; memcpy(live, has_move, sizeof(live));
    for (;;) {
        int num_moves = 0;
        for (int i = 0; i < n; i++) num_moves += live[i];
        if (num_moves == 0) break;
        memcpy(abp, e->pos, sizeof(abp));
        int in_copy[MAXN]; pt mv_copy[MAXN]; int deleted[MAXN];
        memcpy(in_copy, live, sizeof(in_copy)); memcpy(mv_copy, mv, sizeof(mv_copy));
        memset(deleted, 0, sizeof(deleted));
        for (int i = 0; i < n; i++) {
            if (!in_copy[i]) continue;
            if (deleted[i]) continue;
            pt move = mv_copy[i];
            if (in_agent_pos(e, move.r, move.c)) {
                int conflicting = agent_by_pos_lookup(abp, n, move.r, move.c);
                if (conflicting < 0) { e->err |= 2; live[i] = 0; deleted[i] = 1; continue; }
                pt curr_pos = e->pos[i];
                pt curr_conflict_pos = e->pos[conflicting];
                pt conflict_move = live[conflicting] ? mv[conflicting] : curr_conflict_pos;
                if (i == conflicting) {
                    live[i] = 0; deleted[i] = 1;
                } else if (!in_copy[conflicting] ||
                           (curr_conflict_pos.r == conflict_move.r && curr_conflict_pos.c == conflict_move.c)) {
                    live[i] = 0; deleted[i] = 1;
                } else if (in_copy[conflicting]) {
                    if (!live[conflicting]) { e->err |= 4; continue; }   /* would be KeyError */
                    if (mv[conflicting].r == curr_pos.r && mv[conflicting].c == curr_pos.c &&
                        move.r == e->pos[conflicting].r && move.c == e->pos[conflicting].c) {
                        live[conflicting] = 0; live[i] = 0; deleted[i] = 1; deleted[conflicting] = 1;
                    }
                }
            } else {
                update_agent_pos(e, i, move.r, move.c);
                live[i] = 0; deleted[i] = 1;
            }
        }
        int left = 0;
        for (int i = 0; i < n; i++) left += live[i];
        if (left == num_moves) {
            for (int i = 0; i < n; i++) if (live[i]) update_agent_pos(e, i, mv[i].r, mv[i].c);
            break;
        }
    }
}

/* ------------------------------------------------------------------ update_map_fire (map_env.py:721-814) */
static int update_map_fire(env_t* e, int shooter, char fire_char, int clean, pt* updates)
{
    int n = e->n, n_up = 0;
    pt abp[MAXN]; memcpy(abp, e->pos, sizeof(abp));
    pt start = e->pos[shooter];
    pt dir = orientation_vec(e->ori[shooter]);
    pt right = rotate_right(dir);
    pt firing_pos[3] = {
        start,
        { start.r + right.r - dir.r, start.c + right.c - dir.c },
        { start.r - right.r - dir.r, start.c - right.c - dir.c } };
    for (int b = 0; b < 3; b++) {
        pt next = { firing_pos[b].r + dir.r, firing_pos[b].c + dir.c };
        for (int i = 0; i < 5; i++) {           /* fire_len = all_actions["FIRE"] = 5 (cleanup_new.py:39,277,285) */
            if (next.r >= 0 && next.r < e->H && next.c >= 0 && next.c < e->W &&
                e->world_map[next.r][next.c] != '@') {
                e->beam[next.r][next.c] = fire_char;                     /* firing_points -> beam_pos (:789,812) */
                if (clean && e->world_map[next.r][next.c] == 'H') { updates[n_up].r = next.r; updates[n_up].c = next.c; n_up++; }
                if (in_agent_pos(e, next.r, next.c)) {
                    int hit = agent_by_pos_lookup(abp, n, next.r, next.c);
                    if (fire_char == 'F') e->reward_acc[hit] -= 50;      /* Agent.py:178-180,224-226 */
                    break;
                }
                /* blocking_cells: [b"H"] for CLEAN; b"P" for FIRE (never present in world_map) */
                if (clean ? (e->world_map[next.r][next.c] == 'H') : (e->world_map[next.r][next.c] == 'P')) break;
                next.r += dir.r; next.c += dir.c;
            } else break;
        }
    }
    return n_up;
}

/* update_custom_moves (map_env.py:678-693) + custom_action (cleanup_new.py:269-292, harvest_new.py:241-249) */
static void update_custom_moves(env_t* e, const int* acls)
{
    int order[MAXN];
    shuffle_order(e, (uint32_t)e->timesteps, SITE_BEAM_ORDER, 0, e->n, order);
    for (int j = 0; j < e->n; j++) {
        int i = order[j];
        if (acls[i] != A_FIRE && acls[i] != A_CLEAN) continue;
        pt updates[3];
        int nu;
        if (acls[i] == A_FIRE) {
            e->reward_acc[i] -= 1;                                       /* fire_beam(b"F") */
            nu = update_map_fire(e, i, 'F', 0, updates);
        } else {
            nu = update_map_fire(e, i, 'C', 1, updates);
            e->cleaned[i] = nu;                                          /* cleanup_new.py:291 */
        }
        for (int u = 0; u < nu; u++) single_update_map(e, updates[u].r, updates[u].c, 'R');
    }
}

/* ------------------------------------------------------------------ cleanup spawning */
static void compute_probabilities(env_t* e)                 /* cleanup_new.py:351-376 */
{
    double waste_density = 0;
    if (e->s->potential_waste_area > 0) {
        int current_area = 0;
        for (int r = 0; r < e->H; r++) for (int c = 0; c < e->W; c++) current_area += e->world_map[r][c] == 'H';
        int free_area = e->s->potential_waste_area - current_area;
        waste_density = 1 - (double)free_area / (double)e->s->potential_waste_area;
    }
    if (waste_density >= 0.4) { e->p_apple = 0; e->p_waste = 0; }
    else {
        e->p_waste = 0.5;
        if (waste_density <= 0.0) e->p_apple = 0.05;
        else e->p_apple = (1 - (waste_density - 0.0) / (0.4 - 0.0)) * 0.05;
    }
}
static void cleanup_custom_map_update(env_t* e)             /* cleanup_new.py:294-297, 322-349 */
{
    compute_probabilities(e);
    uint32_t t = (uint32_t)e->timesteps;
    pt sp[MAXPTS]; char spc[MAXPTS]; int nsp = 0;
    uint32_t r = 0;
    for (int i = 0; i < e->s->n_apple; i++) {
        int row = e->s->apple_points[i].r, col = e->s->apple_points[i].c;
        if (!in_agent_pos(e, row, col) && e->world_map[row][col] != 'A') {
            double rand_num = draw_f64(e, t, SITE_SPAWN_DRAWS, 0, r); r++;
            if (rand_num < e->p_apple) { sp[nsp].r = row; sp[nsp].c = col; spc[nsp] = 'A'; nsp++; }
        }
    }
    if (e->p_waste != 0) {                                  /* not np.isclose(p, 0); p is 0 or 0.5 */
        int order[MAXPTS];
        shuffle_order(e, t, SITE_WASTE_ORDER, 0, e->s->n_waste, order);
        for (int i = 0; i < e->s->n_waste; i++) {
            int row = e->s->waste_points[order[i]].r, col = e->s->waste_points[order[i]].c;
            if (e->world_map[row][col] != 'H') {
                double rand_num = draw_f64(e, t, SITE_SPAWN_DRAWS, 0, r); r++;
                if (rand_num < e->p_waste) { sp[nsp].r = row; sp[nsp].c = col; spc[nsp] = 'H'; nsp++; break; }
            }
        }
    }
    for (int i = 0; i < nsp; i++) single_update_map(e, sp[i].r, sp[i].c, spc[i]);
}

/* ------------------------------------------------------------------ harvest spawning (harvest_new.py:284-317) */
static void harvest_custom_map_update(env_t* e)
{
    static const double SPAWN_PROB[4] = { 0, 0.005, 0.02, 0.05 };
    uint32_t t = (uint32_t)e->timesteps;
    pt sp[MAXPTS]; int nsp = 0;
    uint32_t r = 0;
    for (int i = 0; i < e->s->n_apple; i++) {
        int row = e->s->apple_points[i].r, col = e->s->apple_points[i].c;
        if (!in_agent_pos(e, row, col) && e->world_map[row][col] != 'A') {
            int num_apples = 0;
            for (int j = -2; j <= 2; j++) for (int k = -2; k <= 2; k++) {
                if (j * j + k * k <= 2) {                    /* APPLE_RADIUS = 2, not squared (harvest_new.py:303) */
                    int x = row + j, y = col + k;
                    if (x >= 0 && x < e->H && y >= 0 && y < e->W && e->world_map[x][y] == 'A') num_apples++;
                }
            }
            double spawn_prob = SPAWN_PROB[num_apples < 3 ? num_apples : 3];
            double rand_num = draw_f64(e, t, SITE_SPAWN_DRAWS, 0, r); r++;
            if (rand_num < spawn_prob) { sp[nsp].r = row; sp[nsp].c = col; nsp++; }
        }
    }
    for (int i = 0; i < nsp; i++) single_update_map(e, sp[i].r, sp[i].c, 'A');
}

/* ------------------------------------------------------------------ feature caches */
static void compute_current(env_t* e, char ch, pt* list, int* cnt)   /* cleanup_new.py:378-394 */
{
    int k = 0;
    for (int i = 0; i < e->H; i++) for (int j = 0; j < e->W; j++)
        if (e->world_map[i][j] == ch) { list[k].r = i; list[k].c = j; k++; }
    *cnt = k;
}
static pt closest_in(const pt* list, int cnt, pt p)         /* np.argmin of L1 distances: first minimum */
{
    pt best = { 0, 0 };                                      /* sentinel [0, 0] when the list is empty */
    int bd = 1 << 30;
    for (int i = 0; i < cnt; i++) {
        int d = abs(list[i].r - p.r) + abs(list[i].c - p.c);
        if (d < bd) { bd = d; best = list[i]; }
    }
    return best;
}
static int count_apples_in_radius5(const env_t* e, pt loc)  /* harvest_new.py:326-336, radius = 5 (not squared) */
{
    int num = 0;
    for (int j = -5; j <= 5; j++) for (int k = -5; k <= 5; k++) {
        if (j * j + k * k <= 5) {
            int r = loc.r + j, c = loc.c + k;
            for (int a = 0; a < e->n_cur_apples; a++)
                if (e->cur_apples[a].r == r && e->cur_apples[a].c == c) { num++; break; }
        }
    }
    return num;
}

/* ------------------------------------------------------------------ color_view (map_env.py:397-411) */
static void color_view(const env_t* e, int a, uint8_t* out /* [15][15][3] */)
{
    int row = e->pos[a].r, col = e->pos[a].c;                /* slice starts at (row, col) in padded coords */
    for (int i = 0; i < OBSW; i++) for (int j = 0; j < OBSW; j++) {
        int vi, vj;
        switch (e->ori[a]) {
        case ORI_UP:   vi = i; vj = j; break;
        case ORI_LEFT: vi = j; vj = OBSW - 1 - i; break;                 /* np.rot90(v)            */
        case ORI_DOWN: vi = OBSW - 1 - i; vj = OBSW - 1 - j; break;      /* np.rot90(v, k=2)       */
        default:       vi = OBSW - 1 - j; vj = i; break;                 /* np.rot90(v, 1, (1,0))  */
        }
        memcpy(out + (i * OBSW + j) * 3, e->color[row + vi][col + vj], 3);
    }
}

/* ------------------------------------------------------------------ reset */
static void setup_agents(env_t* e)                           /* cleanup_new.py:302-320, map_env.py:816-832 */
{
    static const int rot_of[4] = { ORI_LEFT, ORI_RIGHT, ORI_UP, ORI_DOWN };   /* list(ORIENTATIONS.keys()) */
    for (int i = 0; i < e->n; i++) {
        int order[2 * 64];
        shuffle_order(e, 0, SITE_SPAWN_POINT, (uint32_t)i, e->s->n_spawn, order);
        int spawn_index = 0;
        for (int j = 0; j < e->s->n_spawn; j++) {
            pt s = e->s->spawn_points[order[j]];
            int taken = 0;
            for (int k = 0; k < i; k++) if (e->pos[k].r == s.r && e->pos[k].c == s.c) taken = 1;
            if (!taken) spawn_index = j;                     /* no break: the LAST free entry wins */
        }
        e->pos[i] = e->s->spawn_points[order[spawn_index]];
        e->ori[i] = rot_of[draw_u32(e, 0, SITE_SPAWN_ROT, (uint32_t)i, 0) >> 30];
        e->reward_acc[i] = 0; e->cleaned[i] = 0;
    }
}
static void env_reset(env_t* e, uint32_t episode)
{
    e->episode = episode;
    e->timesteps = 0;
    memset(e->beam, 0, sizeof(e->beam));                     /* self.beam_pos = [] (map_env.py:316) */
    setup_agents(e);
    /* reset_map (map_env.py:710-719) */
    for (int r = 0; r < e->H; r++) for (int c = 0; c < e->W; c++) e->world_map[r][c] = ' ';
    memset(e->color, 0, sizeof(e->color));
    for (int i = 0; i < e->s->n_wall; i++) single_update_map(e, e->s->wall[i].r, e->s->wall[i].c, '@');
    if (e->kind == KIND_CLEANUP) {                           /* cleanup_new.py:171-189 */
        for (int i = 0; i < e->s->n_waste_start; i++) single_update_map(e, e->s->waste_start[i].r, e->s->waste_start[i].c, 'H');
        for (int i = 0; i < e->s->n_river; i++) single_update_map(e, e->s->river[i].r, e->s->river[i].c, 'R');
        for (int i = 0; i < e->s->n_stream; i++) single_update_map(e, e->s->stream[i].r, e->s->stream[i].c, 'S');
        compute_current(e, 'A', e->cur_apples, &e->n_cur_apples);
        compute_current(e, 'H', e->cur_wastes, &e->n_cur_wastes);
    } else {                                                 /* harvest_new.py:143-156 */
        for (int i = 0; i < e->s->n_apple; i++) single_update_map(e, e->s->apple_points[i].r, e->s->apple_points[i].c, 'A');
        compute_current(e, 'A', e->cur_apples, &e->n_cur_apples);
    }
    e->m_apples = e->m_low_density = e->m_raw = e->m_transfers = e->m_dirt = 0;
    for (int i = 0; i < MAXN; i++) {
        e->m_agent_a[i] = e->m_agent_b[i] = 0;
        e->sum_raw[i] = e->tsum_raw[i] = e->sum_tr[i] = e->tsum_tr[i] = 0;
    }
    /* custom_map_update at reset (map_env.py:320): draws are addressed with t = 0 */
    if (e->kind == KIND_CLEANUP) cleanup_custom_map_update(e); else harvest_custom_map_update(e);
    /* SeparateContractSubgameStage.reset (two_stage_train.py:159-168) */
    if (e->contract != CONTRACT_NONE) {
        double u0 = draw_f64(e, 0, SITE_CONTRACT, 0, 0);
        if (u0 > e->null_prob) {
            double u1 = draw_f64(e, 0, SITE_CONTRACT, 0, 1);
            e->theta = e->theta_low + (e->theta_high - e->theta_low) * u1;
        } else e->theta = e->theta_low;
    } else e->theta = 0;
}

/* ------------------------------------------------------------------ step */
typedef struct {
    uint8_t* obs;        /* [n][15][15][3] */
    double* rew;         /* [n] rewards after contract transfers (== base rewards without contract) */
    double* base_rew;    /* [n] */
    double* transfers;   /* [n] */
    int32_t* info;       /* [n][4]: eaten_apples, cleaned_squares | eaten_close_apples, total_close_apples, 0 */
    double* feat;        /* [n][F] feature_obs, F = 12+n (cleanup) or 10+2n (harvest); may be NULL */
    uint8_t* done;       /* [1] */
} step_out;

/* use_collective_reward / inequity_averse_reward (map_env.py:289-301).  The env rewards are Python ints, so
 * every difference and both partial sums are exact; alpha * sum, beta * sum, their sum, the division by
 * (num_agents - 1) and the final subtraction are float64 operations in that order. */
static void shape_rewards(const env_t* e, double* r)
{
    int n = e->n;
    if (e->reward_mode & 1) {
        double s = 0;
        for (int i = 0; i < n; i++) s += r[i];
        for (int i = 0; i < n; i++) r[i] = s;
    }
    if (e->reward_mode & 2) {
        double tmp[MAXN];
        for (int i = 0; i < n; i++) {
            double sp = 0, sn = 0;
            for (int j = 0; j < n; j++) {
                double d = r[j] - r[i];
                if (d > 0) sp += d; else if (d < 0) sn += d;
            }
            volatile double dis = e->alpha * sp, adv = e->beta * sn;
            volatile double q = (dis + adv) / (double)(n - 1);
            tmp[i] = r[i] - q;
        }
        for (int i = 0; i < n; i++) r[i] = tmp[i];
    }
}

static void env_step(env_t* e, const int32_t* actions, step_out* o, int want_feat)
{
    int n = e->n;
    int acls[MAXN]; pt avec[MAXN];
    e->timesteps += 1;                                       /* map_env.py:230 */
    memset(e->beam, 0, sizeof(e->beam));                     /* self.beam_pos = [] (:231) */
    for (int i = 0; i < n; i++) {
        avec[i].r = avec[i].c = 0;
        acls[i] = action_class(e, actions[i], &avec[i]);
        if (acls[i] == A_BAD) { e->err |= 8; acls[i] = A_MOVE; }        /* reference would KeyError */
    }
    for (int i = 0; i < n; i++)                              /* un-paint agents (map_env.py:238-240) */
        single_update_world_color_map(e, e->pos[i].r, e->pos[i].c, e->world_map[e->pos[i].r][e->pos[i].c]);
    update_moves(e, acls, avec);
    for (int i = 0; i < n; i++) {                            /* consume (map_env.py:244-247, Agent.py:228-234) */
        int r = e->pos[i].r, c = e->pos[i].c;
        char ch = e->world_map[r][c];
        if (ch == 'A') { e->reward_acc[i] += 1; ch = ' '; }
        single_update_map(e, r, c, ch);
    }
    update_custom_moves(e, acls);
    if (e->kind == KIND_CLEANUP) cleanup_custom_map_update(e); else harvest_custom_map_update(e);
    for (int i = 0; i < n; i++)                              /* paint agents (map_env.py:257-261) */
        single_update_world_color_map(e, e->pos[i].r, e->pos[i].c, (char)('1' + i));
    double base[MAXN];
    for (int i = 0; i < n; i++) {
        color_view(e, i, o->obs + (size_t)i * OBSW * OBSW * 3);
        base[i] = (double)e->reward_acc[i]; e->reward_acc[i] = 0;       /* compute_reward, Agent.py:87-90 */
    }
    shape_rewards(e, base);

    /* CleanupEnv.step / HarvestEnv.step tails (cleanup_new.py:213-253, harvest_new.py:183-224) */
    int eaten[MAXN], cleaned[MAXN], eaten_close[MAXN], total_close[MAXN];
    for (int i = 0; i < n; i++) {
        eaten[i] = eaten_close[i] = total_close[i] = 0;
        cleaned[i] = e->cleaned[i]; e->cleaned[i] = 0;
        if (e->kind == KIND_CLEANUP) { e->m_dirt += cleaned[i]; e->m_agent_a[i] += cleaned[i]; }
    }
    for (int i = 0; i < n; i++) {
        int in_list = 0;
        for (int a = 0; a < e->n_cur_apples; a++)
            if (e->cur_apples[a].r == e->pos[i].r && e->cur_apples[a].c == e->pos[i].c) { in_list = 1; break; }
        if (in_list) {
            eaten[i] += 1;
            if (e->kind == KIND_HARVEST) {
                e->m_agent_a[i] += 1;
                if (count_apples_in_radius5(e, e->pos[i]) < 4) { eaten_close[i] += 1; e->m_low_density += 1; e->m_agent_b[i] += 1; }
            }
            e->m_apples += 1;
        }
    }
    double raw = 0;
    for (int i = 0; i < n; i++) raw += base[i];
    e->m_raw += raw;
    for (int i = 0; i < n; i++) {                            /* total_reward_dict -> sustainability/equality sums */
        e->sum_raw[i] += base[i];
        e->tsum_raw[i] += (double)(e->timesteps - 1) * base[i];
    }
    compute_current(e, 'A', e->cur_apples, &e->n_cur_apples);
    if (e->kind == KIND_CLEANUP) compute_current(e, 'H', e->cur_wastes, &e->n_cur_wastes);
    if (e->kind == KIND_HARVEST)
        for (int i = 0; i < n; i++) total_close[i] = count_apples_in_radius5(e, e->pos[i]);
    int done = e->timesteps == e->horizon;                   /* cleanup_new.py:242 */

    if (want_feat && o->feat) {
        int F = e->kind == KIND_CLEANUP ? 12 + n : 10 + 2 * n;
        for (int i = 0; i < n; i++) {
            double* f = o->feat + (size_t)i * F;
            int cp = (i == 0) ? (n > 1 ? 1 : 0) : 0;         /* compute_closest_pos quirk, cleanup_new.py:405-412 */
            pt ca = closest_in(e->cur_apples, e->n_cur_apples, e->pos[i]);
            f[0] = e->pos[i].r; f[1] = e->pos[i].c; f[2] = e->ori[i];
            f[3] = e->pos[cp].r; f[4] = e->pos[cp].c; f[5] = e->ori[cp];
            f[6] = ca.r; f[7] = ca.c;
            if (e->kind == KIND_CLEANUP) {
                pt cw = closest_in(e->cur_wastes, e->n_cur_wastes, e->pos[i]);
                f[8] = cw.r; f[9] = cw.c; f[10] = e->n_cur_apples; f[11] = e->n_cur_wastes;
                for (int j = 0; j < n; j++) f[12 + j] = cleaned[j];
            } else {
                f[8] = total_close[i]; f[9] = e->n_cur_apples;
                for (int j = 0; j < 2 * n; j++) f[10 + j] = 0.0;
            }
        }
    }

    /* contract transfers + redistribution (contract_list.py:22-27,45-54; two_stage_train.py:69-92) */
    double rews[MAXN], tr[MAXN];
    for (int i = 0; i < n; i++) { rews[i] = base[i]; tr[i] = 0; }
    if (e->contract != CONTRACT_NONE) {
        for (int i = 0; i < n; i++) {
            if (e->contract == CONTRACT_CLEANUP) tr[i] = (-e->theta) * (double)cleaned[i];
            else tr[i] = (total_close[i] < 4 && eaten_close[i] > 0) ? e->theta : 0.0;
        }
        double total_transfers = 0;
        for (int i = 0; i < n; i++) {
            rews[i] -= tr[i];
            total_transfers += tr[i];
            for (int j = 0; j < n; j++) if (i != j) rews[j] += tr[i] / (double)(n - 1);
        }
        e->m_transfers += total_transfers;
        for (int i = 0; i < n; i++) {
            e->sum_tr[i] += rews[i];
            e->tsum_tr[i] += (double)(e->timesteps - 1) * rews[i];
        }
    }
    for (int i = 0; i < n; i++) {
        o->rew[i] = rews[i]; o->base_rew[i] = base[i]; o->transfers[i] = tr[i];
        o->info[i * 4 + 0] = eaten[i];
        o->info[i * 4 + 1] = e->kind == KIND_CLEANUP ? cleaned[i] : eaten_close[i];
        o->info[i * 4 + 2] = total_close[i];
        o->info[i * 4 + 3] = 0;
    }
    o->done[0] = (uint8_t)done;
}

/* ------------------------------------------------------------------ construction */
static int pt_less(pt a, pt b) { return a.r < b.r || (a.r == b.r && a.c < b.c); }
static void sort_pts(pt* p, int n)
{
    for (int i = 1; i < n; i++) { pt x = p[i]; int j = i - 1; while (j >= 0 && pt_less(x, p[j])) { p[j + 1] = p[j]; j--; } p[j + 1] = x; }
}
static int static_init(static_t* s, int kind, int H, int W, const char* ascii)
{
    memset(s, 0, sizeof(*s));
    for (int r = 0; r < H; r++) for (int c = 0; c < W; c++) {
        char ch = ascii[r * W + c];
        s->base_map[r][c] = ch;
        pt p = { (int16_t)r, (int16_t)c };
        if (ch == 'P') {
            if (s->n_spawn >= 64) return -2;
            s->spawn_points[s->n_spawn++] = p;
        } else if (ch == '@') s->wall[s->n_wall++] = p;
        if (kind == KIND_CLEANUP) {
            if (ch == 'B') s->apple_points[s->n_apple++] = p;
            if (ch == 'S') s->stream[s->n_stream++] = p;
            if (ch == 'H') s->waste_start[s->n_waste_start++] = p;
            if (ch == 'H' || ch == 'R') s->waste_points[s->n_waste++] = p;
            if (ch == 'R') s->river[s->n_river++] = p;
        } else if (ch == 'A') s->apple_points[s->n_apple++] = p;
    }
    if (kind == KIND_CLEANUP) {
        /* CleanupEnv.__init__ appends every spawn point a second time (cleanup_new.py:114-115
         * after map_env.py:127-128); the stateless shuffle canonicalises by sorting. */
        int k = s->n_spawn;
        for (int i = 0; i < k; i++) s->spawn_points[s->n_spawn++] = s->spawn_points[i];
        sort_pts(s->spawn_points, s->n_spawn);
        s->potential_waste_area = s->n_waste;            /* count(H) + count(R), cleanup_new.py:98-100 */
    }
    if (s->n_spawn < 1) return -3;
    return 0;
}
static int env_init(env_t* e, const static_t* s, int kind, int n, int H, int W, int horizon, int contract,
                    double theta_low, double theta_high, double null_prob, uint32_t seed, uint32_t env_id)
{
    memset(e, 0, sizeof(*e));
    e->kind = kind; e->n = n; e->H = H; e->W = W; e->horizon = horizon; e->contract = contract;
    e->theta_low = theta_low; e->theta_high = theta_high; e->null_prob = null_prob;
    e->seed = seed; e->env_id = env_id; e->s = s;
    return 0;
}

/* ================================================================== batch C interface (ctypes) */
typedef struct {
    int E, n, H, W, kind, F;
    env_t* envs;
    static_t st;
} batch_t;

void* oracle_create(int kind, int E, int n, int H, int W, const char* ascii, int horizon, int contract,
                    double theta_low, double theta_high, double null_prob, uint32_t seed, uint32_t first_env_id)
{
    if (n < 1 || n > MAXN || H < 1 || H > MAXH || W < 1 || W > MAXW || E < 1) return NULL;
    batch_t* b = (batch_t*)calloc(1, sizeof(batch_t));
    if (!b) return NULL;
    b->E = E; b->n = n; b->H = H; b->W = W; b->kind = kind;
    if (static_init(&b->st, kind, H, W, ascii) != 0) { free(b); return NULL; }
    b->F = kind == KIND_CLEANUP ? 12 + n : 10 + 2 * n;
    b->envs = (env_t*)calloc((size_t)E, sizeof(env_t));
    if (!b->envs) { free(b); return NULL; }
    for (int i = 0; i < E; i++)
        env_init(&b->envs[i], &b->st, kind, n, H, W, horizon, contract, theta_low, theta_high, null_prob,
                 seed, first_env_id + (uint32_t)i);
    return b;
}
/* MapEnv kwargs use_collective_reward, inequity_averse_reward, alpha, beta (map_env.py:69-72) */
int oracle_set_reward_shaping(void* h, int mode, double alpha, double beta)
{
    batch_t* b = (batch_t*)h;
    if (!b || mode < 0 || mode > 3 || ((mode & 2) && b->n < 2)) return -1;     /* assert num_agents > 1, map_env.py:294 */
    for (int i = 0; i < b->E; i++) { b->envs[i].reward_mode = mode; b->envs[i].alpha = alpha; b->envs[i].beta = beta; }
    return 0;
}
void oracle_destroy(void* h) { batch_t* b = (batch_t*)h; if (b) { free(b->envs); free(b); } }

/* reset every env (mask == NULL) or those with mask[i] != 0; obs: [E][n][15][15][3] (reset obs
 * shows no agents: MapEnv.reset never paints them, map_env.py:306-342). */
void oracle_reset(void* h, const uint8_t* mask, const uint32_t* episode, uint8_t* obs)
{
    batch_t* b = (batch_t*)h;
    size_t ob = (size_t)b->n * OBSW * OBSW * 3;
#pragma omp parallel for schedule(static) if (b->E > 1)   /* a one-env batch (the live differential tests) stays on the calling thread */
    for (int i = 0; i < b->E; i++) {
        if (mask && !mask[i]) continue;
        env_t* e = &b->envs[i];
        env_reset(e, episode[i]);
        if (obs) for (int a = 0; a < b->n; a++) color_view(e, a, obs + i * ob + (size_t)a * OBSW * OBSW * 3);
    }
}

void oracle_step(void* h, const int32_t* actions, uint8_t* obs, double* rew, double* base_rew, double* transfers,
                 int32_t* info, double* feat, uint8_t* done)
{
    batch_t* b = (batch_t*)h;
    int n = b->n;
    size_t ob = (size_t)n * OBSW * OBSW * 3;
#pragma omp parallel for schedule(static) if (b->E > 1)   /* a one-env batch (the live differential tests) stays on the calling thread */
    for (int i = 0; i < b->E; i++) {
        step_out o;
        o.obs = obs + i * ob; o.rew = rew + (size_t)i * n; o.base_rew = base_rew + (size_t)i * n;
        o.transfers = transfers + (size_t)i * n; o.info = info + (size_t)i * n * 4;
        o.feat = feat ? feat + (size_t)i * n * b->F : NULL; o.done = done + i;
        env_step(&b->envs[i], actions + (size_t)i * n, &o, feat != NULL);
    }
}

/* state access: map chars [E][H][W], pos [E][n][2], ori [E][n], t [E], theta [E] */
void oracle_get_state(void* h, uint8_t* map, int32_t* pos, int32_t* ori, int32_t* t, double* theta)
{
    batch_t* b = (batch_t*)h;
    for (int i = 0; i < b->E; i++) {
        env_t* e = &b->envs[i];
        for (int r = 0; r < b->H; r++) for (int c = 0; c < b->W; c++) map[((size_t)i * b->H + r) * b->W + c] = (uint8_t)e->world_map[r][c];
        for (int a = 0; a < b->n; a++) { pos[((size_t)i * b->n + a) * 2] = e->pos[a].r; pos[((size_t)i * b->n + a) * 2 + 1] = e->pos[a].c; ori[(size_t)i * b->n + a] = e->ori[a]; }
        t[i] = e->timesteps; theta[i] = e->theta;
    }
}
/* any pointer may be NULL (left unchanged).  Mirrors oracle/ref_harness.RefGridEnv.set_state:
 * the colour grid and the stale apple/waste lists are rebuilt from the new map. */
void oracle_set_state(void* h, const uint8_t* map, const int32_t* pos, const int32_t* ori, const int32_t* t, const double* theta)
{
    batch_t* b = (batch_t*)h;
    for (int i = 0; i < b->E; i++) {
        env_t* e = &b->envs[i];
        if (map) for (int r = 0; r < b->H; r++) for (int c = 0; c < b->W; c++) single_update_map(e, r, c, (char)map[((size_t)i * b->H + r) * b->W + c]);
        else for (int a = 0; a < b->n; a++) single_update_world_color_map(e, e->pos[a].r, e->pos[a].c, e->world_map[e->pos[a].r][e->pos[a].c]);
        if (pos) for (int a = 0; a < b->n; a++) { e->pos[a].r = pos[((size_t)i * b->n + a) * 2]; e->pos[a].c = pos[((size_t)i * b->n + a) * 2 + 1]; }
        if (ori) for (int a = 0; a < b->n; a++) e->ori[a] = ori[(size_t)i * b->n + a];
        if (t) e->timesteps = t[i];
        if (theta) e->theta = theta[i];
        compute_current(e, 'A', e->cur_apples, &e->n_cur_apples);
        if (e->kind == KIND_CLEANUP) compute_current(e, 'H', e->cur_wastes, &e->n_cur_wastes);
        for (int a = 0; a < b->n; a++) single_update_world_color_map(e, e->pos[a].r, e->pos[a].c, (char)('1' + a));
    }
}
/* metrics accumulators, 8 + 6*MAXN doubles per env:
 * [apples, low_density, raw, transfers, dirt, err, 0, 0, agent_a[8], agent_b[8], sum_raw[8], tsum_raw[8], sum_tr[8], tsum_tr[8]] */
void oracle_get_metrics(void* h, double* out)
{
    batch_t* b = (batch_t*)h;
    for (int i = 0; i < b->E; i++) {
        env_t* e = &b->envs[i];
        double* o = out + (size_t)i * (8 + 6 * MAXN);
        o[0] = e->m_apples; o[1] = e->m_low_density; o[2] = e->m_raw; o[3] = e->m_transfers; o[4] = e->m_dirt;
        o[5] = e->err; o[6] = o[7] = 0;
        for (int a = 0; a < MAXN; a++) {
            o[8 + a] = e->m_agent_a[a]; o[8 + MAXN + a] = e->m_agent_b[a];
            o[8 + 2 * MAXN + a] = e->sum_raw[a]; o[8 + 3 * MAXN + a] = e->tsum_raw[a];
            o[8 + 4 * MAXN + a] = e->sum_tr[a]; o[8 + 5 * MAXN + a] = e->tsum_tr[a];
        }
    }
}
/* MapEnv.global_view() (map_env.py:394-395): world_map_color without its padding, uint8 [E][H][W][3].  Agents
 * appear in it from the first step on: MapEnv.reset (map_env.py:306-342) never paints them into the colour map. */
void oracle_global_view(void* h, uint8_t* out)
{
    batch_t* b = (batch_t*)h;
    for (int i = 0; i < b->E; i++) {
        env_t* e = &b->envs[i];
        uint8_t* o = out + (size_t)i * b->H * b->W * 3;
        for (int r = 0; r < b->H; r++)
            for (int c = 0; c < b->W; c++)
                memcpy(o + ((size_t)r * b->W + c) * 3, e->color[r + VIEW][c + VIEW], 3);
    }
}
/* MapEnv.full_map_to_colors() (map_env.py:389-392) = render(mode != 'human'): the char grid with agent ids, then the
 * beams of the last step on top (get_map_with_agents :354-375), through the colour map.  uint8 [E][H][W][3]. */
void oracle_render(void* h, uint8_t* out)
{
    batch_t* b = (batch_t*)h;
    for (int i = 0; i < b->E; i++) {
        env_t* e = &b->envs[i];
        uint8_t* o = out + (size_t)i * b->H * b->W * 3;
        for (int r = 0; r < b->H; r++)
            for (int c = 0; c < b->W; c++) color_of(e->world_map[r][c], o + ((size_t)r * b->W + c) * 3);
        for (int a = 0; a < e->n; a++) color_of((char)('1' + a), o + ((size_t)e->pos[a].r * b->W + e->pos[a].c) * 3);
        for (int r = 0; r < b->H; r++)
            for (int c = 0; c < b->W; c++) if (e->beam[r][c]) color_of(e->beam[r][c], o + ((size_t)r * b->W + c) * 3);
    }
}
/* threads of the OpenMP loops over envs (torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants every core) */
int oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}
int oracle_feature_dim(void* h) { return ((batch_t*)h)->F; }

/* Agreement stage of SeparateContractNegotiateStage.step (two_stage_train.py:266-281).
 * proposals [E] (a0's action[:-1]), accept [E][n] (every agent's action[-1]), decision [E] out.
 * n > 3: chosen = random.sample(range(1, n), 2) = the first two entries of the stateless shuffle of
 * [1 .. n-1] (site NEGOTIATE call 0); prod = 1 * accept[chosen[0]] * accept[chosen[1]] in that order;
 * decision = random.random() (site NEGOTIATE call 1) < prod; theta = proposal if accepted else 0. */
void oracle_negotiate(void* h, const double* proposals, const double* accept, uint8_t* decision)
{
    batch_t* b = (batch_t*)h;
    int n = b->n;
    for (int i = 0; i < b->E; i++) {
        env_t* e = &b->envs[i];
        double prod = 1;
        if (n > 3) {
            int order[MAXN];
            shuffle_order(e, 0, SITE_NEGOTIATE, 0, n - 1, order);
            prod *= accept[(size_t)i * n + 1 + order[0]];
            prod *= accept[(size_t)i * n + 1 + order[1]];
        } else {
            for (int a = 1; a < n; a++) prod *= accept[(size_t)i * n + a];
        }
        double r = draw_f64(e, 0, SITE_NEGOTIATE, 1, 0);
        int dec = r < prod;
        e->theta = dec ? proposals[i] : 0.0;
        if (decision) decision[i] = (uint8_t)dec;
    }
}
