/*
 * ssd_b200.h — C ABI of libssd_b200.so, the B200 (sm_100a) batched simulator for the
 * sequential-social-dilemma environments of Algorithmic-Alignment-Lab/contracts.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own: the
 * interface it exposes for this path is the Python MultiAgentEnv `reset()/step()` of the
 * classes built by `utils/env_creator_functions.py:12-34`.  Each entry point below replaces
 * the batched equivalent of one reference method (cited per function; paths relative to the
 * reference root).  The Python binding a maintainer adds is a ctypes stub — see
 * INTEGRATION.md and contracts_b200/_lib.py.
 *
 * Conventions
 *   - plain C: opaque handle, POD config, raw pointers + sizes, no C++/torch types;
 *   - every function returns 0 on success or a negative SSD_E* code; the message is
 *     available through ssd_last_error(handle) (or ssd_last_error(NULL) for create errors);
 *   - all `_dev` pointers are device pointers owned by the CALLER (e.g. torch tensors via
 *     tensor.data_ptr()); the library owns only the persistent per-env state;
 *   - every call is asynchronous on the given `stream` (a cudaStream_t passed as void*;
 *     NULL = legacy default stream) and performs no host synchronisation;
 *   - a handle is bound to one device and is not thread-safe; calls must be stream-ordered.
 *   - env i of a handle has global id `first_env_id + i`; its random stream is
 *     Philox4x32-10 keyed (seed, global id), so trajectories do not depend on how envs are
 *     sharded over handles / GPUs.
 */
#ifndef SSD_B200_H
#define SSD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSD_ABI_VERSION 3

/* env_kind: which reference class the handle simulates */
#define SSD_ENV_CLEANUP 0           /* environments/cleanup_new.py  CleanupEnv   ('CleanupNew') */
#define SSD_ENV_HARVEST 1           /* environments/harvest_new.py  HarvestEnv   ('HarvestNew') */
#define SSD_ENV_CLEANUP_FEATURES 2  /* environments/cleanup_features.py CleanupFeatures ('Cleanup') */
#define SSD_ENV_HARVEST_FEATURES 3  /* environments/harvest_features.py HarvestFeatures ('Harvest') */
#define SSD_ENV_SELFDRIVE 4         /* environments/self_driving_car_accelerate.py ('SelfDrive') */

/* contract_kind: contract/contract_list.py classes fused into the reward write */
#define SSD_CONTRACT_NONE 0
#define SSD_CONTRACT_CLEANUP 1              /* CleanupContract                 :7-27  */
#define SSD_CONTRACT_HARVEST_LOCAL 2        /* HarvestFeaturemodLocalContract  :29-54 */
#define SSD_CONTRACT_SELFDRIVE_DISTPROP 3   /* SelfdriveContractDistprop       :56-102 */

#define SSD_MAX_AGENTS 8
#define SSD_OBS_BYTES_PER_AGENT 675         /* 15 x 15 x 3 uint8 (map_env.py:397-411) */
#define SSD_METRIC_STRIDE 56                /* doubles per env written by ssd_get_metrics */

#define SSD_OK 0
#define SSD_EINVAL (-1)
#define SSD_ENOMEM (-2)
#define SSD_ECUDA (-3)
#define SSD_EUNSUPPORTED (-4)

typedef struct ssd_handle ssd_handle;

/* ssd_config.flags, cleanup_new / harvest_new: MapEnv kwargs use_collective_reward and
 * inequity_averse_reward (map_env.py:69-70): reward shaping at the end of MapEnv.step (map_env.py:289-301) */
#define SSD_FLAG_COLLECTIVE_REWARD 1
#define SSD_FLAG_INEQUITY_AVERSE 2

typedef struct ssd_config {
    int32_t abi_version;      /* SSD_ABI_VERSION */
    int32_t struct_size;      /* sizeof(ssd_config) as the caller compiled / declared it: ssd_create rejects a struct
                                 that is shorter than the library's (a stale binding would otherwise be read past its end) */
    int32_t env_kind;
    int32_t num_envs;         /* E, envs stepped per launch by this handle */
    int32_t num_agents;       /* n in [1, 8] */
    int32_t map_h, map_w;     /* gridworlds: ascii map size; ignored by selfdrive */
    const char* ascii_map;    /* map_h * map_w chars, row-major (CLEANUP_MAP / HARVEST_MAP layout) */
    int32_t horizon;          /* env `horizon` kwarg (cleanup_new.py:71): done when t == horizon */
    int32_t contract_kind;
    double theta_low;         /* contract_space.low[0]  (a float32 value, e.g. 0) */
    double theta_high;        /* contract_space.high[0] (a float32 value, e.g. float32(0.2)) */
    double null_prob;         /* SeparateContractSubgameStage null_prob (two_stage_train.py:152-166) */
    uint32_t seed;            /* Philox key word 0 */
    uint32_t first_env_id;    /* global id of env 0 (Philox key word 1 = first_env_id + i) */
    int32_t device;           /* CUDA device ordinal */
    int32_t flags;            /* SSD_FLAG_* bits (gridworlds), else 0 */
    double env_params[8];     /* selfdrive: low_bound, high_bound, start_vel, start_vel_ambulance
                                 (SelfAcceleratingCarEnv.__init__, self_driving_car_accelerate.py:19);
                                 cleanup_new / harvest_new: alpha, beta of inequity_averse_reward
                                 (map_env.py:71-72,293-300); else unused */
} ssd_config;

/* Buffers of one step.  Gridworld / feature envs: actions are uint8 [E][n] action ids
 * (Agent.py:8-16,161-162,198-199); selfdrive: float32 [E][n] accelerations.
 * NULL is allowed for every output except obs and rew (rew_dev may be NULL in ssd_step_host_async, whose result
 * block carries the rewards). */
typedef struct ssd_step_io {
    const void* actions_dev;
    uint8_t* obs_dev;         /* gridworlds: uint8 [E][n][15][15][3] (MapEnv.step 'curr_obs', map_env.py:269,284) */
    int64_t obs_env_stride;   /* bytes between consecutive envs in obs_dev (>= n*675; 0 = dense) */
    double* rew_dev;          /* [E][n] rewards AFTER contract transfers (two_stage_train.py:69-90) */
    double* base_rew_dev;     /* [E][n] env rewards before transfers (MapEnv.step rewards, map_env.py:285) */
    double* transfers_dev;    /* [E][n] Contract.compute_transfer values (contract_list.py) */
    uint8_t* info_dev;        /* [E][n][4]: eaten_apples, cleaned_squares|eaten_close_apples, total_close_apples, 0 */
    double* feature_obs_dev;  /* [E][n][F] infos['feature_obs'] (cleanup_new.py:243-251 F=12+n, harvest_new.py:215-222 F=10+2n) */
    uint8_t* done_dev;        /* [E] dones['__all__'] (cleanup_new.py:242) */
    /* Vectorised-sampler mode (what RLlib's rollout worker does with a vector env, utils/ray_config_utils.py:140: reset
     * an env as soon as it is done).  auto_reset != 0: every env that finishes in this step (t == horizon) is reset
     * by the step itself, exactly as ssd_reset with mask = done would (same draws, same episode statistics hand-off):
     * its rewards / dones / infos are those of the final step, its observation is the RESET observation of the next
     * episode.  With neg_proposals_dev / neg_accept_dev (double [E] / [E][n], the policy's negotiation outputs for every
     * env; only the rows of the resetting envs are read) the agreement stage of ssd_negotiate also runs for them
     * (neg_decision_dev uint8 [E], nullable, receives their decisions).  Needs done_dev.  One masked launch behind the step. */
    int32_t auto_reset;
    const double* neg_proposals_dev;
    const double* neg_accept_dev;
    uint8_t* neg_decision_dev;
} ssd_step_io;

/* --- lifetime ------------------------------------------------------------------------------- */
/* Replaces env_creator(name, config) (utils/env_creator_functions.py:12-34) for E envs at once:
 * parses the ascii map like MapEnv.__init__/CleanupEnv.__init__ (map_env.py:92-130,
 * cleanup_new.py:106-125), uploads point lists / palette / spawn-probability thresholds. */
int ssd_create(const ssd_config* cfg, ssd_handle** out);
void ssd_destroy(ssd_handle* h);
const char* ssd_last_error(const ssd_handle* h);

/* --- the hot path ----------------------------------------------------------------------------- */
/* MapEnv.reset + CleanupEnv/HarvestEnv.custom_reset + SeparateContractSubgameStage.reset
 * (map_env.py:306-342, cleanup_new.py:171-209, harvest_new.py:143-176, two_stage_train.py:159-187)
 * for the envs with mask_dev[i] != 0 (all envs when mask_dev is NULL).  Writes the reset
 * observation of those envs (agents are NOT drawn in it, as in the reference). */
int ssd_reset(ssd_handle* h, const uint8_t* mask_dev, uint8_t* obs_dev, int64_t obs_env_stride, void* stream);

/* MapEnv.step (map_env.py:216-304) + env step tail (cleanup_new.py:211-267 / harvest_new.py:181-239)
 * + SeparateContractEnv.step reward redistribution (two_stage_train.py:62-121), one launch for E envs. */
int ssd_step(ssd_handle* h, const ssd_step_io* io, void* stream);
/* Error surface of the step.  ssd_step returns SSD_OK as soon as the launch is enqueued; conditions that only the
 * kernel can see are recorded in a STICKY per-env flag word that lasts until the env is reset and is reported as
 * `err_flags` by ssd_get_metrics (out[5]) / ssd_set_episode_stats (max over envs):
 *   2  the reference would have raised KeyError in update_moves (map_env.py:640, agent_by_pos lookup)
 *   8  an action id outside the env's action space (the agent stays in place; the reference raises KeyError in
 *      Agent.action_map, Agent.py:161-162)
 *   16 ssd_set_state was given a map character that is not one of ' @AHRS'
 * The reference checks none of these eagerly either; callers that want a hard failure read err_flags at episode end
 * (contracts_b200/environments/gridworld.py does and raises).  There is no `rng_mode=injected` / ssd_set_injected_draws:
 * the only random stream is the counter-based Philox one (key = seed, global env id), which the parity harness injects
 * into the unmodified reference instead (oracle/ref_harness.py). */
/* The same step for a caller that holds HOST buffers (a CPU rollout worker, RLlib's sampler): copies
 * actions_host (uint8 [E][n], pinned for asynchronous copies) into io->actions_dev, steps, copies the rewards
 * (double [E][n], after transfers) and dones (uint8 [E], nullable) back, and returns when they are valid.  The device ->
 * host copy starts as soon as the rewards exist — for cleanup_new that is before the observe kernel, so it overlaps
 * it; the observations stay in the device batch tensor io->obs_dev. */
int ssd_step_host(ssd_handle* h, const ssd_step_io* io, const void* actions_host, double* rew_host, uint8_t* done_host,
                  void* stream);

/* Pipelined form of ssd_step_host (what RLlib's BaseEnv.send_actions() / poll() pair is to a vector env,
 * utils/ray_config_utils.py:140 -> rollout worker): submit step k + 1 before waiting for step k.  Two slots.
 *   ssd_step_host_async  uploads actions_host (uint8 [E][n], pinned) on a library-owned copy-in stream into the slot's
 *                        device action buffer (io->actions_dev is ignored), runs the step on `stream`, and — as soon as
 *                        the rewards exist, i.e. overlapping the observe kernel — copies ONE result block to
 *                        result_host (pinned, ssd_host_result_layout().total_bytes) on a copy-out stream.  Returns
 *                        without any host synchronisation; *ticket_out identifies the step (ticket & 1 = slot).
 *   ssd_step_host_wait   blocks until that step's result block is valid in host memory (and the slot can be reused).
 * At most two steps may be in flight; a slot must be waited for before it is submitted again (SSD_EINVAL otherwise).
 *
 * The result block is a LOSSLESS compact form of (rewards after transfers, dones).  Rewards on this path are small
 * integers except in envs where a contract transfer was paid this step, so the block holds
 *   count    uint32         number of records below
 *   done     uint8 [E]      dones['__all__']
 *   rew_i8   int8 [E][n]    the reward, valid for every env that is NOT in the record list
 *   records  [count] x { int32 env; int32 0; double rew[n] }   envs with a reward that is not an integer in
 *                           [-127, 127] (bit pattern compared, so -0.0 also lands here): all n exact float64 rewards;
 *                           order unspecified (warp-aggregated atomics)
 * at the byte offsets ssd_host_result_layout() reports.  Only the leading `count` records travel: the copy size is
 * predicted from the previous steps and ssd_step_host_wait fetches the remainder in the (rare) case of a miss.
 * ssd_host_result_expand() is the host-side convenience that rebuilds the dense float64 [E][n] matrix. */
typedef struct ssd_host_layout {
    int64_t total_bytes;      /* size of a result block (worst case: every env in the record list) */
    int64_t count_offset;     /* uint32 */
    int64_t done_offset;      /* uint8 [E] */
    int64_t rew_i8_offset;    /* int8 [E][n] */
    int64_t records_offset;   /* records, record_bytes each */
    int32_t record_bytes;     /* 8 + 8 n */
    int32_t record_capacity;  /* E */
} ssd_host_layout;
int ssd_host_result_layout(const ssd_handle* h, ssd_host_layout* out);
int ssd_step_host_async(ssd_handle* h, const ssd_step_io* io, const void* actions_host, void* result_host,
                        int64_t* ticket_out, void* stream);
int ssd_step_host_wait(ssd_handle* h, int64_t ticket);
/* The same pipelined form for the feature envs and selfdrive (declared with their io structs below): one kernel per step,
 * io->actions_dev ignored, io->rew_dev may be NULL.  ssd_host_result_layout / ssd_step_host_wait / ssd_host_result_expand
 * serve every env kind; the block's `done` field is uint8 [E] for the gridworlds and the feature envs and uint8 [E][n+1]
 * (per-car dones, then '__all__') for selfdrive, whose actions_host is float32 [E][n]. */
/* host code only: rew_out double [E][n] from a valid result block */
int ssd_host_result_expand(const ssd_handle* h, const void* result_host, double* rew_out);

/* --- contract parameters / negotiation ----------------------------------------------------------- */
/* theta_dev: double [E].  What SeparateContractNegotiateStage does with a0's proposal
 * (two_stage_train.py:258-262) / a caller-chosen contract. */
int ssd_set_contract_params(ssd_handle* h, const double* theta_dev, void* stream);
/* Agreement stage of SeparateContractNegotiateStage.step (two_stage_train.py:266-281):
 * proposals_dev double [E] (a0's action[:-1]), accept_dev double [E][n] (each agent's action[-1]),
 * decision_dev uint8 [E] out.  Chooses 2 of agents 1..n-1 when n > 3, multiplies their accept
 * values, draws once and keeps the proposal as theta or zeroes it.  Every env kind. */
int ssd_negotiate(ssd_handle* h, const uint8_t* mask_dev, const double* proposals_dev, const double* accept_dev,
                  uint8_t* decision_dev, void* stream);
/* mask_dev (nullable, uint8 [E]): negotiate only in the envs with mask_dev[i] != 0 — the envs that were just reset when
 * every env of the batch is at its own point of its episode (a vectorised sampler in steady state). */

/* --- JointEnv output layouts (environments/two_stage_train.py:476-617) ----------------------------- */
/* `global_obs`: MapEnv.global_view() (map_env.py:394-395; base_env.get_global_obs(), cleanup_new.py:299-300,
 * harvest_new.py:178-179, before the / 255): the whole colour map with agents, uint8 [E][H][W][3]. */
int ssd_global_view(ssd_handle* h, uint8_t* out_dev, void* stream);
/* `concatenated_obs`: np.concatenate of the agents' windows along the channel axis (two_stage_train.py:527-533,
 * 604-609): obs_dev uint8 [E][n][15][15][3] (env stride obs_env_stride, 0 = dense) -> out_dev uint8 [E][15][15][3n]. */
int ssd_concat_obs(ssd_handle* h, const uint8_t* obs_dev, int64_t obs_env_stride, uint8_t* out_dev, void* stream);

/* --- render path (environments/map_env.py:389-392,460-475) --------------------------------------------- */
/* MapEnv.beam_pos (map_env.py:231,812) is only ever shown by render(): recording it is optional.  When enabled, every
 * step also stores the cells its beams crossed (library-owned overlay, cleared at each step / reset). */
int ssd_record_beams(ssd_handle* h, int32_t enable);
/* MapEnv.full_map_to_colors() = render(mode='rgb_array') (map_env.py:389-392): the map with the agents and, when recorded,
 * the beams of the last step on top (get_map_with_agents, map_env.py:354-375).  out_dev uint8 [E][H][W][3]. */
int ssd_render(ssd_handle* h, uint8_t* out_dev, void* stream);

/* --- policy-side consumer (environments/Networks/vision_net.py:150-181) --------------------------------- */
#define SSD_POLICY_F32 0
#define SSD_POLICY_F16 1
#define SSD_POLICY_BF16 2
/* The observation as VisionNetwork.forward feeds its first convolution, in one pass over the uint8 batch tensor:
 * image_dev T [E*n][3][15][15] = (`curr_obs / 255` as float32).permute(0, 3, 1, 2) (:159,167) cast to T;
 * contract_dev (nullable) T [E*n][10] = (theta, 0) repeated 5 times (:160-161).  T by `dtype`. */
int ssd_policy_inputs(ssd_handle* h, const uint8_t* obs_dev, int64_t obs_env_stride, int32_t dtype, void* image_dev,
                      void* contract_dev, void* stream);

/* --- NegotiationSolver (environments/two_stage_train.py:619-776); every env kind ------------------- */
#define SSD_SOLVER_RULE_MAX 0        /* decision_rule 'max'      (:751-757) */
#define SSD_SOLVER_RULE_MAJORITY 1   /* decision_rule 'majority' (:758-773) */
/* negotiate() :705-746, the candidate contracts: params_dev double [E][1 + num_samples]; [e][0] = contract_space.low
 * (the null contract), [e][1 + i] = contract_param_space.sample() = float32(uniform(low, high)) from the env's
 * Philox stream.  The caller evaluates its value function V(s, c) per agent for every candidate (compute_vals
 * :693-703 — an RLlib policy forward in the reference). */
int ssd_solver_sample(ssd_handle* h, int32_t num_samples, double* params_dev, void* stream);
/* compute_best_param() :748-776 on vals_dev double [E][1 + num_samples][n]; the chosen parameter becomes the env's
 * contract parameter (reset :675-676) and is also written to best_param_dev double [E] / best_index_dev int32 [E]
 * (either may be NULL). */
int ssd_solver_choose(ssd_handle* h, int32_t num_samples, int32_t rule, const double* params_dev, const double* vals_dev,
                      double* best_param_dev, int32_t* best_index_dev, void* stream);

/* --- state access (parity tests, checkpointing) ---------------------------------------------------- */
/* map: uint8 chars [E][H][W] (MapEnv.world_map); pos: int32 [E][n][2] (row, col); ori: int32 [E][n]
 * (Agent.int_orientation); t: int32 [E] (MapEnv.timesteps); theta: double [E].  NULL pointers are skipped. */
int ssd_get_state(ssd_handle* h, uint8_t* map_dev, int32_t* pos_dev, int32_t* ori_dev, int32_t* t_dev,
                  double* theta_dev, void* stream);
int ssd_set_state(ssd_handle* h, const uint8_t* map_dev, const int32_t* pos_dev, const int32_t* ori_dev,
                  const int32_t* t_dev, const double* theta_dev, void* stream);
/* Episode accumulators behind env.metrics (cleanup_new.py:186-189,213-232,264-266; two_stage_train.py:92-99).
 * out_dev: double [E][SSD_METRIC_STRIDE] =
 *   [apples_eaten, low_density_eaten, raw_env_rewards, transfers, dirt_cleaned, err_flags, 0, 0,
 *    agent_a[8], agent_b[8], sum_raw[8], tsum_raw[8], sum_transferred[8], tsum_transferred[8]] */
int ssd_get_metrics(ssd_handle* h, double* out_dev, void* stream);
/* Episode statistics without a pass over all envs (the MetricsCallback hand-off, utils/logger_utils.py:126-151):
 * while stats_dev (double [SSD_STATS_LEN], caller-owned, caller-zeroed) is set, every ssd_reset adds the accumulators
 * of each FINISHED episode it replaces (envs that had been reset before) to it with device atomics:
 *   [apples_eaten, raw_env_rewards, transfers, dirt_cleaned, sum of transferred rewards, sum of raw rewards,
 *    episodes, max err_flags].  NULL switches it off.  Gridworld kinds. */
#define SSD_STATS_LEN 8
int ssd_set_episode_stats(ssd_handle* h, double* stats_dev);

/* --- SelfAcceleratingCarEnv (env_kind SSD_ENV_SELFDRIVE) -------------------------------------------------
 * environments/self_driving_car_accelerate.py reset :49-79 / step :151-250 (collision_on=False), optionally with
 * SelfdriveContractDistprop (contract/contract_list.py:66-102) + SeparateContractSubgameStage
 * (two_stage_train.py:62-121,159-187) fused.  Agents that are done stop acting (RLlib semantics).  All values are
 * float64; D = 2 (n + 1) + 3 observation elements per car (ssd_feature_dim returns D). */
typedef struct ssd_selfdrive_io {
    const float* actions_dev;   /* [E][n] accelerations */
    double* obs_dev;            /* [E][n][D]  (rows of cars that did not act this step are still written) */
    double* rew_dev;            /* [E][n] after transfers; 0 for cars that did not act */
    double* base_rew_dev;       /* [E][n] nullable */
    double* transfers_dev;      /* [E][n] nullable: value of each car's transfer tuple */
    double* info_dev;           /* [E][n][4] nullable: just_passed, acted, ambulance_rank, ambulance_dist_to_front
                                   (last two in the row of the first acting car, like infos[key_lst[0]]) */
    uint8_t* done_dev;          /* [E][n+1] nullable: per-car dones, then '__all__' */
    int32_t auto_reset;         /* next-step auto-reset: an env whose episode ended ('__all__') in the previous step is reset
                                   by this step instead of being stepped — reset observation, zero rewards, dones cleared,
                                   its actions ignored (what ssd_selfdrive_reset with mask = done[:, n] would do, no launch) */
} ssd_selfdrive_io;
int ssd_selfdrive_reset(ssd_handle* h, const uint8_t* mask_dev, double* obs_dev, void* stream);
int ssd_selfdrive_step(ssd_handle* h, const ssd_selfdrive_io* io, void* stream);
/* host actions in (float32 [E][n], pinned), compact result block out: see ssd_step_host_async */
int ssd_selfdrive_step_host_async(ssd_handle* h, const ssd_selfdrive_io* io, const void* actions_host, void* result_host,
                                  int64_t* ticket_out, void* stream);
/* pos/vel double [E][n], theta / transfers metric double [E], t int32 [E]; NULL pointers are skipped */
int ssd_selfdrive_get_state(ssd_handle* h, double* pos_dev, double* vel_dev, double* theta_dev, double* transfers_dev,
                            int32_t* t_dev, void* stream);
/* uniform float32 accelerations in [lo, hi) for benchmark rollouts, [E][n] */
int ssd_selfdrive_random_actions(ssd_handle* h, uint32_t step_index, float lo, float hi, float* actions_dev, void* stream);

/* --- CleanupFeatures / HarvestFeatures (env_kind SSD_ENV_CLEANUP_FEATURES / SSD_ENV_HARVEST_FEATURES) ----------
 * environments/cleanup_features.py step :156-254 / reset :256-284, environments/harvest_features.py step :173-287 /
 * reset :289-336, optionally with CleanupContract / HarvestFeaturemodLocalContract + SeparateContractSubgameStage
 * fused (two_stage_train.py:62-121,159-187).  ssd_create takes the same ascii map as the gridworld kinds.
 * F = ssd_feature_dim(h) = 12 + n (cleanup) or 10 + 2 n (harvest) float64 features per agent. */
typedef struct ssd_feat_io {
    const uint8_t* actions_dev; /* [E][n] action ids (cleanup 0-8: moves, stay, turns, clean 7, fire 8; harvest 0-7) */
    double* obs_dev;            /* [E][n][F] */
    double* rew_dev;            /* [E][n] after transfers */
    double* base_rew_dev;       /* [E][n] nullable */
    double* transfers_dev;      /* [E][n] nullable */
    uint8_t* info_dev;          /* [E][n][4] nullable: cleanup (cleaned_squares,0,0,0); harvest (eaten_apples, eaten_close_apples,0,0) */
    uint8_t* done_dev;          /* [E] nullable: dones['__all__'] (timesteps == horizon) */
    int32_t auto_reset;         /* next-step auto-reset, as in ssd_selfdrive_io */
} ssd_feat_io;
int ssd_feat_reset(ssd_handle* h, const uint8_t* mask_dev, double* obs_dev, void* stream);
int ssd_feat_step(ssd_handle* h, const ssd_feat_io* io, void* stream);
/* host actions in (uint8 [E][n], pinned), compact result block out: see ssd_step_host_async */
int ssd_feat_step_host_async(ssd_handle* h, const ssd_feat_io* io, const void* actions_host, void* result_host,
                             int64_t* ticket_out, void* stream);
/* pos int32 [E][n][2], ori int32 [E][n], cells uint8 [E][H][W] (1 = apple, 2 = waste in the current lists),
 * theta double [E], t int32 [E]; NULL pointers are skipped */
int ssd_feat_get_state(ssd_handle* h, int32_t* pos_dev, int32_t* ori_dev, uint8_t* cells_dev, double* theta_dev,
                       int32_t* t_dev, void* stream);
/* double [E][40]: dirt_cleaned, raw_env_rewards, transfers, total_apples_eaten, low_density_apples_eaten,
 * err_flags (always 0 since the lists are kept in birth order; was: 16-bit birth stamps wrapped), 0, 0,
 * per agent: sum r [8], sum t*r [8], sum transferred r [8], sum t * transferred r [8] */
int ssd_feat_get_metrics(ssd_handle* h, double* out_dev, void* stream);

/* --- utilities ------------------------------------------------------------------------------------------ */
/* uniform random action ids in [0, num_actions) for benchmark rollouts, uint8 [E][n];
 * drawn from Philox site 13 at counter `step_index` (no reference equivalent: RLlib's policy).
 * step_index == SSD_STEP_AUTO takes (and then bumps) a per-handle device counter instead, so the call can be
 * captured in a CUDA graph and still draw fresh actions at every replay (the last CTA to finish bumps the counter). */
#define SSD_STEP_AUTO 0xFFFFFFFFu
int ssd_random_actions(ssd_handle* h, uint32_t step_index, int32_t num_actions, uint8_t* actions_dev, void* stream);   /* gridworld and feature kinds */
/* host-side Philox4x32-10 (so tests can pin the generator: KATs in tests/test_philox.py) */
void ssd_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
int ssd_abi_version(void);
int ssd_feature_dim(const ssd_handle* h);          /* F of feature_obs_dev */
int64_t ssd_state_bytes_per_env(const ssd_handle* h);
/* Measurement only (bench.py roofline): with timing on, ssd_step brackets its kernels with CUDA events on the
 * launch stream (not capturable in a CUDA graph).  ssd_get_step_times waits for the last step and returns
 * out_ms[0] = logic kernel, out_ms[1] = observe (+ harvest reward) kernel, in milliseconds. */
int ssd_enable_timing(ssd_handle* h, int32_t on);
int ssd_get_step_times(ssd_handle* h, double* out_ms);
int64_t ssd_kernel_launches(const ssd_handle* h);  /* kernels launched through this handle so far */

#ifdef __cplusplus
}
#endif
#endif /* SSD_B200_H */
