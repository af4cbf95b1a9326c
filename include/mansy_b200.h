/*
 * mansy_b200.h -- C ABI of the B200-native tile-based streaming simulator.
 *
 * This is the drop-in boundary for the reference's data-parallel hot path
 * (bitrate_selection/simulators + bitrate_selection/envs of
 * duowuyms/MANSY_ImmersiveVideoStreaming).  The reference is pure Python and has no FFI; each
 * entry point below cites the reference interface it replaces (file:line relative to the
 * reference root).  Signatures use plain pointers and sizes only -- no torch types.  A Python
 * caller binds it with ctypes (mansy_immersivevideostreaming_b200/_capi.py; INTEGRATION.md shows
 * the stub a maintainer of the reference would add).
 *
 * Conventions
 *   - every function returns 0 on success or a negative MANSY_E_* code; the message is
 *     available from mansy_last_error() (thread-local).
 *   - "dev" pointers are device pointers on the handle's GPU, "host" pointers are host memory
 *     (pinned for asynchronous copies).  `stream` is a cudaStream_t passed as void*; all device
 *     work is ordered on it and nothing synchronises unless documented.
 *   - the caller owns observation / reward / done / action buffers; the library owns its copies
 *     of the read-only tables and the per-environment state.  No allocation happens after
 *     mansy_create / mansy_policy_create.
 *   - a handle is not thread-safe; use one host thread per handle.
 */
#ifndef MANSY_B200_H_
#define MANSY_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MANSY_ABI_VERSION 1

/* error codes */
#define MANSY_OK 0
#define MANSY_E_INVALID (-1) /* bad argument / unsupported configuration */
#define MANSY_E_CUDA (-2)    /* CUDA runtime error (message holds cudaGetErrorString) */
#define MANSY_E_NOMEM (-3)
#define MANSY_E_STATE (-4)   /* call not valid in the handle's current state */

/* observation layouts (float32 words per environment row, see config.py) */
#define MANSY_OBS_NONE 0
#define MANSY_OBS_MANSY 1  /* 13 arrays, 777 floats, row stride 784 (envs/mansy_env.py:136-150) */
#define MANSY_OBS_SIMPLE 2 /* 5 arrays, 395 floats, row stride 400 (envs/simple_rl_env.py:103-109) */
#define MANSY_OBS_MANSY_STRIDE 784
#define MANSY_OBS_SIMPLE_STRIDE 400

/* reward modes: envs/mansy_env.py:168-177, envs/simple_rl_env.py:124-127 */
#define MANSY_REWARD_QOE 0
#define MANSY_REWARD_QOE_NORM 1 /* qoe / (w1 + w2 + w3) */

/* doubles per environment in the optional `aux` output of mansy_step */
#define MANSY_AUX_DOUBLES 16
enum {
  MANSY_AUX_CHUNK_SIZE = 0,   /* simulators/simulator.py:100  (exact integer) */
  MANSY_AUX_DOWNLOAD_TIME = 1,/* simulators/network.py:34 */
  MANSY_AUX_REBUFFER = 2,     /* simulators/buffer.py:8-15 */
  MANSY_AUX_BUFFER = 3,       /* buffer level after the push */
  MANSY_AUX_CUR_IDX = 4,      /* network.py cur_idx after the download (exact integer) */
  MANSY_AUX_CUR_TIME = 5,     /* network.py cur_time after the download */
  MANSY_AUX_QOE = 6,          /* utils/qoe.py:33 */
  MANSY_AUX_QOE1 = 7,
  MANSY_AUX_QOE2 = 8,
  MANSY_AUX_QOE3 = 9,
  MANSY_AUX_NEXT_CHUNK = 10,  /* simulator.next_chunk after the step (exact integer) */
  MANSY_AUX_EP_STEP = 11,     /* steps taken in the episode, this one included */
  MANSY_AUX_SAMPLE_ID = 12,   /* sample the stepped episode belongs to */
  MANSY_AUX_REWARD = 13,      /* reward before rounding to float32 */
  MANSY_AUX_GT_MASK_LO = 14,  /* ground-truth mask used by the QoE, low / high 32 bits */
  MANSY_AUX_GT_MASK_HI = 15
};

/* doubles per environment in mansy_episode_stats */
#define MANSY_STATS_DOUBLES 16
enum {
  /* last finished episode (the row envs/mansy_env.py:271-290 `_log` appends) */
  MANSY_STAT_LAST_SUM_QOE = 0,
  MANSY_STAT_LAST_SUM_QOE1 = 1,
  MANSY_STAT_LAST_SUM_QOE2 = 2,
  MANSY_STAT_LAST_SUM_QOE3 = 3,
  MANSY_STAT_LAST_STEPS = 4,
  MANSY_STAT_LAST_SAMPLE = 5,
  /* totals over all finished episodes since create / mansy_stats_clear */
  MANSY_STAT_TOT_SUM_QOE = 6,
  MANSY_STAT_TOT_SUM_QOE1 = 7,
  MANSY_STAT_TOT_SUM_QOE2 = 8,
  MANSY_STAT_TOT_SUM_QOE3 = 9,
  MANSY_STAT_TOT_STEPS = 10,
  MANSY_STAT_TOT_EPISODES = 11,
  MANSY_STAT_TOT_RETURN = 12, /* sum of rewards of finished episodes */
  MANSY_STAT_LAST_RETURN = 13
};

typedef struct mansy_sim *mansy_handle_t;
typedef struct mansy_policy *mansy_policy_t;

/*
 * Read-only inputs, HOST pointers; mansy_create copies them to the device.
 * Replaces the three files Simulator.__init__ re-reads at every reset
 * (simulators/simulator.py:30-38): manifest JSON, viewport pickle, bandwidth pickle.
 */
typedef struct {
  const int32_t *size;       /* [n_videos][n_chunks][5][64] tile bytes      (manifest "size") */
  const float *quality;      /* [n_videos][n_chunks][5][64] tile quality    (manifest "quality") */
  const int32_t *video_time; /* [n_videos]                                  (manifest "Video_Time") */
  const uint64_t *vp_gt;     /* [n_videos*n_users][n_vp_chunks] bit t = tile t (row*8+col) */
  const uint64_t *vp_pred;   /* same shape: predicted viewport */
  const double *vp_acc;      /* same shape: IoU accuracy */
  const int32_t *vp_start;   /* [n_videos*n_users] first chunk id  (simulators/hmdtrace.py:10) */
  const int32_t *vp_end;     /* [n_videos*n_users] last chunk id   (simulators/hmdtrace.py:11) */
  const double *trace;       /* [n_traces][trace_stride] bytes/s per 1-s segment (network.py:9) */
  const int32_t *trace_len;  /* [n_traces] */
  const float *qoe_w;        /* [n_qoe][3] */
  const int32_t *samples;    /* [n_samples][4] (video, user, trace, qoe) indices (utils/common.py:60-98) */
  int32_t n_videos, n_chunks, n_users, n_vp_chunks, n_traces, trace_stride, n_qoe, n_samples;
} mansy_tables_t;

/* Constants of config.yml:68-75,153-157 plus the vector-env geometry. */
typedef struct {
  int32_t n_envs;     /* environments owned by this handle (the local shard) */
  int32_t env_offset; /* global index of local env 0 (multi-GPU sharding; 0 on one GPU) */
  int32_t worker_num; /* global sample stride (envs/mansy_env.py:56,100-101) */
  int32_t seed;       /* env k starts at sample (seed + env_offset + k) % worker_num (mansy_env.py:253-256) */
  int32_t obs_mode;   /* MANSY_OBS_* */
  int32_t reward_mode;/* MANSY_REWARD_* */
  int32_t video_rates[5];
  int32_t startup_download;
  int32_t chunk_length;
  int32_t max_size;
  int32_t max_throughput;
} mansy_cfg_t;

/* Outputs of one step, DEVICE pointers, row i belongs to env_ids[i] (or env i when env_ids is NULL). */
typedef struct {
  float *obs;            /* [n][obs_stride] or NULL (no observation materialised) */
  int64_t obs_stride;    /* floats between rows; >= the layout's stride and a multiple of 4 */
  float *reward;         /* [n] or NULL */
  uint8_t *done;         /* [n] or NULL */
  double *aux;           /* [n][MANSY_AUX_DOUBLES] or NULL (parity / debugging) */
  uint8_t *tile_versions;/* [n][64] chosen bitrate version per tile (utils/common.py:142-193) or NULL */
} mansy_out_t;

const char *mansy_last_error(void);
int mansy_abi_version(void);
/* kernels launched by this library in the calling process (all handles), for bench accounting */
int64_t mansy_kernel_launches(void);

/* Simulator(...) + MANSYEnv/SimpleRLEnv state for n_envs environments
 * (simulators/simulator.py:15-46, envs/mansy_env.py:19-97). */
int mansy_create(const mansy_tables_t *tables, const mansy_cfg_t *cfg, int device, mansy_handle_t *out);
int mansy_destroy(mansy_handle_t h);

/* env.seed(seed) of a vector env: env k gets worker_id (seed + env_offset + k) % worker_num
 * (envs/mansy_env.py:253-256). */
int mansy_seed(mansy_handle_t h, int32_t seed, void *stream);

/* env.reset() for the listed envs (env_ids_dev NULL = all): picks the next sample, rebuilds the
 * episode state and writes the initial observation (envs/mansy_env.py:99-152,
 * envs/simple_rl_env.py:76-111).  obs_dev may be NULL. */
int mansy_reset(mansy_handle_t h, const int32_t *env_ids_dev, int32_t n, float *obs_dev, int64_t obs_stride,
                void *stream);

/* env.step(action) (envs/mansy_env.py:154-248, envs/simple_rl_env.py:113-160) for the listed
 * envs.  actions_dev[i] belongs to env_ids_dev[i].  With auto_reset != 0 an env whose episode
 * ended is reset inside the same launch and its row holds the first observation of the next
 * episode (done is still 1); otherwise the row holds the terminal observation and the env
 * stays finished until mansy_reset. */
int mansy_step(mansy_handle_t h, const int32_t *actions_dev, const int32_t *env_ids_dev, int32_t n,
               int32_t auto_reset, const mansy_out_t *out, void *stream);

/* Same step with HOST buffers: copies actions host->device, steps every env, copies observation,
 * reward and done device->host and synchronises the stream.  Host buffers should be pinned. */
int mansy_step_host(mansy_handle_t h, const int32_t *actions_host, int32_t auto_reset, float *obs_host,
                    float *reward_host, uint8_t *done_host, void *stream);
int mansy_reset_host(mansy_handle_t h, float *obs_host, void *stream);

/* n_steps lockstep steps with in-kernel actions: action = hash(seed, global env, step0 + t) % 15
 * (the simulator-only throughput sweep of run_simple_rl.py-style baselines), auto-reset on.
 * Buffers are [n_steps][n_envs][...] when `per_step_outputs` != 0, otherwise [n_envs][...]
 * overwritten every step. */
int mansy_rollout_random(mansy_handle_t h, int32_t n_steps, uint64_t seed, int64_t step0, int32_t per_step_outputs,
                         const mansy_out_t *out, void *stream);

/* Per-env episode statistics (what `_log` writes per episode, envs/mansy_env.py:271-290, plus
 * running totals for the per-rollout all-gather): stats_dev is [n_envs][MANSY_STATS_DOUBLES]. */
int mansy_episode_stats(mansy_handle_t h, double *stats_dev, void *stream);
int mansy_stats_clear(mansy_handle_t h, void *stream);
/* The six running totals a rollout exchanges (MANSY_STAT_TOT_SUM_QOE .. MANSY_STAT_TOT_EPISODES) packed as
 * totals_dev[n_envs][6] by one small kernel: the send buffer of the per-rollout all-gather. */
int mansy_episode_totals(mansy_handle_t h, double *totals_dev, void *stream);

/*
 * Peer group: the per-rollout all-gather of episode totals as ONE kernel over NVLink peer memory, and a
 * device-side barrier (csrc/mansy_peer.cu).  One process per GPU; every rank creates a mailbox, exports its CUDA
 * IPC handle (MANSY_PEER_HANDLE_BYTES bytes), the caller exchanges the handles out of band (torch.distributed
 * all_gather in rollout.PeerGroup) and passes all `world` of them, rank-ordered, to mansy_peer_connect.
 * Replaces nothing in the reference (it has no multi-device path); it is the B200 form of
 * "a single allgather of episode returns and QoE statistics per rollout" (BASELINE.json north_star).
 *   slot_bytes: bytes each rank contributes per gather = n_envs_local * 6 * 8.
 *   mansy_peer_allgather_stats: *gathered_dev (device pointer on this GPU, valid until the gather after the next
 *   one) holds [world * n_envs_local][6] float64 in rank order once the stream has passed the call.
 *   Every rank must issue the same sequence of barrier / gather calls.  A wait that sees no peer for ~4 s gives
 *   up and raises the flag mansy_peer_timed_out reports (the data of that gather is then incomplete).
 */
#define MANSY_PEER_HANDLE_BYTES 64
typedef struct mansy_peer *mansy_peer_t;
int mansy_peer_create(int32_t world, int32_t rank, int64_t slot_bytes, int device, mansy_peer_t *out);
int mansy_peer_export(mansy_peer_t p, void *handle_out);
int mansy_peer_connect(mansy_peer_t p, const void *all_handles);
int mansy_peer_barrier(mansy_peer_t p, void *stream);
int mansy_peer_allgather_stats(mansy_peer_t p, mansy_handle_t h, void *stream, const double **gathered_dev);
int mansy_peer_timed_out(mansy_peer_t p, int32_t *flag_host);
int mansy_peer_destroy(mansy_peer_t p);

/* (viewport pair, chunk, action) outcome table: chunk bytes, viewport quality and intra-chunk variance -- everything
 * of a step that depends only on the read-only tables and the action (utils/common.py:101-193, simulators/simulator.py:94-101,
 * utils/qoe.py:23-28; the reference's ExpertEnv caches the same statistics, envs/expert_env.py:121-160) -- is tabulated
 * at mansy_create by the step's own gather code, and steps read it back (bit-identical results) unless tile_versions
 * are requested.  enable = 0 makes every step gather again (parity tests compare the two). */
int mansy_set_outcome_table(mansy_handle_t h, int32_t enable);

/* Field-of-view -> tile masks with per-chunk OR and IoU
 * (viewport_prediction/utils/common.py:37-58,83-127; viewport_prediction/predict.py:33-48).
 * gt_xy_dev / pred_xy_dev: float32 [n_chunks][points][2] normalised centres in [0,1];
 * outputs gt/pred uint64 [n_chunks] (bit t = tile t), acc float64 [n_chunks].  pred_xy_dev,
 * pred_mask_dev and acc_dev may be NULL together.  Grid fixed to 8x8 tiles. */
int mansy_viewport_tiles(const float *gt_xy_dev, const float *pred_xy_dev, int64_t n_chunks, int32_t points,
                         int32_t video_width, int32_t video_height, int32_t fov_width, int32_t fov_height,
                         uint64_t *gt_mask_dev, uint64_t *pred_mask_dev, double *acc_dev, void *stream);

/* action -> per-tile bitrate versions for standalone masks (utils/common.py:101-119,142-193);
 * masks_dev uint64[n], actions_dev int32[n], versions_dev uint8[n][64]. */
int mansy_allocate_tile_versions(const uint64_t *masks_dev, const int32_t *actions_dev, int64_t n,
                                 const int32_t video_rates[5], uint8_t *versions_dev, void *stream);

/*
 * Policy / value forward (bitrate_selection/models/mansy.py:26-51,63-66,77-80 and
 * models/simple_rl.py:21-35,46-49,60-63).  Weights are HOST float32 arrays in the reference's
 * state-dict layouts; mansy_policy_create repacks and uploads them.
 */
#define MANSY_NET_IDENTIFIER 3 /* mansy_policy_weights_t.kind: the QoE identifier (models/mansy.py:83-143): the MANSY FeatureNet
                                 layout with fc2 over the 15 action_one_hot floats (row[760..774]) as the 10th / residual
                                 branch, ONE head Linear(128 -> 3) + sigmoid.  branch_w[9] is [128][15], actor_fc_* is
                                 identifier.fc, actor_out_w / actor_out_b are [3][128] / [3]; the critic_* pointers are
                                 ignored (may be NULL).  Forward calls return the 3 sigmoid outputs in logits[:, 0..2]
                                 (columns 3..15 and value are 0). */
typedef struct {
  int32_t kind; /* MANSY_OBS_MANSY, MANSY_OBS_SIMPLE or MANSY_NET_IDENTIFIER */
  /* MANSY: branch weights [128][K_b] for the 10 branches in FeatureNet order, biases [128] */
  const float *branch_w[10];
  const float *branch_b[10];
  const float *actor_fc_w;  /* [128][feature_dim] */
  const float *actor_fc_b;  /* [128] */
  const float *actor_out_w; /* [15][128] */
  const float *actor_out_b; /* [15] */
  const float *critic_fc_w; /* [128][feature_dim] */
  const float *critic_fc_b;
  const float *critic_out_w;/* [1][128] */
  const float *critic_out_b;/* [1] */
} mansy_policy_weights_t;

int mansy_policy_create(const mansy_policy_weights_t *w, int device, mansy_policy_t *out);
int mansy_policy_destroy(mansy_policy_t p);
/* logits_dev [n][16] (15 used; MANSY: raw logits, SimpleRL: softmax probabilities as the reference's
 * Actor returns), value_dev [n]. */
int mansy_policy_forward(mansy_policy_t p, const float *obs_dev, int64_t obs_stride, int32_t n, float *logits_dev,
                         float *value_dev, void *stream);
/* Categorical(logits).sample() with a counter-based generator keyed by (seed, env, step)
 * (run_mansy.py:228-229); also returns log-probabilities. */
int mansy_policy_sample(const float *logits_dev, int32_t n, int32_t is_probs, uint64_t seed, int64_t step,
                        int32_t env_offset, int32_t *actions_dev, float *logp_dev, void *stream);

/*
 * Identifier reward (bitrate_selection/utils/mansy_utils.py:42-49, models/mansy_ppo.py:40-49):
 *   ident[i] = 1 - mean_j (pred[i][j] - qoe_weight[i][j])^2   (float32 like F.mse_loss; qoe_weight = obs row floats 776..778)
 *   mixed[i] = (1 - lamb) * qoe_reward[i] + lamb * ident[i]   (float64 like the numpy expression in PPOPolicy.update)
 * pred_dev [n][16] is the identifier forward's output; ident_dev (float32 [n]) and mixed_dev (float64 [n]) may be NULL.
 */
int mansy_identifier_reward(const float *pred_dev, const float *obs_dev, int64_t obs_stride, const float *qoe_reward_dev,
                            double lamb, int32_t n, float *ident_dev, double *mixed_dev, void *stream);

/*
 * Generalised advantage estimation over a rollout ring, one backward scan per environment: what tianshou 0.4.8's
 * BasePolicy.compute_episodic_return / _gae_return does on the host (pinned README.md:21, source not vendored --
 * parity unpinned; oracle/gae_oracle.py restates the published algorithm):
 *   delta_t = rew_t + gamma * v_next_t * (1 - done_t) - v_t;   gae_t = delta_t + gamma * lam * (1 - done_t) * gae_{t+1}
 *   adv_t = gae_t,  ret_t = gae_t + v_t          (float64 accumulation, outputs float32)
 * reward / value / done are [T][n] (row t = rollout step t), last_value [n] is V(s_T) of the observation after the
 * last step; adv_dev / ret_dev [T][n].
 */
int mansy_gae(const float *reward_dev, const float *value_dev, const uint8_t *done_dev, const float *last_value_dev,
              int32_t n_steps, int32_t n, double gamma, double lam, float *adv_dev, float *ret_dev, void *stream);

/* Tensor-core (tcgen05, TF32 inputs / fp32 accumulate -- the precision class the reference runs its nets
 * in, run_mansy.py:253) version of mansy_policy_forward + mansy_policy_sample in ONE launch: 128 environments
 * per CTA, observation rows are TMA-loaded as the MMA operand.  Any of logits / value / actions / logp may be
 * NULL.  feat_dbg_dev ([n][n_branches*128], branch PROCESSING order: MANSY conv1d2, conv1d3, conv1d4, conv1d1,
 * conv1d5..8, fc1, fc2; SimpleRL conv1d_2, fc3, conv1d_1, fc1, fc2) and hid_dbg_dev ([n][256], actor | critic
 * hidden activations after the residual) are test hooks, normally NULL.  obs_stride must be a multiple of 4. */
int mansy_policy_forward_tc(mansy_policy_t p, const float *obs_dev, int64_t obs_stride, int32_t n, float *logits_dev,
                            float *value_dev, int32_t *actions_dev, float *logp_dev, uint64_t seed, int64_t step,
                            int32_t env_offset, float *feat_dbg_dev, float *hid_dbg_dev, void *stream);

/* mansy_policy_forward_tc for the CURRENT observations of simulator `h` (row i = environment i, all n_envs rows): the two
 * 320-input FeatureNet branches (conv1d2 / conv1d3 over the next chunk's size and quality tables, models/mansy.py:15-16,
 * 41-42; conv1d_2 of models/simple_rl.py) see rows of the simulator's read-only tables, so their share of the
 * actor.fc | critic.fc pre-activation is looked up in a (video, chunk) table computed once per (policy, simulator) in exact
 * fp32 instead of being recomputed per environment step.  The rollout entry points use the same path. */
int mansy_policy_forward_tc_sim(mansy_policy_t p, mansy_handle_t h, const float *obs_dev, int64_t obs_stride, float *logits_dev,
                                float *value_dev, int32_t *actions_dev, float *logp_dev, uint64_t seed, int64_t step, void *stream);

/* Same launch with a profiling hook: timeline_dev (int64[512], may be NULL) receives SM-clock stamps of CTA 0:
 * [0..127] TMA issue per job, [128..255] operand arrival per job, [256..383] MMAs issued per job,
 * [384 + 2i, 385 + 2i] epilogue begin/end of branch i (then the head epilogue), [511] kernel start. */
int mansy_policy_forward_tc_timeline(mansy_policy_t p, const float *obs_dev, int64_t obs_stride, int32_t n, float *logits_dev,
                                     float *value_dev, int32_t *actions_dev, float *logp_dev, uint64_t seed, int64_t step,
                                     int32_t env_offset, float *feat_dbg_dev, float *hid_dbg_dev, int64_t *timeline_dev,
                                     void *stream);
/* How a 128-environment tile is mapped to SMs by the tensor-core forward: 1 = one CTA per tile (persistent
 * over tiles), 4 = a 4-CTA thread-block cluster per tile, each CTA running layer 1 for a subset of the
 * FeatureNet branches and the matching K-slice of actor.fc | critic.fc, partial sums reduce-scattered over
 * distributed shared memory (for batches whose tiles cannot fill the SMs), 0 = chosen from the batch size
 * (default).  In split mode hid_dbg_dev receives the hidden activations WITHOUT the `+ qoe_features`
 * residual (models/mansy.py:65,79), which that kernel applies through the output heads; in the timeline
 * [480, 482..486] are the phase stamps (partial done, partials stored, cluster sync 1, heads issued, cluster sync 2,
 * rows written) of the CTA selected with the environment variable MANSY_TC_TIMELINE_CTA. */
int mansy_policy_tc_set_split(mansy_policy_t p, int32_t split);
/* debugging builds (-DMANSY_MBAR_WATCHDOG) only: host-mapped int32[grid * 16] that every warp of the fused rollout kernel stamps with (step << 16 | phase code) */
int mansy_debug_progress(int32_t *progress_dev);
/* Profiling hook of the fused rollout kernel (mansy_rollout_policy's one-launch path): CTA `cta` stamps the SM clock
 * of its SECOND rollout step into timeline_dev (int64[512], layout as above plus [489] step begin, [487] simulator
 * phase done, [488] cluster sync 3, [490..492] state loaded / step_env done / observation written); NULL turns it off. */
int mansy_debug_fused_timeline(int64_t *timeline_dev, int32_t cta);

/*
 * Rollout loop on the device: what tianshou's Collector.collect(n_step) does around policy(batch) and
 * env.step(act) (test loop bitrate_selection/run_mansy.py:161-175), without a host round trip
 * per step.  Step t (t = t0 .. t0 + n_steps - 1) reads observation slab t % slabs, writes actions / logp /
 * value / reward / done of slab t % slabs and the next observation into slab (t + 1) % slabs (auto-reset on).
 * Slab 0 .. must hold the current observations when the call is made (mansy_reset into slab t0 % slabs).
 * All pointers are DEVICE memory owned by the caller; two kernel launches per step.
 */
typedef struct {
  float *obs;         /* [slabs][n_envs][obs_stride] */
  int64_t obs_stride; /* floats between rows */
  int32_t slabs;
  int32_t *actions;   /* [slabs][n_envs] */
  float *logp;        /* [slabs][n_envs] */
  float *value;       /* [slabs][n_envs] */
  float *reward;      /* [slabs][n_envs] */
  uint8_t *done;      /* [slabs][n_envs] */
  float *logits;      /* [n_envs][16] scratch (overwritten every step) */
} mansy_rollout_t;
#define MANSY_ROLLOUT_FP32_POLICY 1 /* use the exact-fp32 CUDA-core policy kernels instead of tcgen05 */
#define MANSY_ROLLOUT_TIME_KERNELS 2 /* record CUDA events around every policy / step launch */
#define MANSY_ROLLOUT_NO_PDL 4       /* launch without programmatic dependent launch (kernels strictly one after another) */
#define MANSY_ROLLOUT_NO_ZERO_COPY 16 /* mansy_rollout_policy_host: hand actions and per-step scalars to the host with cudaMemcpyAsync even
                                         when the host buffers are device-mapped (default: kernels store them into the buffers directly) */
#define MANSY_ROLLOUT_TWO_KERNELS 8  /* never use the fused policy+step cluster kernel (one launch for all n_steps), which
                                        mansy_rollout_policy picks up to two 128-env tiles per resident 4-CTA cluster (<= 8 448 envs on B200) */
int mansy_rollout_policy(mansy_handle_t h, mansy_policy_t p, const mansy_rollout_t *buffers, int32_t n_steps, int64_t t0,
                         uint64_t seed, int32_t flags, void *stream);
/*
 * The same loop for callers whose rollout storage is HOST memory (tianshou's replay buffer is numpy): every
 * step (1) runs the policy on the device-resident observation, (2) copies the sampled actions to the host and
 * synchronises -- policy(batch) + to_numpy(act), run_mansy.py:169-172 --, (3) copies them back as the input of
 * env.step(act) (run_mansy.py:173), (4) steps, (5) copies the next observation, reward, done, logp and value
 * to host slab t % host_slabs on a second stream, overlapped with the following step (a device slab is not
 * rewritten before its copy has finished).  The call returns when every copy has landed.  Host buffers should
 * be pinned; host_slabs >= n_steps keeps every step's results.
 */
typedef struct {
  int32_t host_slabs;
  float *obs;       /* [host_slabs][n_envs][obs_stride]: observation AFTER the step */
  int32_t *actions; /* [host_slabs][n_envs] */
  float *logp, *value, *reward;
  uint8_t *done;
} mansy_rollout_host_t;
int mansy_rollout_policy_host(mansy_handle_t h, mansy_policy_t p, const mansy_rollout_t *buffers,
                              const mansy_rollout_host_t *host, int32_t n_steps, int64_t t0, uint64_t seed, int32_t flags,
                              void *stream);

/* Creates the CUDA events a timed rollout of n_steps needs ahead of time (event creation is slow; keep it out
 * of a timed region). */
int mansy_rollout_reserve_timing(mansy_handle_t h, int32_t n_steps);
/* After the stream has been synchronised: summed event-timed durations (ms) of the policy and step launches
 * of the last rollout that ran with MANSY_ROLLOUT_TIME_KERNELS, and the number of steps they cover. */
int mansy_rollout_kernel_ms(mansy_handle_t h, double *policy_ms, double *step_ms, int32_t *n_steps);

/* Raw copy of the per-environment state records (128 bytes each, layout: csrc/mansy_sim.cuh
 * EnvState / _capi.ENV_STATE_DTYPE) into state_dev[n_envs*128]; for host-side bookkeeping
 * (current_video / current_user / ... attributes of the gym envs) and tests. */
int mansy_state_snapshot(mansy_handle_t h, void *state_dev, void *stream);

/*
 * MPC expert: ExpertEnv.choose_action (bitrate_selection/envs/expert_env.py:358-422 with ExpertSimulator's virtual
 * downloads, simulators/simulator.py:125-144, and QoEModelExpert, utils/qoe.py:50-60) for every environment of the
 * handle at its current state: exhaustive search over the 15^H action sequences of the next H = min(horizon, chunks
 * left) chunks, first action of the first best sequence.  actions_dev int32 [n_envs]; best_value_dev (float64 [n_envs],
 * may be NULL) receives the winning QoE sum.  horizon 1..6 (run_expert.py:169 default 4).  The state is not modified.
 */
int mansy_expert_actions(mansy_handle_t h, int32_t horizon, int32_t *actions_dev, double *best_value_dev, void *stream);

/* Non-zero once a kernel met a data error (a trace that can never finish a download). */
int mansy_error_flag(mansy_handle_t h, int32_t *flag_host);

/* CPU self-tests of the scalar building blocks the kernels are made of (mansy_core.cuh, the same
 * __host__ __device__ code), so the non-GPU test-suite can check them against the oracle.  Not a
 * product path: nothing in the package calls them. */
int mansy_selftest_allocate(uint64_t mask, int32_t action, const int32_t video_rates[5], uint8_t versions_out[64]);
int mansy_selftest_fov_mask(int32_t x, int32_t y, int32_t width, int32_t height, int32_t fov_w, int32_t fov_h,
                            uint64_t *mask_out, int32_t *valid_out);
int mansy_selftest_centre_to_pixel(float v, int32_t length);
int mansy_selftest_download(const double *thr, int32_t trace_len, int64_t size, int32_t *cur_idx, double *cur_time,
                            double *buf, double *download_time, double *rebuffer);
/* out[i] = ddiv_rcp(a[i], b[i], 1 / b[i]): the reciprocal-based division the kernels use (csrc/mansy_core.cuh); must equal a[i] / b[i] bit for bit */
int mansy_selftest_ddiv_rcp(const double *a, const double *b, int64_t n, double *out);
int mansy_selftest_hashed_action(uint64_t seed, uint64_t env, uint64_t step);

/*
 * MTIO viewport-prediction transformer, inference path (BASELINE config 5):
 * replaces ViewportTransformerMTIO.sample (viewport_prediction/models/mtio.py:106-133) as predict.py:27 calls it.
 * Architecture as the reference constructs it (mtio.py:48-64, customized_transformer.py:39-51, predict.py:73-74):
 * d_model = dim_feedforward = 512, nn.Transformer defaults (8 heads, post-norm, ReLU), 1..4 encoder and decoder
 * layers, DistillLayer between them, 6-wide tokens (in_channel 2 x 3 MTIO heads), sigmoid predictor.
 * Weights are HOST float32 arrays in the layout of the reference's state dict (torch [out][in] matrices); any
 * *_b pointer may be NULL = zeros (under torch >= 2.1 the reference's positional constructor arguments switch the
 * transformer's biases off, customized_transformer.py:47-50).  Everything is copied at create.
 */
#define MANSY_MTIO_MAX_LAYERS 4
typedef struct {
  const float *in_proj_w; /* [1536][512]  q | k | v */
  const float *in_proj_b; /* [1536] or NULL */
  const float *out_w;     /* [512][512] */
  const float *out_b;     /* [512] or NULL */
} mansy_mtio_attn_t;
typedef struct {
  mansy_mtio_attn_t self_attn;
  mansy_mtio_attn_t cross_attn; /* decoder layers only (multihead_attn) */
  const float *lin1_w, *lin1_b; /* [512][512], [512] */
  const float *lin2_w, *lin2_b;
  const float *norm1_w, *norm1_b, *norm2_w, *norm2_b;
  const float *norm3_w, *norm3_b; /* decoder layers only */
} mansy_mtio_layer_t;
typedef struct {
  int32_t n_enc, n_dec;   /* layers (predict.py --block-num, default 2) */
  int32_t his_window;     /* source tokens (default 5), <= 16 */
  int32_t fut_window;     /* autoregressive steps (default 15), <= 31 */
  int32_t pe_rows;        /* rows of `pe` (>= max(his_window, fut_window)) */
  int32_t reserved;
  const float *emb_w, *emb_b; /* embedding.linear [512][6], [512] */
  const float *pe;            /* positional_embedding.pe [pe_rows][512] */
  mansy_mtio_layer_t enc[MANSY_MTIO_MAX_LAYERS], dec[MANSY_MTIO_MAX_LAYERS];
  const float *enc_norm_w, *enc_norm_b, *dec_norm_w, *dec_norm_b;
  const float *conv_w, *conv_b; /* distill_layer.downConv [512][512][3], [512] */
  const float *bn_w, *bn_b, *bn_mean, *bn_var; /* distill_layer.norm (eval mode) */
  const float *pred_w, *pred_b; /* predictor.0 [6][512], [6] */
} mansy_mtio_weights_t;
typedef struct mansy_mtio *mansy_mtio_t;

/* max_batch = samples processed per pass (workspace is sized for it; larger calls run in chunks). */
int mansy_mtio_create(const mansy_mtio_weights_t *weights, int device, int32_t max_batch, mansy_mtio_t *out);
int mansy_mtio_destroy(mansy_mtio_t m);
#define MANSY_MTIO_FP32 1         /* exact-fp32 CUDA-core GEMMs (parity anchor) instead of tcgen05 kind::tf32 */
#define MANSY_MTIO_TIME_KERNELS 2 /* CUDA events around every launch (see mansy_mtio_kernel_ms) */
/*
 * history_dev [n][his_window][2], current_dev [n][1][2] (viewport centres in [0,1]^2, predict.py:24-27) ->
 * pred_dev [n][fut_window][2]: the MTIO-head ensemble, wrapped into the unit square (mtio.py:124-132).
 * tokens_dev ([n][fut_window + 1][6], may be NULL) receives the decoder input tokens (test hook).
 * The encoder runs once per sample and the decoder keeps its keys / values between the autoregressive steps:
 * in eval mode that is the same function as the reference's per-step re-encoding (mtio.py:120-123).
 * n_steps (0 = fut_window) stops the autoregression early: rows [n_steps, fut_window) of pred_dev / tokens_dev are
 * left untouched.  predict.py:39-44 builds its tile masks from the first `dataset_frequency` (5) predicted points
 * only, and a prediction never depends on later steps, so the mask pipeline needs 5 of the 15 steps.
 */
int mansy_mtio_sample(mansy_mtio_t m, const float *history_dev, const float *current_dev, int32_t n, int32_t n_steps, int32_t flags,
                      float *pred_dev, float *tokens_dev, void *stream);
/* Same with HOST buffers (pinned for overlap): copies in, runs, copies out, synchronises the stream. */
int mansy_mtio_sample_host(mansy_mtio_t m, const float *history_host, const float *current_host, int32_t n, int32_t n_steps,
                           int32_t flags, float *pred_host, void *stream);
/* After a MANSY_MTIO_TIME_KERNELS call and a stream synchronise: summed durations (ms) and launch counts of
 * [0] the GEMM kernels, [1] the attention kernels, [2] everything else (embedding, norms, distillation, head). */
int mansy_mtio_kernel_ms(mansy_mtio_t m, double ms[3], int32_t launches[3]);

/* The `--model regression` predictor of predict.py: LinearRegression.sample
 * (viewport_prediction/models/linear_regression.py:16-33), an ordinary least-squares line per sample and coordinate
 * through history + current (his_window + 1 points), extrapolated fut_window steps.  Device pointers, layouts as above. */
int mansy_linreg_sample(const float *history_dev, const float *current_dev, int64_t n, int32_t his_window, int32_t fut_window,
                        float *pred_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MANSY_B200_H_ */
