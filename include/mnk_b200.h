/*
 * mnk_b200.h -- C ABI of libmnk_b200.so: the B200 (sm_100a) implementation of the batched MNK
 * environment step, the self-play opponent turn and the rollout store of
 * michal-szadkowski/rl-selfplay-mnk.
 *
 * The reference has no FFI layer (it is pure Python/PyTorch); each entry point below names the
 * reference function it replaces (paths relative to the reference repo).  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer into caller-owned memory unless its name starts with
 *     `host_`.  The library allocates no persistent device memory and keeps no global state.
 *   - Every call is an asynchronous launch on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream).  No call synchronises except the *_host entry points, which say so.
 *     All calls are CUDA-graph capturable except the *_host ones.
 *   - Return value: 0 = ok; > 0 = a cudaError_t from the launch; < 0 = MNK_ERR_* argument error.
 *   - bool tensors travel as uint8_t (torch.bool storage), 0 or 1.
 *   - There is no CPU fallback: without a CUDA device every compute call returns a cudaError_t.
 *
 * Board state in HBM (struct mnk_state)
 *   Each player's stones are a bitboard with ROW STRIDE n+1: bit (r*(n+1) + c) <=> cell (r, c); the
 *   extra column is a permanently-zero guard that stops horizontal / diagonal lines from wrapping
 *   across a row end.  A plane is `words` = ceil(m*(n+1)/64) uint64 words.  Planes are stored
 *   structure-of-arrays so that one warp reads 32 consecutive envs' word w in one 256-byte request:
 *       bits[(player * words + w) * num_envs + env]           player 0 = black, 1 = white
 *       meta[env] = (move_count << 1) | current_player
 */
#ifndef MNK_B200_H
#define MNK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MNK_B200_VERSION 200 /* major*10000 + minor*100 + patch */

#define MNK_OK 0
#define MNK_ERR_NULL (-1)     /* a required pointer is NULL */
#define MNK_ERR_GEOM (-2)     /* m, n, k unsupported: need 1 <= k <= min(m, n), n <= 32, m*(n+1) <= 512 */
#define MNK_ERR_ALIGN (-3)    /* a buffer is not aligned as documented */
#define MNK_ERR_ARG (-4)      /* other invalid argument (negative count, words mismatch, ...) */

#define MNK_MAX_WORDS 8

/* flags of mnk_step / mnk_step_host */
#define MNK_STEP_ACTIONS_I32 1u /* `actions` is int32_t[] instead of int64_t[]                      */
#define MNK_STEP_AUTORESET 2u   /* envs that finish are reset after their outputs are written:       */
                                /* equals env.step(a) followed by env.reset(dones.nonzero())        */
#define MNK_STEP_PDL 8u         /* dense step only: launch with programmatic stream serialisation so that   */
                                /* back-to-back steps on one stream overlap launch / drain (sm_90+ PDL)     */
#define MNK_STEP_ZEROCOPY 4u    /* mnk_step_host only: host_actions / host_rd are pinned, device-mapped */
                                /* (UVA) buffers; the kernel reads / writes them over PCIe itself,   */
                                /* no staging copies (dev_actions / dev_rd may be NULL)              */
#define MNK_STEP_NOSYNC 16u     /* mnk_step_host only: enqueue, do not synchronise -- the caller waits on the */
                                /* stream before reading host_rd (lets a host loop double-buffer two env groups) */

typedef struct mnk_state {
    int32_t m, n, k;
    int32_t words;     /* must equal mnk_state_words(m, n) */
    int64_t num_envs;  /* envs held by THIS process (the local shard) */
    uint64_t* bits;    /* u64[2][words][num_envs] */
    uint32_t* meta;    /* u32[num_envs]           */
} mnk_state_t;

/* ---- introspection ------------------------------------------------------------------------ */
int mnk_version(void);
const char* mnk_error_string(int code);
/* words per plane for an m x n board, or MNK_ERR_GEOM */
int mnk_state_words(int m, int n);

/* ---- environment: src/env/torch_vector_mnk_env.py ------------------------------------------ */

/* TorchVectorMnkEnv.reset (torch_vector_mnk_env.py:34-42).  idx == NULL resets every env,
 * otherwise the n_idx listed envs (player := black, move_count := 0).  The observe() the
 * reference ends reset with is a separate mnk_observe call. */
int mnk_reset(const mnk_state_t* st, const int64_t* idx, int64_t n_idx, void* stream);

/* TorchVectorMnkEnv.observe (:46-53) fused with TorchSelfPlayWrapper._get_canonical_obs
 * (src/selfplay/torch_self_play_wrapper.py:99-112).
 *   obs   f32[num_envs][2][m][n] or NULL     mask  u8[num_envs][m*n] or NULL
 *   swap  u8[num_envs] or NULL: envs with swap[e] != 0 get their two planes exchanged (the
 *         "me first" canonical view of a white agent, :104-106)
 *   fix_all_masked != 0: rows with no legal cell get mask[e][0] = 1 (:108-110) */
int mnk_observe(const mnk_state_t* st, float* obs, uint8_t* mask, const uint8_t* swap,
                int fix_all_masked, void* stream);

/* TorchVectorMnkEnv.step / step_subset (:55-84) with _check_wins (:106-119) as shift-and-AND
 * line tests.  For each listed env, in the reference's order: place the mover's stone
 * unconditionally, ++move_count, win = mover has >= k in a row anywhere, draw = board full and
 * no win, reward = 1.0 on a win, done = win | draw, toggle the player (also when done).
 *   actions  i64[n_active] (i32 with MNK_STEP_ACTIONS_I32); values outside [0, m*n) place no stone
 *   idx      NULL => dense step of envs 0..num_envs-1 (n_active must equal num_envs);
 *            else i64[n_active] distinct env indices (step_subset)
 *   rewards  f32[num_envs], dones u8[num_envs]: FULL SIZE, zero for unlisted envs
 *   obs/mask as in mnk_observe (over ALL envs, after the step), either may be NULL
 *   illegal  NULL, or int32[2] pre-zeroed: strict mode.  [0] counts moves onto an occupied or
 *            out-of-range cell, [1] = 0x7fffffff - (the smallest offending env index), 0 = none.
 *            The reference's validators (:86-104) are dead code, so the default (NULL) applies
 *            such moves silently. */
int mnk_step(const mnk_state_t* st, const void* actions, const int64_t* idx, int64_t n_active,
             float* rewards, uint8_t* dones, float* obs, uint8_t* mask, int32_t* illegal,
             uint32_t flags, void* stream);

/* `steps` (<= MNK_MAX_SLAB_STEPS) consecutive dense steps of ALL envs in ONE launch: exactly `steps` calls of
 * mnk_step(st, actions + i * action_stride, NULL, num_envs, ...) in a row -- TorchVectorMnkEnv.step K times,
 * src/env/torch_vector_mnk_env.py:55-84 -- for callers that already hold the actions of several steps (a recorded trace, a
 * slab of a host pipeline: mnk_step_host_loop).  Envs are independent and a CTA keeps its tile of 32 envs for the whole
 * launch, so no grid-wide synchronisation is needed between the steps.
 *   actions         step i reads i64 (i32 with MNK_STEP_ACTIONS_I32) [num_envs] at byte offset i * action_stride
 *   rewards_dones   step i writes f32 rewards[num_envs] followed by u8 dones[num_envs] at byte offset i * rd_stride
 *                   (rd_stride >= 5 * num_envs, a multiple of 4)
 *   obs, mask       NULL, or `steps` device pointers each: where step i materialises its f32 observation / u8 mask (entries
 *                   may be NULL)
 *   flags           MNK_STEP_ACTIONS_I32 | MNK_STEP_AUTORESET | MNK_STEP_PDL */
#define MNK_MAX_SLAB_STEPS 16
int mnk_step_slab(const mnk_state_t* st, const void* actions, int64_t action_stride, void* rewards_dones, int64_t rd_stride,
                  int32_t steps, float* const* obs, uint8_t* const* mask, uint32_t flags, void* stream);

/* Same as the dense mnk_step but with HOST buffers (the end-to-end path a caller without device
 * tensors uses): copies host_actions to dev_actions, steps, copies rewards / dones back and
 * SYNCHRONISES the stream.  host_* should be pinned.  dev_* are caller-owned scratch:
 * dev_actions i64|i32[num_envs], dev_rd = 5 * num_envs bytes (f32 rewards then u8 dones);
 * host_rd receives the same 5 * num_envs bytes.  obs / mask stay on the device (may be NULL).
 * With MNK_STEP_ZEROCOPY the step kernel dereferences the pinned host buffers directly (one launch +
 * one synchronise instead of copy + launch + copy + synchronise).  MNK_STEP_NOSYNC leaves the final
 * synchronise to the caller. */
int mnk_step_host(const mnk_state_t* st, const void* host_actions, void* dev_actions, void* dev_rd,
                  void* host_rd, float* obs, uint8_t* mask, uint32_t flags, void* stream);

/* K dense steps per call for callers whose actions are already in pinned host memory (a recorded trace, an
 * evaluation script, a host policy working one slab ahead): TorchVectorMnkEnv.step (:55-84) applied `steps` times,
 * pipelined over slabs of `slab_steps` steps -- H2D of slab i+1 | kernels of slab i | D2H of slab i-1 on three
 * streams, ONE cudaMemcpyAsync per slab and direction and ONE host wait per slab (csrc/mnk_hostloop.cu).
 * Synchronous: returns when host_rd (and host_obs / host_mask) hold all `steps` results. */
typedef struct mnk_host_loop {
    const void* host_actions;   /* pinned  i64|i32 [steps][num_envs]                                              */
    void* host_rd;              /* pinned  [steps][5*num_envs] bytes: per step f32 rewards[num_envs], u8 dones[num_envs] */
    float* host_obs;            /* NULL, or pinned f32 [steps][num_envs][2][m][n]: also bring every observation home */
    uint8_t* host_mask;         /* NULL, or pinned u8  [steps][num_envs][m*n]                                     */
    void* dev_actions;          /* device scratch [buffers][slab_steps][num_envs] actions                         */
    void* dev_rd;               /* device scratch [buffers][slab_steps][5*num_envs] bytes                         */
    float* const* obs_ring;     /* host array of `ring` device pointers f32[num_envs][2][m][n]: step t materialises */
    uint8_t* const* mask_ring;  /*   its observation / mask into slot t % ring (ring == 0: packed mode, no views) */
    int32_t ring;               /* >= buffers * slab_steps when host_obs / host_mask are set                      */
    int64_t steps, slab_steps;
    int32_t buffers;            /* slabs in flight, 2 .. MNK_HOST_LOOP_MAX_BUFFERS (0 = 2): with 3 the host waits for slab  */
                                /* i-2 after queuing slab i, so a slow copy-out no longer delays the next slab's kernels  */
} mnk_host_loop_t;
#define MNK_HOST_LOOP_MAX_BUFFERS 4

/* Streams + events of the pipeline, owned by the caller (create once, reuse across calls on the same device). */
int mnk_host_pipe_create(void** pipe);
int mnk_host_pipe_destroy(void* pipe);
/* flags: MNK_STEP_ACTIONS_I32, MNK_STEP_AUTORESET.  `pipe` may be NULL (a temporary one is made for the call). */
int mnk_step_host_loop(const mnk_state_t* st, const mnk_host_loop_t* job, void* pipe, uint32_t flags, void* stream);

/* `env.boards` (torch_vector_mnk_env.py:17) as a writable f32[num_envs][2][m][n] mirror:
 * unpack = bitboards -> f32 planes, pack = f32 planes (non-zero = stone) -> bitboards. */
int mnk_unpack_boards(const mnk_state_t* st, float* boards, void* stream);
int mnk_pack_boards(const mnk_state_t* st, const float* boards, void* stream);

/* `env.current_player` / `env.move_counts` (:18-19) as i64[num_envs]; either pointer may be NULL. */
int mnk_export_meta(const mnk_state_t* st, int64_t* current_player, int64_t* move_counts, void* stream);
int mnk_import_meta(const mnk_state_t* st, const int64_t* current_player, const int64_t* move_counts,
                    void* stream);

/* ---- policies: src/selfplay/policy.py -------------------------------------------------------- */

/* RandomPolicy.act (policy.py:13-29) straight from the bitboards: a uniformly random empty cell
 * per env (all cells when the board is full), deterministic != 0 => the first empty cell (:26-27).
 * Counter-based Philox4x32-10 keyed by `seed`, indexed by (env_offset + env, counter): results do
 * not depend on how envs are sharded over GPUs.  actions i64[num_envs]. */
int mnk_random_legal(const mnk_state_t* st, uint64_t seed, uint64_t counter, int64_t env_offset,
                     int deterministic, int64_t* actions, void* stream);

/* Categorical(logits=masked logits) of the policy heads (src/alg/architectures/resnet.py:84-94)
 * and its use in NNPolicy.act (policy.py:46-52) / PPOAgent.learn (src/alg/ppo.py:97-100):
 * illegal logits -> -inf, an all-masked row -> all zeros (uniform), sample ~ softmax,
 * log_prob(a) = logit_a - logsumexp, entropy = -sum p log p.  One warp per row.
 *   logits       f32[rows][row_stride] (row_stride >= num_actions, in elements), num_actions <= 512
 *   mask         u8[rows][num_actions] or NULL (everything legal)
 *   counter_base NULL, or a device u64 added to `counter` (a captured CUDA graph bakes `counter`; bumping the
 *                device value between replays gives fresh draws)
 *   given        NULL => draw the action (Gumbel-max on Philox(seed; row_offset + row, counter)),
 *                deterministic != 0 => argmax, first index on ties (policy.py:49-50);
 *                else i64[rows]: evaluate these actions instead of sampling
 *   actions      i64[rows] out (may be NULL when `given` is set)
 *   log_probs    f32[rows] or NULL, entropy f32[rows] or NULL */
int mnk_masked_sample(const float* logits, int64_t row_stride, const uint8_t* mask, int32_t num_actions,
                      int64_t rows, uint64_t seed, uint64_t counter, const uint64_t* counter_base, int64_t row_offset,
                      int deterministic, const int64_t* given, int64_t* actions, float* log_probs, float* entropy, void* stream);

/* ---- self-play wrapper: src/selfplay/torch_self_play_wrapper.py ------------------------------ */

typedef struct mnk_selfplay {
    uint8_t* agent_side;  /* u8[num_envs]  0 = agent plays black, 1 = white (wrapper.agent_side, :13)     */
    uint8_t* pending;     /* u8[num_envs]  envs to reset at the next step (wrapper.pending_resets, :14)   */
    uint32_t* episodes;   /* u32[num_envs] episodes started per env: the Philox counter of the side draw   */
    uint64_t seed;        /* key of the side / opponent draws                                             */
    int64_t env_offset;   /* global id of local env 0 (draws do not depend on the sharding)               */
    const uint64_t* counter_base; /* NULL, or a device u64 added to step_counter of the fused random       */
                          /* opponent (lets a captured CUDA graph draw fresh numbers on every replay)     */
} mnk_selfplay_t;

#define MNK_SP_ACTIONS_I32 1u        /* agent / opponent actions are int32_t[]                              */
#define MNK_SP_RESET_ALL 2u          /* treat every env as pending: this is wrapper.reset() (:19-30)        */
#define MNK_SP_DETERMINISTIC_OPP 4u  /* fused random opponent plays the first empty cell (policy.py:26-27)  */

/* First half of TorchSelfPlayWrapper.step (:32-56).  Envs with pending[e] are reset, get a new side
 * (forced_sides[e] if non-NULL, else the low Philox bit keyed by (seed, env_offset+e, ++episodes[e]))
 * and ignore their action; every other env plays actions[e].  Outputs the agent ply's
 * rewards f32 / terminated u8 and opp_active u8 (0 = opponent idle, 1 = opponent answers,
 * 2 = opponent opens a freshly reset game: its result is discarded, :46).  If opp_obs / opp_mask
 * are given they receive the view the reference hands to opponent_policy.act for EVERY env
 * (channels swapped where the side to move is white, raw legal mask, :83-94); rows with
 * opp_active == 0 are ignored by the second half. */
int mnk_selfplay_agent(const mnk_state_t* st, const mnk_selfplay_t* sp, const void* actions,
                       const int64_t* forced_sides, float* rewards, uint8_t* terminated, uint8_t* opp_active,
                       float* opp_obs, uint8_t* opp_mask, uint32_t flags, void* stream);

/* Second half (:58-67, :99-112): applies opp_actions where opp_active, reward -= r_opp and
 * terminated = done_opp where opp_active == 1, pending = terminated, and writes the agent's
 * canonical observation (planes swapped for a white agent) and mask (all-masked rows get cell 0). */
int mnk_selfplay_opponent(const mnk_state_t* st, const mnk_selfplay_t* sp, const void* opp_actions,
                          const uint8_t* opp_active, float* rewards, uint8_t* terminated, float* obs, uint8_t* mask,
                          uint32_t flags, void* stream);

/* Both halves in ONE launch for a RandomPolicy opponent (policy.py:13-29): the opponent's cell is
 * drawn from the bitboards with Philox(seed; env_offset+e, step_counter). */
int mnk_selfplay_step_random(const mnk_state_t* st, const mnk_selfplay_t* sp, const void* actions,
                             const int64_t* forced_sides, uint64_t step_counter, float* rewards, uint8_t* terminated,
                             float* obs, uint8_t* mask, uint32_t flags, void* stream);

/* ---- rollout storage: src/alg/rollout_buffer.py, src/alg/ppo.py:78-133 ------------------------- */

/* RolloutBuffer.add, observation + mask part (rollout_buffer.py:51,57): stores the agent's canonical
 * planes of the CURRENT state (own stones first; agent_side as in mnk_selfplay_t, NULL = raw) into
 * one packed slot  u64[2][words][num_envs]  -- 32 B per env at 9x9 instead of 729 B. */
int mnk_rollout_store_obs(const mnk_state_t* st, const uint8_t* agent_side, uint64_t* slot, void* stream);

/* RolloutBuffer.get_data_loader's b_obs[batch_idx], b_masks[batch_idx] (:101-110): materialises
 * samples index[i] (NULL = 0..count-1) of the flattened [steps * num_envs] packed rollout
 * (packed = u64[steps][2][words][num_envs]) as f32[count][2][m][n] and u8[count][m*n]
 * (all-masked rows get cell 0, as the wrapper's canonical mask). */
int mnk_rollout_gather(int32_t m, int32_t n, int32_t k, const uint64_t* packed, int64_t num_envs, const int64_t* index,
                       int64_t count, float* obs, uint8_t* mask, void* stream);

/* RolloutBuffer.compute_advantages_and_returns (:60-80): GAE(lambda) over [steps][num_envs] arrays,
 * bit-identical in fp32 to the reference's tensor expressions.  gamma / gae_lambda are the reference's Python
 * floats (doubles): gamma enters the fp32 arithmetic as (float)gamma, their product as (float)(gamma * gae_lambda)
 * -- multiplied in double first, as `gamma * gae_lambda * tensor` evaluates in Python (:76). */
int mnk_gae(const float* rewards, const float* values, const uint8_t* dones, const float* last_values, int64_t steps,
            int64_t num_envs, double gamma, double gae_lambda, float* advantages, float* returns, void* stream);

/* PPOAgent.learn's episode accounting (ppo.py:110-120) without per-step host reads: ep_reward /
 * ep_len f32[num_envs] running sums, totals f64[6] += {episodes, sum reward, sum length, wins,
 * losses, draws} of the episodes that finished this step. */
int mnk_episode_stats(const float* rewards, const uint8_t* dones, int64_t num_envs, float* ep_reward, float* ep_len,
                      double* totals, void* stream);

/* ---- policy / value network: src/alg/architectures/resnet.py ------------------------------------ */

/* The convolutional part of BaseResNetActorCritic.forward (resnet.py:73-81) for 32 channels
 * ("resnet_b_s", configs.py:28-35), eval-mode BatchNorm folded: conv_in, `blocks` residual blocks and
 * the 1x1 convolutions that open the policy (32->2) and value (32->1) heads, as one tcgen05
 * implicit-GEMM kernel reading the packed bitboards (swap[e] != 0 exchanges the planes: the mover's /
 * agent's canonical view, as in mnk_observe).  16-bit floating-point operands ("op16": IEEE fp16 unless the
 * library was built with -DMNK_ACT_BF16 -- see mnk_resnet_operand_dtype), fp32 accumulation in TMEM.
 *   weights  op16 [1+2*blocks][9 taps][4 k-chunks][32 c_out][8 c_in]  (tap = ky*3+kx; 16-byte aligned)
 *   bias     f32  [1+2*blocks][32] (16-byte aligned)   head_w f32 [3][32], head_b f32 [3] (policy0, policy1, value)
 *   policy_feat f32 [num_envs][2*m*n]  (= Flatten(Conv2d(32,2,1)(features))), value_feat f32 [num_envs][m*n]
 *   error    NULL or int32[1], set to 1 if an internal barrier wait timed out (results invalid) */
int mnk_resnet_operand_dtype(void);   /* 0 = IEEE fp16 (default build), 1 = bf16: element type of `weights` below */

int mnk_resnet_tower(const mnk_state_t* st, const uint8_t* swap, const void* weights, const float* bias,
                     const float* head_w, const float* head_b, int32_t blocks, float* policy_feat,
                     float* value_feat, int32_t* error, void* stream);

/* The convolutional body of the reference's wider networks (configs.py:36-65), eval-mode BatchNorm folded, one tcgen05
 * kernel reading the packed bitboards (csrc/mnk_convtower.cu):
 *   residual = 1: BaseResNetActorCritic.forward_body (resnet.py:68-71) -- layer 0 = conv_in, then (conv1, conv2 + skip)
 *                 pairs; `layers` = 1 + 2 * num_blocks.  "resnet_b_l": channels = 80, layers = 11
 *   residual = 0: BaseCnnActorCritic.shared_body (cnn.py:12-29) -- a plain conv3x3 + BN + ReLU stack.
 *                 "cnn_b_s": 56 channels zero-padded to channels = 64, layers = 4; "cnn_b_l": channels = 96, layers = 8
 * followed by the 1x1 convolutions that open the two heads.  channels must be 64, 80 or 96 (MNK_ERR_ARG otherwise);
 * boards up to n <= 22 whose guard-strided env (m (n+1) + n + 2 pixel rows) fits the CTA tile of 512 (384 at 96
 * channels) pixel rows, MNK_ERR_GEOM otherwise.
 *   weights  op16 [layers][9 taps][channels/8 k-chunks][channels c_out][8 c_in]  (tap = ky*3+kx; 16-byte aligned)
 *   bias     f32  [layers][channels] (16-byte aligned)   head_w f32 [3][channels], head_b f32 [3]
 *   policy_feat / value_feat / error as mnk_resnet_tower (error flag value 4) */
int mnk_conv_tower(const mnk_state_t* st, const uint8_t* swap, int32_t channels, int32_t layers, int32_t residual,
                   const void* weights, const float* bias, const float* head_w, const float* head_b,
                   float* policy_feat, float* value_feat, int32_t* error, void* stream);

/* The body of the reference's transformer networks (src/alg/architectures/transformer.py:7-92, configs.py:7-25):
 * cell_embed + pos_embed, `layers` pre-norm nn.TransformerEncoderLayer (ReLU, no dropout) and the two Conv1d(D, ., 1) that
 * open the heads, one tcgen05 kernel per call reading the packed bitboards (csrc/mnk_transformer.cu).  Supported shapes:
 * (embed_dim, heads) = (56, 4) "transformer_b_s" and (96, 8) "transformer_b_l"; boards of at most 128 cells
 * (MNK_ERR_GEOM otherwise).  With DP = embed_dim rounded up to 16, QP = 16 * heads, F = 4 * embed_dim:
 *   weights       op16, per layer mnk_transformer_layer_weight_bytes() bytes: B-operand tiles [K/8][N][8] in consumption
 *                 order -- in_proj (K = DP, N = 3 QP: Q | K | V, head_dim zero-padded to 16; two N-halves at embed_dim 96),
 *                 out_proj (K = QP, N = DP), linear1 (K = DP, N = F; two N-halves at 96), linear2 (K = F, N = DP; two
 *                 K-halves at 96); 16-byte aligned
 *   layer_params  f32 [layers][mnk_transformer_layer_params()]: norm1 weight, bias [DP] | in_proj bias [3 QP] | out_proj
 *                 bias [DP] | norm2 weight, bias [DP] | linear1 bias [F] | linear2 bias [DP]  (padding = 0)
 *   embed f32 [3][DP]: cell_embed weight of channel 0, channel 1, bias;  pos f32 [m*n][DP];  head_w f32 [3][DP], head_b f32 [3]
 *   policy_feat f32 [num_envs][2*m*n], value_feat f32 [num_envs][m*n]; error flag value 8 */
int64_t mnk_transformer_layer_weight_bytes(int32_t embed_dim, int32_t heads);
int64_t mnk_transformer_layer_params(int32_t embed_dim, int32_t heads);
int mnk_transformer_body(const mnk_state_t* st, const uint8_t* swap, int32_t embed_dim, int32_t heads, int32_t layers,
                         const void* weights, const float* layer_params, const float* embed, const float* pos,
                         const float* head_w, const float* head_b, float* policy_feat, float* value_feat,
                         int32_t* error, void* stream);

/* The same tower for boards with 3 <= m <= 10 rows (MNK_ERR_GEOM otherwise), with the three vertical taps fused
 * into the MMA's N dimension (csrc/mnk_resnet_rows.cu): identical arguments and results, except the weight layout
 *   weights_rows  op16 [1+2*blocks][3 kx][4 k-chunks][ky*32 + c_out][8 c_in]   (16-byte aligned)
 * 2.1x fewer shared-memory operand reads per layer; the faster kernel wherever the board fits. */
int mnk_resnet_tower_rows(const mnk_state_t* st, const uint8_t* swap, const void* weights_rows, const float* bias,
                          const float* head_w, const float* head_b, int32_t blocks, float* policy_feat,
                          float* value_feat, int32_t* error, void* stream);

/* The same tower with TRAIN-MODE BatchNorm, as the reference's rollout forward runs it (src/alg/ppo.py:97 calls the
 * network without .eval(); BatchNorm2d of src/alg/architectures/resnet.py:9-21,27-31): every layer normalises with the
 * mean / biased variance of THIS batch over (envs, rows, columns) and updates running_mean / running_var in place
 * (momentum, unbiased variance), exactly one conv layer per kernel launch (csrc/mnk_resnet_train.cu), 3 <= m <= 13.
 *   weights_rows  op16 [1+2*blocks][3 kx][4 k-chunks][ky*32 + c_out][8 c_in]: the UNFOLDED conv weights
 *   bn            per-layer BatchNorm parameters and statistics, f32 [1+2*blocks][32] each (layer 0 = conv_in, then
 *                 conv1 / conv2 of every block); batch_stats (may be NULL) receives f32 [1+2*blocks][64]: the batch
 *                 mean (conv bias included) and the biased batch variance of every layer
 *   scratch       caller-owned device memory, 256-byte aligned, mnk_resnet_tower_train_scratch_bytes(...) bytes
 *                 (fp16 pre-activation and op16 skip activations of the whole batch, ~24 KB per env at 9x9)
 * Statistics are reduced in a fixed order: results are bit-reproducible for a given device and batch. */
typedef struct mnk_bn_train {
    const float* gamma;      /* BatchNorm2d.weight */
    const float* beta;       /* BatchNorm2d.bias   */
    const float* conv_bias;  /* Conv2d.bias (cancels in the normalised output; enters running_mean) */
    float* running_mean;     /* updated in place */
    float* running_var;      /* updated in place */
    float* batch_stats;      /* NULL or f32 [layers][64] */
    float momentum, eps;
} mnk_bn_train_t;

int64_t mnk_resnet_tower_train_scratch_bytes(int32_t m, int32_t n, int64_t num_envs, int32_t blocks);

int mnk_resnet_tower_train(const mnk_state_t* st, const uint8_t* swap, const void* weights_rows, const mnk_bn_train_t* bn,
                           const float* head_w, const float* head_b, int32_t blocks, void* scratch, int64_t scratch_bytes,
                           float* policy_feat, float* value_feat, int32_t* error, void* stream);

/* The heads' tails after the tower (resnet.py:41-63), one kernel, fp32:
 *   logits = Linear(128,A)(ReLU(LN(128)(Linear(2A,128)(ReLU(LN(2A)(policy_feat))))))
 *   values = Tanh(Linear(128,1)(ReLU(LN(128)(Linear(A,128)(ReLU(LN(A)(value_feat)))))))
 * Linear weights are passed TRANSPOSED ([in][out], contiguous); hidden width is 128.
 * `values` may be NULL (then `value_feat` may be too): the value head is skipped -- a policy-only caller such
 * as the frozen opponent (src/selfplay/policy.py:45-52 discards the value). */
typedef struct mnk_heads_weights {
    const float *p_ln1_w, *p_ln1_b; /* [2A]                         */
    const float *p_w1t, *p_b1;      /* [2A][128], [128]             */
    const float *p_ln2_w, *p_ln2_b; /* [128]                        */
    const float *p_w2t, *p_b2;      /* [128][A], [A]                */
    const float *v_ln1_w, *v_ln1_b; /* [A]                          */
    const float *v_w1t, *v_b1;      /* [A][128], [128]              */
    const float *v_ln2_w, *v_ln2_b; /* [128]                        */
    const float *v_w2, *v_b2;       /* [128], [1]                   */
} mnk_heads_weights_t;

int mnk_resnet_heads(const float* policy_feat, const float* value_feat, int64_t rows, int32_t cells,
                     const mnk_heads_weights_t* w, float* logits, float* values, void* stream);

/* The same head tails with the three multi-output Linear layers as tcgen05 GEMMs over tiles of 128 samples
 * (csrc/mnk_heads_mma.cu; boards up to 96 cells, MNK_ERR_GEOM otherwise): op16 operands, fp32 accumulation,
 * LayerNorm / bias / ReLU / Tanh in fp32.  Weight layouts (op16, 16-byte aligned, zero padded; Kp = K rounded up to 16):
 *   w1p [Kp(2A)/8][128][8]: element (k, n) = policy Linear(2A,128).weight[n][k]      w1v [Kp(A)/8][128][8]: value Linear(A,128)
 *   w2  [16][Np][8], Np = A rounded up to 16: element (k, n) = policy Linear(128,A).weight[n][k]
 *   params f32: p_ln1_w[2A] p_ln1_b[2A] v_ln1_w[A] v_ln1_b[A] p_b1 v_b1 p_ln2_w p_ln2_b v_ln2_w v_ln2_b v_w2 (128 each) p_b2[A] v_b2[1]
 * `values` NULL skips the value head; `error` (may be NULL) is raised to 2 if an internal barrier wait timed out. */
int mnk_resnet_heads_mma(const float* policy_feat, const float* value_feat, int64_t rows, int32_t cells, const void* w1p,
                         const void* w1v, const void* w2, const float* params, float* logits, float* values,
                         int32_t* error, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MNK_B200_H */
