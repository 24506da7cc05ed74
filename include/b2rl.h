/* b2rl.h — C ABI of the B200-native learner-side replay path.
 *
 * The reference (seungju-k1m/Distributed_RL) is pure Python and defines no
 * FFI of its own; its boundary for this path is three duck-typed Python
 * surfaces.  Every entry point below therefore cites the
 * reference *Python* interface it replaces (paths relative to the reference
 * root).  The Python mirrors in distributed_rl_b200/ bind these with ctypes
 * (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - every function returns 0 on success, <0 on error; the message is
 *     available (per thread) from b2rl_last_error().
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it
 *     and is asynchronous w.r.t. the host unless stated otherwise.
 *   - pointers named *_dev are device pointers owned by the caller (e.g.
 *     torch tensors' data_ptr()); the library never frees them.
 *   - the sum-tree and the payload arrays are owned by the handle and freed
 *     by b2rl_replay_destroy().
 *   - one producer + one consumer thread may use a handle concurrently as
 *     long as they enqueue on the same stream (stream order replaces the
 *     reference's `lock` flag handshake, APE_X/ReplayMemory.py:151-160).
 */
#ifndef B2RL_H_
#define B2RL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2RL_MAX_FIELDS 8
#define B2RL_OK 0
#define B2RL_ERR_INVALID (-1)
#define B2RL_ERR_CUDA (-2)
#define B2RL_ERR_NOMEM (-3)

typedef struct b2rl_replay b2rl_replay; /* opaque */

/* One replay shard: a ring of `capacity` slots, `n_fields` SoA payload arrays
 * (field f holds capacity x field_bytes[f] bytes) and an fp64 sum-tree + fp32
 * min-tree over the slot priorities.
 * Replaces the storage of baseline/PER.py:48-66 (`memory` list of pickled
 * records + `Tree.prior_torch`) and baseline/utils.py:328-333
 * (PrioritizedMemory: CompressedDeque + SumTree). */
typedef struct {
  int64_t capacity;                      /* slots; the tree is padded to 2^k */
  int32_t n_fields;                      /* 0..B2RL_MAX_FIELDS               */
  int32_t device;                        /* CUDA device ordinal              */
  int64_t field_bytes[B2RL_MAX_FIELDS];  /* bytes per slot of each field     */
} b2rl_replay_desc;

const char* b2rl_last_error(void);
int b2rl_version(void);

/* PER.__init__ (baseline/PER.py:49-66), PrioritizedMemory.__init__
 * (baseline/utils.py:329-332). */
int b2rl_replay_create(const b2rl_replay_desc* desc, b2rl_replay** out);
int b2rl_replay_destroy(b2rl_replay* h);

/* PER.__len__ (baseline/PER.py:80-81): number of valid slots, capacity, and
 * the ring head (next slot to be written). Host-side, no sync. */
int b2rl_replay_size(const b2rl_replay* h, int64_t* size, int64_t* capacity, int64_t* head);

/* Device base pointer of payload field f (capacity x field_bytes[f] bytes): the storage behind
 * PER.memory / Tree.data (baseline/PER.py:7-28), exposed so that consumers (b2rl_conv1_fused) read rows in place. */
int b2rl_replay_field_ptr(const b2rl_replay* h, int32_t field, void** ptr_dev);

/* PER.push (baseline/PER.py:69-75) / PrioritizedMemory.push
 * (baseline/utils.py:334-337): append n records.  fields_src[f] points to
 * n x field_bytes[f] contiguous bytes (host — ideally pinned — or device;
 * copied with cudaMemcpyDefault), prios to n fp32 priorities (same rule).
 * Slots are written at the ring head; when the ring is full the oldest slots
 * are overwritten (FIFO, as PER.remove_to_fit :118-127 drops them — but slot
 * ids stay stable instead of being renumbered).  Leaves and their root paths
 * are refreshed in the same call.  n <= capacity. */
int b2rl_replay_push(b2rl_replay* h, const void* const* fields_src, const float* prios,
                     int64_t n, void* stream);

/* Pipelined ingest — the same PER.push (baseline/PER.py:69-75; drained from Redis by Replay.run,
 * APE_X/ReplayMemory.py:128-139), split so that the host->device copy of the NEXT batch of
 * records can run on its own stream while the learner step of the current batch computes:
 *   reserve   (learner stream) priorities of the next n ring slots := 0, so the records about to
 *             be overwritten can no longer be sampled; *start_slot receives the first slot
 *   copy      (ingest stream, after an event recorded behind `reserve`) payload rows only
 *   commit    (learner stream, after the copy's event) priorities + root paths; advances the ring
 * reserve/commit must alternate; n <= capacity. */
int b2rl_replay_reserve(b2rl_replay* h, int64_t n, int64_t* start_slot, void* stream);
int b2rl_replay_copy_payload(b2rl_replay* h, const void* const* fields_src, int64_t start_slot, int64_t n,
                             void* stream);
int b2rl_replay_commit(b2rl_replay* h, const float* prios, int64_t n, void* stream);

/* The steady-state form of the ingest loop of Replay.run (APE_X/ReplayMemory.py:128-139: drain -> PER.push) as ONE
 * call per learner iteration: (1) the batch whose host->device copy the previous call started is published —
 * `stream` waits for that copy, then its priorities are written (PER.push's priority append, baseline/PER.py:69-75);
 * (2) the ring slots of the next batch are retired (priority 0: unsampleable while they are overwritten) and its
 * payload + priorities are copied on a library-owned copy stream, overlapping whatever is enqueued on `stream`
 * next.  Host buffers (pinned) must stay valid and unmodified until their copy has executed (the call is
 * asynchronous: synchronize `stream` after the NEXT call, or keep enough staging sets).  fields_src == NULL:
 * publish only (flush).
 * Do not interleave with b2rl_replay_reserve / b2rl_replay_commit. */
int b2rl_replay_ingest_pipelined(b2rl_replay* h, const void* const* fields_src, const float* prios_src, int64_t n,
                                 void* stream);

/* PER.remove_to_fit (baseline/PER.py:118-127): drop the `delta` oldest
 * records (priority := 0 so they can never be sampled; size -= delta). */
int b2rl_replay_evict(b2rl_replay* h, int64_t delta, void* stream);

/* Benchmark / property-test helper (no reference counterpart; stands in for a replay pre-filled through
 * PER.push, baseline/PER.py:69-75, with SURVEY.md §8d's synthetic frames): fill slots
 * [0, n) of every field with the counter hash documented in DESIGN.md §4
 * (word w of slot s of field f = lowbias32(seed ^ f*0x9E3779B9 ^ s*2654435761
 * ^ w*2246822519)) and mark them valid.  Priorities are NOT touched: follow
 * with b2rl_tree_build(). */
int b2rl_replay_fill_hash(b2rl_replay* h, int64_t n, uint32_t seed, void* stream);

/* Bulk (re)build: priorities of slots [0, n) := prios_dev[0..n), the rest 0,
 * then every internal node is recomputed (node = left + right in fp64, as
 * baseline/sumtree.py Node._reduce :21-27).  State equals
 * SumTree.extend(prios) (:97-99) / Tree.push (baseline/PER.py:22-28). */
int b2rl_tree_build(b2rl_replay* h, const float* prios_dev, int64_t n, void* stream);

/* PER.sample (baseline/PER.py:92-116) + the IS-weight lines of
 * APE_X/ReplayMemory.py:65-67 and PER.max_weight (:129-133), with the
 * descent rule of SumTree.prioritized_sample / Node._find
 * (baseline/sumtree.py:53-62,128-140): pos = root*u; at each node
 * `pos < left ? left : (pos -= left, right)`.
 *   u01_dev   n fp64 uniforms in [0,1) on the device, or NULL to draw them
 *             from Philox4x32-10 (seed, counter = rng_offset + k)
 *   idx_out   int64[n]  sampled slot ids (with replacement)
 *   prob_out  fp32[n]   p_i / sum(p)               (may be NULL)
 *   w_out     fp32[n]   (1/(N*prob))^beta / max_w  (may be NULL)
 *   max_w_dev NULL: max_w is this shard's own max_j (N*prob_j)^-beta.  Non-NULL:
 *             one device float used instead — the all-reduced MAX over the shards
 *             of a multi-GPU replay (SURVEY.md §8e "priority-max reduction").
 * Sampling from an empty tree is an error. */
int b2rl_tree_sample(b2rl_replay* h, const double* u01_dev, uint64_t seed, uint64_t rng_offset,
                     int64_t n, float beta, const float* max_w_dev, int64_t* idx_out_dev,
                     float* prob_out_dev, float* w_out_dev, void* stream);

/* Same (PER.sample, baseline/PER.py:92-116), but the uniforms come from the handle's DEVICE-RESIDENT Philox stream
 * {seed, counter} (set with b2rl_replay_seed), and the counter is advanced by n
 * on the stream afterwards — so the call can be captured in a CUDA graph and
 * every replay draws fresh numbers.  Draw k of the call uses counter + k. */
int b2rl_replay_seed(b2rl_replay* h, uint64_t seed, uint64_t counter, void* stream);
int b2rl_tree_sample_stream(b2rl_replay* h, int64_t n, float beta, const float* max_w_dev,
                            int64_t* idx_out_dev, float* prob_out_dev, float* w_out_dev, void* stream);

/* b2rl_tree_sample_stream + the part of Replay.buffer (APE_X/ReplayMemory.py:74-93) that unpacks the SCALAR
 * fields of the sampled records (action, reward, done): for every field f with small_fields_out_dev[f] != NULL
 * (field_bytes[f] must be 1, 2, 4 or 8) row idx[k] is copied to small_fields_out_dev[f] + k*field_bytes[f] by the
 * sampling thread itself, so a learner step needs no separate gather launch for them (the frame fields are read
 * in place by b2rl_conv1_fused).  small_fields_out_dev may be NULL (= b2rl_tree_sample_stream). */
int b2rl_tree_sample_fetch(b2rl_replay* h, int64_t n, float beta, const float* max_w_dev,
                           int64_t* idx_out_dev, float* prob_out_dev, float* w_out_dev,
                           void* const* small_fields_out_dev, void* stream);

/* The uniforms b2rl_tree_sample would draw for (seed, rng_offset) in place of torch.multinomial's generator
 * (baseline/PER.py:97) — lets a
 * test replay a device-RNG run through the oracle. */
int b2rl_philox_uniforms(uint64_t seed, uint64_t rng_offset, int64_t n, double* out_dev,
                         void* stream);

/* PER.update (baseline/PER.py:83-90 -> Tree.update :36-42) and
 * PrioritizedMemory.update_priorities (baseline/utils.py:347-350): set
 * priority[idx[k]] = vals[k] for k = 0..n-1 in order — for a duplicated index
 * the LAST occurrence wins — and refresh the touched root paths.
 * Deterministic (no floating-point atomics). */
int b2rl_tree_update(b2rl_replay* h, const int64_t* idx_dev, const float* vals_dev, int64_t n,
                     void* stream);

/* PrioritizedMemory.total_prios (baseline/utils.py:359-360) and
 * PER.max_weight (baseline/PER.py:129-133).  stats_out_dev receives 3
 * doubles: {sum(p), min valid p, max IS weight for `beta`}; max_w_out_dev (may be
 * NULL) receives the max IS weight as one fp32 (the operand of the multi-GPU
 * MAX all-reduce). */
int b2rl_tree_stats(b2rl_replay* h, float beta, double* stats_out_dev, float* max_w_out_dev,
                    void* stream);

/* Tree.prior_torch (baseline/PER.py:17) read-back: priorities of slots
 * [start, start+n) as fp32. */
int b2rl_tree_leaves(b2rl_replay* h, int64_t start, int64_t n, float* out_dev, void* stream);

/* Minibatch assembly of Replay.buffer (APE_X/ReplayMemory.py:61-116,
 * R2D2/ReplayMemory.py:53-122, IMPALA/ReplayMemory.py:30-54): for every
 * field f with out_fields[f] != NULL copy row idx[k] to out_fields[f] + k *
 * field_bytes[f].  Rows that are multiples of 16 B go HBM -> SMEM -> HBM with
 * bulk async copies (TMA); the rest through a vectorised byte kernel. */
int b2rl_replay_gather(b2rl_replay* h, const int64_t* idx_dev, int64_t n,
                       void* const* out_fields_dev, void* stream);

/* Learner.train target section, Ape-X (APE_X/Learner.py:85-121):
 *   a* = argmax_a qn_online[b,:];  y = r + gamma_n * qn_target[b,a*] * notdone
 *   d = clamp(y - q_s[b,action[b]], -1, 1);  prio = (|d| + 1e-7)^alpha
 *   loss = 0.5 * mean(w * d^2);  grad_q = dLoss/dq_s  (dense B x A)
 * q_* are (B, A) fp32 row-major; action int64[B]; reward/notdone/weight
 * fp32[B].  scalars_out_dev receives 3 floats {loss, mean(y), mean(w)}.
 * Any output pointer may be NULL. */
int b2rl_apex_target(const float* q_s_dev, const float* qn_online_dev, const float* qn_target_dev,
                     const int64_t* action_dev, const float* reward_dev, const float* notdone_dev,
                     const float* weight_dev, int32_t B, int32_t A, float gamma_n, float alpha,
                     float* target_out_dev, float* td_out_dev, float* prio_out_dev,
                     float* grad_q_out_dev, float* scalars_out_dev, void* stream);

/* Learner.train target section, R2D2 (R2D2/Learner.py:110-198) with
 * value_transform / value_inv_transform (:22-35).  Time-major layouts:
 * q, q_target (L, B, A); action int64 (L-1, B); reward fp32 (L-1, B);
 * notdone fp64-valued but passed as fp32[B] (0/1); weight fp32[B].
 * Outputs: target, td (L-1, B); prio fp32[B] =
 * (0.9 max_t|td| + 0.1 mean_t|td|)^alpha; grad_q (L, B, A);
 * scalars {loss, mean q(s,a)}. */
int b2rl_r2d2_target(const float* q_dev, const float* q_target_dev, const int64_t* action_dev,
                     const float* reward_dev, const float* notdone_dev, const float* weight_dev,
                     int32_t L, int32_t B, int32_t A, int32_t n_step, double gamma, float alpha,
                     int32_t use_rescaling, float* target_out_dev, float* td_out_dev,
                     float* prio_out_dev, float* grad_q_out_dev, float* scalars_out_dev,
                     void* stream);

/* V-trace of IMPALA (IMPALA/Learner.py:141-215), (T, B) time-major fp32:
 * pi_a = learner prob of the taken action, mu_a = behaviour prob, value =
 * V(s_t), bootstrap[B] = V(s_T) * done, reward.  Outputs vtarget (T, B)
 * (:202) and advantage (T, B) (:207-212). */
int b2rl_vtrace(const float* pi_a_dev, const float* mu_a_dev, const float* value_dev,
                const float* bootstrap_dev, const float* reward_dev, int32_t T, int32_t B,
                float gamma, float c_lambda, float c_bar, float p_bar, float* vtarget_out_dev,
                float* advantage_out_dev, void* stream);

/* Fused gather + first convolution (north-star "TMA staging of sampled transition slices
 * into shared memory for the Q-network's first GEMM"): conv_1 of cfg/ape_x.json / cfg/r2d2.json
 * (8x8, stride 4, 4 -> 32 channels, no bias; baseline/baseNetwork.py:165-172) applied to
 * frames[idx[k]] / 255 (APE_X/Learner.py:61-67,78,85,87) on the tcgen05 tensor cores, for one or
 * two networks (online + target) in one pass; the sampled uint8 frames are never staged in HBM.
 * c_out = 32 (cfg/ape_x.json, cfg/r2d2.json) or 16 (cfg/impala.json:25-37).
 *   b2rl_conv1_pack   w_dev fp32 [c_out][4][8][8] of network `net` -> packed int8 digits
 *                     (bq_out: n_nets*4*c_out*256 bytes) and per-channel scale (scale_out: n_nets*c_out fp32)
 *   b2rl_conv1_fused  frames_dev: rows of 28 224 bytes (e.g. b2rl_replay_field_ptr of the state
 *                     field), idx_dev int64[n] or NULL (rows 0..n-1), out_dev fp32
 *                     [n_nets][n][20][20][c_out] (NHWC), relu != 0 applies ReLU. */
int b2rl_conv1_pack(const float* w_dev, int32_t net, int32_t n_nets, int32_t c_out, int8_t* bq_out_dev,
                    float* scale_out_dev, void* stream);
/* Up to 4 such packs in ONE launch (host arrays of `jobs` entries): the learner step packs the online conv_1 weights
 * for its one-network and its two-network b2rl_conv1_fused launch and the target weights for the latter
 * (APE_X/Learner.py:78,85,87 evaluate conv_1 with both parameter sets every step). */
int b2rl_conv1_pack_jobs(const float* const* w_dev, const int32_t* net, const int32_t* n_nets,
                         int8_t* const* bq_out_dev, float* const* scale_out_dev, int32_t jobs, int32_t c_out,
                         void* stream);
int b2rl_conv1_fused(const uint8_t* frames_dev, int64_t capacity, const int64_t* idx_dev, int64_t n,
                     const int8_t* bq_dev, const float* scale_dev, int32_t n_nets, int32_t c_out,
                     float* out_dev, int32_t relu, void* stream);

/* Weight gradient of conv_1 fused with the gather (the backward half of b2rl_conv1_fused;
 * loss.backward() in APE_X/Learner.py:123-138 for baseline/baseNetwork.py:165-172's first layer):
 *   gw[co][c][ky][kx] (+)= (1/255) * sum_{k,oy,ox} gy[k][oy][ox][co] * frames[idx[k]][c][4oy+ky][4ox+kx]
 * gy_dev: [n][20][20][c_out] fp32 (NHWC); y_relu_dev: NULL, or the post-ReLU output of b2rl_conv1_fused(relu = 1)
 * for the same rows — gy is then the gradient w.r.t. that output and is masked by (y > 0) on the fly (the
 * ReLU's backward); gw_dev: [c_out][4][8][8] fp32; workspace_dev:
 * b2rl_conv1_wgrad_workspace_floats(c_out) floats of scratch (per-SM partial sums, summed in fp64 in a
 * fixed order: the result is deterministic).  idx_dev may be NULL (rows 0..n-1). */
int64_t b2rl_conv1_wgrad_workspace_floats(int32_t c_out);
int b2rl_conv1_wgrad(const uint8_t* frames_dev, int64_t capacity, const int64_t* idx_dev, int64_t n,
                     const float* gy_dev, const float* y_relu_dev, int32_t c_out, float* workspace_dev,
                     float* gw_dev, int32_t accumulate, void* stream);

/* Learner.step (APE_X/Learner.py:123-138; IMPALA/Learner.py:258-266 without the clipping) with
 * torch.optim.RMSprop's update (baseline/utils.py getOptim :124-130; centered for Ape-X,
 * cfg/ape_x.json:27-35) in ONE pass: square_avg / grad_avg / param update, gradient zeroed, and
 * the reference's diagnostic "norm" sqrt(sum_i ||g_i||_2) written to grad_norm_out_dev (may be
 * NULL).  The four pointer arrays and numel are HOST arrays of n_tensors (<= 24) entries holding
 * device pointers of dense tensors with identical element order; sumsq_scratch_dev: n_tensors
 * doubles, zeroed once by the caller (the kernel re-zeroes them). */
int b2rl_rmsprop_step(float* const* params, float* const* grads, float* const* square_avg,
                      float* const* grad_avg, const int64_t* numel, int32_t n_tensors, double lr, double alpha,
                      double eps, int32_t centered, double* sumsq_scratch_dev, float* grad_norm_out_dev,
                      void* stream);
/* The same update issued in two parts: Learner.step (APE_X/Learner.py:123-138) has no gradient clipping, so a
 * parameter can be updated as soon as its own gradient is final — the dense heads' (97 % of the elements) while the
 * convolution stack's backward still runs.  Each part calls b2rl_rmsprop_step on its tensors with
 * grad_norm_out_dev = NULL and sumsq_scratch_dev pointing at its slots of one n-entry scratch; this call then forms
 * the reference's "norm" sqrt(sum_i ||g_i||_2) (:130) over all n slots and re-zeroes them. */
int b2rl_rmsprop_norm_finish(double* sumsq_scratch_dev, int32_t n_tensors, float* grad_norm_out_dev, void* stream);

/* The dense heads of the networks (nn.Linear, bias-free: baseline/baseNetwork.py:77-79; 3136 -> 512
 * adv/val heads cfg/ape_x.json:52-71) at fp32 accuracy on the tensor cores: every fp32 operand is
 * split into two TF32 terms and C (+)= A[M][K] * B[N][K]^T is formed from three tcgen05 products
 * with fp32 accumulation.  `split_pack` turns a row-major fp32 matrix (or its transpose) into the
 * operand image (b_role = 0: the A / M side, 1: the B / N side); `packed_floats` is the size of
 * that image in floats.  Shapes with few output tiles split K over the SMs; their partial tiles go to
 * workspace_dev (b2rl_gemm_workspace_floats floats, 0 = not needed) and are summed in a fixed order,
 * so the result is deterministic. */
int64_t b2rl_gemm_packed_floats(int64_t rows, int64_t k, int32_t b_role);
int b2rl_gemm_split_pack(const float* src_dev, int64_t src_rows, int64_t src_cols, int64_t src_ld,
                         int32_t transpose, int32_t b_role, float* out_dev, void* stream);
/* One piece of an operand (stacked weight matrices of sibling heads, cfg/ape_x.json:52-71): image rows [row_offset, +rows), contraction
 * [k_offset, +k) of a total_rows x total_k operand; offsets (and inner piece sizes) multiples of 32. */
int b2rl_gemm_split_pack_into(const float* src_dev, int64_t src_rows, int64_t src_cols, int64_t src_ld,
                              int32_t transpose, int32_t b_role, float* out_dev, int64_t total_rows,
                              int64_t total_k, int64_t row_offset, int64_t k_offset, void* stream);
/* The two element-wise steps between the conv stack and the heads — the act_3 ReLU and nn.Flatten of
 * cfg/ape_x.json:37-51 (baseline/baseNetwork.py:204-209) — folded into the heads' operand packing.  y_dev is the conv
 * stack's output as it lies in memory (NHWC: [B][HW][C]); the images index features in the NCHW-flatten order
 * f = c*HW + hw the reference's weights use.  transpose = 0: A-role image of x = relu(y) as [B][C*HW] (forward);
 * transpose = 1: B-role image of x^T (the heads' weight gradient).  b2rl_unflatten_relu_mask is their backward:
 * out[b][hw][c] = gx[b][c*HW + hw] * (y[b][hw][c] > 0). */
int b2rl_gemm_pack_act_nhwc(const float* y_dev, int64_t B, int64_t HW, int64_t C, int32_t relu, int32_t transpose,
                            float* out_dev, void* stream);
int b2rl_unflatten_relu_mask(const float* gx_dev, int64_t gx_ld, const float* y_dev, int64_t B, int64_t HW, int64_t C,
                             float* out_dev, void* stream);
int64_t b2rl_gemm_workspace_floats(int64_t M, int64_t N, int64_t K, int64_t ldc);
int b2rl_gemm_tf32x3(const float* a_packed_dev, const float* b_packed_dev, float* c_dev, int64_t M,
                     int64_t N, int64_t K, int64_t ldc, float* workspace_dev, void* stream);

/* Tail of the dueling Q-network after the first dense layer of the two heads (cfg/ape_x.json:52-88: MLP
 * heads 3136-512-A and 3136-512-1, then the Add / Mean / Substract nodes executed by
 * baseline/baseAgent.py:287-309):  r = relu(h);  Q_j = r[:H].Wa[j] + r[H:].Wv - mean_i(r[:H].Wa[i]).
 * h_dev: [M][2H] pre-activations (advantage | value), wa_dev: [A][H], wv_dev: [H], q_dev: [M][A].
 * backward: gh_dev [M][2H] (may be NULL), gwa_dev [A][H] and gwv_dev [H] (both or neither), row_ws_dev:
 * M*(A+1) floats of scratch; sums over the batch run in a fixed order.  H % 32 == 0, H <= 1024, A <= 32. */
int b2rl_dueling_forward(const float* h_dev, int64_t M, int64_t H, const float* wa_dev, int64_t A,
                         const float* wv_dev, float* q_dev, void* stream);
int b2rl_dueling_backward(const float* h_dev, const float* gq_dev, int64_t M, int64_t H, const float* wa_dev,
                          int64_t A, const float* wv_dev, float* gh_dev, float* gwa_dev, float* gwv_dev,
                          float* row_ws_dev, void* stream);
/* The weight-gradient half alone (dL/dW of the heads' second layers, cfg/ape_x.json:52-71; part of loss.backward(),
 * APE_X/Learner.py:123-138), from the row table a previous b2rl_dueling_backward (with gwa_dev = NULL)
 * left in row_ws_dev: lets the caller run it on another stream than the dL/dh half. */
int b2rl_dueling_backward_w(const float* h_dev, const float* row_ws_dev, int64_t M, int64_t H, int64_t A,
                            float* gwa_dev, float* gwv_dev, void* stream);

/* Number of kernels this library has launched in this process (bench.py's
 * `gpu_launches`). */
int64_t b2rl_launch_count(void);

/* Replay-sharded data parallelism (the reference has one learner process and no collective:
 * APE_X/Learner.py:123-138 steps a single model).  Mean all-reduce of the SMALL gradient slice that is left when
 * backward ends (the convolution stack: 0.3 MB) across the learner ranks of one node, as one kernel per rank over
 * NVLink / NVSwitch peer memory: stage -> per-CTA flag to every peer -> read every rank's staged slice through the
 * peer mapping, add in rank order (bit-identical on all ranks), scale, write data_dev in place.
 * stage_ptrs_dev / flag_ptrs_dev: DEVICE arrays of `world` device pointers — rank r's staging buffer
 * (2 * stage_cap_floats floats) and flag pad (world * b2rl_peer_allreduce_max_ctas() uint32, zeroed once), both
 * mapped into this process (CUDA IPC / VMM; the Python host uses torch symmetric memory).  epoch_dev:
 * max_ctas uint32 zeroed once, private to the rank; error_dev: set to 1 if a peer never arrived (bounded spin).
 * Every rank must call it the same number of times with the same n. */
int32_t b2rl_peer_allreduce_max_ctas(void);
int b2rl_peer_allreduce_mean(const uint64_t* stage_ptrs_dev, const uint64_t* flag_ptrs_dev, int32_t rank,
                             int32_t world, int64_t stage_cap_floats, float* data_dev, int64_t n,
                             uint32_t* epoch_dev, uint32_t* error_dev, void* stream);
/* The LARGE slice (the dense heads' 12.9 MB, cfg/ape_x.json:52-71) as reduce-scatter + all-gather in one kernel:
 * the gradient bucket itself is peer-mapped (bucket_ptrs_dev[r] = rank r's slice start, n floats); rank r reduces
 * floats [r * slice_floats, ...) of every rank's bucket in rank order into its own bucket and into its result buffer
 * (result_ptrs_dev[r]: 2 * slice_floats floats), then gathers every peer's result.  Flag pads: 2 * world * max_ctas
 * uint32 per rank, zeroed once.  `ctas` CTAs (<= max_ctas, the same on every rank); launched on the stream that
 * produced the gradients, it overlaps the rest of backward (new work, no reference counterpart). */
int b2rl_peer_allreduce_mean_big(const uint64_t* bucket_ptrs_dev, const uint64_t* result_ptrs_dev,
                                 const uint64_t* flag_ptrs_dev, int32_t rank, int32_t world, int64_t n,
                                 int64_t slice_floats, int32_t ctas, uint32_t* epoch_dev, uint32_t* error_dev,
                                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2RL_H_ */
