/*
 * srl_policy.h -- C-ABI of the helpers a GPU-resident PPO2 consumer needs next to srl_sim_step (same shared library,
 * libsrl_sim_b200.so): the policy step and the observation filter of the collection loop, and the minibatch gradient of the update.  They belong to the
 * CONSUMER of the simulator, not to the environment boundary of include/srl_sim.h.
 *
 * Reference interfaces they replace (paths relative to the reference repo):
 *   srl_policy_act  <- stable-baselines 2.5 `PPO2` runner's `model.step(obs)` with `MlpPolicy` (two separate 64-64 tanh
 *                      towers), chosen by rl_baselines/rl_algorithm/ppo2.py:58-72; one call per env step:
 *                      policy forward, sample, log-probability, value, rollout-buffer writes
 *   srl_ppo2_grad   <- the loss + `tf.gradients` of stable-baselines 2.5 `PPO2.setup_model`, run once per minibatch by `PPO2._train_step`
 *   srl_obs_filter  <- stable-baselines `VecNormalize._obfilt` (norm_obs=True, clip_obs=10), wrapped around the envs
 *                      by rl_baselines/utils.py:224-227: running mean / variance update + normalisation
 *
 * Conventions are those of srl_sim.h: 0 on success, message from srl_sim_last_error(); all pointers are DEVICE pointers;
 * calls are asynchronous on `stream` and capturable into a CUDA graph (everything a launch reads that changes between
 * replays -- weights, filter state, RNG counter -- lives in device memory).
 */
#ifndef SRL_POLICY_H_
#define SRL_POLICY_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* stable-baselines MlpPolicy: separate policy / value towers obs_dim -> 64 -> 64 -> n_out / 1, tanh.
 * Weights in torch.nn.Linear layout: weight [out][in] row-major float32, bias [out]. */
typedef struct srl_mlp_policy {
    uint32_t struct_size;   /* = sizeof(srl_mlp_policy); checked                          */
    int32_t  obs_dim;       /* 1..8                                                       */
    int32_t  n_out;         /* Discrete: number of actions (2..8); Box: action dim (1..8) */
    int32_t  discrete;      /* 1 = Categorical(logits), 0 = Normal(mean, exp(logstd))     */
    const float *pi_w1, *pi_b1, *pi_w2, *pi_b2, *pi_w3, *pi_b3;
    const float *vf_w1, *vf_b1, *vf_w2, *vf_b2, *vf_w3, *vf_b3;
    const float *logstd;    /* Box only: [n_out]                                          */
} srl_mlp_policy;

/* One policy step for n envs.
 *   obs      : f32[n, obs_dim], already normalised
 *   rng      : u64[3] device words {seed, step counter, 0}; env i samples from the Philox stream
 *              (seed, env_offset + i, counter); the launch advances the counter by one when its last CTA retires
 *   obs_buf  : nullable f32[n, obs_dim], copy of obs (rollout buffer row)
 *   act_env  : i32[n] (Discrete) / f32[n, n_out] clipped to [-1, 1] (Box): the action for srl_sim_step
 *   act_buf  : nullable i64[n] / f32[n, n_out] unclipped sample (rollout buffer row)
 *   logp, value : f32[n] */
int srl_policy_act(const srl_mlp_policy* policy, int n, const float* obs, uint64_t* rng, uint64_t env_offset,
                   float* obs_buf, void* act_env, void* act_buf, float* logp, float* value, void* stream);

/* VecNormalize's observation filter for one batch.
 *   state : f64[2 * obs_dim + 1] device words {mean[obs_dim], var[obs_dim], count}
 *   update != 0: fold the batch moments of `obs_raw` into the state first (parallel-variance merge), then
 *   out = clip((obs_raw - mean) / sqrt(var + eps), -clip, clip) in float32.  obs_dim <= 8. */
int srl_obs_filter(int n, int obs_dim, const float* obs_raw, double* state, int update, float clip, float eps,
                   float* obs_norm_out, void* stream);

/* Gradient tensors of an MlpPolicy, same shapes and layout as the weights (torch: `param.grad`, contiguous float32). */
typedef struct srl_mlp_grads {
    uint32_t struct_size;   /* = sizeof(srl_mlp_grads); checked */
    uint32_t reserved;
    float *pi_w1, *pi_b1, *pi_w2, *pi_b2, *pi_w3, *pi_b3;
    float *vf_w1, *vf_b1, *vf_w2, *vf_b2, *vf_w3, *vf_b3;
    float *logstd;          /* Box only */
} srl_mlp_grads;

/* The gradient of stable-baselines' PPO2 loss over one minibatch, in one pass: what `loss.backward()` leaves in `param.grad` for
 *   loss = pg_loss - ent_coef * entropy + vf_coef * vf_loss     (PPO2.setup_model of stable-baselines 2.5, chosen by
 *                                                                 rl_baselines/rl_algorithm/ppo2.py:58-72 of the reference)
 *   A = (adv - mean(adv)) / (std(adv) + 1e-8) over the minibatch, ratio = exp(logp - old_logp),
 *   pg_loss = mean(max(-A ratio, -A clip(ratio, 1 - c, 1 + c))),
 *   vf_loss = 0.5 mean(max((v - ret)^2, (old_value + clip(v - old_value, -c, c) - ret)^2)).
 * Gradient clipping and the optimiser step stay with the caller.
 *   idx        : nullable i64[minibatch] row indices into the rollout arrays (NULL: rows 0 .. minibatch - 1)
 *   obs        : f32[rows, obs_dim]; actions : i64[rows] (Discrete) / f32[rows, n_out] (Box, the unclipped samples)
 *   adv, ret, old_logp, old_value : f32[rows]
 *   workspace  : device memory of at least srl_ppo2_workspace_bytes(...) bytes (per-CTA partial gradients; no state between calls)
 * The gradient tensors are OVERWRITTEN (not accumulated).  Deterministic: partial sums are combined in a fixed order. */
size_t srl_ppo2_workspace_bytes(int obs_dim, int n_out, int discrete, int minibatch);
int srl_ppo2_grad(const srl_mlp_policy* policy, const srl_mlp_grads* grads, int minibatch, const int64_t* idx, const float* obs,
                  const void* actions, const float* adv, const float* ret, const float* old_logp, const float* old_value,
                  float cliprange, float ent_coef, float vf_coef, void* workspace, size_t workspace_bytes, void* stream);

/* GAE(lambda) over one rollout (the backward recursion of stable-baselines' PPO2 runner): rew, value, done (1.0 where the episode ended at that
 * step), adv_out, ret_out are f32[n_steps, n_envs]; last_value f32[n_envs] is the value of the observation after the last step.
 *   delta_t = rew_t + gamma * V_{t+1} * (1 - done_t) - V_t;  adv_t = delta_t + gamma * lam * (1 - done_t) * adv_{t+1};  ret_t = adv_t + V_t */
int srl_ppo2_gae(int n_steps, int n_envs, const float* rew, const float* value, const float* done, const float* last_value, float gamma, float lam,
                 float* adv_out, float* ret_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRL_POLICY_H_ */
