// MobileRobot family -- sm_100a kernels.
//
// Replaces MobileRobotGymEnv.reset/step/_reward/_termination
// (environments/mobile_robot/mobile_robot_env.py:159-222,235-280,336-363) and the three variants
// (mobile_robot_2target_env.py, mobile_robot_1D_env.py, mobile_robot_line_target_env.py) for
// thousands of independent envs held in structure-of-arrays HBM.  The reference teleports a
// fixed-base racecar (`resetBasePositionAndOrientation`, :265); there are no dynamics to integrate,
// so the kernel is the kinematic update + bump revert + reward + fixed-length episodes.
//
// Layout: one thread per env, 16-byte (double2 / int4) coalesced loads and stores of the state
// records; in the fused rollout the state lives in registers for all T steps and only
// action (4 B) in, obs (4-8 B) / reward (4 B) / done (1 B) out touch HBM per env-step.
//
// Arithmetic is float64 like the reference's numpy code, written with the explicit
// round-to-nearest intrinsics (__dadd_rn/__dmul_rn/__dsqrt_rn) so that no multiply-add is fused:
// positions, rewards and done flags are BIT-EXACT against the CPU restatement.
#include <stdio.h>
#include "common.cuh"
#include "philox.cuh"

namespace {

// module constants, mobile_robot_env.py:13-28,101-104
constexpr double MAX_X = 4.0, MAX_Y = 4.0, MIN_X = 0.0, MIN_Y = 0.0;
constexpr double DELTA_POS = 0.1;
constexpr double ROBOT_WIDTH = 0.2, ROBOT_LENGTH = 0.325 * 2;
constexpr double COLLISION_MARGIN = 0.1;
constexpr double LINE_REWARD_DIST_THRESHOLD = 0.1, LINE_ROBOT_OFFSET = 0.2;  // line_target_env.py:3-4

struct MobileEnvRegs {
    double px, py;
    double t0x, t0y, t1x, t1y;
    int counter, current_target, has_bumped;
    uint32_t episode, total_steps;
    double ep_ret, ep_len;
};

__device__ __forceinline__ double uniform_rn(double low, double high, double u) {
    // numpy RandomState.uniform: low + (high - low) * random_sample()
    return __dadd_rn(low, __dmul_rn(high - low, u));
}

template <int KIND>
__device__ __forceinline__ void mobile_reset_env(MobileEnvRegs& e, const double* __restrict__ d6, bool random_target,
                                                 uint64_t seed, uint64_t genv) {
    double d[6];
    if (d6) {
#pragma unroll
        for (int k = 0; k < 6; ++k) d[k] = d6[k];
    } else {
        uint4 r = philox4x32_10(seed, genv, e.episode, PHILOX_PURPOSE_RESET0 + 0);
        d[0] = __dadd_rn(MAX_X / 2, uniform_rn(-MAX_X / 3, MAX_X / 3, philox_u01(r.x, r.y)));  // :168
        d[1] = __dadd_rn(MAX_Y / 2, uniform_rn(-MAX_Y / 3, MAX_Y / 3, philox_u01(r.z, r.w)));  // :169
        constexpr double margin = 0.1 * MAX_X;                                                  // :176
        d[2] = d[3] = d[4] = d[5] = 0.0;
        if (random_target) {
            r = philox4x32_10(seed, genv, e.episode, PHILOX_PURPOSE_RESET0 + 1);
            d[2] = uniform_rn(MIN_X + margin, MAX_X - margin, philox_u01(r.x, r.y));
            d[3] = uniform_rn(MIN_Y + margin, MAX_Y - margin, philox_u01(r.z, r.w));
            if (KIND == SRL_ENV_MOBILE_2TARGET) {
                r = philox4x32_10(seed, genv, e.episode, PHILOX_PURPOSE_RESET0 + 2);
                d[4] = uniform_rn(MIN_X + margin, MAX_X - margin, philox_u01(r.x, r.y));
                d[5] = uniform_rn(MIN_Y + margin, MAX_Y - margin, philox_u01(r.z, r.w));
            }
        }
    }
    e.px = d[0];
    e.py = (KIND == SRL_ENV_MOBILE_1D) ? 0.0 : d[1];  // 1D_env.py:66
    // fixed targets: mobile_robot_env.py:173-174, 2target_env.py:52-53,62-63, 1D_env.py:69, line_target_env.py:56-57
    double t0x = 0.9 * MAX_X, t0y = MAX_Y * 3 / 4, t1x = 0.1 * MAX_X, t1y = MAX_Y * 3 / 4;
    if (KIND == SRL_ENV_MOBILE_1D) t0y = 0.0;
    if (KIND == SRL_ENV_MOBILE_LINE_TARGET) t0y = MAX_X;
    if (random_target) {
        t0x = d[2];
        if (KIND == SRL_ENV_MOBILE || KIND == SRL_ENV_MOBILE_2TARGET) t0y = d[3];
        if (KIND == SRL_ENV_MOBILE_2TARGET) { t1x = d[4]; t1y = d[5]; }
    }
    e.t0x = t0x; e.t0y = t0y; e.t1x = t1x; e.t1y = t1y;
    e.current_target = 0;
    e.counter = 0;
    e.has_bumped = 0;
    e.ep_ret = 0.0;
    e.ep_len = 0.0;
    e.episode += 1;
}

// getSRLState = getGroundTruth() - getTargetPos()  (srl_env.py:39-42, RELATIVE_POS = True)
template <int KIND>
__device__ __forceinline__ void mobile_store_obs(const MobileEnvRegs& e, float* __restrict__ obs, size_t i) {
    const double tx = e.current_target ? e.t1x : e.t0x;
    const double ty = e.current_target ? e.t1y : e.t0y;
    if (KIND == SRL_ENV_MOBILE_1D) {
        obs[i] = (float)__dsub_rn(e.px, tx);
    } else if (KIND == SRL_ENV_MOBILE_LINE_TARGET) {
        const double lx = __dsub_rn(tx, LINE_ROBOT_OFFSET);  // line_target_env.py:35-40, 1-vector broadcast
        reinterpret_cast<float2*>(obs)[i] = make_float2((float)__dsub_rn(e.px, lx), (float)__dsub_rn(e.py, lx));
    } else {
        reinterpret_cast<float2*>(obs)[i] = make_float2((float)__dsub_rn(e.px, tx), (float)__dsub_rn(e.py, ty));
    }
}

__device__ __forceinline__ void mobile_load(const MobileDev& m, int i, MobileEnvRegs& e, bool two_targets) {
    const double2 p = m.pos[i];
    const double2 t0 = m.tgt0[i];
    const int4 mt = m.meta[i];
    const double2 ep = m.ep[i];
    e.px = p.x; e.py = p.y;
    e.t0x = t0.x; e.t0y = t0.y;
    e.t1x = 0.0; e.t1y = 0.0;
    if (two_targets) { const double2 t1 = m.tgt1[i]; e.t1x = t1.x; e.t1y = t1.y; }
    e.counter = mt.x;
    e.current_target = mt.y & 0xff;
    e.has_bumped = (mt.y >> 8) & 1;
    e.episode = (uint32_t)mt.z;
    e.total_steps = (uint32_t)mt.w;
    e.ep_ret = ep.x; e.ep_len = ep.y;
}

__device__ __forceinline__ void mobile_store(const MobileDev& m, int i, const MobileEnvRegs& e, bool store_targets,
                                             bool two_targets) {
    m.pos[i] = make_double2(e.px, e.py);
    if (store_targets) {
        m.tgt0[i] = make_double2(e.t0x, e.t0y);
        if (two_targets) m.tgt1[i] = make_double2(e.t1x, e.t1y);
    }
    m.meta[i] = make_int4(e.counter, e.current_target | (e.has_bumped << 8), (int)e.episode, (int)e.total_steps);
    m.ep[i] = make_double2(e.ep_ret, e.ep_len);
}

template <int KIND>
__global__ void __launch_bounds__(128) mobile_reset_kernel(MobileDev m, int n, const uint8_t* __restrict__ mask,
                                                           const double* __restrict__ draws,
                                                           float* __restrict__ obs, bool random_target, uint64_t seed,
                                                           uint64_t env_offset) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mask && !mask[i]) return;
    MobileEnvRegs e;
    mobile_load(m, i, e, KIND == SRL_ENV_MOBILE_2TARGET);
    mobile_reset_env<KIND>(e, draws ? draws + (size_t)i * 6 : nullptr, random_target, seed, env_offset + (uint64_t)i);
    mobile_store(m, i, e, true, KIND == SRL_ENV_MOBILE_2TARGET);
    if (obs) mobile_store_obs<KIND>(e, obs, (size_t)i);
}

// sqrt_rn(s) <= 0.4  <=>  s <= S_THR_04 for every non-negative double s (sqrt is monotone and correctly rounded; the
// constant is the largest double whose rounded square root does not exceed 0.4, found by exact rational arithmetic).
// It lets the unshaped reward test `np.linalg.norm(.) <= REWARD_DIST_THRESHOLD` (:353) skip the square root BIT-EXACTLY.
constexpr double S_THR_04 = 0x1.47ae147ae147cp-3;

// Fused T-step rollout; T = 1 is the plain lockstep step.  Auto-reset on done reproduces the
// SubprocVecEnv worker loop (rl_baselines/utils.py:216-220): the stored obs is the post-reset one.
//
// Throughput structure (ncu round 1: one dependent global load + a ~900-cycle dependent fp64 chain per step, `long_sb` 56 %):
//  * the action / noise streams do not depend on the state: they are PREFETCHED a chunk of MOBILE_PF steps ahead into
//    registers (double buffering), MOBILE_PF independent 128-byte loads per warp in flight;
//  * each chunk is stepped in TWO PASSES.  Pass 1 is the only truly serial part -- position += action, bump test, revert,
//    step counter, episode end / reset -- and records (target - position) per step.  Pass 2 turns those into distance,
//    reward, observation and the HBM stores; its MOBILE_PF steps are independent, so the long-latency fp64 work
//    (sqrt, conversions) overlaps instead of serialising.
//  * the chunk is kept SHORT (4 steps): with one or two warps per SM nothing hides instruction fetch, and a 16-step
//    unrolled body (54 KB of SASS) ran from L2 -- `no_inst` 29 %, 0.50 ms; 8 steps 0.44 ms; 4 steps 0.33 ms (measured; an extra
//    `prefetch.global.L2` 32 steps ahead made it slower, 0.42 ms, and was dropped).
// Arithmetic per step is unchanged (same operations, same order), so results stay bit-exact.
constexpr int MOBILE_PF = 4;        // steps per chunk: the unrolled chunk body must stay small (see below)

template <bool DISCRETE>
struct ActionChunk {
    int a[DISCRETE ? MOBILE_PF : 1];
    float x[DISCRETE ? 1 : MOBILE_PF], y[DISCRETE ? 1 : MOBILE_PF];
    float nz[MOBILE_PF];
};

template <bool DISCRETE>
__device__ __forceinline__ void load_chunk(ActionChunk<DISCRETE>& c, const void* __restrict__ actions, const float* __restrict__ noise,
                                           int t0, int T, size_t N, size_t i) {
#pragma unroll
    for (int k = 0; k < MOBILE_PF; ++k) {
        const int t = t0 + k;
        if (t < T) {
            const size_t off = (size_t)t * N + i;
            if (actions) {
                if (DISCRETE) c.a[DISCRETE ? k : 0] = __ldg(reinterpret_cast<const int32_t*>(actions) + off);
                else { const float2 v = __ldg(reinterpret_cast<const float2*>(actions) + off); c.x[DISCRETE ? 0 : k] = v.x; c.y[DISCRETE ? 0 : k] = v.y; }
            }
            if (noise) c.nz[k] = __ldg(noise + off);
        }
    }
}

template <int KIND, bool DISCRETE>
__global__ void __launch_bounds__(64) mobile_rollout_kernel(MobileDev m, int n, int T, const void* __restrict__ actions,
                                                            const float* __restrict__ noise, float* __restrict__ obs,
                                                            float* __restrict__ rew, uint8_t* __restrict__ done,
                                                            float* __restrict__ ep_ret, int32_t* __restrict__ ep_len,
                                                            bool random_target, bool shape_reward, bool auto_reset,
                                                            int max_steps, uint64_t seed, uint64_t env_offset) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr bool TWO = (KIND == SRL_ENV_MOBILE_2TARGET);
    constexpr int D = (KIND == SRL_ENV_MOBILE_1D) ? 1 : 2;
    constexpr uint32_t NA = (KIND == SRL_ENV_MOBILE_1D) ? 2u : 4u;
    constexpr double mx = COLLISION_MARGIN + ROBOT_LENGTH / 2, my = COLLISION_MARGIN + ROBOT_WIDTH / 2;  // :257-258
    const uint64_t genv = env_offset + (uint64_t)i;
    MobileEnvRegs e;
    mobile_load(m, i, e, TWO);
    bool targets_dirty = false;
    const size_t N = (size_t)n;
    ActionChunk<DISCRETE> cur, nxt;
    load_chunk<DISCRETE>(cur, actions, noise, 0, T, N, (size_t)i);
    for (int t0 = 0; t0 < T; t0 += MOBILE_PF) {
        load_chunk<DISCRETE>(nxt, actions, noise, t0 + MOBILE_PF, T, N, (size_t)i);   // in flight while `cur` is stepped
        // ---------------- pass 1: the serial state chain (mobile_robot_env.py:237-268) ----------------
        double ddx[MOBILE_PF], ddy[MOBILE_PF];   // what _reward() and getSRLState() are made of: target - position
        uint32_t bump_mask = 0u, done_mask = 0u, obs_done_mask = 0u, reach_mask = 0u;
        const double ep_ret0 = e.ep_ret, ep_len0 = e.ep_len;   // the episode sums belong to pass 2
#pragma unroll
        for (int k = 0; k < MOBILE_PF; ++k) {
            const int t = t0 + k;
            if (t < T) {
                int a_disc = 0;
                float a0 = 0.f, a1 = 0.f;
                if (actions) {
                    if (DISCRETE) a_disc = cur.a[DISCRETE ? k : 0];
                    else { a0 = cur.x[DISCRETE ? 0 : k]; a1 = cur.y[DISCRETE ? 0 : k]; }
                } else {
                    const uint4 r = philox4x32_10(seed, genv, e.total_steps, PHILOX_PURPOSE_ACTION);
                    if (DISCRETE) a_disc = (int)__umulhi(r.x, NA);
                    else { a0 = (float)((double)r.x * (2.0 / 4294967296.0) - 1.0); a1 = (float)((double)r.y * (2.0 / 4294967296.0) - 1.0); }
                }
                // dv = DELTA_POS + np_random.normal(0.0, scale=NOISE_STD), NOISE_STD = 0.0 (:239-241)
                const double dv = noise ? __dadd_rn(DELTA_POS, (double)cur.nz[k]) : DELTA_POS;
                e.total_steps += 1;
                double ax = 0.0, ay = 0.0;
                if (DISCRETE) {
                    if (KIND == SRL_ENV_MOBILE_1D) ax = (a_disc & 1) ? dv : -dv;  // 1D_env.py:115
                    else { const int a = a_disc & 3; ax = (a == 0) ? -dv : (a == 1) ? dv : 0.0; ay = (a == 2) ? -dv : (a == 3) ? dv : 0.0; }  // :242-243
                } else {
                    // float32 action array * python float -> float32 product, then += into float64 (:250,255)
                    const float fdv = (float)dv;
                    ax = (double)__fmul_rn(fmaxf(fminf(a0, 1.0f), -1.0f), fdv);
                    ay = (double)__fmul_rn(fmaxf(fminf(a1, 1.0f), -1.0f), fdv);
                }
                const double prev_x = e.px, prev_y = e.py;  // :254
                e.px = __dadd_rn(e.px, ax);
                if (KIND != SRL_ENV_MOBILE_1D) e.py = __dadd_rn(e.py, ay);
                bool bumped = (e.px < mx) || (e.px > MAX_X - mx);   // :257-263
                if (KIND != SRL_ENV_MOBILE_1D) bumped = bumped || (e.py < my) || (e.py > MAX_Y - my);
                if (bumped) { e.px = prev_x; e.py = prev_y; bump_mask |= 1u << k; }
                e.has_bumped = bumped ? 1 : 0;
                e.counter += 1;  // :268
                const double tx = e.current_target ? e.t1x : e.t0x, ty = e.current_target ? e.t1y : e.t0y;
                if (KIND == SRL_ENV_MOBILE_LINE_TARGET) {
                    const double lx = __dsub_rn(tx, LINE_ROBOT_OFFSET);      // line_target_env.py:35-40,113
                    ddx[k] = __dsub_rn(lx, e.px); ddy[k] = __dsub_rn(lx, e.py);
                } else {
                    ddx[k] = __dsub_rn(tx, e.px); ddy[k] = (KIND == SRL_ENV_MOBILE_1D) ? 0.0 : __dsub_rn(ty, e.py);
                }
                if (TWO) {  // the target switch feeds later steps: decide it here (2target_env.py:170-173)
                    const double sq = __dadd_rn(__dmul_rn(ddx[k], ddx[k]), __dmul_rn(ddy[k], ddy[k]));
                    if (sq <= S_THR_04) {
                        reach_mask |= 1u << k;
                        if (e.current_target < 1) {
                            e.current_target += 1;   // the observation of THIS step is already relative to the new target
                            obs_done_mask |= 1u << k;
                            if (obs) mobile_store_obs<KIND>(e, obs + (size_t)t * N * D, (size_t)i);
                        }
                    }
                }
                if (e.counter > max_steps) {  // _termination (:336-343); `terminated` is never set
                    done_mask |= 1u << k;
                    if (auto_reset) {
                        mobile_reset_env<KIND>(e, nullptr, random_target, seed, genv);
                        targets_dirty = true;
                        obs_done_mask |= 1u << k;
                        if (obs) mobile_store_obs<KIND>(e, obs + (size_t)t * N * D, (size_t)i);   // post-reset observation
                    }
                }
            }
        }
        // ---------------- pass 2: rewards and outputs, independent across the chunk (:345-363) ----------------
        e.ep_ret = ep_ret0; e.ep_len = ep_len0;
#pragma unroll
        for (int k = 0; k < MOBILE_PF; ++k) {
            const int t = t0 + k;
            if (t < T) {
                const size_t off = (size_t)t * N + (size_t)i;
                const bool bumped = (bump_mask >> k) & 1u, is_done = (done_mask >> k) & 1u;
                double reward;
                if (shape_reward) {
                    double distance;
                    if (KIND == SRL_ENV_MOBILE_LINE_TARGET) distance = fabs(ddx[k]);
                    else if (KIND == SRL_ENV_MOBILE_1D) distance = __dsqrt_rn(__dmul_rn(ddx[k], ddx[k]));
                    else distance = __dsqrt_rn(__dadd_rn(__dmul_rn(ddx[k], ddx[k]), __dmul_rn(ddy[k], ddy[k])));  // np.linalg.norm = sqrt(x.dot(x))
                    reward = -distance;
                } else {
                    bool reached;
                    if (KIND == SRL_ENV_MOBILE_LINE_TARGET) reached = fabs(ddx[k]) <= LINE_REWARD_DIST_THRESHOLD;
                    else if (TWO) reached = (reach_mask >> k) & 1u;
                    else if (KIND == SRL_ENV_MOBILE_1D) reached = __dmul_rn(ddx[k], ddx[k]) <= S_THR_04;
                    else reached = __dadd_rn(__dmul_rn(ddx[k], ddx[k]), __dmul_rn(ddy[k], ddy[k])) <= S_THR_04;
                    reward = reached ? 1.0 : 0.0;
                    if (bumped) reward = -1.0;
                }
                e.ep_ret = __dadd_rn(e.ep_ret, reward);
                e.ep_len += 1.0;
                if (rew) rew[off] = (float)reward;
                if (done) done[off] = is_done ? 1 : 0;
                if (is_done) {
                    if (ep_ret) ep_ret[off] = (float)e.ep_ret;
                    if (ep_len) ep_len[off] = (int32_t)e.ep_len;
                    if (auto_reset) { e.ep_ret = 0.0; e.ep_len = 0.0; }
                }
                if (obs && !((obs_done_mask >> k) & 1u)) {
                    // getSRLState = getGroundTruth() - getTargetPos() = -(target - position): negation is exact
                    if (KIND == SRL_ENV_MOBILE_1D) obs[(size_t)t * N + (size_t)i] = -(float)ddx[k];
                    else reinterpret_cast<float2*>(obs + (size_t)t * N * D)[i] = make_float2(-(float)ddx[k], -(float)ddy[k]);
                }
            }
        }
        cur = nxt;
    }
    mobile_store(m, i, e, targets_dirty, TWO);
}

template <int KIND>
int launch_rollout_kind(srl_sim* s, int T, const void* actions, const float* noise, float* obs, float* rew, uint8_t* done,
                        float* ep_ret, int32_t* ep_len, cudaStream_t st) {
    // one warp per CTA below ~19k envs: spreads the (latency-bound) warps over all 148 SMs
    const int block = s->n <= 148 * 128 ? 32 : 64;
    const int grid = (s->n + block - 1) / block;
    if (s->cfg.is_discrete)
        mobile_rollout_kernel<KIND, true><<<grid, block, 0, st>>>(s->mob, s->n, T, actions, noise, obs, rew, done, ep_ret, ep_len,
                                                                   s->cfg.random_target != 0, s->cfg.shape_reward != 0,
                                                                   s->auto_reset != 0, s->max_steps, s->seed, s->cfg.global_env_offset);
    else
        mobile_rollout_kernel<KIND, false><<<grid, block, 0, st>>>(s->mob, s->n, T, actions, noise, obs, rew, done, ep_ret, ep_len,
                                                                    s->cfg.random_target != 0, s->cfg.shape_reward != 0,
                                                                    s->auto_reset != 0, s->max_steps, s->seed, s->cfg.global_env_offset);
    SRL_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

int mobile_alloc(srl_sim* s) {
    const size_t N = (size_t)s->n;
    MobileDev& m = s->mob;
    SRL_CUDA_OK(cudaMalloc(&m.pos, N * sizeof(double2)));
    SRL_CUDA_OK(cudaMalloc(&m.tgt0, N * sizeof(double2)));
    SRL_CUDA_OK(cudaMalloc(&m.tgt1, N * sizeof(double2)));
    SRL_CUDA_OK(cudaMalloc(&m.meta, N * sizeof(int4)));
    SRL_CUDA_OK(cudaMalloc(&m.ep, N * sizeof(double2)));
    SRL_CUDA_OK(cudaMemset(m.pos, 0, N * sizeof(double2)));
    SRL_CUDA_OK(cudaMemset(m.tgt0, 0, N * sizeof(double2)));
    SRL_CUDA_OK(cudaMemset(m.tgt1, 0, N * sizeof(double2)));
    SRL_CUDA_OK(cudaMemset(m.meta, 0, N * sizeof(int4)));
    SRL_CUDA_OK(cudaMemset(m.ep, 0, N * sizeof(double2)));
    return 0;
}

void mobile_free(srl_sim* s) {
    MobileDev& m = s->mob;
    cudaFree(m.pos); cudaFree(m.tgt0); cudaFree(m.tgt1); cudaFree(m.meta); cudaFree(m.ep);
    m = MobileDev{};
}

int mobile_launch_reset(srl_sim* s, const uint8_t* mask, const double* draws, float* obs, cudaStream_t st) {
    const int block = 128, grid = (s->n + block - 1) / block;
    const bool rt = s->cfg.random_target != 0;
    switch (s->kind) {
    case SRL_ENV_MOBILE: mobile_reset_kernel<SRL_ENV_MOBILE><<<grid, block, 0, st>>>(s->mob, s->n, mask, draws, obs, rt, s->seed, s->cfg.global_env_offset); break;
    case SRL_ENV_MOBILE_2TARGET: mobile_reset_kernel<SRL_ENV_MOBILE_2TARGET><<<grid, block, 0, st>>>(s->mob, s->n, mask, draws, obs, rt, s->seed, s->cfg.global_env_offset); break;
    case SRL_ENV_MOBILE_1D: mobile_reset_kernel<SRL_ENV_MOBILE_1D><<<grid, block, 0, st>>>(s->mob, s->n, mask, draws, obs, rt, s->seed, s->cfg.global_env_offset); break;
    default: mobile_reset_kernel<SRL_ENV_MOBILE_LINE_TARGET><<<grid, block, 0, st>>>(s->mob, s->n, mask, draws, obs, rt, s->seed, s->cfg.global_env_offset); break;
    }
    SRL_CUDA_OK(cudaGetLastError());
    return 0;
}

int mobile_launch_rollout(srl_sim* s, int T, const void* actions, const float* noise, float* obs, float* rew, uint8_t* done,
                          float* ep_ret, int32_t* ep_len, cudaStream_t st) {
    switch (s->kind) {
    case SRL_ENV_MOBILE: return launch_rollout_kind<SRL_ENV_MOBILE>(s, T, actions, noise, obs, rew, done, ep_ret, ep_len, st);
    case SRL_ENV_MOBILE_2TARGET: return launch_rollout_kind<SRL_ENV_MOBILE_2TARGET>(s, T, actions, noise, obs, rew, done, ep_ret, ep_len, st);
    case SRL_ENV_MOBILE_1D: return launch_rollout_kind<SRL_ENV_MOBILE_1D>(s, T, actions, noise, obs, rew, done, ep_ret, ep_len, st);
    default: return launch_rollout_kind<SRL_ENV_MOBILE_LINE_TARGET>(s, T, actions, noise, obs, rew, done, ep_ret, ep_len, st);
    }
}

// ---- host-side state access (debug / single-env accessors; not on the hot path) -------------
#include <vector>

int mobile_get_state(srl_sim* s, int field, void* dst, size_t bytes) {
    const size_t N = (size_t)s->n;
    std::vector<double2> a(N), b(N);
    std::vector<int4> mt(N);
    SRL_CUDA_OK(cudaDeviceSynchronize());
    SRL_CUDA_OK(cudaMemcpy(mt.data(), s->mob.meta, N * sizeof(int4), cudaMemcpyDeviceToHost));
    switch (field) {
    case SRL_F_ROBOT_POS:
    case SRL_F_TARGET_POS: {
        if (bytes != N * 3 * sizeof(double)) { srl_set_error("get_state: size mismatch"); return 1; }
        double* o = (double*)dst;
        if (field == SRL_F_ROBOT_POS) {
            SRL_CUDA_OK(cudaMemcpy(a.data(), s->mob.pos, N * sizeof(double2), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < N; ++i) { o[3 * i] = a[i].x; o[3 * i + 1] = a[i].y; o[3 * i + 2] = 0.0; }
        } else {
            SRL_CUDA_OK(cudaMemcpy(a.data(), s->mob.tgt0, N * sizeof(double2), cudaMemcpyDeviceToHost));
            SRL_CUDA_OK(cudaMemcpy(b.data(), s->mob.tgt1, N * sizeof(double2), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < N; ++i) {
                const double2 t = (mt[i].y & 0xff) ? b[i] : a[i];
                o[3 * i] = t.x; o[3 * i + 1] = t.y; o[3 * i + 2] = 0.0;
            }
        }
        return 0;
    }
    case SRL_F_STEP_COUNTER:
        if (bytes != N * sizeof(int32_t)) { srl_set_error("get_state: size mismatch"); return 1; }
        for (size_t i = 0; i < N; ++i) ((int32_t*)dst)[i] = mt[i].x;
        return 0;
    case SRL_F_COUNTERS:
        if (bytes != N * 4 * sizeof(int32_t)) { srl_set_error("get_state: size mismatch"); return 1; }
        for (size_t i = 0; i < N; ++i) {
            int32_t* o = (int32_t*)dst + 4 * i;
            o[0] = mt[i].y & 0xff; o[1] = (mt[i].y >> 8) & 1; o[2] = 0; o[3] = mt[i].z;
        }
        return 0;
    case SRL_F_EPISODE_STATS:
        if (bytes != N * 2 * sizeof(double)) { srl_set_error("get_state: size mismatch"); return 1; }
        SRL_CUDA_OK(cudaMemcpy(dst, s->mob.ep, N * sizeof(double2), cudaMemcpyDeviceToHost));
        return 0;
    default:
        srl_set_error("get_state: field %d not available for MobileRobot", field);
        return 1;
    }
}

int mobile_set_state(srl_sim* s, int field, const void* src, size_t bytes) {
    const size_t N = (size_t)s->n;
    SRL_CUDA_OK(cudaDeviceSynchronize());
    switch (field) {
    case SRL_F_ROBOT_POS:
    case SRL_F_TARGET_POS: {
        if (bytes != N * 3 * sizeof(double)) { srl_set_error("set_state: size mismatch"); return 1; }
        const double* in = (const double*)src;
        std::vector<double2> a(N);
        for (size_t i = 0; i < N; ++i) a[i] = make_double2(in[3 * i], in[3 * i + 1]);
        if (field == SRL_F_ROBOT_POS) {
            SRL_CUDA_OK(cudaMemcpy(s->mob.pos, a.data(), N * sizeof(double2), cudaMemcpyHostToDevice));
        } else {
            std::vector<int4> mt(N);
            std::vector<double2> t0(N), t1(N);
            SRL_CUDA_OK(cudaMemcpy(mt.data(), s->mob.meta, N * sizeof(int4), cudaMemcpyDeviceToHost));
            SRL_CUDA_OK(cudaMemcpy(t0.data(), s->mob.tgt0, N * sizeof(double2), cudaMemcpyDeviceToHost));
            SRL_CUDA_OK(cudaMemcpy(t1.data(), s->mob.tgt1, N * sizeof(double2), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < N; ++i) ((mt[i].y & 0xff) ? t1[i] : t0[i]) = a[i];
            SRL_CUDA_OK(cudaMemcpy(s->mob.tgt0, t0.data(), N * sizeof(double2), cudaMemcpyHostToDevice));
            SRL_CUDA_OK(cudaMemcpy(s->mob.tgt1, t1.data(), N * sizeof(double2), cudaMemcpyHostToDevice));
        }
        return 0;
    }
    case SRL_F_STEP_COUNTER: {
        if (bytes != N * sizeof(int32_t)) { srl_set_error("set_state: size mismatch"); return 1; }
        std::vector<int4> mt(N);
        SRL_CUDA_OK(cudaMemcpy(mt.data(), s->mob.meta, N * sizeof(int4), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < N; ++i) mt[i].x = ((const int32_t*)src)[i];
        SRL_CUDA_OK(cudaMemcpy(s->mob.meta, mt.data(), N * sizeof(int4), cudaMemcpyHostToDevice));
        return 0;
    }
    default:
        srl_set_error("set_state: field %d not settable for MobileRobot", field);
        return 1;
    }
}
