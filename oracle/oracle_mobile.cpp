/*
 * ORACLE (test infrastructure, not product code).
 *
 * MobileRobot family, restated line by line in double precision from the reference Python.
 * The racecar is a fixed-base body that is teleported every step
 * (environments/mobile_robot/mobile_robot_env.py:207-208,265); p.stepSimulation() (:267) has no
 * effect on any `ground_truth` observable, so the restatement is the kinematic update itself.
 *
 *   reset  : mobile_robot_env.py:159-222      (2target: mobile_robot_2target_env.py:29-115,
 *            1D: mobile_robot_1D_env.py:56-100, line target: mobile_robot_line_target_env.py:42-106)
 *   step   : mobile_robot_env.py:235-280      (2target :117-162, 1D :108-147)
 *   reward : mobile_robot_env.py:345-363      (2target :164-185, 1D :149-168, line :108-125)
 *   done   : mobile_robot_env.py:336-343  -- `terminated` is never set, so done <=> counter > 250
 *
 * Compile with -ffp-contract=off: the reference is numpy float64 without fused multiply-add.
 */
#include <math.h>
#include <string.h>
#include "oracle_sim.h"
#include "philox.h"

/* module constants, mobile_robot_env.py:13-28,101-104 */
static const double MAX_X = 4.0, MAX_Y = 4.0, MIN_X = 0.0, MIN_Y = 0.0;
static const double DELTA_POS = 0.1;
static const double ROBOT_WIDTH = 0.2, ROBOT_LENGTH = 0.325 * 2;
static const double COLLISION_MARGIN = 0.1;
static const double REWARD_DIST_THRESHOLD = 0.4;
static const double LINE_REWARD_DIST_THRESHOLD = 0.1, LINE_ROBOT_OFFSET = 0.2; /* line_target_env.py:3-4 */
static const int MOBILE_MAX_STEPS = 250; /* mobile_robot_env.py:13,95 (2target declares 1500 but never applies it) */

int oracle_mobile_obs_dim(int kind) { return kind == SRL_ENV_MOBILE_1D ? 1 : 2; }

int oracle_mobile_action_dim(const srl_sim* s) {
    if (s->cfg.is_discrete) return 1;
    return 2;
}

static int mobile_num_actions(int kind) { return kind == SRL_ENV_MOBILE_1D ? 2 : 4; }

/* numpy RandomState.uniform(low, high) = low + (high - low) * random_sample() */
static double uniform(double low, double high, double u) { return low + (high - low) * u; }

void oracle_mobile_reset_env(srl_sim* s, int i, const double* draws) {
    MobileEnv& e = s->mobile[i];
    const int kind = s->kind;
    double d[6];
    if (draws) {
        memcpy(d, draws, sizeof(d));
    } else {
        const uint64_t genv = s->cfg.global_env_offset + (uint64_t)i;
        uint32_t r[4];
        philox4x32_10(s->seed, genv, e.episode, PHILOX_PURPOSE_RESET0 + 0, r);
        /* mobile_robot_env.py:168-169 */
        d[0] = MAX_X / 2 + uniform(-MAX_X / 3, MAX_X / 3, philox_u01(r[0], r[1]));
        d[1] = MAX_Y / 2 + uniform(-MAX_Y / 3, MAX_Y / 3, philox_u01(r[2], r[3]));
        const double margin = 0.1 * MAX_X; /* :176 */
        philox4x32_10(s->seed, genv, e.episode, PHILOX_PURPOSE_RESET0 + 1, r);
        d[2] = uniform(MIN_X + margin, MAX_X - margin, philox_u01(r[0], r[1]));
        d[3] = uniform(MIN_Y + margin, MAX_Y - margin, philox_u01(r[2], r[3]));
        philox4x32_10(s->seed, genv, e.episode, PHILOX_PURPOSE_RESET0 + 2, r);
        d[4] = uniform(MIN_X + margin, MAX_X - margin, philox_u01(r[0], r[1]));
        d[5] = uniform(MIN_Y + margin, MAX_Y - margin, philox_u01(r[2], r[3]));
    }
    e.pos[0] = d[0];
    e.pos[1] = (kind == SRL_ENV_MOBILE_1D) ? 0.0 : d[1]; /* 1D_env.py:66 */
    e.pos[2] = 0.0;
    /* fixed targets: mobile_robot_env.py:173-174, 2target_env.py:52-53,62-63,
       1D_env.py:69, line_target_env.py:56-57 */
    double t0x = 0.9 * MAX_X, t0y = MAX_Y * 3 / 4;
    double t1x = 0.1 * MAX_X, t1y = MAX_Y * 3 / 4;
    if (kind == SRL_ENV_MOBILE_1D) t0y = 0.0;
    if (kind == SRL_ENV_MOBILE_LINE_TARGET) t0y = MAX_X;
    if (s->cfg.random_target) {
        t0x = d[2];
        if (kind == SRL_ENV_MOBILE || kind == SRL_ENV_MOBILE_2TARGET) t0y = d[3];
        if (kind == SRL_ENV_MOBILE_2TARGET) { t1x = d[4]; t1y = d[5]; }
    }
    e.target[0][0] = t0x; e.target[0][1] = t0y; e.target[0][2] = 0.0;
    e.target[1][0] = t1x; e.target[1][1] = t1y; e.target[1][2] = 0.0;
    e.current_target = 0;
    e.counter = 0;
    e.has_bumped = 0;
    e.ep_ret = 0.0;
    e.ep_len = 0;
    e.episode += 1;
}

/* getSRLState = getGroundTruth() - getTargetPos()  (srl_env.py:39-42, RELATIVE_POS = True) */
void oracle_mobile_obs(const srl_sim* s, int i, float* obs) {
    const MobileEnv& e = s->mobile[i];
    const double* t = e.target[e.current_target];
    switch (s->kind) {
    case SRL_ENV_MOBILE_1D: /* 1D_env.py:38-49 */
        obs[0] = (float)(e.pos[0] - t[0]);
        break;
    case SRL_ENV_MOBILE_LINE_TARGET: { /* line_target_env.py:35-40: target is the 1-vector [x - 0.2], broadcast */
        const double tx = t[0] - LINE_ROBOT_OFFSET;
        obs[0] = (float)(e.pos[0] - tx);
        obs[1] = (float)(e.pos[1] - tx);
        break;
    }
    default:
        obs[0] = (float)(e.pos[0] - t[0]);
        obs[1] = (float)(e.pos[1] - t[1]);
    }
}

void oracle_mobile_step_env(srl_sim* s, int i, const void* actions, const float* noise, float* obs,
                            float* rew, uint8_t* done, float* ep_ret, int32_t* ep_len) {
    MobileEnv& e = s->mobile[i];
    const int kind = s->kind;
    const uint64_t genv = s->cfg.global_env_offset + (uint64_t)i;
    e.has_bumped = 0; /* :237 */
    /* dv = DELTA_POS + np_random.normal(0.0, scale=NOISE_STD), NOISE_STD = 0.0  (:239-241) */
    double dv = DELTA_POS + (noise ? (double)noise[i] : 0.0);
    double real_action[2] = {0.0, 0.0};
    if (s->cfg.is_discrete) {
        int a;
        if (actions) {
            a = ((const int32_t*)actions)[i];
        } else {
            uint32_t r[4];
            philox4x32_10(s->seed, genv, e.total_steps, PHILOX_PURPOSE_ACTION, r);
            a = (int)(((uint64_t)r[0] * (uint64_t)mobile_num_actions(kind)) >> 32);
        }
        if (kind == SRL_ENV_MOBILE_1D) {
            const double dx[2] = {-dv, dv}; /* 1D_env.py:115 */
            real_action[0] = dx[a & 1];
        } else {
            const double dx[4] = {-dv, dv, 0, 0}, dy[4] = {0, 0, -dv, dv}; /* :242-243 */
            real_action[0] = dx[a & 3];
            real_action[1] = dy[a & 3];
        }
    } else {
        /* np.maximum(np.minimum(action, 1), -1) * dv with a float32 action array: the product is
           evaluated in float32 (numpy scalar promotion), then added to the float64 position (:250,255) */
        float a[2];
        if (actions) {
            a[0] = ((const float*)actions)[2 * i];
            a[1] = ((const float*)actions)[2 * i + 1];
        } else {
            uint32_t r[4];
            philox4x32_10(s->seed, genv, e.total_steps, PHILOX_PURPOSE_ACTION, r);
            a[0] = (float)((double)r[0] * (2.0 / 4294967296.0) - 1.0);
            a[1] = (float)((double)r[1] * (2.0 / 4294967296.0) - 1.0);
        }
        for (int k = 0; k < 2; ++k) {
            float c = fminf(a[k], 1.0f);
            c = fmaxf(c, -1.0f);
            real_action[k] = (double)(c * (float)dv);
        }
    }
    e.total_steps += 1;

    double prev[3] = {e.pos[0], e.pos[1], e.pos[2]}; /* :254 */
    e.pos[0] += real_action[0];
    if (kind != SRL_ENV_MOBILE_1D) e.pos[1] += real_action[1]; /* 1D: robot_pos[:1] += (1D_env.py:124) */
    /* Handle collisions (:257-263): x uses ROBOT_LENGTH, y uses ROBOT_WIDTH; first violated axis reverts all */
    {
        const double limit[2] = {MAX_X, MAX_Y}, dim[2] = {ROBOT_LENGTH, ROBOT_WIDTH};
        const int naxes = (kind == SRL_ENV_MOBILE_1D) ? 1 : 2;
        for (int k = 0; k < naxes; ++k) {
            const double margin = COLLISION_MARGIN + dim[k] / 2;
            if (e.pos[k] < margin || e.pos[k] > limit[k] - margin) {
                e.has_bumped = 1;
                e.pos[0] = prev[0]; e.pos[1] = prev[1]; e.pos[2] = prev[2];
                break;
            }
        }
    }
    e.counter += 1; /* :268 */

    /* _reward (:345-363) */
    double distance;
    const double* t = e.target[e.current_target];
    double thr = REWARD_DIST_THRESHOLD;
    if (kind == SRL_ENV_MOBILE_LINE_TARGET) {
        distance = fabs((t[0] - LINE_ROBOT_OFFSET) - e.pos[0]); /* line_target_env.py:113 */
        thr = LINE_REWARD_DIST_THRESHOLD;
    } else if (kind == SRL_ENV_MOBILE_1D) {
        const double dx = t[0] - e.pos[0];
        distance = sqrt(dx * dx); /* np.linalg.norm of a 1-vector: sqrt(x.dot(x)) */
    } else {
        const double dx = t[0] - e.pos[0], dy = t[1] - e.pos[1];
        distance = sqrt(dx * dx + dy * dy); /* np.linalg.norm(., 2) = sqrt(x.dot(x)) */
    }
    double reward = 0;
    if (distance <= thr) {
        reward = 1;
        if (kind == SRL_ENV_MOBILE_2TARGET && e.current_target < 1) e.current_target += 1; /* 2target_env.py:172-173 */
    }
    if (e.has_bumped) reward = -1;
    if (s->cfg.shape_reward) reward = -distance;

    const int max_steps = s->cfg.max_steps > 0 ? s->cfg.max_steps : MOBILE_MAX_STEPS;
    const int is_done = e.counter > max_steps; /* _termination (:336-343) */

    e.ep_ret += reward;
    e.ep_len += 1;
    if (rew) rew[i] = (float)reward;
    if (done) done[i] = (uint8_t)is_done;
    if (is_done) {
        if (ep_ret) ep_ret[i] = (float)e.ep_ret;
        if (ep_len) ep_len[i] = e.ep_len;
        if (s->auto_reset) oracle_mobile_reset_env(s, i, NULL);
    }
    if (obs) oracle_mobile_obs(s, i, obs + (size_t)i * oracle_mobile_obs_dim(kind));
}
