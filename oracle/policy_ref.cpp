/*
 * ORACLE (test infrastructure, not product code).
 *
 * CPU checker of the PPO2 consumer helpers declared in include/srl_policy.h.  The per-env arithmetic is
 * robotics-rl-srl_b200/csrc/policy_core.h compiled for the host; the tests pin it against an independent implementation --
 * torch's nn.Linear / tanh / log_softmax / Normal on the same weights (tests/test_policy_cpu.py) -- and the GPU tests then hold the
 * sm_100a kernels to it.  The filter is restated from stable-baselines 2.5 `RunningMeanStd.update_from_moments` +
 * `VecNormalize._obfilt` (the wrapper rl_baselines/utils.py:224-227 of the reference puts around its envs).
 * Only tests/ load this library.
 */
#include <math.h>
#include <stdint.h>
#include <vector>
#include "../include/srl_policy.h"
#include "../robotics-rl-srl_b200/csrc/policy_core.h"

extern "C" {

/* host pointers everywhere; `counter` is the value the device launch would read from rng[1] */
int policy_ref_act(const srl_mlp_policy* p, int n, const float* obs, uint64_t seed, uint64_t counter, uint64_t env_offset, void* act_env,
                   void* act_buf, float* logp, float* value, float* logits_out /* nullable [n, n_out]: policy tower output */) {
    if (!p || p->struct_size != sizeof(srl_mlp_policy)) return 1;
    const int D = p->obs_dim, A = p->n_out;
    const SrlTowerWeights Wpi = {p->pi_w1, p->pi_b1, p->pi_w2, p->pi_b2, p->pi_w3, p->pi_b3};
    const SrlTowerWeights Wvf = {p->vf_w1, p->vf_b1, p->vf_w2, p->vf_b2, p->vf_w3, p->vf_b3};
    std::vector<float> col(SRL_POLICY_HIDDEN);
    for (int i = 0; i < n; ++i) {
        float out[SRL_POLICY_MAX_OUT], v[1], lp;
        srl_mlp_tower(Wpi, D, A, obs + (size_t)i * D, col.data(), 1, out);
        srl_mlp_tower(Wvf, D, 1, obs + (size_t)i * D, col.data(), 1, v);
        value[i] = v[0];
        if (logits_out) for (int k = 0; k < A; ++k) logits_out[(size_t)i * A + k] = out[k];
        if (p->discrete) {
            const int a = srl_sample_categorical(out, A, seed, env_offset + (uint64_t)i, (uint32_t)counter, &lp);
            ((int32_t*)act_env)[i] = a;
            if (act_buf) ((int64_t*)act_buf)[i] = a;
        } else {
            float smp[SRL_POLICY_MAX_OUT], clp[SRL_POLICY_MAX_OUT];
            srl_sample_gaussian(out, p->logstd, A, seed, env_offset + (uint64_t)i, (uint32_t)counter, smp, clp, &lp);
            for (int k = 0; k < A; ++k) {
                ((float*)act_env)[(size_t)i * A + k] = clp[k];
                if (act_buf) ((float*)act_buf)[(size_t)i * A + k] = smp[k];
            }
        }
        logp[i] = lp;
    }
    return 0;
}

int policy_ref_filter(int n, int D, const float* obs, double* state, int update, float clip, float eps, float* out) {
    if (update) {
        const double count = state[2 * D], bc = (double)n, tot = count + bc;
        for (int d = 0; d < D; ++d) {
            double sum = 0.0, sq = 0.0;
            for (int i = 0; i < n; ++i) sum += (double)obs[(size_t)i * D + d];
            const double bmean = sum / bc;
            for (int i = 0; i < n; ++i) { const double c = (double)obs[(size_t)i * D + d] - bmean; sq += c * c; }
            const double bvar = sq / bc, mean = state[d], var = state[D + d], delta = bmean - mean;
            state[d] = mean + delta * bc / tot;
            state[D + d] = (var * count + bvar * bc + delta * delta * count * bc / tot) / tot;
        }
        state[2 * D] = tot;
    }
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < D; ++d) {
            const float v = (obs[(size_t)i * D + d] - (float)state[d]) / sqrtf((float)state[D + d] + eps);
            out[(size_t)i * D + d] = fminf(fmaxf(v, -clip), clip);
        }
    return 0;
}

}  // extern "C"
