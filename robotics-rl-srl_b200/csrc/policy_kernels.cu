// PPO2 consumer helpers (include/srl_policy.h): the per-step policy forward + sample and the VecNormalize observation filter as ONE
// launch each, so that a captured rollout is three launches per env step (policy, simulator, filter) instead of ~60 small torch
// kernels around the simulator's.  Per-env arithmetic lives in policy_core.h (shared with the CPU checker of the tests).
#include <cuda_runtime.h>
#include "common.cuh"
#include "policy_core.h"
#include "../../include/srl_policy.h"

namespace {

constexpr int POLICY_BLOCK = 64;     // 4096 envs -> 64 CTAs; one env per thread
constexpr int H = SRL_POLICY_HIDDEN;

struct PolicyArgs {
    srl_mlp_policy p;
    int n;
    const float* obs;
    unsigned long long* rng;
    unsigned long long env_offset;
    float* obs_buf; void* act_env; void* act_buf; float* logp; float* value;
};

// shared-memory layout (floats): the 16-byte aligned 64-wide rows first, then the small pieces, then the activation columns
__device__ __forceinline__ void stage(float* dst, const float* __restrict__ src, int count) {
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = __ldg(src + i);
}

__global__ void __launch_bounds__(POLICY_BLOCK) policy_act_kernel(const __grid_constant__ PolicyArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int D = a.p.obs_dim, A = a.p.n_out;
    float* pi_w2 = smem;            float* vf_w2 = pi_w2 + H * H;
    float* pi_w3 = vf_w2 + H * H;   float* vf_w3 = pi_w3 + SRL_POLICY_MAX_OUT * H;
    float* pi_w1 = vf_w3 + H;       float* vf_w1 = pi_w1 + H * SRL_POLICY_MAX_OBS;
    float* pi_b1 = vf_w1 + H * SRL_POLICY_MAX_OBS; float* vf_b1 = pi_b1 + H;
    float* pi_b2 = vf_b1 + H;       float* vf_b2 = pi_b2 + H;
    float* pi_b3 = vf_b2 + H;       float* vf_b3 = pi_b3 + SRL_POLICY_MAX_OUT;
    float* s_logstd = vf_b3 + 4;    float* cols = s_logstd + SRL_POLICY_MAX_OUT;    // cols: H x POLICY_BLOCK, 16-byte aligned by construction
    stage(pi_w2, a.p.pi_w2, H * H); stage(vf_w2, a.p.vf_w2, H * H);
    stage(pi_w3, a.p.pi_w3, A * H); stage(vf_w3, a.p.vf_w3, H);
    stage(pi_w1, a.p.pi_w1, H * D); stage(vf_w1, a.p.vf_w1, H * D);
    stage(pi_b1, a.p.pi_b1, H); stage(vf_b1, a.p.vf_b1, H); stage(pi_b2, a.p.pi_b2, H); stage(vf_b2, a.p.vf_b2, H);
    stage(pi_b3, a.p.pi_b3, A); stage(vf_b3, a.p.vf_b3, 1);
    if (!a.p.discrete) stage(s_logstd, a.p.logstd, A);
    const unsigned long long seed = a.rng[0], counter = a.rng[1];    // read before this CTA arrives: the counter moves only after ALL CTAs arrived
    __syncthreads();
    const int i = blockIdx.x * POLICY_BLOCK + threadIdx.x;
    if (i < a.n) {
        float x[SRL_POLICY_MAX_OBS];
        for (int d = 0; d < D; ++d) {
            x[d] = a.obs[(size_t)i * D + d];
            if (a.obs_buf) a.obs_buf[(size_t)i * D + d] = x[d];
        }
        float* col = cols + threadIdx.x;
        float out[SRL_POLICY_MAX_OUT], v[1];
        const SrlTowerWeights Wpi = {pi_w1, pi_b1, pi_w2, pi_b2, pi_w3, pi_b3};
        const SrlTowerWeights Wvf = {vf_w1, vf_b1, vf_w2, vf_b2, vf_w3, vf_b3};
        srl_mlp_tower(Wpi, D, A, x, col, POLICY_BLOCK, out);
        srl_mlp_tower(Wvf, D, 1, x, col, POLICY_BLOCK, v);
        a.value[i] = v[0];
        const unsigned long long env = a.env_offset + (unsigned long long)i;
        float lp;
        if (a.p.discrete) {
            const int act = srl_sample_categorical(out, A, seed, env, (uint32_t)counter, &lp);
            reinterpret_cast<int32_t*>(a.act_env)[i] = act;
            if (a.act_buf) reinterpret_cast<long long*>(a.act_buf)[i] = (long long)act;
        } else {
            float smp[SRL_POLICY_MAX_OUT], clp[SRL_POLICY_MAX_OUT];
            srl_sample_gaussian(out, s_logstd, A, seed, env, (uint32_t)counter, smp, clp, &lp);
            for (int k = 0; k < A; ++k) {
                reinterpret_cast<float*>(a.act_env)[(size_t)i * A + k] = clp[k];
                if (a.act_buf) reinterpret_cast<float*>(a.act_buf)[(size_t)i * A + k] = smp[k];
            }
        }
        a.logp[i] = lp;
    }
    // the LAST CTA to retire advances the step counter: every CTA has read it by then, and the next launch sees the new value
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long arrived = atomicAdd(a.rng + 2, 1ull);
        if (arrived == (unsigned long long)gridDim.x - 1ull) {
            a.rng[2] = 0ull;
            a.rng[1] = counter + 1ull;
            __threadfence();
        }
    }
}

constexpr int FILTER_BLOCK = 1024;

// One CTA: batch mean and (two-pass, biased) variance in float64, Chan's parallel-variance merge into the running state
// (the update of stable-baselines' RunningMeanStd), then the normalisation of the whole batch in float32.
__global__ void __launch_bounds__(FILTER_BLOCK) obs_filter_kernel(int n, int D, const float* __restrict__ obs, double* state, int update,
                                                                   float clip, float eps, float* __restrict__ out) {
    __shared__ double red[32][SRL_POLICY_MAX_OBS];
    __shared__ double s_mean[SRL_POLICY_MAX_OBS], s_bm[SRL_POLICY_MAX_OBS];
    __shared__ float s_mf[SRL_POLICY_MAX_OBS], s_inv[SRL_POLICY_MAX_OBS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto block_sum = [&](double (&acc)[SRL_POLICY_MAX_OBS], double* result) {   // result[d] valid in every thread after the call
        for (int d = 0; d < D; ++d) {
            double v = acc[d];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][d] = v;
        }
        __syncthreads();
        if (warp == 0) {
            for (int d = 0; d < D; ++d) {
                double v = red[lane][d];
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) result[d] = v;
            }
        }
        __syncthreads();
    };
    if (update) {
        double acc[SRL_POLICY_MAX_OBS];
        for (int d = 0; d < D; ++d) acc[d] = 0.0;
        for (int i = tid; i < n; i += FILTER_BLOCK)
            for (int d = 0; d < D; ++d) acc[d] += (double)obs[(size_t)i * D + d];
        block_sum(acc, s_bm);
        double bm[SRL_POLICY_MAX_OBS];
        for (int d = 0; d < D; ++d) { bm[d] = s_bm[d] / (double)n; acc[d] = 0.0; }
        for (int i = tid; i < n; i += FILTER_BLOCK)
            for (int d = 0; d < D; ++d) { const double c = (double)obs[(size_t)i * D + d] - bm[d]; acc[d] += c * c; }
        __syncthreads();                         // every thread has read s_bm before block_sum overwrites the scratch it shares
        block_sum(acc, s_mean);                  // s_mean temporarily holds the sums of squared deviations
        if (tid == 0) {
            const double count = state[2 * D], bc = (double)n, tot = count + bc;
            for (int d = 0; d < D; ++d) {
                const double bmean = s_bm[d] / bc, bvar = s_mean[d] / bc;
                const double mean = state[d], var = state[D + d], delta = bmean - mean;
                state[d] = mean + delta * bc / tot;
                state[D + d] = (var * count + bvar * bc + delta * delta * count * bc / tot) / tot;
            }
            state[2 * D] = tot;
        }
        __syncthreads();
    }
    if (tid < D) {
        s_mf[tid] = (float)state[tid];
        s_inv[tid] = sqrtf((float)state[D + tid] + eps);
    }
    __syncthreads();
    const int total = n * D;
    for (int e = tid; e < total; e += FILTER_BLOCK) {
        const int d = e % D;
        const float v = (obs[e] - s_mf[d]) / s_inv[d];
        out[e] = fminf(fmaxf(v, -clip), clip);
    }
}

constexpr size_t policy_smem_bytes() {
    return sizeof(float) * (size_t)(2 * H * H + SRL_POLICY_MAX_OUT * H + H + 2 * H * SRL_POLICY_MAX_OBS + 4 * H + SRL_POLICY_MAX_OUT + 4 +
                                    SRL_POLICY_MAX_OUT + H * POLICY_BLOCK);
}

}  // namespace

extern "C" {

int srl_policy_act(const srl_mlp_policy* p, int n, const float* obs, uint64_t* rng, uint64_t env_offset, float* obs_buf,
                   void* act_env, void* act_buf, float* logp, float* value, void* stream) {
    if (!p || !obs || !rng || !act_env || !logp || !value) { srl_set_error("policy_act: null argument"); return 1; }
    if (p->struct_size != sizeof(srl_mlp_policy)) { srl_set_error("policy_act: srl_mlp_policy size mismatch (%u != %zu)", p->struct_size, sizeof(srl_mlp_policy)); return 1; }
    if (n <= 0) { srl_set_error("policy_act: n must be positive"); return 1; }
    if (p->obs_dim < 1 || p->obs_dim > SRL_POLICY_MAX_OBS || p->n_out < 1 || p->n_out > SRL_POLICY_MAX_OUT || (p->discrete && p->n_out < 2)) {
        srl_set_error("policy_act: unsupported shape obs_dim=%d n_out=%d", p->obs_dim, p->n_out); return 1;
    }
    if (!p->pi_w1 || !p->pi_b1 || !p->pi_w2 || !p->pi_b2 || !p->pi_w3 || !p->pi_b3 || !p->vf_w1 || !p->vf_b1 || !p->vf_w2 || !p->vf_b2 ||
        !p->vf_w3 || !p->vf_b3 || (!p->discrete && !p->logstd)) { srl_set_error("policy_act: null weight pointer"); return 1; }
    static bool attr_set[64] = {};     // > 48 KB of dynamic shared memory needs the opt-in once per device context
    constexpr size_t smem = policy_smem_bytes();
    int dev = 0;
    SRL_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        SRL_CUDA_OK(cudaFuncSetAttribute(policy_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    PolicyArgs a;
    a.p = *p; a.n = n; a.obs = obs; a.rng = reinterpret_cast<unsigned long long*>(rng); a.env_offset = env_offset;
    a.obs_buf = obs_buf; a.act_env = act_env; a.act_buf = act_buf; a.logp = logp; a.value = value;
    policy_act_kernel<<<(n + POLICY_BLOCK - 1) / POLICY_BLOCK, POLICY_BLOCK, smem, (cudaStream_t)stream>>>(a);
    SRL_CUDA_OK(cudaGetLastError());
    return 0;
}

int srl_obs_filter(int n, int obs_dim, const float* obs_raw, double* state, int update, float clip, float eps, float* obs_norm_out,
                   void* stream) {
    if (!obs_raw || !state || !obs_norm_out) { srl_set_error("obs_filter: null argument"); return 1; }
    if (n <= 0 || obs_dim < 1 || obs_dim > SRL_POLICY_MAX_OBS) { srl_set_error("obs_filter: unsupported shape n=%d obs_dim=%d", n, obs_dim); return 1; }
    obs_filter_kernel<<<1, FILTER_BLOCK, 0, (cudaStream_t)stream>>>(n, obs_dim, obs_raw, state, update, clip, eps, obs_norm_out);
    SRL_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
