// C-ABI of include/srl_sim.h for the sm_100a library (libsrl_sim_b200.so).
// Plain pointers and sizes only; torch never appears in a signature.  All buffers are DEVICE
// pointers except in srl_sim_rollout_host.  There is no CPU path in this library.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <new>
#include "common.cuh"

static thread_local char g_err[512] = "";

void srl_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

int ensure_stage(srl_sim* s, int slot, size_t bytes) {
    if (s->stage_cap[slot] >= bytes) return 0;
    if (s->stage[slot]) cudaFree(s->stage[slot]);
    s->stage[slot] = nullptr;
    s->stage_cap[slot] = 0;
    SRL_CUDA_OK(cudaMalloc(&s->stage[slot], bytes));
    s->stage_cap[slot] = bytes;
    return 0;
}

int launch_rollout(srl_sim* s, int T, const void* actions, const float* noise, float* obs, float* rew, uint8_t* done,
                   float* ep_ret, int32_t* ep_len, cudaStream_t st) {
    SRL_CUDA_OK(cudaEventRecord(s->ev0, st));
    int rc = srl_is_mobile(s->kind) ? mobile_launch_rollout(s, T, actions, noise, obs, rew, done, ep_ret, ep_len, st)
                                    : kuka_launch_rollout(s, T, actions, noise, obs, rew, done, ep_ret, ep_len, st);
    if (rc) return rc;
    SRL_CUDA_OK(cudaEventRecord(s->ev1, st));
    s->ev_valid = true;
    s->launches += 1;
    return 0;
}

}  // namespace

extern "C" {

int srl_sim_abi_version(void) { return SRL_SIM_ABI_VERSION; }
const char* srl_sim_last_error(void) { return g_err; }

int srl_sim_create(srl_sim** out, int env_kind, int num_envs, int device, const srl_cfg* cfg, const void* model_blob,
                   size_t model_bytes, uint64_t seed) {
    if (!out || !cfg) { srl_set_error("create: null argument"); return 1; }
    *out = nullptr;
    if (cfg->struct_size != sizeof(srl_cfg)) {
        srl_set_error("create: srl_cfg size mismatch (%u != %zu)", cfg->struct_size, sizeof(srl_cfg));
        return 1;
    }
    if (device < 0) { srl_set_error("create: this library has no CPU path (device=%d)", device); return 1; }
    if (num_envs <= 0) { srl_set_error("create: num_envs must be positive"); return 1; }
    if (!srl_is_mobile(env_kind) && !srl_is_kuka(env_kind)) { srl_set_error("create: unknown env kind %d", env_kind); return 1; }
    if (cfg->action_joints && (!srl_is_kuka(env_kind) || cfg->is_discrete)) {
        // kuka_button_gym_env.py:149-161: the (7,) joint action space only exists for continuous Kuka actions
        srl_set_error("create: action_joints needs a Kuka env with is_discrete=0"); return 1;
    }
    if (srl_is_mobile(env_kind) && !cfg->is_discrete && env_kind != SRL_ENV_MOBILE && env_kind != SRL_ENV_MOBILE_LINE_TARGET) {
        // mobile_robot_2target_env.py:128, mobile_robot_1D_env.py:43,118 raise ValueError
        srl_set_error("Only discrete actions is supported");
        return 2;
    }
    int ndev = 0;
    SRL_CUDA_OK(cudaGetDeviceCount(&ndev));
    if (device >= ndev) { srl_set_error("create: device %d out of range (%d visible)", device, ndev); return 1; }
    DeviceGuard guard(device);
    if (!guard.ok) { srl_set_error("create: cudaSetDevice(%d) failed", device); return 1; }
    cudaDeviceProp prop;
    SRL_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        srl_set_error("create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return 1;
    }
    srl_sim* s = new (std::nothrow) srl_sim();
    if (!s) { srl_set_error("create: out of memory"); return 1; }
    memset(s, 0, sizeof(*s));
    s->kind = env_kind;
    s->n = num_envs;
    s->device = device;
    s->cfg = *cfg;
    if (s->cfg.action_repeat < 1) s->cfg.action_repeat = 1;
    if (s->cfg.solver_iterations <= 0) s->cfg.solver_iterations = 150;
    if (s->cfg.timestep <= 0.f) s->cfg.timestep = 1.0f / 240.0f;
    s->seed = seed;
    s->auto_reset = !cfg->no_auto_reset;
    s->max_steps = cfg->max_steps > 0 ? cfg->max_steps : (srl_is_mobile(env_kind) ? 250 : (env_kind == SRL_ENV_KUKA_MOVING_BUTTON || env_kind == SRL_ENV_KUKA_2BUTTON) ? 1500 : 1000);
    if (cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess) {
        srl_set_error("create: cudaEventCreate failed");
        delete s;
        return 1;
    }
    int rc = srl_is_mobile(env_kind) ? mobile_alloc(s) : kuka_alloc(s, model_blob, model_bytes);
    if (rc) { srl_sim_destroy(s); return rc; }
    *out = s;
    return 0;
}

void srl_sim_destroy(srl_sim* s) {
    if (!s) return;
    DeviceGuard guard(s->device);
    cudaDeviceSynchronize();
    if (srl_is_mobile(s->kind)) mobile_free(s); else kuka_free(s);
    for (int k = 0; k < 5; ++k) if (s->stage[k]) cudaFree(s->stage[k]);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    delete s;
}

int srl_sim_num_envs(const srl_sim* s) { return s ? s->n : 0; }
int srl_sim_obs_dim(const srl_sim* s) { return !s ? 0 : srl_is_kuka(s->kind) ? 3 : (s->kind == SRL_ENV_MOBILE_1D ? 1 : 2); }
int srl_sim_action_dim(const srl_sim* s) {
    if (!s) return 0;
    if (s->cfg.is_discrete) return 1;
    return srl_is_mobile(s->kind) ? 2 : (s->cfg.action_joints ? 7 : 3);
}
uint64_t srl_sim_launch_count(const srl_sim* s) { return s ? s->launches : 0; }

float srl_sim_last_kernel_ms(srl_sim* s) {
    if (!s || !s->ev_valid) return -1.0f;
    DeviceGuard guard(s->device);
    if (cudaEventSynchronize(s->ev1) != cudaSuccess) return -1.0f;
    float ms = -1.0f;
    if (cudaEventElapsedTime(&ms, s->ev0, s->ev1) != cudaSuccess) return -1.0f;
    return ms;
}

int srl_sim_reset(srl_sim* s, const uint8_t* mask, const double* reset_draws, float* obs_out, void* stream) {
    if (!s) { srl_set_error("reset: null handle"); return 1; }
    DeviceGuard guard(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = srl_is_mobile(s->kind) ? mobile_launch_reset(s, mask, reset_draws, obs_out, st)
                                    : kuka_launch_reset(s, mask, reset_draws, obs_out, st);
    if (!rc) s->launches += 1;
    return rc;
}

int srl_sim_step(srl_sim* s, const void* actions, const float* noise, float* obs_out, float* rew_out, uint8_t* done_out,
                 float* ep_ret_out, int32_t* ep_len_out, void* stream) {
    if (!s) { srl_set_error("step: null handle"); return 1; }
    if (!actions) { srl_set_error("step: actions must not be NULL (use srl_sim_rollout for in-kernel random actions)"); return 1; }
    DeviceGuard guard(s->device);
    return launch_rollout(s, 1, actions, noise, obs_out, rew_out, done_out, ep_ret_out, ep_len_out, (cudaStream_t)stream);
}

int srl_sim_rollout(srl_sim* s, int T, const void* actions, const float* noise, float* obs_out, float* rew_out,
                    uint8_t* done_out, float* ep_ret_out, int32_t* ep_len_out, void* stream) {
    if (!s) { srl_set_error("rollout: null handle"); return 1; }
    if (T < 0) { srl_set_error("rollout: negative T"); return 1; }
    if (T == 0) return 0;
    DeviceGuard guard(s->device);
    return launch_rollout(s, T, actions, noise, obs_out, rew_out, done_out, ep_ret_out, ep_len_out, (cudaStream_t)stream);
}

int srl_sim_rollout_host(srl_sim* s, int T, const void* actions, const float* noise, float* obs_out, float* rew_out,
                         uint8_t* done_out) {
    if (!s) { srl_set_error("rollout_host: null handle"); return 1; }
    if (T <= 0) { srl_set_error("rollout_host: T must be positive"); return 1; }
    DeviceGuard guard(s->device);
    const size_t N = (size_t)s->n, TN = (size_t)T * N;
    const size_t D = (size_t)srl_sim_obs_dim(s), A = (size_t)srl_sim_action_dim(s);
    const size_t act_bytes = TN * A * 4, noise_bytes = TN * 4, obs_bytes = TN * D * 4, rew_bytes = TN * 4, done_bytes = TN;
    cudaStream_t st = 0;
    void *d_act = nullptr, *d_noise = nullptr, *d_obs = nullptr, *d_rew = nullptr, *d_done = nullptr;
    if (actions) { if (ensure_stage(s, 0, act_bytes)) return 1; d_act = s->stage[0];
                   SRL_CUDA_OK(cudaMemcpyAsync(d_act, actions, act_bytes, cudaMemcpyHostToDevice, st)); }
    if (noise) { if (ensure_stage(s, 1, noise_bytes)) return 1; d_noise = s->stage[1];
                 SRL_CUDA_OK(cudaMemcpyAsync(d_noise, noise, noise_bytes, cudaMemcpyHostToDevice, st)); }
    if (obs_out) { if (ensure_stage(s, 2, obs_bytes)) return 1; d_obs = s->stage[2]; }
    if (rew_out) { if (ensure_stage(s, 3, rew_bytes)) return 1; d_rew = s->stage[3]; }
    if (done_out) { if (ensure_stage(s, 4, done_bytes)) return 1; d_done = s->stage[4]; }
    int rc = launch_rollout(s, T, d_act, (const float*)d_noise, (float*)d_obs, (float*)d_rew, (uint8_t*)d_done, nullptr, nullptr, st);
    if (rc) return rc;
    if (obs_out) SRL_CUDA_OK(cudaMemcpyAsync(obs_out, d_obs, obs_bytes, cudaMemcpyDeviceToHost, st));
    if (rew_out) SRL_CUDA_OK(cudaMemcpyAsync(rew_out, d_rew, rew_bytes, cudaMemcpyDeviceToHost, st));
    if (done_out) SRL_CUDA_OK(cudaMemcpyAsync(done_out, d_done, done_bytes, cudaMemcpyDeviceToHost, st));
    SRL_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int srl_sim_get_state(srl_sim* s, int field, void* dst, size_t bytes) {
    if (!s || !dst) { srl_set_error("get_state: null argument"); return 1; }
    DeviceGuard guard(s->device);
    return srl_is_mobile(s->kind) ? mobile_get_state(s, field, dst, bytes) : kuka_get_state(s, field, dst, bytes);
}

int srl_sim_set_state(srl_sim* s, int field, const void* src, size_t bytes) {
    if (!s || !src) { srl_set_error("set_state: null argument"); return 1; }
    DeviceGuard guard(s->device);
    return srl_is_mobile(s->kind) ? mobile_set_state(s, field, src, bytes) : kuka_set_state(s, field, src, bytes);
}

}  // extern "C"
