// PPO2 consumer helpers (include/srl_policy.h): the per-step policy forward + sample and the VecNormalize observation filter as ONE
// launch each, so that a captured rollout is three launches per env step (policy, simulator, filter) instead of ~60 small torch
// kernels around the simulator's.  Per-env arithmetic lives in policy_core.h (shared with the CPU checker of the tests).
#include <cuda_runtime.h>
#include "common.cuh"
#include "policy_core.h"
#include "../../include/srl_policy.h"

namespace {

constexpr int H = SRL_POLICY_HIDDEN;
constexpr int POLICY_LANES = 4;                   // threads per env in the last layer and for the per-env work
constexpr int POLICY_ENVS = 32;                   // envs per CTA -> 4096 envs = 128 CTAs of 128 threads (round 1: 64 CTAs of 64, one env per thread, 55 us)
constexpr int POLICY_BLOCK = POLICY_ENVS * POLICY_LANES;
constexpr int WS = H + 4;                         // padded row stride of the 64-wide rows: 16-byte aligned, and the 4 rows the lanes of an env read at
                                                  // the same time (o, o + 1, o + 2, o + 3) start 4 banks apart -- LDS.128 without bank conflicts

struct PolicyArgs {
    srl_mlp_policy p;
    int n;
    const float* obs;
    unsigned long long* rng;
    unsigned long long env_offset;
    float* obs_buf; void* act_env; void* act_buf; float* logp; float* value;
};

// Staging of the weights: every thread first ISSUES all of its global loads (registers), then stores them to shared memory -- one exposed
// L2 latency for the whole set instead of one per array (a load -> store loop per array cost ~1 us each, 12 arrays).
template <int PER>
struct RowRegs { float4 v[PER]; };
// 64-wide rows -> padded shared rows, 16 bytes per load; PER = ceil(rows * 16 / POLICY_BLOCK)
template <int PER>
__device__ __forceinline__ void rows_load(RowRegs<PER>& r, const float* __restrict__ src, int rows, bool vec) {
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * POLICY_BLOCK;
        r.v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < rows * (H / 4)) {
            if (vec) r.v[k] = __ldg(reinterpret_cast<const float4*>(src) + i);
            else r.v[k] = make_float4(__ldg(src + 4 * i), __ldg(src + 4 * i + 1), __ldg(src + 4 * i + 2), __ldg(src + 4 * i + 3));   // a view at an odd offset
        }
    }
}
template <int PER>
__device__ __forceinline__ void rows_store(const RowRegs<PER>& r, float* dst, int rows) {
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * POLICY_BLOCK;
        if (i < rows * (H / 4)) *reinterpret_cast<float4*>(dst + (i >> 4) * WS + 4 * (i & 15)) = r.v[k];
    }
}
template <int PER>
struct VecRegs { float v[PER]; };
template <int PER>
__device__ __forceinline__ void vec_load(VecRegs<PER>& r, const float* __restrict__ src, int count) {
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * POLICY_BLOCK;
        r.v[k] = i < count ? __ldg(src + i) : 0.f;
    }
}
template <int PER>
__device__ __forceinline__ void vec_store(const VecRegs<PER>& r, float* dst, int count) {
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * POLICY_BLOCK;
        if (i < count) dst[i] = r.v[k];
    }
}

struct TowerSmem { const float *w1, *b1, *w2, *b2, *w3, *b3; };

__device__ __forceinline__ void load_column(const float* h, float (&a)[H]) {
#pragma unroll
    for (int i4 = 0; i4 < H / 4; ++i4) {
        const float4 v = *reinterpret_cast<const float4*>(h + 4 * i4);
        a[4 * i4] = v.x; a[4 * i4 + 1] = v.y; a[4 * i4 + 2] = v.z; a[4 * i4 + 3] = v.w;
    }
}
// one padded 64-wide row against the activations: the four partial sums and their combination of policy_core.h's srl_mlp_tower
__device__ __forceinline__ float dot_row(const float* row, const float (&a)[H], float bias) {
    float s0 = bias, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i4 = 0; i4 < H / 4; ++i4) {
        const float4 w = *reinterpret_cast<const float4*>(row + 4 * i4);
        s0 = fmaf(w.x, a[4 * i4 + 0], s0); s1 = fmaf(w.y, a[4 * i4 + 1], s1);
        s2 = fmaf(w.z, a[4 * i4 + 2], s2); s3 = fmaf(w.w, a[4 * i4 + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}

// One 64-64 tower for the CTA's 32 envs: same arithmetic, value by value, as srl_mlp_tower (policy_core.h) -- per output the four partial sums
// over i4 = 0..15 in order, then (s0 + s1) + (s2 + s3).  Layers 1 and 2 are REGISTER-TILED: thread t owns 4 envs (4 (t / 16) + 0..3) x 4
// outputs ((t % 16) + 0, 16, 32, 48) -- per 4 inputs it loads 4 weight quads + 4 activation quads (8 LDS.128) for 64 FFMA into 64 independent
// accumulators.  (One env's 16 outputs per thread needed 1 LDS.128 per 4 FFMA and was bound by the shared-memory pipe: ncu, 36 % of the
// stalls on the first FFMA after a weight load, profiles/r02_policy_act_ncu.txt.)  The 8 lanes of a quarter-warp read 8 consecutive weight
// rows (stride WS = 68 words: 8 different 4-bank groups) and one common activation quad (broadcast).  `xs` [32][8] observations, `ha` / `hb`
// [32][WS] activation columns (layer 1 -> ha, layer 2 -> hb), `out` [32][stride] receives the last layer (thread t: env t / 4, outputs t % 4 + 4 k).
__device__ __forceinline__ void tower_tiled(const TowerSmem& W, int D, int n_out, const float* xs, float* ha, float* hb, float* out, int out_stride) {
    const int t = threadIdx.x, eg = t >> 4, og = t & 15;
    {   // layer 1: obs_dim -> 64
        float acc[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e][k] = W.b1[og + 16 * k];
        for (int d = 0; d < D; ++d) {
            float w[4], x[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) w[k] = W.w1[(og + 16 * k) * D + d];
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = xs[(4 * eg + e) * SRL_POLICY_MAX_OBS + d];
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[e][k] = fmaf(w[k], x[e], acc[e][k]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int k = 0; k < 4; ++k) ha[(4 * eg + e) * WS + og + 16 * k] = tanhf(acc[e][k]);
    }
    __syncthreads();
    {   // layer 2: 64 -> 64
        float s[4][4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int k = 0; k < 4; ++k) { s[e][k][0] = W.b2[og + 16 * k]; s[e][k][1] = 0.f; s[e][k][2] = 0.f; s[e][k][3] = 0.f; }
#pragma unroll 2
        for (int i4 = 0; i4 < H / 4; ++i4) {
            float4 w[4], a[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) w[k] = *reinterpret_cast<const float4*>(W.w2 + (og + 16 * k) * WS + 4 * i4);
#pragma unroll
            for (int e = 0; e < 4; ++e) a[e] = *reinterpret_cast<const float4*>(ha + (4 * eg + e) * WS + 4 * i4);
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    s[e][k][0] = fmaf(w[k].x, a[e].x, s[e][k][0]); s[e][k][1] = fmaf(w[k].y, a[e].y, s[e][k][1]);
                    s[e][k][2] = fmaf(w[k].z, a[e].z, s[e][k][2]); s[e][k][3] = fmaf(w[k].w, a[e].w, s[e][k][3]);
                }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int k = 0; k < 4; ++k) hb[(4 * eg + e) * WS + og + 16 * k] = tanhf((s[e][k][0] + s[e][k][1]) + (s[e][k][2] + s[e][k][3]));
    }
    __syncthreads();
    {   // layer 3: 64 -> n_out (<= 8): thread t -> env t / 4, outputs t % 4 and t % 4 + 4
        const int e = t >> 2, u = t & 3;
        if (u < n_out) {
            float a[H];
            load_column(hb + e * WS, a);
            for (int k = u; k < n_out; k += POLICY_LANES) out[e * out_stride + k] = dot_row(W.w3 + k * WS, a, W.b3[k]);
        }
    }
}

__global__ void __launch_bounds__(POLICY_BLOCK) policy_act_kernel(const __grid_constant__ PolicyArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int D = a.p.obs_dim, A = a.p.n_out;
    float* pi_w2 = smem;                 float* vf_w2 = pi_w2 + H * WS;
    float* pi_w3 = vf_w2 + H * WS;       float* vf_w3 = pi_w3 + SRL_POLICY_MAX_OUT * WS;
    float* ha = vf_w3 + WS;              float* hb = ha + POLICY_ENVS * WS;       // [POLICY_ENVS][WS] activation columns (16-byte aligned)
    float* pi_w1 = hb + POLICY_ENVS * WS;            float* vf_w1 = pi_w1 + H * SRL_POLICY_MAX_OBS;
    float* pi_b1 = vf_w1 + H * SRL_POLICY_MAX_OBS;   float* vf_b1 = pi_b1 + H;
    float* pi_b2 = vf_b1 + H;            float* vf_b2 = pi_b2 + H;
    float* pi_b3 = vf_b2 + H;            float* vf_b3 = pi_b3 + SRL_POLICY_MAX_OUT;
    float* s_logstd = vf_b3 + 4;         float* outs = s_logstd + SRL_POLICY_MAX_OUT;   // [POLICY_ENVS][SRL_POLICY_MAX_OUT + 1]: logits / mean, value
    float* xs = outs + POLICY_ENVS * (SRL_POLICY_MAX_OUT + 1);                           // [POLICY_ENVS][SRL_POLICY_MAX_OBS] observations
    {
        constexpr int W2PER = H * (H / 4) / POLICY_BLOCK, W3PER = (SRL_POLICY_MAX_OUT * (H / 4) + POLICY_BLOCK - 1) / POLICY_BLOCK;
        constexpr int W1PER = (H * SRL_POLICY_MAX_OBS + POLICY_BLOCK - 1) / POLICY_BLOCK, XPER = (POLICY_ENVS * SRL_POLICY_MAX_OBS + POLICY_BLOCK - 1) / POLICY_BLOCK;
        static_assert(H * (H / 4) % POLICY_BLOCK == 0 && H <= POLICY_BLOCK && SRL_POLICY_MAX_OUT <= POLICY_BLOCK, "staging shape");
        const bool vec = ((reinterpret_cast<uintptr_t>(a.p.pi_w2) | reinterpret_cast<uintptr_t>(a.p.vf_w2) | reinterpret_cast<uintptr_t>(a.p.pi_w3) |
                           reinterpret_cast<uintptr_t>(a.p.vf_w3)) & 15u) == 0;
        RowRegs<W2PER> r_pw2, r_vw2; RowRegs<W3PER> r_pw3; RowRegs<1> r_vw3;
        VecRegs<W1PER> r_pw1, r_vw1; VecRegs<1> r_pb1, r_vb1, r_pb2, r_vb2, r_pb3, r_vb3, r_ls; VecRegs<XPER> r_x;
        rows_load(r_pw2, a.p.pi_w2, H, vec); rows_load(r_vw2, a.p.vf_w2, H, vec); rows_load(r_pw3, a.p.pi_w3, A, vec); rows_load(r_vw3, a.p.vf_w3, 1, vec);
        vec_load(r_pw1, a.p.pi_w1, H * D); vec_load(r_vw1, a.p.vf_w1, H * D);
        vec_load(r_pb1, a.p.pi_b1, H); vec_load(r_vb1, a.p.vf_b1, H); vec_load(r_pb2, a.p.pi_b2, H); vec_load(r_vb2, a.p.vf_b2, H);
        vec_load(r_pb3, a.p.pi_b3, A); vec_load(r_vb3, a.p.vf_b3, 1);
        vec_load(r_ls, a.p.discrete ? a.p.pi_b3 : a.p.logstd, a.p.discrete ? 0 : A);
        // this CTA's observations: a contiguous run of (up to) 32 D floats
        const int first = blockIdx.x * POLICY_ENVS, nx = min(POLICY_ENVS, a.n - first) * D;
        vec_load(r_x, a.obs + (size_t)first * D, nx);
        rows_store(r_pw2, pi_w2, H); rows_store(r_vw2, vf_w2, H); rows_store(r_pw3, pi_w3, A); rows_store(r_vw3, vf_w3, 1);
        vec_store(r_pw1, pi_w1, H * D); vec_store(r_vw1, vf_w1, H * D);
        vec_store(r_pb1, pi_b1, H); vec_store(r_vb1, vf_b1, H); vec_store(r_pb2, pi_b2, H); vec_store(r_vb2, vf_b2, H);
        vec_store(r_pb3, pi_b3, A); vec_store(r_vb3, vf_b3, 1);
        vec_store(r_ls, s_logstd, a.p.discrete ? 0 : A);
#pragma unroll
        for (int k = 0; k < XPER; ++k) {                     // [env][D] -> [env][8]; envs past n read as zeros (their results are never stored)
            const int j = threadIdx.x + k * POLICY_BLOCK;
            if (j < POLICY_ENVS * D) xs[(j / D) * SRL_POLICY_MAX_OBS + (j % D)] = r_x.v[k];
            if (a.obs_buf && j < nx) a.obs_buf[(size_t)first * D + j] = r_x.v[k];
        }
    }
    const unsigned long long seed = a.rng[0], counter = a.rng[1];    // read before this CTA arrives: the counter moves only after ALL CTAs arrived
    __syncthreads();
    const TowerSmem Wpi = {pi_w1, pi_b1, pi_w2, pi_b2, pi_w3, pi_b3};
    const TowerSmem Wvf = {vf_w1, vf_b1, vf_w2, vf_b2, vf_w3, vf_b3};
    tower_tiled(Wpi, D, A, xs, ha, hb, outs, SRL_POLICY_MAX_OUT + 1);
    tower_tiled(Wvf, D, 1, xs, ha, hb, outs + SRL_POLICY_MAX_OUT, SRL_POLICY_MAX_OUT + 1);   // its layer 1 rewrites `ha`, last read before the previous tower's second barrier
    __syncwarp();                                            // an env's outputs were written by the 4 lanes t / 4 = env of this warp
    const int slot = threadIdx.x >> 2, u = threadIdx.x & 3;
    const int i = blockIdx.x * POLICY_ENVS + slot;
    if (i < a.n && u == 0) {                                 // the lead lane of each env samples and stores
        const float* out = outs + slot * (SRL_POLICY_MAX_OUT + 1);
        float lg[SRL_POLICY_MAX_OUT];
        for (int k = 0; k < A; ++k) lg[k] = out[k];
        a.value[i] = out[SRL_POLICY_MAX_OUT];
        const unsigned long long env = a.env_offset + (unsigned long long)i;
        float lp;
        if (a.p.discrete) {
            const int act = srl_sample_categorical(lg, A, seed, env, (uint32_t)counter, &lp);
            reinterpret_cast<int32_t*>(a.act_env)[i] = act;
            if (a.act_buf) reinterpret_cast<long long*>(a.act_buf)[i] = (long long)act;
        } else {
            float smp[SRL_POLICY_MAX_OUT], clp[SRL_POLICY_MAX_OUT];
            srl_sample_gaussian(lg, s_logstd, A, seed, env, (uint32_t)counter, smp, clp, &lp);
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

// One CTA: batch mean and biased variance in float64, Chan's parallel-variance merge into the running state (the update of stable-baselines'
// RunningMeanStd), then the normalisation of the whole batch in float32.  ONE pass over the batch: the sums S1 = sum(x - m0), S2 = sum((x - m0)^2)
// are taken about the running mean m0 (known before the batch is read), so mean = m0 + S1 / n and var = S2 / n - (S1 / n)^2 lose nothing to
// cancellation (|x - m0| is of the order of the standard deviation; float64 throughout: ~1e-13 of numpy's two-pass result at n = 4096), one
// block reduction of 2 D values instead of two of D with a second read in between, the D state updates by D threads.  D is a template
// parameter: every per-dimension array is registers (a run-time D put them in local memory: the first one-pass version was SLOWER, 24 us
// against 16), and a thread keeps its envs' observations in registers between the statistics and the normalisation (one read of the batch).
// Every phase of this kernel is a latency; there is no throughput to speak of.
template <int D>
__global__ void __launch_bounds__(FILTER_BLOCK) obs_filter_kernel(int n, const float* __restrict__ obs, double* state, int update,
                                                                   float clip, float eps, float* __restrict__ out) {
    constexpr int KEEP = 4;              // envs per thread held in registers (n <= KEEP * FILTER_BLOCK: the trainer's 4096); beyond that re-read
    __shared__ double red[32][2 * D];
    __shared__ double s_sum[2 * D];
    __shared__ float s_mf[D], s_inv[D];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float x[KEEP][D];
#pragma unroll
    for (int k = 0; k < KEEP; ++k) {
        const int i = tid + k * FILTER_BLOCK;
#pragma unroll
        for (int d = 0; d < D; ++d) x[k][d] = i < n ? obs[(size_t)i * D + d] : 0.f;
    }
    if (update) {
        double m0[D], acc[2 * D];
#pragma unroll
        for (int d = 0; d < D; ++d) { m0[d] = state[d]; acc[d] = 0.0; acc[D + d] = 0.0; }
#pragma unroll
        for (int k = 0; k < KEEP; ++k) {
            if (tid + k * FILTER_BLOCK < n) {
#pragma unroll
                for (int d = 0; d < D; ++d) { const double c = (double)x[k][d] - m0[d]; acc[d] += c; acc[D + d] = fma(c, c, acc[D + d]); }
            }
        }
        for (int i = tid + KEEP * FILTER_BLOCK; i < n; i += FILTER_BLOCK) {
#pragma unroll
            for (int d = 0; d < D; ++d) { const double c = (double)obs[(size_t)i * D + d] - m0[d]; acc[d] += c; acc[D + d] = fma(c, c, acc[D + d]); }
        }
#pragma unroll
        for (int q = 0; q < 2 * D; ++q) {
            double v = acc[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][q] = v;
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int q = 0; q < 2 * D; ++q) {
                double v = red[lane][q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) s_sum[q] = v;
            }
        }
        __syncthreads();
        if (tid < D) {
            const int d = tid;
            const double count = state[2 * D], bc = (double)n, tot = count + bc;
            const double s1 = s_sum[d] / bc, bmean = state[d] + s1, bvar = s_sum[D + d] / bc - s1 * s1;
            const double mean = state[d], var = state[D + d], delta = bmean - mean;
            const double nm = mean + delta * bc / tot, nv = (var * count + bvar * bc + delta * delta * count * bc / tot) / tot;
            state[d] = nm; state[D + d] = nv;
            s_mf[d] = (float)nm; s_inv[d] = sqrtf((float)nv + eps);
        }
        __syncthreads();                 // every thread d < D has read the count before thread 0 moves it
        if (tid == 0) state[2 * D] += (double)n;
    } else {
        if (tid < D) { s_mf[tid] = (float)state[tid]; s_inv[tid] = sqrtf((float)state[D + tid] + eps); }
        __syncthreads();
    }
    float mf[D], sd[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { mf[d] = s_mf[d]; sd[d] = s_inv[d]; }
#pragma unroll
    for (int k = 0; k < KEEP; ++k) {
        const int i = tid + k * FILTER_BLOCK;
        if (i < n) {
#pragma unroll
            for (int d = 0; d < D; ++d) out[(size_t)i * D + d] = fminf(fmaxf((x[k][d] - mf[d]) / sd[d], -clip), clip);
        }
    }
    for (int i = tid + KEEP * FILTER_BLOCK; i < n; i += FILTER_BLOCK) {
#pragma unroll
        for (int d = 0; d < D; ++d) out[(size_t)i * D + d] = fminf(fmaxf((obs[(size_t)i * D + d] - mf[d]) / sd[d], -clip), clip);
    }
}

constexpr size_t policy_smem_bytes() {
    return sizeof(float) * (size_t)(2 * H * WS + SRL_POLICY_MAX_OUT * WS + WS + 2 * POLICY_ENVS * WS + 2 * H * SRL_POLICY_MAX_OBS + 4 * H + SRL_POLICY_MAX_OUT + 4 +
                                    SRL_POLICY_MAX_OUT + POLICY_ENVS * (SRL_POLICY_MAX_OUT + 1) + POLICY_ENVS * SRL_POLICY_MAX_OBS);
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
    policy_act_kernel<<<(n + POLICY_ENVS - 1) / POLICY_ENVS, POLICY_BLOCK, smem, (cudaStream_t)stream>>>(a);
    SRL_CUDA_OK(cudaGetLastError());
    return 0;
}

int srl_obs_filter(int n, int obs_dim, const float* obs_raw, double* state, int update, float clip, float eps, float* obs_norm_out,
                   void* stream) {
    if (!obs_raw || !state || !obs_norm_out) { srl_set_error("obs_filter: null argument"); return 1; }
    if (n <= 0 || obs_dim < 1 || obs_dim > SRL_POLICY_MAX_OBS) { srl_set_error("obs_filter: unsupported shape n=%d obs_dim=%d", n, obs_dim); return 1; }
    switch (obs_dim) {
#define SRL_FILTER_CASE(DD) case DD: obs_filter_kernel<DD><<<1, FILTER_BLOCK, 0, (cudaStream_t)stream>>>(n, obs_raw, state, update, clip, eps, obs_norm_out); break;
        SRL_FILTER_CASE(1) SRL_FILTER_CASE(2) SRL_FILTER_CASE(3) SRL_FILTER_CASE(4) SRL_FILTER_CASE(5) SRL_FILTER_CASE(6) SRL_FILTER_CASE(7) SRL_FILTER_CASE(8)
#undef SRL_FILTER_CASE
        default: srl_set_error("obs_filter: unsupported obs_dim %d", obs_dim); return 1;
    }
    SRL_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
