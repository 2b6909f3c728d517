// PPO2 minibatch gradient in one pass (include/srl_policy.h: srl_ppo2_grad) -- the optimiser phase of the GPU-resident PPO2 consumer
// (the reference trains through stable-baselines' PPO2 with MlpPolicy, rl_baselines/rl_algorithm/ppo2.py:58-72).
//
// What it replaces: torch autograd over a 131 072-sample minibatch of config 3 -- ~150 launches, every activation of both 64-64 towers
// written to and read back from HBM / L2 (33 MB per [mb, 64] tensor), 2.2 ms per minibatch step even when captured in a CUDA graph, 71 % of a
// PPO2 update once the collection loop costs 99 us per env step.  Here a persistent CTA keeps the weights (and W2 transposed) in shared
// memory, walks its share of the minibatch in chunks of 64 samples, and for each chunk runs forward -> PPO2 loss derivative -> backward with
// every activation in shared memory and the weight gradients accumulating in registers; the per-CTA partial gradients are summed in a fixed
// order by a second kernel (deterministic: no atomics).  Gradient clipping and Adam stay with torch (a dozen launches over 9.3 k parameters).
//
// The loss is stable-baselines' PPO2 loss as rl_baselines/ppo2.py states it (its `minibatch_step` is the torch reference of the tests):
//   A_n = (adv_n - mean) / (std + 1e-8) over the minibatch (std unbiased), ratio = exp(logp - old_logp),
//   pg = mean(max(-A ratio, -A clip(ratio, 1 - c, 1 + c))), vf = 0.5 mean(max((v - R)^2, (old_v + clip(v - old_v, -c, c) - R)^2)),
//   loss = pg - ent_coef mean(entropy) + vf_coef vf.
#include <cuda_runtime.h>
#include <math.h>
#include "common.cuh"
#include "../../include/srl_policy.h"

namespace {

constexpr int H = 64, WS = 68, CH = 64, NT = 256, MAXO = 8, MAXD = 8;      // 64 samples per chunk, 8 warps per CTA (two per scheduler)
static_assert(NT == 4 * CH && NT == 256, "thread roles: 4 threads per sample in the heads, 128 owners per tower of the W2 gradient patches");

// offsets of the parameter tensors inside a flat gradient vector (the order of the per-CTA partials)
struct Seg { int pw1, pb1, pw2, pb2, pw3, pb3, vw1, vb1, vw2, vb2, vw3, vb3, ls, P; };
__host__ __device__ inline Seg make_seg(int D, int A, int discrete) {
    Seg s; int o = 0;
    s.pw1 = o; o += H * D; s.pb1 = o; o += H; s.pw2 = o; o += H * H; s.pb2 = o; o += H; s.pw3 = o; o += A * H; s.pb3 = o; o += A;
    s.vw1 = o; o += H * D; s.vb1 = o; o += H; s.vw2 = o; o += H * H; s.vb2 = o; o += H; s.vw3 = o; o += H; s.vb3 = o; o += 1;
    s.ls = o; o += discrete ? 0 : A; s.P = o;
    return s;
}

struct GradArgs {
    srl_mlp_policy p;
    int mb;
    const long long* idx; const float* obs; const void* act; const float* adv; const float* ret; const float* old_logp; const float* old_val;
    float clip, ent_coef, vf_coef;
    const double* stats;      // STATS_CTAS x {sum(x - shift), sum((x - shift)^2)}, then the shift: the minibatch's advantages
    float* partial;           // [gridDim.x][P]
};

// ---- advantage statistics of the minibatch: STATS_CTAS partial sums (float64, about the first element: no cancellation), combined in a fixed
//      order by every CTA of the gradient kernel (a single CTA gathering 131 072 elements took 75 us) ----
constexpr int STATS_CTAS = 64, STATS_NT = 256;
__global__ void __launch_bounds__(STATS_NT) adv_stats_kernel(int mb, const long long* __restrict__ idx, const float* __restrict__ adv, double* __restrict__ part) {
    __shared__ double red[STATS_NT / 32][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double shift = (double)adv[idx ? idx[0] : 0];
    double s1 = 0.0, s2 = 0.0;
    for (int e0 = blockIdx.x * STATS_NT + tid; e0 < mb; e0 += 8 * STATS_CTAS * STATS_NT) {
        float x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { const int e = e0 + k * STATS_CTAS * STATS_NT; x[k] = e < mb ? adv[idx ? idx[e] : e] : 0.f; }
#pragma unroll
        for (int k = 0; k < 8; ++k) if (e0 + k * STATS_CTAS * STATS_NT < mb) { const double c = (double)x[k] - shift; s1 += c; s2 = fma(c, c, s2); }
    }
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if (lane == 0) { red[warp][0] = s1; red[warp][1] = s2; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < STATS_NT / 32; ++w) { s1 += red[w][0]; s2 += red[w][1]; }
        part[2 * blockIdx.x] = s1; part[2 * blockIdx.x + 1] = s2;
        if (blockIdx.x == 0) part[2 * STATS_CTAS] = shift;
    }
}

// out[n][o] for the thread's 4 samples (4 eg + e) x 4 units (og + 16 k): bias[o] + sum_i W[o][i] in[n][i], W rows padded to WS words.
// Four partial sums per output over the 16 input quads, as csrc/policy_core.h's tower does.
__device__ __forceinline__ void tile_matvec(const float* W, const float* bias, const float* in, int eg, int og, float (&out)[4][4]) {
    float s[4][4][4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int k = 0; k < 4; ++k) { s[e][k][0] = bias ? bias[og + 16 * k] : 0.f; s[e][k][1] = 0.f; s[e][k][2] = 0.f; s[e][k][3] = 0.f; }
#pragma unroll 2
    for (int i4 = 0; i4 < H / 4; ++i4) {
        float4 w[4], a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = *reinterpret_cast<const float4*>(W + (og + 16 * k) * WS + 4 * i4);
#pragma unroll
        for (int e = 0; e < 4; ++e) a[e] = *reinterpret_cast<const float4*>(in + (4 * eg + e) * WS + 4 * i4);
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
        for (int k = 0; k < 4; ++k) out[e][k] = (s[e][k][0] + s[e][k][1]) + (s[e][k][2] + s[e][k][3]);
}

__device__ __forceinline__ float dot64(const float* row, const float* col) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i4 = 0; i4 < H / 4; ++i4) {
        const float4 w = *reinterpret_cast<const float4*>(row + 4 * i4), a = *reinterpret_cast<const float4*>(col + 4 * i4);
        s0 = fmaf(w.x, a.x, s0); s1 = fmaf(w.y, a.y, s1); s2 = fmaf(w.z, a.z, s2); s3 = fmaf(w.w, a.w, s3);
    }
    return (s0 + s1) + (s2 + s3);
}

struct SampleRegs { float x[MAXD], af[MAXO], adv, ret, olp, ov; int ai, valid; };

__device__ __forceinline__ void load_sample(const GradArgs& a, int s, SampleRegs& r) {
    const int D = a.p.obs_dim, A = a.p.n_out;
    r.valid = s < a.mb;
    const long long g = r.valid ? (a.idx ? a.idx[s] : (long long)s) : 0;
#pragma unroll
    for (int d = 0; d < MAXD; ++d) r.x[d] = (r.valid && d < D) ? a.obs[g * D + d] : 0.f;
    r.ai = 0;
#pragma unroll
    for (int k = 0; k < MAXO; ++k) r.af[k] = 0.f;
    if (r.valid) {
        if (a.p.discrete) r.ai = (int)reinterpret_cast<const long long*>(a.act)[g];
        else {
#pragma unroll
            for (int k = 0; k < MAXO; ++k) if (k < A) r.af[k] = reinterpret_cast<const float*>(a.act)[g * A + k];
        }
    }
    r.adv = r.valid ? a.adv[g] : 0.f; r.ret = r.valid ? a.ret[g] : 0.f; r.olp = r.valid ? a.old_logp[g] : 0.f; r.ov = r.valid ? a.old_val[g] : 0.f;
}

__global__ void __launch_bounds__(NT, 1) ppo2_grad_kernel(const __grid_constant__ GradArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int D = a.p.obs_dim, A = a.p.n_out, t = threadIdx.x;
    const bool discrete = a.p.discrete != 0;
    const Seg seg = make_seg(D, A, a.p.discrete);
    // ---- shared memory map (floats) ----
    float* w2p = sm;                 float* w2v = w2p + H * WS;       float* w2pT = w2v + H * WS;      float* w2vT = w2pT + H * WS;
    float* w3p = w2vT + H * WS;      float* w3v = w3p + MAXO * WS;
    float* hap = w3v + WS;           float* hbp = hap + CH * WS;      float* hav = hbp + CH * WS;      float* hbv = hav + CH * WS;
    float* d2p = hbv + CH * WS;      float* d2v = d2p + CH * WS;      float* d1p = d2v + CH * WS;      float* d1v = d1p + CH * WS;
    float* w1p = d1v + CH * WS;      float* w1v = w1p + H * MAXD;
    float* b1p = w1v + H * MAXD;     float* b1v = b1p + H;            float* b2p = b1v + H;            float* b2v = b2p + H;
    float* b3p = b2v + H;            float* b3v = b3p + MAXO;         float* lsd = b3v + 4;            // logstd
    float* xs = lsd + MAXO;          // [CH][MAXD]
    float* zo = xs + CH * MAXD;      // [CH][MAXO + 1]: logits / mean, value at [MAXO]
    float* d3 = zo + CH * (MAXO + 1);   // [CH][MAXO + 1]: d loss / d logits, d loss / d value
    float* dls = d3 + CH * (MAXO + 1);  // [CH][MAXO]: per-sample d loss / d logstd
    float* saf = dls + CH * MAXO;    // [CH][MAXO] actions (Box)
    float* ssc = saf + CH * MAXO;    // [CH][4]: adv, ret, old_logp, old_val
    int* sai = reinterpret_cast<int*>(ssc + CH * 4);   // [CH][2]: action (Discrete), valid
    // ---- weights, once per CTA ----
    for (int e = t; e < H * H; e += NT) {
        const int j = e >> 6, i = e & 63;
        const float wp = a.p.pi_w2[e], wv = a.p.vf_w2[e];
        w2p[j * WS + i] = wp; w2pT[i * WS + j] = wp; w2v[j * WS + i] = wv; w2vT[i * WS + j] = wv;
    }
    for (int e = t; e < A * H; e += NT) w3p[(e >> 6) * WS + (e & 63)] = a.p.pi_w3[e];
    for (int e = t; e < H; e += NT) {
        w3v[e] = a.p.vf_w3[e];
        b1p[e] = a.p.pi_b1[e]; b1v[e] = a.p.vf_b1[e]; b2p[e] = a.p.pi_b2[e]; b2v[e] = a.p.vf_b2[e];
    }
    for (int e = t; e < H * D; e += NT) { w1p[(e / D) * MAXD + e % D] = a.p.pi_w1[e]; w1v[(e / D) * MAXD + e % D] = a.p.vf_w1[e]; }
    if (t < A) { b3p[t] = a.p.pi_b3[t]; lsd[t] = discrete ? 0.f : a.p.logstd[t]; }
    if (t == 0) b3v[0] = a.p.vf_b3[0];
    __shared__ float s_adv[2];
    if (t < 32) {                  // mean and 1 / (std + 1e-8) (torch.std(): unbiased) from the partial sums, in a fixed order
        double s1 = a.stats[2 * t] + a.stats[2 * (t + 32)], s2 = a.stats[2 * t + 1] + a.stats[2 * (t + 32) + 1];
        for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        if (t == 0) {
            const double n = (double)a.mb, mean = a.stats[2 * STATS_CTAS] + s1 / n;
            const double var = a.mb > 1 ? fmax((s2 - s1 * s1 / n) / (n - 1.0), 0.0) : 0.0;
            s_adv[0] = (float)mean; s_adv[1] = (float)(1.0 / (sqrt(var) + 1e-8));
        }
    }
    __syncthreads();
    const float amean = s_adv[0], ainv = s_adv[1], inv_mb = 1.0f / (float)a.mb;
    // ---- gradient accumulators (registers; every entry of the flat gradient has exactly one owner thread) ----
    const int eg = t >> 4, og = t & 15;          // forward / delta tiles: samples 4 eg + e, units og + 16 k
    const bool own_pi = t < 128;                 // W2 gradient: threads 0..127 own the policy tower's 4 x 8 patches, 128..255 the value tower's
    const int jq = (t & 127) >> 3, iq = t & 7;   // rows 4 jq + jj, columns 8 iq + ii
    float gw2[4][8];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) gw2[jj][ii] = 0.f;
    float gw3p[4] = {0.f, 0.f, 0.f, 0.f}, gw1p[4] = {0.f, 0.f, 0.f, 0.f}, gw1v[4] = {0.f, 0.f, 0.f, 0.f};
    float gw3v = 0.f, gb3 = 0.f, gls = 0.f, gb2 = 0.f, gb1 = 0.f;       // gb2 / gb1: t < 64 the policy tower's unit t, t >= 64 the value tower's unit t - 64
    const int nchunks = (a.mb + CH - 1) / CH;
    SampleRegs cur;
    if (t < CH) load_sample(a, blockIdx.x * CH + t, cur);
    __syncthreads();
    for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
        // ---- 0: this chunk's samples -> shared; the next chunk's loads are issued now and consumed an iteration later ----
        if (t < CH) {
#pragma unroll
            for (int d = 0; d < MAXD; ++d) xs[t * MAXD + d] = cur.x[d];
#pragma unroll
            for (int k = 0; k < MAXO; ++k) saf[t * MAXO + k] = cur.af[k];
            ssc[t * 4] = cur.adv; ssc[t * 4 + 1] = cur.ret; ssc[t * 4 + 2] = cur.olp; ssc[t * 4 + 3] = cur.ov;
            sai[t * 2] = cur.ai; sai[t * 2 + 1] = cur.valid;
            load_sample(a, (c + gridDim.x) * CH + t, cur);
        }
        __syncthreads();
        // ---- 1: forward, both towers ----
        {
            float accp[4][4], accv[4][4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int e = 0; e < 4; ++e) { accp[e][k] = b1p[og + 16 * k]; accv[e][k] = b1v[og + 16 * k]; }
            for (int d = 0; d < D; ++d) {
                float wp[4], wv[4], x[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { wp[k] = w1p[(og + 16 * k) * MAXD + d]; wv[k] = w1v[(og + 16 * k) * MAXD + d]; }
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] = xs[(4 * eg + e) * MAXD + d];
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int k = 0; k < 4; ++k) { accp[e][k] = fmaf(wp[k], x[e], accp[e][k]); accv[e][k] = fmaf(wv[k], x[e], accv[e][k]); }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int k = 0; k < 4; ++k) { hap[(4 * eg + e) * WS + og + 16 * k] = tanhf(accp[e][k]); hav[(4 * eg + e) * WS + og + 16 * k] = tanhf(accv[e][k]); }
        }
        __syncthreads();
        {
            float o[4][4];
            tile_matvec(w2p, b2p, hap, eg, og, o);
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int k = 0; k < 4; ++k) hbp[(4 * eg + e) * WS + og + 16 * k] = tanhf(o[e][k]);
            tile_matvec(w2v, b2v, hav, eg, og, o);
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int k = 0; k < 4; ++k) hbv[(4 * eg + e) * WS + og + 16 * k] = tanhf(o[e][k]);
        }
        __syncthreads();
        {   // heads: thread t -> sample t / 4; outputs t % 4 and t % 4 + 4 of the policy head, the value by lane 3 of the sample's four
            const int n = t >> 2, u = t & 3;
            for (int k = u; k < A; k += 4) zo[n * (MAXO + 1) + k] = b3p[k] + dot64(w3p + k * WS, hbp + n * WS);
            if (u == 3) zo[n * (MAXO + 1) + MAXO] = b3v[0] + dot64(w3v, hbv + n * WS);
        }
        __syncthreads();
        // ---- 2: d loss / d outputs, one thread per sample ----
        if (t < CH) {
            const int n = t;
            float z[MAXO], g[MAXO];
#pragma unroll
            for (int k = 0; k < MAXO; ++k) { z[k] = k < A ? zo[n * (MAXO + 1) + k] : 0.f; g[k] = 0.f; }
            const float v = zo[n * (MAXO + 1) + MAXO];
            const float adv = ssc[n * 4], R = ssc[n * 4 + 1], olp = ssc[n * 4 + 2], ov = ssc[n * 4 + 3];
            const bool valid = sai[n * 2 + 1] != 0;
            const float An = (adv - amean) * ainv, c = a.clip;
            float gv = 0.f, gl[MAXO];
#pragma unroll
            for (int k = 0; k < MAXO; ++k) gl[k] = 0.f;
            if (valid) {
                float logp;
                float p[MAXO], lse = 0.f, Hent = 0.f;
                if (discrete) {
                    float m = z[0];
#pragma unroll
                    for (int k = 1; k < MAXO; ++k) if (k < A) m = fmaxf(m, z[k]);
                    float S = 0.f;
#pragma unroll
                    for (int k = 0; k < MAXO; ++k) { p[k] = k < A ? expf(z[k] - m) : 0.f; S += p[k]; }
                    lse = m + logf(S);
                    const float iS = 1.0f / S;
                    float pz = 0.f, za = 0.f;
                    const int ai = sai[n * 2];
#pragma unroll
                    for (int k = 0; k < MAXO; ++k) { p[k] *= iS; pz = fmaf(p[k], z[k], pz); if (k == ai) za = z[k]; }
                    Hent = lse - pz;
                    logp = za - lse;
                } else {
                    logp = 0.f;
#pragma unroll
                    for (int k = 0; k < MAXO; ++k) if (k < A) {
                        const float is = expf(-lsd[k]), u = (saf[n * MAXO + k] - z[k]) * is;
                        logp += -0.5f * u * u - lsd[k] - 0.91893853320467274178f;
                        p[k] = u;                      // (a - mu) / sigma, reused below
                    }
                }
                const float ratio = expf(logp - olp);
                const float rc = fminf(fmaxf(ratio, 1.0f - c), 1.0f + c);
                const float unclipped = -An * ratio, clipped = -An * rc;
                const float dr = (rc == ratio || unclipped > clipped) ? -An : 0.f;      // max(): the live branch (a tie inside the clip range sums to the same)
                const float dlogp = dr * ratio * inv_mb;
                if (discrete) {
                    const int ai = sai[n * 2];
                    const float ec = a.ent_coef * inv_mb;
#pragma unroll
                    for (int k = 0; k < MAXO; ++k) if (k < A) g[k] = dlogp * ((k == ai ? 1.f : 0.f) - p[k]) + ec * p[k] * ((z[k] - lse) + Hent);
                } else {
#pragma unroll
                    for (int k = 0; k < MAXO; ++k) if (k < A) {
                        const float is = expf(-lsd[k]);
                        g[k] = dlogp * p[k] * is;                                   // d logp / d mu = (a - mu) / sigma^2
                        gl[k] = dlogp * (p[k] * p[k] - 1.0f) - a.ent_coef * inv_mb;     // d logp / d logstd; entropy = sum(logstd) + const
                    }
                }
                const float dv = v - ov, dvc = fminf(fmaxf(dv, -c), c);
                const float e1 = v - R, e2 = (ov + dvc) - R;
                const float l1 = e1 * e1, l2 = e2 * e2;
                gv = (dvc == dv || l1 > l2) ? e1 : (l1 == l2 ? 0.5f * e1 : 0.f);
                gv *= a.vf_coef * inv_mb;
            }
#pragma unroll
            for (int k = 0; k < MAXO; ++k) { d3[n * (MAXO + 1) + k] = g[k]; dls[n * MAXO + k] = gl[k]; }
            d3[n * (MAXO + 1) + MAXO] = gv;
        }
        __syncthreads();
        // ---- 3: delta of the second hidden layer (pre-activation), both towers ----
        {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int n = 4 * eg + e;
                float gz[MAXO];
#pragma unroll
                for (int k = 0; k < MAXO; ++k) gz[k] = d3[n * (MAXO + 1) + k];
                const float gvn = d3[n * (MAXO + 1) + MAXO];
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const int j = og + 16 * k4;
                    float s = 0.f;
#pragma unroll
                    for (int k = 0; k < MAXO; ++k) if (k < A) s = fmaf(w3p[k * WS + j], gz[k], s);
                    const float hp = hbp[n * WS + j], hv = hbv[n * WS + j];
                    d2p[n * WS + j] = s * (1.0f - hp * hp);
                    d2v[n * WS + j] = w3v[j] * gvn * (1.0f - hv * hv);
                }
            }
        }
        __syncthreads();
        // ---- 4: gradients of layers 3 and 2; delta of the first hidden layer ----
        {
            for (int m = 0; m < 4; ++m) {            // policy head weights: entries t + 128 m of [A][64]
                const int e = t + NT * m;
                if (e < A * H) {
                    const int k = e >> 6, j = e & 63;
                    float s = 0.f;
#pragma unroll 8
                    for (int n = 0; n < CH; ++n) s = fmaf(d3[n * (MAXO + 1) + k], hbp[n * WS + j], s);
                    gw3p[m] += s;
                }
            }
            if (t < H) {
                float s = 0.f, sb = 0.f;
#pragma unroll 8
                for (int n = 0; n < CH; ++n) { s = fmaf(d3[n * (MAXO + 1) + MAXO], hbv[n * WS + t], s); sb += d2p[n * WS + t]; }
                gw3v += s; gb2 += sb;
            } else if (t < 2 * H) {
                float sb = 0.f;
#pragma unroll 8
                for (int n = 0; n < CH; ++n) sb += d2v[n * WS + (t - H)];
                gb2 += sb;
            }
            if (t < A || t == MAXO) {                 // head biases: t < A the policy's, t == 8 the value's
                float s = 0.f;
                const int col = t < A ? t : MAXO;
                for (int n = 0; n < CH; ++n) s += d3[n * (MAXO + 1) + col];
                gb3 += s;
            }
            if (!discrete && t >= 16 && t < 16 + A) {
                float s = 0.f;
                for (int n = 0; n < CH; ++n) s += dls[n * MAXO + (t - 16)];
                gls += s;
            }
            // W2 gradient patches: per sample one delta quad and two activation quads
            {
                const float* dsrc = own_pi ? d2p : d2v;
                const float* hsrc = own_pi ? hap : hav;
#pragma unroll 4
                for (int n = 0; n < CH; ++n) {
                    const float4 dq = *reinterpret_cast<const float4*>(dsrc + n * WS + 4 * jq);
                    const float4 h0 = *reinterpret_cast<const float4*>(hsrc + n * WS + 8 * iq), h1 = *reinterpret_cast<const float4*>(hsrc + n * WS + 8 * iq + 4);
                    const float dd[4] = {dq.x, dq.y, dq.z, dq.w}, hh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                        for (int ii = 0; ii < 8; ++ii) gw2[jj][ii] = fmaf(dd[jj], hh[ii], gw2[jj][ii]);
                }
            }
            // delta 1 = (W2^T delta 2) (1 - h1^2): the same tile as the forward pass, on the transposed weights
            float o[4][4];
            tile_matvec(w2pT, nullptr, d2p, eg, og, o);
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float h = hap[(4 * eg + e) * WS + og + 16 * k]; d1p[(4 * eg + e) * WS + og + 16 * k] = o[e][k] * (1.0f - h * h); }
            tile_matvec(w2vT, nullptr, d2v, eg, og, o);
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float h = hav[(4 * eg + e) * WS + og + 16 * k]; d1v[(4 * eg + e) * WS + og + 16 * k] = o[e][k] * (1.0f - h * h); }
        }
        __syncthreads();
        // ---- 5: gradients of layer 1 ----
        {
            if (t < 2 * H) {
                const float* d1 = t < H ? d1p : d1v;
                const int unit = t < H ? t : t - H;
                float sb = 0.f;
#pragma unroll 8
                for (int n = 0; n < CH; ++n) sb += d1[n * WS + unit];
                gb1 += sb;
            }
            for (int m = 0; m < 4; ++m) {
                const int e = t + NT * m;
                if (e < H * D) {
                    const int i = e / D, d = e % D;
                    float sp = 0.f, sv = 0.f;
#pragma unroll 8
                    for (int n = 0; n < CH; ++n) { const float x = xs[n * MAXD + d]; sp = fmaf(d1p[n * WS + i], x, sp); sv = fmaf(d1v[n * WS + i], x, sv); }
                    gw1p[m] += sp; gw1v[m] += sv;
                }
            }
        }
        __syncthreads();
    }
    // ---- this CTA's partial gradient ----
    float* out = a.partial + (size_t)blockIdx.x * seg.P;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) out[(own_pi ? seg.pw2 : seg.vw2) + (4 * jq + jj) * H + 8 * iq + ii] = gw2[jj][ii];
    for (int m = 0; m < 4; ++m) {
        const int e = t + NT * m;
        if (e < A * H) out[seg.pw3 + e] = gw3p[m];
        if (e < H * D) { out[seg.pw1 + e] = gw1p[m]; out[seg.vw1 + e] = gw1v[m]; }
    }
    if (t < H) { out[seg.vw3 + t] = gw3v; out[seg.pb2 + t] = gb2; out[seg.pb1 + t] = gb1; }
    else if (t < 2 * H) { out[seg.vb2 + (t - H)] = gb2; out[seg.vb1 + (t - H)] = gb1; }
    if (t < A) out[seg.pb3 + t] = gb3;
    if (t == MAXO) out[seg.vb3] = gb3;
    if (!discrete && t >= 16 && t < 16 + A) out[seg.ls + (t - 16)] = gls;
}

// GAE(lambda) of one rollout, the reference's backward recursion (stable-baselines PPO2 runner): one thread per env, T sequential steps,
// eight steps' loads in flight.  Separate roundings (no FMA contraction): the same bits as the torch recursion of rl_baselines/ppo2.py.
__global__ void __launch_bounds__(128) gae_kernel(int T, int N, const float* __restrict__ rew, const float* __restrict__ val, const float* __restrict__ done,
                                                   const float* __restrict__ last_val, float gamma, float gl, float* __restrict__ adv, float* __restrict__ ret) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float lastgae = 0.f, nextval = last_val[n];
    for (int t0 = T - 1; t0 >= 0; t0 -= 8) {
        float r[8], v[8], d[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int t = t0 - k;
            const size_t o = (size_t)(t >= 0 ? t : 0) * N + n;
            r[k] = rew[o]; v[k] = val[o]; d[k] = done[o];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int t = t0 - k;
            if (t >= 0) {
                const float nonterminal = __fsub_rn(1.0f, d[k]);
                const float delta = __fsub_rn(__fadd_rn(r[k], __fmul_rn(__fmul_rn(gamma, nextval), nonterminal)), v[k]);
                lastgae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nonterminal), lastgae));
                const size_t o = (size_t)t * N + n;
                adv[o] = lastgae; ret[o] = __fadd_rn(lastgae, v[k]);
                nextval = v[k];
            }
        }
    }
}

struct ReduceArgs { srl_mlp_grads g; Seg seg; int nparts; const float* partial; };

// sum of the per-CTA partials in CTA order (deterministic) -> the parameter's gradient tensor
__global__ void ppo2_reduce_kernel(const __grid_constant__ ReduceArgs r) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r.seg.P) return;
    float s = 0.f;
    for (int c = 0; c < r.nparts; ++c) s += r.partial[(size_t)c * r.seg.P + e];
    const Seg& q = r.seg;
    float* dst; int off;
    if (e < q.pb1) { dst = r.g.pi_w1; off = e - q.pw1; } else if (e < q.pw2) { dst = r.g.pi_b1; off = e - q.pb1; }
    else if (e < q.pb2) { dst = r.g.pi_w2; off = e - q.pw2; } else if (e < q.pw3) { dst = r.g.pi_b2; off = e - q.pb2; }
    else if (e < q.pb3) { dst = r.g.pi_w3; off = e - q.pw3; } else if (e < q.vw1) { dst = r.g.pi_b3; off = e - q.pb3; }
    else if (e < q.vb1) { dst = r.g.vf_w1; off = e - q.vw1; } else if (e < q.vw2) { dst = r.g.vf_b1; off = e - q.vb1; }
    else if (e < q.vb2) { dst = r.g.vf_w2; off = e - q.vw2; } else if (e < q.vw3) { dst = r.g.vf_b2; off = e - q.vb2; }
    else if (e < q.vb3) { dst = r.g.vf_w3; off = e - q.vw3; } else if (e < q.ls) { dst = r.g.vf_b3; off = e - q.vb3; }
    else { dst = r.g.logstd; off = e - q.ls; }
    dst[off] = s;
}

constexpr size_t grad_smem_bytes() {
    return sizeof(float) * (size_t)(4 * H * WS + MAXO * WS + WS + 8 * CH * WS + 2 * H * MAXD + 4 * H + MAXO + 4 + MAXO + CH * MAXD + 2 * CH * (MAXO + 1) +
                                    2 * CH * MAXO + CH * 4 + CH * 2);
}

int grid_ctas(int mb) {
    static int sms[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev < 0 || dev >= 64) return 0;
    if (!sms[dev]) { if (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0; }
    const int chunks = (mb + CH - 1) / CH;
    return chunks < sms[dev] ? chunks : sms[dev];
}

}  // namespace

extern "C" {

size_t srl_ppo2_workspace_bytes(int obs_dim, int n_out, int discrete, int minibatch) {
    if (obs_dim < 1 || obs_dim > MAXD || n_out < 1 || n_out > MAXO || minibatch < 1) return 0;
    const Seg seg = make_seg(obs_dim, n_out, discrete);
    const int ctas = grid_ctas(minibatch);
    return 2048 + sizeof(float) * (size_t)seg.P * (size_t)(ctas > 0 ? ctas : 1);
}

int srl_ppo2_grad(const srl_mlp_policy* p, const srl_mlp_grads* grads, int minibatch, const int64_t* idx, const float* obs, const void* actions,
                  const float* adv, const float* ret, const float* old_logp, const float* old_value, float cliprange, float ent_coef, float vf_coef,
                  void* workspace, size_t workspace_bytes, void* stream) {
    if (!p || !grads || !obs || !actions || !adv || !ret || !old_logp || !old_value || !workspace) { srl_set_error("ppo2_grad: null argument"); return 1; }
    if (p->struct_size != sizeof(srl_mlp_policy) || grads->struct_size != sizeof(srl_mlp_grads)) { srl_set_error("ppo2_grad: struct size mismatch"); return 1; }
    if (p->obs_dim < 1 || p->obs_dim > MAXD || p->n_out < 1 || p->n_out > MAXO || (p->discrete && p->n_out < 2) || minibatch < 1) {
        srl_set_error("ppo2_grad: unsupported shape obs_dim=%d n_out=%d minibatch=%d", p->obs_dim, p->n_out, minibatch); return 1;
    }
    if (!p->pi_w1 || !p->pi_b1 || !p->pi_w2 || !p->pi_b2 || !p->pi_w3 || !p->pi_b3 || !p->vf_w1 || !p->vf_b1 || !p->vf_w2 || !p->vf_b2 || !p->vf_w3 ||
        !p->vf_b3 || (!p->discrete && !p->logstd)) { srl_set_error("ppo2_grad: null weight pointer"); return 1; }
    if (!grads->pi_w1 || !grads->pi_b1 || !grads->pi_w2 || !grads->pi_b2 || !grads->pi_w3 || !grads->pi_b3 || !grads->vf_w1 || !grads->vf_b1 || !grads->vf_w2 ||
        !grads->vf_b2 || !grads->vf_w3 || !grads->vf_b3 || (!p->discrete && !grads->logstd)) { srl_set_error("ppo2_grad: null gradient pointer"); return 1; }
    const int ctas = grid_ctas(minibatch);
    if (ctas <= 0) { srl_set_error("ppo2_grad: no CUDA device"); return 1; }
    const Seg seg = make_seg(p->obs_dim, p->n_out, p->discrete);
    if (workspace_bytes < 2048 + sizeof(float) * (size_t)seg.P * (size_t)ctas) { srl_set_error("ppo2_grad: workspace too small (srl_ppo2_workspace_bytes)"); return 1; }
    static bool attr_set[64] = {};
    int dev = 0;
    SRL_CUDA_OK(cudaGetDevice(&dev));
    constexpr size_t smem = grad_smem_bytes();
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        SRL_CUDA_OK(cudaFuncSetAttribute(ppo2_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    double* stats = reinterpret_cast<double*>(workspace);
    float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 2048);
    adv_stats_kernel<<<STATS_CTAS, STATS_NT, 0, st>>>(minibatch, reinterpret_cast<const long long*>(idx), adv, stats);
    SRL_CUDA_OK(cudaGetLastError());
    GradArgs a;
    a.p = *p; a.mb = minibatch; a.idx = reinterpret_cast<const long long*>(idx); a.obs = obs; a.act = actions; a.adv = adv; a.ret = ret;
    a.old_logp = old_logp; a.old_val = old_value; a.clip = cliprange; a.ent_coef = ent_coef; a.vf_coef = vf_coef; a.stats = stats; a.partial = partial;
    ppo2_grad_kernel<<<ctas, NT, smem, st>>>(a);
    SRL_CUDA_OK(cudaGetLastError());
    ReduceArgs r;
    r.g = *grads; r.seg = seg; r.nparts = ctas; r.partial = partial;
    ppo2_reduce_kernel<<<(seg.P + 255) / 256, 256, 0, st>>>(r);
    SRL_CUDA_OK(cudaGetLastError());
    return 0;
}

int srl_ppo2_gae(int n_steps, int n_envs, const float* rew, const float* value, const float* done, const float* last_value, float gamma, float lam,
                 float* adv_out, float* ret_out, void* stream) {
    if (!rew || !value || !done || !last_value || !adv_out || !ret_out) { srl_set_error("ppo2_gae: null argument"); return 1; }
    if (n_steps < 1 || n_envs < 1) { srl_set_error("ppo2_gae: bad shape"); return 1; }
    gae_kernel<<<(n_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_steps, n_envs, rew, value, done, last_value, gamma, (float)((double)gamma * (double)lam), adv_out, ret_out);
    SRL_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
