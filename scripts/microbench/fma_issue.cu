// Micro-benchmark (B200, sm_100a): how fast can ONE warp per scheduler issue the row update of the Kuka PGS sweep?
//   v[j] += A[j][i] * d_i   (12 x 12 per sweep)
// as 3-register FFMA, as packed FFMA2 (fma.rn.f32x2), and with the clamp chain done as FFMA.SAT -> FADD.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fma_issue fma_issue.cu ; run: ./fma_issue
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NB = 12;

template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k(const float* __restrict__ in, float* __restrict__ out, long long* cyc, int iters) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float A[NB][NB], v[NB], lam[NB], invd[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        v[i] = in[(tid * 7 + i) & 1023]; lam[i] = 0.5f; invd[i] = in[(tid + 31 * i) & 1023] * 1e-3f;
#pragma unroll
        for (int j = 0; j < NB; ++j) A[i][j] = in[(tid + i * NB + j) & 1023] * 1e-3f;
    }
    const long long t0 = clock64();
    if (MODE == 0) {          // scalar FFMA rows, min/max clamp chain (the round-1 sweep shape)
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const float s = fminf(fmaxf(fmaf(-invd[i], v[i], lam[i]), -0.8f), 0.8f);
                const float d = s - lam[i]; lam[i] = s;
#pragma unroll
                for (int j = 0; j < NB; ++j) v[j] = fmaf(A[j][i], d, v[j]);
            }
        }
    } else if (MODE == 1) {   // scalar FFMA rows, FFMA.SAT clamp
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const float s = __saturatef(fmaf(-invd[i], v[i], lam[i]));
                const float d = s - lam[i]; lam[i] = s;
#pragma unroll
                for (int j = 0; j < NB; ++j) v[j] = fmaf(A[j][i], d, v[j]);
            }
        }
    } else if (MODE == 2) {   // packed rows: 6 FFMA2 per row, FFMA.SAT clamp
        float2 Ap[NB][NB / 2], vp[NB / 2];
#pragma unroll
        for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int p = 0; p < NB / 2; ++p) Ap[i][p] = make_float2(A[2 * p][i], A[2 * p + 1][i]);
#pragma unroll
        for (int p = 0; p < NB / 2; ++p) vp[p] = make_float2(v[2 * p], v[2 * p + 1]);
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const float vi = (i & 1) ? vp[i >> 1].y : vp[i >> 1].x;
                const float s = __saturatef(fmaf(-invd[i], vi, lam[i]));
                const float d = s - lam[i]; lam[i] = s;
                const float2 d2 = make_float2(d, d);
#pragma unroll
                for (int p = 0; p < NB / 2; ++p) vp[p] = __ffma2_rn(Ap[i][p], d2, vp[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < NB / 2; ++p) { v[2 * p] = vp[p].x; v[2 * p + 1] = vp[p].y; }
    } else if (MODE == 3) {   // pure FFMA issue rate: 144 independent-ish 3-register FFMAs per iteration, no chain
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NB; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) v[j] = fmaf(A[j][i], lam[i], v[j]);
        }
    } else if (MODE == 4) {   // pure FFMA2 issue rate: 72 per iteration
        float2 Ap[NB][NB / 2], vp[NB / 2], l2[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            l2[i] = make_float2(invd[i], invd[i]);
#pragma unroll
            for (int p = 0; p < NB / 2; ++p) Ap[i][p] = make_float2(A[2 * p][i], A[2 * p + 1][i]);
        }
#pragma unroll
        for (int p = 0; p < NB / 2; ++p) vp[p] = make_float2(v[2 * p], v[2 * p + 1]);
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NB; ++i)
#pragma unroll
                for (int p = 0; p < NB / 2; ++p) vp[p] = __ffma2_rn(Ap[i][p], l2[i], vp[p]);
        }
#pragma unroll
        for (int p = 0; p < NB / 2; ++p) { v[2 * p] = vp[p].x; v[2 * p + 1] = vp[p].y; }
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < NB; ++i) acc += v[i] + lam[i];
    out[tid] = acc;
    if ((threadIdx.x & 31) == 0) cyc[tid >> 5] = t1 - t0;
}

template <int MODE, int threads>
void run(const char* name, int lanes, const float* in, float* out, long long* cyc, int per_iter) {
    const int iters = 2000, grid = 148;
    k<MODE, threads><<<grid, threads>>>(in, out, cyc, 10);
    cudaDeviceSynchronize();
    k<MODE, threads><<<grid, threads>>>(in, out, cyc, iters);
    cudaDeviceSynchronize();
    const int nw = grid * threads / 32;
    long long* h = new long long[nw];
    cudaMemcpy(h, cyc, nw * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0; long long mx = 0;
    for (int i = 0; i < nw; ++i) { mean += h[i]; if (h[i] > mx) mx = h[i]; }
    mean /= nw;
    printf("%-34s warps/scheduler %d: %.1f cycles per sweep (max %.1f), %d instr-equivalents -> %.2f cycles per FMA instruction\n", name, threads / 128,
           mean / iters, (double)mx / iters, per_iter, mean / iters / per_iter);
    delete[] h;
    (void)lanes;
}

int main() {
    float *in, *out; long long* cyc;
    cudaMalloc(&in, 1024 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 16 * 8);
    float h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = 0.001f * (float)((i * 37) % 101) + 0.01f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
#define ALL(T) \
        run<0, T>("FFMA rows + FMNMX clamp chain", 32, in, out, cyc, 144); \
        run<1, T>("FFMA rows + FFMA.SAT clamp", 32, in, out, cyc, 144); \
        run<2, T>("FFMA2 rows + FFMA.SAT clamp", 32, in, out, cyc, 72); \
        run<3, T>("FFMA only (144 / sweep)", 32, in, out, cyc, 144); \
        run<4, T>("FFMA2 only (72 / sweep)", 32, in, out, cyc, 72);
    ALL(128) ALL(256)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
