// Image observations for the batched simulator: one launch builds every env's primitive list from its state, one launch
// ray-casts all frames.  Per-pixel arithmetic and the scene lists live in render_core.h (shared with the CPU checker).
//
// Layout: the primitive lists are [N][SRL_MAX_PRIMS][16 floats] in HBM (2.5 KB per env); the raster kernel runs one CTA of 16 x 16 pixels per
// (tile, env), stages the env's list in shared memory once and writes RGB bytes row-major -- a 224 x 224 frame is 196 tiles, 4096 envs
// are 803 k CTAs, 617 MB of output per call: the kernel is bound by the intersection arithmetic (~30 primitives x ~40 flop per pixel).
#include "common.cuh"
#include "render_core.h"

namespace {

__global__ void mobile_prims_kernel(MobileDev d, int n, int kind, float* __restrict__ prims, int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 pos = d.pos[i], t0 = d.tgt0[i], t1 = d.tgt1[i];
    const int rk = kind == SRL_ENV_MOBILE_2TARGET ? 1 : kind == SRL_ENV_MOBILE_LINE_TARGET ? 2 : kind == SRL_ENV_MOBILE_1D ? 3 : 0;
    SrlPrim* out = reinterpret_cast<SrlPrim*>(prims + (size_t)i * SRL_MAX_PRIMS * SRL_PRIM_WORDS);
    counts[i] = srl_mobile_scene(rk, (float)pos.x, (float)pos.y, (float)t0.x, (float)t0.y, (float)t1.x, (float)t1.y, out);
}

__global__ void __launch_bounds__(256) raster_kernel(const float* __restrict__ prims, const int* __restrict__ counts, SrlCam cam, int W, int H,
                                                      uint8_t* __restrict__ rgb) {
    __shared__ SrlPrim sp[SRL_MAX_PRIMS];
    const int env = blockIdx.y, tiles_x = (W + 15) / 16;
    const int np = counts[env];
    const float* src = prims + (size_t)env * SRL_MAX_PRIMS * SRL_PRIM_WORDS;
    for (int k = threadIdx.x; k < np * SRL_PRIM_WORDS; k += blockDim.x) reinterpret_cast<float*>(sp)[k] = src[k];
    __syncthreads();
    const int x = (blockIdx.x % tiles_x) * 16 + (threadIdx.x & 15), y = (blockIdx.x / tiles_x) * 16 + (threadIdx.x >> 4);
    if (x >= W || y >= H) return;
    uint8_t px[3];
    srl_render_pixel(cam, sp, np, x, y, W, H, px);
    uint8_t* o = rgb + ((size_t)env * H * W + (size_t)y * W + x) * 3;
    o[0] = px[0]; o[1] = px[1]; o[2] = px[2];
}

}  // namespace

int render_launch(srl_sim* s, const srl_camera* cam, int width, int height, uint8_t* rgb, cudaStream_t st) {
    if (!s->render_prims) {
        SRL_CUDA_OK(cudaMalloc(&s->render_prims, (size_t)s->n * SRL_MAX_PRIMS * SRL_PRIM_WORDS * sizeof(float)));
        SRL_CUDA_OK(cudaMalloc(&s->render_counts, (size_t)s->n * sizeof(int)));
    }
    if (srl_is_mobile(s->kind)) {
        mobile_prims_kernel<<<(s->n + 127) / 128, 128, 0, st>>>(s->mob, s->n, s->kind, s->render_prims, s->render_counts);
        SRL_CUDA_OK(cudaGetLastError());
    } else if (kuka_render_prims(s, s->render_prims, s->render_counts, st)) return 1;
    SrlCam c;
    srl_camera_setup(cam->target, cam->distance, cam->yaw, cam->pitch, cam->roll, cam->fov, (float)width / (float)height, c);
    const dim3 grid(((width + 15) / 16) * ((height + 15) / 16), s->n);
    raster_kernel<<<grid, 256, 0, st>>>(s->render_prims, s->render_counts, c, width, height, rgb);
    SRL_CUDA_OK(cudaGetLastError());
    s->launches += 2;
    return 0;
}

void render_free(srl_sim* s) {
    if (s->render_prims) cudaFree(s->render_prims);
    if (s->render_counts) cudaFree(s->render_counts);
    s->render_prims = nullptr; s->render_counts = nullptr;
}
