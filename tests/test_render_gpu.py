"""
Image observations on the GPU: the sm_100a ray-caster (csrc/render_kernels.cu) against the CPU checker (same primitive lists and
per-pixel arithmetic, csrc/render_core.h; the checker itself is pinned against the reference's frames and an independent camera restatement
in tests/test_render_cpu.py), for every registered env id, plus the batched raw_pixels VecEnv and the throughput of a full batch.
"""
import time

import numpy as np
import pytest

from srl_sim import _abi
from srl_sim.model import load_kuka_scene
from srl_sim.render import KUKA_CAMERA, KUKA_CAMERA_2, MOBILE_CAMERA, camera, mobile_fpv_camera

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("env_id", sorted(_abi.ENV_KINDS))
def test_cuda_frames_match_the_cpu_checker(cuda_backend, oracle_backend, env_id):
    """Same seeds, same actions, a handful of steps (float32 kernel vs float64 oracle state: the joint frames differ by ~1e-6 m, so a few
    silhouette pixels may fall on the other side of an edge): at least 99.5 % of the bytes identical, every frame non-trivial."""
    n, T = 6, 12
    kuka = env_id.startswith("Kuka")
    cfg = dict(seed=3, random_target=True)
    if env_id in ("MobileRobot2TargetGymEnv-v0", "MobileRobot1DGymEnv-v0"):
        cfg["is_discrete"] = True
    rs = np.random.RandomState(1)
    n_act = 6 if kuka else (2 if env_id == "MobileRobot1DGymEnv-v0" else 4)
    acts = rs.randint(0, n_act, size=(T, n)).astype(np.int32)
    cams = [KUKA_CAMERA, KUKA_CAMERA_2] if kuka else [MOBILE_CAMERA, mobile_fpv_camera((2.0, 2.0))]
    frames = {}
    for tag, be in (("cuda", cuda_backend), ("oracle", oracle_backend)):
        sim = be.make_sim(env_id, n, model_blob=load_kuka_scene().blob if kuka else None, **cfg)
        sim.reset(stream=be.stream())
        obs = be.zeros((T, n, sim.obs_dim), np.float32); rew = be.zeros((T, n), np.float32); done = be.zeros((T, n), np.uint8)
        sim.rollout(T, be.from_host(acts), None, obs, rew, done, stream=be.stream())
        out = []
        for c in cams:
            buf = be.zeros((n, 224, 224, 3), np.uint8)
            sim.render(camera(**c), 224, 224, buf, stream=be.stream())
            out.append(be.to_host(buf).copy())
        small = be.zeros((n, 33, 50, 3), np.uint8)               # a non-square, non-multiple-of-16 size
        sim.render(camera(**cams[0]), 50, 33, small, stream=be.stream())
        out.append(be.to_host(small).copy())
        frames[tag] = out
        sim.close()
    for a, b in zip(frames["cuda"], frames["oracle"]):
        same = (a == b).mean()
        assert same > 0.995, (env_id, same)
        assert np.abs(a.astype(int) - b.astype(int)).mean() < 0.5
        for k in range(n):
            assert len(np.unique(a[k].reshape(-1, 3), axis=0)) >= 3      # not a blank frame (a MobileRobot frame has 4 to 8 flat colours)
    assert not np.array_equal(frames["cuda"][0][0], frames["cuda"][0][1])   # different envs, different frames


@pytest.mark.parametrize("env_id", ["KukaButtonGymEnv-v0", "Kuka2ButtonGymEnv-v0", "MobileRobotGymEnv-v0", "MobileRobotLineTargetGymEnv-v0"])
def test_tile_culling_never_changes_a_byte(cuda_backend, env_id, monkeypatch):
    """The raster kernel drops, per 32 x 8 tile, the primitives whose bounding sphere cannot reach the tile's rays.  The test is conservative,
    so the frames with and without it (SRL_RENDER_NO_CULL) are the same bytes: many envs, mid-episode states, every camera, odd sizes."""
    be = cuda_backend
    n, T = 64, 40
    kuka = env_id.startswith("Kuka")
    sim = be.make_sim(env_id, n, model_blob=load_kuka_scene().blob if kuka else None, seed=11, random_target=True)
    sim.reset(stream=be.stream())
    acts = np.random.RandomState(5).randint(0, 6 if kuka else 4, size=(T, n)).astype(np.int32)
    obs = be.zeros((T, n, sim.obs_dim), np.float32); rew = be.zeros((T, n), np.float32); done = be.zeros((T, n), np.uint8)
    sim.rollout(T, be.from_host(acts), None, obs, rew, done, stream=be.stream())
    cams = [KUKA_CAMERA, KUKA_CAMERA_2, dict(KUKA_CAMERA, distance=0.6, pitch=-10.0, yaw=200.0)] if kuka else \
        [MOBILE_CAMERA, mobile_fpv_camera((1.0, 3.0)), dict(MOBILE_CAMERA, distance=2.0, pitch=-30.0, yaw=40.0)]
    for c in cams:
        for (w, h) in ((224, 224), (64, 64), (50, 33), (96, 40)):
            out = []
            for no_cull in (False, True):
                if no_cull:
                    monkeypatch.setenv("SRL_RENDER_NO_CULL", "1")
                else:
                    monkeypatch.delenv("SRL_RENDER_NO_CULL", raising=False)
                buf = be.zeros((n, h, w, 3), np.uint8)
                sim.render(camera(**c), w, h, buf, stream=be.stream())
                out.append(be.to_host(buf).copy())
            assert np.array_equal(out[0], out[1]), (env_id, c, w, h, int((out[0] != out[1]).sum()))
    monkeypatch.delenv("SRL_RENDER_NO_CULL", raising=False)
    sim.close()


def test_batched_raw_pixels_vec_env_and_throughput(cuda_lib):
    from srl_sim import backend
    backend.use_library(None, None)
    import torch
    from srl_sim.vec_env import BatchedSRLVecEnv
    venv = BatchedSRLVecEnv("KukaButtonGymEnv-v0", 8, seed=1, srl_model="raw_pixels", multi_view=True)
    o = venv.reset()
    assert o.shape == (8, 224, 224, 6) and o.dtype == np.uint8
    o2, r, d, _ = venv.step([0] * 8)
    assert o2.shape == o.shape
    t = venv.render_tensors()
    assert t.is_cuda and t.dtype == torch.uint8 and tuple(t.shape) == (8, 224, 224, 6)
    venv.close()
    # a full batch: 4096 Kuka envs x one 224 x 224 frame (617 MB of output)
    venv = BatchedSRLVecEnv("KukaButtonGymEnv-v0", 4096, seed=1, srl_model="ground_truth")
    venv.reset()
    for _ in range(2):
        venv.render_tensors()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        f = venv.render_tensors()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print("RENDER THROUGHPUT: 4096 Kuka frames of 224 x 224 in %.2f ms = %.2f M frames/s, %.1f GB/s of pixels" % (1e3 * dt, 4096 / dt / 1e6, f.numel() / dt / 1e9))
    assert tuple(f.shape) == (4096, 224, 224, 3)
    venv.close()
