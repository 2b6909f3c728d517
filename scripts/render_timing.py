"""Device time of srl_sim_render (primitive-list kernel + raster kernel, CUDA events around the call) for full batches of both scenes,
with and without the per-tile primitive culling.  Run on the GPU box; the output goes to profiles/."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
import torch
from srl_sim._abi import load_cuda_library
from srl_sim.backend import Backend
from srl_sim.model import load_kuka_scene
from srl_sim.render import KUKA_CAMERA, KUKA_CAMERA_2, MOBILE_CAMERA, camera

be = Backend(load_cuda_library(), 0)
st = be.stream()


def timed(sim, cam, w, h, n, reps=10):
    buf = be.zeros((n, h, w, 3), np.uint8)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=be.torch_device)
    for _ in range(3):
        sim.render(camera(**cam), w, h, buf, stream=st)
    ms = []
    for _ in range(reps):
        flush.zero_()                                  # cold L2 for the primitive lists
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream()); sim.render(camera(**cam), w, h, buf, stream=st); e1.record(torch.cuda.current_stream())
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return float(np.median(ms)), buf


for env_id, n, cams in (("KukaButtonGymEnv-v0", 4096, (("camera 1", KUKA_CAMERA), ("camera 2", KUKA_CAMERA_2))), ("MobileRobotGymEnv-v0", 4096, (("top-down", MOBILE_CAMERA),))):
    kuka = env_id.startswith("Kuka")
    sim = be.make_sim(env_id, n, model_blob=load_kuka_scene().blob if kuka else None, seed=0, random_target=True)
    sim.reset(stream=st)
    acts = torch.randint(0, 6 if kuka else 4, (64, n), dtype=torch.int32, device=be.torch_device)
    o = be.zeros((64, n, sim.obs_dim), np.float32); r = be.zeros((64, n), np.float32); d = be.zeros((64, n), np.uint8)
    sim.rollout(64, acts, None, o, r, d, stream=st)
    for name, cam in cams:
        for (w, h) in ((224, 224), (64, 64)):
            os.environ.pop("SRL_RENDER_NO_CULL", None)
            ms, a = timed(sim, cam, w, h, n)
            os.environ["SRL_RENDER_NO_CULL"] = "1"
            ms0, b = timed(sim, cam, w, h, n)
            os.environ.pop("SRL_RENDER_NO_CULL", None)
            px = n * w * h
            print("%s %s, %d frames of %d x %d: %.3f ms (%.2f M frames/s, %.1f Gpixel/s, %.1f GB/s of RGB bytes written); without tile culling %.3f ms; same bytes: %s"
                  % (env_id, name, n, w, h, ms, n / ms / 1e3, px / ms / 1e6, 3 * px / ms / 1e6, ms0, bool(torch.equal(a, b))))
    sim.close()
