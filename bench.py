#!/usr/bin/env python
"""
bench.py -- env-steps/sec of the B200-native batched simulator (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload kuka|mobile]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Default workload = BASELINE.json configs[1]: KukaButtonGymEnv-v0, ground_truth, 4096 envs per GPU,
synthetic random discrete actions + N(0, 0.01) step noise.  One bench "step" = ONE fused rollout of
T = 128 env steps over the whole batch (the n_steps of the reference's PPO2 runner,
rl_baselines/rl_algorithm/ppo2.py:58-72): a single kernel launch, 4096 x 128 env-steps.

  value : env-steps/s, actions/noise already resident in HBM, outputs left in HBM (CUDA events, max over ranks)
  e2e   : the same metric through the host-facing C-ABI call (srl_sim_rollout_host): pinned HOST action/noise
          buffers in, pinned HOST obs/reward/done out, copies inside the timed region
  roofline     : algorithmic HBM bytes of one launch / device time of that launch vs the measured copy peak
  cpu_baseline : the CPU oracle (double precision, oracle/liboracle_sim.so, kind "port") on the host cores
  --impl reference : the reference arm.  PyBullet is not installable here, so it times the oracle -- the CPU
          restatement of the reference's step -- with every host thread, on the same config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))

KUKA_ENVS_PER_GPU = 4096      # BASELINE.json configs[1]
MOBILE_ENVS_PER_GPU = 8192    # BASELINE.json configs[3]
# Algorithmic HBM bytes (DESIGN.md "Measurement"): SoA state in + out once per launch, per-step I/O floor.
KUKA_STATE_BYTES = 2 * 224            # 12 float4 + 2 int4 records, read + written once per launch
KUKA_STEP_BYTES = 4 + 4 + 12 + 4 + 1  # action i32 + noise f32 in, obs f32[3] + reward f32 + done u8 out
KUKA_WARP_INST_PER_LAUNCH = 3904811262  # ncu smsp__inst_executed.sum of one 4096-env x 128-step launch (profiles/r01_kuka_kernel_ncu_full.txt)
KUKA_FLOP_PER_STEP = 1.0e5            # ~150 PGS sweeps x 13 rows x 2 x 13 + dynamics
MOBILE_STATE_BYTES = 2 * 80
MOBILE_STEP_BYTES = 4 + 8 + 4 + 1     # action in, obs f32[2] + reward + done out (in-kernel actions: no action read)


def effective_cores():
    """Host threads that can actually run: min(visible CPUs, cgroup CPU quota).  The GPU boxes show 128 CPUs but the
    container's cgroup grants 16 CPUs of time (cpu.max = 1600000 100000); the oracle scales linearly to 16 threads and is
    flat beyond (scripts/cpu_scaling.py), so that is the core count reported with the CPU numbers."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(round(float(quota) / float(period)))))
    except Exception:
        pass
    return n


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-i", str(self.gpu_index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        mx = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def make_inputs(workload, n, T, seed):
    rng = np.random.default_rng(seed)
    if workload == "kuka":
        acts = rng.integers(0, 6, (T, n), dtype=np.int32)
        noise = rng.normal(0, 0.01, (T, n)).astype(np.float32)
        return acts, noise
    return rng.integers(0, 4, (T, n), dtype=np.int32), None


def workload_spec(workload):
    if workload == "kuka":
        return dict(env_id="KukaButtonGymEnv-v0", n=KUKA_ENVS_PER_GPU, T=128, obs_dim=3, state_bytes=KUKA_STATE_BYTES,
                    step_bytes=KUKA_STEP_BYTES, cfg=dict(is_discrete=True, random_target=False, force_down=True, action_repeat=1, max_distance=0.8))
    return dict(env_id="MobileRobotGymEnv-v0", n=MOBILE_ENVS_PER_GPU, T=1024, obs_dim=2, state_bytes=MOBILE_STATE_BYTES,
                step_bytes=MOBILE_STEP_BYTES, cfg=dict(is_discrete=True, random_target=True))


def metric_and_config(workload, world):
    """`metric` / `config` of the JSON line -- shared by the b200 arm and the reference arm (the driver compares the two lines)."""
    spec = workload_spec(workload)
    n, T = spec["n"], spec["T"]
    metric = "env-steps/sec %s ground_truth @%d envs/GPU" % (spec["env_id"], n)
    config = {"workload": "%s ground_truth, %d envs/GPU, one bench step = one fused rollout of T=%d env steps (random discrete actions%s)"
                          % (spec["env_id"], n, T, " + N(0,0.01) step noise" if workload == "kuka" else ""),
              "envs_per_gpu": n, "env_steps_per_bench_step": n * T,
              "l2_flush_between_steps": "write 256 MB + read 256 MB between timed steps, outside the event bracket",
              "parallelism": "env-shard x%d" % world}
    return metric, config


def model_blob(workload):
    if workload != "kuka":
        return None
    from srl_sim.model import load_kuka_scene
    return load_kuka_scene().blob


# ------------------------------------------------------------------------------- CPU oracle legs --------
def _oracle_backend():
    from srl_sim._abi import SimLibrary
    from srl_sim.backend import Backend
    path = os.path.join(ROOT, "oracle", "liboracle_sim.so")
    if not os.path.isfile(path):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return Backend(SimLibrary(path), -1)


class OraclePool(object):
    """The CPU oracle sharded over `threads` host threads (ctypes releases the GIL), one sim handle per thread --
    the same one-env-group-per-worker shape as the reference's SubprocVecEnv (rl_baselines/utils.py:216-220)."""

    def __init__(self, workload, n_total, T, threads, seed=0):
        spec = workload_spec(workload)
        be = _oracle_backend()
        self.T, self.n_total, self.threads = T, n_total, max(1, min(threads, n_total))
        blob = model_blob(workload)
        bounds = np.linspace(0, n_total, self.threads + 1).astype(int)
        self.parts = []
        acts, noise = make_inputs(workload, n_total, T, seed + 100)
        for k in range(self.threads):
            lo, hi = int(bounds[k]), int(bounds[k + 1])
            if hi == lo:
                continue
            sim = be.make_sim(spec["env_id"], hi - lo, seed=seed, model_blob=blob, global_env_offset=lo, **spec["cfg"])
            sim.reset()
            a = np.ascontiguousarray(acts[:, lo:hi]); nz = None if noise is None else np.ascontiguousarray(noise[:, lo:hi])
            obs = np.zeros((T, hi - lo, spec["obs_dim"]), np.float32); rew = np.zeros((T, hi - lo), np.float32)
            done = np.zeros((T, hi - lo), np.uint8)
            self.parts.append((sim, a, nz, obs, rew, done))

    def step(self):
        ths = [threading.Thread(target=lambda p=p: p[0].rollout(self.T, p[1], p[2], p[3], p[4], p[5])) for p in self.parts]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0


def cpu_baseline(workload, cores):
    spec = workload_spec(workload)
    # bounded sample: a slice of the same workload worth ~10-30 s of single-core work
    T = spec["T"] if workload == "kuka" else 256
    n = spec["n"] if workload == "kuka" else spec["n"]
    pool = OraclePool(workload, n, T, cores)
    pool.step()  # warm-up (page in, caches)
    dt = min(pool.step() for _ in range(2))
    return {"value": n * T / dt, "unit": "env-steps/s", "cores": pool.threads, "kind": "port",
            "sample": "%d envs x %d steps of %s, oracle/liboracle_sim.so (float64 CPU restatement, no rendering, no Python in the "
                      "loop), %d host threads" % (n, T, spec["env_id"], pool.threads)}


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's own step on all host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = workload_spec(args.workload)
    cores = effective_cores()
    n, T = spec["n"], spec["T"] if args.workload == "kuka" else 256
    pool = OraclePool(args.workload, n, T, cores)
    for _ in range(args.warmup):
        pool.step()
    times = [pool.step() for _ in range(args.steps)]
    total = sum(times)
    value = n * T * args.steps / total
    sample = ("%d envs x %d steps per step, %d host threads (container CPU quota; %d CPUs visible), CPU oracle "
              "(PyBullet itself is not installable offline)" % (n, T, pool.threads, os.cpu_count() or 1))
    metric, config = metric_and_config(args.workload, 1)
    config["parallelism"] = "%d host threads, one env shard each (rank 0 only)" % pool.threads
    config.pop("l2_flush_between_steps")           # a CPU run: nothing to flush
    if T != spec["T"]:
        config["reference_sample"] = "bounded sample: %d of the %d env steps per bench step" % (T, spec["T"])
    line = {"impl": "reference", "metric": metric, "value": value, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": pool.threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------- GPU arm -------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from srl_sim._abi import load_cuda_library
    from srl_sim.backend import Backend

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    be = Backend(load_cuda_library(), local_rank)
    spec = workload_spec(args.workload)
    n, T, D = spec["n"], spec["T"], spec["obs_dim"]
    sim = be.make_sim(spec["env_id"], n, seed=args.seed, model_blob=model_blob(args.workload), global_env_offset=rank * n, **spec["cfg"])
    st = be.stream()
    sim.reset(stream=st)
    acts_h, noise_h = make_inputs(args.workload, n, T, args.seed + 1000 * rank)
    acts = be.from_host(acts_h)
    noise = None if noise_h is None else be.from_host(noise_h)
    obs = be.zeros((T, n, D), np.float32); rew = be.zeros((T, n), np.float32); done = be.zeros((T, n), np.uint8)
    ep_ret = be.zeros((T, n), np.float32); ep_len = be.zeros((T, n), np.int32)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=be.torch_device)  # > 126 MB L2
    flush_rd = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=be.torch_device)  # 256 MB, only ever read
    sink = torch.zeros((), dtype=torch.float32, device=be.torch_device)

    def step():
        sim.rollout(T, acts, noise, obs, rew, done, ep_ret, ep_len, stream=st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = sim.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        # evict L2 between timed iterations (outside the event bracket): WRITE a buffer larger than L2, then READ another
        # one so that the timed kernel starts from a cold L2 holding clean lines -- otherwise it would also pay the DRAM
        # write-back of up to 126 MB of the flush buffer's dirty lines, which is not its traffic
        flush.fill_(k & 0xff)
        sink += flush_rd.sum()
        ev[k][0].record()
        step()
        ev[k][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    launches = sim.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host buffers through srl_sim_rollout_host ----
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h_acts = pin(acts_h); h_noise = None if noise_h is None else pin(noise_h)
    h_obs = torch.empty((T, n, D), dtype=torch.float32).pin_memory(); h_rew = torch.empty((T, n), dtype=torch.float32).pin_memory()
    h_done = torch.empty((T, n), dtype=torch.uint8).pin_memory()

    def e2e_step():
        sim.rollout_host(T, h_acts, h_noise, h_obs, h_rew, h_done)

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = h_acts.numel() * 4 + (0 if h_noise is None else h_noise.numel() * 4)
    d2h = h_obs.numel() * 4 + h_rew.numel() * 4 + h_done.numel()

    # ---- max over ranks; optional cross-rank episode-return all-gather (the only collective; off the step path) ----
    tms = torch.tensor([total_ms, e2e_s * 1e3, wall * 1e3], device=be.torch_device, dtype=torch.float64)
    d = done.bool()
    ep_stats = torch.stack([ep_ret[d].sum().double(), d.sum().double()])
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(ep_stats) for _ in range(world)]
        dist.all_gather(gathered, ep_stats)
        ep_stats = torch.stack(gathered).sum(0)
    total_ms, e2e_ms, wall_ms = [float(x) for x in tms.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = _peaks()
    units = n * T * world
    value = units * args.steps / (total_ms * 1e-3)
    launch_bytes = n * (spec["state_bytes"] + T * spec["step_bytes"])
    launch_s = (total_ms / args.steps) * 1e-3
    achieved = launch_bytes / launch_s / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "traffic": TRAFFIC_BYTES.get(args.workload), "peak_source": "%s (MEASURED_PEAKS.json hbm_gbs)" % peak_src,
            "algorithmic_bytes_per_launch": launch_bytes, "kernel": "kuka_kernel" if args.workload == "kuka" else "mobile_rollout_kernel"}
    if args.workload == "kuka":
        fl = n * T * KUKA_FLOP_PER_STEP / launch_s / 1e12
        roof["note"] = ("latency/issue-bound fp32 kernel (150 strictly sequential PGS sweeps per env-step), not HBM-bound: "
                        "%.2f TFLOP/s of useful fp32 work; see DESIGN.md 'Measurement'" % fl)
        roof["fp32_tflops"] = fl
        # what actually bounds this kernel: warp-instruction issue slots (one per scheduler per cycle, 4 schedulers per SM).
        # Instructions per launch come from the committed ncu capture (like `traffic`), the time is this run's.
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        clk = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
        issue_peak = sms * 4 * clk * 1e-3
        issue_ach = KUKA_WARP_INST_PER_LAUNCH / launch_s / 1e9
        roof["issue"] = {"achieved": issue_ach, "peak": issue_peak, "unit": "G warp-inst/s", "frac": issue_ach / issue_peak,
                         "warp_inst_per_launch": KUKA_WARP_INST_PER_LAUNCH,
                         "source": "smsp__inst_executed.sum of profiles/r01_kuka_kernel_ncu_full.txt; peak = SMs x 4 schedulers x SM clock"}
    metric, config = metric_and_config(args.workload, world)
    line = {"metric": metric, "value": value, "unit": "env-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.workload == "kuka" else "f64", "data": "synthetic",
            "config": config,
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": units * args.steps / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "roofline": roof,
            "wall_ms_total_incl_flush": wall_ms,
            "episodes_finished": int(ep_stats[1].item()),
            "episode_return_mean": float(ep_stats[0].item() / max(1.0, ep_stats[1].item()))}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.workload, effective_cores())
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed
# `ncu --set full` capture (profiles/); filled in per round, None until measured.
TRAFFIC_BYTES = {"kuka": 108501760, "mobile": 86471680}  # profiles/r01_*_ncu_full.txt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="kuka", choices=["kuka", "mobile"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
