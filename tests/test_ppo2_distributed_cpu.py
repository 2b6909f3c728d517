"""
Data-parallel PPO2 on CPU: world-size-2 `gloo`, the simulator behind the C-ABI replaced by the CPU oracle
(test infrastructure, installed explicitly through srl_sim.backend.use_library).  Checks the two collectives of the trainer:
the per-minibatch gradient all-reduce keeps the replicas bit-identical, and the per-rollout merge of the observation
filter equals a single filter that saw every rank's batches.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ORACLE_LIB, PKG


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, outdir):
    import sys
    sys.path.insert(0, PKG)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from srl_sim import backend
    from srl_sim._abi import SimLibrary
    backend.use_library(SimLibrary(ORACLE_LIB), -1)
    from rl_baselines import ppo2
    hist = ppo2.train("MobileRobotGymEnv-v0", 8, 8 * 16 * 2 * 3, seed=3, env_kwargs=dict(is_discrete=True, shape_reward=True, max_steps=20),
                      hyperparams=dict(n_steps=16), verbose=0, cuda_graph=False, log_dir=os.path.join(outdir, "log"), device=None)
    policy, norm = ppo2.train.last_policy, ppo2.train.last_norm
    flat = torch.cat([p.detach().reshape(-1) for p in policy.parameters()]).numpy()
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), params=flat, mean=norm.mean.numpy(), var=norm.var.numpy(), count=norm.count.numpy(),
             steps=[h[0] for h in hist], ret=[h[1] for h in hist])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ppo2_keeps_replicas_identical(tmp_path, oracle_lib):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert np.array_equal(a["params"], b["params"])                    # averaged gradients => bit-identical replicas
    for k in ("mean", "var", "count"):
        assert np.array_equal(a[k], b[k]), k                           # the merged filter is the same on both ranks
    # 3 updates x 2 ranks x 8 envs x 16 steps, +1 reset batch per rank (the filter also sees the reset observation) + the 1e-4 prior
    assert float(a["count"]) == pytest.approx(3 * 2 * 8 * 16 + 2 * 8 + 1e-4)
    assert list(a["steps"]) == [256, 512, 768]                         # global env steps
    assert np.array_equal(a["ret"], b["ret"], equal_nan=True)                            # all-gathered episode statistics
    assert np.isfinite(a["ret"][-1])                                   # max_steps=20 => 21-step episodes finish from the second rollout on
    assert os.path.isfile(os.path.join(str(tmp_path), "log", "ppo2_model.pt"))   # rank 0 alone writes the run directory


def test_merge_running_moments_equals_one_filter_over_all_batches():
    import sys
    sys.path.insert(0, PKG)
    from rl_baselines.ppo2 import RunningNorm, merge_running_moments
    rng = np.random.default_rng(0)
    dev = torch.device("cpu")
    prior_batch = torch.from_numpy(rng.normal(1.0, 2.0, (50, 3)))
    batches = [[torch.from_numpy(rng.normal(r, 1.0 + r, (16, 3))) for _ in range(5)] for r in range(3)]
    ranks = []
    for r in range(3):
        f = RunningNorm(3, dev); f.update(prior_batch)
        ranks.append(f)
    prior = (ranks[0].mean.clone(), ranks[0].var.clone(), ranks[0].count.clone())
    for f, bs in zip(ranks, batches):
        for x in bs:
            f.update(x)
    # a sum over "ranks" that runs in-process: the all-reduce callback adds the other ranks' statistics
    def make_all_reduce(me):
        def all_reduce(s):
            for j, g in enumerate(ranks_stats):
                if j != me:
                    s.add_(g)
        return all_reduce
    ranks_stats = [torch.cat([f.count.reshape(1), f.count * f.mean, f.count * (f.var + f.mean ** 2)]) for f in ranks]
    for i, f in enumerate(ranks):
        merge_running_moments(f, prior, make_all_reduce(i), world=3)
    whole = RunningNorm(3, dev); whole.update(prior_batch)
    whole.update(torch.cat([x for bs in batches for x in bs]))
    for f in ranks:
        assert torch.allclose(f.mean, whole.mean, rtol=0, atol=1e-12)
        assert torch.allclose(f.var, whole.var, rtol=1e-12, atol=1e-12)
        assert float(f.count) == pytest.approx(float(whole.count), rel=1e-14)


def test_single_process_ppo2_runs_on_the_oracle_backend(use_oracle_backend):
    """The trainer's eager path (no CUDA graphs) end to end on CPU: collection, GAE, minibatch steps, phase timing, history."""
    from rl_baselines import ppo2
    pt = {}
    hist = ppo2.train("MobileRobotGymEnv-v0", 16, 16 * 32 * 3, seed=1, env_kwargs=dict(is_discrete=True, shape_reward=True, max_steps=20),
                      hyperparams=dict(n_steps=32), verbose=0, device=None, phase_times=pt)
    assert [h[0] for h in hist] == [512, 1024, 1536]
    assert all(np.isfinite(h[1]) and h[1] < 0 for h in hist)           # shaped reward = -distance: every finished episode has a negative return
    assert set(pt) == {"collect", "gae", "optimise"} and all(v > 0 for v in pt.values())
    hist_c = ppo2.train("MobileRobotGymEnv-v0", 8, 8 * 16 * 2, seed=2, env_kwargs=dict(is_discrete=False, max_steps=20), hyperparams=dict(n_steps=16),
                        verbose=0, device=None)                        # continuous actions: the Normal policy head
    assert len(hist_c) == 2


def test_fused_policy_step_has_no_cpu_fallback(use_oracle_backend):
    """fused_act=True is the sm_100a kernels of include/srl_policy.h or nothing: on the CPU oracle backend the trainer refuses."""
    from rl_baselines import ppo2
    with pytest.raises(ValueError, match="no CPU fallback"):
        ppo2.train("MobileRobotGymEnv-v0", 8, 8 * 16, seed=0, hyperparams=dict(n_steps=16), verbose=0, device=None, fused_act=True)


def test_train_entry_point_flags_on_the_oracle_backend(use_oracle_backend, tmp_path):
    """`python -m rl_baselines.train` with the reference's flag names (train.py:177-208): `--hyperparam name:value` pairs are typed and
    checked like base_classes.parserHyperParam does, `-joints` needs `-c`, the run directory gets args.json / env_globals.json / the model."""
    import glob
    import json
    from rl_baselines.train import main, parserHyperParam
    assert parserHyperParam(["n_steps:16", "gamma:0.9"]) == {"n_steps": 16, "gamma": 0.9}
    with pytest.raises(AssertionError, match="not in list of valid hyperparameters"):
        parserHyperParam(["batch:3"])
    with pytest.raises(ValueError, match="continuous only"):
        main(["--env", "KukaButtonGymEnv-v0", "-joints", "--device", "-1"])
    hist = main(["--algo", "ppo2", "--env", "MobileRobotGymEnv-v0", "--num-cpu", "8", "--num-timesteps", "300", "--hyperparam", "n_steps:16", "noptepochs:2",
                 "--episode_window", "5", "-r", "--shape-reward", "--log-dir", str(tmp_path), "--device", "-1", "--seed", "4"])
    assert [h[0] for h in hist] == [128, 256]                        # 1.1 x 300 steps (train.py:319) in updates of 8 envs x 16 steps
    run = glob.glob(os.path.join(str(tmp_path), "MobileRobotGymEnv-v0", "ground_truth", "ppo2", "*"))[0]
    args = json.load(open(os.path.join(run, "args.json")))
    assert args["n_steps"] == 16 and args["noptepochs"] == 2 and args["num_cpu"] == 8 and args["seed"] == 4
    assert json.load(open(os.path.join(run, "env_globals.json")))["random_target"] is True
    assert os.path.isfile(os.path.join(run, "ppo2_model.pt"))


def test_monitor_log_best_model_callback_and_obs_filter_file(use_oracle_backend, tmp_path):
    """SURVEY 8(f).1 pieces of the trainer, on the oracle backend: a per-process `<rank>.monitor.csv` in the bench.Monitor format that the
    reference's loadCsv reads (/root/reference/rl_baselines/visualize.py:59-107), the best-model callback
    (/root/reference/rl_baselines/train.py:132-159: mean of the last 100 episodes, saved when it improves and 100 episodes exist) and the
    observation filter in VecNormalize's file format, loadable through createEnvs(load_path_normalise=...)."""
    import json
    import types
    import torch
    from rl_baselines.ppo2 import train
    from rl_baselines.utils import createEnvs
    from srl_sim.monitor import load_monitor_csv
    n, updates = 64, 10
    hist = train("MobileRobotGymEnv-v0", n, n * 128 * updates, seed=3, env_kwargs=dict(is_discrete=True, shape_reward=True), log_dir=str(tmp_path),
                 verbose=0, cuda_graph=False, device=None)
    assert len(hist) == updates
    # ---- monitor file: header, r,l,t rows; every MobileRobot episode lasts 251 steps; loadCsv's view of it ----
    lines = open(str(tmp_path / "0.monitor.csv")).read().split("\n")
    assert json.loads(lines[0][1:])["env_id"] == "MobileRobotGymEnv-v0" and lines[1] == "r,l,t"
    result, timesteps = load_monitor_csv(str(tmp_path))
    n_ep = n * (128 * updates // 251)
    assert len(result) == n_ep and timesteps == 251 * n_ep
    assert [r[0] for r in result[:3]] == [0, 251, 502]
    rows = [l.split(",") for l in lines[2:] if l]
    assert all(int(r[1]) == 251 for r in rows) and all(float(a[2]) <= float(b[2]) for a, b in zip(rows, rows[1:]))
    # ---- best-model callback ----
    meta = json.load(open(str(tmp_path / "best_model.json")))
    assert meta["saves"] >= 1 and meta["min_episodes_before_save"] == 100
    best = torch.load(str(tmp_path / "ppo2_model.pt"))
    final = torch.load(str(tmp_path / "ppo2_model_final.pt"))
    assert set(best) == set(final) == {"policy", "obs_mean", "obs_var", "obs_count"}
    assert abs(meta["best_mean_reward"] - train.best_mean_reward) < 1e-9
    y = [r[1] for r in result]
    assert meta["best_mean_reward"] <= max(np.mean(y[max(0, k - 100):k]) for k in range(100, len(y) + 1)) + 1e-6
    # ---- observation filter through the reference's load path ----
    args = types.SimpleNamespace(env="MobileRobotGymEnv-v0", num_cpu=2, seed=0, num_stack=1, srl_model="ground_truth")
    envs = createEnvs(args, env_kwargs=dict(is_discrete=True), load_path_normalise=str(tmp_path))
    assert np.allclose(envs.obs_rms.mean, final["obs_mean"].cpu().numpy()) or np.allclose(envs.obs_rms.mean, best["obs_mean"].cpu().numpy())
    assert envs.obs_rms.count > 64 * 128
    envs.close()
    with pytest.raises(ValueError):
        train("MobileRobotGymEnv-v0", 4, 4 * 128, env_kwargs=dict(srl_model="joints"), verbose=0, cuda_graph=False, device=None)
