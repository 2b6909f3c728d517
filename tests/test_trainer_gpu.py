"""Consumers of the hot path on the GPU: the PPO2 restatement learns on the batched simulator, and the
reference-shaped entry point `python -m rl_baselines.train` runs for every registered env id (tests/test_pipeline.py:95-111
of the reference asserts exactly that: exit code 0 after 1600 steps with --num-cpu 4)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cuda_graph", [True, False])
def test_ppo2_learns_mobile_robot(cuda_lib, cuda_graph):
    """cuda_graph=True replays the captured 128-step collection loop (policy + simulator launches) once per update;
    cuda_graph=False issues the same launches eagerly.  Both must learn."""
    from srl_sim import backend
    backend.use_library(None, None)
    from rl_baselines.ppo2 import train
    hist = train("MobileRobotGymEnv-v0", 1024, 1024 * 128 * 12, seed=0, env_kwargs=dict(is_discrete=True, shape_reward=True), verbose=0, cuda_graph=cuda_graph)
    rets = [h[1] for h in hist if np.isfinite(h[1])]
    # shaped reward = -distance per step over 251 steps: a random policy scores about -420; learning must clearly beat it
    assert rets[-1] > rets[0] + 60, rets


@pytest.mark.parametrize("env_id", ["KukaButtonGymEnv-v0", "KukaRandButtonGymEnv-v0", "KukaMovingButtonGymEnv-v0", "MobileRobotGymEnv-v0", "MobileRobot2TargetGymEnv-v0",
                                    "MobileRobot1DGymEnv-v0", "MobileRobotLineTargetGymEnv-v0"])
def test_train_entry_point_runs_for_every_env(env_id, cuda_lib, tmp_path):
    from srl_sim import backend
    backend.use_library(None, None)
    from rl_baselines.train import main
    hist = main(["--algo", "ppo2", "--env", env_id, "--num-cpu", "4", "--num-timesteps", "1600", "--log-dir", str(tmp_path)])
    assert len(hist) >= 1
    fps = main(["--algo", "random_agent", "--env", env_id, "--num-cpu", "4", "--num-timesteps", "400"])
    assert fps > 0


def test_train_logs_reload_and_enjoy(cuda_lib, tmp_path):
    """SURVEY 8(f).1 end to end on the GPU: rl_baselines.ppo2.train on the device-resident VecEnv (built by createTensorEnvs, records for
    the lockstep step on) writes args.json / env_globals.json, a bench.Monitor-format `0.monitor.csv` fed by the kernel's episode
    statistics, the best model + observation filter whenever the mean of the last 100 episodes improves
    (/root/reference/rl_baselines/train.py:132-159); the log is parsed the way the reference's loadCsv does
    (/root/reference/rl_baselines/visualize.py:59-107) and replay.enjoy_baselines replays the BEST model."""
    import json
    import torch
    from srl_sim import backend
    backend.use_library(None, None)
    from rl_baselines.ppo2 import train
    from srl_sim.monitor import load_monitor_csv
    n, updates = 512, 8
    hist = train("KukaButtonGymEnv-v0", n, n * 128 * updates, seed=1, env_kwargs=dict(is_discrete=True, max_steps=60), log_dir=str(tmp_path), verbose=0,
                 prefetch_resets=True, fused_act=True)
    assert len(hist) == updates
    result, timesteps = load_monitor_csv(str(tmp_path))
    assert len(result) >= n * updates and result[0][0] == 0 and 0 < timesteps <= n * 128 * updates     # 61-step episodes at most
    rows = [l.split(",") for l in open(str(tmp_path / "0.monitor.csv")).read().split("\n")[2:] if l]
    assert all(1 <= int(r[1]) <= 61 for r in rows) and all(float(r[0]) == int(float(r[0])) for r in rows)   # sparse reward: integer returns
    meta = json.load(open(str(tmp_path / "best_model.json")))
    assert meta["saves"] >= 1
    best = torch.load(str(tmp_path / "ppo2_model.pt"))
    assert {"policy", "obs_mean", "obs_var", "obs_count"} <= set(best) and os.path.isfile(str(tmp_path / "obs_rms.pkl"))
    from replay.enjoy_baselines import main
    n_done, mean_reward = main(["--log-dir", str(tmp_path), "--num-cpu", "16", "--num-timesteps", "1100"])
    assert n_done >= 16 and np.isfinite(mean_reward)


def test_enjoy_replays_a_trained_agent(cuda_lib, tmp_path):
    """replay.enjoy_baselines reloads args.json / env_globals.json / the saved policy and observation filter of a run
    written by rl_baselines.train and reports finished episodes (the reference's test_enjoy.py asserts the exit code)."""
    from srl_sim import backend
    backend.use_library(None, None)
    from rl_baselines.ppo2 import train
    train("MobileRobotGymEnv-v0", 256, 256 * 128 * 2, seed=1, env_kwargs=dict(is_discrete=True), log_dir=str(tmp_path), verbose=0)
    from replay.enjoy_baselines import main
    n_done, mean_reward = main(["--log-dir", str(tmp_path), "--num-cpu", "16", "--num-timesteps", "300"])
    assert n_done == 16 and np.isfinite(mean_reward)       # every MobileRobot episode lasts 251 steps
    n_done, _ = main(["--log-dir", str(tmp_path), "--num-cpu", "4", "--num-timesteps", "260", "--deterministic", "--shape-reward"])
    assert n_done == 4


def test_dataset_generator_and_record_data_on_the_cuda_backend(cuda_lib, tmp_path):
    """SURVEY 8(f).3 on the product path: `python -m environments.dataset_generator` drives single-env objects (N = 1 views on the CUDA
    simulator) with record_data=True, EpisodeSaver writes the two npz schemas (/root/reference/state_representation/episode_saver.py:139-162);
    the recorded MobileRobot states must be the ones the oracle backend records for the same seeds (bit-exact: the env has no physics), and a
    recorded Kuka episode must be consistent with itself (positions within the fp32 tolerance of the oracle's recording)."""
    from srl_sim import backend
    from srl_sim._abi import SimLibrary
    from conftest import ORACLE_LIB
    from environments import dataset_generator
    out = {}
    for tag, lib, dev in (("cuda", None, None), ("oracle", SimLibrary(ORACLE_LIB), -1)):
        backend.use_library(lib, dev)
        try:
            base = str(tmp_path / tag) + "/"
            os.makedirs(base, exist_ok=True)
            n = dataset_generator.main(["--env", "MobileRobotGymEnv-v0", "--num-episode", "3", "--save-path", base, "--seed", "5", "-r", "--name", "mob", "--num-cpu", "1"])
            k = dataset_generator.main(["--env", "KukaButtonGymEnv-v0", "--num-episode", "1", "--save-path", base, "--seed", "2", "--name", "kuka", "--max-distance", "0.8"])
            out[tag] = (n, k, dict(np.load(base + "mob/preprocessed_data.npz")), dict(np.load(base + "mob/ground_truth.npz")),
                        dict(np.load(base + "kuka/preprocessed_data.npz")), dict(np.load(base + "kuka/ground_truth.npz")))
        finally:
            backend.use_library(None, None)
    (n_c, k_c, pre_c, gt_c, kpre_c, kgt_c), (n_o, k_o, pre_o, gt_o, kpre_o, kgt_o) = out["cuda"], out["oracle"]
    assert n_c == n_o == 3 * 251
    for key in ("rewards", "actions", "episode_starts"):
        assert np.array_equal(pre_c[key], pre_o[key]), key
    for key in ("target_positions", "ground_truth_states"):
        assert np.array_equal(gt_c[key], gt_o[key]), key
    assert list(gt_c["images_path"]) == list(gt_o["images_path"]) and gt_c["images_path"][251] == "mob/record_001/frame000000"
    # Kuka: same seeds and actions; float32 kernel vs float64 oracle
    m = min(k_c, k_o)
    assert m > 10 and abs(k_c - k_o) <= 1
    assert np.array_equal(kpre_c["actions"][:m - 1], kpre_o["actions"][:m - 1])
    assert np.abs(kgt_c["ground_truth_states"][:m - 1] - kgt_o["ground_truth_states"][:m - 1]).max() < 1e-3
    assert kgt_c["target_positions"].shape == (1, 3) and np.abs(kgt_c["target_positions"] - kgt_o["target_positions"]).max() < 1e-5
