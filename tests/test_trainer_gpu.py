"""Consumers of the hot path on the GPU: the PPO2 restatement learns on the batched simulator, and the
reference-shaped entry point `python -m rl_baselines.train` runs for every registered env id (tests/test_pipeline.py:95-111
of the reference asserts exactly that: exit code 0 after 1600 steps with --num-cpu 4)."""
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
