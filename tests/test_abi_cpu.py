"""
The C-ABI boundary (include/srl_sim.h) without a GPU: both libraries load and export every declared symbol, the
config struct layouts agree, and the CUDA library refuses (loudly) to run without a device -- there is no CPU path
in the product library.
"""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import CUDA_LIB, ORACLE_LIB, ROOT
from srl_sim import _abi


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "srl_sim.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(srl_sim_\w+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_abi.EXPORTED_SYMBOLS)
    hdr = open(os.path.join(ROOT, "include", "srl_sim.h")).read()
    fields = re.findall(r"^\s+(?:uint32_t|int32_t|float|uint64_t)\s+(\w+);", hdr[hdr.index("typedef struct srl_cfg"):hdr.index("} srl_cfg;")], flags=re.M)
    assert fields == [f[0] for f in _abi.SrlCfg._fields_]
    assert ctypes.sizeof(_abi.SrlCfg) == 64
    for name, kind in _abi.ENV_KINDS.items():
        assert re.search(r"=\s*%d,?\s*/\*\s*%s" % (kind, re.escape(name)), hdr), name


@pytest.mark.parametrize("path", [ORACLE_LIB, CUDA_LIB])
def test_library_loads_and_exports_every_symbol(path, oracle_lib):
    if not os.path.isfile(path):
        pytest.fail("%s has not been built (python __graft_entry__.py build)" % path)
    lib = _abi.SimLibrary(path)            # raises if a symbol is missing or the ABI version differs
    assert lib.lib.srl_sim_abi_version() == _abi.ABI_VERSION
    assert lib.lib.srl_sim_num_envs(None) == 0 and lib.lib.srl_sim_obs_dim(None) == 0


def test_cuda_library_has_no_cpu_path():
    import torch
    lib = _abi.SimLibrary(CUDA_LIB)
    with pytest.raises(_abi.SimError, match="no CPU path"):
        _abi.Sim(lib, "MobileRobotGymEnv-v0", 4, -1)
    if not torch.cuda.is_available():
        with pytest.raises(_abi.SimError):
            _abi.Sim(lib, "MobileRobotGymEnv-v0", 4, 0)       # no device: must fail, never fall back
        from srl_sim.backend import Backend
        with pytest.raises(_abi.SimError, match="no CPU fallback"):
            Backend(lib, 0)


def test_bad_arguments_are_reported_not_thrown(oracle_lib):
    with pytest.raises(_abi.SimError, match="num_envs"):
        _abi.Sim(oracle_lib, "MobileRobotGymEnv-v0", 0, -1)
    with pytest.raises(_abi.SimError, match="model blob"):
        _abi.Sim(oracle_lib, "KukaButtonGymEnv-v0", 2, -1, model_blob=np.zeros(8))
    with pytest.raises(_abi.SimError, match="action_joints"):
        _abi.Sim(oracle_lib, "MobileRobotGymEnv-v0", 2, -1, action_joints=True, is_discrete=False)   # joint actions are Kuka-only ...
    with pytest.raises(_abi.SimError, match="action_joints"):
        from srl_sim.model import load_kuka_scene
        _abi.Sim(oracle_lib, "KukaButtonGymEnv-v0", 2, -1, model_blob=load_kuka_scene().blob, action_joints=True)   # ... and continuous-only
    sim = _abi.Sim(oracle_lib, "MobileRobotGymEnv-v0", 3, -1)
    with pytest.raises(_abi.SimError, match="not available"):
        sim.get_state(_abi.F_JOINT_POS)


def test_vec_env_semantics_on_oracle(use_oracle_backend):
    """stable-baselines VecEnv contract: shapes, auto-reset, Monitor-style info['episode'], createEnvs wrapping."""
    import argparse
    from rl_baselines.utils import createEnvs
    from srl_sim.vec_env import BatchedSRLVecEnv
    env = BatchedSRLVecEnv("MobileRobotGymEnv-v0", 4, seed=0, is_discrete=True)   # config 1: 4 envs, random agent
    obs = env.reset()
    assert obs.shape == (4, 2) and obs.dtype == np.float32 and env.action_space.n == 4
    steps, episodes = 0, []
    for t in range(1600 // 4):                                                     # tests/test_pipeline.py:14 NUM_TIMESTEP
        actions = [env.action_space.sample() for _ in range(4)]
        obs, rew, done, infos = env.step(actions)
        steps += 4
        assert obs.shape == (4, 2) and rew.shape == (4,) and done.dtype == bool and len(infos) == 4
        for i in np.nonzero(done)[0]:
            episodes.append(infos[i]["episode"])
    assert steps == 1600 and len(episodes) == 4 and all(e["l"] == 251 for e in episodes)
    env.close()
    args = argparse.Namespace(env="KukaButtonGymEnv-v0", num_cpu=3, seed=0, num_stack=2, srl_model="ground_truth")
    venv = createEnvs(args, env_kwargs=dict(is_discrete=False, max_steps=10))
    o = venv.reset()
    assert o.shape == (3, 6) and np.abs(o).max() <= 10.0
    for _ in range(12):
        o, r, d, infos = venv.step(np.zeros((3, 3), np.float32))
    assert venv.get_original_obs().shape == (3, 6)
    venv.close()


def test_monitor_files_and_joint_states_on_oracle(use_oracle_backend, tmp_path):
    """bench.Monitor file format (environments/utils.py:53-54, read back like rl_baselines/visualize.py:59-76) from both the
    single-env makeEnv path and the batched VecEnv; Kuka `joints` / `joints_position` states in the batch."""
    import json
    from environments.utils import makeEnv
    from srl_sim.vec_env import BatchedSRLVecEnv

    def load_csv(path):
        with open(path) as f:
            header = json.loads(f.readline()[1:])
            assert f.readline().strip() == "r,l,t"
            rows = [line.strip().split(",") for line in f]
        return header, [(float(r), int(l), float(t)) for r, l, t in rows]

    env = makeEnv("MobileRobotGymEnv-v0", seed=0, rank=3, log_dir=str(tmp_path), env_kwargs=dict(srl_model="ground_truth"))()
    with pytest.raises(RuntimeError):
        env.step(0)                                   # bench.Monitor: step before reset
    env.reset()
    ret, done, n = 0, False, 0
    while not done:
        _, r, done, info = env.step(n % 4)
        ret += r; n += 1
    assert n == 251 and info["episode"]["l"] == 251 and info["episode"]["r"] == ret
    with pytest.raises(RuntimeError):
        env.step(0)                                   # needs reset after done
    env.close()
    header, rows = load_csv(str(tmp_path / "3.monitor.csv"))
    assert header["env_id"] == "MobileRobotGymEnv-v0" and rows == [(float(ret), 251, rows[0][2])]

    venv = BatchedSRLVecEnv("MobileRobotGymEnv-v0", 5, seed=1, log_dir=str(tmp_path / "batch"), global_env_offset=10)
    venv.reset()
    for t in range(260):
        _, _, d, infos = venv.step(np.full(5, t % 4))
    venv.close()
    header, rows = load_csv(str(tmp_path / "batch" / "10.monitor.csv"))
    assert len(rows) == 5 and all(l == 251 for _, l, _ in rows)

    from srl_sim.model import KUKA_INIT_JOINT_POSITIONS
    for model, dim in (("joints", 14), ("joints_position", 17)):
        kenv = BatchedSRLVecEnv("KukaButtonGymEnv-v0", 2, seed=0, srl_model=model, max_steps=5)
        o = kenv.reset()
        assert o.shape == (2, dim) and kenv.observation_space.shape == (dim,)
        assert np.allclose(o[:, -14:], KUKA_INIT_JOINT_POSITIONS, atol=1e-6)
        o, _, _, _ = kenv.step([0, 1])
        assert o.shape == (2, dim)
        kenv.close()
    with pytest.raises(NotImplementedError):
        BatchedSRLVecEnv("MobileRobotGymEnv-v0", 2, srl_model="joints")


def test_dataset_generator_and_episode_saver_on_oracle(use_oracle_backend, tmp_path):
    """environments.dataset_generator + EpisodeSaver (reference dataset_generator.py:37-257, episode_saver.py:90-162): file
    layout, array bookkeeping, and that the fused 3-partition dataset equals the single-partition one (same episode seeds)."""
    import json
    from environments import dataset_generator
    base = str(tmp_path) + "/"
    common = ["--env", "MobileRobotGymEnv-v0", "--num-episode", "5", "--save-path", base, "--seed", "3", "-r"]
    n1 = dataset_generator.main(common + ["--name", "one", "--num-cpu", "1"])
    n3 = dataset_generator.main(common + ["--name", "three", "--num-cpu", "3"])
    assert n1 == n3 == 5 * 251
    for name in ("one", "three"):
        d = base + name
        assert json.load(open(d + "/dataset_config.json")) == {"relative_pos": True, "max_dist": "0.28"}
        assert "MAX_STEPS" in json.load(open(d + "/env_globals.json"))
        assert sorted(os.listdir(d))[-5:] == ["record_%03d" % k for k in range(5)]
        assert not [p for p in os.listdir(base) if "_part-" in p]
    a_pre, b_pre = np.load(base + "one/preprocessed_data.npz"), np.load(base + "three/preprocessed_data.npz")
    a_gt, b_gt = np.load(base + "one/ground_truth.npz"), np.load(base + "three/ground_truth.npz")
    assert a_pre["rewards"].shape == a_pre["actions"].shape == a_pre["episode_starts"].shape == (5 * 251,)
    assert a_pre["episode_starts"].sum() == 5 and a_gt["ground_truth_states"].shape == (5 * 251, 2) and a_gt["target_positions"].shape == (5, 2)
    for k in a_pre.files:
        assert np.array_equal(a_pre[k], b_pre[k]), k
    for k in ("target_positions", "ground_truth_states"):
        assert np.array_equal(a_gt[k], b_gt[k]), k
    assert [p.replace("three/", "one/") for p in b_gt["images_path"]] == list(a_gt["images_path"])
    assert a_gt["images_path"][251] == "one/record_001/frame000000"
    # Kuka with record_data: the default raw_pixels model is accepted for a recording run
    k = dataset_generator.main(["--env", "KukaButtonGymEnv-v0", "--num-episode", "1", "--save-path", base, "--name", "kuka", "--max-distance", "0.8"])
    g = np.load(base + "kuka/ground_truth.npz")
    assert k == len(g["ground_truth_states"]) and g["target_positions"].shape == (1, 3)


def test_dataset_generator_env_testing_mode_without_recording(use_oracle_backend, tmp_path):
    """`dataset_generator --no-record-data` is the reference's env-testing mode (dataset_generator.py:37-121 with record_data False):
    it must run episodes without a saver and write nothing."""
    from environments import dataset_generator
    base = str(tmp_path) + "/"
    n = dataset_generator.main(["--env", "MobileRobotGymEnv-v0", "--num-episode", "2", "--save-path", base, "--name", "dry", "--no-record-data"])
    assert n == 2 * 251
    assert not os.path.exists(base + "dry/preprocessed_data.npz")


def test_reference_shaped_plumbing_matches_batched_env(use_oracle_backend):
    """createEnvs(per_env_objects=True) builds what the reference builds -- num_cpu env objects from makeEnv thunks (seed + rank) behind a
    DummyVecEnv -> VecFrameStack -> VecNormalize (/root/reference/rl_baselines/utils.py:194-229, environments/utils.py:36-57) -- and its
    stepping semantics (auto-reset, post-reset observation) must be those of the env objects themselves."""
    import types
    from environments.utils import makeEnv
    from rl_baselines.utils import DummyVecEnv, createEnvs
    args = types.SimpleNamespace(env="MobileRobotGymEnv-v0", num_cpu=3, seed=5, num_stack=1, srl_model="ground_truth", per_env_objects=True, log_dir=None)
    envs = createEnvs(args, env_kwargs=dict(is_discrete=True))
    assert isinstance(envs.venv.venv, DummyVecEnv) and envs.num_envs == 3
    singles = [makeEnv("MobileRobotGymEnv-v0", 5, i, None, env_kwargs=dict(is_discrete=True, srl_model="ground_truth"))() for i in range(3)]
    o0 = envs.reset()
    raw0 = np.stack([e.reset() for e in singles])
    assert np.array_equal(envs.get_original_obs(), raw0.astype(np.float32)) and o0.shape == (3, 2)   # VecFrameStack holds float32
    rs = np.random.RandomState(0)
    for t in range(260):                                    # crosses the 251-step episode boundary
        a = rs.randint(0, 4, size=3)
        _, r, d, _ = envs.step(list(a))
        for i, e in enumerate(singles):
            o, ri, di, _ = e.step(a[i])
            if di:
                o = e.reset()
            assert np.array_equal(envs.get_original_obs()[i], np.asarray(o, np.float32)) and r[i] == ri and d[i] == di
    assert t == 259
    envs.close()
