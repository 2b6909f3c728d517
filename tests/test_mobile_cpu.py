"""
CPU tests (no GPU): the oracle and the host-side env classes against golden vectors produced by the
REFERENCE MobileRobot classes themselves (tests/golden/gen_mobile_golden.py), the Philox
known-answer vectors, and VecEnv-style auto-reset semantics.
"""
import ctypes
import os

import numpy as np
import pytest

from conftest import GOLDEN

CASES = {
    # tag: (module, class, kwargs, seed)
    "base_disc": ("mobile_robot_env", "MobileRobotGymEnv", dict(is_discrete=True), 0),
    "base_disc_rand": ("mobile_robot_env", "MobileRobotGymEnv", dict(is_discrete=True, random_target=True), 3),
    "base_disc_shaped": ("mobile_robot_env", "MobileRobotGymEnv", dict(is_discrete=True, shape_reward=True, random_target=True), 11),
    "base_cont": ("mobile_robot_env", "MobileRobotGymEnv", dict(is_discrete=False), 5),
    "base_cont_rand_shaped": ("mobile_robot_env", "MobileRobotGymEnv", dict(is_discrete=False, random_target=True, shape_reward=True), 7),
    "two_target": ("mobile_robot_2target_env", "MobileRobot2TargetGymEnv", dict(is_discrete=True), 1),
    "two_target_rand": ("mobile_robot_2target_env", "MobileRobot2TargetGymEnv", dict(is_discrete=True, random_target=True), 2),
    "one_d": ("mobile_robot_1D_env", "MobileRobot1DGymEnv", dict(is_discrete=True), 4),
    "one_d_rand_shaped": ("mobile_robot_1D_env", "MobileRobot1DGymEnv", dict(is_discrete=True, random_target=True, shape_reward=True), 6),
    "line": ("mobile_robot_line_target_env", "MobileRobotLineTargetGymEnv", dict(is_discrete=True), 8),
    "line_rand_cont": ("mobile_robot_line_target_env", "MobileRobotLineTargetGymEnv", dict(is_discrete=False, random_target=True), 9),
}


def replay_golden_case(tag):
    """Replay one golden case through OUR env class (whatever backend is installed); bit-exact."""
    import importlib
    g = np.load(os.path.join(GOLDEN, "mobile_ref_golden.npz"))
    modname, clsname, kwargs, seed = CASES[tag]
    mod = importlib.import_module("environments.mobile_robot." + modname)
    env = getattr(mod, clsname)(srl_model="ground_truth", **kwargs)
    env.seed(seed)
    actions, obs, reward, done = g[tag + "/action"], g[tag + "/obs"], g[tag + "/reward"], g[tag + "/done"]
    rpos, rtgt = g[tag + "/robot_pos"], g[tag + "/target"]
    reset_obs, reset_pos = g[tag + "/reset_obs"], g[tag + "/reset_pos"]
    t, ep = 0, 0
    while t < len(reward):
        o = env.reset()
        assert np.array_equal(np.asarray(o, dtype=np.float64), reset_obs[ep]), (tag, "reset obs", ep)
        assert np.array_equal(env.robot_pos, reset_pos[ep]), (tag, "reset pos", ep)
        d = False
        while not d:
            a = int(actions[t, 0]) if kwargs.get("is_discrete", True) else actions[t].astype(np.float32)
            o, r, d, info = env.step(a)
            assert np.array_equal(np.asarray(o, dtype=np.float64), obs[t]), (tag, "obs", t, o, obs[t])
            assert np.array_equal(env.robot_pos, rpos[t]), (tag, "pos", t)
            tp = env.getTargetPos()
            assert tp[0] == rtgt[t, 0] and (len(tp) == 1 or tp[1] == rtgt[t, 1]), (tag, "target", t)
            if kwargs.get("shape_reward", False):
                # the kernel returns rewards as float32: exact after rounding the float64 golden to f32
                assert np.float32(r) == np.float32(reward[t]), (tag, "reward", t, r, reward[t])
            else:
                assert isinstance(r, int) and r == reward[t], (tag, "reward", t, r, reward[t])
            assert d == bool(done[t]), (tag, "done", t)
            assert info == {}
            t += 1
        ep += 1
    env.close()
    return t


@pytest.mark.parametrize("tag", sorted(CASES))
def test_reference_golden_through_oracle(tag, use_oracle_backend):
    steps = replay_golden_case(tag)
    assert steps in (251, 502)  # every episode is exactly 251 steps (done <=> counter > 250)


def test_philox_known_answers(oracle_lib):
    """Random123 known-answer vectors for philox4x32-10 (counter, key) -> output."""
    fn = oracle_lib.lib.oracle_philox4x32
    fn.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)]
    fn.restype = None
    out = (ctypes.c_uint32 * 4)()
    # ctr = 0, key = 0
    fn(0, 0, 0, 0, out)
    assert [hex(x) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    # ctr = ff.., key = ff..
    fn(0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, out)
    assert [hex(x) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    # ctr = 243f6a88 85a308d3 13198a2e 03707344, key = a4093822 299f31d0 (digits of pi)
    fn(0x299F31D0A4093822, 0x85A308D3243F6A88, 0x13198A2E, 0x03707344, out)
    assert [hex(x) for x in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_gym_seeding_restatement():
    """first two uniforms of seed 0 under gym 0.11's seeding hash."""
    from srl_sim.seeding import np_random
    rng, seed = np_random(0)
    assert seed == 0
    assert rng.uniform(-4 / 3, 4 / 3) == -1.1883731833312448
    assert rng.uniform(-4 / 3, 4 / 3) == 1.2410424993128542
    with pytest.raises(ValueError):
        np_random(-1)


def _mk(backend, kind, n, **cfg):
    return backend.make_sim(kind, n, **cfg)


@pytest.mark.parametrize("kind", ["MobileRobotGymEnv-v0", "MobileRobot2TargetGymEnv-v0", "MobileRobot1DGymEnv-v0",
                                  "MobileRobotLineTargetGymEnv-v0"])
def test_oracle_autoreset_and_episode_stats(oracle_backend, kind):
    """SubprocVecEnv semantics: fixed 251-step episodes, post-reset obs on done, Monitor-style stats."""
    n, T = 7, 600
    sim = _mk(oracle_backend, kind, n, seed=5, random_target=True)
    obs0 = np.zeros((n, sim.obs_dim), np.float32)
    sim.reset(obs_out=obs0)
    rng = np.random.RandomState(0)
    na = 2 if "1D" in kind else 4
    acts = rng.randint(0, na, size=(T, n)).astype(np.int32)
    obs = np.zeros((T, n, sim.obs_dim), np.float32)
    rew = np.zeros((T, n), np.float32)
    done = np.zeros((T, n), np.uint8)
    ep_ret = np.full((T, n), np.nan, np.float32)
    ep_len = np.zeros((T, n), np.int32)
    sim.rollout(T, acts, None, obs, rew, done, ep_ret, ep_len)
    d = done.astype(bool)
    assert d[250].all() and d[501].all() and d.sum() == 2 * n
    assert (ep_len[d] == 251).all()
    assert np.allclose(ep_ret[250], rew[:251].sum(0)) and np.allclose(ep_ret[501], rew[251:502].sum(0))
    assert set(np.unique(rew)) <= {-1.0, 0.0, 1.0}
    # the obs stored on the done step is the post-reset observation of a fresh episode
    pos = sim.get_state(0)
    assert (sim.get_state(2)[:, 0] == T - 502).all()
    assert np.isfinite(obs).all() and np.isfinite(pos).all()
    # lockstep step() == fused rollout()
    sim2 = _mk(oracle_backend, kind, n, seed=5, random_target=True)
    sim2.reset(obs_out=np.zeros_like(obs0))
    o2 = np.zeros((n, sim.obs_dim), np.float32); r2 = np.zeros(n, np.float32); d2 = np.zeros(n, np.uint8)
    for t in range(300):
        sim2.step(acts[t], None, o2, r2, d2)
        assert np.array_equal(o2, obs[t]) and np.array_equal(r2, rew[t]) and np.array_equal(d2, done[t])


def test_oracle_rejects_unsupported_modes(oracle_backend):
    with pytest.raises(ValueError):
        oracle_backend.make_sim("MobileRobot1DGymEnv-v0", 2, is_discrete=False)
    with pytest.raises(ValueError):
        oracle_backend.make_sim("MobileRobot2TargetGymEnv-v0", 2, is_discrete=False)


def test_oracle_sharding_invariance(oracle_backend):
    """Env i's stream is keyed by its GLOBAL index: two shards == one batch."""
    kind, T = "MobileRobotGymEnv-v0", 300
    whole = _mk(oracle_backend, kind, 10, seed=9, random_target=True)
    a = _mk(oracle_backend, kind, 6, seed=9, random_target=True, global_env_offset=0)
    b = _mk(oracle_backend, kind, 4, seed=9, random_target=True, global_env_offset=6)
    outs = []
    for s in (whole, a, b):
        s.reset()
        obs = np.zeros((T, s.num_envs, 2), np.float32); rew = np.zeros((T, s.num_envs), np.float32)
        s.rollout(T, None, None, obs, rew, None)  # in-stream random actions
        outs.append((obs, rew))
    assert np.array_equal(outs[0][0], np.concatenate([outs[1][0], outs[2][0]], axis=1))
    assert np.array_equal(outs[0][1], np.concatenate([outs[1][1], outs[2][1]], axis=1))
