"""
EXPERIMENTAL, gated: next-episode records for the lockstep Kuka step (srl_cfg.prefetch_resets + srl_sim_prefetch_resets, DESIGN.md section 9
item 4).  The feature was written at the end of round 1 after the GPU budget was spent, so it is OFF by default and these tests only run
with SRL_TEST_PREFETCH=1 (first thing to do on a GPU box next round).  What they demand: lockstep stepping with records -- refreshed on the
same stream (every finished episode hits a record), on a side stream (a mix of hits and in-launch resets, depending on timing), or never
(every reset in the launch) -- is BIT-IDENTICAL to the default path, because a record is produced by the very instructions of the in-launch reset.
"""
import os

import numpy as np
import pytest

from srl_sim import _abi
from srl_sim.model import load_kuka_scene

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("SRL_TEST_PREFETCH") != "1", reason="experimental: set SRL_TEST_PREFETCH=1")]


def _lockstep(be, kind, n, T, acts, mode, **cfg):
    import torch
    sim = be.make_sim(kind, n, model_blob=load_kuka_scene().blob, prefetch_resets=mode != "off", **cfg)
    obs = be.zeros((n, 3), np.float32); rew = be.zeros((n,), np.float32); done = be.zeros((n,), np.uint8)
    ep_ret = be.zeros((n,), np.float32); ep_len = be.zeros((n,), np.int32)
    main = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    sim.reset(obs_out=obs, stream=main.cuda_stream)
    out = dict(obs=[be.to_host(obs).copy()], rew=[], done=[], ep_ret=[], ep_len=[])
    a = be.from_host(acts)

    def refresh():
        if mode == "same_stream":
            sim.prefetch_resets(stream=main.cuda_stream)
        elif mode == "side_stream":
            side.wait_stream(main)
            sim.prefetch_resets(stream=side.cuda_stream)
    refresh()
    for t in range(T):
        sim.step(a[t], None, obs, rew, done, ep_ret, ep_len, stream=main.cuda_stream)
        refresh()
        out["obs"].append(be.to_host(obs).copy()); out["rew"].append(be.to_host(rew).copy()); out["done"].append(be.to_host(done).copy())
        d = out["done"][-1].astype(bool)
        out["ep_ret"].append(np.where(d, be.to_host(ep_ret), 0)); out["ep_len"].append(np.where(d, be.to_host(ep_len), 0))
    torch.cuda.synchronize()
    state = {f: sim.get_state(f) for f in (_abi.F_JOINT_POS, _abi.F_JOINT_VEL, _abi.F_EE_CMD, _abi.F_TARGET_POS, _abi.F_COUNTERS, _abi.F_STEP_COUNTER,
                                           _abi.F_BUTTON_GLIDER, _abi.F_EPISODE_STATS, _abi.F_ROBOT_POS)}
    sim.close()
    return {k: np.stack(v) for k, v in out.items()}, state


@pytest.mark.parametrize("kind,cfg", [("KukaButtonGymEnv-v0", dict(is_discrete=True)), ("KukaRandButtonGymEnv-v0", dict(is_discrete=False, random_target=True)),
                                      ("KukaMovingButtonGymEnv-v0", dict(is_discrete=True))])
def test_lockstep_steps_with_next_episode_records_are_bit_identical(cuda_backend, kind, cfg):
    n, T = 200, 150
    rs = np.random.RandomState(4)
    acts = rs.randint(0, 6, size=(T, n)).astype(np.int32) if cfg["is_discrete"] else rs.uniform(-1, 1, size=(T, n, 3)).astype(np.float32)
    base, base_state = _lockstep(cuda_backend, kind, n, T, acts, "off", seed=9, max_steps=30, **cfg)
    assert base["done"].sum() >= 4 * n                                  # 31-step episodes: every env resets at least four times
    for mode in ("same_stream", "side_stream", "never_refreshed"):
        got, state = _lockstep(cuda_backend, kind, n, T, acts, mode, seed=9, max_steps=30, **cfg)
        for k in base:
            assert np.array_equal(base[k], got[k]), (mode, k)
        for f in base_state:
            assert np.array_equal(base_state[f], state[f]), (mode, f)


def test_fused_rollout_and_explicit_resets_with_records(cuda_backend):
    """The records also serve the fused T-step rollout (same kernel), and an explicit srl_sim_reset between two steps leaves a record for
    an episode index the env no longer has: it must be dropped, not used."""
    import torch
    n, T = 96, 120
    acts = np.random.RandomState(5).randint(0, 6, size=(T, n)).astype(np.int32)
    res = {}
    for pf in (False, True):
        be = cuda_backend
        sim = be.make_sim("KukaButtonGymEnv-v0", n, model_blob=load_kuka_scene().blob, seed=3, max_steps=25, prefetch_resets=pf)
        st = be.stream()
        sim.reset(stream=st)
        sim.prefetch_resets(stream=st)
        obs = be.zeros((T, n, 3), np.float32); rew = be.zeros((T, n), np.float32); done = be.zeros((T, n), np.uint8)
        a = be.from_host(acts)
        sim.rollout(40, a[:40], None, obs[:40], rew[:40], done[:40], None, None, stream=st)
        mask = be.from_host((np.arange(n) % 3 == 0).astype(np.uint8))
        sim.reset(mask=mask, stream=st)                                 # these envs move on one episode: their records are now stale
        sim.rollout(T - 40, a[40:], None, obs[40:], rew[40:], done[40:], None, None, stream=st)
        torch.cuda.synchronize()
        res[pf] = (be.to_host(obs).copy(), be.to_host(rew).copy(), be.to_host(done).copy(), sim.get_state(_abi.F_JOINT_POS), sim.get_state(_abi.F_COUNTERS))
        sim.close()
    for x, y in zip(res[False], res[True]):
        assert np.array_equal(x, y)
