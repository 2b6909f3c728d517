"""
Next-episode records for the lockstep Kuka step (srl_cfg.prefetch_resets, include/srl_sim.h): with the option on, every step / rollout
launch uses the first idle slot of every warp to advance one incomplete record of that warp's envs by one random micro-step of reset() per env step, and a step whose env
finishes an episode copies a complete record in instead of running reset() inside the launch.  A record is produced by the very
instructions of the in-launch reset, so whatever mix of record hits and in-launch resets a run sees, it must be BIT-IDENTICAL to the
default path.  Reference semantics: the reset() a SubprocVecEnv worker runs between two steps
(/root/reference/rl_baselines/utils.py:216-220, environments/kuka_gym/kuka_button_gym_env.py:214-281).
"""
import numpy as np
import pytest

from srl_sim import _abi
from srl_sim.model import load_kuka_scene

pytestmark = [pytest.mark.gpu]

STATE_FIELDS = (_abi.F_JOINT_POS, _abi.F_JOINT_VEL, _abi.F_EE_CMD, _abi.F_TARGET_POS, _abi.F_COUNTERS, _abi.F_STEP_COUNTER,
                _abi.F_BUTTON_GLIDER, _abi.F_EPISODE_STATS, _abi.F_ROBOT_POS)


def _lockstep(be, kind, n, T, acts, mode, **cfg):
    """mode: off | helper (records only ever produced by the helper slots) | bulk_first (one bulk fill after reset, then helper slots) |
    bulk_side (a bulk fill on a SIDE stream after every step: the library orders it against the step launches)."""
    import torch
    sim = be.make_sim(kind, n, model_blob=load_kuka_scene().blob, prefetch_resets=mode != "off", **cfg)
    obs = be.zeros((n, 3), np.float32); rew = be.zeros((n,), np.float32); done = be.zeros((n,), np.uint8)
    ep_ret = be.zeros((n,), np.float32); ep_len = be.zeros((n,), np.int32)
    main = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    sim.reset(obs_out=obs, stream=main.cuda_stream)
    out = dict(obs=[be.to_host(obs).copy()], rew=[], done=[], ep_ret=[], ep_len=[])
    a = be.from_host(acts)
    if mode == "bulk_first":
        sim.prefetch_resets(stream=main.cuda_stream)
    hits = []
    for t in range(T):
        if mode != "off":
            rec = sim.get_state(_abi.F_NEXT_RECORD)
            live_ep = sim.get_state(_abi.F_COUNTERS)[:, 3]
            ready = (rec[:, 0] == 1) & (rec[:, 2] == live_ep)
        sim.step(a[t], None, obs, rew, done, ep_ret, ep_len, stream=main.cuda_stream)
        if mode == "bulk_side":
            sim.prefetch_resets(stream=side.cuda_stream)
        out["obs"].append(be.to_host(obs).copy()); out["rew"].append(be.to_host(rew).copy()); out["done"].append(be.to_host(done).copy())
        d = out["done"][-1].astype(bool)
        out["ep_ret"].append(np.where(d, be.to_host(ep_ret), 0)); out["ep_len"].append(np.where(d, be.to_host(ep_len), 0))
        if mode != "off":
            hits.append((int((d & ready).sum()), int(d.sum())))     # finished episodes that found a complete record / all finished episodes
    torch.cuda.synchronize()
    state = {f: sim.get_state(f) for f in STATE_FIELDS}
    sim.close()
    return {k: np.stack(v) for k, v in out.items()}, state, hits


@pytest.mark.parametrize("kind,cfg", [("KukaButtonGymEnv-v0", dict(is_discrete=True)), ("KukaRandButtonGymEnv-v0", dict(is_discrete=False, random_target=True)),
                                      ("KukaMovingButtonGymEnv-v0", dict(is_discrete=True))])
def test_lockstep_steps_with_next_episode_records_are_bit_identical(cuda_backend, kind, cfg):
    n, T = 200, 150
    rs = np.random.RandomState(4)
    acts = rs.randint(0, 6, size=(T, n)).astype(np.int32) if cfg["is_discrete"] else rs.uniform(-1, 1, size=(T, n, 3)).astype(np.float32)
    base, base_state, _ = _lockstep(cuda_backend, kind, n, T, acts, "off", seed=9, max_steps=30, **cfg)
    assert base["done"].sum() >= 4 * n                                  # 31-step episodes: every env resets at least four times
    for mode in ("helper", "bulk_first", "bulk_side"):
        got, state, hits = _lockstep(cuda_backend, kind, n, T, acts, mode, seed=9, max_steps=30, **cfg)
        for k in base:
            assert np.array_equal(base[k], got[k]), (mode, k)
        for f in base_state:
            assert np.array_equal(base_state[f], state[f]), (mode, f)
        found, finished = np.sum(hits, axis=0)
        print("%s %s: %d of %d finished episodes found a complete record" % (kind, mode, found, finished))
        # the mechanism must actually serve records (200 envs = one env per warp: its helper slot needs 5 of the 31 launches of an episode)
        assert found >= 0.9 * finished, (mode, found, finished)


@pytest.mark.parametrize("epw", [1, 3, 7])
def test_helper_slot_completes_a_record_in_five_lockstep_launches(cuda_backend, epw):
    """The first idle slot of every warp advances ONE record of its warp's envs by one micro-step per lockstep launch: with `epw` envs per
    warp, after k launches warp w has completed k // 5 records and the next one carries progress k % 5."""
    be = cuda_backend
    n = 40 * epw
    sim = be.make_sim("KukaButtonGymEnv-v0", n, model_blob=load_kuka_scene().blob, seed=1, prefetch_resets=True, envs_per_warp=epw)
    st = be.stream()
    sim.reset(stream=st)
    a = be.from_host(np.zeros((n,), np.int32))
    obs = be.zeros((n, 3), np.float32); rew = be.zeros((n,), np.float32); done = be.zeros((n,), np.uint8)
    rec = sim.get_state(_abi.F_NEXT_RECORD)
    assert rec[:, 0].sum() == 0 and rec[:, 1].sum() == 0
    for k in range(1, 5 * epw + 3):
        sim.step(a, None, obs, rew, done, None, None, stream=st)
        rec = sim.get_state(_abi.F_NEXT_RECORD).reshape(40, epw, 3)
        complete, progress = rec[:, :, 0].sum(axis=1), rec[:, :, 1].sum(axis=1)
        assert np.all(complete == min(k // 5, epw)), (k, complete)
        assert np.all(progress == (k % 5 if k // 5 < epw else 0)), (k, progress)
    live_ep = sim.get_state(_abi.F_COUNTERS)[:, 3]
    rec = rec.reshape(n, 3)
    assert np.all(rec[:, 0] == 1) and np.all(rec[:, 2] == live_ep)     # produced for the episode index the env's next reset() will draw with
    sim.close()


def test_fused_rollout_and_explicit_resets_with_records(cuda_backend):
    """The records also serve the fused T-step rollout (same kernel; the helper slots complete a record within one launch), and an explicit
    srl_sim_reset between two rollouts leaves records for an episode index the env no longer has: they must be dropped, not used."""
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


def test_records_inside_a_captured_cuda_graph(cuda_backend):
    """Config 3's collection loop is a captured CUDA graph of lockstep launches: the helper slots are part of the step kernel, so the records
    need no extra launches, streams or events inside the graph.  Replays must match eager stepping without records bit for bit."""
    import torch
    be = cuda_backend
    n, T, reps = 256, 16, 6
    acts = np.random.RandomState(11).randint(0, 6, size=(T, n)).astype(np.int32)
    blob = load_kuka_scene().blob

    def run(pf, graph):
        sim = be.make_sim("KukaButtonGymEnv-v0", n, model_blob=blob, seed=5, max_steps=20, prefetch_resets=pf)
        a = be.from_host(acts)
        obs = be.zeros((T, n, 3), np.float32); rew = be.zeros((T, n), np.float32); done = be.zeros((T, n), np.uint8)
        s = torch.cuda.Stream()
        out = []
        with torch.cuda.stream(s):
            sim.reset(stream=s.cuda_stream)
            if pf:
                sim.prefetch_resets(stream=s.cuda_stream)

            def steps():
                for t in range(T):
                    sim.step(a[t], None, obs[t], rew[t], done[t], None, None, stream=torch.cuda.current_stream().cuda_stream)
            if graph:
                s.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    steps()
                for _ in range(reps):
                    g.replay(); s.synchronize()
                    out.append((be.to_host(obs).copy(), be.to_host(rew).copy(), be.to_host(done).copy()))
            else:
                for _ in range(reps):
                    steps(); s.synchronize()
                    out.append((be.to_host(obs).copy(), be.to_host(rew).copy(), be.to_host(done).copy()))
        q = sim.get_state(_abi.F_JOINT_POS)
        sim.close()
        return out, q

    base, q0 = run(False, False)
    got, q1 = run(True, True)
    assert sum(int(r[2].sum()) for r in base) >= 3 * n
    for r0, r1 in zip(base, got):
        for x, y in zip(r0, r1):
            assert np.array_equal(x, y)
    assert np.array_equal(q0, q1)
