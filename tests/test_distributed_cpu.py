"""
World-size-2 `gloo` tests of the N>1 path on CPU: contiguous sharding keyed by the global env index reproduces the
single-process batch, and the only collective (the episode-statistics all-gather) agrees across ranks.
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


def _worker(rank, world, port, env_id, total, T, outdir):
    import sys
    sys.path.insert(0, PKG)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from srl_sim import backend
    from srl_sim._abi import SimLibrary
    from srl_sim.distributed import allgather_episode_stats, rank_world, shard
    from srl_sim.vec_env import BatchedSRLVecEnv
    backend.use_library(SimLibrary(ORACLE_LIB), -1)
    r, w, _ = rank_world()
    off, n = shard(total, r, w)
    env = BatchedSRLVecEnv(env_id, n, seed=11, global_env_offset=off, random_target=True, max_steps=40)
    env.reset()
    out = env.rollout_tensors(T)                       # in-stream random actions: the reference's random agent
    d = out["done"].astype(bool)
    mean_ret, episodes = allgather_episode_stats(out["ep_ret"][d].sum(), d.sum())
    np.savez(os.path.join(outdir, "rank%d.npz" % r), obs=out["obs"], rew=out["rew"], done=out["done"], off=off,
             mean_ret=mean_ret, episodes=episodes)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("env_id,total", [("MobileRobotGymEnv-v0", 13), ("KukaRandButtonGymEnv-v0", 5)])
def test_two_rank_sharding_matches_single_process(env_id, total, tmp_path, oracle_lib):
    T, world = 90, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, env_id, total, T, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert [int(p["off"]) for p in parts] == [0, (total + 1) // 2]
    from srl_sim import backend
    from srl_sim.vec_env import BatchedSRLVecEnv
    backend.use_library(oracle_lib, -1)
    try:
        env = BatchedSRLVecEnv(env_id, total, seed=11, random_target=True, max_steps=40)
        env.reset()
        whole = env.rollout_tensors(T)
    finally:
        backend.use_library(None, None)
    for k in ("obs", "rew", "done"):
        assert np.array_equal(whole[k], np.concatenate([p[k] for p in parts], axis=1)), k
    d = whole["done"].astype(bool)
    assert d.sum() >= total                               # max_steps=40 forces resets
    assert int(parts[0]["episodes"]) == int(parts[1]["episodes"]) == int(d.sum())
    assert float(parts[0]["mean_ret"]) == float(parts[1]["mean_ret"]) == pytest.approx(float(whole["ep_ret"][d].sum() / d.sum()))


def test_shard_arithmetic():
    from srl_sim.distributed import shard
    for total in (1, 7, 8, 4096):
        for world in (1, 2, 3, 8):
            offs = [shard(total, r, world) for r in range(world)]
            assert sum(n for _, n in offs) == total
            assert all(offs[r + 1][0] == offs[r][0] + offs[r][1] for r in range(world - 1))
