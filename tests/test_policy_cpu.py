"""
CPU tests of the PPO2 consumer helpers (include/srl_policy.h): the per-env arithmetic the sm_100a kernels run
(csrc/policy_core.h, compiled for the host as oracle/libpolicy_ref.so) against an INDEPENDENT implementation -- torch's
nn.Linear / tanh / log_softmax / Normal on the same weights, and the trainer's float64 RunningNorm -- plus the export check
of the CUDA library.  The GPU tests (tests/test_policy_gpu.py) then hold the kernels to the same references.
"""
import ctypes
import os
from ctypes import POINTER, c_float, c_int, c_uint64, c_void_p

import numpy as np
import pytest
import torch

from conftest import CUDA_LIB, ROOT
from rl_baselines.ppo2 import MlpPolicy, RunningNorm
from srl_sim.policy import POLICY_EXPORTS, SrlMlpPolicy, policy_struct

REF_LIB = os.path.join(ROOT, "oracle", "libpolicy_ref.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.isfile(REF_LIB):
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "libpolicy_ref.so"])
    lib = ctypes.CDLL(REF_LIB)
    lib.policy_ref_act.restype = c_int
    lib.policy_ref_act.argtypes = [POINTER(SrlMlpPolicy), c_int, c_void_p, c_uint64, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.policy_ref_filter.restype = c_int
    lib.policy_ref_filter.argtypes = [c_int, c_int, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p]
    return lib


def _policy(obs_dim, discrete, n_out, seed):
    torch.manual_seed(seed)
    pol = MlpPolicy(obs_dim, n_actions=n_out) if discrete else MlpPolicy(obs_dim, action_dim=n_out)
    with torch.no_grad():     # the orthogonal init has a 0.01 gain on the last layer: scale it up so that the logits differ visibly
        for tower in (pol.pi, pol.vf):
            tower[-1].weight.mul_(30.0); tower[-1].bias.uniform_(-0.5, 0.5)
        if not discrete:
            pol.logstd.uniform_(-1.0, 0.3)
    return pol


def _ref_act(ref, pol, obs, seed, counter, env_offset=0, with_buf=True):
    st, keep = policy_struct(pol)
    n, A = obs.shape[0], st.n_out
    act_env = np.zeros((n,), np.int32) if pol.discrete else np.zeros((n, A), np.float32)
    act_buf = (np.zeros((n,), np.int64) if pol.discrete else np.zeros((n, A), np.float32)) if with_buf else None
    logp, val, logits = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros((n, A), np.float32)
    o = np.ascontiguousarray(obs.numpy())
    rc = ref.policy_ref_act(ctypes.byref(st), n, o.ctypes.data, seed, counter, env_offset, act_env.ctypes.data,
                            None if act_buf is None else act_buf.ctypes.data, logp.ctypes.data, val.ctypes.data, logits.ctypes.data)
    assert rc == 0
    return act_env, act_buf, logp, val, logits


@pytest.mark.parametrize("obs_dim,n_out", [(3, 6), (2, 4), (1, 2), (8, 8)])
def test_discrete_policy_step_matches_torch(ref, obs_dim, n_out):
    pol = _policy(obs_dim, True, n_out, seed=obs_dim)
    obs = torch.randn(3000, obs_dim) * 2.0
    act_env, act_buf, logp, val, logits = _ref_act(ref, pol, obs, seed=7, counter=5)
    with torch.no_grad():
        t_logits, t_val = pol.pi(obs), pol.vf(obs).squeeze(-1)
        t_logp = torch.log_softmax(t_logits, -1).gather(1, torch.from_numpy(act_buf)[:, None]).squeeze(1)
    assert np.abs(logits - t_logits.numpy()).max() < 2e-5
    assert np.abs(val - t_val.numpy()).max() < 2e-5
    assert np.abs(logp - t_logp.numpy()).max() < 2e-5
    assert np.array_equal(act_env, act_buf.astype(np.int32)) and act_env.min() >= 0 and act_env.max() < n_out
    # the same (seed, env, counter) gives the same sample; another counter or env offset another one
    again = _ref_act(ref, pol, obs, seed=7, counter=5)[0]
    assert np.array_equal(again, act_env)
    assert not np.array_equal(_ref_act(ref, pol, obs, seed=7, counter=6)[0], act_env)
    assert np.array_equal(_ref_act(ref, pol, obs[100:], seed=7, counter=5, env_offset=100)[0], act_env[100:])   # sharding invariance


def test_categorical_sampling_follows_the_softmax(ref):
    pol = _policy(3, True, 6, seed=11)
    one = torch.tensor([[0.3, -1.2, 0.8]])
    n = 120000
    acts = _ref_act(ref, pol, one.repeat(n, 1), seed=3, counter=0)[0]
    with torch.no_grad():
        p = torch.softmax(pol.pi(one), -1).numpy()[0]
    freq = np.bincount(acts, minlength=6) / n
    assert p.min() > 0.01                                           # a non-degenerate distribution, or the test says little
    assert np.abs(freq - p).max() < 5 * np.sqrt(p.max() * (1 - p.max()) / n) + 1e-4, (freq, p)


@pytest.mark.parametrize("obs_dim,n_out", [(3, 3), (3, 7), (2, 2)])
def test_gaussian_policy_step_matches_torch(ref, obs_dim, n_out):
    pol = _policy(obs_dim, False, n_out, seed=20 + n_out)
    obs = torch.randn(40000, obs_dim)
    act_env, act_buf, logp, val, mean = _ref_act(ref, pol, obs, seed=9, counter=2)
    with torch.no_grad():
        t_mean, t_val = pol.pi(obs), pol.vf(obs).squeeze(-1)
        t_logp = torch.distributions.Normal(t_mean, pol.logstd.exp()).log_prob(torch.from_numpy(act_buf)).sum(-1)
    assert np.abs(mean - t_mean.numpy()).max() < 2e-5 and np.abs(val - t_val.numpy()).max() < 2e-5
    assert np.abs(logp - t_logp.numpy()).max() < 2e-4            # the log-probability of the sample actually drawn
    assert np.array_equal(act_env, np.clip(act_buf, -1.0, 1.0))  # the env gets the sample clipped to the Box bounds
    z = (act_buf - mean) / np.exp(pol.logstd.detach().numpy())   # the draws are standard normal, independent across action dims
    assert abs(z.mean()) < 0.015 and abs(z.std() - 1.0) < 0.01          # 4 sigma of 80 000+ draws
    assert np.abs(np.corrcoef(z.T) - np.eye(n_out)).max() < 0.025


def test_observation_filter_matches_the_trainer_running_norm(ref):
    rng = np.random.default_rng(0)
    D = 3
    state = np.concatenate([np.zeros(D), np.ones(D), [1e-4]])
    norm = RunningNorm(D, torch.device("cpu"))
    for it in range(6):
        x = (rng.normal(0.0, 1.0, (777, D)) * [0.3, 2.0, 9.0] + np.array([1.0, -4.0, 0.5]) * (it + 1)).astype(np.float32)
        out = np.zeros_like(x)
        assert ref.policy_ref_filter(x.shape[0], D, x.ctypes.data, state.ctypes.data, 1, 10.0, 1e-8, out.ctypes.data) == 0
        expect = norm(torch.from_numpy(x)).numpy()
        assert np.allclose(state[:D], norm.mean.numpy(), rtol=0, atol=1e-12) and np.allclose(state[D:2 * D], norm.var.numpy(), rtol=1e-12, atol=1e-12)
        assert state[2 * D] == pytest.approx(float(norm.count), rel=1e-15)
        assert np.abs(out - expect).max() < 1e-6
        assert np.abs(out).max() <= 10.0
    frozen = state.copy()
    x = rng.normal(0, 30, (50, D)).astype(np.float32); out = np.zeros_like(x)
    assert ref.policy_ref_filter(50, D, x.ctypes.data, state.ctypes.data, 0, 10.0, 1e-8, out.ctypes.data) == 0     # update = 0: evaluation mode
    assert np.array_equal(state, frozen) and np.abs(out - norm(torch.from_numpy(x), update=False).numpy()).max() < 1e-6
    assert (np.abs(out) == 10.0).any()                            # the clip is reached


def test_cuda_library_exports_the_policy_helpers():
    """Every symbol include/srl_policy.h declares is exported by the sm_100a library (loading needs no GPU)."""
    import re
    hdr = open(os.path.join(ROOT, "include", "srl_policy.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|size_t)\s+(srl_\w+)\s*\(", hdr, flags=re.M)))
    assert declared == sorted(POLICY_EXPORTS)
    if not os.path.isfile(CUDA_LIB):
        pytest.skip("CUDA library not built")
    lib = ctypes.CDLL(CUDA_LIB)
    for name in declared:
        assert hasattr(lib, name), name
    assert ctypes.sizeof(SrlMlpPolicy) == 16 + 13 * 8             # 4 x int32 + 13 pointers, no padding
    from srl_sim.policy import SrlMlpGrads
    assert ctypes.sizeof(SrlMlpGrads) == 8 + 13 * 8
