"""
ctypes binding of ``include/srl_policy.h``: the two per-step helpers of a GPU-resident PPO2 rollout that live in the same
sm_100a library as the simulator -- the policy step (both 64-64 towers, sample, log-probability, value, rollout-buffer writes
in ONE launch) and the VecNormalize observation filter (ONE launch).  With ``srl_sim_step`` a captured rollout is then three
launches per env step instead of ~60 small torch kernels around the simulator's.

Reference pieces replaced: stable-baselines' ``PPO2`` runner ``model.step(obs)`` with ``MlpPolicy`` (selected by
``rl_baselines/rl_algorithm/ppo2.py:58-72``) and ``VecNormalize._obfilt`` (``rl_baselines/utils.py:224-227``).
There is no CPU fallback here either: :class:`FusedPolicy` needs the CUDA library and CUDA tensors.
"""
import ctypes
from ctypes import POINTER, Structure, byref, c_float, c_int, c_int32, c_size_t, c_uint32, c_uint64, c_void_p

HIDDEN, MAX_OBS, MAX_OUT = 64, 8, 8
POLICY_EXPORTS = ["srl_policy_act", "srl_obs_filter", "srl_ppo2_grad", "srl_ppo2_workspace_bytes", "srl_ppo2_gae"]


class SrlMlpPolicy(Structure):
    """struct srl_mlp_policy (include/srl_policy.h)."""
    _fields_ = [("struct_size", c_uint32), ("obs_dim", c_int32), ("n_out", c_int32), ("discrete", c_int32)] + \
               [(name, c_void_p) for name in ("pi_w1", "pi_b1", "pi_w2", "pi_b2", "pi_w3", "pi_b3",
                                              "vf_w1", "vf_b1", "vf_w2", "vf_b2", "vf_w3", "vf_b3", "logstd")]


class SrlMlpGrads(Structure):
    """struct srl_mlp_grads (include/srl_policy.h)."""
    _fields_ = [("struct_size", c_uint32), ("reserved", c_uint32)] + \
               [(name, c_void_p) for name in ("pi_w1", "pi_b1", "pi_w2", "pi_b2", "pi_w3", "pi_b3",
                                              "vf_w1", "vf_b1", "vf_w2", "vf_b2", "vf_w3", "vf_b3", "logstd")]


def bind(cdll):
    """Declare the argument types of the two entry points on a loaded library (raises AttributeError if they are missing)."""
    cdll.srl_policy_act.restype = c_int
    cdll.srl_policy_act.argtypes = [POINTER(SrlMlpPolicy), c_int, c_void_p, c_void_p, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    cdll.srl_obs_filter.restype = c_int
    cdll.srl_obs_filter.argtypes = [c_int, c_int, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p, c_void_p]
    cdll.srl_ppo2_workspace_bytes.restype = c_size_t
    cdll.srl_ppo2_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    cdll.srl_ppo2_grad.restype = c_int
    cdll.srl_ppo2_grad.argtypes = [POINTER(SrlMlpPolicy), POINTER(SrlMlpGrads), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_float, c_float, c_float, c_void_p, c_size_t, c_void_p]
    cdll.srl_ppo2_gae.restype = c_int
    cdll.srl_ppo2_gae.argtypes = [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_void_p]
    return cdll


def policy_struct(policy):
    """``srl_mlp_policy`` over the parameters of an ``rl_baselines.ppo2.MlpPolicy`` (pointers into the live tensors: optimiser
    steps are seen by the next launch).  Returns (struct, keep-alive list)."""
    lin = lambda tower: [m for m in tower if hasattr(m, "weight")]
    pi, vf = lin(policy.pi), lin(policy.vf)
    tensors = []
    for layer in pi + vf:
        for t in (layer.weight, layer.bias):
            if not t.is_contiguous() or str(t.dtype) != "torch.float32":
                raise ValueError("policy parameters must be contiguous float32")
            tensors.append(t)
    obs_dim, n_out = pi[0].weight.shape[1], pi[2].weight.shape[0]
    if pi[0].weight.shape[0] != HIDDEN or pi[1].weight.shape != (HIDDEN, HIDDEN) or vf[2].weight.shape[0] != 1:
        raise ValueError("srl_policy_act implements the 64-64 MlpPolicy")
    if not (1 <= obs_dim <= MAX_OBS and 1 <= n_out <= MAX_OUT):
        raise ValueError("unsupported policy shape obs_dim=%d n_out=%d" % (obs_dim, n_out))
    s = SrlMlpPolicy()
    s.struct_size = ctypes.sizeof(SrlMlpPolicy)
    s.obs_dim, s.n_out, s.discrete = obs_dim, n_out, int(policy.discrete)
    for name, t in zip(("pi_w1", "pi_b1", "pi_w2", "pi_b2", "pi_w3", "pi_b3", "vf_w1", "vf_b1", "vf_w2", "vf_b2", "vf_w3", "vf_b3"), tensors):
        setattr(s, name, t.data_ptr())
    s.logstd = None if policy.discrete else policy.logstd.data_ptr()
    return s, tensors


class FusedPolicy(object):
    """The fused policy step + observation filter on the CUDA library, for one ``MlpPolicy`` and one env batch."""

    def __init__(self, library, policy, filter_state, seed, env_offset=0, clip=10.0, eps=1e-8):
        """
        :param library: (srl_sim._abi.SimLibrary) the loaded CUDA library
        :param policy: (rl_baselines.ppo2.MlpPolicy) on a CUDA device
        :param filter_state: (torch.Tensor) float64 [2 * obs_dim + 1] on the same device: mean, var, count (updated in place)
        :param seed: (int) key of the sampling streams; env ``i`` uses the stream (seed, env_offset + i)
        """
        import torch
        self._lib = bind(library.lib)
        self._library = library
        self.struct, self._keep = policy_struct(policy)
        dev = policy.pi[0].weight.device
        if dev.type != "cuda":
            raise ValueError("FusedPolicy needs a policy on a CUDA device (there is no CPU fallback)")
        if filter_state.dtype != torch.float64 or filter_state.numel() != 2 * self.struct.obs_dim + 1 or not filter_state.is_contiguous():
            raise ValueError("filter_state must be a contiguous float64 tensor of 2 * obs_dim + 1 elements")
        self.filter_state = filter_state
        self.rng = torch.tensor([int(seed) & 0x7FFFFFFFFFFFFFFF, 0, 0], dtype=torch.int64, device=dev)   # {seed, step counter, arrivals}
        self.env_offset, self.clip, self.eps = int(env_offset), float(clip), float(eps)
        self.obs_dim, self.n_out, self.discrete = self.struct.obs_dim, self.struct.n_out, bool(self.struct.discrete)

    def act(self, n, obs, act_env, logp, value, obs_buf=None, act_buf=None, stream=None):
        rc = self._lib.srl_policy_act(byref(self.struct), int(n), obs.data_ptr(), self.rng.data_ptr(), self.env_offset,
                                      None if obs_buf is None else obs_buf.data_ptr(), act_env.data_ptr(),
                                      None if act_buf is None else act_buf.data_ptr(), logp.data_ptr(), value.data_ptr(), stream)
        self._library.check(rc, "srl_policy_act")

    def filter(self, n, obs_raw, obs_norm_out, update=True, stream=None):
        rc = self._lib.srl_obs_filter(int(n), self.obs_dim, obs_raw.data_ptr(), self.filter_state.data_ptr(), int(bool(update)),
                                      self.clip, self.eps, obs_norm_out.data_ptr(), stream)
        self._library.check(rc, "srl_obs_filter")


class FusedPPO2Grad(object):
    """``srl_ppo2_grad``: the gradient of the PPO2 loss over one minibatch in one pass (forward, loss derivative, backward of both towers
    with every activation on chip), written into the policy's ``.grad`` tensors -- what ``loss.backward()`` of
    ``rl_baselines.ppo2``'s minibatch step produces.  Gradient clipping and the optimiser step stay with torch."""

    def __init__(self, library, policy, minibatch):
        import torch
        self._lib = bind(library.lib)
        self._library = library
        self.struct, self._keep = policy_struct(policy)
        dev = policy.pi[0].weight.device
        if dev.type != "cuda":
            raise ValueError("FusedPPO2Grad needs a policy on a CUDA device (there is no CPU fallback)")
        lin = lambda tower: [m for m in tower if hasattr(m, "weight")]
        params = [t for layer in lin(policy.pi) + lin(policy.vf) for t in (layer.weight, layer.bias)]
        if not policy.discrete:
            params.append(policy.logstd)
        self.grads = SrlMlpGrads()
        self.grads.struct_size = ctypes.sizeof(SrlMlpGrads)
        names = ["pi_w1", "pi_b1", "pi_w2", "pi_b2", "pi_w3", "pi_b3", "vf_w1", "vf_b1", "vf_w2", "vf_b2", "vf_w3", "vf_b3"] + ([] if policy.discrete else ["logstd"])
        for name, prm in zip(names, params):
            prm.grad = torch.zeros_like(prm)               # static gradient tensors: the kernel overwrites them, the optimiser reads them (capturable)
            setattr(self.grads, name, prm.grad.data_ptr())
        self.params = params
        self.minibatch = int(minibatch)
        nbytes = int(self._lib.srl_ppo2_workspace_bytes(self.struct.obs_dim, self.struct.n_out, self.struct.discrete, self.minibatch))
        if nbytes <= 0:
            raise ValueError("srl_ppo2_workspace_bytes: unsupported shape")
        self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=dev)

    def gae(self, rew, value, done, last_value, gamma, lam, adv_out, ret_out, stream=None):
        """``srl_ppo2_gae``: GAE(lambda) of a [T, N] rollout in one launch (float32 tensors; ``done`` holds 1.0 where an episode ended)."""
        T, N = rew.shape
        rc = self._lib.srl_ppo2_gae(int(T), int(N), rew.data_ptr(), value.data_ptr(), done.data_ptr(), last_value.data_ptr(), float(gamma), float(lam),
                                    adv_out.data_ptr(), ret_out.data_ptr(), stream)
        self._library.check(rc, "srl_ppo2_gae")

    def __call__(self, idx, obs, actions, adv, ret, old_logp, old_value, cliprange, ent_coef, vf_coef, stream=None):
        """All arguments are CUDA tensors of the whole rollout (``idx``: int64 [minibatch] rows, or None for the first ``minibatch`` rows)."""
        rc = self._lib.srl_ppo2_grad(byref(self.struct), byref(self.grads), self.minibatch, None if idx is None else idx.data_ptr(), obs.data_ptr(),
                                     actions.data_ptr(), adv.data_ptr(), ret.data_ptr(), old_logp.data_ptr(), old_value.data_ptr(),
                                     float(cliprange), float(ent_coef), float(vf_coef), self.workspace.data_ptr(), self.workspace.numel(), stream)
        self._library.check(rc, "srl_ppo2_grad")
