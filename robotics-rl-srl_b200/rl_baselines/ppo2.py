"""
PPO2 consumer of the batched simulator (BASELINE config 3).

The reference trains with stable-baselines' TF-1.8 ``PPO2`` through ``rl_baselines/rl_algorithm/ppo2.py:58-73`` and
``rl_baselines/base_classes.py:214-253``; neither TensorFlow nor stable-baselines exists in this image, so the algorithm is
restated in PyTorch with the reference's hyper-parameters (n_steps=128, nminibatches=4, noptepochs=4, lr=2.5e-4 * f,
ent_coef=0.01, vf_coef=0.5, cliprange=0.2, gamma=0.99, lam=0.95, max_grad_norm=0.5) and stable-baselines' ``MlpPolicy``
(two separate 64-64 tanh networks for policy and value).  Everything -- envs, observation normalisation
(``VecNormalize(norm_obs=True, norm_reward=False)``, rl_baselines/utils.py:224-227), policy, GAE, optimisation -- stays on the
GPU: the env step is ``srl_sim_step`` on torch tensors, no host round trip per step.
This is a CONSUMER of the hot path (library GEMMs via torch are fine here); the product is the simulator underneath.

Data-parallel over GPUs: under ``torchrun`` every rank owns ``num_envs`` envs (global env offset
``rank * num_envs``, so env streams do not depend on the GPU count) and a replica of the policy.  Two collectives, none of
them on the env-step path: ONE all-reduce of the flattened ~10^4-parameter gradient per minibatch, and ONE all-reduce of the
observation filter's sufficient statistics ([2 D + 1] float64) per rollout.  NCCL on GPUs, gloo in the CPU tests.
"""
import json
import os
import time

import numpy as np
import torch
import torch.nn as nn

from srl_sim.vec_env import BatchedSRLVecEnv

PPO2_DEFAULTS = dict(n_steps=128, ent_coef=0.01, learning_rate=2.5e-4, vf_coef=0.5, max_grad_norm=0.5, gamma=0.99, lam=0.95,
                     nminibatches=4, noptepochs=4, cliprange=0.2)   # rl_algorithm/ppo2.py:58-72


class MlpPolicy(nn.Module):
    """stable-baselines 2.5 ``MlpPolicy``: separate pi / vf towers, 2 x 64 tanh, orthogonal init."""

    def __init__(self, obs_dim, n_actions=None, action_dim=None):
        super().__init__()
        self.discrete = n_actions is not None
        out = n_actions if self.discrete else action_dim

        def tower(last, gain):
            layers = [nn.Linear(obs_dim, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh(), nn.Linear(64, last)]
            for m in layers:
                if isinstance(m, nn.Linear):
                    nn.init.orthogonal_(m.weight, np.sqrt(2)); nn.init.zeros_(m.bias)
            nn.init.orthogonal_(layers[-1].weight, gain)
            return nn.Sequential(*layers)
        self.pi, self.vf = tower(out, 0.01), tower(1, 1.0)
        if not self.discrete:
            self.logstd = nn.Parameter(torch.zeros(out))

    def dist(self, obs):
        logits = self.pi(obs)
        if self.discrete:
            return torch.distributions.Categorical(logits=logits, validate_args=False)   # validation synchronises: not allowed in a captured graph
        return torch.distributions.Normal(logits, self.logstd.exp(), validate_args=False)

    def act(self, obs):
        d = self.dist(obs)
        a = d.sample()
        logp = d.log_prob(a) if self.discrete else d.log_prob(a).sum(-1)
        return a, logp, self.vf(obs).squeeze(-1)

    def evaluate(self, obs, actions):
        d = self.dist(obs)
        logp = d.log_prob(actions) if self.discrete else d.log_prob(actions).sum(-1)
        ent = d.entropy() if self.discrete else d.entropy().sum(-1)
        return logp, ent, self.vf(obs).squeeze(-1)


class RunningNorm(object):
    """VecNormalize's observation filter on the device: running mean / var, clip to +-10.  All of its state (the sample
    count included) lives in device tensors and is updated in place, so the update can sit inside a captured CUDA graph."""

    def __init__(self, dim, device, clip=10.0, eps=1e-8):
        # one contiguous float64 record {mean[dim], var[dim], count} (the layout srl_obs_filter updates in place); the attributes are views
        self.state = torch.cat([torch.zeros(dim, dtype=torch.float64), torch.ones(dim, dtype=torch.float64),
                                torch.full((1,), 1e-4, dtype=torch.float64)]).to(device)
        self.mean, self.var, self.count = self.state[:dim], self.state[dim:2 * dim], self.state[2 * dim]
        self.clip, self.eps = clip, eps

    def update(self, x):
        x = x.double()
        bm, bv, bc = x.mean(0), x.var(0, unbiased=False), float(x.shape[0])
        delta, tot = bm - self.mean, self.count + bc
        new_var = (self.var * self.count + bv * bc + delta ** 2 * self.count * bc / tot) / tot
        self.mean.add_(delta * bc / tot)
        self.var.copy_(new_var)
        self.count.copy_(tot)

    def __call__(self, x, update=True):
        if update:
            self.update(x)
        return torch.clamp((x - self.mean.float()) / torch.sqrt(self.var.float() + self.eps), -self.clip, self.clip)


def _dist_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def merge_running_moments(norm, prior, all_reduce_sum, world):
    """Merge the per-rank observation filters after a rollout.  Every rank started the rollout from the same ``prior`` =
    (mean, var, count) and folded its own batches in; the sufficient statistics S = (count, count * mean, count * (var + mean^2))
    are additive, so the filter that saw every rank's batches is  sum_r S_r - (world - 1) * S_prior.  One all-reduce of
    2 D + 1 doubles; the result is identical on every rank."""
    def stats(mean, var, count):
        return torch.cat([count.reshape(1), count * mean, count * (var + mean * mean)])
    d = norm.mean.numel()
    s = stats(norm.mean, norm.var, norm.count)
    all_reduce_sum(s)
    s = s - (world - 1) * stats(*prior)
    count, mean = s[0], s[1:1 + d] / s[0]
    norm.count.copy_(count)
    norm.mean.copy_(mean)
    norm.var.copy_(torch.clamp(s[1 + d:] / count - mean * mean, min=0.0))


def allreduce_mean_gradients(params, dist, world):
    """Average the gradients over ranks with ONE collective: flatten, all-reduce (sum), scale, scatter back."""
    grads = [p.grad for p in params]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def train(env_id, num_envs, num_timesteps, seed=0, env_kwargs=None, log_dir=None, device=0, hyperparams=None, verbose=1, cuda_graph=True,
          phase_times=None, fused_act=None, prefetch_resets=None, episode_window=40, fused_update=None):
    """PPO2.learn on a BatchedSRLVecEnv.  Returns a history of (timesteps, mean episode return, fps).

    ``cuda_graph``: the n_steps-long collection loop (policy forward, action sampling, observation filter, one simulator
    launch per step, buffer writes -- a few dozen small kernels per env step) is captured ONCE into a CUDA graph and replayed
    per update, so a rollout costs one graph launch instead of ~n_steps x 50 kernel launches from Python.
    ``fused_act`` (default: on whenever the envs live on a GPU; measured 1.3x end to end with the records on, profiles/r02_step_launch_timing.txt):
    run the per-step policy work through the library's own kernels (``srl_policy_act``: both towers, sample, log-prob,
    value and the rollout-buffer writes in one launch; ``srl_obs_filter``: the observation filter in one launch; include/srl_policy.h)
    instead of ~60 small torch kernels: an env step of the collection loop is then three launches.  Sampling then uses the library's
    counter-based streams (keyed by seed and global env index) instead of torch's generator.
    ``prefetch_resets`` (default: on whenever the envs live on a GPU; a no-op for the kinds without records): create the envs with ``srl_cfg.prefetch_resets`` -- every lockstep launch then uses the idle slot of each warp as a helper that
    prepares the next-episode records, so that a step whose env finishes an episode copies a record in instead of running reset() inside
    the launch (include/srl_sim.h: srl_sim_prefetch_resets; validated bit-identical on B200 in round 2).
    ``fused_update`` (default: on whenever the envs live on a GPU): the gradient of a minibatch step comes from the library's
    ``srl_ppo2_grad`` (include/srl_policy.h: forward, PPO2 loss derivative and backward of both towers in one pass, every activation on chip)
    instead of torch autograd over [minibatch, 64] tensors; gradient clipping, the data-parallel all-reduce and Adam stay in torch.
    ``phase_times``: optional dict; when given, every update synchronises between its phases and accumulates the wall time of
    ``collect`` / ``gae`` / ``optimise`` in it (a profiling aid: the synchronisations cost throughput)."""
    hp = dict(PPO2_DEFAULTS); hp.update(hyperparams or {})
    torch.manual_seed(seed)
    env_kwargs = dict(env_kwargs or {})
    dist, rank, world = _dist_world()
    if prefetch_resets is None:
        from srl_sim.backend import default_backend
        prefetch_resets = default_backend(device).on_gpu
    if prefetch_resets:
        env_kwargs["prefetch_resets"] = True
    if env_kwargs.get("srl_model", "ground_truth") != "ground_truth":
        # the collection loop feeds the simulator's observation buffer (the 3-D / 2-D ground-truth observation) straight to the policy
        raise ValueError("ppo2.train supports srl_model='ground_truth' only (got %r)" % env_kwargs["srl_model"])
    import types
    from rl_baselines.utils import createTensorEnvs, save_obs_rms
    env = createTensorEnvs(types.SimpleNamespace(env=env_id, num_cpu=num_envs, seed=seed, device=device), env_kwargs=env_kwargs,
                           global_env_offset=rank * num_envs)
    on_gpu = env.backend.on_gpu
    # device -1 is the CPU oracle installed by a test through srl_sim.backend.use_library (its buffers are numpy arrays,
    # shared with torch below); the product backend is always a CUDA device
    dev = env.backend.torch_device if on_gpu else torch.device("cpu")
    e_obs, e_rew, e_done, e_ep_ret, e_ep_len = [x if on_gpu else torch.from_numpy(x) for x in (env._obs, env._rew, env._done, env._ep_ret, env._ep_len)]
    D = env.observation_space.shape[0]
    if env.is_discrete:
        policy = MlpPolicy(D, n_actions=env.action_space.n).to(dev)
    else:
        policy = MlpPolicy(D, action_dim=env.action_space.shape[0]).to(dev)
    params = list(policy.parameters())
    if dist is not None:
        for p in params:                      # same seed => same init; the broadcast makes it independent of library versions
            dist.broadcast(p.data, 0)
        torch.manual_seed(seed + rank)        # action sampling / minibatch permutations differ per rank
    N, T = num_envs, hp["n_steps"]
    # The GAE recursion and the minibatch step (forward, losses, backward, gradient clip, Adam) are captured into CUDA graphs too:
    # ~1000 and ~150 small launches respectively that cost more on the host than on the GPU.  Data-parallel runs keep the eager
    # minibatch step (its gradient all-reduce sits between backward and the optimiser step).
    graph_update = bool(cuda_graph and on_gpu and dist is None and (T * N) % hp["nminibatches"] == 0)
    if graph_update:      # a captured optimiser needs its step counter and its learning rate on the device
        lr_t = torch.tensor(float(hp["learning_rate"]), device=dev, dtype=torch.float32)
        opt = torch.optim.Adam(params, lr=lr_t, eps=1e-5, capturable=True)
    else:
        opt = torch.optim.Adam(params, lr=hp["learning_rate"], eps=1e-5)
    norm = RunningNorm(D, dev)
    n_updates = max(1, int(num_timesteps) // (N * T * world))
    if log_dir and rank == 0:
        os.makedirs(log_dir, exist_ok=True)
        with open(os.path.join(log_dir, "args.json"), "w") as f:       # train.py:282-283
            json.dump(dict(env=env_id, algo="ppo2", num_cpu=N, num_timesteps=num_timesteps, seed=seed, srl_model="ground_truth", **hp), f)
        with open(os.path.join(log_dir, "env_globals.json"), "w") as f:  # train.py:285-315
            json.dump({k: v for k, v in env_kwargs.items() if isinstance(v, (int, float, str, bool))}, f)
    env.sim.reset(obs_out=env._obs, stream=env.backend.stream())
    if dist is not None:                   # the reset batch goes through the same merge, so every rank starts from one filter
        prior = (norm.mean.clone(), norm.var.clone(), norm.count.clone())
        norm.update(e_obs)
        merge_running_moments(norm, prior, dist.all_reduce, world)
        obs = norm(e_obs.clone(), update=False)
    else:
        obs = norm(e_obs.clone())        # the current (filtered) observation; updated IN PLACE by the collection loop
    buf = dict(obs=torch.empty((T, N, D), device=dev), act=torch.empty((T, N) if env.is_discrete else (T, N, env.sim.action_dim), device=dev,
                                                                       dtype=torch.int64 if env.is_discrete else torch.float32),
               logp=torch.empty((T, N), device=dev), val=torch.empty((T, N), device=dev), rew=torch.empty((T, N), device=dev),
               done=torch.empty((T, N), device=dev), ep_ret=torch.empty((T, N), device=dev), ep_len=torch.zeros((T, N), device=dev, dtype=torch.int32))
    last_val = torch.empty(N, device=dev)
    fused = None
    if fused_act is None:
        fused_act = on_gpu
    if fused_act:
        if not on_gpu:
            raise ValueError("fused_act=True needs the CUDA library (there is no CPU fallback)")
        from srl_sim.policy import FusedPolicy
        fused = FusedPolicy(env.backend.library, policy, norm.state, seed=seed, env_offset=rank * num_envs, clip=norm.clip, eps=norm.eps)
        act_dev = torch.zeros(N if env.is_discrete else (N, env.sim.action_dim), device=dev, dtype=torch.int32 if env.is_discrete else torch.float32)
        done_u8 = torch.zeros((T, N), device=dev, dtype=torch.uint8)

    if prefetch_resets and on_gpu:
        env.sim.prefetch_resets(stream=env.backend.stream())   # bulk fill of the first records; from here on the helper slots of every step launch keep them up

    def collect():
        """n_steps lockstep env steps under the current policy; everything stays on the device, nothing synchronises."""
        with torch.no_grad():
            for t in range(T):
                a, logp, v = policy.act(obs)
                buf["obs"][t], buf["act"][t], buf["logp"][t], buf["val"][t] = obs, a, logp, v
                act_dev = a.to(torch.int32) if env.is_discrete else torch.clamp(a, -1, 1).contiguous()
                env.step_tensors(act_dev)                                 # one kernel launch, tensors stay on the GPU
                buf["rew"][t], buf["done"][t], buf["ep_ret"][t], buf["ep_len"][t] = e_rew, e_done.float(), e_ep_ret, e_ep_len
                obs.copy_(norm(e_obs))
            last_val.copy_(policy.vf(obs).squeeze(-1))

    def collect_fused():
        """The same rollout as ``collect`` in three launches per env step: policy step, simulator step, observation filter.  The
        simulator writes reward / done / episode return straight into the rollout buffers."""
        with torch.no_grad():
            st = env.backend.stream()
            for t in range(T):
                fused.act(N, obs, act_dev, buf["logp"][t], buf["val"][t], obs_buf=buf["obs"][t], act_buf=buf["act"][t], stream=st)
                env.sim.step(act_dev, None, env._obs, buf["rew"][t], done_u8[t], buf["ep_ret"][t], buf["ep_len"][t], stream=st)
                fused.filter(N, env._obs, obs, update=True, stream=st)
            buf["done"].copy_(done_u8)
            last_val.copy_(policy.vf(obs).squeeze(-1))

    if fused is not None:
        collect = collect_fused
    graph = None
    if cuda_graph and on_gpu:
        # an even number of simulator launches per replay keeps the MobileRobot state double buffer (swapped by the host at every
        # launch, so the pointers are baked into the captured kernels) in phase
        if T % 2:
            raise ValueError("cuda_graph=True needs an even n_steps")
        side = torch.cuda.Stream(device=dev)      # library handles / workspaces are created outside the capture
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):
                policy.act(obs); norm(e_obs, update=False)
            if fused is not None:         # first launches outside the capture (one-off function attributes); they change nothing that matters:
                fused.act(N, obs, act_dev, buf["logp"][0], buf["val"][0], stream=env.backend.stream())     # scratch rows, one sampling counter
                fused.filter(N, env._obs, obs, update=False, stream=env.backend.stream())                   # re-normalises the current observation
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):       # capture only records the launches: neither the envs nor the filter advance
            collect()

    # ---- advantage estimation and the minibatch step on static buffers (so that both can be captured) ----
    adv, ret = torch.zeros((T, N), device=dev), torch.zeros((T, N), device=dev)
    flat = {k: v.reshape((T * N,) + v.shape[2:]) for k, v in buf.items()}
    flat_adv, flat_ret = adv.reshape(-1), ret.reshape(-1)
    mb = max(1, T * N // hp["nminibatches"])
    idx_static = torch.zeros(mb, dtype=torch.int64, device=dev)

    def gae():
        """GAE(lambda), the reference's backward recursion over the rollout."""
        if fused_grad is not None:     # one launch instead of ~6 small kernels per step of the recursion
            fused_grad.gae(buf["rew"], buf["val"], buf["done"], last_val, hp["gamma"], hp["lam"], adv, ret, stream=env.backend.stream())
            return
        with torch.no_grad():
            lastgae = torch.zeros(N, device=dev)
            for t in reversed(range(T)):
                nonterminal = 1.0 - buf["done"][t]
                nextval = last_val if t == T - 1 else buf["val"][t + 1]
                delta = buf["rew"][t] + hp["gamma"] * nextval * nonterminal - buf["val"][t]
                lastgae = delta + hp["gamma"] * hp["lam"] * nonterminal * lastgae
                adv[t].copy_(lastgae)
            torch.add(adv, buf["val"], out=ret)

    fused_grad = None
    if fused_update is None:
        fused_update = on_gpu
    if fused_update:
        if not on_gpu:
            raise ValueError("fused_update=True needs the CUDA library (there is no CPU fallback)")
        if (T * N) % mb:
            raise ValueError("fused_update=True needs n_steps * num_envs divisible by nminibatches")
        from srl_sim.policy import FusedPPO2Grad
        fused_grad = FusedPPO2Grad(env.backend.library, policy, mb)

    def zero_grads():
        if fused_grad is None:         # the fused gradient kernel overwrites its static .grad tensors: nothing to clear
            opt.zero_grad(set_to_none=True)

    def minibatch_step(idx):
        if fused_grad is not None:     # the kernel overwrites the static .grad tensors
            fused_grad(idx, flat["obs"], flat["act"], flat_adv, flat_ret, flat["logp"], flat["val"], hp["cliprange"], hp["ent_coef"], hp["vf_coef"],
                       stream=env.backend.stream())
            if dist is not None:
                allreduce_mean_gradients(params, dist, world)
            nn.utils.clip_grad_norm_(params, hp["max_grad_norm"])
            opt.step()
            return
        logp, ent, v = policy.evaluate(flat["obs"][idx], flat["act"][idx])
        a_mb = flat_adv[idx]
        a_mb = (a_mb - a_mb.mean()) / (a_mb.std() + 1e-8)
        ratio = torch.exp(logp - flat["logp"][idx])
        pg = torch.max(-a_mb * ratio, -a_mb * torch.clamp(ratio, 1 - hp["cliprange"], 1 + hp["cliprange"])).mean()
        vclip = flat["val"][idx] + torch.clamp(v - flat["val"][idx], -hp["cliprange"], hp["cliprange"])
        vf_loss = 0.5 * torch.max((v - flat_ret[idx]) ** 2, (vclip - flat_ret[idx]) ** 2).mean()
        loss = pg - hp["ent_coef"] * ent.mean() + hp["vf_coef"] * vf_loss
        loss.backward()
        if dist is not None:           # one all-reduce of the flattened gradient per minibatch
            allreduce_mean_gradients(params, dist, world)
        nn.utils.clip_grad_norm_(params, hp["max_grad_norm"])
        opt.step()

    gae_graph = None
    if graph is not None:              # same conditions as the collection graph; elementwise work only, nothing to warm up
        gae_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gae_graph):
            gae()
    mb_graph, mb_warm = None, 0        # the minibatch step is captured after three eager warm-up steps (real ones) on a side stream

    history, ep_returns = [], []
    # Monitor log (environments/utils.py:53-54 wraps every env in bench.Monitor; rl_baselines/visualize.py:59-107 reads the files back): one
    # `<rank>.monitor.csv` per process with the episodes of all its envs, return and length straight from the kernel's episode statistics
    monitor = None
    if log_dir:
        from srl_sim.monitor import MonitorWriter, compute_mean_reward
        os.makedirs(log_dir, exist_ok=True)
        monitor = MonitorWriter(os.path.join(log_dir, str(rank)), env_id=env_id)
    # best-model callback of the reference (rl_baselines/train.py:132-159): every SAVE_INTERVAL callback calls (20 PPO2 updates of 8 envs x
    # 128 steps there; scaled to this batch) the mean return of the last N_EPISODES_EVAL episodes of the monitor logs is computed, and when it
    # beats the best so far (and MIN_EPISODES_BEFORE_SAVE episodes exist) the observation filter and the model are saved
    N_EPISODES_EVAL, MIN_EPISODES_BEFORE_SAVE = 100, 100
    save_interval = max(1, 20 * 8 * 128 // (N * T * world))
    best_mean_reward, n_saved = -10000.0, 0

    def save_model(path):
        torch.save(dict(policy=policy.state_dict(), obs_mean=norm.mean.clone(), obs_var=norm.var.clone(), obs_count=norm.count.clone()), path)
        save_obs_rms(log_dir, norm.mean.detach().cpu().numpy(), norm.var.detach().cpu().numpy(), float(norm.count))

    def tick(name=None, since=0.0):
        if phase_times is None:
            return 0.0
        if on_gpu:
            torch.cuda.synchronize()
        now = time.perf_counter()
        if name is not None:
            phase_times[name] = phase_times.get(name, 0.0) + now - since
        return now
    t_start = time.time()
    for update in range(1, n_updates + 1):
        frac = 1.0 - (update - 1.0) / n_updates
        if graph_update:
            lr_t.fill_(hp["learning_rate"] * frac)                        # learning_rate = lambda f: f * 2.5e-4
        else:
            for g in opt.param_groups:
                g["lr"] = hp["learning_rate"] * frac
        prior = (norm.mean.clone(), norm.var.clone(), norm.count.clone()) if dist is not None else None
        t_ph = tick()
        if graph is not None:
            graph.replay()
        else:
            collect()
        if dist is not None:                   # one all-reduce of 2 D + 1 doubles per rollout
            merge_running_moments(norm, prior, dist.all_reduce, world)
        t_ph = tick("collect", t_ph)
        dmask = buf["done"].bool()
        new_rets = buf["ep_ret"][dmask].tolist()
        ep_returns.extend(new_rets)
        if monitor is not None and new_rets:
            t_now = time.time() - monitor.t_start
            t_idx = dmask.nonzero()[:, 0].float()                          # step index of each finished episode within the rollout
            t_prev = getattr(monitor, "_t_prev", 0.0)
            monitor.write_episodes(new_rets, buf["ep_len"][dmask].tolist(), (t_prev + (t_now - t_prev) * (t_idx + 1.0) / T).tolist())
            monitor._t_prev = t_now
        if gae_graph is not None:
            gae_graph.replay()
        else:
            gae()
        t_ph = tick("gae", t_ph)
        for _ in range(hp["noptepochs"]):
            perm = torch.randperm(T * N, device=dev)
            for s in range(0, T * N, mb):
                if not graph_update:
                    zero_grads()
                    minibatch_step(perm[s:s + mb])
                    continue
                idx_static.copy_(perm[s:s + mb])
                if mb_graph is not None:
                    mb_graph.replay()
                elif mb_warm < 3:          # warm-up iterations run on a side stream (torch's whole-network capture recipe)
                    cur = torch.cuda.current_stream(dev)
                    side = torch.cuda.Stream(device=dev)
                    side.wait_stream(cur)
                    with torch.cuda.stream(side):
                        zero_grads()
                        minibatch_step(idx_static)
                    cur.wait_stream(side)
                    mb_warm += 1
                else:
                    mb_graph = torch.cuda.CUDAGraph()
                    zero_grads()
                    with torch.cuda.graph(mb_graph):   # gradients are allocated from the graph's pool and re-created by every replay
                        minibatch_step(idx_static)
                    mb_graph.replay()                  # the capture only recorded this minibatch: now run it
        t_ph = tick("optimise", t_ph)
        steps = update * N * T * world
        fps = steps / (time.time() - t_start)
        window = ep_returns[-max(episode_window, N):]                      # --episode_window (train.py:182), at least one episode per env
        if dist is not None:
            from srl_sim.distributed import allgather_episode_stats
            mean_ret, n_ep = allgather_episode_stats(float(np.sum(window)), len(window), device=dev if on_gpu else None)
            mean_ret = mean_ret if n_ep else float("nan")
        else:
            mean_ret = float(np.mean(window)) if window else float("nan")
        history.append((steps, mean_ret, fps))
        if log_dir and update % save_interval == 0:
            if dist is not None:
                dist.barrier()             # every rank's monitor file holds this update's episodes
            if rank == 0:
                if dist is None:           # one process: the in-memory list is the monitor file (same episodes, same order)
                    ok, n_episodes = len(ep_returns) > 0, len(ep_returns)
                    eval_reward = float(np.mean(ep_returns[-N_EPISODES_EVAL:])) if ok else 0.0
                else:
                    ok, eval_reward, n_episodes = compute_mean_reward(log_dir, N_EPISODES_EVAL)
                if ok and verbose:
                    print("Best mean reward: {:.2f} - Last mean reward per episode: {:.2f}".format(best_mean_reward, eval_reward))
                if ok and eval_reward > best_mean_reward and n_episodes >= MIN_EPISODES_BEFORE_SAVE:
                    best_mean_reward = eval_reward
                    if verbose:
                        print("Saving new best model")
                    save_model(os.path.join(log_dir, "ppo2_model.pt"))
                    n_saved += 1
        if verbose and rank == 0:
            print("update %d/%d  steps %d  mean episode return %.3f  episodes %d  fps %.0f" % (update, n_updates, steps, mean_ret, len(ep_returns), fps))
    if monitor is not None:
        monitor.close()
    if log_dir and rank == 0:
        save_model(os.path.join(log_dir, "ppo2_model_final.pt"))
        if n_saved == 0:                   # a run too short for the callback to fire (fewer than MIN_EPISODES_BEFORE_SAVE episodes): keep the last model
            save_model(os.path.join(log_dir, "ppo2_model.pt"))
        with open(os.path.join(log_dir, "best_model.json"), "w") as f:
            json.dump(dict(best_mean_reward=best_mean_reward if n_saved else None, saves=n_saved, n_episodes_eval=N_EPISODES_EVAL,
                           min_episodes_before_save=MIN_EPISODES_BEFORE_SAVE, save_interval_updates=save_interval), f)
    env.close()
    train.best_mean_reward, train.n_saved = best_mean_reward, n_saved
    train.last_policy, train.last_norm = policy, norm      # for callers that want the trained objects (tests, enjoy)
    return history
