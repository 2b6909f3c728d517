"""
``Monitor`` -- episode statistics in the file format of the reference's ``bench.Monitor`` (stable-baselines 2.5), which
``environments/utils.py:53-54`` wraps around every env and ``rl_baselines/visualize.py:59-107`` (``loadCsv``) reads back:

    #{"t_start": 1537690000.0, "env_id": "KukaButtonGymEnv-v0"}
    r,l,t
    <episode return>,<episode length>,<seconds since t_start>

``Monitor`` wraps one single-env object (``makeEnv``); ``MonitorWriter`` is the same file fed by the batched VecEnv, whose
episode returns / lengths come out of the kernel (``ep_ret_out`` / ``ep_len_out`` of ``srl_sim_step``).
"""
import json
import os
import time


class MonitorWriter(object):
    EXT = "monitor.csv"

    def __init__(self, filename, env_id=None):
        self.t_start = time.time()
        if not filename.endswith(self.EXT):
            filename = os.path.join(filename, self.EXT) if os.path.isdir(filename) else filename + "." + self.EXT
        os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
        self.filename = filename
        self.f = open(filename, "wt")
        self.f.write("#%s\n" % json.dumps({"t_start": self.t_start, "env_id": env_id}))
        self.f.write("r,l,t\n")
        self.f.flush()

    def write_episode(self, ep_ret, ep_len):
        info = {"r": round(float(ep_ret), 6), "l": int(ep_len), "t": round(time.time() - self.t_start, 6)}
        self.f.write("%s,%d,%s\n" % (info["r"], info["l"], info["t"]))
        self.f.flush()
        return info

    def write_episodes(self, ep_rets, ep_lens, times):
        """Many rows at once (the batched trainer finishes ~1000 episodes per update): `times` are seconds since t_start."""
        self.f.write("".join("%s,%d,%s\n" % (round(float(r), 6), int(l), round(float(t), 6)) for r, l, t in zip(ep_rets, ep_lens, times)))
        self.f.flush()

    def close(self):
        if self.f is not None:
            self.f.close()
            self.f = None


def load_monitor_csv(log_folder):
    """What the reference's ``loadCsv`` (rl_baselines/visualize.py:59-107, non-ES branch) makes of a log folder: every ``*.monitor.csv`` is
    read (two header lines, then ``r,l,t`` rows), the episodes of all files are merged in order of their time stamp, and the result is
    ``([[timesteps before the episode, episode return], ...], total timesteps)``."""
    import glob
    rows = []
    for path in glob.glob(os.path.join(log_folder, "*.monitor.csv")):
        with open(path) as f:
            f.readline(); f.readline()
            for line in f:
                r, l, t = line.split(",")
                rows.append((float(t), int(l), float(r)))
    rows.sort(key=lambda x: x[0])
    result, timesteps = [], 0
    for _, l, r in rows:
        result.append([timesteps, r])
        timesteps += l
    return result, timesteps


def compute_mean_reward(log_folder, n_episodes):
    """``computeMeanReward`` of the reference (rl_baselines/utils.py:123-147): mean return of the last ``n_episodes`` episodes of a log
    folder, ``(ok, mean, episodes available)``; not ok while there is no finished episode."""
    result, _ = load_monitor_csv(log_folder)
    if not result:
        return False, 0.0, 0
    y = [r for _, r in result]
    return True, float(sum(y[-n_episodes:]) / len(y[-n_episodes:])), len(y)


class Monitor(object):
    """Single-env wrapper with ``bench.Monitor`` semantics: accumulates rewards, writes one row per finished episode,
    puts ``info['episode']`` on the terminal step, refuses ``step`` after done / early ``reset`` unless allowed."""

    def __init__(self, env, filename, allow_early_resets=False):
        self.env = env
        self.writer = MonitorWriter(filename, env_id=getattr(getattr(env, "spec", None), "id", None)) if filename is not None else None
        self.allow_early_resets = allow_early_resets
        self.rewards = None
        self.needs_reset = True
        self.episode_rewards, self.episode_lengths, self.total_steps = [], [], 0

    def __getattr__(self, name):   # everything else (spaces, getGroundTruth, seed, ...) is the wrapped env's
        return getattr(self.env, name)

    def reset(self):
        if not self.allow_early_resets and not self.needs_reset:
            raise RuntimeError("Tried to reset an environment before done. If you want to allow early resets, "
                               "wrap your env with Monitor(env, path, allow_early_resets=True)")
        self.rewards = []
        self.needs_reset = False
        return self.env.reset()

    def step(self, action):
        if self.needs_reset:
            raise RuntimeError("Tried to step environment that needs reset")
        ob, rew, done, info = self.env.step(action)
        self.rewards.append(rew)
        if done:
            self.needs_reset = True
            eprew, eplen = sum(self.rewards), len(self.rewards)
            ep = self.writer.write_episode(eprew, eplen) if self.writer else {"r": round(float(eprew), 6), "l": eplen, "t": 0.0}
            self.episode_rewards.append(eprew); self.episode_lengths.append(eplen)
            info = dict(info); info["episode"] = ep
        self.total_steps += 1
        return ob, rew, done, info

    def close(self):
        if self.writer:
            self.writer.close()
        self.env.close()

    def get_total_steps(self):
        return self.total_steps

    def get_episode_rewards(self):
        return self.episode_rewards

    def get_episode_lengths(self):
        return self.episode_lengths
