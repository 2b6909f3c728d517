"""
Protocol model of the experimental next-episode records (csrc/kuka_kernels.cu: KukaNext, op = PREFETCH), on CPU.

The CUDA code could not be run when it was written, so this test pins the part that does not need a GPU: the hand-over STATE MACHINE.  It
replays the kernel's branches literally -- PREFETCH: skip if valid, read the live episode index (possibly an outdated value: the step kernel may
be running), produce the record for it, publish `episode_for = episode_after_reset_end - 1`, then the flag; ROLLOUT at `done`: if valid, clear the
flag, use the record only when `episode_for == live episode`, otherwise reset in the launch -- under random interleavings of step launches, prefetch
launches that start and finish at arbitrary times, and explicit resets.  Invariant: every episode an env starts begins from
reset_state(env, index the env's own reset would have drawn with), whatever the timing, and the live index advances by exactly one per reset.
"""
import random


def reset_state(env, episode):            # stands for reset_begin -> 5 random micro-steps -> reset_end: a pure function of (seed, env, episode)
    return ("state", env, episode)


class Env(object):
    def __init__(self, i):
        self.i, self.episode, self.state = i, 0, None      # live `e.episode`: the index the NEXT reset draws with
        self.valid, self.episode_for, self.record = 0, -1, None
        self.history = []                                   # (episode index drawn with, start state)

    # ---- kernel branches ----
    def in_launch_reset(self):                              # reset_begin(e.episode) ... reset_end: episode += 1
        self.state = reset_state(self.i, self.episode)
        self.history.append((self.episode, self.state))
        self.episode += 1

    def done_in_rollout(self):
        if self.valid:
            match = self.episode_for == self.episode
            self.valid = 0
            if match:                                       # env_load(record): the record already carries reset_end (episode + 1)
                state, episode_after = self.record
                self.history.append((self.episode, state))
                self.state, self.episode = state, episode_after
                return "hit"
        self.in_launch_reset()
        return "miss"

    def prefetch_begin(self):                               # kernel start: `if (nx.valid[i]) return;` then env_load(live state)
        if self.valid:
            return None
        return self.episode                                 # the value read NOW; the env may move on before the record is published

    def prefetch_end(self, episode_read):
        episode_after_reset_end = episode_read + 1
        self.record = (reset_state(self.i, episode_read), episode_after_reset_end)
        self.episode_for = episode_after_reset_end - 1      # nx.episode[i] = (int)e.episode - 1
        self.valid = 1


def test_handover_never_changes_what_an_episode_starts_from():
    rng = random.Random(0)
    hits = misses = 0
    for trial in range(200):
        envs = [Env(i) for i in range(6)]
        for e in envs:
            e.in_launch_reset()                              # the explicit srl_sim_reset before the first step
        in_flight = []                                       # prefetch threads that have read the live episode and not yet published
        for _ in range(400):
            op = rng.random()
            if op < 0.45:                                    # a step launch: some envs finish their episode
                for e in envs:
                    if rng.random() < 0.25:
                        r = e.done_in_rollout()
                        hits += r == "hit"; misses += r == "miss"
            elif op < 0.70:                                  # a prefetch launch starts (per env: its thread reads the flag and the live index)
                for e in envs:
                    if not any(x[0] is e for x in in_flight):    # launches on one stream are serialised per env
                        ep = e.prefetch_begin()
                        if ep is not None:
                            in_flight.append((e, ep))
            elif op < 0.95 and in_flight:                    # some in-flight prefetch thread publishes its record
                e, ep = in_flight.pop(rng.randrange(len(in_flight)))
                e.prefetch_end(ep)
            else:                                            # explicit srl_sim_reset(mask) between two steps (default instantiation: records untouched)
                rng.choice(envs).in_launch_reset()
        for e in envs:
            assert [h[0] for h in e.history] == list(range(len(e.history)))            # one index per reset, in order, none skipped or repeated
            assert all(s == reset_state(e.i, idx) for idx, s in e.history)             # and the start state is the one that index draws
            assert e.episode == len(e.history)
    assert hits > 1000 and misses > 1000                     # both paths were exercised
