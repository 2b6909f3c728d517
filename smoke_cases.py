"""Small hot-path invocations used by __graft_entry__.smoke(): CUDA (through the C-ABI) vs oracle."""


def _rollout(be, kind, n, T, np, acts=None, noise=None, **cfg):
    sim = be.make_sim(kind, n, **cfg)
    sim.reset(stream=be.stream())
    obs = be.zeros((T, n, sim.obs_dim), np.float32)
    rew = be.zeros((T, n), np.float32)
    done = be.zeros((T, n), np.uint8)
    sim.rollout(T, None if acts is None else be.from_host(acts), None if noise is None else be.from_host(noise),
                obs, rew, done, stream=be.stream())
    out = be.to_host(obs).copy(), be.to_host(rew).copy(), be.to_host(done).copy()
    sim.close()
    return out


def run(cuda, oracle, np):
    # MobileRobot: bit-exact
    acts = np.random.RandomState(0).randint(0, 4, size=(300, 64)).astype(np.int32)
    a = _rollout(cuda, "MobileRobotGymEnv-v0", 64, 300, np, acts, seed=1, random_target=True)
    b = _rollout(oracle, "MobileRobotGymEnv-v0", 64, 300, np, acts, seed=1, random_target=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y), "MobileRobot CUDA != oracle"
    try:
        from srl_sim import model as _model
    except ImportError:
        return
    if not hasattr(_model, "load_kuka_scene"):
        return
    # Kuka: fp32 kernel vs fp64 oracle, tolerance 1e-3 m on the gripper-relative observation
    blob = _model.load_kuka_scene().blob
    n, T = 8, 40
    rs = np.random.RandomState(1)
    acts = rs.randint(0, 6, size=(T, n)).astype(np.int32)
    noise = rs.normal(0, 0.01, size=(T, n)).astype(np.float32)
    a = _rollout(cuda, "KukaButtonGymEnv-v0", n, T, np, acts, noise, seed=2, model_blob=blob)
    b = _rollout(oracle, "KukaButtonGymEnv-v0", n, T, np, acts, noise, seed=2, model_blob=blob)
    assert np.abs(a[0] - b[0]).max() < 1e-3, "Kuka obs diverged: %g" % np.abs(a[0] - b[0]).max()
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), "Kuka reward/done flags differ"
