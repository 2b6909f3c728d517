"""
Stub modules that let the REFERENCE env classes (under /root/reference, read-only, only present in
the build container) be imported without pybullet / gym / stable-baselines, so their own Python
arithmetic can be executed to produce golden vectors.  Used only by the gen_*.py scripts in this
directory; nothing here runs on the GPU box or in the product.

What is real and what is stubbed:
  * real: every line of the reference env module being exercised (step/reset/_reward/_termination,
    RNG consumption order, numpy arithmetic);
  * stubbed: `pybullet` (all calls are no-ops returning inert values -- for MobileRobot they have no
    effect on the ground-truth arithmetic), `pybullet_data`, `gym` (Env/spaces shells; `seeding.np_random`
    is the restatement in srl_sim/seeding.py, gym being a third-party dependency that is absent),
    `state_representation.episode_saver`, `srl_zoo.preprocessing`.
"""
import sys
import types

import numpy as np

import os

REFERENCE_ROOT = os.environ.get("SRL_REFERENCE_ROOT", "/root/reference")   # a checkout of araffin/robotics-rl-srl


class _Inert(types.ModuleType):
    """Module whose unknown attributes are no-op callables."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name.isupper():
            return 0

        def _noop(*args, **kwargs):
            return None
        _noop.__name__ = name
        return _noop


def make_pybullet_stub():
    p = _Inert("pybullet")
    counter = {"uid": 0}

    def loadURDF(*a, **k):
        counter["uid"] += 1
        return counter["uid"]

    def getCameraImage(width=1, height=1, **k):
        return (width, height, np.zeros((height, width, 4), dtype=np.uint8), None, None)

    p.loadURDF = loadURDF
    p.loadSDF = lambda *a, **k: [loadURDF()]
    p.connect = lambda *a, **k: 0
    p.getCameraImage = getCameraImage
    p.getQuaternionFromEuler = lambda e: (0.0, 0.0, 0.0, 1.0)
    p.computeViewMatrixFromYawPitchRoll = lambda **k: [0.0] * 16
    p.computeProjectionMatrixFOV = lambda **k: [0.0] * 16
    return p


def install(extra_path, real_pybullet=False):
    """Install the stubs into sys.modules and put the reference + our package on sys.path.
    real_pybullet=True (only meaningful where the PyPI packages exist): keep the REAL `pybullet` / `pybullet_data`, and the
    real `gym` if it imports -- this is the mode that records PyBullet's own arithmetic (gen_kuka_ref_logic_golden.py --real-pybullet)."""
    sys.path.insert(0, extra_path)  # robotics-rl-srl_b200 (for srl_sim.seeding / spaces)
    from srl_sim import seeding as _seeding, spaces as _spaces

    have_gym = False
    if real_pybullet:
        import pybullet  # noqa: F401  (fails loudly when the package is absent)
        import pybullet_data  # noqa: F401
        try:
            import gym  # noqa: F401
            have_gym = True
        except ImportError:
            pass
    else:
        sys.modules["pybullet"] = make_pybullet_stub()
        pd = types.ModuleType("pybullet_data")
        pd.getDataPath = lambda: "/nonexistent/pybullet_data"
        sys.modules["pybullet_data"] = pd

    gym = types.ModuleType("gym")
    gym.Env = _spaces.Env
    gym.spaces = types.ModuleType("gym.spaces")
    gym.spaces.Discrete = _spaces.Discrete
    gym.spaces.Box = _spaces.Box
    gym.utils = types.ModuleType("gym.utils")
    gym.utils.seeding = types.ModuleType("gym.utils.seeding")
    gym.utils.seeding.np_random = _seeding.np_random
    if not have_gym:
        for name, mod in (("gym", gym), ("gym.spaces", gym.spaces), ("gym.utils", gym.utils),
                          ("gym.utils.seeding", gym.utils.seeding)):
            sys.modules[name] = mod

    sr = types.ModuleType("state_representation")
    sr.__path__ = []
    es = types.ModuleType("state_representation.episode_saver")
    es.EpisodeSaver = object
    sys.modules["state_representation"] = sr
    sys.modules["state_representation.episode_saver"] = es
    sz = types.ModuleType("srl_zoo")
    sz.__path__ = []
    pp = types.ModuleType("srl_zoo.preprocessing")
    pp.getNChannels = lambda: 3
    sys.modules["srl_zoo"] = sz
    sys.modules["srl_zoo.preprocessing"] = pp

    # the reference's `environments` package must win over our mirror of the same name
    for k in [k for k in sys.modules if k == "environments" or k.startswith("environments.")]:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
