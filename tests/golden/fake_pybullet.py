"""
A minimal stand-in for the `pybullet` module, backed by the CPU oracle's low-level world (oracle/oracle_kuka.cpp,
`okb_*`).  TEST INFRASTRUCTURE: it exists so that the UNMODIFIED reference env classes
(/root/reference/environments/kuka_gym/*.py) can be imported and run in the build container, making THEIR Python
logic -- action decoding and noise (kuka_button_gym_env.py:293-340), the applyAction call sequence (kuka.py:118-187),
RNG draw order, reset sequencing (:214-281), reward / termination (:422-463) -- the source of the golden vectors in
tests/golden/kuka_ref_logic_golden.npz.  The physics underneath is OUR restatement (real PyBullet is unavailable), so
this pins the env-level logic against the reference code, not the physics engine.

Only the calls the Kuka envs make are implemented; anything else is an inert no-op.
"""
import ctypes
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

POSITION_CONTROL, VELOCITY_CONTROL, TORQUE_CONTROL = 2, 0, 1
DIRECT, GUI, SHARED_MEMORY = 2, 1, 3
ER_TINY_RENDERER, WORLD_FRAME, LINK_FRAME = 1, 1, 2
DEFAULT_KP, DEFAULT_KD, DEFAULT_FORCE = 0.1, 1.0, 100000.0   # setJointMotorControl2 defaults


class _World(object):
    def __init__(self):
        sys.path.insert(0, os.path.join(ROOT, "robotics-rl-srl_b200"))
        from srl_sim.model import load_kuka_scene
        self.scene = load_kuka_scene()
        self.lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle_sim.so"))
        self.lib.okb_create.restype = ctypes.c_void_p
        self.lib.okb_create.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        for name, args in (("okb_reset_world", [ctypes.c_void_p]), ("okb_set_iterations", [ctypes.c_void_p, ctypes.c_int]),
                           ("okb_set_button_base", [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]),
                           ("okb_set_button_base3", [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_double]),
                           ("okb_reset_joint", [ctypes.c_void_p, ctypes.c_int, ctypes.c_double]),
                           ("okb_ik", [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
                           ("okb_step", [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_double] * 4),
                           ("okb_get", [ctypes.c_void_p, ctypes.c_void_p]), ("okb_get2", [ctypes.c_void_p, ctypes.c_void_p]),
                           ("okb_set_ik_damping", [ctypes.c_void_p, ctypes.c_double]),
                           ("okb_set_button2_base", [ctypes.c_void_p, ctypes.c_double, ctypes.c_double])):
            getattr(self.lib, name).argtypes = args
            getattr(self.lib, name).restype = None
        blob = self.scene.blob
        self.h = self.lib.okb_create(blob.ctypes.data, blob.nbytes)
        assert self.h
        self.body_of_joint = {j: b for b, j in enumerate(self.scene.ref_joints)}   # PyBullet joint index -> movable body
        self.reset()

    def reset(self):
        self.lib.okb_reset_world(self.h)
        self.uids = {}
        self.next_uid = 0
        # default joint motors created at load time: velocity target 0 with a small impulse bound; for the Kuka every
        # joint is re-commanded before each step, so only the button's default matters
        self.motors = np.zeros((12, 5))
        self.motors[:, 1:] = [DEFAULT_KP, DEFAULT_KD, 0.0, 0.0]
        self.button = dict(position_control=0, target=0.0, kp=DEFAULT_KP, kd=DEFAULT_KD, force=DEFAULT_FORCE)
        self.n_buttons = 0

    def state2(self):
        out = np.zeros(4)
        self.lib.okb_get2(self.h, out.ctypes.data)
        return out

    def new_uid(self, kind):
        uid = self.next_uid
        self.next_uid += 1
        self.uids[uid] = kind
        return uid

    def state(self):
        out = np.zeros(33)
        self.lib.okb_get(self.h, out.ctypes.data)
        return out


_W = None


def _world():
    global _W
    if _W is None:
        _W = _World()
    return _W


# ---- the pybullet API subset -----------------------------------------------------------------------------------
def connect(*a, **k):
    _world()
    return 0


def disconnect(*a, **k):
    return None


def resetSimulation(*a, **k):
    _world().reset()


def setPhysicsEngineParameter(numSolverIterations=None, **k):
    if numSolverIterations is not None:
        _world().lib.okb_set_iterations(_world().h, int(numSolverIterations))


def setTimeStep(dt):
    assert abs(dt - 1. / 240.) < 1e-12


def setGravity(x, y, z):
    assert (x, y, z) == (0, 0, -10)


def loadURDF(path, *args, **kwargs):
    w = _world()
    name = os.path.basename(str(path))
    if name.startswith("simple_button"):
        pos = args[0] if args else kwargs.get("basePosition")
        w.n_buttons += 1
        if w.n_buttons == 2:                 # Kuka2ButtonGymEnv loads simple_button_2.urdf as a second body (:68)
            w.lib.okb_set_button2_base(w.h, float(pos[0]), float(pos[1]))
            return w.new_uid("button2")
        w.lib.okb_set_button_base(w.h, float(pos[0]), float(pos[1]))
        return w.new_uid("button")
    if name == "table.urdf":
        return w.new_uid("table")
    return w.new_uid("other:" + name)     # plane, distractor objects, sphere: not simulated


def loadSDF(path, *a, **k):
    assert "kuka_with_gripper2" in path
    return [_world().new_uid("kuka")]


def resetBasePositionAndOrientation(uid, pos, orn):
    w = _world()
    if w.uids.get(uid) == "kuka":
        assert np.allclose(pos, [-0.1, 0.0, -0.15])
    elif w.uids.get(uid) == "button":       # KukaMovingButtonGymEnv teleports the button every step (:116-117)
        w.lib.okb_set_button_base3(w.h, float(pos[0]), float(pos[1]), float(pos[2]))


def getNumJoints(uid):
    return 14 if _world().uids.get(uid) == "kuka" else 2


def getJointInfo(uid, i):
    w = _world()
    q_index = -1 if i not in w.body_of_joint else 7 + w.body_of_joint[i]
    return (i, ("joint_%d" % i).encode(), 0 if q_index >= 0 else 4, q_index)


def resetJointState(uid, jointIndex, targetValue, targetVelocity=0):
    w = _world()
    if w.uids.get(uid) == "kuka" and jointIndex in w.body_of_joint:
        w.lib.okb_reset_joint(w.h, w.body_of_joint[jointIndex], float(targetValue))


def setJointMotorControl2(bodyUniqueId=None, jointIndex=None, controlMode=None, targetPosition=0.0, targetVelocity=0.0,
                          force=DEFAULT_FORCE, positionGain=DEFAULT_KP, velocityGain=DEFAULT_KD, maxVelocity=0.0, **k):
    w = _world()
    assert controlMode == POSITION_CONTROL and targetVelocity == 0
    kind = w.uids.get(bodyUniqueId)
    if kind == "kuka":
        if jointIndex in w.body_of_joint:
            w.motors[w.body_of_joint[jointIndex]] = [targetPosition, positionGain, velocityGain, force, maxVelocity]
    elif kind in ("button", "button2"):      # the two-button env arms both with the same command (:137-138)
        assert jointIndex == 1
        w.button = dict(position_control=1, target=targetPosition, kp=positionGain, kd=velocityGain, force=force)


def getQuaternionFromEuler(e):
    r, p, y = [0.5 * v for v in e]
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)


def getEulerFromQuaternion(q):
    return (0.0, 0.0, 0.0)


def calculateInverseKinematics(uid, link, pos, orn=None, lowerLimits=None, upperLimits=None, jointRanges=None, restPoses=None,
                               jointDamping=None, **k):
    """RECALLED pybullet 1.8.6 argument handling (pybullet.c): the null-space task needs all four lists with one entry per
    JOINT of the body (14 here) -- the reference passes 7 (kuka.py:34-40), so it is dropped; `jointDamping` likewise needs 14
    entries (kuka.py:42-43 has them), otherwise the server's default 0.5 per DoF applies."""
    w = _world()
    assert link == 6 and orn is not None and np.allclose(orn, getQuaternionFromEuler([0, -math.pi, 0]))
    n_joints = getNumJoints(uid)
    null_space = all(x is not None and len(x) == n_joints for x in (lowerLimits, upperLimits, jointRanges, restPoses))
    assert not null_space, "null-space IK is not restated (the reference never enables it: 7-entry lists on a 14-joint body)"
    if jointDamping is not None and len(jointDamping) == n_joints:
        assert len(set(jointDamping)) == 1
        damping = float(jointDamping[0])
    else:
        damping = 0.5
    w.lib.okb_set_ik_damping(w.h, damping)
    target = np.asarray(pos, dtype=np.float64).copy()
    out = np.zeros(12)
    w.lib.okb_ik(w.h, target.ctypes.data, out.ctypes.data)
    return tuple(out)                       # one value per movable joint, like PyBullet


def stepSimulation():
    w = _world()
    m = np.ascontiguousarray(w.motors)
    b = w.button
    w.lib.okb_step(w.h, m.ctypes.data, int(b["position_control"]), float(b["target"]), float(b["kp"]), float(b["kd"]), float(b["force"]))


def getLinkState(uid, link, *a, **k):
    w = _world()
    s = w.state()
    kind = w.uids.get(uid)
    if kind == "kuka":
        assert link == 8
        return (tuple(s[25:28]), (0.0, 0.0, 0.0, 1.0))
    if kind == "button":
        assert link == 1
        return (tuple(s[28:31]), (0.0, 0.0, 0.0, 1.0))
    raise AssertionError("getLinkState on %r" % kind)


def getContactPoints(bodyA=None, bodyB=None, linkIndexA=None, *a, **k):
    w = _world()
    s = w.state()
    ka, kb = w.uids.get(bodyA), w.uids.get(bodyB)
    if ka == "button" and kb == "kuka" and linkIndexA == 1:
        return [()] if s[31] else []
    if ka in ("button", "button2") and kb == "kuka" and linkIndexA is None:      # any link of that button body
        return [()] if w.state2()[0 if ka == "button" else 1] else []
    if ka == "table" and kb == "kuka":
        return [()] if s[32] else []
    return []


def applyExternalForce(*a, **k):
    return None


def changeVisualShape(*a, **k):
    return None


def computeViewMatrixFromYawPitchRoll(**k):
    return [0.0] * 16


def computeProjectionMatrixFOV(**k):
    return [0.0] * 16


def getCameraImage(width=1, height=1, **k):
    return (width, height, np.zeros((height, width, 4), dtype=np.uint8), None, None)


def as_module():
    mod = types.ModuleType("pybullet")
    for k, v in globals().items():
        if not k.startswith("_") and k not in ("as_module",):
            setattr(mod, k, v)
    return mod
