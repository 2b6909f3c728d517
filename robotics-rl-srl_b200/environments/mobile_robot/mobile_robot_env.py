"""
``MobileRobotGymEnv`` on the B200-native simulator.

Interface mirrored: environments/mobile_robot/mobile_robot_env.py:38-363 of the reference (same
constructor keywords, attributes, ``reset``/``step``/``getGroundTruth``/``getTargetPos``/``getSRLState``,
module constants and ``getGlobals()``).  Underneath, the PyBullet world is gone: the racecar of the
reference is a fixed-base body teleported every step (:207-208,265), so the env is the kinematic
state machine of :235-280,336-363, which runs as an sm_100a kernel behind ``include/srl_sim.h``.
This class is an N=1 view on that kernel; the np_random draws happen here, in the reference's order,
and are handed to the kernel (``reset_draws`` / ``noise``) so seeded runs reproduce the reference.
"""
import numpy as np

from environments.srl_env import SRLGymEnv
from srl_sim import _abi, spaces
from srl_sim.backend import default_backend

#  Number of steps before termination
MAX_STEPS = 250
REWARD_DIST_THRESHOLD = 0.4  # Min distance to target before finishing an episode
RENDER_HEIGHT = 224
RENDER_WIDTH = 224
N_DISCRETE_ACTIONS = 4

DELTA_POS = 0.1
RELATIVE_POS = True  # Use relative position for ground truth
NOISE_STD = 0.0

# From the racecar urdf of the reference, used for the wall bounding box
ROBOT_WIDTH = 0.2
ROBOT_LENGTH = 0.325 * 2


def getGlobals():
    """
    :return: (dict)
    """
    return globals()


class MobileRobotGymEnv(SRLGymEnv):
    """
    Gym wrapper for Mobile Robot environment (additional keyword arguments are discarded, like the
    reference, so Kuka scripts can pass theirs).

    :param urdf_root: (str) unused (no PyBullet assets are loaded)
    :param renders: (bool) GUI rendering is not available
    :param is_discrete: (bool) Whether to use discrete or continuous actions
    :param name: (str) name of the folder where recorded data would be stored
    :param max_distance: (float) unused by this env (kept for signature compatibility)
    :param shape_reward: (bool) Set to true, reward = -distance_to_goal
    :param srl_model: (str) SRL model ("ground_truth" is the supported observation mode)
    :param record_data: (bool) Set to true, record the states / actions / rewards with ``EpisodeSaver`` (frames by name only)
    :param random_target: (bool) Set the target to a random position
    :param state_dim: (int) When learning states
    :param env_rank: (int) the number ID of the environment
    :param srl_pipe: (Queue, [Queue]) contains the input and output of the SRL model
    :param fpv: (bool) first person view camera: a second frame stacked on the channels of the image observation
    :param device: (int) CUDA device ordinal (extension; default 0)
    """
    _ENV_ID = "MobileRobotGymEnv-v0"

    def __init__(self, urdf_root=None, renders=False, is_discrete=True, name="mobile_robot", max_distance=1.6,
                 shape_reward=False, record_data=False, srl_model="raw_pixels", random_target=False, force_down=True,
                 state_dim=-1, learn_states=False, verbose=False, save_path='srl_zoo/data/', env_rank=0, srl_pipe=None,
                 fpv=False, device=None, **_):
        super(MobileRobotGymEnv, self).__init__(srl_model=srl_model, relative_pos=RELATIVE_POS, env_rank=env_rank,
                                                srl_pipe=srl_pipe)
        self._timestep = 1. / 240.
        self._urdf_root = urdf_root
        self._observation = []
        self._env_step_counter = 0
        self._renders = renders
        self._width = RENDER_WIDTH
        self._height = RENDER_HEIGHT
        self._max_distance = max_distance
        self._shape_reward = shape_reward
        self._random_target = random_target
        self._force_down = force_down
        self._is_discrete = is_discrete
        self.terminated = False
        self.n_contacts = 0
        self.state_dim = state_dim
        self.relative_pos = RELATIVE_POS
        self.saver = None
        self.verbose = verbose
        self.max_steps = MAX_STEPS
        self.robot_pos = np.zeros(3)
        self.target_pos = np.zeros(3)
        # Boundaries of the square env
        self._min_x, self._max_x = 0, 4
        self._min_y, self._max_y = 0, 4
        self.has_bumped = False
        self.collision_margin = 0.1
        self.fpv = fpv
        self.camera_target_pos = (2, 2, 0)        # (:84); the 1-D variant looks at (2, 0, 0)
        self.srl_model = srl_model

        if record_data:   # (:109-111)
            from state_representation.episode_saver import EpisodeSaver
            self.saver = EpisodeSaver(name, max_distance, state_dim, globals_=getGlobals(), relative_pos=RELATIVE_POS,
                                      learn_states=learn_states, path=save_path)

        self.action_space = self._make_action_space()
        if self.srl_model == "ground_truth":
            self.state_dim = self.getGroundTruthDim()
        if self.srl_model == "raw_pixels":
            self.observation_space = spaces.Box(low=0, high=255, shape=(self._height, self._width, 6 if fpv else 3), dtype=np.uint8)
        else:
            self.observation_space = spaces.Box(low=-np.inf, high=np.inf, shape=(self.state_dim,), dtype=np.float32)

        # N=1 view on the batched simulator (gym semantics: no auto-reset)
        self._backend = default_backend(device)
        self._sim = self._backend.make_sim(self._ENV_ID, 1, seed=0, is_discrete=is_discrete,
                                           random_target=random_target, shape_reward=shape_reward,
                                           no_auto_reset=True)
        be = self._backend
        self._act_buf = be.zeros((1,) if is_discrete else (1, self._sim.action_dim), np.int32 if is_discrete else np.float32)
        self._noise_buf = be.zeros((1,), np.float32)
        self._obs_buf = be.zeros((1, self._sim.obs_dim), np.float32)
        self._rew_buf = be.zeros((1,), np.float32)
        self._done_buf = be.zeros((1,), np.uint8)

    # ---- hooks the variants override ---------------------------------------------------------
    def _make_action_space(self):
        if self._is_discrete:
            return spaces.Discrete(N_DISCRETE_ACTIONS)
        return spaces.Box(low=-1, high=1, shape=(2,), dtype=np.float32)

    def _reset_draws(self):
        """Consume np_random exactly like mobile_robot_env.py:168-181; returns the 6 reset values."""
        x_start = self._max_x / 2 + self.np_random.uniform(- self._max_x / 3, self._max_x / 3)
        y_start = self._max_y / 2 + self.np_random.uniform(- self._max_y / 3, self._max_y / 3)
        x_pos, y_pos = 0.0, 0.0
        if self._random_target:
            margin = 0.1 * self._max_x
            x_pos = self.np_random.uniform(self._min_x + margin, self._max_x - margin)
            y_pos = self.np_random.uniform(self._min_y + margin, self._max_y - margin)
        return [x_start, y_start, x_pos, y_pos, 0.0, 0.0]

    # ---- reference surface -------------------------------------------------------------------
    def getTargetPos(self):
        # Return only the [x, y] coordinates
        return self.target_pos[:2]

    @staticmethod
    def getGroundTruthDim():
        return 2

    def getGroundTruth(self):
        # Return only the [x, y] coordinates
        return np.array(self.robot_pos)[:2]

    def _pull_state(self):
        self.robot_pos = self._sim.get_state(_abi.F_ROBOT_POS)[0].copy()
        self.target_pos = self._sim.get_state(_abi.F_TARGET_POS)[0].copy()

    def reset(self):
        self.terminated = False
        draws = self._backend.from_host(np.asarray([self._reset_draws()], dtype=np.float64))
        self._sim.reset(mask=None, reset_draws=draws, obs_out=self._obs_buf, stream=self._backend.stream())
        self._env_step_counter = 0
        self.has_bumped = False
        self._pull_state()
        if self.saver is not None:   # (:216-217)
            self.saver.reset(self._frame_for_saver(), self.getTargetPos(), self.getGroundTruth())
        return self._state_or_image()

    def _frame_for_saver(self):
        return self.getObservation() if self.srl_model == "raw_pixels" else None

    def _state_or_image(self):
        if self.srl_model != "raw_pixels":
            return self.getSRLState(self._observation)
        return np.array(self.getObservation())

    def getObservation(self):
        """Image observation of the reference (:228-233): the frame of the fixed top-down camera (+ the first-person frame with fpv)."""
        self._observation = self.render("rgb_array")
        return self._observation

    def _encode_action(self, action):
        if self._is_discrete:
            return np.asarray([int(action)], dtype=np.int32)
        return np.asarray(action, dtype=np.float32).reshape(1, -1)

    def step(self, action):
        # dv = DELTA_POS + np_random.normal(0.0, scale=NOISE_STD): the draw is consumed like in the reference
        noise = self.np_random.normal(0.0, scale=NOISE_STD)
        be = self._backend
        act = be.from_host(self._encode_action(action))
        nz = be.from_host(np.asarray([noise], dtype=np.float32))
        self._sim.step(act, noise=nz, obs_out=self._obs_buf, rew_out=self._rew_buf, done_out=self._done_buf,
                       stream=be.stream())
        rew = float(be.to_host(self._rew_buf)[0])
        done = bool(be.to_host(self._done_buf)[0])
        self._env_step_counter += 1
        self._pull_state()
        self.has_bumped = bool(self._sim.get_state(_abi.F_COUNTERS)[0, 1])
        reward = rew if self._shape_reward else int(rew)
        if self.saver is not None:   # (:274-275)
            self.saver.step(self._frame_for_saver(), action, reward, done, self.getGroundTruth())
        return self._state_or_image(), reward, done, {}

    def render(self, mode='human', close=False):
        if mode != "rgb_array":
            return np.array([])
        from srl_sim.render import MOBILE_CAMERA, mobile_fpv_camera, render_batch
        cams = [dict(MOBILE_CAMERA, target=self.camera_target_pos)]
        if self.fpv:      # first-person camera stacked on the channels (:316-332)
            cams.append(mobile_fpv_camera(self.robot_pos))
        return self._backend.to_host(render_batch(self._sim, self._backend, cams, RENDER_WIDTH, RENDER_HEIGHT))[0].copy()

    def close(self):
        if getattr(self, "_sim", None) is not None:
            self._sim.close()
            self._sim = None

    def _termination(self):
        return self.terminated or self._env_step_counter > self.max_steps
