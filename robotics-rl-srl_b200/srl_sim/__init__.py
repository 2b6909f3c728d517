"""
srl_sim -- host side of the B200-native batched simulator for the robotics-rl-srl PyBullet envs.

``_abi``     ctypes binding of include/srl_sim.h (libsrl_sim_b200.so, sm_100a CUDA; no CPU fallback)
``backend``  device / buffer plumbing (torch is used only for device memory and streams)
``seeding``  gym 0.11 ``seeding.np_random`` restated (gym is not installed here)
``spaces``   minimal gym.spaces / gym.Env stand-ins
``model``    URDF loader -> flat robot/scene model blob consumed by the Kuka kernels
``vec_env``  BatchedSRLVecEnv: stable-baselines-style VecEnv over one lockstep batch
"""
from ._abi import ENV_KINDS, Sim, SimError, SimLibrary, load_cuda_library  # noqa: F401
