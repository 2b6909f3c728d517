"""
Mirror of environments/utils.py:36-95 (``makeEnv`` / ``_make`` / ``dynamicEnvLoad``): the single-env construction
path used by ``rl_baselines``, ``environments.dataset_generator`` and ``replay.enjoy_baselines``.  gym's registry is
replaced by ``environments.registry.registry``; ``bench.Monitor`` is ``srl_sim.monitor.Monitor`` (same ``*.monitor.csv`` format,
readable by the reference's ``rl_baselines/visualize.py:loadCsv``).
"""
import importlib
import os

from environments.registry import registry


def dynamicEnvLoad(env_id):
    """
    :param env_id: (str)
    :return: (module, str, str) module_env, class_name, env_module_path
    """
    entry_point = registry.spec(env_id)._entry_point
    if callable(entry_point):
        class_name, env_module_path = entry_point.__name__, entry_point.__module__
    else:
        env_module_path, class_name = entry_point.split(':')
    try:
        module_env = importlib.import_module(env_module_path)
    except ImportError:
        raise AssertionError("Error: could not import module {}, ".format(env_module_path) +
                             "Halting execution. Are you sure this is a valid environement?")
    return module_env, class_name, env_module_path


def _make(id_, env_kwargs=None):
    """
    :param id_: (str) The environment ID
    :param env_kwargs: (dict) The extra arguments for the environment
    """
    env_kwargs = env_kwargs or {}
    spec = registry.spec(id_)
    module_env, class_name, _ = dynamicEnvLoad(id_)
    env = getattr(module_env, class_name)(**{**spec._kwargs, **env_kwargs})
    env.spec = spec
    return env


def makeEnv(env_id, seed, rank, log_dir, allow_early_resets=False, env_kwargs=None):
    """
    Instantiate one env (an N=1 view on the simulator), seeded with ``seed + rank`` like the reference.
    """
    def _thunk():
        local_env_kwargs = dict(env_kwargs or {})
        local_env_kwargs["env_rank"] = rank
        env = _make(env_id, env_kwargs=local_env_kwargs)
        env.seed(seed + rank)
        if log_dir is not None:
            from srl_sim.monitor import Monitor
            env = Monitor(env, os.path.join(log_dir, str(rank)), allow_early_resets=allow_early_resets)
        return env
    return _thunk
