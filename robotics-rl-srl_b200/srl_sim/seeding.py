"""
Restatement of ``gym.utils.seeding.np_random`` as shipped in gym 0.11.0 (the version the
reference pins, environment.yml:101).  ``SRLGymEnv.seed`` (environments/srl_env.py:71-78) calls it
to build the per-env legacy ``numpy.random.RandomState``; keeping the exact hashing means an env
seeded with ``seed + rank`` (environments/utils.py:52) draws the same stream as in the reference.

gym is not installed in this image, so the published algorithm is restated here:
seed -> sha512(str(seed))[:8] -> little-endian uint32 words (padded with 4 zero bytes when the
length is already a multiple of 4) -> bigint -> list of uint32 -> RandomState.seed(list).
"""
import hashlib
import os
import struct

import numpy as np


def _bigint_from_bytes(data):
    sizeof_int = 4
    padding = sizeof_int - len(data) % sizeof_int
    data += b"\0" * padding
    int_count = len(data) // sizeof_int
    unpacked = struct.unpack("{}I".format(int_count), data)
    accum = 0
    for i, val in enumerate(unpacked):
        accum += 2 ** (sizeof_int * 8 * i) * val
    return accum


def _int_list_from_bigint(bigint):
    if bigint < 0:
        raise ValueError("Seed must be non-negative, not {}".format(bigint))
    if bigint == 0:
        return [0]
    ints = []
    while bigint > 0:
        bigint, mod = divmod(bigint, 2 ** 32)
        ints.append(mod)
    return ints


def create_seed(a=None, max_bytes=8):
    if a is None:
        a = _bigint_from_bytes(os.urandom(max_bytes))
    elif isinstance(a, str):
        a = a.encode("utf8")
        a += hashlib.sha512(a).digest()
        a = _bigint_from_bytes(a[:max_bytes])
    elif isinstance(a, (int, np.integer)):
        a = int(a) % 2 ** (8 * max_bytes)
    else:
        raise TypeError("Invalid type for seed: {} ({})".format(type(a), a))
    return a


def hash_seed(seed=None, max_bytes=8):
    if seed is None:
        seed = create_seed(max_bytes=max_bytes)
    digest = hashlib.sha512(str(seed).encode("utf8")).digest()
    return _bigint_from_bytes(digest[:max_bytes])


def np_random(seed=None):
    """:return: (numpy.random.RandomState, int) exactly as gym 0.11's ``seeding.np_random``."""
    if seed is not None and not (isinstance(seed, (int, np.integer)) and 0 <= seed):
        raise ValueError("Seed must be a non-negative integer or omitted, not {}".format(seed))
    seed = create_seed(seed)
    rng = np.random.RandomState()
    rng.seed(_int_list_from_bigint(hash_seed(seed)))
    return rng, seed
