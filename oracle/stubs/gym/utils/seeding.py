"""gym.utils.seeding stand-in (test infrastructure only): the np_random of gym 0.12.1, the release the reference pins
(requirements.txt:5).  gym itself is not installed here; this restates its published algorithm - the env seed is hashed
with SHA-512 and the first 8 digest bytes, read as little-endian 32-bit words, seed MT19937 through init_by_array - so
the fixtures recorded through this stub are the draws a real gym==0.12.1 install produces for the same env.seed()."""
import hashlib
import os
import struct

import numpy as np


def _words_to_int(raw):
    raw = raw + b"\0" * (4 - len(raw) % 4)            # gym appends a whole zero word to an aligned buffer
    words = struct.unpack("%dI" % (len(raw) // 4), raw)
    return sum(w << (32 * i) for i, w in enumerate(words))


def create_seed(a=None, max_bytes=8):
    if a is None:
        return _words_to_int(os.urandom(max_bytes))
    if isinstance(a, str):
        b = a.encode("utf8")
        return _words_to_int((b + hashlib.sha512(b).digest())[:max_bytes])
    if isinstance(a, int):
        return a % 2 ** (8 * max_bytes)
    raise TypeError("Invalid type for seed: %r" % (a,))


def hash_seed(seed=None, max_bytes=8):
    if seed is None:
        seed = create_seed(max_bytes=max_bytes)
    return _words_to_int(hashlib.sha512(str(seed).encode("utf8")).digest()[:max_bytes])


def np_random(seed=None):
    if seed is not None and not (isinstance(seed, int) and 0 <= seed):
        raise ValueError("Seed must be a non-negative integer or omitted, not %r" % (seed,))
    seed = create_seed(seed)
    h = hash_seed(seed)
    key = []
    while True:
        key.append(h & 0xffffffff)
        h >>= 32
        if h == 0:
            break
    rng = np.random.RandomState()
    rng.seed(key)
    return rng, seed
