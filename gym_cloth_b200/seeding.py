"""`gym.utils.seeding.np_random` of gym 0.12.1 - the generator behind ClothEnv.seed (cloth_env.py:332-341).

The reference pins gym==0.12.1 (requirements.txt:5); gym is not part of the reference checkout and is not installed in
this image, so the published algorithm of that release is restated here (parity with a real gym install is UNPINNED:
there is no gym here to run it against; the oracle's stub under oracle/stubs/gym/utils/seeding.py restates it separately
and tests/test_abi_cpu.py pins both against hand-computed SHA-512 values):

    seed -> create_seed: int taken modulo 2**64
         -> hash_seed:   first 8 bytes of sha512(str(seed)) as a little-endian integer (padded with one zero word)
         -> np.random.RandomState seeded with that integer split into 32-bit words, least significant first

so `env.seed(1337)` does NOT give RandomState(1337).
"""
import hashlib
import os
import struct

import numpy as np
from numpy.random.bit_generator import ISeedSequence


def _bigint_from_bytes(data):
    sizeof_int = 4
    padding = sizeof_int - len(data) % sizeof_int      # (gym pads a full word when the length is already a multiple)
    data = data + b"\0" * padding
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
        raise ValueError("Invalid type for seed: {} ({})".format(type(a), a))
    return a


def hash_seed(seed=None, max_bytes=8):
    if seed is None:
        seed = create_seed(max_bytes=max_bytes)
    digest = hashlib.sha512(str(seed).encode("utf8")).digest()
    return _bigint_from_bytes(digest[:max_bytes])


def mt_key(seed):
    """The init_by_array key RandomState.seed receives for this env seed."""
    return _int_list_from_bigint(hash_seed(create_seed(seed)))


class _NoEntropy(ISeedSequence):
    """Seed sequence handed to MT19937() so that constructing a generator does not gather and hash OS entropy (3/4 of
    the cost of a RandomState, and 65 536 environments each own one); the state it leaves is overwritten by seed()."""
    _words = np.zeros(624, np.uint32)

    def generate_state(self, n_words, dtype=np.uint32):
        return self._words[:n_words] if (dtype == np.uint32 and n_words <= 624) else np.zeros(n_words, dtype)


_NO_ENTROPY = _NoEntropy()


def np_random(seed=None):
    if seed is not None and not (isinstance(seed, (int, np.integer)) and 0 <= seed):
        raise ValueError("Seed must be a non-negative integer or omitted, not {}".format(seed))
    seed = create_seed(seed)
    rng = np.random.RandomState(np.random.MT19937(_NO_ENTROPY))
    rng.seed(_int_list_from_bigint(hash_seed(seed)))      # legacy init_by_array, as RandomState().seed(list) in gym
    return rng, seed
