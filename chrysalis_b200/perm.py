"""numpy statement of ``csrc/perm.cuh``: the counter-based permutations the device draws for permutation inference.

Host mirror for the tests (and for reproducing a device draw on the host); the product path evaluates the permutation
inside the kernels."""
from __future__ import annotations

import numpy as np

ROUNDS = 8
_M = np.uint32(0xFFFFFFFF)


def _mix(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return x


def permutation(seed: int, draw: int, n: int) -> np.ndarray:
    """pi_draw as an int64 array: position i takes the value of spot ``pi[i]``."""
    bits = 1
    while bits < 32 and (1 << bits) < n:
        bits += 1
    bits = max(bits, 2)
    a = bits // 2
    b = bits - a
    ma, mb = np.uint64((1 << a) - 1), np.uint64((1 << b) - 1)
    lo, hi = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    m32 = np.uint64(0xFFFFFFFF)
    keys = []
    for i in range(ROUNDS):
        inner = (hi + np.uint64(draw) * np.uint64(0x9E3779B9) + np.uint64(i + 1) * np.uint64(0x85EBCA6B)) & m32
        keys.append(_mix(np.array([lo ^ _mix(np.array([inner]))[0]]))[0])
    x = np.arange(n, dtype=np.uint64)
    todo = np.ones(n, dtype=bool)
    while todo.any():
        v = x[todo]
        left, right = v & ma, v >> np.uint64(a)
        for i in range(ROUNDS):
            if i & 1:
                right ^= _mix((left * np.uint64(0x9E3779B1) + keys[i]) & m32) & mb
            else:
                left ^= _mix((right * np.uint64(0x9E3779B1) + keys[i]) & m32) & ma
        v = (right << np.uint64(a)) | left
        x[todo] = v
        todo[todo] = v >= np.uint64(n)
    return x.astype(np.int64)
