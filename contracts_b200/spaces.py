"""gym spaces when gym is installed, otherwise the minimal equivalents the env API needs.

The reference builds `gym.spaces.Box/Discrete/MultiDiscrete/Dict` objects (cleanup_new.py:90-169,
contract_list.py:20,43,67).  gym is not part of this image, so a tiny stand-in with the same
attributes (`low`, `high` stored as float32 like gym 0.21, `shape`, `n`, `sample()`) is used then.
"""
import numpy as np

try:  # pragma: no cover - depends on the user's environment
    from gym.spaces import Box, Dict, Discrete, MultiDiscrete  # noqa: F401
    HAVE_GYM = True
except Exception:  # noqa: BLE001
    HAVE_GYM = False

    class Space:
        def __init__(self, shape=None, dtype=None):
            self.shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            super().__init__(shape, dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    class Discrete(Space):
        def __init__(self, n):
            self.n = int(n)
            super().__init__((), np.int64)

        def sample(self):
            return int(np.random.randint(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

    class MultiDiscrete(Space):
        def __init__(self, nvec):
            self.nvec = np.asarray(nvec, dtype=np.int64)
            super().__init__(self.nvec.shape, np.int64)

        def sample(self):
            return np.array([np.random.randint(k) for k in self.nvec])

    class Dict(Space):
        def __init__(self, spaces):
            self.spaces = dict(spaces)
            super().__init__(None, None)

        def keys(self):
            return self.spaces.keys()

        def __getitem__(self, k):
            return self.spaces[k]

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}
