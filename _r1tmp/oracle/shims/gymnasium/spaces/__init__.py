import numpy as np


class Box:
    """Box space: bounds, shape, contains(), sample() — nothing else is used."""

    def __init__(self, low, high, shape=None, dtype=np.float64, seed=None):
        self.dtype = np.dtype(dtype)
        self.low = np.asarray(low, dtype=self.dtype)
        self.high = np.asarray(high, dtype=self.dtype)
        self.shape = self.low.shape
        self._rng = np.random.default_rng(seed)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def contains(self, x):
        x = np.asarray(x)
        if not np.can_cast(x.dtype, self.dtype):
            return False
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1e6)
        hi = np.where(np.isfinite(self.high), self.high, 1e6)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def __repr__(self):
        return "Box(%r, %r)" % (self.low, self.high)
