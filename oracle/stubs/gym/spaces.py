"""gym.spaces stand-in (test infrastructure only): just Box."""
import numpy as np


class Box(object):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        # gym==0.12.1 casts bounds to `dtype` (float32 by default).
        self.low = np.asarray(low).astype(dtype)
        self.high = np.asarray(high).astype(dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)
        self.np_random = np.random.RandomState()

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        return self.np_random.uniform(low=self.low, high=self.high,
                                      size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)
