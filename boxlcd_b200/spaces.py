"""Minimal gym-0.17-shaped spaces (Box, Dict) so the env classes expose `observation_space.spaces[...]`,
`.shape`, `.dtype` and `action_space.sample()` without gym being installed (reference: world_env.py:128-141)."""
import numpy as np


class Box:
  def __init__(self, low, high, shape, dtype=np.float32):
    self.shape = tuple(shape)
    self.dtype = np.dtype(dtype)
    self.low = np.full(self.shape, low, dtype=self.dtype)
    self.high = np.full(self.shape, high, dtype=self.dtype)
    self.np_random = np.random.RandomState()

  def seed(self, seed=None):
    self.np_random = np.random.RandomState(seed)
    return [seed]

  def sample(self):
    if self.dtype == np.bool_:
      return self.np_random.randint(0, 2, self.shape).astype(np.bool_)
    return self.np_random.uniform(self.low, self.high, self.shape).astype(self.dtype)

  def contains(self, x):
    x = np.asarray(x)
    return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

  def __repr__(self):
    return f'Box{self.shape}'


class Dict:
  def __init__(self, spaces):
    # gym 0.17.3 spaces.Dict sorts a plain dict's keys (gym/spaces/dict.py): callers that iterate `.spaces.items()`
    # (examples/collect.py:28, research/data.py:50) see full_state, lcd, proprio -- and write npz members in that order
    self.spaces = dict(sorted(dict(spaces).items()))

  def __getitem__(self, k):
    return self.spaces[k]

  def sample(self):
    return {k: s.sample() for k, s in self.spaces.items()}

  def __repr__(self):
    return f'Dict({self.spaces})'
