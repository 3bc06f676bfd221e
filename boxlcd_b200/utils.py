"""Small host-side helpers with the same public names and semantics as the reference's `boxLCD/utils.py`
(AttrDict :5-7, args_type :9-16, A :18-31, NamedArray :33-101, dict/list filters :105-113, mapto/rmapto :117-119).
Written for this package; NamedArray keeps an index table instead of searching the key list on every access.
"""
import pathlib
import re
import numpy as np


class AttrDict(dict):
  """dict whose items are also attributes (G.ep_len == G['ep_len'])."""
  def __getattr__(self, key):
    try:
      return self[key]
    except KeyError:
      raise KeyError(key)

  def __setattr__(self, key, value):
    self[key] = value


def args_type(default):
  """argparse `type=` callable for a config default (bool from 'True'/'False', int that tolerates 1e5 / 0.5, Path, else type)."""
  if isinstance(default, bool):
    return lambda s: bool(['False', 'True'].index(s))
  if isinstance(default, int):
    return lambda s: float(s) if ('e' in s or '.' in s) else int(s)
  if isinstance(default, pathlib.Path):
    return lambda s: pathlib.Path(s).expanduser()
  return type(default)


class _ArrayMaker:
  """`A[1, 2, 3]` -> np.array([1, 2, 3])"""
  def __getitem__(self, items):
    return np.array(items)


A = _ArrayMaker()


def mapto(a, lowhigh):
  """[-1, 1] -> [low, high]"""
  return ((a + 1.0) / 2.0 * (lowhigh[1] - lowhigh[0])) + lowhigh[0]


def rmapto(a, lowhigh):
  """[low, high] -> [-1, 1]"""
  return ((a - lowhigh[0]) / (lowhigh[1] - lowhigh[0]) * 2) + -1


class NamedArray:
  """View of `arr[..., N]` addressed by the N key names of `arr_info` (name -> (low, high) bounds).
  With do_map, reads map stored [-1,1] values to bounds and writes map bounds back to [-1,1]."""

  def __init__(self, arr, arr_info, do_map=True):
    self.arr = arr
    self.arr_info = arr_info
    self.do_map = do_map
    self._index = {k: i for i, k in enumerate(arr_info)}

  def _name2idx(self, name):
    return self._index[name]

  def _resolve(self, key):
    if isinstance(key, str):
      return self._index[key], self.arr_info[key]
    if isinstance(key, (list, tuple)):
      return [self._index[k] for k in key], np.array([self.arr_info[k] for k in key]).T
    raise NotImplementedError

  def __getitem__(self, key):
    idx, bounds = self._resolve(key)
    val = self.arr[..., idx]
    return mapto(val, bounds) if self.do_map else val

  __call__ = __getitem__

  def __setitem__(self, key, item):
    idx, bounds = self._resolve(key)
    self.arr[..., idx] = rmapto(item, bounds) if self.do_map else item

  def todict(self):
    return {k: self[k] for k in self.arr_info}


def subdict(d, subkeys): return {k: d[k] for k in subkeys}
def sortdict(d): return subdict(d, sorted(d))
def subdlist(d, subkeys): return [d[k] for k in subkeys]
def filtdict(d, phrase): return {k: v for k, v in d.items() if re.match(phrase, k) is not None}
def nfiltdict(d, phrase): return {k: v for k, v in d.items() if re.match(phrase, k) is None}
def filtlist(xs, phrase): return [x for x in xs if re.match(phrase, x) is not None]
def nfiltlist(xs, phrase): return [x for x in xs if re.match(phrase, x) is None]
def get_angle(sin, cos): return np.arctan2(sin, cos)
def make_rot(angle): return np.array([[np.cos(angle), -np.sin(angle)], [np.sin(angle), np.cos(angle)]])
