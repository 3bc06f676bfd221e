"""The parallel dataset writer must produce archives `np.load` reads exactly like the reference's np.savez_compressed files
(examples/collect.py:41, research/data.py:77)."""
import io
import os
import zipfile
import zlib
import numpy as np
import pytest
from boxlcd_b200 import npz_writer


def test_crc32_combine_matches_zlib():
  rng = np.random.RandomState(0)
  for la, lb in [(0, 5), (5, 0), (1, 1), (1000, 1), (12345, 54321), (1 << 20, 3)]:
    a, b = rng.bytes(la), rng.bytes(lb)
    assert npz_writer.crc32_combine(zlib.crc32(a), zlib.crc32(b), lb) == zlib.crc32(a + b)


def _dataset(n=37, T=11, seed=0):
  rng = np.random.RandomState(seed)
  return {'action': rng.uniform(-1, 1, (n, T, 3)), 'full_state': rng.uniform(-1, 1, (n, T, 16)).astype(np.float32),
          'proprio': rng.uniform(-1, 1, (n, T, 16)).astype(np.float32), 'lcd': rng.uniform(size=(n, T, 16, 32)) < 0.9}


@pytest.mark.parametrize('chunk', [64, 4096, 8 << 20])
@pytest.mark.parametrize('threads', [1, 5])
def test_round_trip_equals_numpy_writer(tmp_path, chunk, threads):
  data = _dataset()
  a, b = tmp_path / 'ours.npz', tmp_path / 'numpy.npz'
  npz_writer.savez_compressed_parallel(a, threads=threads, chunk=chunk, **data)
  np.savez_compressed(b, **data)
  la, lb = np.load(a), np.load(b)
  assert sorted(la.files) == sorted(lb.files) == sorted(data)
  for k in data:
    assert la[k].dtype == lb[k].dtype == data[k].dtype and la[k].shape == data[k].shape and (la[k] == data[k]).all()
  with zipfile.ZipFile(a) as z:
    assert z.testzip() is None                                      # CRCs of every member check out
    assert all(i.compress_type == zipfile.ZIP_DEFLATED for i in z.infolist())
    assert [i.filename for i in z.infolist()] == [k + '.npy' for k in data]
  if chunk >= 4096:
    assert os.path.getsize(a) < 1.1 * os.path.getsize(b) + 4096      # independent chunks cost little compression


def test_suffix_edge_shapes_and_non_contiguous_input(tmp_path):
  data = {'empty': np.zeros((0, 4), np.float32), 'scalar': np.float64(3.5), 'strided': np.arange(40).reshape(5, 8)[:, ::2], 'bools': np.array([True, False])}
  path = npz_writer.savez_compressed_parallel(str(tmp_path / 'x'), threads=2, chunk=16, **data)
  assert path.endswith('x.npz')
  l = np.load(path)
  for k, v in data.items():
    assert l[k].shape == np.asarray(v).shape and (l[k] == v).all()


def test_zip64_records_are_readable(tmp_path):
  """archives past 4 GiB / 65534 members use the ZIP64 end record and central-directory fields; force them on a small file"""
  data = _dataset(9, 5)
  path = npz_writer.savez_compressed_parallel(tmp_path / 'z64.npz', threads=3, chunk=1000, force_zip64=True, **data)
  l = np.load(path)
  for k in data:
    assert (l[k] == data[k]).all()
  with zipfile.ZipFile(path) as z:
    assert z.testzip() is None and len(z.infolist()) == 4
