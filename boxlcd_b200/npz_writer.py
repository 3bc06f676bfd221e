"""Parallel writer for the reference's dataset files.

`np.savez_compressed` (examples/collect.py:41, research/data.py:77) deflates every array on one core; with the rollouts
coming off the GPU at millions of env-steps per second that single zlib stream is the whole wall time of a collection
run.  `savez_compressed_parallel` writes the SAME container -- a zip archive whose members `<key>.npy` are deflated .npy
files, readable by `np.load` exactly like the reference's -- but deflates each member as independent chunks on a thread
pool (zlib releases the GIL), the way pigz does: every chunk but the last is ended with Z_FULL_FLUSH, so the chunks'
raw deflate streams concatenate into one valid stream; chunk CRCs are merged with crc32_combine.

The zip structures are written by hand because `zipfile` cannot take pre-deflated data; ZIP64 fields are used throughout
for the local headers and where needed in the central directory, so members and archives may exceed 4 GiB.
"""
import io
import os
import struct
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

CHUNK = 1 << 20   # independent deflate units: small enough that one 1000-rollout barrel (~80 MB) keeps every core busy


# -- crc32 of a concatenation from the parts' crcs (zlib's crc32_combine, GF(2) matrix method) --------------------------------
def _gf2_times(mat, vec):
  s, i = 0, 0
  while vec:
    if vec & 1:
      s ^= mat[i]
    vec >>= 1
    i += 1
  return s


def _gf2_square(mat):
  return [_gf2_times(mat, mat[n]) for n in range(32)]


def crc32_combine(crc1, crc2, len2):
  """crc32(A + B) from crc32(A), crc32(B) and len(B)"""
  if len2 <= 0:
    return crc1
  odd = [0xEDB88320] + [1 << n for n in range(31)]   # operator for one zero bit
  even = _gf2_square(odd)                              # two zero bits
  odd = _gf2_square(even)                              # four zero bits
  while True:
    even = _gf2_square(odd)                            # first pass: one zero byte
    if len2 & 1:
      crc1 = _gf2_times(even, crc1)
    len2 >>= 1
    if not len2:
      break
    odd = _gf2_square(even)
    if len2 & 1:
      crc1 = _gf2_times(odd, crc1)
    len2 >>= 1
    if not len2:
      break
  return crc1 ^ crc2


# -- one member --------------------------------------------------------------------------------------------------------------
def _npy_header(arr):
  buf = io.BytesIO()
  np.lib.format.write_array_header_1_0(buf, np.lib.format.header_data_from_array_1_0(arr))
  return buf.getvalue()


def _deflate_chunk(args):
  view, last, level = args
  c = zlib.compressobj(level, zlib.DEFLATED, -15)
  out = c.compress(view) + c.flush(zlib.Z_FINISH if last else zlib.Z_FULL_FLUSH)
  return out, zlib.crc32(view), len(view)


def _dos_time(t):
  lt = time.localtime(t)
  return (lt.tm_hour << 11) | (lt.tm_min << 5) | (lt.tm_sec // 2), (max(lt.tm_year - 1980, 0) << 9) | (lt.tm_mon << 5) | lt.tm_mday


def _local_header(name, dtime, ddate, crc, csize, usize):
  extra = struct.pack('<HHQQ', 0x0001, 16, usize, csize)
  return struct.pack('<IHHHHHIIIHH', 0x04034B50, 45, 0, 8, dtime, ddate, crc, 0xFFFFFFFF, 0xFFFFFFFF, len(name), len(extra)) + name + extra


def _central_header(name, dtime, ddate, crc, csize, usize, offset, zip64=False):
  fields, c32, u32, o32 = b'', csize, usize, offset
  if zip64 or usize >= 0xFFFFFFFF or csize >= 0xFFFFFFFF or offset >= 0xFFFFFFFF:   # all three, in the order the spec fixes
    fields = struct.pack('<QQQ', usize, csize, offset)
    c32 = u32 = o32 = 0xFFFFFFFF
  extra = struct.pack('<HH', 0x0001, len(fields)) + fields if fields else b''
  return struct.pack('<IHHHHHHIIIHHHHHII', 0x02014B50, 45, 45, 0, 8, dtime, ddate, crc, c32, u32, len(name), len(extra), 0, 0, 0, 0o600 << 16,
                     o32) + name + extra


def savez_compressed_parallel(file, threads=None, level=6, chunk=CHUNK, force_zip64=False, **arrays):
  """Drop-in for np.savez_compressed(file, **arrays) that deflates on `threads` cores (default: all).
  force_zip64 writes the ZIP64 records even where the 32-bit fields would do (they are otherwise used only past 4 GiB / 65534 members)."""
  path = os.fspath(file)
  if not path.endswith('.npz'):
    path += '.npz'          # np.savez appends the suffix too
  threads = threads or os.cpu_count() or 1
  dtime, ddate = _dos_time(time.time())
  central = []
  with open(path, 'wb') as f, ThreadPoolExecutor(max_workers=threads) as pool:
    for key, val in arrays.items():
      arr = np.asanyarray(val)
      if arr.ndim and not arr.flags.c_contiguous:
        arr = np.ascontiguousarray(arr)   # (ascontiguousarray would turn a 0-d array into shape (1,))
      if arr.dtype.hasobject:
        raise TypeError(f'{key}: object arrays are not supported')
      name = (key + '.npy').encode()
      head = _npy_header(arr)
      body = memoryview(arr.reshape(-1).view(np.uint8)) if arr.size else memoryview(b'')
      # chunk list over header + body without copying the body
      pieces = [memoryview(head)] + [body[o:o + chunk] for o in range(0, len(body), chunk)]
      jobs = [(p, i == len(pieces) - 1, level) for i, p in enumerate(pieces)]
      offset = f.tell()
      f.write(_local_header(name, dtime, ddate, 0, 0, 0))        # patched below once the sizes are known
      crc, csize, usize = 0, 0, 0
      for out, ccrc, clen in pool.map(_deflate_chunk, jobs):
        f.write(out)
        crc = crc32_combine(crc, ccrc, clen) if usize else ccrc
        csize += len(out)
        usize += clen
      end = f.tell()
      f.seek(offset)
      f.write(_local_header(name, dtime, ddate, crc, csize, usize))
      f.seek(end)
      central.append(_central_header(name, dtime, ddate, crc, csize, usize, offset, force_zip64))
    cd_offset = f.tell()
    for rec in central:
      f.write(rec)
    cd_size = f.tell() - cd_offset
    n = len(central)
    if force_zip64 or n >= 0xFFFF or cd_offset >= 0xFFFFFFFF or cd_size >= 0xFFFFFFFF:
      z64 = f.tell()
      f.write(struct.pack('<IQHHIIQQQQ', 0x06064B50, 44, 45, 45, 0, 0, n, n, cd_size, cd_offset))
      f.write(struct.pack('<IIQI', 0x07064B50, 0, z64, 1))
    # saturated 16/32-bit fields send the reader to the ZIP64 record
    n16 = 0xFFFF if force_zip64 else min(n, 0xFFFF)
    size32 = 0xFFFFFFFF if force_zip64 else min(cd_size, 0xFFFFFFFF)
    off32 = 0xFFFFFFFF if force_zip64 else min(cd_offset, 0xFFFFFFFF)
    f.write(struct.pack('<IHHHHIIH', 0x06054B50, 0, 0, n16, n16, size32, off32, 0))
  return path
