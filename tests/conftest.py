import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

HAVE_REFERENCE = os.path.isdir('/root/reference/boxLCD')


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
  config.addinivalue_line('markers', 'reference: needs /root/reference (build container only)')


def _have_cuda():
  try:
    import torch
    return torch.cuda.is_available()
  except Exception:
    return False


def pytest_collection_modifyitems(config, items):
  skip_ref = pytest.mark.skip(reason='/root/reference not present on this machine')
  skip_gpu = None
  for item in items:
    if 'reference' in item.keywords and not HAVE_REFERENCE:
      item.add_marker(skip_ref)
    if 'gpu' in item.keywords:
      if skip_gpu is None:
        skip_gpu = False if _have_cuda() else pytest.mark.skip(reason='no CUDA device: the product path has no CPU fallback (run with -m gpu on a B200)')
      if skip_gpu:
        item.add_marker(skip_gpu)
