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


def pytest_collection_modifyitems(config, items):
  skip_ref = pytest.mark.skip(reason='/root/reference not present on this machine')
  for item in items:
    if 'reference' in item.keywords and not HAVE_REFERENCE:
      item.add_marker(skip_ref)
