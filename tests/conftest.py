"""pytest configuration: registers the `gpu` marker, skips `gpu` tests on machines without a CUDA
device (so a plain `pytest` is green there), and puts the repo root on sys.path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
  import torch
  if torch.cuda.is_available():
    return
  skip = pytest.mark.skip(reason='needs a CUDA device')
  for item in items:
    if 'gpu' in item.keywords:
      item.add_marker(skip)
