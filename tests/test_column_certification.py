"""The certified COLUMN of K2's fast projection (canon_math.cuh project_pixel_fast), emulated in
float32 on the CPU against the canonical chain: a column the scheme certifies must be the canonical
column.  Points are planted at k +- dx (1 +- eps) around column borders in all octants, over radii
1e-3 .. 1e4 and widths 10 .. 8192, with the approximate reciprocal off by up to +-2 ulp
(tests/tools/col_cert_proto.py).  The worst-case bound |fast - canonical| <= 0.53 dx is derived in
DESIGN.md section 4; the GPU side is tests/test_gpu_parity.py::test_planted_column_borders."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import ref_exact as E

F32 = np.float32
_spec = importlib.util.spec_from_file_location(
    'col_cert_proto', os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tools', 'col_cert_proto.py'))
proto = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(proto)


def test_emulated_canonical_chain_is_the_oracle():
  """canonical_fx (numpy float32) truncates to the column the C oracle computes."""
  rng = np.random.default_rng(1)
  h, w, m = 512, 1024, 200000
  xyz = (rng.normal(size=(3, m)) * 10.0 ** rng.uniform(-2, 2, m)).astype(F32)
  out = E.splat(np.concatenate([xyz, np.ones((1, m), F32)])[None], np.ones((1, m, 1), F32), h, w, 20.0, -1.0, 0.0, 0)
  valid = out['valid'][0].astype(bool)
  col = np.floor(proto.canonical_fx(xyz[0], xyz[1], xyz[2], w)).astype(np.int64)
  assert valid.mean() > 0.99
  assert np.array_equal(col[valid], (out['flat'][0] % w)[valid])


@pytest.mark.parametrize('w', [10, 1000, 1024, 4096, 8192])
def test_certified_columns_equal_canonical_columns(w):
  n, frac, wrong, dev, dev_canon = proto.check(w, n_cols=800 if w > 1024 else 1500)
  assert wrong == 0
  assert dev < 0.53, dev          # |fast - canonical| in units of the margin dx: the derived worst case
  assert dev_canon < 0.40, dev_canon
  assert frac > 0.75              # planted borders included, most points are still certified
