"""The certified-row scheme of the fast projection (canon_math.cuh project_pixel_fast, table of
se3ds_geom.cu get_tables), emulated in float32 on the CPU against the bit-exact oracle: no row that the
scheme certifies may differ from the canonical row, for approximate reciprocals / square roots that are
off by up to +-2 ulp.  (The GPU side is covered by test_certified_fast_projection and the parity tests.)"""
import importlib.util
import os

import numpy as np
import pytest

from oracle import ref_exact as E
from se3ds_b200 import synth

F32 = np.float32
_spec = importlib.util.spec_from_file_location(
    'row_cert_proto', os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tools', 'row_cert_proto.py'))
proto = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(proto)


@pytest.mark.parametrize('h,dist', [(5, 'rand'), (33, 'room'), (64, 'rand'), (256, 'room')])
def test_certified_rows_equal_canonical_rows(h, dist):
  w = 2 * h
  inp = synth.make_inputs(1, 1, 1, h, seed=h, dist=dist, sweep=True)
  rgb = inp['rgb'][0, 0].astype(np.int32)[None]
  xyz1, feats = E.equirectangular_to_pointcloud(rgb, inp['depth'][0:1, 0], -1, 20.0)
  xyz = ((xyz1[0, :3] + inp['src_pos'][0, 0][:, None]).astype(F32) - inp['tgt_pos'][0, 0][:, None]).astype(F32)
  out = E.splat(np.concatenate([xyz, np.ones((1, xyz.shape[1]), F32)])[None], feats, h, w, 20.0, -1.0, 0.0, 0)
  valid = out['valid'][0].astype(bool)
  row_canon = out['flat'][0] // w
  certified = 0
  for sr in (-1, 0, 1):
    for ss in (-1, 0, 1):
      row, certain = proto.kernel_rows(xyz[2], out['rad'][0], h, sr, ss)
      sel = valid & certain
      assert np.array_equal(row[sel], row_canon[sel])
      certified = max(certified, float(certain[valid].mean()))
  assert certified > 0.95   # the scheme must also be useful: almost every point is certified


def test_row_table_is_inside_the_rows():
  for h in (3, 64, 512, 2048):
    lo, hi = proto.host_table(h)
    edges = np.cos(np.arange(h + 1) * np.pi / h)
    assert np.all(lo.astype(np.float64) > edges[1:]) and np.all(hi.astype(np.float64) < edges[:-1])
    # rows narrower than two margins can never be certified (next to the poles at high resolution)
    assert np.all((lo < hi) | (edges[:-1] - edges[1:] < 4 * 2 * np.pi * 1e-6))
