"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/se3ds_geom.h declares, and fails loudly (no CPU fallback) without a CUDA device."""
import ctypes
import os
import re

import pytest
import torch

from se3ds_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  text = open(os.path.join(ROOT, 'include', 'se3ds_geom.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(se3ds_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported_and_bound():
  lib = _lib.load()
  names = _declared_symbols()
  assert len(names) >= 12
  for name in names:
    assert hasattr(lib, name), f'{name} declared in se3ds_geom.h but not exported'
    assert name in _lib.SIGNATURES, f'{name} has no ctypes signature'
  assert sorted(_lib.SIGNATURES) == names


def test_header_is_plain_c():
  """include/se3ds_geom.h must be consumable by a C compiler (cgo / JNI / ctypes-style bindings)."""
  import subprocess
  import tempfile
  with tempfile.NamedTemporaryFile('w', suffix='.c', delete=False) as f:
    f.write('#include "se3ds_geom.h"\nint main(void) { se3ds_ws* ws = 0; int (*v)(void) = se3ds_version; (void)ws; (void)v; return SE3DS_OK; }\n')
    path = f.name
  try:
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-pedantic', '-fsyntax-only',
                           '-I', os.path.join(ROOT, 'include'), path])
  finally:
    os.unlink(path)


def test_version_and_status_strings():
  lib = _lib.load()
  assert lib.se3ds_version() == 201
  assert lib.se3ds_status_string(0) == b'ok'
  assert b'shape' in lib.se3ds_status_string(_lib.ERR_BAD_SHAPE)


def test_no_torch_symbols_in_the_abi_library():
  """The boundary is plain C: the shared object must not link against libtorch / libc10."""
  import subprocess
  out = subprocess.run(['ldd', _lib.library_path()], capture_output=True, text=True).stdout
  names = [line.split()[0] for line in out.splitlines() if line.strip()]
  assert names and not any('torch' in n or 'c10' in n for n in names), names


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_fails_loudly_without_gpu():
  h = ctypes.c_void_p()
  rc = _lib.load().se3ds_ws_create(0, 0, 0, ctypes.byref(h))
  assert rc == _lib.ERR_CUDA
  with pytest.raises(_lib.Se3dsError):
    _lib.check(rc)
  from se3ds_b200.utils import pano_utils
  with pytest.raises(_lib.Se3dsError):
    pano_utils.mask_pano(torch.zeros((1, 8, 16, 3)))
  from se3ds_b200 import guidance
  with pytest.raises(_lib.Se3dsError):
    guidance.reproject(torch.zeros((1, 1, 8, 16, 3), dtype=torch.uint8), torch.zeros((1, 1, 8, 16)),
                       torch.zeros((1, 1, 3)), torch.zeros((1, 1, 3)))


def test_argument_errors_match_reference_exception_types():
  """Validation happens before any device work, so it is checkable on CPU
  (reference: pano_utils.py:190-202, point_cloud_utils.py:120-122, models.py:95-96)."""
  from se3ds_b200 import guidance
  from se3ds_b200.utils import pano_utils, point_cloud_utils
  with pytest.raises(ValueError):
    pano_utils.equirectangular_to_pointcloud(torch.zeros((1, 4, 8, 3, 1), dtype=torch.int32), torch.zeros((1, 4, 8)), 0, 20.0)
  with pytest.raises(ValueError):
    pano_utils.equirectangular_to_pointcloud(torch.zeros((1, 4, 8, 3), dtype=torch.uint8), torch.zeros((1, 4, 8)), -1, 20.0)
  with pytest.raises(AssertionError):
    pano_utils.equirectangular_to_pointcloud(torch.zeros((1, 4, 9, 3), dtype=torch.int32), torch.zeros((1, 4, 9)), -1, 20.0)
  with pytest.raises(ValueError):
    point_cloud_utils.project_to_feat(torch.zeros((1, 4, 5)), torch.zeros((1, 5, 3, 1)), 8, 8, 20.0, 0)
  with pytest.raises(ValueError):
    pano_utils.mask_pano(torch.zeros((8, 16, 3)))
  with pytest.raises(ValueError):
    guidance.GuidanceMemory(64, batch_size=2)


def test_intrinsic_matrix():
  from se3ds_b200 import constants
  from se3ds_b200.utils import point_cloud_utils
  k = point_cloud_utils.get_intrinsic_matrix(constants.HFOV)
  assert k.shape == (4, 4) and abs(k[0, 0].item() - 1.0) < 1e-6


def test_product_does_not_import_the_oracle():
  """Only tests/, smoke() and bench.py's CPU-baseline legs may import, link or execute oracle/."""
  pkg = os.path.join(ROOT, 'se3ds_b200')
  bad = re.compile(r'^\s*(from|import)\s+oracle\b|#\s*include.*oracle|CDLL\(.*oracle|oracle[/.]_?(build|ref)', re.M)
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h')):
        text = open(os.path.join(dirpath, f)).read()
        code = '\n'.join(l for l in text.splitlines() if not l.lstrip().startswith(('//', '#', '*', '/*')) or 'include' in l)
        assert not bad.search(code), f'{f} uses oracle/'
