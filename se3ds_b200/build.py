"""Builds se3ds_b200/lib/libse3ds_geom.so (sm_100a, in-tree) with nvcc.

The library is the product: a C-ABI shared object (include/se3ds_geom.h) with no
torch types in its signatures.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, 'csrc', 'se3ds_geom.cu')
DEPS = [SRC, os.path.join(PKG, 'csrc', 'kernels.cuh'), os.path.join(PKG, 'csrc', 'canon_math.cuh'),
        os.path.join(ROOT, 'include', 'se3ds_geom.h')]
OUT = os.path.join(PKG, 'lib', 'libse3ds_geom.so')


def nvcc_path() -> str:
  for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
    if cand and os.path.exists(cand):
      return cand
  raise RuntimeError('nvcc not found')


def command(verbose_ptxas: bool = False):
  cmd = [nvcc_path(), '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         # canonical arithmetic: never contract a*b+c (the kernels use explicit _rn intrinsics too)
         '-fmad=false', '-prec-div=true', '-prec-sqrt=true', '-ftz=false',
         '-shared', '-Xcompiler', '-fPIC,-ffp-contract=off,-O2',
         '-I', os.path.join(ROOT, 'include'), '-o', OUT, SRC]
  if verbose_ptxas:
    cmd[1:1] = ['-Xptxas', '-v']
  return cmd


def needs_build() -> bool:
  if not os.path.exists(OUT):
    return True
  t = os.path.getmtime(OUT)
  return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
  if force or needs_build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = command(verbose)
    if verbose:
      print(' '.join(cmd))
    subprocess.check_call(cmd)
  return OUT


if __name__ == '__main__':
  build(force='--force' in sys.argv, verbose=True)
  print(OUT)
