"""Generates tests/golden/*.npz: small seeded inputs with the outputs of the canonical oracle
(oracle/ref_exact.c).  Provenance: the reference itself cannot run here (TensorFlow is not
installed), so these are ORACLE outputs, frozen so that (a) the oracle cannot drift silently and
(b) the GPU tests have fixed vectors to reproduce bit for bit.  The known-answer vectors that come
from the reference's own tests (ray table, plane KAT, perturbation KATs) are asserted directly in
tests/test_oracle_kat.py.

  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_exact as X  # noqa: E402
from se3ds_b200 import synth  # noqa: E402

CASES = {
    # name: (n, s, p, h, dist, sweep, unproject_void, project_void, mask_first_frame, per_job_bin)
    'eval_metric_16': (2, 2, 1, 16, 'rand', False, -1, -1, True, False),
    'gan_manager_16': (2, 3, 1, 16, 'room', False, 0, -1, True, False),
    'pose_sweep_32': (1, 1, 4, 32, 'room', True, -1, -1, False, True),
    'ragged_5': (2, 2, 2, 5, 'rand', True, -1, -1, True, False),
}


def main():
  for name, (n, s, p, h, dist, sweep, uv, pv, mask, per_job) in CASES.items():
    inp = synth.make_inputs(n, s, p, h, seed=len(name), dist=dist, sweep=sweep)
    o = X.reproject(inp['rgb'], inp['depth'], inp['src_pos'], inp['tgt_pos'], unproject_void=uv, project_void=pv,
                    mask_first_frame=mask, per_job_bin=per_job)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), rgb=inp['rgb'], depth=inp['depth'], src_pos=inp['src_pos'],
                        tgt_pos=inp['tgt_pos'], unproject_void=uv, project_void=pv, mask_first_frame=mask,
                        per_job_bin=per_job, image=o['image'], proj_depth=o['depth'], mask=o['mask'], winner=o['winner'],
                        flat=o['flat'], rad=o['rad'])
    print(name, {k: v.shape for k, v in o.items() if hasattr(v, 'shape')})


if __name__ == '__main__':
  main()
