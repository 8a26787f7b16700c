"""Generate tests/golden/ref_<case>.npz from the REFERENCE's own CUDA kernels.

Run on the B200 box (the reference ops are CUDA-only: src/utils.hpp:7 CHECK_CUDA):

    gpurun -- python tests/golden/make_golden.py          # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/              # then commit

It imports oracle/_ref/_pvcnn_backend.so -- the unmodified reference extension compiled from
/root/reference/experiments/model/pvcnn/modules/functional/src by oracle/build_ref.py -- and runs
every seeded case of tests/cases.py:GOLDEN_CASES through it.  Each .npz stores the inputs and the
reference outputs, so the CPU-only test (tests/test_oracle_golden.py) needs neither a GPU nor
/root/reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import build_ref  # noqa: E402
from tests import cases, runners  # noqa: E402


def main():
    ref = build_ref.load_ref()
    if ref is None:
        raise SystemExit("oracle/_ref/_pvcnn_backend.so missing: run `python oracle/build_ref.py` first")
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name in cases.GOLDEN_CASES:
        inp = cases.build_case(name)
        out = runners.run_backend(name, ref, inp)
        blob = {f"in_{k}": v for k, v in inp.items()}
        blob.update({f"out_{k}": v for k, v in out.items()})
        np.savez_compressed(os.path.join(out_dir, f"ref_{name}.npz"), **blob)
        print(name, {k: v.shape for k, v in out.items()}, flush=True)


if __name__ == "__main__":
    main()
