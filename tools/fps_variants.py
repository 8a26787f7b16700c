"""Time the register-FPS launch variants (points per thread x threads) at N=4096 -> M=1024, B=16."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from bdm_b200 import backend as B
from tests import cases
co = torch.as_tensor(cases.cloud(np.random.default_rng(1), 16, 4096, "shape")).cuda()
for _ in range(3): B.furthest_point_sampling(co, 1024)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): idx = B.furthest_point_sampling(co, 1024)
e1.record(); torch.cuda.synchronize()
print("us per call", e0.elapsed_time(e1) * 100, "checksum", int(idx.sum()))
''' % ROOT
for v in ("16", "8", "4"):
    env = dict(os.environ, BDM_FPS_VARIANT=v)
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print("variant", v, out.stdout.strip(), out.stderr.strip()[-200:])
