import collections, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import bdm_b200.denoiser as D
from bdm_b200 import backend
D.PLAN_AHEAD = False
x, feats, cams = bench.make_inputs(int(os.environ.get("BDM_BATCH", "32")), 1234, "cuda:0")
sampler = bench.build_sampler(feats, cams, "cuda:0", mode="vanilla")
with torch.no_grad():
    sampler.pc2_step(x, 500)
    torch.cuda.synchronize()
    backend.profile_start()
    sampler.pc2_step(x, 500)
    prof = backend.profile_stop()
h = collections.OrderedDict()
for ms, shp in prof["groupnorm_act"]:
    a = h.setdefault(str(shp[0]) + " " + str(shp[5:]), [0, 0.0]); a[0] += 1; a[1] += ms
for k, v in sorted(h.items(), key=lambda kv: -kv[1][1]):
    print(k, v[0], round(v[1] * 1e3, 1), "us")
