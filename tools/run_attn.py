import torch, sys
sys.path.insert(0, '.')
from bdm_b200 import backend as B
q, k, v = (torch.randn(int(__import__("os").environ.get("BDM_BATCH", "16")), 64, 4096, device="cuda") * 0.6 for _ in range(3))
for _ in range(2):
    B.attention(q, k, v)
torch.cuda.synchronize()
