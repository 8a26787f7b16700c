import torch
import bdm_b200.modules.layers as L
import bdm_b200.modules.point_voxel as PV
from bdm_b200.denoiser import PVCNN2_PC2
torch.manual_seed(5)
net = PVCNN2_PC2(num_classes=3, embed_dim=64, extra_feature_channels=6).cuda().eval()
x = torch.randn(4, 9, 2048, device="cuda")
t = torch.tensor([500.0, 3.0, 999.0, 0.0], device="cuda")
def run(fused, sparse, tf32):
    L.FUSED_NORM_ACT = fused; PV.SPARSE_FIRST_CONV = sparse
    torch.backends.cudnn.allow_tf32 = tf32
    with torch.no_grad():
        return net(x, t)
ref = run(False, False, False).double()
pk = ref.abs().max().item()
for tf32 in (False, True):
    for sparse in (False, True):
        for fused in (False, True):
            y = run(fused, sparse, tf32)
            print(f"tf32={tf32} sparse={sparse} fused={fused}: err vs (plain, dense, fp32) = {(y.double()-ref).abs().max().item()/pk:.2e}")
