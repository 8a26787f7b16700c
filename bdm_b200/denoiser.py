"""Point-voxel denoisers of BDM on the B200 hot path: the callers either side of the sparse ops.

Three networks, all with the reference's module tree (attribute names, Sequential positions and
therefore state_dict keys), so reference checkpoints load with `load_state_dict`:

  PVCNN2_PC2   projection-conditioned reconstruction denoiser
               (reference: experiments/model/pvcnn/pvcnn.py:10-150, builders pvcnn_utils.py:14-185)
  PVCNN2_PVD   unconditional prior denoiser -- same architecture, no extra feature channels
               (reference: experiments/pvd/model/pvcnn_generation.py:172-245, pvd/__init__.py:300-332)
  PVCNNFuse    BDM-Merging network: PC^2 encoder + PVD encoder, 4 zero-initialised 1x1-conv
               projections, shared decoder   (reference: experiments/model/pvcnn/pvcnn_fuse.py:14-237)

The dense layers (Conv3d / GroupNorm / attention / 1x1 convs) run on cuDNN / cuBLAS through torch; every
sparse op goes through bdm_b200.modules -> bdm_b200.functional -> libbdm_b200.so.

The layer tables and the two builder functions below restate the reference's construction rules,
including its quirks, because they decide the parameter shapes:
  * in an SA stage after the first, only the FIRST PVConv of the configured `num_blocks` is
    instantiated (pvcnn_utils.py:98-103 `elif k == 0`);
  * the FP-stage attention predicate can never be true (pvcnn_utils.py:149 compares against the length
    of a list it has just shadowed), so FP PVConvs never carry attention;
  * the SA module's input width counts the time embedding only when the stage has no PVConv (:118).
"""
import contextlib
import math
import os

import torch
import torch.nn as nn

from . import functional as F
from .functional import geometry
from .functional import ops as _ops
from .modules import Attention, PointNetAModule, PointNetFPModule, PointNetSAModule, PVConv, SharedMLP
from .modules.point_voxel import _point_stream, coordinate_plan

# ((conv out_channels, num_blocks, voxel_resolution) | None, (num_centers, radius, num_neighbors, mlp widths))
SA_BLOCKS = [
    ((32, 2, 32), (1024, 0.1, 32, (32, 64))),
    ((64, 3, 16), (256, 0.2, 32, (64, 128))),
    ((128, 3, 8), (64, 0.4, 32, (128, 256))),
    (None, (16, 0.8, 32, (256, 256, 512))),
]
# (fp mlp widths, (conv out_channels, num_blocks, voxel_resolution) | None)
FP_BLOCKS = [
    ((256, 256), (256, 3, 8)),
    ((256, 256), (256, 3, 8)),
    ((256, 128), (128, 2, 16)),
    ((128, 128, 64), (64, 2, 32)),
]


_FREQ_TABLES = {}


def _frequency_table(embed_dim, device):
    """exp(-i * ln(1e4)/(half-1)) computed on the host in float64 and cast to fp32, like the reference's
    numpy expression; cached per device so that no host->device copy happens inside a forward (a
    pageable copy would break CUDA-graph capture)."""
    key = (embed_dim, str(device))
    tab = _FREQ_TABLES.get(key)
    if tab is None:
        half = embed_dim // 2
        step = math.log(10000) / (half - 1)
        tab = torch.exp(torch.arange(half, dtype=torch.float64) * -step).float().to(device)
        _FREQ_TABLES[key] = tab
    return tab


def timestep_embedding(embed_dim, timesteps, device):
    """Sinusoidal embedding [B] -> [B, embed_dim]: sin half then cos half.
    reference: pvcnn_utils.py:171-185"""
    assert timesteps.dim() == 1
    freqs = _frequency_table(embed_dim, device)
    arg = timesteps[:, None] * freqs[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)
    if embed_dim % 2 == 1:
        emb = nn.functional.pad(emb, (0, 1), "constant", 0)
    return emb


def _scaled(widths, r):
    return [[int(r * w) for w in ws] if isinstance(ws, (list, tuple)) else int(r * ws) for ws in widths]


def build_sa_layers(sa_blocks, extra_feature_channels, embed_dim=64, use_att=False, dropout=0.1, with_se=False,
                    normalize=True, eps=0, width_multiplier=1, voxel_resolution_multiplier=1):
    """-> (list of stages, input width of every stage, output width, number of final centres)"""
    r, vr = width_multiplier, voxel_resolution_multiplier
    width = extra_feature_channels + 3
    stages, stage_in_widths = [], []
    num_centers = None
    for stage_idx, (conv_cfg, sa_cfg) in enumerate(sa_blocks):
        stage_in_widths.append(width)
        parts = []
        convs_configured = 0
        feat_width = extra_feature_channels if stage_idx == 0 else width
        if conv_cfg is not None:
            c_out, num_blocks, resolution = conv_cfg
            c_out = int(r * c_out)
            for p in range(num_blocks):
                attention = (stage_idx + 1) % 2 == 0 and use_att and p == 0

                def make(c_in, c_out=c_out, attention=attention):
                    if resolution is None:
                        return SharedMLP(c_in, c_out)
                    return PVConv(c_in, c_out, kernel_size=3, resolution=int(vr * resolution), attention=attention,
                                  dropout=dropout, with_se=with_se, with_se_relu=True, normalize=normalize, eps=eps)

                if stage_idx == 0:
                    parts.append(make(width))
                elif convs_configured == 0:
                    parts.append(make(width + embed_dim))
                width = c_out
                convs_configured += 1
            feat_width = width
        num_centers, radius, num_neighbors, mlp_widths = sa_cfg
        mlp_widths = _scaled(mlp_widths, r)
        sa_in = feat_width + (embed_dim if convs_configured == 0 else 0)
        if num_centers is None:
            parts.append(PointNetAModule(in_channels=sa_in, out_channels=mlp_widths, include_coordinates=True))
        else:
            parts.append(PointNetSAModule(num_centers=num_centers, radius=radius, num_neighbors=num_neighbors,
                                          in_channels=sa_in, out_channels=mlp_widths, include_coordinates=True))
        width = parts[-1].out_channels
        stages.append(parts[0] if len(parts) == 1 else nn.Sequential(*parts))
    return stages, stage_in_widths, width, (1 if num_centers is None else num_centers)


def build_fp_layers(fp_blocks, in_channels, sa_in_channels, embed_dim=64, use_att=False, dropout=0.1,
                    with_se=False, normalize=True, eps=0, width_multiplier=1, voxel_resolution_multiplier=1):
    """-> (list of stages, output width).  `use_att` is accepted for signature parity; the reference's
    predicate never enables attention in FP stages (see module docstring)."""
    r, vr = width_multiplier, voxel_resolution_multiplier
    width = in_channels
    stages = []
    for fp_idx, (fp_widths, conv_cfg) in enumerate(fp_blocks):
        fp_widths = tuple(int(r * w) for w in fp_widths)
        parts = [PointNetFPModule(in_channels=width + sa_in_channels[-1 - fp_idx] + embed_dim, out_channels=fp_widths)]
        width = fp_widths[-1]
        if conv_cfg is not None:
            c_out, num_blocks, resolution = conv_cfg
            c_out = int(r * c_out)
            for _ in range(num_blocks):
                if resolution is None:
                    parts.append(SharedMLP(width, c_out))
                else:
                    parts.append(PVConv(width, c_out, kernel_size=3, resolution=int(vr * resolution),
                                        attention=False, dropout=dropout, with_se=with_se, with_se_relu=True,
                                        normalize=normalize, eps=eps))
                width = c_out
        stages.append(parts[0] if len(parts) == 1 else nn.Sequential(*parts))
    return stages, width


def build_classifier(in_channels, num_classes, dropout, width_multiplier=1):
    """SharedMLP(in,128) -> Dropout -> Conv1d(128,num_classes,1)   (pvcnn_utils.py:14-46 with
    out_channels=[128, dropout, num_classes], classifier=True, dim=2)"""
    hidden = int(width_multiplier * 128)
    return nn.Sequential(SharedMLP(in_channels, hidden), nn.Dropout(dropout), nn.Conv1d(hidden, num_classes, 1))


def _time_mlp(embed_dim):
    return nn.Sequential(nn.Linear(embed_dim, embed_dim), nn.LeakyReLU(0.1, inplace=True),
                         nn.Linear(embed_dim, embed_dim))


def _stage_parts(stage):
    return list(stage) if isinstance(stage, nn.Sequential) else [stage]


def _plan_block(block, pts):
    """coordinate-only work of one PVConv block: the voxel plan always; the x-slice binning of the plain
    devoxelization only when the block will not take the channels-last route (which has no use for it)"""
    v = block.voxelization
    norm_coords = coordinate_plan(pts, v.r, v.normalize, v.eps)[0]
    if not block.devoxelizes_channels_last(pts.shape[2]):
        F.devoxelize_plan(norm_coords, v.r)


def plan_geometry_ahead(cache, sa_layers, fp_layers, coords):
    """Issue every coordinate-only op of the forward on the cache's side stream, in dependency order:
    voxel plans of the first stage first (the main stream needs them immediately), then the FPS /
    ball-query pyramid, then the 3-NN searches and the decoder's voxel plans."""
    with cache.side_stream():
        levels = [coords]
        for stage in sa_layers:
            pts = levels[-1]
            for part in _stage_parts(stage):
                if isinstance(part, PVConv):
                    _plan_block(part, pts)
                elif isinstance(part, PointNetSAModule):
                    cen = F.furthest_point_sample(pts, part.num_centers)
                    for grouper in part.groupers:
                        F.ball_query(cen, pts, grouper.radius, grouper.num_neighbors)
                    levels.append(cen)
        if fp_layers is not None and len(levels) == len(sa_layers) + 1:
            cen = levels[-1]
            for fp_idx, stage in enumerate(fp_layers):
                pts = levels[-2 - fp_idx]
                for part in _stage_parts(stage):
                    if isinstance(part, PointNetFPModule):
                        F.three_nn_search(pts, cen)
                    elif isinstance(part, PVConv):
                        _plan_block(part, pts)
                cen = pts


PLAN_AHEAD = os.environ.get("BDM_PLAN_AHEAD", "1") != "0"   # module switch (tests / A-B timing)
TEMB_STREAM = os.environ.get("BDM_TEMB_STREAM", "1") != "0"


def _ahead_enabled(x):
    return PLAN_AHEAD and x.is_cuda and not torch.is_grad_enabled() and not _ops.REFERENCE_CALL_PATTERN


def _encode(sa_layers, features, coords, temb, temb_stream=None):
    """Run the SA pyramid; returns the bottleneck state and the per-stage (coords, input features).
    temb_stream: the stream `temb` is still being computed on (PVCNN2.forward); PVConv blocks only pass the embedding
    along, so the main stream joins in front of the first part that reads it."""
    coords_per_stage, feats_per_stage = [], []
    for i, stage in enumerate(sa_layers):
        feats_per_stage.append(features)
        coords_per_stage.append(coords)
        if temb_stream is not None and i == 0:
            state = (features, coords, temb)
            for part in _stage_parts(stage):
                if temb_stream is not None and not isinstance(part, PVConv):
                    temb_stream = _join(temb_stream, temb)
                state = part(state)
            features, coords, temb = state
            continue
        if temb_stream is not None:
            temb_stream = _join(temb_stream, temb)
        stage_in = features if i == 0 else torch.cat([features, temb], dim=1)
        features, coords, temb = stage((stage_in, coords, temb))
    if temb_stream is not None:
        _join(temb_stream, temb)
    return features, coords, temb, coords_per_stage, feats_per_stage


def _join(stream, tensor):
    main = torch.cuda.current_stream()
    main.wait_stream(stream)
    tensor.record_stream(main)
    return None


def _decode(fp_layers, features, coords, temb, coords_per_stage, skips_per_stage):
    for fp_idx, stage in enumerate(fp_layers):
        joined = torch.cat([features, temb], dim=1)
        joined._bdm_tail_is = temb      # (PointNetFPModule: the interpolated embedding is the tail of the interpolated `joined`)
        features, coords, temb = stage((coords_per_stage[-1 - fp_idx], coords, joined, skips_per_stage[-1 - fp_idx], temb))
    return features


class PVCNN2(nn.Module):
    """U-shaped point-voxel denoiser: 4 set-abstraction stages, global attention, 4 feature-propagation
    stages, per-point classifier.  forward(inputs f32[B,3+S,N], t [B]) -> f32[B,num_classes,N]."""
    sa_blocks = SA_BLOCKS
    fp_blocks = FP_BLOCKS

    def __init__(self, num_classes, embed_dim, use_att=True, dropout=0.1, extra_feature_channels=3,
                 width_multiplier=1, voxel_resolution_multiplier=1):
        super().__init__()
        assert extra_feature_channels >= 0
        self.embed_dim = embed_dim
        self.dropout = dropout
        self.width_multiplier = width_multiplier
        self.in_channels = extra_feature_channels + 3
        common = dict(embed_dim=embed_dim, use_att=use_att, dropout=dropout, with_se=True,
                      width_multiplier=width_multiplier, voxel_resolution_multiplier=voxel_resolution_multiplier)
        sa_layers, sa_in_widths, bottleneck, _ = build_sa_layers(self.sa_blocks, extra_feature_channels, **common)
        self.sa_layers = nn.ModuleList(sa_layers)
        self.global_att = Attention(bottleneck, 8, D=1) if use_att else None
        sa_in_widths[0] = extra_feature_channels  # the last FP stage sees only the extra features
        fp_layers, fp_width = build_fp_layers(self.fp_blocks, bottleneck, sa_in_widths, **common)
        self.fp_layers = nn.ModuleList(fp_layers)
        self.channels_fp_features = fp_width
        self.classifier = build_classifier(fp_width, num_classes, dropout, width_multiplier)
        self.embedf = _time_mlp(embed_dim)

    def forward(self, inputs, t):
        temb_stream = None
        if TEMB_STREAM and _ahead_enabled(inputs) and t.is_cuda:
            # the time MLP is ten tiny launches that depend on t alone: on the second stream, next to the first block
            temb_stream = _point_stream(inputs.device)
            temb_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(temb_stream) if temb_stream is not None else contextlib.nullcontext():
            temb = self.embedf(timestep_embedding(self.embed_dim, t, inputs.device).float())
            temb = temb[:, :, None].expand(-1, -1, inputs.shape[-1])
        coords = inputs[:, :3, :].contiguous()
        with geometry.scope(_ahead_enabled(inputs)) as cache:
            if cache is not None:
                plan_geometry_ahead(cache, self.sa_layers, self.fp_layers, coords)
            features, coords, temb, coords_per_stage, feats_per_stage = _encode(self.sa_layers, inputs, coords, temb,
                                                                                 temb_stream)
            feats_per_stage[0] = inputs[:, 3:, :]   # a view: the FP stage's torch.cat copies it anyway
            feats_per_stage[0]._bdm_slice_of = (inputs, 3)    # (layers.conv_no_bias_concat may widen it to an aligned width)
            if self.global_att is not None:
                features = self.global_att(features)
            features = _decode(self.fp_layers, features, coords, temb, coords_per_stage, feats_per_stage)
            return self.classifier(features)


class PVCNN2_PC2(PVCNN2):
    """reference: experiments/model/pvcnn/pvcnn.py:130-150"""


class PVCNN2_PVD(PVCNN2):
    """reference: experiments/pvd/__init__.py:300-332 (use_att and dropout are positional there)"""

    def __init__(self, num_classes, embed_dim, use_att, dropout, extra_feature_channels=3, width_multiplier=1,
                 voxel_resolution_multiplier=1):
        super().__init__(num_classes, embed_dim, use_att, dropout, extra_feature_channels, width_multiplier,
                         voxel_resolution_multiplier)


class PointCloudModel(nn.Module):
    """(B,N,C) <-> (B,C,N) adapter around PVCNN2_PC2 with the reference's output-layer initialisation.
    reference: experiments/model/point_cloud_model.py:14-65 (model_type='pvcnn' only; fp32 is forced
    there through an autocast context, here simply by running in fp32)."""

    def __init__(self, in_channels=3, out_channels=3, embed_dim=64, dropout=0.1, width_multiplier=1,
                 voxel_resolution_multiplier=1):
        super().__init__()
        self.model = PVCNN2_PC2(num_classes=out_channels, embed_dim=embed_dim, extra_feature_channels=in_channels - 3,
                                dropout=dropout, width_multiplier=width_multiplier,
                                voxel_resolution_multiplier=voxel_resolution_multiplier)
        self.model.classifier[-1].bias.data.normal_(0, 1e-6)
        self.model.classifier[-1].weight.data.normal_(0, 1e-6)

    def forward(self, inputs, t):
        return self.model(inputs.transpose(1, 2), t).transpose(1, 2)

    def forward_channel_first(self, inputs_cf, t):
        """inputs already (B,C,N) (ProjectionConditioner.get_input_channel_first) -> (B,N,out_channels)"""
        return self.model(inputs_cf, t).transpose(1, 2)


class PVCNNFuse(nn.Module):
    """BDM-Merging network.  The two encoders are the *shared* submodules of an existing PC^2 denoiser
    and PVD denoiser (as in the reference, which aliases them: pvcnn_fuse.py:29-35); the decoder,
    classifier and time MLP are initialised from the PC^2 network (:88, :100-104); `projs[k]` are
    conv-LeakyReLU-conv-zero_conv stacks of widths 64/128/256/512 (:111-123).

    forward(recon_inputs_with_cond f32[B,3+S,N], input_from_prior f32[B,3,N], t [B]) -> f32[B,3,N]

    Defined behaviour where the reference has none: the reference feeds the PVD encoder the time
    embedding left over from the PC^2 encoder, which by then is [B,64,16], and then gathers it with
    neighbour indices up to N-1 (pvcnn_fuse.py:176-186 -> modules/ball_query.py:30): an out-of-bounds
    read.  Here the PVD encoder gets the time embedding at full length N."""

    def __init__(self, pvd_net, pc2_net, num_classes=3, embed_dim=64, use_att=True, dropout=0.1,
                 extra_feature_channels=3, width_multiplier=1, voxel_resolution_multiplier=1):
        super().__init__()
        self.pvd_model_sa_layers = pvd_net.sa_layers
        self.pvd_model_global_att = pvd_net.global_att
        self.pc2_model_sa_layers = pc2_net.sa_layers
        self.pc2_model_global_att = pc2_net.global_att
        self.pc2_model_fp_layers = pc2_net.fp_layers
        self.pc2_model_classiifier = pc2_net.classifier  # (sic) the reference's attribute name
        self.pc2_model_embedf = pc2_net.embedf
        self.embed_dim = embed_dim
        common = dict(embed_dim=embed_dim, use_att=use_att, dropout=dropout, with_se=True,
                      width_multiplier=width_multiplier, voxel_resolution_multiplier=voxel_resolution_multiplier)
        _, sa_in_widths, bottleneck, _ = build_sa_layers(SA_BLOCKS, extra_feature_channels, **common)
        sa_in_widths[0] = extra_feature_channels
        fp_layers, fp_width = build_fp_layers(FP_BLOCKS, bottleneck, sa_in_widths, **common)
        self.fusion_decoder_fp_layers = nn.ModuleList(fp_layers)
        self.classifier = build_classifier(fp_width, num_classes, dropout, width_multiplier)
        self.embedf = _time_mlp(embed_dim)
        self.embedf.load_state_dict(self.pc2_model_embedf.state_dict())
        self.fusion_decoder_fp_layers.load_state_dict(self.pc2_model_fp_layers.state_dict())
        self.classifier.load_state_dict(self.pc2_model_classiifier.state_dict())
        projs = []
        for dim in (64, 128, 256, 512):
            conv1, conv2, zero_conv = nn.Conv1d(dim, dim, 1), nn.Conv1d(dim, dim, 1), nn.Conv1d(dim, dim, 1)
            for conv in (conv1, conv2):
                nn.init.normal_(conv.weight, mean=0.0, std=math.sqrt(2 / dim))
                nn.init.constant_(conv.bias, 0)
            for p in zero_conv.parameters():
                p.detach().zero_()
            projs.append(nn.Sequential(conv1, nn.LeakyReLU(0.02, inplace=True), conv2, zero_conv))
        self.projs = nn.ModuleList(projs)

    def forward(self, recon_inputs_with_cond, input_from_prior, t, mode='fusion_nstep'):
        n = recon_inputs_with_cond.shape[-1]
        temb_vec = self.embedf(timestep_embedding(self.embed_dim, t, recon_inputs_with_cond.device).float())
        temb_full = temb_vec[:, :, None].expand(-1, -1, n)
        coords_pc2 = recon_inputs_with_cond[:, :3, :].contiguous()
        coords_pvd = (input_from_prior if mode == 'fusion_nstep' else coords_pc2).clone()

        with geometry.scope(_ahead_enabled(recon_inputs_with_cond)) as cache:
            if cache is not None:
                plan_geometry_ahead(cache, self.pc2_model_sa_layers, self.fusion_decoder_fp_layers, coords_pc2)
                plan_geometry_ahead(cache, self.pvd_model_sa_layers, None, coords_pvd)
            return self._forward(recon_inputs_with_cond, coords_pc2, coords_pvd, temb_full)

    def _forward(self, recon_inputs_with_cond, coords_pc2, coords_pvd, temb_full):
        f_pc2, c_pc2, temb, coords_per_stage, pc2_skips = _encode(self.pc2_model_sa_layers, recon_inputs_with_cond,
                                                                 coords_pc2, temb_full)
        pc2_skips[0] = recon_inputs_with_cond[:, 3:, :].contiguous()
        if self.pc2_model_global_att is not None:
            f_pc2 = self.pc2_model_global_att(f_pc2)

        f_pvd, _, _, _, pvd_skips = _encode(self.pvd_model_sa_layers, coords_pvd.clone(), coords_pvd, temb_full)
        if self.pvd_model_global_att is not None:
            f_pvd = self.pvd_model_global_att(f_pvd)

        features = self.projs[-1](f_pvd) + f_pc2
        fused_skips = [pc2_skips[0]] + [proj(pvd_f) + pc2_f for pc2_f, pvd_f, proj in
                                        zip(pc2_skips[1:], pvd_skips[1:], self.projs)]
        features = _decode(self.fusion_decoder_fp_layers, features, c_pc2, temb, coords_per_stage, fused_skips)
        return self.classifier(features)
