"""Dense building blocks of the point-voxel networks.  These run on cuDNN / cuBLAS through torch
(the sparse hot path is elsewhere); they are restated here only so that the module tree -- class
names, parameter names, layer order -- matches the reference's and its checkpoints load.

reference: modules/shared_mlp.py:10-37, modules/se.py:8-19, modules/pvconv.py:12-63, modules/loss.py:8-10
"""
import contextlib
import os
import threading

import torch
import torch.nn as nn

from .. import functional as F
from ..functional import geometry
from ..functional import ops as _ops

# Fused GroupNorm+Swish kernel of libbdm_b200 for inference on CUDA (tolerance 1e-5 vs the torch pair,
# tests/test_dense_fused_gpu.py).  BDM_FUSED_NORM=0, autograd, CPU tensors or a foreign `_backend`
# (the reference's extension, the test oracle) fall back to nn.GroupNorm followed by Swish.
FUSED_NORM_ACT = os.environ.get("BDM_FUSED_NORM", "1") != "0"
# Fused online-softmax attention kernel (csrc/attention.cu) for the 64-channel attention block, inference
# on CUDA; fp32-equivalent (3xTF32).  BDM_FUSED_ATTENTION=0 keeps torch's matmul / softmax / matmul.
FUSED_ATTENTION = os.environ.get("BDM_FUSED_ATTENTION", "1") != "0"
# The dense second 3x3x3 convolution of a voxel stack on the tcgen05 tensor cores (csrc/conv3_tc05.cu) instead of
# cuDNN's TF32 kernels: the GroupNorm+Swish in front of it writes the convolution's fp16 operand, the convolution
# adds its bias and emits the statistics of the GroupNorm behind it.  Taken only when torch itself would run the
# convolution in TF32 (torch.backends.cudnn.allow_tf32), on grids of at least CONV3_TC05_MIN_R^3 voxels (cuDNN is
# faster on the 8^3 grids: too few 128-row tiles for 148 SMs).  BDM_CONV3_TC05=0 disables.
CONV3_TC05 = os.environ.get("BDM_CONV3_TC05", "1") != "0"
CONV3_TC05_MIN_R = int(os.environ.get("BDM_CONV3_TC05_MIN_R", "16"))
# the convolution's per-unit group statistics go to the next norm as they are ("groups": its prologue folds them) instead
# of through a folding kernel (True); BDM_CONV3_GROUP_STATS=0 selects the latter
CONV3_STATS = "groups" if os.environ.get("BDM_CONV3_GROUP_STATS", "1") != "0" else True
# Tail of a voxel stack whose last convolution made its own statistics: conv -> GroupNorm -> Swish -> SE -> devoxelize.
# The normalised grid is never written: one read-only pass yields the SE squeeze sums and the per-channel (A, B) of
# y = swish(x*A + B), and the devoxelization applies that to the 8 corner values it reads (bit-identical results).
# BDM_FUSED_TAIL_NORM=0 disables.
FUSED_TAIL_NORM = os.environ.get("BDM_FUSED_TAIL_NORM", "1") != "0"
# ... only on grids of at least this many voxels per side: the devoxelization then evaluates Swish 8 times per point
# instead of once per voxel (MUFU-bound: +35 us at R=32 / C=64 / 4096 points against -42 us for the norm pass; on the 16^3
# grids, where 8 * points = 2 * voxels, it loses)
FUSED_TAIL_NORM_MIN_R = int(os.environ.get("BDM_FUSED_TAIL_NORM_MIN_R", "32"))


_TF32_LOCK = threading.RLock()


@contextlib.contextmanager
def matmul_precision_of_convs():
    """Run the enclosed torch.matmul calls with the TF32 policy torch applies to convolutions
    (torch.backends.cudnn.allow_tf32): they stand in for 1x1 / 3x3x3 convolutions of the reference network.
    The matmul switch is process-global, so the (rare) flip is serialised behind a lock and always restored;
    when both policies already agree nothing is touched."""
    want = bool(torch.backends.cudnn.allow_tf32)
    with _TF32_LOCK:
        prev = bool(torch.backends.cuda.matmul.allow_tf32)
        if prev == want:
            yield
            return
        torch.backends.cuda.matmul.allow_tf32 = want
        try:
            yield
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev


# A 1x1 convolution whose input width is not a multiple of 4 runs on cuBLAS's fp32 SIMT kernels.  Splitting the
# reduction into an aligned head (tensor-op kernel) and a <= 3-wide tail pays only when the SIMT kernel is compute
# bound: the tail is a read-modify-write pass over the whole output.  Measured on B200 at 32 shapes
# (tools/gemm_probe.py): Cin=35 plain 123 us / split 214 us, Cin=67 71 / 114, Cin=131 64 / 51, Cin=390 86 / 54.
SPLIT_MIN_CHANNELS = 128


class Swish(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


def is_channels_last_3d(x):
    """a [B,C,D,H,W] tensor whose memory is [B,D,H,W,C] (and not also plain-contiguous)"""
    return x.dim() == 5 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last_3d)


def _fusable(x):
    return (FUSED_NORM_ACT and x.is_cuda and x.dtype == torch.float32 and x.dim() >= 3
            and (x.is_contiguous() or is_channels_last_3d(x))
            and not torch.is_grad_enabled() and not _ops.REFERENCE_CALL_PATTERN
            and hasattr(_ops._B, "groupnorm_act"))


def _groupnorm_act(y, gn, swish, conv_bias=None, **kw):
    """fused norm(+act) on a plain-contiguous or channels-last-3d tensor, layout preserved"""
    if is_channels_last_3d(y):
        if hasattr(_ops._B, "groupnorm_act_cl") and _ops._B.groupnorm_cl_supported(y.shape[1], gn.num_groups) \
                and not kw.get("max_over_last"):
            want_sums = kw.get("channel_sums", False)
            if want_sums and hasattr(_ops._B, "se_gate"):
                want_sums = "tiles"   # raw per-tile sums: SE3d.gate folds them inside its own kernel
            out = _ops._B.groupnorm_act_cl(y.permute(0, 2, 3, 4, 1), gn.num_groups, gn.weight, gn.bias, gn.eps, swish,
                                           conv_bias=conv_bias, channel_sums=want_sums, partials=kw.get("partials"))
            if isinstance(out, tuple):
                return out[0].permute(0, 4, 1, 2, 3), out[1]
            return out.permute(0, 4, 1, 2, 3)
        y = y.contiguous()
    kw.pop("partials", None)
    return _ops._B.groupnorm_act(y, gn.num_groups, gn.weight, gn.bias, gn.eps, swish, conv_bias=conv_bias, **kw)


def norm_act(norm, x, swish=True):
    """swish(norm(x)) for an nn.GroupNorm `norm`, through the fused kernel when possible."""
    if isinstance(norm, nn.GroupNorm) and _fusable(x):
        return _groupnorm_act(x, norm, swish)
    y = norm(x)
    return y * torch.sigmoid(y) if swish else y


_CONVS = (nn.Conv1d, nn.Conv2d, nn.Conv3d)


def _plain_conv3(m):
    return (isinstance(m, nn.Conv3d) and m.kernel_size == (3, 3, 3) and m.stride == (1, 1, 1) and m.padding == (1, 1, 1)
            and m.dilation == (1, 1, 1) and m.groups == 1 and m.padding_mode == 'zeros')


def conv3_tc05_applicable(conv, gn_before, channels, resolution):
    """Can `conv` (fed by GroupNorm `gn_before` + Swish over a channels-last [B,R,R,R,C] grid) take the tcgen05 route?"""
    return (CONV3_TC05 and hasattr(_ops._B, "conv3_tc05") and bool(torch.backends.cudnn.allow_tf32)
            and _plain_conv3(conv) and conv.in_channels == channels and isinstance(gn_before, nn.GroupNorm)
            and channels % gn_before.num_groups == 0 and 256 % channels == 0 and channels >= 8
            and resolution >= CONV3_TC05_MIN_R
            and _ops._B.conv3_tc05_supported(conv.in_channels, conv.out_channels, resolution))


_HALF_PLANES = {}     # (B, C, R, device) -> backend.HalfPlanes: scratch between a norm and its convolution, shared by
                      # every block of that shape (pad rows are zero once and for all; stream order serialises reuse)


def half_planes(b, c, r, device):
    key = (int(b), int(c), int(r), str(device))
    planes = _HALF_PLANES.get(key)
    if planes is None:
        planes = _HALF_PLANES[key] = _ops._B.HalfPlanes(b, c, r, device)
    return planes


def conv3_prepared(conv, gn, group_elems):
    """scales + fp16 weight stages of `conv` (whose input GroupNorm `gn` produces; gn None: the operand's producer
    sets the activation scale itself, bdm_conv3_tc05_fill_planes), cached per parameter version"""
    params = (conv.weight, gn.weight if gn is not None else None, gn.bias if gn is not None else None)
    key = tuple((p.data_ptr(), geometry.tensor_version(p), p.device) if p is not None else None for p in params) + (int(group_elems),)
    cached = getattr(conv, "_tc05_prepared", None)
    if cached is None or cached[0] != key:
        cached = (key, _ops._B.conv3_tc05_prepare(conv.weight.detach().contiguous(),
                                                 gn.weight.detach() if gn is not None and gn.weight is not None else None,
                                                 gn.bias.detach() if gn is not None and gn.bias is not None else None, group_elems))
        conv._tc05_prepared = cached
    return cached[1]


def conv_no_bias_concat(m, parts):
    """m(torch.cat(parts, dim=1)) without the bias and without materialising the concatenation, for a 1x1
    Conv1d `m`: W @ cat(x_i) = sum_i W[:, slice_i] @ x_i, one accumulating GEMM per part (each part's
    reduction split at a multiple of 4 like conv_no_bias).  parts: f32[B, C_i, L] with contiguous [C_i, L]
    blocks (batch stride free).  PointNetFPModule uses it for cat([interpolated, skip]) -- at the last FP
    stage the skip tensor alone is 100 MB."""
    w = m.weight
    # A part that is a channel slice base[:, c0:] of a wider tensor (the network input minus its coordinates: 387 of 390
    # channels) and whose width is not a multiple of 4 is widened downwards to the next multiple -- base[:, c0 - e:] --
    # against e zero columns in the weight: the same sums from ONE aligned tensor-op GEMM instead of an aligned head plus
    # a SIMT tail that re-reads and re-writes the whole output for 3 channels (36 us at the last FP stage).  The caller
    # vouches for the extra channels being ordinary data (`_bdm_slice_of` = (base, c0), set by PVCNN2.forward).
    front = []
    eff = []
    for p in parts:
        e = (-p.shape[1]) % 4
        origin = getattr(p, "_bdm_slice_of", None)
        if (e and origin is not None and p.shape[1] >= SPLIT_MIN_CHANNELS and origin[1] >= e and p.shape[2] % 4 == 0
                and origin[0].is_contiguous() and origin[0].shape[1] == origin[1] + p.shape[1]):
            eff.append(origin[0][:, origin[1] - e:])
            front.append(e)
        else:
            eff.append(p)
            front.append(0)
    parts = eff
    key = (w.data_ptr(), geometry.tensor_version(w), w.device, tuple(p.shape[1] for p in parts), tuple(front))
    cached = getattr(m, "_concat_weight", None)
    if cached is None or cached[0] != key:
        w2 = w.detach().reshape(m.out_channels, m.in_channels)
        if any(front):
            cols, src = [], 0
            for p, e in zip(parts, front):
                if e:
                    cols.append(w2.new_zeros((m.out_channels, e)))
                cols.append(w2[:, src:src + p.shape[1] - e])
                src += p.shape[1] - e
            assert src == m.in_channels
            w2 = torch.cat(cols, dim=1)
        pieces, off = [], 0
        for p in parts:
            ci = p.shape[1]
            head = ci & ~3 if ci >= SPLIT_MIN_CHANNELS else ci   # aligned head (+ tail) for wide parts, whole for narrow ones
            pieces.append((off, head, w2[:, off:off + head].contiguous()))
            if head < ci:
                pieces.append((off + head, ci - head, w2[:, off + head:off + ci].contiguous()))
            off += ci
        assert off == m.in_channels + sum(front)
        cached = (key, pieces)
        m._concat_weight = cached
    nb = parts[0].shape[0]
    with matmul_precision_of_convs():
        y, off_p = None, 0
        bounds = []
        for p in parts:
            bounds.append((off_p, p))
            off_p += p.shape[1]
        for off, width, wpiece in cached[1]:
            src_off, src = next((o, p) for o, p in reversed(bounds) if o <= off)
            xs = src[:, off - src_off: off - src_off + width]
            if y is None:
                y = torch.matmul(wpiece, xs)
            else:
                y.baddbmm_(wpiece.expand(nb, -1, -1), xs)
    return y


# A 1x1 convolution over 3 + C grouped channels (35, 67, 131, 259 in the SA stages) has a reduction length that is not a
# multiple of 4, which sends cuBLAS to its SIMT / split kernels (122 us for 35 -> 32 channels over 32 x 1024 x 32 columns,
# against ~65 us of HBM time).  When the consumer is the fused conv -> norm route, BallQuery allocates the grouped tensor
# with the channel count rounded up to a multiple of 4 (zero planes at the end) and the convolution multiplies by a
# zero-padded weight: same sums, aligned tensor-op GEMM.  BDM_PAD_GROUPED=0 disables.
PAD_GROUPED_CHANNELS = os.environ.get("BDM_PAD_GROUPED", "1") != "0"


def padded_channels(c):
    return (int(c) + 3) & ~3


def pads_grouped_channels(t, first_conv):
    """may a grouped tensor built for `first_conv` from CUDA tensor `t` carry zero-padded channels?"""
    return (PAD_GROUPED_CHANNELS and FUSED_NORM_ACT and t.is_cuda and t.dtype == torch.float32
            and not torch.is_grad_enabled() and not _ops.REFERENCE_CALL_PATTERN and hasattr(_ops._B, "groupnorm_act")
            and _pointwise(first_conv) and first_conv.bias is not None and first_conv.in_channels % 4 != 0)


def _pointwise(m):
    return (isinstance(m, (nn.Conv1d, nn.Conv2d)) and all(k == 1 for k in m.kernel_size)
            and all(v == 1 for v in m.stride) and all(v == 0 for v in m.padding)
            and all(v == 1 for v in m.dilation) and m.groups == 1)


def conv_no_bias(m, x):
    """m(x) without the bias.  A 1x1 convolution whose input width is not a multiple of 4 (390 at the first
    PC^2 layer) makes cuDNN/cuBLAS fall back to an `align1` SIMT GEMM (~0.7 ms there); splitting the
    reduction into an aligned head, which gets the tensor-op kernel the aligned layers get, plus a <=3-wide
    tail costs two GEMM launches and no copy of x.  TF32 follows the conv policy (cudnn.allow_tf32), as for
    every other convolution of the network."""
    cin = m.in_channels
    if _pointwise(m) and x.shape[1] != cin:
        # zero-padded input channels (BallQuery pads 3 + C to a multiple of 4 for exactly this): one aligned
        # tensor-op GEMM against the weight padded with zero columns (cached)
        cpad = x.shape[1]
        assert cpad == padded_channels(cin) and x.is_contiguous()
        w = m.weight
        key = (w.data_ptr(), geometry.tensor_version(w), w.device, cpad)
        cached = getattr(m, "_padded_weight", None)
        if cached is None or cached[0] != key:
            w2 = torch.zeros((m.out_channels, cpad), dtype=w.dtype, device=w.device)
            w2[:, :cin] = w.detach().reshape(m.out_channels, cin)
            cached = (key, w2)
            m._padded_weight = cached
        nb = x.shape[0]
        with matmul_precision_of_convs():
            y = torch.matmul(cached[1], x.reshape(nb, cpad, -1))
        return y.reshape(nb, m.out_channels, *x.shape[2:])
    if isinstance(m, nn.Conv3d) and is_channels_last_3d(x):
        # channels-last activations: hand cuDNN the weight in the same format (cached), so that it neither
        # transposes the weight on every call nor the activations around the kernel
        w = m.weight
        key = (w.data_ptr(), geometry.tensor_version(w), w.device)
        cached = getattr(m, "_cl_weight", None)
        if cached is None or cached[0] != key:
            cached = (key, w.detach().contiguous(memory_format=torch.channels_last_3d))
            m._cl_weight = cached
        return m._conv_forward(x, cached[1], None)
    if not (_pointwise(m) and cin % 4 != 0 and cin >= SPLIT_MIN_CHANNELS and x.is_cuda and x.is_contiguous()):
        return m._conv_forward(x, m.weight, None)
    w = m.weight
    key = (w.data_ptr(), geometry.tensor_version(w), w.device)
    cached = getattr(m, "_split_weight", None)
    if cached is None or cached[0] != key:
        w2 = w.detach().reshape(m.out_channels, cin)
        head = cin & ~3
        cached = (key, w2[:, :head].contiguous(), w2[:, head:].contiguous())
        m._split_weight = cached
    _, w_head, w_tail = cached
    head = w_head.shape[1]
    nb = x.shape[0]
    xr = x.reshape(nb, cin, -1)
    with matmul_precision_of_convs():
        y = torch.matmul(w_head, xr[:, :head])
        y.baddbmm_(w_tail.expand(nb, -1, -1), xr[:, head:])
    return y.reshape(nb, m.out_channels, *x.shape[2:])


class FusedSequential(nn.Sequential):
    """nn.Sequential (same children, same state_dict keys) that, for inference on CUDA, runs
        Conv -> GroupNorm [-> Swish] [-> SE3d | -> max over the last dim]
    through the fused kernel: the conv is issued without its bias (folded analytically into the norm
    statistics), norm + activation are one pass, the SE squeeze comes out of that same pass and a
    trailing max over neighbours replaces the full-size write.  Anything else runs module by module."""

    def forward(self, x, max_over_last=False, first_output=None, defer_gate=False, first_stats=None, first_biased=False,
                defer_norm=False):
        """first_output: the bias-less output of self[0] (a conv) when the caller computed it by other
        means (PVConv's sparse first convolution); `x` is then ignored.
        first_stats: per-channel statistics of first_output made by its producer (sparse_conv3_gather), handed
        to the norm that follows so that it does not read the tensor a second time.
        first_biased: first_output already includes self[0]'s bias (and first_stats are those of the biased tensor).
        defer_norm (with defer_gate): when the stack ends in conv -> GroupNorm -> Swish -> SE3d and the conv's statistics
        are known, return (UN-normalised conv output, gate, coefficients f32[B,C,2]) -- the caller's devoxelization applies
        the norm + Swish to the values it reads (trilinear_devoxelize_cl(norm_coef=)).
        defer_gate: when the stack ends in an SE3d gate, return (ungated grid, gate f32[B,C]) instead of
        multiplying the whole grid -- the caller applies the gate after its (linear) consumer."""
        mods = list(self)
        n = len(mods)
        i = 0
        reduced = False
        gate = None
        norm_coef = None
        pre, pre_stats, pre_biased = first_output, first_stats, bool(first_biased)   # output of mods[i] made by other means
        while i < n:
            m = mods[i]
            fusable = _fusable(x if pre is None else pre)
            if (fusable and isinstance(m, _CONVS) and m.bias is not None and i + 1 < n
                    and isinstance(mods[i + 1], nn.GroupNorm)):
                gn = mods[i + 1]
                swish = i + 2 < n and isinstance(mods[i + 2], Swish)
                nxt = i + (3 if swish else 2)
                y = pre if pre is not None else conv_no_bias(m, x)
                cbias = None if (pre is not None and pre_biased) else m.bias
                stats = pre_stats if pre is not None else None
                pre = pre_stats = None
                pre_biased = False
                # conv -> norm -> Swish [-> Dropout (inference: identity)] -> 3x3x3 conv: tcgen05 route for the second conv
                j = nxt
                if j < n and isinstance(mods[j], nn.Dropout) and not mods[j].training:
                    j += 1
                if (swish and stats is not None and j < n and is_channels_last_3d(y) and y.shape[2] == y.shape[3] == y.shape[4]
                        and conv3_tc05_applicable(mods[j], gn, y.shape[1], y.shape[2])):
                    conv2 = mods[j]
                    nb, nc, r = y.shape[0], y.shape[1], y.shape[2]
                    prepared = conv3_prepared(conv2, gn, (nc // gn.num_groups) * r ** 3)
                    planes = half_planes(nb, nc, r, y.device)
                    _ops._B.groupnorm_swish_half_planar(y.permute(0, 2, 3, 4, 1), gn.num_groups, gn.weight, gn.bias, gn.eps,
                                                        True, cbias, stats, prepared, planes)
                    out2, pre_stats = _ops._B.conv3_tc05(planes, prepared, conv2.out_channels, bias=conv2.bias, stats=CONV3_STATS)
                    pre, pre_biased = out2.permute(0, 4, 1, 2, 3), True
                    i = j
                    continue
                if (nxt < n and isinstance(mods[nxt], SE3d) and swish and defer_norm and defer_gate and nxt == n - 1
                        and FUSED_TAIL_NORM and y.shape[-1] >= FUSED_TAIL_NORM_MIN_R and stats is not None and is_channels_last_3d(y)
                        and hasattr(_ops._B, "groupnorm_cl_sums") and _ops._B.groupnorm_cl_supported(y.shape[1], gn.num_groups)):
                    sums, norm_coef = _ops._B.groupnorm_cl_sums(y.permute(0, 2, 3, 4, 1), gn.num_groups, gn.weight, gn.bias,
                                                                gn.eps, True, cbias, stats)
                    x, gate = y, mods[nxt].gate(y, channel_sums=sums)
                    nxt += 1
                elif nxt < n and isinstance(mods[nxt], SE3d) and swish:
                    y, sums = _groupnorm_act(y, gn, True, conv_bias=cbias, channel_sums=True, partials=stats)
                    if defer_gate and nxt == n - 1:
                        x, gate = y, mods[nxt].gate(y, channel_sums=sums)
                    else:
                        x = mods[nxt](y, channel_sums=sums)
                    nxt += 1
                elif (max_over_last and nxt == n and swish and y.dim() == 4
                      and _ops._B.groupnorm_max_supported(y.shape[-1])):
                    x = _groupnorm_act(y, gn, True, conv_bias=cbias, max_over_last=True)
                    reduced = True
                else:
                    x = _groupnorm_act(y, gn, swish, conv_bias=cbias, partials=stats)
                i = nxt
            elif pre is not None:
                x = pre if (m.bias is None or pre_biased) else pre + m.bias.view(1, -1, *([1] * (pre.dim() - 2)))
                pre = pre_stats = None
                pre_biased = False
                i += 1
            elif fusable and isinstance(m, nn.GroupNorm) and i + 1 < n and isinstance(mods[i + 1], Swish):
                x = norm_act(m, x, True)
                i += 2
            elif (isinstance(m, Attention) and i + 1 < n and isinstance(mods[i + 1], SE3d)
                  and m.fused_applicable(x)):
                # the attention block hands the SE squeeze its per-channel sums (no three chained means over the grid)
                y, sums = m.forward_fused(x, channel_sums=True)
                if defer_gate and i + 1 == n - 1:
                    x, gate = y, mods[i + 1].gate(y, channel_sums=sums)
                else:
                    x = mods[i + 1](y, channel_sums=sums)
                i += 2
            elif defer_gate and i == n - 1 and isinstance(m, SE3d):
                gate = m.gate(x)
                i += 1
            else:
                x = m(x)
                i += 1
        if max_over_last and not reduced:
            x = x.max(dim=-1).values
        if defer_gate:
            return (x, gate, norm_coef) if defer_norm else (x, gate)
        return x


class SharedMLP(nn.Module):
    """Stack of (1x1 conv, GroupNorm(8), Swish) over points [B,C,N] (dim=1) or neighbourhoods
    [B,C,M,U] (dim=2).  Parameters live at `layers.{3i}` (conv) and `layers.{3i+1}` (norm)."""

    def __init__(self, in_channels, out_channels, dim=1):
        super().__init__()
        try:
            conv = {1: nn.Conv1d, 2: nn.Conv2d}[dim]
        except KeyError:
            raise ValueError
        widths = list(out_channels) if isinstance(out_channels, (list, tuple)) else [out_channels]
        stack, c_in = [], in_channels
        for c_out in widths:
            stack += [conv(c_in, c_out, 1), nn.GroupNorm(8, c_out), Swish()]
            c_in = c_out
        self.layers = FusedSequential(*stack)

    def forward(self, inputs):
        # tuples carry (features, *passthrough): only the features go through the MLP
        if isinstance(inputs, (list, tuple)):
            head, *rest = inputs
            return (self.layers(head), *rest)
        return self.layers(inputs)

    def forward_max(self, features):
        """`self(features).max(dim=-1).values` with the max folded into the last norm+activation pass."""
        return self.layers(features, max_over_last=True)


class SE3d(nn.Module):
    """Squeeze-and-excitation gate over a voxel grid; two bias-free Linears at `fc.0` / `fc.2`."""

    def __init__(self, channel, reduction=8, use_relu=False):
        super().__init__()
        hidden = channel // reduction
        self.fc = nn.Sequential(nn.Linear(channel, hidden, bias=False),
                                nn.ReLU(True) if use_relu else Swish(),
                                nn.Linear(hidden, channel, bias=False),
                                nn.Sigmoid())

    def gate(self, inputs, channel_sums=None):
        """the per-(shape, channel) excitation f32[B,C]"""
        if channel_sums is not None:   # squeeze already produced by the fused norm+activation pass
            count = float(inputs.shape[2] * inputs.shape[3] * inputs.shape[4])
            if (hasattr(_ops._B, "se_gate") and channel_sums.is_cuda and not torch.is_grad_enabled()
                    and isinstance(self.fc[0], nn.Linear) and self.fc[0].bias is None and self.fc[2].bias is None):
                return _ops._B.se_gate(channel_sums.contiguous(), count, self.fc[0].weight, self.fc[2].weight,
                                       isinstance(self.fc[1], nn.ReLU))
            if channel_sums.dim() == 3:
                channel_sums = channel_sums.sum(dim=1)
            pooled = channel_sums / count
        else:                          # three chained means (z, y, x) like the reference, for identical rounding
            pooled = inputs.mean(-1).mean(-1).mean(-1)
        return self.fc(pooled)

    def forward(self, inputs, channel_sums=None):
        return inputs * self.gate(inputs, channel_sums)[:, :, None, None, None]


class Attention(nn.Module):
    """Dense single-head self-attention with un-scaled logits over voxels (D=3) or points (D=1),
    residual, GroupNorm, Swish.  Parameters: q, k, v, out (1x1 convs) and norm."""

    def __init__(self, in_ch, num_groups, D=3):
        super().__init__()
        assert in_ch % num_groups == 0
        conv = {3: nn.Conv3d, 1: nn.Conv1d}[D]
        self.q, self.k, self.v, self.out = (conv(in_ch, in_ch, 1) for _ in range(4))
        self.norm = nn.GroupNorm(num_groups, in_ch)
        self.nonlin = Swish()
        self.sm = nn.Softmax(-1)

    # -- inference on CUDA over a channels-last voxel grid: the whole block in five launches ------------------
    def fused_applicable(self, x):
        """x f32[B,64,D,H,W] in channels-last-3d memory, enough query tiles to fill the GPU, 1x1 Conv3d projections"""
        if not (FUSED_ATTENTION and _fusable(x) and is_channels_last_3d(x) and hasattr(_ops._B, "attention_qkv")
                and hasattr(_ops._B, "groupnorm_act_cl")):
            return False
        nb, nc = x.shape[:2]
        tokens = x.shape[2] * x.shape[3] * x.shape[4]
        return (all(isinstance(m, nn.Conv3d) and m.kernel_size == (1, 1, 1) and m.bias is not None
                    for m in (self.q, self.k, self.v, self.out))
                and isinstance(self.norm, nn.GroupNorm) and _ops._B.attention_supported(nc, tokens, nb)
                and _ops._B.groupnorm_cl_supported(nc, self.norm.num_groups))

    def _fused_weights(self):
        """[Wq;Wk;Wv]^T f32[C,3C], their biases f32[3C], Wo^T f32[C,C]; cached per weight version"""
        params = (self.q.weight, self.k.weight, self.v.weight, self.out.weight, self.q.bias, self.k.bias, self.v.bias)
        key = tuple((p.data_ptr(), geometry.tensor_version(p), p.device) for p in params)
        cached = getattr(self, "_fused", None)
        if cached is None or cached[0] != key:
            c = self.q.weight.shape[0]
            wqkv = torch.cat([m.weight.detach().reshape(c, c) for m in (self.q, self.k, self.v)], dim=0).t().contiguous()
            bqkv = torch.cat([m.bias.detach() for m in (self.q, self.k, self.v)]).contiguous()
            cached = (key, wqkv, bqkv, self.out.weight.detach().reshape(c, c).t().contiguous())
            self._fused = cached
        return cached[1:]

    def forward_fused(self, x, channel_sums=False):
        """The block on a channels-last grid (memory [B,T,C], T = voxels) without a single layout copy or bias kernel:
        one GEMM for q | k | v, the attention kernel (biases added as it reads, token-major output), one GEMM that
        adds the residual (x + mixed @ Wo^T), and the norm kernel (out-conv bias folded in, + Swish, + the SE squeeze
        sums when asked).  The reference runs 4 convolutions, 4 bias adds, a residual add, 2 matmuls, a softmax and
        3 layout copies here (modules/pvconv.py:40-63).  -> y (same shape / memory format as x) [, sums]"""
        nb, nc = x.shape[:2]
        spatial = tuple(x.shape[2:])
        x_cl = x.permute(0, 2, 3, 4, 1).reshape(nb, -1, nc)                 # a view of the channels-last memory
        wqkv, bqkv, wo = self._fused_weights()
        with matmul_precision_of_convs():                                   # they stand in for 1x1 convolutions
            qkv = torch.matmul(x_cl, wqkv)                                  # [B,T,3C]
            mixed = _ops._B.attention_qkv(qkv, bqkv)                        # [B,T,C]
            y = torch.baddbmm(x_cl, mixed, wo.expand(nb, -1, -1))           # residual + out projection (bias: below)
        gn = self.norm
        want = channel_sums
        if want and hasattr(_ops._B, "se_gate"):
            want = "tiles"
        out = _ops._B.groupnorm_act_cl(y, gn.num_groups, gn.weight, gn.bias, gn.eps, True, conv_bias=self.out.bias,
                                       channel_sums=want)
        if isinstance(out, tuple):
            return out[0].view((nb,) + spatial + (nc,)).permute(0, 4, 1, 2, 3), out[1]
        return out.view((nb,) + spatial + (nc,)).permute(0, 4, 1, 2, 3)

    # -- inference on CUDA over a short point set (the bottleneck: [B,512,16]): seven launches instead of thirteen ----------
    def small_applicable(self, x):
        return (FUSED_ATTENTION and _fusable(x) and x.dim() == 3 and x.is_contiguous() and x.shape[2] <= 64
                and x.shape[2] % 4 == 0 and hasattr(_ops._B, "groupnorm_act")
                and all(isinstance(m, nn.Conv1d) and _pointwise(m) and m.bias is not None
                        for m in (self.q, self.k, self.v, self.out)) and isinstance(self.norm, nn.GroupNorm))

    def _small_weights(self):
        """[Wq;Wk;Wv] f32[3C,C], their biases f32[1,3C,1], Wo f32[C,C]; cached per weight version"""
        params = (self.q.weight, self.k.weight, self.v.weight, self.out.weight, self.q.bias, self.k.bias, self.v.bias)
        key = tuple((p.data_ptr(), geometry.tensor_version(p), p.device) for p in params)
        cached = getattr(self, "_small", None)
        if cached is None or cached[0] != key:
            c = self.q.weight.shape[0]
            wqkv = torch.cat([m.weight.detach().reshape(c, c) for m in (self.q, self.k, self.v)], dim=0).contiguous()
            bqkv = torch.cat([m.bias.detach() for m in (self.q, self.k, self.v)]).reshape(1, 3 * c, 1).contiguous()
            cached = (key, wqkv, bqkv, self.out.weight.detach().reshape(c, c).contiguous())
            self._small = cached
        return cached[1:]

    def forward_small(self, x):
        """x f32[B,C,T], T <= 64: one GEMM for q | k | v + one bias add, the two attention matmuls and the softmax as the
        reference orders them, one GEMM that adds the residual, and the norm kernel (out-conv bias folded in, + Swish).
        The reference (and the plain route below) runs 4 convolutions, 4 bias kernels, a residual add and a SIMT GEMM for
        the transposed q here (modules/pvconv.py:40-63)."""
        nb, nc, _ = x.shape
        wqkv, bqkv, wo = self._small_weights()
        with matmul_precision_of_convs():                                   # they stand in for 1x1 convolutions
            qkv = torch.matmul(wqkv, x)                                     # [B,3C,T]
        qkv += bqkv
        q, k, v = qkv[:, :nc], qkv[:, nc:2 * nc], qkv[:, 2 * nc:]
        attn = self.sm(torch.matmul(q.transpose(1, 2), k))                  # [B,T,T]
        mixed = torch.matmul(v, attn.transpose(1, 2))                       # [B,C,T]
        with matmul_precision_of_convs():
            y = torch.baddbmm(x, wo.expand(nb, -1, -1), mixed)              # residual + out projection (bias: below)
        return _groupnorm_act(y, self.norm, True, conv_bias=self.out.bias)

    def forward(self, x):
        if self.fused_applicable(x):
            return self.forward_fused(x)
        if self.small_applicable(x):
            return self.forward_small(x)
        nb, nc = x.shape[:2]
        q, k, v = (proj(x).reshape(nb, nc, -1) for proj in (self.q, self.k, self.v))
        if (FUSED_ATTENTION and _fusable(x) and hasattr(_ops._B, "attention")
                and _ops._B.attention_supported(nc, q.shape[2], nb)):
            mixed = _ops._B.attention(q.contiguous(), k.contiguous(), v.contiguous()).reshape(x.shape)
        else:
            attn = self.sm(torch.matmul(q.transpose(1, 2), k))              # [B, T, T]
            mixed = torch.matmul(v, attn.transpose(1, 2)).reshape(x.shape)  # [B, C, ...]
        return norm_act(self.norm, self.out(mixed) + x, True)


class KLLoss(nn.Module):
    def forward(self, x, y):
        return F.kl_loss(x, y)
