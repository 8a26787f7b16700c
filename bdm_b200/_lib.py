"""ctypes binding of libbdm_b200.so (the C-ABI declared in include/bdm_b200.h).

There is NO fallback: if the library is missing or does not load, importing this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("BDM_LIB_PATH") or os.path.join(_HERE, "libbdm_b200.so")   # BDM_LIB_PATH: A/B builds (tools/)

if not os.path.exists(SO_PATH):
    raise ImportError(
        f"{SO_PATH} not found: build it with `python -m bdm_b200.build` (nvcc, sm_100a). "
        "bdm_b200 has no CPU or PyTorch fallback for its kernels.")

lib = ctypes.CDLL(SO_PATH)

_i = ctypes.c_int
_f = ctypes.c_float
_p = ctypes.c_void_p
_z = ctypes.c_size_t

_PROTOS = {
    "bdm_abi_version": (ctypes.c_int, []),
    "bdm_error_string": (ctypes.c_char_p, [_i]),
    "bdm_avg_voxelize_workspace_bytes": (_z, [_i, _i, _i]),
    "bdm_avg_voxelize": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _z, _p]),
    "bdm_voxel_plan": (_i, [_i, _i, _i, _p, _p, _p, _p, _z, _p]),
    "bdm_voxelize_coords": (_i, [_i, _i, _i, _i, _f, _p, _p, _p, _p]),
    "bdm_avg_voxelize_fill": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _z, _p]),
    "bdm_avg_voxelize_compact": (_i, [_i, _i, _i, _i, _p, _p, _p, _z, _p]),
    "bdm_grouping_into": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _i, _p]),
    "bdm_attention_workspace_bytes": (_z, [_i, _i, _i]),
    "bdm_attention": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _z, _p]),
    "bdm_attention_qkv": (_i, [_i, _i, _i, _p, _i, _p, _p, _p, _z, _p]),
    "bdm_sparse_conv3_gather": (_i, [_i, _i, _i, _i, _p, _p, _p, _i, _p, _p, _z, _p]),
    "bdm_sparse_conv3_stats_blocks": (_i, [_i]),
    "bdm_trilinear_devoxelize_cl": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "bdm_trilinear_devoxelize_cl_norm": (_i, [_i, _i, _i, _i, _p, _p, _p, _i, _p, _p, _p, _p]),
    "bdm_groupnorm_cl_sums_tiles": (_i, [_i, _i, ctypes.c_longlong]),
    "bdm_groupnorm_cl_sums": (_i, [_i, _i, ctypes.c_longlong, _i, _f, _i, _p, _p, _p, _p, _p, _i, _p, _p, _p]),
    "bdm_se_gate": (_i, [_i, _i, _i, _i, _f, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _p, _p, _p, _i, _p, _p]),
    "bdm_avg_voxelize_grad": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "bdm_trilinear_devoxelize_workspace_bytes": (_z, [_i, _i, _i]),
    "bdm_trilinear_devoxelize_plan": (_i, [_i, _i, _i, _p, _p, _z, _p]),
    "bdm_trilinear_devoxelize": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _z, _i, _p]),
    "bdm_trilinear_devoxelize_grad": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "bdm_gather_features": (_i, [_i, _i, _i, _i, _p, _p, _p, _p]),
    "bdm_gather_features_grad": (_i, [_i, _i, _i, _i, _p, _p, _p, _p]),
    "bdm_furthest_point_sampling_workspace_bytes": (_z, [_i, _i]),
    "bdm_furthest_point_sampling": (_i, [_i, _i, _i, _p, _p, _p, _z, _p]),
    "bdm_ball_query": (_i, [_i, _i, _i, _f, _i, _p, _p, _p, _p]),
    "bdm_grouping": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "bdm_grouping_grad": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "bdm_three_nearest_neighbors_interpolate": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "bdm_three_nn_search": (_i, [_i, _i, _i, _p, _p, _p, _p, _p]),
    "bdm_three_nn_interpolate": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "bdm_three_nearest_neighbors_interpolate_grad": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "bdm_surface_projection": (_i, [_i, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "bdm_surface_projection_hwc": (_i, [_i, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "bdm_groupnorm_workspace_bytes": (_z, [_i, _i, ctypes.c_longlong]),
    "bdm_groupnorm_tiles": (_i, [_i, _i, ctypes.c_longlong]),
    "bdm_groupnorm_act": (_i, [_i, _i, ctypes.c_longlong, _i, _f, _i, _i, _p, _p, _p, _p, _p, _p, _p, _z, _p]),
    "bdm_groupnorm_last_launches": (_i, []),
    "bdm_groupnorm_cl_supported": (_i, [_i, _i]),
    "bdm_groupnorm_cl_workspace_bytes": (_z, [_i, _i, ctypes.c_longlong]),
    "bdm_groupnorm_cl_tiles": (_i, [_i, _i, ctypes.c_longlong, _i]),
    "bdm_groupnorm_act_cl": (_i, [_i, _i, ctypes.c_longlong, _i, _f, _i, _p, _p, _p, _p, _p, _p, _p, _z, _i, _p]),
    "bdm_surface_projection_cf": (_i, [_i, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "bdm_nn_f64": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "bdm_nn_f64_reduce_blocks": (_i, [_i, _i]),
    "bdm_nn_f64_reduce": (_i, [_i, _i, _i, _i, ctypes.c_double, _p, _p, _p, _p, _p]),
    "bdm_sampler_update": (_i, [ctypes.c_longlong, _i, _p, _p, _p, _p, _i, _p, _p, _p]),
    "bdm_conv3_tc05_supported": (_i, [_i, _i, _i]),
    "bdm_conv3_tc05_plane_rows": (ctypes.c_longlong, [_i, _i]),
    "bdm_conv3_tc05_units": (_i, [_i, _i, _i]),
    "bdm_conv3_tc05_weight_bytes": (_z, [_i, _i]),
    "bdm_conv3_tc05_workspace_bytes": (_z, [_i, _i]),
    "bdm_conv3_tc05_prepare": (_i, [_i, _i, _p, _p, _p, ctypes.c_longlong, _p, _z, _p]),
    "bdm_groupnorm_swish_half_planar": (_i, [_i, _i, _i, _i, _f, _i, _p, _p, _p, _p, _p, _i, _p, _p, ctypes.c_longlong, _p]),
    "bdm_conv3_tc05_fill_planes": (_i, [_i, _i, _i, _i, _p, _p, _z, _p, _p, ctypes.c_longlong, _i, _p, _p]),
    "bdm_conv3_tc05_occ_words": (_i, [_i]),
    "bdm_avg_voxelize_compact_amax": (_i, [_i, _i, _i, _i, _p, _p, _p, _z, _p, _p]),
    "bdm_conv3_tc05": (_i, [_i, _i, _i, _i, _p, ctypes.c_longlong, _p, _p, _p, _p, _p, _z, _p, _p]),
}

EXPORTS = tuple(_PROTOS)

for _name, (_res, _args) in _PROTOS.items():
    _fn = getattr(lib, _name)  # AttributeError here == a symbol of the header is missing: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args

ABI_VERSION = lib.bdm_abi_version()


def check(rc):
    """Turn a non-zero return code of any bdm_* call into a RuntimeError (the reference's wrappers
    raise RuntimeError through TORCH_CHECK, src/utils.hpp:7-18)."""
    if rc != 0:
        raise RuntimeError(f"bdm_b200 error {rc}: {lib.bdm_error_string(rc).decode()}")
