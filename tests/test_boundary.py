"""CPU: the drop-in boundary.  libbdm_b200.so loads and exports every symbol include/bdm_b200.h
declares (no compute call is made: there is no GPU here); the product package never touches oracle/."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bdm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bdm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bdm_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    so = ctypes.CDLL(_lib.SO_PATH)
    missing = [s for s in declared if not hasattr(so, s)]
    assert not missing, f"declared in include/bdm_b200.h but not exported: {missing}"
    assert sorted(_lib.EXPORTS) == declared, "ctypes prototypes and header disagree"
    assert _lib.ABI_VERSION == 1


def test_error_strings():
    from bdm_b200 import _lib
    assert _lib.lib.bdm_error_string(0) == b"success"
    assert b"NULL" in _lib.lib.bdm_error_string(-1)
    assert b"workspace" in _lib.lib.bdm_error_string(-3)


def test_workspace_queries_need_no_gpu():
    from bdm_b200 import _lib
    L = _lib.lib
    assert L.bdm_avg_voxelize_workspace_bytes(16, 4096, 32) >= 16 * (4096 + 2048 + 8192 + 8192)
    assert L.bdm_avg_voxelize_workspace_bytes(1, 4096, 64) == 16          # generic path: no workspace
    assert L.bdm_furthest_point_sampling_workspace_bytes(4, 4096) == 16   # register path
    assert L.bdm_furthest_point_sampling_workspace_bytes(4, 20000) == 4 * 4 * 20000


def test_argument_errors_do_not_need_a_gpu():
    from bdm_b200 import _lib
    L = _lib.lib
    assert L.bdm_grouping(1, 1, 4, 2, 2, None, None, None, None) == -1          # NULL pointers
    assert L.bdm_ball_query(-1, 4, 2, 0.1, 2, None, None, None, None) == -2     # bad size
    assert L.bdm_grouping(0, 1, 4, 2, 2, None, None, None, None) == 0           # empty batch is a no-op
    # the fused routes validate their shape restrictions before touching the device
    assert L.bdm_attention(2, 32, 4096, None, None, None, None, None, 16, None) == -2        # 64 channels only
    assert L.bdm_attention(2, 64, 100, None, None, None, None, None, 16, None) == -2         # multiple of 128 tokens
    assert L.bdm_attention(0, 64, 4096, None, None, None, None, None, 16, None) == 0
    assert L.bdm_sparse_conv3_gather(1, 8, 64, 12, None, None, None, 0, None, None, 0, None) == -2 # r must be a power of two
    assert L.bdm_sparse_conv3_gather(1, 8, 64, 8, None, None, None, 0, None, None, 0, None) == -1  # NULL taps
    assert L.bdm_grouping_into(1, 4, 8, 2, 2, None, None, None, None, 3, 0, None) == -2      # slice outside the tensor
    assert L.bdm_groupnorm_cl_supported(64, 8) == 1 and L.bdm_groupnorm_cl_supported(48, 8) == 0
    assert L.bdm_groupnorm_act_cl(2, 48, 512, 8, 1e-5, 1, None, None, None, None, None, None, None, 0, 0, None) == -2
    assert L.bdm_avg_voxelize_compact(1, 4, 64, 64, None, None, None, 0, None) == -2         # needs the sorted plan (r^3 <= 32768)


def test_conv3_tc05_geometry_and_argument_errors_need_no_gpu():
    """flat padded grid of csrc/conv3_tc05.cu: plane rows, units and the shape restrictions, all host-side"""
    from bdm_b200 import _lib
    L = _lib.lib
    for r in (4, 8, 16, 32):
        q = r + 1
        guard = (q * q + q + 1 + 7) // 8 * 8
        sample_rows = (guard + q ** 3 + 7) // 8 * 8
        last_valid = ((r - 1) * q + (r - 1)) * q + (r - 1)
        assert L.bdm_conv3_tc05_units(128, 128, r) == (last_valid + 1 + 255) // 256          # one MMA per tap: 256 rows per unit
        assert L.bdm_conv3_tc05_units(64, 64, r) in ((last_valid + 1 + 255) // 256, (last_valid + 1 + 254) // 255)  # tap pairing: 255
        for b in (1, 3, 32):
            rows = L.bdm_conv3_tc05_plane_rows(b, r)
            # every sample's positions plus the rows its last 256-row unit reads beyond them fit
            assert rows >= guard + b * sample_rows + guard
            for cc in (64, 128):
                assert rows >= guard + (b - 1) * sample_rows + L.bdm_conv3_tc05_units(cc, cc, r) * 256 + q * q + q + 1
    assert L.bdm_conv3_tc05_supported(64, 64, 32) == 1 and L.bdm_conv3_tc05_supported(32, 128, 16) == 1
    assert L.bdm_conv3_tc05_supported(390, 32, 32) == 0      # reduction length: 32 or a multiple of 64
    assert L.bdm_conv3_tc05_supported(64, 256, 8) == 0       # 256 accumulator columns do not double-buffer in TMEM
    assert L.bdm_conv3_tc05_supported(64, 64, 12) == 0 and L.bdm_conv3_tc05_supported(64, 64, 64) == 0
    assert L.bdm_conv3_tc05_weight_bytes(64, 64) == 256 + 27 * 64 * 64 * 2
    rows = L.bdm_conv3_tc05_plane_rows(2, 16)
    assert L.bdm_conv3_tc05(2, 390, 32, 16, None, rows, None, None, None, None, None, 0, None, None) == -2     # unsupported widths
    assert L.bdm_conv3_tc05(2, 64, 64, 16, None, rows, None, None, None, None, None, 0, None, None) == -1      # NULL operands
    assert L.bdm_conv3_tc05(0, 64, 64, 16, None, rows, None, None, None, None, None, 0, None, None) == 0       # empty batch
    assert L.bdm_conv3_tc05_prepare(48, 64, None, None, None, 1, None, 0, None) == -2
    assert L.bdm_groupnorm_swish_half_planar(2, 64, 12, 8, 1e-5, 1, None, None, None, None, None, 1, None, None, rows, None) == -2
    assert L.bdm_conv3_tc05_fill_planes(2, 64, 1024, 64, None, None, 0, None, None, rows, 0, None, None) == -2  # needs the plan (r <= 32)
    assert L.bdm_groupnorm_cl_sums(2, 48, 512, 8, 1e-5, 1, None, None, None, None, None, 1, None, None, None) == -2
    assert L.bdm_trilinear_devoxelize_cl_norm(0, 64, 128, 16, None, None, None, 1, None, None, None, None) == 0


def test_product_never_imports_oracle():
    """A product path that routes through the oracle voids every parity claim: forbid it textually."""
    offenders = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bdm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "libbdm_oracle" in src \
                        or re.search(r"#include\s+[<\"].*oracle", src):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    """bdm_b200._lib must raise, not fall back, when the .so is absent."""
    import importlib
    import sys
    src = open(os.path.join(ROOT, "bdm_b200", "_lib.py")).read()
    pkg = tmp_path / "fakepkg"
    pkg.mkdir()
    (pkg / "__init__.py").write_text("")
    (pkg / "_lib.py").write_text(src)
    monkeypatch.syspath_prepend(str(tmp_path))
    try:
        importlib.import_module("fakepkg._lib")
        raised = False
    except ImportError:
        raised = True
    finally:
        sys.modules.pop("fakepkg._lib", None)
        sys.modules.pop("fakepkg", None)
    assert raised
