"""Build recipe for oracle/_ref: the UNMODIFIED reference CUDA extension, compiled for sm_100a.

TEST INFRASTRUCTURE ONLY.  Nothing under bdm_b200/ may import this.

The reference's 12-function pybind module `_pvcnn_backend`
(/root/reference/experiments/model/pvcnn/modules/functional/src/bindings.cpp:10-37) is compiled from
the sources *where they lie* under /root/reference (no source is copied into this repository); only
build products are written, and only into oracle/_ref/ (git-ignored, but shipped to the GPU box).

The reference's own recipe (functional/backend.py:12-31) is a torch JIT `load()` that hard-codes
`--compiler-bindir=/usr/bin/gcc-8`; gcc-8 does not exist in this image, so this recipe drops that
flag and pins the arch to sm_100a.  Everything else (flags -O3 -std=c++17, the 13 source files, the
module name) is the same.

Used for (1) generating tests/golden/*.npz on the GPU box (tests/golden/make_golden.py), (2) the
direct GPU-vs-reference parity tests, (3) the "reference kernels recompiled for sm_100a" timing that
bench.py reports beside our own numbers.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/experiments/model/pvcnn/modules/functional/src"
OUT = os.path.join(HERE, "_ref")

FILES = [
    "ball_query/ball_query.cpp", "ball_query/ball_query.cu",
    "grouping/grouping.cpp", "grouping/grouping.cu",
    "interpolate/neighbor_interpolate.cpp", "interpolate/neighbor_interpolate.cu",
    "interpolate/trilinear_devox.cpp", "interpolate/trilinear_devox.cu",
    "sampling/sampling.cpp", "sampling/sampling.cu",
    "voxelization/vox.cpp", "voxelization/vox.cu",
    "bindings.cpp",
]


def so_path():
    return os.path.join(OUT, "_pvcnn_backend.so")


def build(verbose=False):
    """Compile the reference extension if its sources are present; returns the .so path or None."""
    if os.path.exists(so_path()):
        return so_path()
    if not os.path.isdir(REF_SRC):
        return None  # GPU box: only the prebuilt file can be used
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load
    load(name="_pvcnn_backend", extra_cflags=["-O3", "-std=c++17"],
         extra_cuda_cflags=["-lineinfo"],
         sources=[os.path.join(REF_SRC, f) for f in FILES],
         build_directory=OUT, verbose=verbose, is_python_module=False)
    return so_path() if os.path.exists(so_path()) else None


def load_ref():
    """Import the prebuilt reference module (GPU box or here). Returns the module or None."""
    p = so_path()
    if not os.path.exists(p):
        return None
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location("_pvcnn_backend", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("oracle/_ref:", p)
