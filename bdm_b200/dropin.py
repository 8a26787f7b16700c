"""Make an unmodified checkout of mlpc-ucsd/BDM run on the B200 kernels.

The reference binds its native ops in exactly two places, both `from .backend import _backend`:
    experiments/model/pvcnn/modules/functional/backend.py:12-33   (PC^2 copy)
    experiments/pvd/modules/functional/backend.py:6-26            (PVD copy)
Each JIT-compiles the reference's CUDA sources at import.  `install()` pre-registers replacement
modules under those two names in sys.modules, so that when the reference's own
`modules/functional/*.py` execute `from .backend import _backend` they receive `bdm_b200.backend`
(same 12 function names, bindings.cpp:10-37) and nothing of the reference's extension is built or
loaded.  Everything above that line -- the reference's autograd Functions, nn.Modules, PVCNN2_PC2,
PVCNN2_PVD, PVCNN_fuse, main_blending.py / main_merging.py -- runs unchanged.

    import bdm_b200.dropin as dropin
    dropin.install("/path/to/BDM/experiments")      # before importing `model` / `pvd`
    from model.pvcnn.pvcnn import PVCNN2_PC2         # reference code, B200 kernels

`stub_packages=True` additionally registers path-only stand-ins for the `model` and `pvd` packages,
whose real __init__ files import hydra / diffusers / pytorch3d; use it to load only the PVCNN parts
in an environment without those dependencies (the unit tests do).
"""
import importlib.machinery
import os
import sys
import types

BACKEND_MODULE_NAMES = ("model.pvcnn.modules.functional.backend", "pvd.modules.functional.backend")


def _path_only_package(name, path):
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    mod.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
    mod.__spec__.submodule_search_locations = [path]
    return mod


def install(experiments_dir=None, backend=None, stub_packages=False):
    """Register the replacement `backend` modules (and optionally stub parent packages)."""
    if backend is None:
        from . import backend as _b
        backend = _b
    if experiments_dir is not None:
        experiments_dir = os.path.abspath(experiments_dir)
        if experiments_dir not in sys.path:
            sys.path.insert(0, experiments_dir)
        if stub_packages:
            for pkg in ("model", "pvd"):
                if pkg not in sys.modules:
                    sys.modules[pkg] = _path_only_package(pkg, os.path.join(experiments_dir, pkg))
    for name in BACKEND_MODULE_NAMES:
        shim = types.ModuleType(name)
        shim._backend = backend
        shim.__all__ = ['_backend']
        sys.modules[name] = shim
    return backend


def uninstall():
    for name in BACKEND_MODULE_NAMES:
        sys.modules.pop(name, None)
