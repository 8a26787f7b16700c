"""Build libbdm_b200.so (the C-ABI library, include/bdm_b200.h) in-tree with nvcc for sm_100a.

    python -m bdm_b200.build [--force] [-v]

nvcc cross-compiles without a GPU.  The library links cudart statically and has no torch dependency.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libbdm_b200.so")
SOURCES = ["abi.cu", "voxelize.cu", "voxel_coords.cu", "devoxelize.cu", "sampling.cu", "ball_query.cu", "grouping.cu",
           "three_nn.cu", "projection.cu", "knn_eval.cu", "groupnorm.cu", "sparse_conv.cu", "attention.cu", "attention_tc05.cu", "conv3_tc05.cu", "sampler.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "-Xcompiler", "-fPIC", "-cudart", "static"]


def _newest_source_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "bdm_b200.h")]
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False):
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= _newest_source_mtime():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(f"== {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libbdm_b200.so")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", SO] + objs)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
