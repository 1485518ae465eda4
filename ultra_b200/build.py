"""Build the C-ABI library ``ultra_b200/lib/libultra_rspmm.so`` with nvcc for sm_100a (in-tree, no torch headers).

``python -m ultra_b200.build [-f] [-v]``.  nvcc cross-compiles without a GPU; the resulting ``.so`` is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libultra_rspmm.so")
SOURCES = ["rspmm_api.cu", "rspmm_plan.cu", "nbfnet_update.cu", "relation_project.cu", "score_head.cu"]
HEADERS = ["rspmm_kernels.cuh", "rspmm_plan.h", os.path.join("..", "..", "include", "ultra_rspmm.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    for f in os.listdir(LIB_DIR):  # objects of sources that no longer exist (removed kernels) must not travel with the snapshot
        if f.endswith(".o") and f.replace(".o", ".cu") not in SOURCES:
            os.remove(os.path.join(LIB_DIR, f))
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
        return r.stderr

    logs = []
    if jobs:
        with ThreadPoolExecutor(len(jobs)) as ex:
            logs = list(ex.map(run, jobs))
    if jobs or force or _stale(LIB_PATH, objs):
        run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs)
    if ptxas_info:
        print("\n".join(logs))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv, ptxas_info="--ptxas" in sys.argv))
