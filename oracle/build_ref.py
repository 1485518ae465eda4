"""TEST INFRASTRUCTURE ONLY - builds the *unmodified* reference CPU rspmm extension into oracle/_ref/.

Recipe (no reference sources are copied into this repo): the reference's own CPU translation unit
`/root/reference/ultra/rspmm/source/rspmm.cpp` (+ its headers `operator.cuh`, `rspmm.h`) is compiled where it
lies with exactly the flags the reference uses for its CPU-only build
(`ultra/rspmm/rspmm.py:181-202`: `-Ofast -fopenmp -DAT_PARALLEL_OPENMP`, CUDA file dropped because no
`CUDA_OP`).  The resulting pybind11 module `rspmm_ref.so` lands in `oracle/_ref/` (git-ignored, shipped to the
GPU box by gpurun).  `/root/reference` only exists in the build container; on the GPU box the prebuilt `.so` is
loaded by `oracle.ref.load_ref()`.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
"""
import os
import sys

REF_SRC = "/root/reference/ultra/rspmm/source/rspmm.cpp"
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
MODULE_NAME = "rspmm_ref"


def build(verbose: bool = False) -> str | None:
    """Build oracle/_ref/rspmm_ref.so if the reference sources are present. Returns the .so path or None."""
    so_path = os.path.join(OUT_DIR, MODULE_NAME + ".so")
    if not os.path.exists(REF_SRC):
        return so_path if os.path.exists(so_path) else None
    if os.path.exists(so_path) and os.path.getmtime(so_path) >= os.path.getmtime(REF_SRC):
        return so_path
    os.makedirs(OUT_DIR, exist_ok=True)
    import torch  # noqa: F401
    from torch.utils import cpp_extension

    cpp_extension.load(
        MODULE_NAME,
        [REF_SRC],
        extra_cflags=["-Ofast", "-fopenmp", "-DAT_PARALLEL_OPENMP"],
        build_directory=OUT_DIR,
        verbose=verbose,
    )
    return so_path


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
