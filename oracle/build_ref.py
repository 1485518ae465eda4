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


REF_CU = "/root/reference/ultra/rspmm/source/rspmm.cu"
CUDA_MODULE = "rspmm_ref_cuda"


def build_cuda(verbose: bool = False) -> str | None:
    """Also compile the reference's OWN CUDA kernels (rspmm.cu, unmodified, where they lie) for sm_100a into
    oracle/_ref/rspmm_ref_cuda.so: the incumbent GPU implementation, timed beside ours by tools/ref_gpu_bench.py.
    nvcc cross-compiles without a GPU; the reference passes only -O3 to nvcc (rspmm.py:191-194)."""
    so_path = os.path.join(OUT_DIR, CUDA_MODULE + ".so")
    if not os.path.exists(REF_CU):
        return so_path if os.path.exists(so_path) else None
    if os.path.exists(so_path) and os.path.getmtime(so_path) >= os.path.getmtime(REF_CU):
        return so_path
    out = os.path.join(OUT_DIR, "cuda")
    os.makedirs(out, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    import torch  # noqa: F401
    from torch.utils import cpp_extension

    cpp_extension.load(
        CUDA_MODULE,
        [REF_SRC, REF_CU],
        extra_cflags=["-Ofast", "-fopenmp", "-DAT_PARALLEL_OPENMP", "-DCUDA_OP"],
        extra_cuda_cflags=["-O3", "-gencode", "arch=compute_100a,code=sm_100a"],
        build_directory=out,
        verbose=verbose,
    )
    built = os.path.join(out, CUDA_MODULE + ".so")
    if os.path.exists(built):
        os.replace(built, so_path)
    return so_path if os.path.exists(so_path) else None


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
    if "--cuda" in sys.argv:
        print(build_cuda(verbose="-v" in sys.argv))
