"""The drop-in boundary exercised from plain C ON THE GPU (-m gpu): tests/abi/abi_gpu_consumer.c allocates with the CUDA
runtime C API and goes plan_build -> set_edge_weight -> forward -> backward through include/ultra_rspmm.h, checking the
hand case of SURVEY.md section 8c bit for bit.  Round-1 review: the C consumer only touched host-side entry points."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_consumer(out_dir):
    lib_dir = os.path.join(ROOT, "ultra_b200", "lib")
    exe = os.path.join(str(out_dir), "abi_gpu_consumer")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    "-isystem", os.path.join(CUDA, "include"), os.path.join(ROOT, "tests", "abi", "abi_gpu_consumer.c"),
                    "-o", exe, "-L", lib_dir, "-lultra_rspmm", "-L", os.path.join(CUDA, "lib64"), "-lcudart",
                    "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + os.path.join(CUDA, "lib64")],
                   check=True, capture_output=True)
    return exe


def test_plain_c_consumer_on_gpu(cuda_device, tmp_path):
    exe = build_consumer(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "abi gpu consumer ok" in r.stdout
