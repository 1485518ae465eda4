"""CPU suite: the C-ABI library loads and exports every symbol include/ultra_rspmm.h declares, and the Python
boundary mirrors the reference's names and error behaviour (no compute calls - there is no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from ultra_b200 import build, _capi
    build.build()  # nvcc cross-compiles sm_100a without a GPU
    return _capi


def test_header_symbols_exported(capi):
    header = open(os.path.join(ROOT, "include", "ultra_rspmm.h")).read()
    declared = set(re.findall(r"\b(ultra_(?:rspmm|nbfnet|relation|score)_[a-z0-9_]+)\s*\(", header))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    handle = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert capi.lib().ultra_rspmm_version().startswith(b"ultra_rspmm sm_100a")


def test_library_is_sm100a_only(capi):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_workspace_bytes_is_pure_host(capi):
    lib = capi.lib()
    small = lib.ultra_rspmm_plan_workspace_bytes(10, 100, 3)
    big = lib.ultra_rspmm_plan_workspace_bytes(14541, 544230, 474)
    assert 0 < small < big < 1 << 30


def test_reference_surface_names():
    from ultra_b200 import rspmm as mod
    from ultra_b200.rspmm import rspmm as inner
    for s in ("Add", "Min", "Max"):
        for m in ("Mul", "Add"):
            cls = getattr(mod, "RSPMM%s%sFunction" % (s, m))
            assert issubclass(cls, torch.autograd.Function)
            assert hasattr(inner, "RSPMM%s%sFunction" % (s, m))
    import inspect
    sig = inspect.signature(mod.generalized_rspmm)
    assert list(sig.parameters) == ["edge_index", "edge_type", "edge_weight", "relation", "input", "sum", "mul"]
    assert sig.parameters["sum"].default == "add" and sig.parameters["mul"].default == "mul"


def test_unknown_op_raises_value_error():
    from ultra_b200.rspmm import generalized_rspmm
    a = torch.zeros(2, 3, dtype=torch.long)
    with pytest.raises(ValueError, match="No generalized rspmm implementation found for summation `mean`"):
        generalized_rspmm(a, a[0], torch.ones(3), torch.ones(2, 4), torch.ones(3, 4), sum="mean")
    with pytest.raises(ValueError, match="multiplication `rotate`"):
        generalized_rspmm(a, a[0], torch.ones(3), torch.ones(2, 4), torch.ones(3, 4), mul="rotate")


@pytest.mark.parametrize("mutate,match", [
    (lambda a: a.update(edge_index=a["edge_index"].int()), "expected scalar type Long"),
    (lambda a: a.update(edge_index=a["edge_index"][:1]), "size 2 at dimension 0"),
    (lambda a: a.update(edge_type=a["edge_type"][:2]), "argument #2 'edge_type'"),
    (lambda a: a.update(edge_weight=a["edge_weight"][:2]), "argument #3 'edge_weight'"),
    (lambda a: a.update(relation=a["relation"][:, :3]), "argument #4 'relation'"),
    (lambda a: a.update(relation=a["relation"].double()), "same type"),
    (lambda a: a.update(input=a["input"][0]), "argument #5 'input'"),
    (lambda a: a.update(input=a["input"].half(), relation=a["relation"].half(), edge_weight=a["edge_weight"].half()),
     "not implemented for 'Half'"),
])
def test_argument_checks_raise_runtime_error(mutate, match):
    """Same situations as rspmm_forward_check (/root/reference/ultra/rspmm/source/rspmm.cpp:15-28)."""
    from ultra_b200.rspmm import generalized_rspmm
    args = dict(edge_index=torch.zeros(2, 3, dtype=torch.long), edge_type=torch.zeros(3, dtype=torch.long),
                edge_weight=torch.ones(3), relation=torch.ones(2, 4), input=torch.ones(3, 4))
    mutate(args)
    with pytest.raises(RuntimeError, match=match):
        generalized_rspmm(**args)


def test_no_cpu_fallback_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from ultra_b200.rspmm import generalized_rspmm
    a = torch.zeros(2, 3, dtype=torch.long)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        generalized_rspmm(a, a[0], torch.ones(3), torch.ones(2, 4), torch.ones(3, 4))


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or /root/reference)."""
    pkg = os.path.join(ROOT, "ultra_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "oracle/_ref" not in src and "rspmm_ref" not in src, f


def test_tunables_are_host_only_and_validated(capi):
    """ultra_rspmm_set_tunable needs no GPU; unknown names / values are rejected with a message (include/ultra_rspmm.h)."""
    from ultra_b200 import _capi
    for name, value in (("l2_percent", 25), ("host_copy_split", 1), ("bwd_fused_relgrad", 1), ("bwd_fused_relgrad", 0)):
        _capi.set_tunable(name, value)
    with pytest.raises(RuntimeError, match="unknown tunable"):
        _capi.set_tunable("no_such_knob", 1)
    with pytest.raises(RuntimeError, match="unknown tunable"):  # round 1's process-global SM reservation is gone:
        _capi.set_tunable("sm_limit", 140)                      # it is a per-plan field now (plan_set_sm_limit)
    assert _capi.lib().ultra_rspmm_plan_set_sm_limit(None, 140) != 0  # null plan -> error code, no crash
    with pytest.raises(RuntimeError, match="unknown tunable"):  # measured in round 2 (no gain at any size) and removed
        _capi.set_tunable("item_edges", 16)
    for name in ("l2_prefetch_tiles", "l2_prefetch_copies"):    # same: measured (profiles/r02_l2_prefetch_ab.txt), removed
        with pytest.raises(RuntimeError, match="unknown tunable"):
            _capi.set_tunable(name, 4)
    for value in (1, 0, -1):                                     # plan-time order of small graphs; -1 = the default rule
        _capi.set_tunable("plan_relation_runs", value)
    _capi.set_tunable("work_queue", 1)


def test_score_head_entry_point_validates_without_a_gpu(capi):
    """ultra_score_head_f32 rejects unsupported sizes / null or misaligned pointers before touching the device, and the
    Python gate keeps the eager MLP for anything it does not cover (CPU tensors, autograd, other layer sizes)."""
    import torch
    from ultra_b200 import _capi, fused
    lib = _capi.lib()
    assert lib.ultra_score_head_f32(None, None, None, None, None, None, 4, 4, 32, 128, None) == 1
    assert b"dim == 64" in lib.ultra_rspmm_last_error()
    assert lib.ultra_score_head_f32(None, None, None, None, None, None, 4, 0, 64, 128, None) == 1
    assert lib.ultra_score_head_f32(None, None, None, None, None, None, 4, 4, 64, 128, None) == 1
    assert b"non-null" in lib.ultra_rspmm_last_error()
    assert lib.ultra_score_head_f32(None, None, None, None, None, None, 0, 4, 64, 128, None) == 0   # nothing to do
    mlp = torch.nn.Sequential(torch.nn.Linear(128, 128), torch.nn.ReLU(), torch.nn.Linear(128, 1))
    with torch.no_grad():
        assert not fused.score_head_ok(torch.zeros(2, 5, 64), mlp)          # CPU tensor: not this library's business


def test_plain_c_consumer(capi, tmp_path):
    """The boundary is a C ABI: a C99 program compiled against include/ultra_rspmm.h with -pedantic links to the library
    and calls its host-side entry points (what a cgo / JNI / N-API binding on the reference side would do)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "ultra_b200", "lib")
    exe = str(tmp_path / "abi_consumer")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "abi", "abi_consumer.c"), "-o", exe, "-L", lib_dir, "-lultra_rspmm",
                    "-Wl,-rpath," + lib_dir], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    assert r.stdout.startswith("ultra_rspmm sm_100a")


def test_numa_cpulist_parsing_and_graceful_fallback():
    """ultra_b200.numa: sysfs cpulist syntax; on a box without GPU / NUMA information the pinning is a no-op, not an error."""
    from ultra_b200 import numa
    assert numa._parse_cpulist("0-3,8,10-11") == {0, 1, 2, 3, 8, 10, 11}
    assert numa._parse_cpulist("") == set() and numa._parse_cpulist(None) == set()
    info = numa.pin_to_gpu_numa(0)
    assert "node" in info
