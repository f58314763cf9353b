"""The C-ABI shared library builds for sm_100a, loads, and exports every symbol the header
declares with the signature the ctypes binding assumes (no compute calls: runs without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "ncf_b200.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    text = text.split('extern "C" {', 1)[1]
    protos = re.findall(r"\b(?:int|int64_t|const char\*)\s+(ncf_\w+)\s*\(([^;{]*)\)\s*;", text)
    return {name: args for name, args in protos}


def test_library_builds_and_exports_header_symbols():
    from ncf_b200 import _lib, build
    path = build.build()
    assert path.exists()
    lib = ctypes.CDLL(str(path))
    decl = declared_functions()
    assert len(decl) >= 19
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(decl) == set(_lib.SIGNATURES), "ctypes binding and header disagree on the surface"
    for name, args in decl.items():
        n_decl = 0 if args.strip() in ("", "void") else args.count(",") + 1
        assert n_decl == len(_lib.SIGNATURES[name][1]), f"{name}: arity mismatch"


def test_version_error_string_and_pure_host_queries():
    from ncf_b200 import _lib
    lib = _lib.load()
    assert lib.ncf_version() == 1
    assert isinstance(lib.ncf_last_error(), bytes)
    # tower buffer size: f=8, L=3 NeuMF: 64*32+32 + 32*16+16 + 16*8+8 + 16 + 1
    assert lib.ncf_tower_param_count(2, 8, 3) == 64 * 32 + 32 + 32 * 16 + 16 + 16 * 8 + 8 + 16 + 1
    assert lib.ncf_tower_param_count(0, 8, 3) == 64 * 32 + 32 + 32 * 16 + 16 + 16 * 8 + 8 + 8 + 1
    assert lib.ncf_tower_param_count(2, 8, 99) == -1
    assert lib.ncf_csr_workspace_bytes(1000, 10) > 1000 * 4


def test_struct_layout_matches_header():
    from ncf_b200 import _lib
    assert ctypes.sizeof(_lib.NcfModel) == 6 * 4 + 2 * 8 + 4 * 8 + 2 * 8 * 8 + 2 * 8
    assert ctypes.sizeof(_lib.NcfGrads) == 10 * 8
    assert ctypes.sizeof(_lib.NcfAdamState) == 13 * 8
    assert ctypes.sizeof(_lib.NcfAdamHyper) == 16


def test_bad_arguments_are_reported_not_crashed():
    from ncf_b200 import _lib
    lib = _lib.load()
    rc = lib.ncf_eval_rank(None, 4, 5, 6, None, None, None, None, None)  # k > C
    assert rc == -1 and b"k=6" in lib.ncf_last_error()
    m = _lib.NcfModel()
    assert lib.ncf_forward(ctypes.byref(m), None, None, 4, None, None, 0, None) == -1
    with pytest.raises(_lib.NcfError):
        _lib.check(rc, "ncf_eval_rank")
    # peer-memory entries: argument checks come before any CUDA call
    out = ctypes.c_void_p()
    assert lib.ncf_peer_alloc(0, ctypes.byref(out)) == -1
    assert lib.ncf_ipc_export(None, ctypes.create_string_buffer(64)) == -1
    assert lib.ncf_ipc_open(None, ctypes.byref(out)) == -1
    ptrs = (ctypes.c_void_p * 9)(*([16] * 9))
    hyper = _lib.NcfAdamHyper(1e-3, 0.9, 0.999, 1e-8)
    assert lib.ncf_adam_p2p(ptrs, ptrs, 16, 16, 0, 4, 9, 0, 1.0, 16, hyper, None) == -1
    assert b"world" in lib.ncf_last_error()
    assert lib.ncf_adam_p2p(ptrs, ptrs, 16, 16, 2, 4, 2, 0, 1.0, 16, hyper, None) == -1   # lo not a multiple of 4


def test_no_cpu_fallback():
    """Host tensors are refused: the product path never computes on the CPU."""
    import torch
    from ncf_b200 import _lib
    from ncf_b200.models import NCF
    model = NCF(5, 5, 8, 2, 0.0, "NeuMF-end").eval()
    with pytest.raises(_lib.NcfError), torch.no_grad():
        model(torch.zeros(2, dtype=torch.int64), torch.zeros(2, dtype=torch.int64))
