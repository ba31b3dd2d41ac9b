"""The C-ABI boundary without a GPU: librecbox_b200.so builds / loads, exports every entry point include/recbox_b200.h
declares (and nothing the header does not), the ctypes table of recbox_b200/_lib.py covers exactly that set, and the
argument checks that run before any CUDA call answer with RBX_ERR_ARG + a message.  No compute calls."""
import ctypes
import os
import re
import subprocess

import pytest

from recbox_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "recbox_b200.h")


def declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)                 # comments mention entry points too
    return sorted(set(re.findall(r"\b(rbx_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_the_loader_binds():
    names = declared()
    assert len(names) >= 45
    assert sorted(_lib.SIGNATURES) == names


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    for name in declared():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\bT (rbx_[a-z0-9_]+)$", out, flags=re.M)))
    assert exported == declared(), "exported C symbols and the header disagree"
    assert lib.rbx_version() >= 100
    assert isinstance(lib.rbx_last_error(), bytes)


def test_argument_errors_come_before_any_cuda_call():
    lib = _lib.load()
    assert lib.rbx_interact_fwd(None, None, 4, 3, 4, 7, None) == -1            # unknown InnerProductInteraction mode
    assert b"mode 7" in lib.rbx_last_error()
    assert lib.rbx_topk_ip(None, None, 3, 10, 6, 2, 128, None, None, None, 0, None) == -1      # D % 4 != 0
    assert b"D=6" in lib.rbx_last_error()
    assert lib.rbx_shard_set_rank(99) == -1 and lib.rbx_shard_set_rank(-1) == 0
    assert lib.rbx_split_batch_f64(None, 4, 3, 2, None, None, None, None, 0, 0, None, None, None, None, None) == -1   # ld < n_cols
    # pure host helpers
    assert lib.rbx_topk_ws_bytes(1000, 100, 4096) >= 1000 * 4096 * 8
    assert lib.rbx_topk_ws_bytes(1000, 5000, 4096) == 0                          # k above the supported maximum
    assert lib.rbx_unique_ws_bytes(10_000_000) >= 2 * 10_000_000 // 8
    assert lib.rbx_shard_ws_bytes(1 << 20, 8) > 0


def test_empty_inputs_are_accepted_without_a_device():
    lib = _lib.load()
    assert lib.rbx_gather_rows(None, None, None, 0, 16, None) == 0
    assert lib.rbx_interact_fwd(None, None, 0, 3, 4, 0, None) == 0
    assert lib.rbx_topk_ip(None, None, 0, 10, 8, 2, 128, None, None, None, 0, None) == 0
