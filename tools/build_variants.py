#!/usr/bin/env python
"""Build tuning variants of librecbox_b200.so (compile-time knobs of csrc/embed_fm.cu) into
build/variants/; bench.py picks one with RBX_LIB_PATH=...  Used for the sweeps under profiles/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recbox_b200 import _lib  # noqa: E402

VARIANTS = {
    "base": [],
    "fwd_minb4": ["RBX_FWD_MINB=4"],
    "l2_keep": ["RBX_L2_HINTS=1"],
    "l2_stream": ["RBX_L2_HINTS=2"],
    "l2_both": ["RBX_L2_HINTS=3"],
    "l2_both_minb4": ["RBX_L2_HINTS=3", "RBX_FWD_MINB=4"],
    "h3": ["RBX_L2_HINTS=3"], "h6": ["RBX_L2_HINTS=6"], "h7": ["RBX_L2_HINTS=7"], "h2": ["RBX_L2_HINTS=2"],
    "noagg": ["RBX_BWD_WARP_AGG=0"],
    "nofuse": ["RBX_BWD_FUSE_NUM=0"],
    "fuse_noagg": ["RBX_BWD_WARP_AGG=0"],
    "rev": ["RBX_BWD_REVERSE=1"],
    "rev_l2both": ["RBX_BWD_REVERSE=1", "RBX_L2_HINTS=3"],
    "fwd_u13_b2": ["RBX_FWD_U=13", "RBX_FWD_MINB=2"],
    "fwd_u6_b4": ["RBX_FWD_U=6", "RBX_FWD_MINB=4"],
    "fwd_u4_b5": ["RBX_FWD_U=4", "RBX_FWD_MINB=5"],
    "bwd_u8_b3": ["RBX_BWD_U=8", "RBX_BWD_MINB=3"],
    "bwd_u2_b6": ["RBX_BWD_U=2", "RBX_BWD_MINB=6"],
    "bulk": ["RBX_BWD_BULK=1"],
    "bulk_remote": ["RBX_BWD_BULK=2"],
    "bulk_b3": ["RBX_BWD_BULK=1", "RBX_BWD_MINB=3"],
    "bulk_u2": ["RBX_BWD_BULK=1", "RBX_BWD_U=2", "RBX_BWD_MINB=5"],
}

if __name__ == "__main__":
    out_dir = os.path.join(ROOT, "build", "variants")
    os.makedirs(out_dir, exist_ok=True)
    names = sys.argv[1:] or list(VARIANTS)
    for name in names:
        path = _lib.build(force=True, defines=VARIANTS[name], out=os.path.join(out_dir, name + ".so"))
        print(name, "->", path)
