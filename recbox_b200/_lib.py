"""Loader (and in-tree builder) for librecbox_b200.so, the C-ABI CUDA library of include/recbox_b200.h.

The product path has NO fallback: if the shared library is missing or a call fails, the caller gets
an exception (`RbxError`).  Nothing here imports `oracle/`.
"""
import ctypes
import glob
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("RBX_LIB_PATH") or os.path.join(_HERE, "librecbox_b200.so")   # override: tuning variants
HEADER = os.path.join(os.path.dirname(_HERE), "include", "recbox_b200.h")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xptxas=-v",
    "-Xcompiler", "-fPIC", "-shared", "-ldl",
]


class RbxError(RuntimeError):
    pass


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RbxError("nvcc not found; cannot build librecbox_b200.so")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if os.environ.get("RBX_LIB_PATH"):
        return False
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """Compile every csrc/*.cu for sm_100a into recbox_b200/librecbox_b200.so (in-tree, so the
    built library travels with the repo snapshot).  nvcc cross-compiles without a GPU."""
    out = out or LIB_PATH
    if not force and out == LIB_PATH and not needs_build():
        return LIB_PATH
    tmp = out + ".tmp.%d" % os.getpid()
    cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f] + ["-D" + d for d in defines] + ["-o", tmp] + sources()
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RbxError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stderr[-8000:]))
    os.replace(tmp, out)
    if verbose:
        print(proc.stderr)
    return out


_lib = None

_c = ctypes
_P = _c.c_void_p
_I64 = _c.c_int64
_I = _c.c_int
_F = _c.c_float

# name -> argtypes, in the order of include/recbox_b200.h
SIGNATURES = {
    "rbx_version": [],
    "rbx_last_error": [],
    "rbx_device_sm_count": [],
    "rbx_l2_set_persisting_bytes": [_c.c_longlong],
    "rbx_split_batch_f64": [_P, _I64, _I, _I64, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P],
    "rbx_pack_columns": [_P, _P, _P, _P, _P, _I, _I64, _I, _P, _P, _P],
    "rbx_unpack_ids_u16": [_P, _I64, _I, _P, _P, _P],
    "rbx_zero_f32": [_P, _I64, _P],
    "rbx_embed_fm_fwd": [_P] * 15 + [_I64, _I64, _I, _I, _I, _I, _P],
    "rbx_embed_fm_bwd": [_P] * 19 + [_I64, _I64, _I, _I, _I, _I, _P],
    "rbx_gather_rows": [_P, _P, _P, _I64, _I, _P],
    "rbx_scatter_add_rows": [_P, _P, _c.c_int32, _P, _I64, _I, _P],
    "rbx_segment_sum_rows": [_P, _P, _P, _c.c_int32, _P, _I64, _I, _P],
    "rbx_pooled_gather_fwd": [_P, _P, _I64, _P, _I64, _P, _I64, _I, _I, _I, _P],
    "rbx_pooled_gather_bwd": [_P, _I64, _P, _I64, _P, _c.c_int32, _P, _I64, _I, _I, _I, _P],
    "rbx_pool_fwd": [_P, _P, _P, _P, _I64, _I, _I, _I, _P],
    "rbx_pool_bwd": [_P, _P, _P, _I64, _I, _I, _I, _P],
    "rbx_rowdot_fwd": [_P, _P, _P, _I64, _I, _I, _P],
    "rbx_rowdot_bwd": [_P, _P, _P, _P, _P, _I64, _I, _I, _P],
    "rbx_interact_fwd": [_P, _P, _I64, _I, _I, _I, _P],
    "rbx_interact_bwd": [_P, _P, _P, _I64, _I, _I, _I, _P],
    "rbx_shard_ws_bytes": [_I64, _I],
    "rbx_shard_route": [_P, _I64, _I, _P, _c.c_size_t, _P, _P, _P, _P],
    "rbx_shard_permute": [_P, _P, _P, _I64, _I, _P],
    "rbx_shard_unroute": [_P, _P, _P, _I64, _I, _P],
    "rbx_shard_set_rank": [_I],
    "rbx_embed_fm_fwd_sharded": [_P, _P, _I] + [_P] * 12 + [_I64, _I64, _I, _I, _I, _I, _P],
    "rbx_embed_fm_bwd_sharded": [_P, _P, _P, _I] + [_P] * 15 + [_I64, _I64, _I, _I, _I, _I, _P],
    "rbx_embed_fm_fwd_sharded_rowlr": [_P, _I] + [_P] * 12 + [_I64, _I64, _I, _I, _I, _I, _P],
    "rbx_embed_fm_bwd_sharded_rowlr": [_P, _P, _I] + [_P] * 15 + [_I64, _I64, _I, _I, _I, _I, _P],
    "rbx_shard_push_ids": [_P, _P, _P, _P, _I, _I, _I64, _I64, _P],
    "rbx_shard_serve_rows": [_P, _P, _I, _P, _P, _P, _P, _I, _I64, _P],
    "rbx_shard_push_grads": [_P, _P, _P, _P, _P, _I, _I, _I64, _I, _I64, _P],
    "rbx_shard_apply_grads": [_P, _P, _P, _P, _P, _P, _I, _I64, _I, _P, _I, _P],
    "rbx_xs_tile_samples": [_I, _I],
    "rbx_xs_route": [_P, _I64, _I, _I64, _I, _I, _I, _I64, _P, _P, _P, _P, _P, _P, _P],
    "rbx_xs_barrier": [_P, _P, _P, _I, _I, _c.c_uint32, _P],
    "rbx_xs_serve": [_P, _I64, _P, _I64, _I, _P, _P, _I64, _I, _I, _P, _P, _P],
    "rbx_xs_consume": [_P] * 16 + [_I64, _I64, _I, _I, _I, _I, _I, _P],
    "rbx_xs_grad_push": [_P] * 16 + [_I, _P, _P, _P, _I64, _I64, _I, _I, _I, _I, _I, _P, _P, _P],
    "rbx_xs_apply": [_P, _P, _P, _P, _I64, _I, _P, _I64, _P, _I64, _I, _P],
    "rbx_peer_alloc": [_c.c_size_t, _P],
    "rbx_peer_free": [_P],
    "rbx_peer_export": [_P, _P],
    "rbx_peer_open": [_P, _P],
    "rbx_peer_close": [_P],
    "rbx_peer_can_access": [_I, _I],
    "rbx_sqnorm": [_P, _I64, _P, _P],
    "rbx_clip_coef": [_P, _F, _P, _P, _P],
    "rbx_adam_dense": [_P, _P, _P, _P, _I64, _P, _F, _F, _F, _F, _I, _P],
    "rbx_sqnorm_rows": [_P, _P, _P, _I64, _I, _P, _P],
    "rbx_optim_rows": [_P, _P, _P, _P, _P, _P, _I64, _I, _P, _I, _F, _F, _F, _F, _I, _I, _P],
    "rbx_unique_ws_bytes": [_I64],
    "rbx_unique_ids_i64": [_P, _I64, _I64, _P, _c.c_size_t, _P, _P, _P, _P, _P],
    "rbx_unique_ids_i32": [_P, _I64, _I64, _P, _c.c_size_t, _P, _P, _P, _P, _P],
    "rbx_power_sums_fwd": [_P, _P, _I64, _I, _I, _I, _P],
    "rbx_power_sums_bwd": [_P, _P, _P, _I64, _I, _I, _I, _P],
    "rbx_sample_negatives": [_I64, _I, _I64, _c.c_uint64, _P, _P, _P, _P, _P, _P, _P],
    "rbx_gemm_f32": [_P, _I64, _I, _P, _I64, _I, _P, _I64, _I64, _I64, _I64, _P, _I, _P, _I64, _I, _I, _P],
    "rbx_colsum_f32": [_P, _I64, _P, _I64, _I64, _I, _P],
    "rbx_nvls_allreduce_f32": [_P, _I64, _I, _I, _P],
    "rbx_topk_ws_bytes": [_I64, _I, _I64],
    "rbx_topk_ip": [_P, _P, _I64, _I64, _I, _I, _I64, _P, _P, _P, _c.c_size_t, _P],
    "rbx_rank_metrics": [_P, _I, _I64, _P, _P, _P, _P, _I, _P, _P, _I, _P, _P, _P, _P],
}


def load():
    """dlopen the library (building it first if sources are newer).  Raises RbxError if it cannot."""
    global _lib
    if _lib is not None:
        return _lib
    if needs_build():
        build()
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise RbxError("cannot load %s: %s" % (LIB_PATH, e))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise RbxError("%s does not export %s (stale build?)" % (LIB_PATH, name))
        fn.argtypes = argtypes
        fn.restype = {"rbx_last_error": ctypes.c_char_p, "rbx_shard_ws_bytes": ctypes.c_size_t, "rbx_unique_ws_bytes": ctypes.c_size_t, "rbx_topk_ws_bytes": ctypes.c_size_t,
                      "rbx_l2_set_persisting_bytes": ctypes.c_longlong}.get(name, ctypes.c_int)
    _lib = lib
    return lib


def check(rc, lib=None):
    if rc != 0:
        lib = lib or load()
        raise RbxError("librecbox_b200 error %d: %s" % (rc, lib.rbx_last_error().decode()))
