"""Build and load ``libvitcnn.so`` (the C-ABI CUDA library) and declare its ctypes prototypes.

The library is the product: there is no CPU or PyTorch fallback.  ``lib()`` raises if the
shared object is missing, and every wrapper raises ``RuntimeError`` on a non-zero status.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_uint8, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO_PATH = os.path.join(CSRC, "libvitcnn.so")
SOURCES = ["abi.cu", "conv_tc.cu", "conv_var.cu", "pack.cu", "gather_tma.cu", "transformer.cu", "tokens_tc.cu", "tokens_tc2.cu", "tokens_tm.cu", "lidar_stem.cu", "metrics.cu", "wgrad_tc.cu", "wgrad_small.cu", "train.cu", "tokens_bwd.cu"]
HEADERS = ["tokens_tc_common.cuh", "vc_common.cuh", "vc_kernels.h", "vc_tparams.h", "vc_tokens.cuh", os.path.join("..", "..", "include", "vitcnn.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale() -> bool:
    if not os.path.isfile(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.isfile(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into csrc/libvitcnn.so (nvcc cross-compiles without a GPU).
    One object per translation unit under csrc/build/ (compiled in parallel, reused while newer than
    the source and every header), then one link step."""
    if not force and not _stale():
        return SO_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [s for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS if os.path.isfile(os.path.join(CSRC, h)))
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        if (not force and os.path.isfile(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(os.path.join(CSRC, src)), hdr_t)):
            return obj, ""
        res = subprocess.run([nvcc] + cflags + ["-c", "-o", obj, src], cwd=CSRC, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        done = list(ex.map(compile_one, srcs))
    if verbose:
        print("".join(log for _, log in done))
    res = subprocess.run([nvcc] + NVCC_FLAGS + ["-o", SO_PATH] + [o for o, _ in done], cwd=CSRC, capture_output=True,
                         text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return SO_PATH


class VcModel(ctypes.Structure):
    """Mirror of ``struct vc_model`` (include/vitcnn.h)."""
    _fields_ = [
        ("C1", c_int32), ("C2", c_int32), ("P", c_int32), ("K", c_int32),
        ("S1", c_int32), ("S2", c_int32),
        ("nsplit_h", c_int32 * 3), ("nsplit_l", c_int32 * 3),
        ("w_h", c_void_p * 3), ("scale_h", c_void_p * 3), ("bias_h", c_void_p * 3),
        ("w_l", c_void_p * 3), ("scale_l", c_void_p * 3), ("bias_l", c_void_p * 3),
        ("tparams", c_void_p),
        ("lidar_blob", c_void_p),
        ("w_h1_border", c_void_p),
    ]


class VcTrain(ctypes.Structure):
    """Mirror of ``struct vc_train`` (include/vitcnn.h)."""
    _fields_ = [
        ("C1", c_int32), ("C2", c_int32), ("P", c_int32), ("K", c_int32),
        ("params", c_void_p), ("grads", c_void_p),
        ("off", c_int64 * 58),
        ("bn_running_mean", c_void_p * 7), ("bn_running_var", c_void_p * 7), ("bn_num_batches", c_void_p * 7),
        ("bn_eps", c_float), ("bn_momentum", c_float),
        ("blob_segments", c_void_p), ("n_blob_segments", c_int32),
        ("dropout", c_float), ("drop_seed", c_void_p),
    ]


_PROTOS = {
    "vc_abi_version": (c_int32, []),
    "vc_lidar_blob_bytes": (c_int64, []),
    "vc_last_error": (c_char_p, []),
    "vc_sps_rows": (c_int64, [c_int32, c_int32]),
    "vc_launch_count": (c_int64, []),
    "vc_profile_begin": (c_int32, []),
    "vc_profile_end": (c_int32, [POINTER(ctypes.c_double), POINTER(c_int64), c_int32]),
    "vc_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32, c_int32]),
    "vc_tparams_layout": (c_int32, [c_int32, c_int32, POINTER(c_int64), c_int32]),
    "vc_gather_patches_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                        c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                        c_void_p]),
    "vc_scene_index": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                 c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vc_confusion_matrix": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int64, c_int32, ctypes.c_uint64, c_void_p,
                                      c_void_p]),
    "vc_minmax_normalise": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "vc_pack_sps": (c_int32, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int32, c_int32, c_int32,
                              c_void_p, c_int32, c_void_p]),
    "vc_conv_sps": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                              c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "vc_tokens_forward": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "vc_tokens_tc_scratch_bytes": (c_int64, [c_int32]),
    "vc_tokens_forward_tc": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_int64, c_void_p]),
    "vc_wgrad_workspace_bytes": (c_int64, [c_int32, c_int32]),
    "vc_wgrad_sps": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                               c_int64, c_void_p, c_int32, c_int32, c_int64, c_int64, c_int64, c_int32, c_void_p,
                               c_int32, c_void_p]),
    "vc_bn_forward": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_float,
                                c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int32, c_void_p]),
    "vc_bn_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vc_pack_conv_weight": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                      c_void_p, c_void_p]),
    "vc_pack_segments": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
    "vc_ce_loss": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float, c_void_p, c_void_p, c_void_p,
                             c_void_p]),
    "vc_adam_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                               c_float, c_int32, c_float, c_void_p]),
    "vc_adam_step_dev": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_float, c_void_p]),
    "vc_train_workspace_bytes": (c_int64, [POINTER(VcTrain), c_int32]),
    "vc_train_workspace_init": (c_int32, [POINTER(VcTrain), c_int32, c_void_p, c_int64, c_void_p]),
    "vc_train_forward": (c_int32, [POINTER(VcTrain), c_void_p, POINTER(c_int64), c_void_p, POINTER(c_int64), c_int32,
                                   c_void_p, c_int64, c_void_p, c_void_p]),
    "vc_train_forward_gather": (c_int32, [POINTER(VcTrain), c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                          c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "vc_train_backward": (c_int32, [POINTER(VcTrain), c_void_p, c_int32, c_void_p, c_int64, c_void_p]),
    "vc_forward_patches": (c_int32, [POINTER(VcModel), c_void_p, POINTER(c_int64), c_void_p, POINTER(c_int64),
                                     c_int32, c_void_p, c_int64, c_void_p, c_void_p]),
    "vc_scene_workspace_bytes": (c_int64, [POINTER(VcModel), c_int32, c_int32, c_int32]),
    "vc_scene_block": (c_int32, [c_int32, c_int32, c_int32]),
    "vc_scene_shared_depth": (c_int32, [POINTER(VcModel), c_int32, c_int32, c_int32, c_int64, c_int64]),
    "vc_scene_infer": (c_int32, [POINTER(VcModel), c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int32,
                                 c_int32, c_int64, c_int64, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
}

# symbols include/vitcnn.h declares; tests check every one is exported
EXPORTED = tuple(_PROTOS)

_lib = None


def lib() -> ctypes.CDLL:
    """The loaded library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        so_path = os.environ.get("VITCNN_LIB", SO_PATH)       # development: a variant built by tools/build_variants.py
        if not os.path.isfile(so_path):
            raise RuntimeError(
                f"{so_path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the ViT-CNN hot path)")
        L = ctypes.CDLL(so_path)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.vc_abi_version() != 4:
            raise RuntimeError("libvitcnn.so ABI version mismatch")
        _lib = L
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().vc_last_error()
        raise RuntimeError(f"{what} failed with status {status}: {msg.decode() if msg else ''}")


KERNEL_CLASSES = ("index", "pack", "conv_h1", "conv_h2", "conv_h3", "conv_lidar", "tokens", "halo", "bn", "wgrad",
                  "dgrad", "tokens_bwd", "misc")


def profile_begin() -> None:
    check(lib().vc_profile_begin(), "vc_profile_begin")


def profile_end() -> dict:
    """{class: (total_ms, launches)} since profile_begin() on this thread."""
    n = len(KERNEL_CLASSES)
    ms = (ctypes.c_double * n)()
    cnt = (c_int64 * n)()
    check(lib().vc_profile_end(ms, cnt, n), "vc_profile_end")
    return {k: (ms[i], int(cnt[i])) for i, k in enumerate(KERNEL_CLASSES)}


def tparams_layout(P: int, K: int) -> dict:
    """Byte offsets of the token-stage parameter blob (csrc/vc_tparams.h)."""
    L = lib()
    n = L.vc_tparams_layout(P, K, None, 0)
    buf = (c_int64 * n)()
    L.vc_tparams_layout(P, K, buf, n)
    v = list(buf)
    names = ["total", "wfus", "fus_scale", "fus_bias", "cls", "lnf_g", "lnf_b", "whead", "bhead", "pos"]
    out = dict(zip(names, v[:10]))
    per = ["wqkv", "wproj", "wfc1", "wfc2", "ln1_g", "ln1_b", "bqkv", "bproj", "ln2_g", "ln2_b", "bfc1", "bfc2"]
    out["layers"] = []
    k = 10
    while k < n:
        out["layers"].append(dict(zip(per, v[k:k + len(per)])))
        k += len(per)
    return out
