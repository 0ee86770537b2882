"""Build variants of libvitcnn.so that differ in -D switches of ONE translation unit (development tool for same-box
A/B timing: `VITCNN_LIB=<variant.so> python tools/time_tokens.py`).  Usage:
    python tools/build_variants.py tokens_tc.cu base:-DVC_TC_WARP_ARRIVE=0,-DVC_TC_LD16=0 arr:-DVC_TC_LD16=0 ...
Objects of the other translation units are taken from csrc/build/ (run __graft_entry__.build() first).
Outputs: vit-cnn_b200/csrc/variants/libvitcnn_<tag>.so (git-ignored, travels with gpurun)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vitcnn_b200  # noqa: E402,F401
from vitcnn_b200 import _lib  # noqa: E402


def main():
    src, specs = sys.argv[1], sys.argv[2:]
    _lib.build()
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    out_dir = os.path.join(_lib.CSRC, "variants")
    os.makedirs(out_dir, exist_ok=True)
    others = [os.path.join(_lib.CSRC, "build", s[:-3] + ".o") for s in _lib.SOURCES if s != src]
    cflags = [f for f in _lib.NVCC_FLAGS if f != "-shared"]
    for spec in specs:
        tag, _, defs = spec.partition(":")
        obj = os.path.join(out_dir, f"{src[:-3]}_{tag}.o")
        subprocess.run([nvcc] + cflags + [d for d in defs.split(",") if d] + ["-c", "-o", obj, src], cwd=_lib.CSRC, check=True)
        so = os.path.join(out_dir, f"libvitcnn_{tag}.so")
        subprocess.run([nvcc] + _lib.NVCC_FLAGS + ["-o", so, obj] + others, cwd=_lib.CSRC, check=True)
        print("built", so)


if __name__ == "__main__":
    main()
