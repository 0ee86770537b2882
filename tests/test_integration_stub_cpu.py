"""INTEGRATION.md's ctypes stub must match the C ABI: the code block is extracted from the document and
executed with a recording stand-in for the library; the recorded call is checked against the prototype
``vit-cnn_b200/_lib.py`` declares for ``include/vitcnn.h`` (argument count and pointer / integer kinds)."""
import ctypes
import os
import re

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Fn:
    def __init__(self, log, name):
        self.log, self.name, self.restype = log, name, None

    def __call__(self, *args):
        self.log.append((self.name, args))
        return b"" if self.name == "vc_last_error" else 0


class _FakeCDLL:
    def __init__(self, path):
        self.path, self.calls, self._fns = path, [], {}

    def __getattr__(self, name):
        if name.startswith("vc_"):
            return self._fns.setdefault(name, _Fn(self.calls, name))
        raise AttributeError(name)


def _stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"<!-- ctypes-stub:begin.*?-->\s*```python\n(.*?)```\s*<!-- ctypes-stub:end -->", text, re.S)
    assert m, "INTEGRATION.md lost its marked ctypes stub"
    return m.group(1)


def test_stub_call_matches_the_declared_prototype(monkeypatch):
    import vitcnn_b200  # noqa: F401
    from vitcnn_b200 import _lib
    made = []
    monkeypatch.setattr(ctypes, "CDLL", lambda path: made.append(_FakeCDLL(path)) or made[-1])
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: type("S", (), {"cuda_stream": 0})())
    ns = {}
    exec(compile(_stub_source(), "INTEGRATION.md", "exec"), ns)
    H, W, C1, C2, P, n = 13, 14, 6, 1, 5, 3
    img1, img2 = torch.zeros(H, W, C1), torch.zeros(H, W, C2)
    gt, xy = torch.zeros(H, W, dtype=torch.uint8), torch.full((n, 2), 6, dtype=torch.int32)
    for ops in (None, torch.zeros(n, dtype=torch.uint8)):
        hsi, lid, lab = ns["gather"](img1, img2, gt, xy, P, ops=ops)
        assert hsi.shape == (n, C1, P, P) and lid.shape == (n, C2, P, P) and lab.dtype == torch.int64
    assert os.path.basename(made[0].path) == os.path.basename(_lib.SO_PATH)
    calls = [c for c in made[0].calls if c[0] == "vc_gather_patches_f32"]
    assert len(calls) == 2
    _, argtypes = _lib._PROTOS["vc_gather_patches_f32"]
    for _, args in calls:
        assert len(args) == len(argtypes), (len(args), len(argtypes))
        for a, t in zip(args, argtypes):
            if t is ctypes.c_void_p:
                assert a is None or isinstance(a, ctypes.c_void_p), (a, t)
            else:
                assert isinstance(a, int) and not isinstance(a, bool), (a, t)
    # positions: ..., xy, ops, n, P, center_mode, hsi, lidar, labels, stream (include/vitcnn.h)
    args = calls[0][1]
    assert args[9] is None and args[10:13] == (n, P, 1)
    assert calls[1][1][9] is not None


def test_header_declares_what_the_prototypes_bind():
    from vitcnn_b200 import _lib
    header = open(os.path.join(ROOT, "include", "vitcnn.h")).read()
    for name, (_, argtypes) in _lib._PROTOS.items():
        m = re.search(r"^(?:const\s+)?[a-z_0-9]+\s*\*?\s*" + name + r"\s*\(([^;{]*?)\)\s*;", header, re.S | re.M)
        assert m, f"{name} is bound by _lib.py but not declared in include/vitcnn.h"
        params = m.group(1).strip()
        count = 0 if params in ("", "void") else params.count(",") + 1
        assert count == len(argtypes), (name, count, len(argtypes))
