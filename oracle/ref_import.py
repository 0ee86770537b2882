"""Import the reference toolkit's own data / loop modules, unmodified.

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` exists in the dev container and not on the GPU
box; ``tools/install_ref.sh`` (run by ``__graft_entry__.build()``) puts the four files the loops
live in (``utils.py``, ``datasets.py``, ``model_utils.py``, ``losses.py``) into the git-ignored
``baseline/_ref/``, which travels to the box.  Used by ``tests/golden/make_golden.py`` (fixture
generation), by the tests that drive the reference's own ``train()`` / ``val()`` / ``test()`` over
the CUDA module, and by ``bench.py``'s reference arm.

Recipe: SURVEY.md Appendix B - stub the absent third-party modules and every first-party model
file ``model_utils.py`` imports that is not there (13 are never shipped; under ``baseline/_ref``
none is), then import ``utils``, ``datasets`` and ``model_utils`` as they are.
"""
from __future__ import annotations

import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    cands = [os.environ.get("VITCNN_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "model_utils.py")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_root()

_THIRD_PARTY = ["seaborn", "spectral", "visdom", "matplotlib", "matplotlib.pyplot"]
_MISSING_FIRST_PARTY = {
    "model.CascadeMamba": ["CascadeRSMamba_complete"], "model.FICNN_VIT": ["FICNN_VIT"],
    "model.HybridSN": ["HybridSN"], "model.compare_method.MHST.MHST": ["MHST"],
    "model.Multimodality_Mamba.Mutimodality_Mamba7": ["Multimodality_Mamba"],
    "model.RSMamba": ["RSMamba_complete"], "model.SupConResNet": ["SupConResNet"],
    "model.compare_method.HCTnet": ["HCTnet"], "model.Selective": [],
    "model.Selective.fasternet": ["FasterNet"], "model.S2ENet": ["S2ENet"],
    "model.FI_CNN": ["FI_CNN"], "model.ResNet18": ["ResNet18"],
    "model.S2ENet_ResNet18": ["S2ENet_ResNet18"], "model.multiScaleCNN": ["multiScaleCNN"],
    "model.FI_CNN3D": ["FI_CNN3D"], "model.VIT": ["VIT"], "model.proposed": ["proposed"],
    "model.nncnet": ["moco_based_NNCNet"],
}
# model files the reference does ship (importable from /root/reference): stubbed too when the root is
# baseline/_ref, which holds the loop modules only
_SHIPPED_FIRST_PARTY = {
    "model.compare_method.MFT": ["MFT"], "model.compare_method.GLT_Net.GLT_Net": ["GLT"],
    "model.compare_method.spectralformer": ["SpectralFormer"], "model.compare_method.FusAtNet": ["FusAtNet"],
    "model.compare_method.EndNet": ["EndNet"], "model.compare_method.S2EFT": ["ViT"],
    "model.compare_method.DML_Hong": ["Early_fusion_CNN", "Middle_fusion_CNN", "Late_fusion_CNN", "Cross_fusion_CNN"],
}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model_utils.py"))


def import_reference(with_model_utils: bool = True):
    """Returns (utils, datasets, model_utils-or-None) of the reference."""
    if not available():
        raise FileNotFoundError(REFERENCE_ROOT)
    for n in _THIRD_PARTY:
        sys.modules.setdefault(n, types.ModuleType(n))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the repo's own host-side mirrors are called utils/datasets/model_utils inside
    # the package, never at top level, so these names resolve to the reference
    import utils as ref_utils          # noqa
    import datasets as ref_datasets    # noqa
    ref_model_utils = None
    if with_model_utils:
        stubs = dict(_MISSING_FIRST_PARTY)
        for mod, names in _SHIPPED_FIRST_PARTY.items():
            if not os.path.isfile(os.path.join(REFERENCE_ROOT, *mod.split(".")) + ".py"):
                stubs[mod] = names
        if not os.path.isdir(os.path.join(REFERENCE_ROOT, "model")):          # parent packages of the stubs
            for mod in list(stubs):
                parts = mod.split(".")
                for k in range(1, len(parts)):
                    stubs.setdefault(".".join(parts[:k]), [])
        for mod, names in stubs.items():
            if mod not in sys.modules:
                m = types.ModuleType(mod)
                m.__path__ = []
                for k in names:
                    setattr(m, k, type(k, (), {}))
                sys.modules[mod] = m
        import model_utils as ref_model_utils  # noqa
    return ref_utils, ref_datasets, ref_model_utils


class NullDisplay:
    """Stand-in for the visdom client ``train()`` plots to (model_utils.py:940-974)."""

    def line(self, *a, **k):
        return None

    def text(self, *a, **k):
        return None

    def heatmap(self, *a, **k):
        return None
