"""Import the reference toolkit's own data / loop modules (dev container only).

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` exists in the dev container and not
on the GPU box, so this is used by ``tests/golden/make_golden.py`` (fixture
generation) and by not-gpu tests that skip when the tree is absent.

Recipe: SURVEY.md Appendix B - stub the absent third-party modules and the 13
first-party model files the reference imports but does not ship, then import
``utils``, ``datasets`` and ``model_utils`` unmodified.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VITCNN_REFERENCE_ROOT", "/root/reference")

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
        for mod, names in _MISSING_FIRST_PARTY.items():
            if mod not in sys.modules:
                m = types.ModuleType(mod)
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
