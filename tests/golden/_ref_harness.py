"""Import the *real* reference (`/root/reference`, read-only) with `sys.modules` stubs for the
packages it imports but does not exercise on the hot path.

Only used by `tests/golden/make_golden.py` in the build container.  Nothing under `tests/` that
runs with `-m gpu`, `bench.py` or `smoke()` may import this module: `/root/reference` does not
exist on the GPU box.
"""
from __future__ import annotations

import os
import sys
import types

import torch.nn as nn

REFERENCE_ROOT = os.environ.get("KOD_REFERENCE_ROOT", "/root/reference")


def _stub(name: str, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install() -> bool:
    """Make `import kod...` work.  Returns False when the reference tree is absent."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "kod")):
        return False
    if "kod" in sys.modules:
        return True
    sys.dont_write_bytecode = True  # the reference tree is read-only

    _stub("albumentations")
    _stub("albumentations.pytorch", ToTensorV2=object)
    _stub("albumentations.core")
    _stub("albumentations.core.composition", TransformsSeqType=list)

    class _LightningModule(nn.Module):
        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    lp = _stub(
        "lightning.pytorch",
        LightningModule=_LightningModule,
        LightningDataModule=object,
        Trainer=object,
    )
    _stub("lightning", pytorch=lp, Callback=object, seed_everything=lambda *a, **k: None)
    _stub("lightning.pytorch.callbacks", Callback=object)
    _stub("lightning.pytorch.utilities", rank_zero_only=lambda f: f)
    _stub("lightning.pytorch.utilities.types", STEP_OUTPUT=object)
    _stub("lightning.pytorch.loggers", Logger=object)

    class _Evaluator:  # constructible only; its arithmetic is third-party (pycocotools), absent here
        def __init__(self, *a, **k):
            pass

    _stub("vision_evaluation")
    _stub("vision_evaluation.evaluators", CocoMeanAveragePrecisionEvaluator=_Evaluator)

    sys.path.insert(0, REFERENCE_ROOT)
    return True
