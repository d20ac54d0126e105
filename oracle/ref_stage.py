"""oracle/_ref: the reference's OWN python files of the hot path, staged for the `--impl reference` arm of bench.py.

TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.  The reference (craston/object_detection_cib, package `kod`) is
pure Python: there is nothing to compile, so "building" oracle/_ref is `stage()` -- the python files of the `kod`
package are packed UNMODIFIED from /root/reference into ONE build artefact, oracle/_ref/kod_reference.zip (imported
through zipimport).  oracle/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored, so
the GPU box -- which has no /root/reference -- can time the real
`Yolov5Loss` (kod/lightning/experiments/yv5_baseline/loss.py:46-248), `DefaultYolov5Experiment.get_detections`
(exp.py:70-102) and `non_max_suppression` (kod/core/nms.py:9-75) on its host cores.

`install(root)` makes `import kod` work from a reference tree: `sys.modules` stubs for the packages `kod` imports at
module level but does not execute on this path (albumentations, lightning, vision_evaluation: SURVEY.md Appendix C).
`__graft_entry__.build()` calls `stage()` where /root/reference exists; `tests/golden/make_golden.py` uses `install()`
on /root/reference itself.
"""
from __future__ import annotations

import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("KOD_REFERENCE_ROOT", "/root/reference")


STAGED_ZIP = os.path.join(STAGED_ROOT, "kod_reference.zip")


def stage(src_root: str = REFERENCE_ROOT, dst_zip: str = STAGED_ZIP) -> bool:
    """Pack the python files of `<src_root>/kod` into `dst_zip` (paths `kod/...`).  False when the reference is absent."""
    import zipfile

    src = os.path.join(src_root, "kod")
    if not os.path.isdir(src):
        return False
    os.makedirs(os.path.dirname(dst_zip), exist_ok=True)
    stale = os.path.join(os.path.dirname(dst_zip), "kod")
    if os.path.isdir(stale):  # loose files of an earlier staging
        shutil.rmtree(stale)
    tmp = dst_zip + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for d, dirs, names in os.walk(src):
            dirs[:] = sorted(x for x in dirs if x != "__pycache__")
            for n in sorted(names):
                if n.endswith(".py"):
                    full = os.path.join(d, n)
                    z.write(full, os.path.relpath(full, src_root))
    os.replace(tmp, dst_zip)
    return True


def staged() -> bool:
    return os.path.isfile(STAGED_ZIP)


def _stub(name: str, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install(root: str) -> bool:
    """Make `import kod...` resolve to `<root>/kod` (a directory tree) or to the staged zip (`root` = its path).  Returns
    False when neither exists."""
    if not (os.path.isdir(os.path.join(root, "kod")) or (root.endswith(".zip") and os.path.isfile(root))):
        return False
    if "kod" in sys.modules:
        return True
    import torch.nn as nn

    sys.dont_write_bytecode = True  # /root/reference is read-only

    _stub("albumentations")
    _stub("albumentations.pytorch", ToTensorV2=object)
    _stub("albumentations.core")
    _stub("albumentations.core.composition", TransformsSeqType=list)

    class _LightningModule(nn.Module):
        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    lp = _stub("lightning.pytorch", LightningModule=_LightningModule, LightningDataModule=object, Trainer=object)
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

    sys.path.insert(0, root)
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference tree not found: nothing staged")
