"""Import the UNMODIFIED reference hot-path modules from a reference checkout.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py and by
tests/test_oracle_vs_reference.py to pin the oracle.  The reference tree exists
only in the build container (/root/reference); on the GPU box this returns None
and the committed golden vectors are the pin.

`pycocotools` is not installed; the reference imports it at module scope
(utils/sbp_utils.py:8-9, utils/spm_utils.py:9-10) but the numeric path never
calls it, so three stub modules are registered before the import.
"""
import importlib
import os
import sys
import types

_CANDIDATES = (os.environ.get("POSE_REF"), "/root/reference")


def reference_root():
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "utils", "sbp_utils.py")):
            return c
    return None


def _stub_pycocotools():
    if "pycocotools" in sys.modules:
        return
    pkg = types.ModuleType("pycocotools")
    coco = types.ModuleType("pycocotools.coco")
    cocoeval = types.ModuleType("pycocotools.cocoeval")
    coco.COCO = type("COCO", (), {"__init__": lambda self, *a, **k: None})
    cocoeval.COCOeval = type("COCOeval", (), {"__init__": lambda self, *a, **k: None})
    pkg.coco, pkg.cocoeval = coco, cocoeval
    sys.modules.update({"pycocotools": pkg, "pycocotools.coco": coco, "pycocotools.cocoeval": cocoeval})


def load_reference():
    """Returns a namespace with sbp_utils, spm_utils, sbp_pis_utils, SBPLoss, SPMLoss -- or None."""
    root = reference_root()
    if root is None:
        return None
    sys.dont_write_bytecode = True          # the reference tree is read-only
    _stub_pycocotools()
    # the reference's top-level package names are generic (`utils`, `models`); import them under a
    # temporary sys.path entry and make sure nothing of ours shadows them
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "utils" or k.startswith("utils.")
             or k == "models" or k.startswith("models.")}
    sys.path.insert(0, root)
    try:
        ns = types.SimpleNamespace()
        ns.sbp_utils = importlib.import_module("utils.sbp_utils")
        ns.spm_utils = importlib.import_module("utils.spm_utils")
        ns.sbp_pis_utils = importlib.import_module("utils.sbp_pis_utils")
        ns.SBPLoss = importlib.import_module("models.loss.sbp_loss").SBPLoss
        ns.SPMLoss = importlib.import_module("models.loss.spm_loss").SPMLoss
        ns.root = root
    finally:
        sys.path.remove(root)
        for k in list(sys.modules):
            if k == "utils" or k.startswith("utils.") or k == "models" or k.startswith("models."):
                sys.modules.pop(k)
        sys.modules.update(saved)
    return ns
