"""pose_b200 -- B200-native (sm_100a) heatmap hot path for Simple Baselines and SPM pose estimation.

Importable as `pose_b200` (the repo-root shim) -- the directory name carries a hyphen.  Mirrors the reference's
hot-path modules:

    utils/sbp_utils.py      -> pose_b200.sbp_utils      (SBPHeatmapGenerator, nms_sbp, DecodeSBP, SBPmAPCOCO)
    utils/sbp_pis_utils.py  -> pose_b200.sbp_pis_utils  (SBPmAPPIS)
    utils/spm_utils.py      -> pose_b200.spm_utils      (SPM*Generator, nms_spm, DecodeSPM, SPMmAPCOCO)
    models/loss/sbp_loss.py -> pose_b200.sbp_loss       (SBPLoss)
    models/detector/sbp.py:35-37 (the 1x1 head) + SBPLoss / DecodeSBP -> pose_b200.sbp_head (sbp_head_fused, HeadFusedSBPLoss)
    models/loss/spm_loss.py -> pose_b200.spm_loss       (SPMLoss)
    pycocotools COCOeval    -> pose_b200.coco_eval      (KeypointEval: OKS matching + AP behind the metric classes' result())

All arithmetic runs in hand-written CUDA kernels behind the C ABI in include/pose_b200.h
(libpose_b200.so, loaded with ctypes).  There is no CPU, PyTorch-op or Triton fallback: without the
built library or without a CUDA device every entry point raises.
"""
from ._cabi import LIB_PATH, PoseB200Error, launch_count, lib  # noqa: F401
from .coco_eval import CocoKeypointsGT, KeypointEval  # noqa: F401
from .sbp_head import HeadFusedSBPLoss, head_tuning, sbp_head_fused  # noqa: F401
from .sbp_loss import SBPLoss, sbp_fused  # noqa: F401
from .sbp_pis_utils import SBPmAPPIS  # noqa: F401
from .sbp_utils import (DecodeSBP, SBPHeatmapGenerator, SBPmAPCOCO, backproject_packed, backproject_rows, decode_batch,  # noqa: F401
                        nms_sbp, packed_to_results)
from .spm_loss import SPMLoss, spm_fused, spm_loss_fused  # noqa: F401
from .spm_utils import (DecodeSPM, SPMDisplacementGenerator, SPMHeatmapGenerator, SPMMaskGenerator, SPMmAPCOCO,  # noqa: F401
                        get_spm_keypoints, get_spm_keypoints_chained, nms_spm, spm_decode_batch, spm_render_batch)

__version__ = "0.1.0"
