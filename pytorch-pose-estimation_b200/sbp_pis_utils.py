"""PIS variant of the SBP metric (utils/sbp_pis_utils.py:10-47): same decode + back-projection kernels with K=11,
18 zeros appended to every keypoints row (:40).  The HandleGrip / FallingDown heuristics and drawing are
application logic outside the hot path and are not part of this package."""
from .sbp_utils import SBPmAPCOCO


class SBPmAPPIS(SBPmAPCOCO):
    _pad = 18

    def __init__(self, json_path, input_size, conf_threshold):
        super().__init__(json_path, input_size, conf_threshold)
