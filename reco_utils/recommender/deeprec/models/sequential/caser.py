"""Import-only placeholder: CaserModel is one of the comparison baselines of the reference
(reco_utils/recommender/deeprec/models/sequential/caser.py); it is outside the CLSR hot path this
repository accelerates, but examples/00_quick_start/sequential.py imports it unconditionally."""
from reco_utils.recommender.deeprec.models.sequential.sequential_base_model import SequentialBaseModel

__all__ = ["CaserModel"]


class CaserModel(SequentialBaseModel):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("CaserModel is not part of the B200 CLSR build; use --model CLSR")
