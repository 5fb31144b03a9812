"""Import-only placeholder: DIENModel is one of the comparison baselines of the reference
(reco_utils/recommender/deeprec/models/sequential/dien.py); it is outside the CLSR hot path this
repository accelerates, but examples/00_quick_start/sequential.py imports it unconditionally."""
from reco_utils.recommender.deeprec.models.sequential.sequential_base_model import SequentialBaseModel

__all__ = ["DIENModel"]


class DIENModel(SequentialBaseModel):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("DIENModel is not part of the B200 CLSR build; use --model CLSR")
