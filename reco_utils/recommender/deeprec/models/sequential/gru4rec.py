"""Import-only placeholder: GRU4RecModel is one of the comparison baselines of the reference
(reco_utils/recommender/deeprec/models/sequential/gru4rec.py); it is outside the CLSR hot path this
repository accelerates, but examples/00_quick_start/sequential.py imports it unconditionally."""
from reco_utils.recommender.deeprec.models.sequential.sequential_base_model import SequentialBaseModel

__all__ = ["GRU4RecModel"]


class GRU4RecModel(SequentialBaseModel):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("GRU4RecModel is not part of the B200 CLSR build; use --model CLSR")
