"""B200-native drop-in for the slice of ``reco_utils`` that the CLSR quick-start uses
(reference: reco_utils/__init__.py; module paths imported by examples/00_quick_start/sequential.py:16-34).
The CLSR model executes in hand-written sm_100a CUDA (package ``clsr_b200``); everything else on
these paths is host logic that keeps the reference's names, arguments and error behaviour."""
__title__ = "clsr-b200 reco_utils"
__version__ = "0.1.0"
