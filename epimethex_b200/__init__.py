"""epimethex_b200 — B200-native engine for EpiMethEx's per-gene methylation-vs-expression association
path (``epimethex.analysis`` -> ``CG_data_individually.csv``).  The statistics run in libepx.so
(hand-written sm_100a CUDA behind the C ABI of include/epx.h); this package is the host-side mirror of
the reference's R interface.  There is no CPU fallback."""
from .analysis import OUTPUT_COLUMNS, epimethex_analysis, prepare  # noqa: F401
from ._native import EpxError, Session, make_options  # noqa: F401

__all__ = ["epimethex_analysis", "prepare", "Session", "make_options", "EpxError", "OUTPUT_COLUMNS"]
