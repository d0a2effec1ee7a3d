"""photohive_dsp_b200 -- B200-native implementation of PhotoHive_DSP's get_report() hot path.

Drop-in surface (same names as the reference package): ``get_report``, ``set_bounding_boxes``, ``Report``.
Additive: ``get_reports`` / ``Context`` for batches of 8-bit images on one GPU, ``shard`` helpers for N GPUs.
"""
from .core import Report, get_report, get_reports, set_bounding_boxes  # noqa: F401
from .batch import BatchReports, Context, PhotoHiveError, make_params  # noqa: F401
