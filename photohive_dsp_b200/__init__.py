"""photohive_dsp_b200 -- B200-native implementation of PhotoHive_DSP's get_report() hot path.

Drop-in surface (same names as the reference package): ``get_report``, ``set_bounding_boxes``, ``Report``.
Additive: ``get_reports`` / ``Context`` for batches of 8-bit images on one GPU, ``MultiContext`` for one host batch over
several GPUs of one process, ``shard`` for one process per GPU.

Attributes resolve lazily so that ``python -m photohive_dsp_b200.build`` can (re)build the shared library
before anything tries to load it.
"""
_CORE = {"Report", "get_report", "get_reports", "set_bounding_boxes"}
_BATCH = {"BatchReports", "Context", "MultiContext", "PhotoHiveError", "make_params"}
__all__ = sorted(_CORE | _BATCH)


def __getattr__(name):
    if name in _CORE:
        from . import core
        return getattr(core, name)
    if name in _BATCH:
        from . import batch
        return getattr(batch, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
