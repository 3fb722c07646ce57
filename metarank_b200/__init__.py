"""metarank_b200 — B200-native engine for Metarank's /rank hot path.

The product is libmrgpu.so (csrc/, C ABI in include/mr_b200.h); this package is the
thin Python mirror of the reference's plugin surfaces used by tests and bench.py.
"""
from ._capi import MrError, build  # noqa: F401
from .booster import B200Booster, Context, LightGBMBooster, XGBoostBooster  # noqa: F401
