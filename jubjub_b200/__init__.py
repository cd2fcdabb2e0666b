"""jubjub_b200 -- B200-native batched Jubjub curve engine (host-side mirror of the reference's
jubjub::{Fq, Fr, AffinePoint, ExtendedPoint, ...} surface over the C ABI in include/jubjub_b200.h)."""
from ._lib import (JJ_ASYNC, JJ_CANON, JJ_CHECK_SUBGROUP, JJ_CONST_TIME, JJ_DEVICE_PTRS, JJ_OUT_AFFINE, JJ_OUT_BYTES, JJ_PRE_ZIP216,  # noqa: F401
                   JJ_SCALAR_MONT, JJ_SUBTRACT, JJ_TORSION_LADDER, LIB_PATH)
from .engine import FQ, FR, DeviceArray, Engine, JubjubError, default_engine  # noqa: F401
from .sharding import equal_shards, gather_offsets, shard_range  # noqa: F401

__all__ = ["Engine", "DeviceArray", "JubjubError", "default_engine", "FQ", "FR", "LIB_PATH"]

from . import types  # noqa: E402,F401  (reference-style batch types: jubjub_b200.types.Fq, ExtendedPoint, ...)
