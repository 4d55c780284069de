"""`OriginUniversal` under the reference's module path (codes/models/modules/origin_universal.py)."""
from .universal import OriginUniversal  # noqa: F401
