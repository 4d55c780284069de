"""`IspUniversal` under the reference's module path (codes/models/modules/isp_universal.py)."""
from .universal import IspUniversal  # noqa: F401
