"""bgge_b200: B200-native getK() / BGGE() (hot paths of italo-granato/BGGE) behind the reference API."""
from .getk import getK, materialize          # noqa: F401
from .bgge import BGGE, BGGE_batch, BGGEFit, setDEC   # noqa: F401
from . import _lib              # noqa: F401

__all__ = ["getK", "materialize", "BGGE", "BGGE_batch", "BGGEFit", "setDEC"]
