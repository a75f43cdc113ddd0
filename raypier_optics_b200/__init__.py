"""raypier_optics_b200 -- B200-native (sm_100a CUDA) non-sequential ray-tracing core with
the API of ``raypier.core`` for the hot path."""
from . import _abi  # noqa: F401

__all__ = ["core", "scene", "engine", "configs"]
