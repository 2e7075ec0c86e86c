"""fastenhancer_b200 -- B200-native engine for FastEnhancer's per-frame streaming hot path."""
from .config import FEConfig, PRESETS, preset  # noqa: F401

__version__ = "0.1.0"
