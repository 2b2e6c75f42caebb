"""vampire_b200 -- B200-native (sm_100a) 2D->3D feature path of Vampire: lift + pool + render."""
from .config import PathConfig, R50_256x704, R50_512x1408, MINI, NAMED  # noqa: F401

__version__ = "0.1.0"
