"""voroffset_b200: B200-native dexel morphology (3D dilation / erosion / opening / closing of a
CompressedVolume, plus the vor2d per-row variant) behind the reference's own operator interface."""
from .volume import CompressedVolume, DexelImage  # noqa: F401
