"""B200-native stereo disparity engine behind mvStereoVision3's disparity.h / Stereosystem path.

The product is libmvsv.so (hand-written sm_100a CUDA behind the C ABI in include/mvsv.h);
this package is the thin Python binding used by tests/ and bench.py.  There is no CPU fallback:
importing `mvstereovision3_b200.api` without the built library raises.
"""
__version__ = "0.1.0"
