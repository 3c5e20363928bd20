"""unidet3d_b200: Blackwell-native (sm_100a) implementation of UniDet3D's forward hot path."""
__version__ = "0.1.0"
