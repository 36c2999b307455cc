"""streamformer_b200 — B200-native (sm_100a) implementation of the StreamFormer video-encoder hot path."""
__version__ = "0.1.0"
