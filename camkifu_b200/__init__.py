"""camkifu_b200 — B200-native (sm_100a) stone-detection hot path of CamKifu behind the reference's plugin API."""
__version__ = "0.1.0"
