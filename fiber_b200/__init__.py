"""fiber_b200 — B200-native (sm_100a) implementation of FIBER's fusion-in-the-backbone hot path."""
__version__ = "0.1.0"
