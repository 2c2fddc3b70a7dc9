"""thinktwice_b200 — B200-native implementation of the ThinkTwice per-frame forward path."""
__version__ = '0.1.0'
