"""B200-native Conformer-encoder forward path for mpc001/auto_avsr (hot path only)."""
__version__ = "0.1.0"
