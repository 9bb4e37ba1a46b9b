"""B200-native KD-loss hot path."""
