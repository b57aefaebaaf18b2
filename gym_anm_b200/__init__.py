"""B200-native batched gym-anm step() path."""
