"""MPDOSimulator - B200-native build of the noisy-gate update path (drop-in API of the reference package)."""
__version__ = "1.0.0+b200"
