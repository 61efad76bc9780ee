"""drloco_b200: B200-native batched DeepMimic walker environment (drop-in for DRLoco's VecEnv path)."""
__version__ = "0.1.0"
