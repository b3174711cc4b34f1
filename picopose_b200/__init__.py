"""picopose_b200: B200-native correspondence hot path of PicoPose."""
__version__ = "0.1.0"
