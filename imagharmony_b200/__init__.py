"""imagharmony_b200: B200-native (sm_100a) SDXL denoise hot path with IMAGHarmony decoupled IP cross-attention."""
__version__ = "0.1.0"
