"""B200-native engine for the perturbation-explainability hot path of
Michal2711/Audio-Deepfake-Explainability: spectrogram occlusion sweep + frequency-band perturbation
pushed through the SpecTTTra-alpha classifier and reduced to a time-frequency importance map.

Host Python (this package) mirrors the reference's entry points; all device work goes through the
C-ABI shared library built from ``csrc/`` (``include/b200xai.h``).  There is no CPU fallback: anything
that computes raises ``RuntimeError`` if ``libb200xai.so`` is missing.
"""
__version__ = "0.1.0"
