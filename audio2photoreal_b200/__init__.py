"""audio2photoreal_b200 -- B200-native drop-in for audio2photoreal's diffusion sampling hot path.

Only what the path needs: csrc/ (sm_100a kernels + the C-ABI), and the host-side mirror of the
reference's FiLMTransformer / ClassifierFreeSampleModel / SpacedDiffusion interface.
"""
__version__ = "0.1.0"
