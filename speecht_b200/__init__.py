"""speecht_b200 -- B200-native Wav2Letter hot path behind the speechT Python surface.

Modules mirror the reference package (speecht.*): vocabulary, speech_input, speech_model, preprocessing,
evaluation, training, execution.  Everything numerical runs in libspeecht_b200.so (hand-written sm_100a CUDA,
C ABI in include/speecht_b200.h); torch is used for device memory, streams and torch.distributed only.
"""
__version__ = '0.1.0'
