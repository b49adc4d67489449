"""reconstructor_b200 -- B200-native exhaustive pair matching + epipolar filtering.

Drop-in for one hot path of smileyenot983/reconstructor (SequentialReconstructor::matchFeatures
and the two plugins it calls).  Layout:

  csrc/    hand-written sm_100a CUDA kernels + the C ABI (include/pairmatch_b200.h)
  cpp/     C++ host-side mirror of the reference's plugin classes over the C ABI
  api.py   ctypes binding (tests, bench)
  shard.py multi-GPU partition of the pair list
  synth.py synthetic image sets
"""
__all__ = ["api", "shard", "synth"]
