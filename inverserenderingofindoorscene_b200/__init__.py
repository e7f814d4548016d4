"""B200-native render layer + bilateral solver (hot path of lzqsd/InverseRenderingOfIndoorScene).

Host side: Python/PyTorch mirrors of the reference's own classes (same names and signatures):
  models.renderingLayer, models.output2env, models.LSregress, models.LSregressDiffSpec,
  BilateralLayer.BilateralFunction, BilateralLayer.BilateralLayer.
Device side: hand-written sm_100a CUDA kernels behind the C ABI of include/irb.h
(libirb_b200.so), called through ctypes.  No Triton, no CPU fallback.
"""
__version__ = '0.1.0'
