"""maven_b200: B200-native (sm_100a) CLIP contrastive training step of multimodal-supernovae.

Drop-in mirrors of the reference's operator surface for that path only:
  maven_b200.transformer_utils  <->  src/transformer_utils.py
  maven_b200.models_multimodal  <->  src/models_multimodal.py (ConvMixer, LightCurveImageCLIP, MLP)
  maven_b200.loss               <->  src/loss.py (clip_loss, clip_loss_multimodal)
All arithmetic runs in hand-written CUDA kernels reached through the C ABI in include/maven_sm100.h
(libmaven_sm100.so).  There is no CPU fallback: CPU tensors raise.
"""
__version__ = "0.1.0"
