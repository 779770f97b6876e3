"""ldmae_b200 -- B200-native (sm_100a) implementation of the LDMAE latent-diffusion hot path.

Mirrors the reference's module API for that path only:
  ldmae_b200.models.lightningdit   <-> LDMAE/models/lightningdit.py   (LightningDiT, LightningDiT_models)
  ldmae_b200.transport             <-> LDMAE/transport/               (create_transport, Transport, Sampler)
  ldmae_b200.tokenizer.models_mae  <-> LDMAE/tokenizer/models_mae.py  (mae_for_ldmae_f8d16_prev, decode, decode_to_images)
All compute runs in libldmae_b200.so (ldmae_b200/csrc, C ABI in include/ldmae_b200.h).
"""
__version__ = "0.1.0"
