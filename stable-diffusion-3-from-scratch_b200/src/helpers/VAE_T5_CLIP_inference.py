"""Synthetic stand-in for the reference's frozen encoders (reference:
src/helpers/VAE_T5_CLIP_inference.py:151-165 text_to_embedding; diff_model.py:377,381,467,477
are the attributes sample_imgs touches).

The FLUX VAE and the Gemma / ModernBERT / CLIP text encoders are out of scope
(BASELINE.json north_star): they need Hugging Face weights and a network.  This class is
duck-type compatible with what diff_model.sample_imgs uses and produces deterministic
synthetic embeddings of the right shapes, with an identity "VAE" (latents in, latents out).
"""
import hashlib
from types import SimpleNamespace

import torch


class _IdentityVAE:
    def __init__(self, device, latent_channels=16):
        self.config = SimpleNamespace(latent_channels=latent_channels, shift_factor=0.0, scaling_factor=1.0)
        self.dtype = torch.float32
        self.device = device

    def decode(self, z):
        return SimpleNamespace(sample=z)


class VAE_T5_CLIP_inference:
    def __init__(self, device, latent_channels=16, text_tokens=154, text_dim=2304, pooled_dim=768):
        self.device = device
        self.VAE = _IdentityVAE(device, latent_channels)
        self.text_tokens, self.text_dim, self.pooled_dim = text_tokens, text_dim, pooled_dim

    @torch.no_grad()
    def text_to_embedding(self, text):
        """(hidden [1,154,2304], pooled [1,768]) fp16 like the reference, seeded by the prompt."""
        if isinstance(text, (list, tuple)):
            text = " ".join(map(str, text))
        seed = int.from_bytes(hashlib.sha256(str(text).encode()).digest()[:4], "little")
        g = torch.Generator().manual_seed(seed)
        hidden = torch.randn((1, self.text_tokens, self.text_dim), generator=g)
        pooled = torch.randn((1, self.pooled_dim), generator=g)
        return hidden.to(torch.float16), pooled.to(torch.float16)
