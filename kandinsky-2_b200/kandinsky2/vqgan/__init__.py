"""MoVQ image decoder of the hot path (reference: kandinsky2/vqgan/autoencoder.py:160-201, movq_modules.py)."""
from .autoencoder import MOVQ  # noqa: F401
