"""Module boundary of the reference's kandinsky2/model package for the denoising hot path."""
from .unet import Text2ImUNet, InpaintText2ImUNet  # noqa: F401
from .model_creation import create_model, create_gaussian_diffusion  # noqa: F401
