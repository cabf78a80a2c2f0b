"""Deterministic synthetic weights keyed by parameter name (no checkpoints exist offline).

Every tensor is drawn from its own CPU generator seeded by (seed, crc32(key)), so the same state dict is
produced in the build container (where the reference runs and the golden vectors are written) and on the
GPU box.  zero_module()'d tensors of the reference (nn.py:73-79: ResBlock out conv, attention proj_out,
the output conv) are filled like any other weight -- otherwise the network outputs exact zeros
(SURVEY.md headline fact 4) and parity would be vacuous.
"""
import zlib

import torch


def synth_tensor(key, shape, seed=0):
    g = torch.Generator(device="cpu").manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "bias":
        return 0.05 * torch.randn(shape, generator=g)
    if len(shape) == 1:  # norm gains
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if leaf == "embedding" or key.endswith("embedding.weight"):  # VQ codebook
        return torch.randn(shape, generator=g)
    fan_in = 1
    for d in shape[1:]:
        fan_in *= d
    return torch.randn(shape, generator=g) / fan_in ** 0.5


def synth_state_dict(spec, seed=0, dtype=torch.float32):
    """spec: iterable of (key, shape) -> {key: tensor}."""
    return {k: synth_tensor(k, s, seed).to(dtype) for k, s in spec}
