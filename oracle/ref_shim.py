"""Import the UNMODIFIED reference modules from /root/reference (build container only).

`import kandinsky2` itself fails on a modern huggingface_hub (kandinsky2/__init__.py:2 cached_download), so
the package is registered as an empty namespace whose __path__ points at the reference tree, and
pytorch_lightning (vqgan/autoencoder.py:3) is stubbed.  Nothing from the reference is copied.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("K2_REFERENCE_ROOT", "/root/reference")
REF_NAME = "kandinsky2"


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "kandinsky2", "model"))


class reference_modules:
    """Context manager: inside it, `kandinsky2.*` resolves to the reference; the previous sys.modules
    entries (e.g. this repo's own kandinsky2 package) are restored on exit."""

    def __enter__(self):
        if not available():
            raise RuntimeError(f"reference tree not found under {REF_ROOT}")
        self._saved = {k: v for k, v in sys.modules.items() if k == REF_NAME or k.startswith(REF_NAME + ".")}
        for k in self._saved:
            del sys.modules[k]
        self._saved_pl = sys.modules.get("pytorch_lightning")
        pkg = types.ModuleType(REF_NAME)
        pkg.__path__ = [os.path.join(REF_ROOT, "kandinsky2")]
        sys.modules[REF_NAME] = pkg
        for sub in ("model", "vqgan"):
            m = types.ModuleType(f"{REF_NAME}.{sub}")
            m.__path__ = [os.path.join(REF_ROOT, "kandinsky2", sub)]
            sys.modules[f"{REF_NAME}.{sub}"] = m
        if "pytorch_lightning" not in sys.modules:
            import torch.nn as nn
            pl = types.ModuleType("pytorch_lightning")
            pl.LightningModule = nn.Module
            sys.modules["pytorch_lightning"] = pl
        return self

    def load(self, name):
        return importlib.import_module(f"{REF_NAME}.{name}")

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k == REF_NAME or k.startswith(REF_NAME + ".")]:
            del sys.modules[k]
        sys.modules.update(self._saved)
        if self._saved_pl is None:
            sys.modules.pop("pytorch_lightning", None)
        return False



class cuda_as_cpu:
    """Context manager for the reference code that hard-codes the device ("cuda" in model/samplers.py:78-79,101,226,344-
    345,369,495): inside it `Tensor.to("cuda")`, `torch.full(..., device="cuda")` and `torch.randn(..., device="cuda")`
    land on the CPU, so the UNMODIFIED sampler classes run in the build container.  Test infrastructure only."""

    @staticmethod
    def _is_cuda(d):
        import torch
        return (isinstance(d, str) and d.startswith("cuda")) or (isinstance(d, torch.device) and d.type == "cuda")

    def __enter__(self):
        import torch
        self._torch = torch
        self._to, self._full, self._randn = torch.Tensor.to, torch.full, torch.randn
        is_cuda, to0, full0, randn0 = self._is_cuda, self._to, self._full, self._randn

        def to(t, *a, **k):
            a = tuple("cpu" if is_cuda(x) else x for x in a)
            if is_cuda(k.get("device")):
                k["device"] = "cpu"
            return to0(t, *a, **k)

        def full(*a, **k):
            if is_cuda(k.get("device")):
                k["device"] = "cpu"
            return full0(*a, **k)

        def randn(*a, **k):
            if is_cuda(k.get("device")):
                k["device"] = "cpu"
            return randn0(*a, **k)

        torch.Tensor.to, torch.full, torch.randn = to, full, randn
        return self

    def __exit__(self, *exc):
        t = self._torch
        t.Tensor.to, t.full, t.randn = self._to, self._full, self._randn
        return False
