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
