"""splitvae_b200: B200-native (sm_100a) train hot path of SPLIT-VAE / SPLIT-GMVAE (51616/split-vae).

Python is host plumbing only; the arithmetic lives in libsplitvae.so (include/splitvae.h).
Importing the package does not load the library; the first use of the product path does, and fails
loudly if it has not been built or if no sm_100 GPU is present (there is no CPU fallback)."""

__all__ = ["LGVae", "LGGMVae", "GMVae", "Engine", "Augmentator"]


def __getattr__(name):
    if name in ("LGVae", "LGGMVae", "GMVae"):
        from . import model
        return getattr(model, name)
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name == "Augmentator":
        from .augmentation import Augmentator
        return Augmentator
    raise AttributeError(name)
