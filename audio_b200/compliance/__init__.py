"""Drop-in ``torchaudio.compliance`` surface backed by libb200audio.so (Kaldi-compatible features)."""
from . import kaldi  # noqa: F401

__all__ = ["kaldi"]
