"""B200-native Conformer-encoder forward path for mpc001/auto_avsr (hot path only).

    import auto_avsr_b200; auto_avsr_b200.install()      # before importing the reference's E2E
    from auto_avsr_b200 import ConformerEncoder           # or use the class directly

Importing the compute modules requires the built CUDA library (``__graft_entry__.build()``); there is no
CPU or PyTorch fallback.  ``auto_avsr_b200.synthetic`` (weights / inputs generators) imports without it.
"""
__version__ = "0.1.0"

_LAZY = {"ConformerEncoder", "Encoder", "EncoderLayer", "ConvolutionModule", "RelPositionMultiHeadedAttention",
         "PositionwiseFeedForward", "LayerNorm", "RelPositionalEncoding"}


def install():
    from .install import install as _install
    _install()


def __getattr__(name):
    if name in _LAZY:
        from . import espnet_dropin
        return getattr(espnet_dropin, name)
    raise AttributeError(name)
