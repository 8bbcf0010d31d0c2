"""guided_attention_b200 -- B200-native cross-attention guidance path behind the Guided-Attention hook API.

Importing the package is cheap and GPU-free; the CUDA library (`csrc/libguidedattn.so`, built by
`__graft_entry__.build()`) is loaded on first use by `guided_attention_b200._cabi` and its absence is a hard error:
there is no CPU or PyTorch fallback for the hot path.
"""
__version__ = "0.1.0"
