"""qinco_b200 — B200-native QINCo / QINCo2 beam-search residual-quantisation encode/decode.

Importing the package does not load the CUDA library; creating a model does, and fails
loudly if libqinco_b200.so is missing (there is no CPU fallback).
"""
__version__ = "0.1.0"
