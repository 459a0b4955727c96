"""Activation factory mirroring ``src/utils/activation.py:9-35`` of the reference.

The returned ``nn.Module`` instances exist for API parity (``model.activation`` etc.); on the hot path the
activation is fused into the GEMM epilogues of libax2d (all five kinds, see ``csrc/common.cuh``).
"""
import torch.nn as nn

_FACTORY = {"relu": nn.ReLU, "leakyrelu": nn.LeakyReLU, "elu": nn.ELU, "gelu": nn.GELU, "silu": nn.SiLU}


def get_activation_function(activation_type: str) -> nn.Module:
    if activation_type not in _FACTORY:
        supported = ", ".join(_FACTORY.keys())
        raise ValueError(f"Invalid activation type: {activation_type}. Supported: {supported}")
    module = _FACTORY[activation_type]()
    module.ax2d_name = activation_type
    return module


def get_activation_by_name(name: str) -> nn.Module:
    return get_activation_function(name)
