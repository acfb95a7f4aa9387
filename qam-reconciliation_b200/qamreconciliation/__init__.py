"""qamreconciliation -- B200 (sm_100a) implementation of the reverse-reconciliation hot path of
moriglia/qam-reconciliation behind the reference's class API (reference __init__.py:1-4).

Host code is Python + torch tensors over the C ABI of libqamrecon.so (include/qamrecon.h).
There is no CPU fallback: constructing any compute class without a CUDA device raises.
"""
from .decoder import Decoder
from .matrix import Matrix
from .noisemapper import NoiseMapper, NoiseDemapper, NoiseMapperFlipSign, NoiseMapperAntiFlipSign
from .alphabet import PAMAlphabet

__all__ = ["Decoder", "Matrix", "NoiseMapper", "NoiseDemapper", "NoiseMapperFlipSign", "NoiseMapperAntiFlipSign",
           "PAMAlphabet"]
