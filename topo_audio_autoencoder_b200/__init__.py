"""B200-native simplicial-complex stage behind the reference's Python signatures.

Importing this package loads libtopo_b200.so (sm_100a CUDA kernels behind a C ABI,
include/topo_b200.h).  If the library is missing the import fails: there is no fallback path.
"""
from ._lib import lib, LIB_PATH, TopoError            # noqa: F401
from .rectifier import (ConstraintMatrices, SimplexIndices, RectifiedProbs,          # noqa: F401
                        enforce_constraints, rectify_batch)
from .complex_builder import SparseSimplicialMatrices, build_sparse_matrices        # noqa: F401
from .gate import HardConcrete, BinaryGumbel, hard_concrete                         # noqa: F401
from .custom_sccn import GradientSCCN, GradientSCCNLayer, BatchedComplex, Conv      # noqa: F401
from .encoder_complex import ComplexHead, ComplexStage, active_sets                 # noqa: F401
from .decoder import AudioDecoder, DecoderTail                                      # noqa: F401
from .frontend import ConvFrontEnd                                                  # noqa: F401
from .audio2complex import AudioAutoencoder, AudioEncoder                           # noqa: F401
from .loss import AutoencoderLoss, spectral_distance                                # noqa: F401
from .trainer import Trainer                                                        # noqa: F401

__all__ = [
    "ConstraintMatrices", "SimplexIndices", "RectifiedProbs", "enforce_constraints", "rectify_batch",
    "SparseSimplicialMatrices", "build_sparse_matrices", "HardConcrete", "BinaryGumbel", "hard_concrete",
    "GradientSCCN", "GradientSCCNLayer", "BatchedComplex", "Conv", "ComplexHead", "ComplexStage", "active_sets",
    "AudioDecoder", "DecoderTail", "ConvFrontEnd", "AudioAutoencoder", "AudioEncoder", "AutoencoderLoss",
    "spectral_distance", "Trainer",
]
