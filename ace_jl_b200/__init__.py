"""ace_jl_b200 -- B200-native drop-in for the basis/model evaluation hot path of ACE.jl.

Host side (this package): a Python mirror of the reference's construction and evaluation interface.
Device side: ``csrc/libaceb200.so`` (hand-written sm_100a CUDA behind the C ABI of include/aceb200.h).
Importing the package does not need a GPU; evaluating anything does.
"""
from .transforms import (Lambda, agnesitransform, idtransform, morsetransform, parse_exstr,  # noqa: F401
                         polytransform)
from .orthpolys import OrthPolyBasis, discrete_jacobi  # noqa: F401
from .onepbasis import Categorical1pBasis, Product1pBasis, Rn1pBasis, Ylm1pBasis  # noqa: F401
from .selectors import (CategorySparseBasis, MaxBasis, NoConstant, SimpleSparseBasis, SparseBasis,  # noqa: F401
                        gensparse, init1pspec)
from .properties import EuclideanMatrix, EuclideanVector, Invariant, SymmetricEuclideanMatrix  # noqa: F401
from .symmetrygroups import NoSym, O3  # noqa: F401
from .pibasis import PIBasis, PIBasisSpec  # noqa: F401
from .symmbasis import SparseCSC, SymmetricBasis  # noqa: F401
from .api import (ACEConfig, B200Batch, B200Evaluator, LinearACEModel, adjoint_EVAL_D, evaluate, evaluate_d,  # noqa: F401
                  evaluate_ed, grad_config, grad_params, grad_params_config, rrule_evaluate, set_params)
from .structure import B200Structure, neighbourlist, pack_neighbours, reverse_pairs  # noqa: F401
from . import utils  # noqa: F401
from . import fio  # noqa: F401
from .fio import load_model, read_dict, save_model, write_dict  # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
