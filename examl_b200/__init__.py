"""examl_b200 -- B200-native implementation of ExaML's per-site likelihood hot path
(newview / evaluate / makenewz) behind the reference's own entry points.

  examl_b200/csrc/      CUDA kernels + C ABI (include/examl_b200.h) -> libexaml_b200.so,
                        and the C drop-in for the reference's host code (axml_backend.c)
  examl_b200.capi       ctypes binding of the C ABI
  examl_b200.likelihood evaluateGeneric / newviewGeneric / makenewzGeneric host mirror
  examl_b200.phylo      node rings, traversal descriptors, Newick
  examl_b200.model      model constants (EIGN/EV/EI/tipVector/gammaRates)
  examl_b200.data       synthetic alignments, ExaML binary alignment files
"""
from .phylo import Tree, Entry, TIP_TIP, TIP_INNER, INNER_INNER  # noqa: F401

__version__ = "0.1.0"
