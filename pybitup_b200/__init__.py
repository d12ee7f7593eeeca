"""pybitup_b200: B200-native batched-chain MCMC engine behind jcoheur/pybitup's Python surface.

The sub-module names are the reference's (``pybitup/__init__.py``), so ``from pybitup_b200 import
bayesian_inference as bi`` and ``pybitup_b200.solve_problem.Sampling(json).sample(models)`` read like the
reference's own scripts.  Importing the package loads the CUDA library (``_capi``): it fails loudly if the library
has not been built -- there is no CPU fallback.
"""
from . import _capi                                    # noqa: F401  (raises ImportError if the CUDA library is missing)
from . import distributions, bayesian_inference, metropolis_hastings_algorithms, inference_problem, solve_problem  # noqa: F401
from . import diagnostics, distributed, propagation, synthetic                                                     # noqa: F401
from .engine import ChainEngine, ProblemSpec, SamplerConfig                                                        # noqa: F401

__all__ = ["distributions", "bayesian_inference", "metropolis_hastings_algorithms", "inference_problem",
           "solve_problem", "diagnostics", "distributed", "propagation", "synthetic",
           "ChainEngine", "ProblemSpec", "SamplerConfig"]
