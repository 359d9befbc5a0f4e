"""gamcsampler.jl_b200 -- B200-native hot path of papamarkou/GAMCSampler.jl behind a C ABI.

csrc/      hand-written sm_100a CUDA kernels + the C-ABI library (libgamc_b200.so, built in-tree)
_lib.py    ctypes binding of include/gamc_b200.h (no CPU fallback: import fails loudly without the .so)
engine.py  one handle = one GPU
frontend.py host-side mirror of the reference's sampler / tuner / schedule / job API
stats.py   chain summaries (acceptance, ESS) and the examples' CSV writers
distributed.py  one process per GPU plumbing (torch.distributed) for row-sharded jobs
julia/     the ccall front-end a GAMCSampler.jl maintainer would add (not executable in this image)
"""
from . import _lib
from .engine import Engine
from .frontend import *  # noqa: F401,F403
from .frontend import __all__ as _fe_all
from .stats import ess, mcvar_imse, summary, write_chain_csv, write_run_csvs, table  # noqa: F401

__all__ = ["Engine", "ess", "mcvar_imse", "summary", "write_chain_csv", "write_run_csvs", "table"] + list(_fe_all)
__version__ = "0.1.0"
