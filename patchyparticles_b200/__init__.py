"""B200-native Monte Carlo engine for the PatchyParticles roto-translation hot path.

The product is the C-ABI library ``libpatchy_b200.so`` (hand-written sm_100a CUDA, see
``csrc/`` and ``include/pp_b200.h``) and the C host in ``host/``.  This package is the Python
mirror of that ABI used by the tests and by ``bench.py``.  There is no CPU fallback: importing
:class:`Engine` without the built library, or creating one without a CUDA device, raises.
"""
from .engine import Engine, Group, Params, Observables, load_library, library_path, LibraryMissing  # noqa: F401
from .rrr import read_rrr, write_rrr, read_rrb, write_rrb, orientation_from_rows  # noqa: F401
from .configs import fcc_configuration, random_fluid  # noqa: F401

__all__ = ["Engine", "Group", "Params", "Observables", "load_library", "library_path", "LibraryMissing",
           "read_rrr", "write_rrr", "read_rrb", "write_rrb", "orientation_from_rows", "fcc_configuration", "random_fluid"]
