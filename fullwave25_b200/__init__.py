"""fullwave25_b200 -- B200-native (sm_100a) time-stepping engine for Fullwave 2.5.

Drop-in for the pre-compiled CUDA executable that ``fullwave.Solver.run`` launches
(/root/reference/fullwave/solver/launcher.py:160-254): same inputs (the ``.dat`` protocol of
/root/reference/fullwave/solver/input_file_writer.py), same output (``genout.dat`` frames), computed by
hand-written CUDA kernels behind the C-ABI in ``include/fw25.h``.  No CPU fallback.
"""

from .problem import Problem  # noqa: F401

__version__ = "0.1.0"
