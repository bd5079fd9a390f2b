"""cudabrot_b200 -- B200-native Buddhabrot hot path behind a C ABI (include/buddha.h).

The product is `libbuddha.so` (hand-written sm_100a CUDA, csrc/) and the drop-in `bin/cudabrot`
command line.  This package only binds the C ABI for tests and benchmarks; it never computes
anything itself and has no CPU fallback.
"""
from . import capi
from .capi import (F_BURNING_SHIP, F_EXACT_BINNING, F_FORCE_TILED, F_NO_SHORTCUT, F_SIMPLE_KERNEL,
                   BuddhaError)
from .renderer import Renderer, merge_in_process, write_pgm

__all__ = ["capi", "Renderer", "merge_in_process", "write_pgm", "BuddhaError",
           "F_NO_SHORTCUT", "F_SIMPLE_KERNEL", "F_EXACT_BINNING", "F_FORCE_TILED", "F_BURNING_SHIP"]
