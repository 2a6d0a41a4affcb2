"""seqwin_b200 -- B200-native (sm_100a) drop-in for the hot path of treangenlab/Seqwin.

Only the data-parallel path ntHash -> window minimizers -> minimizer pan-genome graph lives here:

* ``seqwin_b200._core``  -- ctypes shim with the three entry points of the reference extension
  ``seqwin.graph._core`` (``_build_native``, ``_get_penalty_native``, ``_filter_kmers_native``).
* ``seqwin_b200.graph``  -- mirror of ``seqwin/graph/__init__.py`` (``KmerGraph`` and dtypes).
* ``seqwin_b200.csrc``   -- the CUDA kernels and the C ABI (``include/seqwin_b200.h``).

There is no CPU fallback: every compute entry point raises if ``libseqwin_b200.so`` is missing or
no sm_100 device is usable.
"""
from ._version import __version__  # noqa: F401
