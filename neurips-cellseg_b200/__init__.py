"""cellseg_b200 — B200-native inference glue, post-processing and F1 metric path of
JunMa11/NeurIPS-CellSeg (reference: baseline/predict.py:87-122, baseline/compute_metric.py).

Compute lives in libcellseg_sm100.so (hand-written sm_100a CUDA behind the C ABI of
include/cellseg_b200.h); this package is the host-side mirror of the reference's
call sites.  Submodules are imported lazily so that `import cellseg_b200` works on
a box without a GPU (the build check and the CPU test tier)."""
__version__ = "0.1.0"
