"""fugu_b200 — B200-native (sm_100a) query hot path for fugu: posting decode -> AND/OR -> BM25 -> top-k.

The product is the C-ABI library libfugu_gpu.so (include/fugu_gpu.h); this package is the thin
Python host mirror used by the tests and the benchmark. There is no CPU fallback.
"""
__all__ = ["_native", "synth"]
