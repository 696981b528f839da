"""
evidence_b200 — B200-native (sm_100a) batched Keplerian radial-velocity log-likelihood and
unit-cube prior transform for the `evidence` package's nested-sampling runners.

Only the likelihood hot path is rebuilt: ``rvmodel`` (device model),
``priors`` (batched on-device transform), the vectorised UltraNest runner, a batch-of-1 adapter
for PolyChord, and the ctypes layer over the C-ABI of include/rvlnl.h.  There is no CPU
fallback: compute entry points raise if librvlnl.so or an sm_100 device is missing.
"""
__version__ = "0.1.0"
