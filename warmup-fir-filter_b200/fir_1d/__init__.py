"""GPU-backed stand-in for the reference's ``fir_1d`` package: same import
paths and call signatures, the arithmetic runs in libfirgpu (sm_100a)."""
