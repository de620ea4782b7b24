"""firgpu — B200-native (sm_100a) fixed-point FIR path of warmup-fir-filter.

``firgpu.ops`` holds the array-level calls, ``firgpu.quant`` the host-side
contract (validation texts, float64 tap quantisation), ``firgpu.shard`` the
partition planner for 1/2/4/8 GPUs, ``firgpu._lib`` the ctypes binding of
``libfirgpu.so`` (C ABI: ``include/firgpu.h``).  The reference's own import
paths (``fir_1d.model.python...``, ``fir_2d...``, ``pipeline_fir_1d``) live next to this
package and call into it.  No CPU fallback exists.
"""
from .ops import (DeviceArray, FirStream, PinnedArray, Recording, compare_metrics, fir1d_fixed, fir1d_fixed_complex, fir1d_fixed_iq,  # noqa: F401
                  fir1d_fixed_multi, fir1d_fixed_multi_gpu, fir1d_fixed_plan, fir1d_fixed_q, fir1d_ideal, fir2d_fixed, fir2d_fixed_q, fir2d_ideal, record)

__all__ = ["DeviceArray", "FirStream", "PinnedArray", "Recording", "compare_metrics", "fir1d_fixed", "fir1d_fixed_q", "fir1d_fixed_iq", "fir1d_fixed_complex",
           "fir1d_fixed_multi", "fir1d_fixed_multi_gpu", "fir1d_fixed_plan", "fir1d_ideal", "fir2d_fixed_q", "fir2d_fixed", "fir2d_ideal", "record"]
