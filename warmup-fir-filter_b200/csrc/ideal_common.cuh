// Helpers shared by the register-only kernels of the floating-point model
// (fir1d_ideal_fast.cu, fir2d_ideal_fast.cu).
#pragma once
#include "common.cuh"

namespace firgpu {
namespace {

// exact integer -> floating point for 0 <= v <= 255 without a conversion instruction
template <typename F> __device__ __forceinline__ F pixel_to(int v);
template <> __device__ __forceinline__ double pixel_to<double>(int v)
{
    return __dadd_rn(__hiloint2double(0x43300000, v), -4503599627370496.0);      // (2^52 + v) - 2^52
}
template <> __device__ __forceinline__ float pixel_to<float>(int v)
{
    return __fadd_rn(__uint_as_float(0x4b000000u | (uint32_t)v), -8388608.0f);   // (2^23 + v) - 2^23
}

template <typename InT> struct PairOf;
template <> struct PairOf<uint8_t> { using type = uint16_t; };
template <> struct PairOf<int16_t> { using type = uint32_t; };
template <> struct PairOf<double> { using type = double2; };

}  // namespace
}  // namespace firgpu
