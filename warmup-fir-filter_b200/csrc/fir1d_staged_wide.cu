// Instantiates the wide (signed int16, MAC == 2) datapath of the staged 1-D
// kernel in its own translation unit.
#define FIRGPU_STAGED_WIDE_TU 1
#include "fir1d_staged.cuh"

namespace firgpu {
namespace staged {

int launch_staged_wide(const StagedParams &p, int n_taps, int dilation, bool clamp, cudaStream_t s)
{
    if (dilation == 1) return launch_taps<int16_t, int16_t, 1>(p, n_taps, clamp ? 1 : 0, 2, s);
    if (dilation == 2) return launch_taps<int16_t, int16_t, 2>(p, n_taps, clamp ? 1 : 0, 2, s);
    return 0;
}

}  // namespace staged
}  // namespace firgpu
