// Instantiates the wide (signed int16, MAC == 2) datapath of the staged 1-D
// kernel in its own translation unit.
#define FIRGPU_STAGED_WIDE_TU 1
#include "fir1d_staged.cuh"

namespace firgpu {
namespace staged {

int launch_staged_wide(const StagedParams &p, int n_taps, int dilation, bool clamp, cudaStream_t s)
{
    return launch_dil<int16_t, int16_t>(p, n_taps, dilation, clamp, 2, s);
}

}  // namespace staged
}  // namespace firgpu
