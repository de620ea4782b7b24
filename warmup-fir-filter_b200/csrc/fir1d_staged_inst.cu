// Instantiates the staged 1-D kernels for ONE (input, output) container pair,
// selected by -DFIRGPU_STAGED_IN=... -DFIRGPU_STAGED_OUT=... -DFIRGPU_STAGED_NAME=...
#include "fir1d_staged.cuh"

namespace firgpu {
namespace staged {

int FIRGPU_STAGED_NAME(const StagedParams &p, int n_taps, int dilation, bool hi16, int mac, cudaStream_t s)
{
    return launch_dil<FIRGPU_STAGED_IN, FIRGPU_STAGED_OUT>(p, n_taps, dilation, hi16, mac, s);
}

}  // namespace staged
}  // namespace firgpu
