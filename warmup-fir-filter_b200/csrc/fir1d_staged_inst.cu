// Instantiates the staged 1-D kernels for ONE (input, output, dilation) combination,
// selected by -DFIRGPU_STAGED_IN=... -DFIRGPU_STAGED_OUT=... -DFIRGPU_STAGED_DIL=... -DFIRGPU_STAGED_NAME=...
#include "fir1d_staged.cuh"

namespace firgpu {
namespace staged {

int FIRGPU_STAGED_NAME(const StagedParams &p, int n_taps, int post, int mac, cudaStream_t s)
{
    return launch_taps<FIRGPU_STAGED_IN, FIRGPU_STAGED_OUT, FIRGPU_STAGED_DIL>(p, n_taps, post, mac, s);
}

}  // namespace staged
}  // namespace firgpu
