// Instantiates the long-filter buckets of the staged 1-D kernel for ONE container
// or one input / output pair (-DFIRGPU_LONG_T=... [-DFIRGPU_LONG_OUT=...] -DFIRGPU_LONG_NAME=...).
#include "fir1d_staged.cuh"

#ifndef FIRGPU_LONG_OUT
#define FIRGPU_LONG_OUT FIRGPU_LONG_T
#endif

namespace firgpu {
namespace staged {

int FIRGPU_LONG_NAME(const StagedParams &p, int bucket, int post, cudaStream_t s)
{
    return launch_long<FIRGPU_LONG_T, FIRGPU_LONG_OUT>(p, bucket, post, s);
}

}  // namespace staged
}  // namespace firgpu
