// Instantiates the tiled 2-D kernels for ONE container type, or one input / output pair
// (-DFIRGPU_TILED_T=... [-DFIRGPU_TILED_OUT=...] -DFIRGPU_TILED_NAME=...).
#include "fir2d_tiled.cuh"

#ifndef FIRGPU_TILED_OUT
#define FIRGPU_TILED_OUT FIRGPU_TILED_T
#endif

namespace firgpu {
namespace tiled {

int FIRGPU_TILED_NAME(const CUtensorMap &map, const TiledParams &p, int kh, int kw, int post, bool sep, cudaStream_t s)
{
    return launch_tiled_shape<FIRGPU_TILED_T, FIRGPU_TILED_OUT>(map, p, kh, kw, post, sep, s);
}

}  // namespace tiled
}  // namespace firgpu
