// placeholder, replaced below
#include "yc_common.cuh"
namespace yc {
int launch_head_tcgen05(const yc_head_desc *d, int rows_total, const int *row_off, cudaStream_t stream)
{
    set_error("tcgen05 head path not built yet");
    return YC_ERR_UNSUPPORTED;
}
}
