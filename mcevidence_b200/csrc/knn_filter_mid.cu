// Instantiations of the K2 filter kernel for ndim 17..32 (generated list; see knn_filter.cuh).
#include "knn_filter.cuh"

namespace mce {

int filter_dispatch_mid(int op, int d, const FilterConfig& cfg, const FilterParams* p, int list_len, cudaStream_t st, int* qpc,
       int* smem, int* ctas) {
    MCE_FILTER_CASE(17, 2, 8, 1)
    MCE_FILTER_CASE(17, 4, 4, 1)
    MCE_FILTER_CASE(18, 2, 8, 1)
    MCE_FILTER_CASE(18, 4, 4, 1)
    MCE_FILTER_CASE(19, 2, 8, 1)
    MCE_FILTER_CASE(19, 4, 4, 1)
    MCE_FILTER_CASE(20, 2, 8, 1)
    MCE_FILTER_CASE(20, 4, 4, 1)
    MCE_FILTER_CASE(20, 1, 8, 1)
    MCE_FILTER_CASE(20, 2, 4, 1)
    MCE_FILTER_CASE(20, 3, 8, 1)
    MCE_FILTER_CASE(20, 2, 8, 0)
    MCE_FILTER_CASE(21, 2, 8, 1)
    MCE_FILTER_CASE(21, 4, 4, 1)
    MCE_FILTER_CASE(22, 2, 8, 1)
    MCE_FILTER_CASE(22, 4, 4, 1)
    MCE_FILTER_CASE(23, 2, 8, 1)
    MCE_FILTER_CASE(23, 4, 4, 1)
    MCE_FILTER_CASE(24, 2, 8, 1)
    MCE_FILTER_CASE(24, 4, 4, 1)
    MCE_FILTER_CASE(25, 2, 8, 1)
    MCE_FILTER_CASE(25, 4, 4, 1)
    MCE_FILTER_CASE(26, 2, 8, 1)
    MCE_FILTER_CASE(26, 4, 4, 1)
    MCE_FILTER_CASE(27, 2, 8, 1)
    MCE_FILTER_CASE(27, 4, 4, 1)
    MCE_FILTER_CASE(27, 1, 8, 1)
    MCE_FILTER_CASE(27, 2, 4, 1)
    MCE_FILTER_CASE(27, 3, 8, 1)
    MCE_FILTER_CASE(27, 2, 8, 0)
    MCE_FILTER_CASE(28, 2, 8, 1)
    MCE_FILTER_CASE(28, 4, 4, 1)
    MCE_FILTER_CASE(29, 2, 8, 1)
    MCE_FILTER_CASE(29, 4, 4, 1)
    MCE_FILTER_CASE(30, 2, 8, 1)
    MCE_FILTER_CASE(30, 4, 4, 1)
    MCE_FILTER_CASE(31, 2, 8, 1)
    MCE_FILTER_CASE(31, 4, 4, 1)
    MCE_FILTER_CASE(32, 2, 8, 1)
    MCE_FILTER_CASE(32, 4, 4, 1)
    return -1;
}

}  // namespace mce
