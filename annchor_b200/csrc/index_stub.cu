// stubs for entry points that are declared in include/annb.h but not built yet
#include "common.cuh"
using namespace annb;
ANNB_API int annb_bruteforce_knn(annb_ctx *, const annb_dataset *, int, int64_t, int64_t *, double *)
{
    set_error("annb_bruteforce_knn not implemented yet");
    return ANNB_ESTATE;
}
