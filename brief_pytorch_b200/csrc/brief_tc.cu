// placeholder, replaced below
#include "brief_kernels.h"
namespace brief {
bool tc_supported(int, int, int, int) { return false; }
int tc_fpad(int f) { return ((f + 1 + 15) / 16) * 16; }
size_t tc_wpack_bytes(int, int) { return 0; }
size_t tc_eval_smem(int, int) { return 0; }
size_t tc_fit_smem(int, int) { return 0; }
cudaError_t launch_tc_eval(const EvalArgs&, int, int, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_tc_fit(const FitArgs&, int, int, cudaStream_t) { return cudaErrorNotSupported; }
}
