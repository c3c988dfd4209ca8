// placeholder replaced below in this round: tcgen05 3xTF32 kernel
#include "common.cuh"
namespace b200 {
bool tf32x3_supported(const GemmProblem&) { return false; }
size_t tf32x3_workspace_bytes(const GemmProblem&) { return 0; }
int launch_tf32x3(const GemmProblem&, cudaStream_t, void*, size_t) { return (int)cudaErrorNotSupported; }
double tf32_pipe_peak(int) { return -1.0; }
}
