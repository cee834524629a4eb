// Fast-path step kernels for n_dim in {6, 8} (see tb_mcmc_fast.cuh).
#include "tb_mcmc_fast.cuh"

namespace tb {
template int launch_fast<6>(const StepArgs& a, cudaStream_t st);
template int launch_fast<8>(const StepArgs& a, cudaStream_t st);
}  // namespace tb
