// Fast-path step kernels for n_dim in {2, 3} (see tb_mcmc_fast.cuh).
#include "tb_mcmc_fast.cuh"

namespace tb {
template int launch_fast<2>(const StepArgs& a, cudaStream_t st);
template int launch_fast<3>(const StepArgs& a, cudaStream_t st);
}  // namespace tb
