// Fast-path step kernels for n_dim in {4, 5} (see tb_mcmc_fast.cuh).
#include "tb_mcmc_fast.cuh"

namespace tb {
template int launch_fast<4>(const StepArgs& a, cudaStream_t st);
template int launch_fast<5>(const StepArgs& a, cudaStream_t st);
}  // namespace tb
