// Fast-path step kernels for n_dim in {10} (see tb_mcmc_fast.cuh).
#include "tb_mcmc_fast.cuh"

namespace tb {
template int launch_fast<10>(const StepArgs& a, cudaStream_t st);
}  // namespace tb
