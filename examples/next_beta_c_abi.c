/* Plain-C caller of libtempest_b200.so: the reweighting step of one Persistent Sampling iteration
 * (tempest/steps/reweight.py:341-495) without Python or torch -- raw device pointers, sizes and a stream.
 *
 *   nvcc -x c -I include examples/next_beta_c_abi.c -L tempest_b200/lib -ltempest_b200 -lcudart -o next_beta_demo
 *   LD_LIBRARY_PATH=tempest_b200/lib ./next_beta_demo
 *
 * Three stored generations drawn at beta = 0 (warm-up), log-likelihoods -0.5*chi^2-like; finds the next
 * temperature for an ESS target of 2N and normalised weights at it. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "tempest_b200.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "%s failed: %d\n", #x, rc_); return 1; } } while (0)

int main(void) {
  const int64_t n_gen = 1 << 16, T = 3, n = n_gen * T;
  double* h_logl = (double*)malloc(sizeof(double) * n);
  unsigned long long s = 88172645463325252ull;
  for (int64_t i = 0; i < n; ++i) {          /* xorshift uniforms -> sum of squares of 4 pseudo-normals */
    double acc = 0.0;
    for (int k = 0; k < 4; ++k) {
      double u = 0.0;
      for (int j = 0; j < 12; ++j) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; u += (double)(s >> 11) / 9007199254740992.0; }
      acc += (u - 6.0) * (u - 6.0);
    }
    h_logl[i] = -0.5 * acc;
  }
  const double h_beta[3] = {0.0, 0.0, 0.0}, h_logz[3] = {0.0, 0.0, 0.0};
  const double h_logn[3] = {log((double)n_gen), log((double)n_gen), log((double)n_gen)};
  double *logl, *C, *w, *gb, *gz, *gn, *res;
  void *ws;
  cudaStream_t st;
  CHECK(cudaStreamCreate(&st));
  CHECK(cudaMalloc((void**)&logl, sizeof(double) * n));
  CHECK(cudaMalloc((void**)&C, sizeof(double) * n));
  CHECK(cudaMalloc((void**)&w, sizeof(double) * n));
  CHECK(cudaMalloc((void**)&gb, sizeof(h_beta)));
  CHECK(cudaMalloc((void**)&gz, sizeof(h_logz)));
  CHECK(cudaMalloc((void**)&gn, sizeof(h_logn)));
  CHECK(cudaMalloc((void**)&res, sizeof(double) * 16));
  CHECK(cudaMalloc(&ws, tb_next_beta_workspace_bytes()));
  CHECK(cudaMemset(ws, 0, tb_next_beta_workspace_bytes()));
  CHECK(cudaMemcpy(logl, h_logl, sizeof(double) * n, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(gb, h_beta, sizeof(h_beta), cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(gz, h_logz, sizeof(h_logz), cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(gn, h_logn, sizeof(h_logn), cudaMemcpyHostToDevice));

  CHECK(tb_mixture_build(logl, C, n, gb, gz, gn, (int32_t)T, st));                 /* state_manager.py:466-471 */
  /* flags = 1: every stored beta is 0, skip the (rounding-decided) probe at beta_prev */
  CHECK(tb_next_beta(logl, C, n, 0.0, 2.0 * (double)n_gen, 1, ws, res, NULL, 0, st)); /* reweight.py:123-297 */
  double h_res[16];
  CHECK(cudaMemcpyAsync(h_res, res, sizeof(h_res), cudaMemcpyDeviceToHost, st));
  CHECK(cudaStreamSynchronize(st));
  printf("next beta %.10g  ESS %.3f (target %.1f)  logZ %.6f  probes %d\n", h_res[0], h_res[4], 2.0 * (double)n_gen,
         h_res[5], (int)h_res[6]);
  CHECK(tb_weights(logl, C, n, h_res[0], res + 1, w, st));                        /* w / sum w at that beta */
  CHECK(cudaStreamSynchronize(st));
  free(h_logl);
  return 0;
}
