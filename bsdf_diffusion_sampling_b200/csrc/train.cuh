// Training-step launch parameters (train.cu).
#pragma once
#include <cuda_runtime.h>

namespace bsdfdiff {

struct TrainParams {
    int domain, in_dim, hidden, n_hidden;
    long long n;
    const float* x0;          // [n,2] base samples
    const float* x1;          // [n,2] targets (omega_o)
    const float* wi;          // [n,2] conditioning (omega_i, domain coordinates)
    const float* alpha;       // [n] or null = linspace(0, 1, n)
    float* weights;           // fp32 master weights, torch layout: W1 [H,in], W2.. [H,H], Wout [2,H], concatenated
    float* grad;              // same layout; zero on entry; holds dL/dW on exit when apply_update == 0, zero otherwise
    float* adam_m;            // first / second moment estimates (same layout)
    float* adam_v;
    float lr, beta1, beta2, eps;
    long long step;           // 1-based optimiser step (bias correction)
    int apply_update;
    float* loss;              // device scalar, zero on entry
    unsigned int* ticket;     // device counter, zero on entry (and on exit)
};

int launch_flow_matching_step(const TrainParams& P, cudaStream_t stream);
int launch_base_nll_step(const TrainParams& P, cudaStream_t stream);      // weights = the 308-float base blob, x1 = omega_o

}  // namespace bsdfdiff
