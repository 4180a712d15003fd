// TEST INFRASTRUCTURE ONLY: storage for the emulated CUDA built-in variables (see cuda_sim.h).
#include "cuda_sim.h"
dim3 threadIdx, blockIdx, blockDim, gridDim;
