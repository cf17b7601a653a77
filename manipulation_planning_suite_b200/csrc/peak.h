#pragma once
#include <cuda_runtime.h>
cudaError_t mps_run_fma_peak(int num_sms, cudaStream_t stream, float* tflops);
