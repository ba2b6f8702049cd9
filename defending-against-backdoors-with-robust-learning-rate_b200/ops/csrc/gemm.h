// tcgen05 / TMA GEMM + implicit-GEMM convolution launchers (see gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
