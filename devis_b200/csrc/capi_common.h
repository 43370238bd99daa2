// capi_common.h -- bookkeeping shared by the translation units of libdevis_msda.so (defined in msda_capi.cu):
// every kernel launch goes through devis_capi_check_launch(), which counts it (devis_msda_launch_count) and turns a
// launch failure into DEVIS_MSDA_ERR_CUDA with the CUDA error kept for devis_msda_last_cuda_error().
#pragma once
#include <cuda_runtime.h>

int devis_capi_cuda_fail(cudaError_t e);
int devis_capi_check_launch(int family);   // family: DEVIS_MSDA_KERNEL_* (include/devis_msda.h)
int devis_capi_helper_blocks();            // 8 blocks per SM of the current device, for grid-stride helper kernels
