// 1x1 conv with BN+ReLU pre-activation, A operand through tensor memory (see tn_conv1x1_ts.cu).
#pragma once
#include "tn_conv_gemm.h"

namespace tn {

// true when launch_conv1x1_ts can run this problem (TMA-eligible 1x1/stride-1 conv with a scale/shift prologue, 64 < Cout <= 128,
// K <= 512, 32-byte aligned output rows); TN_NO_TS=1 disables the path (A/B against the shared-memory transform kernel)
bool conv1x1_ts_eligible(const ConvGemmParams& p);
cudaError_t launch_conv1x1_ts(const ConvGemmParams& p, int num_sms, cudaStream_t stream);

}  // namespace tn
