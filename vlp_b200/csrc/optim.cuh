// vlp_b200 — fused multi-tensor BertAdam step (see optim.cu).
#pragma once
#include "../../include/vlpk.h"
#include "common.cuh"

namespace vlpk {

constexpr int ADAM_CHUNK = 4096;  // elements of one tensor handled by one work item

struct AdamHyper {
  float lr;        // schedule already applied (optimization.py:165-170)
  float b1, omb1;  // beta1, 1 - beta1 (difference taken in double like the reference's Python scalar)
  float b2, omb2;
  float eps;
  float max_grad_norm;  // <= 0: no clipping
};

int launch_bertadam(const VlpkAdamTensor* tensors_host, const VlpkAdamTensor* tensors_dev, const int32_t* prefix_host,
                    const int32_t* prefix_dev, int n_tensors, float* sqnorm_dev, const AdamHyper& h, cudaStream_t s);

}  // namespace vlpk
