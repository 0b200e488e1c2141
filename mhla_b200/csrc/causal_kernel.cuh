// Causal chunked MHLA forward (variant C) -- replaces mhla_nlp/fla/ops/mhla/naive.py:10-83.
#pragma once
#include <string>
#include <cuda_runtime.h>
#include "../../include/mhla_b200.h"
#include "ptx.cuh"

namespace mhla {

inline size_t causal_workspace_bytes(const mhla_causal_desc* d) {
  (void)d;
  return 0;
}

inline int causal_forward(const mhla_causal_desc* d, cudaStream_t stream, int* launches, std::string* err) {
  (void)d; (void)stream; (void)launches; (void)err;
  return MHLA_ERR_UNSUPPORTED_SHAPE;
}

}  // namespace mhla
