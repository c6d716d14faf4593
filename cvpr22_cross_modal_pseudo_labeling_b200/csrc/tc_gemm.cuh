// tc_gemm.cuh -- entry points of the persistent tcgen05 GEMM (tc_gemm.cu) used by embed_match.cu.
#pragma once
#include "common.cuh"

namespace b200 {

// SOFTMAX scoring for class matrices of up to 512 columns (one or two column blocks of <= 256)
bool softmax_gemm_applies(int n_cols);
int softmax_gemm_launch(const void* A_bf16, const void* E_bf16, int64_t n_rows, int n_cols, int dim, int ld,
                        float score_thresh, float* probs, float* logits, int32_t* top_label, float* top_prob,
                        cudaStream_t st);

// SOFTMAX scoring for class matrices of any width (two GEMM passes per row tile, no logits round trip);
// n_cols < 65536
int softmax_wide_launch(const void* A_bf16, const void* E_bf16, int64_t n_rows, int n_cols, int dim, float score_thresh,
                        float* probs, float* logits, int32_t* top_label, float* top_prob, cudaStream_t st);

}  // namespace b200
