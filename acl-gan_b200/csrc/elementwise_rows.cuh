// Row-structured fast paths of the plane-sized element-wise passes (elementwise_rows.cu).  Each returns -100 when the
// arguments are outside what the fast path covers (the caller then runs the generic kernels of elementwise.cu).
#pragma once
#include "common.cuh"

namespace aclgan {
int rows_norm_apply(const aclgan_apply_args* a, cudaStream_t st);
int rows_norm_finalize_apply(const aclgan_norm_finalize_args* f, const aclgan_apply_args* a, cudaStream_t st);
int rows_bwd_reduce(const aclgan_block_bwd_args* a, cudaStream_t st);
int rows_bwd_apply(const aclgan_block_bwd_args* a, cudaStream_t st);
int rows_bwd_finalize_apply(const aclgan_norm_bwd_finalize_args* f, const aclgan_block_bwd_args* a, cudaStream_t st);
}  // namespace aclgan
