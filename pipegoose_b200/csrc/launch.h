// C launch API shared between the CUDA translation units (nvcc) and the torch bindings (g++).
// Every entry point returns 0 on success and enqueues on the given stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_MAX_PEERS 8

typedef struct PgGemmDesc {
  const void* A;  // bf16
  const void* B;  // bf16
  int M, N, K;
  int lda, ldb;
  int a_mn;  // 0: A is [M, K] row-major (K-major); 1: A is [K, M] row-major (MN-major)
  int b_mn;  // 0: B is [N, K] row-major (K-major); 1: B is [K, N] row-major (MN-major)
  void* out;
  int ldc;
  const void* bias;      // bf16 [N]
  const void* residual;  // bf16 [M, ldr]
  int ldr;
  void* aux;  // bf16 [M, ldc]
  int flags;  // EpiFlags
  int block_n;   // 0 = auto
  int max_ctas;  // 0 = all SMs
  int num_chunks, chunk_rows, first_chunk;
  const uint32_t* chunk_flags;
  uint32_t flag_value;
  void* out_peer[PG_MAX_PEERS];
  uint32_t* arrive_ctr[PG_MAX_PEERS];
} PgGemmDesc;

int pg_gemm_bf16(const PgGemmDesc* d, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
