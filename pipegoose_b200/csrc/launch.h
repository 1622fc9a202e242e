// C launch API shared between the CUDA translation units (nvcc) and the torch bindings (g++).
// Every entry point returns 0 on success and enqueues on the given stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_MAX_PEERS 8

// In-kernel data-parallel gradient reduce-scatter (ZeRO-1): the flat fp32 gradient buffer of every data-parallel rank is
// peer-mapped; elements [start, start + world * seg) are owned slice-wise (rank r owns [start + r*seg, start + (r+1)*seg)).
// A kernel that produces a gradient (wgrad GEMM epilogue, embedding backward, gradient fold) adds it with
// red.global.add straight into the OWNER's buffer — no local copy, no separate reduction kernel, nothing exposed
// after backward except one peer barrier.  world <= 1: plain local accumulation.
typedef struct PgGradRS {
  float* peer[PG_MAX_PEERS];  // flat gradient buffer of data-parallel rank r (peer[my rank] is the local buffer)
  const float* local;         // this rank's flat gradient buffer (gradient pointers handed to kernels point into it)
  long long start;            // first element of the in-kernel reduce-scatter region
  long long seg;              // elements per owner slice (multiple of 4)
  int world;
  int scalar_red;             // 1: four scalar red.global.add.f32 instead of one red.global.add.v4.f32
} PgGradRS;

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
  int k_splits;  // 0 = auto (fp32 outputs only), 1 = never split, >1 = forced
  int cta_pair;  // 0 = auto, -1 = 1-CTA tiles only, 1 = CTA-pair (cta_group::2, 256-row) tiles
  int max_ctas;  // 0 = all SMs
  int num_chunks, chunk_rows, first_chunk;
  const uint32_t* chunk_flags;
  uint32_t flag_value;
  void* out_peer[PG_MAX_PEERS];
  uint32_t* arrive_ctr[PG_MAX_PEERS];
  // GEMM -> reduce-scatter with the reduction fused into the local chunk's epilogue (see GemmArgs)
  const void* rs_in[PG_MAX_PEERS];
  const uint32_t* rs_wait_ctr;
  uint32_t rs_wait_value;
  // all-gather -> GEMM communication CTAs
  int n_comm;
  const void* ag_src[PG_MAX_PEERS];
  void* ag_dst;
  uint64_t ag_chunk_bytes;
  uint32_t* ag_peer_flag[PG_MAX_PEERS];
  const uint32_t* ag_ready;
  uint32_t ag_epoch;
  int my_rank;
  const void* a_local;  // all-gather -> GEMM: this rank's shard [chunk_rows, K] (read in place)
  // grouped GEMM / MoE combine
  int b_chunk_rows, bias_chunk_stride;
  const int* row_ret;
  const float* row_scale;
  int scatter_rows_per_src;
  // fp32 output (wgrad): add the tile into the owners' gradient buffers (see PgGradRS); world <= 1: off
  PgGradRS grad_rs;
  // lm_head: per-row online-softmax partials [M, 2 * ceil(N / block_n), 2] fp32 written by the epilogue (null: off);
  // columns >= ce_valid are padding.  block_n is forced to 256 so that the caller can size the buffer.
  float* ce_part;
  int ce_valid;
} PgGemmDesc;

// ---- attention_sm100.cu
int pg_attention_fwd(const void* qkv, const float* slopes, void* out, float* lse, int B, int S, int H, int D,
                     float softmax_scale, cudaStream_t s);
int pg_attention_bwd(const void* qkv, const float* slopes, const void* out, const float* lse, const void* dout,
                     void* dqkv, float* dq_acc, float* delta, int B, int S, int H, int D, float softmax_scale,
                     cudaStream_t s);

// ---- moe.cu
int pg_moe_route(const void* x, const void* wg, const void* bg, const float* jitter, int n, int h, int E, int top_k,
                 int capacity, float* probs, int* topk_idx, float* topk_prob, int* pos, int* counts,
                 float* prob_sum, float* zsum, float* lse_out, cudaStream_t s);
int pg_moe_dispatch(const void* x, const int* topk_idx, const float* topk_prob, const int* pos,
                    void* const* peer_buf, int* const* peer_row_ret, float* const* peer_row_scale,
                    uint32_t* const* peer_arrive, int n, int h, int top_k, int E_local, int T, int C, int my_rank,
                    int scale_by_prob, int blocks, cudaStream_t s);
int pg_fill_i32(int* p, int v, int64_t n, cudaStream_t s);

// ---- comm.cu
// out[rows, cols] = sum_src staging[src][rows, cols] (+ bias) (+ residual), after every source's
// arrival counter reached `expected`
int pg_rs_reduce(const void* staging, int num_src, int64_t src_stride_elems, const uint32_t* arrive_ctr,
                 uint32_t expected, const void* bias, const void* residual, void* out, int rows, int cols,
                 cudaStream_t s);
// flat fp32 gradient bucket: in-place all-reduce / reduce-scatter average over NVLink peers
// mc_buf != null: the reduce half is ONE multimem.ld_reduce per 16 bytes (summed inside the NVSwitch), the all-gather
// half ONE multimem.st
// [offset, offset + n) may be a run of buckets of bucket_elems elements (0: one bucket): each is reduced slice-wise
int pg_allreduce_f32(float* const* peer_bufs, float* mc_buf, int world, int rank, int64_t offset_elems, int64_t n,
                     int64_t bucket_elems, float scale, int reduce_scatter_only, uint32_t* const* peer_flags,
                     uint32_t epoch, int blocks, cudaStream_t s);
// ZeRO-1 parameter all-gather: regions [0, head) and then [head + k*bucket, ...) up to total; every rank pushes its
// 1/world slice of each region to all peers (mc_buf != null: ONE multimem.st per 16 bytes through the NVSwitch)
int pg_allgather_bf16(void* const* peer_bufs, void* mc_buf, int world, int rank, int64_t head_elems,
                      int64_t bucket_elems, int64_t total_elems, uint32_t* const* peer_flags, uint32_t epoch,
                      cudaStream_t s);
int pg_barrier_peers(uint32_t* const* peer_flags, int world, int rank, uint32_t epoch, cudaStream_t s);
// reduce-scatter without a GEMM: push row block c of x [num_chunks * chunk_elems] to rank c's staging slot and bump its
// arrival counter by `blocks` (one per CTA); the owner runs pg_rs_reduce on the slots
int pg_rs_push(const void* x, int num_chunks, int first_chunk, int64_t chunk_elems, void* const* out_peer,
               uint32_t* const* arrive_ctr, int blocks, cudaStream_t s);
// NVLS
int pg_multimem_selftest(const float* mc_in, float* mc_out, float* out, int64_t n, cudaStream_t s);
int pg_rs_reduce_mc(const void* mc_partial, const uint32_t* arrive_ctr, int num_src, uint32_t expected, const void* bias,
                    const void* residual, void* out, int rows, int cols, cudaStream_t s);
// symmetric memory (VMM + multicast; symm_vmm.cu)
int pg_vmm_probe(int world, int* mc_supported, int64_t* gran);
int pg_vmm_alloc(int64_t nbytes, void** ptr, int* fd, uint64_t* handle);
int pg_vmm_import(int fd, int64_t nbytes, void** ptr, uint64_t* handle);
int pg_vmm_unmap(void* ptr, int64_t nbytes, uint64_t handle);
int pg_mc_create(int world, int64_t nbytes, int* fd, uint64_t* mc_handle);
int pg_mc_import(int fd, uint64_t* mc_handle);
int pg_mc_add_device(uint64_t mc_handle);
int pg_mc_bind(uint64_t mc_handle, uint64_t mem_handle, int64_t nbytes, void** mc_ptr);
// symmetric memory (cudaIpc)
int pg_symm_alloc(int64_t nbytes, void** ptr, void* handle64);
int pg_symm_open(const void* handle64, void** ptr);
int pg_symm_close(void* ptr);
int pg_symm_free(void* ptr);

int pg_gemm_bf16(const PgGemmDesc* d, cudaStream_t stream);
// persistent GEMM grids use at most `ctas` CTAs (0 = all SMs): leaves SMs to overlapped communication kernels
void pg_set_gemm_cta_cap(int ctas);

// ---- elementwise.cu
int pg_layernorm_fwd(const void* x, const int64_t* ids, int vocab_start, int vocab_end,
                     const void* gamma, const void* beta, void* y, float* mean, float* rstd,
                     int rows, int h, float eps, int apply_ln, cudaStream_t s);
int pg_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                     const float* rstd, const void* dx_extra, void* dx, float* dgamma, float* dbeta,
                     int rows, int h, cudaStream_t s);
int pg_colsum(const void* x, int ld, float* out, int rows, int cols, cudaStream_t s);
int pg_embedding_bwd(const void* dx, const int64_t* ids, float* dw, int rows, int h,
                     int vocab_start, int vocab_end, const PgGradRS* grad_rs, cudaStream_t s);
// dst (a range of the local flat gradient buffer) += scale * src, added into the owners' buffers (src: bf16 or fp32)
int pg_grad_rs_accum(const void* src, int src_is_f32, float* dst_local, int64_t n, float scale,
                     const PgGradRS* grad_rs, cudaStream_t s);
int pg_ce_stats(const void* logits, int ld, const int64_t* targets, float* stats, int rows,
                int vocab_local, int vocab_start, cudaStream_t s);
// merge the GEMM epilogue's partials [rows, nparts, 2] into stats [rows, 3] = (max, sumexp, logit[target] or 0)
int pg_ce_combine(const float* part, int nparts, const void* logits, int ld, const int64_t* targets, float* stats,
                  int rows, int vocab_local, int vocab_start, cudaStream_t s);
int pg_ce_finalize(void* logits, int ld, const int64_t* targets, const float* gstats,
                   float* loss_rows, int rows, int vocab_local, int vocab_start,
                   const float* grad_scale, int64_t ignore_index, int write_grad, cudaStream_t s);
int pg_adam(float* master, float* m, float* v, const float* grad, void* param_bf16, int64_t n,
            float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2,
            float grad_scale, int adamw, int zero_grad, cudaStream_t s);
// ZeRO-1 form: ONE launch updates this rank's slice of nb equally sized buckets — optimizer state element b * seg + w
// belongs to flat element first + b * bucket_stride + w (w < seg)
int pg_adam_strided(float* master, float* m, float* v, float* grad, void* param_bf16, int64_t seg, int64_t bucket_stride,
                    int64_t nb, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2,
                    float grad_scale, int adamw, int zero_grad, cudaStream_t s);
int pg_sgd(float* master, float* mom, const float* grad, void* param_bf16, int64_t n, float lr,
           float momentum, float wd, float grad_scale, int first_step, cudaStream_t s);
int pg_accum_bf16_to_f32(const void* src, float* dst, int64_t n, float scale, int accumulate,
                         cudaStream_t s);

#ifdef __cplusplus
}
#endif
