// torch <-> kernel glue.  Only argument checking and pointer plumbing lives here; every op
// enqueues on the current CUDA stream and raises if the launch API reports a failure.
#include <torch/extension.h>
#include <c10/cuda/CUDAStream.h>
#include <c10/cuda/CUDAGuard.h>

#include "launch.h"

namespace {

cudaStream_t cur_stream() { return c10::cuda::getCurrentCUDAStream().stream(); }

void check_bf16_2d(const torch::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == torch::kBFloat16, name, " must be bf16");
  TORCH_CHECK(t.dim() == 2 && t.stride(1) == 1, name, " must be 2-D with unit inner stride");
  TORCH_CHECK(t.stride(0) % 8 == 0, name, " leading dimension must be a multiple of 8");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(t.data_ptr()) % 16 == 0, name, " must be 16B aligned");
}

// out[M,N] = epilogue(op(A) * op(B)).  a_mn/b_mn select the stored layout (see launch.h).
void gemm(const torch::Tensor& a, const torch::Tensor& b, torch::Tensor out, bool a_mn, bool b_mn,
          const c10::optional<torch::Tensor>& bias, const c10::optional<torch::Tensor>& residual,
          const c10::optional<torch::Tensor>& aux, int64_t flags, int64_t block_n,
          int64_t max_ctas, int64_t num_chunks, int64_t first_chunk, int64_t chunk_flags_ptr,
          int64_t flag_value, std::vector<int64_t> out_peer_ptrs,
          std::vector<int64_t> arrive_ctr_ptrs) {
  check_bf16_2d(a, "a");
  check_bf16_2d(b, "b");
  c10::cuda::CUDAGuard guard(a.device());
  PgGemmDesc d;
  memset(&d, 0, sizeof(d));
  const int64_t M = a_mn ? a.size(1) : a.size(0);
  const int64_t K = a_mn ? a.size(0) : a.size(1);
  const int64_t N = b_mn ? b.size(1) : b.size(0);
  const int64_t Kb = b_mn ? b.size(0) : b.size(1);
  TORCH_CHECK(K == Kb, "gemm: contraction dims differ: ", K, " vs ", Kb);
  TORCH_CHECK(N % 8 == 0, "gemm: N must be a multiple of 8");
  d.A = a.data_ptr();
  d.B = b.data_ptr();
  d.M = (int)M; d.N = (int)N; d.K = (int)K;
  d.lda = (int)a.stride(0);
  d.ldb = (int)b.stride(0);
  d.a_mn = a_mn; d.b_mn = b_mn;
  TORCH_CHECK(out.is_cuda() && out.dim() == 2 && out.stride(1) == 1, "out must be 2-D cuda");
  const bool f32 = out.scalar_type() == torch::kFloat32;
  TORCH_CHECK(f32 || out.scalar_type() == torch::kBFloat16, "out must be bf16 or fp32");
  if (out_peer_ptrs.empty()) {
    TORCH_CHECK(out.size(0) == M && out.size(1) == N, "out shape mismatch");
  }
  d.out = out.data_ptr();
  d.ldc = (int)out.stride(0);
  TORCH_CHECK(d.ldc % 8 == 0, "out leading dimension must be a multiple of 8");
  d.flags = (int)flags | (f32 ? 8 : 0);
  if (bias.has_value()) {
    TORCH_CHECK(bias->scalar_type() == torch::kBFloat16 && bias->numel() == N && bias->is_contiguous(), "bias must be contiguous bf16 [N]");
    d.bias = bias->data_ptr();
    d.flags |= 1;
  }
  if (residual.has_value()) {
    check_bf16_2d(*residual, "residual");
    TORCH_CHECK(residual->size(0) == M && residual->size(1) == N, "residual shape mismatch");
    d.residual = residual->data_ptr();
    d.ldr = (int)residual->stride(0);
    d.flags |= 4;
  }
  if (aux.has_value()) {
    check_bf16_2d(*aux, "aux");
    TORCH_CHECK(aux->size(0) == M && aux->size(1) == N && aux->stride(0) == out.stride(0), "aux must match out");
    d.aux = aux->data_ptr();
  }
  d.block_n = (int)block_n;
  d.max_ctas = (int)max_ctas;
  d.num_chunks = (int)num_chunks;
  d.first_chunk = (int)first_chunk;
  if (num_chunks > 1) {
    TORCH_CHECK(M % num_chunks == 0, "M must divide into chunks");
    d.chunk_rows = (int)(M / num_chunks);
    TORCH_CHECK(num_chunks <= PG_MAX_PEERS, "too many chunks");
  }
  d.chunk_flags = reinterpret_cast<const uint32_t*>(chunk_flags_ptr);
  d.flag_value = (uint32_t)flag_value;
  for (size_t i = 0; i < out_peer_ptrs.size() && i < PG_MAX_PEERS; ++i) d.out_peer[i] = reinterpret_cast<void*>(out_peer_ptrs[i]);
  for (size_t i = 0; i < arrive_ctr_ptrs.size() && i < PG_MAX_PEERS; ++i) d.arrive_ctr[i] = reinterpret_cast<uint32_t*>(arrive_ctr_ptrs[i]);
  TORCH_CHECK(pg_gemm_bf16(&d, cur_stream()) == 0, "pg_gemm_bf16 failed");
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("gemm", &gemm, "tcgen05 bf16 GEMM with fused epilogue / chunked collectives",
        py::arg("a"), py::arg("b"), py::arg("out"), py::arg("a_mn") = false, py::arg("b_mn") = false,
        py::arg("bias") = py::none(), py::arg("residual") = py::none(), py::arg("aux") = py::none(),
        py::arg("flags") = 0, py::arg("block_n") = 0, py::arg("max_ctas") = 0,
        py::arg("num_chunks") = 1, py::arg("first_chunk") = 0, py::arg("chunk_flags_ptr") = 0,
        py::arg("flag_value") = 0, py::arg("out_peer_ptrs") = std::vector<int64_t>{},
        py::arg("arrive_ctr_ptrs") = std::vector<int64_t>{});
}
