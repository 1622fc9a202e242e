// torch <-> kernel glue.  Only argument checking and pointer plumbing lives here; every op
// enqueues on the current CUDA stream and raises if the launch API reports a failure.
#include <torch/extension.h>
#include <c10/cuda/CUDAStream.h>
#include <c10/cuda/CUDAGuard.h>

#include "launch.h"
#include <cmath>

namespace {

cudaStream_t cur_stream() { return c10::cuda::getCurrentCUDAStream().stream(); }

bool parse_grad_rs(const py::dict& d, PgGradRS* out);

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
          std::vector<int64_t> arrive_ctr_ptrs, const py::dict& ag) {
  check_bf16_2d(a, "a");
  check_bf16_2d(b, "b");
  c10::cuda::CUDAGuard guard(a.device());
  PgGemmDesc d;
  memset(&d, 0, sizeof(d));
  const int64_t M = a_mn ? a.size(1) : a.size(0);
  const int64_t K = a_mn ? a.size(0) : a.size(1);
  const int64_t groups = ag.contains("b_chunk_rows") ? num_chunks : 1;  // stacked (grouped) B
  const int64_t N = b_mn ? b.size(1) : b.size(0) / groups;
  const int64_t Kb = b_mn ? b.size(0) / groups : b.size(1);
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
    TORCH_CHECK(bias->scalar_type() == torch::kBFloat16 && bias->numel() % N == 0 && bias->is_contiguous(), "bias must be contiguous bf16 [N] (or [groups, N])");
    d.bias = bias->data_ptr();
    d.flags |= 1;
  }
  if (residual.has_value()) {
    check_bf16_2d(*residual, "residual");
    // fused reduce-scatter: the residual is the local token shard [M / num_chunks, N]
    TORCH_CHECK((residual->size(0) == M || (ag.contains("rs_in") && residual->size(0) * num_chunks == M)) &&
                    residual->size(1) == N, "residual shape mismatch");
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
  if (ag.contains("b_chunk_rows")) {
    d.b_chunk_rows = ag["b_chunk_rows"].cast<int>();
    d.bias_chunk_stride = ag.contains("bias_chunk_stride") ? ag["bias_chunk_stride"].cast<int>() : 0;
    if (ag.contains("row_ret")) {
      d.row_ret = reinterpret_cast<const int*>(ag["row_ret"].cast<int64_t>());
      d.row_scale = reinterpret_cast<const float*>(ag["row_scale"].cast<int64_t>());
      d.scatter_rows_per_src = ag["rows_per_src"].cast<int>();
    }
  }
  if (ag.contains("rs_in")) {
    auto in = ag["rs_in"].cast<std::vector<int64_t>>();
    for (size_t i = 0; i < in.size() && i < PG_MAX_PEERS; ++i) d.rs_in[i] = reinterpret_cast<const void*>(in[i]);
    d.rs_wait_ctr = reinterpret_cast<const uint32_t*>(ag["rs_wait"].cast<int64_t>());
    d.rs_wait_value = (uint32_t)ag["rs_wait_value"].cast<int64_t>();
    d.my_rank = ag["rank"].cast<int>();
  }
  if (ag.contains("ce_part")) {
    // lm_head + cross entropy: the epilogue also writes the online-softmax partials (fp32 [M, 2 * ceil(N/256), 2])
    TORCH_CHECK(!f32 && num_chunks >= 1, "ce_part needs a bf16 logits output");
    d.ce_part = reinterpret_cast<float*>(ag["ce_part"].cast<int64_t>());
    d.ce_valid = ag.contains("ce_valid") ? ag["ce_valid"].cast<int>() : 0;
  }
  if (ag.contains("grad_rs")) {
    TORCH_CHECK(f32, "grad_rs needs an fp32 output");
    parse_grad_rs(ag["grad_rs"].cast<py::dict>(), &d.grad_rs);
  }
  if (ag.contains("k_splits")) d.k_splits = ag["k_splits"].cast<int>();
  if (ag.contains("cta_pair")) d.cta_pair = ag["cta_pair"].cast<int>();
  if (ag.contains("n_comm")) {
    d.n_comm = ag["n_comm"].cast<int>();
    d.ag_dst = reinterpret_cast<void*>(ag["dst"].cast<int64_t>());
    d.ag_chunk_bytes = (uint64_t)ag["chunk_bytes"].cast<int64_t>();
    d.ag_ready = reinterpret_cast<const uint32_t*>(ag["ready"].cast<int64_t>());
    d.ag_epoch = (uint32_t)ag["epoch"].cast<int64_t>();
    d.my_rank = ag["rank"].cast<int>();
    d.a_local = reinterpret_cast<const void*>(ag["local"].cast<int64_t>());
    auto src = ag["src"].cast<std::vector<int64_t>>();
    auto pf = ag["peer_flag"].cast<std::vector<int64_t>>();
    for (size_t i = 0; i < src.size() && i < PG_MAX_PEERS; ++i) d.ag_src[i] = reinterpret_cast<const void*>(src[i]);
    for (size_t i = 0; i < pf.size() && i < PG_MAX_PEERS; ++i) d.ag_peer_flag[i] = reinterpret_cast<uint32_t*>(pf[i]);
  }
  TORCH_CHECK(pg_gemm_bf16(&d, cur_stream()) == 0, "pg_gemm_bf16 failed");
}


// {"peer": [ptr per data-parallel rank], "local": ptr, "start": elems, "seg": elems, "scalar": 0/1} -> PgGradRS
bool parse_grad_rs(const py::dict& d, PgGradRS* out) {
  memset(out, 0, sizeof(*out));
  if (!d.contains("peer")) return false;
  auto peers = d["peer"].cast<std::vector<int64_t>>();
  TORCH_CHECK(peers.size() >= 2 && peers.size() <= PG_MAX_PEERS, "grad_rs: 2..", PG_MAX_PEERS, " peers");
  for (size_t i = 0; i < peers.size(); ++i) out->peer[i] = reinterpret_cast<float*>(peers[i]);
  out->local = reinterpret_cast<const float*>(d["local"].cast<int64_t>());
  out->start = d["start"].cast<int64_t>();
  out->seg = d["seg"].cast<int64_t>();
  out->world = (int)peers.size();
  out->scalar_red = d.contains("scalar") ? d["scalar"].cast<int>() : 0;
  TORCH_CHECK(out->seg > 0 && out->seg % 4 == 0 && out->start % 4 == 0, "grad_rs: start / seg must be multiples of 4");
  return true;
}

#define PG_CUDA(t) TORCH_CHECK((t).is_cuda() && (t).is_contiguous(), #t " must be a contiguous CUDA tensor")
#define PG_BF16(t) TORCH_CHECK((t).scalar_type() == torch::kBFloat16, #t " must be bf16")
#define PG_F32(t) TORCH_CHECK((t).scalar_type() == torch::kFloat32, #t " must be fp32")

const void* opt_ptr(const c10::optional<torch::Tensor>& t) { return t.has_value() ? t->data_ptr() : nullptr; }

// y = LN(x) (or LN(table[ids]) / table[ids] when ids is given); returns nothing, fills y/mean/rstd
void layernorm_fwd(const torch::Tensor& x, const c10::optional<torch::Tensor>& ids, int64_t vocab_start,
                   int64_t vocab_end, const torch::Tensor& gamma, const torch::Tensor& beta, torch::Tensor y,
                   c10::optional<torch::Tensor> mean, c10::optional<torch::Tensor> rstd, double eps, bool apply_ln) {
  PG_CUDA(x); PG_BF16(x); PG_CUDA(y); PG_BF16(y); PG_CUDA(gamma); PG_BF16(gamma); PG_CUDA(beta); PG_BF16(beta);
  c10::cuda::CUDAGuard guard(x.device());
  const int h = (int)y.size(-1);
  const int rows = (int)(y.numel() / h);
  TORCH_CHECK(x.size(-1) == h, "hidden size mismatch");
  const int64_t* idp = nullptr;
  if (ids.has_value()) { PG_CUDA(*ids); TORCH_CHECK(ids->scalar_type() == torch::kInt64 && ids->numel() == rows, "ids must be int64 [rows]"); idp = ids->data_ptr<int64_t>(); }
  else { TORCH_CHECK(x.numel() == y.numel(), "x/y size mismatch"); }
  float* mp = mean.has_value() ? mean->data_ptr<float>() : nullptr;
  float* rp = rstd.has_value() ? rstd->data_ptr<float>() : nullptr;
  TORCH_CHECK(pg_layernorm_fwd(x.data_ptr(), idp, (int)vocab_start, (int)vocab_end, gamma.data_ptr(), beta.data_ptr(),
                               y.data_ptr(), mp, rp, rows, h, (float)eps, apply_ln, cur_stream()) == 0, "layernorm_fwd failed");
}

void layernorm_bwd(const torch::Tensor& dy, const torch::Tensor& x, const torch::Tensor& gamma, const torch::Tensor& mean,
                   const torch::Tensor& rstd, const c10::optional<torch::Tensor>& dx_extra, torch::Tensor dx,
                   c10::optional<torch::Tensor> dgamma, c10::optional<torch::Tensor> dbeta) {
  PG_CUDA(dy); PG_BF16(dy); PG_CUDA(x); PG_BF16(x); PG_CUDA(dx); PG_BF16(dx); PG_CUDA(mean); PG_F32(mean); PG_CUDA(rstd); PG_F32(rstd);
  c10::cuda::CUDAGuard guard(x.device());
  const int h = (int)x.size(-1);
  const int rows = (int)(x.numel() / h);
  float* dg = nullptr; float* db = nullptr;
  if (dgamma.has_value()) { PG_F32(*dgamma); PG_F32(*dbeta); dg = dgamma->data_ptr<float>(); db = dbeta->data_ptr<float>(); }
  TORCH_CHECK(pg_layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr<float>(), rstd.data_ptr<float>(),
                               opt_ptr(dx_extra), dx.data_ptr(), dg, db, rows, h, cur_stream()) == 0, "layernorm_bwd failed");
}

void colsum(const torch::Tensor& x, torch::Tensor out) {
  check_bf16_2d(x, "x"); PG_CUDA(out); PG_F32(out);
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(out.numel() == x.size(1), "out must have one entry per column");
  TORCH_CHECK(pg_colsum(x.data_ptr(), (int)x.stride(0), out.data_ptr<float>(), (int)x.size(0), (int)x.size(1), cur_stream()) == 0, "colsum failed");
}

void embedding_bwd(const torch::Tensor& dx, const torch::Tensor& ids, torch::Tensor dw, int64_t vocab_start, int64_t vocab_end,
                   const py::dict& grad_rs) {
  PG_CUDA(dx); PG_BF16(dx); PG_CUDA(ids); PG_CUDA(dw); PG_F32(dw);
  c10::cuda::CUDAGuard guard(dx.device());
  const int h = (int)dx.size(-1);
  const int rows = (int)(dx.numel() / h);
  TORCH_CHECK(ids.scalar_type() == torch::kInt64 && ids.numel() == rows, "ids must be int64 [rows]");
  PgGradRS grs;
  const bool inl = parse_grad_rs(grad_rs, &grs);
  TORCH_CHECK(!inl || h % 8 == 0, "embedding_bwd with grad_rs: hidden size must be a multiple of 8");
  TORCH_CHECK(pg_embedding_bwd(dx.data_ptr(), ids.data_ptr<int64_t>(), dw.data_ptr<float>(), rows, h, (int)vocab_start, (int)vocab_end,
                               inl ? &grs : nullptr, cur_stream()) == 0, "embedding_bwd failed");
}

void ce_stats(const torch::Tensor& logits, const torch::Tensor& targets, torch::Tensor stats, int64_t vocab_start) {
  check_bf16_2d(logits, "logits"); PG_CUDA(targets); PG_CUDA(stats); PG_F32(stats);
  c10::cuda::CUDAGuard guard(logits.device());
  const int rows = (int)logits.size(0);
  TORCH_CHECK(targets.scalar_type() == torch::kInt64 && targets.numel() == rows && stats.numel() == rows * 3, "bad targets/stats");
  TORCH_CHECK(pg_ce_stats(logits.data_ptr(), (int)logits.stride(0), targets.data_ptr<int64_t>(), stats.data_ptr<float>(), rows, (int)logits.size(1), (int)vocab_start, cur_stream()) == 0, "ce_stats failed");
}

void ce_combine(const torch::Tensor& part, const torch::Tensor& logits, const torch::Tensor& targets, torch::Tensor stats,
                int64_t vocab_start) {
  check_bf16_2d(logits, "logits"); PG_CUDA(part); PG_F32(part); PG_CUDA(targets); PG_CUDA(stats); PG_F32(stats);
  c10::cuda::CUDAGuard guard(logits.device());
  const int rows = (int)logits.size(0);
  TORCH_CHECK(part.dim() == 3 && part.size(0) == rows && part.size(2) == 2, "part must be [rows, nparts, 2]");
  TORCH_CHECK(targets.scalar_type() == torch::kInt64 && targets.numel() == rows && stats.numel() == rows * 3, "bad targets/stats");
  TORCH_CHECK(pg_ce_combine(part.data_ptr<float>(), (int)part.size(1), logits.data_ptr(), (int)logits.stride(0), targets.data_ptr<int64_t>(),
                            stats.data_ptr<float>(), rows, (int)logits.size(1), (int)vocab_start, cur_stream()) == 0, "ce_combine failed");
}

void ce_finalize(torch::Tensor logits, const torch::Tensor& targets, const torch::Tensor& gstats, torch::Tensor loss_rows,
                 int64_t vocab_start, const c10::optional<torch::Tensor>& grad_scale, int64_t ignore_index, bool write_grad) {
  const float* gsp = nullptr;
  if (grad_scale.has_value()) { PG_CUDA(*grad_scale); PG_F32(*grad_scale); gsp = grad_scale->data_ptr<float>(); }
  check_bf16_2d(logits, "logits"); PG_CUDA(targets); PG_CUDA(gstats); PG_F32(gstats); PG_CUDA(loss_rows); PG_F32(loss_rows);
  c10::cuda::CUDAGuard guard(logits.device());
  const int rows = (int)logits.size(0);
  TORCH_CHECK(pg_ce_finalize(logits.data_ptr(), (int)logits.stride(0), targets.data_ptr<int64_t>(), gstats.data_ptr<float>(), loss_rows.data_ptr<float>(),
                             rows, (int)logits.size(1), (int)vocab_start, gsp, ignore_index, write_grad, cur_stream()) == 0, "ce_finalize failed");
}

void adam_step(torch::Tensor master, torch::Tensor m, torch::Tensor v, const torch::Tensor& grad, c10::optional<torch::Tensor> param_bf16,
               double lr, double beta1, double beta2, double eps, double wd, int64_t step, double grad_scale, bool adamw,
               bool zero_grad) {
  PG_CUDA(master); PG_F32(master); PG_CUDA(m); PG_F32(m); PG_CUDA(v); PG_F32(v); PG_CUDA(grad); PG_F32(grad);
  c10::cuda::CUDAGuard guard(master.device());
  const int64_t n = master.numel();
  TORCH_CHECK(m.numel() == n && v.numel() == n && grad.numel() == n, "adam: size mismatch");
  void* pb = nullptr;
  if (param_bf16.has_value()) { PG_CUDA(*param_bf16); PG_BF16(*param_bf16); TORCH_CHECK(param_bf16->numel() == n, "adam: bf16 param size mismatch"); pb = param_bf16->data_ptr(); }
  const double bc1 = 1.0 - std::pow(beta1, (double)step), bc2 = 1.0 - std::pow(beta2, (double)step);
  TORCH_CHECK(pg_adam(master.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), grad.data_ptr<float>(), pb, n, (float)lr, (float)beta1, (float)beta2,
                      (float)eps, (float)wd, (float)bc1, (float)bc2, (float)grad_scale, adamw, zero_grad, cur_stream()) == 0, "adam failed");
}

// ZeRO-1: this rank's slice (seg elements at flat offset first + b * bucket_stride) of nb equally sized buckets in ONE launch;
// master / m / v hold the nb slices back to back
void adam_step_strided(torch::Tensor master, torch::Tensor m, torch::Tensor v, torch::Tensor flat_grad, c10::optional<torch::Tensor> flat_param,
                       int64_t first, int64_t seg, int64_t bucket_stride, int64_t nb, double lr, double beta1, double beta2, double eps,
                       double wd, int64_t step, double grad_scale, bool adamw, bool zero_grad) {
  PG_CUDA(master); PG_F32(master); PG_CUDA(m); PG_F32(m); PG_CUDA(v); PG_F32(v); PG_CUDA(flat_grad); PG_F32(flat_grad);
  c10::cuda::CUDAGuard guard(master.device());
  TORCH_CHECK(master.numel() == seg * nb && m.numel() == seg * nb && v.numel() == seg * nb, "adam_strided: state size mismatch");
  TORCH_CHECK(first >= 0 && first + (nb - 1) * bucket_stride + seg <= flat_grad.numel(), "adam_strided: range outside the flat buffer");
  void* pb = nullptr;
  if (flat_param.has_value()) { PG_CUDA(*flat_param); PG_BF16(*flat_param); TORCH_CHECK(flat_param->numel() == flat_grad.numel(), "flat sizes differ");
    pb = reinterpret_cast<char*>(flat_param->data_ptr()) + 2 * first; }
  const double bc1 = 1.0 - std::pow(beta1, (double)step), bc2 = 1.0 - std::pow(beta2, (double)step);
  TORCH_CHECK(pg_adam_strided(master.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), flat_grad.data_ptr<float>() + first, pb, seg,
                              bucket_stride, nb, (float)lr, (float)beta1, (float)beta2, (float)eps, (float)wd, (float)bc1, (float)bc2,
                              (float)grad_scale, adamw, zero_grad, cur_stream()) == 0, "adam_strided failed");
}

void sgd_step(torch::Tensor master, c10::optional<torch::Tensor> mom, const torch::Tensor& grad, c10::optional<torch::Tensor> param_bf16,
              double lr, double momentum, double wd, double grad_scale, bool first_step) {
  PG_CUDA(master); PG_F32(master); PG_CUDA(grad); PG_F32(grad);
  c10::cuda::CUDAGuard guard(master.device());
  const int64_t n = master.numel();
  float* mp = nullptr;
  if (mom.has_value()) { PG_CUDA(*mom); PG_F32(*mom); mp = mom->data_ptr<float>(); }
  TORCH_CHECK(momentum == 0.0 || mp != nullptr, "sgd: momentum buffer required");
  void* pb = nullptr;
  if (param_bf16.has_value()) { PG_CUDA(*param_bf16); PG_BF16(*param_bf16); pb = param_bf16->data_ptr(); }
  TORCH_CHECK(pg_sgd(master.data_ptr<float>(), mp, grad.data_ptr<float>(), pb, n, (float)lr, (float)momentum, (float)wd, (float)grad_scale, first_step, cur_stream()) == 0, "sgd failed");
}

void accum_bf16_to_f32(const torch::Tensor& src, torch::Tensor dst, double scale, bool accumulate) {
  PG_CUDA(src); PG_BF16(src); PG_CUDA(dst); PG_F32(dst);
  c10::cuda::CUDAGuard guard(src.device());
  TORCH_CHECK(src.numel() == dst.numel(), "size mismatch");
  TORCH_CHECK(pg_accum_bf16_to_f32(src.data_ptr(), dst.data_ptr<float>(), src.numel(), (float)scale, accumulate, cur_stream()) == 0, "accum failed");
}


// dst (a view into the local flat fp32 gradient buffer) += scale * src through the in-kernel reduce-scatter
void grad_rs_accum(const torch::Tensor& src, torch::Tensor dst, double scale, const py::dict& grad_rs) {
  PG_CUDA(src); PG_CUDA(dst); PG_F32(dst);
  TORCH_CHECK(src.scalar_type() == torch::kBFloat16 || src.scalar_type() == torch::kFloat32, "src must be bf16 or fp32");
  TORCH_CHECK(src.numel() == dst.numel(), "size mismatch");
  c10::cuda::CUDAGuard guard(src.device());
  PgGradRS grs;
  TORCH_CHECK(parse_grad_rs(grad_rs, &grs), "grad_rs_accum needs a grad_rs descriptor");
  TORCH_CHECK(pg_grad_rs_accum(src.data_ptr(), src.scalar_type() == torch::kFloat32, dst.data_ptr<float>(), src.numel(), (float)scale,
                               &grs, cur_stream()) == 0, "grad_rs_accum failed");
}

void attention_fwd(const torch::Tensor& qkv, const torch::Tensor& slopes, torch::Tensor out, torch::Tensor lse,
                   int64_t B, int64_t S, int64_t H, int64_t D, double softmax_scale) {
  PG_CUDA(qkv); PG_BF16(qkv); PG_CUDA(out); PG_BF16(out); PG_CUDA(slopes); PG_F32(slopes); PG_CUDA(lse); PG_F32(lse);
  c10::cuda::CUDAGuard guard(qkv.device());
  TORCH_CHECK(qkv.numel() == B * S * H * 3 * D && out.numel() == B * S * H * D && lse.numel() == B * H * S && slopes.numel() == H, "attention_fwd: shape mismatch");
  TORCH_CHECK(pg_attention_fwd(qkv.data_ptr(), slopes.data_ptr<float>(), out.data_ptr(), lse.data_ptr<float>(), (int)B, (int)S, (int)H, (int)D, (float)softmax_scale, cur_stream()) == 0, "attention_fwd failed");
}

void attention_bwd(const torch::Tensor& qkv, const torch::Tensor& slopes, const torch::Tensor& out, const torch::Tensor& lse,
                   const torch::Tensor& dout, torch::Tensor dqkv, int64_t B, int64_t S, int64_t H, int64_t D,
                   double softmax_scale) {
  PG_CUDA(qkv); PG_BF16(qkv); PG_CUDA(out); PG_BF16(out); PG_CUDA(dout); PG_BF16(dout); PG_CUDA(dqkv); PG_BF16(dqkv); PG_CUDA(lse); PG_F32(lse);
  c10::cuda::CUDAGuard guard(qkv.device());
  TORCH_CHECK(dout.numel() == out.numel() && dqkv.numel() == qkv.numel(), "attention_bwd: shape mismatch");
  auto dq_acc = torch::empty({B * S, H * D}, qkv.options().dtype(torch::kFloat32));
  auto delta = torch::empty({B, H, S}, qkv.options().dtype(torch::kFloat32));
  TORCH_CHECK(pg_attention_bwd(qkv.data_ptr(), slopes.data_ptr<float>(), out.data_ptr(), lse.data_ptr<float>(), dout.data_ptr(), dqkv.data_ptr(),
                               dq_acc.data_ptr<float>(), delta.data_ptr<float>(), (int)B, (int)S, (int)H, (int)D, (float)softmax_scale, cur_stream()) == 0, "attention_bwd failed");
}

void moe_route(const torch::Tensor& x, const torch::Tensor& wg, const c10::optional<torch::Tensor>& bg,
               const c10::optional<torch::Tensor>& jitter, int64_t top_k, int64_t capacity, torch::Tensor probs,
               torch::Tensor topk_idx, torch::Tensor topk_prob, torch::Tensor pos, torch::Tensor counts,
               torch::Tensor prob_sum, torch::Tensor zsum, torch::Tensor lse) {
  PG_CUDA(x); PG_BF16(x); PG_CUDA(wg); PG_BF16(wg); PG_CUDA(probs); PG_F32(probs);
  c10::cuda::CUDAGuard guard(x.device());
  const int n = (int)x.size(0), h = (int)x.size(1), E = (int)wg.size(0);
  const float* jp = jitter.has_value() ? jitter->data_ptr<float>() : nullptr;
  TORCH_CHECK(pg_moe_route(x.data_ptr(), wg.data_ptr(), opt_ptr(bg), jp, n, h, E, (int)top_k, (int)capacity,
                           probs.data_ptr<float>(), topk_idx.data_ptr<int>(), topk_prob.data_ptr<float>(), pos.data_ptr<int>(),
                           counts.data_ptr<int>(), prob_sum.data_ptr<float>(), zsum.data_ptr<float>(), lse.data_ptr<float>(), cur_stream()) == 0, "moe_route failed");
}

void moe_dispatch(const torch::Tensor& x, const torch::Tensor& topk_idx, const torch::Tensor& topk_prob, const torch::Tensor& pos,
                  std::vector<int64_t> peer_buf, std::vector<int64_t> peer_row_ret, std::vector<int64_t> peer_row_scale,
                  std::vector<int64_t> peer_arrive, int64_t top_k, int64_t E_local, int64_t C, int64_t my_rank, bool scale_by_prob, int64_t blocks) {
  PG_CUDA(x); PG_BF16(x);
  c10::cuda::CUDAGuard guard(x.device());
  const int T = (int)peer_buf.size();
  void* bufs[PG_MAX_PEERS]; int* rets[PG_MAX_PEERS]; float* scs[PG_MAX_PEERS]; uint32_t* arr[PG_MAX_PEERS];
  for (int i = 0; i < T; ++i) { bufs[i] = reinterpret_cast<void*>(peer_buf[i]); rets[i] = reinterpret_cast<int*>(peer_row_ret[i]);
    scs[i] = reinterpret_cast<float*>(peer_row_scale[i]); arr[i] = reinterpret_cast<uint32_t*>(peer_arrive[i]); }
  TORCH_CHECK(pg_moe_dispatch(x.data_ptr(), topk_idx.data_ptr<int>(), topk_prob.data_ptr<float>(), pos.data_ptr<int>(), bufs, rets, scs, arr,
                              (int)x.size(0), (int)x.size(1), (int)top_k, (int)E_local, T, (int)C, (int)my_rank, scale_by_prob, (int)blocks, cur_stream()) == 0, "moe_dispatch failed");
}

// ---------------------------------------------------------------- symmetric memory / collectives
py::tuple symm_alloc(int64_t nbytes) {
  void* ptr = nullptr;
  char handle[64];
  TORCH_CHECK(pg_symm_alloc(nbytes, &ptr, handle) == 0, "symm_alloc failed");
  return py::make_tuple(reinterpret_cast<int64_t>(ptr), py::bytes(handle, 64));
}
int64_t symm_open(const std::string& handle) {
  TORCH_CHECK(handle.size() == 64, "bad ipc handle");
  void* ptr = nullptr;
  TORCH_CHECK(pg_symm_open(handle.data(), &ptr) == 0, "symm_open failed");
  return reinterpret_cast<int64_t>(ptr);
}
void symm_close(int64_t ptr) { pg_symm_close(reinterpret_cast<void*>(ptr)); }
void symm_free(int64_t ptr) { pg_symm_free(reinterpret_cast<void*>(ptr)); }

torch::Tensor tensor_from_ptr(int64_t ptr, int64_t nbytes, int64_t device_index) {
  auto opts = torch::TensorOptions().dtype(torch::kUInt8).device(torch::kCUDA, (int)device_index);
  return torch::from_blob(reinterpret_cast<void*>(ptr), {nbytes}, [](void*) {}, opts);
}

void rs_reduce(int64_t staging_ptr, int64_t num_src, int64_t src_stride, int64_t ctr_ptr, int64_t expected,
               const c10::optional<torch::Tensor>& bias, const c10::optional<torch::Tensor>& residual, torch::Tensor out) {
  PG_CUDA(out); PG_BF16(out);
  c10::cuda::CUDAGuard guard(out.device());
  TORCH_CHECK(pg_rs_reduce(reinterpret_cast<const void*>(staging_ptr), (int)num_src, src_stride,
                           reinterpret_cast<const uint32_t*>(ctr_ptr), (uint32_t)expected, opt_ptr(bias), opt_ptr(residual),
                           out.data_ptr(), (int)out.size(0), (int)out.size(1), cur_stream()) == 0, "rs_reduce failed");
}

void allreduce_f32(std::vector<int64_t> peer_bufs, int64_t rank, int64_t offset, int64_t n, double scale, bool rs_only,
                   std::vector<int64_t> peer_flags, int64_t epoch, int64_t blocks, int64_t mc_buf, int64_t bucket_elems) {
  float* bufs[PG_MAX_PEERS]; uint32_t* flags[PG_MAX_PEERS];
  const int world = (int)peer_bufs.size();
  TORCH_CHECK(world <= PG_MAX_PEERS && peer_flags.size() == peer_bufs.size(), "bad peer lists");
  for (int i = 0; i < world; ++i) { bufs[i] = reinterpret_cast<float*>(peer_bufs[i]); flags[i] = reinterpret_cast<uint32_t*>(peer_flags[i]); }
  TORCH_CHECK(pg_allreduce_f32(bufs, reinterpret_cast<float*>(mc_buf), world, (int)rank, offset, n, bucket_elems, (float)scale, rs_only, flags,
                               (uint32_t)epoch, (int)blocks, cur_stream()) == 0, "allreduce_f32 failed");
}

void allgather_bf16(std::vector<int64_t> peer_bufs, int64_t rank, int64_t bucket_elems, int64_t total_elems,
                    std::vector<int64_t> peer_flags, int64_t epoch, int64_t head_elems, int64_t mc_buf) {
  void* bufs[PG_MAX_PEERS]; uint32_t* flags[PG_MAX_PEERS];
  const int world = (int)peer_bufs.size();
  TORCH_CHECK(world <= PG_MAX_PEERS && peer_flags.size() == peer_bufs.size(), "bad peer lists");
  for (int i = 0; i < world; ++i) { bufs[i] = reinterpret_cast<void*>(peer_bufs[i]); flags[i] = reinterpret_cast<uint32_t*>(peer_flags[i]); }
  TORCH_CHECK(pg_allgather_bf16(bufs, reinterpret_cast<void*>(mc_buf), world, (int)rank, head_elems, bucket_elems, total_elems, flags,
                                (uint32_t)epoch, cur_stream()) == 0, "allgather_bf16 failed");
}

void rs_push(const torch::Tensor& x, int64_t num_chunks, int64_t first_chunk, std::vector<int64_t> out_peer,
             std::vector<int64_t> arrive_ctr, int64_t blocks) {
  PG_CUDA(x); PG_BF16(x);
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK((int64_t)out_peer.size() == num_chunks && (int64_t)arrive_ctr.size() == num_chunks && x.numel() % num_chunks == 0, "rs_push: bad lists");
  void* outs[PG_MAX_PEERS]; uint32_t* ctrs[PG_MAX_PEERS];
  for (int i = 0; i < num_chunks; ++i) { outs[i] = reinterpret_cast<void*>(out_peer[i]); ctrs[i] = reinterpret_cast<uint32_t*>(arrive_ctr[i]); }
  TORCH_CHECK(pg_rs_push(x.data_ptr(), (int)num_chunks, (int)first_chunk, x.numel() / num_chunks, outs, ctrs, (int)blocks, cur_stream()) == 0, "rs_push failed");
}

void multimem_selftest(int64_t mc_in, int64_t mc_out, torch::Tensor out) {
  PG_CUDA(out); PG_F32(out);
  c10::cuda::CUDAGuard guard(out.device());
  TORCH_CHECK(pg_multimem_selftest(reinterpret_cast<const float*>(mc_in), reinterpret_cast<float*>(mc_out), out.data_ptr<float>(),
                                   out.numel(), cur_stream()) == 0, "multimem_selftest failed");
}

void rs_reduce_mc(int64_t mc_partial, int64_t ctr_ptr, int64_t num_src, int64_t expected, const c10::optional<torch::Tensor>& bias,
                  const c10::optional<torch::Tensor>& residual, torch::Tensor out) {
  PG_CUDA(out); PG_BF16(out);
  c10::cuda::CUDAGuard guard(out.device());
  TORCH_CHECK(pg_rs_reduce_mc(reinterpret_cast<const void*>(mc_partial), reinterpret_cast<const uint32_t*>(ctr_ptr), (int)num_src,
                              (uint32_t)expected, opt_ptr(bias), opt_ptr(residual), out.data_ptr(), (int)out.size(0), (int)out.size(1),
                              cur_stream()) == 0, "rs_reduce_mc failed");
}

// ---- VMM + multicast symmetric memory (symm_vmm.cu)
py::tuple vmm_probe(int64_t world) {
  int mc = 0; int64_t gran = 0;
  const int rc = pg_vmm_probe((int)world, &mc, &gran);
  return py::make_tuple(rc == 0, mc != 0, gran);
}
py::tuple vmm_alloc(int64_t nbytes) {
  void* ptr = nullptr; int fd = -1; uint64_t h = 0;
  TORCH_CHECK(pg_vmm_alloc(nbytes, &ptr, &fd, &h) == 0, "vmm_alloc failed");
  return py::make_tuple(reinterpret_cast<int64_t>(ptr), fd, (int64_t)h);
}
py::tuple vmm_import(int64_t fd, int64_t nbytes) {
  void* ptr = nullptr; uint64_t h = 0;
  TORCH_CHECK(pg_vmm_import((int)fd, nbytes, &ptr, &h) == 0, "vmm_import failed");
  return py::make_tuple(reinterpret_cast<int64_t>(ptr), (int64_t)h);
}
void vmm_unmap(int64_t ptr, int64_t nbytes, int64_t handle) { pg_vmm_unmap(reinterpret_cast<void*>(ptr), nbytes, (uint64_t)handle); }
py::tuple mc_create(int64_t world, int64_t nbytes) {
  int fd = -1; uint64_t h = 0;
  TORCH_CHECK(pg_mc_create((int)world, nbytes, &fd, &h) == 0, "mc_create failed");
  return py::make_tuple(fd, (int64_t)h);
}
int64_t mc_import(int64_t fd) {
  uint64_t h = 0;
  TORCH_CHECK(pg_mc_import((int)fd, &h) == 0, "mc_import failed");
  return (int64_t)h;
}
void mc_add_device(int64_t mc) { TORCH_CHECK(pg_mc_add_device((uint64_t)mc) == 0, "mc_add_device failed"); }
int64_t mc_bind(int64_t mc, int64_t mem, int64_t nbytes) {
  void* p = nullptr;
  TORCH_CHECK(pg_mc_bind((uint64_t)mc, (uint64_t)mem, nbytes, &p) == 0, "mc_bind failed");
  return reinterpret_cast<int64_t>(p);
}

void barrier_peers(std::vector<int64_t> peer_flags, int64_t rank, int64_t epoch) {
  uint32_t* flags[PG_MAX_PEERS];
  const int world = (int)peer_flags.size();
  for (int i = 0; i < world; ++i) flags[i] = reinterpret_cast<uint32_t*>(peer_flags[i]);
  TORCH_CHECK(pg_barrier_peers(flags, world, (int)rank, (uint32_t)epoch, cur_stream()) == 0, "barrier_peers failed");
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("gemm", &gemm, "tcgen05 bf16 GEMM with fused epilogue / chunked collectives",
        py::arg("a"), py::arg("b"), py::arg("out"), py::arg("a_mn") = false, py::arg("b_mn") = false,
        py::arg("bias") = py::none(), py::arg("residual") = py::none(), py::arg("aux") = py::none(),
        py::arg("flags") = 0, py::arg("block_n") = 0, py::arg("max_ctas") = 0,
        py::arg("num_chunks") = 1, py::arg("first_chunk") = 0, py::arg("chunk_flags_ptr") = 0,
        py::arg("flag_value") = 0, py::arg("out_peer_ptrs") = std::vector<int64_t>{},
        py::arg("arrive_ctr_ptrs") = std::vector<int64_t>{}, py::arg("ag") = py::dict());
  m.def("layernorm_fwd", &layernorm_fwd);
  m.def("layernorm_bwd", &layernorm_bwd);
  m.def("colsum", &colsum);
  m.def("embedding_bwd", &embedding_bwd, py::arg("dx"), py::arg("ids"), py::arg("dw"), py::arg("vocab_start"), py::arg("vocab_end"),
        py::arg("grad_rs") = py::dict());
  m.def("grad_rs_accum", &grad_rs_accum);
  m.def("ce_stats", &ce_stats);
  m.def("ce_combine", &ce_combine);
  m.def("ce_finalize", &ce_finalize);
  m.def("adam_step", &adam_step, py::arg("master"), py::arg("m"), py::arg("v"), py::arg("grad"), py::arg("param_bf16"), py::arg("lr"),
        py::arg("beta1"), py::arg("beta2"), py::arg("eps"), py::arg("wd"), py::arg("step"), py::arg("grad_scale"), py::arg("adamw"),
        py::arg("zero_grad") = false);
  m.def("adam_step_strided", &adam_step_strided);
  m.def("sgd_step", &sgd_step);
  m.def("accum_bf16_to_f32", &accum_bf16_to_f32);
  m.def("attention_fwd", &attention_fwd, py::arg("qkv"), py::arg("slopes"), py::arg("out"), py::arg("lse"), py::arg("B"),
        py::arg("S"), py::arg("H"), py::arg("D"), py::arg("softmax_scale") = 0.0);
  m.def("attention_bwd", &attention_bwd, py::arg("qkv"), py::arg("slopes"), py::arg("out"), py::arg("lse"), py::arg("dout"),
        py::arg("dqkv"), py::arg("B"), py::arg("S"), py::arg("H"), py::arg("D"), py::arg("softmax_scale") = 0.0);
  m.def("moe_route", &moe_route);
  m.def("moe_dispatch", &moe_dispatch);
  m.def("symm_alloc", &symm_alloc);
  m.def("symm_open", &symm_open);
  m.def("symm_close", &symm_close);
  m.def("symm_free", &symm_free);
  m.def("tensor_from_ptr", &tensor_from_ptr);
  m.def("rs_reduce", &rs_reduce);
  m.def("allreduce_f32", &allreduce_f32, py::arg("peer_bufs"), py::arg("rank"), py::arg("offset"), py::arg("n"),
        py::arg("scale"), py::arg("rs_only"), py::arg("peer_flags"), py::arg("epoch"), py::arg("blocks") = 0, py::arg("mc_buf") = 0, py::arg("bucket_elems") = 0);
  m.def("set_gemm_cta_cap", [](int64_t n) { pg_set_gemm_cta_cap((int)n); });
  m.def("allgather_bf16", &allgather_bf16, py::arg("peer_bufs"), py::arg("rank"), py::arg("bucket_elems"), py::arg("total_elems"),
        py::arg("peer_flags"), py::arg("epoch"), py::arg("head_elems") = 0, py::arg("mc_buf") = 0);
  m.def("rs_push", &rs_push);
  m.def("multimem_selftest", &multimem_selftest);
  m.def("rs_reduce_mc", &rs_reduce_mc);
  m.def("vmm_probe", &vmm_probe);
  m.def("vmm_alloc", &vmm_alloc);
  m.def("vmm_import", &vmm_import);
  m.def("vmm_unmap", &vmm_unmap);
  m.def("mc_create", &mc_create);
  m.def("mc_import", &mc_import);
  m.def("mc_add_device", &mc_add_device);
  m.def("mc_bind", &mc_bind);
  m.def("barrier_peers", &barrier_peers);
}
