// Host side of the tcgen05 GEMM: TMA tensor-map construction (driver entry point resolved at
// run time so the library links without libcuda), tile-shape selection, launch.
#include "gemm_sm100.cuh"
#include "launch.h"
#include "pdl_launch.cuh"

#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace pg {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || p == nullptr) {
      fprintf(stderr, "pipegoose_b200: cuTensorMapEncodeTiled not available (%s)\n",
              cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major [rows, cols] matrix (cols contiguous, leading
// dimension ld elements) with a [box_rows, 64] box and 128-byte swizzle.
int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols,
                      uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -1;
  // driver entry points need the primary context bound to the calling thread (autograd worker
  // threads have not necessarily touched the runtime yet)
  cudaFree(nullptr);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr,
            "pipegoose_b200: cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%llu cols=%llu ld=%llu "
            "box=%ux%u\n",
            (int)r, ptr, (unsigned long long)rows, (unsigned long long)cols,
            (unsigned long long)ld, box_rows, box_cols);
    return -1;
  }
  return 0;
}

struct TmapKey {
  const void* ptr;
  uint64_t rows, cols, ld;
  uint32_t box_rows;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= k.rows * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= k.cols * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (k.ld * 31 + k.box_rows) + (h << 6) + (h >> 2);
    return h;
  }
};

static int cached_tmap(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols,
                       uint64_t ld, uint32_t box_rows) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  TmapKey key{ptr, rows, cols, ld, box_rows};
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return 0;
  }
  if (make_tmap_bf16_2d(out, ptr, rows, cols, ld, box_rows, 64) != 0) return -1;
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, bool A_MN, bool B_MN, bool CTA2>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tal,
                      const GemmArgs& args, int max_ctas, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CTA2>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, CTA2>;
  constexpr int BM_T = CTA2 ? 2 * BM : BM;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      fprintf(stderr, "pipegoose_b200: cudaFuncSetAttribute failed: %s\n", cudaGetErrorString(e));
      return -1;
    }
    attr_set = true;
  }
  const int chunk_rows = args.num_chunks > 1 ? args.chunk_rows : args.M;
  const int tiles = ((chunk_rows + BM_T - 1) / BM_T) * ((args.N + BN - 1) / BN) * args.num_chunks *
                    (args.k_splits > 1 ? args.k_splits : 1);
  int units = (max_ctas - args.n_comm) / (CTA2 ? 2 : 1);  // CTAs or CTA pairs that run GEMM tiles
  if (units < 1) units = 1;
  if (tiles < units) units = tiles;
  if (units < 1) units = 1;
  const int grid = units * (CTA2 ? 2 : 1) + args.n_comm;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(kGemmThreads, 1, 1);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTA2 ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tal, args);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "pipegoose_b200: gemm launch failed: %s\n", cudaGetErrorString(e));
    return -1;
  }
  return 0;
}

// Tile width (and, for fp32-accumulating outputs, a K split) by a wave-quantisation cost model:
// cost = waves * (k-blocks per work item * BN / efficiency(BN)) + one exposed epilogue.
// `cta2`: in -1 never / 0 auto / 1 always use CTA pairs (cta_group::2, 256-row tiles); out: the choice.
// Measured on B200 (profiles/gemm_check_r1_v3_cta_pair.json): pairs win where the L2->SM operand feed is the
// limit (long K, or many tiles per SM: +5..+16%), and lose a few percent on short problems (coarser tiles,
// deeper prologue) and on the fp32 read-modify-write epilogue of the small wgrads.
static int pick_bn(int M_rows, int N, int K, int chunks, int ctas, bool allow_split, bool b_mn, bool out_f32,
                   int* k_splits, int* cta2) {
  const int cand[4] = {256, 192, 128, 64};
  const float eff1[4] = {1.0f, 0.93f, 0.82f, 0.55f};   // 1-CTA tiles 128 x BN
  const float eff2[4] = {1.0f, 0.80f, 0.70f, 0.0f};    // CTA-pair tiles 256 x BN (relative to each other)
  const int num_kb = (K + BK - 1) / BK;
  int mode = *cta2;
  if (mode == 0) {
    const long tiles1 = (long)((M_rows + BM - 1) / BM) * ((N + 255) / 256) * chunks;
    const bool pair = out_f32 ? (tiles1 >= 6L * ctas) : (K >= 2048 || tiles1 >= 3L * ctas);
    mode = (pair && M_rows >= 2 * BM) ? 1 : -1;
  }
  if (mode > 0 && (ctas < 2 || (chunks > 1 && (M_rows % (2 * BM)) != 0))) mode = -1;
  const int pair = mode > 0 ? 1 : 0;
  const int bm = pair ? 2 * BM : BM;
  const int units = pair ? ctas / 2 : ctas;
  int best = 256;
  float best_cost = 1e30f;
  *k_splits = 1;
  *cta2 = pair;
  for (int i = 0; i < 4; ++i) {
    const int bn = cand[i];
    const float eff = pair ? eff2[i] : eff1[i];
    if (eff <= 0.f) continue;
    if (pair && b_mn && (bn / 2) % 64 != 0) continue;  // MN-major B is staged in 64-column atoms
    if (bn > 64 && N <= bn / 2) continue;
    const long tiles = (long)((M_rows + bm - 1) / bm) * ((N + bn - 1) / bn) * chunks;
    const int max_split = allow_split ? 8 : 1;
    for (int s = 1; s <= max_split; ++s) {
      const int per = (num_kb + s - 1) / s;
      if (s > 1 && (per < 16 || (long)(s - 1) * per >= num_kb)) break;
      const long waves = (tiles * s + units - 1) / units;
      // the epilogue of the last wave is exposed; atomics cost about twice a plain store
      const float epi = (s > 1 ? 16.0f : 8.0f) * (float)bn;
      const float cost = (float)waves * (float)per * (float)bn / eff + epi;
      if (cost < best_cost * 0.97f) {
        best_cost = cost;
        best = bn;
        *k_splits = s;
      }
    }
  }
  return best;
}

}  // namespace pg

using namespace pg;

static int g_cta_cap = 0;
extern "C" void pg_set_gemm_cta_cap(int ctas) { g_cta_cap = ctas > 0 ? ctas : 0; }

extern "C" int pg_gemm_bf16(const PgGemmDesc* d, cudaStream_t stream) {
  if (d->M <= 0 || d->N <= 0 || d->K <= 0) return 0;
  GemmArgs args;
  memset(&args, 0, sizeof(args));
  args.M = d->M;
  args.N = d->N;
  args.K = d->K;
  args.out = d->out;
  args.ldc = d->ldc;
  args.bias = reinterpret_cast<const __nv_bfloat16*>(d->bias);
  args.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual);
  args.ldr = d->ldr;
  args.aux = d->aux;
  args.flags = d->flags;
  args.num_chunks = d->num_chunks > 1 ? d->num_chunks : 1;
  args.chunk_rows = d->chunk_rows;
  args.first_chunk = d->first_chunk;
  args.chunk_flags = d->chunk_flags;
  args.flag_value = d->flag_value;
  for (int i = 0; i < kMaxPeers; ++i) {
    args.out_peer[i] = d->out_peer[i];
    args.arrive_ctr[i] = d->arrive_ctr[i];
    args.rs_in[i] = reinterpret_cast<const __nv_bfloat16*>(d->rs_in[i]);
  }
  args.rs_wait_ctr = d->rs_wait_ctr;
  args.rs_wait_value = d->rs_wait_value;
  args.ce_part = d->ce_part;
  args.ce_valid = d->ce_valid > 0 ? d->ce_valid : d->N;
  args.grad_rs = d->grad_rs;
  if (args.grad_rs.world > 1) {
    if (!(d->flags & EPI_OUT_F32) || args.num_chunks > 1 || (d->ldc % 4) != 0) {
      fprintf(stderr, "pipegoose_b200: gradient reduce-scatter epilogue needs a plain fp32 output with ldc %% 4 == 0\n");
      return -1;
    }
    args.flags |= EPI_ACCUM;  // contributions are ADDED into the owners' buffers (zeroed by the optimizer step)
  }
  args.b_chunk_rows = d->b_chunk_rows;
  args.bias_chunk_stride = d->bias_chunk_stride;
  args.row_ret = d->row_ret;
  args.row_scale = d->row_scale;
  args.scatter_rows_per_src = d->scatter_rows_per_src > 0 ? d->scatter_rows_per_src : 1;
  args.n_comm = d->n_comm;
  args.ag_dst = d->ag_dst;
  args.ag_chunk_bytes = d->ag_chunk_bytes;
  args.ag_ready = d->ag_ready;
  args.ag_epoch = d->ag_epoch;
  args.my_rank = d->my_rank;
  for (int i = 0; i < kMaxPeers; ++i) {
    args.ag_src[i] = d->ag_src[i];
    args.ag_peer_flag[i] = d->ag_peer_flag[i];
  }
  if (args.num_chunks > 1 && (args.chunk_rows % BM) != 0) {
    fprintf(stderr, "pipegoose_b200: chunk_rows (%d) must be a multiple of %d\n", args.chunk_rows, BM);
    return -1;
  }
  int max_ctas = d->max_ctas > 0 ? d->max_ctas : num_sms();
  if (d->max_ctas <= 0 && g_cta_cap > 0 && g_cta_cap < max_ctas) max_ctas = g_cta_cap;
  const int chunk_rows = args.num_chunks > 1 ? args.chunk_rows : args.M;
  if ((args.flags & EPI_DGELU) && (args.flags & EPI_RESIDUAL)) {
    fprintf(stderr, "pipegoose_b200: EPI_DGELU and EPI_RESIDUAL are mutually exclusive\n");
    return -1;
  }
  // split-K only where the epilogue can add partial sums: plain fp32 outputs (wgrad into main grads)
  const bool allow_split = (args.flags & EPI_OUT_F32) && args.num_chunks == 1 && args.bias == nullptr &&
                           d->k_splits != 1 && (d->ldc % 4) == 0;
  int auto_splits = 1;
  int cta2 = d->cta_pair;
  if (args.n_comm & 1) cta2 = -1;  // comm CTAs must fill whole clusters
  int bn = pick_bn(chunk_rows, args.N, args.K, args.num_chunks, max_ctas - args.n_comm, allow_split, d->b_mn != 0,
                   (args.flags & EPI_OUT_F32) != 0, &auto_splits, &cta2);
  if (args.ce_part != nullptr && d->block_n <= 0) {
    // the partials are laid out per 256-column tile (two per row and tile): fix the tile width, keep the pair choice
    bn = 256;
    auto_splits = 1;
  }
  if (d->block_n > 0) {
    bn = d->block_n;
    auto_splits = 1;
    cta2 = d->cta_pair > 0 ? 1 : 0;
    if (cta2 && (bn == 64 || (d->b_mn && (bn / 2) % 64 != 0) || (args.n_comm & 1) ||
                 (args.num_chunks > 1 && args.chunk_rows % (2 * BM) != 0))) {
      fprintf(stderr, "pipegoose_b200: block_n %d cannot run as a CTA pair here\n", bn);
      return -1;
    }
  }
  args.k_splits = (allow_split && d->k_splits > 1) ? d->k_splits : auto_splits;
  {
    const int num_kb = (args.K + BK - 1) / BK;
    if (args.k_splits > num_kb) args.k_splits = num_kb;
    if (args.k_splits > 1) {
      const int per = (num_kb + args.k_splits - 1) / args.k_splits;
      args.k_splits = (num_kb + per - 1) / per;  // no empty split
    }
  }
  if (args.k_splits > 1) {
    if (!(args.flags & EPI_ACCUM)) {
      // partial sums are added atomically: an overwriting GEMM starts from zero
      cudaMemset2DAsync(args.out, (size_t)args.ldc * 4, 0, (size_t)args.N * 4, (size_t)args.M, stream);
    }
    args.flags |= EPI_ATOMIC | EPI_ACCUM;
  }

  CUtensorMap ta, tb, tal;
  args.a_local_chunk = -1;
  // A: K-major  -> matrix [M, K] (ld = lda), box [128 rows, 64]
  //    MN-major -> matrix [K, M] (ld = lda), box [64 k-rows, 64]
  if (!d->a_mn) {
    if (cached_tmap(&ta, d->A, d->M, d->K, d->lda, BM) != 0) return -1;
  } else {
    if (cached_tmap(&ta, d->A, d->K, d->M, d->lda, BK) != 0) return -1;
  }
  const uint64_t b_groups = d->b_chunk_rows > 0 ? (uint64_t)args.num_chunks : 1;  // stacked expert weights
  if (!d->b_mn) {
    if (cached_tmap(&tb, d->B, d->N * b_groups, d->K, d->ldb, cta2 ? bn / 2 : bn) != 0) return -1;
  } else {
    if (cached_tmap(&tb, d->B, d->K * b_groups, d->N, d->ldb, BK) != 0) return -1;
  }

  tal = ta;
  if (d->a_local != nullptr && !d->a_mn && args.num_chunks > 1) {
    // the local shard of an all-gather -> GEMM is read in place: [chunk_rows, K] matrix
    if (cached_tmap(&tal, d->a_local, args.chunk_rows, d->K, d->K, BM) != 0) return -1;
    args.a_local_chunk = d->my_rank;
  }

#define PG_DISPATCH_BN(AMN, BMN)                                                               \
  if (cta2) {                                                                                  \
    switch (bn) {                                                                              \
      case 256: return launch_cfg<256, AMN, BMN, true>(ta, tb, tal, args, max_ctas, stream);    \
      case 192:                                                                                \
        if constexpr (!BMN) return launch_cfg<192, AMN, BMN, true>(ta, tb, tal, args, max_ctas, stream); \
        break;                                                                                 \
      case 128: return launch_cfg<128, AMN, BMN, true>(ta, tb, tal, args, max_ctas, stream);    \
      default: break;                                                                          \
    }                                                                                          \
    fprintf(stderr, "pipegoose_b200: bad CTA-pair block_n %d\n", bn);                          \
    return -1;                                                                                 \
  }                                                                                            \
  switch (bn) {                                                                                \
    case 256: return launch_cfg<256, AMN, BMN, false>(ta, tb, tal, args, max_ctas, stream);     \
    case 192: return launch_cfg<192, AMN, BMN, false>(ta, tb, tal, args, max_ctas, stream);     \
    case 128: return launch_cfg<128, AMN, BMN, false>(ta, tb, tal, args, max_ctas, stream);     \
    case 64: return launch_cfg<64, AMN, BMN, false>(ta, tb, tal, args, max_ctas, stream);       \
    default: fprintf(stderr, "pipegoose_b200: bad block_n %d\n", bn); return -1;               \
  }
  if (!d->a_mn && !d->b_mn) {
    PG_DISPATCH_BN(false, false)
  } else if (!d->a_mn && d->b_mn) {
    PG_DISPATCH_BN(false, true)
  } else if (d->a_mn && d->b_mn) {
    PG_DISPATCH_BN(true, true)
  } else {
    PG_DISPATCH_BN(true, false)
  }
#undef PG_DISPATCH_BN
}
