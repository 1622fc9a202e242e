// Persistent, warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor) -> 128B-swizzled smem ring -> tcgen05.mma (one issuing
//   thread) -> fp32 accumulators double-buffered in TMEM -> tcgen05.ld epilogue.
//
// One kernel serves the three operand layouts a transformer training step needs
//   NT : Y  = X  · Wᵀ   (A K-major,  B K-major)    forward
//   NN : dX = dY · W    (A K-major,  B MN-major)   dgrad
//   TN : dW = dYᵀ · X   (A MN-major, B MN-major)   wgrad
// and the fused compute+collective forms used by tensor parallelism
//   all-gather -> GEMM : M is split into `num_chunks` row chunks that become readable
//                        at different times (a comm CTA group or a peer publishes a
//                        flag per chunk); tiles are visited chunk by chunk, starting at
//                        the local chunk, and the TMA producer acquires the chunk flag.
//   GEMM -> reduce-scatter : the epilogue stores each tile straight into the owner
//                        rank's staging slot through its NVLink peer mapping and bumps a
//                        per-owner arrival counter with a system-scope release.
// The epilogue applies bias / tanh-GELU (also emitting the pre-activation) /
// GELU-backward / residual add / fp32 accumulate, so none of those is a separate pass.
#pragma once
#include "grad_rs.cuh"
#include "ptx.cuh"

namespace pg {

constexpr int kMaxPeers = 8;

enum EpiFlags : int {
  EPI_BIAS = 1,        // acc += bias[col]
  EPI_GELU = 2,        // out = gelu_tanh(acc); if aux_out != null, aux_out = acc (pre-activation)
  EPI_RESIDUAL = 4,    // out = acc + residual[row, col]
  EPI_OUT_F32 = 8,     // out is fp32
  EPI_ACCUM = 16,      // out += acc (fp32 out only)
  EPI_DGELU = 32,      // out = acc * gelu_tanh'(aux_in[row, col])
  EPI_SCATTER = 64,    // MoE combine: out row -> out_peer[src][row_ret[row]], scaled by row_scale[row]
  EPI_ATOMIC = 128,    // fp32 out += acc with red.global.add (split-K partial sums)
};
// (GemmArgs::ce_part != null) lm_head: besides the bf16 logits the epilogue emits, per row and per half tile, the online
// softmax partials (max, sum exp(x - max)) of the logits it holds in registers — the cross entropy never re-reads the
// [tokens, vocab] logits for its statistics (csrc/elementwise.cu: ce_combine merges the partials)

struct GemmArgs {
  int M, N, K;
  int k_splits;  // >1: the K loop is split over k_splits work items per tile (EPI_ATOMIC epilogue)
  void* out;
  int ldc;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* residual;
  int ldr;
  void* aux;  // EPI_GELU: pre-activation output; EPI_DGELU: pre-activation input (bf16, ld = ldc)
  int flags;
  // ---- chunked traversal (fused collectives); num_chunks == 1 for a plain GEMM
  int num_chunks;
  int chunk_rows;   // rows per chunk (multiple of 128 when num_chunks > 1)
  int first_chunk;  // chunk visited first (the local one)
  // all-gather -> GEMM: chunk c is readable once chunk_flags[c] >= flag_value
  const uint32_t* chunk_flags;
  uint32_t flag_value;
  // comm CTAs (blockIdx.x < n_comm) pull every rank's shard into the local gathered A matrix with
  // bulk async copies (peer global -> smem -> local global) and publish chunk_flags (+1 per CTA)
  int n_comm;
  const void* ag_src[kMaxPeers];     // rank r's shard (NVLink peer mapping; ag_src[my_rank] is local)
  void* ag_dst;                      // local gathered matrix [num_chunks * chunk_rows, K]
  uint64_t ag_chunk_bytes;           // bytes of one shard
  uint32_t* ag_peer_flag[kMaxPeers]; // &peer_r.shard_ready[my_rank]: tell rank r that my shard is readable
  const uint32_t* ag_ready;          // local shard_ready[src] (written by src), wait >= ag_epoch
  uint32_t ag_epoch;
  int my_rank;
  int a_local_chunk;  // >= 0: tiles of this chunk load A through tma_a_local (rows relative to the chunk)
  // grouped GEMM (one chunk per expert): B / bias of chunk c start b_chunk_rows*c rows / bias_chunk_stride*c
  // elements further (stacked expert weights)
  int b_chunk_rows;
  int bias_chunk_stride;
  // EPI_SCATTER: row r (inside its chunk) came from rank r / scatter_rows_per_src; it returns to row
  // row_ret[r] of that rank's buffer out_peer[rank] (row_ret < 0: empty slot), scaled by row_scale[r]
  const int* row_ret;
  const float* row_scale;
  int scatter_rows_per_src;
  // GEMM -> reduce-scatter: rows of chunk c go to out_peer[c] (row index relative to chunk),
  // then arrive_ctr[c] (+1 per finished tile, release.sys)
  void* out_peer[kMaxPeers];
  uint32_t* arrive_ctr[kMaxPeers];
  // fused reduction of the LOCAL chunk (visited last): once every peer's arrival counter reached rs_wait_value,
  // the epilogue of this rank's own tiles adds the peers' staged partial tiles (+ bias + residual) and writes the
  // final rows to `out`; no separate reduce kernel, the own partial never goes through memory
  const __nv_bfloat16* rs_in[kMaxPeers];  // local staging slot of source s (rs_in[my_rank] unused); null = off
  const uint32_t* rs_wait_ctr;            // local arrival counters, one per source
  uint32_t rs_wait_value;
  float* ce_part;  // [M, 2 * n_blks, 2] fp32 (max, sumexp) partials; slot 2 * n_blk + half (null: off)
  int ce_valid;    // logits columns >= ce_valid are vocabulary padding: excluded from the statistics
  // wgrad -> data-parallel reduce-scatter (fp32 outputs): every 16-byte group of the tile is added into the gradient
  // buffer of the rank that owns its ZeRO-1 slice (local L2 atomics / NVLink peer atomics); grad_rs.world <= 1: off
  PgGradRS grad_rs;
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 384;  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue

// CTA2: the two CTAs of a 2-CTA cluster (one TPC) compute ONE 256 x BN tile with tcgen05.mma.cta_group::2:
// each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 rows), the leader CTA's MMA thread
// issues the pair-wide instruction, each CTA's TMEM receives its 128 accumulator rows.  Per SM that is
// 32 KB of operands per 128x256x64 MACs instead of 48 KB: the L2->SM feed is what bounds the 1-CTA kernel.
template <int BN, bool CTA2 = false>
struct GemmCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBRows = CTA2 ? BN / 2 : BN;  // rows (N) of B this CTA stages
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kMaxStages = (208 * 1024) / kStageBytes;
  static constexpr int kStages = kMaxStages > 8 ? 8 : kMaxStages;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kEpiStageBytes = 8 * 32 * 64;  // 8 epilogue warps x [32 rows x 64 B]
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

PG_DEVICE float gelu_tanh(float x) {
  // Bloom's GELU: x * 0.5 * (1 + tanh(0.79788456 x (1 + 0.044715 x^2)))
  float u = 0.79788456f * x * (1.0f + 0.044715f * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * (1.0f + t);
}
PG_DEVICE float gelu_tanh_grad(float x) {
  float u = 0.79788456f * x * (1.0f + 0.044715f * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * ((1.0f - t * t) * (0.79788456f + 0.1070322243f * x * x)) + 0.5f * (1.0f + t);
}

struct TileCoord {
  int m_blk, n_blk, chunk;
};

// Tiles are enumerated chunk-major (chunk order rotated to start at first_chunk); inside a
// chunk, groups of 8 row-blocks sweep all column-blocks so that concurrently resident CTAs
// share A and B tiles through L2.
PG_DEVICE TileCoord map_tile(int t, int m_blks_per_chunk, int n_blks, int num_chunks,
                             int first_chunk) {
  const int per_chunk = m_blks_per_chunk * n_blks;
  const int ci = t / per_chunk;
  int r = t - ci * per_chunk;
  int chunk = first_chunk + ci;
  if (chunk >= num_chunks) chunk -= num_chunks;
  constexpr int G = 8;
  const int per_group = G * n_blks;
  const int g = r / per_group;
  const int first_m = g * G;
  const int gsz = min(m_blks_per_chunk - first_m, G);
  r -= g * per_group;
  TileCoord c;
  c.n_blk = r / gsz;
  c.m_blk = chunk * m_blks_per_chunk + first_m + (r - c.n_blk * gsz);
  c.chunk = chunk;
  return c;
}

// Launch-time register budget: 384 threads x 128 registers = 48 K of the SM's 64 K — the rest of the register file (and
// ~15 KB of shared memory) stays free, so that a small communication CTA (the co-resident gradient reducer,
// csrc/comm.cu) can run NEXT TO a persistent GEMM CTA instead of waiting for a kernel boundary.  The roles then
// re-balance: warps 0-3 (TMA producer, MMA issuer, TMEM allocator) drop to 40 registers, the two epilogue warpgroups
// rise to 168 — 128*40 + 256*168 = 48 K, i.e. the epilogue keeps the budget it had when the CTA owned the whole file.
#ifndef PG_GEMM_REBALANCE
#define PG_GEMM_REBALANCE 1   // 0: the round-1 budget (384 x 168 registers, the CTA owns the SM's register file)
#endif
constexpr int kGemmLaunchRegs = 128;
constexpr int kGemmLightRegs = 40;
constexpr int kGemmEpilogueRegs = 168;

template <int BN, bool A_MN, bool B_MN, bool CTA2>
__global__ void
#if PG_GEMM_REBALANCE
__launch_bounds__(512, 1)  /* 65536 / 512 = 128 registers at launch; the block has 384 threads */
#else
__launch_bounds__(kGemmThreads, 1)
#endif
    gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a,
                     const __grid_constant__ CUtensorMap tma_b,
                     const __grid_constant__ CUtensorMap tma_a_local, const GemmArgs args) {
  using Cfg = GemmCfg<BN, CTA2>;
  constexpr int BM_T = CTA2 ? 2 * BM : BM;  // rows of one work tile (CTA pair: 256)
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::kABytes;
  uint8_t* smem_epi = smem + STAGES * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::kEpiStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) pdl_launch_dependents();  // the next kernel may start its prologue under this one's tail

  const int chunk_rows = (args.num_chunks > 1) ? args.chunk_rows : args.M;
  const int m_blks_per_chunk = (chunk_rows + BM_T - 1) / BM_T;
  // work distribution unit: a CTA (1-CTA kernel) or a CTA pair; both CTAs of a pair walk the same tiles
  uint32_t cta_rank = 0;
  if constexpr (CTA2) cta_rank = cluster_ctarank();
  const int unit0 = CTA2 ? (static_cast<int>(blockIdx.x) - args.n_comm) / 2 : static_cast<int>(blockIdx.x) - args.n_comm;
  const int unit_stride = CTA2 ? (static_cast<int>(gridDim.x) - args.n_comm) / 2 : static_cast<int>(gridDim.x) - args.n_comm;
  const int row_in_tile0 = static_cast<int>(cta_rank) * BM;  // this CTA's first row inside the work tile
  const int n_blks = (args.N + BN - 1) / BN;
  const int num_tiles = m_blks_per_chunk * n_blks * args.num_chunks;
  const int num_kb = (args.K + BK - 1) / BK;
  const int k_splits = args.k_splits > 1 ? args.k_splits : 1;
  const int kb_per_split = (num_kb + k_splits - 1) / k_splits;
  const int num_work = num_tiles * k_splits;  // work item t: tile t % num_tiles, K split t / num_tiles

  if (static_cast<int>(blockIdx.x) < args.n_comm) {
    // ============================ communication CTA ============================
    // One thread drives a ring of bulk copies: peer shard -> shared memory -> local gathered
    // matrix, 16 KB per copy, 8 loads in flight per CTA (~1 MB in flight over 8 CTAs: the NVLink
    // bandwidth-delay product).
    pdl_wait();  // predecessors (the kernel that produced this rank's shard) are complete and visible
    constexpr uint32_t kSlice = 16 * 1024;
    constexpr int kCommStages = 12;  // smem ring slots
    constexpr int kLoadsInFlight = 10;  // => up to kCommStages - kLoadsInFlight stores may still be reading smem
    uint64_t* cbar = reinterpret_cast<uint64_t*>(smem + kCommStages * kSlice);
    if (threadIdx.x == 0) {
      for (int i = 0; i < kCommStages; ++i) mbar_init(&cbar[i], 1);
      fence_mbar_init();
      if (blockIdx.x == 0) {
        // my shard was produced by the previous kernel on this stream: publish it to every peer
        fence_acq_rel_sys();
        for (int p = 0; p < args.num_chunks; ++p)
          if (p != args.my_rank) st_release_sys(args.ag_peer_flag[p], args.ag_epoch);
      }
      uint32_t ph[kCommStages];
      for (int i = 0; i < kCommStages; ++i) ph[i] = 0;
      const uint64_t nslices = (args.ag_chunk_bytes + kSlice - 1) / kSlice;
      const uint64_t first = blockIdx.x, step = args.n_comm;
      const uint64_t mine = (nslices > first) ? (nslices - first + step - 1) / step : 0;
      uint64_t ring = 0;  // slices issued so far over all shards -> ring slot = ring % kCommStages
      for (int i = 1; i <= args.num_chunks; ++i) {
        // remote shards first (rank+1, rank+2, ...); the local one last: the GEMM reads it in place,
        // the copy into the gathered matrix is only needed by later consumers (wgrad)
        int src = args.my_rank + i;
        if (src >= args.num_chunks) src -= args.num_chunks;
        if (src != args.my_rank) {
          spin_until_ge(args.ag_ready + src, args.ag_epoch, 1);
        }
        const uint8_t* from = reinterpret_cast<const uint8_t*>(args.ag_src[src]);
        uint8_t* to = reinterpret_cast<uint8_t*>(args.ag_dst) + static_cast<uint64_t>(src) * args.ag_chunk_bytes;
        uint64_t issued = 0, stored = 0;
        while (stored < mine) {
          while (issued < mine && issued - stored < kLoadsInFlight) {
            const int st = static_cast<int>((ring + issued) % kCommStages);
            const uint64_t off = (first + issued * step) * kSlice;
            const uint32_t bytes = static_cast<uint32_t>(min(static_cast<uint64_t>(kSlice), args.ag_chunk_bytes - off));
            // the store that last used this slot is >= kCommStages - kLoadsInFlight stores old
            tma_store_wait_read<kCommStages - kLoadsInFlight>();
            mbar_arrive_expect_tx(&cbar[st], bytes);
            bulk_load_1d(smem + st * kSlice, from + off, bytes, &cbar[st]);
            ++issued;
          }
          const int st = static_cast<int>((ring + stored) % kCommStages);
          const uint64_t off = (first + stored * step) * kSlice;
          const uint32_t bytes = static_cast<uint32_t>(min(static_cast<uint64_t>(kSlice), args.ag_chunk_bytes - off));
          mbar_wait(&cbar[st], ph[st]);
          ph[st] ^= 1;
          bulk_store_1d(to + off, smem + st * kSlice, bytes);
          tma_store_commit();
          ++stored;
        }
        ring += mine;
        tma_store_wait<0>();  // this CTA's part of the shard is written
        fence_proxy_async_global();
        __threadfence();
        atomicAdd(const_cast<uint32_t*>(args.chunk_flags) + src, 1u);
      }
    }
    return;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], CTA2 ? 16 : 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (CTA2) {
      tmem_alloc_cta2(tmem_ptr_smem, Cfg::kTmemCols);
      tmem_relinquish_cta2();
    } else {
      tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  pdl_wait();  // everything above overlapped the previous kernel's tail; operands / epilogue inputs are read below
  const uint32_t tmem_base = *tmem_ptr_smem;

  // each register re-balancing instruction dominates the code of its warpgroup(s): ptxas allocates per region
  if (warp < 4) {
#if PG_GEMM_REBALANCE
    setmaxnreg_dec<kGemmLightRegs>();
#endif
  if (warp == 0) {
    // ============================ TMA producer ============================
    int stage = 0;
    uint32_t phase = 0;
    int seen_chunk = -1;
    for (int t = unit0; t < num_work; t += unit_stride) {
      const TileCoord tc = map_tile(t % num_tiles, m_blks_per_chunk, n_blks, args.num_chunks, args.first_chunk);
      const int kb_begin = (t / num_tiles) * kb_per_split;
      const int kb_end = min(num_kb, kb_begin + kb_per_split);
      if (args.chunk_flags != nullptr && tc.chunk != seen_chunk && tc.chunk != args.a_local_chunk) {
        if (lane == 0) {
          spin_until_ge(args.chunk_flags + tc.chunk, args.flag_value, 2);
          fence_proxy_async_global();
        }
        __syncwarp();
        seen_chunk = tc.chunk;
      }
      int m0 = tc.m_blk * BM_T + row_in_tile0;
      const int n0 = tc.n_blk * BN + static_cast<int>(cta_rank) * Cfg::kBRows;  // this CTA's part of the B tile
      // all-gather -> GEMM: the local shard is read in place from its (peer-visible) staging buffer
      const CUtensorMap* map_a = &tma_a;
      if (args.a_local_chunk >= 0 && tc.chunk == args.a_local_chunk) {
        map_a = &tma_a_local;
        m0 -= tc.chunk * chunk_rows;
      }
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem_a + stage * Cfg::kABytes;
          uint8_t* sb = smem_b + stage * Cfg::kBBytes;
          if constexpr (CTA2) {
            // both CTAs' bytes are counted on the LEADER's full barrier (the only one the MMA thread waits on)
            const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            if constexpr (!A_MN) {
              tma_load_2d_cta2(sa, map_a, bar, kb * BK, m0);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i) tma_load_2d_cta2(sa + i * (BK * 128), map_a, bar, m0 + i * 64, kb * BK);
            }
            if constexpr (!B_MN) {
              tma_load_2d_cta2(sb, &tma_b, bar, kb * BK, n0 + tc.chunk * args.b_chunk_rows);
            } else {
#pragma unroll
              for (int i = 0; i < Cfg::kBRows / 64; ++i)
                tma_load_2d_cta2(sb + i * (BK * 128), &tma_b, bar, n0 + i * 64, kb * BK + tc.chunk * args.b_chunk_rows);
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            if constexpr (!A_MN) {
              tma_load_2d(sa, map_a, &full_bar[stage], kb * BK, m0);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i)
                tma_load_2d(sa + i * (BK * 128), map_a, &full_bar[stage], m0 + i * 64, kb * BK);
            }
            if constexpr (!B_MN) {
              tma_load_2d(sb, &tma_b, &full_bar[stage], kb * BK, n0 + tc.chunk * args.b_chunk_rows);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_2d(sb + i * (BK * 128), &tma_b, &full_bar[stage], n0 + i * 64, kb * BK + tc.chunk * args.b_chunk_rows);
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ============================ MMA issuer (the pair's leader CTA only) ============================
    constexpr uint32_t idesc = make_idesc_bf16(BM_T, BN, A_MN, B_MN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = unit0; t < num_work; t += unit_stride, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      const int kb_begin = (t / num_tiles) * kb_per_split;
      const int kb_end = min(num_kb, kb_begin + kb_per_split);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem_a + stage * Cfg::kABytes);
          const uint32_t sb = smem_u32(smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: 16 elements = 32 bytes inside the 128B swizzle row; 8-row groups 1024B apart.
            // MN-major: 16 k-rows of 128B; 64-wide MN atoms BK*128 B apart (LBO); 8 k-rows 1024B (SBO).
            const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * (UMMA_K * 128), BK * 128, 1024)
                                     : make_smem_desc_sw128(sa + k * (UMMA_K * 2), 0, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * (UMMA_K * 128), BK * 128, 1024)
                                     : make_smem_desc_sw128(sb + k * (UMMA_K * 2), 0, 1024);
            if constexpr (CTA2) {
              umma_f16_cta2(tmem_d, da, db, idesc, ((kb - kb_begin) | k) != 0 ? 1u : 0u);
            } else {
              umma_f16(tmem_d, da, db, idesc, ((kb - kb_begin) | k) != 0 ? 1u : 0u);
            }
          }
          if constexpr (CTA2) {
            umma_commit_cta2(&empty_bar[stage], 3);  // frees the stage in BOTH CTAs
            if (kb == kb_end - 1) umma_commit_cta2(&tmem_full[acc], 3);
          } else {
            umma_commit(&empty_bar[stage]);
            if (kb == kb_end - 1) umma_commit(&tmem_full[acc]);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  }
  } else {
#if PG_GEMM_REBALANCE
    setmaxnreg_inc<kGemmEpilogueRegs>();
#endif
    // ============================ epilogue ============================
    // 8 warps: warp e = 4 + q + 4h drains TMEM lanes [32q, 32q+32) (a warp may only touch the lane
    // quarter warp%4) and takes every second 32-column chunk (c = h, h+2, ...).  Every global access
    // is coalesced through a per-warp [32 rows x 64 B] staging buffer (16-byte chunks XOR-swizzled by
    // row pair): 4 lanes cover one 64-byte row segment, 8 rows per instruction.  Epilogue inputs
    // (residual / GELU' pre-activation) are prefetched one chunk ahead into registers.
    const int e = warp - 4;
    const int q = e & 3;  // == warp % 4
    const int h = e >> 2;
    constexpr int NCH = BN / 32;
    const bool out_f32 = (args.flags & EPI_OUT_F32) != 0;
    const __nv_bfloat16* in_ptr = nullptr;
    int in_ld = 0;
    if (!out_f32) {
      if (args.flags & EPI_DGELU) {
        in_ptr = reinterpret_cast<const __nv_bfloat16*>(args.aux);
        in_ld = args.ldc;
      } else if (args.flags & EPI_RESIDUAL) {
        in_ptr = args.residual;
        in_ld = args.ldr;
      }
    }
    uint8_t* stg_warp = smem_epi + e * 2048;
    uint8_t* stg_own = stg_warp + lane * 64;         // this thread's row (64 B)
    const int own_sw = (lane >> 1) & 3;               // swizzle of this thread's row
    const int co_r = lane >> 2;                       // coalesced pattern: row inside a group of 8
    const int co_ch = lane & 3;                       // 16-byte chunk of the 64-byte segment
    bool rs_ready = false;                            // peers' partial tiles of the local chunk have landed
    int it = 0;
    for (int t = unit0; t < num_work; t += unit_stride, ++it) {
      const TileCoord tc = map_tile(t % num_tiles, m_blks_per_chunk, n_blks, args.num_chunks, args.first_chunk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row = tc.m_blk * BM_T + row_in_tile0 + q * 32 + lane;
      const int n0 = tc.n_blk * BN;
      // destination row pointer (possibly on a peer GPU)
      uint8_t* out_base = reinterpret_cast<uint8_t*>(args.out);
      int out_row = row;
      int row_limit = args.M;
      if (args.num_chunks > 1) {
        row_limit = (tc.chunk + 1) * chunk_rows;
        if (args.out_peer[0] != nullptr && !(args.flags & EPI_SCATTER)) {
          out_base = reinterpret_cast<uint8_t*>(args.out_peer[tc.chunk]);
          out_row = row - tc.chunk * chunk_rows;
        }
      }
      // GEMM -> reduce-scatter with the reduction fused into the local chunk's epilogue
      const bool rs_mode = args.rs_wait_ctr != nullptr && args.num_chunks > 1;
      const bool rs_local = rs_mode && tc.chunk == args.my_rank;
      if (rs_local) {
        out_base = reinterpret_cast<uint8_t*>(args.out);
        out_row = row - tc.chunk * chunk_rows;
        if (!rs_ready) {
          // peers computed my chunk FIRST; by the time I reach it their tiles have normally landed
          if (lane == 0) {
            for (int s2 = 0; s2 < args.num_chunks; ++s2)
              if (s2 != args.my_rank) spin_until_ge(args.rs_wait_ctr + s2, args.rs_wait_value, 3);
          }
          __syncwarp();
          rs_ready = true;
        }
      }
      const bool row_ok = row < row_limit && row < args.M;
      const int warp_row0 = tc.m_blk * BM_T + row_in_tile0 + q * 32;  // first row handled by this warp
      const int warp_out_row0 = out_row - lane;               // its destination row index
      const int rows_ok = max(0, min(32, min(row_limit, args.M) - warp_row0));
      const int local_row0 = warp_row0 - tc.chunk * chunk_rows;
      // bias / residual of a fused reduce-scatter belong to the final (local) rows, not to partials sent to peers;
      // the residual is then the local token shard [chunk_rows, N]
      const bool use_in = in_ptr != nullptr && (!rs_mode || rs_local);
      const bool use_bias = (args.flags & EPI_BIAS) && (!rs_mode || rs_local);
      uint4 pre[4];
      auto prefetch_in = [&](int c) {
        const int col = n0 + c * 32 + co_ch * 8;
        const __nv_bfloat16* base = in_ptr + static_cast<size_t>(rs_local ? local_row0 : warp_row0) * in_ld;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = i * 8 + co_r;
          pre[i] = make_uint4(0, 0, 0, 0);
          if (rr < rows_ok && col < args.N) pre[i] = ld_global_nc_v4(base + static_cast<size_t>(rr) * in_ld + col);
        }
      };
      if (use_in && h < NCH) prefetch_in(h);
      float ce_m = -INFINITY, ce_s = 0.f;  // online softmax partial of this thread's row over this warp's chunks
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = h; c < NCH; c += 2) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c * 32, v);
        if (use_in) {
          // park the prefetched inputs in the staging buffer, start fetching the next chunk's
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + co_r;
            *reinterpret_cast<uint4*>(stg_warp + rr * 64 + ((co_ch ^ ((rr >> 1) & 3)) << 4)) = pre[i];
          }
          __syncwarp();
          if (c + 2 < NCH) prefetch_in(c + 2);
        }
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        const int ncols = max(0, min(32, args.N - col0));  // multiple of 8
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        const bool active = row_ok && ncols > 0;
        if (active && use_bias) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g * 8 < ncols) {
              const uint4 bb = ld_global_nc_v4(args.bias + tc.chunk * args.bias_chunk_stride + col0 + g * 8);
              const uint32_t bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 bf = unpack_bf16x2(bw[j]);
                f[g * 8 + 2 * j] += bf.x;
                f[g * 8 + 2 * j + 1] += bf.y;
              }
            }
          }
        }
        if (out_f32) {
          // fp32 output (wgrad into the main-grad buffer): each thread owns a full 128-byte line
          if (active && args.grad_rs.world > 1) {
            // gradient reduce-scatter fused into the wgrad: the line goes to the owner of its ZeRO-1 slice
            const float* o = reinterpret_cast<const float*>(out_base) + static_cast<size_t>(out_row) * args.ldc + col0;
            const long long off = (o - args.grad_rs.local) - args.grad_rs.start;
            const long long own0 = off / args.grad_rs.seg;
            const bool one_owner = off - own0 * args.grad_rs.seg + 32 <= args.grad_rs.seg;
            float* t0 = grs_target(args.grad_rs, o);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (g * 4 < ncols) {
                const float4 val = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
                grs_add4_at(args.grad_rs, one_owner ? t0 + g * 4 : grs_target(args.grad_rs, o + g * 4), val);
              }
            }
            continue;
          }
          if (active) {
            float* o = reinterpret_cast<float*>(out_base) + static_cast<size_t>(out_row) * args.ldc + col0;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (g * 4 < ncols) {
                float4 val = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
                if (args.flags & EPI_ATOMIC) {
                  red_add_v4_f32(o + g * 4, val);  // split-K partial sums
                } else {
                  if (args.flags & EPI_ACCUM) {
                    const float4 old = *reinterpret_cast<const float4*>(o + g * 4);
                    val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
                  }
                  *reinterpret_cast<float4*>(o + g * 4) = val;
                }
              }
            }
          }
          continue;
        }
        // coalesced store of the staged [32 x 64 B] block to rows dst_row0.. of a bf16 matrix
        auto store_block = [&](__nv_bfloat16* base, int dst_row0, int ld) {
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + co_r;
            if (rr < rows_ok && co_ch * 8 < ncols) {
              const uint4 val = *reinterpret_cast<const uint4*>(stg_warp + rr * 64 + ((co_ch ^ ((rr >> 1) & 3)) << 4));
              st_global_v4(base + static_cast<size_t>(dst_row0 + rr) * ld + col0 + co_ch * 8, val);
            }
          }
          __syncwarp();
        };
        auto stage_own = [&]() {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(stg_own + ((g ^ own_sw) << 4)) =
                make_uint4(pack_bf16x2(f[g * 8 + 0], f[g * 8 + 1]), pack_bf16x2(f[g * 8 + 2], f[g * 8 + 3]),
                           pack_bf16x2(f[g * 8 + 4], f[g * 8 + 5]), pack_bf16x2(f[g * 8 + 6], f[g * 8 + 7]));
        };
        if (args.ce_part != nullptr) {
          // statistics of the values as they are STORED (rounded to bf16), so that the later softmax over the stored
          // logits is normalised exactly
          const int lim = min(ncols, args.ce_valid - col0);
          if (lim > 0) {
            float cm = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              f[i] = __bfloat162float(__float2bfloat16_rn(f[i]));
              if (i < lim) cm = fmaxf(cm, f[i]);
            }
            const float nm = fmaxf(ce_m, cm);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < lim) sum += exp2f((f[i] - nm) * 1.4426950408889634f);
            ce_s = ce_s * exp2f((ce_m - nm) * 1.4426950408889634f) + sum;
            ce_m = nm;
          }
        }
        if (args.flags & EPI_GELU) {
          if (args.aux != nullptr) {
            // the pre-activation goes out through the same staged, coalesced path
            stage_own();
            store_block(reinterpret_cast<__nv_bfloat16*>(args.aux), warp_row0, args.ldc);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = gelu_tanh(f[i]);
        }
        if (use_in) {
          const bool dgelu = (args.flags & EPI_DGELU) != 0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 z = *reinterpret_cast<const uint4*>(stg_own + ((g ^ own_sw) << 4));
            const uint32_t zw[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 zf = unpack_bf16x2(zw[j]);
              if (dgelu) {
                f[g * 8 + 2 * j] *= gelu_tanh_grad(zf.x);
                f[g * 8 + 2 * j + 1] *= gelu_tanh_grad(zf.y);
              } else {
                f[g * 8 + 2 * j] += zf.x;
                f[g * 8 + 2 * j + 1] += zf.y;
              }
            }
          }
        }
        if (rs_local) {
          // add every peer's staged partial of this chunk (local memory, written through NVLink and acquired above)
#pragma unroll 1
          for (int src = 0; src < args.num_chunks; ++src) {
            if (src == args.my_rank) continue;
            const __nv_bfloat16* base = args.rs_in[src] + static_cast<size_t>(local_row0) * args.N + col0 + co_ch * 8;
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = i * 8 + co_r;
              uint4 val = make_uint4(0, 0, 0, 0);
              if (rr < rows_ok && co_ch * 8 < ncols) val = ld_global_v4(base + static_cast<size_t>(rr) * args.N);
              *reinterpret_cast<uint4*>(stg_warp + rr * 64 + ((co_ch ^ ((rr >> 1) & 3)) << 4)) = val;
            }
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 z = *reinterpret_cast<const uint4*>(stg_own + ((g ^ own_sw) << 4));
              const uint32_t zw[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 zf = unpack_bf16x2(zw[j]);
                f[g * 8 + 2 * j] += zf.x;
                f[g * 8 + 2 * j + 1] += zf.y;
              }
            }
          }
          __syncwarp();
        }
        if (active && (args.flags & EPI_SCATTER)) {
          const float sc = args.row_scale[row];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] *= sc;
        }
        stage_own();
        if (args.flags & EPI_SCATTER) {
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + co_r;
            if (rr < rows_ok && co_ch * 8 < ncols) {
              const int gr = warp_row0 + rr;
              const int ret = args.row_ret[gr];
              if (ret >= 0) {
                const uint4 val = *reinterpret_cast<const uint4*>(stg_warp + rr * 64 + ((co_ch ^ ((rr >> 1) & 3)) << 4));
                const int src = (gr - tc.chunk * chunk_rows) / args.scatter_rows_per_src;
                st_global_v4(reinterpret_cast<__nv_bfloat16*>(args.out_peer[src]) +
                                 static_cast<size_t>(ret) * args.ldc + col0 + co_ch * 8, val);
              }
            }
          }
          __syncwarp();
        } else {
          store_block(reinterpret_cast<__nv_bfloat16*>(out_base), warp_out_row0, args.ldc);
        }
      }
      if (args.ce_part != nullptr && row_ok) {
        float2* dst = reinterpret_cast<float2*>(args.ce_part) + static_cast<size_t>(row) * (2 * n_blks) + 2 * tc.n_blk + h;
        *dst = make_float2(ce_m, ce_s);
      }
      // accumulator drained -> hand the TMEM stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTA2) {
          mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));  // the leader's MMA thread waits on it
        } else {
          mbar_arrive(&tmem_empty[acc]);
        }
      }
      if (args.num_chunks > 1 && args.arrive_ctr[0] != nullptr && !(args.flags & EPI_SCATTER) && !rs_local) {
        // all eight epilogue warps have stored their part of this tile -> publish to the owner
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 4 && lane == 0) {
          fence_acq_rel_sys();
          red_add_release_sys(args.arrive_ctr[tc.chunk], 1u);
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync();  // no CTA of the pair leaves (or frees TMEM) while the other still signals it
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CTA2) {
      tmem_dealloc_cta2(tmem_base, Cfg::kTmemCols);
    } else {
      tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
  }
}

}  // namespace pg
