// Causal ALiBi flash attention for sm_100a on Bloom's fused QKV layout.
//
//   qkv : [B*S, H*3*D] bf16, each row laid out [head][q|k|v][D]      out : [B*S, H*D] bf16
//   scores = q.k / sqrt(D) + slope[head] * (key_pos - query_pos)  (== slope * key_pos under softmax)
//
// Forward: one CTA per (128-query block, head, batch).  A TMA warp streams Q once and K/V blocks
// through a 2-stage ring; one MMA thread issues S = Q K^T and O_j = P_j V_j with tcgen05 (fp32
// accumulators in TMEM, S double-buffered so S_{j+1} overlaps the softmax of block j); four
// softmax warps own one query row per thread (tcgen05.ld 32x32b gives a thread a whole row, so
// row max / sum need no shuffles), write P as bf16 into 128B-swizzled shared memory for the PV
// MMA and keep the running output in registers (rescaled online).  [S, S] never exists.
//
// Backward: one CTA per (128-key block, head, batch) looping over the query blocks at or below
// the diagonal: S and dP = dO V^T on the tensor cores, P / dS recomputed per row in registers and
// staged (bf16, swizzled) for dV += P^T dO, dK += dS^T Q (TMEM accumulators across the loop) and
// dQ_i = dS K (read back per block and reduced into an fp32 buffer with vector red.global.add).
#include "launch.h"
#include "pdl_launch.cuh"
#include "ptx.cuh"
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace pg {

int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);

constexpr int ATT_BM = 128;  // queries per CTA
constexpr int ATT_BN = 128;  // keys per block
constexpr float kLog2e = 1.4426950408889634f;

// byte offset of 16-byte chunk `c16` (0..15 -> 128 bf16 columns) of row `r` inside a
// [128 rows x 128 cols] bf16 tile stored as two 64-column 128B-swizzled atoms of 16 KB
PG_DEVICE uint32_t sw128_tile_off(int r, int c16) {
  const int atom = c16 >> 3, c = c16 & 7;
  return atom * 16384 + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4);
}

PG_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

PG_DEVICE void red_add_v4_f32(float* addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b),
               "f"(c), "f"(d)
               : "memory");
}

template <int D>
struct AttFwdCfg {
  static constexpr int kTileBytes = ATT_BM * D * 2;       // Q, K or V tile
  static constexpr int kPBytes = ATT_BM * ATT_BN * 2;     // 32 KB
  static constexpr int kStages = 2;
  // D == 64: one S buffer and 256 TMEM columns so that TWO CTAs fit per SM (112 KB smem each) and
  // overlap each other's softmax / MMA phases; D == 128: S double-buffered inside one CTA per SM.
  static constexpr int kSBufs = (D == 64) ? 1 : 2;
  static constexpr int kTmemCols = (D == 64) ? 256 : 512;
  static constexpr int kSmemBytes = kTileBytes * (1 + 2 * kStages) + kPBytes + 256;  // base is 1024-aligned
  static constexpr int kThreads = 192;
};

struct AttFwdArgs {
  __nv_bfloat16* out;
  float* lse;  // [B, H, S] natural log
  const float* slopes;
  int B, S, H, ld_out;
  float scale_log2;  // log2(e) / sqrt(D)
};

template <int D>
__global__ void __launch_bounds__(192, (D == 64) ? 2 : 1)
    attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttFwdArgs args) {
  using Cfg = AttFwdCfg<D>;
  constexpr int ST = Cfg::kStages;
  constexpr int SB = Cfg::kSBufs;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::kTileBytes;
  uint8_t* sV = sK + ST * Cfg::kTileBytes;
  uint8_t* sP = sV + ST * Cfg::kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::kPBytes);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // ST
  uint64_t* kv_empty = kv_full + ST;  // ST
  uint64_t* s_full = kv_empty + ST;   // 2
  uint64_t* s_free = s_full + 2;      // 2
  uint64_t* p_ready = s_free + 2;     // 1
  uint64_t* o_full = p_ready + 1;     // 1
  uint64_t* o_free = o_full + 1;      // 1
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) pdl_launch_dependents();
  const int num_qb = (args.S + ATT_BM - 1) / ATT_BM;
  const int qb = num_qb - 1 - blockIdx.x;  // longest (most key blocks) first
  const int head = blockIdx.y, batch = blockIdx.z;
  const int num_kb = qb + 1;
  const int row0 = batch * args.S + qb * ATT_BM;  // global row of the first query
  const int col_q = head * 3 * D, col_k = col_q + D, col_v = col_q + 2 * D;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < ST; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 128);
    }
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    mbar_init(o_free, 128);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // prologue above overlapped the predecessor; Q/K/V, dO, lse, delta are read below
  const uint32_t tmem_S0 = tmem_base;              // SB x 128 columns
  const uint32_t tmem_O = tmem_base + SB * 128;    // D columns

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Cfg::kTileBytes);
#pragma unroll
      for (int a = 0; a < D / 64; ++a) tma_load_2d(sQ + a * 16384, &tm_qkv, q_full, col_q + a * 64, row0);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < num_kb; ++j) {
      mbar_wait(&kv_empty[stage], phase ^ 1);
      if (lane == 0) {
        const int krow = batch * args.S + j * ATT_BN;
        mbar_arrive_expect_tx(&kv_full[stage], 2 * Cfg::kTileBytes);
#pragma unroll
        for (int a = 0; a < D / 64; ++a) {
          tma_load_2d(sK + stage * Cfg::kTileBytes + a * 16384, &tm_qkv, &kv_full[stage], col_k + a * 64, krow);
          tma_load_2d(sV + stage * Cfg::kTileBytes + a * 16384, &tm_qkv, &kv_full[stage], col_v + a * 64, krow);
        }
      }
      __syncwarp();
      if (++stage == ST) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(128, ATT_BN, false, false);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, D, false, true);
    const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
    mbar_wait(q_full, 0);
    auto issue_s = [&](int j, int stage) {
      // S[j&1] = Q K_j^T
      const uint32_t aK = smem_u32(sK + stage * Cfg::kTileBytes);
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {
        const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
        umma_f16(tmem_S0 + (j % SB) * 128, make_smem_desc_sw128(aQ + off, 0, 1024),
                 make_smem_desc_sw128(aK + off, 0, 1024), idesc_s, k != 0 ? 1u : 0u);
      }
      umma_commit(&s_full[j % SB]);
    };
    int stage = 0;
    uint32_t phase = 0;
    // prologue: S_0
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    if (lane == 0) issue_s(0, 0);
    __syncwarp();
    for (int j = 0; j < num_kb; ++j) {
      // S_{j+1} while the softmax warps work on S_j
      if (j + 1 < num_kb) {
        const int nstage = (stage + 1 == ST) ? 0 : stage + 1;
        const uint32_t nphase = (stage + 1 == ST) ? phase ^ 1 : phase;
        mbar_wait(&kv_full[nstage], nphase);
        if (j + 1 >= SB) mbar_wait(&s_free[(j + 1) % SB], (((j + 1) / SB) - 1) & 1);
        tc_fence_after();
        if (lane == 0) issue_s(j + 1, nstage);
        __syncwarp();
      }
      // O_j = P_j V_j
      mbar_wait(p_ready, j & 1);
      if (j >= 1) mbar_wait(o_free, (j - 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t aV = smem_u32(sV + stage * Cfg::kTileBytes);
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; ++k) {
          const uint64_t da = make_smem_desc_sw128(aP + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024);
          const uint64_t db = make_smem_desc_sw128(aV + k * 2048, 16384, 1024);
          umma_f16(tmem_O, da, db, idesc_o, k != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[stage]);
      }
      __syncwarp();
      if (++stage == ST) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    // ===================== softmax / output warps (one query row per thread) =====================
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;                 // row inside the tile
    const int qpos = qb * ATT_BM + r;             // position inside the sequence
    const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
    const float slope2 = args.slopes[head] * kLog2e;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
    float acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = 0.f;

    for (int j = 0; j < num_kb; ++j) {
      mbar_wait(&s_full[j % SB], (j / SB) & 1);
      tc_fence_after();
      const uint32_t tS = tmem_S0 + (j % SB) * 128 + lane_addr;
      const bool diag = (j == qb);
      const int kbase = j * ATT_BN - qpos;  // key_pos - query_pos for column 0
      // x = s*scale_log2 + slope2*(kpos - qpos) = fma(s, scale_log2, fma(slope2, i, bias_c)): two FMAs per score,
      // the column index is an immediate; the causal mask only costs instructions on the diagonal block
      const float bias0 = slope2 * static_cast<float>(kbase);
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(tS + c * 32, v);
        tmem_ld_wait();
        const float bias_c = fmaf(slope2, static_cast<float>(c * 32), bias0);
        if (diag) {
          const int lim = -(kbase + c * 32);  // column i is visible iff i <= lim
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = fmaf(__uint_as_float(v[i]), args.scale_log2, fmaf(slope2, static_cast<float>(i), bias_c));
            mx = fmaxf(mx, i > lim ? -INFINITY : x);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            mx = fmaxf(mx, fmaf(__uint_as_float(v[i]), args.scale_log2, fmaf(slope2, static_cast<float>(i), bias_c)));
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = fast_exp2(m_run - m_new);  // m_run = -inf on the first block -> 0
      // fold in O_{j-1} (computed against m_run) before P_j overwrites the P buffer
      if (j >= 1) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_O + lane_addr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[c * 32 + i] = acc[c * 32 + i] * alpha_prev + __uint_as_float(v[i]);
        }
        tc_fence_before();
        mbar_arrive(o_free);
      }
      // pass 2: P_j = exp2(x - m_new) -> bf16 -> swizzled smem; row sum
      float rs = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(tS + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
        const float b2 = fmaf(slope2, static_cast<float>(c * 32), bias0) - m_new;  // exponent bias of this chunk
        if (diag) {
          const int lim = -(kbase + c * 32);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p[2];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const int col = 2 * i + h2;
              const float e = fast_exp2(fmaf(__uint_as_float(v[col]), args.scale_log2, fmaf(slope2, static_cast<float>(col), b2)));
              p[h2] = col > lim ? 0.f : e;
            }
            rs += p[0] + p[1];
            pk[i] = pack_bf16x2(p[0], p[1]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p[2];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const int col = 2 * i + h2;
              p[h2] = fast_exp2(fmaf(__uint_as_float(v[col]), args.scale_log2, fmaf(slope2, static_cast<float>(col), b2)));
            }
            rs += p[0] + p[1];
            pk[i] = pack_bf16x2(p[0], p[1]);
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t off = sw128_tile_off(r, c * 4 + g);
          *reinterpret_cast<uint4*>(sP + off) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        }
      }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      alpha_prev = alpha;
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(p_ready);
      mbar_arrive(&s_free[j % SB]);
    }
    // last block's O
    mbar_wait(o_full, (num_kb - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    const bool row_ok = qpos < args.S;
    __nv_bfloat16* orow = args.out + static_cast<size_t>(row0 + r) * args.ld_out + head * D;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_O + lane_addr + c * 32, v);
      tmem_ld_wait();
      float o[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = (acc[c * 32 + i] * alpha_prev + __uint_as_float(v[i])) * inv_l;
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          st_global_v4(orow + c * 32 + g * 8,
                       make_uint4(pack_bf16x2(o[8 * g], o[8 * g + 1]), pack_bf16x2(o[8 * g + 2], o[8 * g + 3]),
                                  pack_bf16x2(o[8 * g + 4], o[8 * g + 5]), pack_bf16x2(o[8 * g + 6], o[8 * g + 7])));
      }
    }
    if (row_ok)
      args.lse[(static_cast<size_t>(batch) * args.H + head) * args.S + qpos] = (m_run + log2f(l_run)) * 0.6931471805599453f;
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// -----------------------------------------------------------------------------------------------
// backward
// -----------------------------------------------------------------------------------------------
// delta[b,h,i] = sum_d dO[i,d] * O[i,d].  One warp per token row: every lane loads 16-byte chunks of the
// [H*D] row (fully coalesced), the D/8 lanes that share a head reduce with shuffles.
__global__ void __launch_bounds__(256) attention_delta_kernel(const __nv_bfloat16* __restrict__ dout,
                                                              const __nv_bfloat16* __restrict__ out,
                                                              float* __restrict__ delta, int B, int S, int H,
                                                              int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B * S) return;
  const int cph = D / 8;                 // 16-byte chunks (= lanes) per head: 8 or 16
  const int nchunks = H * cph;
  const __nv_bfloat16* a = dout + static_cast<size_t>(row) * H * D;
  const __nv_bfloat16* b = out + static_cast<size_t>(row) * H * D;
  const int bidx = row / S, spos = row % S;
  for (int c0 = 0; c0 < nchunks; c0 += 32) {
    const int c = c0 + lane;
    float s = 0.f;
    if (c < nchunks) {
      const uint4 x = ld_global_nc_v4(a + c * 8), y = ld_global_nc_v4(b + c * 8);
      const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(xw[j]), g = unpack_bf16x2(yw[j]);
        s += f.x * g.x + f.y * g.y;
      }
    }
    for (int o = cph >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (c < nchunks && (lane & (cph - 1)) == 0) {
      const int head = c / cph;
      delta[(static_cast<size_t>(bidx) * H + head) * S + spos] = s;
    }
  }
}

template <int D>
struct AttBwdCfg {
  static constexpr int kTileBytes = ATT_BM * D * 2;
  static constexpr int kPBytes = ATT_BM * ATT_BN * 2;
  static constexpr int kQStages = (D == 64) ? 2 : 1;
  // K, V resident; Q, dO ring; P, dS
  static constexpr int kSmemBytes = kTileBytes * (2 + 2 * kQStages) + 2 * kPBytes + 1024 + 256;
};

struct AttBwdArgs {
  const float* lse;
  const float* delta;
  const float* slopes;
  float* dq_acc;            // [B*S, H*D] fp32, zero-initialised
  __nv_bfloat16* dqkv;      // [B*S, H*3*D]
  int B, S, H;
  float scale, scale_log2;
};

// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..9 = compute.  Compute warp w owns TMEM lanes
// [32*(w%4), +32) (one query row per thread) and the 64-column half (w-2)/4 of every S/dP tile: two warps per
// scheduler hide each other's TMEM/MUFU latency and halve the serial softmax/dS time between MMA batches.
constexpr int kBwdThreads = 320;

template <int D>
__global__ void __launch_bounds__(kBwdThreads, 1)
    attention_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                         const AttBwdArgs args) {
  using Cfg = AttBwdCfg<D>;
  constexpr int QS = Cfg::kQStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + Cfg::kTileBytes;
  uint8_t* sQ = sV + Cfg::kTileBytes;
  uint8_t* sdO = sQ + QS * Cfg::kTileBytes;
  uint8_t* sP = sdO + QS * Cfg::kTileBytes;
  uint8_t* sdS = sP + Cfg::kPBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + Cfg::kPBytes);
  uint64_t* kv_full = bars;            // 1
  uint64_t* q_full = bars + 1;         // QS
  uint64_t* q_empty = q_full + QS;     // QS
  uint64_t* s_full = q_empty + QS;     // 1  (S and dP both computed)
  uint64_t* s_free = s_full + 1;       // 1  (threads finished reading S/dP, count 128)
  uint64_t* pds_ready = s_free + 1;    // 1  (P and dS staged, count 128)
  uint64_t* dq_full = pds_ready + 1;   // 1  (dV/dK/dQ MMAs of this block retired)
  uint64_t* dq_free = dq_full + 1;     // 1  (threads read dQ out of TMEM, count 128)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(dq_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) pdl_launch_dependents();
  const int num_qb = (args.S + ATT_BM - 1) / ATT_BM;
  const int jb = blockIdx.x;  // key block
  const int head = blockIdx.y, batch = blockIdx.z;
  const int n_iter = num_qb - jb;  // query blocks jb .. num_qb-1
  const int col_q = head * 3 * D, col_k = col_q + D, col_v = col_q + 2 * D;
  const int krow0 = batch * args.S + jb * ATT_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    mbar_init(kv_full, 1);
    for (int i = 0; i < QS; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 256);
    mbar_init(pds_ready, 256);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 256);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // prologue above overlapped the predecessor; Q/K/V, dO, lse, delta are read below
  // TMEM columns: S [0,128) dP [128,256) dV [256,256+D) dK [256+D,256+2D); dQ aliases S when D == 128
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 256 + D;
  const uint32_t tdQ = (D == 64) ? tmem_base + 384 : tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * Cfg::kTileBytes);
#pragma unroll
      for (int a = 0; a < D / 64; ++a) {
        tma_load_2d(sK + a * 16384, &tm_qkv, kv_full, col_k + a * 64, krow0);
        tma_load_2d(sV + a * 16384, &tm_qkv, kv_full, col_v + a * 64, krow0);
      }
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int qrow = batch * args.S + (jb + it) * ATT_BM;
      mbar_wait(&q_empty[stage], phase ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&q_full[stage], 2 * Cfg::kTileBytes);
#pragma unroll
        for (int a = 0; a < D / 64; ++a) {
          tma_load_2d(sQ + stage * Cfg::kTileBytes + a * 16384, &tm_qkv, &q_full[stage], col_q + a * 64, qrow);
          tma_load_2d(sdO + stage * Cfg::kTileBytes + a * 16384, &tm_do, &q_full[stage], head * D + a * 64, qrow);
        }
      }
      __syncwarp();
      if (++stage == QS) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(128, ATT_BN, false, false);   // S, dP: K-major x K-major
    constexpr uint32_t idesc_dkv = make_idesc_bf16(128, D, true, true);        // dV, dK: P^T/dS^T (MN) x dO/Q (MN)
    constexpr uint32_t idesc_dq = make_idesc_bf16(128, D, false, true);        // dQ: dS (K-major) x K (MN)
    const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP), adS = smem_u32(sdS);
    mbar_wait(kv_full, 0);
    int stage = 0;
    uint32_t phase = 0;
    auto issue_s_dp = [&](int st) {
      const uint32_t aQ = smem_u32(sQ + st * Cfg::kTileBytes), adO = smem_u32(sdO + st * Cfg::kTileBytes);
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {
        const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
        umma_f16(tS, make_smem_desc_sw128(aQ + off, 0, 1024), make_smem_desc_sw128(aK + off, 0, 1024), idesc_s,
                 k != 0 ? 1u : 0u);
      }
#pragma unroll
      for (int k = 0; k < D / 16; ++k) {
        const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
        umma_f16(tdP, make_smem_desc_sw128(adO + off, 0, 1024), make_smem_desc_sw128(aV + off, 0, 1024), idesc_s,
                 k != 0 ? 1u : 0u);
      }
      umma_commit(s_full);
    };
    mbar_wait(&q_full[0], 0);
    tc_fence_after();
    if (lane == 0) issue_s_dp(0);
    __syncwarp();
    for (int it = 0; it < n_iter; ++it) {
      mbar_wait(pds_ready, it & 1);
      if (it >= 1) mbar_wait(dq_free, (it - 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t aQ = smem_u32(sQ + stage * Cfg::kTileBytes), adO = smem_u32(sdO + stage * Cfg::kTileBytes);
        // dV += P^T dO ; dK += dS^T Q   (contraction over the 128 query rows)
#pragma unroll
        for (int k = 0; k < ATT_BM / 16; ++k) {
          umma_f16(tdV, make_smem_desc_sw128(aP + k * 2048, 16384, 1024),
                   make_smem_desc_sw128(adO + k * 2048, 16384, 1024), idesc_dkv, (it | k) != 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < ATT_BM / 16; ++k) {
          umma_f16(tdK, make_smem_desc_sw128(adS + k * 2048, 16384, 1024),
                   make_smem_desc_sw128(aQ + k * 2048, 16384, 1024), idesc_dkv, (it | k) != 0 ? 1u : 0u);
        }
        // dQ_i = dS K   (contraction over the 128 keys)
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; ++k) {
          umma_f16(tdQ, make_smem_desc_sw128(adS + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024),
                   make_smem_desc_sw128(aK + k * 2048, 16384, 1024), idesc_dq, k != 0 ? 1u : 0u);
        }
        umma_commit(dq_full);
        umma_commit(&q_empty[stage]);
      }
      __syncwarp();
      if (++stage == QS) {
        stage = 0;
        phase ^= 1;
      }
      if (it + 1 < n_iter) {
        // S/dP of the next query block: needs its Q/dO tile, S/dP TMEM drained by the threads and
        // (D == 128: dQ aliases S) the dQ read-out finished
        mbar_wait(&q_full[stage], phase);
        mbar_wait(s_free, it & 1);
        if (D == 128) mbar_wait(dq_free, it & 1);
        tc_fence_after();
        if (lane == 0) issue_s_dp(stage);
        __syncwarp();
      }
    }
  } else {
    // ===================== compute warps: one query row per thread =====================
    const int q4 = warp & 3;
    const int half = (warp - 2) >> 2;  // which 64 columns of S/dP (and which part of dQ/dK/dV) this warp handles
    const int r = q4 * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
    const float slope2 = args.slopes[head] * kLog2e;
    for (int it = 0; it < n_iter; ++it) {
      const int qb = jb + it;
      const int qpos = qb * ATT_BM + r;
      const bool row_ok = qpos < args.S;
      const size_t stat_idx = (static_cast<size_t>(batch) * args.H + head) * args.S + qpos;
      const float lse2 = row_ok ? args.lse[stat_idx] * kLog2e : 0.f;
      const float dlt = row_ok ? args.delta[stat_idx] : 0.f;
      const bool diag = (it == 0);
      const int kbase = jb * ATT_BN - qpos;
      // p = exp2(s*scale_log2 + slope2*(kpos-qpos) - lse2) = exp2(fma(s, scale_log2, fma(slope2, i, bias_c)));
      // dS = p * (dP - delta) * scale = p * fma(dP, scale, -delta*scale).  Rows past the sequence get p = 0
      // through an infinite negative bias; the causal mask only costs instructions on the diagonal block.
      const float bias0 = row_ok ? fmaf(slope2, static_cast<float>(kbase), -lse2) : -INFINITY;
      const float nds = -dlt * args.scale;
      mbar_wait(s_full, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = half * 2; c < half * 2 + 2; ++c) {
        uint32_t vs[32], vd[32];
        __syncwarp();
        tmem_ld_32x32(tS + lane_addr + c * 32, vs);
        tmem_ld_32x32(tdP + lane_addr + c * 32, vd);
        tmem_ld_wait();
        uint32_t pp[16], ds[16];
        const float bias_c = fmaf(slope2, static_cast<float>(c * 32), bias0);
        const int lim = diag ? -(kbase + c * 32) : 64;  // column i of this chunk is visible iff i <= lim
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p[2], g[2];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int col = 2 * i + h2;
            float e = fast_exp2(fmaf(__uint_as_float(vs[col]), args.scale_log2, fmaf(slope2, static_cast<float>(col), bias_c)));
            if (diag) e = col > lim ? 0.f : e;
            p[h2] = e;
            g[h2] = e * fmaf(__uint_as_float(vd[col]), args.scale, nds);
          }
          pp[i] = pack_bf16x2(p[0], p[1]);
          ds[i] = pack_bf16x2(g[0], g[1]);
        }
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          const uint32_t off = sw128_tile_off(r, c * 4 + g4);
          *reinterpret_cast<uint4*>(sP + off) = make_uint4(pp[4 * g4], pp[4 * g4 + 1], pp[4 * g4 + 2], pp[4 * g4 + 3]);
          *reinterpret_cast<uint4*>(sdS + off) = make_uint4(ds[4 * g4], ds[4 * g4 + 1], ds[4 * g4 + 2], ds[4 * g4 + 3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(s_free);
      mbar_arrive(pds_ready);
      // dQ_i partial for this key block -> fp32 accumulator in global memory
      mbar_wait(dq_full, it & 1);
      tc_fence_after();
      float* dq_row = args.dq_acc + static_cast<size_t>(batch * args.S + qpos) * (args.H * D) + head * D;
#pragma unroll 1
      for (int c = half * (D / 64); c < (half + 1) * (D / 64); ++c) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(tdQ + lane_addr + c * 32, v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            red_add_v4_f32(dq_row + c * 32 + g * 4, __uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                           __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
        }
      }
      tc_fence_before();
      mbar_arrive(dq_free);
    }
    // dK, dV of this key block (all MMAs retired: the last dq_full covered them)
    const int kpos = jb * ATT_BN + r;
    const bool krow_ok = kpos < args.S;
    __nv_bfloat16* dk_row = args.dqkv + static_cast<size_t>(batch * args.S + kpos) * (args.H * 3 * D) + col_k;
    __nv_bfloat16* dv_row = dk_row + D;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const uint32_t t = which == 0 ? tdK : tdV;
      __nv_bfloat16* dst = which == 0 ? dk_row : dv_row;
#pragma unroll 1
      for (int c = half * (D / 64); c < (half + 1) * (D / 64); ++c) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(t + lane_addr + c * 32, v);
        tmem_ld_wait();
        if (krow_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            st_global_v4(dst + c * 32 + g * 8,
                         make_uint4(pack_bf16x2(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1])),
                                    pack_bf16x2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])),
                                    pack_bf16x2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])),
                                    pack_bf16x2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7]))));
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dq fp32 [rows, H*D] -> bf16 into dqkv[rows, H*3*D] (q slot of every head)
__global__ void __launch_bounds__(256) attention_dq_convert_kernel(const float* __restrict__ dq,
                                                                   __nv_bfloat16* __restrict__ dqkv, int64_t rows,
                                                                   int H, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t idx = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  const int64_t total = rows * H * D;
  if (idx >= total) return;
  const int64_t row = idx / (H * D);
  const int rem = static_cast<int>(idx % (H * D));
  const int head = rem / D, d = rem % D;
  const float4 a = *reinterpret_cast<const float4*>(dq + idx);
  const float4 b = *reinterpret_cast<const float4*>(dq + idx + 4);
  st_global_v4(dqkv + row * (H * 3 * D) + head * 3 * D + d,
               make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w)));
}

}  // namespace pg

using namespace pg;

static int att_tmap(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols) {
  struct Key {
    const void* p;
    uint64_t r, c;
    bool operator==(const Key& o) const { return p == o.p && r == o.r && c == o.c; }
  };
  struct H {
    size_t operator()(const Key& k) const { return reinterpret_cast<size_t>(k.p) ^ (k.r * 1315423911u) ^ (k.c << 20); }
  };
  static std::unordered_map<Key, CUtensorMap, H> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  Key key{ptr, rows, cols};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return 0;
  }
  if (make_tmap_bf16_2d(out, ptr, rows, cols, cols, ATT_BM, 64) != 0) return -1;
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

#define PG_CHECK_LAUNCH(name)                                                       \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      fprintf(stderr, "pipegoose_b200: %s launch failed: %s\n", name, cudaGetErrorString(e__)); \
      return -1;                                                                    \
    }                                                                               \
  } while (0)

template <int D>
static int launch_att_fwd(const CUtensorMap& tm, const AttFwdArgs& a, cudaStream_t s) {
  using Cfg = AttFwdCfg<D>;
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(attention_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) return -1;
    set = true;
  }
  dim3 grid((a.S + ATT_BM - 1) / ATT_BM, a.H, a.B);
  if (launch_pdl(attention_fwd_kernel<D>, grid, dim3(Cfg::kThreads), Cfg::kSmemBytes, s, tm, a) != cudaSuccess) return -1;
  PG_CHECK_LAUNCH("attention_fwd");
  return 0;
}

// softmax_scale <= 0: 1/sqrt(D).  An explicit scale lets a head that was zero-padded to a supported width (e.g. D=80
// -> 128: the padded columns contribute nothing to Q.K) keep the scale of its true width.
extern "C" int pg_attention_fwd(const void* qkv, const float* slopes, void* out, float* lse, int B, int S, int H,
                                int D, float softmax_scale, cudaStream_t s) {
  if (D != 64 && D != 128) return -1;
  CUtensorMap tm;
  if (att_tmap(&tm, qkv, static_cast<uint64_t>(B) * S, static_cast<uint64_t>(H) * 3 * D) != 0) return -1;
  AttFwdArgs a;
  a.out = (__nv_bfloat16*)out;
  a.lse = lse;
  a.slopes = slopes;
  a.B = B; a.S = S; a.H = H;
  a.ld_out = H * D;
  a.scale_log2 = kLog2e * (softmax_scale > 0.f ? softmax_scale : 1.0f / sqrtf(static_cast<float>(D)));
  return D == 64 ? launch_att_fwd<64>(tm, a, s) : launch_att_fwd<128>(tm, a, s);
}

template <int D>
static int launch_att_bwd(const CUtensorMap& tq, const CUtensorMap& tdo, const AttBwdArgs& a, cudaStream_t s) {
  using Cfg = AttBwdCfg<D>;
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(attention_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) return -1;
    set = true;
  }
  dim3 grid((a.S + ATT_BN - 1) / ATT_BN, a.H, a.B);
  if (launch_pdl(attention_bwd_kernel<D>, grid, dim3(kBwdThreads), Cfg::kSmemBytes, s, tq, tdo, a) != cudaSuccess) return -1;
  PG_CHECK_LAUNCH("attention_bwd");
  return 0;
}

// dq_acc: fp32 workspace [B*S, H*D]; delta: fp32 [B, H, S]
extern "C" int pg_attention_bwd(const void* qkv, const float* slopes, const void* out, const float* lse,
                                const void* dout, void* dqkv, float* dq_acc, float* delta, int B, int S, int H,
                                int D, float softmax_scale, cudaStream_t s) {
  if (D != 64 && D != 128) return -1;
  const int64_t rows = static_cast<int64_t>(B) * S;
  CUtensorMap tq, tdo;
  if (att_tmap(&tq, qkv, rows, static_cast<uint64_t>(H) * 3 * D) != 0) return -1;
  if (att_tmap(&tdo, dout, rows, static_cast<uint64_t>(H) * D) != 0) return -1;
  if (cudaMemsetAsync(dq_acc, 0, rows * H * D * sizeof(float), s) != cudaSuccess) return -1;
  {
    if (launch_pdl(attention_delta_kernel, dim3((unsigned)((rows * 32 + 255) / 256)), dim3(256), 0, s,
                   (const __nv_bfloat16*)dout, (const __nv_bfloat16*)out, delta, B, S, H, D) != cudaSuccess) return -1;
    PG_CHECK_LAUNCH("attention_delta");
  }
  AttBwdArgs a;
  a.lse = lse; a.delta = delta; a.slopes = slopes; a.dq_acc = dq_acc;
  a.dqkv = (__nv_bfloat16*)dqkv;
  a.B = B; a.S = S; a.H = H;
  a.scale = softmax_scale > 0.f ? softmax_scale : 1.0f / sqrtf(static_cast<float>(D));
  a.scale_log2 = a.scale * kLog2e;
  const int rc = D == 64 ? launch_att_bwd<64>(tq, tdo, a, s) : launch_att_bwd<128>(tq, tdo, a, s);
  if (rc != 0) return rc;
  const int64_t total = rows * H * D;
  if (launch_pdl(attention_dq_convert_kernel, dim3((unsigned)((total / 8 + 255) / 256)), dim3(256), 0, s, (const float*)dq_acc,
                 (__nv_bfloat16*)dqkv, (int64_t)rows, H, D) != cudaSuccess) return -1;
  PG_CHECK_LAUNCH("attention_dq_convert");
  return 0;
}
