// K3/K6/K9/K10c — D = act(A . W^T + bias + residual), bf16 operands, fp32 accumulation in TMEM.
//
// Persistent warp-specialised tcgen05 kernel (one CTA per SM, cta_group::1, 384 threads):
//   warp 0 (elected lane) TMA producer : A tile [128 x 64] + W tile [BN x 64] per stage, 128B swizzle
//   warp 1 (elected lane) MMA issuer   : 4 x tcgen05.mma 128 x BN x 16 per stage, accumulator in TMEM
//   warp 2             TMEM allocator (2 accumulator stages x BN columns)
//   warps 4-11         epilogue     : warp = (TMEM lane quarter, 32-column half of a 64-column slab);
//                                     tcgen05.ld 32x32b (thread = row), bias / residual / activation in
//                                     registers, bf16 pack -> 128B-swizzled staging slab in shared memory ->
//                                     one TMA store per [128 x 64] slab (double-buffered, bulk async groups).
// Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), static tile scheduler
// (tile = blockIdx.x + i * gridDim.x, N fastest so the CTAs resident at one time share A tiles through L2).
// Out-of-bounds rows/cols/K are zero-filled by TMA on load and clipped by TMA on store.
// fp32 / atomic outputs, unaligned D and BN = 32 tiles use predicated direct stores instead of the TMA store.
//
// Round-1 ncu finding that shaped this epilogue (profiles/r1_ncu_gemm_a.md): with 4 epilogue warps storing
// 16-byte pieces of 32 different rows per instruction, K = 768 GEMMs ran at 20-28 % tensor-pipe activity
// (epilogue-bound) while K = 3072 reached 60-66 %.
#include "common.cuh"
#include "host_util.h"

namespace opsg {

size_t gemm_skinny_workspace_bytes(int N, int K);
int launch_gemm_skinny(const opsg_bf16* A, int lda, const opsg_bf16* W, int ldw, void* D, int ldd, int M, int N, int K,
                       const float* bias, const opsg_bf16* residual, int ldr, int act, int out_mode, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream);
int launch_gemm_2cta(const opsg_bf16* A, int lda, const opsg_bf16* W, int ldw, opsg_bf16* D, int ldd, int M, int N, int K,
                     const float* bias, int bias_along_m, const opsg_bf16* residual, int ldr, int act, const GemmLnFold* ln,
                     cudaStream_t stream);

constexpr int kBM = 128;
constexpr int kBK = 64;            // 64 bf16 = 128 B = one swizzle span
constexpr int kGemmThreads = 384;
constexpr int kEpiThreads = 256;
constexpr int kSlabBytes = kBM * 128;   // [128 rows x 64 bf16] staging slab

struct GemmParams {
  void* D;
  const float* bias;
  const __nv_bfloat16* residual;
  int M, N, K;
  int ldd, ldr;
  int bias_along_m;
  int act;
  int out_mode;
  int k_splits;
  int m_tiles, n_tiles;
  int tma_store;
  int stream_k;      // 1: contiguous (tile, k-block) ranges per CTA, fp32 partial tiles to `ws` (small-M weight streaming)
  int max_segs;      // stream-K: partial-tile slots per CTA
  float* ws;         // stream-K workspace: [grid * max_segs][128][BN] fp32
};

// One unit of work of a CTA: a (m_t, n_t) output tile and the K blocks [kb0, kb1) it accumulates.
//  * static mode : whole tiles (or k_splits slices), tile index = blockIdx.x + i * gridDim.x;
//  * stream-K    : the flattened (tile, k-block) space is cut into gridDim.x equal contiguous ranges, so every CTA
//                  streams the same number of weight bytes whatever N is; a range that crosses tile boundaries yields
//                  several segments, each stored as an fp32 partial tile in slot blockIdx.x * max_segs + j.
struct GemmSeg { int m_t, n_t, kb0, kb1, slot; };
struct GemmSegIter {
  long long u, u_end;
  int tile, j;
  __device__ __forceinline__ void init(const GemmParams& p, int kb_total) {
    j = 0;
    if (p.stream_k) {
      const long long U = static_cast<long long>(p.m_tiles) * p.n_tiles * kb_total;
      u = U * blockIdx.x / gridDim.x;
      u_end = U * (blockIdx.x + 1) / gridDim.x;
    } else {
      tile = blockIdx.x;
    }
  }
  __device__ __forceinline__ bool next(const GemmParams& p, int kb_total, int kb_per_split, GemmSeg& s) {
    if (p.stream_k) {
      if (u >= u_end) return false;
      const int t = static_cast<int>(u / kb_total);
      s.kb0 = static_cast<int>(u % kb_total);
      s.kb1 = static_cast<int>(min(static_cast<long long>(kb_total), s.kb0 + (u_end - u)));
      s.n_t = t % p.n_tiles;
      s.m_t = t / p.n_tiles;
      s.slot = blockIdx.x * p.max_segs + j;
      u += s.kb1 - s.kb0;
      ++j;
      return true;
    }
    if (tile >= p.m_tiles * p.n_tiles * p.k_splits) return false;
    s.n_t = tile % p.n_tiles;
    const int rest = tile / p.n_tiles;
    s.m_t = rest % p.m_tiles;
    const int split = rest / p.m_tiles;
    s.kb0 = min(kb_total, split * kb_per_split);
    s.kb1 = min(kb_total, s.kb0 + kb_per_split);
    s.slot = split;      // k_splits > 1 with OUT_F32: index of this split's partial slice
    tile += gridDim.x;
    return true;
  }
};

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = (BN >= 64) ? 2 * kSlabBytes : 0;
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kTotal = STAGES * kStageBytes + kStagingBytes + kBarrierBytes + 1024;  // +1024: manual alignment
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmD, const GemmParams p) {
  using S = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * S::kABytes;
  uint8_t* staging = smem + STAGES * S::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + S::kStagingBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmD);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiThreads);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_base_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  pdl_wait_then_trigger();          // everything above overlaps the previous kernel (programmatic dependent launch)
  const int kb_total = (p.K + kBK - 1) / kBK;
  const int kb_per_split = (kb_total + p.k_splits - 1) / p.k_splits;

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    GemmSegIter it;
    GemmSeg sg;
    it.init(p, kb_total);
    while (it.next(p, kb_total, kb_per_split, sg)) {
      const int n_t = sg.n_t, m_t = sg.m_t, kb0 = sg.kb0, kb1 = sg.kb1;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&full_bar[stage], S::kStageBytes);
          tma_load_2d(smem_a + stage * S::kABytes, &tmA, &full_bar[stage], kb * kBK, m_t * kBM);
          tma_load_2d(smem_b + stage * S::kBBytes, &tmB, &full_bar[stage], kb * kBK, n_t * BN);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, elect.sync lane issues tcgen05.mma + commits) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    GemmSegIter it;
    GemmSeg sg;
    it.init(p, kb_total);
    while (it.next(p, kb_total, kb_per_split, sg)) {
      const int kb0 = sg.kb0, kb1 = sg.kb1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t a_desc = umma_desc_k_sw128(smem_u32(smem_a + stage * S::kABytes));
          const uint64_t b_desc = umma_desc_k_sw128(smem_u32(smem_b + stage * S::kBBytes));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)      // +32 bytes per K step = +2 in the descriptor's 16-byte address field
            umma_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          tc_commit(&empty_bar[stage]);
          if (kb + 1 == kb1) tc_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (kb0 >= kb1) {                           // split without K blocks: still hand an (unused) accumulator over
        if (elect_one_sync()) tc_commit(&tmem_full[acc]);
        __syncwarp();
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    constexpr int SLABW = (BN >= 64) ? 64 : BN;        // columns per slab
    constexpr int NSLAB = BN / SLABW;
    const int ew = warp - 4;
    const int q = ew & 3;                   // TMEM lane quarter this warp may access (warp id % 4)
    const int half = ew >> 2;               // which 32 columns of the slab
    const bool has_cols = half * 32 < SLABW;
    const int row_in_tile = q * 32 + lane;
    const bool elected = (threadIdx.x == 4 * 32);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t buf = 0;
    const __nv_bfloat16* resid = p.residual;

    GemmSegIter it;
    GemmSeg sg;
    it.init(p, kb_total);
    while (it.next(p, kb_total, kb_per_split, sg)) {
      const int n_t = sg.n_t, m_t = sg.m_t;
      const bool has_k = sg.kb0 < sg.kb1;
      const int row = m_t * kBM + row_in_tile;
      const bool row_ok = row < p.M;
      const float bias_m = (p.bias && p.bias_along_m && row_ok) ? p.bias[row] : 0.f;

      // residual prefetch (registers) for slab s: this thread's 32 columns of its own row.  Rows / columns that
      // cannot use 16-byte loads (ragged N, unaligned views) are read element-wise at add time instead.
      uint4 rres[4];
      bool rfast = false;
      auto fetch_residual = [&](int slab) {
        const int c0 = n_t * BN + slab * SLABW + half * 32;
        rfast = false;
        if (!resid || !row_ok || !has_cols || c0 + 32 > p.N) return;
        const __nv_bfloat16* r = resid + static_cast<size_t>(row) * p.ldr + c0;
        if ((reinterpret_cast<uintptr_t>(r) & 15) != 0) return;
        rfast = true;
#pragma unroll
        for (int j = 0; j < 4; ++j) rres[j] = __ldg(reinterpret_cast<const uint4*>(r) + j);
      };
      fetch_residual(0);

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;

#pragma unroll 1
      for (int slab = 0; slab < NSLAB; ++slab) {
        const int col0 = n_t * BN + slab * SLABW + half * 32;
        float f[32];
        if (has_cols) {
          uint32_t v[32];
          tmem_ld32(taddr + slab * SLABW + half * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = has_k ? __uint_as_float(v[j]) : 0.f;
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = 0.f;
        }
        if (slab == NSLAB - 1) {            // accumulator fully read -> hand the TMEM stage back to the MMA warp
          tc_fence_before();
          mbar_arrive(&tmem_empty[acc]);
        }
        const bool col_ok = has_cols && col0 < p.N;
        const bool full = col0 + 32 <= p.N;
        if (col_ok) {
          if (p.bias) {
            if (p.bias_along_m) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] += bias_m;
            } else if (full && ((reinterpret_cast<uintptr_t>(p.bias + col0) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
                f[j * 4 + 0] += b4.x; f[j * 4 + 1] += b4.y; f[j * 4 + 2] += b4.z; f[j * 4 + 3] += b4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) f[j] += __ldg(p.bias + col0 + j);
            }
          }
          if (resid && !rfast && row_ok) {
            const __nv_bfloat16* r = resid + static_cast<size_t>(row) * p.ldr + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) f[j] += __bfloat162float(r[j]);
          }
          if (rfast) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              f[j * 8 + 0] += bf16_lo(rres[j].x); f[j * 8 + 1] += bf16_hi(rres[j].x);
              f[j * 8 + 2] += bf16_lo(rres[j].y); f[j * 8 + 3] += bf16_hi(rres[j].y);
              f[j * 8 + 4] += bf16_lo(rres[j].z); f[j * 8 + 5] += bf16_hi(rres[j].z);
              f[j * 8 + 6] += bf16_lo(rres[j].w); f[j * 8 + 7] += bf16_hi(rres[j].w);
            }
          }
        }
        if (slab + 1 < NSLAB) fetch_residual(slab + 1);
        if (col_ok) {
          if (p.act == OPSG_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_erf_fast(f[j]);
          } else if (p.act == OPSG_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
        }

        if (p.stream_k) {
          // fp32 partial tile -> workspace slot (plain 16-byte stores; 128 contiguous bytes per thread)
          if (row_ok && has_cols) {
            float* d = p.ws + (static_cast<size_t>(sg.slot) * kBM + row_in_tile) * BN + slab * SLABW + half * 32;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<float4*>(d)[j] = make_float4(f[j * 4], f[j * 4 + 1], f[j * 4 + 2], f[j * 4 + 3]);
          }
        } else if (p.tma_store) {
          // staging slab `buf`: its previous TMA store (two slabs ago) must have drained before we overwrite it
          if (elected) tma_store_wait_read<1>();
          named_bar_sync(1, kEpiThreads);
          if (has_cols) {
            uint8_t* rowp = staging + buf * kSlabBytes + row_in_tile * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int chunk = (half * 4 + g) ^ (row_in_tile & 7);
              uint4 u;
              u.x = pack_bf16x2(f[g * 8 + 0], f[g * 8 + 1]);
              u.y = pack_bf16x2(f[g * 8 + 2], f[g * 8 + 3]);
              u.z = pack_bf16x2(f[g * 8 + 4], f[g * 8 + 5]);
              u.w = pack_bf16x2(f[g * 8 + 6], f[g * 8 + 7]);
              *reinterpret_cast<uint4*>(rowp + chunk * 16) = u;
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(2, kEpiThreads);
          if (elected) {     // always commit (possibly empty) so that group counting stays one-per-slab
            if (has_k && n_t * BN + slab * SLABW < p.N)
              tma_store_2d(staging + buf * kSlabBytes, &tmD, n_t * BN + slab * SLABW, m_t * kBM);
            tma_store_commit();
          }
          buf ^= 1;
        } else if (row_ok && col_ok && (has_k || (p.out_mode == OPSG_OUT_F32 && p.k_splits > 1))) {
          if (p.out_mode == OPSG_OUT_BF16) {
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.D) + static_cast<size_t>(row) * p.ldd + col0;
            if (full && ((reinterpret_cast<uintptr_t>(d) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 u;
                u.x = pack_bf16x2(f[j * 8 + 0], f[j * 8 + 1]);
                u.y = pack_bf16x2(f[j * 8 + 2], f[j * 8 + 3]);
                u.z = pack_bf16x2(f[j * 8 + 4], f[j * 8 + 5]);
                u.w = pack_bf16x2(f[j * 8 + 6], f[j * 8 + 7]);
                reinterpret_cast<uint4*>(d)[j] = u;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) d[j] = __float2bfloat16(f[j]);
            }
          } else if (p.out_mode == OPSG_OUT_F32) {
            // k_splits > 1: split s writes its fp32 partial to slice s of D ([k_splits][M][ldd]; opsg_splitk_reduce sums
            // the slices in split order -- deterministic, unlike OPSG_OUT_F32_ATOMIC)
            float* d = reinterpret_cast<float*>(p.D) + (static_cast<size_t>(p.k_splits > 1 ? sg.slot : 0) * p.M + row) * p.ldd + col0;
            if (full && ((reinterpret_cast<uintptr_t>(d) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                reinterpret_cast<float4*>(d)[j] = make_float4(f[j * 4], f[j * 4 + 1], f[j * 4 + 2], f[j * 4 + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) d[j] = f[j];
            }
          } else {  // OPSG_OUT_F32_ATOMIC (split-K)
            float* d = reinterpret_cast<float*>(p.D) + static_cast<size_t>(row) * p.ldd + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) atomicAdd(d + j, f[j]);
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tma_store && elected) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BN, int STAGES>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD, GemmParams p,
                       cudaStream_t stream) {
  using S = GemmSmem<BN, STAGES>;
  static_assert(S::kTotal <= 232448, "shared memory budget exceeded");
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[device_slot()];
  if (!configured) {
    int rc = check_cuda(cudaFuncSetAttribute(gemm_bf16_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             S::kTotal), "cudaFuncSetAttribute(gemm)");
    if (rc) return rc;
    configured = true;
  }
  if (BN < 64) p.tma_store = 0;
  p.n_tiles = (p.N + BN - 1) / BN;
  p.m_tiles = (p.M + kBM - 1) / kBM;
  const int total = p.m_tiles * p.n_tiles * p.k_splits;
  const int grid = total < opsg_num_sms() ? total : opsg_num_sms();
  launch_kernel(gemm_bf16_kernel<BN, STAGES>, grid, kGemmThreads, S::kTotal, stream, tmA, tmB, tmD, p);
  OPSG_CHECK_LAUNCH("gemm_bf16_kernel");
  return OPSG_OK;
}

// ------------------------------------------------------------------------------------------------
// stream-K fix-up: D[m, n] = act(bias[n] + sum over the partial tiles covering (tile n / 256) + residual[m, n]).
// The CTA -> range mapping is recomputed with the same integer arithmetic as GemmSegIter.
// ------------------------------------------------------------------------------------------------
struct StreamKReduceParams {
  const float* ws;
  void* D;
  const float* bias;
  const __nv_bfloat16* residual;
  int M, N, ldd, ldr, act, out_f32;
  int n_tiles, kb_total, grid, max_segs;
};

__global__ void __launch_bounds__(256) streamk_reduce_kernel(const StreamKReduceParams p) {
  pdl_wait_then_trigger();
  constexpr int BN = 256;
  const int n_t = blockIdx.x;
  const int c4 = threadIdx.x & 63;                       // 4-column group inside the tile
  const int col = n_t * BN + c4 * 4;
  const long long U = static_cast<long long>(p.n_tiles) * p.kb_total;
  const long long lo = static_cast<long long>(n_t) * p.kb_total, hi = lo + p.kb_total;
  // CTAs whose range [u0, u1) intersects [lo, hi)
  int c_first = static_cast<int>(lo * p.grid / U);
  while (c_first > 0 && U * c_first / p.grid > lo) --c_first;
  for (int row = blockIdx.y * 4 + (threadIdx.x >> 6); row < p.M; row += gridDim.y * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = c_first; c < p.grid; ++c) {
      const long long u0 = U * c / p.grid, u1 = U * (c + 1) / p.grid;
      if (u0 >= hi) break;
      if (u1 <= lo || u1 <= u0) continue;
      const int j = n_t - static_cast<int>(u0 / p.kb_total);            // index of this tile among the CTA's segments
      const float4 v = *reinterpret_cast<const float4*>(p.ws + (static_cast<size_t>(c * p.max_segs + j) * kBM + row) * BN + c4 * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (col >= p.N) continue;
    float f[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (col + e >= p.N) break;
      float x = f[e];
      if (p.bias) x += __ldg(p.bias + col + e);
      if (p.residual) x += __bfloat162float(p.residual[static_cast<size_t>(row) * p.ldr + col + e]);
      if (p.act == OPSG_ACT_GELU) x = gelu_erf(x);
      else if (p.act == OPSG_ACT_RELU) x = fmaxf(x, 0.f);
      if (p.out_f32) reinterpret_cast<float*>(p.D)[static_cast<size_t>(row) * p.ldd + col + e] = x;
      else reinterpret_cast<__nv_bfloat16*>(p.D)[static_cast<size_t>(row) * p.ldd + col + e] = __float2bfloat16(x);
    }
  }
}

static void streamk_shape(int N, int K, int sms, int* n_tiles, int* kb_total, int* grid, int* max_segs) {
  *n_tiles = (N + 255) / 256;
  *kb_total = (K + kBK - 1) / kBK;
  const long long U = static_cast<long long>(*n_tiles) * *kb_total;
  *grid = static_cast<int>(U < sms ? U : sms);
  const long long per = (U + *grid - 1) / *grid;
  *max_segs = static_cast<int>((per + *kb_total - 1) / *kb_total) + 1;
}

}  // namespace opsg

using namespace opsg;

extern "C" int opsg_gemm_bf16_ln(const opsg_bf16* A, int lda, const opsg_bf16* W, int ldw, opsg_bf16* D, int ldd, int M, int N,
                                 int K, const float* bias, const opsg_bf16* residual, int ldr, int act, const float* a_stats,
                                 const float* a_colsum, const float* r_stats, const float* r_gamma, const float* r_beta,
                                 float* stats_out, float eps, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(A && W && D, "gemm_ln: null pointer");
  OPSG_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_ln: bad shape M=%d N=%d K=%d", M, N, K);
  OPSG_CHECK_ARG(lda >= K && ldw >= K && ldd >= N, "gemm_ln: leading dimension too small");
  OPSG_CHECK_ARG((lda % 8) == 0 && (ldw % 8) == 0 && (ldd % 8) == 0 && (N % 8) == 0, "gemm_ln: lda/ldw/ldd/N must be multiples of 8");
  OPSG_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)D & 15) == 0, "gemm_ln: A/W/D must be 16-byte aligned");
  OPSG_CHECK_ARG(act >= OPSG_ACT_NONE && act <= OPSG_ACT_RELU, "gemm_ln: bad activation");
  OPSG_CHECK_ARG(!a_stats || a_colsum, "gemm_ln: a_stats without a_colsum");
  OPSG_CHECK_ARG(!r_stats || (residual && r_gamma && r_beta), "gemm_ln: r_stats needs residual, r_gamma and r_beta");
  OPSG_CHECK_ARG(!residual || (ldr >= N && (ldr % 8) == 0 && ((uintptr_t)residual & 15) == 0), "gemm_ln: bad residual layout");
  OPSG_CHECK_ARG(!r_stats || ((((uintptr_t)r_gamma | (uintptr_t)r_beta) & 15) == 0), "gemm_ln: r_gamma / r_beta must be 16-byte aligned");
  GemmLnFold ln{a_stats, a_colsum, r_stats, r_gamma, r_beta, stats_out, eps};
  rc = launch_gemm_2cta(A, lda, W, ldw, D, ldd, M, N, K, bias, 0, residual, ldr, act, &ln, reinterpret_cast<cudaStream_t>(stream));
  if (rc == OPSG_E_UNSUPPORTED) return set_error(rc, "gemm_ln: unsupported layout");
  return rc;
}

extern "C" size_t opsg_gemm_streamk_workspace_bytes(int N, int K) {
  if (N <= 0 || K <= 0) return 0;
  int n_tiles, kb_total, grid, max_segs;
  int sms = opsg_num_sms();
  if (sms <= 0) sms = kNumSMsB200;
  streamk_shape(N, K, sms, &n_tiles, &kb_total, &grid, &max_segs);
  const size_t a = static_cast<size_t>(grid) * max_segs * kBM * 256 * sizeof(float);
  const size_t b = gemm_skinny_workspace_bytes(N, K);
  return a > b ? a : b;
}

extern "C" int opsg_gemm_bf16_streamk(const opsg_bf16* A, int lda, const opsg_bf16* W, int ldw, void* D, int ldd, int M,
                                      int N, int K, const float* bias, const opsg_bf16* residual, int ldr, int act,
                                      int out_mode, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(A && W && D && workspace, "gemm_streamk: null pointer");
  OPSG_CHECK_ARG(M > 0 && M <= kBM && N > 0 && K > 0, "gemm_streamk: needs 0 < M <= 128 (got M=%d N=%d K=%d)", M, N, K);
  OPSG_CHECK_ARG(lda >= K && ldw >= K && ldd >= N, "gemm_streamk: leading dimension too small");
  OPSG_CHECK_ARG((lda % 8) == 0 && (ldw % 8) == 0, "gemm_streamk: lda/ldw must be multiples of 8 elements (TMA)");
  OPSG_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)workspace & 15) == 0,
                 "gemm_streamk: A/W/workspace must be 16-byte aligned");
  OPSG_CHECK_ARG(out_mode == OPSG_OUT_BF16 || out_mode == OPSG_OUT_F32, "gemm_streamk: out_mode must be BF16 or F32");
  OPSG_CHECK_ARG(act >= OPSG_ACT_NONE && act <= OPSG_ACT_RELU, "gemm_streamk: bad activation");
  OPSG_CHECK_ARG(!residual || ldr >= N, "gemm_streamk: ldr too small");
  // K-sliced kernel with the activation slice resident in tensor memory (gemm_skinny.cu); a layout it does not take (N or a
  // leading dimension not a multiple of 4) falls through to stream-K below
  {
    rc = launch_gemm_skinny(A, lda, W, ldw, D, ldd, M, N, K, bias, residual, ldr, act, out_mode, workspace, workspace_bytes,
                            reinterpret_cast<cudaStream_t>(stream));
    if (rc != OPSG_E_UNSUPPORTED) return rc;
  }
  int n_tiles, kb_total, grid, max_segs;
  streamk_shape(N, K, opsg_num_sms(), &n_tiles, &kb_total, &grid, &max_segs);
  OPSG_CHECK_ARG(workspace_bytes >= static_cast<size_t>(grid) * max_segs * kBM * 256 * sizeof(float),
                 "gemm_streamk: workspace too small (%zu bytes)", workspace_bytes);
  CUtensorMap tmA, tmB;
  rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBM, kBK);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, 256, kBK);
  if (rc) return rc;
  GemmParams p;
  p.D = nullptr; p.bias = nullptr; p.residual = nullptr;
  p.M = M; p.N = N; p.K = K; p.ldd = 0; p.ldr = 0; p.bias_along_m = 0; p.act = OPSG_ACT_NONE;
  p.out_mode = OPSG_OUT_F32; p.k_splits = 1; p.m_tiles = 1; p.n_tiles = n_tiles; p.tma_store = 0;
  p.stream_k = 1; p.max_segs = max_segs; p.ws = reinterpret_cast<float*>(workspace);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  using S = GemmSmem<256, 4>;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[device_slot()];
  if (!configured) {
    rc = check_cuda(cudaFuncSetAttribute(gemm_bf16_kernel<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal),
                    "cudaFuncSetAttribute(gemm streamk)");
    if (rc) return rc;
    configured = true;
  }
  launch_kernel(gemm_bf16_kernel<256, 4>, grid, kGemmThreads, S::kTotal, st, tmA, tmB, tmA, p);
  OPSG_CHECK_LAUNCH("gemm_bf16_kernel(stream-K)");
  StreamKReduceParams r;
  r.ws = p.ws; r.D = D; r.bias = bias; r.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  r.M = M; r.N = N; r.ldd = ldd; r.ldr = ldr; r.act = act; r.out_f32 = out_mode == OPSG_OUT_F32;
  r.n_tiles = n_tiles; r.kb_total = kb_total; r.grid = grid; r.max_segs = max_segs;
  const int gy = (M + 3) / 4 < 8 ? (M + 3) / 4 : 8;
  launch_kernel(streamk_reduce_kernel, dim3(n_tiles, gy), 256, 0, st, r);
  OPSG_CHECK_LAUNCH("streamk_reduce_kernel");
  return OPSG_OK;
}

extern "C" int opsg_gemm_bf16(const opsg_bf16* A, int lda, const opsg_bf16* W, int ldw, void* D, int ldd, int M, int N,
                              int K, const float* bias, int bias_along_m, const opsg_bf16* residual, int ldr, int act,
                              int out_mode, int k_splits, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(A && W && D, "gemm: null pointer");
  OPSG_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  OPSG_CHECK_ARG(lda >= K && ldw >= K && ldd >= N, "gemm: leading dimension too small");
  OPSG_CHECK_ARG((lda % 8) == 0 && (ldw % 8) == 0, "gemm: lda/ldw must be multiples of 8 elements (TMA)");
  OPSG_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0, "gemm: A/W must be 16-byte aligned");
  OPSG_CHECK_ARG(out_mode >= OPSG_OUT_BF16 && out_mode <= OPSG_OUT_F32_ATOMIC, "gemm: bad out_mode");
  OPSG_CHECK_ARG(act >= OPSG_ACT_NONE && act <= OPSG_ACT_RELU, "gemm: bad activation");
  OPSG_CHECK_ARG(k_splits >= 1, "gemm: k_splits must be >= 1");
  if (k_splits > 1 && out_mode == OPSG_OUT_F32) {
    OPSG_CHECK_ARG(!bias && !residual && act == OPSG_ACT_NONE, "gemm: split-K partial slices take no bias / residual / activation");
  } else if (k_splits > 1)
    OPSG_CHECK_ARG(out_mode == OPSG_OUT_F32_ATOMIC && !bias && !residual && act == OPSG_ACT_NONE,
                   "gemm: split-K needs OPSG_OUT_F32_ATOMIC and no bias/residual/act");
  OPSG_CHECK_ARG(!residual || ldr >= N, "gemm: ldr too small");

  // large bf16-output problems: CTA-pair kernel (cta_group::2, 256 x 256 tiles; gemm_2cta.cu)
  if (out_mode == OPSG_OUT_BF16 && k_splits == 1) {
    rc = launch_gemm_2cta(A, lda, W, ldw, reinterpret_cast<opsg_bf16*>(D), ldd, M, N, K, bias, bias_along_m, residual, ldr, act,
                          nullptr, reinterpret_cast<cudaStream_t>(stream));
    if (rc != OPSG_E_UNSUPPORTED) return rc;
  }

  // tile width: wide tiles for big problems, narrower ones when the grid would not fill the machine
  const int m_tiles = (M + kBM - 1) / kBM;
  int bn = 256;
  const int sms = opsg_num_sms();
  // (a grid within an eighth of the machine keeps the wider tile: 140 CTAs of 128 x 256 beat 280 of 128 x 128 in two waves)
  while (bn > 32 && (N <= bn / 2 || m_tiles * ((N + bn - 1) / bn) * k_splits < sms - sms / 8)) bn >>= 1;

  CUtensorMap tmA, tmB, tmD;
  rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBM, kBK);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)bn, kBK);
  if (rc) return rc;
  const int tma_store = (out_mode == OPSG_OUT_BF16 && bn >= 64 && (ldd % 8) == 0 && ((uintptr_t)D & 15) == 0) ? 1 : 0;
  if (tma_store) {
    rc = make_tmap_bf16_2d(&tmD, D, (uint64_t)M, (uint64_t)N, (uint64_t)ldd, kBM, 64);
    if (rc) return rc;
  } else {
    tmD = tmA;   // unused by the kernel
  }

  GemmParams p;
  p.D = D; p.bias = bias; p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.M = M; p.N = N; p.K = K; p.ldd = ldd; p.ldr = ldr; p.bias_along_m = bias_along_m; p.act = act;
  p.out_mode = out_mode; p.k_splits = k_splits; p.m_tiles = 0; p.n_tiles = 0; p.tma_store = tma_store;
  p.stream_k = 0; p.max_segs = 0; p.ws = nullptr;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (bn) {
    case 256: return launch_gemm<256, 4>(tmA, tmB, tmD, p, st);
    case 128: return launch_gemm<128, 6>(tmA, tmB, tmD, p, st);
    case 64: return launch_gemm<64, 8>(tmA, tmB, tmD, p, st);
    default: return launch_gemm<32, 8>(tmA, tmB, tmD, p, st);
  }
}
