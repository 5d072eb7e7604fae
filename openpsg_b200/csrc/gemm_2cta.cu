// K6 (large shapes) — D = act(A . W^T + bias + residual) with tcgen05 cta_group::2: a CTA PAIR (two SMs of one TPC,
// cluster 2x1) computes one 256 x 256 output tile.
//
// Why: in cta_group::1 both MMA operands of a 128 x 256 x 16 step come from one SM's shared memory (4 KB of A + 8 KB
// of W^T) and the instruction is paced by that read, 171 cycles instead of the tensor pipe's 128
// (scripts/ubench/mma_cost.cu, profiles/r1_ubench.md) — a 75 % ceiling for every cta_group::1 GEMM.  In a pair each CTA
// supplies its own 128 rows of A and only HALF of the W^T tile (128 of the 256 N rows): 8 KB per 128-cycle step.
//
// Roles per CTA (384 threads), same as gemm.cu:
//   warp 0  TMA producer : own A tile [128 x 64] + own half of W^T [128 x 64] per stage (32 KB), completion bytes of BOTH
//                          CTAs are signalled on the LEADER's full barrier (cta rank 0; peer bit of the address cleared)
//   warp 1  MMA issuer   : leader only — tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16), accumulator rows 0-127 in
//                          the leader's TMEM, rows 128-255 in the peer's; tcgen05.commit multicasts to both CTAs' barriers
//   warp 2  TMEM allocator (tcgen05.alloc.cta_group::2, both CTAs)
//   warps 4-11 epilogue  : own 128 rows: tcgen05.ld -> bias / residual / activation -> bf16 -> swizzled smem -> TMA store;
//                          the accumulator stage is handed back by arriving on the LEADER's tmem_empty barrier (remote
//                          mbarrier arrive from the peer)
#include "common.cuh"
#include "host_util.h"

namespace opsg {
namespace g2 {

constexpr int kBM = 128;           // rows per CTA (256 per pair)
constexpr int kBN = 256;           // tile N (each CTA stages 128 of the 256 W^T rows)
constexpr int kBK = 64;
constexpr int kThreads = 384;
constexpr int kEpiThreads = 256;
constexpr int kStages = 6;
constexpr int kABytes = kBM * kBK * 2;          // 16384
constexpr int kBBytes = (kBN / 2) * kBK * 2;    // 16384
constexpr int kStageBytes = kABytes + kBBytes;  // 32768 per CTA
constexpr int kSlabBytes = kBM * 128;           // [128 rows x 64 bf16] staging slab
constexpr int kSmemTotal = kStages * kStageBytes + 2 * kSlabBytes + 1024 + 1024;
static_assert(kSmemTotal <= 232448, "shared memory budget exceeded");
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the even (leader) CTA

struct Params {
  const float* bias;
  const __nv_bfloat16* residual;
  int M, N, K;
  int ldr;
  int bias_along_m;
  int act;
  int m2_tiles, n_tiles;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result, uint32_t ncols) {   // whole warp, both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA tile load whose completion bytes land on the LEADER CTA's mbarrier (same smem offset, peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
// D[tmem, both CTAs] (+)= A[smem, both CTAs] * B[smem halves of both CTAs]
__device__ __forceinline__ void umma_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive (once complete) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_2cta_mcast(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
               : "memory");
}
// arrive on the barrier at this smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmD, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint8_t* staging = smem + kStages * kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + 2 * kSlabBytes);   // used in the leader CTA only
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2] used in the leader CTA only (both CTAs' epilogues arrive there)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * (kEpiThreads / 32));     // one arrival per epilogue warp of both CTAs
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_base_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // barriers of both CTAs initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int num_pairs = gridDim.x >> 1;
  const int pair = blockIdx.x >> 1;
  const int total_tiles = p.m2_tiles * p.n_tiles;
  const int kb_total = (p.K + kBK - 1) / kBK;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < total_tiles; tile += num_pairs) {
      const int n_t = tile % p.n_tiles, m2_t = tile / p.n_tiles;
      const int row_a = m2_t * 2 * kBM + static_cast<int>(rank) * kBM;           // this CTA's 128 rows of A
      const int row_b = n_t * kBN + static_cast<int>(rank) * (kBN / 2);          // this CTA's half of the W^T tile
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes);         // bytes of both CTAs land on this barrier
          tma_load_2d_2cta(smem_a + stage * kABytes, &tmA, &full_bar[stage], kb * kBK, row_a);
          tma_load_2d_2cta(smem_b + stage * kBBytes, &tmB, &full_bar[stage], kb * kBK, row_b);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader CTA) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(2 * kBM, kBN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < total_tiles; tile += num_pairs) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kBN;
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t a_desc = umma_desc_k_sw128(smem_u32(smem_a + stage * kABytes));
          const uint64_t b_desc = umma_desc_k_sw128(smem_u32(smem_b + stage * kBBytes));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_ss_2cta(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit_2cta_mcast(&empty_bar[stage]);                     // frees the stage in both CTAs
          if (kb + 1 == kb_total) tc_commit_2cta_mcast(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    constexpr int NSLAB = kBN / 64;
    const int ew = warp - 4;
    const int q = ew & 3;
    const int half = ew >> 2;
    const int row_in_tile = q * 32 + lane;
    const bool elected = (threadIdx.x == 4 * 32);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t buf = 0;
    const __nv_bfloat16* resid = p.residual;

    for (int tile = pair; tile < total_tiles; tile += num_pairs) {
      const int n_t = tile % p.n_tiles, m2_t = tile / p.n_tiles;
      const int row0 = m2_t * 2 * kBM + static_cast<int>(rank) * kBM;
      const int row = row0 + row_in_tile;
      const bool row_ok = row < p.M;
      const float bias_m = (p.bias && p.bias_along_m && row_ok) ? p.bias[row] : 0.f;

      // Residual slab [128 rows x 64 cols] of this CTA: loaded COOPERATIVELY (8 consecutive threads = one 128-byte row
      // segment, 4 x 16 bytes per thread) one slab ahead into registers, parked in the staging slab, then every thread
      // picks up its own row from shared memory.  (Each thread loading 64 bytes of its own row — the first version —
      // made every load instruction touch 32 different rows: 128 L1 wavefronts per warp and slab instead of 16, and the
      // residual GEMMs ran at 790 TFLOP/s while the plain ones reached 1540.)
      const int et = threadIdx.x - 4 * 32;                       // 0..255 inside the epilogue
      const bool res_fast = resid && (p.ldr % 8) == 0 && (p.N % 8) == 0 && ((reinterpret_cast<uintptr_t>(resid) & 15) == 0);
      uint4 rres[4];
      auto fetch_residual = [&](int slab) {
        if (!res_fast) return;
        const int c0 = n_t * kBN + slab * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int idx = j * kEpiThreads + et, r = idx >> 3, ch = idx & 7;
          const int grow = row0 + r, gcol = c0 + ch * 8;
          rres[j] = (grow < p.M && gcol + 8 <= p.N)
                        ? __ldg(reinterpret_cast<const uint4*>(resid + static_cast<size_t>(grow) * p.ldr + gcol))
                        : make_uint4(0, 0, 0, 0);
        }
      };
      fetch_residual(0);

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kBN;

#pragma unroll 1
      for (int slab = 0; slab < NSLAB; ++slab) {
        const int col0 = n_t * kBN + slab * 64 + half * 32;
        float f[32];
        {
          uint32_t v[32];
          tmem_ld32(taddr + slab * 64 + half * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        }
        if (slab == NSLAB - 1) {            // accumulator fully read -> hand the TMEM stage back to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
        }
        const bool col_ok = col0 < p.N;
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
          if (resid && !res_fast && row_ok) {
            const __nv_bfloat16* r = resid + static_cast<size_t>(row) * p.ldr + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) f[j] += __bfloat162float(r[j]);
          }
        }
        uint8_t* rowp = staging + buf * kSlabBytes + row_in_tile * 128;
        // staging slab `buf`: its previous TMA store (two slabs ago) must have drained before we overwrite it
        if (elected) tma_store_wait_read<1>();
        named_bar_sync(1, kEpiThreads);
        if (res_fast) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {                          // park the prefetched residual slab (swizzled like the output)
            const int idx = j * kEpiThreads + et, r = idx >> 3, ch = idx & 7;
            *reinterpret_cast<uint4*>(staging + buf * kSlabBytes + r * 128 + ((ch ^ (r & 7)) * 16)) = rres[j];
          }
          if (slab + 1 < NSLAB) fetch_residual(slab + 1);
          named_bar_sync(3, kEpiThreads);
          if (col_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 u = *reinterpret_cast<const uint4*>(rowp + (((half * 4 + g) ^ (row_in_tile & 7)) * 16));
              f[g * 8 + 0] += bf16_lo(u.x); f[g * 8 + 1] += bf16_hi(u.x);
              f[g * 8 + 2] += bf16_lo(u.y); f[g * 8 + 3] += bf16_hi(u.y);
              f[g * 8 + 4] += bf16_lo(u.z); f[g * 8 + 5] += bf16_hi(u.z);
              f[g * 8 + 6] += bf16_lo(u.w); f[g * 8 + 7] += bf16_hi(u.w);
            }
          }
        }
        if (col_ok) {
          if (p.act == OPSG_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_erf_fast(f[j]);
          } else if (p.act == OPSG_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
        }
        {
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
          if (n_t * kBN + slab * 64 < p.N && row0 < p.M)
            tma_store_2d(staging + buf * kSlabBytes, &tmD, n_t * kBN + slab * 64, row0);
          tma_store_commit();
        }
        buf ^= 1;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (elected) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer may still be reading our TMEM / arriving on our barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace g2

// Launch helper used by opsg_gemm_bf16 (gemm.cu).  Returns OPSG_E_UNSUPPORTED when the shape should use the 1-CTA kernel.
int launch_gemm_2cta(const opsg_bf16* A, int lda, const opsg_bf16* W, int ldw, opsg_bf16* D, int ldd, int M, int N, int K,
                     const float* bias, int bias_along_m, const opsg_bf16* residual, int ldr, int act, cudaStream_t stream) {
  using namespace g2;
  const int sms = opsg_num_sms();
  const int m2_tiles = (M + 2 * kBM - 1) / (2 * kBM);
  const int n_tiles = (N + kBN - 1) / kBN;
  if (N < kBN || m2_tiles * n_tiles < sms / 2 || (ldd % 8) != 0 || (reinterpret_cast<uintptr_t>(D) & 15) != 0)
    return OPSG_E_UNSUPPORTED;
  CUtensorMap tmA, tmB, tmD;
  int rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBM, kBK);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, kBN / 2, kBK);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmD, D, (uint64_t)M, (uint64_t)N, (uint64_t)ldd, kBM, 64);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    rc = check_cuda(cudaFuncSetAttribute(gemm2_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal),
                    "cudaFuncSetAttribute(gemm 2cta)");
    if (rc) return rc;
    configured = true;
  }
  Params p;
  p.bias = bias; p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.M = M; p.N = N; p.K = K; p.ldr = ldr; p.bias_along_m = bias_along_m; p.act = act;
  p.m2_tiles = m2_tiles; p.n_tiles = n_tiles;
  int pairs = sms / 2;
  if (m2_tiles * n_tiles < pairs) pairs = m2_tiles * n_tiles;
  gemm2_bf16_kernel<<<2 * pairs, kThreads, kSmemTotal, stream>>>(tmA, tmB, tmD, p);
  OPSG_CHECK_LAUNCH("gemm2_bf16_kernel");
  return OPSG_OK;
}

}  // namespace opsg
