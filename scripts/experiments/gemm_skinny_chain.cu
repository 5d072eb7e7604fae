// K6c — chain of small-M weight-streaming GEMMs in ONE persistent kernel (LLM decode: M = selected pairs <= 128; the
// out_proj -> fc1 -> fc2 -> next layer's q/k/v run of a decoder layer, or any single one of them; reference v4:305-312 ->
// HF OPT / Llama decoder layers).
//
// What a decode step costs is not the weight stream (157 MB per OPT-2.7B layer = 24 us at HBM speed) but what surrounds
// it (profiles/r2_decode_timeline.md): per GEMM a kernel boundary (~2 us), ~3 us until the activation slice has arrived, a
// pipeline that starts empty, a second kernel that sums the K slices, and a LayerNorm kernel in front of two of the four.
// This kernel keeps the K-sliced layout of gemm_skinny.cu (CTA = (K slice, lane); the slice of the activations resident
// in TENSOR MEMORY as the A operand of TS-mode MMAs, only weights streamed) and removes the surroundings:
//   * PHASES.  A launch carries up to 4 GEMMs ("phases") that depend on each other.  The weight-producer warp walks all
//     of them without ever waiting for anything but a free ring slot -- weights are constants -- so while a phase
//     drains, reduces and the next one fetches its activations, the 128 KB ring fills with the NEXT phase's weights: the
//     HBM stream does not stop at a phase boundary.  A phase boundary is a counter in global memory (finalised tiles of the
//     previous phase, release / acquire), not a kernel boundary.
//   * REDUCTION IN THE KERNEL.  Every CTA writes its fp32 partial tile to the L2-resident workspace with TMA stores (the
//     scattered 16-byte stores of gemm_skinny.cu slowed its own weight stream by a third) and bumps the tile's counter;
//     the CTA that delivers the LAST slice hands the tile to its four finaliser warps, which sum the S partials in slice
//     order (deterministic), apply bias / activation / residual and write bf16 / fp32.  Nobody waits for anybody: no second
//     kernel, no cluster handshake (a DSMEM exchange inside 4/8/16-CTA clusters was built and measured 30-60 % SLOWER
//     than two kernels: profiles/r2_decode_chain.md).
//   * LAYERNORM / RMSNORM FOLDED INTO THE ACTIVATION LOAD.  The activation K-blocks pass through registers on their way
//     to tensor memory (thread = row); a phase may ask for y = (x - mean) rstd gamma + beta there.  The row statistics come
//     from the PRODUCING phase's finalisers, which emit (sum, sum of squares) of the bf16-rounded values they write per
//     (row, tile); the consumer adds a row's partials in tile order.  No LayerNorm kernel, no normalised copy in HBM.
//   * 16 KB weight requests (3-D TMA box = two K-blocks of a 64-row tile): one issuing warp sustains 6.4 TB/s with these
//     against 4.4 TB/s with 8 KB ones (profiles/r2_decode_timeline.md).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host_util.h"

namespace opsg {
namespace skq {

constexpr int kBN = 64;                 // weight rows per n-tile (= MMA N)
constexpr int kBK = 64;                 // 64 bf16 = 128 B = one swizzle span
constexpr int kBlockBytes = kBN * 128;  // one K-block of one n-tile
constexpr int kKreq = 2;                // K-blocks per weight request
constexpr int kSlotBytes = kKreq * kBlockBytes;
constexpr int kThreads = 384;           // warps: 0 W producer, 1 MMA, 2-5 A copy + epilogue, 6 A producer, 7 partial store, 8-11 finalisers
constexpr int kMaxStages = 12;
constexpr int kMaxKS = 12;              // K-blocks per slice: 12 x 32 TMEM columns of A + 2 x 64 accumulator columns = 512
constexpr int kAStages = 2;
constexpr int kFq = 4;                  // depth of the store-warp -> finaliser queue
constexpr int kSmemLimit = 232448;
constexpr int kBarrierBytes = 1024;
constexpr int kAccCols = 2 * kBN;
constexpr int kMaxPhases = OPSG_CHAIN_MAX_PHASES;
constexpr int kCountersPerPhase = 1024; // tile counters (N <= 65536)
constexpr int kNormBytes = kMaxKS * kBK * 2 * 4;   // gamma / beta of one K slice

struct alignas(64) Phase {
  CUtensorMap tmA;      // activations [M, K] bf16, box {64, MR}
  CUtensorMap tmW;      // weights, K-block view {64, N, K/64}, box {64, 64, 2}
  CUtensorMap tmP;      // partial workspace fp32 {N, M, S}, box {32, MR, 1}: stores of this CTA's slice
  CUtensorMap tmF;      // the same workspace, box {32, R, S}, R = ceil(M / S): loads of this CTA's rows of all slices
  const float* ws;      // the same workspace for the finalisers' loads
  const float* bias;
  const __nv_bfloat16* residual;
  void* D;
  const float* norm_gamma;
  const float* norm_beta;
  const float2* stats_in;   // [M][n_stats_in] partial (sum, sumsq) of the A rows (norm != 0)
  float2* stats_out;        // [M][n_tiles] partial (sum, sumsq) of the rows this phase writes, or null
  int M, N, K, ldd, ldr, act, out_f32;
  int S, KS, G, n_tiles, kb_total;
  int norm;                 // 0 none, 1 LayerNorm, 2 RMSNorm, applied to A on its way to tensor memory
  int n_stats_in;
  float norm_eps;
  int wait_prev;            // A / stats_in / residual are written by the previous phase of this launch
};

#ifdef OPSG_TRACE
// development build (make trace): %globaltimer stamps, trace[cta][tile < 32][16]
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TR(tile, slot) do { if (L.trace && (tile) < 32) L.trace[(static_cast<size_t>(blockIdx.x) * 32 + (tile)) * 16 + (slot)] = gtime(); } while (0)
#else
#define TR(tile, slot) do { } while (0)
#endif

struct Launch {
  long long* trace;
  Phase ph[kMaxPhases];
  int n_phases, MR, stages;
  int fin_bytes;            // one finaliser buffer: 2 halves x max over phases of S x R x 128 B
  int* counters;            // [kMaxPhases][kCountersPerPhase] zero between launches (finalisers reset what they used)
  int* sync;                // [kMaxPhases][2]: finalised tiles of the phase / CTAs that have seen them all; zero between launches
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1,
                                            int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* smem_src, const CUtensorMap* map, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// this CTA's part of a phase
struct Part {
  int slice, g, kb0, nkb, nreq;
  bool active;
};
__device__ __forceinline__ Part part_of(const Phase& P) {
  Part t;
  t.slice = blockIdx.x % P.S;
  t.g = blockIdx.x / P.S;
  t.active = t.g < P.G;
  t.kb0 = t.slice * P.KS;
  t.nkb = min(P.KS, P.kb_total - t.kb0);
  t.nreq = (t.nkb + kKreq - 1) / kKreq;
  return t;
}

__global__ void __launch_bounds__(kThreads, 1) skinny_chain_kernel(const __grid_constant__ Launch L) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes_kb = L.MR * 128;                        // one activation K-block / one 32-column fp32 half tile
  uint8_t* smem_a = smem;                                   // [kAStages][MR][128 B]
  uint8_t* smem_w = smem_a + kAStages * a_bytes_kb;         // [stages][16 KB]
  uint8_t* smem_p = smem_w + L.stages * kSlotBytes;         // [2 buffers][2 halves][MR][128 B] fp32 partial tile for the TMA store
  uint8_t* smem_f = smem_p + 4 * a_bytes_kb;                // [2 buffers][2 halves][S][R][128 B] fp32 partials of this CTA's rows
  float* smem_norm = reinterpret_cast<float*>(smem_f + 2 * L.fin_bytes);   // [2][kMaxKS * 64] gamma, beta of the K slice
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_norm) + kNormBytes);
  uint64_t* full_bar = bars;                               // [stages]   weight request landed
  uint64_t* empty_bar = full_bar + kMaxStages;             // [stages]   its MMAs have completed
  uint64_t* a_full = empty_bar + kMaxStages;               // [kAStages] activation K-block landed in the ring
  uint64_t* a_empty = a_full + kAStages;                   // [kAStages] ... and has been copied to TMEM
  uint64_t* a_ready = a_empty + kAStages;                  // [kMaxKS]   K-block kb of A is in TMEM (once per phase)
  uint64_t* tmem_full = a_ready + kMaxKS;                  // [2]
  uint64_t* tmem_empty = tmem_full + 2;                    // [2]
  uint64_t* stage_full = tmem_empty + 2;                   // [2] partial tile written to the staging buffer
  uint64_t* stage_free = stage_full + 2;                   // [2] ... and stored
  uint64_t* fq_full = stage_free + 2;                      // [kFq] finaliser queue
  uint64_t* fq_empty = fq_full + kFq;                      // [kFq]
  uint64_t* fin_full = fq_empty + kFq;                     // [2] all slices of a tile's rows landed in the finaliser buffer
  uint64_t* fin_empty = fin_full + 2;                      // [2] ... and have been summed
  uint64_t* fin_hdr = fin_empty + 2;                       // [2] fin_item written (before the data: bias / residual loads start early)
  int* fq_item = reinterpret_cast<int*>(fin_hdr + 2);      // [kFq][2] (phase, tile); phase < 0: end
  int* fin_item = fq_item + 2 * kFq;                       // [2][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fin_item + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < L.n_phases; ++i) {
      tma_prefetch_desc(&L.ph[i].tmA);
      tma_prefetch_desc(&L.ph[i].tmW);
      tma_prefetch_desc(&L.ph[i].tmP);
      tma_prefetch_desc(&L.ph[i].tmF);
    }
    for (int s = 0; s < L.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAStages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 4);
    }
    for (int s = 0; s < kMaxKS; ++s) mbar_init(&a_ready[s], 4);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
      mbar_init(&stage_full[s], 4);
      mbar_init(&stage_free[s], 1);
    }
    for (int s = 0; s < kFq; ++s) {
      mbar_init(&fq_full[s], 1);
      mbar_init(&fq_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&fin_full[s], 1);
      mbar_init(&fin_hdr[s], 1);
      mbar_init(&fin_empty[s], 3);                           // three summing warps
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a = tmem_base + kAccCols;
  pdl_wait_then_trigger();
  if (threadIdx.x == 0) TR(31, 0);

  if (warp == 0) {
    // ===================== weight producer: all phases back to back, waits for ring slots only =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int pi = 0; pi < L.n_phases; ++pi) {
      const Phase& P = L.ph[pi];
      const Part t = part_of(P);
      if (!t.active) continue;
      for (int n_t = t.g; n_t < P.n_tiles; n_t += P.G) {
        for (int q = 0; q < t.nreq; ++q) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one_sync()) {
            mbar_expect_tx(&full_bar[stage], kSlotBytes);
            tma_load_3d(smem_w + stage * kSlotBytes, &P.tmW, &full_bar[stage], 0, n_t * kBN, t.kb0 + q * kKreq);
          }
          __syncwarp();
          if (++stage == L.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 6) {
    // ===================== activation producer =====================
    int cnt = 0;                                             // K-blocks sent through the ring so far
    for (int pi = 0; pi < L.n_phases; ++pi) {
      const Phase& P = L.ph[pi];
      const Part t = part_of(P);
      if (!t.active) continue;
      if (P.wait_prev) {
        // every tile of the previous phase has been finalised (its output = this phase's A / residual / statistics)
        if (lane == 0) {
          int* done = L.sync + 2 * (pi - 1);
          const int need = L.ph[pi - 1].n_tiles * L.ph[pi - 1].S;      // every slice owner finalises its rows of every tile
          const long long t0 = clock64();
          while (ld_acquire_gpu(done) < need) {
            if (clock64() - t0 > OPSG_WAIT_LIMIT_CYCLES) {
              printf("opsg: chain phase %d wait timed out (block %d: %d of %d tiles)\n", pi, blockIdx.x, ld_acquire_gpu(done), need);
              __trap();
            }
          }
          fence_proxy_async_all();                           // the TMA loads below read what generic stores wrote
          // the last CTA to get here clears the two counters for the next launch
          const int seen = atomicAdd(done + 1, 1);
          if (seen == P.S * P.G - 1) {
            done[0] = 0;
            done[1] = 0;
          }
        }
        __syncwarp();
      }
      for (int kb = 0; kb < t.nkb; ++kb, ++cnt) {
        const int b = cnt % kAStages;
        mbar_wait(&a_empty[b], ((cnt / kAStages) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&a_full[b], static_cast<uint32_t>(a_bytes_kb));
          tma_load_2d(smem_a + b * a_bytes_kb, &P.tmA, &a_full[b], (t.kb0 + kb) * kBK, 0);   // rows >= M, columns >= K: zeros
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, kBN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t a_parity = 0;                                   // a_ready completes once per phase this CTA takes part in
#ifdef OPSG_TRACE
    int trace_tile = 0;
#endif
    for (int pi = 0; pi < L.n_phases; ++pi) {
      const Phase& P = L.ph[pi];
      const Part t = part_of(P);
      if (!t.active) continue;
      bool first_tile = true;
      for (int n_t = t.g; n_t < P.n_tiles; n_t += P.G) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBN;
        for (int q = 0; q < t.nreq; ++q) {
          const int kbq = q * kKreq;
          const int nb = min(kKreq, t.nkb - kbq);
          if (first_tile)
            for (int j = 0; j < nb; ++j) mbar_wait(&a_ready[kbq + j], a_parity);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t w_addr = smem_u32(smem_w + stage * kSlotBytes);
            for (int j = 0; j < nb; ++j) {
              const uint64_t b_desc = umma_desc_k_sw128(w_addr + j * kBlockBytes);
              const int kb = kbq + j;
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k)               // A: 8 TMEM columns (16 packed bf16) per K = 16 step
                umma_ts(d_tmem, tmem_a + kb * 32 + k * 8, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            tc_commit(&empty_bar[stage]);
            if (q + 1 == t.nreq) tc_commit(&tmem_full[acc]);
          }
#ifdef OPSG_TRACE
          if (q + 1 == t.nreq && lane == 0) { TR(trace_tile, 11); ++trace_tile; }
          if (q == 0 && lane == 0) TR(trace_tile, 12);
#endif
          __syncwarp();
          if (++stage == L.stages) { stage = 0; phase ^= 1; }
        }
        first_tile = false;
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      a_parity ^= 1;
    }
  } else if (warp >= 2 && warp < 6) {
    // ===================== warps 2-5: A slice -> (norm) -> TMEM, then accumulators -> staging =====================
    const int q = warp & 3;                                  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    const int t128 = threadIdx.x - 64;                       // 0..127 over the four warps
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    int cnt = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;                                              // tiles written to the staging buffers so far
    for (int pi = 0; pi < L.n_phases; ++pi) {
      const Phase& P = L.ph[pi];
      const Part t = part_of(P);
      if (!t.active) continue;
      float mean = 0.f, rstd = 1.f;
      for (int kb = 0; kb < t.nkb; ++kb, ++cnt) {
        const int b = cnt % kAStages;
        mbar_wait(&a_full[b], (cnt / kAStages) & 1);
        if (kb == 0 && P.norm) {
          // The first K-block has landed: the producer has seen the previous phase complete.  Row statistics = the
          // producing tiles' partials added in tile order; gamma / beta of this K slice go to shared memory.
          float s = 0.f, ss = 0.f;
          if (row < P.M) {
            const float2* st = P.stats_in + static_cast<size_t>(row) * P.n_stats_in;
            for (int i = 0; i < P.n_stats_in; ++i) {
              const float2 u = __ldcg(st + i);
              s += u.x;
              ss += u.y;
            }
          }
          const float inv_n = 1.f / static_cast<float>(P.K);
          if (P.norm == 1) {
            mean = s * inv_n;
            rstd = rsqrtf(fmaxf(ss * inv_n - mean * mean, 0.f) + P.norm_eps);
          } else {
            mean = 0.f;
            rstd = rsqrtf(ss * inv_n + P.norm_eps);
          }
          named_bar_sync(2, 128);                            // the previous phase's readers of smem_norm are done
          for (int i = t128; i < t.nkb * kBK; i += 128) {
            smem_norm[i] = __ldg(P.norm_gamma + t.kb0 * kBK + i);
            smem_norm[kMaxKS * kBK + i] = P.norm_beta ? __ldg(P.norm_beta + t.kb0 * kBK + i) : 0.f;
          }
          named_bar_sync(2, 128);
        }
        uint32_t v[32];
        if (row < L.MR) {
          const uint8_t* src = smem_a + b * a_bytes_kb + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {                      // undo the 128-byte swizzle: 16-byte chunk j sits at j ^ (row % 8)
            const uint4 u = *reinterpret_cast<const uint4*>(src + ((j ^ (row & 7)) * 16));
            v[4 * j] = u.x; v[4 * j + 1] = u.y; v[4 * j + 2] = u.z; v[4 * j + 3] = u.w;
          }
          if (P.norm) {
            const float* gm = smem_norm + kb * kBK;
            const float* bt = smem_norm + kMaxKS * kBK + kb * kBK;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float x0 = (bf16_lo(v[j]) - mean) * rstd * gm[2 * j] + bt[2 * j];
              const float x1 = (bf16_hi(v[j]) - mean) * rstd * gm[2 * j + 1] + bt[2 * j + 1];
              v[j] = row < P.M ? pack_bf16x2(x0, x1) : 0u;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        uint32_t lo[16], hi[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { lo[j] = v[j]; hi[j] = v[16 + j]; }
        tmem_st16(tmem_a + lane_base + kb * 32, lo);
        tmem_st16(tmem_a + lane_base + kb * 32 + 16, hi);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&a_empty[b]);
          mbar_arrive(&a_ready[kb]);
        }
      }
      if (lane == 0)
        for (int kb = t.nkb; kb < kMaxKS; ++kb) mbar_arrive(&a_ready[kb]);     // every a_ready flips once per phase
      if (threadIdx.x == 64) TR(31, 1);

      for (int n_t = t.g; n_t < P.n_tiles; n_t += P.G, ++it) {
        const int sb = it & 1;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_base + acc * kBN;
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr, v0);
        tmem_ld32(taddr + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        if (threadIdx.x == 64) TR(it, 1);
        mbar_wait(&stage_free[sb], ((it >> 1) & 1) ^ 1);
        if (threadIdx.x == 64) TR(it, 2);
        if (row < L.MR) {                                    // two [MR x 128 B] halves in the TMA store's swizzled layout
          uint8_t* d0 = smem_p + (sb * 2) * a_bytes_kb + row * 128;
          uint8_t* d1 = d0 + a_bytes_kb;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            *reinterpret_cast<uint4*>(d0 + ((j ^ (row & 7)) * 16)) = make_uint4(v0[4 * j], v0[4 * j + 1], v0[4 * j + 2], v0[4 * j + 3]);
            *reinterpret_cast<uint4*>(d1 + ((j ^ (row & 7)) * 16)) = make_uint4(v1[4 * j], v1[4 * j + 1], v1[4 * j + 2], v1[4 * j + 3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&stage_full[sb]);
      }
    }
  } else if (warp == 7) {
    // ===================== partial-tile store (one tile of look-ahead), tile counters, hand-over to the finalisers ==========
    int it = 0;
    int fq = 0;
    int prev_pi = -1, prev_nt = 0, prev_sb = 0;              // tile whose store has been issued but not yet published
    auto publish = [&](bool drain) {
      // the store of (prev_pi, prev_nt) is complete and visible: free its staging buffer, count this slice, queue the tile
      if (prev_pi < 0) return;
      if (lane == 0) {
        if (drain) tma_store_wait_all<0>(); else tma_store_wait_all<1>();
        mbar_arrive(&stage_free[prev_sb]);
        TR(fq, 4);
        fence_proxy_async_all();
        red_release_gpu_add(L.counters + prev_pi * kCountersPerPhase + prev_nt, 1);
        TR(fq, 5);
      }
      const int slot = fq % kFq;
      mbar_wait(&fq_empty[slot], ((fq / kFq) & 1) ^ 1);
      if (lane == 0) TR(fq, 6);
      if (lane == 0) {
        fq_item[2 * slot] = prev_pi;
        fq_item[2 * slot + 1] = prev_nt;
        mbar_arrive(&fq_full[slot]);
      }
      __syncwarp();
      ++fq;
      prev_pi = -1;
    };
    for (int pi = 0; pi < L.n_phases; ++pi) {
      const Phase& P = L.ph[pi];
      const Part t = part_of(P);
      if (!t.active) continue;
      for (int n_t = t.g; n_t < P.n_tiles; n_t += P.G, ++it) {
        const int sb = it & 1;
        mbar_wait(&stage_full[sb], (it >> 1) & 1);
        if (lane == 0) TR(it, 3);
        if (lane == 0) {
          tma_store_3d(smem_p + (sb * 2) * a_bytes_kb, &P.tmP, n_t * kBN, 0, t.slice);
          if (n_t * kBN + 32 < P.N) tma_store_3d(smem_p + (sb * 2 + 1) * a_bytes_kb, &P.tmP, n_t * kBN + 32, 0, t.slice);
          tma_store_commit();
        }
        __syncwarp();
        publish(false);                                      // the PREVIOUS tile: its store ran under this tile's wait
        prev_pi = pi;
        prev_nt = n_t;
        prev_sb = sb;
      }
      publish(true);                                         // the next phase cannot start before this one is fully published
    }
    {                                                        // end marker
      const int slot = fq % kFq;
      mbar_wait(&fq_empty[slot], ((fq / kFq) & 1) ^ 1);
      if (lane == 0) {
        fq_item[2 * slot] = -1;
        fq_item[2 * slot + 1] = 0;
        mbar_arrive(&fq_full[slot]);
      }
      __syncwarp();
    }
  } else if (warp == 8) {
    // ===================== finaliser loader =====================
    // Every CTA that contributed a slice to a tile finalises ITS share of the tile's rows (rows [slice R, (slice + 1) R),
    // R = ceil(M / S)).  Once all S slices are visible this warp pulls those rows of every slice into shared memory with
    // two TMA loads: per-thread loads kept ~4 KB in flight per SM and a tile took 9 us (profiles/r2_decode_chain.md).
    for (int fq = 0;; ++fq) {
      const int slot = fq % kFq;
      mbar_wait(&fq_full[slot], (fq / kFq) & 1);
      const int pi = fq_item[2 * slot], n_t = fq_item[2 * slot + 1];
      __syncwarp();
      if (lane == 0) mbar_arrive(&fq_empty[slot]);
      const int b = fq & 1;
      mbar_wait(&fin_empty[b], ((fq >> 1) & 1) ^ 1);
      if (pi < 0) {
        if (lane == 0) {
          fin_item[2 * b] = -1;
          mbar_arrive(&fin_hdr[b]);
        }
        break;
      }
      if (lane == 0) {
        TR(fq, 7);
        fin_item[2 * b] = pi;
        fin_item[2 * b + 1] = n_t;
        mbar_arrive(&fin_hdr[b]);
        const Phase& P = L.ph[pi];
        const int* counter = L.counters + pi * kCountersPerPhase + n_t;
        const long long t0 = clock64();
        while (ld_acquire_gpu(counter) < P.S) {
          if (clock64() - t0 > OPSG_WAIT_LIMIT_CYCLES) {
            printf("opsg: chain finaliser timed out (block %d phase %d tile %d: %d of %d slices)\n", blockIdx.x, pi, n_t,
                   ld_acquire_gpu(counter), P.S);
            __trap();
          }
        }
        fence_proxy_async_all();
        TR(fq, 8);
        const int rows_per = (P.M + P.S - 1) / P.S;
        const int half_bytes = P.S * rows_per * 128;
        const bool two = n_t * kBN + 32 < P.N;
        mbar_expect_tx(&fin_full[b], static_cast<uint32_t>(two ? 2 * half_bytes : half_bytes));
        uint8_t* dst = smem_f + b * L.fin_bytes;
        tma_load_3d(dst, &P.tmF, &fin_full[b], n_t * kBN, (blockIdx.x % P.S) * rows_per, 0);
        if (two) tma_load_3d(dst + half_bytes, &P.tmF, &fin_full[b], n_t * kBN + 32, (blockIdx.x % P.S) * rows_per, 0);
      }
      __syncwarp();
    }
  } else {
    // ===================== warps 9-11: sum the slices in slice order (deterministic), epilogue =====================
    const int t96 = threadIdx.x - 288;
    for (int fq = 0;; ++fq) {
      const int b = fq & 1;
      mbar_wait(&fin_hdr[b], (fq >> 1) & 1);
      const int pi = fin_item[2 * b], n_t = fin_item[2 * b + 1];
      if (pi < 0) break;
      const Phase& P = L.ph[pi];
      const int slice = blockIdx.x % P.S;
      const int rows_per = (P.M + P.S - 1) / P.S;
      const int row_begin = slice * rows_per, row_end = min(P.M, row_begin + rows_per);
      const int half_bytes = P.S * rows_per * 128;
      const uint8_t* src = smem_f + b * L.fin_bytes;
      // 16 lanes (four columns each) per row, 6 rows per pass over the 96 threads, 4 passes per batch: the batch's bias /
      // residual loads are issued together (an L2 round trip each when issued one by one)
      const int c16 = t96 & 15, sub = t96 >> 4;
      const int col = n_t * kBN + c16 * 4;
      const bool col_live = col < P.N;                       // N % 4 == 0 (host)
      const uint8_t* src_c = src + (c16 >> 3) * half_bytes + (c16 & 7) * 16;
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (P.bias && col_live) bv = __ldg(reinterpret_cast<const float4*>(P.bias + col));
      for (int r0 = 0; r0 < rows_per; r0 += 24) {
        uint2 rv[4];
        bool live[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int r = r0 + h * 6 + sub, row = row_begin + r;
          live[h] = r < rows_per && row < row_end && col_live;
          rv[h] = make_uint2(0u, 0u);
          if (live[h] && P.residual)
            rv[h] = __ldcg(reinterpret_cast<const uint2*>(P.residual + static_cast<size_t>(row) * P.ldr + col));
        }
        if (r0 == 0) {
          mbar_wait(&fin_full[b], (fq >> 1) & 1);              // the partial rows have landed
          if (t96 == 0) TR(fq, 13);
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int r = r0 + h * 6 + sub, row = row_begin + r;
          if (r0 + h * 6 >= rows_per) break;                 // warp-uniform
          float g[4] = {0.f, 0.f, 0.f, 0.f};
          if (live[h]) {
            for (int s0 = 0; s0 < P.S; ++s0) {
              const float4 u = *reinterpret_cast<const float4*>(src_c + (s0 * rows_per + r) * 128);
              g[0] += u.x; g[1] += u.y; g[2] += u.z; g[3] += u.w;
            }
            g[0] += bv.x; g[1] += bv.y; g[2] += bv.z; g[3] += bv.w;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (P.act == OPSG_ACT_GELU) g[e] = gelu_erf(g[e]);
              else if (P.act == OPSG_ACT_RELU) g[e] = fmaxf(g[e], 0.f);
            }
            g[0] += bf16_lo(rv[h].x); g[1] += bf16_hi(rv[h].x); g[2] += bf16_lo(rv[h].y); g[3] += bf16_hi(rv[h].y);
            if (P.out_f32) {
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(P.D) + static_cast<size_t>(row) * P.ldd + col) =
                  make_float4(g[0], g[1], g[2], g[3]);
            } else {
              uint2 o;
              o.x = pack_bf16x2(g[0], g[1]);
              o.y = pack_bf16x2(g[2], g[3]);
              *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(P.D) + static_cast<size_t>(row) * P.ldd + col) = o;
              g[0] = bf16_lo(o.x); g[1] = bf16_hi(o.x); g[2] = bf16_lo(o.y); g[3] = bf16_hi(o.y);   // what a reader will see
            }
          }
          if (P.stats_out) {                                 // (sum, sumsq) of the row's 64 columns: fixed shuffle tree
            float sm = g[0] + g[1] + g[2] + g[3];
            float ss = g[0] * g[0] + g[1] * g[1] + g[2] * g[2] + g[3] * g[3];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
              sm += __shfl_xor_sync(0xffffffffu, sm, o);
              ss += __shfl_xor_sync(0xffffffffu, ss, o);
            }
            if (c16 == 0 && r < rows_per && row < row_end) P.stats_out[static_cast<size_t>(row) * P.n_tiles + n_t] = make_float2(sm, ss);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&fin_empty[b]);             // the buffer may be refilled
      if (t96 == 0) TR(fq, 9);
      const bool has_waiter = pi + 1 < L.n_phases;
      if (has_waiter) {
        fence_proxy_async_all();                             // the next phase reads these stores through TMA
        __threadfence();
      }
      named_bar_sync(3, 96);
      if (t96 == 0) {
        // the last of the S finalisers of this tile clears its counter for the next launch
        int* counter = L.counters + pi * kCountersPerPhase + n_t;
        if (atomicAdd(counter, 1) == 2 * P.S - 1) *counter = 0;
        if (has_waiter) atomicAdd(L.sync + 2 * pi, 1);       // one more finalised share of phase pi
        TR(fq, 10);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TR(31, 2);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// K slicing for a problem: S slices of KS <= kMaxKS K-blocks (KS even: a weight request is two K-blocks), every slice
// non-empty, S * G CTAs
static int slicing(int K, int N, int sms, int* S, int* KS, int* G) {
  const int kb_total = K / kBK;
  int s = (kb_total + kMaxKS - 1) / kMaxKS;
  int ks = (kb_total + s - 1) / s;
  if (ks & 1) ++ks;
  if (ks > kMaxKS) return OPSG_E_UNSUPPORTED;
  s = (kb_total + ks - 1) / ks;
  if (s > sms) return OPSG_E_UNSUPPORTED;
  const int n_tiles = (N + kBN - 1) / kBN;
  int g = sms / s;
  if (g > n_tiles) g = n_tiles;
  *S = s; *KS = ks; *G = g;
  return OPSG_OK;
}

static long long* g_chain_trace = nullptr;   // set by opsg_debug_chain_trace (trace build only)

struct DeviceState {
  int* counters = nullptr;   // kMaxPhases * kCountersPerPhase tile counters + 2 * kMaxPhases phase counters, all zero at rest
  bool configured = false;
};
static DeviceState g_state[64];

static int device_state(cudaStream_t stream, DeviceState** out) {
  DeviceState& st = g_state[device_slot()];
  if (!st.counters) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
      cudaGetLastError();
      return OPSG_E_UNSUPPORTED;                             // no allocation inside a capture: the first call must be eager
    }
    const size_t bytes = (static_cast<size_t>(kMaxPhases) * kCountersPerPhase + 2 * kMaxPhases) * sizeof(int);
    int rc = check_cuda(cudaMalloc(&st.counters, bytes), "cudaMalloc(gemm chain counters)");
    if (rc) return rc;
    rc = check_cuda(cudaMemset(st.counters, 0, bytes), "cudaMemset(gemm chain counters)");
    if (rc) return rc;
  }
  if (!st.configured) {
    int rc = check_cuda(cudaFuncSetAttribute(skinny_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit),
                        "cudaFuncSetAttribute(gemm chain)");
    if (rc) return rc;
    st.configured = true;
  }
  *out = &st;
  return OPSG_OK;
}

}  // namespace skq

size_t gemm_chain_workspace_bytes(int N, int K) {          // for any M <= 128
  int S, KS, G;
  int sms = opsg_num_sms();
  if (sms <= 0) sms = 148;
  if (K % skq::kBK || skq::slicing(K, N, sms, &S, &KS, &G)) return 0;
  return static_cast<size_t>(S) * 128 * ((N + 3) / 4 * 4) * sizeof(float);
}

// Fills one phase; returns OPSG_E_UNSUPPORTED for layouts the kernel does not take.
static int fill_phase(skq::Phase* P, const opsg_chain_gemm* g, int M, int mr, void* workspace, size_t workspace_bytes, int sms,
                      int wait_prev) {
  using namespace skq;
  const int N = g->N, K = g->K;
  if ((K % kBK) != 0 || (N % 4) != 0 || (g->ldd % 4) != 0 || (g->residual && (g->ldr % 4) != 0)) return OPSG_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(g->D) & 15) != 0 || (g->bias && (reinterpret_cast<uintptr_t>(g->bias) & 15) != 0) ||
      (g->residual && (reinterpret_cast<uintptr_t>(g->residual) & 7) != 0) || N > kCountersPerPhase * kBN)
    return OPSG_E_UNSUPPORTED;
  int rc = slicing(K, N, sms, &P->S, &P->KS, &P->G);
  if (rc) return rc;
  if (workspace_bytes < static_cast<size_t>(P->S) * M * N * sizeof(float))
    return set_error(OPSG_E_INVALID, "gemm_chain: workspace too small (%zu bytes)", workspace_bytes);
  P->kb_total = K / kBK;
  P->n_tiles = (N + kBN - 1) / kBN;
  P->M = M; P->N = N; P->K = K; P->ldd = g->ldd; P->ldr = g->ldr; P->act = g->act; P->out_f32 = g->out_mode == OPSG_OUT_F32;
  P->ws = reinterpret_cast<const float*>(workspace);
  P->bias = g->bias; P->residual = reinterpret_cast<const __nv_bfloat16*>(g->residual); P->D = g->D;
  P->norm = g->norm; P->norm_gamma = g->norm_gamma; P->norm_beta = g->norm_beta; P->norm_eps = g->norm_eps;
  P->stats_in = reinterpret_cast<const float2*>(g->stats_in); P->n_stats_in = g->n_stats_in;
  P->stats_out = reinterpret_cast<float2*>(g->stats_out);
  P->wait_prev = wait_prev;
  rc = make_tmap_bf16_2d(&P->tmA, g->A, (uint64_t)M, (uint64_t)K, (uint64_t)g->lda, mr, kBK);
  if (rc) return rc;
  rc = make_tmap_bf16_kblocks(&P->tmW, g->W, (uint64_t)N, (uint64_t)P->kb_total, (uint64_t)g->ldw, kBN, kKreq);
  if (rc) return rc;
  rc = make_tmap_f32_3d(&P->tmP, workspace, (uint64_t)N, (uint64_t)M, (uint64_t)P->S, 32, mr, 1, true);
  if (rc) return rc;
  return make_tmap_f32_3d(&P->tmF, workspace, (uint64_t)N, (uint64_t)M, (uint64_t)P->S, 32, (M + P->S - 1) / P->S, P->S, false);
}

// Runs `n` dependent GEMMs (phase i + 1 reads what phase i wrote) as one launch.  OPSG_E_UNSUPPORTED: caller falls back.
int launch_gemm_chain(const opsg_chain_gemm* gemms, int n, int M, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  using namespace skq;
  if (n < 1 || n > kMaxPhases || M < 1 || M > 128) return OPSG_E_UNSUPPORTED;
  const char* env = getenv("OPSG_GEMM_CHAIN");               // read per call: tests compare both paths in one process
  if (env && atoi(env) == 0) return OPSG_E_UNSUPPORTED;
  const int sms = opsg_num_sms();
  DeviceState* st = nullptr;
  int rc = device_state(stream, &st);
  if (rc) return rc;
  Launch L;
  memset(&L, 0, sizeof(L));
  const int mr = (M + 7) / 8 * 8;
  for (int i = 0; i < n; ++i) {
    rc = fill_phase(&L.ph[i], &gemms[i], M, mr, workspace, workspace_bytes, sms, i > 0);
    if (rc) return rc;
    if (gemms[i].norm && (!gemms[i].stats_in || !gemms[i].norm_gamma || gemms[i].n_stats_in < 1))
      return set_error(OPSG_E_INVALID, "gemm_chain: phase %d asks for a norm without statistics / gamma", i);
  }
  L.n_phases = n; L.MR = mr;
  L.trace = g_chain_trace;
  L.counters = st->counters;
  L.sync = st->counters + kMaxPhases * kCountersPerPhase;
  int fin_bytes = 0;
  for (int i = 0; i < n; ++i) {
    const int b = 2 * L.ph[i].S * ((M + L.ph[i].S - 1) / L.ph[i].S) * 128;
    fin_bytes = b > fin_bytes ? b : fin_bytes;
  }
  fin_bytes = (fin_bytes + 1023) / 1024 * 1024;
  L.fin_bytes = fin_bytes;
  const int fixed = 1024 + (kAStages + 4) * mr * 128 + 2 * fin_bytes + kNormBytes + kBarrierBytes;
  int stages = (kSmemLimit - fixed) / kSlotBytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 3) return OPSG_E_UNSUPPORTED;
  L.stages = stages;
  int grid = 0;
  for (int i = 0; i < n; ++i) grid = L.ph[i].S * L.ph[i].G > grid ? L.ph[i].S * L.ph[i].G : grid;
  if (n > 1 && grid > sms) return OPSG_E_UNSUPPORTED;         // phases wait for each other: every CTA must be resident
  launch_kernel(skinny_chain_kernel, grid, kThreads, fixed + stages * kSlotBytes, stream, L);
  OPSG_CHECK_LAUNCH("skinny_chain_kernel");
  return OPSG_OK;
}

}  // namespace opsg

extern "C" size_t opsg_gemm_chain_workspace_bytes(const opsg_chain_gemm* gemms, int n) {
  size_t need = 0;
  for (int i = 0; i < n; ++i) {
    const size_t b = opsg::gemm_chain_workspace_bytes(gemms[i].N, gemms[i].K);
    if (b == 0) return 0;
    need = b > need ? b : need;
  }
  return need;
}

extern "C" int opsg_gemm_chain(const opsg_chain_gemm* gemms, int n, int M, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(gemms && workspace && n >= 1, "gemm_chain: null pointer / empty chain");
  for (int i = 0; i < n; ++i) {
    OPSG_CHECK_ARG(gemms[i].A && gemms[i].W && gemms[i].D && gemms[i].N > 0 && gemms[i].K > 0, "gemm_chain: phase %d: bad argument", i);
    OPSG_CHECK_ARG(gemms[i].lda >= gemms[i].K && gemms[i].ldw >= gemms[i].K && gemms[i].ldd >= gemms[i].N &&
                   (gemms[i].lda % 8) == 0 && (gemms[i].ldw % 8) == 0, "gemm_chain: phase %d: bad leading dimension", i);
    OPSG_CHECK_ARG(gemms[i].act >= OPSG_ACT_NONE && gemms[i].act <= OPSG_ACT_RELU, "gemm_chain: phase %d: bad activation", i);
    OPSG_CHECK_ARG(gemms[i].norm >= 0 && gemms[i].norm <= 2, "gemm_chain: phase %d: bad norm", i);
  }
  return opsg::launch_gemm_chain(gemms, n, M, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

#ifdef OPSG_TRACE
extern "C" void opsg_debug_chain_trace(long long* buf) { opsg::skq::g_chain_trace = buf; }
#endif
