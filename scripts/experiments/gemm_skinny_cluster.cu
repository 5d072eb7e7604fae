// K6c — small-M weight-streaming GEMM with the K-slice reduction INSIDE a thread-block cluster (LLM decode: M = selected
// pairs <= 128; q/k/v, out_proj, fc1 / gate-up, fc2 / down, lm_head of one decode step; reference v4:305-312 -> HF OPT /
// Llama decoder layers).  Same entry point as gemm_skinny.cu (opsg_gemm_bf16_streamk), tried first.
//
// gemm_skinny.cu cuts K into slices, writes every slice's fp32 partial rows to a workspace in L2 and sums them in a second
// kernel.  profiles/r2_decode_timeline.md: that costs a kernel boundary (~2 us) + 2-3 us of reduction per GEMM -- 18 us of a
// 117 us decoder layer -- and the scattered 16-byte partial-row stores slow the weight stream itself by a third.  Here:
//   * a CLUSTER of CS CTAs (4 / 8 / 16) owns a set of n-tiles; CTA rank r holds K slice r of the activations resident in
//     tensor memory (A operand of the TS-mode MMAs, as before) and streams only its K slice of the n-tile's weights;
//   * when a tile's accumulator is complete every CTA SENDS the columns it does not own to their owners through
//     distributed shared memory (st.shared::cluster + a releasing remote mbarrier arrive) and the owner of columns
//     [r W, (r + 1) W), W = BN / CS, adds the CS partials in rank order (deterministic), applies bias / activation /
//     residual and writes the bf16 (or fp32) output: no workspace, no second kernel, no fp32 traffic through L2;
//   * a weight request is 16-20 KB (a 3-D TMA box of several consecutive K-blocks of the n-tile): one issuing warp
//     sustains 6.4 TB/s with requests of this size against 4.4 TB/s with 8 KB ones (profiles/r2_decode_timeline.md);
//   * n-tiles are 64 or 32 weight rows wide, whichever balances the tile count over the resident clusters better.
// Falls back (OPSG_E_UNSUPPORTED -> gemm_skinny.cu) when K is not a multiple of 64 or no cluster size gives every rank a
// non-empty slice of <= 12 K-blocks.
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"

namespace opsg {
namespace skc {

constexpr int kBK = 64;                 // 64 bf16 = 128 B = one swizzle span
constexpr int kThreads = 352;           // warp 0 W producer, 1 MMA, 2-5 A copy + senders, 6 A producer, 7-10 receivers
constexpr int kMaxStages = 16;
constexpr int kMaxKS = 12;              // K-blocks per slice: 12 x 32 TMEM columns of A + 2 x 64 accumulator columns = 512
constexpr int kMaxAStages = 4;
constexpr int kSmemLimit = 232448;
constexpr int kBarrierBytes = 1024;
constexpr int kAccCols = 128;           // two accumulator buffers of <= 64 columns

struct Params {
  void* D;
  const float* bias;
  const __nv_bfloat16* residual;
  int M, N, K;
  int ldd, ldr, act, out_f32;
  int KS;               // K-blocks per slice (rank r owns K-blocks [r KS, (r + 1) KS))
  int kb_total;
  int n_tiles, n_clusters;
  int MR;               // M rounded up to 8 rows
  int stages, kreq;     // ring slots, K-blocks per weight request
  int a_stages;
  int dbg;              // development: bit 0 skips the cluster exchange, bit 1 the global loads / stores of the epilogue
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
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_addr), "r"(cta));
  return ra;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t raddr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// remote arrive that publishes (orders) this thread's earlier memory operations at cluster scope
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
    if (clock64() - t0 > OPSG_WAIT_LIMIT_CYCLES) {
      printf("opsg: cluster mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
// bulk copy of `bytes` (multiple of 16) from this CTA's shared memory into another CTA of the cluster; completes (complete_tx)
// on an mbarrier of the destination CTA.  dst / bar are shared::cluster addresses (map_to_cta).
__device__ __forceinline__ void bulk_copy_to_cta(uint32_t dst, uint32_t src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "r"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// 3-D tile load, coordinates {c0 = element inside the K-block, c1 = weight row, c2 = K-block}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1,
                                            int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
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

// TMEM -> registers: this thread's lane, BN consecutive fp32 columns
template <int BN>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[BN]) {
  uint32_t t[32];
#pragma unroll
  for (int c = 0; c < BN; c += 32) {
    tmem_ld32(taddr + c, t);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[c + j] = t[j];
  }
}

// Receive buffers: recv[buf][source rank][row][W floats], 16-byte chunks of a row XOR-swizzled so that the 8 lanes of a
// quarter-warp hit 8 different 16-byte bank groups (rows are W * 4 = 16 / 32 / 64 bytes apart).
template <int W>
__device__ __forceinline__ int recv_off(int row, int chunk) {
  constexpr int CH = W / 4;                       // 16-byte chunks per row
  const int sw = (row / (8 / CH)) & (CH - 1);
  return row * (W * 4) + ((chunk ^ sw) * 16);
}

template <int BN, int CS>
__global__ void __launch_bounds__(kThreads, 1)
skinny_cluster_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const Params p) {
  constexpr int W = BN / CS;                      // output columns of a tile finalised by one CTA
  static_assert(W >= 4 && W <= 16 && (W & (W - 1)) == 0, "BN / CS must be 4, 8 or 16");
  constexpr int kBlockBytes = BN * 128;           // one K-block of one n-tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes_kb = p.MR * 128;
  const int slot_bytes = kBlockBytes * p.kreq;
  const int xch_src_bytes = p.MR * W * 4;         // one (source, destination) pair's rows of a tile: MR x W floats
  const int send_buf_bytes = (CS - 1) * xch_src_bytes;
  const int recv_buf_bytes = CS * xch_src_bytes;
  uint8_t* smem_a = smem;
  uint8_t* smem_w = smem + p.a_stages * a_bytes_kb;
  uint8_t* smem_send = smem_w + p.stages * slot_bytes;       // [2][CS - 1 destinations][MR][W]  partial columns to send
  uint8_t* smem_recv = smem_send + 2 * send_buf_bytes;       // [2][CS sources][MR][W]           partial columns received
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_recv + 2 * recv_buf_bytes);
  uint64_t* full_bar = bars;                               // [stages]   weight request landed
  uint64_t* empty_bar = full_bar + kMaxStages;             // [stages]   its MMAs have completed
  uint64_t* a_full = empty_bar + kMaxStages;               // [a_stages] activation K-block landed in the ring
  uint64_t* a_empty = a_full + kMaxAStages;                // [a_stages] ... and has been copied to TMEM
  uint64_t* a_ready = a_empty + kMaxAStages;               // [kMaxKS]   K-block kb of A is in TMEM (single use)
  uint64_t* tmem_full = a_ready + kMaxKS;                  // [2]
  uint64_t* tmem_empty = tmem_full + 2;                    // [2]
  uint64_t* recv_full = tmem_empty + 2;                    // [2] all peers' partial columns of a tile have arrived
  uint64_t* send_ok = recv_full + 2;                       // [2] all peers have consumed what was sent into their buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(send_ok + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());   // = K slice
  const int cluster_id = blockIdx.x / CS;
  const int kb0 = rank * p.KS;
  const int nkb = min(p.KS, p.kb_total - kb0);             // >= 1 by construction (host)
  const int nreq = (nkb + p.kreq - 1) / p.kreq;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 4);
    }
    for (int s = 0; s < kMaxKS; ++s) mbar_init(&a_ready[s], 4);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
      mbar_init(&recv_full[s], 1);                         // the local arrive.expect_tx (own slot written) + the peers' bytes
      mbar_init(&send_ok[s], CS * 4);                      // every receiver warp of every CTA (this one included) per tile
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
  cluster_sync_all();                // every CTA's barriers exist before any remote arrive
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a = tmem_base + kAccCols;
  pdl_wait_then_trigger();

  if (warp == 0) {
    // ===================== weight producer (TMA, one request = kreq K-blocks of the n-tile) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int n_t = cluster_id; n_t < p.n_tiles; n_t += p.n_clusters) {
      for (int q = 0; q < nreq; ++q) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(slot_bytes));
          tma_load_3d(smem_w + stage * slot_bytes, &tmW, &full_bar[stage], 0, n_t * BN, kb0 + q * p.kreq);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 6) {
    // ===================== activation producer: the slice's K-blocks through the ring, once =====================
    for (int kb = 0; kb < nkb; ++kb) {
      const int b = kb % p.a_stages;
      mbar_wait(&a_empty[b], ((kb / p.a_stages) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_expect_tx(&a_full[b], static_cast<uint32_t>(a_bytes_kb));
        tma_load_2d(smem_a + b * a_bytes_kb, &tmA, &a_full[b], (kb0 + kb) * kBK, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    bool first_tile = true;
    for (int n_t = cluster_id; n_t < p.n_tiles; n_t += p.n_clusters) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int q = 0; q < nreq; ++q) {
        const int kbq = q * p.kreq;
        const int nb = min(p.kreq, nkb - kbq);
        if (first_tile)
          for (int j = 0; j < nb; ++j) mbar_wait(&a_ready[kbq + j], 0);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t w_addr = smem_u32(smem_w + stage * slot_bytes);
          for (int j = 0; j < nb; ++j) {
            const uint64_t b_desc = umma_desc_k_sw128(w_addr + j * kBlockBytes);
            const int kb = kbq + j;
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_ts(d_tmem, tmem_a + kb * 32 + k * 8, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (q + 1 == nreq) tc_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      first_tile = false;
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp < 6) {
    // ===================== warps 2-5: A slice -> TMEM, then the senders of the cluster reduction =====================
    const int q = warp & 3;                                  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    for (int kb = 0; kb < nkb; ++kb) {
      const int b = kb % p.a_stages;
      mbar_wait(&a_full[b], (kb / p.a_stages) & 1);
      uint32_t v[32];
      if (row < p.MR) {
        const uint8_t* src = smem_a + b * a_bytes_kb + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {                        // undo the 128-byte swizzle: 16-byte chunk j sits at j ^ (row % 8)
          const uint4 u = *reinterpret_cast<const uint4*>(src + ((j ^ (row & 7)) * 16));
          v[4 * j] = u.x; v[4 * j + 1] = u.y; v[4 * j + 2] = u.z; v[4 * j + 3] = u.w;
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

    // ---- senders: accumulator -> staging slots -> bulk copies into the owners' receive buffers ----
    const uint32_t recv_full_addr = smem_u32(recv_full);
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;
    for (int n_t = cluster_id; n_t < p.n_tiles; n_t += p.n_clusters, ++it) {
      const int buf = it & 1;
      const uint32_t use_parity = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      uint32_t v[BN];
      tmem_ld_cols<BN>(tmem_base + lane_base + acc * BN, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }

      // The columns owned by peer d go to the local staging slot of d, this CTA's own columns straight into slot `rank` of
      // its receive buffer, with plain shared-memory stores; one thread then pushes every staging slot into d's receive
      // buffer with a bulk shared::cta -> shared::cluster copy that completes on d's recv_full barrier.  (Per-lane
      // st.shared::cluster stores of the same 24 KB took ~5500 cycles per tile.)
      uint8_t* my_send = smem_send + buf * send_buf_bytes;
      uint8_t* my_recv = smem_recv + buf * recv_buf_bytes;
      // every owner (this CTA's receiver warps included) has consumed what was sent two tiles ago: the peers' buffers
      // `buf`, this staging buffer and the own slot are free
      if (it >= 2) mbar_wait_acquire_cluster(&send_ok[buf], use_parity ^ 1);
      if (row < p.MR) {
#pragma unroll
        for (int d = 0; d < CS; ++d) {
          uint8_t* slot = d == rank ? my_recv + rank * xch_src_bytes : my_send + (d - (d > rank ? 1 : 0)) * xch_src_bytes;
#pragma unroll
          for (int c = 0; c < W / 4; ++c)
            *reinterpret_cast<uint4*>(slot + recv_off<W>(row, c)) =
                make_uint4(v[d * W + 4 * c], v[d * W + 4 * c + 1], v[d * W + 4 * c + 2], v[d * W + 4 * c + 3]);
        }
      }
      fence_proxy_async_smem();                              // staging writes -> visible to the bulk-copy engine
      named_bar_sync(1, 128);                                // all four sender warps have written their rows
      if (warp == 2 && elect_one_sync()) {
        if (!(p.dbg & 1)) {
          mbar_expect_tx(&recv_full[buf], static_cast<uint32_t>(send_buf_bytes));
#pragma unroll
          for (int d = 0; d < CS; ++d) {
            if (d == rank) continue;
            const uint32_t src = smem_u32(my_send + (d - (d > rank ? 1 : 0)) * xch_src_bytes);
            const uint32_t dst = map_to_cta(smem_u32(my_recv + rank * xch_src_bytes), d);
            bulk_copy_to_cta(dst, src, static_cast<uint32_t>(xch_src_bytes), map_to_cta(recv_full_addr + buf * 8, d));
          }
        } else {
          mbar_arrive(&recv_full[buf]);
        }
      }
      __syncwarp();
    }
  }
  if (warp >= 7) {
    // ===================== warps 7-10: receive, reduce in rank order, epilogue =====================
    const int row = (warp - 7) * 32 + lane;
    const bool row_live = row < p.M;
    const uint32_t send_ok_addr = smem_u32(send_ok);
    int it = 0;
    for (int n_t = cluster_id; n_t < p.n_tiles; n_t += p.n_clusters, ++it) {
      const int buf = it & 1;
      const uint32_t use_parity = (it >> 1) & 1;
      // bias / residual of the columns this CTA finalises: requested before the wait (an L2 round trip)
      const int col0 = n_t * BN + rank * W;
      float4 bias_v[W / 4];
      uint2 res_v[W / 4];
#pragma unroll
      for (int c = 0; c < W / 4; ++c) {
        const int col = col0 + 4 * c;
        bias_v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        res_v[c] = make_uint2(0u, 0u);
        if (row_live && col < p.N && !(p.dbg & 2)) {
          if (p.bias) bias_v[c] = __ldg(reinterpret_cast<const float4*>(p.bias + col));
          if (p.residual) res_v[c] = *reinterpret_cast<const uint2*>(p.residual + static_cast<size_t>(row) * p.ldr + col);
        }
      }
      const uint8_t* my_recv = smem_recv + buf * recv_buf_bytes;
      mbar_wait_acquire_cluster(&recv_full[buf], use_parity);
      float f[W];
#pragma unroll
      for (int j = 0; j < W; ++j) f[j] = 0.f;
      if (row_live) {
#pragma unroll
        for (int s = 0; s < CS; ++s) {                       // rank order: the sum does not depend on arrival order
          if ((p.dbg & 1) && s != rank) continue;
#pragma unroll
          for (int c = 0; c < W / 4; ++c) {
            const float4 u = *reinterpret_cast<const float4*>(my_recv + s * xch_src_bytes + recv_off<W>(row, c));
            f[4 * c] += u.x; f[4 * c + 1] += u.y; f[4 * c + 2] += u.z; f[4 * c + 3] += u.w;
          }
        }
      }
      // buffer `buf` of this CTA may be overwritten (tile it + 2) once every lane of this warp has read it: tell every
      // sender of the cluster, this CTA's own included
      __syncwarp();
      if (lane == 0) {
        fence_acq_rel_cluster();
#pragma unroll
        for (int pr = 0; pr < CS; ++pr) mbar_arrive_remote_release(map_to_cta(send_ok_addr + buf * 8, pr));
      }

      if (row_live && !(p.dbg & 2)) {
#pragma unroll
        for (int c = 0; c < W / 4; ++c) {
          const int col = col0 + 4 * c;
          if (col >= p.N) break;                             // N % 4 == 0 (host)
          float g[4] = {f[4 * c] + bias_v[c].x, f[4 * c + 1] + bias_v[c].y, f[4 * c + 2] + bias_v[c].z,
                        f[4 * c + 3] + bias_v[c].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (p.act == OPSG_ACT_GELU) g[e] = gelu_erf(g[e]);
            else if (p.act == OPSG_ACT_RELU) g[e] = fmaxf(g[e], 0.f);
          }
          g[0] += bf16_lo(res_v[c].x); g[1] += bf16_hi(res_v[c].x); g[2] += bf16_lo(res_v[c].y); g[3] += bf16_hi(res_v[c].y);
          if (p.out_f32) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.D) + static_cast<size_t>(row) * p.ldd + col) =
                make_float4(g[0], g[1], g[2], g[3]);
          } else {
            uint2 o;
            o.x = pack_bf16x2(g[0], g[1]);
            o.y = pack_bf16x2(g[2], g[3]);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.D) + static_cast<size_t>(row) * p.ldd + col) = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                // peers may still be writing into this CTA's buffers / arriving on its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct Plan {
  int cs, bn, ks, n_tiles, n_clusters, kreq, stages, a_stages, smem_bytes;
};

template <int BN, int CS>
static int max_clusters(int* out) {
  static int cached[64] = {};
  static bool configured[64] = {};
  const int dev = device_slot();
  auto kernel = skinny_cluster_kernel<BN, CS>;
  if (!configured[dev]) {
    int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit),
                        "cudaFuncSetAttribute(gemm skinny cluster, smem)");
    if (rc) return rc;
    if (CS > 8) {
      rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1),
                      "cudaFuncSetAttribute(gemm skinny cluster, cluster size)");
      if (rc) return rc;
    }
    configured[dev] = true;
  }
  if (cached[dev] == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * 64);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemLimit;           // the largest layout: a safe lower bound for every problem
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = -1;
    }
    cached[dev] = n;
  }
  *out = cached[dev];
  return OPSG_OK;
}

template <int BN, int CS>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmW, const Params& p, int smem_bytes, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.n_clusters * CS);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  (void)cudaLaunchKernelEx(&cfg, skinny_cluster_kernel<BN, CS>, tmA, tmW, p);
  OPSG_CHECK_LAUNCH("skinny_cluster_kernel");
  return OPSG_OK;
}

static int layout(int bn, int cs, int ks, int mr, Plan* pl) {
  // weight request = kreq K-blocks, <= 20 KB, kreq a divisor of ks when one gives >= 12 KB (no bytes fetched past the slice)
  const int block = bn * 128;
  int kreq = 20480 / block;
  if (kreq > ks) kreq = ks;
  for (int d = kreq; d >= 1; --d)
    if (ks % d == 0 && d * block >= 12288) { kreq = d; break; }
  if (getenv("OPSG_SKINNY_KREQ")) kreq = atoi(getenv("OPSG_SKINNY_KREQ"));
  pl->kreq = kreq;
  pl->a_stages = ks < 2 ? ks : 2;
  const int fixed = 1024 + pl->a_stages * mr * 128 + 2 * (2 * cs - 1) * mr * (bn / cs) * 4 + kBarrierBytes;
  int stages = (kSmemLimit - fixed) / (kreq * block);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 3) return OPSG_E_UNSUPPORTED;
  pl->stages = stages;
  pl->smem_bytes = fixed + stages * kreq * block;
  return OPSG_OK;
}

}  // namespace skc

// D = act(A . W^T + bias) + residual for M <= 128, K-slice reduction inside thread-block clusters.
// OPSG_E_UNSUPPORTED: the caller falls back to the workspace version (gemm_skinny.cu).
int launch_gemm_skinny_cluster(const opsg_bf16* A, int lda, const opsg_bf16* W, int ldw, void* D, int ldd, int M, int N, int K,
                               const float* bias, const opsg_bf16* residual, int ldr, int act, int out_mode,
                               cudaStream_t stream) {
  using namespace skc;
  const char* env = getenv("OPSG_SKINNY_CLUSTER");            // read per call: tests compare both paths in one process
  if (env && atoi(env) == 0) return OPSG_E_UNSUPPORTED;
  if ((K % kBK) != 0 || (N % 4) != 0 || (ldd % 4) != 0 || (residual && (ldr % 4) != 0)) return OPSG_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(D) & 15) != 0 || (bias && (reinterpret_cast<uintptr_t>(bias) & 15) != 0) ||
      (residual && (reinterpret_cast<uintptr_t>(residual) & 7) != 0))
    return OPSG_E_UNSUPPORTED;
  const int kb_total = K / kBK;
  Plan pl = {};
  for (int cs : {4, 8, 16}) {
    const int ks = (kb_total + cs - 1) / cs;
    if (ks <= kMaxKS && (cs - 1) * ks < kb_total) { pl.cs = cs; pl.ks = ks; break; }
  }
  if (pl.cs == 0) return OPSG_E_UNSUPPORTED;
  const int mr = (M + 7) / 8 * 8;

  // tile width: 64 weight rows, or 32 when that shortens the longest CTA's share of the stream (waves x rows) by >= 7 %
  int best_bn = 0, best_nc = 0;
  long best_cost = 0;
  const char* bn_env = getenv("OPSG_SKINNY_BN");           // development switch
  const int bn_forced = bn_env ? atoi(bn_env) : 0;
  for (int bn : {64, 32}) {
    if (bn / pl.cs < 4) continue;
    if (bn_forced && bn != bn_forced && bn_forced / pl.cs >= 4) continue;
    int nc = 0, rc = OPSG_OK;
    if (bn == 64 && pl.cs == 4) rc = max_clusters<64, 4>(&nc);
    else if (bn == 64 && pl.cs == 8) rc = max_clusters<64, 8>(&nc);
    else if (bn == 64 && pl.cs == 16) rc = max_clusters<64, 16>(&nc);
    else if (bn == 32 && pl.cs == 4) rc = max_clusters<32, 4>(&nc);
    else if (bn == 32 && pl.cs == 8) rc = max_clusters<32, 8>(&nc);
    if (rc) return rc;
    if (nc <= 0) continue;
    const int tiles = (N + bn - 1) / bn;
    if (nc > tiles) nc = tiles;
    const int waves = (tiles + nc - 1) / nc;
    nc = (tiles + waves - 1) / waves;                       // same makespan with fewer clusters: less activation traffic
    const long cost = static_cast<long>(waves) * bn;
    if (best_bn == 0 || cost * 100 < best_cost * 93) { best_bn = bn; best_nc = nc; best_cost = cost; }
  }
  if (best_bn == 0) return OPSG_E_UNSUPPORTED;
  pl.bn = best_bn;
  pl.n_clusters = best_nc;
  if (const char* e = getenv("OPSG_SKINNY_NC")) {             // development switch
    const int tiles = (N + pl.bn - 1) / pl.bn;
    pl.n_clusters = atoi(e) / pl.cs < tiles ? atoi(e) / pl.cs : tiles;
  }
  pl.n_tiles = (N + pl.bn - 1) / pl.bn;
  int rc = layout(pl.bn, pl.cs, pl.ks, mr, &pl);
  if (rc) return rc;

  Params p;
  p.D = D; p.bias = bias; p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.M = M; p.N = N; p.K = K; p.ldd = ldd; p.ldr = ldr; p.act = act; p.out_f32 = out_mode == OPSG_OUT_F32;
  p.KS = pl.ks; p.kb_total = kb_total; p.n_tiles = pl.n_tiles; p.n_clusters = pl.n_clusters; p.MR = mr;
  p.stages = pl.stages; p.kreq = pl.kreq; p.a_stages = pl.a_stages;
  p.dbg = getenv("OPSG_SKINNY_DBG") ? atoi(getenv("OPSG_SKINNY_DBG")) : 0;
  if (getenv("OPSG_SKINNY_DEBUG"))
    fprintf(stderr, "skinny_cluster M=%d N=%d K=%d cs=%d bn=%d ks=%d tiles=%d clusters=%d kreq=%d stages=%d smem=%d\n", M, N, K,
            pl.cs, pl.bn, pl.ks, pl.n_tiles, pl.n_clusters, pl.kreq, pl.stages, pl.smem_bytes);
  CUtensorMap tmA, tmW;
  rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, mr, kBK);
  if (rc) return rc;
  rc = make_tmap_bf16_kblocks(&tmW, W, (uint64_t)N, (uint64_t)kb_total, (uint64_t)ldw, pl.bn, pl.kreq);
  if (rc) return rc;
  if (pl.bn == 64 && pl.cs == 4) return launch<64, 4>(tmA, tmW, p, pl.smem_bytes, stream);
  if (pl.bn == 64 && pl.cs == 8) return launch<64, 8>(tmA, tmW, p, pl.smem_bytes, stream);
  if (pl.bn == 64 && pl.cs == 16) return launch<64, 16>(tmA, tmW, p, pl.smem_bytes, stream);
  if (pl.bn == 32 && pl.cs == 4) return launch<32, 4>(tmA, tmW, p, pl.smem_bytes, stream);
  if (pl.bn == 32 && pl.cs == 8) return launch<32, 8>(tmA, tmW, p, pl.smem_bytes, stream);
  return OPSG_E_UNSUPPORTED;
}

}  // namespace opsg
