// K6s — small-M weight-streaming GEMM (LLM decode: M = selected pairs <= 128; OPT q/k/v, out_proj, fc1, fc2, lm_head of
// one decode step; reference v4:305-312 -> HF OPT decoder layers).
//
// The job is to pull N x K bf16 weights out of HBM once, at HBM speed; the arithmetic is free.  What bounds a CTA's
// share of the stream is the number of WEIGHT bytes it keeps in flight (Little's law: ~2 us x 44 GB/s per SM ~ 90 KB).
// The tiled kernel (gemm.cu) re-streams the [128 x 64] activation tile with every K block: with 128 x 32 tiles 80 % of
// the bytes in flight are activations and every decode GEMM costs ~27 us whatever its size (profiles/r1_llm_decode.md).
// Here:
//   * K is cut into S slices of <= 12 K-blocks; a CTA owns one slice and keeps ITS slice of the activations resident in
//     TENSOR MEMORY as the MMA's A operand (128 lanes = rows, 32 columns of packed bf16 pairs per K-block): loaded once
//     (TMA -> small shared-memory ring -> tcgen05.st by the warp that owns the lane quadrant), never re-read from L2,
//     and no shared-memory operand traffic for A at all;
//   * CTA (slice s, lane g of G = floor(SMs / S)) streams only weights: n-tiles g, g+G, ... of 64 weight rows, one 8 KB
//     K-block per pipeline stage, 23 stages (184 KB of weights in flight per SM);
//   * tcgen05.mma A-from-TMEM ("TS"), M = 128, N = 64: 32 cycles per K = 16 step, fp32 accumulators double-buffered;
//     Measured (scripts/kbench.py streamk, profiles/r1_llm_decode.md): the stream runs at ~3.8 TB/s whatever the number of
//     stages -- one SM's TMA unit keeps only ~32 KB of requests outstanding, ~14 B/clk at DRAM latency; weights pre-tiled so that
//     every stage is one contiguous 8 KB read change nothing and a cp.async loader (4 warps, ~140 KB in flight)
//     was slower (2.2-3 TB/s), so the TMA producer stays;
//   * the epilogue warps write this slice's fp32 partial rows to a workspace [S][M][N] with plain stores, and
//     skinny_finalize_kernel sums the S partials in a fixed order (deterministic, no atomics) and applies
//     bias / activation / residual.  The workspace (<= 16 MB for the decoder GEMMs) lives in L2 between the two kernels.
//     (Tried: the CTA that delivers a tile's last partial reduces it in place of the second kernel -- fence + atomic +
//     fence per tile on the epilogue's critical path made the GEMMs 3x slower; it would need its own warps.)
#include "common.cuh"
#include "host_util.h"

namespace opsg {
namespace sk {

constexpr int kBN = 64;                 // weight rows per n-tile (= MMA N)
constexpr int kBK = 64;                 // 64 bf16 = 128 B = one swizzle span
constexpr int kStageBytes = kBN * 128;  // one K-block of one n-tile
constexpr int kThreads = 224;           // warp 0 W producer, warp 1 MMA, warps 2-5 epilogue (lane quadrants 2,3,0,1), warp 6 A producer
constexpr int kMaxStages = 24;
constexpr int kMaxKS = 12;              // K-blocks per slice: 12 x 32 TMEM columns of A + 2 x 64 accumulator columns = 512
constexpr int kMaxAStages = 12;          // (ring depth is Params::a_stages)
//             // shared-memory ring the activation K-blocks pass through on their way to TMEM
constexpr int kSmemLimit = 232448;
constexpr int kBarrierBytes = 1024;
constexpr int kAccCols = 2 * kBN;

struct Params {
  float* ws;            // [S][M][N] fp32 partials
  int M, N, K;
  int S, KS;            // K slices, K-blocks per slice
  int G;                // CTAs per slice
  int n_tiles;
  int MR;               // M rounded up to 8 rows (rows of an activation K-block in the ring: MR * 128 B)
  int stages;
  int a_stages;         // depth of the shared-memory ring the activation K-blocks pass through on their way to TMEM
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
skinny_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes_kb = p.MR * 128;                       // one activation K-block in the ring (multiple of 1024)
  uint8_t* smem_a = smem;
  uint8_t* smem_w = smem + p.a_stages * a_bytes_kb;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_w + p.stages * kStageBytes);
  uint64_t* full_bar = bars;                               // [stages]      W K-block landed
  uint64_t* empty_bar = full_bar + kMaxStages;             // [stages]      its MMAs have completed
  uint64_t* a_full = empty_bar + kMaxStages;               // [kAStages]    activation K-block landed in the ring
  uint64_t* a_empty = a_full + kMaxAStages;                // [a_stages]    ... and has been copied to TMEM
  uint64_t* a_ready = a_empty + kMaxAStages;               // [kMaxKS]      K-block kb of A is in TMEM (single use)
  uint64_t* tmem_full = a_ready + kMaxKS;                  // [2]
  uint64_t* tmem_empty = tmem_full + 2;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int slice = blockIdx.x % p.S;
  const int g = blockIdx.x / p.S;
  const int kb_total = (p.K + kBK - 1) / kBK;
  const int kb0 = slice * p.KS;
  const int nkb = min(p.KS, kb_total - kb0);               // >= 1 by construction (host)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 4);                           // one arrival per copying warp
    }
    for (int s = 0; s < kMaxKS; ++s) mbar_init(&a_ready[s], 4);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);                        // one arrival per epilogue warp
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
  const uint32_t tmem_a = tmem_base + kAccCols;            // A slice: 32 columns per K-block
  pdl_wait_then_trigger();          // everything above overlaps the previous kernel (programmatic dependent launch)

  if (warp == 0) {
    // ===================== weight producer (TMA) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int n_t = g; n_t < p.n_tiles; n_t += p.G) {
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_2d(smem_w + stage * kStageBytes, &tmW, &full_bar[stage], (kb0 + kb) * kBK, n_t * kBN);
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
        tma_load_2d(smem_a + b * a_bytes_kb, &tmA, &a_full[b], (kb0 + kb) * kBK, 0);   // rows >= M, columns >= K: zeros
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, kBN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    bool first_tile = true;
    for (int n_t = g; n_t < p.n_tiles; n_t += p.G) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kBN;
      for (int kb = 0; kb < nkb; ++kb) {
        if (first_tile) mbar_wait(&a_ready[kb], 0);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t b_desc = umma_desc_k_sw128(smem_u32(smem_w + stage * kStageBytes));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)                 // A: 8 TMEM columns (16 packed bf16) per K = 16 step
            umma_ts(d_tmem, tmem_a + kb * 32 + k * 8, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(&empty_bar[stage]);
          if (kb + 1 == nkb) tc_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      first_tile = false;
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== warps 2-5: A slice -> TMEM, then the epilogue =====================
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

    int acc = 0;
    uint32_t acc_phase = 0;
    float* ws_row = p.ws + (static_cast<size_t>(slice) * p.M + row) * p.N;
    for (int n_t = g; n_t < p.n_tiles; n_t += p.G) {
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
      if (row < p.M) {
        const int col0 = n_t * kBN;
        float* dst = ws_row + col0;
        if (col0 + kBN <= p.N) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            reinterpret_cast<uint4*>(dst)[j] = make_uint4(v0[4 * j], v0[4 * j + 1], v0[4 * j + 2], v0[4 * j + 3]);
            reinterpret_cast<uint4*>(dst + 32)[j] = make_uint4(v1[4 * j], v1[4 * j + 1], v1[4 * j + 2], v1[4 * j + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (col0 + j < p.N) dst[j] = __uint_as_float(v0[j]);
            if (col0 + 32 + j < p.N) dst[32 + j] = __uint_as_float(v1[j]);
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct FinalizeParams {
  const float* ws;
  void* D;
  const float* bias;
  const __nv_bfloat16* residual;
  int M, N, S, ldd, ldr, act, out_f32;
};

// out[m][n] = act(sum_s ws[s][m][n] + bias[n]) + residual[m][n]; 4 columns per thread, slices summed in index order
__global__ void __launch_bounds__(256) skinny_finalize_kernel(const FinalizeParams p) {
  pdl_wait_then_trigger();
  const int n4 = p.N >> 2;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(p.M) * n4) return;
  const int row = static_cast<int>(idx / n4), col = static_cast<int>(idx % n4) * 4;
  const size_t slice_stride = static_cast<size_t>(p.M) * p.N;
  const float* src = p.ws + static_cast<size_t>(row) * p.N + col;
  float4 a = *reinterpret_cast<const float4*>(src);
  int s0 = 1;
  for (; s0 + 8 <= p.S; s0 += 8) {                         // many slices (long K): eight in flight, added in slice order
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4*>(src + (s0 + u) * slice_stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
  }
  for (; s0 < p.S; ++s0) {
    const float4 v = *reinterpret_cast<const float4*>(src + s0 * slice_stride);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  float f[4] = {a.x, a.y, a.z, a.w};
  if (p.bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    f[0] += b.x; f[1] += b.y; f[2] += b.z; f[3] += b.w;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (p.act == OPSG_ACT_GELU) f[e] = gelu_erf(f[e]);
    else if (p.act == OPSG_ACT_RELU) f[e] = fmaxf(f[e], 0.f);
  }
  if (p.residual) {
    const uint2 r = *reinterpret_cast<const uint2*>(p.residual + static_cast<size_t>(row) * p.ldr + col);
    f[0] += bf16_lo(r.x); f[1] += bf16_hi(r.x); f[2] += bf16_lo(r.y); f[3] += bf16_hi(r.y);
  }
  if (p.out_f32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.D) + static_cast<size_t>(row) * p.ldd + col) =
        make_float4(f[0], f[1], f[2], f[3]);
  } else {
    uint2 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.D) + static_cast<size_t>(row) * p.ldd + col) = o;
  }
}

// K slicing for a problem: S slices of KS <= kMaxKS K-blocks, every slice non-empty.  When K is so long that the
// slices alone outnumber half the SMs (PatchEmbed: K = 65536 -> 1024 K-blocks) every SM gets one slice.
static void slicing(int K, int sms, int* S, int* KS) {
  const int kb_total = (K + kBK - 1) / kBK;
  int ks = kMaxKS;
  if ((kb_total + kMaxKS - 1) / kMaxKS > sms / 2) ks = (kb_total + sms - 1) / sms;
  if (ks > kMaxKS) ks = kMaxKS;                            // more slices than SMs: several waves of CTAs would be needed
  int s = (kb_total + ks - 1) / ks;
  ks = (kb_total + s - 1) / s;
  s = (kb_total + ks - 1) / ks;
  *S = s;
  *KS = ks;
}

}  // namespace sk

size_t gemm_skinny_workspace_bytes(int N, int K) {      // for any M <= 128
  int S, KS;
  int sms = opsg_num_sms();
  if (sms <= 0) sms = 148;
  sk::slicing(K, sms, &S, &KS);
  return static_cast<size_t>(S) * 128 * ((N + 3) / 4 * 4) * sizeof(float);
}

// D = act(A . W^T + bias) + residual for M <= 128.  Returns OPSG_E_UNSUPPORTED when the layout needs the stream-K path.
int launch_gemm_skinny(const opsg_bf16* A, int lda, const opsg_bf16* W, int ldw, void* D, int ldd, int M, int N, int K,
                       const float* bias, const opsg_bf16* residual, int ldr, int act, int out_mode, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
  using namespace sk;
  if ((N % 4) != 0 || (ldd % 4) != 0 || (residual && (ldr % 4) != 0)) return OPSG_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(D) & 15) != 0 || (bias && (reinterpret_cast<uintptr_t>(bias) & 15) != 0) ||
      (residual && (reinterpret_cast<uintptr_t>(residual) & 7) != 0))
    return OPSG_E_UNSUPPORTED;
  Params p;
  const int sms = opsg_num_sms();
  slicing(K, sms, &p.S, &p.KS);
  p.MR = (M + 7) / 8 * 8;
  if (p.S > sms) return OPSG_E_UNSUPPORTED;
  p.M = M; p.N = N; p.K = K;
  p.n_tiles = (N + kBN - 1) / kBN;
  p.G = sms / p.S;
  if (p.G > p.n_tiles) p.G = p.n_tiles;
  p.ws = reinterpret_cast<float*>(workspace);
  if (workspace_bytes < static_cast<size_t>(p.S) * M * N * sizeof(float))
    return set_error(OPSG_E_INVALID, "gemm_skinny: workspace too small (%zu bytes)", workspace_bytes);
  p.a_stages = 3;        // deeper rings (6, 10) and more producer warps measured no faster: profiles/r2_decode_timeline.md
  if (p.a_stages > p.KS) p.a_stages = p.KS;
  const int a_bytes = p.a_stages * p.MR * 128;
  int stages = (kSmemLimit - 1024 - a_bytes - kBarrierBytes) / kStageBytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 4) return OPSG_E_UNSUPPORTED;
  p.stages = stages;
  const int smem_bytes = 1024 + a_bytes + stages * kStageBytes + kBarrierBytes;
  CUtensorMap tmA, tmW;
  int rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, p.MR, kBK);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmW, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, kBN, kBK);
  if (rc) return rc;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[device_slot()];
  if (!configured) {
    rc = check_cuda(cudaFuncSetAttribute(skinny_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit),
                    "cudaFuncSetAttribute(gemm skinny)");
    if (rc) return rc;
    configured = true;
  }
  launch_kernel(skinny_gemm_kernel, p.S * p.G, kThreads, smem_bytes, stream, tmA, tmW, p);
  OPSG_CHECK_LAUNCH("skinny_gemm_kernel");
  FinalizeParams f;
  f.ws = p.ws; f.D = D; f.bias = bias; f.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  f.M = M; f.N = N; f.S = p.S; f.ldd = ldd; f.ldr = ldr; f.act = act; f.out_f32 = out_mode == OPSG_OUT_F32;
  const long long total = static_cast<long long>(M) * (N / 4);
  launch_kernel(skinny_finalize_kernel, (total + 255) / 256, 256, 0, stream, f);
  OPSG_CHECK_LAUNCH("skinny_finalize_kernel");
  return OPSG_OK;
}

}  // namespace opsg
