// K4 (tcgen05) — Q-Former self-attention over the S = n_query + T <= 64 rows of every pair (row a5 of SURVEY.md §8;
// InstructBlipQFormerMultiHeadAttention.forward, HF modeling_instructblip.py:504-536, as called B = N^2 times by
// relation_transformer_head_v4.py:179-185).
//
// One (pair, head) problem is 49 x 49 x 64: far too small for a 128-row tensor-core tile, so TWO pairs are stacked into one
// tile with a block-diagonal key mask (the trick K5 uses for its pair masks): pair slot s owns tile rows / keys
// [64 s, 64 s + n_query + T).  Work unit = (2 pairs, head); a persistent CTA walks a contiguous range of units.
//
//   warp 0      TMA producer: per unit 12 box loads (q | k | v) x (slot 0, 1) x (query rows, text rows) of the head's 64
//               columns straight out of the fused qkv activation (split row layout of opsg_qformer_embed_ln; layer 0 takes
//               the pair-independent query rows from the shared table) into 128-byte-swizzled tiles, 3 stages
//   warp 1      S = Q K^T    : 4 tcgen05.mma (SS, M = 128, N = 128, K = 16) into one of FOUR 128-column TMEM buffers
//   warp 2      TMEM allocator, then O = P V : 8 tcgen05.mma (A = P from TMEM, B = V row-major = MN-major shared-memory
//               operand, N = 64) into columns [64, 128) of the buffer
//   warps 4-19  four softmax / epilogue warpgroups, one per TMEM buffer, thread = one tile row: loads ITS pair's 64 score
//               columns (one TMEM round trip), key mask (query keys always, text keys by the attention mask: HF's -10000
//               bias underflows to weight 0), max, exp2, bf16 P in place (+ zeros over the other pair's half), row sum in a
//               register; after the PV product: O / sum -> bf16 -> swizzled staging -> TMA stores (query rows; text rows only
//               when a later layer reads them).  (Writing each row's 128 bytes straight to global memory instead -- 16 bytes
//               per lane per store, 32 different lines per instruction -- was measured 30 % slower.)
// Nothing is re-read: per unit 37 KB in, 12 KB out; up to four units are in flight per SM.
#include <math.h>

#include "common.cuh"
#include "host_util.h"

#ifndef SA_BUFS
#define SA_BUFS 3
#endif

namespace opsg {

constexpr int kSaBufs = SA_BUFS;           // TMEM buffers (128 columns each) = softmax warpgroups
constexpr int kSaThreads = (4 + 4 * kSaBufs) * 32;
constexpr int kSaStages = 3;               // shared-memory operand stages
constexpr int kSaTile = 128 * 128;         // one operand tile: 128 rows x 128 B

struct SaParams {
  const int32_t* text_mask;  // [B, T]
  __nv_bfloat16* out;        // [R_out, d_model]
  int B, n_query, T, num_heads, d_model;
  int text_queries;          // 0: only the query rows are computed / stored (last layer)
  int shared_query;          // query rows of q / k / v come from the shared table (layer 0)
  int total_units;           // ceil(B / 2) * num_heads
  float scale_log2e;
};

struct SaSmem {
  static constexpr int kStage = 3 * kSaTile;                       // Q | K | V
  static constexpr int kOffOp = 0;
  static constexpr int kOffO = kOffOp + kSaStages * kStage;        // staging tiles for the TMA stores, one per buffer
  static constexpr int kOffBar = kOffO + kSaBufs * kSaTile;
  static constexpr int kTotal = kOffBar + 256 + 1024;
};
static_assert(SaSmem::kTotal <= 232448, "shared memory budget exceeded");

__device__ __forceinline__ void sa_tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
      "%25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
      "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
      "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void sa_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float sa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NQ>
__global__ void __launch_bounds__(kSaThreads, 1)
self_attn_pairs_kernel(const __grid_constant__ CUtensorMap tmQry,   // qkv, box [n_query x 64]
                       const __grid_constant__ CUtensorMap tmTxt,   // qkv, box [T x 64]
                       const __grid_constant__ CUtensorMap tmShr,   // shared query table, box [n_query x 64]
                       const __grid_constant__ CUtensorMap tmOq,    // out, box [n_query x 64]
                       const __grid_constant__ CUtensorMap tmOt,    // out, box [T x 64]
                       const SaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sOp = smem + SaSmem::kOffOp;
  uint8_t* sO = smem + SaSmem::kOffO;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SaSmem::kOffBar);
  uint64_t* full = bars;                       // [<= 4] operands of a unit landed
  uint64_t* empty = bars + 4;                  // [<= 4] both MMAs that read the stage have completed (count 2)
  uint64_t* s_full = bars + 8;                 // [4]
  uint64_t* p_ready = bars + 12;               // [4] (count 128)
  uint64_t* o_full = bars + 16;                // [4]
  uint64_t* s_free = bars + 20;                // [4] (count 128)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int nq = NQ;
  const int T = p.T;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQry);
    tma_prefetch_desc(&tmTxt);
    tma_prefetch_desc(&tmShr);
    tma_prefetch_desc(&tmOq);
    tma_prefetch_desc(&tmOt);
    for (int s = 0; s < kSaStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); }
    for (int b = 0; b < kSaBufs; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&p_ready[b], 128);
      mbar_init(&o_full[b], 1);
      mbar_init(&s_free[b], 128);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // rows no TMA box ever writes (the padding of each 64-row pair slot) must hold finite values: a zero probability times
  // a NaN bit pattern left in V would poison the product
  for (int idx = threadIdx.x; idx < kSaStages * SaSmem::kStage / 16; idx += kSaThreads)
    reinterpret_cast<uint4*>(sOp)[idx] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait_then_trigger();

  const int per = (p.total_units + gridDim.x - 1) / gridDim.x;
  const int u_begin = blockIdx.x * per;
  const int u_end = min(p.total_units, u_begin + per);
  const int n_units = max(0, u_end - u_begin);
  const int RQ = p.B * nq;                               // first text row of the split layout

  if (warp == 0) {
    // ===================== TMA producer =====================
    for (int i = 0; i < n_units; ++i) {
      const int u = u_begin + i;
      const int tile = u / p.num_heads, head = u % p.num_heads;
      const int st = i % kSaStages;
      mbar_wait(&empty[st], ((i / kSaStages) & 1) ^ 1);
      if (elect_one_sync()) {
        const int n_pairs = min(2, p.B - 2 * tile);
        const int q_txt = (p.text_queries && T > 0) ? 1 : 0;          // the text rows of Q are only needed as queries
        const uint32_t bytes = static_cast<uint32_t>(n_pairs) * (3u * nq + (2u + q_txt) * T) * 128u;
        mbar_expect_tx(&full[st], bytes);
        uint8_t* base = sOp + st * SaSmem::kStage;
        for (int s = 0; s < n_pairs; ++s) {
          const int pair = 2 * tile + s;
#pragma unroll
          for (int part = 0; part < 3; ++part) {
            uint8_t* dst = base + part * kSaTile + s * 64 * 128;
            const int col = part * p.d_model + head * 64;
            if (p.shared_query) tma_load_2d(dst, &tmShr, &full[st], col, 0);
            else tma_load_2d(dst, &tmQry, &full[st], col, pair * nq);
            if (T > 0 && (part > 0 || q_txt)) tma_load_2d(dst + nq * 128, &tmTxt, &full[st], col, RQ + pair * T);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== S = Q K^T =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
    for (int j = 0; j < n_units; ++j) {
      const int st = j % kSaStages, b = j % kSaBufs;
      mbar_wait(&full[st], (j / kSaStages) & 1);
      mbar_wait(&s_free[b], ((j / kSaBufs) & 1) ^ 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t a_desc = umma_desc_k_sw128(smem_u32(sOp + st * SaSmem::kStage));
        const uint64_t b_desc = umma_desc_k_sw128(smem_u32(sOp + st * SaSmem::kStage + kSaTile));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_base + b * 128, a_desc + 2 * k, b_desc + 2 * k, idesc, k > 0 ? 1u : 0u);
        tc_commit(&s_full[b]);
        tc_commit(&empty[st]);
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ===================== O = P V =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);          // B operand MN-major: V stored [key][dim]
    for (int i = 0; i < n_units; ++i) {
      const int st = i % kSaStages, b = i % kSaBufs;
      mbar_wait(&full[st], (i / kSaStages) & 1);         // V landed with Q and K (already complete: the scores exist)
      mbar_wait(&p_ready[b], (i / kSaBufs) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t v_desc = umma_desc_mn_sw128(smem_u32(sOp + st * SaSmem::kStage + 2 * kSaTile), 8192u);
        const uint32_t pa = tmem_base + b * 128;
#pragma unroll
        for (int k = 0; k < 8; ++k)                                     // 16 keys per step = 2048 B of V rows = 8 P columns
          umma_ts(pa + 64, pa + k * 8, v_desc + k * (2048 >> 4), idesc, k > 0 ? 1u : 0u);
        tc_commit(&o_full[b]);
        tc_commit(&empty[st]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================== softmax + epilogue: warpgroup b serves TMEM buffer b =====================
    const int b = (warp - 4) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;                         // tile row
    const int slot = r >> 6;                             // pair slot
    const uint32_t tB = tmem_base + b * 128 + (static_cast<uint32_t>(q * 32) << 16);
    const int S = nq + T;
    const bool elected = (warp & 3) == 0 && lane == 0;
    uint8_t* stage_row = sO + b * kSaTile + r * 128;
    for (int i = b; i < n_units; i += kSaBufs) {
      const uint32_t parity = (i / kSaBufs) & 1;
      const int u = u_begin + i;
      const int tile = u / p.num_heads, head = u % p.num_heads;
      const int pair = 2 * tile + slot;
      const bool pair_ok = pair < p.B;
      // text keys that take part (bit t = key NQ + t is a real, unmasked token); query keys are always valid, keys >= S never exist
      uint32_t keep = 0;
      if (pair_ok)
        for (int t = 0; t < T; ++t)
          keep |= static_cast<uint32_t>(__ldg(p.text_mask + static_cast<size_t>(pair) * T + t) != 0) << t;
      mbar_wait(&s_full[b], parity);
      tc_fence_after();
      uint32_t lo[32], hi[32];                           // scores of keys [0, 32) and [32, 64) of this row's pair
      tmem_ld32(tB + slot * 64, lo);
      tmem_ld32(tB + slot * 64 + 32, hi);
      tmem_ld_wait();
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {                  // keys < 32 are query keys (NQ = 33): never masked
        m0 = fmaxf(m0, __uint_as_float(lo[j]));
        m1 = fmaxf(m1, __uint_as_float(lo[j + 1]));
        m2 = fmaxf(m2, __uint_as_float(lo[j + 2]));
        m3 = fmaxf(m3, __uint_as_float(lo[j + 3]));
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float sc = __uint_as_float(hi[j]);
        if (32 + j >= NQ) {                              // text key / padding: -inf unless the attention mask keeps it
          sc = ((keep >> (32 + j - NQ)) & 1u) ? sc : -INFINITY;
          hi[j] = __float_as_uint(sc);
        }
        if ((j & 3) == 0) m0 = fmaxf(m0, sc);
        else if ((j & 3) == 1) m1 = fmaxf(m1, sc);
        else if ((j & 3) == 2) m2 = fmaxf(m2, sc);
        else m3 = fmaxf(m3, sc);
      }
      const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));     // finite: the query keys are never masked
      const float mxs = mx * p.scale_log2e;
      uint32_t pk[32];
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float e0 = sa_ex2(fmaf(__uint_as_float(lo[2 * j]), p.scale_log2e, -mxs));
        const float e1 = sa_ex2(fmaf(__uint_as_float(lo[2 * j + 1]), p.scale_log2e, -mxs));
        pk[j] = pack_bf16x2(e0, e1);
        s0 += e0;
        s1 += e1;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {                      // keys 32 + 8 g ...: groups past the last text key are skipped
        if (32 + g * 8 < S) {                            // (warp-uniform)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = g * 4 + jj;
            const float e0 = sa_ex2(fmaf(__uint_as_float(hi[2 * j]), p.scale_log2e, -mxs));       // exp2(-inf) = +0
            const float e1 = sa_ex2(fmaf(__uint_as_float(hi[2 * j + 1]), p.scale_log2e, -mxs));
            pk[16 + j] = pack_bf16x2(e0, e1);
            s0 += e0;
            s1 += e1;
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) pk[16 + g * 4 + jj] = 0u;
        }
      }
      const float sum = s0 + s1;
      // P of this pair's 64 keys -> columns [32 slot, 32 slot + 32); zeros over the other pair's keys
      sa_tmem_st32(tB + slot * 32, pk);
#pragma unroll
      for (int j = 0; j < 32; ++j) pk[j] = 0u;
      sa_tmem_st32(tB + (slot ^ 1) * 32, pk);
      sa_tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[b]);

      // ---- epilogue: O / sum -> bf16, this row's 128 bytes straight to global memory (one full line per thread) ----
      mbar_wait(&o_full[b], parity);
      tc_fence_after();
      tmem_ld32(tB + 64, lo);
      tmem_ld32(tB + 96, hi);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[b]);                           // the buffer may take the scores of unit i + 4
      const float inv = 1.f / sum;
      if (elected) tma_store_wait_read<0>();             // this warpgroup's previous stores have drained the staging tile
      named_bar_sync(1 + b, 128);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint32_t (&o)[32] = g < 4 ? lo : hi;
        const int j = (g & 3) * 8;
        uint4 u4;
        u4.x = pack_bf16x2(__uint_as_float(o[j + 0]) * inv, __uint_as_float(o[j + 1]) * inv);
        u4.y = pack_bf16x2(__uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
        u4.z = pack_bf16x2(__uint_as_float(o[j + 4]) * inv, __uint_as_float(o[j + 5]) * inv);
        u4.w = pack_bf16x2(__uint_as_float(o[j + 6]) * inv, __uint_as_float(o[j + 7]) * inv);
        *reinterpret_cast<uint4*>(stage_row + ((g ^ (r & 7)) * 16)) = u4;
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + b, 128);
      if (elected) {
        const int n_pairs = min(2, p.B - 2 * tile);
        for (int s2 = 0; s2 < n_pairs; ++s2) {
          const int pr = 2 * tile + s2;
          tma_store_2d(sO + b * kSaTile + s2 * 64 * 128, &tmOq, head * 64, pr * nq);
          if (p.text_queries && T > 0) tma_store_2d(sO + b * kSaTile + (s2 * 64 + nq) * 128, &tmOt, head * 64, RQ + pr * T);
        }
        tma_store_commit();
      }
    }
    if (elected) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace opsg

using namespace opsg;

// Returns OPSG_E_UNSUPPORTED for shapes the tile layout does not cover (the caller falls back to the mma.sync kernel).
int launch_self_attn_pairs(const opsg_bf16* qkv, const opsg_bf16* shared_query_qkv, const int32_t* text_mask, int B, int n_query,
                           int T, int num_heads, int head_dim, int text_queries, opsg_bf16* ctx_out, cudaStream_t stream) {
  if (head_dim != 64 || n_query != 33 || T < 0 || T > 31 || (T > 0 && !text_mask)) return OPSG_E_UNSUPPORTED;   // the head's shape (v4:155-159)
  if ((((uintptr_t)qkv | (uintptr_t)ctx_out | (uintptr_t)shared_query_qkv) & 15) != 0) return OPSG_E_UNSUPPORTED;
  const int d = num_heads * head_dim;
  const int R = B * (n_query + T);
  CUtensorMap tmQry, tmTxt, tmShr, tmOq, tmOt;
  const int R_out = text_queries ? R : B * n_query;
  int rc = make_tmap_bf16_2d(&tmQry, qkv, (uint64_t)R, (uint64_t)3 * d, (uint64_t)3 * d, n_query, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmTxt, qkv, (uint64_t)R, (uint64_t)3 * d, (uint64_t)3 * d, T > 0 ? T : 1, 64);
  if (rc) return rc;
  rc = shared_query_qkv ? make_tmap_bf16_2d(&tmShr, shared_query_qkv, (uint64_t)n_query, (uint64_t)3 * d, (uint64_t)3 * d, n_query, 64)
                        : OPSG_OK;
  if (rc) return rc;
  if (!shared_query_qkv) tmShr = tmQry;
  rc = make_tmap_bf16_2d(&tmOq, ctx_out, (uint64_t)R_out, (uint64_t)d, (uint64_t)d, n_query, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmOt, ctx_out, (uint64_t)R_out, (uint64_t)d, (uint64_t)d, T > 0 ? T : 1, 64);
  if (rc) return rc;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[device_slot()];
  if (!configured) {
    rc = check_cuda(cudaFuncSetAttribute(self_attn_pairs_kernel<33>, cudaFuncAttributeMaxDynamicSharedMemorySize, SaSmem::kTotal),
                    "cudaFuncSetAttribute(self_attn_pairs)");
    if (rc) return rc;
    configured = true;
  }
  SaParams p;
  p.text_mask = text_mask;
  p.out = reinterpret_cast<__nv_bfloat16*>(ctx_out);
  p.B = B; p.n_query = n_query; p.T = T; p.num_heads = num_heads; p.d_model = d;
  p.text_queries = text_queries ? 1 : 0;
  p.shared_query = shared_query_qkv ? 1 : 0;
  p.total_units = ((B + 1) / 2) * num_heads;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(head_dim));
  const int sms = opsg_num_sms();
  const int grid = p.total_units < sms ? p.total_units : sms;
  launch_kernel(self_attn_pairs_kernel<33>, grid, kSaThreads, SaSmem::kTotal, stream, tmQry, tmTxt, tmShr, tmOq, tmOt, p);
  OPSG_CHECK_LAUNCH("self_attn_pairs_kernel");
  return OPSG_OK;
}
