// K10a (tcgen05) — causal multi-head attention of the LLM PREFILL (row a10 of SURVEY.md §8): the [32 projected relation rows ;
// left-padded instruction tokens] prompt of every selected pair (relation_transformer_head_v4.py:294-312; HF OPT
// modeling_opt.py:135-181, HF Llama modeling_llama.py:199-262), all pairs batched.
//
// A prompt is q_len <= 64 tokens, so TWO sequences are stacked into one 128-row tensor-core tile (block-diagonal, as K4 does
// with pairs): sequence slot s owns tile rows / keys [64 s, 64 s + q_len).  Work unit = (2 sequences, head).
//   warp 0      TMA producer: Q rows from the fused qkv activation, K / V rows from the static caches, head_dim 64 / 80 / 128 as
//               one or two 64-column panels (the second panel of head_dim 80 carries 16 live columns), 2 stages
//   warp 1      S = Q K^T : head_dim / 16 tcgen05.mma (SS, M = 128, N = 128) into one of two 256-column TMEM buffers
//   warp 2      TMEM allocator, then O = P V : 8 tcgen05.mma (A = P from TMEM, B = V row-major = MN-major operand over both
//               panels, N = head_dim) into columns [64, 64 + head_dim) of the buffer
//   warps 4-11  two softmax / epilogue warpgroups, thread = one query row: its sequence's 64 score columns in one TMEM round
//               trip, mask = causal & key-validity (left padding), max, exp2, bf16 P in place (+ zeros over the other
//               sequence's half), row sum in a register; then O / sum -> bf16 -> global.  Rows that see no key at all (queries
//               sitting on padding) are written as zeros: finite, never attended to.
// Longer prompts (q_len > 64), a non-zero first position or other head sizes take the warp-level kernel (attention_small.cu).
#include <math.h>

#include "common.cuh"
#include "host_util.h"

namespace opsg {

constexpr int kPfThreads = 384;
constexpr int kPfBufs = 2;
constexpr int kPfStages = 2;
constexpr int kPfPanel = 128 * 128;        // one 64-column panel of an operand tile: 128 rows x 128 B

struct PfParams {
  const uint8_t* key_mask;   // [nseq, max_ctx], 0 = padded key
  __nv_bfloat16* out;        // [nseq * q_len, ld_out]
  int nseq, q_len, num_heads, max_ctx, ld_out;
  int total_units;           // ceil(nseq / 2) * num_heads
  float scale_log2e;
};

__device__ __forceinline__ void pf_tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
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
__device__ __forceinline__ void pf_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void pf_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float pf_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int HD>
struct PfSmem {
  static constexpr int kNP = HD > 64 ? 2 : 1;                        // 64-column panels per operand
  static constexpr int kOperand = kNP * kPfPanel;
  static constexpr int kStage = 3 * kOperand;                        // Q | K | V
  static constexpr int kOffBar = kPfStages * kStage;
  static constexpr int kTotal = kOffBar + 256 + 1024;
};

template <int HD>
__global__ void __launch_bounds__(kPfThreads, 1)
llm_prefill_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const PfParams p) {
  using SM = PfSmem<HD>;
  constexpr int NP = SM::kNP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* full = bars;                       // [2]
  uint64_t* empty = bars + 2;                  // [2] (count 2: the QK^T and the PV commits)
  uint64_t* s_full = bars + 4;                 // [2]
  uint64_t* p_ready = bars + 6;                // [2] (count 128)
  uint64_t* o_full = bars + 8;                 // [2]
  uint64_t* s_free = bars + 10;                // [2] (count 128)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < kPfStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); }
    for (int b = 0; b < kPfBufs; ++b) {
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
  // rows no TMA box writes (the padding of each 64-row sequence slot) must be finite: 0 x NaN would poison P V
  for (int idx = threadIdx.x; idx < kPfStages * SM::kStage / 16; idx += kPfThreads)
    reinterpret_cast<uint4*>(smem)[idx] = make_uint4(0u, 0u, 0u, 0u);
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
  const int L = p.q_len;

  if (warp == 0) {
    // ===================== TMA producer =====================
    for (int i = 0; i < n_units; ++i) {
      const int u = u_begin + i;
      const int tile = u / p.num_heads, head = u % p.num_heads;
      const int st = i % kPfStages;
      mbar_wait(&empty[st], ((i / kPfStages) & 1) ^ 1);
      if (elect_one_sync()) {
        const int n_seq = min(2, p.nseq - 2 * tile);
        mbar_expect_tx(&full[st], static_cast<uint32_t>(n_seq) * 3u * NP * L * 128u);
        uint8_t* base = smem + st * SM::kStage;
        for (int s = 0; s < n_seq; ++s) {
          const int seq = 2 * tile + s;
#pragma unroll
          for (int pn = 0; pn < NP; ++pn) {
            const int col = head * HD + pn * 64;
            tma_load_2d(base + 0 * SM::kOperand + pn * kPfPanel + s * 64 * 128, &tmQ, &full[st], col, seq * L);
            tma_load_2d(base + 1 * SM::kOperand + pn * kPfPanel + s * 64 * 128, &tmK, &full[st], col, seq * p.max_ctx);
            tma_load_2d(base + 2 * SM::kOperand + pn * kPfPanel + s * 64 * 128, &tmV, &full[st], col, seq * p.max_ctx);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== S = Q K^T =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
    for (int j = 0; j < n_units; ++j) {
      const int st = j % kPfStages, b = j % kPfBufs;
      mbar_wait(&full[st], (j / kPfStages) & 1);
      mbar_wait(&s_free[b], ((j / kPfBufs) & 1) ^ 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t qa = smem_u32(smem + st * SM::kStage), ka = qa + SM::kOperand;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint64_t a_desc = umma_desc_k_sw128(qa + (k / 4) * kPfPanel) + 2 * (k % 4);
          const uint64_t b_desc = umma_desc_k_sw128(ka + (k / 4) * kPfPanel) + 2 * (k % 4);
          umma_ss(tmem_base + b * 256, a_desc, b_desc, idesc, k > 0 ? 1u : 0u);
        }
        tc_commit(&s_full[b]);
        tc_commit(&empty[st]);
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ===================== O = P V =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, HD, 0, 1);          // B operand MN-major: V stored [key][dim]
    for (int i = 0; i < n_units; ++i) {
      const int st = i % kPfStages, b = i % kPfBufs;
      mbar_wait(&full[st], (i / kPfStages) & 1);
      mbar_wait(&p_ready[b], (i / kPfBufs) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        // MN-major, 128-byte swizzle: 8-key groups 1024 B apart, the two 64-dim blocks one panel apart (LBO)
        const uint64_t v_desc = umma_desc_mn_sw128(smem_u32(smem + st * SM::kStage + 2 * SM::kOperand), kPfPanel);
        const uint32_t pa = tmem_base + b * 256;
#pragma unroll
        for (int k = 0; k < 8; ++k)
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
    const int r = q * 32 + lane;
    const int slot = r >> 6, qi = r & 63;                // sequence slot, query position inside the prompt
    const uint32_t tB = tmem_base + b * 256 + (static_cast<uint32_t>(q * 32) << 16);
    for (int i = b; i < n_units; i += kPfBufs) {
      const uint32_t parity = (i / kPfBufs) & 1;
      const int u = u_begin + i;
      const int tile = u / p.num_heads, head = u % p.num_heads;
      const int seq = 2 * tile + slot;                   // (warp-uniform: a warp's rows belong to one slot)
      const bool seq_ok = seq < p.nseq;
      // key-validity bits of this sequence (left padding), one ballot per 32 keys
      const uint8_t* km = p.key_mask + static_cast<size_t>(seq_ok ? seq : 0) * p.max_ctx;
      const uint32_t w0 = __ballot_sync(0xffffffffu, seq_ok && lane < L && __ldg(km + lane) != 0);
      const uint32_t w1 = __ballot_sync(0xffffffffu, seq_ok && lane + 32 < L && __ldg(km + min(lane + 32, p.max_ctx - 1)) != 0);
      const uint64_t causal = (qi >= 63) ? ~0ull : ((2ull << qi) - 1ull);
      const uint64_t valid = (static_cast<uint64_t>(w0) | (static_cast<uint64_t>(w1) << 32)) & causal;
      mbar_wait(&s_full[b], parity);
      tc_fence_after();
      uint32_t lo[32], hi[32];
      tmem_ld32(tB + slot * 64, lo);
      tmem_ld32(tB + slot * 64 + 32, hi);
      tmem_ld_wait();
      const uint32_t vlo = static_cast<uint32_t>(valid), vhi = static_cast<uint32_t>(valid >> 32);
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float a = ((vlo >> j) & 1u) ? __uint_as_float(lo[j]) : -INFINITY;
        const float c = ((vhi >> j) & 1u) ? __uint_as_float(hi[j]) : -INFINITY;
        lo[j] = __float_as_uint(a);
        hi[j] = __float_as_uint(c);
        m0 = fmaxf(m0, a);
        m1 = fmaxf(m1, c);
      }
      float mx = fmaxf(m0, m1);
      if (mx == -INFINITY) mx = 0.f;                     // a query on padding: no key at all
      const float mxs = mx * p.scale_log2e;
      uint32_t pk[32];
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float e0 = pf_ex2(fmaf(__uint_as_float(lo[2 * j]), p.scale_log2e, -mxs));           // exp2(-inf) = +0
        const float e1 = pf_ex2(fmaf(__uint_as_float(lo[2 * j + 1]), p.scale_log2e, -mxs));
        pk[j] = pack_bf16x2(e0, e1);
        s0 += e0;
        s1 += e1;
      }
      if (L > 32) {                                      // (uniform)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float e0 = pf_ex2(fmaf(__uint_as_float(hi[2 * j]), p.scale_log2e, -mxs));
          const float e1 = pf_ex2(fmaf(__uint_as_float(hi[2 * j + 1]), p.scale_log2e, -mxs));
          pk[16 + j] = pack_bf16x2(e0, e1);
          s0 += e0;
          s1 += e1;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[16 + j] = 0u;
      }
      const float sum = s0 + s1;
      pf_tmem_st32(tB + slot * 32, pk);                  // P of this sequence's keys; zeros over the other sequence's keys
#pragma unroll
      for (int j = 0; j < 32; ++j) pk[j] = 0u;
      pf_tmem_st32(tB + (slot ^ 1) * 32, pk);
      pf_tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[b]);

      // ---- epilogue: O / sum -> bf16 -> global (16 bytes per store) ----
      mbar_wait(&o_full[b], parity);
      tc_fence_after();
      const bool store = seq_ok && qi < L;
      const float inv = sum > 0.f ? 1.f / sum : 0.f;     // fully masked rows -> zeros (finite)
      __nv_bfloat16* dst = p.out + (static_cast<size_t>(seq_ok ? seq : 0) * L + min(qi, L - 1)) * p.ld_out + head * HD;
#pragma unroll
      for (int c = 0; c < HD / 16; ++c) {
        uint32_t o[16];
        pf_tmem_ld16(tB + 64 + c * 16, o);
        tmem_ld_wait();
        if (store) {
          uint4 a, d;
          a.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
          a.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
          a.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
          a.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
          d.x = pack_bf16x2(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
          d.y = pack_bf16x2(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
          d.z = pack_bf16x2(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
          d.w = pack_bf16x2(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
          reinterpret_cast<uint4*>(dst + c * 16)[0] = a;
          reinterpret_cast<uint4*>(dst + c * 16)[1] = d;
        }
      }
      tc_fence_before();
      mbar_arrive(&s_free[b]);                           // the buffer may take the scores of unit i + 2
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int HD>
static int launch_prefill_hd(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const PfParams& p,
                             cudaStream_t stream) {
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[device_slot()];
  if (!configured) {
    int rc = check_cuda(cudaFuncSetAttribute(llm_prefill_attn_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             PfSmem<HD>::kTotal), "cudaFuncSetAttribute(llm_prefill_attn)");
    if (rc) return rc;
    configured = true;
  }
  const int sms = opsg_num_sms();
  const int grid = p.total_units < sms ? p.total_units : sms;
  launch_kernel(llm_prefill_attn_kernel<HD>, grid, kPfThreads, PfSmem<HD>::kTotal, stream, tmQ, tmK, tmV, p);
  OPSG_CHECK_LAUNCH("llm_prefill_attn_kernel");
  return OPSG_OK;
}

}  // namespace opsg

using namespace opsg;

// Returns OPSG_E_UNSUPPORTED for shapes the tile layout does not cover (the caller falls back to the mma.sync kernel).
int launch_llm_prefill_attn_tc(const opsg_bf16* q, int ld_q, const opsg_bf16* k_cache, const opsg_bf16* v_cache, int max_ctx,
                               const uint8_t* key_mask, int nseq, int q_len, int q_pos0, int num_heads, int head_dim, float scale,
                               opsg_bf16* out, int ld_out, cudaStream_t stream) {
  if (q_pos0 != 0 || q_len < 2 || q_len > 64 || (head_dim != 64 && head_dim != 80 && head_dim != 128)) return OPSG_E_UNSUPPORTED;
  const int d = num_heads * head_dim;
  if ((ld_q % 8) || (ld_out % 8) || ld_q < d || ld_out < d ||
      ((((uintptr_t)q | (uintptr_t)k_cache | (uintptr_t)v_cache | (uintptr_t)out) & 15) != 0))
    return OPSG_E_UNSUPPORTED;
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_tmap_bf16_2d(&tmQ, q, (uint64_t)nseq * q_len, (uint64_t)ld_q, (uint64_t)ld_q, q_len, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmK, k_cache, (uint64_t)nseq * max_ctx, (uint64_t)d, (uint64_t)d, q_len, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmV, v_cache, (uint64_t)nseq * max_ctx, (uint64_t)d, (uint64_t)d, q_len, 64);
  if (rc) return rc;
  PfParams p;
  p.key_mask = key_mask;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.nseq = nseq; p.q_len = q_len; p.num_heads = num_heads; p.max_ctx = max_ctx; p.ld_out = ld_out;
  p.total_units = ((nseq + 1) / 2) * num_heads;
  p.scale_log2e = 1.4426950408889634f * scale;
  if (head_dim == 64) return launch_prefill_hd<64>(tmQ, tmK, tmV, p, stream);
  if (head_dim == 80) return launch_prefill_hd<80>(tmQ, tmK, tmV, p, stream);
  return launch_prefill_hd<128>(tmQ, tmK, tmV, p, stream);
}
