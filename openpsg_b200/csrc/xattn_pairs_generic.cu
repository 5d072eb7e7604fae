// K5, generic variant — the self-contained kernel opsg_xattn_pairs falls back to when the caller passes no prebuilt mask-bias
// tiles or n_query puts more than 8 pairs into a 128-row tile (xattn_pairs.cu handles the head's own shape, n_query = 33):
// pair-query x image-feature masked cross-attention on tcgen05 tensor cores, masks applied by the softmax warps from the bit words.
//
// All pairs' query rows are stacked along M (row = pair * n_query + r); K and V are projected once per image
// and shared by every pair, so per head the whole thing is  softmax(Q[M x 64] . K^T[64 x L] + mask(pair)) . V.
// Work unit = (128-row tile, head); a persistent CTA walks a contiguous, head-major range of units (at most two
// heads per CTA, each head's K [256 x 64] and V^T [80 x 256] stay resident in their own shared-memory set).
//
// 384 threads:
//   warp 0 (1 thread)  TMA producer : K / V^T once per head, Q tile per unit (2 stages)
//   warp 1 (1 thread)  MMA issuer   : S_b = Q.K^T    (128 x 256 x 64, SS, 4 MMAs)   -> TMEM cols [256b, 256b+256)
//                                     O_b = P_b.[V|1] (128 x 80 x 256, TS, 16 MMAs)  -> TMEM cols [256b+128, 256b+208)
//   warp 2             TMEM allocator (512 columns = two S buffers)
//   warps 4-7 / 8-11   softmax warpgroup 0 / 1: units alternate between the two warpgroups and the two TMEM
//                      buffers, so the MMAs of unit i+1 overlap the softmax of unit i.  Thread = score row:
//                      row max over the pair's keys (mask = bits[i] | bits[j], never materialised in HBM),
//                      p = exp2(s*scale - max), P written back IN PLACE over S as packed bf16 (tcgen05.st; it is
//                      the A operand of the PV MMA straight from TMEM), row sum produced by the MMA itself through a
//                      ones row appended to V^T, O/rowsum -> bf16 -> 128B-swizzled smem -> one TMA store per unit.
// Mask semantics follow HF's `(1 - m) * finfo.min` additive bias: masked keys get weight exactly 0 and a pair
// whose union mask is empty attends uniformly to all L keys.
//
// Why this shape (profiles/r1_ncu_xattn_a.md): v1 ran one softmax warp per scheduler strictly serial with the
// MMAs (9.7 % tensor-pipe activity, issue slots 28 % busy, 11.5 k cycles per unit for 1024 cycles of MMA work).
#include "common.cuh"
#include "host_util.h"

namespace opsg {
namespace xa2 {

constexpr int kXaThreads = 384;
constexpr int kXaKeys = 256;       // max keys (one N=256 MMA)
constexpr int kXaHd = 64;
constexpr int kXaPvN = 80;         // PV MMA N: 64 value dims + 1 ones row (row sum) + 15 zero rows

struct XattnParams {
  const uint32_t* bits;
  const int32_t* pair_index;
  int words, num_objects, n_query, L, num_heads, d_model;
  int rows;          // B * n_query
  int m_tiles;
  int total_units;
  int flags;         // bit 0: row sum by FADD in registers instead of the ones row (debug / A-B)
  float scale_log2e;
};

struct XaSmem {
  static constexpr int kKSet = kXaKeys * 128;           // 32768: K [256 keys x 64] SW128 K-major
  static constexpr int kVBlk = kXaPvN * 128;            // 10240: one 64-key block of [V^T | 1 | 0] (80 rows x 128 B)
  static constexpr int kVBox = 64 * 128;                // 8192 : bytes TMA writes into a block (rows 0-63)
  static constexpr int kVSet = 4 * kVBlk;               // 40960
  static constexpr int kQ = 128 * 128;                  // 16384 per stage
  static constexpr int kOst = 128 * 128;                // 16384 per warpgroup: O staging for the TMA store
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kOffK + 2 * kKSet;       // 65536
  static constexpr int kOffQ = kOffV + 2 * kVSet;       // 147456
  static constexpr int kOffO = kOffQ + 2 * kQ;          // 180224
  static constexpr int kOffBar = kOffO + 2 * kOst;      // 212992
  static constexpr int kTotal = kOffBar + 256 + 1024;
};
static_assert(XaSmem::kTotal <= 232448, "shared memory budget exceeded");

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// registers -> TMEM: this thread's lane (row), 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}

__global__ void __launch_bounds__(kXaThreads, 1)
xattn_pairs_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ CUtensorMap tmO,
                   const XattnParams p) {
  pdl_wait_then_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem + XaSmem::kOffK;
  uint8_t* sV = smem + XaSmem::kOffV;
  uint8_t* sQ = smem + XaSmem::kOffQ;
  uint8_t* sO = smem + XaSmem::kOffO;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + XaSmem::kOffBar);
  uint64_t* q_full = bars;          // [2]
  uint64_t* q_empty = bars + 2;     // [2]
  uint64_t* kv_full = bars + 4;     // [2]  one per head set
  uint64_t* s_full = bars + 6;      // [2]
  uint64_t* p_ready = bars + 8;     // [2]
  uint64_t* o_full = bars + 10;     // [2]
  uint64_t* s_free = bars + 12;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmVt);
    tma_prefetch_desc(&tmO);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&q_full[b], 1);
      mbar_init(&q_empty[b], 1);
      mbar_init(&kv_full[b], 1);
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
  // rows 64..79 of every [V^T | 1 | 0] block: row 64 = ones (the PV MMA then also produces the row sum), rest zero.
  // Row 64 has (row & 7) == 0 and all its 16-byte chunks are equal, so the 128B swizzle does not matter here.
  for (int idx = threadIdx.x; idx < 8 * 128; idx += kXaThreads) {
    const int blk = idx >> 7, within = idx & 127;                // 8 blocks (2 sets x 4), 16 rows x 8 chunks each
    const uint32_t one2 = (within < 8) ? 0x3F803F80u : 0u;        // bf16 1.0 pairs in row 64
    uint8_t* dst = sV + (blk >> 2) * XaSmem::kVSet + (blk & 3) * XaSmem::kVBlk + XaSmem::kVBox + within * 16;
    *reinterpret_cast<uint4*>(dst) = make_uint4(one2, one2, one2, one2);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // contiguous, head-major unit range of this CTA (spans at most two heads: host guarantees per <= m_tiles)
  const int per = (p.total_units + gridDim.x - 1) / gridDim.x;
  const int u_begin = blockIdx.x * per;
  const int u_end = min(p.total_units, u_begin + per);
  const int n_units = max(0, u_end - u_begin);
  const int head0 = u_begin / p.m_tiles;

  if (threadIdx.x == 0) {
    // ===================== TMA producer =====================
    int cur_head = -1;
    for (int i = 0; i < n_units; ++i) {
      const int u = u_begin + i;
      const int head = u / p.m_tiles, mt = u % p.m_tiles;
      if (head != cur_head) {
        const int set = head - head0;
        mbar_expect_tx(&kv_full[set], XaSmem::kKSet + 4 * XaSmem::kVBox);
        tma_load_2d(sK + set * XaSmem::kKSet, &tmK, &kv_full[set], head * kXaHd, 0);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
          tma_load_2d(sV + set * XaSmem::kVSet + kb * XaSmem::kVBlk, &tmVt, &kv_full[set], kb * 64, head * kXaHd);
        cur_head = head;
      }
      const int b = i & 1;
      mbar_wait(&q_empty[b], ((i >> 1) & 1) ^ 1);
      mbar_expect_tx(&q_full[b], XaSmem::kQ);
      tma_load_2d(sQ + b * XaSmem::kQ, &tmQ, &q_full[b], head * kXaHd, mt * 128);
    }
  } else if (threadIdx.x == 32) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, kXaKeys);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, kXaPvN);
    bool kv_ready[2] = {false, false};
    auto issue_qk = [&](int j) {
      const int set = (u_begin + j) / p.m_tiles - head0;
      const int b = j & 1;
      if (!kv_ready[set]) {
        mbar_wait(&kv_full[set], 0);
        kv_ready[set] = true;
      }
      mbar_wait(&q_full[b], (j >> 1) & 1);
      mbar_wait(&s_free[b], ((j >> 1) & 1) ^ 1);     // O of unit j-2 has been read out of this buffer
      tc_fence_after();
      const uint32_t a = smem_u32(sQ + b * XaSmem::kQ);
      const uint32_t kk = smem_u32(sK + set * XaSmem::kKSet);
      const uint32_t d = tmem_base + b * 256;
#pragma unroll
      for (int k = 0; k < kXaHd / 16; ++k)
        umma_ss(d, umma_desc_k_sw128(a + k * 32), umma_desc_k_sw128(kk + k * 32), idesc_qk, k > 0 ? 1u : 0u);
      tc_commit(&q_empty[b]);
      tc_commit(&s_full[b]);
    };
    auto issue_pv = [&](int i) {
      const int set = (u_begin + i) / p.m_tiles - head0;
      const int b = i & 1;
      mbar_wait(&p_ready[b], (i >> 1) & 1);          // P_b complete in TMEM, S_b fully consumed
      tc_fence_after();
      const uint32_t pa = tmem_base + b * 256;        // P: 128 columns of packed bf16 pairs (keys 2c, 2c+1)
      const uint32_t od = tmem_base + b * 256 + 128;  // O: 80 fp32 columns
      const uint32_t vb = smem_u32(sV + set * XaSmem::kVSet);
#pragma unroll
      for (int k = 0; k < kXaKeys / 16; ++k)
        umma_ts(od, pa + k * 8, umma_desc_k_sw128(vb + (k >> 2) * XaSmem::kVBlk + (k & 3) * 32), idesc_pv, k > 0 ? 1u : 0u);
      tc_commit(&o_full[b]);
    };
    if (n_units > 0) issue_qk(0);
    if (n_units > 1) issue_qk(1);
    for (int i = 0; i < n_units; ++i) {
      issue_pv(i);
      if (i + 2 < n_units) issue_qk(i + 2);
    }
  } else if (warp >= 4) {
    // ===================== softmax + epilogue warpgroups =====================
    const int wg = (warp - 4) >> 2;                      // 0 / 1 = TMEM buffer = unit parity
    const int q = warp & 3;                              // TMEM lane quarter of this warp (warp id % 4)
    const int r = q * 32 + lane;                         // row inside the tile
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + wg * 256 + lane_off;
    const uint32_t tO = tS + 128;
    const bool elected = (warp == 4 + 4 * wg) && lane == 0;
    uint8_t* stage_row = sO + wg * XaSmem::kOst + r * 128;
    const int full_words = p.L >> 5;
    const uint32_t tail_mask = (p.L & 31) ? ((1u << (p.L & 31)) - 1u) : 0u;
    const bool sum_in_regs = (p.flags & 1) != 0;

    for (int i = wg; i < n_units; i += 2) {
      const uint32_t parity = (i >> 1) & 1;
      const int u = u_begin + i;
      const int head = u / p.m_tiles, mt = u % p.m_tiles;
      const int row = mt * 128 + r;
      const bool valid = row < p.rows;
      // pair mask = bits[i] | bits[j], restricted to the L real keys
      uint32_t m[8];
      bool empty = true;
      {
        int oi = 0, oj = 0;
        if (valid) {
          const int pair = row / p.n_query;
          const int pidx = p.pair_index ? p.pair_index[pair] : pair;
          oi = pidx / p.num_objects;
          oj = pidx % p.num_objects;
        }
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const uint32_t keyok = (w < full_words) ? 0xffffffffu : (w == full_words ? tail_mask : 0u);
          uint32_t v = 0;
          if (valid && w < p.words)
            v = __ldg(p.bits + static_cast<size_t>(oi) * p.words + w) | __ldg(p.bits + static_cast<size_t>(oj) * p.words + w);
          m[w] = v & keyok;
          empty = empty && (m[w] == 0);
        }
        if (empty) {   // finfo.min on every key -> uniform attention over the L real keys
#pragma unroll
          for (int w = 0; w < 8; ++w) m[w] = (w < full_words) ? 0xffffffffu : (w == full_words ? tail_mask : 0u);
        }
      }
      const float sc = empty ? 0.f : p.scale_log2e;       // empty: every real key gets exp2(0) = 1

      mbar_wait(&s_full[wg], parity);
      tc_fence_after();
      // pass 1: row max over unmasked keys
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t v[32];
        tmem_ld32(tS + c * 32, v);
        tmem_ld_wait();
        const uint32_t mw = m[c];
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, ((mw >> j) & 1u) ? __uint_as_float(v[j]) : -INFINITY);
      }
      if (mx == -INFINITY) mx = 0.f;
      const float mxs = mx * sc;
      // pass 2: p = exp2(s*scale - max*scale) on the pair's keys, 0 elsewhere; packed bf16 P overwrites S in place
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t v[32];
        tmem_ld32(tS + c * 32, v);
        tmem_ld_wait();
        const uint32_t mw = m[c];
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float e0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), sc, -mxs));
          const float e1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sc, -mxs));
          const float p0 = ((mw >> (2 * j)) & 1u) ? e0 : 0.f;
          const float p1 = ((mw >> (2 * j + 1)) & 1u) ? e1 : 0.f;
          if (sum_in_regs) sum += p0 + p1;
          pk[j] = pack_bf16x2(p0, p1);
        }
        tmem_st16(tS + c * 16, pk);      // P cols [16c, 16c+16) only overlap S chunks <= c (already consumed)
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[wg]);

      // epilogue: O / rowsum -> bf16 -> swizzled staging row -> TMA store of the [128 x 64] tile
      mbar_wait(&o_full[wg], parity);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld32(tO, o0);
      tmem_ld32(tO + 32, o1);
      const uint32_t osum = tmem_ld1(tO + 64);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[wg]);                          // buffer may be overwritten by QK^T of unit i+2
      const float inv = 1.f / (sum_in_regs ? sum : __uint_as_float(osum));
      if (elected) tma_store_wait_read<0>();             // previous unit's store has drained this staging tile
      named_bar_sync(1 + wg, 128);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint32_t* src = (g < 4) ? (o0 + g * 8) : (o1 + (g - 4) * 8);
        uint4 u4;
        u4.x = pack_bf16x2(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
        u4.y = pack_bf16x2(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
        u4.z = pack_bf16x2(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
        u4.w = pack_bf16x2(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
        *reinterpret_cast<uint4*>(stage_row + ((g ^ (r & 7)) * 16)) = u4;
      }
      fence_proxy_async_smem();
      named_bar_sync(3 + wg, 128);
      if (elected) {
        tma_store_2d(sO + wg * XaSmem::kOst, &tmO, head * kXaHd, mt * 128);
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

}  // namespace xa2
}  // namespace opsg

using namespace opsg;
using namespace opsg::xa2;

extern "C" int opsg_xattn_pairs_v2(const opsg_bf16* q, const opsg_bf16* k, int ld_k, const opsg_bf16* vt, int ld_vt,
                                const uint32_t* bits, int words, const int32_t* pair_index, int num_objects, int B,
                                int n_query, int L, int num_heads, int head_dim, opsg_bf16* ctx_out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(q && k && vt && bits && ctx_out, "xattn_pairs: null pointer");
  OPSG_CHECK_ARG(B > 0 && n_query > 0 && L > 0 && num_heads > 0 && num_objects > 0, "xattn_pairs: bad shape");
  if (head_dim != kXaHd) return set_error(OPSG_E_UNSUPPORTED, "xattn_pairs: head_dim %d unsupported (64 only)", head_dim);
  if (L > kXaKeys) return set_error(OPSG_E_UNSUPPORTED, "xattn_pairs: L=%d image tokens > %d unsupported", L, kXaKeys);
  OPSG_CHECK_ARG(words >= (L + 31) / 32 && words <= 8, "xattn_pairs: words=%d inconsistent with L=%d", words, L);
  const int d_model = num_heads * head_dim;
  OPSG_CHECK_ARG(ld_k >= d_model && ld_k % 8 == 0 && ld_vt >= L && ld_vt % 8 == 0, "xattn_pairs: bad leading dims");
  OPSG_CHECK_ARG(((uintptr_t)ctx_out & 15) == 0, "xattn_pairs: ctx_out must be 16-byte aligned");
  const int rows = B * n_query;
  CUtensorMap tmQ, tmK, tmVt, tmO;
  rc = make_tmap_bf16_2d(&tmQ, q, (uint64_t)rows, (uint64_t)d_model, (uint64_t)d_model, 128, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmK, k, (uint64_t)L, (uint64_t)d_model, (uint64_t)ld_k, kXaKeys, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmVt, vt, (uint64_t)d_model, (uint64_t)L, (uint64_t)ld_vt, 64, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmO, ctx_out, (uint64_t)rows, (uint64_t)d_model, (uint64_t)d_model, 128, 64);
  if (rc) return rc;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[device_slot()];
  if (!configured) {
    rc = check_cuda(cudaFuncSetAttribute(xattn_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XaSmem::kTotal),
                    "cudaFuncSetAttribute(xattn)");
    if (rc) return rc;
    configured = true;
  }
  XattnParams p;
  p.bits = bits; p.pair_index = pair_index;
  p.words = words; p.num_objects = num_objects; p.n_query = n_query; p.L = L; p.num_heads = num_heads; p.d_model = d_model;
  p.rows = rows; p.m_tiles = (rows + 127) / 128; p.total_units = p.m_tiles * num_heads;
  p.flags = 0;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(head_dim));
  const int grid = p.total_units < opsg_num_sms() ? p.total_units : opsg_num_sms();
  // each CTA's contiguous unit range must span at most two heads (two resident K/V sets)
  const int per = (p.total_units + grid - 1) / grid;
  if (per > p.m_tiles) return set_error(OPSG_E_UNSUPPORTED, "xattn_pairs: unit range %d spans more than two heads", per);
  launch_kernel(xattn_pairs_kernel, grid, kXaThreads, XaSmem::kTotal, reinterpret_cast<cudaStream_t>(stream), tmQ, tmK, tmVt, tmO, p);
  OPSG_CHECK_LAUNCH("xattn_pairs_kernel");
  return OPSG_OK;
}
