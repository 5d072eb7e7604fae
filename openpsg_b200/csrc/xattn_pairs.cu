// K5 — pair-query x image-feature masked cross-attention on tcgen05 tensor cores (the north-star kernel; reference
// HF modeling_instructblip.py:499-536 as called at relation_transformer_head_v4.py:179-185).
//
// All pairs' query rows are stacked along M (row = pair * n_query + r); K and V are projected once per image and shared by
// every pair, so per head the whole thing is  softmax(Q[M x 64] . K^T[64 x L] + mask(pair)) . V.
// Work unit = (128-row tile, head).  The grid is num_heads x ctas_per_head persistent CTAs; a CTA walks m-tiles of ONE head
// with that head's K [256 x 64] and [V^T | 1 | 0] [80 x 256] resident in shared memory.
//
// 512 threads (<= 128 registers):
//   warp 0      TMA producer: K / V^T once, then per unit the Q tile + the unit's mask-bias operand tiles (2 stages)
//   warp 1      QK^T issuer:  S_b = Q.K^T + A_aug.B_aug^T  (128 x 256 x (64+16), SS, 5 MMAs) -- its own in-order queue, so it
//               is never blocked behind a PV that waits for P
//   warp 2      TMEM allocator (512 columns = two 256-column buffers), then PV issuer: O_b = P_b.[V | 1] (128 x 80 x 256, TS, 16 MMAs;
//               the ones row makes the MMA deliver the row sum as column 64)
//   warp 3      per-head mean of V (the output of uniform-attention rows)
//   warps 4-7   softmax of TMEM buffer 0, warps 8-11 of buffer 1: thread = one score row x all 256 keys, so the row max
//               needs no cross-thread exchange.  p = exp2(s*scale - max) with no per-element mask work; packed bf16 P
//               overwrites the consumed scores in place (columns [0, 128)) and is the TMEM A operand of the PV MMA
//   warps 12-15 epilogue of both buffers: O / row sum out of TMEM columns [128, 209) (which frees the buffer for the QK^T of unit i+2
//               as soon as it has been read), 1 / row sum, uniform rows, bf16, swizzled staging tile, one TMA store per unit
// so a softmax warp goes straight from the exponentials of unit i to the scores of unit i+2.
//
// Mask inside the MMA: the pair mask enters the scores as an additive bias computed BY THE TENSOR CORE -- a fifth k-step
// multiplies A_aug[r, t] = 1 for the tile-local pair slot t of row r with B_aug[key, t] = 0 if the key belongs to
// bits[i_t] | bits[j_t] else -16384 (bf16-exact; keys >= L get -16384 in every slot).  Both operand tiles are built once per
// image by xattn_bias_tiles_kernel (they depend on the masks and the pair order only, not on the head or the layer), stored
// in the MMA's no-swizzle core-matrix order and fetched with one bulk copy per unit next to the Q tile.
// Sparse softmax: the image tokens reach this kernel sorted by owning object (token_order_kernel), so a pair's visible keys
// are two short runs; 16-key chunks in which no row of a warp sees a key are skipped (no tcgen05.ld, no max, no
// exponentials; their P columns are cleared).  The MMAs stay dense.
// Mask semantics follow HF's `(1 - m) * finfo.min` additive bias: masked keys get weight exactly 0 (exp2(-16384 * scale)
// underflows to +0) and a pair whose union mask is empty attends uniformly to all L keys (those rows are written as the fp32
// mean of V).
//
// History: profiles/r1_ncu_xattn_a.md (v1: 9.7 % tensor-pipe activity, serial softmax), r1_ncu_xattn_h.md (v2: 20 %, ALU pipe
// busy with per-element mask work), r1_ncu_xattn_final3.md (v3: mask as an MMA k-step, 34 %), r2_ncu_xattn.md (v4: split
// issuer warps, thread-per-row softmax, key order + chunk skipping: 43 % of elapsed at cfg2, 54 % at the cfg5 shape).
// Tried in round 2 and dropped: row sums from the softmax threads + PV at N = 64 (55.3 us instead of 56.4 under ncu, but 6 %
// slower on dense masks and 40 % tensor-pipe activity instead of 43: the unit time is not set by the PV MMAs).
#include "common.cuh"
#include "host_util.h"

extern "C" int opsg_xattn_pairs_v2(const opsg_bf16* q, const opsg_bf16* k, int ld_k, const opsg_bf16* vt, int ld_vt,
                                   const uint32_t* bits, int words, const int32_t* pair_index, int num_objects, int B,
                                   int n_query, int L, int num_heads, int head_dim, opsg_bf16* ctx_out, void* stream);

namespace opsg {

int launch_xattn_pairs_long(const opsg_bf16* q, const opsg_bf16* k, int ld_k, const opsg_bf16* vt, int ld_vt, const uint32_t* bits,
                            int words, const int32_t* pair_index, int num_objects, int B, int n_query, int L, int num_heads,
                            int head_dim, opsg_bf16* ctx_out, cudaStream_t stream);

constexpr int kXaThreads = 512;
constexpr int kXaKeys = 256;       // max keys (one N=256 MMA)
constexpr int kXaHd = 64;
constexpr int kXaPvN = 80;         // PV MMA N: 64 value dims + 1 ones row (row sum) + 15 zero rows
constexpr int kXaSlots = 8;        // tile-local pair slots carried by the mask k-step

// Per-unit clock stamps (scripts/xattn_trace.py) are compiled in only with -DOPSG_TRACE (make trace): a runtime-null trace
// pointer still costs a parameter load and a branch inside the elected MMA-issue blocks and the softmax loop.
#ifdef OPSG_TRACE
#define XA_TRACE (p.trace)
#else
#define XA_TRACE (static_cast<long long*>(nullptr))
#endif

struct XattnParams {
  const uint8_t* tiles;      // [m_tiles][kXaTileBytes] A_aug | B_aug in core-matrix order (xattn_bias_tiles_kernel)
  const uint8_t* row_flags;  // [m_tiles * 128] 1 = the row's pair has an empty union mask (uniform attention)
  const uint16_t* chunk_vis; // [m_tiles * 4] per 32-row quarter: bit c = some row of the quarter sees a key of 16-key chunk c
  int L, num_heads, d_model;
  int rows;          // B * n_query
  int m_tiles;
  int total_units;
  int ctas_per_head; // > 0: grid = num_heads x ctas_per_head, every CTA walks m-tiles of a single head
  long long* trace;  // debug: per-unit clock64 stamps of CTA 0 ([unit][8]); NULL in production
  float scale_log2e;
};

constexpr int kXaAugA = 128 * 32;                       // 4096 : A_aug [128 rows x 16 slots] bf16
constexpr int kXaAugB = kXaKeys * 32;                   // 8192 : B_aug [256 keys x 16 slots] bf16
constexpr int kXaTileBytes = kXaAugA + kXaAugB;         // 12288 per 128-row tile

// K-major operand WITHOUT swizzle: 8-row x 16-byte core matrices stored contiguously (128 B); lbo = byte distance
// between the two core matrices of a K = 16 step, sbo = byte distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_k_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

struct XaSmem {
  static constexpr int kK = kXaKeys * 128;              // 32768: K [256 keys x 64] SW128 K-major
  static constexpr int kVBlk = kXaPvN * 128;            // 10240: one 64-key block of [V^T | 1 | 0] (80 rows x 128 B)
  static constexpr int kVBox = 64 * 128;                // 8192 : bytes TMA writes into a block (rows 0-63)
  static constexpr int kV = 4 * kVBlk;                  // 40960
  static constexpr int kQ = 128 * 128;                  // 16384 per stage
  static constexpr int kOst = 128 * 128;                // 16384 per TMEM buffer: O staging for the TMA store
  static constexpr int kAug = 12288;                    // per stage: A_aug (4096) | B_aug (8192), core-matrix order
  static constexpr int kMax = 2 * 2 * 128 * 4;          // 2048 : partial row maxima [buffer][half][row]
  static constexpr int kVbar = 2 * 64 * 4;              // 512  : mean of V per head parity
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kOffK + kK;
  static constexpr int kOffQ = kOffV + kV;
  static constexpr int kOffO = kOffQ + 2 * kQ;
  static constexpr int kOffAug = kOffO + 2 * kOst;
  static constexpr int kOffMax = kOffAug + 2 * kAug;
  static constexpr int kOffVbar = kOffMax + kMax;
  static constexpr int kOffBar = kOffVbar + kVbar;
  static constexpr int kTotal = kOffBar + 256 + 1024;
};
static_assert(XaSmem::kOffV % 1024 == 0 && XaSmem::kOffQ % 1024 == 0 && XaSmem::kOffO % 1024 == 0 &&
              XaSmem::kOffAug % 1024 == 0, "swizzled tiles need 1024-byte alignment");
static_assert(XaSmem::kTotal <= 232448, "shared memory budget exceeded");

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp2 on the FMA / ALU pipes for x <= 0 (FA-4 style MUFU offload): round-to-nearest split x = n + f by the
// 1.5*2^23 magic constant, degree-3 minimax 2^f on [-0.5, 0.5] (max relative error 7.6e-5, 50x below the bf16
// rounding of P), exponent added through the integer bits of t.  x is clamped to -125 so the exponent field cannot
// wrap: a masked score (bias -16384) gives 2^-125 instead of +0, which no fp32 accumulation can see.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;
  const float n = t - 12582912.f;
  const float f = x - n;
  float p = fmaf(0.05518026649951935f, f, 0.24261191487312317f);
  p = fmaf(p, f, 0.6932594180107117f);
  p = fmaf(p, f, 0.9999279975891113f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------------
// Mask-bias operand tiles, once per image (shared by all heads and both layers).
// Tile t (128 stacked query rows) -> 12288 bytes in the no-swizzle K-major core-matrix order the MMA reads:
//   A_aug: 16 row groups x [k-core 0: 8 rows x 16 B (slots 0-7) | k-core 1: 128 B of zeros]
//   B_aug: 32 key groups x [k-core 0: 8 keys x 16 B (slots 0-7) | k-core 1: zeros]
// followed (after all tiles) by one flag byte per row: 1 = the row's pair mask is empty.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
xattn_bias_tiles_kernel(const uint32_t* __restrict__ bits, int words, const int32_t* __restrict__ pair_index,
                        int num_objects, int num_pairs, int n_query, int L, int rows, uint8_t* __restrict__ tiles,
                        uint8_t* __restrict__ row_flags, uint16_t* __restrict__ chunk_vis) {
  pdl_wait_then_trigger();
  const int mt = blockIdx.x;
  const int t = threadIdx.x;
  const int first_pair = (mt * 128) / n_query;
  const int last_pair = min(num_pairs - 1, (mt * 128 + 127) / n_query);
  const int nslots = last_pair - first_pair + 1;              // <= kXaSlots (host-checked)
  __shared__ uint32_t s_union[kXaSlots][8];                   // union mask words per slot, real keys only
  __shared__ uint32_t s_any[kXaSlots];
  if (t < kXaSlots * 8) {
    const int slot = t >> 3, w = t & 7;
    uint32_t u = 0;
    if (slot < nslots) {
      const int pair = first_pair + slot;
      const int pidx = max(0, pair_index ? pair_index[pair] : pair);
      const int oi = min(pidx / num_objects, num_objects - 1), oj = pidx % num_objects;
      const uint32_t keyok = (w < (L >> 5)) ? 0xffffffffu : (w == (L >> 5) ? ((1u << (L & 31)) - 1u) : 0u);
      if (w < words) u = (bits[static_cast<size_t>(oi) * words + w] | bits[static_cast<size_t>(oj) * words + w]) & keyok;
    }
    s_union[slot][w] = u;
  }
  __syncthreads();
  if (t < kXaSlots) {
    uint32_t any = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) any |= s_union[t][w];
    s_any[t] = any;
  }
  __syncthreads();
  if (t < 4) {   // 16-key chunks in which any row of 32-row quarter t sees at least one key
    uint32_t vis = 0;
    for (int r = t * 32; r < t * 32 + 32; ++r) {
      const int row = mt * 128 + r;
      if (row >= rows) break;
      const int slot = row / n_query - first_pair;
#pragma unroll
      for (int w = 0; w < 8; ++w)
        vis |= ((s_union[slot][w] & 0xFFFFu) ? (1u << (2 * w)) : 0u) | ((s_union[slot][w] >> 16) ? (2u << (2 * w)) : 0u);
    }
    chunk_vis[mt * 4 + t] = static_cast<uint16_t>(vis);
  }
  uint8_t* tile = tiles + static_cast<size_t>(mt) * kXaTileBytes;
  {  // B_aug: thread = key
    const int key = t;
    uint32_t pk[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int s = 0; s < kXaSlots; ++s) {
      const bool allowed = (s_union[s][key >> 5] >> (key & 31)) & 1u;     // keys >= L were cleared by keyok
      if (s < nslots && !allowed) pk[s >> 1] |= (s & 1) ? 0xC6800000u : 0x0000C680u;     // bf16(-16384)
    }
    uint8_t* dst = tile + kXaAugA + (key >> 3) * 256 + (key & 7) * 16;
    *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(dst + 128) = make_uint4(0, 0, 0, 0);
  }
  if (t < 128) {  // A_aug: thread = row
    const int row = mt * 128 + t;
    uint32_t pk[4] = {0u, 0u, 0u, 0u};
    uint8_t flag = 0;
    if (row < rows) {
      const int slot = row / n_query - first_pair;
#pragma unroll
      for (int s = 0; s < kXaSlots; ++s)
        if (s == slot) pk[s >> 1] = (s & 1) ? 0x3F800000u : 0x00003F80u;                 // bf16(1.0)
      flag = s_any[slot] == 0 ? 1 : 0;
    }
    uint8_t* dst = tile + (t >> 3) * 256 + (t & 7) * 16;
    *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(dst + 128) = make_uint4(0, 0, 0, 0);
    row_flags[row] = flag;
  }
}


// ------------------------------------------------------------------------------------------------
// Key order for K5: image tokens sorted by owning object (owner = first listed object whose mask holds the token; tokens
// nobody owns go last), stable in the token index.  Attention is invariant under a permutation of the keys, and with the
// keys of an object contiguous a pair's visible keys are two short runs, so most 16-key chunks of a 32-row quarter are
// invisible and skip the softmax work entirely (the panoptic map partitions the image: an object owns ~L/N tokens).
// One CTA of 256 threads (L <= 256): perm[new] = old token index; bits_sorted = the object masks in the new order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
token_order_kernel(const uint32_t* __restrict__ bits, int words, int num_objects, int L, int32_t* __restrict__ perm,
                   uint32_t* __restrict__ bits_sorted) {
  pdl_wait_then_trigger();
  __shared__ int s_owner[kXaKeys];
  __shared__ int s_perm[kXaKeys];
  const int t = threadIdx.x;
  int owner = 0x7fffffff;
  if (t < L) {
    owner = num_objects;
    for (int o = 0; o < num_objects; ++o)
      if ((bits[static_cast<size_t>(o) * words + (t >> 5)] >> (t & 31)) & 1u) { owner = o; break; }
  }
  s_owner[t] = owner;
  __syncthreads();
  if (t < L) {
    int rank = 0;
    for (int u = 0; u < L; ++u) {
      const int ou = s_owner[u];
      rank += (ou < owner || (ou == owner && u < t)) ? 1 : 0;
    }
    s_perm[rank] = t;
    perm[rank] = t;
  }
  __syncthreads();
  const int src = t < L ? s_perm[t] : 0;
  const int lane = t & 31, w = t >> 5;
  for (int o = 0; o < num_objects; ++o) {
    const bool on = t < L && ((bits[static_cast<size_t>(o) * words + (src >> 5)] >> (src & 31)) & 1u);
    const uint32_t word = __ballot_sync(0xffffffffu, on);
    if (lane == 0 && w < words) bits_sorted[static_cast<size_t>(o) * words + w] = word;
  }
}

// POLY_MOD: 0 = every exponential on the MUFU; m > 0 = elements with (index % m == 1) use ex2_poly (FMA / ALU pipes).
//
// v4 warp roles (512 threads, <= 128 registers each):
//   warp 0      TMA producer
//   warp 1      QK^T issuer   (its own in-order queue: never blocked behind a PV that waits for P)
//   warp 2      TMEM allocator, then PV issuer (waits for P in two 128-key halves: the first 8 PV MMAs run under the
//               exponentials of the second half)
//   warp 3      per-head mean of V
//   warps 4-7   softmax of TMEM buffer 0, warps 8-11 of buffer 1: thread = one score row x all 256 keys (no cross-thread
//               max exchange); P overwrites the consumed scores in place, ascending, columns [0, 128)
//   warps 12-15 epilogue of BOTH buffers: O / rowsum out of TMEM columns [128, 209) (frees the buffer for the QK^T of
//               unit i+2 as soon as it has been read), uniform rows, bf16, swizzled staging, TMA store
// so a softmax warp goes straight from the exponentials of unit i to the scores of unit i+2.
template <int POLY_MOD>
__global__ void __launch_bounds__(kXaThreads, 1)
xattn_pairs_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ CUtensorMap tmO,
                   const XattnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem + XaSmem::kOffK;
  uint8_t* sV = smem + XaSmem::kOffV;
  uint8_t* sQ = smem + XaSmem::kOffQ;
  uint8_t* sO = smem + XaSmem::kOffO;
  uint8_t* sAug = smem + XaSmem::kOffAug;                                    // [stage] A_aug | B_aug
  float* sVbar = reinterpret_cast<float*>(smem + XaSmem::kOffVbar);          // [head parity][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + XaSmem::kOffBar);
  uint64_t* q_full = bars;          // [2]  Q tile + mask-bias tiles of a unit
  uint64_t* q_empty = bars + 2;     // [2]
  uint64_t* s_full = bars + 4;      // [2]
  uint64_t* p_ready = bars + 6;     // [2]
  uint64_t* o_full = bars + 10;     // [2]
  uint64_t* s_free = bars + 12;     // [2]
  uint64_t* vbar_ready = bars + 14; // [2]
  uint64_t* k_full = bars + 16;
  uint64_t* k_empty = bars + 17;
  uint64_t* v_full = bars + 18;
  uint64_t* v_empty = bars + 19;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  auto gtime = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return static_cast<long long>(t); };
  if (XA_TRACE && threadIdx.x == 0) XA_TRACE[2048 + 4 * blockIdx.x + 0] = gtime();      // debug: per-CTA wall-clock stamps (ns)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmVt);
    tma_prefetch_desc(&tmO);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&q_full[b], 1);
      mbar_init(&q_empty[b], 1);
      mbar_init(&s_full[b], 1);
      mbar_init(&p_ready[b], 128);
      mbar_init(&o_full[b], 1);
      mbar_init(&s_free[b], 128);
      mbar_init(&vbar_ready[b], 1);
    }
    mbar_init(k_full, 1); mbar_init(k_empty, 1);
    mbar_init(v_full, 1); mbar_init(v_empty, 2);     // v_empty: last PV of the head (tcgen05.commit) + the V-mean warp
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // rows 64..79 of every [V^T | 1 | 0] block: row 64 = ones (the PV MMA then also produces the row sum), rest zero.
  // Row 64 has (row & 7) == 0 and all its 16-byte chunks are equal, so the 128B swizzle does not matter here.
  for (int idx = threadIdx.x; idx < 4 * 128; idx += kXaThreads) {
    const int blk = idx >> 7, within = idx & 127;                // 4 blocks, 16 rows x 8 chunks each
    const uint32_t one2 = (within < 8) ? 0x3F803F80u : 0u;        // bf16 1.0 pairs in row 64
    *reinterpret_cast<uint4*>(sV + blk * XaSmem::kVBlk + XaSmem::kVBox + within * 16) = make_uint4(one2, one2, one2, one2);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait_then_trigger();          // everything above overlaps the previous kernel (programmatic dependent launch)

  // contiguous, head-major unit range of this CTA
  // Unit range of this CTA.  With ctas_per_head > 0 the grid is num_heads x ctas_per_head and a CTA stays inside ONE head
  // (its K / V tiles are loaded once, no reload bubble); otherwise a contiguous head-major range.
  int u_begin, u_end;
  if (p.ctas_per_head > 0) {
    const int head = blockIdx.x / p.ctas_per_head, c = blockIdx.x % p.ctas_per_head;
    u_begin = head * p.m_tiles + static_cast<int>(static_cast<long long>(c) * p.m_tiles / p.ctas_per_head);
    u_end = head * p.m_tiles + static_cast<int>(static_cast<long long>(c + 1) * p.m_tiles / p.ctas_per_head);
  } else {
    const int per = (p.total_units + gridDim.x - 1) / gridDim.x;
    u_begin = blockIdx.x * per;
    u_end = min(p.total_units, u_begin + per);
  }
  const int n_units = max(0, u_end - u_begin);
  const int head0 = u_begin / p.m_tiles;
  auto head_of = [&](int j) { return (u_begin + j) / p.m_tiles; };

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, elect.sync lane issues) =====================
    int cur_head = -1, loads = 0;
    for (int i = 0; i < n_units; ++i) {
      const int u = u_begin + i;
      const int head = u / p.m_tiles, mt = u % p.m_tiles;
      const bool new_head = head != cur_head;
      if (new_head) {      // K first: QK^T of this unit only needs K (V may still be in use by the previous unit's PV)
        if (loads > 0) mbar_wait(k_empty, (loads - 1) & 1);
        if (elect_one_sync()) {
          mbar_expect_tx(k_full, XaSmem::kK);
          tma_load_2d(sK, &tmK, k_full, head * kXaHd, 0);
        }
        __syncwarp();
      }
      const int b = i & 1;
      mbar_wait(&q_empty[b], ((i >> 1) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_expect_tx(&q_full[b], XaSmem::kQ + kXaTileBytes);
        tma_load_2d(sQ + b * XaSmem::kQ, &tmQ, &q_full[b], head * kXaHd, mt * 128);
        bulk_load_1d(sAug + b * XaSmem::kAug, p.tiles + static_cast<size_t>(mt) * kXaTileBytes, kXaTileBytes, &q_full[b]);
      }
      __syncwarp();
      if (new_head) {
        if (loads > 0) mbar_wait(v_empty, (loads - 1) & 1);
        if (elect_one_sync()) {
          mbar_expect_tx(v_full, 4 * XaSmem::kVBox);
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) tma_load_2d(sV + kb * XaSmem::kVBlk, &tmVt, v_full, kb * 64, head * kXaHd);
        }
        __syncwarp();
        ++loads;
        cur_head = head;
      }
    }
  } else if (warp == 1) {
    // ===================== QK^T issuer =====================
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, kXaKeys);
    const uint32_t lbo = 128u, sbo = 256u;      // no-swizzle core matrices: 128 B between the k-cores, 256 B between 8-row groups
    int k_waits = 0, cur_head = -1;
    for (int j = 0; j < n_units; ++j) {
      const int head = head_of(j);
      const int b = j & 1;
      if (head != cur_head) {
        mbar_wait(k_full, k_waits & 1);
        ++k_waits;
        cur_head = head;
      }
      mbar_wait(&q_full[b], (j >> 1) & 1);
      mbar_wait(&s_free[b], ((j >> 1) & 1) ^ 1);     // P of unit j-2 consumed by its PV and O read out of this buffer
      tc_fence_after();
      const bool release_k = j + 1 < n_units && head_of(j + 1) != head;
      if (elect_one_sync()) {
        if (XA_TRACE && blockIdx.x == 0) XA_TRACE[j * 8 + 0] = clock64();
        if (XA_TRACE && j == 0) XA_TRACE[2048 + 4 * blockIdx.x + 2] = gtime();
        const uint64_t a_desc = umma_desc_k_sw128(smem_u32(sQ + b * XaSmem::kQ));
        const uint64_t k_desc = umma_desc_k_sw128(smem_u32(sK));
        const uint32_t aug = smem_u32(sAug + b * XaSmem::kAug);
        const uint32_t d = tmem_base + b * 256;
#pragma unroll
        for (int k = 0; k < kXaHd / 16; ++k)           // +32 bytes per K step = +2 in the 16-byte address field
          umma_ss(d, a_desc + 2 * k, k_desc + 2 * k, idesc_qk, k > 0 ? 1u : 0u);
        umma_ss(d, umma_desc_k_noswz(aug, lbo, sbo), umma_desc_k_noswz(aug + kXaAugA, lbo, sbo), idesc_qk, 1u);  // + mask bias
        tc_commit(&q_empty[b]);
        tc_commit(&s_full[b]);
        if (release_k) tc_commit(k_empty);
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ===================== PV issuer =====================
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, kXaPvN);
    int v_waits = 0, cur_head = -1;
    for (int i = 0; i < n_units; ++i) {
      const int head = head_of(i);
      const int b = i & 1;
      const uint32_t parity = (i >> 1) & 1;
      if (head != cur_head) {
        mbar_wait(v_full, v_waits & 1);
        ++v_waits;
        cur_head = head;
      }
      const bool release_v = i + 1 < n_units && head_of(i + 1) != head;
      const uint32_t pa = tmem_base + b * 256;          // P: key k-step k in columns [8k, 8k + 8)
      const uint32_t od = tmem_base + b * 256 + 128;    // O: 80 fp32 columns [128, 208)
      const uint64_t v_desc = umma_desc_k_sw128(smem_u32(sV));
      mbar_wait(&p_ready[b], parity);                  // P complete in TMEM, S fully consumed
      tc_fence_after();
      if (elect_one_sync()) {
        if (XA_TRACE && blockIdx.x == 0) XA_TRACE[i * 8 + 1] = clock64();
#pragma unroll
        for (int k = 0; k < kXaKeys / 16; ++k)
          umma_ts(od, pa + k * 8, v_desc + (k >> 2) * (XaSmem::kVBlk >> 4) + (k & 3) * 2, idesc_pv, k > 0 ? 1u : 0u);
        tc_commit(&o_full[b]);
        if (release_v) tc_commit(v_empty);
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // ===================== per-head mean of V over the L real keys (rows whose pair mask is empty) =====================
    // This warp READS the resident V tile, so it is a second consumer of it: the producer may only overwrite V with the
    // next head's tile after the head's last PV MMA *and* this warp have released it (v_empty counts both).
    int v_waits = 0, cur_head = -1;
    for (int i = 0; i < n_units; ++i) {
      const int head = head_of(i);
      if (head == cur_head) continue;
      cur_head = head;
      const int hs = head - head0;
      mbar_wait(v_full, v_waits & 1);
      ++v_waits;
#pragma unroll 1
      for (int rr = 0; rr < 2; ++rr) {
        const int n = lane + rr * 32;                        // value dim; the swizzle only permutes chunks inside a row
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int kb = 0; kb < 4; ++kb) {
          const uint4* rowp = reinterpret_cast<const uint4*>(sV + kb * XaSmem::kVBlk + n * 128);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 x = rowp[c];
            a0 += bf16_lo(x.x) + bf16_hi(x.x);
            a1 += bf16_lo(x.y) + bf16_hi(x.y);
            a2 += bf16_lo(x.z) + bf16_hi(x.z);
            a3 += bf16_lo(x.w) + bf16_hi(x.w);
          }
        }
        sVbar[(hs & 1) * 64 + n] = ((a0 + a1) + (a2 + a3)) / static_cast<float>(p.L);     // keys >= L are zero-filled by TMA
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&vbar_ready[hs & 1]);
        const int last_head = head_of(n_units - 1);
        if (head != last_head) mbar_arrive(v_empty);
      }
    }
  } else if (warp < 12) {
    // ===================== softmax warps: thread = one row x 256 keys =====================
    const int b = (warp - 4) >> 2;                       // TMEM buffer = unit parity
    const int q = warp & 3;                              // TMEM lane quarter of this warp (warp id % 4)
    const uint32_t tS = tmem_base + b * 256 + (static_cast<uint32_t>(q * 32) << 16);
    const bool tr_thread = blockIdx.x == 0 && q == 0 && lane == 0;
    auto exp2_sel = [&](float x, int idx) -> float {
      if (POLY_MOD > 0 && (idx % (POLY_MOD > 0 ? POLY_MOD : 1)) == 1) return ex2_poly(x);
      return ex2_approx(x);
    };
    for (int i = b; i < n_units; i += 2) {
      const uint32_t parity = (i >> 1) & 1;
      const int mt = (u_begin + i) % p.m_tiles;
      // 16-key chunks in which no row of this warp sees a key contribute exact zeros: no tcgen05.ld, no max, no
      // exponentials -- their P columns are just cleared.  The MMAs stay dense.  (The image tokens reach this kernel
      // sorted by owning object, opsg_token_order, so a pair's visible keys are two short contiguous runs.)
      const uint32_t vis = __ldg(p.chunk_vis + mt * 4 + q);
      mbar_wait(&s_full[b], parity);
      tc_fence_after();
      const bool tr = XA_TRACE && tr_thread;
      if (tr) XA_TRACE[i * 8 + 2] = clock64();

      const uint32_t zero8[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      if (__popc(vis) <= 4) {
        // ---- sparse rows (the common case with keys in object order): the <= 4 visible chunks are read ONCE into
        // registers (one TMEM round trip), max and exponentials run from registers, every invisible P chunk is cleared
        // with the widest store that does not touch a visible one ----
        uint32_t r0[16], r1[16], r2[16], r3[16];
        uint32_t rem = vis;
        const int c0 = rem ? __ffs(rem) - 1 : -1; rem &= rem - 1;
        const int c1 = rem ? __ffs(rem) - 1 : -1; rem &= rem - 1;
        const int c2 = rem ? __ffs(rem) - 1 : -1; rem &= rem - 1;
        const int c3 = rem ? __ffs(rem) - 1 : -1;
        if (c0 >= 0) tmem_ld16(tS + c0 * 16, r0);
        if (c1 >= 0) tmem_ld16(tS + c1 * 16, r1);
        if (c2 >= 0) tmem_ld16(tS + c2 * 16, r2);
        if (c3 >= 0) tmem_ld16(tS + c3 * 16, r3);
        tmem_ld_wait();
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        auto max16 = [&](const uint32_t (&v)[16]) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            m0 = fmaxf(m0, __uint_as_float(v[j]));
            m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
            m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
            m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
          }
        };
        if (c0 >= 0) max16(r0);
        if (c1 >= 0) max16(r1);
        if (c2 >= 0) max16(r2);
        if (c3 >= 0) max16(r3);
        float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        if (mx == -INFINITY) mx = 0.f;                   // no visible key: uniform / out-of-range rows only
        const float mxs = mx * p.scale_log2e;
        if (tr) XA_TRACE[i * 8 + 3] = clock64();
        // every score this thread needs is in registers: the P columns may be written in any order
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t v4 = (vis >> (4 * g)) & 0xFu;
          if (v4 == 0u) {
            uint32_t z[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) z[j] = 0u;
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
                "%25, %26, %27, %28, %29, %30, %31, %32};"
                ::"r"(tS + g * 32), "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]),
                "r"(z[8]), "r"(z[9]), "r"(z[10]), "r"(z[11]), "r"(z[12]), "r"(z[13]), "r"(z[14]), "r"(z[15]), "r"(z[16]),
                "r"(z[17]), "r"(z[18]), "r"(z[19]), "r"(z[20]), "r"(z[21]), "r"(z[22]), "r"(z[23]), "r"(z[24]), "r"(z[25]),
                "r"(z[26]), "r"(z[27]), "r"(z[28]), "r"(z[29]), "r"(z[30]), "r"(z[31])
                : "memory");
          } else {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
              if (!((v4 >> cc) & 1u)) tmem_st8(tS + (4 * g + cc) * 8, zero8);
          }
        }
        auto exp_store_r = [&](const uint32_t (&v)[16], int c) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float e0 = exp2_sel(fmaf(__uint_as_float(v[2 * j]), p.scale_log2e, -mxs), 2 * j);
            const float e1 = exp2_sel(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2e, -mxs), 2 * j + 1);
            pk[j] = pack_bf16x2(e0, e1);
          }
          tmem_st8(tS + c * 8, pk);
        };
        if (c0 >= 0) exp_store_r(r0, c0);
        if (c1 >= 0) exp_store_r(r1, c1);
        if (c2 >= 0) exp_store_r(r2, c2);
        if (c3 >= 0) exp_store_r(r3, c3);
      } else {
      // ---- pass 1: row max of the biased scores over the visible chunks (next chunk's tcgen05.ld in flight) ----
      uint32_t va[16], vb[16];
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
      auto max16 = [&](const uint32_t (&v)[16]) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          m0 = fmaxf(m0, __uint_as_float(v[j]));
          m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
          m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
          m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
        }
      };
      {
        uint32_t rem = vis;
        int c = rem ? __ffs(rem) - 1 : -1;
        rem &= rem - 1;
        if (c >= 0) tmem_ld16(tS + c * 16, va);
        while (c >= 0) {
          tmem_ld_wait();
          int cn = rem ? __ffs(rem) - 1 : -1;
          rem &= rem - 1;
          if (cn >= 0) tmem_ld16(tS + cn * 16, vb);
          max16(va);
          c = cn;
          if (c < 0) break;
          tmem_ld_wait();
          cn = rem ? __ffs(rem) - 1 : -1;
          rem &= rem - 1;
          if (cn >= 0) tmem_ld16(tS + cn * 16, va);
          max16(vb);
          c = cn;
        }
      }
      float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      if (mx == -INFINITY) mx = 0.f;                     // no visible key: uniform / out-of-range rows only
      const float mxs = mx * p.scale_log2e;
      if (tr) XA_TRACE[i * 8 + 3] = clock64();

      // ---- pass 2: p = exp2(s*scale - max*scale), packed bf16 P in place over consumed score columns ----
      auto exp_store = [&](const uint32_t (&v)[16], int c) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float e0 = exp2_sel(fmaf(__uint_as_float(v[2 * j]), p.scale_log2e, -mxs), 2 * j);
          const float e1 = exp2_sel(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2e, -mxs), 2 * j + 1);
          pk[j] = pack_bf16x2(e0, e1);
        }
        tmem_st8(tS + c * 8, pk);
      };
      // Chunks are walked in ascending order: P chunk c lands on columns [8c, 8c + 8), inside score chunk c / 2 <= c,
      // which is either already in registers or invisible; the prefetched chunk lies above.
      {
        uint32_t rem = vis;
        int nxt = rem ? __ffs(rem) - 1 : 16;             // next visible chunk (16 = none)
        rem &= rem - 1;
        if (nxt < 16) tmem_ld16(tS + nxt * 16, va);
        bool use_a = true;
#pragma unroll 1
        for (int c = 0; c < 16; ++c) {
          if (c == nxt) {
            tmem_ld_wait();
            nxt = rem ? __ffs(rem) - 1 : 16;
            rem &= rem - 1;
            if (use_a) {
              if (nxt < 16) tmem_ld16(tS + nxt * 16, vb);
              exp_store(va, c);
            } else {
              if (nxt < 16) tmem_ld16(tS + nxt * 16, va);
              exp_store(vb, c);
            }
            use_a = !use_a;
          } else {
            tmem_st8(tS + c * 8, zero8);
          }
        }
      }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[b]);
      if (tr) XA_TRACE[i * 8 + 4] = clock64();
    }
  } else {
    // ===================== epilogue warps (both buffers): O / rowsum -> bf16 -> swizzled staging -> TMA store =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const bool elected = warp == 12 && lane == 0;
    for (int i = 0; i < n_units; ++i) {
      const int b = i & 1;
      const uint32_t parity = (i >> 1) & 1;
      const int u = u_begin + i;
      const int head = u / p.m_tiles, mt = u % p.m_tiles;
      const bool uniform = __ldg(p.row_flags + mt * 128 + r) != 0;     // empty pair mask -> uniform attention (HF finfo.min)
      const uint32_t tO = tmem_base + b * 256 + 128 + (static_cast<uint32_t>(q * 32) << 16);
      mbar_wait(&o_full[b], parity);
      tc_fence_after();
      if (XA_TRACE && blockIdx.x == 0 && elected) XA_TRACE[i * 8 + 5] = clock64();
      uint32_t o0[32], o1[32];
      tmem_ld32(tO, o0);
      tmem_ld32(tO + 32, o1);
      const uint32_t osum = tmem_ld1(tO + 64);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[b]);                           // buffer may be overwritten by QK^T of unit i+2
      const float inv = 1.f / __uint_as_float(osum);
      uint8_t* stage_row = sO + b * XaSmem::kOst + r * 128;
      if (elected) tma_store_wait_read<1>();             // the store of unit i-2 has drained this staging tile
      named_bar_sync(1, 128);
      if (uniform) {
        const int hs = head - head0;
        mbar_wait(&vbar_ready[hs & 1], (hs >> 1) & 1);
        const float* vb = sVbar + (hs & 1) * 64;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          uint4 u4;
          u4.x = pack_bf16x2(vb[g * 8 + 0], vb[g * 8 + 1]);
          u4.y = pack_bf16x2(vb[g * 8 + 2], vb[g * 8 + 3]);
          u4.z = pack_bf16x2(vb[g * 8 + 4], vb[g * 8 + 5]);
          u4.w = pack_bf16x2(vb[g * 8 + 6], vb[g * 8 + 7]);
          *reinterpret_cast<uint4*>(stage_row + ((g ^ (r & 7)) * 16)) = u4;
        }
      } else {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint32_t (&o)[32] = g < 4 ? o0 : o1;
          const int j = (g & 3) * 8;
          uint4 u4;
          u4.x = pack_bf16x2(__uint_as_float(o[j + 0]) * inv, __uint_as_float(o[j + 1]) * inv);
          u4.y = pack_bf16x2(__uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
          u4.z = pack_bf16x2(__uint_as_float(o[j + 4]) * inv, __uint_as_float(o[j + 5]) * inv);
          u4.w = pack_bf16x2(__uint_as_float(o[j + 6]) * inv, __uint_as_float(o[j + 7]) * inv);
          *reinterpret_cast<uint4*>(stage_row + ((g ^ (r & 7)) * 16)) = u4;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(2, 128);
      if (elected) {
        tma_store_2d(sO + b * XaSmem::kOst, &tmO, head * kXaHd, mt * 128);
        tma_store_commit();
        if (XA_TRACE && blockIdx.x == 0) XA_TRACE[i * 8 + 6] = clock64();
      }
    }
    if (elected) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (XA_TRACE && threadIdx.x == 0) XA_TRACE[2048 + 4 * blockIdx.x + 3] = gtime();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace opsg

using namespace opsg;

static long long* g_xattn_trace = nullptr;
// debug hook (not part of the public header): device buffer of >= 8 * units_per_cta int64 for per-unit clock stamps
extern "C" void opsg_debug_xattn_trace(void* dev_buffer) { g_xattn_trace = reinterpret_cast<long long*>(dev_buffer); }

extern "C" int opsg_token_order(const uint32_t* bits, int words, int num_objects, int L, int32_t* perm_out,
                                uint32_t* bits_sorted_out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(bits && perm_out && bits_sorted_out, "token_order: null pointer");
  OPSG_CHECK_ARG(num_objects > 0 && L > 0 && words >= (L + 31) / 32 && words <= 8, "token_order: bad shape");
  if (L > kXaKeys) return set_error(OPSG_E_UNSUPPORTED, "token_order: L=%d image tokens > %d unsupported", L, kXaKeys);
  launch_kernel(token_order_kernel, 1, 256, 0, reinterpret_cast<cudaStream_t>(stream), bits, words, num_objects, L, perm_out,
                bits_sorted_out);
  OPSG_CHECK_LAUNCH("token_order_kernel");
  return OPSG_OK;
}

extern "C" size_t opsg_xattn_bias_tiles_bytes(int B, int n_query) {
  if (B <= 0 || n_query <= 0) return 0;
  const size_t m_tiles = (static_cast<size_t>(B) * n_query + 127) / 128;
  return m_tiles * (kXaTileBytes + 128 + 16);       // tiles | row flags | chunk visibility (4 x uint16 per tile, padded)
}

extern "C" int opsg_xattn_bias_tiles(const uint32_t* bits, int words, const int32_t* pair_index, int num_objects, int B,
                                     int n_query, int L, void* tiles_out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(bits && tiles_out, "xattn_bias_tiles: null pointer");
  OPSG_CHECK_ARG(B > 0 && n_query > 0 && L > 0 && num_objects > 0, "xattn_bias_tiles: bad shape");
  if (L > kXaKeys) return set_error(OPSG_E_UNSUPPORTED, "xattn_bias_tiles: L=%d image tokens > %d unsupported", L, kXaKeys);
  OPSG_CHECK_ARG(words >= (L + 31) / 32 && words <= 8, "xattn_bias_tiles: words=%d inconsistent with L=%d", words, L);
  if (127 / n_query + 2 > kXaSlots)
    return set_error(OPSG_E_UNSUPPORTED, "xattn_bias_tiles: n_query=%d puts more than %d pairs in a 128-row tile", n_query, kXaSlots);
  OPSG_CHECK_ARG(((uintptr_t)tiles_out & 15) == 0, "xattn_bias_tiles: output must be 16-byte aligned");
  const int rows = B * n_query;
  const int m_tiles = (rows + 127) / 128;
  uint8_t* tiles = reinterpret_cast<uint8_t*>(tiles_out);
  launch_kernel(xattn_bias_tiles_kernel, m_tiles, 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      bits, words, pair_index, num_objects, B, n_query, L, rows, tiles, tiles + static_cast<size_t>(m_tiles) * kXaTileBytes,
      reinterpret_cast<uint16_t*>(tiles + static_cast<size_t>(m_tiles) * (kXaTileBytes + 128)));
  OPSG_CHECK_LAUNCH("xattn_bias_tiles_kernel");
  return OPSG_OK;
}

extern "C" int opsg_xattn_pairs(const opsg_bf16* q, const opsg_bf16* k, int ld_k, const opsg_bf16* vt, int ld_vt,
                                const uint32_t* bits, int words, const int32_t* pair_index, int num_objects, int B,
                                int n_query, int L, int num_heads, int head_dim, const void* bias_tiles,
                                opsg_bf16* ctx_out, void* stream) {
  // more image tokens than one 256-key score tile: the online-softmax kernel (xattn_pairs_long.cu)
  if (L > kXaKeys)
    return launch_xattn_pairs_long(q, k, ld_k, vt, ld_vt, bits, words, pair_index, num_objects, B, n_query, L, num_heads, head_dim,
                                   ctx_out, reinterpret_cast<cudaStream_t>(stream));
  // without precomputed mask-bias tiles (or for shapes they do not cover) the self-contained v2 kernel runs
  if (!bias_tiles || n_query <= 0 || 127 / n_query + 2 > kXaSlots)
    return opsg_xattn_pairs_v2(q, k, ld_k, vt, ld_vt, bits, words, pair_index, num_objects, B, n_query, L, num_heads,
                               head_dim, ctx_out, stream);
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(q && k && vt && ctx_out, "xattn_pairs: null pointer");
  OPSG_CHECK_ARG(B > 0 && n_query > 0 && L > 0 && num_heads > 0 && num_objects > 0, "xattn_pairs: bad shape");
  if (head_dim != kXaHd) return set_error(OPSG_E_UNSUPPORTED, "xattn_pairs: head_dim %d unsupported (64 only)", head_dim);
  if (L > kXaKeys) return set_error(OPSG_E_UNSUPPORTED, "xattn_pairs: L=%d image tokens > %d unsupported", L, kXaKeys);
  const int d_model = num_heads * head_dim;
  OPSG_CHECK_ARG(ld_k >= d_model && ld_k % 8 == 0 && ld_vt >= L && ld_vt % 8 == 0, "xattn_pairs: bad leading dims");
  OPSG_CHECK_ARG(((uintptr_t)ctx_out & 15) == 0 && ((uintptr_t)bias_tiles & 15) == 0, "xattn_pairs: ctx_out / bias_tiles must be 16-byte aligned");
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
  // every exponential on the MUFU (POLY_MOD = 0); moving every 4th / 3rd / 2nd one to an FMA-pipe polynomial measured slower
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const XattnParams);
  static const KernelFn kernels[] = {xattn_pairs_kernel<0>};
  const KernelFn kernel = kernels[0];
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[device_slot()];
  if (!configured) {
    for (KernelFn f : kernels) {
      rc = check_cuda(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, XaSmem::kTotal),
                      "cudaFuncSetAttribute(xattn)");
      if (rc) return rc;
    }
    configured = true;
  }
  XattnParams p;
  p.m_tiles = (rows + 127) / 128;
  p.tiles = reinterpret_cast<const uint8_t*>(bias_tiles);
  p.row_flags = p.tiles + static_cast<size_t>(p.m_tiles) * kXaTileBytes;
  p.chunk_vis = reinterpret_cast<const uint16_t*>(p.tiles + static_cast<size_t>(p.m_tiles) * (kXaTileBytes + 128));
  p.L = L; p.num_heads = num_heads; p.d_model = d_model;
  p.rows = rows; p.total_units = p.m_tiles * num_heads;
  p.trace = g_xattn_trace;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(head_dim));
  const int sms = opsg_num_sms();
  int grid = p.total_units < sms ? p.total_units : sms;
  p.ctas_per_head = 0;
  if (num_heads <= sms && p.m_tiles >= 2 * (sms / num_heads)) {      // enough tiles per head: pin every CTA to one head
    p.ctas_per_head = sms / num_heads;
    grid = p.ctas_per_head * num_heads;
  }
  launch_kernel(kernel, grid, kXaThreads, XaSmem::kTotal, reinterpret_cast<cudaStream_t>(stream), tmQ, tmK, tmVt, tmO, p);
  OPSG_CHECK_LAUNCH("xattn_pairs_kernel");
  return OPSG_OK;
}
