// K5 for MORE THAN 256 image tokens — the envelope the tcgen05 kernels (xattn_pairs.cu, xattn_pairs_generic.cu) do not cover:
// they keep one 256-key score tile per unit in tensor memory.  The reference has no such limit (`prepare_inference`,
// relation_transformer_head_v4.py:408-435, tokenises whatever feature map it is given; the PSG / COCO images of its dataset give
// <= 240 tokens at 1333 x 800), so larger inputs take this kernel instead of an error: FlashAttention-2 style online softmax
// over 64-key blocks on warp-level mma.sync, masks applied from the bit words (bits[i] | bits[j], never materialised).
//
//   grid = (ceil(rows / 64), heads); 4 warps x 16 stacked query rows; K block [64 keys x 64] and V^T block [64 dims x 64 keys]
//   double-buffered in shared memory by cp.async; S = Q K^T (B fragments of K straight from ldmatrix), P stays in registers as
//   the A fragments of O += P V (B fragments of V from the V^T rows).  Same semantics as the other K5 kernels (HF
//   modeling_instructblip.py:499-536 with encoder_attention_mask = pair_masks, v4:168-170,183-184): masked keys weigh exactly 0,
//   P is rounded to bf16 before the PV product, a row whose pair mask is empty gets the fp32 mean of V over all L keys.
// Not a fast path: ~1/6 of the tensor-core kernels' rate; it exists so that large images run.
#include "common.cuh"
#include "host_util.h"

namespace opsg {
namespace xal {

constexpr int kHd = 64, kBlk = 64, kRows = 64, kThreads = 128;
constexpr int kPitch = (kHd + 8) * 2;                    // padded row pitch (bytes): ldmatrix phases hit distinct banks
constexpr int kTile = kBlk * kPitch;                     // one operand block (K: [key][dim]; V^T: [dim][key])

struct Params {
  const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* vt; __nv_bfloat16* out;
  const uint32_t* bits; const int32_t* pair_index;
  int ld_k, ld_vt, words, num_objects, n_query, L, d_model, rows;
  float scale_log2e;
};

__device__ __forceinline__ void cp16(void* dst, const void* src, bool valid) {
  const uint32_t sz = valid ? 16u : 0u;                  // src-size 0 -> 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__global__ void __launch_bounds__(kThreads) xattn_pairs_long_kernel(const Params p) {
  pdl_wait_then_trigger();
  __shared__ __align__(16) uint8_t sm[2][2][kTile];      // [stage][K | V^T]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, quad = lane & 3, grp = lane >> 2;
  const int head = blockIdx.y;
  const int row0 = blockIdx.x * kRows + warp * 16;       // first stacked query row of this warp
  const int n_blocks = (p.L + kBlk - 1) / kBlk;

  auto load_block = [&](int blk, int st) {               // all 128 threads: 64 rows x 8 chunks for K and for V^T
    const int kb0 = blk * kBlk;
    for (int i = threadIdx.x; i < kBlk * 8; i += kThreads) {
      const int r = i >> 3, c = i & 7;
      const int key = kb0 + r;                           // K: row = key, 16-byte chunk c = dims 8c ..
      cp16(sm[st][0] + r * kPitch + c * 16, p.k + static_cast<size_t>(min(key, p.L - 1)) * p.ld_k + head * kHd + c * 8, key < p.L);
      const int k8 = kb0 + c * 8;                        // V^T: row = dim r, chunk c = keys kb0 + 8c .. (ld_vt % 8 == 0)
      cp16(sm[st][1] + r * kPitch + c * 16, p.vt + static_cast<size_t>(head * kHd + r) * p.ld_vt + min(k8, p.ld_vt - 8), k8 < p.ld_vt);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_block(0, 0);

  // the two rows of this thread (grp and grp + 8 of the warp's 16), their pairs' mask rows
  int rr[2];
  const uint32_t* mrow_i[2];
  const uint32_t* mrow_j[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    rr[e] = row0 + grp + e * 8;
    const int pr = min(rr[e], p.rows - 1) / p.n_query;
    const int idx = p.pair_index ? p.pair_index[pr] : pr;
    mrow_i[e] = p.bits + static_cast<size_t>(idx / p.num_objects) * p.words;
    mrow_j[e] = p.bits + static_cast<size_t>(idx % p.num_objects) * p.words;
  }
  // Q as A fragments: a0 (row grp, dims kb*16 + quad*2 ..), a1 (row grp + 8), a2 / a3 (dims + 8)
  uint32_t qa[kHd / 16][4];
#pragma unroll
  for (int kb = 0; kb < kHd / 16; ++kb)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const __nv_bfloat16* src = p.q + static_cast<size_t>(min(rr[e], p.rows - 1)) * p.d_model + head * kHd + kb * 16 + quad * 2;
      qa[kb][e] = __ldg(reinterpret_cast<const uint32_t*>(src));
      qa[kb][e + 2] = __ldg(reinterpret_cast<const uint32_t*>(src + 8));
    }
  float o[kHd / 8][4];
#pragma unroll
  for (int n = 0; n < kHd / 8; ++n) { o[n][0] = 0.f; o[n][1] = 0.f; o[n][2] = 0.f; o[n][3] = 0.f; }
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  // ldmatrix rows: K (keys = n): matrix lane/8 = (keys +0 dims +0) (keys +0 dims +8) (keys +8 dims +0) (keys +8 dims +8);
  // V^T (dims = n, keys = k): (dims +0 keys +0) (dims +0 keys +8) (dims +8 keys +0) (dims +8 keys +8) -- the same pattern
  const int f_r = ((lane >> 4) & 1) * 8 + (lane & 7), f_c = ((lane >> 3) & 1) * 16;

  for (int blk = 0; blk < n_blocks; ++blk) {
    const int st = blk & 1;
    if (blk + 1 < n_blocks) load_block(blk + 1, st ^ 1);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    // mask words of this block's 64 keys for both rows (bits beyond L are cleared explicitly)
    uint32_t mw[2][2];
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int wi = blk * 2 + w;
        const int left = p.L - wi * 32;
        const uint32_t lim = left >= 32 ? 0xffffffffu : (left > 0 ? (1u << left) - 1u : 0u);
        mw[e][w] = wi < p.words ? ((__ldg(mrow_i[e] + wi) | __ldg(mrow_j[e] + wi)) & lim) : 0u;
      }
    // ---- S = Q K^T over the block's 8 key tiles -------------------------------------------------------------------------
    float s[kBlk / 8][4];
#pragma unroll
    for (int nt = 0; nt < kBlk / 8; nt += 2) {
      float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kb = 0; kb < kHd / 16; ++kb) {
        uint32_t b[4];
        ldsm4(b, sm[st][0] + (nt * 8 + f_r) * kPitch + kb * 32 + f_c);
        mma16816(c0, qa[kb], b[0], b[1]);
        mma16816(c1, qa[kb], b[2], b[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[nt][j] = c0[j]; s[nt + 1][j] = c1[j]; }
    }
    // ---- mask, online softmax ---------------------------------------------------------------------------------------------
    float bm[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < kBlk / 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = j >> 1;                                              // j = 0, 1: row grp; 2, 3: row grp + 8
        const int bit = (nt & 3) * 8 + quad * 2 + (j & 1);                 // key = blk*64 + nt*8 + quad*2 + (j & 1)
        const bool vis = (mw[e][nt >> 2] >> bit) & 1u;
        s[nt][j] = vis ? s[nt][j] : -INFINITY;
        bm[e] = fmaxf(bm[e], s[nt][j]);
      }
    float alpha[2], mn[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      bm[e] = fmaxf(bm[e], __shfl_xor_sync(0xffffffffu, bm[e], 1));
      bm[e] = fmaxf(bm[e], __shfl_xor_sync(0xffffffffu, bm[e], 2));
      mn[e] = fmaxf(m[e], bm[e]);
      alpha[e] = (m[e] == -INFINITY) ? 0.f : ex2((m[e] - mn[e]) * p.scale_log2e);     // mn >= m; both -inf only if nothing seen yet
      m[e] = mn[e];
    }
    uint32_t pa[kBlk / 16][4];
    float ls[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < kBlk / 8; ++nt) {
      float e4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = j >> 1;
        e4[j] = (mn[e] == -INFINITY) ? 0.f : ex2((s[nt][j] - mn[e]) * p.scale_log2e);  // exp2(-inf) = 0 for masked keys
        ls[e] += e4[j];
      }
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(e4[0], e4[1]);           // a0 / a2: row grp
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(e4[2], e4[3]);           // a1 / a3: row grp + 8
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      ls[e] += __shfl_xor_sync(0xffffffffu, ls[e], 1);
      ls[e] += __shfl_xor_sync(0xffffffffu, ls[e], 2);
      l[e] = l[e] * alpha[e] + ls[e];
    }
#pragma unroll
    for (int n = 0; n < kHd / 8; ++n) { o[n][0] *= alpha[0]; o[n][1] *= alpha[0]; o[n][2] *= alpha[1]; o[n][3] *= alpha[1]; }
    // ---- O += P V -----------------------------------------------------------------------------------------------------------
#pragma unroll
    for (int kt = 0; kt < kBlk / 16; ++kt) {
#pragma unroll
      for (int n = 0; n < kHd / 8; n += 2) {
        uint32_t b[4];
        ldsm4(b, sm[st][1] + (n * 8 + f_r) * kPitch + kt * 32 + f_c);
        mma16816(o[n], pa[kt], b[0], b[1]);
        mma16816(o[n + 1], pa[kt], b[2], b[3]);
      }
    }
    __syncthreads();                                                        // the stage is re-filled two iterations later
  }

  // ---- epilogue: O / l, or the mean of V for rows whose pair mask is empty (uniform attention over all L keys) -------------
  const bool any_empty = __any_sync(0xffffffffu, l[0] == 0.f || l[1] == 0.f);
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    if (rr[e] >= p.rows) continue;
    __nv_bfloat16* dst = p.out + static_cast<size_t>(rr[e]) * p.d_model + head * kHd;
    if (l[e] != 0.f) {
      const float inv = 1.f / l[e];
#pragma unroll
      for (int n = 0; n < kHd / 8; ++n)
        *reinterpret_cast<uint32_t*>(dst + n * 8 + quad * 2) = pack_bf16x2(o[n][e * 2] * inv, o[n][e * 2 + 1] * inv);
    }
  }
  if (any_empty) {                                                          // rare: every lane sums 2 dims x all keys of V^T
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      // rows are handled one at a time by the whole quad group that owns them (uniform inside a quad: same row)
      if (rr[e] >= p.rows || l[e] != 0.f) continue;
      __nv_bfloat16* dst = p.out + static_cast<size_t>(rr[e]) * p.d_model + head * kHd;
      for (int n = 0; n < kHd / 8; ++n) {
        float a0 = 0.f, a1 = 0.f;
        const __nv_bfloat16* v0 = p.vt + static_cast<size_t>(head * kHd + n * 8 + quad * 2) * p.ld_vt;
        for (int key = 0; key < p.L; ++key) { a0 += __bfloat162float(v0[key]); a1 += __bfloat162float(v0[p.ld_vt + key]); }
        *reinterpret_cast<uint32_t*>(dst + n * 8 + quad * 2) = pack_bf16x2(a0 / p.L, a1 / p.L);
      }
    }
  }
}

}  // namespace xal

int launch_xattn_pairs_long(const opsg_bf16* q, const opsg_bf16* k, int ld_k, const opsg_bf16* vt, int ld_vt, const uint32_t* bits,
                            int words, const int32_t* pair_index, int num_objects, int B, int n_query, int L, int num_heads,
                            int head_dim, opsg_bf16* ctx_out, cudaStream_t stream) {
  using namespace xal;
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(q && k && vt && bits && ctx_out, "xattn_pairs: null pointer");
  OPSG_CHECK_ARG(B > 0 && n_query > 0 && L > 0 && num_heads > 0 && num_objects > 0, "xattn_pairs: bad shape");
  if (head_dim != kHd) return set_error(OPSG_E_UNSUPPORTED, "xattn_pairs: head_dim %d unsupported (64 only)", head_dim);
  const int d_model = num_heads * head_dim;
  OPSG_CHECK_ARG(words * 32 >= L, "xattn_pairs: %d mask words do not cover L=%d keys", words, L);
  OPSG_CHECK_ARG(ld_k >= d_model && ld_k % 8 == 0 && ld_vt >= L && ld_vt % 8 == 0, "xattn_pairs: bad leading dims");
  OPSG_CHECK_ARG((((uintptr_t)q | (uintptr_t)k | (uintptr_t)vt | (uintptr_t)ctx_out) & 15) == 0, "xattn_pairs: operands must be 16-byte aligned");
  Params p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.k = reinterpret_cast<const __nv_bfloat16*>(k);
  p.vt = reinterpret_cast<const __nv_bfloat16*>(vt); p.out = reinterpret_cast<__nv_bfloat16*>(ctx_out);
  p.bits = bits; p.pair_index = pair_index;
  p.ld_k = ld_k; p.ld_vt = ld_vt; p.words = words; p.num_objects = num_objects; p.n_query = n_query; p.L = L;
  p.d_model = d_model; p.rows = B * n_query;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(head_dim));
  launch_kernel(xattn_pairs_long_kernel, dim3((p.rows + kRows - 1) / kRows, num_heads), kThreads, 0, stream, p);
  OPSG_CHECK_LAUNCH("xattn_pairs_long_kernel");
  return OPSG_OK;
}

}  // namespace opsg
