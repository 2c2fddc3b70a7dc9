// tcgen05 implicit-GEMM convolution on SCALED-SPLIT FP16 operands ("f16s") for sm_100a — the round-2 successor of the
// 3xTF32 kernel in gemm_conv_tc.cu (same tiling, same epilogue contract, same fp32-class accuracy, half the operand bytes
// and twice the tensor-core rate).
//
// Operand format.  An fp32 value x is carried as two halves:  hi = RN_f16(x),  lo = RN_f16((x - hi) * 2^11).
//   x - hi is exact in fp32 and |x - hi| <= 2^-11 |x|, so lo is O(|x|) again (no fp16 underflow of the correction) and
//   x = hi + lo * 2^-11 to 2^-22 relative (fp16 has the same 11 significant bits as TF32): the same representation error as
//   the TF32 hi / lo split, at 4 bytes per element instead of 8.  Activations live in HBM in this form next to (or instead
//   of) their fp32 copy: a "split" tensor is two half planes [2][N][H][W][ld] written by the PRODUCING kernel's epilogue, so
//   the consumer's TMA loads deliver tensor-core-ready operands — no split warps, no shared-memory rewrite (the 3xTF32
//   kernel's measured roof was shared-memory traffic: 176 KB per 32-float K slab, profiles/r1_tc_full_summary.md).
//   Range: |x| must stay below 65504 (hi saturates; counted in g_f16s_saturated, read with tt_f16s_saturation_count).
// Products.  Per K step (16 halves) two instructions, exactly the 3-product scheme of the TF32 kernel:
//   main | corr  +=  A_hi x [B_hi ; B_lo']      one N = 2 BN tcgen05.mma.kind::f16 (B_lo' follows B_hi in shared memory)
//   corr         +=  A_lo' x B_hi
//   result = main + 2^-11 corr  (the dropped lo x lo term is 2^-22 relative).  kind::f16 K = 16 per instruction, so a K of
//   128 costs the 8 truncating accumulations that 64 cost in TF32: chunks (TMEM ping-pong drains into fp32 registers, see
//   gemm_conv_tc.cu) are twice as long at the same accumulated rounding error.
// Everything else follows gemm_conv_tc.cu: TMA box loads per tap (zero fill = padding, element strides = stride 2), CTA pairs
// with multicast weight halves, persistent tiles, 8 epilogue warps with coalesced stores, split-K and GATHER (sparse conv).
// Warp roles: warp 0 = TMA / gather producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue.
#include <cuda.h>
#include <cuda_fp16.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

extern long long g_tt_launches;
extern int g_tt_debug;

__device__ unsigned int g_f16s_saturated = 0;     // number of values an epilogue had to clamp to the fp16 range

namespace {

constexpr int BM = 128;           // UMMA M
constexpr int KE = 64;            // K elements per stage (one 128-byte swizzle row of halves)
constexpr int NTHREADS = 320;     // warp 0 producer, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;
constexpr int MAX_KVOL = 32;
constexpr float H_MAX = 65504.f;

// kind::f16 instruction descriptor: c_format F32 (1) @4, a/b_format F16 (0) @7/@10, K-major, N >> 3 @17, M >> 4 @24
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
TT_DEVICE void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
TT_DEVICE void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
TT_DEVICE void tma_load_4d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
      : "memory");
}


// ---- cta_group::2 (CTA-pair MMA): one instruction of the pair's leader drives both SMs' tensor cores over M = 256 (128 rows from each
// CTA's own shared memory) and an N whose B rows are split between the two CTAs — each SM reads its A tile and only HALF of B.
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the pair's even-ranked CTA (cute: Sm100MmaPeerBitMask)
TT_DEVICE void umma_f16_2cta(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// completion of every tcgen05 operation this thread issued so far -> the barrier at this offset in BOTH CTAs of the pair
TT_DEVICE void tcgen05_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// TMA loads into this CTA's shared memory that signal the LEADER's mbarrier (the pair's MMA issuer waits on one barrier for both CTAs' operands)
TT_DEVICE void tma2_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
TT_DEVICE void tma2_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
TT_DEVICE void tma2_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// one arrival on the barrier at this offset in the pair's leader CTA (a plain local arrive when executed by the leader itself)
TT_DEVICE void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}

// x -> (hi, lo') with saturation to the fp16 range; returns true when x had to be clamped
TT_DEVICE bool split_h(float x, __half& hi, __half& lo) {
  const float c = fminf(fmaxf(x, -H_MAX), H_MAX);
  hi = __float2half_rn(c);
  lo = __float2half_rn((c - __half2float(hi)) * LO_SCALE);
  return c != x;
}
TT_DEVICE void split4(const float4& o, uint2& hi, uint2& lo, bool& sat) {
  __half h[4], l[4];
  sat |= split_h(o.x, h[0], l[0]);
  sat |= split_h(o.y, h[1], l[1]);
  sat |= split_h(o.z, h[2], l[2]);
  sat |= split_h(o.w, h[3], l[3]);
  hi.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
  hi.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
  lo.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
  lo.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
}

TT_DEVICE float4 load_split4(const __half* hi_ptr, long long plane) {
  const uint2 h = *reinterpret_cast<const uint2*>(hi_ptr);
  const uint2 l = *reinterpret_cast<const uint2*>(hi_ptr + plane);
  const __half2 h0 = *reinterpret_cast<const __half2*>(&h.x), h1 = *reinterpret_cast<const __half2*>(&h.y);
  const __half2 l0 = *reinterpret_cast<const __half2*>(&l.x), l1 = *reinterpret_cast<const __half2*>(&l.y);
  const float2 a = __half22float2(h0), b = __half22float2(h1), c = __half22float2(l0), e = __half22float2(l1);
  return make_float4(fmaf(c.x, LO_INV, a.x), fmaf(c.y, LO_INV, a.y), fmaf(e.x, LO_INV, b.x), fmaf(e.y, LO_INV, b.y));
}

struct HArgs {
  tt_conv_desc d;
  const float* bias;
  const float* res;
  const float* res2;
  const __half* res_s;   // residuals given as split planes instead of fp32 (same element offsets; lo' plane res*_plane further)
  const __half* res2_s;
  long long res_plane, res2_plane;
  float* y;              // fp32 output (may be NULL when only the split planes are wanted)
  __half* ys;            // split output: hi plane at the same element offsets as y; lo' plane ys_plane halves further (or NULL)
  long long ys_plane;
  int TH, TW;            // tile rectangle (TH * TW <= 128); flat mode: TH = 1, TW = 128 over N*H*W pixels
  int tiles_w, tiles_h;
  int flat;
  int n_slabs;           // ceil(Cin / 64)
  int chunk;             // stages accumulated inside TMEM before the epilogue folds them into fp32 registers
  int total_pix;
  int dbg;
  int m_tiles, n_tiles;
  // GATHER (tap-major sparse convolution)
  const __half* gx;      // split feature rows: hi plane [.][gx_ld], lo' plane gx_plane halves further
  long long gx_plane;
  int gx_ld;
  const int* pairs_in;
  const int* pairs_out;
  const int* pair_count;
  int kvol, pair_cap;
  const int* nbr;        // output-stationary: [cap_rows][kvol] input row under each tap of an output row, -1 = none
  const int* out_count;  // output-stationary: number of output rows (device scalar)
  int cap_rows;
  int tps;               // output-stationary: taps per 64-half K slab (2 when Cin == 32, else 1)
  int splits, k_per;     // split-K: K stages per split
  int corr_once;         // correction accumulator kept in TMEM for the whole K walk of a work item and drained once (see the MMA issuer)
};

// GM (gather mode): 0 = dense convolution (TMA box loads); 1 = tap-major sparse convolution (work item = one tap's pair tile,
// red.add epilogue); 2 = output-stationary sparse convolution (work item = an output-row tile, K runs over ALL taps through the
// neighbour table, plain fused epilogue — no atomics, no init / finish passes; pays for empty (row, tap) slots with zero rows).
template <int BN, int STAGES, int GM, bool PAIR, int EW, bool CO, int CG>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
conv_f16s_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const HArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int A_BYTES = BM * KE * 2;                    // 16 KB per plane
  constexpr int B_BYTES = BN * KE * 2;                    // per plane
  // CG == 2 (cta_group::2): the stage holds this CTA's A planes, ONE 128-row weight block (the hi plane in the leader, the lo' plane in
  // its peer: together the N = 2 BN operand [B_hi ; B_lo'] of the combined instruction) and this CTA's half of B_hi for the lo' x hi product
  static_assert(CG == 1 || (PAIR && GM == 0 && !CO), "cta_group::2 path: dense convolutions on CTA pairs only");
  constexpr int STAGE_BYTES = CG == 2 ? 2 * A_BYTES + B_BYTES + B_BYTES / 2 : 2 * A_BYTES + 2 * B_BYTES;
  constexpr int ACC_COLS = 2 * BN;                        // main | corr
  constexpr int TMEM_COLS = 2 * ACC_COLS;                 // ping-pong
  constexpr int SLAB = 16;
  constexpr int PITCH = SLAB + 4;
  constexpr int HN = BN / (EW / 4);                        // columns owned by one epilogue warp (EW / 4 warps per TMEM lane quarter)
  constexpr int RPP = EW == 16 ? 16 : 32;                 // rows per transposition phase: the slab buffers stay 20 KB in total
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* tile_all = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);              // [8 warps][32 rows][PITCH]
  long long* row_y = reinterpret_cast<long long*>(tile_all + 8 * 32 * PITCH);
  long long* row_r1 = row_y + BM;
  long long* row_r2 = row_r1 + BM;
  int* row_flag = reinterpret_cast<int*>(row_r2 + BM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(row_flag + BM);
  uint64_t* full = bars;                                  // [STAGES]  producer -> MMA (operands landed)
  uint64_t* empty = bars + STAGES;                        // [STAGES]  MMA -> producer
  uint64_t* acc_full = bars + 2 * STAGES;                 // [2]       MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * STAGES + 2;            // [2]       epilogue -> MMA
  uint64_t* corr_empty = bars + 2 * STAGES + 4;           // [2]       epilogue -> MMA (corr_once: correction buffer of work item t & 1 drained)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);
  int* sp_first = reinterpret_cast<int*>(tmem_slot + 2);
  int* sp_cnt = sp_first + MAX_KVOL + 1;

  const tt_conv_desc& d = p.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool GATHER = GM != 0;
  const int taps = GM == 2 ? p.kvol : d.KH * d.KW;
  // output-stationary with Cin = 32: a 64-half K slab holds TWO taps (p.tps = 2), so no half of a slab is padding
  const int k_iters = (GM == 2 && p.tps == 2) ? (p.kvol + 1) / 2 : taps * p.n_slabs;
  const int os_rows = GM == 2 ? min(*p.out_count, p.cap_rows) : 0;     // output-stationary: rows that exist (device count)
  // PAIR: CTA pairs (cluster of 2) walk two M tiles of the same N tile in lockstep and multicast weight halves to each other;
  // !PAIR: every CTA is on its own (loads the whole weight slab itself, no cross-CTA barriers) — measured alternative
  constexpr int CL = PAIR ? 2 : 1;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int pair0 = blockIdx.x / CL, pair_step = gridDim.x / CL;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], GATHER ? 33 : 1); mbar_init(&empty[s], CG == 2 ? 1 : CL); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], CG == 2 ? 2 * EW : EW); mbar_init(&corr_empty[b], EW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (GM == 1 && threadIdx.x == 64) {
    int acc = 0;
    for (int t = 0; t < p.kvol; ++t) {
      const int c = min(p.pair_count[t], p.pair_cap);
      sp_cnt[t] = c;
      sp_first[t] = acc;
      acc += ((c + BM - 1) / BM + CL - 1) / CL;
    }
    sp_first[p.kvol] = acc;
  }
  if (warp == 1) {
    if constexpr (CG == 2) {                                  // the same columns in both CTAs of the pair (issued by both)
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int os_tiles = (os_rows + BM - 1) / BM;
  const int total_pairs = (GM == 1 ? sp_first[p.kvol] : ((GM == 2 ? os_tiles : p.m_tiles) + CL - 1) / CL) * p.n_tiles * p.splits;
  auto decode = [&](int pt, int& nt, int& mt, int& tap, int& count, int& kb, int& ke) {
    const int ks = pt % p.splits;
    pt /= p.splits;
    kb = ks * p.k_per;
    ke = min(k_iters, kb + p.k_per);
    nt = pt % p.n_tiles;
    const int pm = pt / p.n_tiles;
    tap = 0; count = 0;
    if (GM == 2) {
      mt = CL * pm + (int)rank;
      count = os_rows;
    } else if (GM == 1) {
      while (sp_first[tap + 1] <= pm) ++tap;
      mt = CL * (pm - sp_first[tap]) + (int)rank;
      count = sp_cnt[tap];
    } else {
      mt = CL * pm + (int)rank;
    }
  };

  // weight slab of one stage: hi plane then lo' plane, BN rows each.  PAIR: this CTA loads its half of each plane and multicasts
  // it to both CTAs of the pair; !PAIR: it loads both halves for itself.
  auto load_weights = [&](uint8_t* st, uint64_t* bar, int k0, int tap, int n0) {
    uint8_t* bh = st + 2 * A_BYTES;
    if (PAIR) {
      tma_load_4d_mc(bh + rank * (B_BYTES / 2), &map_b, bar, k0, tap, n0 + (int)rank * (BN / 2), 0, 3);
      tma_load_4d_mc(bh + B_BYTES + rank * (B_BYTES / 2), &map_b, bar, k0, tap, n0 + (int)rank * (BN / 2), 1, 3);
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        tma_load_4d(bh + h * (B_BYTES / 2), &map_b, bar, k0, tap, n0 + h * (BN / 2), 0);
        tma_load_4d(bh + B_BYTES + h * (B_BYTES / 2), &map_b, bar, k0, tap, n0 + h * (BN / 2), 1);
      }
    }
  };

  if (warp == 0 && GATHER) {
    // ===================================================================== gather producer (whole warp) + weight TMA
    const uint32_t txb = 2 * B_BYTES;
    int ig = 0;
    for (int pt = pair0; pt < total_pairs; pt += pair_step) {
      int nt, mt, tap, count, kb, ke;
      decode(pt, nt, mt, tap, count, kb, ke);
      const int n0 = nt * BN;
      const int* pin = GM == 1 ? p.pairs_in + (long long)tap * p.pair_cap : nullptr;
      int rows[4];
      int nxa[4] = {-1, -1, -1, -1}, nxb[4] = {-1, -1, -1, -1};   // neighbour indices of the next stage / tap (GM == 2)
      if (GM == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int m = mt * BM + lane + 32 * i;
          rows[i] = m < count ? __ldg(pin + m) : -1;
        }
      }
      for (int it = kb; it < ke; ++it, ++ig) {
        const int s = ig % STAGES;
        if (GM == 2 && p.tps == 2) {
          // ---- two taps per K slab (Cin = 32): chunks 0..3 of a row come from the neighbour under tap 2 it, chunks 4..7 from tap 2 it + 1
          // (the neighbour indices of stage it + 1 are fetched while stage it's copies are issued: their latency was exposed once per stage)
          int ra[4], rb[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (it == kb) {
              const int m = mt * BM + lane + 32 * i;
              const int* nb = p.nbr + (long long)m * p.kvol + 2 * it;
              ra[i] = m < count ? __ldg(nb) : -1;
              rb[i] = (m < count && 2 * it + 1 < p.kvol) ? __ldg(nb + 1) : -1;
            } else {
              ra[i] = nxa[i]; rb[i] = nxb[i];
            }
          }
          if (it + 1 < ke) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int m = mt * BM + lane + 32 * i;
              const int* nb = p.nbr + (long long)m * p.kvol + 2 * (it + 1);
              nxa[i] = m < count ? __ldg(nb) : -1;
              nxb[i] = (m < count && 2 * (it + 1) + 1 < p.kvol) ? __ldg(nb + 1) : -1;
            }
          }
          mbar_wait(&empty[s], ((ig / STAGES) & 1) ^ 1);
          uint8_t* st = smem + s * STAGE_BYTES;
          if (lane == 0) {
            mbar_expect_tx(&full[s], txb);
            load_weights(st, &full[s], it * KE, 0, n0);              // weights viewed as [Cout][kvol * 32]: K runs across the taps
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = lane + 32 * i;
            const uint32_t dst = smem_u32(st) + r * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int row = j < 4 ? ra[i] : rb[i];
              const bool ok = row >= 0;
              const __half* src = p.gx + (long long)max(row, 0) * p.gx_ld + (j & 3) * 8;
              const uint32_t o = (uint32_t)((j ^ (r & 7)) << 4);
              cp_async16(dst + o, ok ? (const void*)src : (const void*)p.gx, ok ? 16 : 0);
              cp_async16(dst + A_BYTES + o, ok ? (const void*)(src + p.gx_plane) : (const void*)p.gx, ok ? 16 : 0);
            }
          }
          cp_async_arrive_noinc(&full[s]);
          continue;
        }
        const int ktap = GM == 2 ? it / p.n_slabs : tap;            // output-stationary: the K walk visits every tap
        const int slab = GM == 2 ? it - ktap * p.n_slabs : it;
        if (GM == 2 && (slab == 0 || it == kb)) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int m = mt * BM + lane + 32 * i;
            if (it == kb) rows[i] = m < count ? __ldg(p.nbr + (long long)m * p.kvol + ktap) : -1;       // -1: no input site under this tap
            else rows[i] = nxa[i];                                                                       // fetched one tap ahead
          }
          if (ktap + 1 < taps) {                              // next tap's neighbours: in flight while this tap's copies are issued
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int m = mt * BM + lane + 32 * i;
              nxa[i] = m < count ? __ldg(p.nbr + (long long)m * p.kvol + ktap + 1) : -1;
            }
          }
        }
        mbar_wait(&empty[s], ((ig / STAGES) & 1) ^ 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        if (lane == 0) {
          mbar_expect_tx(&full[s], txb);
          load_weights(st, &full[s], slab * KE, ktap, n0);
        }
        const int c0 = slab * KE;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = lane + 32 * i;
          const __half* src = p.gx + (long long)max(rows[i], 0) * p.gx_ld + c0;
          const uint32_t dst = smem_u32(st) + r * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {                       // 16-byte chunk j (8 halves) of row r lives at chunk j ^ (r & 7)
            const bool ok = rows[i] >= 0 && c0 + j * 8 < d.Cin;
            const uint32_t o = (uint32_t)((j ^ (r & 7)) << 4);
            cp_async16(dst + o, ok ? (const void*)(src + j * 8) : (const void*)p.gx, ok ? 16 : 0);
            cp_async16(dst + A_BYTES + o, ok ? (const void*)(src + p.gx_plane + j * 8) : (const void*)p.gx, ok ? 16 : 0);
          }
        }
        cp_async_arrive_noinc(&full[s]);
      }
    }
  } else if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      const uint32_t a_box = (uint32_t)(p.flat ? BM : p.TH * p.TW) * KE * 2;   // bytes one activation box (one plane) delivers
      const uint32_t tx = ((p.dbg & 2) ? 0u : 2 * a_box) + (CG == 2 ? (uint32_t)(B_BYTES + B_BYTES / 2) : 2 * B_BYTES);
      int ig = 0;
      for (int pt = pair0; pt < total_pairs; pt += pair_step) {
        int nt, mt, tap0, count, kb, ke;
        decode(pt, nt, mt, tap0, count, kb, ke);
        const int n0 = nt * BN;
        int cw0, ch0, cn;
        if (p.flat) { cw0 = mt * BM; ch0 = 0; cn = 0; }
        else {
          const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h;
          cn = mt / (p.tiles_w * p.tiles_h);
          cw0 = tw * p.TW * d.stride - d.pad; ch0 = th * p.TH * d.stride - d.pad;
        }
        for (int it = kb; it < ke; ++it, ++ig) {
          const int s = ig % STAGES;
          mbar_wait(&empty[s], ((ig / STAGES) & 1) ^ 1);
          uint8_t* st = smem + s * STAGE_BYTES;
          const int tap = it / p.n_slabs, slab = it - tap * p.n_slabs;
          const int kh = tap / d.KW, kw = tap - kh * d.KW;
          if constexpr (CG == 2) {
            // every load of BOTH CTAs completes on the leader's barrier: the leader announces the pair's bytes, the peer only loads
            // (its complete_tx may land before the announcement: the barrier's pending arrival keeps the phase open)
            if (rank == 0) mbar_expect_tx(&full[s], 2 * tx);
            if (!(p.dbg & 2)) {
              if (p.flat) {
                tma2_load_3d(st, &map_a, &full[s], slab * KE, cw0, 0);
                tma2_load_3d(st + A_BYTES, &map_a, &full[s], slab * KE, cw0, 1);
              } else {
                const int cw = cw0 + kw * d.dil, ch = ch0 + kh * d.dil;
                tma2_load_5d(st, &map_a, &full[s], slab * KE, cw, ch, cn, 0);
                tma2_load_5d(st + A_BYTES, &map_a, &full[s], slab * KE, cw, ch, cn, 1);
              }
            }
            uint8_t* bb = st + 2 * A_BYTES;                           // 128-row block: plane `rank` (hi in the leader, lo' in the peer)
            tma2_load_4d(bb, &map_b, &full[s], slab * KE, tap, n0, (int)rank);
            tma2_load_4d(bb + B_BYTES / 2, &map_b, &full[s], slab * KE, tap, n0 + BN / 2, (int)rank);
            tma2_load_4d(bb + B_BYTES, &map_b, &full[s], slab * KE, tap, n0 + (int)rank * (BN / 2), 0);   // this CTA's half of B_hi
            continue;
          }
          if ((p.dbg & 8) && ig >= STAGES) {                   // diagnosis: no loads after the ring's first fill (stale operands)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[s])) : "memory");
            continue;
          }
          mbar_expect_tx(&full[s], tx);
          if (!(p.dbg & 2)) {
            if (p.flat) {
              tma_load_3d(st, &map_a, &full[s], slab * KE, cw0, 0);
              tma_load_3d(st + A_BYTES, &map_a, &full[s], slab * KE, cw0, 1);
            } else {
              const int cw = cw0 + kw * d.dil, ch = ch0 + kh * d.dil;
              tma_load_5d(st, &map_a, &full[s], slab * KE, cw, ch, cn, 0);
              tma_load_5d(st + A_BYTES, &map_a, &full[s], slab * KE, cw, ch, cn, 1);
            }
          }
          load_weights(st, &full[s], slab * KE, tap, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one thread; cta_group::2: of the pair's leader only)
    if (lane == 0 && (CG == 1 || rank == 0)) {
      constexpr uint32_t idesc = make_idesc_f16(CG * BM, BN);
      constexpr uint32_t idesc2 = make_idesc_f16(CG * BM, 2 * BN);
      int ig = 0, cg = 0, tg = 0;
      // corr_once: the correction products (hi x lo' + lo' x hi, weighted 2^-11 in the result) tolerate the truncating accumulation
      // of the tensor core over the WHOLE K walk — their rounding error is 2^-11 times smaller than the main product's — so only
      // the main accumulator is folded into fp32 registers every `chunk` stages; the correction accumulator of work item t lives in
      // buffer t & 1 until the item's last chunk.  Halves the TMEM read traffic of the drains (TMEM reads are 64 B / clk / SM:
      // 128 KB per K = 128 chunk was 2048 clk against 1536 clk of MMA) at the price of a third N = BN instruction per K step.
      for (int pt = pair0; pt < total_pairs; pt += pair_step, ++tg) {
        int nt, mt, tap0, count, kb, ke;
        decode(pt, nt, mt, tap0, count, kb, ke);
        int it = kb;
        const int nch = (ke - kb + p.chunk - 1) / p.chunk;
        const uint32_t t_corr1 = tmem_base + (uint32_t)((tg & 1) * ACC_COLS + BN);
        if constexpr (CO) {
          mbar_wait(&corr_empty[tg & 1], ((tg >> 1) & 1) ^ 1);
          tcgen05_fence_after();
        }
        bool first_of_item = true;
        for (int c = 0; c < nch; ++c, ++cg) {
          const int b = cg & 1;
          mbar_wait(&acc_empty[b], ((cg >> 1) & 1) ^ 1);
          tcgen05_fence_after();
          const uint32_t t_main = tmem_base + (uint32_t)(b * ACC_COLS), t_corr = t_main + BN;
          const int it_end = min(it + p.chunk, ke);
          bool first = true;
          for (; it < it_end; ++it, ++ig) {
            const int s = ig % STAGES;
            mbar_wait(&full[s], (ig / STAGES) & 1);
            tcgen05_fence_after();
            const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + A_BYTES;
            const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
            for (int kk = 0; kk < KE / 16; ++kk) {           // UMMA K = 16 halves = 32 bytes inside the 128-byte swizzle row
              if (p.dbg & 4) break;
              const uint32_t off = kk * 32;
              const uint32_t acc = (first && kk == 0) ? 0u : 1u;
              if constexpr (CO) {
                umma_f16(t_main, make_smem_desc(a_hi + off), make_smem_desc(b_hi + off), idesc, acc);                           // hi*hi
                umma_f16(t_corr1, make_smem_desc(a_hi + off), make_smem_desc(b_lo + off), idesc, (first_of_item && kk == 0) ? 0u : 1u);   // hi*lo'
                umma_f16(t_corr1, make_smem_desc(a_lo + off), make_smem_desc(b_hi + off), idesc, 1);                            // lo'*hi
              } else if constexpr (CG == 2) {
                // M = 256 over the pair; B rows split between the CTAs: [B_hi (leader) ; B_lo' (peer)] and the two halves of B_hi
                umma_f16_2cta(t_main, make_smem_desc(a_hi + off), make_smem_desc(b_hi + off), idesc2, acc);            // hi*hi | hi*lo'
                umma_f16_2cta(t_corr, make_smem_desc(a_lo + off), make_smem_desc(b_hi + B_BYTES + off), idesc, 1);     // lo'*hi
              } else {
                umma_f16(t_main, make_smem_desc(a_hi + off), make_smem_desc(b_hi + off), idesc2, acc);   // hi*hi | hi*lo'
                umma_f16(t_corr, make_smem_desc(a_lo + off), make_smem_desc(b_hi + off), idesc, 1);      // lo'*hi
              }
            }
            first = false;
            first_of_item = false;
            if constexpr (CG == 2) tcgen05_commit_2cta(&empty[s]);
            else if (PAIR) tcgen05_commit_mc(&empty[s], 3);
            else tcgen05_commit(&empty[s]);
          }
          if constexpr (CG == 2) tcgen05_commit_2cta(&acc_full[b]); else tcgen05_commit(&acc_full[b]);
        }
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..9)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const long long yns = d.y_nstride ? d.y_nstride : (long long)d.yH * d.yW * d.y_ld;
    const int HWo = d.OH * d.OW;
    float* tile = tile_all + (warp - 2) * RPP * PITCH;
    bool sat = false;
    int cg = 0, tg = 0;
    for (int pt = pair0; pt < total_pairs; pt += pair_step, ++tg) {
      int nt, mt, tap, count, kb, ke;
      decode(pt, nt, mt, tap, count, kb, ke);
      const int nch = (ke - kb + p.chunk - 1) / p.chunk;
      const bool real_tile = GATHER ? mt * BM < count : mt < p.m_tiles;
      const int n0 = nt * BN + half * HN;
      if (GM == 2) {
        const int m = mt * BM + r;
        const bool valid = m < count;
        const long long r1 = (long long)m * d.res_ld + d.res_coff;
        if (half == 0) {
          row_y[r] = (long long)m * d.y_ld + d.y_coff;
          row_r1[r] = r1;
          row_r2[r] = 0;
          row_flag[r] = valid ? 1 : 0;
        }
      } else if (GM == 1) {
        const int m = mt * BM + r;
        const bool valid = m < count;
        if (half == 0) {
          const int orow = valid ? __ldg(p.pairs_out + (long long)tap * p.pair_cap + m) : 0;
          row_y[r] = (long long)orow * d.y_ld + d.y_coff;
          row_r1[r] = 0;
          row_r2[r] = 0;
          row_flag[r] = valid ? 1 : 0;
        }
      } else {
        bool valid;
        int nimg, oh, ow;
        long long rrow;
        if (p.flat) {
          const int pix = mt * BM + r;                       // (the host checked N * OH * OW < 2^31: 32-bit index math)
          valid = real_tile && pix < p.total_pix;
          nimg = pix / HWo;
          const int rem = pix - nimg * HWo;
          oh = rem / d.OW; ow = rem - oh * d.OW;
          rrow = pix;
        } else {
          const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h;
          nimg = mt / (p.tiles_w * p.tiles_h);
          const int rh = r / p.TW, rw = r - rh * p.TW;
          oh = th * p.TH + rh; ow = tw * p.TW + rw;
          valid = real_tile && (r < p.TH * p.TW) && oh < d.OH && ow < d.OW;
          rrow = ((long long)nimg * d.OH + oh) * d.OW + ow;
        }
        const long long r1 = (d.res_mode == TT_RES_UP2_NEAREST
                                  ? ((long long)nimg * d.res_H + (oh * d.res_H) / d.OH) * d.res_W + (ow * d.res_W) / d.OW
                                  : rrow) * d.res_ld + d.res_coff;
        const long long r2 = rrow * d.res2_ld + d.res2_coff;
        if (half == 0) {
          row_y[r] = nimg * yns + ((long long)(oh * d.oy_mul + d.oy_add) * d.yW + ow * d.ox_mul + d.ox_add) * d.y_ld + d.y_coff;
          row_r1[r] = r1;
          row_r2[r] = r2;
          row_flag[r] = (valid ? 1 : 0) | (nimg << 1);
        }
        if (valid && n0 < d.Cout) {
          if (p.res) {
#pragma unroll
            for (int j = 0; j < HN; j += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res + r1 + n0 + j));
          } else if (p.res_s) {
#pragma unroll
            for (int j = 0; j < HN; j += 64) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res_s + r1 + n0 + j));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res_s + p.res_plane + r1 + n0 + j));
            }
          }
          if (p.res2) {
#pragma unroll
            for (int j = 0; j < HN; j += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res2 + r2 + n0 + j));
          } else if (p.res2_s) {
#pragma unroll
            for (int j = 0; j < HN; j += 64) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res2_s + r2 + n0 + j));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res2_s + p.res2_plane + r2 + n0 + j));
            }
          }
        }
      }
      float sum[HN];
#pragma unroll
      for (int j = 0; j < HN; ++j) sum[j] = 0.f;
      for (int c = 0; c < nch; ++c, ++cg) {
        const int b = cg & 1;
        mbar_wait(&acc_full[b], (cg >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t t_main = tmem_base + lane_addr + (uint32_t)(b * ACC_COLS + half * HN);
        if constexpr (CO) {                                    // main accumulator only: the correction stays in TMEM until the last chunk
          if constexpr (EW == 16) {
#pragma unroll
            for (int c0 = 0; c0 < HN; c0 += 16) {
              uint32_t v[16];
              tmem_ld16(t_main + c0, v);
#pragma unroll
              for (int j = 0; j < 16; ++j) sum[c0 + j] += __uint_as_float(v[j]);
            }
          } else {
#pragma unroll
            for (int c0 = 0; c0 < HN; c0 += 32) {
              uint32_t v[32];
              tmem_ld32(t_main + c0, v);
#pragma unroll
              for (int j = 0; j < 32; ++j) sum[c0 + j] += __uint_as_float(v[j]);                          // fp32 RN
            }
          }
        } else if constexpr (EW == 16) {                       // register budget of the 576-thread variant: 16 columns at a time
#pragma unroll
          for (int c0 = 0; c0 < HN; c0 += 16) {
            uint32_t v[16], u[16];
            tmem_ld16(t_main + c0, v);
            tmem_ld16(t_main + BN + c0, u);
#pragma unroll
            for (int j = 0; j < 16; ++j) sum[c0 + j] += fmaf(__uint_as_float(u[j]), LO_INV, __uint_as_float(v[j]));
          }
        } else if (!(p.dbg & 16)) {                            // (knob 16: diagnosis — barriers only, no TMEM reads)
#pragma unroll
          for (int c0 = 0; c0 < HN; c0 += 32) {
            uint32_t v[32], u[32];
            tmem_ld32(t_main + c0, v);
            tmem_ld32(t_main + BN + c0, u);
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[c0 + j] += fmaf(__uint_as_float(u[j]), LO_INV, __uint_as_float(v[j]));   // fp32 RN
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_leader(&acc_empty[b]);        // the pair's MMA issuer waits for BOTH CTAs' drains
          else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[b])) : "memory");
        }
      }
      if constexpr (CO) {
        // every MMA of this work item is complete (the last chunk's commit covered them): fold the correction accumulator in, once
        const uint32_t t_corr = tmem_base + lane_addr + (uint32_t)((tg & 1) * ACC_COLS + BN + half * HN);
        if constexpr (EW == 16) {
#pragma unroll
          for (int c0 = 0; c0 < HN; c0 += 16) {
            uint32_t u[16];
            tmem_ld16(t_corr + c0, u);
#pragma unroll
            for (int j = 0; j < 16; ++j) sum[c0 + j] = fmaf(__uint_as_float(u[j]), LO_INV, sum[c0 + j]);
          }
        } else {
#pragma unroll
          for (int c0 = 0; c0 < HN; c0 += 32) {
            uint32_t u[32];
            tmem_ld32(t_corr + c0, u);
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[c0 + j] = fmaf(__uint_as_float(u[j]), LO_INV, sum[c0 + j]);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&corr_empty[tg & 1])) : "memory");
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
      const int sub = lane >> 2, cl = (lane & 3) * 4;
      // the 4 rows this lane stores (row groups of 8) are the same for every column slab: their table entries are read ONCE
      // per tile into registers (the profile showed the per-slab table reloads and the residual round trips as the epilogue's
      // long-scoreboard stalls on the short-K layers, where the epilogue sets the pace: profiles/r2_summary.md)
      int fl[4];
      long long yo[4], o1[4], o2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = q * 32 + i * 8 + sub;
        fl[i] = row_flag[rr];
        yo[i] = row_y[rr];
        o1[i] = row_r1[rr];
        o2[i] = row_r2[rr];
      }
#pragma unroll
      for (int sl = 0; sl < HN / SLAB; ++sl) {
        const int col = n0 + sl * SLAB + cl;
        const bool live = col < d.Cout && !(p.dbg & 1);
        // residual / bias loads of this slab are issued BEFORE the shared-memory transposition so that the two latencies overlap
        // (a deeper software pipeline — loads one slab ahead — pushed the staging arrays into local memory and cost 25 %: measured)
        float4 ra[4], rb[4];
        float4 bv0 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          rb[i] = ra[i];
          if (live && (fl[i] & 1)) {
            if (p.res) ra[i] = *reinterpret_cast<const float4*>(p.res + o1[i] + col);
            else if (p.res_s) ra[i] = load_split4(p.res_s + o1[i] + col, p.res_plane);
            if constexpr (EW != 16) {                           // (the 16-warp variant is not launched for layers with a second residual)
              if (p.res2) rb[i] = *reinterpret_cast<const float4*>(p.res2 + o2[i] + col);
              else if (p.res2_s) rb[i] = load_split4(p.res2_s + o2[i] + col, p.res2_plane);
            }
          }
        }
        if (live && p.bias && !d.bias_n_mod) bv0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));   // one bias vector: once per slab
#pragma unroll
        for (int ph = 0; ph < 32 / RPP; ++ph) {
          __syncwarp();
          if (lane / RPP == ph) {                                // (RPP == 32: every lane) this lane's row goes through the buffer now
#pragma unroll
            for (int j = 0; j < SLAB; j += 4)
              *reinterpret_cast<float4*>(&tile[(lane % RPP) * PITCH + j]) =
                  make_float4(sum[sl * SLAB + j], sum[sl * SLAB + j + 1], sum[sl * SLAB + j + 2], sum[sl * SLAB + j + 3]);
          }
          __syncwarp();
          if (live) {
#pragma unroll
            for (int ii = 0; ii < RPP / 8; ++ii) {
              const int i = ph * (RPP / 8) + ii;
              if (!(fl[i] & 1)) continue;
              const float4 acc4 = *reinterpret_cast<const float4*>(&tile[(ii * 8 + sub) * PITCH + cl]);
              float4 bv = bv0;
              if (p.bias && d.bias_n_mod) bv = __ldg(reinterpret_cast<const float4*>(p.bias + (long long)((fl[i] >> 1) % d.bias_n_mod) * d.Cout + col));
              float4 o = make_float4(acc4.x + bv.x + ra[i].x + rb[i].x, acc4.y + bv.y + ra[i].y + rb[i].y,
                                     acc4.z + bv.z + ra[i].z + rb[i].z, acc4.w + bv.w + ra[i].w + rb[i].w);
              const long long eo = yo[i] + col;
              if (GM == 1 || p.splits > 1) {                     // taps / K splits race on an output row: red.add (fp32 only)
                tt_red_add_v4(p.y + eo, o.x, o.y, o.z, o.w);
              } else {
                o = make_float4(tt_act(o.x, d.act), tt_act(o.y, d.act), tt_act(o.z, d.act), tt_act(o.w, d.act));
                if (p.y) *reinterpret_cast<float4*>(p.y + eo) = o;
                if (p.ys) {
                  uint2 hi, lo;
                  split4(o, hi, lo, sat);
                  *reinterpret_cast<uint2*>(p.ys + eo) = hi;
                  *reinterpret_cast<uint2*>(p.ys + p.ys_plane + eo) = lo;
                }
              }
            }
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
    }
    if (sat) atomicAdd(&g_f16s_saturated, 1u);
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    if constexpr (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// split-K / sparse companions.  init: y = bias + res + res2 before the tensor-core kernel red.adds its partial sums;
// finish: y = act(y) and, when asked, the split planes of the finished rows.
__global__ void f16s_splitk_init_kernel(const tt_conv_desc d, const float* __restrict__ bias, const float* __restrict__ res,
                                        const float* __restrict__ res2, const __half* __restrict__ res_s, long long res_plane,
                                        const __half* __restrict__ res2_s, long long res2_plane, float* __restrict__ y) {
  const int HWo = d.OH * d.OW, C4 = d.Cout / 4;
  const long long total = (long long)d.N * HWo * C4;
  const long long yns = d.y_nstride ? d.y_nstride : (long long)d.yH * d.yW * d.y_ld;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const long long pix = i / C4;
    const int n = (int)(pix / HWo);
    float4 v = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (res) { const float4 t = *reinterpret_cast<const float4*>(res + pix * d.res_ld + d.res_coff + c); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
    if (res2) { const float4 t = *reinterpret_cast<const float4*>(res2 + pix * d.res2_ld + d.res2_coff + c); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
    if (res_s) { const float4 t = load_split4(res_s + pix * d.res_ld + d.res_coff + c, res_plane); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
    if (res2_s) { const float4 t = load_split4(res2_s + pix * d.res2_ld + d.res2_coff + c, res2_plane); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
    *reinterpret_cast<float4*>(y + n * yns + (pix - (long long)n * HWo) * d.y_ld + d.y_coff + c) = v;
  }
}
__global__ void f16s_splitk_finish_kernel(const tt_conv_desc d, float* __restrict__ y, __half* __restrict__ ys, long long ys_plane) {
  const int HWo = d.OH * d.OW, C4 = d.Cout / 4;
  const long long total = (long long)d.N * HWo * C4;
  const long long yns = d.y_nstride ? d.y_nstride : (long long)d.yH * d.yW * d.y_ld;
  bool sat = false;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const long long pix = i / C4;
    const int n = (int)(pix / HWo);
    const long long eo = n * yns + (pix - (long long)n * HWo) * d.y_ld + d.y_coff + c;
    float4 v = *reinterpret_cast<const float4*>(y + eo);
    v = make_float4(tt_act(v.x, d.act), tt_act(v.y, d.act), tt_act(v.z, d.act), tt_act(v.w, d.act));
    if (d.act != TT_ACT_NONE) *reinterpret_cast<float4*>(y + eo) = v;
    if (ys) {
      uint2 hi, lo;
      split4(v, hi, lo, sat);
      *reinterpret_cast<uint2*>(ys + eo) = hi;
      *reinterpret_cast<uint2*>(ys + ys_plane + eo) = lo;
    }
  }
  if (sat) atomicAdd(&g_f16s_saturated, 1u);
}

// generic fp32 rows -> split planes (producers that are not convolutions: pooling, resampling, gating, staging of inputs)
__global__ void split_rows_kernel(const float* __restrict__ x, long long x_ld, __half* __restrict__ ys, long long ys_plane,
                                  long long ys_ld, long long rows, int cols4, const int* __restrict__ row_count) {
  const long long total = (row_count ? min((long long)*row_count, rows) : rows) * cols4;
  bool sat = false;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols4;
    const int c = (int)(i - r * cols4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + r * x_ld + c);
    uint2 hi, lo;
    split4(v, hi, lo, sat);
    *reinterpret_cast<uint2*>(ys + r * ys_ld + c) = hi;
    *reinterpret_cast<uint2*>(ys + ys_plane + r * ys_ld + c) = lo;
  }
  // not counted in g_f16s_saturated: this converter also sweeps arena regions no kernel has written yet (whole-buffer refreshes),
  // whose bit patterns are arbitrary; values that matter were produced — and counted — by a convolution epilogue
}

// split planes -> fp32 rows (a non-convolution kernel needs the fp32 value of a tensor that is stored as planes only)
__global__ void merge_rows_kernel(const __half* __restrict__ xs, long long xs_plane, long long xs_ld, float* __restrict__ y, long long y_ld,
                                  long long rows, int cols4) {
  const long long total = rows * cols4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols4;
    const int c = (int)(i - r * cols4) * 4;
    *reinterpret_cast<float4*>(y + r * y_ld + c) = load_split4(xs + r * xs_ld + c, xs_plane);
  }
}

// camera images NCHW fp32 (C <= 8) -> zero-bordered channels-last split planes with 8 halves per pixel: the layout the row-packed
// stem convolution reads (one 7-pixel tap row = 56 halves of a 64-half K slab).  One thread per pixel: coalesced plane reads,
// one 16-byte store per plane.
__global__ void image_to_split8_kernel(const float* __restrict__ x, __half* __restrict__ ys, long long ys_plane, int C, int H, int W,
                                       int out_H, int out_W, int top, int left, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const long long t = i / W;
    const int h = (int)(t % H), n = (int)(t / H);
    __half hi[8], lo[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float v = c < C ? __ldg(x + (((long long)n * C + c) * H + h) * W + w) : 0.f;
      split_h(v, hi[c], lo[c]);
    }
    const long long o = (((long long)n * out_H + h + top) * out_W + w + left) * 8;
    *reinterpret_cast<uint4*>(ys + o) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(ys + ys_plane + o) = *reinterpret_cast<const uint4*>(lo);
  }
}

// nn.Upsample(scale_factor=2, bilinear, align_corners=True) (lss.py:267) from fp32 straight into split planes: the same arithmetic
// as upsample2x_ac_kernel (elementwise.cu), no fp32 copy of the 4x larger map
__global__ void upsample2x_split_kernel(const float4* __restrict__ x, __half* __restrict__ ys, long long ys_plane, int N, int H, int W, int C4) {
  const int OH = 2 * H, OW = 2 * W;
  const float sh = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f;
  const float sw = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  const long long total = (long long)N * OH * OW * C4;
  bool sat = false;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C4;
    long long t = i / C4;
    const int ow = t % OW; t /= OW;
    const int oh = t % OH;
    const int n = t / OH;
    const float fy = sh * oh, fx = sw * ow;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const float4 a = __ldg(&x[((long long)(n * H + y0) * W + x0) * C4 + c]);
    const float4 b = __ldg(&x[((long long)(n * H + y0) * W + x1) * C4 + c]);
    const float4 cc = __ldg(&x[((long long)(n * H + y1) * W + x0) * C4 + c]);
    const float4 dd = __ldg(&x[((long long)(n * H + y1) * W + x1) * C4 + c]);
    float4 r;
    r.x = hy * (hx * a.x + lx * b.x) + ly * (hx * cc.x + lx * dd.x);
    r.y = hy * (hx * a.y + lx * b.y) + ly * (hx * cc.y + lx * dd.y);
    r.z = hy * (hx * a.z + lx * b.z) + ly * (hx * cc.z + lx * dd.z);
    r.w = hy * (hx * a.w + lx * b.w) + ly * (hx * cc.w + lx * dd.w);
    uint2 hi, lo;
    split4(r, hi, lo, sat);
    *reinterpret_cast<uint2*>(ys + i * 4) = hi;
    *reinterpret_cast<uint2*>(ys + ys_plane + i * 4) = lo;
  }
  if (sat) atomicAdd(&g_f16s_saturated, 1u);
}

// Thin 1x1 convolutions (Cin <= 64, Cout <= 64) over large maps: 128 x 64 tensor-core tiles spend their time in per-tile overheads
// there (one K stage per tile; seg / seg->feature heads at 224 x 448: ~10 % of the HBM roof), while the arithmetic is a per-pixel
// 64 x 64 mat-vec at most.  One thread = one pixel: it streams its channel row from the split planes (full 128-byte lines), keeps
// the COUT accumulators in registers (fp32 FFMA against fp32 weights broadcast from shared memory: exact products, no operand
// split needed) and stores its output row (fp32 and / or planes).
template <int COUT>
__global__ void __launch_bounds__(256) pointwise_f16s_kernel(const __half* __restrict__ xs, long long xs_plane, int x_ld, int Cin,
                                                              const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                                                              float* __restrict__ y, __half* __restrict__ ys, long long ys_plane, int y_ld,
                                                              int act, long long npix, int co0) {
  constexpr int RP = COUT * 2 + 16;                                  // staging row pitch in bytes (+16: conflict-free 16-byte accesses)
  __shared__ __align__(16) float ws[64 * COUT];                      // [k][co0 + co], rows beyond Cin / columns beyond Cout are zero
  __shared__ float bs[COUT];
  __shared__ __align__(16) uint8_t stage[8][32 * RP];                // per warp: 32 pixels x COUT halves (one plane at a time)
  const int K8 = (Cin + 7) / 8 * 8;
  for (int i = threadIdx.x; i < 64 * COUT; i += blockDim.x) {
    const int k = i / COUT, co = i - k * COUT;
    ws[i] = (k < Cin && co0 + co < Cout) ? w[(long long)k * Cout + co0 + co] : 0.f;
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) bs[i] = (bias && co0 + i < Cout) ? bias[co0 + i] : 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bool sat = false;
  // whole warps walk 32 consecutive pixels at a time (the tail warp clamps its loads and predicates its stores)
  for (long long base = (blockIdx.x * (long long)(blockDim.x >> 5) + warp) * 32; base < npix; base += (long long)gridDim.x * blockDim.x) {
    const long long pix = base + lane;
    const bool live = pix < npix;
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = bs[co];
    const __half* xr = xs + (live ? pix : npix - 1) * x_ld;
    for (int k0 = 0; k0 < K8; k0 += 8) {
      const uint4 h = *reinterpret_cast<const uint4*>(xr + k0);
      const uint4 l = *reinterpret_cast<const uint4*>(xr + xs_plane + k0);
      const __half2* hp = reinterpret_cast<const __half2*>(&h);
      const __half2* lp = reinterpret_cast<const __half2*>(&l);
      float xv[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = __half22float2(hp[j]), b = __half22float2(lp[j]);
        xv[2 * j] = fmaf(b.x, LO_INV, a.x);
        xv[2 * j + 1] = fmaf(b.y, LO_INV, a.y);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4* wr = reinterpret_cast<const float4*>(&ws[(k0 + j) * COUT]);
#pragma unroll
        for (int c4 = 0; c4 < COUT / 4; ++c4) {
          const float4 wv = wr[c4];                                    // same address in every lane: shared-memory broadcast
          acc[4 * c4] = fmaf(xv[j], wv.x, acc[4 * c4]);
          acc[4 * c4 + 1] = fmaf(xv[j], wv.y, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(xv[j], wv.z, acc[4 * c4 + 2]);
          acc[4 * c4 + 3] = fmaf(xv[j], wv.w, acc[4 * c4 + 3]);
        }
      }
    }
    // ---- outputs.  fp32 rows: direct 16-byte stores.  Planes: this lane's COUT halves go through a per-warp staging buffer so that the
    // global stores are 16 bytes per lane with 4 (COUT 32) / 2 (COUT 16) lanes on one pixel row: full 32-byte sectors instead of 8-byte shards
    uint2 hv[COUT / 4], lv[COUT / 4];
#pragma unroll
    for (int c4 = 0; c4 < COUT / 4; ++c4) {
      const float4 o = make_float4(tt_act(acc[4 * c4], act), tt_act(acc[4 * c4 + 1], act), tt_act(acc[4 * c4 + 2], act), tt_act(acc[4 * c4 + 3], act));
      if (y && live && co0 + 4 * c4 < Cout) *reinterpret_cast<float4*>(y + pix * y_ld + co0 + 4 * c4) = o;
      split4(o, hv[c4], lv[c4], sat);
    }
    if (ys) {
      constexpr int LPR = COUT / 8;                                    // lanes per pixel row (16-byte chunks of COUT halves)
      constexpr int RPI = 32 / LPR;                                    // pixel rows per store instruction
      const int chunk = lane % LPR, rsub = lane / LPR;
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
        __syncwarp();
#pragma unroll
        for (int c4 = 0; c4 < COUT / 4; ++c4) *reinterpret_cast<uint2*>(&stage[warp][lane * RP + c4 * 8]) = pl ? lv[c4] : hv[c4];
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 32 / RPI; ++it) {
          const int prow = it * RPI + rsub;
          const long long gp = base + prow;
          if (gp < npix && co0 + chunk * 8 < Cout)
            *reinterpret_cast<uint4*>(ys + (pl ? ys_plane : 0) + gp * y_ld + co0 + chunk * 8) =
                *reinterpret_cast<const uint4*>(&stage[warp][prow * RP + chunk * 16]);
        }
      }
    }
  }
  if (sat) atomicAdd(&g_f16s_saturated, 1u);
}

// sparse rows: y = act(y + res) for the first *count rows, plus their split planes
__global__ void f16s_sparse_finish_kernel(float* __restrict__ y, int y_ld, const float* __restrict__ res, int res_ld, int C,
                                          const int* __restrict__ count, int cap, int act, __half* __restrict__ ys, long long ys_plane) {
  const int C4 = C / 4;
  const long long total = (long long)min(*count, cap) * C4;
  bool sat = false;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C4;
    const int c = (int)(i - r * C4) * 4;
    float4 v = *reinterpret_cast<const float4*>(y + r * y_ld + c);
    if (res) { const float4 t = *reinterpret_cast<const float4*>(res + r * res_ld + c); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
    v = make_float4(tt_act(v.x, act), tt_act(v.y, act), tt_act(v.z, act), tt_act(v.w, act));
    *reinterpret_cast<float4*>(y + r * y_ld + c) = v;
    if (ys) {
      uint2 hi, lo;
      split4(v, hi, lo, sat);
      *reinterpret_cast<uint2*>(ys + r * y_ld + c) = hi;
      *reinterpret_cast<uint2*>(ys + ys_plane + r * y_ld + c) = lo;
    }
  }
  if (sat) atomicAdd(&g_f16s_saturated, 1u);
}
__global__ void f16s_sparse_init_kernel(float* __restrict__ y, int y_ld, const float* __restrict__ bias, int C,
                                        const int* __restrict__ count, int cap) {
  const int C4 = C / 4;
  const long long total = (long long)min(*count, cap) * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C4;
    const int c = (int)(i - r * C4) * 4;
    *reinterpret_cast<float4*>(y + r * y_ld + c) = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

int num_sms_cached() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); }
  return n;
}
constexpr int EPI_BYTES = 8 * 32 * 20 * 4 + BM * (3 * 8 + 4) + 512;

template <int BN, int STAGES, int GATHER, bool PAIR, int EW, bool CO, int CG = 1>
cudaError_t launch_f16s_co(cudaLaunchConfig_t& cfg, const CUtensorMap& ma, const CUtensorMap& mb, const HArgs& a) {
  constexpr int smem = STAGES * (2 * BM * KE * 2 + (CG == 2 ? 3 : 4) * BN * KE) + 1024 + EPI_BYTES;
  static bool set = false;
  if (!set) { cudaFuncSetAttribute(conv_f16s_kernel<BN, STAGES, GATHER, PAIR, EW, CO, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); set = true; }
  cfg.dynamicSmemBytes = smem;
  cfg.blockDim = dim3(64 + 32 * EW);
  return cudaLaunchKernelEx(&cfg, conv_f16s_kernel<BN, STAGES, GATHER, PAIR, EW, CO, CG>, ma, mb, a);
}
template <int BN, int STAGES, int GATHER, bool PAIR, int EW = 8>
cudaError_t launch_f16s_v(cudaLaunchConfig_t& cfg, const CUtensorMap& ma, const CUtensorMap& mb, const HArgs& a) {
  return a.corr_once ? launch_f16s_co<BN, STAGES, GATHER, PAIR, EW, true>(cfg, ma, mb, a) : launch_f16s_co<BN, STAGES, GATHER, PAIR, EW, false>(cfg, ma, mb, a);
}
// work items / grid for `tiles` M tiles; debug bit 0x100000 selects the unpaired variant
template <int BN, int STAGES, int GATHER>
cudaError_t launch_f16s(cudaLaunchConfig_t& cfg, const CUtensorMap& ma, const CUtensorMap& mb, const HArgs& a, long long m_units) {
  const bool pair = !(g_tt_debug & 0x100000);
  const int cl = pair ? 2 : 1;
  const int num_sms = num_sms_cached() - ((g_tt_debug >> 8) & 0xFF);
  const long long items = ((m_units + cl - 1) / cl) * a.n_tiles * a.splits;
  const long long max_items = num_sms / cl;
  cfg.gridDim = dim3((unsigned)(cl * (items < max_items ? items : max_items)));
  cfg.attrs[0].val.clusterDim.x = cl;
  // 16 epilogue warps for the 1x1 layers (dense BN = 128 tiles): with one to four K stages per tile the epilogue sets the pace, and
  // twice the warps hide its store / residual latencies (measured +17..22 % on 64->256 / 256->256 @112x224, +7 % on 2048->512); the
  // 3x3 trunk layers keep 8 (256->256 @112x224: -6 % with 16).  Debug bit 0x400000 turns the variant off.
  if constexpr (GATHER == 0) {
    if (pair && a.d.KH * a.d.KW == 1 && !(g_tt_debug & 0x400000) && !a.res2 && !a.res2_s) return launch_f16s_v<BN, STAGES, GATHER, true, 16>(cfg, ma, mb, a);
  }
  // cta_group::2 (one MMA of the pair's leader over M = 256, each SM reading only half of the weight rows).  Measured (profiles/r2_summary.md):
  // parity green, no gain on the long-K layers (341 -> 336 us; 8 x images: slower) — their pace is set by the barrier ring and the epilogue's
  // store phase, not by the operand reads.  Kept behind debug bit 0x2000000 as the starting point of a deeper-pipelined pair kernel.
  if constexpr (GATHER == 0 && BN == 128) {
    if (pair && (g_tt_debug & 0x2000000) && !a.corr_once) return launch_f16s_co<BN, STAGES, GATHER, true, 8, false, 2>(cfg, ma, mb, a);
  }
  return pair ? launch_f16s_v<BN, STAGES, GATHER, true>(cfg, ma, mb, a) : launch_f16s_v<BN, STAGES, GATHER, false>(cfg, ma, mb, a);
}

bool encode_weights(CUtensorMap* mb, const void* w_split, int Cin, int taps, int Cout, int BN, bool flat_k = false) {
  int cin_pad = (Cin + 7) & ~7;
  if (flat_k) { cin_pad *= taps; taps = 1; }                     // K = tap * Cin + c as ONE dimension (taps are contiguous per Cout row)
  cuuint64_t dims[4] = {(cuuint64_t)cin_pad, (cuuint64_t)taps, (cuuint64_t)Cout, 2};
  cuuint64_t str[3] = {(cuuint64_t)cin_pad * 2, (cuuint64_t)taps * cin_pad * 2, (cuuint64_t)Cout * taps * cin_pad * 2};
  cuuint32_t box[4] = {KE, 1, (cuuint32_t)(BN / 2), 1};          // each CTA of a pair loads (and multicasts) half of a plane's slab
  return encode_map(mb, w_split, 4, dims, str, box, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
}

}  // namespace

extern "C" {

/* number of epilogue threads that clamped a value to +-65504 since the last call (synchronises `stream`) */
int tt_f16s_saturation_count(unsigned int* out_host, int reset, tt_stream_t stream) {
  TT_REQUIRE(out_host, "tt_f16s_saturation_count", "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemcpyFromSymbolAsync(out_host, g_f16s_saturated, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) {
    tt_set_error("tt_f16s_saturation_count: %s", cudaGetErrorString(cudaGetLastError()));
    return TT_ERR_CUDA;
  }
  if (reset) {
    const unsigned int z = 0;
    if (cudaMemcpyToSymbolAsync(g_f16s_saturated, &z, sizeof(z), 0, cudaMemcpyHostToDevice, st) != cudaSuccess) return TT_ERR_CUDA;
    cudaStreamSynchronize(st);
  }
  return TT_OK;
}

int tt_split_f16(const float* x, long long x_ld, void* y_split, long long y_plane, long long y_ld, long long rows, int cols,
                 const int* row_count, tt_stream_t stream) {
  TT_REQUIRE(x && y_split && rows >= 0 && cols > 0, "tt_split_f16", "bad argument");
  TT_REQUIRE(cols % 4 == 0 && x_ld % 4 == 0 && y_ld % 4 == 0 && y_plane % 4 == 0 &&
                 ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(y_split) & 7) == 0),
             "tt_split_f16", "needs 4-element alignment");
  if (rows == 0) return TT_OK;
  const long long total = rows * (cols / 4);
  const int nb = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  split_rows_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(x, x_ld, static_cast<__half*>(y_split), y_plane, y_ld, rows, cols / 4, row_count);
  ++g_tt_launches;
  TT_CHECK_LAUNCH("tt_split_f16");
  return TT_OK;
}

int tt_image_to_split8(const float* x, void* y_split, long long y_plane, int N, int C, int H, int W, int out_H, int out_W, int top,
                       int left, tt_stream_t stream) {
  TT_REQUIRE(x && y_split && C >= 1 && C <= 8 && top >= 0 && left >= 0 && top + H <= out_H && left + W <= out_W && y_plane % 8 == 0 &&
                 (reinterpret_cast<uintptr_t>(y_split) & 15) == 0,
             "tt_image_to_split8", "bad arguments");
  const long long total = (long long)N * H * W;
  if (total == 0) return TT_OK;
  const int nb = (int)((total + 255) / 256 > 148 * 32 ? 148 * 32 : (total + 255) / 256);
  image_to_split8_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(x, static_cast<__half*>(y_split), y_plane, C, H, W, out_H, out_W, top, left, total);
  ++g_tt_launches;
  TT_CHECK_LAUNCH("tt_image_to_split8");
  return TT_OK;
}

int tt_upsample2x_bilinear_ac_split(const float* x, void* y_split, long long y_plane, int N, int H, int W, int C, tt_stream_t stream) {
  TT_REQUIRE(x && y_split && C % 4 == 0 && y_plane % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y_split) & 7) == 0,
             "tt_upsample2x_bilinear_ac_split", "C must be a multiple of 4, aligned pointers");
  const long long total = (long long)N * 4 * H * W * (C / 4);
  if (total == 0) return TT_OK;
  const int nb = (int)((total + 255) / 256 > 148 * 32 ? 148 * 32 : (total + 255) / 256);
  upsample2x_split_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x), static_cast<__half*>(y_split), y_plane, N, H, W, C / 4);
  ++g_tt_launches;
  TT_CHECK_LAUNCH("tt_upsample2x_bilinear_ac_split");
  return TT_OK;
}

int tt_pointwise_f16s(const tt_conv_desc* d, const tt_f16s_io* io, const float* w_kmajor, tt_stream_t stream) {
  TT_REQUIRE(d && io && io->x_split && w_kmajor && (io->y || io->y_split), "tt_pointwise_f16s", "null argument");
  const long long xhs = d->x_hstride ? d->x_hstride : (long long)d->W * d->x_ld;
  const long long xns = d->x_nstride ? d->x_nstride : (long long)d->H * xhs;
  const bool ok = d->KH == 1 && d->KW == 1 && d->stride == 1 && d->pad == 0 && d->groups == 1 && d->Cin <= 64 && d->Cout <= 64 && d->Cout % 4 == 0 &&
                  d->x_ld % 8 == 0 && (d->Cin + 7) / 8 * 8 <= d->x_ld && d->y_ld % 4 == 0 && d->y_coff % 4 == 0 && io->x_plane % 8 == 0 && io->y_plane % 4 == 0 &&
                  xhs == (long long)d->W * d->x_ld && xns == (long long)d->H * xhs && d->y_nstride == 0 && d->oy_mul == 1 && d->ox_mul == 1 &&
                  d->oy_add == 0 && d->ox_add == 0 && d->yH == d->OH && d->yW == d->OW && d->OH == d->H && d->OW == d->W &&
                  d->res_mode == TT_RES_NONE && !io->res && !io->res_split && !io->res2 && !io->res2_split && d->bias_n_mod == 0 &&
                  ((reinterpret_cast<uintptr_t>(io->x_split) | reinterpret_cast<uintptr_t>(io->y)) & 15) == 0 && (reinterpret_cast<uintptr_t>(io->y_split) & 7) == 0;
  if (!ok) {
    tt_set_error("tt_pointwise_f16s: needs a dense 1x1 stride-1 conv with Cin, Cout <= 64 (Cout %% 4), no residual, contiguous pixels");
    return TT_ERR_UNSUPPORTED;
  }
  const long long npix = (long long)d->N * d->H * d->W;
  if (npix == 0) return TT_OK;
  const int nb = (int)((npix + 255) / 256 > 148 * 8 ? 148 * 8 : (npix + 255) / 256);
  TT_REQUIRE(d->Cout % 8 == 0 || !io->y_split, "tt_pointwise_f16s", "plane output needs Cout % 8 == 0");
  const __half* xs = static_cast<const __half*>(io->x_split);
  __half* ys = static_cast<__half*>(io->y_split);
  float* y = io->y ? io->y + d->y_coff : nullptr;
  if (ys) ys += d->y_coff;
  cudaStream_t st = (cudaStream_t)stream;
  // 32 output channels per pass (register budget: 2 CTAs per SM); wider layers take two passes over the (small) input rows
  for (int co0 = 0; co0 < d->Cout; co0 += 32) {
    if (d->Cout <= 16) pointwise_f16s_kernel<16><<<nb, 256, 0, st>>>(xs, io->x_plane, d->x_ld, d->Cin, w_kmajor, io->bias, d->Cout, y, ys, io->y_plane, d->y_ld, d->act, npix, co0);
    else pointwise_f16s_kernel<32><<<nb, 256, 0, st>>>(xs, io->x_plane, d->x_ld, d->Cin, w_kmajor, io->bias, d->Cout, y, ys, io->y_plane, d->y_ld, d->act, npix, co0);
    if (co0) ++g_tt_launches;
  }
  ++g_tt_launches;
  TT_CHECK_LAUNCH("tt_pointwise_f16s");
  return TT_OK;
}

int tt_merge_f16(const void* x_split, long long x_plane, long long x_ld, float* y, long long y_ld, long long rows, int cols,
                 tt_stream_t stream) {
  TT_REQUIRE(x_split && y && rows >= 0 && cols > 0, "tt_merge_f16", "bad argument");
  TT_REQUIRE(cols % 4 == 0 && x_ld % 4 == 0 && y_ld % 4 == 0 && x_plane % 4 == 0 && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) &&
                 ((reinterpret_cast<uintptr_t>(x_split) & 7) == 0),
             "tt_merge_f16", "needs 4-element alignment");
  if (rows == 0) return TT_OK;
  const long long total = rows * (cols / 4);
  const int nb = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  merge_rows_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(static_cast<const __half*>(x_split), x_plane, x_ld, y, y_ld, rows, cols / 4);
  ++g_tt_launches;
  TT_CHECK_LAUNCH("tt_merge_f16");
  return TT_OK;
}

int tt_conv2d_f16s(const tt_conv_desc* d, const tt_f16s_io* io, tt_stream_t stream) {
  TT_REQUIRE(d && io && io->x_split && io->w_split && (io->y || io->y_split), "tt_conv2d_f16s", "null argument");
  const void* x_split = io->x_split; const long long x_plane = io->x_plane;
  const void* w_split = io->w_split; const float* bias = io->bias;
  const float* res = io->res; const float* res2 = io->res2;
  float* y = io->y; void* y_split = io->y_split; const long long y_plane = io->y_plane;
  TT_REQUIRE(d->res_mode == TT_RES_NONE || res != nullptr || io->res_split != nullptr, "tt_conv2d_f16s", "res_mode set without a residual");
  TT_REQUIRE(!(res && io->res_split) && !(res2 && io->res2_split), "tt_conv2d_f16s", "a residual is given either as fp32 or as split planes");
  const long long xhs = d->x_hstride ? d->x_hstride : (long long)d->W * d->x_ld;
  const long long xns = d->x_nstride ? d->x_nstride : (long long)d->H * xhs;
  const bool ok = d->groups == 1 && (d->stride == 1 || d->stride == 2) && d->Cout % 4 == 0 && d->x_ld % 8 == 0 && xhs % 8 == 0 &&
                  xns % 8 == 0 && x_plane % 8 == 0 && d->y_ld % 4 == 0 && d->y_coff % 4 == 0 && d->y_nstride % 4 == 0 && y_plane % 4 == 0 &&
                  (d->res_mode == TT_RES_NONE || (d->res_ld % 4 == 0 && d->res_coff % 4 == 0)) && d->res2_ld % 4 == 0 && d->res2_coff % 4 == 0 &&
                  io->res_plane % 4 == 0 && io->res2_plane % 4 == 0 &&
                  ((reinterpret_cast<uintptr_t>(x_split) | reinterpret_cast<uintptr_t>(w_split) | reinterpret_cast<uintptr_t>(y) |
                    reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(res2) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0 &&
                  ((reinterpret_cast<uintptr_t>(y_split) | reinterpret_cast<uintptr_t>(io->res_split) | reinterpret_cast<uintptr_t>(io->res2_split)) & 7) == 0 &&
                  d->OH == (d->H + 2 * d->pad - d->dil * (d->KH - 1) - 1) / d->stride + 1 &&
                  d->OW == (d->W + 2 * d->pad - d->dil * (d->KW - 1) - 1) / d->stride + 1;
  if (!ok) {
    tt_set_error("tt_conv2d_f16s: unsupported shape / alignment (groups 1, stride 1|2, x_ld %% 8, y_ld %% 4, 16-byte aligned pointers)");
    return TT_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int taps = d->KH * d->KW;
  const long long npix_in = (long long)d->N * d->H * d->W;
  HArgs a;
  memset(&a, 0, sizeof(a));
  a.d = *d;
  a.bias = bias; a.res = res; a.res2 = res2; a.y = y;
  a.res_s = static_cast<const __half*>(io->res_split); a.res_plane = io->res_plane;
  a.res2_s = static_cast<const __half*>(io->res2_split); a.res2_plane = io->res2_plane;
  a.ys = static_cast<__half*>(y_split); a.ys_plane = y_plane;
  a.chunk = (g_tt_debug & 0x40000) ? 1 : (g_tt_debug & 0x20000) ? 4 : 2;
  a.n_slabs = (d->Cin + KE - 1) / KE;
  if ((long long)d->N * d->OH * d->OW >= (1ll << 31) - BM || npix_in >= (1ll << 31) - BM) {
    tt_set_error("tt_conv2d_f16s: more than 2^31 pixels");
    return TT_ERR_UNSUPPORTED;
  }
  a.total_pix = d->N * d->OH * d->OW;
  a.flat = (taps == 1 && d->pad == 0 && d->stride == 1 && xhs == (long long)d->W * d->x_ld && xns == (long long)d->H * xhs) ? 1 : 0;
  CUtensorMap ma, mb;
  int grid_x;
  if (a.flat) {
    a.TH = 1; a.TW = BM; a.tiles_w = a.tiles_h = 1;
    cuuint64_t dims[3] = {(cuuint64_t)d->Cin, (cuuint64_t)npix_in, 2};
    cuuint64_t str[2] = {(cuuint64_t)d->x_ld * 2, (cuuint64_t)x_plane * 2};
    cuuint32_t box[3] = {KE, BM, 1};
    if (!encode_map(&ma, x_split, 3, dims, str, box, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16)) return TT_ERR_CUDA;
    grid_x = tt_cdiv(npix_in, BM);
  } else {
    int best_th = 1, best_tw = 1;
    double best = -1;
    for (int tw = 1; tw <= 128 && tw * d->stride <= 256; ++tw) {
      const int th = 128 / tw;
      if (th < 1) break;
      const double cover = (double)d->OW * d->OH / ((double)tt_cdiv(d->OW, tw) * tw * tt_cdiv(d->OH, th) * th);
      const double eff = cover * (tw * th) / 128.0;
      if (eff > best + 1e-9) { best = eff; best_th = th; best_tw = tw; }
    }
    a.TH = best_th; a.TW = best_tw;
    a.tiles_w = tt_cdiv(d->OW, a.TW); a.tiles_h = tt_cdiv(d->OH, a.TH);
    cuuint64_t dims[5] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N, 2};
    cuuint64_t str[4] = {(cuuint64_t)d->x_ld * 2, (cuuint64_t)xhs * 2, (cuuint64_t)xns * 2, (cuuint64_t)x_plane * 2};
    cuuint32_t box[5] = {KE, (cuuint32_t)(a.TW * d->stride), (cuuint32_t)(a.TH * d->stride), 1, 1};
    if (!encode_map(&ma, x_split, 5, dims, str, box, d->stride, CU_TENSOR_MAP_DATA_TYPE_FLOAT16)) return TT_ERR_CUDA;
    grid_x = a.tiles_w * a.tiles_h * d->N;
  }
  const int BN = d->Cout > 64 ? 128 : 64;
  if (!encode_weights(&mb, w_split, d->Cin, taps, d->Cout, BN)) return TT_ERR_CUDA;
  a.dbg = g_tt_debug & 0xFF;
  a.m_tiles = grid_x;
  a.n_tiles = tt_cdiv(d->Cout, BN);
  const int num_sms = num_sms_cached();
  const int k_iters = taps * a.n_slabs;
  a.splits = 1;
  a.k_per = k_iters;
  {
    const long long tile_pairs = (long long)((a.m_tiles + 1) / 2) * a.n_tiles;
    const bool plain = y != nullptr && d->oy_mul == 1 && d->ox_mul == 1 && d->oy_add == 0 && d->ox_add == 0 && d->yH == d->OH && d->yW == d->OW &&
                       d->res_mode != TT_RES_UP2_NEAREST && d->bias_n_mod == 0 && !(g_tt_debug & 128);
    if (plain && tile_pairs * 2 <= num_sms / 2 && k_iters >= 8) {
      int sp = (int)((num_sms / 2 + tile_pairs - 1) / tile_pairs);
      if (sp > k_iters / 4) sp = k_iters / 4;
      if (sp > 32) sp = 32;
      if (sp >= 2) {
        a.k_per = tt_cdiv(k_iters, sp);
        a.splits = tt_cdiv(k_iters, a.k_per);
      }
    }
  }
  const long long total4 = (long long)d->N * d->OH * d->OW * (d->Cout / 4);
  const int nb_ew = (int)((total4 + 255) / 256 > 1184 ? 1184 : (total4 + 255) / 256);
  if (a.splits > 1) {
    f16s_splitk_init_kernel<<<nb_ew, 256, 0, st>>>(*d, bias, d->res_mode != TT_RES_NONE ? res : nullptr, res2,
                                                   d->res_mode != TT_RES_NONE ? a.res_s : nullptr, a.res_plane, a.res2_s, a.res2_plane, y);
    ++g_tt_launches;
    TT_CHECK_LAUNCH("tt_conv2d_f16s(split-k init)");
    a.bias = nullptr; a.res = nullptr; a.res2 = nullptr; a.res_s = nullptr; a.res2_s = nullptr;
    a.d.act = TT_ACT_NONE;
    a.d.res_mode = TT_RES_NONE;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NTHREADS);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // Correction accumulator drained once per work item (corr_once) where the TMEM drains set the pace and the extra N = BN
  // instruction's operand reads do not: 1x1 layers with 2 .. 4 chunks per tile (K 256 .. 512).  Measured (profiles/r2_summary.md):
  // 1x1 256->256 @112x224 +23 %; 3x3 256->256 -9 .. -18 % (shared-memory operand reads bound the long-K layers); one-chunk layers
  // gain nothing.  Debug bits: 0x800000 forces it on everywhere, 0x1000000 off.
  a.corr_once = (g_tt_debug & 0x1000000) ? 0 : ((g_tt_debug & 0x800000) || (taps == 1 && BN == 128 && a.splits == 1 && k_iters >= 4 && k_iters <= 8)) ? 1 : 0;
  const cudaError_t lerr = BN == 128 ? launch_f16s<128, 3, 0>(cfg, ma, mb, a, a.m_tiles) : launch_f16s<64, 4, 0>(cfg, ma, mb, a, a.m_tiles);
  if (lerr != cudaSuccess) { tt_set_error("tt_conv2d_f16s: cluster launch failed: %s", cudaGetErrorString(lerr)); return TT_ERR_CUDA; }
  ++g_tt_launches;
  TT_CHECK_LAUNCH("tt_conv2d_f16s");
  if (a.splits > 1 && (d->act != TT_ACT_NONE || y_split)) {
    f16s_splitk_finish_kernel<<<nb_ew, 256, 0, st>>>(*d, y, static_cast<__half*>(y_split), y_plane);
    ++g_tt_launches;
    TT_CHECK_LAUNCH("tt_conv2d_f16s(split-k finish)");
  }
  return TT_OK;
}

int tt_sparse_conv_f16s(const tt_sparse_conv_desc* d, const void* feats_in_split, long long in_plane, const void* w_split,
                        const float* bias, const float* res, const int* pairs_in, const int* pairs_out, const int* pair_count,
                        const int* out_count, float* feats_out, void* out_split, long long out_plane, tt_stream_t stream) {
  TT_REQUIRE(d && feats_in_split && w_split && pairs_in && pairs_out && pair_count && out_count && feats_out, "tt_sparse_conv_f16s", "null argument");
  if (d->Cin % 8 || d->Cout % 4 || d->in_ld % 8 || d->out_ld % 4 || d->Cin < 32 || d->Cout < 32 || d->kvol > MAX_KVOL || in_plane % 8 || out_plane % 4 ||
      (d->res_ld % 4) ||
      ((reinterpret_cast<uintptr_t>(feats_in_split) | reinterpret_cast<uintptr_t>(w_split) | reinterpret_cast<uintptr_t>(feats_out) |
        reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(bias)) & 15) || (reinterpret_cast<uintptr_t>(out_split) & 7)) {
    tt_set_error("tt_sparse_conv_f16s: needs Cin %% 8, Cout %% 4, Cin / Cout >= 32, in_ld %% 8, out_ld %% 4, 16-byte aligned pointers");
    return TT_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (d->cap_out <= 0) return TT_OK;
  const long long tot4 = (long long)d->cap_out * (d->Cout / 4);
  const int nb = (int)((tot4 + 255) / 256 > 1184 ? 1184 : (tot4 + 255) / 256);
  f16s_sparse_init_kernel<<<nb, 256, 0, st>>>(feats_out, d->out_ld, bias, d->Cout, out_count, d->cap_out);
  ++g_tt_launches;
  TT_CHECK_LAUNCH("tt_sparse_conv_f16s(init)");
  HArgs a;
  memset(&a, 0, sizeof(a));
  a.d.N = a.d.H = a.d.W = a.d.OH = a.d.OW = a.d.yH = a.d.yW = 1;
  a.d.KH = a.d.KW = a.d.stride = a.d.dil = a.d.groups = 1;
  a.d.oy_mul = a.d.ox_mul = 1;
  a.d.Cin = d->Cin; a.d.x_ld = d->in_ld; a.d.Cout = d->Cout; a.d.y_ld = d->out_ld;
  a.d.act = TT_ACT_NONE;
  a.y = feats_out;
  a.chunk = (g_tt_debug & 0x40000) ? 1 : (g_tt_debug & 0x20000) ? 4 : 2;
  a.corr_once = (g_tt_debug & 0x800000) ? 1 : 0;
  a.n_slabs = (d->Cin + KE - 1) / KE;
  a.dbg = g_tt_debug & 0xFF;
  a.gx = static_cast<const __half*>(feats_in_split); a.gx_plane = in_plane; a.gx_ld = d->in_ld;
  a.pairs_in = pairs_in; a.pairs_out = pairs_out; a.pair_count = pair_count;
  a.kvol = d->kvol; a.pair_cap = d->pair_cap;
  a.splits = 1; a.k_per = a.n_slabs;
  const int BN = d->Cout > 64 ? 128 : 64;
  a.n_tiles = tt_cdiv(d->Cout, BN);
  CUtensorMap mb;
  if (!encode_weights(&mb, w_split, d->Cin, d->kvol, d->Cout, BN)) return TT_ERR_CUDA;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NTHREADS);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const long long cap_units = (long long)d->kvol * tt_cdiv(d->pair_cap, BM);      // upper bound of gathered M tiles over all taps
  const cudaError_t lerr = BN == 128 ? launch_f16s<128, 3, 1>(cfg, mb, mb, a, cap_units) : launch_f16s<64, 4, 1>(cfg, mb, mb, a, cap_units);
  if (lerr != cudaSuccess) { tt_set_error("tt_sparse_conv_f16s: cluster launch failed: %s", cudaGetErrorString(lerr)); return TT_ERR_CUDA; }
  ++g_tt_launches;
  TT_CHECK_LAUNCH("tt_sparse_conv_f16s");
  f16s_sparse_finish_kernel<<<nb, 256, 0, st>>>(feats_out, d->out_ld, res, d->res_ld, d->Cout, out_count, d->cap_out, d->act,
                                                 static_cast<__half*>(out_split), out_plane);
  ++g_tt_launches;
  TT_CHECK_LAUNCH("tt_sparse_conv_f16s(finish)");
  return TT_OK;
}

int tt_sparse_conv_os_f16s(const tt_sparse_conv_desc* d, const tt_f16s_io* io, const int* nbr, const int* out_count, tt_stream_t stream) {
  TT_REQUIRE(d && io && io->x_split && io->w_split && (io->y || io->y_split) && nbr && out_count, "tt_sparse_conv_os_f16s", "null argument");
  TT_REQUIRE(!(io->res && io->res_split) && !io->res2 && !io->res2_split, "tt_sparse_conv_os_f16s", "one residual, fp32 or split planes");
  if (d->Cin % 8 || d->Cout % 4 || d->in_ld % 8 || d->out_ld % 4 || d->Cin < 32 || d->Cout < 32 || d->kvol < 1 || io->x_plane % 8 || io->y_plane % 4 ||
      (d->res_ld % 4) || io->res_plane % 4 ||
      ((reinterpret_cast<uintptr_t>(io->x_split) | reinterpret_cast<uintptr_t>(io->w_split) | reinterpret_cast<uintptr_t>(io->y) |
        reinterpret_cast<uintptr_t>(io->res) | reinterpret_cast<uintptr_t>(io->bias)) & 15) ||
      ((reinterpret_cast<uintptr_t>(io->y_split) | reinterpret_cast<uintptr_t>(io->res_split)) & 7)) {
    tt_set_error("tt_sparse_conv_os_f16s: needs Cin %% 8, Cout %% 4, Cin / Cout >= 32, in_ld %% 8, out_ld %% 4, 16-byte aligned pointers");
    return TT_ERR_UNSUPPORTED;
  }
  if (d->cap_out <= 0) return TT_OK;
  HArgs a;
  memset(&a, 0, sizeof(a));
  a.d.N = a.d.H = a.d.W = a.d.OH = a.d.OW = a.d.yH = a.d.yW = 1;
  a.d.KH = a.d.KW = a.d.stride = a.d.dil = a.d.groups = 1;
  a.d.oy_mul = a.d.ox_mul = 1;
  a.d.Cin = d->Cin; a.d.x_ld = d->in_ld; a.d.Cout = d->Cout; a.d.y_ld = d->out_ld;
  a.d.act = d->act;
  a.d.res_mode = (io->res || io->res_split) ? TT_RES_SAME : TT_RES_NONE;
  a.d.res_ld = d->res_ld;
  a.bias = io->bias; a.res = io->res;
  a.res_s = static_cast<const __half*>(io->res_split); a.res_plane = io->res_plane;
  a.y = io->y; a.ys = static_cast<__half*>(io->y_split); a.ys_plane = io->y_plane;
  a.chunk = (g_tt_debug & 0x40000) ? 1 : (g_tt_debug & 0x20000) ? 4 : 2;
  a.corr_once = (g_tt_debug & 0x800000) ? 1 : 0;
  a.n_slabs = (d->Cin + KE - 1) / KE;
  a.dbg = g_tt_debug & 0xFF;
  a.gx = static_cast<const __half*>(io->x_split); a.gx_plane = io->x_plane; a.gx_ld = d->in_ld;
  a.kvol = d->kvol; a.nbr = nbr; a.out_count = out_count; a.cap_rows = d->cap_out;
  a.tps = (d->Cin == 32 && !(g_tt_debug & 0x200000)) ? 2 : 1;
  a.splits = 1; a.k_per = a.tps == 2 ? (d->kvol + 1) / 2 : d->kvol * a.n_slabs;
  const int BN = d->Cout > 64 ? 128 : 64;
  a.n_tiles = tt_cdiv(d->Cout, BN);
  CUtensorMap mb;
  if (!encode_weights(&mb, io->w_split, d->Cin, d->kvol, d->Cout, BN, a.tps == 2)) return TT_ERR_CUDA;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NTHREADS);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const long long cap_units = tt_cdiv(d->cap_out, BM);
  const cudaError_t lerr = BN == 128 ? launch_f16s<128, 3, 2>(cfg, mb, mb, a, cap_units) : launch_f16s<64, 4, 2>(cfg, mb, mb, a, cap_units);
  if (lerr != cudaSuccess) { tt_set_error("tt_sparse_conv_os_f16s: cluster launch failed: %s", cudaGetErrorString(lerr)); return TT_ERR_CUDA; }
  ++g_tt_launches;
  TT_CHECK_LAUNCH("tt_sparse_conv_os_f16s");
  return TT_OK;
}

}  // extern "C"
