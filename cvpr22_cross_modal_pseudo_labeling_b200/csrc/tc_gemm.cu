// tc_gemm.cu -- persistent, warp-specialised tcgen05 GEMM for the embedding head on sm_100a:
//
//   D[M, N] = A[M, K] . B[N, K]^T        bf16 operands (both K-major), fp32 accumulation in TMEM
//
// with two epilogues:
//   LINEAR   D + bias -> bf16 (and / or fp32): the emb_pred projection of FastRCNNPredictor
//            (modeling/roi_heads/box_head/roi_box_predictors.py:63-66, nn.Linear(in, EMB_DIM)) -- the GEMM
//            SURVEY 8f-2 asks to run on the tensor cores in front of the scoring GEMM;
//   SOFTMAX  row softmax / top-1 foreground label / optional raw logits: the scoring of
//            roi_box_predictors.py:67 + box_head/inference.py:62 (b200_embed_match, N <= 512).
//
// Why a second kernel next to embed_match_kernel (kept for the COLMAX mode): that one runs one CTA per
// 128-row tile with all N columns in one TMEM allocation -- MMA and epilogue of a CTA never overlap, and
// four epilogue warps walk 2 x N columns with two exps per element (cfg #5, 501 x 512: 0.60 ms = 16 % of
// the measured bf16 peak).  Here:
//   * persistent CTAs (one per SM), each walks row tiles m, m + grid, ...; a row tile is processed as
//     column blocks of <= 256 columns, so TMEM holds TWO accumulator buffers (2 x 256 columns):
//     the MMAs of unit j + 1 run while the epilogue warps drain unit j (acc_full / acc_empty mbarriers);
//   * warps 0, 1 = TMA producers (A box 128 x 64 / B box BN x 64, 128-byte swizzle, ring of smem stages),
//     warp 2 = MMA issuer (one lane, tcgen05.mma.kind::f16, M = 128, N = block width, K = 16 x 4 per stage),
//     warps 3-18 = epilogue: 4 warps per TMEM lane quadrant, each owns a quarter of the block's columns
//     (with 8 epilogue warps -- two per scheduler -- the sweeps of a 128 x 256 block took longer than its MMAs);
//   * SOFTMAX over up to two column blocks without a second GEMM pass and with ONE exp per element:
//     per block, sweep A = row max + best foreground column, sweep B = e = exp(l - block max) and the
//     row sum.  A block that is not the row's last parks e as fp16 in shared memory ("stash") and frees
//     its TMEM buffer at once; the last block writes e back into its own TMEM columns (tcgen05.st).
//     When the last block's sweep B is done the row's max / sum over all blocks are known:
//     probabilities = e * exp(block max - row max) / sum, streamed from the stash with full-row
//     coalescing (a warp per row) and from TMEM for the last block.
// Numerics: fp32 accumulation of bf16 products (as embed_match_kernel); exp2f of (l - max) * log2(e);
// e in (0, 1] is cut to fp16 precision (relative 2^-10: far inside the 2e-2 absolute bar of BASELINE.json).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace b200 {
namespace {

constexpr int kBM = 128;    // rows per tile (= TMEM lanes)
constexpr int kBK = 64;     // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;  // K of one tcgen05.mma.kind::f16
constexpr int kEpiWarps = 16;
constexpr int kCH = kEpiWarps / 4;  // epilogue warps per TMEM lane quadrant: each owns 1 / kCH of a block's columns
constexpr int kFirstEpiWarp = 3;  // warps 0, 1: TMA producers (A / B operand), warp 2: MMA issuer
constexpr int kGemmThreads = 32 * (kFirstEpiWarp + kEpiWarps);
constexpr int kMaxStages = 8;
constexpr int kMaxBN = 256;
constexpr int kStashPitch = 2 * kMaxBN + 16;  // bytes per stash row: 256 fp16 + 16 (rows shift by 4 banks)
constexpr float kLog2e = 1.4426950408889634f;

struct GemmParams {
  long long M;
  int N, K;
  int BN;          // columns per block (multiple of 16, <= 256)
  int nblk;        // column blocks per row tile
  int stages;
  int buf_stride;  // TMEM columns per accumulator buffer
  int tmem_cols;   // allocation = 2 * buf_stride (power of two >= 32)
  uint32_t a_bytes, stage_bytes;
  // LINEAR
  const float* bias;
  __nv_bfloat16* out_bf16;
  float* out_f32;
  int ldo;
  // SOFTMAX
  float score_thresh;
  float* probs;
  float* logits;
  int32_t* top_label;
  float* top_prob;
  int ld;
  int passes;  // SOFTMAX_WIDE: 2 (statistics, then probabilities) or 1 (no probabilities wanted)
};

enum { kEpiLinear = 0, kEpiSoftmax = 1, kEpiSoftmaxWide = 2 };

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// K-major operand tile with 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// mbarrier wait that parks the thread in hardware until the phase completes (suspend-time hint), instead of
// re-issuing try_wait from a spin loop: with 19 warps per SM the polling of the waiting roles took a fifth
// of the issue slots away from the epilogue warps (profiles/r02_match_cfg5_probs.txt)
__device__ __forceinline__ void mbar_wait_park(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@!p bra WAIT_%=;\n\t}" ::"r"(smem_u32(bar)),
      "r"(phase), "r"(0x989680u)
      : "memory");
}
// 2^x for x <= 0: one MUFU.EX2 (exp2f() adds range fix-ups for arguments the softmax never produces)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bar_named(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// per-row statistics of the softmax epilogue, [block][column half][row]
// (12 KB: with the 66 KB stash this still leaves three 48 KB pipeline stages at 512 columns)
struct SoftStats {
  float sum[2][kCH][kBM];
  float best[2][kCH][kBM];       // best foreground logit of the part (-inf if it has none)
  uint16_t bestc[2][kCH][kBM];   // its column
  float bg[2][kBM];              // logit of column 0 (the background row) in block 0, -inf in block 1
  float fac[2][kBM];  // probability = e * fac[block][row]
};

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const long long n_mtiles = (p.M + kBM - 1) / kBM;
  const int NP = (p.N + 15) / 16 * 16;
  // work units: SOFTMAX keeps all column blocks of a row tile on one CTA (row statistics); LINEAR hands out
  // (row tile, column block) units one by one -- 500 tiles x 3 blocks over 148 CTAs is 10.1 units each,
  // whole row tiles would be 3.4 (one CTA in three does a fourth: 84 % at best)
  const long long n_outer = EPI == kEpiLinear ? n_mtiles * p.nblk : n_mtiles;
  // SOFTMAX_WIDE (more than two column blocks: the parked-block scheme would need all of them in shared
  // memory): every row tile is multiplied TWICE -- pass 0 leaves only the row statistics, pass 1 recomputes
  // the blocks and writes probabilities.  The [M, N] logits make no HBM round trip (the first version of the
  // wide path wrote them, read them back in a softmax kernel and wrote the probabilities: 12 bytes per
  // element; now 4, or 8 when the caller also wants the logits), at twice the tensor work of a class matrix
  // that is wide but not tall.
  const int n_inner = EPI == kEpiLinear ? 1 : EPI == kEpiSoftmax ? p.nblk : p.passes * p.nblk;
  // dynamic shared memory may start at any 16-byte boundary: realign for the 128 B swizzle
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  // SOFTMAX only: fp16 stash of e = exp(l - max), row statistics, per-warp transpose tiles for the logits
  unsigned char* stash = tiles + (size_t)p.stages * p.stage_bytes;
  SoftStats* stats = reinterpret_cast<SoftStats*>(stash + (EPI == kEpiSoftmax ? kBM * kStashPitch : 0));
  float* xpose = reinterpret_cast<float*>(stats + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 2);  // one arrive.expect_tx per producer warp
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp < 2) {
    // ===== TMA producers: warp 0 streams the A tiles, warp 1 the B tiles.  One thread issues a bulk tensor
    // copy every ~250-500 cycles whatever its size (scripts/micro/tma_box.cu): with both operands on one
    // thread the 32 requests of a 128 x 512 tile took longer than its MMAs =====
    if (lane == 0) {
      uint32_t it = 0;
      const uint32_t my_bytes = warp == 0 ? p.a_bytes : p.stage_bytes - p.a_bytes;
      const CUtensorMap* map = warp == 0 ? &map_a : &map_b;
      for (long long o = blockIdx.x; o < n_outer; o += gridDim.x)
        for (int in = 0; in < n_inner; ++in) {
          const long long mt = EPI == kEpiLinear ? o / p.nblk : o;
          const int blk = EPI == kEpiLinear ? (int)(o % p.nblk) : in % p.nblk;
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % p.stages;
            const uint32_t ph = (it / p.stages) & 1;
            mbar_wait_park(&empty_bar[s], ph ^ 1);  // first pass falls through (barrier still in phase 0)
            unsigned char* dst = tiles + (size_t)s * p.stage_bytes + (warp == 0 ? 0u : p.a_bytes);
            mbar_arrive_expect_tx(&full_bar[s], my_bytes);
            tma_load_2d(dst, map, &full_bar[s], kb * kBK, warp == 0 ? (int)(mt * kBM) : blk * p.BN);
          }
        }
    }
  } else if (warp == 2) {
    // ===== MMA issuer (one elected lane) =====
    if (lane == 0) {
      uint32_t it = 0, j = 0;
      for (long long o = blockIdx.x; o < n_outer; o += gridDim.x)
        for (int in = 0; in < n_inner; ++in, ++j) {
          const int blk = EPI == kEpiLinear ? (int)(o % p.nblk) : in % p.nblk;
          const int buf = j & 1;
          mbar_wait_park(&acc_empty[buf], ((j >> 1) & 1) ^ 1);  // the epilogue has drained this buffer
          tc_fence_after();
          const int bn = min(p.BN, NP - blk * p.BN);
          const uint32_t idesc = umma_idesc(bn);
          const uint32_t d_addr = tmem_base + (uint32_t)(buf * p.buf_stride);
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % p.stages;
            mbar_wait_park(&full_bar[s], (it / p.stages) & 1);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(tiles + (size_t)s * p.stage_bytes);
            const uint32_t b_addr = a_addr + p.a_bytes;
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k)
              umma_bf16(d_addr, umma_desc_sw128(a_addr + k * kUmmaK * 2), umma_desc_sw128(b_addr + k * kUmmaK * 2), idesc,
                        (kb | k) ? 1u : 0u);
            umma_commit(&empty_bar[s]);  // stage free once these MMAs have read it
          }
          umma_commit(&acc_full[buf]);  // accumulator of this unit complete
        }
    }
  } else {
    // ===== epilogue: thread <-> TMEM lane <-> row of the tile; two warps per lane quadrant =====
    const int quad = warp & 3;          // a warp may only touch TMEM lanes [32 * (warp % 4), +32)
    const int ch = (warp - kFirstEpiWarp) >> 2;  // which part of the block's columns
    const int trow = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t j = 0;
    for (long long o = blockIdx.x; o < n_outer; o += gridDim.x) {
      const long long mt = EPI == kEpiLinear ? o / p.nblk : o;
      const long long row = mt * kBM + trow;
      const bool row_ok = row < p.M;
      // SOFTMAX_WIDE: running statistics of this thread's part of the row over the blocks of pass 0, then the
      // row's maximum (x log2 e) and 1 / sum for pass 1
      float w_rm = -INFINITY, w_rs = 0.f, w_best = -INFINITY, w_bg = -INFINITY, w_moff = 0.f, w_inv = 0.f;
      int w_bestc = 0;
      for (int in = 0; in < n_inner; ++in, ++j) {
        const int blk = EPI == kEpiLinear ? (int)(o % p.nblk) : in % p.nblk;
        const int buf = j & 1;
        const int col0 = blk * p.BN;
        const int bn = min(p.BN, NP - col0);
        const int part = ((bn / 16 + kCH - 1) / kCH) * 16;
        const int c_lo = min(ch * part, bn), c_hi = min(c_lo + part, bn);
        const uint32_t taddr = lane_addr + (uint32_t)(buf * p.buf_stride);
        mbar_wait_park(&acc_full[buf], (j >> 1) & 1);
        tc_fence_after();
        float v[16];
        if (EPI == kEpiLinear) {
          for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
            tmem_ld16(taddr + c0, v);
            const int c = col0 + c0;
            if (p.bias) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c + i < p.N) v[i] += __ldg(p.bias + c + i);
            }
            if (!row_ok) continue;
            if (p.out_bf16) {
              __nv_bfloat16* o = p.out_bf16 + row * p.ldo + c;
              if (c + 16 <= p.N && (p.ldo & 7) == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                uint4 w0, w1;
                __nv_bfloat162 t;
                t = __floats2bfloat162_rn(v[0], v[1]);   w0.x = *reinterpret_cast<uint32_t*>(&t);
                t = __floats2bfloat162_rn(v[2], v[3]);   w0.y = *reinterpret_cast<uint32_t*>(&t);
                t = __floats2bfloat162_rn(v[4], v[5]);   w0.z = *reinterpret_cast<uint32_t*>(&t);
                t = __floats2bfloat162_rn(v[6], v[7]);   w0.w = *reinterpret_cast<uint32_t*>(&t);
                t = __floats2bfloat162_rn(v[8], v[9]);   w1.x = *reinterpret_cast<uint32_t*>(&t);
                t = __floats2bfloat162_rn(v[10], v[11]); w1.y = *reinterpret_cast<uint32_t*>(&t);
                t = __floats2bfloat162_rn(v[12], v[13]); w1.z = *reinterpret_cast<uint32_t*>(&t);
                t = __floats2bfloat162_rn(v[14], v[15]); w1.w = *reinterpret_cast<uint32_t*>(&t);
                reinterpret_cast<uint4*>(o)[0] = w0;
                reinterpret_cast<uint4*>(o)[1] = w1;
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (c + i < p.N) o[i] = __float2bfloat16_rn(v[i]);
              }
            }
            if (p.out_f32) {
              float* o = p.out_f32 + row * p.ldo + c;
              if (c + 16 <= p.N && (p.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (c + i < p.N) o[i] = v[i];
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        } else if (EPI == kEpiSoftmaxWide) {
          const int pass = in / p.nblk;
          // logits / probabilities leave through a warp-private transpose tile: half a warp writes 64
          // contiguous bytes of one row (2 rows per store instruction), whatever the row pitch
          float* tt = xpose + (warp - kFirstEpiWarp) * (32 * 17);
          const int rsub = lane >> 4, csub = lane & 15;
          auto store_chunk = [&](float* outp, int c) {
#pragma unroll
            for (int i = 0; i < 16; ++i) tt[lane * 17 + i] = v[i];
            __syncwarp();
#pragma unroll 4
            for (int rr = 0; rr < 32; rr += 2) {
              const long long orow = mt * kBM + quad * 32 + rr + rsub;
              if (orow < p.M && c + csub < p.N) outp[orow * p.ld + c + csub] = tt[(rr + rsub) * 17 + csub];
            }
            __syncwarp();
          };
          if (pass == 0) {
            for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
              tmem_ld16(taddr + c0, v);
              const int c = col0 + c0;
              if (p.logits) store_chunk(p.logits, c);
              if (c + 16 > p.N) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (c + i >= p.N) v[i] = -INFINITY;
              }
              if (c == 0) {
                w_bg = v[0];
                v[0] = -INFINITY;
              }
              const float before = w_best;
              int bi = 0;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                bi = v[i] > w_best ? i : bi;
                w_best = fmaxf(w_best, v[i]);
              }
              w_bestc = w_best > before ? c + bi : w_bestc;
              if (c == 0) v[0] = w_bg;   // the background logit counts in the sum
              const float nm = fmaxf(w_rm, fmaxf(w_best, w_bg));
              if (nm > w_rm) {           // the running sum follows the running maximum
                w_rs *= ex2_approx((w_rm - nm) * kLog2e);
                w_rm = nm;
              }
              const float moff = w_rm * kLog2e;
#pragma unroll
              for (int i = 0; i < 16; ++i) w_rs += ex2_approx(fmaf(v[i], kLog2e, -moff));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            if (blk == p.nblk - 1) {
              // the four parts of the row -> its maximum, 1 / sum and top-1 foreground label
              stats->sum[0][ch][trow] = w_rs;
              stats->sum[1][ch][trow] = w_rm;
              stats->best[0][ch][trow] = w_best;
              stats->best[1][ch][trow] = __int_as_float(w_bestc);
              bar_named(1 + quad, 32 * kCH);
              float rmx = -INFINITY;
#pragma unroll
              for (int h = 0; h < kCH; ++h) rmx = fmaxf(rmx, stats->sum[1][h][trow]);
              float tot = 0.f;
#pragma unroll
              for (int h = 0; h < kCH; ++h) tot += stats->sum[0][h][trow] * ex2_approx((stats->sum[1][h][trow] - rmx) * kLog2e);
              w_moff = rmx * kLog2e;
              w_inv = 1.0f / tot;
              if (ch == 0 && row_ok && p.top_label) {
                float bb = -INFINITY;
                int bc = 0;
#pragma unroll
                for (int h = 0; h < kCH; ++h)
                  if (stats->best[0][h][trow] > bb) {   // parts in column order per block; see below for blocks
                    bb = stats->best[0][h][trow];
                    bc = __float_as_int(stats->best[1][h][trow]);
                  } else if (stats->best[0][h][trow] == bb && __float_as_int(stats->best[1][h][trow]) < bc) {
                    bc = __float_as_int(stats->best[1][h][trow]);   // equal maxima: the lower column wins
                  }
                const float bp = p.N > 1 ? ex2_approx(fmaf(bb, kLog2e, -w_moff)) * w_inv : 0.f;
                p.top_label[row] = (p.N > 1 && bp > p.score_thresh) ? bc : 0;
                if (p.top_prob) p.top_prob[row] = bp;
              }
              bar_named(1 + quad, 32 * kCH);   // the statistics may be overwritten by the next row tile
              w_rm = -INFINITY, w_rs = 0.f, w_best = -INFINITY, w_bg = -INFINITY, w_bestc = 0;
            }
          } else {
            for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
              tmem_ld16(taddr + c0, v);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = ex2_approx(fmaf(v[i], kLog2e, -w_moff)) * w_inv;
              store_chunk(p.probs, col0 + c0);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
          }
        } else {
          const bool last = blk == p.nblk - 1;
          // ---- sweep A: row max and best foreground column (>= 1; first maximum wins) over my columns:
          // compare, max, select of the in-chunk index per element; the chunk is noted once per chunk
          // (tcgen05.ld is warp-collective with one address: a per-lane reload of "my best chunk" is impossible)
          float best = -INFINITY, bg = -INFINITY;
          int besti = 0, bestchunk = 0;
          for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
            tmem_ld16(taddr + c0, v);
            const int c = col0 + c0;
            if (c + 16 > p.N) {  // the row's last chunk: padded columns do not count
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c + i >= p.N) v[i] = -INFINITY;
            }
            if (c == 0) {  // column 0 is the background row: in the row max, not among the labels
              bg = v[0];
              v[0] = -INFINITY;
            }
            const float before = best;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              besti = v[i] > best ? i : besti;
              best = fmaxf(best, v[i]);
            }
            bestchunk = best > before ? c : bestchunk;
          }
          const int bestc = bestchunk + besti;
          stats->best[blk][ch][trow] = best;
          stats->bestc[blk][ch][trow] = (uint16_t)bestc;
          if (ch == 0) stats->bg[blk][trow] = bg;
          bar_named(1 + quad, 32 * kCH);  // the two column halves of this quadrant
          float mxb = stats->bg[blk][trow];
#pragma unroll
          for (int h = 0; h < kCH; ++h) mxb = fmaxf(mxb, stats->best[blk][h][trow]);
          // ---- sweep B: e = exp(l - block max) cut to fp16 precision, row sum of the CUT values (so the
          // probabilities still add up to one to fp32 accuracy); raw logits out; e parked in the stash, or --
          // for the row's last block, whose turn at the stash comes later -- written back to its TMEM columns
          float sum = 0.f;
          const float moff = mxb * kLog2e;
          for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
            tmem_ld16(taddr + c0, v);
            const int c = col0 + c0;
            if (p.logits) {
              // a thread holds 16 columns of ITS row; transposed through a warp-private tile so that half a
              // warp writes 64 contiguous bytes of one row (2 rows per store instruction)
              float* tt = xpose + (warp - kFirstEpiWarp) * (32 * 17);
#pragma unroll
              for (int i = 0; i < 16; ++i) tt[lane * 17 + i] = v[i];
              __syncwarp();
              const int rsub = lane >> 4, csub = lane & 15;
#pragma unroll 4
              for (int rr = 0; rr < 32; rr += 2) {
                const long long orow = mt * kBM + quad * 32 + rr + rsub;
                if (orow < p.M && c + csub < p.N) p.logits[orow * p.ld + c + csub] = tt[(rr + rsub) * 17 + csub];
              }
              __syncwarp();
            }
            // e truncated to 11 significant bits (one LOP): exactly representable in fp16, so the parked
            // values, the TMEM copy and the row sum all see the same numbers
            if (c + 16 > p.N) {  // the row's last chunk: padded columns contribute nothing
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c + i >= p.N) v[i] = -INFINITY;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float e = __uint_as_float(__float_as_uint(ex2_approx(fmaf(v[i], kLog2e, -moff))) & 0xFFFFE000u);
              sum += e;
              v[i] = e;
            }
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __half2 t = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
              w[i] = *reinterpret_cast<const uint32_t*>(&t);
            }
            if (last && p.nblk > 1) {
              tmem_st16(taddr + c0, v);
            } else {
              uint4* dst = reinterpret_cast<uint4*>(stash + (size_t)trow * kStashPitch + 2 * c0);
              dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
              dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
            }
          }
          stats->sum[blk][ch][trow] = sum;
          if (!last || p.nblk == 1) {
            // this block's TMEM columns are no longer needed: the MMAs of the unit after next may overwrite them
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            if (!last) continue;
          } else {
            tmem_wait_st();
          }
          bar_named(1 + quad, 32 * kCH);  // sums and stash rows of this quadrant are visible
          // ---- the row's statistics over all blocks ----
          float bmax[2] = {-INFINITY, -INFINITY}, bsum[2] = {0.f, 0.f};
          for (int b = 0; b < p.nblk; ++b) {
            bmax[b] = stats->bg[b][trow];
#pragma unroll
            for (int h = 0; h < kCH; ++h) {
              bmax[b] = fmaxf(bmax[b], stats->best[b][h][trow]);
              bsum[b] += stats->sum[b][h][trow];
            }
          }
          const float rmx = fmaxf(bmax[0], bmax[1]);
          float tot = 0.f, fb[2] = {0.f, 0.f};
          for (int b = 0; b < p.nblk; ++b) {
            fb[b] = ex2_approx((bmax[b] - rmx) * kLog2e);
            tot += fb[b] * bsum[b];
          }
          const float inv = 1.0f / tot;
          if (ch == 0) {
            for (int b = 0; b < p.nblk; ++b) stats->fac[b][trow] = fb[b] * inv;
            if (row_ok && p.top_label) {
              float bb = -INFINITY;
              int bc = 0;
              for (int b = 0; b < p.nblk; ++b)
                for (int h = 0; h < kCH; ++h)
                  if (stats->best[b][h][trow] > bb) {  // blocks / halves in column order: the first maximum wins
                    bb = stats->best[b][h][trow];
                    bc = stats->bestc[b][h][trow];
                  }
              const float bp = p.N > 1 ? ex2_approx((bb - rmx) * kLog2e) * inv : 0.f;
              p.top_label[row] = (p.N > 1 && bp > p.score_thresh) ? bc : 0;
              if (p.top_prob) p.top_prob[row] = bp;
            }
          }
          bar_named(1 + quad, 32 * kCH);  // fac[] written
          // ---- probabilities: e * fac, streamed from the stash a warp per row, lanes along the columns
          // (every store instruction writes 128 contiguous bytes of one row, whatever the row pitch) ----
          // (a lane takes the column pairs (2l, 2l + 1) + 64k: one 32-bit shared load, two conversions, two
          // multiplies, two stores per pair -- a row pitch like 501 floats leaves no wider aligned store)
          const uint32_t stash_a = smem_u32(stash);
          auto stream = [&](int b, int ncol) {
            if (!p.probs) return;
            for (int rr = ch; rr < 32; rr += kCH) {
              const int r = quad * 32 + rr;
              const long long orow = mt * kBM + r;
              if (orow >= p.M) break;
              const float f = stats->fac[b][r];
              const uint32_t src = stash_a + (uint32_t)r * kStashPitch + 4u * (uint32_t)lane;
              float* o = p.probs + orow * p.ld + b * p.BN + 2 * lane;
#pragma unroll
              for (int k = 0; k < kMaxBN / 64; ++k) {
                const int c = 2 * lane + 64 * k;
                if (c < ncol) {
                  uint32_t w;
                  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(src + 128u * k));
                  const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&w));
                  o[64 * k] = e.x * f;
                  if (c + 1 < ncol) o[64 * k + 1] = e.y * f;
                }
              }
            }
          };
          if (p.nblk == 1) {
            stream(0, p.N);
          } else {
            stream(0, p.BN);
            bar_named(1 + quad, 32 * kCH);  // the parked block has left the stash
            // the last block: TMEM -> stash (exact: the values are fp16 already), then the same streaming
            if (p.probs) {
              for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
                tmem_ld16(taddr + c0, v);
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const __half2 t = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                  w[i] = *reinterpret_cast<const uint32_t*>(&t);
                }
                uint4* dst = reinterpret_cast<uint4*>(stash + (size_t)trow * kStashPitch + 2 * c0);
                dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            bar_named(1 + quad, 32 * kCH);
            stream(blk, p.N - col0);
          }
          // the statistics and stash rows of this quadrant may be overwritten by the next tile
          bar_named(1 + quad, 32 * kCH);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows, K] bf16 row-major -> box [box_rows x 64] with 128-byte swizzle, zero fill out of bounds
int make_map(CUtensorMap* map, const void* base, long long rows, int K, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("tc_gemm: cuTensorMapEncodeTiled entry point unavailable");
    return B200_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tc_gemm: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

template <int EPI>
int launch(GemmParams& p, const void* A, const void* B, size_t extra_smem, cudaStream_t st) {
  const int NP = (p.N + 15) / 16 * 16;
  p.BN = NP < kMaxBN ? NP : kMaxBN;
  p.nblk = (NP + p.BN - 1) / p.BN;
  p.buf_stride = 32;
  while (p.buf_stride < p.BN) p.buf_stride <<= 1;
  p.tmem_cols = 2 * p.buf_stride;
  p.a_bytes = kBM * kBK * 2;
  p.stage_bytes = p.a_bytes + (uint32_t)p.BN * kBK * 2;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const size_t budget = (size_t)224 * 1024 - 1024 - extra_smem;
  int stages = (int)(budget / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) {
    set_error("tc_gemm: shared memory budget leaves %d pipeline stages", stages);
    return B200_ERR_UNSUPPORTED;
  }
  (void)num_kb;
  p.stages = stages;
  const size_t smem = (size_t)stages * p.stage_bytes + 1024 + extra_smem;
  CUtensorMap map_a, map_b;
  int rc = make_map(&map_a, A, p.M, p.K, kBM);
  if (rc != B200_OK) return rc;
  rc = make_map(&map_b, B, p.N, p.K, p.BN);
  if (rc != B200_OK) return rc;
  static SmemHighWater hw;
  rc = ensure_dynamic_smem(tc_gemm_kernel<EPI>, smem, &hw, "tc_gemm: smem attribute");
  if (rc != B200_OK) return rc;
  const long long n_units = (p.M + kBM - 1) / kBM * (EPI == kEpiLinear ? p.nblk : 1);
  const long long grid = n_units < sm_count() ? n_units : sm_count();
  tc_gemm_kernel<EPI><<<(unsigned)grid, kGemmThreads, smem, st>>>(map_a, map_b, p);
  B200_CHECK_LAUNCH("tc_gemm_kernel");
  return B200_OK;
}

}  // namespace

bool softmax_gemm_applies(int n_cols) { return n_cols >= 1 && n_cols <= 2 * kMaxBN; }

int softmax_gemm_launch(const void* A_bf16, const void* E_bf16, int64_t n_rows, int n_cols, int dim, int ld,
                        float score_thresh, float* probs, float* logits, int32_t* top_label, float* top_prob,
                        cudaStream_t st) {
  GemmParams p = {};
  p.M = n_rows;
  p.N = n_cols;
  p.K = dim;
  p.score_thresh = score_thresh;
  p.probs = probs;
  p.logits = logits;
  p.top_label = top_label;
  p.top_prob = top_prob;
  p.ld = ld;
  const size_t extra = sizeof(SoftStats) + 16 + (size_t)kBM * kStashPitch + (logits ? sizeof(float) * kEpiWarps * 32 * 17 : 0);
  return launch<kEpiSoftmax>(p, A_bf16, E_bf16, extra, st);
}

int softmax_wide_launch(const void* A_bf16, const void* E_bf16, int64_t n_rows, int n_cols, int dim, float score_thresh,
                        float* probs, float* logits, int32_t* top_label, float* top_prob, cudaStream_t st) {
  GemmParams p = {};
  p.M = n_rows;
  p.N = n_cols;
  p.K = dim;
  p.score_thresh = score_thresh;
  p.probs = probs;
  p.logits = logits;
  p.top_label = top_label;
  p.top_prob = top_prob;
  p.ld = n_cols;
  p.passes = probs ? 2 : 1;
  const size_t extra = sizeof(SoftStats) + 16 + sizeof(float) * kEpiWarps * 32 * 17;
  return launch<kEpiSoftmaxWide>(p, A_bf16, E_bf16, extra, st);
}

}  // namespace b200

extern "C" int b200_linear_bf16(const void* A_bf16, const void* W_bf16, const float* bias, int64_t n_rows, int n_out,
                                int dim, void* out_bf16, float* out_f32, void* stream) {
  using namespace b200;
  B200_REQUIRE(n_rows >= 0 && n_out >= 0 && dim > 0, "linear_bf16: bad shape");
  if (n_rows == 0 || n_out == 0) return B200_OK;
  B200_REQUIRE(A_bf16 && W_bf16 && (out_bf16 || out_f32), "linear_bf16: null operand / no output");
  B200_REQUIRE(aligned16(A_bf16) && aligned16(W_bf16), "linear_bf16: operands must be 16-byte aligned");
  B200_REQUIRE(dim % 8 == 0, "linear_bf16: dim must be a multiple of 8 (16-byte row pitch for TMA)");
  B200_REQUIRE(n_rows < ((int64_t)1 << 31), "linear_bf16: n_rows must fit int32 TMA coordinates");
  GemmParams p = {};
  p.M = n_rows;
  p.N = n_out;
  p.K = dim;
  p.bias = bias;
  p.out_bf16 = static_cast<__nv_bfloat16*>(out_bf16);
  p.out_f32 = out_f32;
  p.ldo = n_out;
  return launch<kEpiLinear>(p, A_bf16, W_bf16, 0, static_cast<cudaStream_t>(stream));
}
