// embed_match.cu -- region -> class-embedding scoring on the sm_100a tensor cores.
//
//   logits[M, N] = A[M, K] . E[N, K]^T      bf16 operands, fp32 accumulation in TMEM
//
// replaces (reference, paths under maskrcnn_benchmark/)
//   modeling/roi_heads/box_head/roi_box_predictors.py:67  einsum('pe,ce->pc', cls_emb, cls_score)
//   modeling/roi_heads/box_head/inference.py:62           F.softmax(class_logits, -1)
//   modeling/detector/st_generalized_rcnn.py:245-255      einsum('pd,wd->pw'), max over regions, sigmoid
// with the consumer fused into the epilogue, so the [M, N] scores never make an extra
// HBM round trip (at N = 66 the op is bound by reading A once).
//
// One CTA per 128-row tile, all N (<= 512) columns:
//   warp 0      TMA producer: cp.async.bulk.tensor 2D boxes of A [128 x 64] and E [N x 64]
//               (128-byte swizzle) into a ring of shared-memory stages, mbarrier completion.
//   warp 1      allocates TMEM and issues tcgen05.mma (M=128, N<=256 per instruction, K=16),
//               tcgen05.commit releases each stage and finally signals the epilogue.
//   warps 2-5   epilogue: thread = one row of the tile = one TMEM lane; tcgen05.ld 16 columns
//               at a time; row softmax + top-1 foreground label (SOFTMAX mode) or per-column
//               max over the rows of the column's image via 64-bit atomicMax (COLMAX mode).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace b200 {
namespace {

constexpr int kBM = 128;       // rows per CTA (= TMEM lanes)
constexpr int kBK = 64;        // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;     // K of one tcgen05.mma.kind::f16
constexpr int kThreads = 192;  // 6 warps: TMA, MMA, 4 epilogue
constexpr int kMaxStages = 8;

struct MatchParams {
  long long M;
  int N, K;
  int NP;        // N rounded up to 16
  int n_halves;  // 1, or 2 when NP > 256
  int b_rows;    // rows of the E tile in shared memory (NP, or 512 when two boxes are loaded)
  int stages;
  int tmem_cols;
  int ld;        // row pitch (floats) of the probs / logits outputs (>= N; column blocks of a wider matrix)
  int mode;
  float score_thresh;
  float* probs;
  float* logits;
  int32_t* top_label;
  float* top_prob;
  const int32_t* row_seg;
  const int32_t* col_seg;
  const int32_t* row_seg_start;
  unsigned long long* col_best;
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// K-major operand tile with 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset = 1024 B between 8-row groups
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                        // layout type: SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
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

__device__ __forceinline__ uint32_t orderable(float f) {
  uint32_t u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

__global__ void __launch_bounds__(kThreads)
embed_match_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_e,
                   const MatchParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float stage_tile[4 * 32 * 17];  // epilogue transpose tiles, one per epilogue warp

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * kBM;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const uint32_t a_bytes = kBM * kBK * 2;
  const uint32_t b_bytes = (uint32_t)p.b_rows * kBK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  // dynamic shared memory may start at any 16-byte boundary: realign for the 128 B swizzle
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % p.stages;
        const uint32_t ph = (kb / p.stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);  // first pass falls through (barrier still in phase 0)
        unsigned char* a_dst = tiles + (size_t)s * stage_bytes;
        unsigned char* b_dst = a_dst + a_bytes;
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        tma_load_2d(a_dst, &map_a, &full_bar[s], kb * kBK, (int)row0);
        tma_load_2d(b_dst, &map_e, &full_bar[s], kb * kBK, 0);
        if (p.n_halves == 2) tma_load_2d(b_dst + 256 * kBK * 2, &map_e, &full_bar[s], kb * kBK, 256);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane) =====
    if (lane == 0) {
      const int n0 = p.n_halves == 2 ? 256 : p.NP;
      const int n1 = p.NP - 256;
      const uint32_t idesc0 = umma_idesc(n0), idesc1 = umma_idesc(n1 > 0 ? n1 : 16);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % p.stages;
        const uint32_t ph = (kb / p.stages) & 1;
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_u32(tiles + (size_t)s * stage_bytes);
        const uint32_t b_addr = a_addr + a_bytes;
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k) {
          const uint32_t acc = (kb | k) ? 1u : 0u;
          const uint64_t da = umma_desc_sw128(a_addr + k * kUmmaK * 2);
          umma_bf16(tmem_base, da, umma_desc_sw128(b_addr + k * kUmmaK * 2), idesc0, acc);
          if (p.n_halves == 2)
            umma_bf16(tmem_base + 256, da, umma_desc_sw128(b_addr + 256 * kBK * 2 + k * kUmmaK * 2), idesc1, acc);
        }
        umma_commit(&empty_bar[s]);  // stage free once these MMAs have read it
      }
      umma_commit(&tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===== epilogue: thread <-> TMEM lane <-> row of the tile =====
    const int quad = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int trow = quad * 32 + lane;
    const long long row = row0 + trow;
    const bool row_ok = row < p.M;
    mbar_wait(&tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int N = p.N;
    float v[16];

    if (p.mode == B200_MATCH_SOFTMAX) {
      // pass 1 (online softmax): running row max, rescaled exp-sum, best foreground column (>= 1)
      float mx = -INFINITY, best = -INFINITY, sum = 0.f;
      int best_c = 0;
      for (int c0 = 0; c0 < N; c0 += 16) {
        tmem_ld16(taddr + c0, v);
        float cmx = mx;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = c0 + i;
          if (c < N) {
            cmx = fmaxf(cmx, v[i]);
            if (c >= 1 && v[i] > best) {
              best = v[i];
              best_c = c;
            }
          }
        }
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < N) part += __expf(v[i] - cmx);
        sum = sum * __expf(mx - cmx) + part;  // exp(-inf) = 0 on the first chunk
        mx = cmx;
      }
      const float inv = 1.0f / sum;
      // pass 2: outputs.  A thread holds 16 consecutive columns of ITS row; written directly,
      // a warp store would touch 32 rows x 4 bytes (32 sectors).  The 32x16 chunk is transposed
      // through a warp-private shared tile so that half a warp writes 64 contiguous bytes of one
      // row: 2 rows per store instruction, ~5 sectors instead of 32.
      if (p.probs || p.logits) {
        float* tile = stage_tile + (warp - 2) * (32 * 17);
        const int rsub = lane >> 4, csub = lane & 15;
        for (int which = 0; which < 2; ++which) {
          float* outp = which == 0 ? p.logits : p.probs;
          if (!outp) continue;
          for (int c0 = 0; c0 < N; c0 += 16) {
            tmem_ld16(taddr + c0, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) tile[lane * 17 + i] = which == 0 ? v[i] : __expf(v[i] - mx) * inv;
            __syncwarp();
            const int c = c0 + csub;
#pragma unroll 4
            for (int rr = 0; rr < 32; rr += 2) {
              const long long orow = row0 + quad * 32 + rr + rsub;
              if (orow < p.M && c < N) outp[orow * p.ld + c] = tile[(rr + rsub) * 17 + csub];
            }
            __syncwarp();
          }
        }
      }
      if (row_ok && p.top_label) {
        const float bp = N > 1 ? __expf(best - mx) * inv : 0.f;
        p.top_label[row] = (N > 1 && bp > p.score_thresh) ? best_c : 0;
        if (p.top_prob) p.top_prob[row] = bp;
      }
    } else {
      // COLMAX: per column, max over the rows of the column's image; ties -> first row
      const int seg = row_ok ? p.row_seg[row] : -1;
      const int lrow = row_ok ? (int)(row - p.row_seg_start[seg]) : 0;
      const int seg_lo = __reduce_min_sync(0xffffffffu, row_ok ? seg : 0x7fffffff);
      const int seg_hi = __reduce_max_sync(0xffffffffu, seg);
      for (int c0 = 0; c0 < N; c0 += 16) {
        tmem_ld16(taddr + c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = c0 + i;
          if (c >= N) break;
          if (p.logits && row_ok) p.logits[row * p.ld + c] = v[i];
          const int cs = p.col_seg[c];
          if (cs < seg_lo || cs > seg_hi) continue;  // warp-uniform: no row of this warp can match
          unsigned long long key = 0;
          if (seg == cs) key = ((unsigned long long)orderable(v[i]) << 32) | (0xFFFFFFFFu - (uint32_t)lrow);
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, d);
            key = o > key ? o : key;
          }
          if (lane == 0 && key) atomicMax(p.col_best + c, key);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

__global__ void colmax_decode_kernel(const unsigned long long* __restrict__ col_best, int n,
                                     int32_t* __restrict__ row_idx, float* __restrict__ max_score,
                                     float* __restrict__ sig) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const unsigned long long key = col_best[c];
  if (key == 0) {  // no row of the column's image was seen
    if (row_idx) row_idx[c] = -1;
    if (max_score) max_score[c] = -INFINITY;
    if (sig) sig[c] = 0.f;
    return;
  }
  uint32_t u = (uint32_t)(key >> 32);
  u ^= (u >> 31) ? 0x80000000u : 0xFFFFFFFFu;
  const float s = __uint_as_float(u);
  if (row_idx) row_idx[c] = (int32_t)(0xFFFFFFFFu - (uint32_t)key);
  if (max_score) max_score[c] = s;
  if (sig) sig[c] = 1.0f / (1.0f + expf(-s));  // torch.sigmoid (st_generalized_rcnn.py:255)
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
    set_error("embed_match: cuTensorMapEncodeTiled entry point unavailable");
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
    set_error("embed_match: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

}  // namespace
}  // namespace b200

namespace b200 {
namespace {

// Row softmax over already computed logits [M, N] (row pitch N): what the fused epilogue does for
// N <= 512, as a second pass for wider class matrices (b200_embed_match_wide).  One warp per row;
// same arithmetic as the epilogue (__expf, first maximum wins ties).
__global__ void __launch_bounds__(256) row_softmax_kernel(const float* __restrict__ logits, long long M, int N,
                                                          float score_thresh, float* __restrict__ probs,
                                                          int32_t* __restrict__ top_label,
                                                          float* __restrict__ top_prob) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* x = logits + row * N;
  float mx = -INFINITY, best = -INFINITY;
  int best_c = 0x7fffffff;
  for (int c = lane; c < N; c += 32) {
    const float v = x[c];
    mx = fmaxf(mx, v);
    if (c >= 1 && v > best) {
      best = v;
      best_c = c;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    const float ob = __shfl_xor_sync(0xffffffffu, best, d);
    const int oc = __shfl_xor_sync(0xffffffffu, best_c, d);
    if (ob > best || (ob == best && oc < best_c)) {
      best = ob;
      best_c = oc;
    }
  }
  float sum = 0.f;
  for (int c = lane; c < N; c += 32) sum += __expf(x[c] - mx);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  const float inv = 1.0f / sum;
  if (probs)
    for (int c = lane; c < N; c += 32) probs[row * N + c] = __expf(x[c] - mx) * inv;
  if (lane == 0 && top_label) {
    const float bp = N > 1 ? __expf(best - mx) * inv : 0.f;
    top_label[row] = (N > 1 && bp > score_thresh) ? best_c : 0;
    if (top_prob) top_prob[row] = bp;
  }
}

int embed_match_launch(const void* A_bf16, const void* E_bf16, int64_t n_rows, int n_cols, int dim, int ld, int mode,
                       float score_thresh, float* probs, float* logits, int32_t* top_label, float* top_prob,
                       const int32_t* row_seg, const int32_t* col_seg, const int32_t* row_seg_start,
                       uint64_t* col_best, void* stream);

bool g_match_legacy = false;

}  // namespace
}  // namespace b200

// Test hook (not part of the reference-facing ABI): SOFTMAX through the one-CTA-per-tile kernel.
extern "C" void b200_debug_match(int legacy) { b200::g_match_legacy = legacy != 0; }

extern "C" int b200_embed_match(const void* A_bf16, const void* E_bf16, int64_t n_rows, int n_cols, int dim,
                                int mode, float score_thresh, float* probs, float* logits, int32_t* top_label,
                                float* top_prob, const int32_t* row_seg, const int32_t* col_seg,
                                const int32_t* row_seg_start, uint64_t* col_best, void* stream) {
  return b200::embed_match_launch(A_bf16, E_bf16, n_rows, n_cols, dim, n_cols, mode, score_thresh, probs, logits,
                                  top_label, top_prob, row_seg, col_seg, row_seg_start, col_best, stream);
}

extern "C" int b200_embed_match_wide(const void* A_bf16, const void* E_bf16, int64_t n_rows, int n_cols, int dim,
                                     float score_thresh, float* probs, float* logits, int32_t* top_label,
                                     float* top_prob, void* stream) {
  using namespace b200;
  B200_REQUIRE(n_rows >= 0 && n_cols >= 0 && dim > 0, "embed_match_wide: bad shape");
  if (n_rows == 0 || n_cols == 0) return B200_OK;
  B200_REQUIRE(A_bf16 && E_bf16 && dim % 8 == 0, "embed_match_wide: null operand or dim %% 8 != 0");
  B200_REQUIRE(aligned16(A_bf16) && aligned16(E_bf16), "embed_match_wide: operands must be 16-byte aligned");
  B200_REQUIRE(n_rows < ((int64_t)1 << 31), "embed_match_wide: n_rows must fit int32 TMA coordinates");
  if (!g_match_legacy) {
    // one persistent launch: statistics pass + probability pass over recomputed blocks, the logits make no
    // HBM round trip (tc_gemm.cu, SOFTMAX_WIDE); `logits` is an optional output, no longer scratch
    if (!probs && !logits && !top_label) return B200_OK;
    return softmax_wide_launch(A_bf16, E_bf16, n_rows, n_cols, dim, score_thresh, probs, logits, top_label, top_prob,
                               static_cast<cudaStream_t>(stream));
  }
  // legacy (test hook b200_debug_match): logits by column blocks of <= 512 written at the full row pitch, then
  // a row-softmax pass over them
  B200_REQUIRE(logits, "embed_match_wide (legacy path): the [n_rows, n_cols] logits buffer is required (output and scratch)");
  for (int c0 = 0; c0 < n_cols; c0 += 512) {
    const int nb = n_cols - c0 < 512 ? n_cols - c0 : 512;
    const char* eb = static_cast<const char*>(E_bf16) + (size_t)c0 * dim * 2;
    int rc = embed_match_launch(A_bf16, eb, n_rows, nb, dim, n_cols, B200_MATCH_SOFTMAX, score_thresh, nullptr,
                                logits + c0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
    if (rc != B200_OK) return rc;
  }
  if (probs || top_label) {
    const long long blocks = (n_rows + 7) / 8;
    row_softmax_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, n_rows, n_cols, score_thresh, probs, top_label, top_prob);
    B200_CHECK_LAUNCH("row_softmax_kernel");
  }
  return B200_OK;
}

namespace b200 {
namespace {
int embed_match_launch(const void* A_bf16, const void* E_bf16, int64_t n_rows, int n_cols, int dim, int ld, int mode,
                       float score_thresh, float* probs, float* logits, int32_t* top_label, float* top_prob,
                       const int32_t* row_seg, const int32_t* col_seg, const int32_t* row_seg_start,
                       uint64_t* col_best, void* stream) {
  B200_REQUIRE(mode == B200_MATCH_SOFTMAX || mode == B200_MATCH_COLMAX, "embed_match: bad mode %d", mode);
  B200_REQUIRE(n_rows >= 0 && n_cols >= 0 && dim > 0, "embed_match: bad shape");
  if (n_rows == 0 || n_cols == 0) return B200_OK;
  B200_REQUIRE(A_bf16 && E_bf16, "embed_match: null operand");
  B200_REQUIRE(aligned16(A_bf16) && aligned16(E_bf16), "embed_match: operands must be 16-byte aligned");
  B200_REQUIRE(dim % 8 == 0, "embed_match: dim must be a multiple of 8 (16-byte row pitch for TMA)");
  if (n_cols > 512) {
    set_error("embed_match: n_cols %d > 512 (one TMEM allocation); split the columns", n_cols);
    return B200_ERR_UNSUPPORTED;
  }
  B200_REQUIRE(n_rows < ((int64_t)1 << 31), "embed_match: n_rows must fit int32 TMA coordinates");
  if (mode == B200_MATCH_COLMAX)
    B200_REQUIRE(row_seg && col_seg && row_seg_start && col_best, "embed_match: COLMAX needs row_seg/col_seg/"
                                                                  "row_seg_start/col_best");
  // SOFTMAX: the persistent kernel with overlapped epilogue (tc_gemm.cu); g_match_legacy (test hook) keeps
  // the one-CTA-per-tile kernel below reachable, which also serves COLMAX
  if (mode == B200_MATCH_SOFTMAX && !g_match_legacy && softmax_gemm_applies(n_cols))
    return softmax_gemm_launch(A_bf16, E_bf16, n_rows, n_cols, dim, ld, score_thresh, probs, logits, top_label, top_prob,
                               static_cast<cudaStream_t>(stream));
  MatchParams p;
  p.M = n_rows;
  p.N = n_cols;
  p.K = dim;
  p.NP = (n_cols + 15) / 16 * 16;
  p.n_halves = p.NP > 256 ? 2 : 1;
  p.b_rows = p.n_halves == 2 ? 512 : p.NP;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.NP) p.tmem_cols <<= 1;
  p.ld = ld;
  p.mode = mode;
  p.score_thresh = score_thresh;
  p.probs = probs;
  p.logits = logits;
  p.top_label = top_label;
  p.top_prob = top_prob;
  p.row_seg = row_seg;
  p.col_seg = col_seg;
  p.row_seg_start = row_seg_start;
  p.col_best = reinterpret_cast<unsigned long long*>(col_best);
  const size_t stage_bytes = (size_t)kBM * kBK * 2 + (size_t)p.b_rows * kBK * 2;
  const int num_kb = (dim + kBK - 1) / kBK;
  // ring depth: ~52 KB per CTA when the E tile is small (N <= 128: four CTAs and their 128
  // TMEM columns share an SM, so one tile's epilogue overlaps the others' loads), else ~100 KB
  int stages = (int)(((p.tmem_cols <= 128 ? 52 : 100) * 1024) / stage_bytes);
  if (stages < 2) stages = 2;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > num_kb) stages = num_kb < 1 ? 1 : num_kb;
  p.stages = stages;
  const size_t smem = stage_bytes * stages + 1024;

  CUtensorMap map_a, map_e;
  int rc = make_map(&map_a, A_bf16, n_rows, dim, kBM);
  if (rc != B200_OK) return rc;
  rc = make_map(&map_e, E_bf16, n_cols, dim, p.n_halves == 2 ? 256 : p.NP);
  if (rc != B200_OK) return rc;
  static SmemHighWater hw;
  rc = ensure_dynamic_smem(embed_match_kernel, smem, &hw, "embed_match: smem attribute");
  if (rc != B200_OK) return rc;
  const long long grid = (n_rows + kBM - 1) / kBM;
  embed_match_kernel<<<(unsigned)grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(map_a, map_e, p);
  B200_CHECK_LAUNCH("embed_match_kernel");
  return B200_OK;
}
}  // namespace
}  // namespace b200

extern "C" int b200_colmax_decode(const uint64_t* col_best, int n_cols, int32_t* row_idx, float* max_score,
                                  float* sigmoid_score, void* stream) {
  using namespace b200;
  B200_REQUIRE(n_cols >= 0, "colmax_decode: bad n_cols");
  if (n_cols == 0) return B200_OK;
  B200_REQUIRE(col_best, "colmax_decode: null col_best");
  colmax_decode_kernel<<<(n_cols + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const unsigned long long*>(col_best), n_cols, row_idx, max_score, sigmoid_score);
  B200_CHECK_LAUNCH("colmax_decode_kernel");
  return B200_OK;
}
