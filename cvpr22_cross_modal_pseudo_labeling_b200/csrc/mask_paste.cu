// mask_paste.cu -- Masker.paste_mask_in_image for all boxes of a batch in one launch.
//
// Replaces the per-box Python loop of Masker.forward_single_image (reference
// modeling/roi_heads/mask_head/inference.py:124-186): expand_masks (zero border of `padding`
// pixels, :114-122), expand_boxes (same relative growth, :96-111), truncation of the box to
// int32, bilinear resize of the (M+2p)^2 mask to the box size (F.interpolate, align_corners =
// False), `> thresh`, and the paste into a zero image.  One thread produces 16 consecutive
// pixels of one image row of one box's full-size mask and stores them as one 128-bit word.
#include "common.cuh"

namespace b200 {
namespace {

// at::native area_pixel_compute_source_index (align_corners = false, not cubic) followed by
// guard_index_and_lambda: source index and weight of output pixel `dst` along one axis
__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  if (s < 0.f) s = 0.f;
  i0 = min((int)floorf(s), in_size - 1);
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = fminf(fmaxf(__fsub_rn(s, (float)i0), 0.f), 1.f);
  l0 = __fsub_rn(1.f, l1);
}

constexpr int kPasteThreads = 128;
constexpr int kPxPerThread = 16;
constexpr int kRowsPerCta = 8;  // amortises the box set-up over several image rows

__global__ void __launch_bounds__(kPasteThreads)
paste_masks_kernel(const float* __restrict__ masks, const float* __restrict__ boxes, int M, int padding, int im_h,
                   int im_w, int row_words, float thresh, uint8_t* __restrict__ out) {
  const int n = blockIdx.y;
  const float* bx = boxes + (size_t)n * 4;
  // expand_boxes (:96-111) with scale = (M + 2p) / M, then .to(int32): truncation toward zero
  const int Mp = M + 2 * padding;
  const float scale = (float)Mp / (float)M;
  const float w_half = __fmul_rn(__fmul_rn(__fsub_rn(bx[2], bx[0]), 0.5f), scale);
  const float h_half = __fmul_rn(__fmul_rn(__fsub_rn(bx[3], bx[1]), 0.5f), scale);
  const float x_c = __fmul_rn(__fadd_rn(bx[2], bx[0]), 0.5f), y_c = __fmul_rn(__fadd_rn(bx[3], bx[1]), 0.5f);
  const int b0 = (int)__fsub_rn(x_c, w_half), b2 = (int)__fadd_rn(x_c, w_half);
  const int b1 = (int)__fsub_rn(y_c, h_half), b3 = (int)__fadd_rn(y_c, h_half);
  const int w = max(b2 - b0 + 1, 1), h = max(b3 - b1 + 1, 1);
  const int x_0 = max(b0, 0), x_1 = min(b2 + 1, im_w), y_0 = max(b1, 0), y_1 = min(b3 + 1, im_h);
  const float sx = (float)Mp / (float)w, sy = (float)Mp / (float)h;
  const float* mk = masks + (size_t)n * M * M;
  // padded mask value: zero border of `padding` pixels around the M x M mask
  auto at = [&](int py, int px) -> float {
    const int my = py - padding, mx = px - padding;
    return (my >= 0 && my < M && mx >= 0 && mx < M) ? __ldg(mk + my * M + mx) : 0.f;
  };
  const int y_end = min((int)(blockIdx.x + 1) * kRowsPerCta, im_h);
  for (int y = blockIdx.x * kRowsPerCta; y < y_end; ++y) {
  const bool row_in = y >= y_0 && y < y_1;
  int yi0 = 0, yi1 = 0;
  float yl0 = 0.f, yl1 = 0.f;
  if (row_in) src_index(sy, y - b1, Mp, yi0, yi1, yl0, yl1);
  uint8_t* orow = out + ((size_t)n * im_h + y) * im_w;
  // 16-byte words are aligned to the ADDRESS, not to x = 0 (row pitch = im_w is arbitrary): the
  // first and last word of a row may be partial, every other one is a single 128-bit store
  const int mis = (int)(reinterpret_cast<uintptr_t>(orow) & 15);
  for (int wd = threadIdx.x; wd < row_words; wd += kPasteThreads) {
    const int xb = wd * kPxPerThread - mis;
    uint32_t packed[4] = {0u, 0u, 0u, 0u};
    if (row_in && xb < x_1 && xb + kPxPerThread > x_0) {
#pragma unroll
      for (int k = 0; k < kPxPerThread; ++k) {
        const int x = xb + k;
        if (x >= x_0 && x < x_1) {
          int xi0, xi1;
          float xl0, xl1;
          src_index(sx, x - b0, Mp, xi0, xi1, xl0, xl1);
          // at::native cpu_upsample_linear: h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11)
          const float top = __fadd_rn(__fmul_rn(xl0, at(yi0, xi0)), __fmul_rn(xl1, at(yi0, xi1)));
          const float bot = __fadd_rn(__fmul_rn(xl0, at(yi1, xi0)), __fmul_rn(xl1, at(yi1, xi1)));
          const float v = __fadd_rn(__fmul_rn(yl0, top), __fmul_rn(yl1, bot));
          if (v > thresh) packed[k >> 2] |= 1u << (8 * (k & 3));
        }
      }
    }
    if (xb >= 0 && xb + kPxPerThread <= im_w) {
      *reinterpret_cast<uint4*>(orow + xb) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    } else {
      for (int k = 0; k < kPxPerThread; ++k)
        if (xb + k >= 0 && xb + k < im_w) orow[xb + k] = (uint8_t)((packed[k >> 2] >> (8 * (k & 3))) & 1u);
    }
  }
  }
}

// ---------------------------------------------------------------------------------------------
// Mask targets of the student without the full-image masks ("paste o crop"):
// generate_pseudo_label pastes every pseudo-label's M0 x M0 mask into a full-image boolean mask
// (Masker, st_generalized_rcnn.py:267-272), and the mask loss then crops that image at each positive proposal
// and resizes the crop to M x M (project_masks_on_boxes, mask_head/loss.py:11-42: BinaryMaskList.crop with
// Python round(), F.interpolate bilinear align_corners = False, `.type_as(bool)` = "> 0").  A target pixel is
// a bilinear blend of four binary image pixels, each of which is a thresholded bilinear blend of four mask
// values: 16 taps, evaluated here directly from the M0 x M0 masks.  One thread per target pixel.
// ---------------------------------------------------------------------------------------------
struct PastedBox {
  int b0, b1, x_0, x_1, y_0, y_1;
  float sx, sy;
};

__device__ __forceinline__ PastedBox pasted_box(const float* bx, int M, int padding, int im_h, int im_w) {
  const int Mp = M + 2 * padding;
  const float scale = (float)Mp / (float)M;
  const float w_half = __fmul_rn(__fmul_rn(__fsub_rn(bx[2], bx[0]), 0.5f), scale);
  const float h_half = __fmul_rn(__fmul_rn(__fsub_rn(bx[3], bx[1]), 0.5f), scale);
  const float x_c = __fmul_rn(__fadd_rn(bx[2], bx[0]), 0.5f), y_c = __fmul_rn(__fadd_rn(bx[3], bx[1]), 0.5f);
  PastedBox p;
  p.b0 = (int)__fsub_rn(x_c, w_half);
  p.b1 = (int)__fsub_rn(y_c, h_half);
  const int b2 = (int)__fadd_rn(x_c, w_half), b3 = (int)__fadd_rn(y_c, h_half);
  const int w = max(b2 - p.b0 + 1, 1), h = max(b3 - p.b1 + 1, 1);
  p.x_0 = max(p.b0, 0);
  p.x_1 = min(b2 + 1, im_w);
  p.y_0 = max(p.b1, 0);
  p.y_1 = min(b3 + 1, im_h);
  p.sx = (float)Mp / (float)w;
  p.sy = (float)Mp / (float)h;
  return p;
}

// the pasted full-image mask of one label at image pixel (y, x): paste_masks_kernel's per-pixel expression
__device__ __forceinline__ bool pasted_pixel(const float* __restrict__ mk, const PastedBox& pb, int M, int padding, int y,
                                             int x, float thresh) {
  if (y < pb.y_0 || y >= pb.y_1 || x < pb.x_0 || x >= pb.x_1) return false;
  const int Mp = M + 2 * padding;
  auto at = [&](int py, int px) -> float {
    const int my = py - padding, mx = px - padding;
    return (my >= 0 && my < M && mx >= 0 && mx < M) ? __ldg(mk + my * M + mx) : 0.f;
  };
  int yi0, yi1, xi0, xi1;
  float yl0, yl1, xl0, xl1;
  src_index(pb.sy, y - pb.b1, Mp, yi0, yi1, yl0, yl1);
  src_index(pb.sx, x - pb.b0, Mp, xi0, xi1, xl0, xl1);
  const float top = __fadd_rn(__fmul_rn(xl0, at(yi0, xi0)), __fmul_rn(xl1, at(yi0, xi1)));
  const float bot = __fadd_rn(__fmul_rn(xl0, at(yi1, xi0)), __fmul_rn(xl1, at(yi1, xi1)));
  return __fadd_rn(__fmul_rn(yl0, top), __fmul_rn(yl1, bot)) > thresh;
}

__global__ void __launch_bounds__(256)
mask_targets_kernel(const float* __restrict__ masks, const float* __restrict__ label_boxes, const int32_t* __restrict__ match,
                    const float* __restrict__ proposals, long long n_props, int M0, int padding, int im_h, int im_w,
                    float thresh, int M, float* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_props * M * M) return;
  const long long p = e / (M * M);
  const int oy = (int)(e % (M * M)) / M, ox = (int)(e % M);
  const int k = match[p];
  if (k < 0) {  // no label matched: an all-zero target
    out[e] = 0.f;
    return;
  }
  // BinaryMaskList.crop (structures/segmentation_mask.py:118-137): Python round() = half to even
  const float* pr = proposals + p * 4;
  int xmin = __float2int_rn(pr[0]), ymin = __float2int_rn(pr[1]), xmax = __float2int_rn(pr[2]), ymax = __float2int_rn(pr[3]);
  xmin = min(max(xmin, 0), im_w - 1);
  ymin = min(max(ymin, 0), im_h - 1);
  xmax = max(min(max(xmax, 0), im_w), xmin + 1);
  ymax = max(min(max(ymax, 0), im_h), ymin + 1);
  const int cw = xmax - xmin, ch = ymax - ymin;
  // BinaryMaskList.resize (:139-158): bilinear, align_corners = False, scale = input / output
  int yi0, yi1, xi0, xi1;
  float yl0, yl1, xl0, xl1;
  src_index((float)ch / (float)M, oy, ch, yi0, yi1, yl0, yl1);
  src_index((float)cw / (float)M, ox, cw, xi0, xi1, xl0, xl1);
  const float* mk = masks + (size_t)k * M0 * M0;
  const PastedBox pb = pasted_box(label_boxes + (size_t)k * 4, M0, padding, im_h, im_w);
  const float v00 = pasted_pixel(mk, pb, M0, padding, ymin + yi0, xmin + xi0, thresh) ? 1.f : 0.f;
  const float v01 = pasted_pixel(mk, pb, M0, padding, ymin + yi0, xmin + xi1, thresh) ? 1.f : 0.f;
  const float v10 = pasted_pixel(mk, pb, M0, padding, ymin + yi1, xmin + xi0, thresh) ? 1.f : 0.f;
  const float v11 = pasted_pixel(mk, pb, M0, padding, ymin + yi1, xmin + xi1, thresh) ? 1.f : 0.f;
  const float top = __fadd_rn(__fmul_rn(xl0, v00), __fmul_rn(xl1, v01));
  const float bot = __fadd_rn(__fmul_rn(xl0, v10), __fmul_rn(xl1, v11));
  // `.type_as(bool)` of the interpolated value: any non-zero blend is True
  out[e] = __fadd_rn(__fmul_rn(yl0, top), __fmul_rn(yl1, bot)) != 0.f ? 1.f : 0.f;
}

}  // namespace
}  // namespace b200

extern "C" int b200_mask_targets(const float* masks, const float* label_boxes, const int32_t* match, const float* proposals,
                                 int64_t n_proposals, int mask_size, int padding, int im_h, int im_w, float thresh,
                                 int target_size, float* out, void* stream) {
  using namespace b200;
  B200_REQUIRE(n_proposals >= 0 && mask_size > 0 && padding >= 0 && im_h > 0 && im_w > 0 && target_size > 0,
               "mask_targets: bad shape");
  B200_REQUIRE(thresh >= 0.f, "mask_targets: thresh must be >= 0");
  if (n_proposals == 0) return B200_OK;
  B200_REQUIRE(masks && label_boxes && match && proposals && out, "mask_targets: null pointer");
  const long long total = (long long)n_proposals * target_size * target_size;
  B200_REQUIRE(total < ((long long)1 << 38), "mask_targets: too many target pixels for one launch");
  mask_targets_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      masks, label_boxes, match, proposals, (long long)n_proposals, mask_size, padding, im_h, im_w, thresh, target_size, out);
  B200_CHECK_LAUNCH("mask_targets_kernel");
  return B200_OK;
}

extern "C" int b200_paste_masks(const float* masks, const float* boxes, int64_t n_boxes, int mask_size, int padding,
                                int im_h, int im_w, float thresh, uint8_t* out, void* stream) {
  using namespace b200;
  B200_REQUIRE(n_boxes >= 0 && mask_size > 0 && padding >= 0 && im_h > 0 && im_w > 0, "paste_masks: bad shape");
  B200_REQUIRE(thresh >= 0.f, "paste_masks: thresh must be >= 0 (the un-thresholded debug mode is not provided)");
  if (n_boxes == 0) return B200_OK;
  B200_REQUIRE(masks && boxes && out, "paste_masks: null pointer");
  B200_REQUIRE(n_boxes <= 65535 && im_h <= 2147483647, "paste_masks: too many boxes for one launch");
  const int row_words = (im_w + kPxPerThread - 1) / kPxPerThread + 1;  // + 1: rows need not start 16-byte aligned
  paste_masks_kernel<<<dim3((unsigned)((im_h + kRowsPerCta - 1) / kRowsPerCta), (unsigned)n_boxes), kPasteThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      masks, boxes, mask_size, padding, im_h, im_w, row_words, thresh, out);
  B200_CHECK_LAUNCH("paste_masks_kernel");
  return B200_OK;
}
