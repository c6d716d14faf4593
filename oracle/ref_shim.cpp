// oracle/ref_shim.cpp -- builds the REFERENCE's own CPU RoIAlign / NMS into
// oracle/_ref/libref_cpu.so, compiling the sources where they lie under
// /root/reference (nothing is copied into this repo; see build_oracle.py).
//
// TEST INFRASTRUCTURE ONLY (checker + CPU baseline); never on the product path.
//
// The reference targets a 2019 ATen.  Two things keep it from compiling against
// torch 2.11 unmodified: AT_DISPATCH_FLOATING_TYPES(x.type(), ...) needs
// ::detail::scalar_type(DeprecatedTypeProperties) (csrc/cpu/ROIAlign_cpu.cpp:242,
// csrc/cpu/nms_cpu.cpp:71), which newer headers dropped.  Supplying that one
// overload here lets the untouched files build.
#include <torch/extension.h>
#include <cstring>

namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) {
  return t.scalarType();
}
}  // namespace detail

// -I<reference>/maskrcnn_benchmark/csrc is on the command line.  The reference's
// cpu/vision.h pulls <torch/extension.h> (pybind11 + Python.h): the build adds
// those include dirs; the library is only ever dlopen'ed inside a Python
// process that has already imported torch.
#include "cpu/ROIAlign_cpu.cpp"
#include "cpu/nms_cpu.cpp"

extern "C" {

// input [B,C,H,W] fp32, rois [R,5] fp32 -> out [R,C,PH,PW] fp32
int ref_roi_align_forward(const float* input, int B, int C, int H, int W,
                          const float* rois, int R, float scale, int PH, int PW,
                          int sampling_ratio, float* out) {
  try {
    auto in_t = at::from_blob(const_cast<float*>(input), {B, C, H, W}, at::kFloat);
    auto roi_t = at::from_blob(const_cast<float*>(rois), {R, 5}, at::kFloat);
    at::Tensor o = ROIAlign_forward_cpu(in_t, roi_t, scale, PH, PW, sampling_ratio);
    std::memcpy(out, o.data_ptr<float>(), sizeof(float) * o.numel());
    return 0;
  } catch (...) {
    return -1;
  }
}

// dets [N,4], scores [N] -> keep (ascending original indices); returns count
long ref_nms(const float* dets, const float* scores, long N, float threshold,
             long* keep) {
  try {
    auto d = at::from_blob(const_cast<float*>(dets), {N, 4}, at::kFloat);
    auto s = at::from_blob(const_cast<float*>(scores), {N}, at::kFloat);
    at::Tensor k = nms_cpu(d, s, threshold).contiguous();
    std::memcpy(keep, k.data_ptr<int64_t>(), sizeof(int64_t) * k.numel());
    return (long)k.numel();
  } catch (...) {
    return -1;
  }
}

}  // extern "C"
