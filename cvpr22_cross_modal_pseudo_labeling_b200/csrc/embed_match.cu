// embed_match.cu -- placeholder until the tcgen05 kernel lands (next commit).
#include "common.cuh"

extern "C" int b200_embed_match(const void*, const void*, int64_t, int, int, int, float, float*, float*, int32_t*,
                                float*, const int32_t*, const int32_t*, const int32_t*, uint64_t*, void*) {
  b200::set_error("embed_match: not built yet");
  return B200_ERR_UNSUPPORTED;
}
extern "C" int b200_colmax_decode(const uint64_t*, int, int32_t*, float*, float*, void*) {
  b200::set_error("colmax_decode: not built yet");
  return B200_ERR_UNSUPPORTED;
}
