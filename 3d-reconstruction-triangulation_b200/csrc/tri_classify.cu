// tri_classify.cu -- kernel 3 (placeholder until the batched classifier lands in this round)
#include "tri_engine.cuh"
extern "C" int tri_classify(tri_engine*, int, unsigned, int, const int32_t*, const double*, int, double*, int8_t*, uint8_t*,
                            tri_classify_stats*) {
  return tri::fail(TRI_ERR_ARG, "tri_classify: not built yet");
}
