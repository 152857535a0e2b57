// arrow_bridge.cpp -- Arrow C Stream level of the C ABI (pbgpu_range_op).  Placeholder until the
// ingest / materialise path lands: returns PBGPU_EINVAL without touching the streams' contents.
#include "../../include/pbgpu.h"

namespace pbgpu { int set_error(int code, const char *fmt, ...); }

extern "C" int pbgpu_range_op(struct ArrowArrayStream *, struct ArrowArrayStream *, const PbRangeOptions *,
                              struct ArrowArrayStream *) {
  return pbgpu::set_error(PBGPU_EINVAL, "pbgpu_range_op: Arrow-level entry point not built yet");
}
