#include "frame_parse.h"

namespace pcdn {

bool parse_frame(const uint8_t* raw, uint32_t len, ParsedFrame* out) { return parse_frame_core(raw, len, out); }

uint32_t prune_topics(const uint8_t* in, uint32_t n, uint32_t n_valid, uint16_t* out) {
  uint32_t k = 0;
  for (uint32_t i = 0; i < n; i++)
    if (topic_kept(in, i, n_valid)) out[k++] = in[i];
  return k;
}

}  // namespace pcdn
