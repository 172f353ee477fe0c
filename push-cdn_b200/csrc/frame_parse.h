// frame_parse.h — host ingress parse of a cdn-proto frame body (see frame_parse_core.h for the
// walk itself, which is shared with the device parse kernel).  Nothing is copied: the broker
// forwards the raw bytes verbatim (R1) and the direct-lookup kernel reads the recipient key
// straight out of the frame in HBM.
#pragma once
#include <cstdint>

#include "frame_parse_core.h"

namespace pcdn {

// returns true and fills `out`, or false = Error::Deserialize (the peer is disconnected)
bool parse_frame(const uint8_t* raw, uint32_t len, ParsedFrame* out);

// Topic::prune (cdn-proto/src/def.rs:36-49): consecutive dedup, keep valid topics (t < n_valid,
// 0 = every u8 valid); returns the pruned count (0 = Err "supplied no valid topics")
uint32_t prune_topics(const uint8_t* in, uint32_t n, uint32_t n_valid, uint16_t* out);

}  // namespace pcdn
