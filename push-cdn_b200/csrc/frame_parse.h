// frame_parse.h — ingress parse of a cdn-proto frame body, restating what the broker's receive
// loops need from Message::deserialize (cdn-proto/src/message.rs:212-312): the union tag and, for
// the routed kinds, WHERE inside the raw bytes the topics / recipient lie.  Nothing is copied — the
// broker forwards the raw bytes verbatim (R1), and the direct-lookup kernel reads the recipient key
// straight out of the frame in HBM.
//
// Wire layout: Cap'n Proto stream framing (u32 LE nseg-1, nseg × u32 LE words, pad to 8) followed by
// the segments; Message = 1 data word (u16 union tag @0) + 1 pointer (messages_capnp.rs:175);
// Direct/Broadcast = 0 data + 2 pointers (:1438,:1687); far pointers (single and double) are
// followed because capnp-rust spills payloads that do not fit its 1024-word first segment.
#pragma once
#include <cstdint>

namespace pcdn {

struct ParsedFrame {
  int kind = -1;           // capnp union tag 0..8
  uint32_t f0_off = 0;     // byte offset in raw of field 0 (topics list / recipient / sync blob)
  uint32_t f0_len = 0;
  uint32_t f1_off = 0;     // Direct.message / Broadcast.message
  uint32_t f1_len = 0;
};

// returns true and fills `out`, or false = Error::Deserialize (the peer is disconnected)
bool parse_frame(const uint8_t* raw, uint32_t len, ParsedFrame* out);

// Topic::prune (cdn-proto/src/def.rs:36-49): consecutive dedup, keep valid topics (t < n_valid,
// 0 = every u8 valid); returns the pruned count (0 = Err "supplied no valid topics")
uint32_t prune_topics(const uint8_t* in, uint32_t n, uint32_t n_valid, uint16_t* out);

}  // namespace pcdn
