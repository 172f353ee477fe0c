// frame_parse_core.h — the Cap'n Proto walk shared by the host parser (frame_parse.cpp) and the
// device parse kernel (k_parse in kernels.cu): one source, two compilers, identical decisions.
//
// Restates what the broker needs from Message::deserialize (cdn-proto/src/message.rs:212-312):
// stream framing (u32 LE nseg-1, nseg × u32 LE words, pad to 8), root struct Message = 1 data word
// (u16 union tag @0) + 1 pointer (messages_capnp.rs:175), Direct/Broadcast = 0 data + 2 pointers
// (:1438,:1687), byte lists, single and double far pointers.  No allocation, no recursion, segment
// sizes are re-read from the table on demand (frames have 1-2 segments in practice).
#pragma once
#include <stdint.h>

#include "hash.h"  // PCDN_HD

namespace pcdn {

struct ParsedFrame {
  int kind;          // capnp union tag 0..8, -1 = malformed
  uint32_t f0_off;   // byte offset in raw of field 0 (topics list / recipient / sync blob)
  uint32_t f0_len;
  uint32_t f1_off;   // Direct.message / Broadcast.message
  uint32_t f1_len;
};

namespace fpc {

PCDN_HD uint32_t rd32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
PCDN_HD uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

struct Msg {
  const uint8_t* raw;
  uint32_t len;
  uint32_t nseg;
  uint32_t table;  // byte offset of segment 0
};

// byte offset and size (words) of segment `seg`; false when out of range
PCDN_HD bool seg_info(const Msg& m, uint32_t seg, uint32_t* off, uint32_t* words) {
  if (seg >= m.nseg) return false;
  uint64_t pos = m.table;
  for (uint32_t i = 0; i < seg; i++) pos += (uint64_t)rd32(m.raw + 4 + 4 * i) * 8;
  *off = (uint32_t)pos;
  *words = rd32(m.raw + 4 + 4 * seg);
  return true;
}
PCDN_HD bool rd(const Msg& m, uint32_t seg, uint64_t word, uint64_t* v) {
  uint32_t off, words;
  if (!seg_info(m, seg, &off, &words) || word >= words) return false;
  *v = rd64(m.raw + off + word * 8);
  return true;
}

struct Near {
  uint64_t w;
  uint32_t seg;
  uint64_t base;  // word the signed offset counts from, or the object start when `absolute`
  bool absolute;
};

PCDN_HD bool resolve(const Msg& m, uint32_t seg, uint64_t word, Near* out) {
  uint64_t w;
  if (!rd(m, seg, word, &w)) return false;
  if ((w & 3) != 2) { out->w = w; out->seg = seg; out->base = word + 1; out->absolute = false; return true; }
  const uint32_t tseg = (uint32_t)(w >> 32);
  const uint64_t pad = (uint32_t)w >> 3;
  if (!((w >> 2) & 1)) {  // single far: landing pad holds the real pointer
    uint64_t p;
    if (!rd(m, tseg, pad, &p) || (p & 3) == 2) return false;
    out->w = p; out->seg = tseg; out->base = pad + 1; out->absolute = false;
    return true;
  }
  uint64_t far2, tag;     // double far: pad = far pointer to the object + tag word
  if (!rd(m, tseg, pad, &far2) || !rd(m, tseg, pad + 1, &tag)) return false;
  if ((far2 & 3) != 2 || ((far2 >> 2) & 1)) return false;
  out->w = tag; out->seg = (uint32_t)(far2 >> 32); out->base = (uint32_t)far2 >> 3; out->absolute = true;
  return true;
}
PCDN_HD uint64_t target(const Near& n) {
  if (n.absolute) return n.base;
  return (uint64_t)((int64_t)n.base + ((int32_t)(uint32_t)n.w >> 2));
}
// byte list (Data / List(UInt8)) → (offset, length) inside raw
PCDN_HD bool byte_list(const Msg& m, uint32_t seg, uint64_t word, uint32_t* off, uint32_t* len) {
  Near n;
  if (!resolve(m, seg, word, &n)) return false;
  *off = 0; *len = 0;
  if (n.w == 0) return true;            // null pointer: empty default
  if ((n.w & 3) != 1) return false;
  const uint32_t hi = (uint32_t)(n.w >> 32);
  if ((hi & 7) != 2) return false;      // element size BYTE only (what the clients emit)
  const uint32_t count = hi >> 3;
  const uint64_t start = target(n), words = ((uint64_t)count + 7) / 8;
  uint32_t soff, swords;
  if (!seg_info(m, n.seg, &soff, &swords) || start > swords || words > swords - start) return false;
  *off = soff + (uint32_t)start * 8;
  *len = count;
  return true;
}

}  // namespace fpc

// true + `out` filled, or false = Error::Deserialize (the reference disconnects the peer)
PCDN_HD bool parse_frame_core(const uint8_t* raw, uint32_t len, ParsedFrame* out) {
  using namespace fpc;
  out->kind = -1; out->f0_off = out->f0_len = out->f1_off = out->f1_len = 0;
  if (len < 8) return false;
  const uint64_t nseg = (uint64_t)rd32(raw) + 1;
  if (nseg >= 512) return false;        // capnp-rust: "Too many segments"
  const uint64_t table = (4 + 4 * nseg + 7) & ~7ull;
  if (table > len) return false;
  uint64_t pos = table;
  for (uint32_t i = 0; i < (uint32_t)nseg; i++) {
    const uint64_t bytes = (uint64_t)rd32(raw + 4 + 4 * i) * 8;
    if (bytes > len - pos) return false;  // premature end of message
    pos += bytes;
  }
  Msg m{raw, len, (uint32_t)nseg, (uint32_t)table};
  Near root;
  if (!resolve(m, 0, 0, &root)) return false;
  if (root.w == 0) { out->kind = 0; return true; }
  if ((root.w & 3) != 0) return false;
  const uint64_t st = target(root);
  const uint32_t dw = (uint32_t)(root.w >> 32) & 0xFFFF, pw = (uint32_t)(root.w >> 48);
  uint32_t roff, rwords;
  if (!seg_info(m, root.seg, &roff, &rwords) || st > rwords || (uint64_t)dw + pw > rwords - st) return false;
  uint32_t tag = 0;
  if (dw) tag = (uint32_t)(rd64(raw + roff + st * 8) & 0xFFFF);
  if (tag > 8) return false;            // "message not in schema"
  out->kind = (int)tag;
  if (!pw) return true;
  const uint64_t ptr = st + dw;
  if (tag == 3 || tag == 4) {
    Near v;
    if (!resolve(m, root.seg, ptr, &v)) return false;
    if (v.w == 0) return true;
    if ((v.w & 3) != 0) return false;
    const uint64_t vs = target(v);
    const uint32_t vd = (uint32_t)(v.w >> 32) & 0xFFFF, vp = (uint32_t)(v.w >> 48);
    uint32_t voff, vwords;
    if (!seg_info(m, v.seg, &voff, &vwords) || vs > vwords || (uint64_t)vd + vp > vwords - vs) return false;
    if (vp >= 1 && !byte_list(m, v.seg, vs + vd, &out->f0_off, &out->f0_len)) return false;
    if (vp >= 2 && !byte_list(m, v.seg, vs + vd + 1, &out->f1_off, &out->f1_len)) return false;
    return true;
  }
  if (tag >= 5) return byte_list(m, root.seg, ptr, &out->f0_off, &out->f0_len);
  return true;
}

// Union tag only (device-parse mode: the host just decides which receive path a frame takes).
// -1 when the root struct cannot be resolved.
PCDN_HD int peek_kind_core(const uint8_t* raw, uint32_t len) {
  using namespace fpc;
  if (len < 8) return -1;
  const uint64_t nseg = (uint64_t)rd32(raw) + 1;
  if (nseg >= 512) return -1;
  const uint64_t table = (4 + 4 * nseg + 7) & ~7ull;
  if (table > len) return -1;
  uint64_t pos = table;
  for (uint32_t i = 0; i < (uint32_t)nseg; i++) {
    const uint64_t bytes = (uint64_t)rd32(raw + 4 + 4 * i) * 8;
    if (bytes > len - pos) return -1;
    pos += bytes;
  }
  Msg m{raw, len, (uint32_t)nseg, (uint32_t)table};
  Near root;
  if (!resolve(m, 0, 0, &root)) return -1;
  if (root.w == 0) return 0;
  if ((root.w & 3) != 0) return -1;
  const uint64_t st = target(root);
  const uint32_t dw = (uint32_t)(root.w >> 32) & 0xFFFF, pw = (uint32_t)(root.w >> 48);
  uint32_t roff, rwords;
  if (!seg_info(m, root.seg, &roff, &rwords) || st > rwords || (uint64_t)dw + pw > rwords - st) return -1;
  const uint32_t tag = dw ? (uint32_t)(rd64(raw + roff + st * 8) & 0xFFFF) : 0;
  return tag > 8 ? -1 : (int)tag;
}

// Topic::prune (cdn-proto/src/def.rs:36-49) as a predicate over position i of the wire topic list:
// Vec::dedup() drops an element equal to its predecessor, retain() drops invalid topics.
PCDN_HD bool topic_kept(const uint8_t* topics, uint32_t i, uint32_t n_valid) {
  if (i > 0 && topics[i] == topics[i - 1]) return false;
  if (n_valid && topics[i] >= n_valid) return false;
  return true;
}

}  // namespace pcdn
