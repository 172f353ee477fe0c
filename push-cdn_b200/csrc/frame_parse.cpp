#include "frame_parse.h"

#include <cstring>

namespace pcdn {
namespace {

struct Segs {
  const uint8_t* raw;
  uint32_t n;
  uint32_t off[512];    // byte offset of each segment in raw
  uint32_t words[512];
};

inline bool rd(const Segs& s, uint32_t seg, uint64_t word, uint64_t* v) {
  if (seg >= s.n || word >= s.words[seg]) return false;
  std::memcpy(v, s.raw + s.off[seg] + word * 8, 8);
  return true;
}

// A resolved pointer: the near pointer word plus the position its offset counts from
// (or, for a double-far pointer, the object start itself).
struct Near {
  uint64_t w;
  uint32_t seg;
  uint64_t base;  // word index the signed offset is relative to
  bool absolute;  // true: `base` already is the object start (double-far)
};

bool resolve(const Segs& s, uint32_t seg, uint64_t word, Near* out) {
  uint64_t w;
  if (!rd(s, seg, word, &w)) return false;
  if ((w & 3) != 2) { *out = Near{w, seg, word + 1, false}; return true; }
  uint32_t tseg = (uint32_t)(w >> 32);
  uint64_t pad = (uint32_t)w >> 3;
  if (!((w >> 2) & 1)) {                       // single far: pad holds the real pointer
    uint64_t p;
    if (!rd(s, tseg, pad, &p) || (p & 3) == 2) return false;
    *out = Near{p, tseg, pad + 1, false};
    return true;
  }
  uint64_t far2, tag;                          // double far: pad = far pointer + tag word
  if (!rd(s, tseg, pad, &far2) || !rd(s, tseg, pad + 1, &tag)) return false;
  if ((far2 & 3) != 2 || ((far2 >> 2) & 1)) return false;
  *out = Near{tag, (uint32_t)(far2 >> 32), (uint32_t)far2 >> 3, true};
  return true;
}

inline uint64_t target(const Near& n) {
  if (n.absolute) return n.base;
  return (uint64_t)((int64_t)n.base + ((int32_t)(uint32_t)n.w >> 2));
}

// byte list (Data / List(UInt8)) → (offset, length) inside raw
bool byte_list(const Segs& s, uint32_t seg, uint64_t word, uint32_t* off, uint32_t* len) {
  Near n;
  if (!resolve(s, seg, word, &n)) return false;
  *off = 0; *len = 0;
  if (n.w == 0) return true;                    // null pointer: empty default
  if ((n.w & 3) != 1) return false;
  uint32_t hi = (uint32_t)(n.w >> 32);
  if ((hi & 7) != 2) return false;              // element size BYTE only (what the clients emit)
  uint32_t count = hi >> 3;
  uint64_t start = target(n), words = ((uint64_t)count + 7) / 8;
  if (n.seg >= s.n || start > s.words[n.seg] || words > s.words[n.seg] - start) return false;
  *off = s.off[n.seg] + (uint32_t)start * 8;
  *len = count;
  return true;
}

}  // namespace

bool parse_frame(const uint8_t* raw, uint32_t len, ParsedFrame* out) {
  *out = ParsedFrame();
  if (len < 8) return false;
  uint32_t nm1;
  std::memcpy(&nm1, raw, 4);
  uint64_t nseg = (uint64_t)nm1 + 1;
  if (nseg >= 512) return false;
  uint64_t table = (4 + 4 * nseg + 7) & ~7ull;
  if (table > len) return false;
  Segs s;
  s.raw = raw;
  s.n = (uint32_t)nseg;
  uint64_t pos = table;
  for (uint32_t i = 0; i < s.n; i++) {
    uint32_t w;
    std::memcpy(&w, raw + 4 + 4 * i, 4);
    if ((uint64_t)w * 8 > len - pos) return false;
    s.off[i] = (uint32_t)pos;
    s.words[i] = w;
    pos += (uint64_t)w * 8;
  }
  Near root;
  if (!resolve(s, 0, 0, &root)) return false;
  if (root.w == 0) { out->kind = 0; return true; }
  if ((root.w & 3) != 0) return false;
  uint64_t st = target(root);
  uint32_t dw = (uint32_t)(root.w >> 32) & 0xFFFF, pw = (uint32_t)(root.w >> 48);
  if (root.seg >= s.n || st > s.words[root.seg] || (uint64_t)dw + pw > s.words[root.seg] - st) return false;
  uint16_t tag = 0;
  if (dw) { uint64_t d = 0; rd(s, root.seg, st, &d); tag = (uint16_t)d; }
  if (tag > 8) return false;
  out->kind = tag;
  if (!pw) return true;
  uint64_t ptr = st + dw;
  switch (tag) {
    case 3: case 4: {
      Near v;
      if (!resolve(s, root.seg, ptr, &v)) return false;
      if (v.w == 0) return true;
      if ((v.w & 3) != 0) return false;
      uint64_t vs = target(v);
      uint32_t vd = (uint32_t)(v.w >> 32) & 0xFFFF, vp = (uint32_t)(v.w >> 48);
      if (v.seg >= s.n || vs > s.words[v.seg] || (uint64_t)vd + vp > s.words[v.seg] - vs) return false;
      if (vp >= 1 && !byte_list(s, v.seg, vs + vd, &out->f0_off, &out->f0_len)) return false;
      if (vp >= 2 && !byte_list(s, v.seg, vs + vd + 1, &out->f1_off, &out->f1_len)) return false;
      return true;
    }
    case 5: case 6: case 7: case 8:
      return byte_list(s, root.seg, ptr, &out->f0_off, &out->f0_len);
    default:
      return true;
  }
}

uint32_t prune_topics(const uint8_t* in, uint32_t n, uint32_t n_valid, uint16_t* out) {
  uint32_t k = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (i > 0 && in[i] == in[i - 1]) continue;          // Vec::dedup(): consecutive duplicates
    if (n_valid && in[i] >= n_valid) continue;          // retain(|t| Topic::try_from(t).is_ok())
    out[k++] = in[i];
  }
  return k;
}

}  // namespace pcdn
