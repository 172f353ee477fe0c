// hash.h — the direct-map key hash, identical on host (table maintenance) and device (lookup).
//
// A key of any length is read as little-endian 64-bit words w_0..w_{n-1} (zero padded).  Every word
// is mixed on its own with a position- and seed-dependent salt and the results are SUMMED, so a warp
// can hash one key with one word per lane and a shuffle reduction; the sum is then finalised with
// the length.  The seed is per engine (pcdn_config.hash_seed) so that bucket placement is not
// predictable from the public keys alone.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PCDN_HD __host__ __device__ __forceinline__
#else
#define PCDN_HD inline
#endif

namespace pcdn {

PCDN_HD uint64_t fmix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}
// contribution of 64-bit word `w` at word position `i`
PCDN_HD uint64_t key_word_mix(uint64_t w, uint32_t i, uint64_t seed) {
  return fmix64(w ^ (seed + (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ULL));
}
PCDN_HD uint64_t key_hash_finish(uint64_t acc, uint32_t len) {
  return fmix64(acc ^ ((uint64_t)len * 0xD6E8FEB86659FD93ULL));
}
// 32-bit non-zero fingerprint and the two candidate buckets (partial-key cuckoo: the alternate
// bucket is a function of (bucket, tag), so evictions never re-read the key)
PCDN_HD uint32_t key_tag(uint64_t h) { return (uint32_t)(h >> 32) | 1u; }
PCDN_HD uint32_t key_bucket(uint64_t h, uint32_t bucket_mask) { return (uint32_t)h & bucket_mask; }
PCDN_HD uint32_t alt_bucket(uint32_t b, uint32_t tag, uint32_t bucket_mask) {
  return (b ^ (tag * 0x5bd1e995u)) & bucket_mask;
}

#if !defined(__CUDA_ARCH__)
// host: sequential form of the same hash
inline uint64_t key_hash_host(const uint8_t* key, uint32_t len, uint64_t seed) {
  uint64_t acc = 0;
  uint32_t nw = (len + 7) / 8;
  for (uint32_t i = 0; i < nw; i++) {
    uint64_t w = 0;
    uint32_t rem = len - i * 8;
    for (uint32_t b = 0; b < (rem < 8 ? rem : 8); b++) w |= (uint64_t)key[i * 8 + b] << (8 * b);
    acc += key_word_mix(w, i, seed);
  }
  return key_hash_finish(acc, len);
}
#endif

// route encoding stored in the table
enum : uint32_t { ROUTE_NONE = 0xFFFFFFFFu, ROUTE_REMOTE = 0x80000000u };

struct CuckooEntry {  // 16 bytes, 4 per bucket (one 64-byte bucket = two sectors)
  uint32_t tag;       // 0 = empty
  uint32_t key_slot;  // index into the key arena
  uint32_t route;     // local user: connection id; remote: ROUTE_REMOTE | owner index
  uint32_t key_len;
};

}  // namespace pcdn
