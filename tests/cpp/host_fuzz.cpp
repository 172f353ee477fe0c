// host_fuzz.cpp — the engine's HOST-side code (frame parser shared with the device kernel, table
// mirror, cuckoo insert/erase) under AddressSanitizer + UBSan, no GPU and no CUDA involved.
//   1. parse_frame / prune_topics on seed frames (read from a file of length-prefixed records) with
//      random mutations, truncations and extensions, each in an exact-size heap buffer so that any
//      read past `len` is an ASan error; results must stay inside the buffer.
//   2. Connections: a long random sequence of add/remove/kick/subscribe/unsubscribe/broker/sync calls
//      with invariants checked against a trivial model (who is connected, who is subscribed).
// Built and run by tests/test_host_sanitizers.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <set>
#include <string>
#include <vector>

#include "frame_parse.h"
#include "host_state.h"
#include "pcdn_fanout.h"

using namespace pcdn;

#define REQUIRE(c)                                                                  \
  do {                                                                              \
    if (!(c)) { fprintf(stderr, "REQUIRE failed %s:%d: %s\n", __FILE__, __LINE__, #c); exit(2); } \
  } while (0)

static std::vector<std::vector<uint8_t>> read_seeds(const char* path) {
  std::vector<std::vector<uint8_t>> out;
  FILE* f = fopen(path, "rb");
  REQUIRE(f != nullptr);
  for (;;) {
    uint32_t n;
    if (fread(&n, 4, 1, f) != 1) break;
    std::vector<uint8_t> v(n);
    if (n) REQUIRE(fread(v.data(), 1, n, f) == n);
    out.push_back(std::move(v));
  }
  fclose(f);
  return out;
}

static void fuzz_parser(const std::vector<std::vector<uint8_t>>& seeds, uint32_t iters, std::mt19937_64& rng) {
  uint64_t ok = 0, bad = 0;
  for (uint32_t it = 0; it < iters; it++) {
    std::vector<uint8_t> v = seeds[rng() % seeds.size()];
    const int nmut = (int)(rng() % 5);
    for (int k = 0; k < nmut && !v.empty(); k++) {
      size_t i = (rng() % 3) ? rng() % std::min<size_t>(v.size(), 96) : rng() % v.size();
      static const uint8_t vals[] = {0, 1, 2, 3, 4, 0x7F, 0x80, 0xFF};
      v[i] = (rng() % 2) ? vals[rng() % 8] : (uint8_t)rng();
    }
    if (rng() % 6 == 0) v.resize(rng() % (v.size() + 1));
    if (rng() % 16 == 0) { size_t extra = rng() % 40; for (size_t k = 0; k < extra; k++) v.push_back((uint8_t)rng()); }
    // exact-size heap copy: ASan red zones sit right behind the last byte
    uint8_t* buf = (uint8_t*)malloc(v.size() ? v.size() : 1);
    if (!v.empty()) memcpy(buf, v.data(), v.size());
    ParsedFrame pf;
    const bool good = parse_frame(buf, (uint32_t)v.size(), &pf);
    if (good) {
      ok++;
      REQUIRE(pf.kind >= 0 && pf.kind <= 8);
      REQUIRE((uint64_t)pf.f0_off + pf.f0_len <= v.size());
      REQUIRE((uint64_t)pf.f1_off + pf.f1_len <= v.size());
      if (pf.kind >= 4 && pf.kind <= 6) {
        uint16_t topics[256];
        const uint32_t n_in = std::min<uint32_t>(pf.f0_len, 256);
        const uint32_t n = prune_topics(buf + pf.f0_off, n_in, (uint32_t)(rng() % 4), topics);
        REQUIRE(n <= n_in);
      }
    } else {
      bad++;
    }
    free(buf);
  }
  printf("parser: %llu accepted, %llu rejected\n", (unsigned long long)ok, (unsigned long long)bad);
  REQUIRE(ok > iters / 20 && bad > iters / 20);
}

// n_shards > 1: the same model, with the connection ids spread over shard ranges of 8192 ids
// (usable part: `per_shard` ids each) — ids must stay inside the usable part of their shard and the
// shards must stay balanced (least-loaded hand-out).
static void fuzz_connections(uint32_t iters, std::mt19937_64& rng, uint32_t n_shards = 1, uint32_t per_shard = 300) {
  Geometry g;
  g.n_shards = n_shards; g.shard_N = 8192; g.shard_max_conns = per_shard;
  g.max_conns = (n_shards - 1) * g.shard_N + per_shard; g.N = n_shards * g.shard_N;
  g.W = g.N / 32; g.T = 64; g.max_keys = 1024; g.max_key_len = 40; g.key_stride = 48;
  g.nbuckets = 512; g.bucket_mask = 511; g.max_owners = 8; g.seed = 12345;
  const uint32_t capacity = n_shards * per_shard;
  HostTables t(g);
  Connections c(t, "me/me");
  std::map<std::string, std::set<uint16_t>> model;  // connected users → topics
  std::map<std::string, uint32_t> conn_of;
  auto key_of = [&](uint32_t i) { std::string k = "user-" + std::to_string(i); k.resize(8 + i % 30, 'x'); return k; };
  const char* brokers[3] = {"b0/p0", "b1/p1", "b2/p2"};
  std::set<int> bconn;
  for (uint32_t it = 0; it < iters; it++) {
    const uint32_t op = (uint32_t)(rng() % 100);
    const std::string k = key_of((uint32_t)(rng() % 400));
    uint16_t tp[4];
    const uint32_t nt = (uint32_t)(rng() % 4);
    for (uint32_t i = 0; i < nt; i++) tp[i] = (uint16_t)(rng() % 64);
    if (op < 30) {
      uint32_t conn = 0;
      const int rc = c.add_user(k, tp, nt, &conn);
      if (rc == 0) {
        REQUIRE(conn < g.max_conns && conn % g.shard_N < per_shard);
        model[k] = std::set<uint16_t>(tp, tp + nt);
        conn_of[k] = conn;
      } else {
        REQUIRE(rc == PCDN_ENOSPC || rc == PCDN_EAGAIN);
        if (rc == PCDN_ENOSPC && !model.count(k)) REQUIRE(model.size() + bconn.size() >= capacity || t.n_keys() >= g.max_keys - 8);
      }
    } else if (op < 45) {
      c.remove_user(k);
      model.erase(k); conn_of.erase(k);
    } else if (op < 65) {
      REQUIRE(c.subscribe_user_to(k, tp, nt) == 0);
      if (model.count(k)) model[k].insert(tp, tp + nt);
    } else if (op < 80) {
      c.unsubscribe_user_from(k, tp, nt);
      if (model.count(k)) for (uint32_t i = 0; i < nt; i++) model[k].erase(tp[i]);
    } else if (op < 86) {
      const int b = (int)(rng() % 3);
      uint32_t conn;
      if (c.add_broker(brokers[b], &conn) == 0) bconn.insert(b);
    } else if (op < 90) {
      const int b = (int)(rng() % 3);
      c.remove_broker(brokers[b]);
      bconn.erase(b);
    } else if (op < 95) {
      std::vector<UserSyncEntry> e;
      for (int i = 0; i < 3; i++) e.push_back(UserSyncEntry{key_of((uint32_t)(400 + rng() % 50)), rng() % 5, (rng() % 4) != 0, brokers[rng() % 3]});
      c.apply_user_sync(brokers[rng() % 3], e);
    } else {
      std::vector<UserSyncEntry> out;
      c.get_partial_user_sync(out);
      std::vector<TopicSyncEntry> ts;
      c.get_partial_topic_sync(ts);
    }
    if (it % 64 == 0) {  // mirror == model
      REQUIRE(c.num_users() == model.size());
      for (uint16_t topic = 0; topic < 64; topic += 7) {
        std::vector<uint32_t> got;
        c.interested(&topic, 1, /*to_users_only=*/true, got);
        std::set<uint32_t> want;
        for (auto& kv : model) if (kv.second.count(topic)) want.insert(conn_of[kv.first]);
        REQUIRE(std::set<uint32_t>(got.begin(), got.end()) == want);
      }
      std::vector<uint32_t> load(n_shards, 0);
      for (auto& kv : conn_of) {
        uint32_t conn = 0;
        REQUIRE(c.route(kv.first, &conn) == 1 && conn == kv.second);
        load[conn / g.shard_N]++;
      }
      t.clear_dirty();
    }
  }
  printf("connections (%u shard%s): %u ops, %zu users connected at the end\n", n_shards, n_shards > 1 ? "s" : "", iters, model.size());
}

// 3. Cuckoo table driven far past its design load (8 buckets x 4 slots, 64 keys allowed): the first
//    refused insert must leave every earlier key resolvable with its own route (the eviction walk is
//    rolled back) and must give the key-arena slot back (a later erase + insert succeeds again).
static void overfill_cuckoo(std::mt19937_64& rng) {
  Geometry g;
  g.shard_N = 8192; g.shard_max_conns = 64; g.max_conns = 64; g.N = 8192; g.W = 256; g.T = 1; g.max_keys = 64; g.max_key_len = 16; g.key_stride = 16;
  g.nbuckets = 8; g.bucket_mask = 7; g.max_owners = 4; g.seed = 0x1234567ull;
  for (int round = 0; round < 50; round++) {
    HostTables t(g);
    std::vector<std::string> in;
    int refused = 0;
    for (uint32_t i = 0; i < 64; i++) {
      std::string k(12, '\0');
      for (auto& c : k) c = (char)rng();
      int rc = t.route_upsert((const uint8_t*)k.data(), 12, 1000 + i);
      if (rc == 0) { in.push_back(k); continue; }
      REQUIRE(rc == PCDN_ENOSPC);
      refused++;
      for (size_t j = 0; j < in.size(); j++) {
        uint32_t r = 0;
        REQUIRE(t.route_find((const uint8_t*)in[j].data(), 12, &r));
      }
      uint32_t r;
      REQUIRE(!t.route_find((const uint8_t*)k.data(), 12, &r));
      REQUIRE(t.n_keys() == in.size());
    }
    REQUIRE(refused > 0);  // 32 slots, up to 64 inserts: the table must have refused some
    // routes are intact value-wise too
    std::set<uint32_t> seen;
    for (auto& k : in) { uint32_t r = 0; REQUIRE(t.route_find((const uint8_t*)k.data(), 12, &r)); REQUIRE(seen.insert(r).second); }
    // erase one, insert a fresh key into the freed space
    t.route_erase((const uint8_t*)in[0].data(), 12);
    REQUIRE(t.n_keys() == in.size() - 1);
  }
  printf("cuckoo overfill: rollback ok\n");
}

int main(int argc, char** argv) {
  REQUIRE(argc >= 2);
  const uint32_t iters = argc > 2 ? (uint32_t)atoi(argv[2]) : 200000;
  std::mt19937_64 rng(argc > 3 ? (uint64_t)atoll(argv[3]) : 1);
  auto seeds = read_seeds(argv[1]);
  REQUIRE(!seeds.empty());
  fuzz_parser(seeds, iters, rng);
  fuzz_connections(iters / 4, rng);
  fuzz_connections(iters / 8, rng, 3, 110);
  overfill_cuckoo(rng);
  printf("host_fuzz ok\n");
  return 0;
}
