// ORACLE — test infrastructure only (see oracle/README.md).  CPU restatement of the Push-CDN
// cdn-broker routing hot path with the reference's own data structures.  Nothing here is used by
// the product path; tests/, smoke() and bench.py's CPU legs are the only callers.
//
// Each function cites the reference file:line it restates (paths relative to the reference repo).
// The reference cannot be built here (Rust; no cargo/rustc), so this restatement is pinned by
// replaying the reference's own tests against it (tests/test_oracle_reference.py, tests/scenarios.py):
//   cdn-broker/src/tests/broadcast.rs:26-167, cdn-broker/src/tests/direct.rs:27-173,
//   cdn-broker/src/connections/broadcast/relational_map.rs:132-346,
//   cdn-broker/src/connections/versioned_map.rs:277-376, cdn-broker/src/connections/mod.rs:410-526,
//   tests/src/tests/subscribe.rs:19-121, tests/src/tests/double_connect.rs:16-58.
//
// Topic is u8 in the reference (cdn-proto/src/message.rs:26).  The oracle carries topics as u16 so
// that the BASELINE "extended" configs (1 K / 4 K topics) can be checked too; wire-exact tests
// stay below 256.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "capnp_lite.hpp"

namespace {

using Key = std::string;     // UserPublicKey = Arc<Vec<u8>>, bytewise identity (connection/mod.rs:20, R8)
using Topic = uint16_t;

// BrokerIdentifier, ordered lexicographically by (public, private) — derive(Ord),
// cdn-proto/src/discovery/mod.rs:80-87; parsed from "pub/priv" (:104-129).
struct BrokerId {
  std::string pub, priv;
  bool operator==(const BrokerId& o) const { return pub == o.pub && priv == o.priv; }
  bool operator!=(const BrokerId& o) const { return !(*this == o); }
  bool operator<(const BrokerId& o) const { return pub != o.pub ? pub < o.pub : priv < o.priv; }
  bool operator>(const BrokerId& o) const { return o < *this; }
  std::string str() const { return pub + "/" + priv; }
  static BrokerId parse(const std::string& s) {
    BrokerId b;
    size_t a = s.find('/');
    if (a == std::string::npos) { b.pub = s; return b; }
    b.pub = s.substr(0, a);
    size_t c = s.find('/', a + 1);
    b.priv = s.substr(a + 1, c == std::string::npos ? std::string::npos : c - a - 1);
    return b;
  }
};
struct BrokerIdHash {
  size_t operator()(const BrokerId& b) const { return std::hash<std::string>()(b.str()); }
};

// ---- RelationalMap: cdn-broker/src/connections/broadcast/relational_map.rs:13-116 -------------
template <class K, class V, class KH = std::hash<K>>
struct RelationalMap {
  std::unordered_map<K, std::unordered_set<V>, KH> key_to_values;
  std::unordered_map<V, std::unordered_set<K, KH>> value_to_keys;

  // get_keys_by_value :39-47
  std::vector<K> get_keys_by_value(const V& v) const {
    std::vector<K> out;
    auto it = value_to_keys.find(v);
    if (it != value_to_keys.end()) out.assign(it->second.begin(), it->second.end());
    return out;
  }
  // get_values :51-54
  std::vector<V> get_values() const {
    std::vector<V> out;
    for (auto& kv : value_to_keys) out.push_back(kv.first);
    return out;
  }
  // associate_key_with_values :57-68
  void associate_key_with_values(const K& k, const std::vector<V>& vs) {
    auto& s = key_to_values[k];  // entry().or_default() — created even when vs is empty
    s.insert(vs.begin(), vs.end());
    for (auto& v : vs) value_to_keys[v].insert(k);
  }
  // dissociate_keys_from_value :71-96
  void dissociate_keys_from_value(const K& k, const std::vector<V>& vs) {
    for (auto& v : vs) {
      auto it = value_to_keys.find(v);
      if (it != value_to_keys.end()) {
        it->second.erase(k);
        if (it->second.empty()) value_to_keys.erase(it);
      }
    }
    auto kt = key_to_values.find(k);
    if (kt != key_to_values.end()) {
      for (auto& v : vs) kt->second.erase(v);
      if (kt->second.empty()) key_to_values.erase(kt);
    }
  }
  // remove_key :100-115
  void remove_key(const K& k) {
    auto kt = key_to_values.find(k);
    if (kt == key_to_values.end()) return;
    std::unordered_set<V> vs = std::move(kt->second);
    key_to_values.erase(kt);
    for (auto& v : vs) {
      auto it = value_to_keys.find(v);
      if (it != value_to_keys.end()) {
        it->second.erase(k);
        if (it->second.empty()) value_to_keys.erase(it);
      }
    }
  }
};

// ---- VersionedMap: cdn-broker/src/connections/versioned_map.rs:27-270 -------------------------
template <class K, class V, class C, class KH = std::hash<K>>
struct VersionedMap {
  struct VersionedValue { uint64_t version; std::optional<V> value; };
  std::unordered_map<K, VersionedValue, KH> underlying_map;
  std::unordered_set<K, KH> locally_modified_keys;
  C conflict_identity;

  explicit VersionedMap(C c = C()) : conflict_identity(std::move(c)) {}
  bool is_empty() const { return underlying_map.empty(); }
  // get :76-81
  const V* get(const K& k) const {
    auto it = underlying_map.find(k);
    if (it == underlying_map.end() || !it->second.value) return nullptr;
    return &*it->second.value;
  }
  // modify_local :84-113
  void modify_local(const K& k, std::optional<V> v) {
    auto it = underlying_map.find(k);
    if (it != underlying_map.end()) {
      if (!locally_modified_keys.count(k)) it->second.version += 1;
      it->second.value = std::move(v);
    } else {
      underlying_map.emplace(k, VersionedValue{1, std::move(v)});
    }
    locally_modified_keys.insert(k);
  }
  void insert(const K& k, const V& v) { modify_local(k, v); }          // :117-119
  void remove(const K& k) { modify_local(k, std::nullopt); }           // :123-125
  // remove_if_equals :128-136
  void remove_if_equals(const K& k, const V& v) {
    auto it = underlying_map.find(k);
    if (it != underlying_map.end() && it->second.value && *it->second.value == v) remove(k);
  }
  // remove_by_value_no_modify :141-155
  void remove_by_value_no_modify(const V& v) {
    std::vector<K> keys;
    for (auto& kv : underlying_map)
      if (kv.second.value && *kv.second.value == v) keys.push_back(kv.first);
    for (auto& k : keys) underlying_map.erase(k);
  }
  // get_full :159-165
  VersionedMap get_full() const {
    VersionedMap m(conflict_identity);
    m.underlying_map = underlying_map;
    return m;
  }
  // diff :169-195
  VersionedMap diff() {
    auto mod = std::move(locally_modified_keys);
    locally_modified_keys.clear();
    VersionedMap d(conflict_identity);
    for (auto& k : mod) {
      auto it = underlying_map.find(k);
      if (it != underlying_map.end()) {
        d.underlying_map.emplace(k, it->second);
        if (!it->second.value) underlying_map.erase(it);
      }
    }
    return d;
  }
  // merge :202-269 — returns the (key, new value) changes
  std::vector<std::pair<K, std::optional<V>>> merge(const VersionedMap& remote) {
    std::vector<std::pair<K, std::optional<V>>> changes;
    for (auto& rkv : remote.underlying_map) {
      const K& rk = rkv.first;
      const VersionedValue& rv = rkv.second;
      auto it = underlying_map.find(rk);
      if (it != underlying_map.end()) {
        bool take = false;
        if (rv.version > it->second.version) take = true;
        else if (rv.version == it->second.version) take = remote.conflict_identity > conflict_identity;
        if (take) {
          if (rv.value) { it->second.value = rv.value; it->second.version = rv.version; }
          else underlying_map.erase(it);
          locally_modified_keys.erase(rk);
          changes.emplace_back(rk, rv.value);
        }
      } else if (rv.value) {
        underlying_map.emplace(rk, rv);
        changes.emplace_back(rk, rv.value);
      }
    }
    return changes;
  }
};

enum SubscriptionStatus : uint8_t { Subscribed = 0, Unsubscribed = 1 };  // broadcast/mod.rs:19-23
using DirectMap = VersionedMap<Key, BrokerId, BrokerId>;                 // direct/mod.rs:14
using TopicSyncMap = VersionedMap<Topic, SubscriptionStatus, uint32_t>;  // broadcast/mod.rs:25

// A Connection is modelled as the byte stream its writer task emits
// (cdn-proto/src/connection/protocols/mod.rs:156-186): FIFO queue → write_length_delimited.
struct Conn {
  int id = -1;
  bool is_broker = false;
  bool closed = false;   // send_message_raw fails once the channel is closed (:239-251)
  bool removed = false;  // stands for the aborted receive task (AbortHandle)
  std::vector<uint8_t> stream;
  std::vector<uint32_t> frame_lens;
};

struct Broker {
  int conn;
  TopicSyncMap topic_sync_map{0};
};

// ---- Connections: cdn-broker/src/connections/mod.rs:40-388 ------------------------------------
struct Oracle {
  BrokerId identity;
  std::unordered_map<Key, int> users;                       // :45
  std::unordered_map<BrokerId, Broker, BrokerIdHash> brokers;  // :47
  DirectMap direct_map;                                     // :50
  RelationalMap<Key, Topic> bm_users;                       // BroadcastMap broadcast/mod.rs:30-36
  RelationalMap<BrokerId, Topic, BrokerIdHash> bm_brokers;
  TopicSyncMap topic_sync_map{0};
  std::unordered_set<Topic> previous_subscribed_topics;
  std::vector<std::unique_ptr<Conn>> conns;
  uint32_t n_valid_topics = 0;  // Topic::try_from validity for prune (def.rs:25-49); 0 = all valid
  uint64_t bytes_sent = 0, deliveries = 0;

  explicit Oracle(const std::string& id) : identity(BrokerId::parse(id)), direct_map(identity) {}

  int new_conn(bool is_broker) {
    conns.emplace_back(new Conn());
    conns.back()->id = (int)conns.size() - 1;
    conns.back()->is_broker = is_broker;
    return conns.back()->id;
  }

  // get_broker_identifier_of_user :69-71
  const BrokerId* get_broker_identifier_of_user(const Key& u) const { return direct_map.get(u); }

  // get_interested_by_topic :94-124
  void get_interested_by_topic(const std::vector<Topic>& topics, bool to_users_only,
                               std::vector<BrokerId>& brokers_out, std::vector<Key>& users_out) const {
    std::unordered_set<BrokerId, BrokerIdHash> br;
    std::unordered_set<Key> us;
    for (Topic t : topics) {
      for (auto& u : bm_users.get_keys_by_value(t)) us.insert(u);
      if (!to_users_only)
        for (auto& b : bm_brokers.get_keys_by_value(t)) br.insert(b);
    }
    brokers_out.assign(br.begin(), br.end());
    users_out.assign(us.begin(), us.end());
  }

  // remove_user :330-351
  void remove_user(const Key& k) {
    auto it = users.find(k);
    if (it != users.end()) { conns[it->second]->removed = true; users.erase(it); }
    bm_users.remove_key(k);
    direct_map.remove_if_equals(k, identity);
  }
  // add_user :278-304
  int add_user(const Key& k, const std::vector<Topic>& topics) {
    remove_user(k);
    int c = new_conn(false);
    users[k] = c;
    direct_map.insert(k, identity);
    bm_users.associate_key_with_values(k, topics);
    return c;
  }
  // remove_broker :308-324
  void remove_broker(const BrokerId& b) {
    auto it = brokers.find(b);
    if (it != brokers.end()) { conns[it->second.conn]->removed = true; brokers.erase(it); }
    bm_brokers.remove_key(b);
  }
  // add_broker :252-274
  int add_broker(const BrokerId& b) {
    remove_broker(b);
    int c = new_conn(true);
    brokers.emplace(b, Broker{c});
    return c;
  }
  void subscribe_user_to(const Key& k, const std::vector<Topic>& t) { bm_users.associate_key_with_values(k, t); }    // :365
  void unsubscribe_user_from(const Key& k, const std::vector<Topic>& t) { bm_users.dissociate_keys_from_value(k, t); }  // :383
  void subscribe_broker_to(const BrokerId& b, const std::vector<Topic>& t) { bm_brokers.associate_key_with_values(b, t); }  // :354
  void unsubscribe_broker_from(const BrokerId& b, const std::vector<Topic>& t) { bm_brokers.dissociate_keys_from_value(b, t); }  // :372

  // apply_user_sync :154-162
  void apply_user_sync(const DirectMap& remote) {
    auto changed = direct_map.merge(remote);
    for (auto& kv : changed) remove_user(kv.first);
  }
  // apply_topic_sync :165-191
  void apply_topic_sync(const BrokerId& b, const TopicSyncMap& remote) {
    auto it = brokers.find(b);
    if (it == brokers.end()) { remove_broker(b); return; }
    auto changed = it->second.topic_sync_map.merge(remote);
    for (auto& kv : changed) {
      if (kv.second && *kv.second == Subscribed) subscribe_broker_to(b, {kv.first});
      else unsubscribe_broker_from(b, {kv.first});
    }
  }
  // get_partial_topic_sync :205-237
  std::optional<TopicSyncMap> get_partial_topic_sync() {
    std::unordered_set<Topic> now;
    for (Topic t : bm_users.get_values()) now.insert(t);
    std::vector<Topic> added, removed;
    for (Topic t : now) if (!previous_subscribed_topics.count(t)) added.push_back(t);
    for (Topic t : previous_subscribed_topics) if (!now.count(t)) removed.push_back(t);
    if (added.empty() && removed.empty()) return std::nullopt;
    previous_subscribed_topics = now;
    for (Topic t : added) topic_sync_map.insert(t, Subscribed);
    for (Topic t : removed) topic_sync_map.insert(t, Unsubscribed);
    return topic_sync_map.diff();
  }
  // get_full_topic_sync :194-200
  std::optional<TopicSyncMap> get_full_topic_sync() const {
    if (topic_sync_map.underlying_map.empty()) return std::nullopt;
    return topic_sync_map;  // clone (includes locally_modified_keys, irrelevant to merge)
  }

  // write_length_delimited protocols/mod.rs:354-394: u32 BE length, then the bytes; BYTES_SENT += len
  bool send_message_raw(int c, const uint8_t* raw, uint32_t len) {
    Conn& cn = *conns[c];
    if (cn.closed) return false;
    uint8_t hdr[4] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len};
    cn.stream.insert(cn.stream.end(), hdr, hdr + 4);
    cn.stream.insert(cn.stream.end(), raw, raw + len);
    cn.frame_lens.push_back(len);
    bytes_sent += len;
    deliveries += 1;
    return true;
  }
  // try_send_to_user tasks/user/sender.rs:16-32
  void try_send_to_user(const Key& u, const uint8_t* raw, uint32_t len) {
    auto it = users.find(u);
    if (it == users.end()) return;
    if (!send_message_raw(it->second, raw, len)) remove_user(u);
  }
  // try_send_to_broker tasks/broker/sender.rs:17-45
  void try_send_to_broker(const BrokerId& b, const uint8_t* raw, uint32_t len) {
    auto it = brokers.find(b);
    if (it == brokers.end()) return;
    if (!send_message_raw(it->second.conn, raw, len)) remove_broker(b);
  }
  // handle_direct_message tasks/broker/handler.rs:197-237
  void handle_direct_message(const Key& user, const uint8_t* raw, uint32_t len, bool to_user_only) {
    const BrokerId* b = get_broker_identifier_of_user(user);
    if (!b) return;
    BrokerId owner = *b;  // cloned in the reference before the lock is released
    if (owner == identity) try_send_to_user(user, raw, len);
    else if (!to_user_only) try_send_to_broker(owner, raw, len);
  }
  // handle_broadcast_message tasks/broker/handler.rs:240-272
  void handle_broadcast_message(const std::vector<Topic>& topics, const uint8_t* raw, uint32_t len,
                                bool to_users_only) {
    std::vector<BrokerId> ib; std::vector<Key> iu;
    get_interested_by_topic(topics, to_users_only, ib, iu);
    for (auto& b : ib) try_send_to_broker(b, raw, len);
    for (auto& u : iu) try_send_to_user(u, raw, len);
  }
  // Topic::prune cdn-proto/src/def.rs:36-49: consecutive dedup, keep valid, error when empty
  bool prune(std::vector<Topic>& t) const {
    t.erase(std::unique(t.begin(), t.end()), t.end());
    if (n_valid_topics)
      t.erase(std::remove_if(t.begin(), t.end(), [&](Topic x) { return x >= n_valid_topics; }), t.end());
    return !t.empty();
  }
  // one iteration of user_receive_loop tasks/user/handler.rs:104-161.  <0 = the loop returns Err
  // (the caller, handle_user_connection :61-69, then removes the user).
  int user_receive(const Key& sender, const uint8_t* raw, uint32_t len) {
    capnp_lite::Message m;
    if (!capnp_lite::deserialize(raw, len, m)) return -7;
    std::vector<Topic> t(m.topics.begin(), m.topics.end());
    switch (m.kind) {
      case capnp_lite::Direct:
        handle_direct_message(Key(m.recipient.begin(), m.recipient.end()), raw, len, false);
        return 0;
      case capnp_lite::Broadcast:
        if (!prune(t)) return -8;
        handle_broadcast_message(t, raw, len, false);
        return 0;
      case capnp_lite::Subscribe:
        if (!prune(t)) return -8;
        subscribe_user_to(sender, t);
        return 0;
      case capnp_lite::Unsubscribe:
        if (!prune(t)) return -8;
        unsubscribe_user_from(sender, t);
        return 0;
      default:
        return -9;  // "invalid message received" :160
    }
  }
  // one iteration of broker_receive_loop tasks/broker/handler.rs:130-192 (Direct/Broadcast only)
  int broker_receive(const uint8_t* raw, uint32_t len) {
    capnp_lite::Message m;
    if (!capnp_lite::deserialize(raw, len, m)) return -7;
    if (m.kind == capnp_lite::Direct) {
      handle_direct_message(Key(m.recipient.begin(), m.recipient.end()), raw, len, true);
      return 0;
    }
    if (m.kind == capnp_lite::Broadcast) {
      std::vector<Topic> t(m.topics.begin(), m.topics.end());  // no prune :157
      handle_broadcast_message(t, raw, len, true);
      return 0;
    }
    return 1;  // sync kinds are applied through the explicit calls; others ignored :191
  }
};

inline std::vector<Topic> tv(const uint16_t* t, uint32_t n) { return std::vector<Topic>(t, t + n); }
inline Key kv(const uint8_t* k, uint32_t n) { return Key((const char*)k, n); }

// test-only standalone maps (relational_map.rs / versioned_map.rs unit tests use &str keys)
using StrRel = RelationalMap<std::string, uint64_t>;
using StrVer = VersionedMap<std::string, std::string, uint64_t>;

}  // namespace

// =============================================================================== C API (ctypes)
extern "C" {

void* orc_create(const char* identity, uint32_t n_valid_topics) {
  Oracle* o = new Oracle(identity);
  o->n_valid_topics = n_valid_topics;
  return o;
}
void orc_destroy(void* h) { delete (Oracle*)h; }

int orc_add_user(void* h, const uint8_t* k, uint32_t kl, const uint16_t* t, uint32_t n) {
  return ((Oracle*)h)->add_user(kv(k, kl), tv(t, n));
}
void orc_remove_user(void* h, const uint8_t* k, uint32_t kl) { ((Oracle*)h)->remove_user(kv(k, kl)); }
void orc_subscribe_user_to(void* h, const uint8_t* k, uint32_t kl, const uint16_t* t, uint32_t n) {
  ((Oracle*)h)->subscribe_user_to(kv(k, kl), tv(t, n));
}
void orc_unsubscribe_user_from(void* h, const uint8_t* k, uint32_t kl, const uint16_t* t, uint32_t n) {
  ((Oracle*)h)->unsubscribe_user_from(kv(k, kl), tv(t, n));
}
int orc_add_broker(void* h, const char* id) { return ((Oracle*)h)->add_broker(BrokerId::parse(id)); }
void orc_remove_broker(void* h, const char* id) { ((Oracle*)h)->remove_broker(BrokerId::parse(id)); }
void orc_subscribe_broker_to(void* h, const char* id, const uint16_t* t, uint32_t n) {
  ((Oracle*)h)->subscribe_broker_to(BrokerId::parse(id), tv(t, n));
}
void orc_unsubscribe_broker_from(void* h, const char* id, const uint16_t* t, uint32_t n) {
  ((Oracle*)h)->unsubscribe_broker_from(BrokerId::parse(id), tv(t, n));
}

// user sync: build a remote DirectMap incrementally, then apply
void* orc_dmap_new(const char* identity) { return new DirectMap(BrokerId::parse(identity)); }
void orc_dmap_free(void* m) { delete (DirectMap*)m; }
void orc_dmap_put(void* m, const uint8_t* k, uint32_t kl, uint64_t version, const char* owner) {
  DirectMap::VersionedValue v{version, std::nullopt};
  if (owner) v.value = BrokerId::parse(owner);
  ((DirectMap*)m)->underlying_map[kv(k, kl)] = v;
}
void orc_apply_user_sync(void* h, void* m) { ((Oracle*)h)->apply_user_sync(*(DirectMap*)m); }

// user sync between two oracles: get_full_user_sync mod.rs:131-137 / get_partial_user_sync :141-148
// (None when empty) → apply_user_sync :154.  Returns 0 when there was nothing to send.
int orc_user_sync(void* from, void* to, int full, int apply) {
  Oracle* f = (Oracle*)from;
  std::optional<DirectMap> m;
  if (full) { if (!f->direct_map.underlying_map.empty()) m = f->direct_map; }
  else { DirectMap d = f->direct_map.diff(); if (!d.is_empty()) m = d; }
  if (!m) return 0;
  if (apply) ((Oracle*)to)->apply_user_sync(*m);
  return 1;
}
// topic sync between two oracles (cdn-broker/src/connections/mod.rs:410-526 tests)
// mode 0 = partial, 1 = full.  Returns 0 if there was nothing to sync (None), 1 if applied.
int orc_topic_sync(void* from, void* to, const char* from_id_in_to, int mode, int apply) {
  Oracle* f = (Oracle*)from;
  std::optional<TopicSyncMap> m = mode == 0 ? f->get_partial_topic_sync() : f->get_full_topic_sync();
  if (!m) return 0;
  if (apply) ((Oracle*)to)->apply_topic_sync(BrokerId::parse(from_id_in_to), *m);
  return 1;
}
// direct TopicSyncMap delivery as the test harness does (tests/mod.rs:352-363): a fresh map with
// the given topics Subscribed, its diff applied for broker `id`
void orc_apply_topic_list(void* h, const char* id, const uint16_t* t, uint32_t n) {
  TopicSyncMap m(0);
  for (uint32_t i = 0; i < n; i++) m.insert(t[i], Subscribed);
  ((Oracle*)h)->apply_topic_sync(BrokerId::parse(id), m.diff());
}

void orc_handle_broadcast_message(void* h, const uint16_t* t, uint32_t n, const uint8_t* raw,
                                  uint32_t len, int to_users_only) {
  ((Oracle*)h)->handle_broadcast_message(tv(t, n), raw, len, to_users_only != 0);
}
void orc_handle_direct_message(void* h, const uint8_t* k, uint32_t kl, const uint8_t* raw,
                               uint32_t len, int to_user_only) {
  ((Oracle*)h)->handle_direct_message(kv(k, kl), raw, len, to_user_only != 0);
}
int orc_user_receive(void* h, const uint8_t* k, uint32_t kl, const uint8_t* raw, uint32_t len) {
  return ((Oracle*)h)->user_receive(kv(k, kl), raw, len);
}
int orc_broker_receive(void* h, const uint8_t* raw, uint32_t len) {
  return ((Oracle*)h)->broker_receive(raw, len);
}

// connection streams
uint32_t orc_num_conns(void* h) { return (uint32_t)((Oracle*)h)->conns.size(); }
uint64_t orc_stream_len(void* h, int c) { return ((Oracle*)h)->conns[c]->stream.size(); }
const uint8_t* orc_stream_ptr(void* h, int c) { return ((Oracle*)h)->conns[c]->stream.data(); }
uint32_t orc_stream_frames(void* h, int c) { return (uint32_t)((Oracle*)h)->conns[c]->frame_lens.size(); }
void orc_stream_clear(void* h, int c) {
  auto& cn = *((Oracle*)h)->conns[c];
  cn.stream.clear(); cn.frame_lens.clear();
}
void orc_stream_clear_all(void* h) {
  for (auto& c : ((Oracle*)h)->conns) { c->stream.clear(); c->frame_lens.clear(); }
}
void orc_close_conn(void* h, int c) { ((Oracle*)h)->conns[c]->closed = true; }
int orc_conn_removed(void* h, int c) { return ((Oracle*)h)->conns[c]->removed; }
int orc_user_conn(void* h, const uint8_t* k, uint32_t kl) {
  auto& u = ((Oracle*)h)->users;
  auto it = u.find(kv(k, kl));
  return it == u.end() ? -1 : it->second;
}
int orc_broker_conn(void* h, const char* id) {
  auto& b = ((Oracle*)h)->brokers;
  auto it = b.find(BrokerId::parse(id));
  return it == b.end() ? -1 : it->second.conn;
}
uint32_t orc_num_users(void* h) { return (uint32_t)((Oracle*)h)->users.size(); }
uint64_t orc_bytes_sent(void* h) { return ((Oracle*)h)->bytes_sent; }
uint64_t orc_deliveries(void* h) { return ((Oracle*)h)->deliveries; }
// get_interested_by_topic as conn ids (users then brokers); returns count
uint32_t orc_interested(void* h, const uint16_t* t, uint32_t n, int to_users_only, int* out, uint32_t cap) {
  Oracle* o = (Oracle*)h;
  std::vector<BrokerId> ib; std::vector<Key> iu;
  o->get_interested_by_topic(tv(t, n), to_users_only != 0, ib, iu);
  uint32_t k = 0;
  for (auto& u : iu) { auto it = o->users.find(u); if (it != o->users.end() && k < cap) out[k++] = it->second; }
  for (auto& b : ib) { auto it = o->brokers.find(b); if (it != o->brokers.end() && k < cap) out[k++] = it->second.conn; }
  return k;
}
// number of (brokers, user keys) returned by get_interested_by_topic — incl. connection-less keys
void orc_interested_counts(void* h, const uint16_t* t, uint32_t n, int to_users_only, uint32_t* nb, uint32_t* nu) {
  std::vector<BrokerId> ib; std::vector<Key> iu;
  ((Oracle*)h)->get_interested_by_topic(tv(t, n), to_users_only != 0, ib, iu);
  *nb = (uint32_t)ib.size(); *nu = (uint32_t)iu.size();
}
// route of a key: 0 none, 1 local user (conn), 2 remote broker (conn or -1)
int orc_route(void* h, const uint8_t* k, uint32_t kl, int* conn) {
  Oracle* o = (Oracle*)h;
  const BrokerId* b = o->get_broker_identifier_of_user(kv(k, kl));
  *conn = -1;
  if (!b) return 0;
  if (*b == o->identity) {
    auto it = o->users.find(kv(k, kl));
    if (it == o->users.end()) return 0;
    *conn = it->second; return 1;
  }
  auto it = o->brokers.find(*b);
  if (it != o->brokers.end()) *conn = it->second.conn;
  return 2;
}

// ---- capnp-lite ------------------------------------------------------------------------------
// serialize kind with (f0 = topics|recipient, payload); returns length, writes up to cap bytes
uint64_t orc_serialize(int kind, const uint8_t* f0, uint32_t f0_len, const uint8_t* payload,
                       uint32_t payload_len, uint8_t* out, uint64_t cap) {
  capnp_lite::Message m;
  m.kind = kind;
  if (kind == capnp_lite::Direct) m.recipient.assign(f0, f0 + f0_len);
  else m.topics.assign(f0, f0 + f0_len);
  m.payload.assign(payload, payload + payload_len);
  auto b = capnp_lite::serialize(m);
  if (b.size() <= cap) std::memcpy(out, b.data(), b.size());
  return b.size();
}
// deserialize: returns kind or -1; copies field0/payload (caps) and reports their lengths
int orc_deserialize(const uint8_t* raw, uint64_t len, uint8_t* f0, uint32_t f0_cap, uint32_t* f0_len,
                    uint8_t* payload, uint32_t p_cap, uint32_t* p_len) {
  capnp_lite::Message m;
  if (!capnp_lite::deserialize(raw, len, m)) return -1;
  const capnp_lite::Bytes& a = (m.kind == capnp_lite::Direct) ? m.recipient : m.topics;
  *f0_len = (uint32_t)a.size(); *p_len = (uint32_t)m.payload.size();
  if (a.size() <= f0_cap && !a.empty()) std::memcpy(f0, a.data(), a.size());
  if (m.payload.size() <= p_cap && !m.payload.empty()) std::memcpy(payload, m.payload.data(), m.payload.size());
  return m.kind;
}

// ---- standalone map handles for the reference's unit tests ------------------------------------
void* orc_rel_new() { return new StrRel(); }
void orc_rel_free(void* r) { delete (StrRel*)r; }
void orc_rel_assoc(void* r, const char* k, const uint64_t* v, uint32_t n) {
  ((StrRel*)r)->associate_key_with_values(k, std::vector<uint64_t>(v, v + n));
}
void orc_rel_dissoc(void* r, const char* k, const uint64_t* v, uint32_t n) {
  ((StrRel*)r)->dissociate_keys_from_value(k, std::vector<uint64_t>(v, v + n));
}
void orc_rel_remove_key(void* r, const char* k) { ((StrRel*)r)->remove_key(k); }
// keys by value joined with '\n' into out; returns count
uint32_t orc_rel_keys_by_value(void* r, uint64_t v, char* out, uint32_t cap) {
  auto ks = ((StrRel*)r)->get_keys_by_value(v);
  std::sort(ks.begin(), ks.end());
  std::string s;
  for (auto& k : ks) { s += k; s += '\n'; }
  if (s.size() < cap) std::memcpy(out, s.c_str(), s.size() + 1);
  return (uint32_t)ks.size();
}
uint32_t orc_rel_values(void* r, uint64_t* out, uint32_t cap) {
  auto vs = ((StrRel*)r)->get_values();
  std::sort(vs.begin(), vs.end());
  for (uint32_t i = 0; i < vs.size() && i < cap; i++) out[i] = vs[i];
  return (uint32_t)vs.size();
}
uint32_t orc_rel_values_of_key(void* r, const char* k, uint64_t* out, uint32_t cap) {
  auto& m = ((StrRel*)r)->key_to_values;
  auto it = m.find(k);
  if (it == m.end()) return 0xFFFFFFFFu;
  std::vector<uint64_t> vs(it->second.begin(), it->second.end());
  std::sort(vs.begin(), vs.end());
  for (uint32_t i = 0; i < vs.size() && i < cap; i++) out[i] = vs[i];
  return (uint32_t)vs.size();
}
uint32_t orc_rel_num_keys(void* r) { return (uint32_t)((StrRel*)r)->key_to_values.size(); }
uint32_t orc_rel_num_values(void* r) { return (uint32_t)((StrRel*)r)->value_to_keys.size(); }

void* orc_ver_new(uint64_t identity) { return new StrVer(identity); }
void orc_ver_free(void* v) { delete (StrVer*)v; }
void orc_ver_insert(void* v, const char* k, const char* val) { ((StrVer*)v)->insert(k, val); }
void orc_ver_remove(void* v, const char* k) { ((StrVer*)v)->remove(k); }
// get: returns 1 and copies value, or 0
int orc_ver_get(void* v, const char* k, char* out, uint32_t cap) {
  const std::string* s = ((StrVer*)v)->get(k);
  if (!s) return 0;
  if (s->size() < cap) std::memcpy(out, s->c_str(), s->size() + 1);
  return 1;
}
void* orc_ver_get_full(void* v) { return new StrVer(((StrVer*)v)->get_full()); }
void* orc_ver_diff(void* v) { return new StrVer(((StrVer*)v)->diff()); }
uint32_t orc_ver_merge(void* v, void* remote) { return (uint32_t)((StrVer*)v)->merge(*(StrVer*)remote).size(); }
void orc_ver_remove_by_value_no_modify(void* v, const char* val) { ((StrVer*)v)->remove_by_value_no_modify(val); }
uint32_t orc_ver_len(void* v) { return (uint32_t)((StrVer*)v)->underlying_map.size(); }

}  // extern "C"
