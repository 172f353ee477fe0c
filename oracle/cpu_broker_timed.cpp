// ORACLE — test/bench infrastructure only (CPU baseline; never linked into the product).
//
// cpu_broker_timed.cpp — timed C++ restatement of the reference cdn-broker's CPU broadcast path,
// used as bench.py's `cpu_baseline` and `--impl reference` arm because the reference itself (Rust)
// cannot be built here.  It reproduces the *shape* of the reference's work per message:
//
//  stage 1  Connections::get_interested_by_topic (cdn-broker/src/connections/mod.rs:94-124):
//           RelationalMap::get_keys_by_value clones every subscribed key (Arc bump) into a Vec
//           (relational_map.rs:39-47), each is inserted into a HashSet<UserPublicKey> (std SipHash-1-3
//           over the key bytes), the set is collected into a Vec.
//  stage 2  handle_broadcast_message's sequential loop (tasks/broker/handler.rs:268-271): per
//           recipient try_send_to_user = HashMap<UserPublicKey,Connection> probe (SipHash) +
//           Connection clone + Bytes clone (2 Arc bumps) + unbounded channel push
//           (tasks/user/sender.rs:16-32, cdn-proto/src/connection/protocols/mod.rs:239-251).
//           One receive loop handles one sender's messages sequentially; different senders run on
//           different worker threads — here: one thread per in-flight message, all host threads.
//  stage 3  per-connection writer tasks (protocols/mod.rs:156-186,354-394): pop, write u32 BE length,
//           write the bytes (= one memcpy of the frame per recipient into that connection's buffer),
//           on all threads that are not running a receive loop, CONCURRENTLY with stages 1+2 (the
//           reference's writer tasks run on the same tokio pool as the receive loops).
//  Threads are started once (persistent pool); a step is timed as a whole (wall clock); the total
//  over the timed steps and the median step are both reported.
//
// It is deliberately generous to the CPU: no tokio scheduling, no syscalls/TLS, no allocator
// contention between stages, perfect static load balance.
//
// usage: cpu_broker_timed <n_conns> <payload_bytes> <msgs_per_step> <steps> <warmup> <threads> [model 0|1]
// prints one JSON object.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

// SipHash-1-3 (Rust's DefaultHasher) over a byte string
static inline uint64_t rotl(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
struct Sip13 {
  uint64_t k0 = 0x0706050403020100ULL, k1 = 0x0f0e0d0c0b0a0908ULL;
  size_t operator()(const std::shared_ptr<const std::string>& sp) const {
    const std::string& s = *sp;
    uint64_t v0 = k0 ^ 0x736f6d6570736575ULL, v1 = k1 ^ 0x646f72616e646f6dULL, v2 = k0 ^ 0x6c7967656e657261ULL,
             v3 = k1 ^ 0x7465646279746573ULL;
    auto round = [&]() {
      v0 += v1; v1 = rotl(v1, 13); v1 ^= v0; v0 = rotl(v0, 32);
      v2 += v3; v3 = rotl(v3, 16); v3 ^= v2;
      v0 += v3; v3 = rotl(v3, 21); v3 ^= v0;
      v2 += v1; v1 = rotl(v1, 17); v1 ^= v2; v2 = rotl(v2, 32);
    };
    const uint8_t* p = (const uint8_t*)s.data();
    size_t n = s.size(), i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t m; memcpy(&m, p + i, 8); v3 ^= m; round(); v0 ^= m; }
    uint64_t b = (uint64_t)n << 56;
    for (size_t j = 0; i + j < n; j++) b |= (uint64_t)p[i + j] << (8 * j);
    v3 ^= b; round(); v0 ^= b;
    v2 ^= 0xff; round(); round(); round();
    return (size_t)(v0 ^ v1 ^ v2 ^ v3);
  }
};
struct KeyEq {
  bool operator()(const std::shared_ptr<const std::string>& a, const std::shared_ptr<const std::string>& b) const {
    return *a == *b;
  }
};
using Key = std::shared_ptr<const std::string>;       // UserPublicKey = Arc<Vec<u8>>
using Bytes = std::shared_ptr<const std::vector<uint8_t>>;  // Bytes = Arc<Vec<u8>> (+permit)

struct Connection {  // the sending half: an unbounded MPSC queue
  std::mutex mu;
  std::vector<Bytes> q;
};

int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: %s n_conns payload msgs_per_step steps warmup threads\n", argv[0]); return 2; }
  const size_t N = strtoull(argv[1], 0, 10), K = strtoull(argv[2], 0, 10), M = strtoull(argv[3], 0, 10);
  const int steps = atoi(argv[4]), warmup = atoi(argv[5]);
  int T = atoi(argv[6]);
  // model 0 = "phases": all threads run the receive loops, then all threads run the writer tasks
  //           (partitioned); model 1 = "overlap": writer tasks run concurrently with the receive loops
  //           and are woken per connection.  bench.py calibrates both and reports the FASTER one.
  const int model = argc > 7 ? atoi(argv[7]) : 1;
  if (T <= 0) T = (int)std::thread::hardware_concurrency();
  if (T <= 0) T = 1;
  const size_t L = 8 * (6 + 1 + (K + 7) / 8), F = 4 + L;  // single-segment Broadcast, one topic
  const size_t slot = (F + 63) / 64 * 64, depth = 2;

  // state: users map, topic 0 → set of keys (all subscribed), 32-byte keys
  std::unordered_map<Key, std::shared_ptr<Connection>, Sip13, KeyEq> users;
  std::unordered_set<Key, Sip13, KeyEq> topic0;
  std::vector<std::shared_ptr<Connection>> conns(N);
  users.reserve(N * 2); topic0.reserve(N * 2);
  for (size_t i = 0; i < N; i++) {
    std::string k(32, 0);
    uint64_t x = i * 0x9E3779B97F4A7C15ULL + 1;
    for (int j = 0; j < 4; j++) { x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ULL; memcpy(&k[j * 8], &x, 8); }
    Key key = std::make_shared<const std::string>(std::move(k));
    conns[i] = std::make_shared<Connection>();
    users.emplace(key, conns[i]);
    topic0.insert(key);
  }
  std::vector<uint8_t> out(N * slot * depth);  // per-connection output buffers (stand-in for sockets)
  std::vector<uint32_t> wr(N, 0);

  std::vector<Bytes> msgs(M);
  for (size_t m = 0; m < M; m++) {
    auto v = std::make_shared<std::vector<uint8_t>>(L);
    for (size_t i = 0; i < L; i++) (*v)[i] = (uint8_t)(i * 131 + m);
    msgs[m] = v;
  }

  // Persistent worker threads (the reference's tokio runtime is started once), and the writer tasks
  // run CONCURRENTLY with the receive loops: routers = one thread per in-flight message (a receive
  // loop handles its sender's messages sequentially), every other thread runs writer tasks.  A writer
  // task is WOKEN when its connection's queue becomes non-empty (the first push after it went idle
  // schedules the connection on its writer thread's ready list, like tokio waking the task), drains
  // the queue, and goes idle again — writers never poll connections that have nothing queued.
  const int R = model == 0 ? T : (int)std::min<size_t>(M, (size_t)std::max(1, T / 2));  // router threads
  const int Wt = model == 0 ? T : std::max(1, T - R);                                   // writer threads
  struct Step { std::atomic<size_t> next{0}; std::atomic<int> routers_left{0}; };
  Step st;
  struct Ready { std::mutex mu; std::condition_variable cv; std::vector<uint32_t> q; char pad[64]; };
  std::vector<Ready> ready(Wt);
  std::vector<std::atomic<uint8_t>> scheduled(N);
  for (auto& x : scheduled) x.store(0);
  std::vector<uint32_t> conn_index_of;  // Connection* → index
  std::unordered_map<const Connection*, uint32_t> cidx;
  cidx.reserve(N * 2);
  for (size_t i = 0; i < N; i++) cidx.emplace(conns[i].get(), (uint32_t)i);
  std::atomic<int> phase{0};          // bumped by the main thread to start a step
  std::atomic<int> done{0};
  std::atomic<bool> quit{false};
  std::vector<uint64_t> cs(T, 0), dl(T, 0), by(T, 0);
  std::vector<double> router_busy(T, 0), writer_busy(T, 0);
  auto now = [] { return std::chrono::steady_clock::now(); };
  // the users map hands out the connection; its index (for the writer partition) rides along
  std::unordered_map<Key, std::pair<std::shared_ptr<Connection>, uint32_t>, Sip13, KeyEq> users_ix;
  users_ix.reserve(N * 2);
  for (auto& kv : users) users_ix.emplace(kv.first, std::make_pair(kv.second, cidx[kv.second.get()]));

  auto router = [&](int t) {
    auto a = now();
    for (;;) {
      size_t m = st.next.fetch_add(1);
      if (m >= M) break;
      // stage 1
      std::vector<Key> cloned;
      cloned.reserve(topic0.size());
      for (const Key& k : topic0) cloned.push_back(k);            // get_keys_by_value: clone
      std::unordered_set<Key, Sip13, KeyEq> recipients;
      for (Key& k : cloned) recipients.insert(std::move(k));     // HashSet insert
      std::vector<Key> list(recipients.begin(), recipients.end());  // into_iter().collect()
      // stage 2
      for (const Key& k : list) {
        auto it2 = users_ix.find(k);                               // get_user_connection
        if (it2 == users_ix.end()) continue;
        std::shared_ptr<Connection> c = it2->second.first;         // Connection clone
        const uint32_t ci = it2->second.second;
        Bytes b = msgs[m];                                         // message.clone()
        {
          std::lock_guard<std::mutex> g(c->mu);
          c->q.push_back(std::move(b));                            // send_message_raw
        }
        if (model == 1 && !scheduled[ci].exchange(1, std::memory_order_acq_rel)) {   // wake the writer task
          Ready& r = ready[(size_t)ci * Wt / N];
          bool was_empty;
          {
            std::lock_guard<std::mutex> g(r.mu);
            was_empty = r.q.empty();
            r.q.push_back(ci);
          }
          if (was_empty) r.cv.notify_one();   // the writer thread sleeps while it has nothing to do
        }
      }
    }
    if (st.routers_left.fetch_sub(1) == 1 && model == 1)
      for (auto& r : ready) { std::lock_guard<std::mutex> g(r.mu); r.cv.notify_all(); }   // let idle writers see the end of the step
    router_busy[t] += std::chrono::duration<double>(now() - a).count();
  };
  auto writer = [&](int t, int wi) {
    auto a = now();
    Ready& r = ready[wi];
    std::vector<uint32_t> woken;
    std::vector<Bytes> take;
    for (;;) {
      bool last;
      {
        std::unique_lock<std::mutex> g(r.mu);
        r.cv.wait(g, [&] { return !r.q.empty() || st.routers_left.load() == 0; });   // blocked, not spinning: idle threads must not steal SMT cycles
        last = st.routers_left.load() == 0;   // read BEFORE taking the list: a list taken after the routers finished is complete
        woken.swap(r.q);
      }
      for (uint32_t c : woken) {
        Connection& cn = *conns[c];
        scheduled[c].store(0, std::memory_order_release);   // a push from now on wakes the task again
        {
          std::lock_guard<std::mutex> g(cn.mu);
          take.swap(cn.q);
        }
        for (Bytes& msg : take) {
          uint8_t* dst = &out[((size_t)c * depth + (wr[c]++ % depth)) * slot];
          uint32_t len = (uint32_t)msg->size();
          dst[0] = len >> 24; dst[1] = len >> 16; dst[2] = len >> 8; dst[3] = len;  // write_u32 (BE)
          memcpy(dst + 4, msg->data(), len);                                          // write_all
          cs[t] += dst[4 + (len >> 1)];
          dl[t]++; by[t] += 4 + len;
        }
        take.clear();
      }
      const bool idle = woken.empty();
      woken.clear();
      if (last && idle) break;
    }
    writer_busy[t] += std::chrono::duration<double>(now() - a).count();
  };
  // model 0: stage 3 after a barrier — every thread walks its partition of the connections
  std::atomic<int> routed{0};
  std::mutex gm;
  std::condition_variable gcv, bcv, dcv;   // step start, stage barrier, step done
  auto phase_writer = [&](int t) {
    auto a = now();
    const size_t lo = N * (size_t)t / T, hi = N * (size_t)(t + 1) / T;
    for (size_t c = lo; c < hi; c++) {
      Connection& cn = *conns[c];
      for (Bytes& msg : cn.q) {
        uint8_t* dst = &out[(c * depth + (wr[c]++ % depth)) * slot];
        uint32_t len = (uint32_t)msg->size();
        dst[0] = len >> 24; dst[1] = len >> 16; dst[2] = len >> 8; dst[3] = len;  // write_u32 (BE)
        memcpy(dst + 4, msg->data(), len);                                          // write_all
        cs[t] += dst[4 + (len >> 1)];
        dl[t]++; by[t] += 4 + len;
      }
      cn.q.clear();
    }
    writer_busy[t] += std::chrono::duration<double>(now() - a).count();
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < T; t++)
    pool.emplace_back([&, t] {
      int seen = 0;
      for (;;) {
        {
          std::unique_lock<std::mutex> g(gm);
          gcv.wait(g, [&] { return phase.load() != seen || quit.load(); });
        }
        if (quit.load()) return;
        seen = phase.load();
        if (model == 0) {
          router(t);
          {   // barrier between the stages (threads without a message to route sleep here)
            std::unique_lock<std::mutex> g(gm);
            if (routed.fetch_add(1) + 1 == T) bcv.notify_all();
            else bcv.wait(g, [&] { return routed.load() >= T; });
          }
          phase_writer(t);
        } else if (t < R) router(t); else writer(t, t - R);
        {
          std::lock_guard<std::mutex> g(gm);
          if (done.fetch_add(1) + 1 == T) dcv.notify_all();
        }
      }
    });

  std::vector<double> step_s;
  uint64_t deliveries = 0, bytes = 0, checksum = 0;
  double sec = 0;
  for (int it = 0; it < warmup + steps; it++) {
    if (it == warmup) {
      for (int t = 0; t < T; t++) { dl[t] = by[t] = cs[t] = 0; router_busy[t] = writer_busy[t] = 0; }
    }
    st.next.store(0); st.routers_left.store(R); done.store(0); routed.store(0);
    auto a = now();
    {
      std::lock_guard<std::mutex> g(gm);
      phase.fetch_add(1);
    }
    gcv.notify_all();
    {
      std::unique_lock<std::mutex> g(gm);
      dcv.wait(g, [&] { return done.load() >= T; });
    }
    const double d = std::chrono::duration<double>(now() - a).count();
    if (it >= warmup) { step_s.push_back(d); sec += d; }
  }
  {
    std::lock_guard<std::mutex> g(gm);
    quit.store(true);
  }
  gcv.notify_all();
  for (auto& x : pool) x.join();
  for (int t = 0; t < T; t++) { deliveries += dl[t]; bytes += by[t]; checksum += cs[t]; }
  double t12 = 0, t3 = 0;
  for (int t = 0; t < T; t++) { t12 = std::max(t12, router_busy[t]); t3 = std::max(t3, writer_busy[t]); }
  std::vector<double> sorted = step_s;
  std::sort(sorted.begin(), sorted.end());
  const double med = sorted.empty() ? 0 : sorted[sorted.size() / 2];
  const double step_bytes = steps > 0 ? (double)bytes / steps : 0;
  printf("{\"n_conns\": %zu, \"payload\": %zu, \"frame_bytes\": %zu, \"msgs_per_step\": %zu, \"steps\": %d, \"threads\": %d, "
         "\"router_threads\": %d, \"writer_threads\": %d, "
         "\"deliveries\": %llu, \"bytes\": %llu, \"seconds\": %.6f, \"stage12_s\": %.6f, \"stage3_s\": %.6f, "
         "\"gbps\": %.4f, \"gbps_median_step\": %.4f, \"median_step_s\": %.6f, \"deliveries_per_s\": %.1f, \"checksum\": %llu, "
         "\"model\": \"%s\"}\n",
         N, K, F, M, steps, T, R, Wt, (unsigned long long)deliveries, (unsigned long long)bytes, sec, t12, t3,
         bytes / sec / 1e9, med > 0 ? step_bytes / med / 1e9 : 0.0, med, deliveries / sec, (unsigned long long)checksum,
         model == 0 ? "persistent threads; receive loops then writer tasks (two phases)" : "persistent threads; writer tasks woken per connection, concurrent with the receive loops");
  return 0;
}
