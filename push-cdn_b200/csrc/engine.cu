// engine.cu — host runtime of the fan-out engine and the C ABI (include/pcdn_fanout.h).
//
// One engine = one CUDA device, one stream, the routing tables (host mirror in host_state.*, device
// copy in DevState), the per-connection output rings and a small pool of batch slots.  A batch is
// staged in pinned memory while it is open, copied to the device on flush and routed by the kernel
// pipeline of kernels.cuh; results (span table, counters) come back through pinned memory.
// There is no CPU data path: without a device every routing call fails with PCDN_ENODEV.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "frame_parse.h"
#include "host_state.h"
#include "kernels.cuh"
#include "pcdn_fanout.h"

using namespace pcdn;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return fail(PCDN_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));             \
  } while (0)

template <class T>
int dev_alloc(T** p, size_t n) {
  *p = nullptr;
  if (!n) n = 1;
  cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
  if (e != cudaSuccess) return fail(PCDN_ENOMEM, std::string("cudaMalloc ") + std::to_string(n * sizeof(T)) + " B: " + cudaGetErrorString(e));
  return 0;
}
template <class T>
int pin_alloc(T** p, size_t n) {
  *p = nullptr;
  if (!n) n = 1;
  cudaError_t e = cudaMallocHost((void**)p, n * sizeof(T));
  if (e != cudaSuccess) return fail(PCDN_ENOMEM, std::string("cudaMallocHost ") + std::to_string(n * sizeof(T)) + " B: " + cudaGetErrorString(e));
  return 0;
}

// pinned + mapped: the device writes through *dev_alias (same bytes the host reads through *p)
template <typename T>
int pin_alloc_mapped(T** p, T** dev_alias, size_t n) {
  *p = nullptr;
  if (!n) n = 1;
  cudaError_t e = cudaHostAlloc((void**)p, n * sizeof(T), cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)dev_alias, (void*)*p, 0);
  if (e != cudaSuccess) return fail(PCDN_ENOMEM, std::string("cudaHostAlloc(mapped) ") + std::to_string(n * sizeof(T)) + " B: " + cudaGetErrorString(e));
  return 0;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

enum SlotState { SLOT_FREE = 0, SLOT_OPEN = 1, SLOT_INFLIGHT = 2 };
// engines up to this many connection slots publish spans directly into mapped host memory
constexpr uint32_t kDirectPublishMaxConns = 65536;  // = kSmallCtrlConns: the engines the fused control kernel serves

struct Slot {
  int state = SLOT_FREE;
  uint64_t batch_id = 0;
  bool polled = false, device_input = false;
  // host staging while open
  uint8_t* h_arena = nullptr;   // pinned
  size_t arena_used = 0;
  std::vector<uint8_t> kind, flags;
  std::vector<uint32_t> slot_off16, raw_len, aux_off, aux_len, bcast_index;
  std::vector<uint16_t> topics;
  uint32_t n_direct = 0;
  uint64_t ingress_bytes = 0;   // pool permits held by this batch
  std::chrono::steady_clock::time_point t_launch;
  bool devparse = false;        // some messages carry MSGF_DEVPARSE (k_parse runs first)
  int8_t* h_msg_status = nullptr;  // pinned
  uint32_t n_msg_errors = 0;
  uint8_t* h_desc = nullptr;    // pinned descriptor block
  // device
  uint8_t* d_arena = nullptr;
  uint8_t* d_desc = nullptr;
  Work w{};
  BatchIn in{};
  // results
  BatchStats* h_stats = nullptr;  // pinned: final counters (after the pack)
  BatchStats* d_stats_pub = nullptr;  // direct publish: device alias of h_stats (mapped)
  // span table / overflow list of this batch: written by the device straight into mapped host memory
  // (few spans expected) or staged in HBM and copied out while the pack runs (up to 2 per connection)
  bool spans_mapped = false;
  Span* d_spans_map = nullptr; Span* d_spans_dev = nullptr;
  uint32_t* d_ovf_map = nullptr; uint32_t* d_ovf_dev = nullptr;
  BatchStats* h_early = nullptr;  // pinned: counters as of k_offsets (n_spans, n_overflow are final there)
  Span* h_spans = nullptr;        // pinned
  uint32_t* h_overflow = nullptr; // pinned
  cudaEvent_t ev_done = nullptr;   // pack + final counters complete (pack stream)
  cudaEvent_t ev_ctrl = nullptr;   // match/plan/offsets complete (main stream)
  cudaEvent_t ev_early = nullptr;  // early counters are in h_early (copy stream)
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool timed = false;
};

}  // namespace

struct pcdn_engine {
  std::mutex mu;
  pcdn_config cfg{};
  std::string identity;
  Geometry geo{};
  std::unique_ptr<HostTables> tables;
  std::unique_ptr<Connections> conns;
  bool has_device = false;
  bool direct_publish = false;  // spans / overflow list written by the device into mapped host memory
  uint8_t* h_rings = nullptr;   // PCDN_FLAG_HOST_RINGS: host address of the (mapped, pinned) rings
  int n_sms = 148;
  // main stream: uploads, table updates, direct/match/plan/offsets, release.  pack stream: k_pack, so
  // that the control kernels of batch n+1 overlap the HBM-bound pack of batch n.  copy stream: D2H.
  cudaStream_t stream = nullptr, pack_stream = nullptr, copy_stream = nullptr;
  bool own_stream = false;
  DevState dev{};
  std::vector<Slot> slots;
  int open_slot = -1;
  uint64_t next_batch_id = 1;
  std::vector<uint64_t> inflight;  // submit order
  size_t desc_cap = 0, topics_cap = 0;
  // journal staging (pinned + device), reuse guarded by an event
  uint8_t* jstage_h = nullptr; uint8_t* jstage_d = nullptr; size_t jstage_cap = 0;
  cudaEvent_t ev_journal = nullptr; bool ev_journal_pending = false;
  std::vector<Upd32> h_u32; std::vector<UpdSlot> h_slot; std::vector<uint32_t> h_kslot; std::vector<uint8_t> h_kbytes;
  bool timing = false;
  uint64_t inflight_bytes = 0;  // Limiter analogue: accepted frame bytes whose batch is not released yet
  pcdn_stats stats{};
  std::vector<void*> dev_allocs, pin_allocs;
  // buffers behind pcdn_get_*_sync
  std::vector<UserSyncEntry> sync_users;
  std::vector<pcdn_user_sync_entry> sync_users_c;
  std::vector<TopicSyncEntry> sync_topics;
  std::vector<pcdn_topic_sync_entry> sync_topics_c;
};

namespace {

// Upload changed table words/slots/keys and apply them on the engine stream (K4).  Stream order
// gives R12: every earlier batch sees the old tables, every later batch the new ones.
int flush_journal(pcdn_engine* e) {
  HostTables& t = *e->tables;
  if (!e->has_device) { t.clear_dirty(); return 0; }
  const Geometry& g = e->geo;
  bool any = !t.dirty_sub.empty() || !t.dirty_brk.empty() || !t.dirty_owner.empty() || !t.dirty_slots.empty() ||
             !t.dirty_keys.empty();
  if (!any) return 0;
  cudaStream_t st = e->stream;
  // keys first (slots reference them)
  const bool full_keys = t.dirty_keys.size() > (size_t)g.max_keys / 16 + 64;
  if (full_keys) CUDA_TRY(cudaMemcpyAsync(e->dev.keys, t.keys.data(), t.keys.size(), cudaMemcpyHostToDevice, st));
  e->h_u32.clear(); e->h_slot.clear(); e->h_kslot.clear(); e->h_kbytes.clear();
  bool full_sub = t.dirty_sub.size() > t.sub.size() / 16 + 64;
  if (full_sub) CUDA_TRY(cudaMemcpyAsync(e->dev.sub, t.sub.data(), t.sub.size() * 4, cudaMemcpyHostToDevice, st));
  else for (uint32_t i : t.dirty_sub) e->h_u32.push_back(Upd32{0, i, t.sub[i]});
  for (uint32_t i : t.dirty_brk) e->h_u32.push_back(Upd32{1, i, t.brk[i]});
  for (uint32_t i : t.dirty_owner) e->h_u32.push_back(Upd32{2, i, t.owner_conn[i]});
  bool full_slots = t.dirty_slots.size() > t.cuckoo.size() / 16 + 64;
  if (full_slots) CUDA_TRY(cudaMemcpyAsync(e->dev.cuckoo, t.cuckoo.data(), t.cuckoo.size() * sizeof(CuckooEntry), cudaMemcpyHostToDevice, st));
  else for (uint32_t i : t.dirty_slots) e->h_slot.push_back(UpdSlot{i, t.cuckoo[i]});
  if (!full_keys) for (uint32_t k : t.dirty_keys) {
    e->h_kslot.push_back(k);
    size_t at = e->h_kbytes.size();
    e->h_kbytes.resize(at + g.key_stride);
    std::memcpy(&e->h_kbytes[at], &t.keys[(size_t)k * g.key_stride], g.key_stride);
  }
  // One pinned staging block [Upd32 | UpdSlot | key slots | key bytes] → one H2D copy → apply kernels.
  // The staging block is reused by the next flush; an event (not a stream sync) guards it, so table
  // churn at control-plane rate never stalls the batches already queued on the stream.
  const size_t b_u32 = align_up(e->h_u32.size() * sizeof(Upd32), 16), b_slot = align_up(e->h_slot.size() * sizeof(UpdSlot), 16);
  const size_t b_ks = align_up(e->h_kslot.size() * 4, 16), b_kb = align_up(e->h_kbytes.size(), 16);
  const size_t total = b_u32 + b_slot + b_ks + b_kb;
  if (total) {
    if (e->ev_journal_pending) { CUDA_TRY(cudaEventSynchronize(e->ev_journal)); e->ev_journal_pending = false; }
    if (total > e->jstage_cap) {
      const size_t ncap = std::max(total, e->jstage_cap * 2 + (1 << 16));
      CUDA_TRY(cudaStreamSynchronize(st));
      if (e->jstage_h) cudaFreeHost(e->jstage_h);
      if (e->jstage_d) cudaFree(e->jstage_d);
      e->jstage_h = nullptr; e->jstage_d = nullptr; e->jstage_cap = 0;
      CUDA_TRY(cudaMallocHost((void**)&e->jstage_h, ncap));
      CUDA_TRY(cudaMalloc((void**)&e->jstage_d, ncap));
      e->jstage_cap = ncap;
    }
    uint8_t* h = e->jstage_h;
    if (b_u32) std::memcpy(h, e->h_u32.data(), e->h_u32.size() * sizeof(Upd32));
    if (b_slot) std::memcpy(h + b_u32, e->h_slot.data(), e->h_slot.size() * sizeof(UpdSlot));
    if (b_ks) std::memcpy(h + b_u32 + b_slot, e->h_kslot.data(), e->h_kslot.size() * 4);
    if (b_kb) std::memcpy(h + b_u32 + b_slot + b_ks, e->h_kbytes.data(), e->h_kbytes.size());
    CUDA_TRY(cudaMemcpyAsync(e->jstage_d, h, total, cudaMemcpyHostToDevice, st));
    uint8_t* d = e->jstage_d;
    launch_apply_updates(e->dev, (const Upd32*)d, (uint32_t)e->h_u32.size(), (const UpdSlot*)(d + b_u32), (uint32_t)e->h_slot.size(),
                         (const uint32_t*)(d + b_u32 + b_slot), d + b_u32 + b_slot + b_ks, (uint32_t)e->h_kslot.size(), st);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(e->ev_journal, st));
    e->ev_journal_pending = true;
  }
  // whole-table uploads come from pageable vectors (staged by the runtime before the call returns);
  // they only happen on bulk loads, where one synchronisation is irrelevant
  if (full_keys || full_sub || full_slots) CUDA_TRY(cudaStreamSynchronize(st));
  t.clear_dirty();
  return 0;
}

void slot_reset_open(Slot& s) {
  s.arena_used = 0; s.n_direct = 0; s.devparse = false; s.n_msg_errors = 0; s.ingress_bytes = 0;
  s.kind.clear(); s.flags.clear(); s.slot_off16.clear(); s.raw_len.clear(); s.aux_off.clear(); s.aux_len.clear();
  s.bcast_index.clear(); s.topics.clear();
  s.polled = false; s.device_input = false; s.timed = false;
}

int acquire_open_slot(pcdn_engine* e) {
  if (e->open_slot >= 0) return 0;
  for (size_t i = 0; i < e->slots.size(); i++)
    if (e->slots[i].state == SLOT_FREE) {
      e->open_slot = (int)i;
      e->slots[i].state = SLOT_OPEN;
      slot_reset_open(e->slots[i]);
      return 0;
    }
  return fail(PCDN_EAGAIN, "all batch slots are in flight: poll and release a batch first");
}

// run the kernel pipeline for slot `s` whose BatchIn is ready on the device
int launch_pipeline(pcdn_engine* e, Slot& s, uint32_t n_direct) {
  // Default: the pack runs on the main stream.  A/B switch (pack_variant bit 3): run it on the
  // high-priority pack stream so the next batch's control kernels overlap it — measured SLOWER for
  // the bulk-store pack (profiles/r1_sweep_overlap.txt), so it stays opt-in.
  const bool dp = e->direct_publish;
  cudaStream_t st = e->stream, ps = (!dp && (e->cfg.pack_variant & 8)) ? e->pack_stream : e->stream, cs = e->copy_stream;
  const bool has_direct = n_direct > 0;
  s.timed = e->timing;
  if (++s.w.stamp == 0) s.w.stamp = 1;  // validity stamp of this batch's direct buckets
  // latency path of the smallest geometry: match + plan + offsets in one cluster launch that also
  // zeroes / publishes the counters (kernels.cu: k_ctrl_small)
  const bool fused = dp && e->geo.N <= kSmallCtrlConns && s.in.n_msgs <= kSmallCtrlMsgs;
  // Spans go straight into mapped host memory when few are expected (16-byte PCIe writes: a table
  // of 16 K spans measured 20 us slower than the staged copy): the smallest geometry, or a batch
  // without broadcasts and with few messages (at most one span per message).  Otherwise they are staged in HBM and copied
  // out with one DMA of the exact size while the pack runs.
  s.spans_mapped = dp && (e->geo.N <= 8192 || (s.in.n_bcast == 0 && s.in.n_msgs <= 4096));
  s.w.spans = s.spans_mapped ? s.d_spans_map : s.d_spans_dev;
  s.w.overflow = s.spans_mapped ? s.d_ovf_map : s.d_ovf_dev;
  const bool zero_in_kernel = fused && !s.devparse;  // (k_parse counts into the batch counters before the fused kernel)
  if (s.timed) CUDA_TRY(cudaEventRecord(s.ev[0], st));
  if (!zero_in_kernel) launch_batch_begin(e->dev, s.w, s.in, has_direct, st);
  if (s.devparse) launch_parse(e->dev, s.w, s.in, st);
  if (has_direct && !fused) launch_direct(e->dev, s.w, s.in, st);  // fused: lookup + sort inside k_ctrl_small
  if (s.timed) CUDA_TRY(cudaEventRecord(s.ev[1], st));
  if (fused) {
    launch_ctrl_small(e->dev, s.w, s.in, has_direct, zero_in_kernel, s.d_stats_pub, st);
    if (s.timed) { CUDA_TRY(cudaEventRecord(s.ev[2], st)); CUDA_TRY(cudaEventRecord(s.ev[3], st)); }
  } else {
    launch_match(e->dev, s.w, s.in, st);
    if (s.timed) CUDA_TRY(cudaEventRecord(s.ev[2], st));
    launch_plan(e->dev, s.w, s.in, st);
    launch_offsets(e->dev, s.w, s.in, has_direct, st);
    if (s.timed) CUDA_TRY(cudaEventRecord(s.ev[3], st));
  }
  if (!s.spans_mapped) {
    CUDA_TRY(cudaEventRecord(s.ev_ctrl, st));
    // the span table is final once k_offsets is done: its counters go home while the pack runs
    CUDA_TRY(cudaStreamWaitEvent(cs, s.ev_ctrl, 0));
    CUDA_TRY(cudaMemcpyAsync(s.h_early, s.w.stats, sizeof(BatchStats), cudaMemcpyDeviceToHost, cs));
    CUDA_TRY(cudaEventRecord(s.ev_early, cs));
    // pack on its own stream (packs of successive batches stay ordered among themselves)
    if (ps != st) CUDA_TRY(cudaStreamWaitEvent(ps, s.ev_ctrl, 0));
  }
  // (mapped spans: k_offsets wrote spans / overflow into host memory; everything stays on one
  //  stream and the host waits for ev_done only)
  if (s.timed) CUDA_TRY(cudaEventRecord(s.ev[4], ps));
  launch_pack(e->dev, s.w, s.in, e->cfg.pack_variant, e->n_sms, ps);
  if (s.timed) CUDA_TRY(cudaEventRecord(s.ev[5], ps));
  CUDA_TRY(cudaGetLastError());
  if (!fused) CUDA_TRY(cudaMemcpyAsync(s.h_stats, s.w.stats, sizeof(BatchStats), cudaMemcpyDeviceToHost, ps));
  CUDA_TRY(cudaEventRecord(s.ev_done, ps));
  s.state = SLOT_INFLIGHT;
  s.t_launch = std::chrono::steady_clock::now();
  s.batch_id = e->next_batch_id++;
  s.polled = false;
  e->inflight.push_back(s.batch_id);
  return 0;
}

// close the open batch: upload staging, launch
int flush_open(pcdn_engine* e, uint64_t* batch_id) {
  if (batch_id) *batch_id = 0;
  if (e->open_slot < 0) return 0;
  Slot& s = e->slots[e->open_slot];
  const uint32_t n = (uint32_t)s.kind.size();
  if (n == 0) { s.state = SLOT_FREE; e->open_slot = -1; return 0; }
  if (!e->has_device) return fail(PCDN_ENODEV, "host-only engine cannot route messages");
  int rc = flush_journal(e);
  if (rc) return rc;
  cudaStream_t st = e->stream;
  // descriptor block layout (offsets 16-byte aligned)
  size_t o_kind = 0, o_flags = align_up(o_kind + n, 16), o_slot = align_up(o_flags + n, 16);
  size_t o_len = o_slot + (size_t)n * 4, o_aoff = o_len + (size_t)n * 4, o_alen = o_aoff + (size_t)n * 4;
  size_t o_bidx = o_alen + (size_t)n * 4, o_top = align_up(o_bidx + s.bcast_index.size() * 4, 16);
  size_t total = align_up(o_top + s.topics.size() * 2, 16);
  if (total > e->desc_cap) return fail(PCDN_ENOSPC, "descriptor block overflow");
  // Small batches ride in ONE host→device copy: the descriptor block is appended to the frame arena
  // when it fits there (one DMA + one API call less on the latency path); otherwise two copies.
  const size_t doff = align_up(s.arena_used, 256);
  const bool one_copy = doff + total <= (size_t)e->cfg.max_batch_bytes + 64 && doff + total <= (64u << 10);
  uint8_t* hd = one_copy ? s.h_arena + doff : s.h_desc;
  uint8_t* dd = one_copy ? s.d_arena + doff : s.d_desc;
  std::memcpy(hd + o_kind, s.kind.data(), n);
  std::memcpy(hd + o_flags, s.flags.data(), n);
  std::memcpy(hd + o_slot, s.slot_off16.data(), (size_t)n * 4);
  std::memcpy(hd + o_len, s.raw_len.data(), (size_t)n * 4);
  std::memcpy(hd + o_aoff, s.aux_off.data(), (size_t)n * 4);
  std::memcpy(hd + o_alen, s.aux_len.data(), (size_t)n * 4);
  if (!s.bcast_index.empty()) std::memcpy(hd + o_bidx, s.bcast_index.data(), s.bcast_index.size() * 4);
  if (!s.topics.empty()) std::memcpy(hd + o_top, s.topics.data(), s.topics.size() * 2);
  if (one_copy) {
    CUDA_TRY(cudaMemcpyAsync(s.d_arena, s.h_arena, doff + total, cudaMemcpyHostToDevice, st));
  } else {
    CUDA_TRY(cudaMemcpyAsync(s.d_arena, s.h_arena, align_up(s.arena_used, 16), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s.d_desc, s.h_desc, total, cudaMemcpyHostToDevice, st));
  }
  s.in.n_msgs = n;
  s.in.n_bcast = (uint32_t)s.bcast_index.size();
  s.in.arena = s.d_arena;
  s.in.kind = dd + o_kind;
  s.in.flags = dd + o_flags;
  s.in.slot_off16 = (const uint32_t*)(dd + o_slot);
  s.in.raw_len = (const uint32_t*)(dd + o_len);
  s.in.aux_off = (const uint32_t*)(dd + o_aoff);
  s.in.aux_len = (const uint32_t*)(dd + o_alen);
  s.in.bcast_index = (const uint32_t*)(dd + o_bidx);
  s.in.topics = (const uint16_t*)(dd + o_top);
  s.device_input = false;
  rc = launch_pipeline(e, s, s.n_direct);
  if (rc) return rc;
  if (batch_id) *batch_id = s.batch_id;
  e->open_slot = -1;
  e->stats.batches++;
  e->stats.msgs += n;
  return 0;
}

// R12: a table mutation must not be visible to messages already handed to the engine
int before_state_change(pcdn_engine* e) {
  int rc = 0;
  if (e->open_slot >= 0 && !e->slots[e->open_slot].kind.empty()) rc = flush_open(e, nullptr);
  // connection-id quarantine (host_state.h): ids freed now may be named by spans of batches <= fence_now
  e->conns->fence_now = e->next_batch_id - 1;
  e->conns->oldest_unreleased = e->inflight.empty() ? ~0ull : e->inflight.front();
  return rc;
}

// append one message to the open batch (flushing a full batch first)
int append_msg(pcdn_engine* e, uint8_t kind, uint8_t flags, const uint16_t* topics, uint32_t n_topics,
               const uint8_t* recipient, uint32_t recipient_len, const uint8_t* raw, uint32_t raw_len) {
  if (!e->has_device) return fail(PCDN_ENODEV, "host-only engine cannot route messages");
  if (raw_len > 0x1FFFFFFFu) return fail(PCDN_EINVAL, "message larger than MAX_MESSAGE_SIZE (cdn-proto/src/lib.rs:25)");
  if (kind != PCDN_KIND_BROADCAST && kind != PCDN_KIND_DIRECT) return fail(PCDN_EINVAL, "kind must be broadcast or direct");
  const pcdn_config& c = e->cfg;
  if (c.global_memory_pool_size) {
    // limiter/mod.rs:56-68: the frame's length in permits must be available before it is accepted
    if (raw_len > c.global_memory_pool_size) return fail(PCDN_EINVAL, "message larger than the global memory pool");
    if (e->inflight_bytes + raw_len > c.global_memory_pool_size)
      return fail(PCDN_EAGAIN, "global memory pool exhausted: release a batch first");
  }
  const size_t slot_bytes = align_up(4 + (size_t)raw_len, 16);
  size_t need = slot_bytes + (kind == PCDN_KIND_DIRECT ? align_up(recipient_len, 16) : 0);
  if (need + 64 > c.max_batch_bytes) return fail(PCDN_ENOSPC, "message does not fit max_batch_bytes");
  if (kind == PCDN_KIND_DIRECT && recipient_len > c.max_key_len) {
    // longer than any key in the table: cannot match (bytewise identity, R8) → dropped silently,
    // but batch order bookkeeping still wants the message; route it as "no recipient"
    recipient_len = 0;
  }
  for (int attempt = 0; attempt < 2; attempt++) {
    int rc = acquire_open_slot(e);
    if (rc) return rc;
    Slot& s = e->slots[e->open_slot];
    bool full = s.kind.size() >= c.max_batch_msgs || s.arena_used + need + 64 > c.max_batch_bytes ||
                (kind == PCDN_KIND_BROADCAST && s.bcast_index.size() >= c.max_batch_bcast) ||
                s.topics.size() + n_topics > e->topics_cap;
    if (!full) break;
    if (attempt == 1) return fail(PCDN_ENOSPC, "message does not fit an empty batch");
    if ((rc = flush_open(e, nullptr))) return rc;
  }
  Slot& s = e->slots[e->open_slot];
  const uint32_t m = (uint32_t)s.kind.size();
  const size_t off = s.arena_used;  // 16-byte aligned
  uint8_t* dst = s.h_arena + off;
  std::memset(dst, 0, 4);
  if (raw_len) std::memcpy(dst + 4, raw, raw_len);
  std::memset(dst + 4 + raw_len, 0, slot_bytes - 4 - raw_len);
  s.arena_used += slot_bytes;
  s.kind.push_back(kind);
  s.flags.push_back(flags);
  s.slot_off16.push_back((uint32_t)(off / 16));
  s.raw_len.push_back(raw_len);
  s.ingress_bytes += raw_len;
  e->inflight_bytes += raw_len;
  e->stats.bytes_in += raw_len;
  if (kind == PCDN_KIND_BROADCAST) {
    s.aux_off.push_back((uint32_t)s.topics.size());
    s.aux_len.push_back(n_topics);
    for (uint32_t i = 0; i < n_topics; i++) s.topics.push_back(topics[i]);
    s.bcast_index.push_back(m);
  } else {
    // recipient key: read it in place when it lies inside the frame at a 4-byte aligned offset
    size_t koff;
    if (recipient_len && recipient >= raw && recipient + recipient_len <= raw + raw_len &&
        ((off + 4 + (size_t)(recipient - raw)) & 3) == 0) {
      koff = off + 4 + (size_t)(recipient - raw);
    } else {
      koff = s.arena_used;
      size_t kb = align_up(recipient_len, 16);
      if (recipient_len) std::memcpy(s.h_arena + koff, recipient, recipient_len);
      std::memset(s.h_arena + koff + recipient_len, 0, kb - recipient_len);
      s.arena_used += kb;
    }
    s.aux_off.push_back((uint32_t)koff);
    s.aux_len.push_back(recipient_len);
    s.n_direct++;
  }
  return 0;
}

Slot* find_slot(pcdn_engine* e, uint64_t id) {
  for (auto& s : e->slots)
    if (s.state == SLOT_INFLIGHT && s.batch_id == id) return &s;
  return nullptr;
}

void destroy_engine(pcdn_engine* e) {
  if (e->has_device) {
    cudaSetDevice(e->cfg.device);
    cudaStreamSynchronize(e->stream);
    if (e->pack_stream) cudaStreamSynchronize(e->pack_stream);
    if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
    for (auto& s : e->slots) {
      if (s.ev_done) cudaEventDestroy(s.ev_done);
      if (s.ev_ctrl) cudaEventDestroy(s.ev_ctrl);
      if (s.ev_early) cudaEventDestroy(s.ev_early);
      for (auto& ev : s.ev) if (ev) cudaEventDestroy(ev);
    }
    for (void* p : e->dev_allocs) cudaFree(p);
    for (void* p : e->pin_allocs) cudaFreeHost(p);
    if (e->jstage_h) cudaFreeHost(e->jstage_h);
    if (e->jstage_d) cudaFree(e->jstage_d);
    if (e->ev_journal) cudaEventDestroy(e->ev_journal);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->pack_stream) cudaStreamDestroy(e->pack_stream);
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  }
  delete e;
}

#define DEV_ALLOC(ptr, n)                                 \
  do {                                                    \
    int _rc = dev_alloc(&(ptr), (n));                     \
    if (_rc) return _rc;                                  \
    e->dev_allocs.push_back((void*)(ptr));                \
  } while (0)
#define PIN_ALLOC(ptr, n)                                 \
  do {                                                    \
    int _rc = pin_alloc(&(ptr), (n));                     \
    if (_rc) return _rc;                                  \
    e->pin_allocs.push_back((void*)(ptr));                \
  } while (0)

int init_device(pcdn_engine* e) {
  const pcdn_config& c = e->cfg;
  const Geometry& g = e->geo;
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0)
    return fail(PCDN_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(err));
  if (c.device >= ndev) return fail(PCDN_ENODEV, "device ordinal out of range");
  CUDA_TRY(cudaSetDevice(c.device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, c.device));
  e->n_sms = prop.multiProcessorCount;
  if (c.stream) { e->stream = (cudaStream_t)c.stream; e->own_stream = false; }
  else { CUDA_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)); e->own_stream = true; }
  CUDA_TRY(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&e->ev_journal, cudaEventDisableTiming));
  {
    // highest priority: when a pack and the (small) control kernels of the next batch become
    // runnable together, the pack's persistent CTAs must be placed first and evenly over the SMs
    int lo = 0, hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&e->pack_stream, cudaStreamNonBlocking, hi));
  }
  e->has_device = true;
  e->direct_publish = g.N <= kDirectPublishMaxConns && !(c.flags & PCDN_FLAG_STAGED_SPANS);

  DevState& d = e->dev;
  d.N = g.N; d.W = g.W; d.T = g.T; d.nblk = g.W / kBlockWords;
  d.bucket_mask = g.bucket_mask; d.key_stride = g.key_stride; d.seed = g.seed;
  d.ring_bytes = c.ring_bytes_per_conn; d.ring_units = (uint32_t)(c.ring_bytes_per_conn / kUnit);
  d.cm_enable = (c.pack_variant & 2) ? 0 : 1;
  d.fat_tile_bytes = (128u << 10) << ((c.pack_variant >> 4) & 15u);  // A/B: bits 4-7 double the tile
  d.n_valid_topics = c.n_valid_topics;
  d.max_key_len = c.max_key_len;
  DEV_ALLOC(d.sub, (size_t)g.T * g.W);
  DEV_ALLOC(d.brk, g.W);
  DEV_ALLOC(d.owner_conn, g.max_owners);
  DEV_ALLOC(d.cuckoo, (size_t)g.nbuckets * 4);
  DEV_ALLOC(d.keys, (size_t)g.max_keys * g.key_stride);
  DEV_ALLOC(d.ptail, g.N);
  DEV_ALLOC(d.used, g.N);
  if (c.flags & PCDN_FLAG_HOST_RINGS) {
    // egress hand-off: the pack stores straight into host memory the socket writers read
    int _rc = pin_alloc_mapped(&e->h_rings, &d.rings, (size_t)g.max_conns * c.ring_bytes_per_conn);
    if (_rc) return _rc;
    e->pin_allocs.push_back((void*)e->h_rings);
  } else {
    DEV_ALLOC(d.rings, (size_t)g.max_conns * c.ring_bytes_per_conn);
  }
  CUDA_TRY(cudaMemsetAsync(d.sub, 0, (size_t)g.T * g.W * 4, e->stream));
  CUDA_TRY(cudaMemsetAsync(d.brk, 0, (size_t)g.W * 4, e->stream));
  CUDA_TRY(cudaMemsetAsync(d.owner_conn, 0xFF, (size_t)g.max_owners * 4, e->stream));
  CUDA_TRY(cudaMemsetAsync(d.cuckoo, 0, (size_t)g.nbuckets * 4 * sizeof(CuckooEntry), e->stream));
  CUDA_TRY(cudaMemsetAsync(d.keys, 0, (size_t)g.max_keys * g.key_stride, e->stream));
  CUDA_TRY(cudaMemsetAsync(d.ptail, 0, (size_t)g.N * 4, e->stream));
  CUDA_TRY(cudaMemsetAsync(d.used, 0, (size_t)g.N * 4, e->stream));

  const uint32_t M = c.max_batch_msgs, MB = c.max_batch_bcast;
  e->topics_cap = (size_t)M * 4 + 4096;
  e->desc_cap = align_up((size_t)M * 2 + 64, 16) + (size_t)M * 20 + 64 + e->topics_cap * 2 + 64;
  const size_t cap_fat = (size_t)c.max_batch_deliveries;
  const size_t cap_thin = std::min<size_t>(c.max_batch_deliveries, (size_t)M * (kFatMin - 1));
  const size_t ntiles = sort_tiles(M);
  e->slots.resize(c.batch_slots);
  for (Slot& s : e->slots) {
    PIN_ALLOC(s.h_arena, c.max_batch_bytes + 64);
    PIN_ALLOC(s.h_desc, e->desc_cap);
    DEV_ALLOC(s.d_arena, c.max_batch_bytes + 64);
    DEV_ALLOC(s.d_desc, e->desc_cap);
    Work& w = s.w;
    DEV_ALLOC(w.B, (size_t)MB * g.W);
    DEV_ALLOC(w.wpre, (size_t)MB * g.W);
    DEV_ALLOC(w.cnt, (size_t)MB * d.nblk);
    DEV_ALLOC(w.base, (size_t)MB * d.nblk);
    DEV_ALLOC(w.done, MB);
    CUDA_TRY(cudaMemsetAsync(w.done, 0, (size_t)MB * 4, e->stream));
    DEV_ALLOC(w.D, M);
    DEV_ALLOC(w.dconn, M);
    DEV_ALLOC(w.eb_fat, (size_t)M + 1);
    DEV_ALLOC(w.eb_thin, (size_t)M + 1);
    DEV_ALLOC(w.tbase, (size_t)M + 1);
    DEV_ALLOC(w.scan_tmp, 4 * ((size_t)M / 256 + 2));
    DEV_ALLOC(w.cls, M);
    DEV_ALLOC(w.cm_rank, (size_t)M + 1);
    DEV_ALLOC(w.cm_list, MB);
    DEV_ALLOC(w.jidx, M);
    DEV_ALLOC(w.efat, cap_fat);
    DEV_ALLOC(w.ecm, cap_fat);
    DEV_ALLOC(w.ethin, cap_thin);
    w.cap_fat = (uint32_t)std::min<size_t>(cap_fat, 0xFFFFFFFFu);
    w.cap_thin = (uint32_t)std::min<size_t>(cap_thin, 0xFFFFFFFFu);
    for (int k = 0; k < 2; k++) { DEV_ALLOC(w.skey[k], M); DEV_ALLOC(w.sval[k], M); }
    DEV_ALLOC(w.hist, 256 * ntiles);
    DEV_ALLOC(w.hist_tmp, 256 * ntiles / 1024 + 2);
    DEV_ALLOC(w.dstart, (size_t)g.N + 1);
    DEV_ALLOC(w.dend, (size_t)g.N + 1);
    DEV_ALLOC(w.dstamp, (size_t)g.N + 1);
    CUDA_TRY(cudaMemsetAsync(w.dstamp, 0, ((size_t)g.N + 1) * 4, e->stream));
    w.stamp = 0;
    DEV_ALLOC(w.batch_units, g.N);
    DEV_ALLOC(s.d_spans_dev, (size_t)2 * g.N);
    DEV_ALLOC(s.d_ovf_dev, g.N);
    if (e->direct_publish) {
      int _rc = pin_alloc_mapped(&s.h_spans, &s.d_spans_map, (size_t)2 * g.N);
      if (_rc) return _rc;
      e->pin_allocs.push_back((void*)s.h_spans);
      if ((_rc = pin_alloc_mapped(&s.h_overflow, &s.d_ovf_map, (size_t)g.N))) return _rc;
      e->pin_allocs.push_back((void*)s.h_overflow);
    } else {
      PIN_ALLOC(s.h_spans, (size_t)2 * g.max_conns);
      PIN_ALLOC(s.h_overflow, g.max_conns);
    }
    w.spans = s.d_spans_dev; w.overflow = s.d_ovf_dev;
    DEV_ALLOC(w.msg_status, M);
    PIN_ALLOC(s.h_msg_status, M);
    DEV_ALLOC(w.stats, 1);
    if (e->direct_publish) {
      int _rc = pin_alloc_mapped(&s.h_stats, &s.d_stats_pub, 1);
      if (_rc) return _rc;
      e->pin_allocs.push_back((void*)s.h_stats);
    } else {
      PIN_ALLOC(s.h_stats, 1);
    }
    PIN_ALLOC(s.h_early, 1);
    CUDA_TRY(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s.ev_ctrl, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s.ev_early, cudaEventDisableTiming));
    for (auto& ev : s.ev) CUDA_TRY(cudaEventCreate(&ev));
  }
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return 0;
}

}  // namespace

// ================================================================================== C ABI
#define LOCK std::lock_guard<std::mutex> _g(e->mu)
#define GUARD_BEGIN try {
#define GUARD_END                                                          \
  } catch (const std::bad_alloc&) { return fail(PCDN_ENOMEM, "host allocation failed"); } \
  catch (const std::exception& ex) { return fail(PCDN_EINVAL, ex.what()); }

extern "C" {

uint32_t pcdn_abi_version(void) { return PCDN_ABI_VERSION; }
const char* pcdn_last_error(void) { return g_err.c_str(); }

void pcdn_config_default(pcdn_config* c) {
  std::memset(c, 0, sizeof(*c));
  c->struct_size = sizeof(pcdn_config);
  c->device = 0;
  c->max_conns = 1 << 16;
  c->max_topics = 256;
  c->max_keys = 1 << 17;
  c->max_key_len = 128;
  c->ring_bytes_per_conn = 1 << 16;
  c->max_batch_msgs = 4096;
  c->max_batch_bcast = 1024;
  c->max_batch_bytes = 64ull << 20;
  c->max_batch_deliveries = 16ull << 20;
  c->batch_slots = 4;
  c->n_valid_topics = 0;
  c->hash_seed = 0;
  c->identity = "/";
}

int pcdn_create(const pcdn_config* cfg, pcdn_engine** out) {
  GUARD_BEGIN
  if (!cfg || !out) return fail(PCDN_EINVAL, "null argument");
  if (cfg->struct_size != sizeof(pcdn_config)) return fail(PCDN_EINVAL, "pcdn_config.struct_size mismatch (ABI)");
  if (!cfg->max_conns || !cfg->max_topics || cfg->max_topics > 65536 || !cfg->max_keys || !cfg->max_key_len ||
      !cfg->max_batch_msgs || !cfg->batch_slots)
    return fail(PCDN_EINVAL, "zero or out-of-range capacity in pcdn_config");
  if (cfg->max_batch_bcast == 0 || cfg->max_batch_bcast > 65535) return fail(PCDN_EINVAL, "max_batch_bcast must be 1..65535");
  if (cfg->ring_bytes_per_conn % PCDN_RECORD_ALIGN || cfg->ring_bytes_per_conn == 0 || cfg->ring_bytes_per_conn > (1ull << 31))
    return fail(PCDN_EINVAL, "ring_bytes_per_conn must be a multiple of 32, at most 2 GiB");
  if (cfg->max_key_len > 4096) return fail(PCDN_EINVAL, "max_key_len > 4096");
  pcdn_engine* e = new pcdn_engine();
  e->cfg = *cfg;
  e->identity = cfg->identity ? cfg->identity : "/";
  e->cfg.identity = e->identity.c_str();
  Geometry& g = e->geo;
  g.max_conns = cfg->max_conns;
  g.N = (uint32_t)align_up(cfg->max_conns, 32 * kBlockWords);
  g.W = g.N / 32;
  g.T = cfg->max_topics;
  g.max_keys = cfg->max_keys;
  g.max_key_len = cfg->max_key_len;
  g.key_stride = (uint32_t)align_up(cfg->max_key_len, 16);
  uint32_t nb = 1;
  while ((uint64_t)nb * 2 < cfg->max_keys) nb <<= 1;  // 4 slots per bucket → load factor <= 50 %
  g.nbuckets = nb;
  g.bucket_mask = nb - 1;
  g.max_owners = 4096;
  g.seed = cfg->hash_seed ? cfg->hash_seed : 0x243F6A8885A308D3ULL;
  e->tables.reset(new HostTables(g));
  e->conns.reset(new Connections(*e->tables, e->identity.c_str()));
  if (cfg->device >= 0) {
    int rc = init_device(e);
    if (rc) { destroy_engine(e); return rc; }
  }
  *out = e;
  return 0;
  GUARD_END
}

void pcdn_destroy(pcdn_engine* e) {
  if (e) destroy_engine(e);
}

// ---- state ------------------------------------------------------------------------------------
int pcdn_add_user(pcdn_engine* e, const uint8_t* key, uint32_t key_len, const uint16_t* topics, uint32_t n,
                  pcdn_conn* out_conn) {
  GUARD_BEGIN
  LOCK;
  int rc = before_state_change(e);
  if (rc) return rc;
  rc = e->conns->add_user(std::string((const char*)key, key_len), topics, n, out_conn);
  if (rc) return fail(rc, "add_user failed (capacity, key length or topic id)");
  return 0;
  GUARD_END
}
int pcdn_add_users_bulk(pcdn_engine* e, const uint8_t* keys, uint32_t key_len, uint32_t key_stride, uint32_t n_users,
                        const uint16_t* topics, const uint32_t* topic_offsets, pcdn_conn* out_conns) {
  GUARD_BEGIN
  LOCK;
  int rc = before_state_change(e);
  if (rc) return rc;
  for (uint32_t i = 0; i < n_users; i++) {
    uint32_t conn;
    const uint16_t* t = topics ? topics + topic_offsets[i] : nullptr;
    uint32_t nt = topics ? topic_offsets[i + 1] - topic_offsets[i] : 0;
    rc = e->conns->add_user(std::string((const char*)keys + (size_t)i * key_stride, key_len), t, nt, &conn);
    if (rc) return fail(rc, "add_users_bulk failed at user " + std::to_string(i));
    if (out_conns) out_conns[i] = conn;
  }
  return 0;
  GUARD_END
}
int pcdn_remove_user(pcdn_engine* e, const uint8_t* key, uint32_t key_len) {
  GUARD_BEGIN
  LOCK;
  int rc = before_state_change(e);
  if (rc) return rc;
  rc = e->conns->remove_user(std::string((const char*)key, key_len));
  return rc ? fail(rc, "remove_user failed") : 0;
  GUARD_END
}
int pcdn_subscribe_user_to(pcdn_engine* e, const uint8_t* key, uint32_t key_len, const uint16_t* topics, uint32_t n) {
  GUARD_BEGIN
  LOCK;
  int rc = before_state_change(e);
  if (rc) return rc;
  rc = e->conns->subscribe_user_to(std::string((const char*)key, key_len), topics, n);
  return rc ? fail(rc, "topic id out of range") : 0;
  GUARD_END
}
int pcdn_unsubscribe_user_from(pcdn_engine* e, const uint8_t* key, uint32_t key_len, const uint16_t* topics, uint32_t n) {
  GUARD_BEGIN
  LOCK;
  int rc = before_state_change(e);
  if (rc) return rc;
  return e->conns->unsubscribe_user_from(std::string((const char*)key, key_len), topics, n);
  GUARD_END
}
int pcdn_add_broker(pcdn_engine* e, const char* identifier, pcdn_conn* out_conn) {
  GUARD_BEGIN
  LOCK;
  int rc = before_state_change(e);
  if (rc) return rc;
  rc = e->conns->add_broker(identifier, out_conn);
  return rc ? fail(rc, "add_broker failed") : 0;
  GUARD_END
}
int pcdn_remove_broker(pcdn_engine* e, const char* identifier) {
  GUARD_BEGIN
  LOCK;
  int rc = before_state_change(e);
  if (rc) return rc;
  return e->conns->remove_broker(identifier);
  GUARD_END
}
int pcdn_subscribe_broker_to(pcdn_engine* e, const char* identifier, const uint16_t* topics, uint32_t n) {
  GUARD_BEGIN
  LOCK;
  int rc = before_state_change(e);
  if (rc) return rc;
  rc = e->conns->subscribe_broker_to(identifier, topics, n);
  return rc ? fail(rc, "topic id out of range") : 0;
  GUARD_END
}
int pcdn_unsubscribe_broker_from(pcdn_engine* e, const char* identifier, const uint16_t* topics, uint32_t n) {
  GUARD_BEGIN
  LOCK;
  int rc = before_state_change(e);
  if (rc) return rc;
  return e->conns->unsubscribe_broker_from(identifier, topics, n);
  GUARD_END
}
int pcdn_apply_user_sync(pcdn_engine* e, const char* remote_identity, const pcdn_user_sync_entry* entries, uint32_t n) {
  GUARD_BEGIN
  LOCK;
  int rc = before_state_change(e);
  if (rc) return rc;
  std::vector<UserSyncEntry> v;
  v.reserve(n);
  for (uint32_t i = 0; i < n; i++)
    v.push_back(UserSyncEntry{std::string((const char*)entries[i].key, entries[i].key_len), entries[i].version,
                              entries[i].owner != nullptr, entries[i].owner ? entries[i].owner : ""});
  rc = e->conns->apply_user_sync(remote_identity, v);
  return rc ? fail(rc, "apply_user_sync failed") : 0;
  GUARD_END
}

int pcdn_get_user_sync(pcdn_engine* e, int full, const pcdn_user_sync_entry** out, uint32_t* n) {
  GUARD_BEGIN
  LOCK;
  if (full) e->conns->get_full_user_sync(e->sync_users);
  else e->conns->get_partial_user_sync(e->sync_users);
  e->sync_users_c.clear();
  for (const UserSyncEntry& u : e->sync_users)
    e->sync_users_c.push_back(pcdn_user_sync_entry{(const uint8_t*)u.key.data(), (uint32_t)u.key.size(), u.version,
                                                   u.has_owner ? u.owner.c_str() : nullptr});
  *out = e->sync_users_c.data();
  *n = (uint32_t)e->sync_users_c.size();
  return 0;
  GUARD_END
}
int pcdn_apply_topic_sync(pcdn_engine* e, const char* identifier, uint32_t remote_identity,
                          const pcdn_topic_sync_entry* entries, uint32_t n) {
  GUARD_BEGIN
  LOCK;
  int rc = before_state_change(e);
  if (rc) return rc;
  std::vector<TopicSyncEntry> v;
  v.reserve(n);
  for (uint32_t i = 0; i < n; i++) v.push_back(TopicSyncEntry{entries[i].topic, entries[i].status, entries[i].version});
  rc = e->conns->apply_topic_sync(identifier, remote_identity, v);
  return rc ? fail(rc, "topic id out of range") : 0;
  GUARD_END
}
int pcdn_get_topic_sync(pcdn_engine* e, int full, const pcdn_topic_sync_entry** out, uint32_t* n) {
  GUARD_BEGIN
  LOCK;
  if (full) e->conns->get_full_topic_sync(e->sync_topics);
  else e->conns->get_partial_topic_sync(e->sync_topics);
  e->sync_topics_c.clear();
  for (const TopicSyncEntry& t : e->sync_topics) {
    pcdn_topic_sync_entry c{};
    c.topic = t.topic; c.status = t.status; c.version = t.version;
    e->sync_topics_c.push_back(c);
  }
  *out = e->sync_topics_c.data();
  *n = (uint32_t)e->sync_topics_c.size();
  return 0;
  GUARD_END
}

// ---- data in ----------------------------------------------------------------------------------
int pcdn_handle_broadcast_message(pcdn_engine* e, const uint16_t* topics, uint32_t n_topics, const uint8_t* raw,
                                  uint32_t raw_len, int to_users_only) {
  GUARD_BEGIN
  LOCK;
  return append_msg(e, PCDN_KIND_BROADCAST, to_users_only ? PCDN_TO_USERS_ONLY : 0, topics, n_topics, nullptr, 0, raw, raw_len);
  GUARD_END
}
int pcdn_handle_direct_message(pcdn_engine* e, const uint8_t* recipient, uint32_t recipient_len, const uint8_t* raw,
                               uint32_t raw_len, int to_user_only) {
  GUARD_BEGIN
  LOCK;
  return append_msg(e, PCDN_KIND_DIRECT, to_user_only ? PCDN_TO_USERS_ONLY : 0, nullptr, 0, recipient, recipient_len, raw, raw_len);
  GUARD_END
}

// device-parse mode: Broadcast / Direct frames are only tag-peeked and copied; k_parse does the rest
static int append_frame_devparse(pcdn_engine* e, int kind, bool from_broker, const uint8_t* raw, uint32_t raw_len) {
  uint8_t flags = MSGF_DEVPARSE | (from_broker ? MSGF_USERS_ONLY : 0);
  if (kind == PCDN_KIND_BROADCAST && !from_broker) flags |= MSGF_PRUNE;  // user-origin only (handler.rs:157 vs user/handler.rs:133)
  int rc = append_msg(e, (uint8_t)kind, flags, nullptr, 0, nullptr, 0, raw, raw_len);
  if (rc == 0) e->slots[e->open_slot].devparse = true;
  return rc;
}

static int user_receive_locked(pcdn_engine* e, const uint8_t* sender_key, uint32_t key_len, const uint8_t* raw, uint32_t raw_len) {
  if (e->cfg.flags & PCDN_FLAG_DEVICE_PARSE) {
    const int k = peek_kind_core(raw, raw_len);
    if (k == PCDN_KIND_DIRECT || k == PCDN_KIND_BROADCAST) return append_frame_devparse(e, k, false, raw, raw_len);
  }
  ParsedFrame pf;
  if (!parse_frame(raw, raw_len, &pf)) return fail(PCDN_EPARSE, "failed to deserialize message");
  uint16_t topics[65536 / 8];
  switch (pf.kind) {
    case PCDN_KIND_DIRECT:
      return append_msg(e, PCDN_KIND_DIRECT, 0, nullptr, 0, raw + pf.f0_off, pf.f0_len, raw, raw_len);
    case PCDN_KIND_BROADCAST:
    case PCDN_KIND_SUBSCRIBE:
    case PCDN_KIND_UNSUBSCRIBE: {
      if (pf.f0_len > sizeof(topics) / 2) return fail(PCDN_EPARSE, "topic list too long");
      uint32_t n = prune_topics(raw + pf.f0_off, pf.f0_len, e->cfg.n_valid_topics, topics);
      if (n == 0) return fail(PCDN_EPRUNE, "supplied no valid topics");
      if (pf.kind == PCDN_KIND_BROADCAST) return append_msg(e, PCDN_KIND_BROADCAST, 0, topics, n, nullptr, 0, raw, raw_len);
      int rc = before_state_change(e);
      if (rc) return rc;
      std::string key((const char*)sender_key, key_len);
      rc = pf.kind == PCDN_KIND_SUBSCRIBE ? e->conns->subscribe_user_to(key, topics, n)
                                          : e->conns->unsubscribe_user_from(key, topics, n);
      return rc ? fail(rc, "topic id out of range") : 0;
    }
    default:
      return fail(PCDN_EKIND, "invalid message received");
  }
}

static int broker_receive_locked(pcdn_engine* e, const uint8_t* raw, uint32_t raw_len) {
  if (e->cfg.flags & PCDN_FLAG_DEVICE_PARSE) {
    const int k = peek_kind_core(raw, raw_len);
    if (k == PCDN_KIND_DIRECT || k == PCDN_KIND_BROADCAST) return append_frame_devparse(e, k, true, raw, raw_len);
  }
  ParsedFrame pf;
  if (!parse_frame(raw, raw_len, &pf)) return fail(PCDN_EPARSE, "failed to deserialize message");
  if (pf.kind == PCDN_KIND_DIRECT)
    return append_msg(e, PCDN_KIND_DIRECT, PCDN_TO_USERS_ONLY, nullptr, 0, raw + pf.f0_off, pf.f0_len, raw, raw_len);
  if (pf.kind == PCDN_KIND_BROADCAST) {
    uint16_t topics[65536 / 8];
    if (pf.f0_len > sizeof(topics) / 2) return fail(PCDN_EPARSE, "topic list too long");
    for (uint32_t i = 0; i < pf.f0_len; i++) topics[i] = raw[pf.f0_off + i];  // broker-origin: no prune (handler.rs:157)
    return append_msg(e, PCDN_KIND_BROADCAST, PCDN_TO_USERS_ONLY, topics, pf.f0_len, nullptr, 0, raw, raw_len);
  }
  return 1;
}

int pcdn_user_receive(pcdn_engine* e, const uint8_t* sender_key, uint32_t key_len, const uint8_t* raw, uint32_t raw_len) {
  GUARD_BEGIN
  LOCK;
  return user_receive_locked(e, sender_key, key_len, raw, raw_len);
  GUARD_END
}

int pcdn_broker_receive(pcdn_engine* e, const char* /*identifier*/, const uint8_t* raw, uint32_t raw_len) {
  GUARD_BEGIN
  LOCK;
  return broker_receive_locked(e, raw, raw_len);
  GUARD_END
}

// ---- multi-threaded ingest of many frames -----------------------------------------------------
extern "C++" {
namespace {

struct FramePlan {
  int8_t kind;        // 3 / 4 routable; -1 = needs the sequential path (state change, other kinds); -2 = protocol error
  int32_t rc;
  uint32_t f0_off, f0_len, ntopics;
  uint32_t msg_idx, topic_off, bcast_pos;
  uint64_t arena_off;
};

template <class F>
void parallel_for(uint32_t n, uint32_t nthreads, F f) {
  if (n < 2048 || nthreads <= 1) { f(0u, n); return; }
  std::vector<std::thread> th;
  const uint32_t per = (n + nthreads - 1) / nthreads;
  for (uint32_t t = 1; t < nthreads; t++) {
    const uint32_t lo = t * per, hi = std::min(n, lo + per);
    if (lo < hi) th.emplace_back([=] { f(lo, hi); });
  }
  f(0u, std::min(n, per));
  for (auto& x : th) x.join();
}

uint32_t ingest_threads() {
  static uint32_t n = [] {
    if (const char* e = std::getenv("PCDN_INGEST_THREADS")) return (uint32_t)std::max(1, atoi(e));
    return std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  }();
  return n;
}

// One receive-loop iteration per frame, in order; returns the number of frames consumed (a capacity
// condition — no free batch slot, memory pool exhausted — stops early) or a negative error when
// nothing could be consumed.  Large calls run in three phases per run of routable frames:
//   A (parallel)  parse (host mode) or tag peek (device-parse mode) of every frame
//   scan (serial) a few integer adds per frame: placement in the open batch, capacity, ordering
//   B (parallel)  copy of the raw bytes into the pinned arena + descriptor fill by index
// Frames that change state (Subscribe/Unsubscribe) or need the exact synchronous error path end a
// run and go through user_receive_locked / broker_receive_locked, so R12 ordering is untouched.
int receive_frames_locked(pcdn_engine* e, const pcdn_frame* frames, uint32_t n, int32_t* rc_out) {
  const pcdn_config& c = e->cfg;
  const bool dev = (c.flags & PCDN_FLAG_DEVICE_PARSE) != 0;
  const uint32_t T = ingest_threads();
  if (n < 2048 || T <= 1 || !e->has_device) {
    for (uint32_t i = 0; i < n; i++) {
      const pcdn_frame& f = frames[i];
      int rc = f.origin ? broker_receive_locked(e, f.raw, f.raw_len) : user_receive_locked(e, f.sender, f.sender_len, f.raw, f.raw_len);
      if (rc_out) rc_out[i] = rc;
      if (rc == PCDN_EAGAIN || rc == PCDN_ECUDA || rc == PCDN_ENODEV) return i ? (int)i : rc;
    }
    return (int)n;
  }
  std::vector<FramePlan> plan(n);
  // ---- phase A
  parallel_for(n, T, [&](uint32_t lo, uint32_t hi) {
    for (uint32_t i = lo; i < hi; i++) {
      const pcdn_frame& f = frames[i];
      FramePlan& p = plan[i];
      p.kind = -1; p.rc = 0; p.f0_off = p.f0_len = p.ntopics = 0;
      if (f.raw_len > 0x1FFFFFFFu || align_up(4 + (size_t)f.raw_len, 16) + 64 > c.max_batch_bytes ||
          (c.global_memory_pool_size && f.raw_len > c.global_memory_pool_size)) continue;  // sequential path reports it (PCDN_EINVAL / PCDN_ENOSPC for that frame)
      if (dev) {
        const int k = peek_kind_core(f.raw, f.raw_len);
        if (k == PCDN_KIND_DIRECT || k == PCDN_KIND_BROADCAST) p.kind = (int8_t)k;
        continue;
      }
      ParsedFrame pf;
      if (!parse_frame(f.raw, f.raw_len, &pf)) { p.kind = -2; p.rc = PCDN_EPARSE; continue; }
      if (pf.kind == PCDN_KIND_DIRECT) {
        p.kind = 3; p.f0_off = pf.f0_off; p.f0_len = pf.f0_len > c.max_key_len ? 0 : pf.f0_len;
      } else if (pf.kind == PCDN_KIND_BROADCAST) {
        if (pf.f0_len > 8192) { p.kind = -2; p.rc = PCDN_EPARSE; continue; }
        uint32_t cnt = pf.f0_len;
        if (!f.origin) {  // user-origin: Topic::prune
          cnt = 0;
          for (uint32_t k = 0; k < pf.f0_len; k++) cnt += topic_kept(f.raw + pf.f0_off, k, c.n_valid_topics) ? 1u : 0u;
          if (cnt == 0) { p.kind = -2; p.rc = PCDN_EPRUNE; continue; }
        }
        p.kind = 4; p.f0_off = pf.f0_off; p.f0_len = pf.f0_len; p.ntopics = cnt;
      }
    }
  });
  uint32_t i = 0;
  while (i < n) {
    FramePlan& p0 = plan[i];
    if (p0.kind == -2) { if (rc_out) rc_out[i] = p0.rc; i++; continue; }
    if (p0.kind == -1) {
      const pcdn_frame& f = frames[i];
      int rc = f.origin ? broker_receive_locked(e, f.raw, f.raw_len) : user_receive_locked(e, f.sender, f.sender_len, f.raw, f.raw_len);
      if (rc_out) rc_out[i] = rc;
      if (rc == PCDN_EAGAIN || rc == PCDN_ECUDA || rc == PCDN_ENODEV) return i ? (int)i : rc;
      i++;
      continue;
    }
    // ---- a run of routable frames starting at i: placement scan
    int rc = acquire_open_slot(e);
    if (rc) return i ? (int)i : rc;
    Slot& s = e->slots[e->open_slot];
    uint32_t nm = (uint32_t)s.kind.size(), nb = (uint32_t)s.bcast_index.size(), nt = (uint32_t)s.topics.size(), nd = 0;
    uint64_t used = s.arena_used, ingress = 0;
    const uint32_t m0 = nm, b0 = nb, t0 = nt;
    uint32_t j = i;
    bool full = false;
    for (; j < n; j++) {
      FramePlan& p = plan[j];
      if (p.kind == -2) continue;
      if (p.kind == -1) break;
      const pcdn_frame& f = frames[j];
      const uint64_t sb = align_up(4 + (size_t)f.raw_len, 16);
      if (nm >= c.max_batch_msgs || used + sb + 64 > c.max_batch_bytes || (p.kind == 4 && nb >= c.max_batch_bcast) ||
          (uint64_t)nt + p.ntopics > e->topics_cap) { full = true; break; }
      if (c.global_memory_pool_size && e->inflight_bytes + ingress + f.raw_len > c.global_memory_pool_size) { full = true; break; }
      p.msg_idx = nm++; p.arena_off = used; used += sb; ingress += f.raw_len;
      if (p.kind == 4) { p.bcast_pos = nb++; p.topic_off = nt; nt += dev ? 0 : p.ntopics; }
      else nd++;
    }
    if (j == i) {  // nothing fits: the open batch is full (or the pool is) — launch it and retry, or give up
      if (s.kind.empty()) return i ? (int)i : fail(PCDN_EAGAIN, "global memory pool exhausted: release a batch first");
      if ((rc = flush_open(e, nullptr))) return i ? (int)i : rc;
      continue;
    }
    // ---- phase B: descriptors by index + raw bytes
    s.kind.resize(nm); s.flags.resize(nm); s.slot_off16.resize(nm); s.raw_len.resize(nm); s.aux_off.resize(nm); s.aux_len.resize(nm);
    s.bcast_index.resize(nb); s.topics.resize(nt);
    parallel_for(j - i, T, [&](uint32_t lo, uint32_t hi) {
      for (uint32_t q = i + lo; q < i + hi; q++) {
        const FramePlan& p = plan[q];
        if (p.kind < 0) continue;
        const pcdn_frame& f = frames[q];
        const uint32_t m = p.msg_idx;
        const size_t sb = align_up(4 + (size_t)f.raw_len, 16);
        uint8_t* dst = s.h_arena + p.arena_off;
        std::memset(dst, 0, 4);
        if (f.raw_len) std::memcpy(dst + 4, f.raw, f.raw_len);
        std::memset(dst + 4 + f.raw_len, 0, sb - 4 - f.raw_len);
        s.kind[m] = (uint8_t)p.kind;
        uint8_t fl = f.origin ? MSGF_USERS_ONLY : 0;
        if (dev) fl |= MSGF_DEVPARSE | ((p.kind == 4 && !f.origin) ? MSGF_PRUNE : 0);
        s.flags[m] = fl;
        s.slot_off16[m] = (uint32_t)(p.arena_off / 16);
        s.raw_len[m] = f.raw_len;
        if (p.kind == 4) {
          s.bcast_index[p.bcast_pos] = m;
          s.aux_off[m] = p.topic_off;
          s.aux_len[m] = dev ? 0 : p.ntopics;
          if (!dev) {
            uint32_t k = p.topic_off;
            for (uint32_t t = 0; t < p.f0_len; t++)
              if (f.origin || topic_kept(f.raw + p.f0_off, t, c.n_valid_topics)) s.topics[k++] = f.raw[p.f0_off + t];
          }
        } else {
          s.aux_off[m] = dev ? 0 : (uint32_t)(p.arena_off + 4 + p.f0_off);  // recipient read in place (word aligned)
          s.aux_len[m] = dev ? 0 : p.f0_len;
        }
      }
    });
    if (rc_out)
      for (uint32_t q = i; q < j; q++) rc_out[q] = plan[q].kind == -2 ? plan[q].rc : 0;
    s.arena_used = used;
    s.n_direct += nd;
    s.ingress_bytes += ingress;
    e->inflight_bytes += ingress;
    e->stats.bytes_in += ingress;
    if (dev && nm > m0) s.devparse = true;
    (void)b0; (void)t0;
    i = j;
    if (full) {
      if ((rc = flush_open(e, nullptr))) return (int)i;
    }
  }
  return (int)n;
}

}  // namespace
}  // extern "C++"

int pcdn_receive_frames(pcdn_engine* e, const pcdn_frame* frames, uint32_t n, int32_t* rc_out) {
  GUARD_BEGIN
  LOCK;
  return receive_frames_locked(e, frames, n, rc_out);
  GUARD_END
}

int pcdn_flush(pcdn_engine* e, uint64_t* batch_id) {
  GUARD_BEGIN
  LOCK;
  return flush_open(e, batch_id);
  GUARD_END
}

// All-or-nothing check of an explicit batch against every per-batch capacity, BEFORE anything is
// staged: a refused pcdn_submit leaves no message behind that a later flush would deliver (and a
// retry would deliver twice).
static int validate_explicit_batch(pcdn_engine* e, const pcdn_msg* msgs, uint32_t n) {
  const pcdn_config& c = e->cfg;
  if (!e->has_device) return fail(PCDN_ENODEV, "host-only engine cannot route messages");
  if (n > c.max_batch_msgs) return fail(PCDN_ENOSPC, "batch larger than max_batch_msgs");
  if (n && !msgs) return fail(PCDN_EINVAL, "null message array");
  uint64_t bytes = 0, topics = 0, ingress = 0, nb = 0;
  for (uint32_t i = 0; i < n; i++) {
    const pcdn_msg& m = msgs[i];
    if (m.kind != PCDN_KIND_BROADCAST && m.kind != PCDN_KIND_DIRECT)
      return fail(PCDN_EINVAL, "message " + std::to_string(i) + ": kind must be broadcast or direct");
    if (m.flags & ~(uint8_t)PCDN_TO_USERS_ONLY)
      return fail(PCDN_EINVAL, "message " + std::to_string(i) + ": unknown bits in pcdn_msg.flags");
    if (m.raw_len > 0x1FFFFFFFu) return fail(PCDN_EINVAL, "message larger than MAX_MESSAGE_SIZE (cdn-proto/src/lib.rs:25)");
    if ((m.raw_len && !m.raw) || (m.kind == PCDN_KIND_BROADCAST && m.n_topics && !m.topics) ||
        (m.kind == PCDN_KIND_DIRECT && m.recipient_len && !m.recipient))
      return fail(PCDN_EINVAL, "message " + std::to_string(i) + ": null pointer with non-zero length");
    if (c.global_memory_pool_size && m.raw_len > c.global_memory_pool_size)
      return fail(PCDN_EINVAL, "message larger than the global memory pool");
    bytes += align_up(4 + (size_t)m.raw_len, 16);
    if (m.kind == PCDN_KIND_DIRECT) bytes += align_up(std::min<uint32_t>(m.recipient_len, c.max_key_len), 16);  // worst case: key staged beside the frame
    else { nb++; topics += m.n_topics; }
    ingress += m.raw_len;
  }
  if (bytes + 64 > c.max_batch_bytes) return fail(PCDN_ENOSPC, "batch does not fit max_batch_bytes");
  if (nb > c.max_batch_bcast) return fail(PCDN_ENOSPC, "batch has more broadcasts than max_batch_bcast");
  if (topics > e->topics_cap) return fail(PCDN_ENOSPC, "batch has more topic entries than the descriptor block holds");
  if (c.global_memory_pool_size && e->inflight_bytes + ingress > c.global_memory_pool_size)
    return fail(PCDN_EAGAIN, "global memory pool exhausted: release a batch first");
  return 0;
}

// drop the open batch (nothing of it has been launched) and give its permits back
static void abandon_open(pcdn_engine* e) {
  if (e->open_slot < 0) return;
  Slot& s = e->slots[e->open_slot];
  e->inflight_bytes -= std::min(e->inflight_bytes, s.ingress_bytes);
  e->stats.bytes_in -= std::min(e->stats.bytes_in, s.ingress_bytes);
  slot_reset_open(s);
  s.state = SLOT_FREE;
  e->open_slot = -1;
}

int pcdn_submit(pcdn_engine* e, const pcdn_msg* msgs, uint32_t n, uint64_t* batch_id) {
  GUARD_BEGIN
  LOCK;
  if (batch_id) *batch_id = 0;
  int rc = validate_explicit_batch(e, msgs, n);  // before anything is staged or launched
  if (rc) return rc;
  rc = flush_open(e, nullptr);  // keep explicit batches separate from the implicit open one
  if (rc) return rc;
  if (n == 0) return 0;
  if ((rc = acquire_open_slot(e))) return rc;  // PCDN_EAGAIN: nothing staged
  for (uint32_t i = 0; i < n; i++) {
    const pcdn_msg& m = msgs[i];
    uint64_t before = e->next_batch_id;
    rc = append_msg(e, m.kind, m.flags, m.topics, m.n_topics, m.recipient, m.recipient_len, m.raw, m.raw_len);
    if (rc == 0 && e->next_batch_id != before) rc = fail(PCDN_ENOSPC, "batch exceeded a per-batch capacity and was split");
    if (rc) { abandon_open(e); return rc; }  // unreachable after validation; never leave a half batch open
  }
  return flush_open(e, batch_id);
  GUARD_END
}

int pcdn_submit_device(pcdn_engine* e, const pcdn_device_batch* b, uint64_t* batch_id) {
  GUARD_BEGIN
  LOCK;
  if (batch_id) *batch_id = 0;
  if (!e->has_device) return fail(PCDN_ENODEV, "host-only engine cannot route messages");
  if (!b || b->n_msgs == 0 || b->n_msgs > e->cfg.max_batch_msgs || b->n_bcast > e->cfg.max_batch_bcast ||
      b->n_bcast > b->n_msgs)
    return fail(PCDN_EINVAL, "device batch exceeds configured capacities");
  int rc = flush_open(e, nullptr);
  if (rc) return rc;
  if ((rc = acquire_open_slot(e))) return rc;
  Slot& s = e->slots[e->open_slot];
  if ((rc = flush_journal(e))) { s.state = SLOT_FREE; e->open_slot = -1; return rc; }
  s.in.n_msgs = b->n_msgs;
  s.in.n_bcast = b->n_bcast;
  s.in.arena = (const uint8_t*)b->arena;
  s.in.kind = b->kind;
  s.in.flags = b->flags;
  s.in.slot_off16 = b->slot_off16;
  s.in.raw_len = b->raw_len;
  s.in.aux_off = b->aux_off;
  s.in.aux_len = b->aux_len;
  s.in.topics = b->topics;
  s.in.bcast_index = b->bcast_index;
  s.device_input = true;
  s.devparse = false;
  rc = launch_pipeline(e, s, b->n_msgs - b->n_bcast);
  if (rc) { s.state = SLOT_FREE; e->open_slot = -1; return rc; }
  if (batch_id) *batch_id = s.batch_id;
  e->open_slot = -1;
  e->stats.batches++;
  e->stats.msgs += b->n_msgs;
  return 0;
  GUARD_END
}

// ---- data out ---------------------------------------------------------------------------------
int pcdn_next_batch(pcdn_engine* e, uint64_t* batch_id) {
  LOCK;
  *batch_id = e->inflight.empty() ? 0 : e->inflight.front();
  return 0;
}

int pcdn_poll(pcdn_engine* e, uint64_t batch_id, pcdn_batch_result* out, int block) {
  GUARD_BEGIN
  // The blocking waits happen OUTSIDE the engine lock, so ingest threads keep appending to the next
  // batch while an egress thread waits for this one (one poller per batch).
  cudaEvent_t ev_early = nullptr, ev_done = nullptr;
  bool wait = false, mapped = false;
  {
    std::lock_guard<std::mutex> g(e->mu);
    Slot* s = find_slot(e, batch_id);
    if (!s) return fail(PCDN_ENOENT, "unknown batch id");
    if (!s->polled) {
      if (!block) {
        cudaError_t q = cudaEventQuery(s->ev_done);
        if (q == cudaErrorNotReady) return 1;
        CUDA_TRY(q);
      }
      ev_early = s->ev_early; ev_done = s->ev_done; wait = true; mapped = s->spans_mapped;
    }
  }
  uint32_t nsp = 0, nov = 0;
  bool devparse = false;
  if (wait) {
    // 1. counters as of k_offsets → exact size of the span table; its D2H overlaps the pack
    //    (mapped spans: the table is already in host memory when ev_done fires)
    if (!mapped) {
      CUDA_TRY(cudaEventSynchronize(ev_early));
      std::lock_guard<std::mutex> g(e->mu);
      Slot* s = find_slot(e, batch_id);
      if (!s) return fail(PCDN_ENOENT, "batch released while it was being polled");
      nsp = std::min<uint32_t>(s->h_early->n_spans, 2 * e->geo.max_conns);
      nov = std::min<uint32_t>(s->h_early->n_overflow, e->geo.max_conns);
      devparse = s->devparse;
      if (nsp) CUDA_TRY(cudaMemcpyAsync(s->h_spans, s->w.spans, (size_t)nsp * sizeof(Span), cudaMemcpyDeviceToHost, e->copy_stream));
      if (nov) CUDA_TRY(cudaMemcpyAsync(s->h_overflow, s->w.overflow, (size_t)nov * 4, cudaMemcpyDeviceToHost, e->copy_stream));
    } else {
      std::lock_guard<std::mutex> g(e->mu);
      Slot* s = find_slot(e, batch_id);
      if (!s) return fail(PCDN_ENOENT, "batch released while it was being polled");
      devparse = s->devparse;
    }
    // 2. the pack itself (ring bytes are valid after this)
    CUDA_TRY(cudaEventSynchronize(ev_done));
  }
  std::lock_guard<std::mutex> _g(e->mu);
  Slot* s = find_slot(e, batch_id);
  if (!s) return fail(PCDN_ENOENT, "batch released while it was being polled");
  if (wait && !s->polled) {
    if (devparse) CUDA_TRY(cudaMemcpyAsync(s->h_msg_status, s->w.msg_status, s->in.n_msgs, cudaMemcpyDeviceToHost, e->copy_stream));
    if (nsp || nov || devparse) CUDA_TRY(cudaStreamSynchronize(e->copy_stream));
    if (s->devparse) { s->n_msg_errors = 0; for (uint32_t i = 0; i < s->in.n_msgs; i++) s->n_msg_errors += s->h_msg_status[i] != 0; }
    const BatchStats& bs = *s->h_stats;
    s->polled = true;
    e->stats.deliveries += bs.n_deliveries;
    e->stats.bytes_out += bs.bytes_out;
    if (s->timed) {
      float t[4] = {0, 0, 0, 0};
      for (int i = 0; i < 3; i++) cudaEventElapsedTime(&t[i], s->ev[i], s->ev[i + 1]);
      cudaEventElapsedTime(&t[3], s->ev[4], s->ev[5]);
      e->stats.ms_direct += t[0];
      e->stats.ms_match += t[1];
      e->stats.ms_plan += t[2];
      e->stats.ms_pack += t[3];
      e->stats.ms_total += t[0] + t[1] + t[2] + t[3];
      e->stats.timed_batches++;
    }
  }
  if (out) {
    const BatchStats& bs = *s->h_stats;
    out->batch_id = batch_id;
    out->n_msgs = s->in.n_msgs;
    out->n_spans = std::min<uint32_t>(bs.n_spans, 2 * e->geo.max_conns);
    out->spans = reinterpret_cast<const pcdn_span*>(s->h_spans);
    out->n_deliveries = bs.n_deliveries;
    out->bytes_out = bs.bytes_out;
    out->n_overflow = std::min<uint32_t>(bs.n_overflow, e->geo.max_conns);
    out->overflow_conns = s->h_overflow;
    out->n_direct_dropped = bs.n_direct_dropped;
    out->status = bs.status ? (uint32_t)(-PCDN_E2BIG) : 0;
    out->msg_status = s->devparse ? s->h_msg_status : nullptr;
    out->n_msg_errors = s->n_msg_errors;
    out->reserved = 0;
  }
  return 0;
  GUARD_END
}

int pcdn_read(pcdn_engine* e, pcdn_conn conn, uint32_t ring_off, uint32_t len, void* dst) {
  GUARD_BEGIN
  LOCK;
  if (!e->has_device) return fail(PCDN_ENODEV, "host-only engine");
  if (conn >= e->geo.max_conns || (uint64_t)ring_off + len > e->cfg.ring_bytes_per_conn)
    return fail(PCDN_EINVAL, "read outside the connection's ring");
  if (e->h_rings) {
    std::memcpy(dst, e->h_rings + (size_t)conn * e->cfg.ring_bytes_per_conn + ring_off, len);
    return 0;
  }
  CUDA_TRY(cudaMemcpyAsync(dst, e->dev.rings + (size_t)conn * e->cfg.ring_bytes_per_conn + ring_off, len,
                           cudaMemcpyDeviceToHost, e->copy_stream));
  CUDA_TRY(cudaStreamSynchronize(e->copy_stream));
  return 0;
  GUARD_END
}

int pcdn_release_batch(pcdn_engine* e, uint64_t batch_id) {
  GUARD_BEGIN
  LOCK;
  Slot* s = find_slot(e, batch_id);
  if (!s) return fail(PCDN_ENOENT, "unknown batch id");
  if (e->inflight.empty() || e->inflight.front() != batch_id)
    return fail(PCDN_EINVAL, "batches must be released oldest first");
  // A slot released without having been polled may still have its host→device staging copy queued:
  // its pinned staging buffers must not be refilled before that copy ran (device-input batches have
  // no host staging and stay fully asynchronous — the pipelined submit_device/release loop).
  if (!s->polled && !s->device_input) CUDA_TRY(cudaEventSynchronize(s->ev_done));
  // ring space may be reused only after the pack that filled it has finished
  CUDA_TRY(cudaStreamWaitEvent(e->stream, s->ev_done, 0));
  launch_release(e->dev, s->w.batch_units, s->w.stats, e->stream);
  CUDA_TRY(cudaGetLastError());
  e->inflight.erase(e->inflight.begin());
  s->state = SLOT_FREE;
  // the last 'clone' of every frame of this batch is gone: permits back to the pool (pool.rs:44-52)
  e->inflight_bytes -= std::min(e->inflight_bytes, s->ingress_bytes);
  s->ingress_bytes = 0;
  e->stats.released_batches++;
  {
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - s->t_launch).count();
    e->stats.latency_ms_sum += ms;
    const uint64_t us = (uint64_t)(ms * 1000.0);
    int bucket = 0;
    while (bucket < 15 && us >= (16ull << bucket)) bucket++;
    e->stats.latency_hist_us[bucket]++;
  }
  return 0;
  GUARD_END
}

// ---- introspection ----------------------------------------------------------------------------
int pcdn_get_stats(pcdn_engine* e, pcdn_stats* out) {
  LOCK;
  e->stats.inflight_bytes = e->inflight_bytes;
  *out = e->stats;
  return 0;
}
int pcdn_set_timing(pcdn_engine* e, int on) {
  LOCK;
  e->timing = on != 0;
  return 0;
}
int pcdn_ring_info(pcdn_engine* e, void** dev_base, uint64_t* ring_bytes, uint32_t* max_conns) {
  LOCK;
  if (dev_base) *dev_base = e->has_device ? (void*)e->dev.rings : nullptr;
  if (ring_bytes) *ring_bytes = e->cfg.ring_bytes_per_conn;
  if (max_conns) *max_conns = e->geo.max_conns;
  return 0;
}
int pcdn_host_rings(pcdn_engine* e, const void** host_base) {
  LOCK;
  if (host_base) *host_base = e->h_rings;
  return e->h_rings ? 0 : fail(PCDN_ENOENT, "rings live in device memory (PCDN_FLAG_HOST_RINGS not set)");
}
int pcdn_num_users(pcdn_engine* e, uint32_t* users, uint32_t* brokers) {
  LOCK;
  if (users) *users = e->conns->num_users();
  if (brokers) *brokers = e->conns->num_brokers();
  return 0;
}
int pcdn_debug_interested(pcdn_engine* e, const uint16_t* topics, uint32_t n_topics, int to_users_only, pcdn_conn* out,
                          uint32_t cap, uint32_t* n) {
  GUARD_BEGIN
  LOCK;
  std::vector<uint32_t> v;
  e->conns->interested(topics, n_topics, to_users_only != 0, v);
  *n = (uint32_t)v.size();
  for (uint32_t i = 0; i < v.size() && i < cap; i++) out[i] = v[i];
  return 0;
  GUARD_END
}
int pcdn_debug_route(pcdn_engine* e, const uint8_t* key, uint32_t key_len, int* kind, pcdn_conn* conn) {
  GUARD_BEGIN
  LOCK;
  *kind = e->conns->route(std::string((const char*)key, key_len), conn);
  return 0;
  GUARD_END
}
int pcdn_parse_frame(const uint8_t* raw, uint32_t raw_len, uint16_t* topics_out, uint32_t* n_topics, uint32_t* field_off,
                     uint32_t* field_len) {
  ParsedFrame pf;
  if (!parse_frame(raw, raw_len, &pf)) return fail(PCDN_EPARSE, "failed to deserialize message");
  if (field_off) *field_off = pf.f0_off;
  if (field_len) *field_len = pf.f0_len;
  if (n_topics) *n_topics = 0;
  if ((pf.kind == PCDN_KIND_BROADCAST || pf.kind == PCDN_KIND_SUBSCRIBE || pf.kind == PCDN_KIND_UNSUBSCRIBE) && topics_out &&
      n_topics) {
    uint32_t n = std::min<uint32_t>(pf.f0_len, 256);
    for (uint32_t i = 0; i < n; i++) topics_out[i] = raw[pf.f0_off + i];
    *n_topics = n;
  }
  return pf.kind;
}

}  // extern "C"
