// engine.cu — host runtime of the fan-out engine and the C ABI (include/pcdn_fanout.h).
//
// One engine = one logical broker: ONE host mirror of the routing tables (host_state.*) and ONE
// connection-id space, spread over one or more connection SHARDS.  A shard = one CUDA device with
// its streams, its slice of the subscription bitmap, a replica of the direct map, the output rings
// of its connections and its share of every batch slot.  A batch is staged in pinned memory while it
// is open, brought to every shard on flush (one H2D for a single shard; H2D to shard 0 + ONE
// ncclBroadcast over NVLink for several), routed there by the kernel pipeline of kernels.cuh, and
// its results (span table, counters) come back through pinned memory per shard.
// There is no CPU data path: without a device every routing call fails with PCDN_ENODEV.
#include "engine_internal.h"

namespace pcdn_detail {
thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
}  // namespace pcdn_detail

namespace {

// Upload changed table words/slots/keys and apply them on every local shard's stream (K4).  Stream
// order gives R12: every earlier batch sees the old tables, every later batch the new ones.  A
// shard's bitmap / broker mask are its word slice of the global arrays; owner_conn, the cuckoo
// slots and the key arena are replicated.
int flush_journal(pcdn_engine* e) {
  HostTables& t = *e->tables;
  if (!e->has_device) { t.clear_dirty(); return 0; }
  const Geometry& g = e->geo;
  bool any = !t.dirty_sub.empty() || !t.dirty_brk.empty() || !t.dirty_owner.empty() || !t.dirty_slots.empty() ||
             !t.dirty_keys.empty();
  if (!any) return 0;
  const uint32_t Ws = e->shard_W(), W = g.W;
  const bool full_keys = t.dirty_keys.size() > (size_t)g.max_keys / 16 + 64;
  const bool full_sub = t.dirty_sub.size() > t.sub.size() / 16 + 64;
  const bool full_slots = t.dirty_slots.size() > t.cuckoo.size() / 16 + 64;
  // parts common to all shards
  e->h_slot.clear(); e->h_kslot.clear(); e->h_kbytes.clear();
  if (!full_slots) for (uint32_t i : t.dirty_slots) e->h_slot.push_back(UpdSlot{i, t.cuckoo[i]});
  if (!full_keys) for (uint32_t k : t.dirty_keys) {
    e->h_kslot.push_back(k);
    size_t at = e->h_kbytes.size();
    e->h_kbytes.resize(at + g.key_stride);
    std::memcpy(&e->h_kbytes[at], &t.keys[(size_t)k * g.key_stride], g.key_stride);
  }
  for (Shard& sh : e->shards) {
    DeviceGuard dg(sh.device);
    cudaStream_t st = sh.stream;
    const uint32_t w0 = sh.gindex * Ws;
    // keys first (slots reference them)
    if (full_keys) CUDA_TRY(cudaMemcpyAsync(sh.dev.keys, t.keys.data(), t.keys.size(), cudaMemcpyHostToDevice, st));
    sh.h_u32.clear();
    if (full_sub) {
      if (Ws == W) CUDA_TRY(cudaMemcpyAsync(sh.dev.sub, t.sub.data(), t.sub.size() * 4, cudaMemcpyHostToDevice, st));
      else CUDA_TRY(cudaMemcpy2DAsync(sh.dev.sub, (size_t)Ws * 4, t.sub.data() + w0, (size_t)W * 4, (size_t)Ws * 4, g.T, cudaMemcpyHostToDevice, st));
    } else {
      for (uint32_t i : t.dirty_sub) {
        const uint32_t row = i / W, wd = i % W;
        if (wd >= w0 && wd < w0 + Ws) sh.h_u32.push_back(Upd32{0, row * Ws + (wd - w0), t.sub[i]});
      }
    }
    for (uint32_t i : t.dirty_brk)
      if (i >= w0 && i < w0 + Ws) sh.h_u32.push_back(Upd32{1, i - w0, t.brk[i]});
    for (uint32_t i : t.dirty_owner) sh.h_u32.push_back(Upd32{2, i, t.owner_conn[i]});
    if (full_slots) CUDA_TRY(cudaMemcpyAsync(sh.dev.cuckoo, t.cuckoo.data(), t.cuckoo.size() * sizeof(CuckooEntry), cudaMemcpyHostToDevice, st));
    // One pinned staging block [Upd32 | UpdSlot | key slots | key bytes] → one H2D copy → apply kernels.
    // The staging block is reused by the next flush; an event (not a stream sync) guards it, so table
    // churn at control-plane rate never stalls the batches already queued on the stream.
    const size_t b_u32 = align_up(sh.h_u32.size() * sizeof(Upd32), 16), b_slot = align_up(e->h_slot.size() * sizeof(UpdSlot), 16);
    const size_t b_ks = align_up(e->h_kslot.size() * 4, 16), b_kb = align_up(e->h_kbytes.size(), 16);
    const size_t total = b_u32 + b_slot + b_ks + b_kb;
    if (total) {
      if (sh.ev_journal_pending) { CUDA_TRY(cudaEventSynchronize(sh.ev_journal)); sh.ev_journal_pending = false; }
      if (total > sh.jstage_cap) {
        const size_t ncap = std::max(total, sh.jstage_cap * 2 + (1 << 16));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (sh.jstage_h) cudaFreeHost(sh.jstage_h);
        if (sh.jstage_d) cudaFree(sh.jstage_d);
        sh.jstage_h = nullptr; sh.jstage_d = nullptr; sh.jstage_cap = 0;
        CUDA_TRY(cudaMallocHost((void**)&sh.jstage_h, ncap));
        CUDA_TRY(cudaMalloc((void**)&sh.jstage_d, ncap));
        sh.jstage_cap = ncap;
      }
      uint8_t* h = sh.jstage_h;
      if (b_u32) std::memcpy(h, sh.h_u32.data(), sh.h_u32.size() * sizeof(Upd32));
      if (b_slot) std::memcpy(h + b_u32, e->h_slot.data(), e->h_slot.size() * sizeof(UpdSlot));
      if (b_ks) std::memcpy(h + b_u32 + b_slot, e->h_kslot.data(), e->h_kslot.size() * 4);
      if (b_kb) std::memcpy(h + b_u32 + b_slot + b_ks, e->h_kbytes.data(), e->h_kbytes.size());
      CUDA_TRY(cudaMemcpyAsync(sh.jstage_d, h, total, cudaMemcpyHostToDevice, st));
      uint8_t* d = sh.jstage_d;
      launch_apply_updates(sh.dev, (const Upd32*)d, (uint32_t)sh.h_u32.size(), (const UpdSlot*)(d + b_u32), (uint32_t)e->h_slot.size(),
                           (const uint32_t*)(d + b_u32 + b_slot), d + b_u32 + b_slot + b_ks, (uint32_t)e->h_kslot.size(), st);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaEventRecord(sh.ev_journal, st));
      sh.ev_journal_pending = true;
    }
    // whole-table uploads come from pageable vectors (staged by the runtime before the call returns);
    // they only happen on bulk loads, where one synchronisation is irrelevant
    if (full_keys || full_sub || full_slots) CUDA_TRY(cudaStreamSynchronize(st));
  }
  t.clear_dirty();
  return 0;
}

void slot_reset_open(Slot& s) {
  s.arena_used = 0; s.n_direct = 0; s.devparse = false; s.ingress_bytes = 0; s.n_msgs = 0;
  s.kind.clear(); s.flags.clear(); s.slot_off16.clear(); s.raw_len.clear(); s.aux_off.clear(); s.aux_len.clear();
  s.bcast_index.clear(); s.topics.clear();
  s.device_input = false; s.counted = false;
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

// The adaptive pack-stream overlap applies to batches whose previous output was at most this many bytes.
// With every step queued ahead, about half of the overlapped packs lose the launch race against the next
// control stage and run ~40 % longer (profiles/r2_timeline_overlap.txt); the overlap hides at most the
// ~0.08 ms control stage, so it stops paying once a pack takes more than ~0.25 ms (1.5 GB of stores).
static constexpr unsigned long long kOverlapMaxBytes = 3ull << 29;
static constexpr uint32_t kTimelineBatches = 64;

// run the kernel pipeline of one shard for slot `si`, whose BatchIn is ready (or will be, once
// ev_ingest fires) in that shard's memory
int launch_shard_pipeline(pcdn_engine* e, Shard& sh, uint32_t si, uint32_t n_direct, bool devparse, bool wait_ingest, bool unblock = false) {
  DeviceGuard dg(sh.device);
  ShardSlot& s = sh.slots[si];
  // Default: the pack runs on the main stream.  A/B switch (pack_variant bit 3): run it on the
  // high-priority pack stream so the next batch's control kernels overlap it — measured SLOWER for
  // the bulk-store pack (profiles/r1_sweep_overlap.txt), so it stays opt-in.
  const bool dp = sh.direct_publish;
  // Batches of nothing but many direct messages DO take the pack stream: their control kernels are
  // latency-bound (three dependent random reads per message), their pack is a separate bandwidth-bound
  // launch, and the two overlap well — C4: 1.98 → 2.10 G msgs/s (profiles/r2_cfg_C4_sweep_v*.json).
  const bool direct_only = s.in.n_bcast == 0 && n_direct >= kThinSeparateMin;
  // Broadcast batches: running the next batch's control stage beside this batch's pack helps when the pack is
  // short and message-major (sparse fan-out, config 5: 0.293 -> 0.282 ms) and hurts a long pack, which then shares
  // SMs and HBM with it (config 5 dense: 5.29 -> 6.70 ms; C2 likewise) - profiles/r2_overlap_adaptive.txt.  Which
  // one this batch will be is decided on the device, so the class mix and size of the most recent COMPLETED batch
  // of this shard predict it (a workload changes its mix rarely; a wrong guess costs one batch a few percent).
  if (!dp && s.in.n_bcast > 0) {
    for (int k = 0; k < 2; k++) {
      const int pv = sh.prev_slot[k];
      if (pv < 0 || pv == (int)si) continue;
      const ShardSlot& o = sh.slots[pv];
      if (cudaEventQuery(o.ev_done) != cudaSuccess) { cudaGetLastError(); continue; }
      const BatchStats& ps_ = *o.h_stats;
      sh.fat_only = ps_.status == 0 && ps_.n_cm == 0 && ps_.n_fat_tiles > 0 && ps_.bytes_out <= kOverlapMaxBytes;
      break;
    }
  }
  const bool fat_overlap = s.in.n_bcast > 0 && sh.fat_only && !(e->cfg.pack_variant & 32);   // (bit 5: A/B switch, never overlap broadcast batches)
  sh.prev_slot[1] = sh.prev_slot[0];
  sh.prev_slot[0] = (int)si;
  cudaStream_t st = sh.stream, ps = (!dp && ((e->cfg.pack_variant & 8) || direct_only || fat_overlap)) ? sh.pack_stream : sh.stream, cs = sh.copy_stream;
  const bool has_direct = n_direct > 0;
  if (wait_ingest) CUDA_TRY(cudaStreamWaitEvent(st, s.ev_ingest, 0));
  s.timed = e->timing && !e->timeline_async;
  cudaEvent_t* tev = s.ev;
  bool timed = s.timed;
  if (e->timeline_async) {   // diagnostic: stage events from a ring of kTimelineBatches sets, read back when the engine goes away
    const uint32_t k = sh.tl_n++ % kTimelineBatches;
    tev = &sh.tl_ev[(size_t)k * 6];
    sh.tl_batch[k] = e->next_batch_id;
    timed = true;
  }
  s.polled = false;
  s.n_msg_errors = 0;
  if (++s.w.stamp == 0) s.w.stamp = 1;  // validity stamp of this batch's direct buckets / look-back words
  if (sh.dev.pool && (s.w.stamp & 0x3FFFFFFFu) == 0) {   // the look-back words carry 30 bits of it
    s.w.stamp++;
    CUDA_TRY(cudaMemsetAsync(s.w.lb_state, 0, ((size_t)sh.dev.N / 256 + 1) * 8, st));
  }
  s.w.pool_unblock = unblock ? 1u : 0u;
  // latency path of the smallest geometry: match + plan + offsets in one cluster launch that also
  // zeroes / publishes the counters (kernels.cu: k_ctrl_small)
  // (one cluster of 8 CTAs: worth it while the whole match is a few passes — a 128-message batch on a
  //  65536-slot engine is 1024 (message, block) items and runs 10x faster through the regular kernels)
  const bool fused = dp && sh.dev.N <= kSmallCtrlConns && s.in.n_msgs <= kSmallCtrlMsgs &&
                     (uint64_t)s.in.n_bcast * sh.dev.nblk <= kSmallCtrlItems;
  // Spans go straight into mapped host memory when few are expected (16-byte PCIe writes: a table
  // of 16 K spans measured 20 us slower than the staged copy): the smallest geometry, or a batch
  // without broadcasts and with few messages (at most one span per message).  Otherwise they are staged in HBM and copied
  // out with one DMA of the exact size while the pack runs.
  s.spans_mapped = dp && (sh.dev.N <= 8192 || (s.in.n_bcast == 0 && s.in.n_msgs <= 4096));
  if (sh.dev.pool && !fused) s.spans_mapped = false;   // k_pool_finish patches the table: keep it in HBM until it is final
  s.w.spans = s.spans_mapped ? s.d_spans_map : s.d_spans_dev;
  s.w.overflow = s.spans_mapped ? s.d_ovf_map : s.d_ovf_dev;
  const bool zero_in_kernel = fused && !devparse;  // (k_parse counts into the batch counters before the fused kernel)
  if (timed) CUDA_TRY(cudaEventRecord(tev[0], st));
  if (!zero_in_kernel) launch_batch_begin(sh.dev, s.w, s.in, has_direct, st);
  if (devparse) launch_parse(sh.dev, s.w, s.in, st);
  if (has_direct && !fused) launch_direct(sh.dev, s.w, s.in, n_direct, st);  // fused: lookup + sort inside k_ctrl_small
  if (timed) CUDA_TRY(cudaEventRecord(tev[1], st));
  if (fused) {
    launch_ctrl_small(sh.dev, s.w, s.in, has_direct, zero_in_kernel, s.d_stats_pub, st);
    if (timed) { CUDA_TRY(cudaEventRecord(tev[2], st)); CUDA_TRY(cudaEventRecord(tev[3], st)); }
  } else {
    launch_match(sh.dev, s.w, s.in, st);
    if (timed) CUDA_TRY(cudaEventRecord(tev[2], st));
    launch_plan(sh.dev, s.w, s.in, st);
    launch_offsets(sh.dev, s.w, s.in, has_direct, st);
    if (timed) CUDA_TRY(cudaEventRecord(tev[3], st));
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
  s.on_pack_stream = ps != st;
  if (e->timeline_async) sh.tl_ps[(sh.tl_n - 1) % kTimelineBatches] = (int)s.on_pack_stream;
  if (timed) CUDA_TRY(cudaEventRecord(tev[4], ps));
  launch_pack(sh.dev, s.w, s.in, n_direct, e->cfg.pack_variant, sh.n_sms, ps);
  if (timed) CUDA_TRY(cudaEventRecord(tev[5], ps));
  CUDA_TRY(cudaGetLastError());
  if (!fused) CUDA_TRY(cudaMemcpyAsync(s.h_stats, s.w.stats, sizeof(BatchStats), cudaMemcpyDeviceToHost, ps));
  CUDA_TRY(cudaEventRecord(s.ev_done, ps));
  return 0;
}

// every local shard runs the pipeline; the slot becomes the newest in-flight batch
int launch_pipeline(pcdn_engine* e, uint32_t si, uint32_t n_direct, bool wait_ingest) {
  Slot& s = e->slots[si];
  for (Shard& sh : e->shards) {
    int rc = launch_shard_pipeline(e, sh, si, n_direct, s.devparse, wait_ingest);
    if (rc) return rc;
  }
  s.state = SLOT_INFLIGHT;
  s.t_launch = std::chrono::steady_clock::now();
  s.batch_id = e->next_batch_id++;
  s.counted = false;
  e->inflight.push_back(s.batch_id);
  return 0;
}

// Sharded engines: bring `bytes` of slot si's pinned staging (or, for device input, `src_root` on the
// root GPU) into every shard's d_arena.  NCCL mode: H2D on the root's ingest stream, then ONE
// ncclBroadcast per region over all shards of the broker (grouped over the local shards), all on the
// ingest streams — ahead of the main streams, so it overlaps the pack of the previous batch.  The
// region may only be overwritten once the pack that last read this slot's arena is done (ev_done).
struct IngestRegion { const void* root_src; size_t dst_off; size_t bytes; };  // root_src: device pointer on the root, or nullptr = staged bytes
// Host-staged batch: `bytes` of slot si's pinned staging → every shard's d_arena.
int ingest_staged(pcdn_engine* e, uint32_t si, const uint8_t* h_src, size_t bytes) {
  if (e->ingest == PCDN_INGEST_HOST) {   // every shard copies from the pinned staging itself
    for (Shard& sh : e->shards) {
      DeviceGuard dg(sh.device);
      ShardSlot& ss = sh.slots[si];
      CUDA_TRY(cudaStreamWaitEvent(sh.ingest_stream, ss.ev_done, 0));
      CUDA_TRY(cudaMemcpyAsync(ss.d_arena, h_src, bytes, cudaMemcpyHostToDevice, sh.ingest_stream));
      CUDA_TRY(cudaEventRecord(ss.ev_ingest, sh.ingest_stream));
    }
    return 0;
  }
  const NcclApi* nc = e->nccl;
  for (Shard& sh : e->shards) {
    DeviceGuard dg(sh.device);
    ShardSlot& ss = sh.slots[si];
    CUDA_TRY(cudaStreamWaitEvent(sh.ingest_stream, ss.ev_done, 0));
    if (sh.gindex == 0) CUDA_TRY(cudaMemcpyAsync(ss.d_arena, h_src, bytes, cudaMemcpyHostToDevice, sh.ingest_stream));
  }
  NCCL_TRY(nc, nc->GroupStart());
  for (Shard& sh : e->shards) {
    DeviceGuard dg(sh.device);
    ShardSlot& ss = sh.slots[si];
    int rc = nc->Broadcast(ss.d_arena, ss.d_arena, bytes, kNcclUint8, 0, sh.comm, sh.ingest_stream);
    if (rc) { nc->GroupEnd(); return fail(PCDN_ECUDA, std::string("ncclBroadcast: ") + nc->GetErrorString(rc)); }
  }
  NCCL_TRY(nc, nc->GroupEnd());
  for (Shard& sh : e->shards) {
    DeviceGuard dg(sh.device);
    CUDA_TRY(cudaEventRecord(sh.slots[si].ev_ingest, sh.ingest_stream));
  }
  return 0;
}

// Device-resident batch (its arrays lie on the root GPU, global shard 0).  The root first GATHERS the
// descriptor arrays — and the frames too unless they are large (`arena_in_place`) — into its own slot
// region with a few device-to-device copies, so that ONE ncclBroadcast of one contiguous range (two
// with in-place frames) replicates the batch; nine separate broadcasts per step cost C5-sparse 8 % at
// 8 GPUs.  `wait_submit`: order the ingest after everything queued on the root's main stream (the
// caller's producer kernels); false when the caller says the buffers are already complete, so the
// broadcast of batch n+1 overlaps the pack of batch n.
int ingest_device(pcdn_engine* e, uint32_t si, const IngestRegion* regs, int nregs, bool arena_in_place, bool wait_submit) {
  // regs[0] = frames at offset 0, regs[1..] = descriptor arrays behind them (ascending dst_off)
  const size_t desc_lo = regs[1].dst_off, all_hi = regs[nregs - 1].dst_off + regs[nregs - 1].bytes;
  if (e->ingest == PCDN_INGEST_HOST) {   // single process: peer copies from the root's buffers; the root reads in place
    for (Shard& sh : e->shards) {
      DeviceGuard dg(sh.device);
      ShardSlot& ss = sh.slots[si];
      CUDA_TRY(cudaStreamWaitEvent(sh.ingest_stream, ss.ev_done, 0));
      if (wait_submit) CUDA_TRY(cudaStreamWaitEvent(sh.ingest_stream, e->shards[0].ev_submit, 0));
      for (int r = 0; r < nregs; r++) {
        if (!regs[r].bytes || (sh.gindex == 0 && r == 0 && arena_in_place)) continue;
        if (sh.gindex == 0) CUDA_TRY(cudaMemcpyAsync(ss.d_arena + regs[r].dst_off, regs[r].root_src, regs[r].bytes, cudaMemcpyDeviceToDevice, sh.ingest_stream));
        else CUDA_TRY(cudaMemcpyPeerAsync(ss.d_arena + regs[r].dst_off, sh.device, regs[r].root_src, e->shards[0].device, regs[r].bytes, sh.ingest_stream));
      }
      CUDA_TRY(cudaEventRecord(ss.ev_ingest, sh.ingest_stream));
    }
    return 0;
  }
  const NcclApi* nc = e->nccl;
  for (Shard& sh : e->shards) {
    DeviceGuard dg(sh.device);
    ShardSlot& ss = sh.slots[si];
    CUDA_TRY(cudaStreamWaitEvent(sh.ingest_stream, ss.ev_done, 0));
    if (sh.gindex != 0) continue;
    if (wait_submit) CUDA_TRY(cudaStreamWaitEvent(sh.ingest_stream, sh.ev_submit, 0));
    for (int r = arena_in_place ? 1 : 0; r < nregs; r++)
      if (regs[r].bytes)
        CUDA_TRY(cudaMemcpyAsync(ss.d_arena + regs[r].dst_off, regs[r].root_src, regs[r].bytes, cudaMemcpyDeviceToDevice, sh.ingest_stream));
  }
  NCCL_TRY(nc, nc->GroupStart());
  for (Shard& sh : e->shards) {
    DeviceGuard dg(sh.device);
    ShardSlot& ss = sh.slots[si];
    int rc = 0;
    if (arena_in_place) {
      if (regs[0].bytes) {
        void* buf = sh.gindex == 0 ? const_cast<void*>(regs[0].root_src) : (void*)ss.d_arena;   // the root sends from the caller's buffer
        rc = nc->Broadcast(buf, buf, regs[0].bytes, kNcclUint8, 0, sh.comm, sh.ingest_stream);
      }
      if (!rc) rc = nc->Broadcast(ss.d_arena + desc_lo, ss.d_arena + desc_lo, all_hi - desc_lo, kNcclUint8, 0, sh.comm, sh.ingest_stream);
    } else {
      rc = nc->Broadcast(ss.d_arena, ss.d_arena, all_hi, kNcclUint8, 0, sh.comm, sh.ingest_stream);
    }
    if (rc) { nc->GroupEnd(); return fail(PCDN_ECUDA, std::string("ncclBroadcast: ") + nc->GetErrorString(rc)); }
  }
  NCCL_TRY(nc, nc->GroupEnd());
  for (Shard& sh : e->shards) {
    DeviceGuard dg(sh.device);
    CUDA_TRY(cudaEventRecord(sh.slots[si].ev_ingest, sh.ingest_stream));
  }
  return 0;
}

// close the open batch: upload staging, launch
int flush_open(pcdn_engine* e, uint64_t* batch_id) {
  if (batch_id) *batch_id = 0;
  if (e->open_slot < 0) return 0;
  const uint32_t si = (uint32_t)e->open_slot;
  Slot& s = e->slots[si];
  const uint32_t n = (uint32_t)s.kind.size();
  if (n == 0) { s.state = SLOT_FREE; e->open_slot = -1; return 0; }
  if (!e->has_device) return fail(PCDN_ENODEV, "host-only engine cannot route messages");
  int rc = flush_journal(e);
  if (rc) return rc;
  // descriptor block layout (offsets 16-byte aligned)
  size_t o_kind = 0, o_flags = align_up(o_kind + n, 16), o_slot = align_up(o_flags + n, 16);
  size_t o_len = o_slot + (size_t)n * 4, o_aoff = o_len + (size_t)n * 4, o_alen = o_aoff + (size_t)n * 4;
  size_t o_bidx = o_alen + (size_t)n * 4, o_top = align_up(o_bidx + s.bcast_index.size() * 4, 16);
  size_t total = align_up(o_top + s.topics.size() * 2, 16);
  if (total > e->desc_cap) return fail(PCDN_ENOSPC, "descriptor block overflow");
  // Small batches ride in ONE host→device copy: the descriptor block is appended to the frame arena
  // when it fits there (one DMA + one API call less on the latency path); otherwise two copies.
  // Sharded engines always use the appended layout: the batch is one ingest region.
  const size_t doff = align_up(s.arena_used, 256);
  const bool one_copy = e->sharded || (doff + total <= (size_t)e->cfg.max_batch_bytes + 64 && doff + total <= (64u << 10));
  uint8_t* hd = one_copy ? s.h_arena + doff : s.h_desc;
  std::memcpy(hd + o_kind, s.kind.data(), n);
  std::memcpy(hd + o_flags, s.flags.data(), n);
  std::memcpy(hd + o_slot, s.slot_off16.data(), (size_t)n * 4);
  std::memcpy(hd + o_len, s.raw_len.data(), (size_t)n * 4);
  std::memcpy(hd + o_aoff, s.aux_off.data(), (size_t)n * 4);
  std::memcpy(hd + o_alen, s.aux_len.data(), (size_t)n * 4);
  if (!s.bcast_index.empty()) std::memcpy(hd + o_bidx, s.bcast_index.data(), s.bcast_index.size() * 4);
  if (!s.topics.empty()) std::memcpy(hd + o_top, s.topics.data(), s.topics.size() * 2);
  if (e->sharded) {
    std::memset(s.h_arena + s.arena_used, 0, doff - s.arena_used);
    if ((rc = ingest_staged(e, si, s.h_arena, doff + total))) return rc;
  } else {
    Shard& sh = e->shards[0];
    DeviceGuard dg(sh.device);
    ShardSlot& ss = sh.slots[si];
    if (one_copy) {
      CUDA_TRY(cudaMemcpyAsync(ss.d_arena, s.h_arena, doff + total, cudaMemcpyHostToDevice, sh.stream));
    } else {
      CUDA_TRY(cudaMemcpyAsync(ss.d_arena, s.h_arena, align_up(s.arena_used, 16), cudaMemcpyHostToDevice, sh.stream));
      CUDA_TRY(cudaMemcpyAsync(ss.d_desc, s.h_desc, total, cudaMemcpyHostToDevice, sh.stream));
    }
  }
  for (Shard& sh : e->shards) {
    ShardSlot& ss = sh.slots[si];
    uint8_t* dd = one_copy ? ss.d_arena + doff : ss.d_desc;
    ss.in.n_msgs = n;
    ss.in.n_bcast = (uint32_t)s.bcast_index.size();
    ss.in.arena = ss.d_arena;
    ss.in.kind = dd + o_kind;
    ss.in.flags = dd + o_flags;
    ss.in.slot_off16 = (const uint32_t*)(dd + o_slot);
    ss.in.raw_len = (const uint32_t*)(dd + o_len);
    ss.in.aux_off = (const uint32_t*)(dd + o_aoff);
    ss.in.aux_len = (const uint32_t*)(dd + o_alen);
    ss.in.bcast_index = (const uint32_t*)(dd + o_bidx);
    ss.in.topics = (const uint16_t*)(dd + o_top);
  }
  s.device_input = false;
  s.n_msgs = n;
  rc = launch_pipeline(e, si, s.n_direct, e->sharded);
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
    // (multi-process groups always stage it beside the frame: the layout must not depend on how a
    // process happens to hold the bytes)
    size_t koff;
    if (recipient_len && recipient >= raw && recipient + recipient_len <= raw + raw_len &&
        ((off + 4 + (size_t)(recipient - raw)) & 3) == 0 && e->world_shards == e->shards.size()) {
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

}  // namespace
int pcdn_detail::find_slot_index(pcdn_engine* e, uint64_t id) {
  for (size_t i = 0; i < e->slots.size(); i++)
    if (e->slots[i].state == SLOT_INFLIGHT && e->slots[i].batch_id == id) return (int)i;
  return -1;
}
namespace {

void destroy_shard(pcdn_engine* e, Shard& sh) {
  if (!sh.stream && sh.dev_allocs.empty() && sh.pin_allocs.empty()) return;  // never initialised (create failed earlier)
  if (cudaSetDevice(sh.device) != cudaSuccess) { cudaGetLastError(); return; }
  if (sh.stream) cudaStreamSynchronize(sh.stream);
  if (sh.pack_stream) cudaStreamSynchronize(sh.pack_stream);
  if (sh.copy_stream) cudaStreamSynchronize(sh.copy_stream);
  if (sh.ingest_stream) cudaStreamSynchronize(sh.ingest_stream);
  if (sh.comm && e->nccl) { e->nccl->CommDestroy(sh.comm); sh.comm = nullptr; }
  if (e->timeline && e->timeline_async && sh.ev_base) {
    const uint32_t n = std::min(sh.tl_n, kTimelineBatches), first = sh.tl_n - n;
    for (uint32_t j = first; j < sh.tl_n; j++) {
      const uint32_t k = j % kTimelineBatches;
      float t[6];
      for (int q = 0; q < 6; q++) if (cudaEventElapsedTime(&t[q], sh.ev_base, sh.tl_ev[(size_t)k * 6 + q]) != cudaSuccess) { t[q] = -1.f; cudaGetLastError(); }
      std::fprintf(e->timeline, "shard %u batch %llu pack_stream %d ctrl_begin %.4f direct_end %.4f match_end %.4f offsets_end %.4f pack_begin %.4f pack_end %.4f\n",
                   sh.gindex, (unsigned long long)sh.tl_batch[k], sh.tl_ps[k], t[0], t[1], t[2], t[3], t[4], t[5]);
    }
    std::fflush(e->timeline);
  }
  for (auto& ev : sh.tl_ev) if (ev) cudaEventDestroy(ev);
  for (auto& s : sh.slots) {
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    if (s.ev_ctrl) cudaEventDestroy(s.ev_ctrl);
    if (s.ev_early) cudaEventDestroy(s.ev_early);
    if (s.ev_ingest) cudaEventDestroy(s.ev_ingest);
    for (auto& ev : s.ev) if (ev) cudaEventDestroy(ev);
  }
  for (void* p : sh.dev_allocs) cudaFree(p);
  for (void* p : sh.pin_allocs) cudaFreeHost(p);
  if (sh.jstage_h) cudaFreeHost(sh.jstage_h);
  if (sh.jstage_d) cudaFree(sh.jstage_d);
  if (sh.ev_journal) cudaEventDestroy(sh.ev_journal);
  if (sh.ev_submit) cudaEventDestroy(sh.ev_submit);
  if (sh.ev_base) cudaEventDestroy(sh.ev_base);
  if (sh.copy_stream) cudaStreamDestroy(sh.copy_stream);
  if (sh.pack_stream) cudaStreamDestroy(sh.pack_stream);
  if (sh.ingest_stream) cudaStreamDestroy(sh.ingest_stream);
  if (sh.own_stream && sh.stream) cudaStreamDestroy(sh.stream);
}

void destroy_engine(pcdn_engine* e) {
  if (e->has_device) {
    int prev = -1;
    cudaGetDevice(&prev);
    for (Shard& sh : e->shards) destroy_shard(e, sh);
    for (Slot& s : e->slots) {
      if (s.h_arena) cudaFreeHost(s.h_arena);
      if (s.h_desc) cudaFreeHost(s.h_desc);
    }
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();  // a failed create must not leave its error behind for the next engine's launches
  }
  if (e->timeline) std::fclose(e->timeline);
  delete e;
}

#define DEV_ALLOC(ptr, n)                                 \
  do {                                                    \
    int _rc = dev_alloc(&(ptr), (n));                     \
    if (_rc) return _rc;                                  \
    sh.dev_allocs.push_back((void*)(ptr));                \
  } while (0)
#define PIN_ALLOC(ptr, n)                                 \
  do {                                                    \
    int _rc = pin_alloc(&(ptr), (n));                     \
    if (_rc) return _rc;                                  \
    sh.pin_allocs.push_back((void*)(ptr));                \
  } while (0)
#define PIN_ALLOC_MAPPED(ptr, alias, n)                   \
  do {                                                    \
    int _rc = pin_alloc_mapped(&(ptr), &(alias), (n));    \
    if (_rc) return _rc;                                  \
    sh.pin_allocs.push_back((void*)(ptr));                \
  } while (0)

// device side of one shard: streams, its table slices / replicas, rings, per-slot scratch + results
int init_shard(pcdn_engine* e, Shard& sh, int ndev, void* user_stream) {
  const pcdn_config& c = e->cfg;
  const Geometry& g = e->geo;
  if (sh.device < 0 || sh.device >= ndev) return fail(PCDN_ENODEV, "device ordinal " + std::to_string(sh.device) + " out of range (" + std::to_string(ndev) + " CUDA devices)");
  CUDA_TRY(cudaSetDevice(sh.device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, sh.device));
  sh.n_sms = prop.multiProcessorCount;
  if (user_stream) { sh.stream = (cudaStream_t)user_stream; sh.own_stream = false; }
  else { CUDA_TRY(cudaStreamCreateWithFlags(&sh.stream, cudaStreamNonBlocking)); sh.own_stream = true; }
  CUDA_TRY(cudaStreamCreateWithFlags(&sh.copy_stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&sh.ev_journal, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&sh.ev_submit, cudaEventDisableTiming));
  if (e->timeline) { CUDA_TRY(cudaEventCreate(&sh.ev_base)); CUDA_TRY(cudaEventRecord(sh.ev_base, sh.stream)); }
  if (e->timeline_async) {
    sh.tl_ev.assign((size_t)kTimelineBatches * 6, nullptr);
    sh.tl_batch.assign(kTimelineBatches, 0); sh.tl_ps.assign(kTimelineBatches, 0);
    for (auto& ev : sh.tl_ev) CUDA_TRY(cudaEventCreate(&ev));
  }
  {
    // highest priority: when a pack and the (small) control kernels of the next batch become
    // runnable together, the pack's persistent CTAs must be placed first and evenly over the SMs
    int lo = 0, hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&sh.pack_stream, cudaStreamNonBlocking, hi));
    if (e->sharded) CUDA_TRY(cudaStreamCreateWithPriority(&sh.ingest_stream, cudaStreamNonBlocking, hi));
  }
  const uint32_t Ns = g.shard_N, Ws = Ns / 32;
  sh.direct_publish = Ns <= kDirectPublishMaxConns && !(c.flags & PCDN_FLAG_STAGED_SPANS);

  DevState& d = sh.dev;
  d.N = Ns; d.W = Ws; d.T = g.T; d.nblk = Ws / kBlockWords;
  d.bucket_mask = g.bucket_mask; d.key_stride = g.key_stride; d.seed = g.seed;
  d.ring_bytes = c.ring_bytes_per_conn; d.ring_units = (uint32_t)(c.ring_bytes_per_conn / kUnit);
  d.cm_enable = (c.pack_variant & 2) ? 0 : 1;
  d.fat_tile_bytes = (128u << 10) << ((c.pack_variant >> 4) & 15u);  // A/B: bits 4-7 double the tile
  d.fat_grab = 1u << ((c.pack_variant >> 16) & 7u);                   // A/B: bits 16-18 = log2 tiles per cursor update
  d.n_valid_topics = c.n_valid_topics;
  d.max_key_len = c.max_key_len;
  d.conn_base = sh.gindex * Ns;
  d.span_runs = (c.flags & PCDN_FLAG_SPAN_RUNS) ? 1u : 0u;
  d.pool = (c.flags & PCDN_FLAG_OUTPUT_POOL) ? 1u : 0u;
  d.pool_units = (uint32_t)(e->pool_bytes / kUnit);
  d.count_drops = sh.gindex == 0 ? 1u : 0u;
  DEV_ALLOC(d.sub, (size_t)g.T * Ws);
  DEV_ALLOC(d.brk, Ws);
  DEV_ALLOC(d.owner_conn, g.max_owners);
  DEV_ALLOC(d.cuckoo, (size_t)g.nbuckets * 4);
  DEV_ALLOC(d.keys, (size_t)g.max_keys * g.key_stride);
  DEV_ALLOC(d.ptail, Ns);
  DEV_ALLOC(d.used, Ns);
  const size_t out_bytes = d.pool ? (size_t)e->pool_bytes : (size_t)g.shard_max_conns * c.ring_bytes_per_conn;
  if (c.flags & PCDN_FLAG_HOST_RINGS) {
    // egress hand-off: the pack stores straight into host memory the socket writers read
    PIN_ALLOC_MAPPED(sh.h_rings, d.rings, out_bytes);
  } else {
    DEV_ALLOC(d.rings, out_bytes);
  }
  if (d.pool) {
    DEV_ALLOC(d.pool_state, 1);
    launch_pool_init(d, sh.stream);
  }
  CUDA_TRY(cudaMemsetAsync(d.sub, 0, (size_t)g.T * Ws * 4, sh.stream));
  CUDA_TRY(cudaMemsetAsync(d.brk, 0, (size_t)Ws * 4, sh.stream));
  CUDA_TRY(cudaMemsetAsync(d.owner_conn, 0xFF, (size_t)g.max_owners * 4, sh.stream));
  CUDA_TRY(cudaMemsetAsync(d.cuckoo, 0, (size_t)g.nbuckets * 4 * sizeof(CuckooEntry), sh.stream));
  CUDA_TRY(cudaMemsetAsync(d.keys, 0, (size_t)g.max_keys * g.key_stride, sh.stream));
  CUDA_TRY(cudaMemsetAsync(d.ptail, 0, (size_t)Ns * 4, sh.stream));
  CUDA_TRY(cudaMemsetAsync(d.used, 0, (size_t)Ns * 4, sh.stream));

  const uint32_t M = c.max_batch_msgs, MB = c.max_batch_bcast;
  const size_t cap_fat = (size_t)c.max_batch_deliveries;
  const size_t cap_thin = std::min<size_t>(c.max_batch_deliveries, (size_t)M * (kFatMin - 1));
  sh.slots.resize(c.batch_slots);
  for (ShardSlot& s : sh.slots) {
    DEV_ALLOC(s.d_arena, e->arena_cap);
    if (!e->sharded) DEV_ALLOC(s.d_desc, e->desc_cap);
    Work& w = s.w;
    DEV_ALLOC(w.B, (size_t)MB * Ws);
    DEV_ALLOC(w.wpre, (size_t)MB * Ws);
    DEV_ALLOC(w.cnt, (size_t)MB * d.nblk);
    DEV_ALLOC(w.base, (size_t)MB * d.nblk);
    DEV_ALLOC(w.done, MB);
    CUDA_TRY(cudaMemsetAsync(w.done, 0, (size_t)MB * 4, sh.stream));
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
    DEV_ALLOC(w.dcount, (size_t)Ns + 2);
    DEV_ALLOC(w.dloc, (size_t)Ns + 2);
    DEV_ALLOC(w.dtile, (size_t)Ns / 1024 + 3);
    DEV_ALLOC(w.dlist, M);
    DEV_ALLOC(w.hot_list, (size_t)M / kHotMin + 2);
    DEV_ALLOC(w.hot_bitmap, (size_t)kHotCtas * ((size_t)M / 32 + 1));
    DEV_ALLOC(w.scan_done, 1);
    CUDA_TRY(cudaMemsetAsync(w.scan_done, 0, 4, sh.stream));
    DEV_ALLOC(w.edir, M);
    DEV_ALLOC(w.dstart, (size_t)Ns + 1);
    DEV_ALLOC(w.dend, (size_t)Ns + 1);
    DEV_ALLOC(w.dstamp, (size_t)Ns + 1);
    CUDA_TRY(cudaMemsetAsync(w.dstamp, 0, ((size_t)Ns + 1) * 4, sh.stream));
    w.stamp = 0;
    DEV_ALLOC(w.batch_units, Ns);
    if (d.pool) {
      DEV_ALLOC(w.cbase, Ns);
      DEV_ALLOC(w.lb_state, (size_t)Ns / 256 + 1);
      DEV_ALLOC(w.lb_tot, (size_t)Ns / 256 + 1);
      CUDA_TRY(cudaMemsetAsync(w.lb_state, 0, ((size_t)Ns / 256 + 1) * 8, sh.stream));
    }
    w.pool_unblock = 0;
    const size_t span_entries = (size_t)2 * Ns * (d.span_runs ? 3 : 2) / 2;  // SpanRun = 24 B, Span = 16 B
    DEV_ALLOC(s.d_spans_dev, span_entries);
    DEV_ALLOC(s.d_ovf_dev, Ns);
    if (sh.direct_publish) {
      PIN_ALLOC_MAPPED(s.h_spans, s.d_spans_map, span_entries);
      PIN_ALLOC_MAPPED(s.h_overflow, s.d_ovf_map, (size_t)Ns);
    } else {
      PIN_ALLOC(s.h_spans, span_entries);
      PIN_ALLOC(s.h_overflow, g.shard_max_conns);
    }
    w.spans = s.d_spans_dev; w.overflow = s.d_ovf_dev;
    DEV_ALLOC(w.msg_status, M);
    PIN_ALLOC(s.h_msg_status, M);
    DEV_ALLOC(w.stats, 1);
    if (sh.direct_publish) PIN_ALLOC_MAPPED(s.h_stats, s.d_stats_pub, 1);
    else PIN_ALLOC(s.h_stats, 1);
    PIN_ALLOC(s.h_early, 1);
    CUDA_TRY(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s.ev_ctrl, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s.ev_early, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s.ev_ingest, cudaEventDisableTiming));
    for (auto& ev : s.ev) CUDA_TRY(cudaEventCreate(&ev));
  }
  CUDA_TRY(cudaStreamSynchronize(sh.stream));
  return 0;
}

// the ingest communicator: one NCCL rank per shard of the broker, ranks of this process = its shards
int init_nccl(pcdn_engine* e) {
  const char* why = "";
  e->nccl = nccl_api(&why);
  if (!e->nccl) return fail(PCDN_ENODEV, std::string("PCDN_INGEST_NCCL needs libnccl.so.2: ") + why);
  NcclUniqueId id;
  if (e->cfg.nccl_unique_id) std::memcpy(&id, e->cfg.nccl_unique_id, sizeof(id));
  else NCCL_TRY(e->nccl, e->nccl->GetUniqueId(&id));
  NCCL_TRY(e->nccl, e->nccl->GroupStart());
  for (Shard& sh : e->shards) {
    cudaSetDevice(sh.device);
    int rc = e->nccl->CommInitRank(&sh.comm, (int)e->world_shards, id, (int)sh.gindex);
    if (rc) { e->nccl->GroupEnd(); return fail(PCDN_ECUDA, std::string("ncclCommInitRank: ") + e->nccl->GetErrorString(rc)); }
  }
  NCCL_TRY(e->nccl, e->nccl->GroupEnd());
  for (Shard& sh : e->shards) {
    int n = 0;
    if (e->nccl->CommCount && e->nccl->CommCount(sh.comm, &n) == 0) sh.nccl_ranks = n;
  }
  return 0;
}

int init_device(pcdn_engine* e) {
  const pcdn_config& c = e->cfg;
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0)
    return fail(PCDN_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(err));
  int prev = -1;
  cudaGetDevice(&prev);
  const uint32_t M = c.max_batch_msgs;
  e->topics_cap = (size_t)M * 4 + 4096;
  e->desc_cap = align_up((size_t)M * 2 + 64, 16) + (size_t)M * 20 + 64 + e->topics_cap * 2 + 64;
  // sharded engines keep frames + descriptor block in one ingest region (and device-input batches
  // need room for the descriptor arrays behind the frames)
  e->arena_cap = c.max_batch_bytes + 64 + (e->sharded ? 256 + e->desc_cap : 0);
  e->has_device = true;
  for (size_t i = 0; i < e->shards.size(); i++) {
    int rc = init_shard(e, e->shards[i], ndev, i == 0 ? c.stream : nullptr);
    if (rc) return rc;
  }
  e->slots.resize(c.batch_slots);
  for (Slot& s : e->slots) {
    int rc = pin_alloc(&s.h_arena, e->arena_cap);
    if (rc) return rc;
    if (!e->sharded && (rc = pin_alloc(&s.h_desc, e->desc_cap))) return rc;
  }
  if (e->sharded && e->ingest == PCDN_INGEST_NCCL) {
    int rc = init_nccl(e);
    if (rc) return rc;
  }
  if (prev >= 0) cudaSetDevice(prev);
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
  // ---- connection shards
  const uint32_t n_local = cfg->n_devices ? cfg->n_devices : 1;
  const uint32_t world = cfg->world_shards ? cfg->world_shards : n_local;
  if (cfg->n_devices && !cfg->devices) return fail(PCDN_EINVAL, "n_devices > 0 but devices == NULL");
  if (n_local > 64 || world > 1024 || cfg->first_shard + n_local > world)
    return fail(PCDN_EINVAL, "shard layout: first_shard + n_devices must be <= world_shards");
  if (cfg->ingest != PCDN_INGEST_NCCL && cfg->ingest != PCDN_INGEST_HOST) return fail(PCDN_EINVAL, "unknown pcdn_config.ingest");
  const bool host_only = cfg->n_devices ? false : cfg->device < 0;
  if (world > n_local && cfg->ingest == PCDN_INGEST_NCCL && !cfg->nccl_unique_id && !host_only)
    return fail(PCDN_EINVAL, "a multi-process group needs pcdn_config.nccl_unique_id (pcdn_nccl_unique_id)");
  const uint64_t shard_N = align_up(cfg->max_conns, 32 * kBlockWords);
  if (shard_N * world > 0xFFFF0000ull) return fail(PCDN_EINVAL, "connection id space (max_conns x shards) exceeds 32 bits");
  if (cfg->n_devices && cfg->ingest == PCDN_INGEST_NCCL && world > 1)
    for (uint32_t i = 0; i < n_local; i++)
      for (uint32_t j = 0; j < i; j++)
        if (cfg->devices[i] == cfg->devices[j])
          return fail(PCDN_EINVAL, "PCDN_INGEST_NCCL needs one GPU per shard (use PCDN_INGEST_HOST for shards that share a device)");
  pcdn_engine* e = new pcdn_engine();
  if (const char* tl = std::getenv("PCDN_TIMELINE")) {
    e->timeline = std::fopen(tl, "a");
    e->timing = e->timeline != nullptr;
    const char* as = std::getenv("PCDN_TIMELINE_ASYNC");
    e->timeline_async = e->timeline && as && as[0] == '1';
  }
  e->cfg = *cfg;
  e->identity = cfg->identity ? cfg->identity : "/";
  e->cfg.identity = e->identity.c_str();
  e->world_shards = world;
  e->first_shard = cfg->first_shard;
  e->sharded = world > 1;
  e->ingest = cfg->ingest;
  if (cfg->n_devices) e->devices.assign(cfg->devices, cfg->devices + cfg->n_devices);
  else e->devices.assign(1, cfg->device);
  e->cfg.devices = e->devices.data();
  e->cfg.nccl_unique_id = nullptr;  // consumed below; never dereferenced after pcdn_create returns
  Geometry& g = e->geo;
  g.n_shards = world;
  g.shard_N = (uint32_t)shard_N;
  g.shard_max_conns = cfg->max_conns;
  g.max_conns = (world - 1) * g.shard_N + cfg->max_conns;
  g.N = world * g.shard_N;
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
  if (cfg->flags & PCDN_FLAG_OUTPUT_POOL) {
    e->pool_bytes = cfg->pool_bytes ? cfg->pool_bytes : (uint64_t)cfg->max_conns * cfg->ring_bytes_per_conn;
    e->pool_bytes = e->pool_bytes / PCDN_RECORD_ALIGN * PCDN_RECORD_ALIGN;
    if (e->pool_bytes < 4096 || e->pool_bytes / PCDN_RECORD_ALIGN >= 0xFFFFFFFFull) {
      delete e;
      return fail(PCDN_EINVAL, "pool_bytes must be between 4 KiB and 128 GiB");
    }
  }
  e->tables.reset(new HostTables(g));
  e->conns.reset(new Connections(*e->tables, e->identity.c_str()));
  if (!host_only) {
    e->shards.resize(n_local);
    for (uint32_t i = 0; i < n_local; i++) { e->shards[i].device = e->devices[i]; e->shards[i].gindex = cfg->first_shard + i; }
    e->cfg.nccl_unique_id = cfg->nccl_unique_id;
    int rc = init_device(e);
    e->cfg.nccl_unique_id = nullptr;
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

// MessageHookDef::on_message_received on the parsed message (def.rs:79-92).  Returns 0 = process,
// 1 = skip, negative = error (the receive loop ends).  The hook may shrink / rewrite the topic list
// (a private copy) and re-point the recipient.
static int run_hook(pcdn_engine* e, int origin, const ParsedFrame& pf, const uint8_t* sender, uint32_t sender_len,
                    const uint8_t* raw, uint32_t raw_len, std::vector<uint8_t>& topic_copy, const uint8_t** f0, uint32_t* f0_len) {
  *f0 = raw + pf.f0_off; *f0_len = pf.f0_len;
  if (!e->hook[origin]) return 0;
  pcdn_hook_message m{};
  m.kind = (uint8_t)pf.kind; m.origin = (uint8_t)origin;
  m.raw = raw; m.raw_len = raw_len; m.sender = sender; m.sender_len = sender_len;
  const bool has_topics = pf.kind == PCDN_KIND_BROADCAST || pf.kind == PCDN_KIND_SUBSCRIBE || pf.kind == PCDN_KIND_UNSUBSCRIBE;
  if (has_topics) {
    if (pf.f0_len > 65535) return fail(PCDN_EPARSE, "topic list too long");
    topic_copy.assign(raw + pf.f0_off, raw + pf.f0_off + pf.f0_len);
    m.topics = topic_copy.data(); m.n_topics = (uint16_t)pf.f0_len;
  } else if (pf.kind == PCDN_KIND_DIRECT) {
    m.recipient = raw + pf.f0_off; m.recipient_len = pf.f0_len;
  }
  const int r = e->hook[origin](e->hook_user[origin], &m);
  if (r < 0) return fail(PCDN_EHOOK, "hook failed: " + std::to_string(r));
  if (r == PCDN_HOOK_SKIP) return 1;
  if (has_topics) {
    if (m.n_topics > topic_copy.size() || m.topics != topic_copy.data()) return fail(PCDN_EHOOK, "hook returned an invalid topic list");
    *f0 = topic_copy.data(); *f0_len = m.n_topics;
  } else if (pf.kind == PCDN_KIND_DIRECT) {
    if (m.recipient_len && !m.recipient) return fail(PCDN_EHOOK, "hook returned a null recipient");
    *f0 = m.recipient; *f0_len = m.recipient_len;
  }
  return 0;
}

static int user_receive_locked(pcdn_engine* e, const uint8_t* sender_key, uint32_t key_len, const uint8_t* raw, uint32_t raw_len) {
  if ((e->cfg.flags & PCDN_FLAG_DEVICE_PARSE) && !e->hook[0]) {
    const int k = peek_kind_core(raw, raw_len);
    if (k == PCDN_KIND_DIRECT || k == PCDN_KIND_BROADCAST) return append_frame_devparse(e, k, false, raw, raw_len);
  }
  ParsedFrame pf;
  if (!parse_frame(raw, raw_len, &pf)) return fail(PCDN_EPARSE, "failed to deserialize message");
  std::vector<uint8_t> hooked_topics;
  const uint8_t* f0; uint32_t f0_len;
  int hr = run_hook(e, 0, pf, sender_key, key_len, raw, raw_len, hooked_topics, &f0, &f0_len);
  if (hr < 0) return hr;
  if (hr == 1) return 0;  // Ok(HookResult::SkipMessage) => continue
  uint16_t topics[65536 / 8];
  switch (pf.kind) {
    case PCDN_KIND_DIRECT:
      return append_msg(e, PCDN_KIND_DIRECT, 0, nullptr, 0, f0, f0_len, raw, raw_len);
    case PCDN_KIND_BROADCAST:
    case PCDN_KIND_SUBSCRIBE:
    case PCDN_KIND_UNSUBSCRIBE: {
      if (f0_len > sizeof(topics) / 2) return fail(PCDN_EPARSE, "topic list too long");
      uint32_t n = prune_topics(f0, f0_len, e->cfg.n_valid_topics, topics);
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

static int broker_receive_locked(pcdn_engine* e, const uint8_t* identifier, uint32_t identifier_len, const uint8_t* raw, uint32_t raw_len) {
  if ((e->cfg.flags & PCDN_FLAG_DEVICE_PARSE) && !e->hook[1]) {
    const int k = peek_kind_core(raw, raw_len);
    if (k == PCDN_KIND_DIRECT || k == PCDN_KIND_BROADCAST) return append_frame_devparse(e, k, true, raw, raw_len);
  }
  ParsedFrame pf;
  if (!parse_frame(raw, raw_len, &pf)) return fail(PCDN_EPARSE, "failed to deserialize message");
  std::vector<uint8_t> hooked_topics;
  const uint8_t* f0; uint32_t f0_len;
  int hr = run_hook(e, 1, pf, identifier, identifier_len, raw, raw_len, hooked_topics, &f0, &f0_len);
  if (hr < 0) return hr;
  if (hr == 1) return 0;  // Ok(HookResult::SkipMessage) => continue
  if (pf.kind == PCDN_KIND_DIRECT)
    return append_msg(e, PCDN_KIND_DIRECT, PCDN_TO_USERS_ONLY, nullptr, 0, f0, f0_len, raw, raw_len);
  if (pf.kind == PCDN_KIND_BROADCAST) {
    uint16_t topics[65536 / 8];
    if (f0_len > sizeof(topics) / 2) return fail(PCDN_EPARSE, "topic list too long");
    for (uint32_t i = 0; i < f0_len; i++) topics[i] = f0[i];  // broker-origin: no prune (handler.rs:157)
    return append_msg(e, PCDN_KIND_BROADCAST, PCDN_TO_USERS_ONLY, topics, f0_len, nullptr, 0, raw, raw_len);
  }
  return 1;
}

int pcdn_user_receive(pcdn_engine* e, const uint8_t* sender_key, uint32_t key_len, const uint8_t* raw, uint32_t raw_len) {
  GUARD_BEGIN
  LOCK;
  return user_receive_locked(e, sender_key, key_len, raw, raw_len);
  GUARD_END
}

int pcdn_broker_receive(pcdn_engine* e, const char* identifier, const uint8_t* raw, uint32_t raw_len) {
  GUARD_BEGIN
  LOCK;
  return broker_receive_locked(e, (const uint8_t*)identifier, identifier ? (uint32_t)std::strlen(identifier) : 0, raw, raw_len);
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
  if (n < 2048 || T <= 1 || !e->has_device || e->hook[0] || e->hook[1]) {  // a hook sees every parsed message, in order
    for (uint32_t i = 0; i < n; i++) {
      const pcdn_frame& f = frames[i];
      int rc = f.origin ? broker_receive_locked(e, f.sender, f.sender_len, f.raw, f.raw_len) : user_receive_locked(e, f.sender, f.sender_len, f.raw, f.raw_len);
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
      int rc = f.origin ? broker_receive_locked(e, f.sender, f.sender_len, f.raw, f.raw_len) : user_receive_locked(e, f.sender, f.sender_len, f.raw, f.raw_len);
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

int pcdn_set_message_hook(pcdn_engine* e, int origin, pcdn_message_hook cb, void* user) {
  GUARD_BEGIN
  LOCK;
  if (origin != 0 && origin != 1) return fail(PCDN_EINVAL, "origin must be 0 (user) or 1 (broker)");
  e->hook[origin] = cb;
  e->hook_user[origin] = cb ? user : nullptr;
  return 0;
  GUARD_END
}

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
  const uint32_t n = b->n_msgs, nb = b->n_bcast;
  // region layout of the batch inside a receiving shard's arena (sharded engines)
  const size_t o_arena = 0, o_kind = align_up(b->arena_bytes, 256), o_flags = align_up(o_kind + n, 16), o_slot = align_up(o_flags + n, 16);
  const size_t o_len = o_slot + (size_t)n * 4, o_aoff = o_len + (size_t)n * 4, o_alen = o_aoff + (size_t)n * 4;
  const size_t o_bidx = o_alen + (size_t)n * 4, o_top = align_up(o_bidx + (size_t)nb * 4, 16);
  const size_t total = align_up(o_top + (size_t)b->n_topics_total * 2, 16);
  if (e->sharded) {
    if (total > e->arena_cap) return fail(PCDN_ENOSPC, "device batch does not fit the shards' ingest region (max_batch_bytes)");
    if (e->ingest == PCDN_INGEST_HOST && e->world_shards != e->shards.size())
      return fail(PCDN_EINVAL, "device-resident batches in a multi-process group need PCDN_INGEST_NCCL");
  }
  int rc = flush_open(e, nullptr);
  if (rc) return rc;
  if ((rc = acquire_open_slot(e))) return rc;
  const uint32_t si = (uint32_t)e->open_slot;
  Slot& s = e->slots[si];
  auto give_up = [&](int code) { s.state = SLOT_FREE; e->open_slot = -1; return code; };
  if ((rc = flush_journal(e))) return give_up(rc);
  const bool ready = (b->hints & PCDN_BATCH_READY) != 0;
  const bool arena_in_place = b->arena_bytes > (4u << 20);   // large frame arenas are broadcast from the caller's buffer
  if (e->sharded) {
    // the batch lives on the root GPU (global shard 0): unless the caller says its buffers are complete,
    // everything queued so far on the root's main stream (the caller's producer kernels when it shares
    // that stream) comes before the broadcast
    if (e->owns_root() && !ready) {
      Shard& root = e->shards[0];
      DeviceGuard dg(root.device);
      CUDA_TRY(cudaEventRecord(root.ev_submit, root.stream));
    }
    const IngestRegion regs[9] = {
        {b->arena, o_arena, (size_t)b->arena_bytes}, {b->kind, o_kind, n}, {b->flags, o_flags, n},
        {b->slot_off16, o_slot, (size_t)n * 4}, {b->raw_len, o_len, (size_t)n * 4}, {b->aux_off, o_aoff, (size_t)n * 4},
        {b->aux_len, o_alen, (size_t)n * 4}, {b->bcast_index, o_bidx, (size_t)nb * 4}, {b->topics, o_top, (size_t)b->n_topics_total * 2}};
    if ((rc = ingest_device(e, si, regs, 9, arena_in_place, !ready))) return give_up(rc);
  }
  for (Shard& sh : e->shards) {
    ShardSlot& ss = sh.slots[si];
    ss.in.n_msgs = n;
    ss.in.n_bcast = nb;
    if (!e->sharded) {  // where the caller put it
      ss.in.arena = (const uint8_t*)b->arena;
      ss.in.kind = b->kind; ss.in.flags = b->flags; ss.in.slot_off16 = b->slot_off16; ss.in.raw_len = b->raw_len;
      ss.in.aux_off = b->aux_off; ss.in.aux_len = b->aux_len; ss.in.topics = b->topics; ss.in.bcast_index = b->bcast_index;
    } else {            // the replicated copy in this shard's slot region (the root keeps large frames in place)
      uint8_t* d = ss.d_arena;
      ss.in.arena = (sh.gindex == 0 && arena_in_place) ? (const uint8_t*)b->arena : d + o_arena;
      ss.in.kind = d + o_kind; ss.in.flags = d + o_flags;
      ss.in.slot_off16 = (const uint32_t*)(d + o_slot); ss.in.raw_len = (const uint32_t*)(d + o_len);
      ss.in.aux_off = (const uint32_t*)(d + o_aoff); ss.in.aux_len = (const uint32_t*)(d + o_alen);
      ss.in.bcast_index = (const uint32_t*)(d + o_bidx); ss.in.topics = (const uint16_t*)(d + o_top);
    }
  }
  s.device_input = true;
  s.devparse = false;
  s.n_msgs = n;
  rc = launch_pipeline(e, si, n - nb, e->sharded);
  if (rc) return give_up(rc);
  if (batch_id) *batch_id = s.batch_id;
  e->open_slot = -1;
  e->stats.batches++;
  e->stats.msgs += n;
  return 0;
  GUARD_END
}

// ---- data out ---------------------------------------------------------------------------------
int pcdn_next_batch(pcdn_engine* e, uint64_t* batch_id) {
  LOCK;
  *batch_id = e->inflight.empty() ? 0 : e->inflight.front();
  return 0;
}

extern "C++" {
namespace {

void fill_result(pcdn_engine* e, const Slot& s, const Shard& sh, const ShardSlot& ss, uint64_t batch_id, pcdn_batch_result* out) {
  const BatchStats& bs = *ss.h_stats;
  const uint32_t cap = e->geo.shard_max_conns;
  out->batch_id = batch_id;
  out->n_msgs = s.n_msgs;
  // a batch the device refused (scatter-list capacity, output pool) wrote nothing: its provisional
  // counters (the offsets pass counts before the pool says no) are not deliveries
  const bool refused = bs.status != 0;
  out->n_spans = refused ? 0 : std::min<uint32_t>(bs.n_spans, 2 * cap);
  out->spans = sh.dev.span_runs ? nullptr : reinterpret_cast<const pcdn_span*>(ss.h_spans);
  out->runs = sh.dev.span_runs ? reinterpret_cast<const pcdn_span_run*>(ss.h_spans) : nullptr;
  out->n_runs = (sh.dev.span_runs && !refused) ? std::min<uint32_t>(bs.n_runs, 2 * e->geo.shard_N) : 0;

  out->n_deliveries = refused ? 0 : bs.n_deliveries;
  out->bytes_out = refused ? 0 : bs.bytes_out;
  out->n_overflow = refused ? 0 : std::min<uint32_t>(bs.n_overflow, cap);
  out->overflow_conns = ss.h_overflow;
  out->n_direct_dropped = bs.n_direct_dropped;
  // device status: 1 = scatter list capacity, 3 = larger than the whole output pool (E2BIG); 2 = no room in the pool right now (EAGAIN)
  out->status = bs.status == 0 ? 0 : (bs.status == 2 ? (uint32_t)(-PCDN_EAGAIN) : (uint32_t)(-PCDN_E2BIG));
  out->pool_base = sh.dev.pool ? bs.pool_base : 0;
  out->msg_status = s.devparse ? ss.h_msg_status : nullptr;
  out->n_msg_errors = ss.n_msg_errors;
  out->reserved = sh.gindex;
}

// wait for (or test) one shard's share of a batch and fetch its results; returns 1 when !block and
// the shard is not done yet.  The blocking waits happen OUTSIDE the engine lock, so ingest threads
// keep appending to the next batch while an egress thread waits for this one (one poller per
// shard and batch).
}  // namespace
int pcdn_detail::poll_one(pcdn_engine* e, uint64_t batch_id, uint32_t li, int block) {
  cudaEvent_t ev_early = nullptr, ev_done = nullptr;
  bool mapped = false;
  int device = 0;
  {
    std::lock_guard<std::mutex> g(e->mu);
    const int si = find_slot_index(e, batch_id);
    if (si < 0) return fail(PCDN_ENOENT, "unknown batch id");
    if (li >= e->shards.size()) return fail(PCDN_EINVAL, "no such local shard");
    ShardSlot& ss = e->shards[li].slots[si];
    if (ss.polled) return 0;
    device = e->shards[li].device;
    if (!block) {
      DeviceGuard dg(device);
      cudaError_t q = cudaEventQuery(ss.ev_done);
      if (q == cudaErrorNotReady) return 1;
      CUDA_TRY(q);
    }
    ev_early = ss.ev_early; ev_done = ss.ev_done; mapped = ss.spans_mapped;
  }
  DeviceGuard dg(device);
  uint32_t nsp = 0, nov = 0;
  bool devparse = false;
  // 1. counters as of k_offsets → exact size of the span table; its D2H overlaps the pack
  //    (mapped spans: the table is already in host memory when ev_done fires)
  if (!mapped) CUDA_TRY(cudaEventSynchronize(ev_early));
  {
    std::lock_guard<std::mutex> g(e->mu);
    const int si = find_slot_index(e, batch_id);
    if (si < 0) return fail(PCDN_ENOENT, "batch released while it was being polled");
    Shard& sh = e->shards[li];
    ShardSlot& ss = sh.slots[si];
    devparse = e->slots[si].devparse;
    if (!mapped && !ss.polled) {
      const bool runs = sh.dev.span_runs != 0;
      nsp = std::min<uint32_t>(runs ? ss.h_early->n_runs : ss.h_early->n_spans, 2 * e->geo.shard_N);
      nov = std::min<uint32_t>(ss.h_early->n_overflow, e->geo.shard_max_conns);
      if (nsp) CUDA_TRY(cudaMemcpyAsync(ss.h_spans, ss.w.spans, (size_t)nsp * (runs ? sizeof(SpanRun) : sizeof(Span)), cudaMemcpyDeviceToHost, sh.copy_stream));
      if (nov) CUDA_TRY(cudaMemcpyAsync(ss.h_overflow, ss.w.overflow, (size_t)nov * 4, cudaMemcpyDeviceToHost, sh.copy_stream));
    }
  }
  // 2. the pack itself (ring bytes are valid after this)
  CUDA_TRY(cudaEventSynchronize(ev_done));
  std::lock_guard<std::mutex> _g(e->mu);
  const int si = find_slot_index(e, batch_id);
  if (si < 0) return fail(PCDN_ENOENT, "batch released while it was being polled");
  Shard& sh = e->shards[li];
  ShardSlot& ss = sh.slots[si];
  if (!ss.polled) {
    if (devparse) CUDA_TRY(cudaMemcpyAsync(ss.h_msg_status, ss.w.msg_status, ss.in.n_msgs, cudaMemcpyDeviceToHost, sh.copy_stream));
    if (nsp || nov || devparse) CUDA_TRY(cudaStreamSynchronize(sh.copy_stream));
    if (devparse) { ss.n_msg_errors = 0; for (uint32_t i = 0; i < ss.in.n_msgs; i++) ss.n_msg_errors += ss.h_msg_status[i] != 0; }
    const BatchStats& bs = *ss.h_stats;
    ss.polled = true;
    if (!bs.status) {
      e->stats.deliveries += bs.n_deliveries;
      e->stats.bytes_out += bs.bytes_out;
    }
    if (ss.timed && li == 0) {  // stage times of the first local shard (shards run the same pipeline side by side)
      float t[4] = {0, 0, 0, 0};
      for (int i = 0; i < 3; i++) cudaEventElapsedTime(&t[i], ss.ev[i], ss.ev[i + 1]);
      cudaEventElapsedTime(&t[3], ss.ev[4], ss.ev[5]);
      e->stats.ms_direct += t[0];
      e->stats.ms_match += t[1];
      e->stats.ms_plan += t[2];
      e->stats.ms_pack += t[3];
      e->stats.ms_total += t[0] + t[1] + t[2] + t[3];
      e->stats.timed_batches++;
    }
  }
  return 0;
}

}  // extern "C++"

int pcdn_poll_shard(pcdn_engine* e, uint64_t batch_id, uint32_t local_shard, pcdn_batch_result* out, int block) {
  GUARD_BEGIN
  if (!e->has_device) return fail(PCDN_ENODEV, "host-only engine");
  int rc = poll_one(e, batch_id, local_shard, block);
  if (rc) return rc;
  std::lock_guard<std::mutex> _g(e->mu);
  const int si = find_slot_index(e, batch_id);
  if (si < 0) return fail(PCDN_ENOENT, "batch released while it was being polled");
  if (out) fill_result(e, e->slots[si], e->shards[local_shard], e->shards[local_shard].slots[si], batch_id, out);
  return 0;
  GUARD_END
}

int pcdn_poll(pcdn_engine* e, uint64_t batch_id, pcdn_batch_result* out, int block) {
  GUARD_BEGIN
  if (!e->has_device) return fail(PCDN_ENODEV, "host-only engine");
  const uint32_t nl = (uint32_t)e->shards.size();
  if (nl == 1) return pcdn_poll_shard(e, batch_id, 0, out, block);
  for (uint32_t li = 0; li < nl; li++) {
    int rc = poll_one(e, batch_id, li, block);
    if (rc) return rc;  // error, or 1 = some shard still running
  }
  std::lock_guard<std::mutex> _g(e->mu);
  const int si = find_slot_index(e, batch_id);
  if (si < 0) return fail(PCDN_ENOENT, "batch released while it was being polled");
  Slot& s = e->slots[si];
  if (out) {
    // summed counters + the shards' span tables concatenated (ascending shard = ascending id range)
    s.merged_spans.clear(); s.merged_overflow.clear(); s.merged_runs.clear();
    pcdn_batch_result tot{};
    for (uint32_t li = 0; li < nl; li++) {
      pcdn_batch_result r{};
      fill_result(e, s, e->shards[li], e->shards[li].slots[si], batch_id, &r);
      // (output pool: every shard's offsets are relative to ITS region — the merged view carries absolute unit offsets)
      const size_t s0 = s.merged_spans.size(), r0 = s.merged_runs.size();
      if (r.spans) s.merged_spans.insert(s.merged_spans.end(), r.spans, r.spans + r.n_spans);
      if (r.runs) s.merged_runs.insert(s.merged_runs.end(), r.runs, r.runs + r.n_runs);
      if (r.pool_base) {
        for (size_t i = s0; i < s.merged_spans.size(); i++) s.merged_spans[i].ring_off += r.pool_base;
        for (size_t i = r0; i < s.merged_runs.size(); i++) s.merged_runs[i].ring_off += r.pool_base;
      }
      tot.n_spans += r.n_spans;
      s.merged_overflow.insert(s.merged_overflow.end(), r.overflow_conns, r.overflow_conns + r.n_overflow);
      tot.n_deliveries += r.n_deliveries; tot.bytes_out += r.bytes_out; tot.n_direct_dropped += r.n_direct_dropped;
      if (r.status == (uint32_t)(-PCDN_EAGAIN) || (r.status && !tot.status)) tot.status = r.status;   // "retry" wins over "too big"
      if (li == 0) { tot.msg_status = r.msg_status; tot.n_msg_errors = r.n_msg_errors; }  // identical on every shard
    }
    tot.batch_id = batch_id; tot.n_msgs = s.n_msgs;
    tot.spans = s.merged_runs.empty() && !(e->cfg.flags & PCDN_FLAG_SPAN_RUNS) ? s.merged_spans.data() : nullptr;
    tot.runs = (e->cfg.flags & PCDN_FLAG_SPAN_RUNS) ? s.merged_runs.data() : nullptr;
    tot.n_runs = (uint32_t)s.merged_runs.size();
    tot.n_overflow = (uint32_t)s.merged_overflow.size(); tot.overflow_conns = s.merged_overflow.data();
    *out = tot;
  }
  return 0;
  GUARD_END
}

int pcdn_read(pcdn_engine* e, pcdn_conn conn, uint32_t ring_off, uint32_t len, void* dst) {
  GUARD_BEGIN
  LOCK;
  if (!e->has_device) return fail(PCDN_ENODEV, "host-only engine");
  const uint32_t gs = conn / e->geo.shard_N, local = conn % e->geo.shard_N;
  if (gs < e->first_shard || gs >= e->first_shard + e->shards.size())
    return fail(PCDN_ENOENT, "connection lives on a shard of another process");
  Shard& sh = e->shards[gs - e->first_shard];
  size_t at;
  if (sh.dev.pool) {  // ring_off = absolute 32-byte unit inside the shard's output pool (pool_base + span offset)
    at = (size_t)ring_off * PCDN_RECORD_ALIGN;
    if (local >= e->geo.shard_max_conns || at + len > e->pool_bytes) return fail(PCDN_EINVAL, "read outside the output pool");
  } else {
    if (local >= e->geo.shard_max_conns || (uint64_t)ring_off + len > e->cfg.ring_bytes_per_conn)
      return fail(PCDN_EINVAL, "read outside the connection's ring");
    at = (size_t)local * e->cfg.ring_bytes_per_conn + ring_off;
  }
  if (sh.h_rings) {
    std::memcpy(dst, sh.h_rings + at, len);
    return 0;
  }
  DeviceGuard dg(sh.device);
  CUDA_TRY(cudaMemcpyAsync(dst, sh.dev.rings + at, len,
                           cudaMemcpyDeviceToHost, sh.copy_stream));
  CUDA_TRY(cudaStreamSynchronize(sh.copy_stream));
  return 0;
  GUARD_END
}

int pcdn_retry_batch(pcdn_engine* e, uint64_t batch_id) {
  GUARD_BEGIN
  LOCK;
  if (!e->has_device) return fail(PCDN_ENODEV, "host-only engine");
  if (!(e->cfg.flags & PCDN_FLAG_OUTPUT_POOL)) return fail(PCDN_EINVAL, "only batches of an output-pool engine can be refused for space");
  const int si = find_slot_index(e, batch_id);
  if (si < 0) return fail(PCDN_ENOENT, "unknown batch id");
  if (e->inflight.empty() || e->inflight.front() != batch_id)
    return fail(PCDN_EINVAL, "only the oldest unreleased batch can be retried (release the older ones first)");
  Slot& s = e->slots[si];
  int rc = flush_journal(e);
  if (rc) return rc;
  uint32_t n = 0;
  for (Shard& sh : e->shards) {
    ShardSlot& ss = sh.slots[si];
    {
      DeviceGuard dg(sh.device);
      CUDA_TRY(cudaEventSynchronize(ss.ev_done));
    }
    if (ss.h_stats->status != 2) continue;   // this shard packed its share (or refused it for good): leave it alone
    if ((rc = launch_shard_pipeline(e, sh, (uint32_t)si, s.device_input ? s.n_msgs - ss.in.n_bcast : s.n_direct, s.devparse, false, true))) return rc;
    n++;
  }
  if (!n) return fail(PCDN_EINVAL, "the batch was not refused for space");
  return 0;
  GUARD_END
}

int pcdn_release_batch(pcdn_engine* e, uint64_t batch_id) {
  GUARD_BEGIN
  LOCK;
  const int si = find_slot_index(e, batch_id);
  if (si < 0) return fail(PCDN_ENOENT, "unknown batch id");
  if (e->inflight.empty() || e->inflight.front() != batch_id)
    return fail(PCDN_EINVAL, "batches must be released oldest first");
  Slot& s = e->slots[si];
  for (Shard& sh : e->shards) {
    DeviceGuard dg(sh.device);
    ShardSlot& ss = sh.slots[si];
    // A slot released without having been polled may still have its host→device staging copy queued:
    // its pinned staging buffers must not be refilled before that copy ran (device-input batches have
    // no host staging and stay fully asynchronous — the pipelined submit_device/release loop).
    if (!ss.polled && !s.device_input) CUDA_TRY(cudaEventSynchronize(e->sharded ? ss.ev_ingest : ss.ev_done));
    if (e->timeline && ss.timed && !e->timeline_async) {   // diagnostic: where this batch's stages ran on the shard's clock (blocks until the pack is done)
      CUDA_TRY(cudaEventSynchronize(ss.ev_done));
      float t[6];
      for (int k = 0; k < 6; k++) if (cudaEventElapsedTime(&t[k], sh.ev_base, ss.ev[k]) != cudaSuccess) { t[k] = -1.f; cudaGetLastError(); }
      std::fprintf(e->timeline, "shard %u batch %llu pack_stream %d ctrl_begin %.4f direct_end %.4f match_end %.4f offsets_end %.4f pack_begin %.4f pack_end %.4f\n",
                   sh.gindex, (unsigned long long)batch_id, (int)ss.on_pack_stream, t[0], t[1], t[2], t[3], t[4], t[5]);
      std::fflush(e->timeline);
    }
    // ring space may be reused only after the pack that filled it has finished
    CUDA_TRY(cudaStreamWaitEvent(sh.stream, ss.ev_done, 0));
    launch_release(sh.dev, ss.w.batch_units, ss.w.stats, sh.stream);
    CUDA_TRY(cudaGetLastError());
  }
  e->inflight.erase(e->inflight.begin());
  s.state = SLOT_FREE;
  // the last 'clone' of every frame of this batch is gone: permits back to the pool (pool.rs:44-52)
  e->inflight_bytes -= std::min(e->inflight_bytes, s.ingress_bytes);
  s.ingress_bytes = 0;
  e->stats.released_batches++;
  {
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - s.t_launch).count();
    e->stats.latency_ms_sum += ms;
    const uint64_t us = (uint64_t)(ms * 1000.0);
    int bucket = 0;
    while (bucket < 15 && us >= (16ull << bucket)) bucket++;
    e->stats.latency_hist_us[bucket]++;
  }
  return 0;
  GUARD_END
}

// ---- connection shards -------------------------------------------------------------------------
int pcdn_nccl_unique_id(void* out128) {
  GUARD_BEGIN
  if (!out128) return fail(PCDN_EINVAL, "null argument");
  const char* why = "";
  const NcclApi* nc = nccl_api(&why);
  if (!nc) return fail(PCDN_ENODEV, std::string("libnccl.so.2 not available: ") + why);
  NcclUniqueId id;
  NCCL_TRY(nc, nc->GetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
  return 0;
  GUARD_END
}
int pcdn_num_shards(pcdn_engine* e, uint32_t* n_local, uint32_t* n_world) {
  LOCK;
  if (n_local) *n_local = e->has_device ? (uint32_t)e->shards.size() : 0;
  if (n_world) *n_world = e->world_shards;
  return 0;
}
int pcdn_shard_info(pcdn_engine* e, uint32_t local_shard, pcdn_shard_desc* out) {
  GUARD_BEGIN
  LOCK;
  if (!out) return fail(PCDN_EINVAL, "null argument");
  std::memset(out, 0, sizeof(*out));
  out->shard_stride = e->geo.shard_N;
  out->ring_bytes = (e->cfg.flags & PCDN_FLAG_OUTPUT_POOL) ? e->pool_bytes : e->cfg.ring_bytes_per_conn;
  if (!e->has_device) {  // host-only mirror: geometry only
    if (local_shard != 0) return fail(PCDN_EINVAL, "no such local shard");
    out->global_index = e->first_shard; out->device = -1; out->conn_base = e->first_shard * e->geo.shard_N;
    return 0;
  }
  if (local_shard >= e->shards.size()) return fail(PCDN_EINVAL, "no such local shard");
  const Shard& sh = e->shards[local_shard];
  out->global_index = sh.gindex;
  out->device = sh.device;
  out->conn_base = sh.dev.conn_base;
  out->rings_dev = sh.dev.rings;
  out->rings_host = sh.h_rings;
  out->nccl_ranks = (uint32_t)sh.nccl_ranks;
  std::vector<uint32_t> v;
  out->n_conns = e->conns->shard_load(sh.gindex);
  return 0;
  GUARD_END
}

// ---- introspection ----------------------------------------------------------------------------
int pcdn_get_stats(pcdn_engine* e, pcdn_stats* out) {
  LOCK;
  e->stats.inflight_bytes = e->inflight_bytes;
  e->stats.kernel_launches = kernel_launches();
  *out = e->stats;
  return 0;
}
int pcdn_set_timing(pcdn_engine* e, int on) {
  LOCK;
  e->timing = on != 0 || e->timeline != nullptr;
  return 0;
}
int pcdn_ring_info(pcdn_engine* e, void** dev_base, uint64_t* ring_bytes, uint32_t* max_conns) {
  LOCK;
  if (dev_base) *dev_base = e->has_device ? (void*)e->shards[0].dev.rings : nullptr;  // first local shard (pcdn_shard_info for the others)
  if (ring_bytes) *ring_bytes = e->cfg.ring_bytes_per_conn;
  if (max_conns) *max_conns = e->geo.shard_max_conns;
  return 0;
}
int pcdn_host_rings(pcdn_engine* e, const void** host_base) {
  LOCK;
  uint8_t* h = e->has_device ? e->shards[0].h_rings : nullptr;
  if (host_base) *host_base = h;
  return h ? 0 : fail(PCDN_ENOENT, "rings live in device memory (PCDN_FLAG_HOST_RINGS not set)");
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
