/* A C host driving the engine on a GPU exactly the way the Rust broker would:
 * state calls, handle_broadcast_message / handle_direct_message with opaque frames, flush, poll,
 * read the spans (once through pcdn_read, once in place with PCDN_FLAG_HOST_RINGS), release.
 * Checks the cdn-proto framing (u32 BE length || raw, protocols/mod.rs:366-385), recipients, FIFO. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pcdn_fanout.h"

#define CHECK(cond)                                                         \
  do {                                                                      \
    if (!(cond)) {                                                          \
      fprintf(stderr, "FAILED %s:%d: %s (last error: %s)\n", __FILE__, __LINE__, #cond, pcdn_last_error()); \
      return 1;                                                             \
    }                                                                       \
  } while (0)

static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

/* shards: 0 = one GPU; 2 = the SAME calls on an engine whose connections are spread over two shards
 * (pcdn_config.devices): GPUs 0 and 1 with the library's ncclBroadcast ingest when the box has two,
 * else both shards on GPU 0 with PCDN_INGEST_HOST */
static int run(uint32_t flags, int shards) {
  pcdn_config cfg;
  pcdn_config_default(&cfg);
  cfg.device = 0;
  cfg.max_conns = 16;
  cfg.max_keys = 64;
  cfg.ring_bytes_per_conn = 1 << 16;
  cfg.identity = "pub/priv";
  cfg.flags = flags;
  pcdn_engine* e = NULL;
  const int32_t two_gpus[2] = {0, 1}, one_gpu[2] = {0, 0};
  if (shards) {
    cfg.n_devices = 2; cfg.devices = two_gpus; cfg.ingest = PCDN_INGEST_NCCL;
    if (pcdn_create(&cfg, &e) != 0) {   /* a one-GPU box */
      cfg.devices = one_gpu; cfg.ingest = PCDN_INGEST_HOST;
      CHECK(pcdn_create(&cfg, &e) == 0);
    }
    uint32_t nl = 0, nw = 0;
    CHECK(pcdn_num_shards(e, &nl, &nw) == 0 && nl == 2 && nw == 2);
  } else {
    CHECK(pcdn_create(&cfg, &e) == 0);
  }

  const uint8_t k0[8] = {0}, k1[8] = {1}, k2[8] = {2};
  const uint16_t t0[1] = {0}, t1[1] = {1};
  pcdn_conn c0, c1, c2;
  CHECK(pcdn_add_user(e, k0, 8, t0, 1, &c0) == 0);
  CHECK(pcdn_add_user(e, k1, 8, t0, 1, &c1) == 0);
  CHECK(pcdn_add_user(e, k2, 8, t1, 1, &c2) == 0);

  uint8_t m1[1000], m2[37], m3[5];   /* opaque to the engine (R1): any bytes */
  for (size_t i = 0; i < sizeof m1; i++) m1[i] = (uint8_t)(i * 7 + 1);
  for (size_t i = 0; i < sizeof m2; i++) m2[i] = (uint8_t)(200 - i);
  memcpy(m3, "hello", 5);
  CHECK(pcdn_handle_broadcast_message(e, t0, 1, m1, sizeof m1, 0) == 0);   /* -> c0, c1 */
  CHECK(pcdn_handle_direct_message(e, k1, 8, m2, sizeof m2, 0) == 0);      /* -> c1 */
  CHECK(pcdn_handle_broadcast_message(e, t1, 1, m3, sizeof m3, 0) == 0);   /* -> c2 */
  CHECK(pcdn_handle_direct_message(e, (const uint8_t*)"nobody!!", 8, m3, sizeof m3, 0) == 0);  /* dropped */

  uint64_t batch = 0;
  CHECK(pcdn_flush(e, &batch) == 0 && batch != 0);
  pcdn_batch_result r;
  CHECK(pcdn_poll(e, batch, &r, 1) == 0);
  CHECK(r.status == 0 && r.n_msgs == 4 && r.n_deliveries == 4 && r.n_direct_dropped == 1 && r.n_overflow == 0);
  CHECK(r.bytes_out == (4 + sizeof m1) * 2 + (4 + sizeof m2) + (4 + sizeof m3));
  CHECK(r.n_spans == 3);

  const void* hbase = NULL;
  const int in_place = pcdn_host_rings(e, &hbase) == 0;
  CHECK(in_place == ((flags & PCDN_FLAG_HOST_RINGS) != 0));
  if (shards) {   /* connections went to the least-loaded shard: ids 0, stride, 1 */
    pcdn_shard_desc d0, d1;
    CHECK(pcdn_shard_info(e, 0, &d0) == 0 && pcdn_shard_info(e, 1, &d1) == 0);
    CHECK(d0.conn_base == 0 && d1.conn_base == d0.shard_stride && d0.n_conns == 2 && d1.n_conns == 1);
    CHECK(c1 == d1.conn_base && c0 == 0 && c2 == 1);
  }
  int seen = 0;
  for (uint32_t i = 0; i < r.n_spans; i++) {
    const pcdn_span* s = &r.spans[i];
    uint8_t* buf = (uint8_t*)malloc(s->len);
    if (in_place && !shards) memcpy(buf, (const uint8_t*)hbase + (size_t)s->conn * cfg.ring_bytes_per_conn + s->ring_off, s->len);
    else CHECK(pcdn_read(e, s->conn, s->ring_off, s->len, buf) == 0);
    if (s->conn == c0) {
      CHECK(s->n_records == 1 && be32(buf) == sizeof m1 && memcmp(buf + 4, m1, sizeof m1) == 0);
      seen |= 1;
    } else if (s->conn == c1) {  /* FIFO per connection = batch order (R9): m1 then m2 */
      CHECK(s->n_records == 2 && be32(buf) == sizeof m1 && memcmp(buf + 4, m1, sizeof m1) == 0);
      const size_t second = (4 + sizeof m1 + PCDN_RECORD_ALIGN - 1) / PCDN_RECORD_ALIGN * PCDN_RECORD_ALIGN;
      CHECK(be32(buf + second) == sizeof m2 && memcmp(buf + second + 4, m2, sizeof m2) == 0);
      seen |= 2;
    } else if (s->conn == c2) {
      CHECK(s->n_records == 1 && be32(buf) == sizeof m3 && memcmp(buf + 4, m3, sizeof m3) == 0);
      seen |= 4;
    } else {
      CHECK(!"span for an unknown connection");
    }
    free(buf);
  }
  CHECK(seen == 7);
  CHECK(pcdn_release_batch(e, batch) == 0);
  pcdn_destroy(e);
  return 0;
}

int main(void) {
  if (run(0, 0)) return 1;
  if (run(PCDN_FLAG_HOST_RINGS, 0)) return 1;
  if (run(0, 2)) return 1;
  if (run(PCDN_FLAG_HOST_RINGS, 2)) return 1;
  printf("gpu_roundtrip ok\n");
  return 0;
}
