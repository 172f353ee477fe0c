/* Plain-C consumer of include/pcdn_fanout.h: what a cgo / Rust-bindgen / JNI binding sees.
 * Built with gcc -std=c11 -pedantic -Werror by tests/test_c_abi.py and run against a host-only
 * engine (device = -1: state calls + debug lookups work without a GPU; routing answers PCDN_ENODEV). */
#include <stdio.h>
#include <string.h>

#include "pcdn_fanout.h"

#define CHECK(cond)                                                         \
  do {                                                                      \
    if (!(cond)) {                                                          \
      fprintf(stderr, "FAILED %s:%d: %s (last error: %s)\n", __FILE__, __LINE__, #cond, pcdn_last_error()); \
      return 1;                                                             \
    }                                                                       \
  } while (0)

int main(void) {
  pcdn_config cfg;
  pcdn_config_default(&cfg);
  CHECK(cfg.struct_size == sizeof(pcdn_config));
  cfg.device = -1;
  cfg.max_conns = 64;
  cfg.max_keys = 64;
  cfg.identity = "pub/priv";
  pcdn_engine* e = NULL;
  CHECK(pcdn_create(&cfg, &e) == 0 && e != NULL);
  CHECK(pcdn_abi_version() != 0);

  const uint8_t alice[8] = {1, 0, 0, 0, 0, 0, 0, 0}, bob[8] = {2, 0, 0, 0, 0, 0, 0, 0};
  const uint16_t t0[1] = {0}, t01[2] = {0, 1};
  pcdn_conn a = 0, b = 0, br = 0;
  CHECK(pcdn_add_user(e, alice, 8, t0, 1, &a) == 0);
  CHECK(pcdn_add_user(e, bob, 8, t01, 2, &b) == 0 && a != b);
  CHECK(pcdn_add_broker(e, "other/broker", &br) == 0);
  CHECK(pcdn_subscribe_broker_to(e, "other/broker", t0, 1) == 0);

  uint32_t users = 0, brokers = 0;
  CHECK(pcdn_num_users(e, &users, &brokers) == 0 && users == 2 && brokers == 1);

  /* Connections::get_interested_by_topic on the mirror */
  pcdn_conn out[8];
  uint32_t n = 0;
  CHECK(pcdn_debug_interested(e, t0, 1, /*to_users_only=*/0, out, 8, &n) == 0 && n == 3);
  CHECK(pcdn_debug_interested(e, t0, 1, /*to_users_only=*/1, out, 8, &n) == 0 && n == 2);
  const uint16_t t1[1] = {1};
  CHECK(pcdn_debug_interested(e, t1, 1, 0, out, 8, &n) == 0 && n == 1 && out[0] == b);

  /* direct map */
  int kind = -1;
  pcdn_conn rc = 0;
  CHECK(pcdn_debug_route(e, alice, 8, &kind, &rc) == 0 && kind == 1 && rc == a);
  CHECK(pcdn_remove_user(e, alice, 8) == 0);
  CHECK(pcdn_debug_route(e, alice, 8, &kind, &rc) == 0 && kind == 0);

  /* routing needs the device: must fail loudly, never fall back to a CPU path */
  const uint8_t raw[4] = {0, 0, 0, 0};
  CHECK(pcdn_handle_broadcast_message(e, t0, 1, raw, 4, 0) == PCDN_ENODEV);
  CHECK(strlen(pcdn_last_error()) > 0);

  pcdn_destroy(e);
  printf("abi_smoke ok\n");
  return 0;
}
