/* Eight host threads hammer the state ABI (add / kick / remove / (un)subscribe / lookups) of a host-only
 * engine; the whole library (engine.cu + kernels.cu + host sources) is built with -fsanitize=thread by
 * tests/test_host_sanitizers.py.  SURVEY 8b "Threading": the engine serialises callers internally. */
#include <pthread.h>
#include <stdio.h>
#include <string.h>
#include "pcdn_fanout.h"
static pcdn_engine* e;
static void* worker(void* arg) {
  long id = (long)arg;
  for (int it = 0; it < 20000; it++) {
    uint8_t key[8]; memset(key, 0, 8); key[0] = (uint8_t)((it * 7 + id) % 200); key[1] = (uint8_t)id;
    uint16_t t[2] = {(uint16_t)(it % 16), (uint16_t)((it + 3) % 16)};
    pcdn_conn c;
    switch ((it + id) % 6) {
      case 0: pcdn_add_user(e, key, 8, t, 2, &c); break;
      case 1: pcdn_remove_user(e, key, 8); break;
      case 2: pcdn_subscribe_user_to(e, key, 8, t, 2); break;
      case 3: pcdn_unsubscribe_user_from(e, key, 8, t, 1); break;
      case 4: { pcdn_conn out[64]; uint32_t n; pcdn_debug_interested(e, t, 2, 0, out, 64, &n); break; }
      default: { int kind; pcdn_debug_route(e, key, 8, &kind, &c); uint32_t u, b; pcdn_num_users(e, &u, &b); }
    }
  }
  return NULL;
}
int main(void) {
  pcdn_config cfg; pcdn_config_default(&cfg); cfg.device = -1; cfg.max_conns = 2048; cfg.max_keys = 4096; cfg.identity = "a/b";
  if (pcdn_create(&cfg, &e)) { printf("create failed %s\n", pcdn_last_error()); return 1; }
  pthread_t th[8];
  for (long i = 0; i < 8; i++) pthread_create(&th[i], NULL, worker, (void*)i);
  for (int i = 0; i < 8; i++) pthread_join(th[i], NULL);
  uint32_t u, b; pcdn_num_users(e, &u, &b); printf("tsan driver ok, %u users\n", u);
  pcdn_destroy(e); return 0;
}
