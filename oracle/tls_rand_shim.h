/* Force-included (-include) ahead of the upstream sources for the _ref/*_tls.so build only:
 * pulls in <stdlib.h> first (libstdc++'s <cstdlib> #undefs rand), then renames rand() to the
 * thread-local replica in tls_rand.c.  The upstream sources themselves are untouched. */
#ifndef S2TC_TLS_RAND_SHIM_H
#define S2TC_TLS_RAND_SHIM_H
#include <stdlib.h>
#ifdef __cplusplus
extern "C" {
#endif
int s2tc_tls_rand(void);
#ifdef __cplusplus
}
#endif
#define rand s2tc_tls_rand
#endif
