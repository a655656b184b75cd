// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for ADOL-C <adolc/taping.h>: taping is a no-op.
#ifndef ORACLE_STUB_ADOLC_TAPING_H
#define ORACLE_STUB_ADOLC_TAPING_H
#include "adolc/adouble.h"
inline int trace_on(int, int = 0) { return 0; }
inline void trace_off(int = 0) {}
#endif
