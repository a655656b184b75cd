// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in umbrella header for ADOL-C.
#ifndef ORACLE_STUB_ADOLC_H
#define ORACLE_STUB_ADOLC_H
#include "adolc/adouble.h"
#include "adolc/taping.h"
#include "adolc/drivers/drivers.h"
#endif
