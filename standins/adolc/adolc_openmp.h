// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in umbrella header for ADOL-C.
#ifndef ORACLE_STUB_ADOLC_OPENMP_H
#define ORACLE_STUB_ADOLC_OPENMP_H
#include "adolc/adouble.h"
#include "adolc/taping.h"
#include "adolc/drivers/drivers.h"
// real ADOL-C: `firstprivate(ADOLC_OpenMP_Handler)`, a clause for `#pragma omp parallel`; nothing to privatise without taping
#ifndef ADOLC_OPENMP
#define ADOLC_OPENMP
#endif
#endif
