// TEST INFRASTRUCTURE ONLY (oracle build). See gsl_interp.h in this directory.
#include "gsl/gsl_interp.h"
