#include "gsl/gsl_sf.h"
