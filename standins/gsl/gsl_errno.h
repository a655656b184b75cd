// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for <gsl/gsl_errno.h>.
#ifndef ORACLE_STUB_GSL_ERRNO_H
#define ORACLE_STUB_GSL_ERRNO_H
enum { GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_CONTINUE = -2, GSL_EDOM = 1, GSL_ERANGE = 2, GSL_EINVAL = 4 };
typedef void gsl_error_handler_t(const char *, const char *, int, int);
inline gsl_error_handler_t *gsl_set_error_handler_off(void) { return 0; }
inline gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *) { return 0; }
inline const char *gsl_strerror(const int) { return "oracle gsl stub"; }
#endif
