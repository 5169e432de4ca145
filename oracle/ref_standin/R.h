/* R.h (STAND-IN, test infrastructure -- see RcppArmadillo.h in this directory).  The reference translation unit
 * needs nothing from it beyond what Rmath.h declares. */
#ifndef STANDIN_R_H
#define STANDIN_R_H
#include <math.h>
#endif
