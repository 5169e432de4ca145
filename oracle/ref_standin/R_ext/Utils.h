/* R_ext/Utils.h (STAND-IN, test infrastructure).  R_CheckUserInterrupt is reached only through
 * Rcpp::checkUserInterrupt(), a no-op here. */
#ifndef STANDIN_R_EXT_UTILS_H
#define STANDIN_R_EXT_UTILS_H
#endif
