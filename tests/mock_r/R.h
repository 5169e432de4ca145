/* R.h (MOCK) -- see Rinternals.h in this directory. */
#ifndef MOCK_R_H
#define MOCK_R_H
#include <math.h>
#include <stdlib.h>
#include <string.h>
#endif
