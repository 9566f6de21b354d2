/* Stand-in for R's <R.h> -- TEST INFRASTRUCTURE ONLY; see Rinternals.h in this directory. */
#ifndef MINI_R_H
#define MINI_R_H
#include <stdlib.h>
#include <string.h>
#endif
