/* tests/_rstub/R.h — see Rinternals.h in this directory (a fake of R's C API for compiling the .Call shim) */
#ifndef RSTUB_R_H
#define RSTUB_R_H
#include "Rinternals.h"
#endif
