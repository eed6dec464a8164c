/* tests/_rstub/R_ext/Rdynload.h — registration records of R's dynamic loader (fake; see ../Rinternals.h) */
#ifndef RSTUB_RDYNLOAD_H
#define RSTUB_RDYNLOAD_H
#include "../Rinternals.h"
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct { const char *name; DL_FUNC fun; int numArgs; void *types; } R_CMethodDef;
typedef struct rstub_dllinfo { const R_CallMethodDef *call; int dynamic; } DllInfo;
int R_registerRoutines(DllInfo *info, const R_CMethodDef *c, const R_CallMethodDef *call, const void *f, const void *e);
Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value);
#endif
