/* R_ext/Rdynload.h (MOCK) -- registration types used by R_init_<pkg>; see ../Rinternals.h. */
#ifndef MOCK_RDYNLOAD_H
#define MOCK_RDYNLOAD_H
#ifdef __cplusplus
extern "C" {
#endif
typedef void *(*DL_FUNC)(void);
typedef struct {
  const char *name;
  DL_FUNC fun;
  int numArgs;
} R_CallMethodDef;
typedef struct mock_dllinfo {
  const R_CallMethodDef *call_methods;
  int n_call_methods;
  int dynamic_symbols; /* -1 = never set */
} DllInfo;
typedef int Rboolean;
#ifndef FALSE
#define FALSE 0
#define TRUE 1
#endif
int R_registerRoutines(DllInfo *, const void *c_methods, const R_CallMethodDef *call_methods, const void *f_methods,
                       const void *ext_methods);
Rboolean R_useDynamicSymbols(DllInfo *, Rboolean);
#ifdef __cplusplus
}
#endif
#endif
