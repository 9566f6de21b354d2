/* Stand-in for R's <R_ext/Rdynload.h> -- TEST INFRASTRUCTURE ONLY; see ../Rinternals.h. */
#ifndef MINI_RDYNLOAD_H
#define MINI_RDYNLOAD_H
typedef void *(*DL_FUNC)(void);
typedef struct {
    const char *name;
    DL_FUNC fun;
    int numArgs;
} R_CallMethodDef;
#endif
