/* utils.h -- logging, timing and allocation helpers of the LSSP API (reference include/utils.h). */
#ifndef LSSP_UTILS_H
#define LSSP_UTILS_H

#include <assert.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "config.h"

extern int lssp_verbosity;

#define lssp_unused(x) (void)(x)

void lssp_set_log(FILE *io);
int lssp_comp_int_asc(const void *p, const void *n);
int lssp_comp_int_des(const void *p, const void *n);
double lssp_get_time();
double lssp_get_mem_usage(double *peak);
int lssp_printf(const char *fmt, ...);
void lssp_error(int code, const char *fmt, ...);   /* prints and exit(code) when code != 0 */
void lssp_warning(const char *fmt, ...);

template <typename T> T *lssp_malloc(const int n)
{
    assert(n >= 0);
    if (n == 0) return NULL;
    T *p = (T *)malloc(n * sizeof(T));
    if (p == NULL) lssp_error(1, "lssp: failed to malloc %g MB memory: %s %d.\n", n * sizeof(T) / 1048576., __FILE__, __LINE__);
    return p;
}

template <typename T> T *lssp_calloc(const int n)
{
    assert(n >= 0);
    if (n == 0) return NULL;
    T *p = (T *)calloc(n, sizeof(T));
    if (p == NULL) lssp_error(1, "lssp: failed to calloc %g MB memory: %s %d.\n", n * sizeof(T) / 1048576., __FILE__, __LINE__);
    return p;
}

template <typename T> T *lssp_realloc(T *old, const int n)
{
    assert(n >= 0);
    if (n == 0) return NULL;
    T *p = (T *)realloc((void *)old, n * sizeof(T));
    if (p == NULL) lssp_error(1, "lssp: failed to realloc %g MB memory: %s %d.\n", n * sizeof(T) / 1048576., __FILE__, __LINE__);
    return p;
}

template <typename T> void lssp_free(T *&p)
{
    if (p == NULL) return;
    free(p);
    p = NULL;
}

template <typename T> void lssp_memcpy_on(T *dst, const T *src, const int n)
{
    assert(n >= 0);
    if (n == 0) return;
    assert(dst != NULL && src != NULL);
    memcpy(dst, src, n * sizeof(T));
}

template <typename T> T *lssp_copy_on(const T *src, const int n)
{
    assert(n >= 0);
    if (n == 0) return NULL;
    assert(src != NULL);
    T *dst = lssp_malloc<T>(n);
    lssp_memcpy_on<T>(dst, src, n);
    return dst;
}

#endif
