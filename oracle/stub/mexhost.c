/*
 * oracle/stub/mexhost.c -- TEST INFRASTRUCTURE, not product code.
 *
 * A "fake MATLAB host": implements the mx / mex functions declared in oracle/stub/mex.h and a
 * generic gateway, mexhost_call(), through which a test (Python/ctypes or C) can invoke any
 * mexFunction-style entry point with plain double arrays, exactly the way SPH_Poiseuille.m calls
 * feval(mex_name, ...) (reference call sites: SPH_Poiseuille.m:167,169,366,380,388,398,406,419,428).
 *
 * Semantics mirrored from MATLAB:
 *   - inputs are borrowed, const;
 *   - every output is a fresh zero-initialised array owned by the caller on return;
 *   - mexErrMsgIdAndTxt does not return: it records (id, message), frees everything the MEX call
 *     allocated through mxMalloc/mxCreate* (MATLAB's auto-free on error) and longjmps back here.
 *
 * Used by: oracle/_ref/libsph_ref.so (reference sources compiled unmodified) and by the stub-host
 * build of this repo's MEX shims (tests only).
 */
#include "mex.h"

#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct mxArray_tag {
    int is_char;
    char *str;      /* owned when is_char */
    size_t m, n;
    double *data;   /* owned unless borrowed */
    int borrowed;   /* input wrapper around caller memory */
    int tracked;    /* index+1 in the per-call registry, 0 if not tracked */
};

/* ---- per-call allocation registry (so an error unwinding frees what the MEX call made) ---- */
typedef struct {
    void **ptr;
    int *kind; /* 0 raw memory, 1 mxArray */
    size_t count, cap;
} registry_t;

static __thread registry_t g_reg;
static __thread jmp_buf g_jmp;
static __thread int g_active = 0;
static __thread char g_errid[256];
static __thread char g_errmsg[1024];

static void reg_add(void *p, int kind, mxArray *arr)
{
    if (!g_active || !p) return;
    if (g_reg.count == g_reg.cap) {
        size_t ncap = g_reg.cap ? g_reg.cap * 2 : 64;
        g_reg.ptr = (void **)realloc(g_reg.ptr, ncap * sizeof(void *));
        g_reg.kind = (int *)realloc(g_reg.kind, ncap * sizeof(int));
        g_reg.cap = ncap;
    }
    g_reg.ptr[g_reg.count] = p;
    g_reg.kind[g_reg.count] = kind;
    g_reg.count++;
    if (arr) arr->tracked = (int)g_reg.count;
}

static void reg_remove(void *p)
{
    size_t i;
    if (!g_active || !p) return;
    for (i = g_reg.count; i-- > 0;) {
        if (g_reg.ptr[i] == p) {
            g_reg.ptr[i] = NULL;
            return;
        }
    }
}

static void reg_replace(void *oldp, void *newp)
{
    size_t i;
    if (!g_active) return;
    for (i = g_reg.count; i-- > 0;) {
        if (g_reg.ptr[i] == oldp) {
            g_reg.ptr[i] = newp;
            return;
        }
    }
    reg_add(newp, 0, NULL);
}

static void free_array_raw(mxArray *a)
{
    if (!a) return;
    if (a->str) free(a->str);
    if (a->data && !a->borrowed) free(a->data);
    free(a);
}

static void reg_free_all(void)
{
    size_t i;
    for (i = 0; i < g_reg.count; ++i) {
        if (!g_reg.ptr[i]) continue;
        if (g_reg.kind[i] == 1) free_array_raw((mxArray *)g_reg.ptr[i]);
        else free(g_reg.ptr[i]);
    }
    g_reg.count = 0;
}

/* ------------------------------------ mx API ------------------------------------ */
double *mxGetDoubles(const mxArray *a) { return a ? a->data : NULL; }
double mxGetScalar(const mxArray *a) { return (a && a->data && a->m * a->n > 0) ? a->data[0] : 0.0; }
size_t mxGetM(const mxArray *a) { return a ? a->m : 0; }
size_t mxGetN(const mxArray *a) { return a ? a->n : 0; }
size_t mxGetNumberOfElements(const mxArray *a) { return a ? a->m * a->n : 0; }
int mxIsDouble(const mxArray *a) { return a && !a->is_char; }
int mxIsChar(const mxArray *a) { return a && a->is_char; }

int mxGetString(const mxArray *a, char *buf, mwSize buflen)
{
    size_t len;
    if (!a || !a->is_char || !buf || buflen == 0) return 1;
    len = strlen(a->str);
    if (len + 1 > buflen) {
        memcpy(buf, a->str, buflen - 1);
        buf[buflen - 1] = '\0';
        return 1;
    }
    memcpy(buf, a->str, len + 1);
    return 0;
}

mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity flag)
{
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    size_t cnt = m * n;
    (void)flag;
    a->m = m;
    a->n = n;
    a->data = (double *)calloc(cnt ? cnt : 1, sizeof(double));
    reg_add(a, 1, a);
    return a;
}

mxArray *mxCreateDoubleScalar(double v)
{
    mxArray *a = mxCreateDoubleMatrix(1, 1, mxREAL);
    a->data[0] = v;
    return a;
}

mxArray *mxCreateString(const char *s)
{
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    size_t len = strlen(s);
    a->is_char = 1;
    a->str = (char *)malloc(len + 1);
    memcpy(a->str, s, len + 1);
    a->m = 1;
    a->n = len;
    reg_add(a, 1, a);
    return a;
}

mxArray *mxDuplicateArray(const mxArray *src)
{
    mxArray *a;
    if (!src) return NULL;
    if (src->is_char) return mxCreateString(src->str);
    a = mxCreateDoubleMatrix(src->m, src->n, mxREAL);
    if (src->m * src->n > 0) memcpy(a->data, src->data, src->m * src->n * sizeof(double));
    return a;
}

void mxDestroyArray(mxArray *a)
{
    if (!a) return;
    reg_remove(a);
    free_array_raw(a);
}

void *mxMalloc(size_t n)
{
    void *p = malloc(n ? n : 1);
    reg_add(p, 0, NULL);
    return p;
}

void *mxCalloc(size_t n, size_t sz)
{
    void *p = calloc(n ? n : 1, sz ? sz : 1);
    reg_add(p, 0, NULL);
    return p;
}

void *mxRealloc(void *p, size_t n)
{
    void *old = p;
    void *q = realloc(p, n ? n : 1);
    reg_replace(old, q);
    return q;
}

void mxFree(void *p)
{
    if (!p) return;
    reg_remove(p);
    free(p);
}

void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...)
{
    va_list ap;
    snprintf(g_errid, sizeof(g_errid), "%s", id ? id : "");
    va_start(ap, fmt);
    vsnprintf(g_errmsg, sizeof(g_errmsg), fmt ? fmt : "", ap);
    va_end(ap);
    if (!g_active) {
        fprintf(stderr, "mexErrMsgIdAndTxt outside mexhost_call: %s: %s\n", g_errid, g_errmsg);
        abort();
    }
    longjmp(g_jmp, 1);
}

/* ------------------------------- the generic gateway ------------------------------- */
typedef struct {
    int is_char;        /* 1: str is the value; 0: (data, m, n) column-major doubles */
    const char *str;
    size_t m, n;
    const double *data;
} mexhost_arg;

typedef struct {
    size_t m, n;
    double *data;       /* malloc'ed; release with mexhost_free_outputs */
} mexhost_out;

typedef void (*mexhost_fn)(int, mxArray **, int, const mxArray **);

#define MEXHOST_MAX_ARGS 64

/*
 * Returns 0 on success, 1 when the MEX code raised mexErrMsgIdAndTxt (errid/errmsg filled),
 * 2 on a gateway usage error.
 */
int mexhost_call(mexhost_fn fn, int nlhs, mexhost_out *outs, int nrhs, const mexhost_arg *args,
                 char *errid, size_t errid_len, char *errmsg, size_t errmsg_len)
{
    mxArray in_store[MEXHOST_MAX_ARGS];
    const mxArray *prhs[MEXHOST_MAX_ARGS];
    mxArray *plhs[MEXHOST_MAX_ARGS];
    int i;

    if (errid && errid_len) errid[0] = '\0';
    if (errmsg && errmsg_len) errmsg[0] = '\0';
    if (!fn || nrhs < 0 || nrhs > MEXHOST_MAX_ARGS || nlhs < 0 || nlhs > MEXHOST_MAX_ARGS) return 2;
    if (g_active) return 2; /* not re-entrant */

    memset(in_store, 0, sizeof(in_store));
    for (i = 0; i < nrhs; ++i) {
        in_store[i].is_char = args[i].is_char;
        in_store[i].str = (char *)args[i].str;
        in_store[i].m = args[i].is_char ? 1 : args[i].m;
        in_store[i].n = args[i].is_char ? (args[i].str ? strlen(args[i].str) : 0) : args[i].n;
        in_store[i].data = (double *)args[i].data;
        in_store[i].borrowed = 1;
        prhs[i] = &in_store[i];
    }
    for (i = 0; i < MEXHOST_MAX_ARGS; ++i) plhs[i] = NULL;

    g_reg.count = 0;
    g_active = 1;
    if (setjmp(g_jmp) != 0) {
        g_active = 0;
        reg_free_all();
        if (errid && errid_len) snprintf(errid, errid_len, "%s", g_errid);
        if (errmsg && errmsg_len) snprintf(errmsg, errmsg_len, "%s", g_errmsg);
        return 1;
    }
    fn(nlhs, plhs, nrhs, prhs);

    for (i = 0; i < nlhs; ++i) {
        if (plhs[i]) {
            outs[i].m = plhs[i]->m;
            outs[i].n = plhs[i]->n;
            outs[i].data = plhs[i]->data;
            plhs[i]->data = NULL;
            reg_remove(plhs[i]);
            free_array_raw(plhs[i]);
        } else {
            outs[i].m = outs[i].n = 0;
            outs[i].data = NULL;
        }
    }
    /* Anything still registered was leaked by the MEX code (MATLAB would free it too). */
    reg_free_all();
    g_active = 0;
    return 0;
}

void mexhost_free_outputs(int nlhs, mexhost_out *outs)
{
    int i;
    for (i = 0; i < nlhs; ++i) {
        if (outs[i].data) free(outs[i].data);
        outs[i].data = NULL;
    }
}
