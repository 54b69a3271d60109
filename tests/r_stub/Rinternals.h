/* Minimal DECLARATIONS of the R C API used by r/mcb200_glue.c, for a syntax / type check of the glue in an image
 * without R (tests/test_host.py::test_r_glue_type_checks_against_the_abi).  Not R's header; nothing is defined. */
#ifndef MCB_R_STUB_H
#define MCB_R_STUB_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned int SEXPTYPE;
enum { LGLSXP = 10, INTSXP = 13, REALSXP = 14, VECSXP = 19 };
extern SEXP R_NilValue, R_DimSymbol;
extern int NA_INTEGER;
extern double NA_REAL;
R_xlen_t XLENGTH(SEXP);
int* INTEGER(SEXP);
double* REAL(SEXP);
int* LOGICAL(SEXP);
char* R_alloc(size_t, int);
SEXP allocVector(SEXPTYPE, R_xlen_t);
SEXP allocMatrix(SEXPTYPE, int, int);
SEXP PROTECT(SEXP);
void UNPROTECT(int);
SEXP mkNamed(SEXPTYPE, const char**);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP ScalarReal(double);
int asInteger(SEXP);
double asReal(SEXP);
int asLogical(SEXP);
SEXP getAttrib(SEXP, SEXP);
int isNull(SEXP);
SEXP GetOption1(SEXP);
SEXP install(const char*);
SEXP coerceVector(SEXP, SEXPTYPE);
void Rf_error(const char*, ...);
#endif
