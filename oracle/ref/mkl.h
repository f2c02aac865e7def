// Build scaffolding for compiling the UNMODIFIED reference (oracle A) in this image:
// maps the Intel MKL entry points the reference calls onto the LP64 OpenBLAS that ships
// inside the scipy wheel (symbols carry a scipy_ prefix).  Not product code; see
// oracle/ref/build_ref.sh.  (SURVEY.md appendix E.)
#pragma once
typedef int MKL_INT;
extern "C" {
enum CBLAS_LAYOUT {CblasRowMajor=101, CblasColMajor=102};
enum CBLAS_TRANSPOSE {CblasNoTrans=111, CblasTrans=112, CblasConjTrans=113};
enum CBLAS_UPLO {CblasUpper=121, CblasLower=122};
enum CBLAS_DIAG {CblasNonUnit=131, CblasUnit=132};
enum CBLAS_SIDE {CblasLeft=141, CblasRight=142};
#define LAPACK_ROW_MAJOR 101
void scipy_cblas_sscal(int, float, float*, int);
void scipy_cblas_saxpy(int, float, const float*, int, float*, int);
void scipy_cblas_scopy(int, const float*, int, float*, int);
double scipy_cblas_dsdot(int, const float*, int, const float*, int);
float scipy_cblas_sdsdot(int, float, const float*, int, const float*, int);
double scipy_cblas_ddot(int, const double*, int, const double*, int);
void scipy_cblas_sgemm(CBLAS_LAYOUT, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int,int,int, float, const float*, int, const float*, int, float, float*, int);
void scipy_cblas_dgemm(CBLAS_LAYOUT, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int,int,int, double, const double*, int, const double*, int, double, double*, int);
void scipy_cblas_sgemv(CBLAS_LAYOUT, CBLAS_TRANSPOSE, int,int, float, const float*, int, const float*, int, float, float*, int);
void scipy_cblas_strmm(CBLAS_LAYOUT, CBLAS_SIDE, CBLAS_UPLO, CBLAS_TRANSPOSE, CBLAS_DIAG, int,int, float, const float*, int, float*, int);
int scipy_LAPACKE_spotrf(int, char, int, float*, int);
int scipy_LAPACKE_spotrs(int, char, int, int, const float*, int, float*, int);
int scipy_LAPACKE_strtri(int, char, char, int, float*, int);
void scipy_openblas_set_num_threads(int);
}
#define cblas_sscal scipy_cblas_sscal
#define cblas_saxpy scipy_cblas_saxpy
#define cblas_scopy scipy_cblas_scopy
#define cblas_dsdot scipy_cblas_dsdot
#define cblas_sdsdot scipy_cblas_sdsdot
#define cblas_ddot scipy_cblas_ddot
#define cblas_sgemm scipy_cblas_sgemm
#define cblas_dgemm scipy_cblas_dgemm
#define cblas_sgemv scipy_cblas_sgemv
#define cblas_strmm scipy_cblas_strmm
#define LAPACKE_spotrf scipy_LAPACKE_spotrf
#define LAPACKE_spotrs scipy_LAPACKE_spotrs
#define LAPACKE_strtri scipy_LAPACKE_strtri
static inline int mkl_set_num_threads_local(int n){ scipy_openblas_set_num_threads(n); return 0; }
