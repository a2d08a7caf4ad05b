/* Minimal stand-in for the GSL entry points the reference's util.cpp calls
 * (/root/reference/util.cpp:202-215, 229-260).  TEST INFRASTRUCTURE ONLY: it exists so that the
 * UNMODIFIED reference sources can be compiled where they lie into oracle/_ref/ (GSL is not
 * installed in this image).  Nothing in the product links against it.
 *
 * gsl_linalg_LU_decomp -> LAPACK dgetrf_ (partial pivoting, like GSL), sign from the pivots
 * gsl_linalg_LU_det    -> signum * prod(diag(LU)) in index order (GSL's definition)
 * gsl_eigen_symmv      -> LAPACK dsyev_, eigenvectors returned as COLUMNS of evec (GSL's layout)
 */
#ifndef PIPSORT_ORACLE_GSL_SHIM_H
#define PIPSORT_ORACLE_GSL_SHIM_H

#include <stdlib.h>
#include <string.h>

extern "C" {
void dgetrf_(const int* m, const int* n, double* a, const int* lda, int* ipiv, int* info);
void dsyev_(const char* jobz, const char* uplo, const int* n, double* a, const int* lda, double* w,
            double* work, const int* lwork, int* info);
}

typedef struct {
    size_t size1, size2, tda;
    double* data;
} gsl_matrix;
typedef struct {
    size_t size;
    double* data;
} gsl_vector;
typedef struct {
    size_t size;
    int* data;
} gsl_permutation;
typedef struct {
    size_t size;
} gsl_eigen_symmv_workspace;

static inline gsl_matrix* gsl_matrix_calloc(size_t n1, size_t n2) {
    gsl_matrix* m = (gsl_matrix*)malloc(sizeof(gsl_matrix));
    m->size1 = n1;
    m->size2 = n2;
    m->tda = n2;
    m->data = (double*)calloc(n1 * n2 ? n1 * n2 : 1, sizeof(double));
    return m;
}
static inline void gsl_matrix_free(gsl_matrix* m) {
    if (m) {
        free(m->data);
        free(m);
    }
}
static inline void gsl_matrix_set(gsl_matrix* m, size_t i, size_t j, double x) { m->data[i * m->tda + j] = x; }
static inline double gsl_matrix_get(const gsl_matrix* m, size_t i, size_t j) { return m->data[i * m->tda + j]; }

static inline gsl_vector* gsl_vector_calloc(size_t n) {
    gsl_vector* v = (gsl_vector*)malloc(sizeof(gsl_vector));
    v->size = n;
    v->data = (double*)calloc(n ? n : 1, sizeof(double));
    return v;
}
static inline void gsl_vector_free(gsl_vector* v) {
    if (v) {
        free(v->data);
        free(v);
    }
}
static inline double gsl_vector_get(const gsl_vector* v, size_t i) { return v->data[i]; }

static inline gsl_permutation* gsl_permutation_alloc(size_t n) {
    gsl_permutation* p = (gsl_permutation*)malloc(sizeof(gsl_permutation));
    p->size = n;
    p->data = (int*)calloc(n ? n : 1, sizeof(int));
    return p;
}

/* The row-major gsl_matrix handed to a column-major LAPACK routine is the transpose; det(A^T) =
 * det(A) and the LD matrices are symmetric, so the factorisation is of the same matrix. */
static inline int gsl_linalg_LU_decomp(gsl_matrix* A, gsl_permutation* p, int* signum) {
    int n = (int)A->size1, lda = (int)A->tda, info = 0;
    dgetrf_(&n, &n, A->data, &lda, p->data, &info);
    int s = 1;
    for (int i = 0; i < n; i++)
        if (p->data[i] != i + 1) s = -s;
    *signum = s;
    return 0;
}
static inline double gsl_linalg_LU_det(gsl_matrix* LU, int signum) {
    double det = (double)signum;
    for (size_t i = 0; i < LU->size1; i++) det *= LU->data[i * LU->tda + i];
    return det;
}

static inline gsl_eigen_symmv_workspace* gsl_eigen_symmv_alloc(size_t n) {
    gsl_eigen_symmv_workspace* w = (gsl_eigen_symmv_workspace*)malloc(sizeof(gsl_eigen_symmv_workspace));
    w->size = n;
    return w;
}
static inline void gsl_eigen_symmv_free(gsl_eigen_symmv_workspace* w) { free(w); }

/* GSL reads the lower triangle of the row-major A; seen column-major that is the upper triangle. */
static inline int gsl_eigen_symmv(gsl_matrix* A, gsl_vector* eval, gsl_matrix* evec, gsl_eigen_symmv_workspace*) {
    int n = (int)A->size1, lda = (int)A->tda, info = 0, lwork = -1;
    double wq = 0;
    dsyev_("V", "U", &n, A->data, &lda, eval->data, &wq, &lwork, &info);
    lwork = (int)wq;
    double* work = (double*)malloc(sizeof(double) * (size_t)(lwork > 1 ? lwork : 1));
    dsyev_("V", "U", &n, A->data, &lda, eval->data, work, &lwork, &info);
    free(work);
    /* column-major Z(i,j) = A->data[j*lda+i] = component i of eigenvector j -> evec(i,j) */
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) evec->data[(size_t)i * evec->tda + j] = A->data[(size_t)j * lda + i];
    return info;
}

#endif
