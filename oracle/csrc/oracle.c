/* ORACLE (test infrastructure only; never linked into or called from the product library).
 *
 * Plain-C restatement of the CPU algorithms on libParanumal's elliptic hot path, used by
 * tests/ as the checker and by bench.py's cpu_baseline / --impl reference legs.
 * Each function cites the reference file:line it follows (paths relative to the
 * libParanumal 0.5.0 tree).
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC oracle.c -o liboracle.so  (oracle/build.py)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int dlong;
typedef long long hlong;
typedef double dfloat;

/* ------------------------------------------------------------------------------------------
 * std::sort as implemented by libstdc++ (bits/stl_algo.h: __introsort_loop, threshold 16,
 * median-of-three to first, unguarded partition, final insertion sort; heap-sort fallback at
 * depth 2*floor(log2 n)).  The reference's shared-node discovery sorts with an UNSTABLE
 * std::sort on |baseId| and then picks the rand()%size-th node of each tie group
 * (libs/ogs/ogsSetup.cpp:245-275), so the tie order of this exact algorithm decides which
 * copy of a node becomes the positive ("owner") one.  The algorithm depends only on comparator
 * outcomes, so sorting a permutation array with the same strict-weak order reproduces it.
 * Third-party dependency restated: GNU libstdc++ 13 (GCC 13.3), not vendored in the reference.
 * ------------------------------------------------------------------------------------------ */
typedef struct { const int64_t* k1; const int64_t* k2; } keys_t;

static inline int lessp(const keys_t* K, int64_t a, int64_t b) {
  if (K->k1[a] < K->k1[b]) return 1;
  if (K->k1[a] > K->k1[b]) return 0;
  if (K->k2) return K->k2[a] < K->k2[b];
  return 0;
}
#define SWAP(p, q) do { int64_t _t = *(p); *(p) = *(q); *(q) = _t; } while (0)

static void push_heap_(int64_t* first, int64_t hole, int64_t top, int64_t value, const keys_t* K) {
  int64_t parent = (hole - 1) / 2;
  while (hole > top && lessp(K, first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}
static void adjust_heap_(int64_t* first, int64_t hole, int64_t len, int64_t value, const keys_t* K) {
  const int64_t top = hole;
  int64_t child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (lessp(K, first[child], first[child - 1])) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  push_heap_(first, hole, top, value, K);
}
static void heap_sort_(int64_t* first, int64_t* last, const keys_t* K) {
  int64_t len = last - first;
  if (len >= 2) {
    int64_t parent = (len - 2) / 2;
    for (;;) {
      int64_t v = first[parent];
      adjust_heap_(first, parent, len, v, K);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {
    --last;
    int64_t v = *last;
    *last = *first;
    adjust_heap_(first, 0, last - first, v, K);
  }
}
static void move_median_to_first_(int64_t* result, int64_t* a, int64_t* b, int64_t* c, const keys_t* K) {
  if (lessp(K, *a, *b)) {
    if (lessp(K, *b, *c)) SWAP(result, b);
    else if (lessp(K, *a, *c)) SWAP(result, c);
    else SWAP(result, a);
  } else if (lessp(K, *a, *c)) SWAP(result, a);
  else if (lessp(K, *b, *c)) SWAP(result, c);
  else SWAP(result, b);
}
static int64_t* unguarded_partition_(int64_t* first, int64_t* last, int64_t* pivot, const keys_t* K) {
  for (;;) {
    while (lessp(K, *first, *pivot)) ++first;
    --last;
    while (lessp(K, *pivot, *last)) --last;
    if (!(first < last)) return first;
    SWAP(first, last);
    ++first;
  }
}
static void introsort_loop_(int64_t* first, int64_t* last, long depth, const keys_t* K) {
  while (last - first > 16) {
    if (depth == 0) { heap_sort_(first, last, K); return; }
    --depth;
    int64_t* mid = first + (last - first) / 2;
    move_median_to_first_(first, first + 1, mid, last - 1, K);
    int64_t* cut = unguarded_partition_(first + 1, last, first, K);
    introsort_loop_(cut, last, depth, K);
    last = cut;
  }
}
static void unguarded_linear_insert_(int64_t* last, const keys_t* K) {
  int64_t val = *last;
  int64_t* next = last - 1;
  while (lessp(K, val, *next)) { *last = *next; last = next; --next; }
  *last = val;
}
static void insertion_sort_(int64_t* first, int64_t* last, const keys_t* K) {
  if (first == last) return;
  for (int64_t* i = first + 1; i != last; ++i) {
    if (lessp(K, *i, *first)) {
      int64_t val = *i;
      memmove(first + 1, first, (size_t)(i - first) * sizeof(int64_t));
      *first = val;
    } else unguarded_linear_insert_(i, K);
  }
}
/* perm[] must be initialised (normally 0..n-1 = the order the records arrive in). */
void oracle_libstdcxx_sort_perm(int64_t n, const int64_t* k1, const int64_t* k2, int64_t* perm) {
  keys_t K = {k1, k2};
  if (n <= 1) return;
  long lg = 0;
  for (int64_t m = n; m > 1; m >>= 1) lg++;
  introsort_loop_(perm, perm + n, 2 * lg, &K);
  if (n > 16) {
    insertion_sort_(perm, perm + 16, &K);
    for (int64_t* i = perm + 16; i != perm + n; ++i) unguarded_linear_insert_(i, &K);
  } else insertion_sort_(perm, perm + n, &K);
}

/* ------------------------------------------------------------------------------------------
 * ellipticPartialAxHex3D  (solvers/elliptic/okl/ellipticAxHex3D.okl:158-295), same operation
 * order as the OKL kernel: k-layer sweep, qr/qs/qt, geometric factors, transposed D applies.
 * GlobalToLocal == NULL gives the element-local twin ellipticAxHex3D (:28-152).
 * ------------------------------------------------------------------------------------------ */
#define MAXNQ 16
void oracle_ax_hex3d(int Nq, dlong Nelements, const dlong* elementList, const dlong* GlobalToLocal,
                     const dfloat* wJ, const dfloat* ggeo, const dfloat* D, dfloat lambda,
                     const dfloat* q, dfloat* Aq) {
  const int Nq2 = Nq * Nq, Np = Nq * Nq * Nq;
#pragma omp parallel
  {
  dfloat* u = (dfloat*)malloc(sizeof(dfloat) * Np);
  dfloat* Au = (dfloat*)malloc(sizeof(dfloat) * Np);
  dfloat s_Gqr[MAXNQ][MAXNQ], s_Gqs[MAXNQ][MAXNQ], r_Gqt[MAXNQ][MAXNQ], r_Auk[MAXNQ][MAXNQ];
#pragma omp for schedule(static)
  for (dlong ei = 0; ei < Nelements; ++ei) {
    const dlong e = elementList ? elementList[ei] : ei;
    for (int n = 0; n < Np; ++n) {
      if (GlobalToLocal) {
        const dlong id = GlobalToLocal[(size_t)e * Np + n];
        u[n] = (id != -1) ? q[id] : 0.0;
      } else u[n] = q[(size_t)e * Np + n];
      Au[n] = 0.0;
    }
    for (int k = 0; k < Nq; ++k) {
      for (int j = 0; j < Nq; ++j)
        for (int i = 0; i < Nq; ++i) {
          const size_t gbase = (size_t)e * 6 * Np + k * Nq2 + j * Nq + i;
          const dfloat G00 = ggeo[gbase + 0 * Np], G01 = ggeo[gbase + 1 * Np], G02 = ggeo[gbase + 2 * Np];
          const dfloat G11 = ggeo[gbase + 3 * Np], G12 = ggeo[gbase + 4 * Np], G22 = ggeo[gbase + 5 * Np];
          const dfloat GwJ = wJ[(size_t)e * Np + k * Nq2 + j * Nq + i];
          dfloat qt = 0, qr = 0, qs = 0;
          for (int m = 0; m < Nq; ++m) qt += D[k * Nq + m] * u[m * Nq2 + j * Nq + i];
          for (int m = 0; m < Nq; ++m) {
            qr += D[i * Nq + m] * u[k * Nq2 + j * Nq + m];
            qs += D[j * Nq + m] * u[k * Nq2 + m * Nq + i];
          }
          s_Gqs[j][i] = (G01 * qr + G11 * qs + G12 * qt);
          s_Gqr[j][i] = (G00 * qr + G01 * qs + G02 * qt);
          r_Gqt[j][i] = (G02 * qr + G12 * qs + G22 * qt);
          r_Auk[j][i] = GwJ * lambda * u[k * Nq2 + j * Nq + i];
        }
      for (int j = 0; j < Nq; ++j)
        for (int i = 0; i < Nq; ++i) {
          dfloat Auk = r_Auk[j][i];
          for (int m = 0; m < Nq; ++m) {
            Auk += D[m * Nq + j] * s_Gqs[m][i];
            Au[m * Nq2 + j * Nq + i] += D[k * Nq + m] * r_Gqt[j][i];
            Auk += D[m * Nq + i] * s_Gqr[j][m];
          }
          Au[k * Nq2 + j * Nq + i] += Auk;
        }
    }
    for (int n = 0; n < Np; ++n) Aq[(size_t)e * Np + n] = Au[n];
  }
  free(u); free(Au);
  }
}

/* ogsOperator_t::Gather host template, Add  (libs/ogs/ogsOperator.cpp:64-113): sequential
 * left-to-right sum per row starting from 0. */
void oracle_gather_add(dlong Nrows, const dlong* rowStarts, const dlong* colIds, const dfloat* v, dfloat* gv) {
#pragma omp parallel for schedule(static)
  for (dlong n = 0; n < Nrows; ++n) {
    dfloat val = 0.0;
    for (dlong g = rowStarts[n]; g < rowStarts[n + 1]; ++g) val += v[colIds[g]];
    gv[n] = val;
  }
}
/* ogsOperator_t::Scatter host (libs/ogs/ogsOperator.cpp:219-262). */
void oracle_scatter(dlong Nrows, const dlong* rowStarts, const dlong* colIds, const dfloat* gv, dfloat* v) {
#pragma omp parallel for schedule(static)
  for (dlong n = 0; n < Nrows; ++n)
    for (dlong g = rowStarts[n]; g < rowStarts[n + 1]; ++g) v[colIds[g]] = gv[n];
}

/* Single-rank elliptic_t::Operator C0 branch (solvers/elliptic/src/ellipticOperator.cpp:31-106):
 * Aq = gather_T( Ax_L( scatter(q) ) ) with the scatter folded into the GlobalToLocal read. */
void oracle_operator(int Nq, dlong Nelements, const dlong* GlobalToLocal, const dfloat* wJ, const dfloat* ggeo,
                     const dfloat* D, dfloat lambda, dlong Nrows, const dlong* rowStartsT, const dlong* colIdsT,
                     const dfloat* q, dfloat* AqL, dfloat* Aq) {
  oracle_ax_hex3d(Nq, Nelements, NULL, GlobalToLocal, wJ, ggeo, D, lambda, q, AqL);
  oracle_gather_add(Nrows, rowStartsT, colIdsT, AqL, Aq);
}

static dfloat dot_(dlong N, const dfloat* a, const dfloat* b) {
  dfloat s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (dlong n = 0; n < N; ++n) s += a[n] * b[n];
  return s;
}

/* LinearSolver::pcg::Solve (libs/linearSolver/linearSolverPCG.cpp:67-151) with the Jacobi
 * preconditioner (solvers/elliptic/src/ellipticPreconJacobi.cpp:42-51) or none (invDiag==NULL),
 * single rank, ABS/REL-INITRESID stopping.  x and r are overwritten like the reference.
 * resHist (optional, length maxit+1) receives sqrt(rdotr) initial + after every iteration. */
int oracle_pcg(int Nq, dlong Nelements, const dlong* GlobalToLocal, const dfloat* wJ, const dfloat* ggeo,
               const dfloat* D, dfloat lambda, dlong N, const dlong* rowStartsT, const dlong* colIdsT,
               const dfloat* invDiag, dfloat* x, dfloat* r, dfloat tol, int maxit, int flexible, dfloat* resHist) {
  const int Np = Nq * Nq * Nq;
  dfloat* AqL = (dfloat*)malloc(sizeof(dfloat) * (size_t)Nelements * Np);
  dfloat* p = (dfloat*)calloc((size_t)N, sizeof(dfloat));
  dfloat* z = (dfloat*)calloc((size_t)N, sizeof(dfloat));
  dfloat* Ap = (dfloat*)calloc((size_t)N, sizeof(dfloat));
  dfloat rdotz1 = 0, rdotz2 = 0, alpha = 0, beta = 0, pAp = 0, rdotr0 = 0, TOL = 0;

  oracle_operator(Nq, Nelements, GlobalToLocal, wJ, ggeo, D, lambda, N, rowStartsT, colIdsT, x, AqL, Ap);
#pragma omp parallel for schedule(static)
  for (dlong n = 0; n < N; ++n) r[n] = -1.0 * Ap[n] + 1.0 * r[n];
  rdotr0 = sqrt(dot_(N, r, r));
  rdotr0 = rdotr0 * rdotr0;
  TOL = fmax(tol * tol * rdotr0, tol * tol);
  if (resHist) resHist[0] = sqrt(rdotr0);
  memset(Ap, 0, sizeof(dfloat) * (size_t)N);

  int iter;
  for (iter = 0; iter < maxit; ++iter) {
    if (((iter == 0) && (rdotr0 == 0.0)) || ((iter > 0) && (rdotr0 <= TOL))) break;
    if (invDiag) {
#pragma omp parallel for schedule(static)
      for (dlong n = 0; n < N; ++n) z[n] = 1.0 * invDiag[n] * r[n];
    } else memcpy(z, r, sizeof(dfloat) * (size_t)N);
    rdotz2 = rdotz1;
    rdotz1 = dot_(N, r, z);
    if (flexible) {
      dfloat zdotAp = dot_(N, z, Ap);
      beta = (iter == 0) ? 0.0 : -alpha * zdotAp / rdotz2;
    } else beta = (iter == 0) ? 0.0 : rdotz1 / rdotz2;
#pragma omp parallel for schedule(static)
    for (dlong n = 0; n < N; ++n) p[n] = 1.0 * z[n] + beta * p[n];
    oracle_operator(Nq, Nelements, GlobalToLocal, wJ, ggeo, D, lambda, N, rowStartsT, colIdsT, p, AqL, Ap);
    pAp = dot_(N, p, Ap);
    alpha = rdotz1 / pAp;
    dfloat rr = 0.0;
#pragma omp parallel for reduction(+ : rr) schedule(static)
    for (dlong n = 0; n < N; ++n) {
      dfloat rn = r[n];
      x[n] += alpha * p[n];
      rn -= alpha * Ap[n];
      rr += rn * rn;
      r[n] = rn;
    }
    rdotr0 = rr;
    if (resHist) resHist[iter + 1] = sqrt(rdotr0);
  }
  free(AqL); free(p); free(z); free(Ap);
  return iter;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
