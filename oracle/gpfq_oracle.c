/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the GPFQ hot path (the oracle's fast leg).
 *
 * Follows elybrand/quantized_neural_networks scripts/quantized_network.py:
 *   bit_round        :40-57   nearest alphabet element, first minimal index on ties
 *   one greedy step  :59-89   dead-direction guard, perpendicular guard, Q(<Xq,u+wX>/||Xq||^2)
 *   neuron walk      :91-121  u += w_t X_t - q_t Xq_t over t = 0..N0-1
 *   filter walk      :185-233 same walk over the C-order flattened (kh,kw) slice
 * with the dtype ladder of NumPy >= 2: w*X_t rounded to fp32 per element,
 * q*Xq_t and all sums in fp64, ||Xq_t|| rounded to fp32 (snrm2) then squared in fp64.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library;
 * the product never links it.  It differs from the NumPy restatement only in the summation
 * order of the two dot products (BLAS ddot vs. a plain loop), i.e. at the 1e-16 level;
 * tests/test_oracle.py pins it to the NumPy restatement.
 *
 * Build: gcc -O2 -pthread -fPIC -shared -ffp-contract=off -o libgpfq_oracle.so gpfq_oracle.c -lm
 * (-ffp-contract=off: the reference's update is mul, sub, add -- never fused.)
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define DEAD_NORM 1e-16
#define PERP_DOT 1e-10

static double bit_round(double t, const double *alphabet, int K) {
    int best = 0;
    double bd = fabs(alphabet[0] - t);
    for (int k = 1; k < K; ++k) {
        double d = fabs(alphabet[k] - t);
        if (d < bd) { bd = d; best = k; }
    }
    return alphabet[best];
}

/* fp32-rounded Euclidean norm of one row (what scipy.linalg.norm -> snrm2 returns) */
static double row_norm(const float *x, long m) {
    double s = 0.0;
    for (long i = 0; i < m; ++i) s += (double)x[i] * (double)x[i];
    return (double)(float)sqrt(s);
}

/* One neuron.  w has stride ldw (column of a row-major (N0,N1) matrix), q has stride ldq. */
static void walk(const float *X, const float *Xq, long N0, long m, const float *w, long ldw,
                 const double *alphabet, int K, const double *norms, double *q, long ldq,
                 double *u) {
    memset(u, 0, sizeof(double) * (size_t)m);
    for (long t = 0; t < N0; ++t) {
        const float *x = X + t * m, *xq = Xq + t * m;
        const float wt = w[t * ldw];
        double qt;
        if (norms[t] < DEAD_NORM) {
            qt = 0.0;
        } else {
            double d = 0.0;
            for (long i = 0; i < m; ++i) d += (double)xq[i] * u[i];
            if (fabs(d) < PERP_DOT) {
                qt = bit_round((double)wt, alphabet, K);
            } else {
                double s = 0.0;
                for (long i = 0; i < m; ++i) {
                    float wx = wt * x[i];            /* fp32 product */
                    s += (double)xq[i] * (u[i] + (double)wx);
                }
                qt = bit_round(s / (norms[t] * norms[t]), alphabet, K);
            }
        }
        q[t * ldq] = qt;
        for (long i = 0; i < m; ++i) {
            float wx = wt * x[i];
            double qx = qt * (double)xq[i];
            double diff = (double)wx - qx;
            u[i] += diff;
        }
    }
}

/* Dense layer fan-out (:549-562), neurons j0..j1-1, one worker thread per core (the reference
 * uses one pool process per core).  Also serves one conv channel: pass the (kh*kw, n_patches)
 * patch matrices, W + c*F with ldw = C*F, N1 = F (:699-718). */
typedef struct {
    const float *X, *Xq, *W;
    long N0, m, ldw, j1, ldq;
    const double *alphabet, *norms;
    int K;
    double *Q;
    long *next;
    int *fail;
} job_t;

static void *worker(void *arg) {
    job_t *jb = (job_t *)arg;
    double *u = (double *)malloc(sizeof(double) * (size_t)(jb->m > 0 ? jb->m : 1));
    if (!u) { __atomic_store_n(jb->fail, 1, __ATOMIC_RELAXED); return NULL; }
    for (;;) {
        long j = __atomic_fetch_add(jb->next, 1, __ATOMIC_RELAXED);
        if (j >= jb->j1) break;
        walk(jb->X, jb->Xq, jb->N0, jb->m, jb->W + j, jb->ldw, jb->alphabet, jb->K, jb->norms,
             jb->Q + j, jb->ldq, u);
    }
    free(u);
    return NULL;
}

int gpfq_oracle_layer(const float *X, const float *Xq, long N0, long m, const float *W, long ldw,
                      long j0, long j1, const double *alphabet, int K, double *Q, long ldq,
                      int nthreads) {
    if (N0 <= 0 || m < 0 || K <= 0 || j1 < j0) return 1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    double *norms = (double *)malloc(sizeof(double) * (size_t)N0);
    if (!norms) return 2;
    for (long t = 0; t < N0; ++t) norms[t] = row_norm(Xq + t * m, m);
    long next = j0;
    int fail = 0;
    job_t jb = {X, Xq, W, N0, m, ldw, j1, ldq, alphabet, norms, K, Q, &next, &fail};
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nthreads - 1; ++i)
        if (pthread_create(&th[started], NULL, worker, &jb) == 0) ++started;
    worker(&jb);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    free(norms);
    return fail ? 2 : 0;
}
