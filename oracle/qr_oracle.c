/*
 * qr_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reverse-reconciliation hot path of
 * moriglia/qam-reconciliation, written from the algorithm description so the
 * CUDA path can be checked against it at sizes where the compiled reference
 * (oracle/_ref, built by oracle/build_ref.py) is too slow (its Decoder
 * constructor is quadratic, its demapper calls scipy per erf).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  Nothing under
 * qam-reconciliation_b200/ links, imports or executes it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py runs this file
 * against the compiled reference in this container and against the committed
 * fixtures in tests/golden/ (generated from the compiled reference by
 * tests/golden/make_golden.py): integer outputs and the whole decoder are
 * bit-identical; the erf-based mapper agrees to <=1e-12 (the reference calls
 * scipy.special.erf, this file calls libm erf -- they differ in the last ulp).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference checkout).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define QRO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Gray labelling -- qamreconciliation/bicm.pyx:26-41.                        */
/* Reflected Gray code built recursively there; closed form here: column k   */
/* of symbol i is 1 iff (i>>k) mod 4 is 1 or 2 (the same rule the reference  */
/* itself uses at noisemapper.pyx:208-215 and :521-530).                      */
QRO_API void qro_gray_table(int bps, uint8_t *s_to_b)
{
    int order = 1 << bps;
    for (int i = 0; i < order; ++i)
        for (int k = 0; k < bps; ++k) {
            int m = (i >> k) & 3;
            s_to_b[i * bps + k] = (uint8_t)(m == 1 || m == 2);
        }
}

/* PAM alphabet -- qamreconciliation/alphabet.pyx:35-76.                     */
/* constellation[i] = (i-(M-1)/2)*step (:62), variance = sum p_i |a_i|^2     */
/* (:66-67), thresholds: interior a_i-step/2 (:69-71), ends 100*a_0 and      */
/* 100*a_{M-1} (:72-73).  probs==NULL means uniform (:46-47).                */
QRO_API int qro_alphabet(int bps, double step, const double *probs,
                         double *constellation, double *thresholds,
                         double *probs_out, double *variance)
{
    if (bps <= 0) return -1;
    int order = 1 << bps;
    for (int i = 0; i < order; ++i)
        probs_out[i] = probs ? probs[i] : 1.0 / order;
    for (int i = 0; i < order; ++i)
        constellation[i] = (i - (order - 1) / 2.0) * step;
    double var = 0.0;
    for (int i = 0; i < order; ++i) {
        double a = fabs(constellation[i]);
        var += probs_out[i] * (a * a);
    }
    *variance = var;
    for (int i = 1; i < order; ++i)
        thresholds[i] = constellation[i] - step / 2;
    thresholds[0] = constellation[0] * 100;
    thresholds[order] = constellation[order - 1] * 100;
    return 0;
}

/* Decision-region search -- qamreconciliation/noisemapper.pyx:27-44.        */
/* The reference recurses on array slices; this is the same recursion on an  */
/* (offset,length) window so NaN and out-of-range inputs fall through the    */
/* comparisons in the same order.                                            */
static long region_search(const double *dom, long len, double val)
{
    long base = 0;
    for (;;) {
        if (len == 1) return base;
        if (val < dom[0]) return base;
        if (val > dom[len - 1]) return base + len - 1;
        long mid = len / 2 - 1;
        if (val < dom[mid]) {            /* left window dom[0:mid] */
            len = mid;
            continue;
        }
        if (val >= dom[mid + 1]) {       /* right window dom[mid+1:] */
            base += mid + 1;
            dom += mid + 1;
            len -= mid + 1;
            continue;
        }
        return base + mid;
    }
}

/* NoiseMapper.hard_decide_index -- noisemapper.pyx:349-359.                 */
QRO_API void qro_hard_decide_index(const double *thresholds, int order,
                                   const double *y, long n, long *idx)
{
    for (long j = 0; j < n; ++j) {
        long r = region_search(thresholds, order + 1, y[j]);
        if (r == order) r = order - 1;
        idx[j] = r;
    }
}

/* PAMAlphabet.demap_symbols_to_bits -- alphabet.pyx:98-107.                 */
QRO_API void qro_symbols_to_bits(const uint8_t *s_to_b, int bps,
                                 const long *idx, long n, uint8_t *bits)
{
    for (long i = 0; i < n; ++i)
        memcpy(bits + i * bps, s_to_b + idx[i] * bps, (size_t)bps);
}

/* Matrix.eval_syndrome -- qamreconciliation/matrix.pyx:55-60.               */
QRO_API void qro_eval_syndrome(const long *vid, const long *cid, long E,
                               const uint8_t *word, long C, uint8_t *synd)
{
    memset(synd, 0, (size_t)C);
    for (long e = 0; e < E; ++e)
        synd[cid[e]] ^= word[vid[e]];
}

/* utils.count_errors_from_lappr -- qamreconciliation/utils.pyx:27-40.       */
QRO_API long qro_count_errors(const double *lappr, const uint8_t *word, long n)
{
    long count = 0;
    for (long i = 0; i < n; ++i)
        count += (lappr[i] >= 0) ? word[i] : 1 - word[i];
    return count;
}

/* ------------------------------------------------------------------------- */
/* Noise mapper                                                              */
typedef struct {
    int order, bps;
    double noise_var, sigma;
    double *constellation;     /* [order]   */
    double *thresholds;        /* [order+1] */
    double *probabilities;     /* [order]   */
    uint8_t *sign_config;      /* [order]   */
    double *F_Y_thresholds;    /* [order+1] */
    double *delta_F_Y;         /* [order]   */
    double *fwrd;              /* [order][order]  P(xhat=a_i | x=a_j) at [j][i] */
    double *back;              /* [order][order] */
    double *bare_llr_table;    /* [order][bps]   */
    double *inf_erf_table;     /* [order][order] */
} qro_mapper;

/* __F_Z -- noisemapper.pyx:66-67 (the reference evaluates erf via scipy).   */
static double gauss_cdf(double z, double mu, double sigma)
{
    return 0.5 * (1 + erf((z - mu) / (sqrt(2) * sigma)));
}

/* _single_F_Y -- noisemapper.pyx:278-286: mixture CDF, summed k=0 upward.   */
static double mixture_cdf(const qro_mapper *m, double y)
{
    double res = gauss_cdf(y, m->constellation[0], m->sigma) * m->probabilities[0];
    for (int i = 1; i < m->order; ++i)
        res += gauss_cdf(y, m->constellation[i], m->sigma) * m->probabilities[i];
    return res;
}

QRO_API void qro_mapper_destroy(qro_mapper *m)
{
    if (!m) return;
    free(m->constellation); free(m->thresholds); free(m->probabilities);
    free(m->sign_config); free(m->F_Y_thresholds); free(m->delta_F_Y);
    free(m->fwrd); free(m->back); free(m->bare_llr_table); free(m->inf_erf_table);
    free(m);
}

/* NoiseMapper.__cinit__ tables -- noisemapper.pyx:103-236 (without the      */
/* dense _F_Y grid of :135-144, which only g_inv reads).                     */
QRO_API qro_mapper *qro_mapper_create(int bps, const double *constellation,
                                      const double *thresholds, const double *probabilities,
                                      double noise_var, const uint8_t *sign_config)
{
    if (noise_var <= 0 || bps <= 0) return NULL;
    int M = 1 << bps;
    qro_mapper *m = (qro_mapper *)calloc(1, sizeof(*m));
    m->order = M; m->bps = bps; m->noise_var = noise_var; m->sigma = sqrt(noise_var);
    m->constellation = (double *)malloc(sizeof(double) * M);
    m->thresholds = (double *)malloc(sizeof(double) * (M + 1));
    m->probabilities = (double *)malloc(sizeof(double) * M);
    m->sign_config = (uint8_t *)calloc(M, 1);
    m->F_Y_thresholds = (double *)malloc(sizeof(double) * (M + 1));
    m->delta_F_Y = (double *)malloc(sizeof(double) * M);
    m->fwrd = (double *)malloc(sizeof(double) * M * M);
    m->back = (double *)malloc(sizeof(double) * M * M);
    m->bare_llr_table = (double *)malloc(sizeof(double) * M * bps);
    m->inf_erf_table = (double *)malloc(sizeof(double) * M * M);
    memcpy(m->constellation, constellation, sizeof(double) * M);
    memcpy(m->thresholds, thresholds, sizeof(double) * (M + 1));
    memcpy(m->probabilities, probabilities, sizeof(double) * M);
    if (sign_config) memcpy(m->sign_config, sign_config, M);

    /* :149-153 and :156-162 */
    m->F_Y_thresholds[0] = 0;
    m->F_Y_thresholds[M] = 1;
    for (int i = 1; i < M; ++i) m->F_Y_thresholds[i] = mixture_cdf(m, thresholds[i]);
    for (int i = 0; i < M; ++i) m->delta_F_Y[i] = m->F_Y_thresholds[i + 1] - m->F_Y_thresholds[i];

    /* forward transition probabilities :167-182 (libc erf in the reference) */
    double s2 = sqrt(2) * m->sigma;
    for (int j = 0; j < M; ++j) {
        m->fwrd[j * M + 0] = 0.5 * (erf((thresholds[1] - constellation[j]) / s2) + 1);
        m->fwrd[j * M + M - 1] = 0.5 * (1 - erf((thresholds[M - 1] - constellation[j]) / s2));
        for (int i = 1; i < M - 1; ++i)
            m->fwrd[j * M + i] = 0.5 * (erf((thresholds[i + 1] - constellation[j]) / s2) -
                                        erf((thresholds[i] - constellation[j]) / s2));
    }
    /* backward transition probabilities :185-194 */
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) {
            double tot = 0;
            for (int k = 0; k < M; ++k) tot += probabilities[k] * m->fwrd[k * M + i];
            m->back[i * M + j] = probabilities[j] * m->fwrd[j * M + i] / tot;
        }
    /* bare LLR table :198-220 */
    for (int j = 0; j < M; ++j)
        for (int k = 0; k < bps; ++k) {
            double num = 0, den = 0;
            for (int i = 0; i < M; ++i) {
                int q = i >> k;
                if ((q * (q + 1)) & 3) den += m->fwrd[j * M + i];
                else num += m->fwrd[j * M + i];
            }
            m->bare_llr_table[j * bps + k] = (den == 0) ? 1e300 : log(num / den);
        }
    /* inf_erf_table :223-235 */
    for (int j = 0; j < M; ++j) {
        m->inf_erf_table[0 * M + j] = -1;
        for (int i = 1; i < M; ++i)
            m->inf_erf_table[i * M + j] = erf((thresholds[i] - constellation[j]) / s2);
    }
    return m;
}

QRO_API void qro_mapper_tables(const qro_mapper *m, double *F_Y_thresholds, double *delta_F_Y,
                               double *fwrd, double *back, double *bare, double *inf_erf)
{
    int M = m->order;
    memcpy(F_Y_thresholds, m->F_Y_thresholds, sizeof(double) * (M + 1));
    memcpy(delta_F_Y, m->delta_F_Y, sizeof(double) * M);
    memcpy(fwrd, m->fwrd, sizeof(double) * M * M);
    memcpy(back, m->back, sizeof(double) * M * M);
    memcpy(bare, m->bare_llr_table, sizeof(double) * M * m->bps);
    memcpy(inf_erf, m->inf_erf_table, sizeof(double) * M * M);
}

/* NoiseMapper.g + map_noise -- noisemapper.pyx:289-292, :373-388.           */
QRO_API void qro_map_noise(const qro_mapper *m, const double *y, const long *idx,
                           long n, double *out)
{
    for (long j = 0; j < n; ++j) {
        long i = idx[j];
        double F = mixture_cdf(m, y[j]);
        out[j] = m->sign_config[i] ? (m->F_Y_thresholds[i + 1] - F) / m->delta_F_Y[i]
                                   : (F - m->F_Y_thresholds[i]) / m->delta_F_Y[i];
    }
}

/* NoiseMapper.g_inv_search -- noisemapper.pyx:310-345: bracket by doubling  */
/* from +-1, bisect until the bracket is <= accuracy wide, return midpoint.  */
QRO_API double qro_g_inv_search(const qro_mapper *m, double n_hat, int i, double accuracy)
{
    double target = m->sign_config[i]
                        ? m->F_Y_thresholds[i + 1] - n_hat * m->delta_F_Y[i]
                        : n_hat * m->delta_F_Y[i] + m->F_Y_thresholds[i];
    double lo, hi;
    if (target > .5) {
        hi = 1; lo = 0;
        while (mixture_cdf(m, hi) < target) { lo = hi; hi *= 2.; }
    } else {
        lo = -1; hi = 0;
        while (mixture_cdf(m, lo) > target) { hi = lo; lo *= 2.; }
    }
    while ((hi - lo) > accuracy) {
        double mid = (hi + lo) / 2;
        if (mixture_cdf(m, mid) > target) hi = mid;
        else lo = mid;
    }
    return (hi + lo) / 2;
}

/* NoiseMapper.demap_lappr -- noisemapper.pyx:450-540.  NOTE the reference   */
/* divides the exponent by 2*sigma^2 only for k>j (:511-515) and not for k<j */
/* (:503-507); restated as is (corrected!=0 divides in both, as the sibling  */
/* formulas at :666-675 do).                                                  */
static void demap_one(const qro_mapper *m, double n, long j, int corrected, double *lappr)
{
    double N[16], D[16];
    const double two_s2 = 2 * m->noise_var;
    const double *a = m->constellation;
    for (int k = 0; k < m->bps; ++k) { N[k] = 0; D[k] = 0; }
    for (int i = 0; i < m->order; ++i) {
        double yh = qro_g_inv_search(m, n, i, 1e-9);
        double s = 0;
        for (long k = 0; k < j; ++k) {
            double ex = (2 * yh - a[k] - a[j]) * (a[k] - a[j]);
            if (corrected) ex = ex / two_s2;
            s += exp(ex) * m->probabilities[k];
        }
        s += m->probabilities[j];
        for (long k = j + 1; k < m->order; ++k)
            s += exp((2 * yh - a[k] - a[j]) * (a[k] - a[j]) / two_s2) * m->probabilities[k];
        int q = i;
        for (int k = 0; k < m->bps; ++k) {
            if ((q * (q + 1)) & 3) D[k] += m->delta_F_Y[i] / s;
            else N[k] += m->delta_F_Y[i] / s;
            q >>= 1;
        }
    }
    for (int k = 0; k < m->bps; ++k) lappr[k] = log(N[k]) - log(D[k]);
}

/* NoiseMapper.demap_lappr_array -- noisemapper.pyx:544-559.                 */
QRO_API void qro_demap_lappr_array(const qro_mapper *m, const double *n, const long *j,
                                   long count, int corrected, double *lappr)
{
    for (long s = 0; s < count; ++s)
        demap_one(m, n[s], j[s], corrected, lappr + s * m->bps);
}

/* NoiseMapper.bare_llr -- noisemapper.pyx:423-432.                          */
QRO_API void qro_bare_llr(const qro_mapper *m, const long *symb, long count, double *llr)
{
    for (long s = 0; s < count; ++s)
        memcpy(llr + s * m->bps, m->bare_llr_table + symb[s] * m->bps, sizeof(double) * m->bps);
}

/* Direct-reconciliation LLR -- sims/reconciliation.pyx:25-51 (+ :54-72).    */
QRO_API void qro_direct_llr(const double *y, long count, const double *constellation,
                            int bps, double two_variance, double *lappr)
{
    int order = 1 << bps;
    for (long s = 0; s < count; ++s) {
        double N[16], D[16];
        for (int l = 0; l < bps; ++l) { N[l] = 0; D[l] = 0; }
        for (int i = 0; i < order; ++i) {
            /* the reference's `**2` compiles to a run-time libm pow(d, 2.0), which is not
             * always the correctly rounded d*d; keep the call (volatile defeats folding) */
            volatile double two = 2.0;
            double d = y[s] - constellation[i];
            double term = exp(-pow(d, two) / two_variance);
            int q = i;
            for (int l = 0; l < bps; ++l) {
                if ((q * (q + 1)) & 3) D[l] += term;
                else N[l] += term;
                q >>= 1;
            }
        }
        for (int l = 0; l < bps; ++l) lappr[s * bps + l] = log(N[l]) - log(D[l]);
    }
}

/* ------------------------------------------------------------------------- */
/* Syndrome sum-product decoder -- qamreconciliation/decoder.pyx             */
typedef struct {
    long E, C, N;
    long *c_ptr, *c_edge, *c_var;   /* per check: edges in ascending edge id, and their variables */
    long *v_ptr, *v_edge;           /* per variable: edges in ascending edge id */
    long max_cdeg;
    double *scratch;                /* 2*(max_cdeg-1) doubles, cf. decoder.pyx:131-135 */
} qro_decoder;

QRO_API void qro_decoder_destroy(qro_decoder *d)
{
    if (!d) return;
    free(d->c_ptr); free(d->c_edge); free(d->c_var); free(d->v_ptr); free(d->v_edge);
    free(d->scratch); free(d);
}

/* Decoder.__cinit__ -- decoder.pyx:93-146.  The reference scans the edge    */
/* list once per node (:60-89, quadratic); a counting sort yields the same   */
/* per-node lists (ascending edge id, :73-76) in O(E).                        */
QRO_API qro_decoder *qro_decoder_create(const long *vid, const long *cid, long E)
{
    qro_decoder *d = (qro_decoder *)calloc(1, sizeof(*d));
    long C = 0, N = 0;
    for (long e = 0; e < E; ++e) {
        if (cid[e] + 1 > C) C = cid[e] + 1;
        if (vid[e] + 1 > N) N = vid[e] + 1;
    }
    d->E = E; d->C = C; d->N = N;
    d->c_ptr = (long *)calloc(C + 1, sizeof(long));
    d->v_ptr = (long *)calloc(N + 1, sizeof(long));
    d->c_edge = (long *)malloc(sizeof(long) * E);
    d->c_var = (long *)malloc(sizeof(long) * E);
    d->v_edge = (long *)malloc(sizeof(long) * E);
    for (long e = 0; e < E; ++e) { d->c_ptr[cid[e] + 1]++; d->v_ptr[vid[e] + 1]++; }
    for (long c = 0; c < C; ++c) {
        if (d->c_ptr[c + 1] > d->max_cdeg) d->max_cdeg = d->c_ptr[c + 1];
        d->c_ptr[c + 1] += d->c_ptr[c];
    }
    for (long v = 0; v < N; ++v) d->v_ptr[v + 1] += d->v_ptr[v];
    long *cfill = (long *)calloc(C, sizeof(long));
    long *vfill = (long *)calloc(N, sizeof(long));
    for (long e = 0; e < E; ++e) {
        long c = cid[e], v = vid[e];
        d->c_edge[d->c_ptr[c] + cfill[c]] = e;
        d->c_var[d->c_ptr[c] + cfill[c]] = v;
        cfill[c]++;
        d->v_edge[d->v_ptr[v] + vfill[v]] = e;
        vfill[v]++;
    }
    free(cfill); free(vfill);
    d->scratch = (double *)malloc(sizeof(double) * 2 * (d->max_cdeg > 1 ? d->max_cdeg : 2));
    return d;
}

QRO_API void qro_decoder_info(const qro_decoder *d, long *N, long *C, long *E)
{
    *N = d->N; *C = d->C; *E = d->E;
}

/* __sgn and __box_plus -- decoder.pyx:37-45, evaluated left to right:       */
/* sgn(a)sgn(b)min(|a|,|b|) + log(1+exp(-|a+b|)) - log(1+exp(-|a-b|)).       */
static inline int sgn(double x) { return (0.0 < x) - (x < 0.0); }

static inline double box_plus(double a, double b)
{
    double fa = fabs(a), fb = fabs(b);
    double mn = (fb < fa) ? fb : fa;
    return sgn(a) * sgn(b) * mn + log(1 + exp(-fabs(a + b))) - log(1 + exp(-fabs(a - b)));
}

/* __process_check_node -- decoder.pyx:322-369: forward/backward box-plus    */
/* recursion over the check's edges in ascending edge id, syndrome bit as a  */
/* +-1 prefactor on every outgoing message.                                   */
static void check_node(const qro_decoder *d, long c, const uint8_t *synd,
                       double *c2v, const double *v2c)
{
    const long *ed = d->c_edge + d->c_ptr[c];
    long deg = d->c_ptr[c + 1] - d->c_ptr[c];
    double *F = d->scratch;
    double *B = F + deg - 2;           /* B[0] aliases F[deg-2]; B[0] is never used */
    F[0] = v2c[ed[0]];
    B[deg - 1] = v2c[ed[deg - 1]];
    for (long i = 1; i < deg - 1; ++i) F[i] = box_plus(F[i - 1], v2c[ed[i]]);
    for (long i = deg - 2; i > 0; --i) B[i] = box_plus(B[i + 1], v2c[ed[i]]);
    double pre = synd[c] ? -1.0 : 1.0;
    c2v[ed[0]] = pre * B[1];
    for (long i = 1; i < deg - 1; ++i) c2v[ed[i]] = pre * box_plus(F[i - 1], B[i + 1]);
    c2v[ed[deg - 1]] = pre * F[deg - 2];
}

/* __process_var_node -- decoder.pyx:285-298: posterior = channel LLR plus   */
/* the incoming messages added one by one in ascending edge id; outgoing =   */
/* posterior minus the message that came in on that edge.                     */
static void var_node(const qro_decoder *d, long v, const double *llr,
                     const double *c2v, double *v2c, double *post)
{
    const long *ed = d->v_edge + d->v_ptr[v];
    long deg = d->v_ptr[v + 1] - d->v_ptr[v];
    post[v] = llr[v];
    for (long i = 0; i < deg; ++i) post[v] += c2v[ed[i]];
    for (long i = 0; i < deg; ++i) v2c[ed[i]] = post[v] - c2v[ed[i]];
}

/* __check_lappr(_node) -- decoder.pyx:235-257: every check must see         */
/* synd XOR parity(#negative posteriors) == 0; strict < 0 (:244).            */
static int checks_satisfied(const qro_decoder *d, const double *post, const uint8_t *synd)
{
    for (long c = 0; c < d->C; ++c) {
        uint8_t parity = synd[c];
        for (long p = d->c_ptr[c]; p < d->c_ptr[c + 1]; ++p)
            if (post[d->c_var[p]] < 0) parity ^= 1;
        if ((parity ^ 1) == 0) return 0;
    }
    return 1;
}

QRO_API int qro_check_lappr(const qro_decoder *d, const double *post, const uint8_t *synd)
{
    return checks_satisfied(d, post, synd);
}

QRO_API void qro_process_check_node(const qro_decoder *d, long c, const uint8_t *synd,
                                    double *c2v, const double *v2c)
{
    check_node(d, c, synd, c2v, v2c);
}

QRO_API void qro_process_var_node(const qro_decoder *d, long v, const double *llr,
                                  const double *c2v, double *v2c, double *post)
{
    var_node(d, v, llr, c2v, v2c, post);
}

/* Decoder._decode -- decoder.pyx:391-436.  Returns success; *iters as the   */
/* reference reports it.  `post` receives the final posteriors.              */
QRO_API int qro_decode(const qro_decoder *d, const double *llr, const uint8_t *synd,
                       int max_iterations, double *post, int *iters)
{
    if (checks_satisfied(d, llr, synd)) {            /* :402-405 */
        memcpy(post, llr, sizeof(double) * d->N);
        *iters = 0;
        return 1;
    }
    double *c2v = (double *)calloc(d->E, sizeof(double));   /* :408 */
    double *v2c = (double *)malloc(sizeof(double) * d->E);
    for (long v = 0; v < d->N; ++v) var_node(d, v, llr, c2v, v2c, post);   /* :420-421 */
    int ok = 0;
    *iters = max_iterations;
    for (int it = 0; it < max_iterations; ++it) {                         /* :424-433 */
        for (long c = 0; c < d->C; ++c) check_node(d, c, synd, c2v, v2c);
        for (long v = 0; v < d->N; ++v) var_node(d, v, llr, c2v, v2c, post);
        if (checks_satisfied(d, post, synd)) { ok = 1; *iters = it + 1; break; }
    }
    free(c2v); free(v2c);
    return ok;
}

/* Convenience for the CPU baseline: decode `frames` frames one after the    */
/* other (llr [frames][N], synd [frames][C]).                                 */
QRO_API void qro_decode_frames(const qro_decoder *d, const double *llr, const uint8_t *synd,
                               long frames, int max_iterations, double *post,
                               uint8_t *success, int *iters)
{
    for (long f = 0; f < frames; ++f)
        success[f] = (uint8_t)qro_decode(d, llr + f * d->N, synd + f * d->C, max_iterations,
                                         post + f * d->N, iters + f);
}

/* ------------------------------------------------------------------------- */
/* Rest of the NoiseMapper surface                 */

/* F_Z -- noisemapper.pyx:70-80.                                              */
QRO_API void qro_F_Z(const double *z, long n, double mu, double sigma, double *out)
{
    for (long i = 0; i < n; ++i) out[i] = gauss_cdf(z[i], mu, sigma);
}

/* NoiseMapper.F_Y -- noisemapper.pyx:264-275: UNIFORM weights (sum / order), */
/* not the alphabet's probabilities; summed i = 0 upward, divided last.       */
QRO_API void qro_F_Y(const qro_mapper *m, const double *y, long n, double *out)
{
    for (long j = 0; j < n; ++j) {
        double res = gauss_cdf(y[j], m->constellation[0], m->sigma);
        for (int i = 1; i < m->order; ++i) res += gauss_cdf(y[j], m->constellation[i], m->sigma);
        out[j] = res / m->order;
    }
}

/* the dense grid of noisemapper.pyx:135-144: numpy.linspace(low, high, n) is  */
/* arange(n) * step + low with the last point set to `high`; F_Y on it.        */
QRO_API void qro_grid(const qro_mapper *m, double y_low, double y_high, long n, double *y, double *F)
{
    const double step = (y_high - y_low) / (double)(n - 1);
    for (long i = 0; i < n; ++i) {
        volatile double t = (double)i * step;   /* two roundings, as numpy does */
        y[i] = t + y_low;
    }
    if (n > 1) y[n - 1] = y_high;
    qro_F_Y(m, y, n, F);
}

/* __interp -- noisemapper.pyx:47-63.                                         */
QRO_API double qro_interp(const double *dom, const double *cod, long n, double val)
{
    if (val >= dom[n - 1]) return cod[n - 1];
    long index = region_search(dom, n, val);
    if (index == n - 1) return cod[index];
    if (dom[index + 1] == dom[index]) return cod[index];
    volatile double num = (cod[index + 1] - cod[index]) * (val - dom[index]);
    return cod[index] + num / (dom[index + 1] - dom[index]);
}

/* g_inv -- noisemapper.pyx:295-307 (and the subclasses' :786-797, :805-816 through `sign`:      */
/* sign[i] != 0 selects the decreasing branch).  demap_noise -- :391-404.                          */
QRO_API void qro_g_inv(const qro_mapper *m, const uint8_t *sign, const double *gridF, const double *gridY,
                       long npts, const double *n_hat, const long *idx, long n, double *out)
{
    for (long j = 0; j < n; ++j) {
        const long i = idx[j];
        volatile double prod = n_hat[j] * m->delta_F_Y[i];
        const double target = sign[i] ? m->F_Y_thresholds[i + 1] - prod : prod + m->F_Y_thresholds[i];
        out[j] = qro_interp(gridF, gridY, npts, target);
    }
}

/* g with an explicit sign vector (NoiseMapper.g :289-292, FlipSign :776-780, AntiFlipSign :799-802) */
QRO_API void qro_map_noise_sign(const qro_mapper *m, const uint8_t *sign, const double *y, const long *idx,
                                long n, double *out)
{
    for (long j = 0; j < n; ++j) {
        long i = idx[j];
        double F = mixture_cdf(m, y[j]);
        out[j] = sign[i] ? (m->F_Y_thresholds[i + 1] - F) / m->delta_F_Y[i]
                         : (F - m->F_Y_thresholds[i]) / m->delta_F_Y[i];
    }
}

/* demap_lappr_simplified(_array) -- noisemapper.pyx:563-621 ("formulation 1").                    */
QRO_API void qro_demap_simplified(const qro_mapper *m, const uint8_t *sign, const double *gridF,
                                  const double *gridY, long npts, const double *n_hat, const long *tx,
                                  long n, double *lappr)
{
    const double two_s2 = 2 * m->noise_var;
    for (long s = 0; s < n; ++s) {
        double N[16], D[16];
        const double a_j = m->constellation[tx[s]];
        for (int k = 0; k < m->bps; ++k) { N[k] = 0; D[k] = 0; }
        for (long i = 0; i < m->order; ++i) {
            double yh;
            qro_g_inv(m, sign, gridF, gridY, npts, &n_hat[s], &i, 1, &yh);
            volatile double sq = (yh - a_j) * (yh - a_j);
            const double e = exp(-sq / two_s2);
            long q = i;
            for (int k = 0; k < m->bps; ++k) {
                if ((q * (q + 1)) & 3) D[k] += e;
                else N[k] += e;
                q >>= 1;
            }
        }
        for (int k = 0; k < m->bps; ++k) lappr[s * m->bps + k] = log(N[k]) - log(D[k]);
    }
}

/* demap_lappr_sofisticated(_array) -- noisemapper.pyx:624-766 ("formulation 3"), as written:      */
/* every hypothetical sample is g_inv(n, j) (j, not i: :656-657), the exponent is divided by        */
/* 2 sigma^2 in both branches (:666-675), A_i = beta_i S - dF_i B (:733).                           */
QRO_API void qro_demap_sofisticated(const qro_mapper *m, const uint8_t *sign, const double *gridF,
                                    const double *gridY, long npts, const double *n_hat, const long *tx,
                                    long n, double *lappr)
{
    const int M = m->order;
    const double two_s2 = 2 * m->noise_var, sqrt2sigma = sqrt(two_s2);
    const double *a = m->constellation, *p = m->probabilities;
    for (long s = 0; s < n; ++s) {
        const long j = tx[s];
        const double a_j = a[j];
        double yh, beta[256], dFZ[256], N[16], D[16], S = 0, B = 0;
        qro_g_inv(m, sign, gridF, gridY, npts, &n_hat[s], &j, 1, &yh);
        for (int i = 0; i < M; ++i) {
            double e = p[j];
            for (long q = 0; q < j; ++q) {
                volatile double pr = (2 * yh - a[q] - a_j) * (a[q] - a_j);
                volatile double t = p[q] * exp(pr / two_s2);
                e += t;
            }
            for (long q = j + 1; q < M; ++q) {
                volatile double pr = (2 * yh - a[q] - a[j]) * (a[q] - a_j);
                volatile double t = p[q] * exp(pr / two_s2);
                e += t;
            }
            beta[i] = m->delta_F_Y[i] / e;
            B += beta[i];
            dFZ[i] = 0.5 * (erf((yh - a_j) / sqrt2sigma) - m->inf_erf_table[i * M + j]);
            S += dFZ[i];
        }
        for (int k = 0; k < m->bps; ++k) { N[k] = 0; D[k] = 0; }
        for (int i = 0; i < M; ++i) {
            volatile double t1 = beta[i] * S, t2 = dFZ[i] * B;
            const double A = t1 - t2;
            long q = i;
            for (int k = 0; k < m->bps; ++k) {
                if ((q * (q + 1)) & 3) D[k] += A;
                else N[k] += A;
                q >>= 1;
            }
        }
        for (int k = 0; k < m->bps; ++k) lappr[s * m->bps + k] = log(N[k]) - log(D[k]);
    }
}

/* ------------------------------------------------------------------------- */
/* Mutual-information Monte Carlo                  */
/* montecarlo_information -- mutual_information.pyx:212-300, for GIVEN         */
/* samples (x_ind, y): the reference draws them with numpy's global RNG       */
/* (:236-239); everything after the draw is restated here, accumulated in     */
/* sample order and divided by N last (:293-295).  which[3] as at :217.       */
static double mi_inner(const qro_mapper *m, double yh, long xi)
{
    const double *a = m->constellation, *p = m->probabilities;
    const double x = a[xi], two_s2 = 2.0 * m->noise_var;
    double tmp = p[xi];
    for (long q = 0; q < m->order; ++q) {
        if (q == xi) continue;
        volatile double pr = (2 * yh - x - a[q]) * (a[q] - x);
        volatile double t = p[q] * exp(pr / two_s2);
        tmp += t;
    }
    return tmp;
}

QRO_API void qro_information(const qro_mapper *m, const uint8_t *sign_g, const double *gridF,
                             const double *gridY, long npts, const double *p_Xhat, const long *x_ind,
                             const double *y, long N, const uint8_t *which, double *out3)
{
    const int M = m->order;
    const double *a = m->constellation, *p = m->probabilities;
    const double two_s2 = 2.0 * m->noise_var;
    double I0 = 0, I1 = 0, I2 = 0;
    for (long s = 0; s < N; ++s) {
        const long xi = x_ind[s];
        long xh = region_search(m->thresholds, M + 1, y[s]);
        if (xh == M) xh = M - 1;
        double nv;
        qro_map_noise_sign(m, sign_g, &y[s], &xh, 1, &nv);
        const double x = a[xi];
        if (which[0]) I0 += log2(p_Xhat[xh] / m->fwrd[xi * M + xh]);
        if (which[1]) {
            double tmp = p[xi];
            for (long k = 0; k < M; ++k) {
                if (k == xi) continue;
                volatile double pr = (2 * y[s] - a[k] - x) * (a[k] - x);
                volatile double t = p[k] * exp(pr / two_s2);
                tmp += t;
            }
            I1 += log2(tmp);
        }
        if (which[2]) {
            double acc = 0;
            for (long k = 0; k < M; ++k) {
                if (k == xh) continue;
                double yh;
                qro_g_inv(m, sign_g, gridF, gridY, npts, &nv, &k, 1, &yh);
                acc += m->delta_F_Y[k] / mi_inner(m, yh, xi);
            }
            const double yh = qro_g_inv_search(m, nv, (int)xh, 1e-9);
            volatile double r = mi_inner(m, yh, xi) / m->delta_F_Y[xh];
            acc *= r;
            acc += 1;
            acc *= p_Xhat[xh];
            I2 -= log2(acc);
        }
    }
    out3[0] = I0 / N; out3[1] = I1 / N; out3[2] = I2 / N;
}
