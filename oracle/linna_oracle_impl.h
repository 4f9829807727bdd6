/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference algorithm.
 *
 * This header is included twice by linna_oracle.c, once with REAL=float / SUF=f32
 * and once with REAL=double / SUF=f64.  Every function cites the reference lines it
 * restates (paths relative to /root/reference).  Nothing in the product imports,
 * links or executes this code: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py do, as the checker or the timed
 * CPU baseline.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* op descriptor layout in the int array: 5 ints per op */
#define OP_STRIDE 5
#define OP_KIND(o) ops[(o) * OP_STRIDE + 0]
#define OP_IN(o) ops[(o) * OP_STRIDE + 1]
#define OP_MID(o) ops[(o) * OP_STRIDE + 2]
#define OP_OUT(o) ops[(o) * OP_STRIDE + 3]
#define OP_ACT(o) ops[(o) * OP_STRIDE + 4]

/* y[n] = b[n] + sum_k W[n][k] x[k]      (torch nn.Linear: linna/nn.py:25-31, :77-88) */
/* Dot product with 8 interleaved partial sums combined pairwise -- the summation shape of the
 * SIMD BLAS kernels the reference runs on (torch -> MKL/oneDNN sgemm), so that the float32
 * oracle carries the same O(sqrt K) rounding growth instead of a sequential sum's O(K). */
static REAL FN(dot)(const REAL *a, const REAL *b, int K)
{
    REAL s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int k = 0;
    for (; k + 8 <= K; k += 8)
        for (int j = 0; j < 8; ++j) s[j] += a[k + j] * b[k + j];
    for (int j = 0; k < K; ++k, ++j) s[j] += a[k] * b[k];
    return ((s[0] + s[4]) + (s[2] + s[6])) + ((s[1] + s[5]) + (s[3] + s[7]));
}

static void FN(linear)(const REAL *W, const REAL *b, const REAL *x, REAL *y, int K, int N)
{
    for (int n = 0; n < N; ++n) {
        REAL acc = FN(dot)(W + (size_t)n * K, x, K);
        y[n] = b ? acc + b[n] : acc;
    }
}

/* gx[k] += sum_n W[n][k] gy[n] */
static void FN(linear_bwd_x)(const REAL *W, const REAL *gy, REAL *gx, int K, int N)
{
    for (int n = 0; n < N; ++n) {
        REAL g = gy[n];
        if (g == (REAL)0) continue;
        const REAL *w = W + (size_t)n * K;
        for (int k = 0; k < K; ++k) gx[k] += w[k] * g;
    }
}

/* gW[n][k] += gy[n] x[k] ; gb[n] += gy[n] */
static void FN(linear_bwd_w)(const REAL *x, const REAL *gy, REAL *gW, REAL *gb, int K, int N)
{
    for (int n = 0; n < N; ++n) {
        REAL g = gy[n];
        if (gb) gb[n] += g;
        if (g == (REAL)0) continue;
        REAL *w = gW + (size_t)n * K;
        for (int k = 0; k < K; ++k) w[k] += g * x[k];
    }
}

static size_t FN(op_nweights)(const int *ops, int o)
{
    if (OP_KIND(o) == 0) return (size_t)OP_IN(o) * OP_OUT(o) + OP_OUT(o);
    size_t n = (size_t)OP_IN(o) * OP_MID(o) + OP_MID(o) + (size_t)OP_MID(o) * OP_OUT(o) + OP_OUT(o);
    if (OP_IN(o) != OP_OUT(o)) n += (size_t)OP_IN(o) * OP_OUT(o);
    return n;
}

static int FN(max_width)(const int *ops, int nops, int n_in, int n_out)
{
    int w = n_in > n_out ? n_in : n_out;
    for (int o = 0; o < nops; ++o) {
        if (OP_IN(o) > w) w = OP_IN(o);
        if (OP_OUT(o) > w) w = OP_OUT(o);
        if (OP_MID(o) > w) w = OP_MID(o);
    }
    return w;
}

/* Network forward for ONE sample, saving every activation (needed by the backward).
 * acts: (nops+1) pointers-worth of storage laid out by the caller via offsets:
 *   act[0] = xhat, act[o+1] = output of op o;  hid[o] = res-block hidden (post relu).
 * ChtoModelv2.forward linna/nn.py:110-133, ResBlock_batchnorm.forward linna/nn.py:45-56,
 * ChtoModelv2_linear.forward linna/nn.py:184-196, ChtoModelsimple.forward :351-374. */
static void FN(net_forward)(const int *ops, int nops, const REAL *weights, int has_linear, int n_in, int n_out,
                            REAL **act, REAL **hid, const REAL *alpha)
{
    const REAL *wp = weights;
    for (int o = 0; o < nops; ++o) {
        int K = OP_IN(o), N = OP_OUT(o), C = OP_MID(o);
        const REAL *x = act[o];
        REAL *y = act[o + 1];
        if (OP_KIND(o) == 0) {
            const REAL *W = wp, *b = wp + (size_t)K * N;
            FN(linear)(W, b, x, y, K, N);
            if (OP_ACT(o))
                for (int n = 0; n < N; ++n) y[n] = y[n] > 0 ? y[n] : (REAL)0;  /* F.relu nn.py:121,125,126 */
        } else {
            const REAL *W1 = wp, *b1 = W1 + (size_t)K * C, *W2 = b1 + C, *b2 = W2 + (size_t)C * N;
            const REAL *Ws = (K != N) ? b2 + N : 0;
            REAL *h = hid[o];
            FN(linear)(W1, b1, x, h, K, C);
            for (int c = 0; c < C; ++c) h[c] = h[c] > 0 ? h[c] : (REAL)0;       /* nn.py:53 */
            FN(linear)(W2, b2, h, y, C, N);
            for (int n = 0; n < N; ++n) {                                       /* nn.py:54 */
                REAL s = 0;
                if (Ws) s = FN(dot)(Ws + (size_t)n * K, x, K);
                else s = x[n];
                REAL z = y[n] * alpha[o] + s;
                y[n] = z > 0 ? z : (REAL)0;
            }
        }
        wp += FN(op_nweights)(ops, o);
    }
    if (has_linear) {  /* s = layer8(s_in) + 1E-3*linearlayer(s)  nn.py:193 */
        const REAL *Wl = wp, *bl = wp + (size_t)n_in * n_out;
        REAL *y = act[nops];
        const REAL *x = act[0];
        for (int n = 0; n < n_out; ++n) {
            REAL acc = bl[n];
            for (int k = 0; k < n_in; ++k) acc += Wl[(size_t)n * n_in + k] * x[k];
            y[n] += (REAL)1e-3 * acc;
        }
    }
}

/* Backward of net_forward for one sample.  gact[nops] holds d/d(yhat) on entry; on exit gact[0]
 * holds d/d(xhat).  If gweights != NULL, parameter gradients are ACCUMULATED into it (same
 * flat layout as weights).  This restates what torch.autograd does for the reference graph. */
static void FN(net_backward)(const int *ops, int nops, const REAL *weights, REAL *gweights, int has_linear,
                             int n_in, int n_out, REAL **act, REAL **hid, REAL **gact, REAL *ghid,
                             const REAL *alpha)
{
    /* offsets of each op's weights */
    size_t off[64];
    size_t tot = 0;
    for (int o = 0; o < nops; ++o) { off[o] = tot; tot += FN(op_nweights)(ops, o); }
    for (int o = 0; o < nops; ++o) for (int k = 0; k < OP_IN(o); ++k) gact[o][k] = 0;
    if (has_linear) {
        const REAL *Wl = weights + tot;
        const REAL *gy = gact[nops];
        for (int n = 0; n < n_out; ++n) {
            REAL g = (REAL)1e-3 * gy[n];
            for (int k = 0; k < n_in; ++k) gact[0][k] += Wl[(size_t)n * n_in + k] * g;
            if (gweights) {
                REAL *gWl = gweights + tot, *gbl = gWl + (size_t)n_in * n_out;
                for (int k = 0; k < n_in; ++k) gWl[(size_t)n * n_in + k] += g * act[0][k];
                gbl[n] += g;
            }
        }
    }
    for (int o = nops - 1; o >= 0; --o) {
        int K = OP_IN(o), N = OP_OUT(o), C = OP_MID(o);
        const REAL *wp = weights + off[o];
        REAL *gw = gweights ? gweights + off[o] : 0;
        const REAL *x = act[o], *y = act[o + 1];
        REAL *gy = gact[o + 1], *gx = gact[o];
        if (OP_KIND(o) == 0) {
            if (OP_ACT(o)) for (int n = 0; n < N; ++n) if (!(y[n] > 0)) gy[n] = 0;
            FN(linear_bwd_x)(wp, gy, gx, K, N);
            if (gw) FN(linear_bwd_w)(x, gy, gw, gw + (size_t)K * N, K, N);
        } else {
            const REAL *W1 = wp, *W2 = W1 + (size_t)K * C + C;
            const REAL *Ws = (K != N) ? W2 + (size_t)C * N + N : 0;
            const REAL *h = hid[o];
            for (int n = 0; n < N; ++n) if (!(y[n] > 0)) gy[n] = 0;     /* relu of nn.py:54 */
            /* skip branch */
            if (Ws) {
                FN(linear_bwd_x)(Ws, gy, gx, K, N);
                if (gw) FN(linear_bwd_w)(x, gy, gw + (size_t)K * C + C + (size_t)C * N + N, 0, K, N);
            } else {
                for (int n = 0; n < N; ++n) gx[n] += gy[n];
            }
            /* 0.1 * layer2(h) branch */
            for (int c = 0; c < C; ++c) ghid[c] = 0;
            for (int n = 0; n < N; ++n) {
                REAL g = alpha[o] * gy[n];
                if (g == (REAL)0) continue;
                for (int c = 0; c < C; ++c) ghid[c] += W2[(size_t)n * C + c] * g;
                if (gw) {
                    REAL *gW2 = gw + (size_t)K * C + C;
                    for (int c = 0; c < C; ++c) gW2[(size_t)n * C + c] += g * h[c];
                    gW2[(size_t)C * N + n] += g;
                }
            }
            for (int c = 0; c < C; ++c) if (!(h[c] > 0)) ghid[c] = 0;   /* relu of nn.py:53 */
            FN(linear_bwd_x)(W1, ghid, gx, K, C);
            if (gw) FN(linear_bwd_w)(x, ghid, gw, gw + (size_t)K * C, K, C);
        }
    }
}

typedef struct {
    REAL **act, **hid, **gact;
    REAL *ghid;
    REAL *store;
} FN(ws_t);

static int FN(ws_alloc)(FN(ws_t) * ws, const int *ops, int nops, int n_in, int n_out)
{
    size_t tot = n_in;
    int maxmid = 1;
    for (int o = 0; o < nops; ++o) { tot += OP_OUT(o) + OP_MID(o); if (OP_MID(o) > maxmid) maxmid = OP_MID(o); }
    ws->store = (REAL *)calloc(2 * tot + maxmid + 16, sizeof(REAL));
    ws->act = (REAL **)calloc(3 * (nops + 1), sizeof(REAL *));
    if (!ws->store || !ws->act) return -1;
    ws->hid = ws->act + (nops + 1);
    ws->gact = ws->hid + (nops + 1);
    REAL *p = ws->store;
    ws->act[0] = p; p += n_in;
    for (int o = 0; o < nops; ++o) { ws->act[o + 1] = p; p += OP_OUT(o); ws->hid[o] = p; p += OP_MID(o); }
    ws->gact[0] = p; p += n_in;
    for (int o = 0; o < nops; ++o) { ws->gact[o + 1] = p; p += OP_OUT(o) + OP_MID(o); }
    ws->ghid = p;
    return 0;
}

static void FN(ws_free)(FN(ws_t) * ws) { free(ws->store); free(ws->act); }

/* lnP(u) for n rows, optional gradient d lnP/du, optional model vectors.
 *   Transform.__call__ + gauss2unif      linna/util.py:323-347, :291-300
 *   X_transform_class.__call__           linna/util.py:483-497
 *   Y_transform_class.__call__           linna/util.py:532-542
 *   Y_invtransform_data.__call__         linna/util.py:457-458
 *   gaussianlogliklihood                 linna/util.py:953-955   (dense d C^-1 d^T, as the reference)
 *   lnprior                              linna/util.py:1160-1165
 *   Log_prob.__call__                    linna/util.py:990-1021  (lnL/T + lnprior; NaN -> -inf)
 * prior_kind: 0 gauss, 1 flat. */
int FN(linna_oracle_lnp)(const int *ops, int nops, const REAL *weights, int has_linear, int n_in, int n_out,
                         const REAL *alpha, const int *prior_kind, const REAL *prior_a1, const REAL *prior_a2,
                         const int *log10_flag, const REAL *x_mean, const REAL *x_std, const REAL *y_mean,
                         const REAL *y_std, int ypositive, const REAL *sigma, const REAL *data,
                         const REAL *invcov, REAL temperature, const REAL *u, long n, REAL *lnp_out,
                         REAL *grad_out, REAL *m_out, REAL *yhat_out, REAL *theta_out)
{
    FN(ws_t) ws;
    if (nops > 60 || FN(ws_alloc)(&ws, ops, nops, n_in, n_out)) return -1;
    REAL *theta = (REAL *)malloc(sizeof(REAL) * (size_t)(n_in + 4 * n_out));
    REAL *yv = theta + n_in, *d = yv + n_out, *cd = d + n_out, *mv = cd + n_out;
    const REAL SQRT2 = (REAL)1.4142135623730951, LN10 = (REAL)2.302585092994046;
    for (long r = 0; r < n; ++r) {
        const REAL *ur = u + (size_t)r * n_in;
        REAL lnprior = 0;
        for (int i = 0; i < n_in; ++i) {
            if (prior_kind[i] == 0) theta[i] = ur[i] * prior_a2[i] + prior_a1[i];
            else theta[i] = ((REAL)0.5 * ((REAL)1 + (REAL)erf((double)(ur[i] / SQRT2)))) * (prior_a2[i] - prior_a1[i]) + prior_a1[i];
            lnprior += ur[i] * ur[i];
            REAL tp = log10_flag[i] ? (REAL)log10((double)theta[i]) : theta[i];
            ws.act[0][i] = (tp - x_mean[i]) / x_std[i];
        }
        lnprior *= (REAL)-0.5;
        if (theta_out) memcpy(theta_out + (size_t)r * n_in, theta, sizeof(REAL) * n_in);
        FN(net_forward)(ops, nops, weights, has_linear, n_in, n_out, ws.act, ws.hid, alpha);
        const REAL *yhat = ws.act[nops];
        if (yhat_out) memcpy(yhat_out + (size_t)r * n_out, yhat, sizeof(REAL) * n_out);
        for (int j = 0; j < n_out; ++j) {
            REAL y = yhat[j] * y_std[j] + y_mean[j];
            if (ypositive) y = (REAL)exp((double)y);
            yv[j] = y;
            mv[j] = y * sigma[j];
            d[j] = mv[j] - data[j];
        }
        if (m_out) memcpy(m_out + (size_t)r * n_out, mv, sizeof(REAL) * n_out);
        /* (d @ invcov) @ d.T * (-0.5) */
        REAL chi2 = 0;
        for (int j = 0; j < n_out; ++j) cd[j] = 0;
        for (int i0 = 0; i0 < n_out; i0 += 8) {           /* blocked over i: 8-term partial sums */
            for (int j = 0; j < n_out; ++j) {
                REAL s8 = 0;
                for (int i = i0; i < i0 + 8 && i < n_out; ++i) s8 += d[i] * invcov[(size_t)i * n_out + j];
                cd[j] += s8;
            }
        }
        chi2 = FN(dot)(cd, d, n_out);
        REAL lnp = (REAL)-0.5 * chi2 / temperature + lnprior;
        if (lnp != lnp) lnp = -(REAL)INFINITY;
        lnp_out[r] = lnp;
        if (grad_out) {
            /* d lnL/d d = -(1/2T) (C^-1 + C^-T) d */
            REAL *gy = ws.gact[nops];
            for (int i = 0; i < n_out; ++i) {
                REAL s = 0;
                const REAL *row = invcov + (size_t)i * n_out;
                for (int j = 0; j < n_out; ++j) s += row[j] * d[j];
                REAL gm = (REAL)-0.5 * (s + cd[i]) / temperature;
                REAL gyv = gm * sigma[i];                      /* through util.py:458 */
                if (ypositive) gyv *= yv[i];                   /* d exp */
                gy[i] = gyv * y_std[i];                        /* through util.py:542 */
            }
            FN(net_backward)(ops, nops, weights, 0, has_linear, n_in, n_out, ws.act, ws.hid, ws.gact, ws.ghid, alpha);
            for (int i = 0; i < n_in; ++i) {
                REAL g = ws.gact[0][i] / x_std[i];
                if (log10_flag[i]) g /= (theta[i] * LN10);
                if (prior_kind[i] == 0) g *= prior_a2[i];
                else g *= (prior_a2[i] - prior_a1[i]) * (REAL)0.3989422804014327 * (REAL)exp((double)((REAL)-0.5 * ur[i] * ur[i]));
                grad_out[(size_t)r * n_in + i] = g - ur[i];
            }
        }
    }
    free(theta);
    FN(ws_free)(&ws);
    return 0;
}

/* One optimiser step of the emulator training inner loop.
 *   Predictor.train inner loop             linna/predictor_gpu.py:273-288
 *   Auxilleryfunc.__call__ / Loss_fn       linna/util.py:1070-1088, :1105-1116
 *   torch.optim.AdamW (lr, wd=1e-4)        linna/predictor_gpu.py:267
 * icov_n = inverse of the normalised covariance (util.py:1060-1064, computed by the caller in
 * float64 and cast), data_n = normalised data vector (util.py:1069).  Y holds targets in
 * physical units.  If do_update == 0 only loss/gradients are produced. */
int FN(linna_oracle_train_step)(const int *ops, int nops, REAL *weights, size_t nweights, int has_linear,
                                int n_in, int n_out, const REAL *alpha, const int *log10_flag,
                                const REAL *x_mean, const REAL *x_std, const REAL *y_mean, const REAL *y_std,
                                int ypositive, const REAL *sigma, const REAL *data_n, const REAL *icov_n,
                                const REAL *X, const REAL *Y, long B, REAL *adam_m, REAL *adam_v, long step,
                                REAL lr, REAL beta1, REAL beta2, REAL eps, REAL wd, int do_update,
                                REAL *loss_out, REAL *grads_out, REAL *loss_rows, REAL *chisq_md,
                                REAL *chisq_nnd, REAL *yhat_out)
{
    FN(ws_t) ws;
    if (nops > 60 || FN(ws_alloc)(&ws, ops, nops, n_in, n_out)) return -1;
    REAL *g = grads_out ? grads_out : (REAL *)malloc(sizeof(REAL) * nweights);
    memset(g, 0, sizeof(REAL) * nweights);
    REAL *t = (REAL *)malloc(sizeof(REAL) * (size_t)(5 * n_out));
    REAL *delta = t + n_out, *cd = delta + n_out, *msk = cd + n_out, *tmp = msk + n_out;
    double loss = 0;
    for (long r = 0; r < B; ++r) {
        const REAL *xr = X + (size_t)r * n_in, *yr = Y + (size_t)r * n_out;
        for (int i = 0; i < n_in; ++i) {
            REAL tp = log10_flag[i] ? (REAL)log10((double)xr[i]) : xr[i];
            ws.act[0][i] = (tp - x_mean[i]) / x_std[i];
        }
        FN(net_forward)(ops, nops, weights, has_linear, n_in, n_out, ws.act, ws.hid, alpha);
        const REAL *yhat = ws.act[nops];
        if (yhat_out) memcpy(yhat_out + (size_t)r * n_out, yhat, sizeof(REAL) * n_out);
        for (int j = 0; j < n_out; ++j) {
            REAL v = yr[j] / sigma[j];                                  /* util.py:432 */
            v = ypositive ? ((REAL)log((double)v) - y_mean[j]) / y_std[j] : (v - y_mean[j]) / y_std[j]; /* :567-570 */
            t[j] = v;
            msk[j] = (yr[j] == (REAL)1e-30 || yr[j] == (REAL)1e10 || data_n[j] == (REAL)1e-30) ? (REAL)0 : (REAL)1; /* :1072 */
        }
        /* chisqnnd: pred vs data */
        REAL c_nnd = 0, c_md = 0, c_mnn = 0;
        for (int pass = 0; pass < 3; ++pass) {
            for (int j = 0; j < n_out; ++j) {
                REAL dv = pass == 0 ? (yhat[j] - data_n[j]) : pass == 1 ? (t[j] - data_n[j]) : (t[j] - yhat[j]);
                tmp[j] = dv * msk[j];
            }
            for (int j = 0; j < n_out; ++j) cd[j] = 0;
            for (int i = 0; i < n_out; ++i) {
                REAL di = tmp[i];
                if (di == (REAL)0) continue;
                const REAL *row = icov_n + (size_t)i * n_out;
                for (int j = 0; j < n_out; ++j) cd[j] += di * row[j];
            }
            REAL s = 0;
            for (int j = 0; j < n_out; ++j) s += cd[j] * tmp[j];
            if (pass == 0) c_nnd = s; else if (pass == 1) c_md = s; else { c_mnn = s; memcpy(delta, tmp, sizeof(REAL) * n_out); }
        }
        if (c_md < (REAL)0.5 * n_out) c_md = (REAL)0.5 * n_out;        /* util.py:1086 */
        REAL lrow = c_mnn / c_md;                                       /* util.py:1087 */
        if (loss_rows) loss_rows[r] = lrow;
        if (chisq_md) chisq_md[r] = c_md;
        if (chisq_nnd) chisq_nnd[r] = c_nnd;
        loss += lrow;
        /* d loss / d yhat = -(C^-1 + C^-T) delta * mask / (c_md * B) ; cd currently = delta @ icov */
        REAL *gy = ws.gact[nops];
        for (int i = 0; i < n_out; ++i) {
            REAL s = 0;
            const REAL *row = icov_n + (size_t)i * n_out;
            for (int j = 0; j < n_out; ++j) s += row[j] * delta[j];
            gy[i] = -(s + cd[i]) * msk[i] / (c_md * (REAL)B);
        }
        FN(net_backward)(ops, nops, weights, g, has_linear, n_in, n_out, ws.act, ws.hid, ws.gact, ws.ghid, alpha);
    }
    *loss_out = (REAL)(loss / (double)B);
    if (do_update) {
        /* torch.optim.AdamW single-tensor update (decoupled decay, amsgrad=False) */
        REAL bc1 = (REAL)1 - (REAL)pow((double)beta1, (double)step);
        REAL bc2 = (REAL)1 - (REAL)pow((double)beta2, (double)step);
        REAL step_size = lr / bc1, bc2s = (REAL)sqrt((double)bc2);
        for (size_t i = 0; i < nweights; ++i) {
            REAL p = weights[i] * ((REAL)1 - lr * wd);
            adam_m[i] = beta1 * adam_m[i] + ((REAL)1 - beta1) * g[i];
            adam_v[i] = beta2 * adam_v[i] + ((REAL)1 - beta2) * g[i] * g[i];
            REAL denom = (REAL)sqrt((double)adam_v[i]) / bc2s + eps;
            weights[i] = p - step_size * (adam_m[i] / denom);
        }
    }
    if (!grads_out) free(g);
    free(t);
    FN(ws_free)(&ws);
    return 0;
}

#undef OP_STRIDE
#undef OP_KIND
#undef OP_IN
#undef OP_MID
#undef OP_OUT
#undef OP_ACT
#undef FN
#undef CAT
#undef CAT_
