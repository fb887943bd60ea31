/*
 * zigflac_lpc.h -- LPC subframes for the CPU oracle (TEST INFRASTRUCTURE, included by zigflac_oracle.c only).
 *
 * NOT A RESTATEMENT.  The reference has no LPC: encoder.zig:626-640 declares an unused `Prediction` enum, the union arm
 * at :694-699 is commented out, readme.md:24-27 lists linear prediction as "progressing".  BASELINE.json's config 4
 * ("LPC order 12 with quantised coefficients, once the reference LPC path lands") therefore has nothing to be
 * byte-identical to.  This file DEFINES the arithmetic of the LPC extension ("zf-LPC v1") so that the CUDA path has an
 * exact CPU statement to be checked against bit for bit; the bar against the outside world is the FLAC format itself:
 * the independent decoder (flac_decode.c) must return the PCM, and the stream must not be larger than the FIXED-only one.
 * It is switched on by zo_config.lpc_order > 0; with 0 nothing in here runs and the oracle is the reference's path.
 *
 * zf-LPC v1, per candidate channel (samples x[0..N) after the wasted-bits shift, bps bits each), max order M <= 32:
 *   1. window     W[i] = 16384 - floor((2i - (N-1))^2 * 16384 / (N-1)^2)          (Welch, integer, 0..16384)
 *                 xw[i] = ((x[i] >> sh) * W[i]) >> 14,   sh = max(0, bps - 24)     (arithmetic shifts)
 *   2. autocorrelation  R[l] = sum_{i>=l} xw[i] * xw[i-l],  l = 0..M               (exact, int64)
 *   3. Levinson-Durbin in IEEE double, one rounding per operation, in the order written in zl_levinson()
 *   4. order      the smallest o whose prediction error is within (1 + delta)^(omax - o) of the error at omax,
 *                 delta = (bps + P) * 2 ln 2 / N  (what a coefficient and a warm-up sample cost against N/2 log2 err)
 *   5. quantise   precision P = 14 (bps <= 17) or 15 bits, shift = P - exponent(max |c|) clamped to 15, rounding with
 *                 error feedback, floor(v + 0.5)
 *   6. residual   r[i] = x[i] - ((sum_j q[j] * x[i-1-j]) >> shift),  int64 accumulation; must fit 32 bits
 *   7. Rice       the reference's partition / parameter search (rice.zig) with pred_order = o
 *   8. choice     LPC iff  rice_bits < N * bps  and  rice_bits + o * (bps + P) + 9  <  the FIXED / VERBATIM alternative's
 *                 bits + order * bps; that cost (overheads included) is also what the stereo decision adds up
 */
#ifndef ZIGFLAC_LPC_H
#define ZIGFLAC_LPC_H

#define ZL_MAX_ORDER 32
#define ZL_MIN_BLOCK 64 /* shorter frames are left to the FIXED predictors */

typedef struct {
    int valid;
    unsigned order, shift, precision;
    int32_t q[ZL_MAX_ORDER];
} zl_model;

static void zl_window(uint32_t n, uint16_t *w) {
    if (n == 1) { w[0] = 16384; return; }
    const uint64_t den = (uint64_t)(n - 1) * (n - 1);
    for (uint32_t i = 0; i < n; i++) {
        const int64_t d = 2 * (int64_t)i - (int64_t)(n - 1);
        w[i] = (uint16_t)(16384u - (uint32_t)(((uint64_t)(d * d) << 14) / den));
    }
}

/* exponent e of v = m * 2^e, m in [0.5, 1), for a positive normal double (what frexp returns, from the bits) */
static int zl_exponent(double v) {
    uint64_t bits;
    memcpy(&bits, &v, 8);
    return (int)((bits >> 52) & 0x7ff) - 1022;
}

/* Steps 3-5.  R[0..M] exact autocorrelation; returns the quantised model. */
static void zl_levinson(const int64_t *R, unsigned M, uint32_t n, unsigned bps, zl_model *m) {
    double lpc[ZL_MAX_ORDER], coef[ZL_MAX_ORDER][ZL_MAX_ORDER], error[ZL_MAX_ORDER];
    m->valid = 0;
    if (R[0] == 0) return;
    double err = (double)R[0];
    unsigned omax = 0;
    for (unsigned i = 0; i < M; i++) {
        double r = -(double)R[i + 1];
        for (unsigned j = 0; j < i; j++) r = r - lpc[j] * (double)R[i - j];
        r = r / err;
        lpc[i] = r;
        for (unsigned j = 0; j < (i >> 1); j++) {
            const double tmp = lpc[j];
            lpc[j] = lpc[j] + r * lpc[i - 1 - j];
            lpc[i - 1 - j] = lpc[i - 1 - j] + r * tmp;
        }
        if (i & 1) lpc[i >> 1] = lpc[i >> 1] + lpc[i >> 1] * r;
        err = err * (1.0 - r * r);
        if (!(err > 0.0)) break; /* perfectly predictable or numerically spent: keep the orders below */
        for (unsigned j = 0; j <= i; j++) coef[i][j] = -lpc[j];
        error[i] = err;
        omax = i + 1;
    }
    if (omax == 0) return;
    const unsigned P = bps <= 17 ? 14u : 15u;
    const double delta = ((double)(bps + P) * 1.3862943611198906) / (double)n;
    unsigned o = omax;
    double thr = error[omax - 1];
    for (unsigned k = omax - 1; k >= 1; k--) {
        thr = thr * (1.0 + delta);
        if (error[k - 1] <= thr) o = k;
    }
    const double *c = coef[o - 1];
    double cmax = 0.0;
    for (unsigned j = 0; j < o; j++) {
        const double a = c[j] < 0.0 ? -c[j] : c[j];
        if (a > cmax) cmax = a;
    }
    if (!(cmax > 0.0) || !(cmax < 1e300)) return;
    int shift = (int)P - zl_exponent(cmax);
    if (shift > 15) shift = 15;
    if (shift < 0) return;
    const double qmax = (double)((1 << (P - 1)) - 1), qmin = -(double)(1 << (P - 1));
    const double scale = (double)(1 << shift);
    double e = 0.0;
    for (unsigned j = 0; j < o; j++) {
        e = e + c[j] * scale;
        double v = floor(e + 0.5);
        if (v > qmax) v = qmax;
        if (v < qmin) v = qmin;
        m->q[j] = (int32_t)v;
        e = e - v;
    }
    m->order = o;
    m->shift = (unsigned)shift;
    m->precision = P;
    m->valid = 1;
}

/* Steps 1-6 for one plane; residuals_dst[i >= order] receives the residual.  Returns 0 when LPC is not applicable. */
static int zl_analyse(const int32_t *s32, const int64_t *s64, size_t len, unsigned bps, unsigned max_order,
                      int32_t *residuals_dst, zl_model *m) {
    m->valid = 0;
    if (len < ZL_MIN_BLOCK || max_order == 0) return 0;
    unsigned M = max_order > ZL_MAX_ORDER ? ZL_MAX_ORDER : max_order;
    static _Thread_local uint16_t w[65536];
    static _Thread_local int64_t xw[65536];
    static _Thread_local uint32_t w_len = 0;
    if (w_len != (uint32_t)len) { zl_window((uint32_t)len, w); w_len = (uint32_t)len; }
    const unsigned sh = bps > 24 ? bps - 24 : 0;
    for (size_t i = 0; i < len; i++) {
        const int64_t x = s64 ? s64[i] : (int64_t)s32[i];
        xw[i] = ((x >> sh) * (int64_t)w[i]) >> 14;
    }
    int64_t R[ZL_MAX_ORDER + 1];
    for (unsigned l = 0; l <= M; l++) {
        int64_t acc = 0;
        for (size_t i = l; i < len; i++) acc += xw[i] * xw[i - l];
        R[l] = acc;
    }
    zl_levinson(R, M, (uint32_t)len, bps, m);
    if (!m->valid) return 0;
    for (size_t i = m->order; i < len; i++) {
        int64_t sum = 0;
        for (unsigned j = 0; j < m->order; j++) {
            const int64_t x = s64 ? s64[i - 1 - j] : (int64_t)s32[i - 1 - j];
            sum += (int64_t)m->q[j] * x;
        }
        const int64_t cur = s64 ? s64[i] : (int64_t)s32[i];
        const int64_t r = cur - (sum >> m->shift);
        if (r > INT32_MAX || r < -(int64_t)INT32_MAX) { m->valid = 0; return 0; }
        residuals_dst[i] = (int32_t)r;
    }
    return 1;
}

#endif
