/*
 * odpd_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's recurrent-backbone forward + I/Q MSE + backward
 * (lab-emi/OpenDPD).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product (opendpd_b200/) never does.
 *
 * Parity pin: the reference holds NO golden vectors or numeric tests for this path,
 * so this oracle is pinned against outputs of the reference itself, generated in the authoring
 * container by oracle/make_golden.py (fixtures in tests/golden/, checked by tests/test_oracle_golden.py).
 *
 * Every function cites the reference file:line it restates.  The arithmetic of nn.GRU / nn.LSTM lives in
 * PyTorch ATen (third-party, torch>=2.4 per pyproject.toml:34; container has 2.11.0):
 *   GRU cell  (aten/src/ATen/native/RNN.cpp, GRUCell):  r=σ(Wir x+bir+Whr h+bhr), z=σ(...),
 *             n=tanh(Win x+bin + r*(Whn h+bhn)), h'=(h-n)*z+n      gate order r,z,n
 *   LSTM cell (LSTMCell): gates i,f,g,o ; c'=f*c+i*g ; h'=o*tanh(c')
 *
 * Compiled twice (REAL=float / REAL=double) into one shared object, see oracle/Makefile.
 * Flat parameter layout = named_parameters() order of the reference module (documented in include/odpd.h).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REAL
#define REAL float
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#ifdef REAL_IS_DOUBLE
#define SUF(n) CAT(n, _f64)
#define R_EXP exp
#define R_TANH tanh
#define R_SQRT sqrt
#define R_FABS fabs
#define R_ATAN2 atan2
#define R_SIN sin
#define R_COS cos
#define R_POW pow
#else
#define SUF(n) CAT(n, _f32)
#define R_EXP expf
#define R_TANH tanhf
#define R_SQRT sqrtf
#define R_FABS fabsf
#define R_ATAN2 atan2f
#define R_SIN sinf
#define R_COS cosf
#define R_POW powf
#endif

enum { CELL_GRU = 0, CELL_LSTM = 1, CELL_DGRU = 2, CELL_DELTAGRU = 3, CELL_TRES = 4, CELL_PGJANET = 5,
       CELL_DVRJANET = 6, CELL_GMP = 7, CELL_QGRU = 8, CELL_QGRU_AMP1 = 9, CELL_QGRU_QAT = 10, CELL_QGRU_AMP1_QAT = 11, CELL_RVTDCNN = 13, CELL_BOJANET = 14, CELL_TCNN = 15, CELL_NEURALTX = 16, CELL_APNRRU = 17, CELL_MCLDNN = 18, CELL_DELTAJANET = 19, CELL_TRES_QAT = 20 };

typedef struct {
    int cell, B, T, H, K;
    REAL thx, thh;
    const REAL *params;
    int want_gx, want_gp;
} Ctx;

static inline REAL sigm(REAL v) { return (REAL)1 / ((REAL)1 + R_EXP(-v)); }

/* y[r] (+)= sum_c W[r*ld+c]*v[c] */
static inline REAL dotv(const REAL *w, const REAL *v, int n) {
    REAL s = 0;
    for (int i = 0; i < n; ++i) s += w[i] * v[i];
    return s;
}

/* ---------------------------------------------------------------- features
 * gru.py:45 (raw I,Q) · dgru.py:61-68 (I,Q,a,a^3,sin,cos) · deltagru.py:61-72 (same order)
 * qgru.py:61-66 (I,Q,a^2,a^4) · qgru_amp1.py:61-70 (I,Q,a,a^3) · deltagru_tcnskip.py:91-100 (I,Q,a,a^3,I_nxt,Q_nxt)
 * ATen pow(x,2)=x*x, pow(x,3)=(x*x)*x, each op separately rounded (no FMA). */
static int n_features(int cell) {
    switch (cell) {
    case CELL_GRU: case CELL_LSTM: return 2;
    case CELL_QGRU: case CELL_QGRU_AMP1: case CELL_QGRU_QAT: case CELL_QGRU_AMP1_QAT: return 4;
    default: return 6;
    }
}

static int base_cell(int cell) { return cell == CELL_QGRU_QAT ? CELL_QGRU : (cell == CELL_QGRU_AMP1_QAT ? CELL_QGRU_AMP1 : cell); }

static void features_fwd(int cell, const REAL *x, int T, int t, REAL *f) {
    cell = base_cell(cell);
    volatile REAL i = x[2 * t], q = x[2 * t + 1];
    volatile REAL ii = i * i, qq = q * q;
    volatile REAL a2 = ii + qq;
    f[0] = i; f[1] = q;
    if (cell == CELL_GRU || cell == CELL_LSTM) return;
    if (cell == CELL_QGRU) { volatile REAL a4 = a2 * a2; f[2] = a2; f[3] = a4; return; }
    volatile REAL a = R_SQRT(a2);
    volatile REAL aa = a * a;
    volatile REAL a3 = aa * a;
    f[2] = a; f[3] = a3;
    if (cell == CELL_QGRU_AMP1) return;
    if (cell == CELL_TRES) { int tn = (t + 1) % T; f[4] = x[2 * tn]; f[5] = x[2 * tn + 1]; return; }
    f[4] = q / a; /* sin */
    f[5] = i / a; /* cos */
}

/* accumulate d(loss)/dx from d(loss)/dfeatures; gx is the whole (T,2) row because TRES rolls. */
static void features_bwd(int cell, const REAL *x, int T, int t, const REAL *gf, REAL *gx) {
    cell = base_cell(cell);
    REAL i = x[2 * t], q = x[2 * t + 1];
    REAL gi = gf[0], gq = gf[1];
    if (cell == CELL_GRU || cell == CELL_LSTM) { gx[2 * t] += gi; gx[2 * t + 1] += gq; return; }
    REAL a2 = i * i + q * q;
    if (cell == CELL_QGRU) {
        REAL ga2 = gf[2] + (REAL)2 * a2 * gf[3];
        gx[2 * t] += gi + (REAL)2 * i * ga2; gx[2 * t + 1] += gq + (REAL)2 * q * ga2; return;
    }
    REAL a = R_SQRT(a2);
    REAL ga = gf[2] + (REAL)3 * a * a * gf[3];
    if (cell == CELL_DGRU || cell == CELL_DELTAGRU) {
        REAL gsin = gf[4], gcos = gf[5];
        ga -= (q * gsin + i * gcos) / a2;
        gi += gcos / a; gq += gsin / a;
    } else if (cell == CELL_TRES) {
        int tn = (t + 1) % T;
        gx[2 * tn] += gf[4]; gx[2 * tn + 1] += gf[5];
    }
    gx[2 * t] += gi + ga * i / a; gx[2 * t + 1] += gq + ga * q / a;
}

/* ================================================================ GRU family: gru / dgru / qgru / qgru_amp1
 * gru.py:45-48, dgru.py:59-74, qgru.py:59-71, qgru_amp1.py:59-76 + ATen GRUCell.
 * params: W_ih(3H,F) W_hh(3H,H) b_ih(3H) b_hh(3H) fc_out.W(2,H[+6]) fc_out.b(2) [fc_hid.W(H,H) fc_hid.b(H)] */
static void seq_gru_family(const Ctx *c, const REAL *x, const REAL *gout_or_null, REAL *out, REAL *gx, REAL *gp,
                           int phase) {
    const int H = c->H, T = c->T, F = n_features(c->cell), dg = (c->cell == CELL_DGRU);
    const int O = dg ? H + F : H;
    const REAL *Wih = c->params, *Whh = Wih + 3 * H * F, *bih = Whh + 3 * H * H, *bhh = bih + 3 * H;
    const REAL *Wo = bhh + 3 * H, *bo = Wo + 2 * O, *Wh = bo + 2, *bh = Wh + H * H;
    /* saved per step: f(F) r z n hgn h g  */
    static __thread REAL *sv = NULL; static __thread size_t sv_n = 0;
    const int S = F + 6 * H;
    size_t need = (size_t)T * S;
    if (sv_n < need) { free(sv); sv = (REAL *)malloc(need * sizeof(REAL)); sv_n = need; }
    if (phase == 0) {
        REAL h[64] = {0}, cat[64 + 8];
        for (int t = 0; t < T; ++t) {
            REAL *s = sv + (size_t)t * S, *f = s, *r = s + F, *z = r + H, *n = z + H, *hgn = n + H, *hs = hgn + H,
                 *g = hs + H;
            features_fwd(c->cell, x, T, t, f);
            REAL hn[64];
            for (int j = 0; j < H; ++j) {
                REAL xr = dotv(Wih + (size_t)j * F, f, F) + bih[j];
                REAL xz = dotv(Wih + (size_t)(H + j) * F, f, F) + bih[H + j];
                REAL xn = dotv(Wih + (size_t)(2 * H + j) * F, f, F) + bih[2 * H + j];
                REAL hr = dotv(Whh + (size_t)j * H, h, H) + bhh[j];
                REAL hz = dotv(Whh + (size_t)(H + j) * H, h, H) + bhh[H + j];
                REAL hnn = dotv(Whh + (size_t)(2 * H + j) * H, h, H) + bhh[2 * H + j];
                r[j] = sigm(hr + xr); z[j] = sigm(hz + xz); hgn[j] = hnn;
                n[j] = R_TANH(xn + hnn * r[j]);
                hn[j] = (h[j] - n[j]) * z[j] + n[j];
            }
            memcpy(h, hn, sizeof(REAL) * H); memcpy(hs, h, sizeof(REAL) * H);
            if (dg) {
                for (int j = 0; j < H; ++j) { REAL p = dotv(Wh + (size_t)j * H, h, H) + bh[j]; g[j] = p > 0 ? p : 0; cat[j] = g[j]; }
                for (int k = 0; k < F; ++k) cat[H + k] = f[k];
            } else memcpy(cat, h, sizeof(REAL) * H);
            out[2 * t] = dotv(Wo, cat, O) + bo[0];
            out[2 * t + 1] = dotv(Wo + O, cat, O) + bo[1];
        }
        return;
    }
    /* backward */
    REAL *gWih = gp, *gWhh = gWih + 3 * H * F, *gbih = gWhh + 3 * H * H, *gbhh = gbih + 3 * H;
    REAL *gWo = gbhh + 3 * H, *gbo = gWo + 2 * O, *gWh = gbo + 2, *gbh = gWh + H * H;
    REAL gH[64] = {0};
    for (int t = T - 1; t >= 0; --t) {
        REAL *s = sv + (size_t)t * S, *f = s, *r = s + F, *z = r + H, *n = z + H, *hgn = n + H, *hs = hgn + H, *g = hs + H;
        const REAL *hp = t ? (sv + (size_t)(t - 1) * S + F + 4 * H) : NULL;
        const REAL go0 = gout_or_null[2 * t], go1 = gout_or_null[2 * t + 1];
        REAL gf[8] = {0};
        if (dg) {
            REAL dpre[64];
            for (int j = 0; j < H; ++j) {
                REAL dgj = Wo[j] * go0 + Wo[O + j] * go1;
                dpre[j] = g[j] > 0 ? dgj : 0;
                gWo[j] += go0 * g[j]; gWo[O + j] += go1 * g[j];
            }
            for (int k = 0; k < F; ++k) {
                gf[k] += Wo[H + k] * go0 + Wo[O + H + k] * go1;
                gWo[H + k] += go0 * f[k]; gWo[O + H + k] += go1 * f[k];
            }
            for (int j = 0; j < H; ++j) {
                gbh[j] += dpre[j];
                for (int k = 0; k < H; ++k) { gWh[j * H + k] += dpre[j] * hs[k]; gH[k] += Wh[j * H + k] * dpre[j]; }
            }
        } else {
            for (int j = 0; j < H; ++j) {
                gH[j] += Wo[j] * go0 + Wo[O + j] * go1;
                gWo[j] += go0 * hs[j]; gWo[O + j] += go1 * hs[j];
            }
        }
        gbo[0] += go0; gbo[1] += go1;
        REAL dx3[192], dh3[192], ghp[64];
        for (int j = 0; j < H; ++j) {
            REAL hpj = hp ? hp[j] : 0;
            REAL gz = gH[j] * (hpj - n[j]), gn = gH[j] * ((REAL)1 - z[j]);
            ghp[j] = gH[j] * z[j];
            REAL an = gn * ((REAL)1 - n[j] * n[j]);
            REAL az = gz * z[j] * ((REAL)1 - z[j]);
            REAL ar = an * hgn[j] * r[j] * ((REAL)1 - r[j]);
            dx3[j] = ar; dx3[H + j] = az; dx3[2 * H + j] = an;
            dh3[j] = ar; dh3[H + j] = az; dh3[2 * H + j] = an * r[j];
        }
        for (int k = 0; k < 3 * H; ++k) {
            gbih[k] += dx3[k]; gbhh[k] += dh3[k];
            for (int q = 0; q < F; ++q) { gWih[k * F + q] += dx3[k] * f[q]; gf[q] += Wih[k * F + q] * dx3[k]; }
            for (int q = 0; q < H; ++q) { gWhh[k * H + q] += dh3[k] * (hp ? hp[q] : 0); ghp[q] += Whh[k * H + q] * dh3[k]; }
        }
        memcpy(gH, ghp, sizeof(REAL) * H);
        if (gx) features_bwd(c->cell, x, T, t, gf, gx);
    }
}

/* ================================================================ LSTM: lstm.py:45-48 + ATen LSTMCell, (h0,c0)=(0,0)
 * params: W_ih(4H,2) W_hh(4H,H) b_ih(4H) b_hh(4H) fc_out.W(2,H) fc_out.b(2) */
static void seq_lstm(const Ctx *c, const REAL *x, const REAL *gout, REAL *out, REAL *gx, REAL *gp, int phase) {
    const int H = c->H, T = c->T, F = 2;
    const REAL *Wih = c->params, *Whh = Wih + 4 * H * F, *bih = Whh + 4 * H * H, *bhh = bih + 4 * H;
    const REAL *Wo = bhh + 4 * H, *bo = Wo + 2 * H;
    static __thread REAL *sv = NULL; static __thread size_t sv_n = 0;
    const int S = 7 * H; /* i f g o c h tanh(c) */
    size_t need = (size_t)T * S;
    if (sv_n < need) { free(sv); sv = (REAL *)malloc(need * sizeof(REAL)); sv_n = need; }
    if (phase == 0) {
        REAL h[64] = {0}, cc[64] = {0};
        for (int t = 0; t < T; ++t) {
            REAL *s = sv + (size_t)t * S; const REAL *f = x + 2 * t; REAL hn[64];
            for (int j = 0; j < H; ++j) {
                REAL a[4];
                for (int g = 0; g < 4; ++g)
                    a[g] = (dotv(Wih + (size_t)(g * H + j) * F, f, F) + bih[g * H + j]) +
                           (dotv(Whh + (size_t)(g * H + j) * H, h, H) + bhh[g * H + j]);
                REAL ig = sigm(a[0]), fg = sigm(a[1]), gg = R_TANH(a[2]), og = sigm(a[3]);
                REAL cn = fg * cc[j] + ig * gg, tc = R_TANH(cn);
                s[j] = ig; s[H + j] = fg; s[2 * H + j] = gg; s[3 * H + j] = og; s[4 * H + j] = cn; s[6 * H + j] = tc;
                cc[j] = cn; hn[j] = og * tc;
            }
            memcpy(h, hn, sizeof(REAL) * H); memcpy(s + 5 * H, h, sizeof(REAL) * H);
            out[2 * t] = dotv(Wo, h, H) + bo[0]; out[2 * t + 1] = dotv(Wo + H, h, H) + bo[1];
        }
        return;
    }
    REAL *gWih = gp, *gWhh = gWih + 4 * H * F, *gbih = gWhh + 4 * H * H, *gbhh = gbih + 4 * H, *gWo = gbhh + 4 * H,
         *gbo = gWo + 2 * H;
    REAL gH[64] = {0}, gC[64] = {0};
    for (int t = T - 1; t >= 0; --t) {
        REAL *s = sv + (size_t)t * S; const REAL *f = x + 2 * t;
        const REAL *sp = t ? sv + (size_t)(t - 1) * S : NULL;
        REAL go0 = gout[2 * t], go1 = gout[2 * t + 1], d[256], ghp[64] = {0};
        for (int j = 0; j < H; ++j) {
            gH[j] += Wo[j] * go0 + Wo[H + j] * go1;
            gWo[j] += go0 * s[5 * H + j]; gWo[H + j] += go1 * s[5 * H + j];
        }
        gbo[0] += go0; gbo[1] += go1;
        for (int j = 0; j < H; ++j) {
            REAL ig = s[j], fg = s[H + j], gg = s[2 * H + j], og = s[3 * H + j], tc = s[6 * H + j];
            REAL cp = sp ? sp[4 * H + j] : 0;
            REAL go = gH[j] * tc;
            REAL gc = gC[j] + gH[j] * og * ((REAL)1 - tc * tc);
            d[j] = gc * gg * ig * ((REAL)1 - ig);
            d[H + j] = gc * cp * fg * ((REAL)1 - fg);
            d[2 * H + j] = gc * ig * ((REAL)1 - gg * gg);
            d[3 * H + j] = go * og * ((REAL)1 - og);
            gC[j] = gc * fg;
        }
        REAL gf[2] = {0, 0};
        for (int k = 0; k < 4 * H; ++k) {
            gbih[k] += d[k]; gbhh[k] += d[k];
            for (int q = 0; q < F; ++q) { gWih[k * F + q] += d[k] * f[q]; gf[q] += Wih[k * F + q] * d[k]; }
            for (int q = 0; q < H; ++q) { gWhh[k * H + q] += d[k] * (sp ? sp[5 * H + q] : 0); ghp[q] += Whh[k * H + q] * d[k]; }
        }
        memcpy(gH, ghp, sizeof(REAL) * H);
        if (gx) { gx[2 * t] += gf[0]; gx[2 * t + 1] += gf[1]; }
    }
}

/* ================================================================ Delta GRU cells
 * deltagru.py:60-77,149-266 (biases folded into M_0, :164-170) and deltagru_tcnskip.py:89-103,232-304
 * (bias-free x2h/h2h, TCN skip :32-49, out=fc_out(h)+skip :102).  Backward: SURVEY.md §8a-D.
 * deltagru params: W_ih(3H,6) W_hh(3H,H) b_ih(3H) b_hh(3H) fc_out.W(2,H) fc_out.b(2)
 * tres     params: x2h.W(3H,6) h2h.W(3H,H) fc_out.W(2,H) tcn.0.W(3,2,3) tcn.2.W(2,3,1) */
static inline REAL hardswish(REAL v) {
    REAL t = v + (REAL)3; t = t < 0 ? 0 : (t > 6 ? 6 : t);
    return v * t / (REAL)6;
}
static inline REAL hardswish_grad(REAL v) { /* ATen hardswish_backward */
    return v < (REAL)-3 ? 0 : (v <= (REAL)3 ? v / (REAL)3 + (REAL)0.5 : (REAL)1);
}

/* Optional diagnostic for the guard-band test of the delta-h mask (SURVEY.md §7 hard part 1): when set, the forward also stores
 * |delta_h| - thh (signed distance of the compare `abs(delta) >= th`, deltagru.py:176-183) for every (b,t,unit) into a caller
 * buffer (B,T,H); addressed through the sequence's mask_h row, so it needs want_masks. */
static REAL *SUF(g_dh_margin) = NULL;
static const uint64_t *SUF(g_dh_margin_mask_base) = NULL;
void SUF(odpd_oracle_set_dh_margin)(REAL *buf, const uint64_t *mask_h_base) { SUF(g_dh_margin) = buf; SUF(g_dh_margin_mask_base) = mask_h_base; }

static void seq_delta(const Ctx *c, const REAL *x, const REAL *gout, REAL *out, REAL *gx, REAL *gp, int phase,
                      uint64_t *mask_x, uint64_t *mask_h, int64_t *stats) {
    const int H = c->H, T = c->T, F = 6, tres = (c->cell == CELL_TRES);
    const REAL *Wih = c->params, *Whh = Wih + 3 * H * F;
    const REAL *bih = tres ? NULL : Whh + 3 * H * H, *bhh = tres ? NULL : bih + 3 * H;
    const REAL *Wo = tres ? Whh + 3 * H * H : bhh + 3 * H;
    const REAL *bo = tres ? NULL : Wo + 2 * H;
    const REAL *w0 = tres ? Wo + 2 * H : NULL, *w2 = tres ? w0 + 18 : NULL;
    /* saved per step: f(6) dx(6) dh(H) r z n Mnh h c1(3) c2(2) + masks */
    static __thread REAL *sv = NULL; static __thread size_t sv_n = 0;
    static __thread uint64_t *mk = NULL; static __thread size_t mk_n = 0;
    const int S = 12 + 6 * H + 5;
    size_t need = (size_t)T * S;
    if (sv_n < need) { free(sv); sv = (REAL *)malloc(need * sizeof(REAL)); sv_n = need; }
    if (mk_n < (size_t)2 * T) { free(mk); mk = (uint64_t *)malloc(sizeof(uint64_t) * 2 * T); mk_n = 2 * T; }
    if (phase == 0) {
        REAL h[64] = {0}, hp[64] = {0}, xp[6] = {0}, M[192], Mnh[64];
        for (int j = 0; j < H; ++j) {
            M[j] = tres ? 0 : bih[j] + bhh[j]; M[H + j] = tres ? 0 : bih[H + j] + bhh[H + j];
            M[2 * H + j] = tres ? 0 : bih[2 * H + j]; Mnh[j] = tres ? 0 : bhh[2 * H + j];
        }
        int64_t zx = 0, zh = 0;
        for (int t = 0; t < T; ++t) {
            REAL *s = sv + (size_t)t * S, *f = s, *dx = s + 6, *dh = s + 12, *r = dh + H, *z = r + H, *n = z + H,
                 *mn = n + H, *hs = mn + H, *cv = hs + H;
            features_fwd(c->cell, x, T, t, f);
            uint64_t mx = 0, mh = 0;
            for (int k = 0; k < F; ++k) {
                REAL d = f[k] - xp[k], a = R_FABS(d);
                if (a < c->thx) d = 0;
                if (a >= c->thx) { xp[k] = f[k]; mx |= (uint64_t)1 << k; }
                dx[k] = d; zx += (d == 0);
            }
            for (int j = 0; j < H; ++j) {
                REAL d = h[j] - hp[j], a = R_FABS(d);
                if (mask_h && SUF(g_dh_margin))
                    SUF(g_dh_margin)[((size_t)(mask_h - SUF(g_dh_margin_mask_base)) + t) * H + j] = a - c->thh;
                if (a < c->thh) d = 0;
                if (a >= c->thh) { hp[j] = h[j]; mh |= (uint64_t)1 << j; }
                dh[j] = d; zh += (d == 0);
            }
            mk[2 * t] = mx; mk[2 * t + 1] = mh;
            if (mask_x) mask_x[t] = mx;
            if (mask_h) mask_h[t] = mh;
            REAL hn[64];
            for (int j = 0; j < H; ++j) {
                REAL mxr = dotv(Wih + (size_t)j * F, dx, F) + M[j];
                REAL mxz = dotv(Wih + (size_t)(H + j) * F, dx, F) + M[H + j];
                REAL mxn = dotv(Wih + (size_t)(2 * H + j) * F, dx, F) + M[2 * H + j];
                M[j] = mxr + dotv(Whh + (size_t)j * H, dh, H);
                M[H + j] = mxz + dotv(Whh + (size_t)(H + j) * H, dh, H);
                M[2 * H + j] = mxn;
                Mnh[j] = dotv(Whh + (size_t)(2 * H + j) * H, dh, H) + Mnh[j];
                r[j] = sigm(M[j]); z[j] = sigm(M[H + j]);
                n[j] = R_TANH(M[2 * H + j] + r[j] * Mnh[j]);
                mn[j] = Mnh[j];
                hn[j] = ((REAL)1 - z[j]) * n[j] + z[j] * h[j];
            }
            memcpy(h, hn, sizeof(REAL) * H); memcpy(hs, h, sizeof(REAL) * H);
            REAL o0 = dotv(Wo, h, H), o1 = dotv(Wo + H, h, H);
            if (tres) {
                for (int co = 0; co < 3; ++co) {
                    REAL a = 0;
                    for (int ci = 0; ci < 2; ++ci)
                        for (int k = 0; k < 3; ++k) {
                            int tt = t + (k - 1) * 16;
                            if (tt >= 0 && tt < T) a += w0[(co * 2 + ci) * 3 + k] * x[2 * tt + ci];
                        }
                    cv[co] = a;
                }
                for (int o = 0; o < 2; ++o) {
                    REAL a = 0;
                    for (int ch = 0; ch < 3; ++ch) a += w2[o * 3 + ch] * hardswish(cv[ch]);
                    cv[3 + o] = a;
                }
                o0 += hardswish(cv[3]); o1 += hardswish(cv[4]);
            } else { o0 += bo[0]; o1 += bo[1]; }
            out[2 * t] = o0; out[2 * t + 1] = o1;
        }
        if (stats) { stats[0] = zx; stats[1] = (int64_t)T * F; stats[2] = zh; stats[3] = (int64_t)T * H; }
        return;
    }
    REAL *gWih = gp, *gWhh = gWih + 3 * H * F;
    REAL *gbih = tres ? NULL : gWhh + 3 * H * H, *gbhh = tres ? NULL : gbih + 3 * H;
    REAL *gWo = tres ? gWhh + 3 * H * H : gbhh + 3 * H;
    REAL *gbo = tres ? NULL : gWo + 2 * H;
    REAL *gw0 = tres ? gWo + 2 * H : NULL, *gw2 = tres ? gw0 + 18 : NULL;
    REAL gH[64] = {0}, gM[192] = {0}, gMnh[64] = {0}, gxp[6] = {0}, ghp[64] = {0};
    for (int t = T - 1; t >= 0; --t) {
        REAL *s = sv + (size_t)t * S, *f = s, *dx = s + 6, *dh = s + 12, *r = dh + H, *z = r + H, *n = z + H, *mn = n + H,
             *hs = mn + H, *cv = hs + H;
        const REAL *hprev = t ? sv + (size_t)(t - 1) * S + 12 + 5 * H : NULL;
        REAL go0 = gout[2 * t], go1 = gout[2 * t + 1];
        for (int j = 0; j < H; ++j) {
            gH[j] += Wo[j] * go0 + Wo[H + j] * go1;
            gWo[j] += go0 * hs[j]; gWo[H + j] += go1 * hs[j];
        }
        if (tres) {
            REAL g2[2] = {go0 * hardswish_grad(cv[3]), go1 * hardswish_grad(cv[4])};
            REAL ga1[3] = {0, 0, 0};
            for (int o = 0; o < 2; ++o)
                for (int ch = 0; ch < 3; ++ch) { gw2[o * 3 + ch] += g2[o] * hardswish(cv[ch]); ga1[ch] += w2[o * 3 + ch] * g2[o]; }
            for (int co = 0; co < 3; ++co) {
                REAL gc1 = ga1[co] * hardswish_grad(cv[co]);
                for (int ci = 0; ci < 2; ++ci)
                    for (int k = 0; k < 3; ++k) {
                        int tt = t + (k - 1) * 16;
                        if (tt >= 0 && tt < T) {
                            gw0[(co * 2 + ci) * 3 + k] += gc1 * x[2 * tt + ci];
                            if (gx) gx[2 * tt + ci] += w0[(co * 2 + ci) * 3 + k] * gc1;
                        }
                    }
            }
        } else { gbo[0] += go0; gbo[1] += go1; }
        REAL ghprev[64];
        for (int j = 0; j < H; ++j) {
            REAL hpj = hprev ? hprev[j] : 0;
            REAL gz = gH[j] * (hpj - n[j]), gn = gH[j] * ((REAL)1 - z[j]);
            ghprev[j] = gH[j] * z[j];
            REAL ga = gn * ((REAL)1 - n[j] * n[j]);
            gM[j] += ga * mn[j] * r[j] * ((REAL)1 - r[j]);
            gM[H + j] += gz * z[j] * ((REAL)1 - z[j]);
            gM[2 * H + j] += ga;
            gMnh[j] += ga * r[j];
        }
        REAL gdx[6] = {0}, gdh[64] = {0};
        for (int k = 0; k < 3 * H; ++k) {
            REAL gk = gM[k], gk_h = (k < 2 * H) ? gM[k] : gMnh[k - 2 * H];
            for (int q = 0; q < F; ++q) { gWih[k * F + q] += gk * dx[q]; gdx[q] += Wih[k * F + q] * gk; }
            for (int q = 0; q < H; ++q) { gWhh[k * H + q] += gk_h * dh[q]; gdh[q] += Whh[k * H + q] * gk_h; }
        }
        uint64_t mx = mk[2 * t], mh = mk[2 * t + 1];
        REAL gf[6];
        for (int k = 0; k < F; ++k) {
            if ((mx >> k) & 1) { gf[k] = gxp[k] + gdx[k]; gxp[k] = -gdx[k]; } else gf[k] = 0;
        }
        for (int j = 0; j < H; ++j) {
            if ((mh >> j) & 1) { ghprev[j] += ghp[j] + gdh[j]; ghp[j] = -gdh[j]; }
        }
        memcpy(gH, ghprev, sizeof(REAL) * H);
        if (gx) features_bwd(c->cell, x, T, t, gf, gx);
    }
    if (!tres)
        for (int j = 0; j < H; ++j) {
            gbih[j] += gM[j]; gbhh[j] += gM[j]; gbih[H + j] += gM[H + j]; gbhh[H + j] += gM[H + j];
            gbih[2 * H + j] += gM[2 * H + j]; gbhh[2 * H + j] += gMnh[j];
        }
}

/* ================================================================ PGJANET: pgjanet.py:24-77, h = h_0[0] = 0
 * params: W_a(H,H+1) b_a W_p1(H,H+1) b_p1 W_p2(H,H+1) b_p2 W_f(H,2H) b_f W_g(H,2H) b_g W_o(2,H) b_o(2) */
static void seq_pgjanet(const Ctx *c, const REAL *x, const REAL *gout, REAL *out, REAL *gx, REAL *gp, int phase) {
    const int H = c->H, T = c->T, H1 = H + 1, H2 = 2 * H;
    const REAL *Wa = c->params, *ba = Wa + H * H1, *Wp1 = ba + H, *bp1 = Wp1 + H * H1, *Wp2 = bp1 + H, *bp2 = Wp2 + H * H1;
    const REAL *Wf = bp2 + H, *bf = Wf + H * H2, *Wg = bf + H, *bg = Wg + H * H2, *Wo = bg + H, *bo = Wo + 2 * H;
    static __thread REAL *sv = NULL; static __thread size_t sv_n = 0;
    const int S = 4 + 7 * H; /* a cos sin pad | an p1 p2 u f g h */
    size_t need = (size_t)T * S;
    if (sv_n < need) { free(sv); sv = (REAL *)malloc(need * sizeof(REAL)); sv_n = need; }
    if (phase == 0) {
        REAL h[64] = {0};
        for (int t = 0; t < T; ++t) {
            REAL *s = sv + (size_t)t * S, *an = s + 4, *p1 = an + H, *p2 = p1 + H, *u = p2 + H, *fg = u + H, *gg = fg + H,
                 *hs = gg + H;
            REAL i = x[2 * t], q = x[2 * t + 1];
            REAL a = R_SQRT(i * i + q * q), th = R_ATAN2(q, i), ct = R_COS(th), st = R_SIN(th);
            s[0] = a; s[1] = ct; s[2] = st;
            for (int j = 0; j < H; ++j) {
                an[j] = R_TANH(dotv(Wa + (size_t)j * H1, h, H) + Wa[j * H1 + H] * a + ba[j]);
                p1[j] = R_TANH(dotv(Wp1 + (size_t)j * H1, h, H) + Wp1[j * H1 + H] * ct + bp1[j]);
                p2[j] = R_TANH(dotv(Wp2 + (size_t)j * H1, h, H) + Wp2[j * H1 + H] * st + bp2[j]);
                u[j] = an[j] * p1[j] * p2[j] * ((REAL)1 - an[j]) * ((REAL)1 - p1[j]) * ((REAL)1 - p2[j]);
            }
            REAL hn[64];
            for (int j = 0; j < H; ++j) {
                fg[j] = sigm(dotv(Wf + (size_t)j * H2, h, H) + dotv(Wf + (size_t)j * H2 + H, u, H) + bf[j]);
                gg[j] = R_TANH(dotv(Wg + (size_t)j * H2, h, H) + dotv(Wg + (size_t)j * H2 + H, u, H) + bg[j]);
                hn[j] = fg[j] * h[j] + ((REAL)1 - fg[j]) * gg[j];
            }
            memcpy(h, hn, sizeof(REAL) * H); memcpy(hs, h, sizeof(REAL) * H);
            out[2 * t] = dotv(Wo, h, H) + bo[0]; out[2 * t + 1] = dotv(Wo + H, h, H) + bo[1];
        }
        return;
    }
    REAL *gWa = gp, *gba = gWa + H * H1, *gWp1 = gba + H, *gbp1 = gWp1 + H * H1, *gWp2 = gbp1 + H, *gbp2 = gWp2 + H * H1;
    REAL *gWf = gbp2 + H, *gbf = gWf + H * H2, *gWg = gbf + H, *gbg = gWg + H * H2, *gWo = gbg + H, *gbo = gWo + 2 * H;
    REAL gH[64] = {0};
    for (int t = T - 1; t >= 0; --t) {
        REAL *s = sv + (size_t)t * S, *an = s + 4, *p1 = an + H, *p2 = p1 + H, *u = p2 + H, *fg = u + H, *gg = fg + H, *hs = gg + H;
        REAL hp[64];
        for (int j = 0; j < H; ++j) hp[j] = t ? (sv + (size_t)(t - 1) * S + 4 + 6 * H)[j] : 0;
        REAL go0 = gout[2 * t], go1 = gout[2 * t + 1];
        for (int j = 0; j < H; ++j) {
            gH[j] += Wo[j] * go0 + Wo[H + j] * go1;
            gWo[j] += go0 * hs[j]; gWo[H + j] += go1 * hs[j];
        }
        gbo[0] += go0; gbo[1] += go1;
        REAL ghp[64], gu[64] = {0};
        REAL af[64], ag[64];
        for (int j = 0; j < H; ++j) {
            REAL gf_ = gH[j] * (hp[j] - gg[j]), gg_ = gH[j] * ((REAL)1 - fg[j]);
            ghp[j] = gH[j] * fg[j];
            af[j] = gf_ * fg[j] * ((REAL)1 - fg[j]); ag[j] = gg_ * ((REAL)1 - gg[j] * gg[j]);
        }
        for (int j = 0; j < H; ++j) {
            gbf[j] += af[j]; gbg[j] += ag[j];
            for (int k = 0; k < H; ++k) {
                gWf[j * H2 + k] += af[j] * hp[k]; gWf[j * H2 + H + k] += af[j] * u[k];
                gWg[j * H2 + k] += ag[j] * hp[k]; gWg[j * H2 + H + k] += ag[j] * u[k];
                ghp[k] += Wf[j * H2 + k] * af[j] + Wg[j * H2 + k] * ag[j];
                gu[k] += Wf[j * H2 + H + k] * af[j] + Wg[j * H2 + H + k] * ag[j];
            }
        }
        REAL ga = 0, gc = 0, gs = 0;
        for (int j = 0; j < H; ++j) {
            REAL A = an[j] * ((REAL)1 - an[j]), P1 = p1[j] * ((REAL)1 - p1[j]), P2 = p2[j] * ((REAL)1 - p2[j]);
            REAL gan = gu[j] * ((REAL)1 - (REAL)2 * an[j]) * P1 * P2;
            REAL gp1_ = gu[j] * A * ((REAL)1 - (REAL)2 * p1[j]) * P2;
            REAL gp2_ = gu[j] * A * P1 * ((REAL)1 - (REAL)2 * p2[j]);
            REAL aa = gan * ((REAL)1 - an[j] * an[j]), a1 = gp1_ * ((REAL)1 - p1[j] * p1[j]), a2 = gp2_ * ((REAL)1 - p2[j] * p2[j]);
            gba[j] += aa; gbp1[j] += a1; gbp2[j] += a2;
            for (int k = 0; k < H; ++k) {
                gWa[j * H1 + k] += aa * hp[k]; gWp1[j * H1 + k] += a1 * hp[k]; gWp2[j * H1 + k] += a2 * hp[k];
                ghp[k] += Wa[j * H1 + k] * aa + Wp1[j * H1 + k] * a1 + Wp2[j * H1 + k] * a2;
            }
            gWa[j * H1 + H] += aa * s[0]; gWp1[j * H1 + H] += a1 * s[1]; gWp2[j * H1 + H] += a2 * s[2];
            ga += Wa[j * H1 + H] * aa; gc += Wp1[j * H1 + H] * a1; gs += Wp2[j * H1 + H] * a2;
        }
        memcpy(gH, ghp, sizeof(REAL) * H);
        if (gx) {
            REAL i = x[2 * t], q = x[2 * t + 1], a = s[0], a2 = a * a;
            REAL gth = -s[2] * gc + s[1] * gs;
            gx[2 * t] += ga * i / a - gth * q / a2;
            gx[2 * t + 1] += ga * q / a + gth * i / a2;
        }
    }
}

/* ================================================================ DVRJANET: dvrjanet.py:43-102, dvr_block :32-41
 * params: cs(K) W_ph(H,H) W_pth(H,1) W_ah(H,H) W_ax(H,1) W_f(H,H) b_f W_ccos(H,2H) b W_csin(H,2H) b W_o1(1,H) b W_o2(1,H) b */
static void seq_dvrjanet(const Ctx *c, const REAL *x, const REAL *gout, REAL *out, REAL *gx, REAL *gp, int phase) {
    const int H = c->H, T = c->T, K = c->K, H2 = 2 * H;
    const REAL *cs = c->params, *Wph = cs + K, *Wpt = Wph + H * H, *Wah = Wpt + H, *Wax = Wah + H * H, *Wf = Wax + H, *bf = Wf + H * H;
    const REAL *Wc = bf + H, *bc = Wc + H * H2, *Ws = bc + H, *bs = Ws + H * H2, *Wo1 = bs + H, *bo1 = Wo1 + H, *Wo2 = bo1 + 1,
               *bo2 = Wo2 + H;
    static __thread REAL *sv = NULL; static __thread size_t sv_n = 0;
    const int S = 4 + 10 * H; /* a th | pa at ct st f gc gs hI hQ tht */
    size_t need = (size_t)T * S;
    if (sv_n < need) { free(sv); sv = (REAL *)malloc(need * sizeof(REAL)); sv_n = need; }
    if (phase == 0) {
        REAL hI[64] = {0}, hQ[64] = {0};
        for (int t = 0; t < T; ++t) {
            REAL *s = sv + (size_t)t * S, *pa = s + 4, *at = pa + H, *ct = at + H, *st = ct + H, *fg = st + H, *gc = fg + H,
                 *gs = gc + H, *sI = gs + H, *sQ = sI + H;
            REAL i = x[2 * t], q = x[2 * t + 1];
            REAL a = R_SQRT(i * i + q * q), th = R_ATAN2(q, i);
            s[0] = a; s[1] = th;
            REAL sm[64], vc[64], vs[64];
            for (int j = 0; j < H; ++j) sm[j] = hI[j] + hQ[j];
            for (int j = 0; j < H; ++j) {
                REAL tht = Wpt[j] * th + dotv(Wph + (size_t)j * H, sm, H);
                pa[j] = Wax[j] * a + dotv(Wah + (size_t)j * H, sm, H);
                REAL acc = 0;
                for (int k = 1; k <= K; ++k) acc = acc + R_FABS(pa[j] - (REAL)((double)k / K)) * cs[k - 1];
                at[j] = acc; ct[j] = R_COS(tht); st[j] = R_SIN(tht);
                fg[j] = sigm(dotv(Wf + (size_t)j * H, sm, H) + bf[j]);
                vc[j] = at[j] * ct[j]; vs[j] = at[j] * st[j];
            }
            REAL nI[64], nQ[64];
            for (int j = 0; j < H; ++j) {
                gc[j] = R_TANH(dotv(Wc + (size_t)j * H2, hI, H) + dotv(Wc + (size_t)j * H2 + H, vc, H) + bc[j]);
                gs[j] = R_TANH(dotv(Ws + (size_t)j * H2, hQ, H) + dotv(Ws + (size_t)j * H2 + H, vs, H) + bs[j]);
                nI[j] = fg[j] * hI[j] + ((REAL)1 - fg[j]) * gc[j];
                nQ[j] = fg[j] * hQ[j] + ((REAL)1 - fg[j]) * gs[j];
            }
            memcpy(hI, nI, sizeof(REAL) * H); memcpy(hQ, nQ, sizeof(REAL) * H);
            memcpy(sI, hI, sizeof(REAL) * H); memcpy(sQ, hQ, sizeof(REAL) * H);
            out[2 * t] = dotv(Wo1, hI, H) + bo1[0]; out[2 * t + 1] = dotv(Wo2, hQ, H) + bo2[0];
        }
        return;
    }
    REAL *gcs = gp, *gWph = gcs + K, *gWpt = gWph + H * H, *gWah = gWpt + H, *gWax = gWah + H * H, *gWf = gWax + H, *gbf = gWf + H * H;
    REAL *gWc = gbf + H, *gbc = gWc + H * H2, *gWs = gbc + H, *gbs = gWs + H * H2, *gWo1 = gbs + H, *gbo1 = gWo1 + H, *gWo2 = gbo1 + 1,
         *gbo2 = gWo2 + H;
    REAL gI[64] = {0}, gQ[64] = {0};
    for (int t = T - 1; t >= 0; --t) {
        REAL *s = sv + (size_t)t * S, *pa = s + 4, *at = pa + H, *ct = at + H, *st = ct + H, *fg = st + H, *gc = fg + H, *gs = gc + H,
             *sI = gs + H, *sQ = sI + H;
        REAL hI[64], hQ[64], sm[64], vc[64], vs[64];
        for (int j = 0; j < H; ++j) {
            hI[j] = t ? (sv + (size_t)(t - 1) * S + 4 + 7 * H)[j] : 0;
            hQ[j] = t ? (sv + (size_t)(t - 1) * S + 4 + 8 * H)[j] : 0;
            sm[j] = hI[j] + hQ[j]; vc[j] = at[j] * ct[j]; vs[j] = at[j] * st[j];
        }
        REAL go0 = gout[2 * t], go1 = gout[2 * t + 1];
        for (int j = 0; j < H; ++j) {
            gI[j] += Wo1[j] * go0; gQ[j] += Wo2[j] * go1;
            gWo1[j] += go0 * sI[j]; gWo2[j] += go1 * sQ[j];
        }
        gbo1[0] += go0; gbo2[0] += go1;
        REAL gIp[64], gQp[64], af[64], ac[64], as_[64], gvc[64] = {0}, gvs[64] = {0};
        for (int j = 0; j < H; ++j) {
            REAL gf_ = gI[j] * (hI[j] - gc[j]) + gQ[j] * (hQ[j] - gs[j]);
            af[j] = gf_ * fg[j] * ((REAL)1 - fg[j]);
            ac[j] = gI[j] * ((REAL)1 - fg[j]) * ((REAL)1 - gc[j] * gc[j]);
            as_[j] = gQ[j] * ((REAL)1 - fg[j]) * ((REAL)1 - gs[j] * gs[j]);
            gIp[j] = gI[j] * fg[j]; gQp[j] = gQ[j] * fg[j];
        }
        for (int j = 0; j < H; ++j) {
            gbc[j] += ac[j]; gbs[j] += as_[j]; gbf[j] += af[j];
            for (int k = 0; k < H; ++k) {
                gWc[j * H2 + k] += ac[j] * hI[k]; gWc[j * H2 + H + k] += ac[j] * vc[k];
                gWs[j * H2 + k] += as_[j] * hQ[k]; gWs[j * H2 + H + k] += as_[j] * vs[k];
                gIp[k] += Wc[j * H2 + k] * ac[j]; gvc[k] += Wc[j * H2 + H + k] * ac[j];
                gQp[k] += Ws[j * H2 + k] * as_[j]; gvs[k] += Ws[j * H2 + H + k] * as_[j];
            }
        }
        REAL gsm[64] = {0}, gth = 0, ga = 0;
        for (int j = 0; j < H; ++j) {
            REAL gat = gvc[j] * ct[j] + gvs[j] * st[j];
            REAL gtht = -gvc[j] * at[j] * st[j] + gvs[j] * at[j] * ct[j];
            REAL gpa = 0;
            for (int k = 1; k <= K; ++k) {
                REAL d = pa[j] - (REAL)((double)k / K);
                gcs[k - 1] += gat * R_FABS(d);
                gpa += cs[k - 1] * (d > 0 ? (REAL)1 : (d < 0 ? (REAL)-1 : (REAL)0));
            }
            gpa *= gat;
            gWpt[j] += gtht * s[1]; gWax[j] += gpa * s[0];
            gth += Wpt[j] * gtht; ga += Wax[j] * gpa;
            for (int k = 0; k < H; ++k) {
                gWph[j * H + k] += gtht * sm[k]; gWah[j * H + k] += gpa * sm[k]; gWf[j * H + k] += af[j] * sm[k];
                gsm[k] += Wph[j * H + k] * gtht + Wah[j * H + k] * gpa + Wf[j * H + k] * af[j];
            }
        }
        for (int j = 0; j < H; ++j) { gI[j] = gIp[j] + gsm[j]; gQ[j] = gQp[j] + gsm[j]; }
        if (gx) {
            REAL i = x[2 * t], q = x[2 * t + 1], a = s[0], a2 = a * a;
            gx[2 * t] += ga * i / a - gth * q / a2;
            gx[2 * t + 1] += ga * q / a + gth * i / a2;
        }
    }
}

/* ================================================================ GMP: gmp.py:18-51 (M=11 taps, degree 5; real weights)
 * y[j] = sum_m w[m] x[j+m-10] + sum_{p<4,k<11,m<11} w[11+p*121+k*11+m] x[j+m-10] |x[j+k+m-20]|^(p+1), x[n<0]=0 */
static void seq_gmp(const Ctx *c, const REAL *x, const REAL *gout, REAL *out, REAL *gx, REAL *gp, int phase) {
    const int T = c->T; const REAL *w = c->params;
    for (int j = 0; j < T; ++j) {
        REAL yr = 0, yi = 0;
        REAL g0 = phase ? gout[2 * j] : 0, g1 = phase ? gout[2 * j + 1] : 0;
        for (int m = 0; m < 11; ++m) {
            int a = j + m - 10; if (a < 0) continue;
            REAL ar = x[2 * a], ai = x[2 * a + 1];
            if (!phase) { yr += w[m] * ar; yi += w[m] * ai; }
            else { gp[m] += g0 * ar + g1 * ai; if (gx) { gx[2 * a] += g0 * w[m]; gx[2 * a + 1] += g1 * w[m]; } }
        }
        for (int p = 0; p < 4; ++p)
            for (int k = 0; k < 11; ++k)
                for (int m = 0; m < 11; ++m) {
                    int a = j + m - 10, cc = j + k + m - 20, idx = 11 + p * 121 + k * 11 + m;
                    if (a < 0) continue;           /* x[a]=0 -> term and all its gradients vanish */
                    REAL ar = x[2 * a], ai = x[2 * a + 1];
                    REAL cr = cc >= 0 ? x[2 * cc] : 0, ci = cc >= 0 ? x[2 * cc + 1] : 0;
                    REAL amp = R_SQRT(cr * cr + ci * ci);
                    REAL pw = amp; for (int e = 0; e < p; ++e) pw *= amp;   /* amp^(p+1) */
                    if (!phase) { yr += w[idx] * ar * pw; yi += w[idx] * ai * pw; continue; }
                    gp[idx] += (g0 * ar + g1 * ai) * pw;
                    if (gx) {
                        gx[2 * a] += g0 * w[idx] * pw; gx[2 * a + 1] += g1 * w[idx] * pw;
                        if (cc >= 0 && amp > 0) {
                            REAL pm = 1; for (int e = 0; e < p; ++e) pm *= amp;   /* amp^p */
                            REAL coef = w[idx] * (g0 * ar + g1 * ai) * (REAL)(p + 1) * pm / amp;
                            gx[2 * cc] += coef * cr; gx[2 * cc + 1] += coef * ci;
                        }
                    }
                }
        if (!phase) { out[2 * j] = yr; out[2 * j + 1] = yi; }
    }
}


/* ================================================================ fake-quantised GRU (QAT), config 5
 * quant/modules/gru.py GRUCell.forward :32-61 with the module swaps of quant/quant_envs.py:132-305:
 *   nn.Linear -> INT_Linear (quant_layers.py:70-82: weight and input each fake-quantised, float bias, 16-bit out_quantizer only
 *   for fc_out and only in eval), Sigmoid/Tanh/Add/Mul -> Quant_* (quant_ops.py:14-66).
 * Fake quantiser (quantizers.py:56-81): s = 2^round(log2|scale|); q(v) = s * rne(clamp(v/s, -2^(b-1), 2^(b-1)-1)); backward = STE
 * through round, zero outside the clamp (inclusive bounds, torch.clamp).  The 13 scale parameters get zero gradient.
 * params (named_parameters order): x2h.W(3H,4) x2h.b(3H) s0 s1 s2 | h2h.W(3H,H) h2h.b(3H) s3 s4 s5 | s6(sigmoid) s7(tanh) s8(add) s9(mul)
 *                                  | fc_out.W(2,H) fc_out.b(2) s10 s11 s12.   K packs the bit widths: bits_w | bits_a<<8 | eval<<16. */
typedef struct { REAL s, qn, qp; } Quant;
static Quant mkq(REAL scale, int bits) {
    Quant q; q.s = (REAL)pow(2.0, nearbyint(log2(fabs((double)scale)))); q.qn = -(REAL)pow(2.0, bits - 1); q.qp = (REAL)pow(2.0, bits - 1) - 1; return q;
}
static inline REAL qf(const Quant *q, REAL v, int *inrange) {
    REAL u = v / q->s;
    if (inrange) *inrange = (u >= q->qn && u <= q->qp);
    u = u < q->qn ? q->qn : (u > q->qp ? q->qp : u);
    return (REAL)nearbyint((double)u) * q->s;
}
static void seq_qgru_qat(const Ctx *c, const REAL *x, const REAL *gout, REAL *out, REAL *gx, REAL *gp, int phase) {
    const int H = c->H, T = c->T, F = 4, bw = c->K & 255, ba = (c->K >> 8) & 255, eval = (c->K >> 16) & 1;
    const REAL *p = c->params;
    const REAL *Wx = p, *bx = Wx + 3 * H * F, *sx = bx + 3 * H, *Wh = sx + 3, *bh = Wh + 3 * H * H, *sh = bh + 3 * H, *sop = sh + 3,
               *Wo = sop + 4, *bo = Wo + 2 * H, *so = bo + 2;
    const Quant qxw = mkq(sx[0], bw), qxa = mkq(sx[1], ba), qhw = mkq(sh[0], bw), qha = mkq(sh[1], ba), qsig = mkq(sop[0], ba),
                qtanh = mkq(sop[1], ba), qadd = mkq(sop[2], ba), qmul = mkq(sop[3], ba), qow = mkq(so[0], bw), qoa = mkq(so[1], ba),
                qoo = mkq(so[2], 16);
    static __thread REAL *hs = NULL; static __thread size_t hs_n = 0;
    if (hs_n < (size_t)(T + 1) * H) { free(hs); hs = (REAL *)malloc(sizeof(REAL) * (size_t)(T + 1) * H); hs_n = (size_t)(T + 1) * H; }
    REAL Wxq[192 * 4], Whq[192 * 64], Woq[128]; int cWx[192 * 4], cWh[192 * 64], cWo[128];
    for (int i = 0; i < 3 * H * F; ++i) Wxq[i] = qf(&qxw, Wx[i], &cWx[i]);
    for (int i = 0; i < 3 * H * H; ++i) Whq[i] = qf(&qhw, Wh[i], &cWh[i]);
    for (int i = 0; i < 2 * H; ++i) Woq[i] = qf(&qow, Wo[i], &cWo[i]);
    REAL *gWx = gp, *gbx = gp ? gWx + 3 * H * F : NULL, *gWh = gp ? gbx + 3 * H + 3 : NULL, *gbh = gp ? gWh + 3 * H * H : NULL,
         *gWo = gp ? gbh + 3 * H + 3 + 4 : NULL, *gbo = gp ? gWo + 2 * H : NULL;
    if (phase == 0) for (int j = 0; j < H; ++j) hs[j] = 0;
    REAL gH[64] = {0};
    for (int step = 0; step < T; ++step) {
        const int t = phase == 0 ? step : T - 1 - step;
        const REAL *hp = hs + (size_t)t * H;
        REAL f[8], fq[4]; int cf[4];
        features_fwd(c->cell, x, T, t, f);
        for (int k = 0; k < F; ++k) fq[k] = qf(&qxa, f[k], &cf[k]);
        REAL hq[64]; int chq[64];
        for (int k = 0; k < H; ++k) hq[k] = qf(&qha, hp[k], &chq[k]);
        REAL xg[192], hg[192];
        for (int r = 0; r < 3 * H; ++r) {
            xg[r] = dotv(Wxq + (size_t)r * F, fq, F) + bx[r];
            hg[r] = dotv(Whq + (size_t)r * H, hq, H) + bh[r];
        }
        REAL sr[64], sz[64], tn[64], rr[64], zz[64], nn[64], m1[64], hn[64];
        int c_ar[64], c_az[64], c_an[64], c_r[64], c_z[64], c_n[64], c_m1[64], c_m2[64], c_m3[64], c_h[64];
        for (int j = 0; j < H; ++j) {
            REAL ar = qf(&qadd, xg[j] + hg[j], &c_ar[j]); sr[j] = sigm(ar); rr[j] = qf(&qsig, sr[j], &c_r[j]);
            REAL az = qf(&qadd, xg[H + j] + hg[H + j], &c_az[j]); sz[j] = sigm(az); zz[j] = qf(&qsig, sz[j], &c_z[j]);
            m1[j] = qf(&qmul, rr[j] * hg[2 * H + j], &c_m1[j]);
            REAL an = qf(&qadd, xg[2 * H + j] + m1[j], &c_an[j]); tn[j] = R_TANH(an); nn[j] = qf(&qtanh, tn[j], &c_n[j]);
            REAL m2 = qf(&qmul, zz[j] * hp[j], &c_m2[j]);
            REAL m3 = qf(&qmul, ((REAL)1 - zz[j]) * nn[j], &c_m3[j]);
            hn[j] = qf(&qadd, m2 + m3, &c_h[j]);
        }
        if (phase == 0) {
            memcpy(hs + (size_t)(t + 1) * H, hn, sizeof(REAL) * H);
            REAL qo[64];
            for (int j = 0; j < H; ++j) qo[j] = qf(&qoa, hn[j], NULL);
            for (int o = 0; o < 2; ++o) {
                REAL y = dotv(Woq + (size_t)o * H, qo, H) + bo[o];
                out[2 * t + o] = eval ? qf(&qoo, y, NULL) : y;
            }
            continue;
        }
        /* ---- backward of step t (everything above was recomputed from the saved h_{t-1}) */
        const REAL go[2] = {gout[2 * t], gout[2 * t + 1]};
        for (int j = 0; j < H; ++j) {
            int cq; REAL qo = qf(&qoa, hn[j], &cq);
            for (int o = 0; o < 2; ++o) { if (cWo[o * H + j]) gWo[o * H + j] += go[o] * qo; }
            if (cq) gH[j] += Woq[j] * go[0] + Woq[H + j] * go[1];
        }
        gbo[0] += go[0]; gbo[1] += go[1];
        REAL g_xg[192], g_hg[192], ghp[64];
        for (int j = 0; j < H; ++j) {
            REAL gs3 = c_h[j] ? gH[j] : 0;
            REAL gm2 = c_m2[j] ? gs3 : 0, gm3 = c_m3[j] ? gs3 : 0;
            REAL gz = gm2 * hp[j] - gm3 * nn[j];
            ghp[j] = gm2 * zz[j];
            REAL gn = gm3 * ((REAL)1 - zz[j]);
            REAL gtn = c_n[j] ? gn : 0;
            REAL gan = gtn * ((REAL)1 - tn[j] * tn[j]);
            REAL gsn = c_an[j] ? gan : 0;                       /* d/d(xg_n + m1) */
            REAL gm1 = c_m1[j] ? gsn : 0;
            REAL gr = gm1 * hg[2 * H + j];
            REAL gsr = c_r[j] ? gr : 0;
            REAL gar = gsr * sr[j] * ((REAL)1 - sr[j]);
            REAL gsum_r = c_ar[j] ? gar : 0;
            REAL gsz = c_z[j] ? gz : 0;
            REAL gaz = gsz * sz[j] * ((REAL)1 - sz[j]);
            REAL gsum_z = c_az[j] ? gaz : 0;
            g_xg[j] = gsum_r; g_hg[j] = gsum_r; g_xg[H + j] = gsum_z; g_hg[H + j] = gsum_z;
            g_xg[2 * H + j] = gsn; g_hg[2 * H + j] = gm1 * rr[j];
        }
        REAL gfq[4] = {0, 0, 0, 0}, ghq[64] = {0};
        for (int r = 0; r < 3 * H; ++r) {
            gbx[r] += g_xg[r]; gbh[r] += g_hg[r];
            for (int k = 0; k < F; ++k) { if (cWx[r * F + k]) gWx[r * F + k] += g_xg[r] * fq[k]; gfq[k] += Wxq[r * F + k] * g_xg[r]; }
            for (int k = 0; k < H; ++k) { if (cWh[r * H + k]) gWh[r * H + k] += g_hg[r] * hq[k]; ghq[k] += Whq[r * H + k] * g_hg[r]; }
        }
        for (int k = 0; k < H; ++k) gH[k] = ghp[k] + (chq[k] ? ghq[k] : 0);
        if (gx) { REAL gf[8] = {0}; for (int k = 0; k < F; ++k) gf[k] = cf[k] ? gfq[k] : 0; features_bwd(c->cell, x, T, t, gf, gx); }
    }
}

/* ================================================================ RVTDCNN: rvtdcnn.py:36-62
 * features (I,Q,a,a^2,a^3) (:41-46); window of timestep t = samples (t-3+r) mod T, r = 0..3 (`pad = x[:, -(window_size-1):, :]`, :51-53);
 * Conv2d(1->3, 3x3, padding (1,0)) on the 4x5 image -> tanh -> 36 values in (channel,row,column) order (:57-58); fc_hid(36->H) tanh (:59);
 * fc_out(H->2) (:60).  params: Conv2d.weight(3,1,3,3) Conv2d.bias(3) fc_hid.weight(H,36) fc_hid.bias(H) fc_out.weight(2,H) fc_out.bias(2). */
static void seq_rvtdcnn(const Ctx *c, const REAL *x, const REAL *gout, REAL *out, REAL *gx, REAL *gp, int phase) {
    const int T = c->T, H = c->H;
    const REAL *Wc = c->params, *bc = Wc + 27, *Wh = bc + 3, *bh = Wh + 36 * H, *Wo = bh + H, *bo = Wo + 2 * H;
    for (int t = 0; t < T; ++t) {
        REAL w[4][5], z[36], hk[64];
        int sidx[4];
        for (int r = 0; r < 4; ++r) {
            int s = ((t - 3 + r) % T + T) % T;
            sidx[r] = s;
            volatile REAL i = x[2 * s], q = x[2 * s + 1];
            volatile REAL ii = i * i, qq = q * q;
            volatile REAL a2 = ii + qq;
            volatile REAL a = R_SQRT(a2);
            volatile REAL aa = a * a;
            volatile REAL a3 = aa * a;
            w[r][0] = i; w[r][1] = q; w[r][2] = a; w[r][3] = a2; w[r][4] = a3;
        }
        for (int ch = 0; ch < 3; ++ch)
            for (int r = 0; r < 4; ++r)
                for (int cc = 0; cc < 3; ++cc) {
                    REAL acc = bc[ch];
                    for (int dr = 0; dr < 3; ++dr) {
                        int rr = r + dr - 1;
                        if (rr < 0 || rr > 3) continue;
                        for (int dc = 0; dc < 3; ++dc) acc += Wc[ch * 9 + dr * 3 + dc] * w[rr][cc + dc];
                    }
                    z[ch * 12 + r * 3 + cc] = R_TANH(acc);
                }
        REAL o0 = bo[0], o1 = bo[1];
        for (int k = 0; k < H; ++k) {
            hk[k] = R_TANH(bh[k] + dotv(Wh + 36 * k, z, 36));
            o0 += Wo[k] * hk[k]; o1 += Wo[H + k] * hk[k];
        }
        if (!phase) { out[2 * t] = o0; out[2 * t + 1] = o1; continue; }
        const REAL g0 = gout[2 * t], g1 = gout[2 * t + 1];
        REAL dz[36] = {0}, dw[4][5] = {{0}};
        gp[30 + 37 * H + 2 * H] += g0; gp[30 + 37 * H + 2 * H + 1] += g1;
        for (int k = 0; k < H; ++k) {
            gp[30 + 37 * H + k] += g0 * hk[k]; gp[30 + 37 * H + H + k] += g1 * hk[k];
            const REAL da = (g0 * Wo[k] + g1 * Wo[H + k]) * ((REAL)1 - hk[k] * hk[k]);
            gp[30 + 36 * H + k] += da;
            for (int m = 0; m < 36; ++m) { gp[30 + 36 * k + m] += da * z[m]; dz[m] += da * Wh[36 * k + m]; }
        }
        for (int ch = 0; ch < 3; ++ch)
            for (int r = 0; r < 4; ++r)
                for (int cc = 0; cc < 3; ++cc) {
                    const int m = ch * 12 + r * 3 + cc;
                    const REAL d = dz[m] * ((REAL)1 - z[m] * z[m]);
                    gp[27 + ch] += d;
                    for (int dr = 0; dr < 3; ++dr) {
                        int rr = r + dr - 1;
                        if (rr < 0 || rr > 3) continue;
                        for (int dc = 0; dc < 3; ++dc) {
                            gp[ch * 9 + dr * 3 + dc] += d * w[rr][cc + dc];
                            dw[rr][cc + dc] += d * Wc[ch * 9 + dr * 3 + dc];
                        }
                    }
                }
        if (gx)
            for (int r = 0; r < 4; ++r) {
                const REAL i = w[r][0], q = w[r][1], a = w[r][2], a2 = w[r][3];
                const REAL ga = dw[r][2] + (REAL)3 * a2 * dw[r][4];
                const REAL sc = (REAL)2 * dw[r][3] + ga / a;
                gx[2 * sidx[r]] += dw[r][0] + i * sc;
                gx[2 * sidx[r] + 1] += dw[r][1] + q * sc;
            }
    }
}

/* ================================================================ BOJANET: bojanet.py:54-106
 * window t = samples t+m-15, m = 0..15, zero before the frame (:73-77); I_fir = fir_I(I)-fir_Q(Q), Q_fir = fir_Q(I)+fir_I(Q) (:80-83);
 * mag = sqrt(I_fir^2+Q_fir^2)+1e-8, sin = Q_fir/mag, cos = I_fir/mag (:30-39); L = [mag(6) | mag^2(6)] (:85-86);
 * f = sigm(W_fi L + b + W_fh h), g = tanh(W_gi L + b + W_gh h), h = f h + (1-f) g (:87-94); unit j rotated by filter j mod 6 (:41-52);
 * out_I = W_out_I(h cos) - W_out_Q(h sin), out_Q = W_out_Q(h sin) + W_out_I(h cos) (:98-102).
 * params: fir_I(6,16) fir_Q(6,16) W_fi(H,12) b_fi(H) W_fh(H,H) W_gi(H,12) b_gi(H) W_gh(H,H) W_out_I(1,H) b(1) W_out_Q(1,H) b(1). */
static void seq_bojanet(const Ctx *c, const REAL *x, const REAL *gout, REAL *out, REAL *gx, REAL *gp, int phase) {
    const int T = c->T, H = c->H;
    const REAL eps = (REAL)1e-8;
    const REAL *FI = c->params, *FQ = FI + 96, *Wfi = FQ + 96, *bfi = Wfi + 12 * H, *Wfh = bfi + H, *Wgi = Wfh + H * H, *bgi = Wgi + 12 * H,
               *Wgh = bgi + H, *WoI = Wgh + H * H, *boI = WoI + H, *WoQ = boI + 1, *boQ = WoQ + H;
    const size_t oWfi = 192, obfi = oWfi + 12 * H, oWfh = obfi + H, oWgi = oWfh + (size_t)H * H, obgi = oWgi + 12 * H, oWgh = obgi + H,
                 oWoI = oWgh + (size_t)H * H, oboI = oWoI + H, oWoQ = oboI + 1, oboQ = oWoQ + H;
    REAL *fir = (REAL *)malloc(sizeof(REAL) * (size_t)T * 12), *fr = (REAL *)malloc(sizeof(REAL) * (size_t)T * 24);
    REAL *act = (REAL *)malloc(sizeof(REAL) * (size_t)T * 3 * H);     /* f | g | h per step */
    REAL h[64] = {0};
    for (int t = 0; t < T; ++t) {
        for (int p = 0; p < 6; ++p) {
            REAL a = 0, b = 0, cI = 0, d = 0;     /* fir_I(I), fir_Q(Q), fir_Q(I), fir_I(Q) */
            for (int m = 0; m < 16; ++m) {
                int s = t + m - 15; if (s < 0) continue;
                a += FI[p * 16 + m] * x[2 * s]; b += FQ[p * 16 + m] * x[2 * s + 1];
                cI += FQ[p * 16 + m] * x[2 * s]; d += FI[p * 16 + m] * x[2 * s + 1];
            }
            const REAL fi = a - b, fq = cI + d;
            const REAL mag = R_SQRT(fi * fi + fq * fq) + eps;
            fir[12 * t + p] = fi; fir[12 * t + 6 + p] = fq;
            fr[24 * t + p] = mag; fr[24 * t + 6 + p] = mag * mag; fr[24 * t + 12 + p] = fq / mag; fr[24 * t + 18 + p] = fi / mag;
        }
        REAL hn[64];
        for (int j = 0; j < H; ++j) {
            const REAL f = sigm(bfi[j] + dotv(Wfi + 12 * j, fr + 24 * t, 12) + dotv(Wfh + H * j, h, H));
            const REAL g = R_TANH(bgi[j] + dotv(Wgi + 12 * j, fr + 24 * t, 12) + dotv(Wgh + H * j, h, H));
            hn[j] = f * h[j] + ((REAL)1 - f) * g;
            act[3 * H * t + j] = f; act[3 * H * t + H + j] = g; act[3 * H * t + 2 * H + j] = hn[j];
        }
        memcpy(h, hn, sizeof(REAL) * H);
        if (!phase) {
            REAL a = boI[0], q = boQ[0];
            for (int j = 0; j < H; ++j) { a += WoI[j] * (h[j] * fr[24 * t + 18 + j % 6]); q += WoQ[j] * (h[j] * fr[24 * t + 12 + j % 6]); }
            out[2 * t] = a - q; out[2 * t + 1] = q + a;
        }
    }
    if (phase) {
        REAL rec[64] = {0};
        REAL *dfir = (REAL *)calloc((size_t)T * 12, sizeof(REAL));
        for (int t = T - 1; t >= 0; --t) {
            const REAL da = gout[2 * t] + gout[2 * t + 1], dq = gout[2 * t + 1] - gout[2 * t];
            const REAL *ft = fr + 24 * t, *at = act + 3 * H * t;
            REAL dsn[6] = {0}, dcs[6] = {0}, dL[12] = {0}, af[64], ag[64], nrec[64] = {0};
            gp[oboI] += da; gp[oboQ] += dq;
            for (int j = 0; j < H; ++j) {
                const REAL hv = at[2 * H + j], cs = ft[18 + j % 6], sn = ft[12 + j % 6];
                gp[oWoI + j] += da * hv * cs; gp[oWoQ + j] += dq * hv * sn;
                const REAL dI = da * WoI[j], dQ = dq * WoQ[j];
                dcs[j % 6] += dI * hv; dsn[j % 6] += dQ * hv;
                const REAL dh = dI * cs + dQ * sn + rec[j];
                const REAL f = at[j], g = at[H + j], hp = t > 0 ? at[-3 * H + 2 * H + j] : 0;
                af[j] = dh * (hp - g) * f * ((REAL)1 - f);
                ag[j] = dh * ((REAL)1 - f) * ((REAL)1 - g * g);
                nrec[j] = dh * f;
            }
            for (int j = 0; j < H; ++j) {
                gp[obfi + j] += af[j]; gp[obgi + j] += ag[j];
                for (int k = 0; k < 12; ++k) { gp[oWfi + 12 * j + k] += af[j] * ft[k]; gp[oWgi + 12 * j + k] += ag[j] * ft[k]; dL[k] += af[j] * Wfi[12 * j + k] + ag[j] * Wgi[12 * j + k]; }
                for (int k = 0; k < H; ++k) {
                    const REAL hp = t > 0 ? at[-3 * H + 2 * H + k] : 0;
                    gp[oWfh + (size_t)H * j + k] += af[j] * hp; gp[oWgh + (size_t)H * j + k] += ag[j] * hp;
                    nrec[k] += af[j] * Wfh[H * j + k] + ag[j] * Wgh[H * j + k];
                }
            }
            memcpy(rec, nrec, sizeof(REAL) * H);
            for (int p = 0; p < 6; ++p) {
                const REAL mag = ft[p], sn = ft[12 + p], cs = ft[18 + p], fi = fir[12 * t + p], fq = fir[12 * t + 6 + p];
                const REAL dmag = dL[p] + (REAL)2 * mag * dL[6 + p] - (dsn[p] * sn + dcs[p] * cs) / mag;
                const REAL root = mag - eps;
                dfir[12 * t + p] = dcs[p] / mag + dmag * fi / root;
                dfir[12 * t + 6 + p] = dsn[p] / mag + dmag * fq / root;
            }
        }
        for (int t = 0; t < T; ++t)
            for (int p = 0; p < 6; ++p) {
                const REAL di = dfir[12 * t + p], dq = dfir[12 * t + 6 + p];
                for (int m = 0; m < 16; ++m) {
                    int s = t + m - 15; if (s < 0) continue;
                    gp[p * 16 + m] += di * x[2 * s] + dq * x[2 * s + 1];
                    gp[96 + p * 16 + m] += dq * x[2 * s] - di * x[2 * s + 1];
                    if (gx) { gx[2 * s] += di * FI[p * 16 + m] + dq * FQ[p * 16 + m]; gx[2 * s + 1] += dq * FI[p * 16 + m] - di * FQ[p * 16 + m]; }
                }
            }
        free(dfir);
    }
    free(fir); free(fr); free(act);
}

/* ================================================================ TCNN (tcnn.py:83-97) and NeuralTX (neuraltx.py:107-124)
 * stack: Conv1d(F->C,k=1,bias) hardswish; 4 x [depthwise Conv1d(k=5, dilation d=1,2,4,8, zero padding 2d, no bias) hardswish]; Conv1d(C->2,k=1).
 * TCNN: F=6 features (I,Q,a,a^3,sin,cos), out = stack + (I,Q).  NeuralTX (its fft over a length-1 axis is the identity): 5-tap FIR
 * I_f = conv_I(I)-conv_Q(Q), Q_f = conv_Q(I)+conv_I(Q) (zero 'same' padding), F=4 features (I_f,Q_f,a,a^3), out = stack + IQ_match(I_f,Q_f) + (I_f,Q_f).
 * hardswish'(v) = 0 (v<-3), v/3+1/2 (-3<=v<=3), 1 (v>3)  (ATen hardswish_backward). */
static inline REAL hswf(REAL v) { REAL r = v + (REAL)3; r = r < 0 ? 0 : (r > (REAL)6 ? (REAL)6 : r); return v * r / (REAL)6; }
static inline REAL hswg(REAL v) { return v < (REAL)-3 ? 0 : (v <= (REAL)3 ? v / (REAL)3 + (REAL)0.5 : (REAL)1); }
static void seq_tcn(const Ctx *c, const REAL *x, const REAL *gout, REAL *out, REAL *gx, REAL *gp, int phase) {
    const int T = c->T, C = c->H, ntx = c->cell == CELL_NEURALTX, F = ntx ? 4 : 6;
    const size_t oW0 = ntx ? 10 : 0, ob0 = oW0 + (size_t)C * F, oDw = ob0 + C, oW10 = oDw + (size_t)20 * C, oIQ = oW10 + (size_t)2 * C;
    const REAL *P = c->params, *W0 = P + oW0, *b0 = P + ob0, *W10 = P + oW10, *IQ = P + oIQ;
    REAL *feat = (REAL *)calloc((size_t)T * 8, sizeof(REAL)), *iq = (REAL *)calloc((size_t)T * 2, sizeof(REAL));
    REAL *pre = (REAL *)calloc((size_t)5 * C * T, sizeof(REAL));
#define PRE(l, ch, t) pre[((size_t)(l) * C + (ch)) * T + (t)]
    for (int t = 0; t < T; ++t) {
        if (!ntx) { features_fwd(CELL_DGRU, x, T, t, feat + 8 * t); iq[2 * t] = x[2 * t]; iq[2 * t + 1] = x[2 * t + 1]; }
        else {
            REAL a = 0, b = 0;
            for (int k = 0; k < 5; ++k) { int q = t + k - 2; if (q < 0 || q >= T) continue; a += P[k] * x[2 * q] - P[5 + k] * x[2 * q + 1]; b += P[5 + k] * x[2 * q] + P[k] * x[2 * q + 1]; }
            const REAL am = R_SQRT(a * a + b * b);
            feat[8 * t] = a; feat[8 * t + 1] = b; feat[8 * t + 2] = am; feat[8 * t + 3] = am * am * am; iq[2 * t] = a; iq[2 * t + 1] = b;
        }
    }
    for (int ch = 0; ch < C; ++ch)
        for (int t = 0; t < T; ++t) { REAL v = b0[ch]; for (int m = 0; m < F; ++m) v += W0[ch * F + m] * feat[8 * t + m]; PRE(0, ch, t) = v; }
    for (int l = 0; l < 4; ++l) {
        const int d = 1 << l; const REAL *w = P + oDw + (size_t)5 * C * l;
        for (int ch = 0; ch < C; ++ch)
            for (int t = 0; t < T; ++t) {
                REAL v = 0;
                for (int k = 0; k < 5; ++k) { int q = t + (k - 2) * d; if (q >= 0 && q < T) v += w[ch * 5 + k] * hswf(PRE(l, ch, q)); }
                PRE(l + 1, ch, t) = v;
            }
    }
    if (!phase) {
        for (int t = 0; t < T; ++t)
            for (int o = 0; o < 2; ++o) {
                REAL v = 0;
                for (int ch = 0; ch < C; ++ch) v += W10[o * C + ch] * hswf(PRE(4, ch, t));
                v += iq[2 * t + o];
                if (ntx) v += IQ[2 * o] * iq[2 * t] + IQ[2 * o + 1] * iq[2 * t + 1];
                out[2 * t + o] = v;
            }
    } else {
        REAL *ga = (REAL *)calloc((size_t)C * T, sizeof(REAL)), *gb = (REAL *)calloc((size_t)C * T, sizeof(REAL));
        REAL *diq = (REAL *)calloc((size_t)T * 2, sizeof(REAL));
        for (int ch = 0; ch < C; ++ch)
            for (int t = 0; t < T; ++t) {
                const REAL g0 = gout[2 * t], g1 = gout[2 * t + 1], a4 = hswf(PRE(4, ch, t));
                gp[oW10 + ch] += g0 * a4; gp[oW10 + C + ch] += g1 * a4;
                ga[(size_t)ch * T + t] = (g0 * W10[ch] + g1 * W10[C + ch]) * hswg(PRE(4, ch, t));
            }
        for (int l = 3; l >= 0; --l) {
            const int d = 1 << l; const REAL *w = P + oDw + (size_t)5 * C * l;
            memset(gb, 0, sizeof(REAL) * (size_t)C * T);
            for (int ch = 0; ch < C; ++ch)
                for (int t = 0; t < T; ++t) {
                    const REAL g = ga[(size_t)ch * T + t];
                    for (int k = 0; k < 5; ++k) {
                        int q = t + (k - 2) * d; if (q < 0 || q >= T) continue;
                        gp[oDw + (size_t)5 * C * l + ch * 5 + k] += g * hswf(PRE(l, ch, q));
                        gb[(size_t)ch * T + q] += g * w[ch * 5 + k];
                    }
                }
            for (int ch = 0; ch < C; ++ch)
                for (int t = 0; t < T; ++t) ga[(size_t)ch * T + t] = gb[(size_t)ch * T + t] * hswg(PRE(l, ch, t));
        }
        for (int t = 0; t < T; ++t) {
            REAL gf[8] = {0};
            const REAL g0 = gout[2 * t], g1 = gout[2 * t + 1];
            for (int ch = 0; ch < C; ++ch) {
                const REAL g = ga[(size_t)ch * T + t];
                gp[ob0 + ch] += g;
                for (int m = 0; m < F; ++m) { gp[oW0 + ch * F + m] += g * feat[8 * t + m]; gf[m] += g * W0[ch * F + m]; }
            }
            if (!ntx) {
                if (gx) { features_bwd(CELL_DGRU, x, T, t, gf, gx); gx[2 * t] += g0; gx[2 * t + 1] += g1; }
            } else {
                const REAL a = iq[2 * t], b = iq[2 * t + 1], am = feat[8 * t + 2];
                gp[oIQ] += g0 * a; gp[oIQ + 1] += g0 * b; gp[oIQ + 2] += g1 * a; gp[oIQ + 3] += g1 * b;
                const REAL gam = gf[2] + (REAL)3 * am * am * gf[3];
                diq[2 * t] = gf[0] + g0 + g0 * IQ[0] + g1 * IQ[2] + gam * a / am;
                diq[2 * t + 1] = gf[1] + g1 + g0 * IQ[1] + g1 * IQ[3] + gam * b / am;
            }
        }
        if (ntx)
            for (int t = 0; t < T; ++t)
                for (int k = 0; k < 5; ++k) {
                    int q = t + k - 2; if (q < 0 || q >= T) continue;
                    const REAL di = diq[2 * t], dq = diq[2 * t + 1];
                    gp[k] += di * x[2 * q] + dq * x[2 * q + 1];
                    gp[5 + k] += dq * x[2 * q] - di * x[2 * q + 1];
                    if (gx) { gx[2 * q] += di * P[k] + dq * P[5 + k]; gx[2 * q + 1] += dq * P[k] - di * P[5 + k]; }
                }
        free(ga); free(gb); free(diq);
    }
#undef PRE
    free(feat); free(iq); free(pre);
}

/* ================================================================ APNRRU: apnrru.py:52-135 (RRU cell :5-31)
 * 16-tap complex FIR with 3 real-weight filter pairs (windows zero before the frame, :67-71, :84-87) + the raw sample = 4 complex inputs;
 * r = conj(x_t)/|x_t| (:74-77) rotates the inputs (:93-95) and, every step, the complex state h_I + j h_Q (:101-102);
 * u = [inputs(8), h_I', h_Q', h_A(3)], hnew = [h_I', h_Q', h_A];  v = sigmoid(C hnew) + Z * tanh(W_h tanh(W_u u + b) + b) (:22-31);
 * (v[:H] + j v[H:2H]) is rotated back by conj(r) (:115-119), h_A = v[2H:];  out_I = o_I(h_I) - o_Q(h_Q), out_Q = o_Q(h_Q) + o_I(h_I) (:123-125).
 * params: fir_I(3,16) fir_Q(3,16) C(1) Z(1,S) W_u(16,S+8) b_u(16) W_h(S,16) b_h(S) o_I(1,H) o_Q(1,H),  S = 2H+3. */
static void seq_apnrru(const Ctx *c, const REAL *x, const REAL *gout, REAL *out, REAL *gx, REAL *gp, int phase) {
    const int T = c->T, H = c->H, S = 2 * H + 3, U = S + 8;
    const size_t oC = 96, oZ = 97, oWu = oZ + S, obu = oWu + (size_t)16 * U, oWh = obu + 16, obh = oWh + (size_t)16 * S, oI = obh + S, oQ = oI + H;
    const REAL *P = c->params, *FI = P, *FQ = P + 48, Cc = P[oC], *Z = P + oZ, *Wu = P + oWu, *bu = P + obu, *Wh = P + oWh, *bh = P + obh,
               *WI = P + oI, *WQ = P + oQ;
    /* per step: fir(6) rr ri mag | u(U) | v1(16) | v2(S) | sg(S) | vv(S) | hd(2H) */
    const int ST = 9 + U + 16 + 3 * S + 2 * H;
    REAL *sv = (REAL *)calloc((size_t)T * ST, sizeof(REAL));
    REAL hI[32] = {0}, hQ[32] = {0}, hA[3] = {0};
    for (int t = 0; t < T; ++t) {
        REAL *q = sv + (size_t)t * ST, *u = q + 9, *v1 = u + U, *v2 = v1 + 16, *sg = v2 + S, *vv = sg + S, *hd = vv + S;
        REAL a[4], b[4];
        for (int p = 0; p < 3; ++p) {
            REAL fi = 0, fq = 0;
            for (int m = 0; m < 16; ++m) {
                int s = t + m - 15; if (s < 0) continue;
                fi += FI[p * 16 + m] * x[2 * s] - FQ[p * 16 + m] * x[2 * s + 1];
                fq += FQ[p * 16 + m] * x[2 * s] + FI[p * 16 + m] * x[2 * s + 1];
            }
            a[p] = fi; b[p] = fq; q[p] = fi; q[3 + p] = fq;
        }
        a[3] = x[2 * t]; b[3] = x[2 * t + 1];
        const REAL mag = R_SQRT(x[2 * t] * x[2 * t] + x[2 * t + 1] * x[2 * t + 1]);
        const REAL rr = x[2 * t] / mag, ri = -x[2 * t + 1] / mag;
        q[6] = rr; q[7] = ri; q[8] = mag;
        for (int k = 0; k < 4; ++k) { u[2 * k] = rr * a[k] - ri * b[k]; u[2 * k + 1] = ri * a[k] + rr * b[k]; }
        for (int j = 0; j < H; ++j) { u[8 + j] = hI[j] * rr - hQ[j] * ri; u[8 + H + j] = hI[j] * ri + hQ[j] * rr; }
        for (int k = 0; k < 3; ++k) u[8 + 2 * H + k] = hA[k];
        for (int i = 0; i < 16; ++i) v1[i] = R_TANH(bu[i] + dotv(Wu + (size_t)i * U, u, U));
        for (int s = 0; s < S; ++s) {
            v2[s] = R_TANH(bh[s] + dotv(Wh + (size_t)s * 16, v1, 16));
            sg[s] = sigm(Cc * u[8 + s]);
            vv[s] = sg[s] + Z[s] * v2[s];
        }
        REAL oi = 0, oq = 0;
        for (int j = 0; j < H; ++j) {
            hI[j] = rr * vv[j] + ri * vv[H + j]; hQ[j] = rr * vv[H + j] - ri * vv[j];
            hd[j] = hI[j]; hd[H + j] = hQ[j];
            oi += WI[j] * hI[j]; oq += WQ[j] * hQ[j];
        }
        for (int k = 0; k < 3; ++k) hA[k] = vv[2 * H + k];
        if (!phase) { out[2 * t] = oi - oq; out[2 * t + 1] = oq + oi; }
    }
    if (phase) {
        REAL ghI[32] = {0}, ghQ[32] = {0}, ghA[3] = {0};
        REAL *dfir = (REAL *)calloc((size_t)T * 6, sizeof(REAL));
        for (int t = T - 1; t >= 0; --t) {
            const REAL *q = sv + (size_t)t * ST, *u = q + 9, *v1 = u + U, *v2 = v1 + 16, *sg = v2 + S, *vv = sg + S, *hd = vv + S;
            const REAL *hp = t > 0 ? q - ST + (ST - 2 * H) : NULL;      /* de-rotated state of step t-1 */
            const REAL rr = q[6], ri = q[7], mag = q[8];
            const REAL da = gout[2 * t] + gout[2 * t + 1], dq = gout[2 * t + 1] - gout[2 * t];
            REAL gv[80], ghn[80], ga2[80], gv1[16] = {0}, gu[96] = {0}, grr = 0, gri = 0;
            for (int j = 0; j < H; ++j) {
                gp[oI + j] += da * hd[j]; gp[oQ + j] += dq * hd[H + j];
                const REAL gI = ghI[j] + da * WI[j], gQ = ghQ[j] + dq * WQ[j];
                gv[j] = gI * rr - gQ * ri; gv[H + j] = gI * ri + gQ * rr;
                grr += gI * vv[j] + gQ * vv[H + j]; gri += gI * vv[H + j] - gQ * vv[j];
            }
            for (int k = 0; k < 3; ++k) gv[2 * H + k] = ghA[k];
            for (int s = 0; s < S; ++s) {
                const REAL ds = sg[s] * ((REAL)1 - sg[s]);
                gp[oC] += gv[s] * ds * u[8 + s];
                ghn[s] = gv[s] * ds * Cc;
                gp[oZ + s] += gv[s] * v2[s];
                ga2[s] = gv[s] * Z[s] * ((REAL)1 - v2[s] * v2[s]);
                gp[obh + s] += ga2[s];
                for (int i = 0; i < 16; ++i) { gp[oWh + (size_t)s * 16 + i] += ga2[s] * v1[i]; gv1[i] += ga2[s] * Wh[(size_t)s * 16 + i]; }
            }
            for (int i = 0; i < 16; ++i) {
                const REAL ga1 = gv1[i] * ((REAL)1 - v1[i] * v1[i]);
                gp[obu + i] += ga1;
                for (int k = 0; k < U; ++k) { gp[oWu + (size_t)i * U + k] += ga1 * u[k]; gu[k] += ga1 * Wu[(size_t)i * U + k]; }
            }
            for (int s = 0; s < S; ++s) ghn[s] += gu[8 + s];
            for (int j = 0; j < H; ++j) {
                const REAL gI = ghn[j], gQ = ghn[H + j], pI = hp ? hp[j] : 0, pQ = hp ? hp[H + j] : 0;
                ghI[j] = gI * rr + gQ * ri; ghQ[j] = -gI * ri + gQ * rr;
                grr += gI * pI + gQ * pQ; gri += -gI * pQ + gQ * pI;
            }
            for (int k = 0; k < 3; ++k) ghA[k] = ghn[2 * H + k];
            REAL gxi = 0, gxq = 0;
            for (int k = 0; k < 4; ++k) {
                const REAL gre = gu[2 * k], gim = gu[2 * k + 1];
                const REAL ak = k < 3 ? q[k] : x[2 * t], bk = k < 3 ? q[3 + k] : x[2 * t + 1];
                const REAL gak = gre * rr + gim * ri, gbk = -gre * ri + gim * rr;
                grr += gre * ak + gim * bk; gri += -gre * bk + gim * ak;
                if (k < 3) { dfir[6 * t + k] = gak; dfir[6 * t + 3 + k] = gbk; } else { gxi += gak; gxq += gbk; }
            }
            const REAL I = x[2 * t], Q = x[2 * t + 1], m3 = mag * mag * mag;
            gxi += grr * ((REAL)1 / mag - I * I / m3) + gri * (Q * I / m3);
            gxq += grr * (-I * Q / m3) + gri * ((REAL)-1 / mag + Q * Q / m3);
            if (gx) { gx[2 * t] += gxi; gx[2 * t + 1] += gxq; }
        }
        for (int t = 0; t < T; ++t)
            for (int p = 0; p < 3; ++p) {
                const REAL di = dfir[6 * t + p], dq = dfir[6 * t + 3 + p];
                for (int m = 0; m < 16; ++m) {
                    int s = t + m - 15; if (s < 0) continue;
                    gp[p * 16 + m] += di * x[2 * s] + dq * x[2 * s + 1];
                    gp[48 + p * 16 + m] += dq * x[2 * s] - di * x[2 * s + 1];
                    if (gx) { gx[2 * s] += di * FI[p * 16 + m] + dq * FQ[p * 16 + m]; gx[2 * s + 1] += dq * FI[p * 16 + m] - di * FQ[p * 16 + m]; }
                }
            }
        free(dfir);
    }
    free(sv);
}

/* ================================================================ MCLDNN: mcldnn.py:83-113
 * features (I,Q,a,a^2,a^3) (:88-93); window of timestep t = samples (t-4+m) mod T, m = 0..4 (`pad = x[:, -(memory_length-1):, :]`, :96-99), laid out
 * as a 5x5 image (feature, memory); conv2d_1 (1->C, 3x3, pad 1) (:101); conv1d over memory with the features as 5 groups (5 -> 5C, k=3, pad 1)
 * re-viewed as (C,5,5): channel oc lands at [oc / 5][oc % 5] (:102-103); concatenated along the height (:104), transposed and convolved by
 * conv2d_2 (10 -> 1, 3x3, pad 1) over (C,5) (:106); LSTM(5C -> 8) (:108); fc_out(8->16), fc_out_2(16->2), no activation in between (:109-110).
 * params: conv2d_1.w(C,1,3,3) .b(C) conv1d.w(5C,1,3) .b(5C) conv2d_2.w(1,10,3,3) .b(1) lstm.w_ih(32,5C) w_hh(32,8) b_ih(32) b_hh(32)
 *         fc_out.w(16,8) .b(16) fc_out_2.w(2,16) .b(2).   H = C (conv channels). */
static void mcl_window(const REAL *x, int T, int t, REAL w[5][5], int sidx[5]) {
    for (int m = 0; m < 5; ++m) {
        int s = ((t - 4 + m) % T + T) % T;
        sidx[m] = s;
        volatile REAL i = x[2 * s], q = x[2 * s + 1];
        volatile REAL ii = i * i, qq = q * q;
        volatile REAL a2 = ii + qq;
        volatile REAL a = R_SQRT(a2);
        volatile REAL aa = a * a;
        volatile REAL a3 = aa * a;
        w[0][m] = i; w[1][m] = q; w[2][m] = a; w[3][m] = a2; w[4][m] = a3;
    }
}
#define MZ(c, r, m) Zb[((c) * 10 + (r)) * 5 + (m)]
static void mcl_front(const REAL *P, int C, REAL w[5][5], REAL *Zb, REAL *O) {
    const REAL *W1 = P, *b1 = W1 + 9 * C, *Wc = b1 + C, *bc = Wc + 15 * C, *W2 = bc + 5 * C, *b2 = W2 + 90;
    for (int c = 0; c < C; ++c)
        for (int f = 0; f < 5; ++f)
            for (int m = 0; m < 5; ++m) {
                REAL v = b1[c];
                for (int df = 0; df < 3; ++df)
                    for (int dm = 0; dm < 3; ++dm) {
                        int ff = f + df - 1, mm = m + dm - 1;
                        if (ff < 0 || ff > 4 || mm < 0 || mm > 4) continue;
                        v += W1[c * 9 + df * 3 + dm] * w[ff][mm];
                    }
                MZ(c, f, m) = v;
            }
    for (int oc = 0; oc < 5 * C; ++oc)
        for (int m = 0; m < 5; ++m) {
            REAL v = bc[oc];
            for (int dm = 0; dm < 3; ++dm) { int mm = m + dm - 1; if (mm < 0 || mm > 4) continue; v += Wc[oc * 3 + dm] * w[oc / C][mm]; }
            MZ(oc / 5, 5 + oc % 5, m) = v;
        }
    for (int c = 0; c < C; ++c)
        for (int m = 0; m < 5; ++m) {
            REAL v = b2[0];
            for (int r = 0; r < 10; ++r)
                for (int dc = 0; dc < 3; ++dc)
                    for (int dm = 0; dm < 3; ++dm) {
                        int cc = c + dc - 1, mm = m + dm - 1;
                        if (cc < 0 || cc >= C || mm < 0 || mm > 4) continue;
                        v += W2[r * 9 + dc * 3 + dm] * MZ(cc, r, mm);
                    }
            O[c * 5 + m] = v;
        }
}
static void seq_mcldnn(const Ctx *c, const REAL *x, const REAL *gout, REAL *out, REAL *gx, REAL *gp, int phase) {
    const int T = c->T, C = c->H, IN = 5 * C;
    const size_t oW1 = 0, ob1 = 9 * C, oWc = ob1 + C, obc = oWc + 15 * C, oW2 = obc + 5 * C, ob2 = oW2 + 90, oWih = ob2 + 1, oWhh = oWih + (size_t)32 * IN,
                 obih = oWhh + 256, obhh = obih + 32, oF1 = obhh + 32, oF1b = oF1 + 128, oF2 = oF1b + 16, oF2b = oF2 + 32;
    const REAL *P = c->params, *Wih = P + oWih, *Whh = P + oWhh, *bih = P + obih, *bhh = P + obhh, *F1 = P + oF1, *F1b = P + oF1b, *F2 = P + oF2,
               *F2b = P + oF2b;
    REAL *Zb = (REAL *)malloc(sizeof(REAL) * (size_t)C * 50), *inp = (REAL *)malloc(sizeof(REAL) * (size_t)T * IN);
    REAL *act = (REAL *)malloc(sizeof(REAL) * (size_t)T * 48);      /* i f g o c h per step */
    REAL h[8] = {0}, cs[8] = {0};
    for (int t = 0; t < T; ++t) {
        REAL w[5][5]; int sidx[5];
        mcl_window(x, T, t, w, sidx);
        mcl_front(P, C, w, Zb, inp + (size_t)t * IN);
        REAL g4[32];
        for (int r = 0; r < 32; ++r) g4[r] = bih[r] + bhh[r] + dotv(Wih + (size_t)r * IN, inp + (size_t)t * IN, IN) + dotv(Whh + r * 8, h, 8);
        REAL *a = act + 48 * t;
        for (int j = 0; j < 8; ++j) {
            const REAL ig = sigm(g4[j]), fg = sigm(g4[8 + j]), gg = R_TANH(g4[16 + j]), og = sigm(g4[24 + j]);
            cs[j] = fg * cs[j] + ig * gg;
            a[j] = ig; a[8 + j] = fg; a[16 + j] = gg; a[24 + j] = og; a[32 + j] = cs[j];
        }
        for (int j = 0; j < 8; ++j) { h[j] = a[24 + j] * R_TANH(cs[j]); a[40 + j] = h[j]; }
        if (!phase) {
            REAL y1[16];
            for (int k = 0; k < 16; ++k) y1[k] = F1b[k] + dotv(F1 + 8 * k, h, 8);
            out[2 * t] = F2b[0] + dotv(F2, y1, 16); out[2 * t + 1] = F2b[1] + dotv(F2 + 16, y1, 16);
        }
    }
    if (phase) {
        REAL dh[8] = {0}, dc[8] = {0};
        REAL *dZ = (REAL *)malloc(sizeof(REAL) * (size_t)C * 50), *din = (REAL *)malloc(sizeof(REAL) * IN);
        const REAL *W1 = P + oW1, *Wc = P + oWc, *W2 = P + oW2;
        for (int t = T - 1; t >= 0; --t) {
            const REAL *a = act + 48 * t, *ap = t > 0 ? a - 48 : NULL;
            const REAL g0 = gout[2 * t], g1 = gout[2 * t + 1];
            REAL y1[16], dy1[16], ga[32];
            for (int k = 0; k < 16; ++k) { y1[k] = F1b[k] + dotv(F1 + 8 * k, a + 40, 8); dy1[k] = g0 * F2[k] + g1 * F2[16 + k]; }
            gp[oF2b] += g0; gp[oF2b + 1] += g1;
            for (int k = 0; k < 16; ++k) {
                gp[oF2 + k] += g0 * y1[k]; gp[oF2 + 16 + k] += g1 * y1[k]; gp[oF1b + k] += dy1[k];
                for (int j = 0; j < 8; ++j) { gp[oF1 + 8 * k + j] += dy1[k] * a[40 + j]; dh[j] += dy1[k] * F1[8 * k + j]; }
            }
            for (int j = 0; j < 8; ++j) {
                const REAL ig = a[j], fg = a[8 + j], gg = a[16 + j], og = a[24 + j], tc = R_TANH(a[32 + j]), cp = ap ? ap[32 + j] : 0;
                const REAL dcc = dc[j] + dh[j] * og * ((REAL)1 - tc * tc);
                ga[j] = dcc * gg * ig * ((REAL)1 - ig);
                ga[8 + j] = dcc * cp * fg * ((REAL)1 - fg);
                ga[16 + j] = dcc * ig * ((REAL)1 - gg * gg);
                ga[24 + j] = dh[j] * tc * og * ((REAL)1 - og);
                dc[j] = dcc * fg;
            }
            for (int j = 0; j < 8; ++j) dh[j] = 0;
            for (int k = 0; k < IN; ++k) din[k] = 0;
            for (int r = 0; r < 32; ++r) {
                gp[obih + r] += ga[r]; gp[obhh + r] += ga[r];
                for (int k = 0; k < IN; ++k) { gp[oWih + (size_t)r * IN + k] += ga[r] * inp[(size_t)t * IN + k]; din[k] += ga[r] * Wih[(size_t)r * IN + k]; }
                for (int j = 0; j < 8; ++j) { gp[oWhh + r * 8 + j] += ga[r] * (ap ? ap[40 + j] : 0); dh[j] += ga[r] * Whh[r * 8 + j]; }
            }
            /* front backward: recompute Z of this timestep */
            REAL w[5][5], dw[5][5] = {{0}}, Otmp[64]; int sidx[5];
            mcl_window(x, T, t, w, sidx);
            mcl_front(P, C, w, Zb, Otmp);
            memset(dZ, 0, sizeof(REAL) * (size_t)C * 50);
            for (int cc = 0; cc < C; ++cc)
                for (int m = 0; m < 5; ++m) {
                    const REAL d = din[cc * 5 + m];
                    gp[ob2] += d;
                    for (int r = 0; r < 10; ++r)
                        for (int dcx = 0; dcx < 3; ++dcx)
                            for (int dm = 0; dm < 3; ++dm) {
                                int c2 = cc + dcx - 1, mm = m + dm - 1;
                                if (c2 < 0 || c2 >= C || mm < 0 || mm > 4) continue;
                                gp[oW2 + r * 9 + dcx * 3 + dm] += d * MZ(c2, r, mm);
                                dZ[(c2 * 10 + r) * 5 + mm] += d * W2[r * 9 + dcx * 3 + dm];
                            }
                }
            for (int cc = 0; cc < C; ++cc)
                for (int f = 0; f < 5; ++f)
                    for (int m = 0; m < 5; ++m) {
                        const REAL d = dZ[(cc * 10 + f) * 5 + m];
                        gp[ob1 + cc] += d;
                        for (int df = 0; df < 3; ++df)
                            for (int dm = 0; dm < 3; ++dm) {
                                int ff = f + df - 1, mm = m + dm - 1;
                                if (ff < 0 || ff > 4 || mm < 0 || mm > 4) continue;
                                gp[oW1 + cc * 9 + df * 3 + dm] += d * w[ff][mm];
                                dw[ff][mm] += d * W1[cc * 9 + df * 3 + dm];
                            }
                    }
            for (int oc = 0; oc < 5 * C; ++oc)
                for (int m = 0; m < 5; ++m) {
                    const REAL d = dZ[((oc / 5) * 10 + 5 + oc % 5) * 5 + m];
                    gp[obc + oc] += d;
                    for (int dm = 0; dm < 3; ++dm) {
                        int mm = m + dm - 1; if (mm < 0 || mm > 4) continue;
                        gp[oWc + oc * 3 + dm] += d * w[oc / C][mm];
                        dw[oc / C][mm] += d * Wc[oc * 3 + dm];
                    }
                }
            if (gx)
                for (int m = 0; m < 5; ++m) {
                    const REAL i = w[0][m], q = w[1][m], am = w[2][m], a2 = w[3][m];
                    const REAL gam = dw[2][m] + (REAL)3 * a2 * dw[4][m];
                    const REAL sc = (REAL)2 * dw[3][m] + gam / am;
                    gx[2 * sidx[m]] += dw[0][m] + i * sc; gx[2 * sidx[m] + 1] += dw[1][m] + q * sc;
                }
        }
        free(dZ); free(din);
    }
    free(Zb); free(inp); free(act);
}
#undef MZ

/* ================================================================ DeltaJANET: deltajanet.py:49-60 (features), :203-262 (layer)
 * The outer module builds its layer with thx = thh = 0 (:22-26), so every delta passes its threshold and the states track every step:
 * dx_t = f_t - f_{t-1}, dh_t = h_{t-1} - h_{t-2} (zero states before the frame); M += W_ih dx + W_hh dh, M_0 = b_ih + b_hh (:160-167);
 * f = sigm(M_f), g = sigm(M_g) (both sigmoids, :246-247), h = (1-f) g + f h (:248); out = fc_out(h).
 * The accumulator is summed in the reference's order: (W_ih dx + M) + W_hh dh (:196-202).
 * params: rnn.weight_ih_l0(2H,6) weight_hh_l0(2H,H) bias_ih_l0(2H) bias_hh_l0(2H) fc_out.weight(2,H) fc_out.bias(2). */
static void seq_deltajanet(const Ctx *c, const REAL *x, const REAL *gout, REAL *out, REAL *gx, REAL *gp, int phase) {
    const int T = c->T, H = c->H, G = 2 * H;
    const size_t oWih = 0, oWhh = (size_t)G * 6, obih = oWhh + (size_t)G * H, obhh = obih + G, oWo = obhh + G, obo = oWo + 2 * H;
    const REAL *P = c->params, *Wih = P, *Whh = P + oWhh, *bih = P + obih, *bhh = P + obhh, *Wo = P + oWo, *bo = P + obo;
    REAL *ft = (REAL *)calloc((size_t)(T + 1) * 8, sizeof(REAL));          /* ft[t+1] = features of step t, ft[0] = 0 */
    REAL *act = (REAL *)calloc((size_t)(T + 2) * 3 * H, sizeof(REAL));     /* rows t+2: f | g | h of step t; rows 0,1 = zero state */
    REAL M[64];
    for (int r = 0; r < G; ++r) M[r] = bih[r] + bhh[r];
    for (int t = 0; t < T; ++t) {
        features_fwd(CELL_DGRU, x, T, t, ft + 8 * (t + 1));
        REAL dx[6], dh[32];
        for (int k = 0; k < 6; ++k) dx[k] = ft[8 * (t + 1) + k] - ft[8 * t + k];
        const REAL *h1 = act + (size_t)(t + 1) * 3 * H + 2 * H, *h2 = act + (size_t)t * 3 * H + 2 * H;
        for (int k = 0; k < H; ++k) dh[k] = h1[k] - h2[k];
        for (int r = 0; r < G; ++r) { REAL mx = dotv(Wih + 6 * r, dx, 6) + M[r]; M[r] = mx + dotv(Whh + (size_t)H * r, dh, H); }
        REAL *a = act + (size_t)(t + 2) * 3 * H;
        for (int j = 0; j < H; ++j) {
            const REAL f = sigm(M[j]), g = sigm(M[H + j]);
            a[j] = f; a[H + j] = g; a[2 * H + j] = ((REAL)1 - f) * g + f * h1[j];
        }
        if (!phase) { out[2 * t] = bo[0] + dotv(Wo, a + 2 * H, H); out[2 * t + 1] = bo[1] + dotv(Wo + H, a + 2 * H, H); }
    }
    if (phase) {
        REAL gM[64] = {0}, gH[32] = {0}, pend[32] = {0};
        REAL *GMs = (REAL *)calloc((size_t)(T + 1) * G, sizeof(REAL));      /* running adjoint of M at step t (row T = 0) */
        for (int t = T - 1; t >= 0; --t) {
            const REAL *a = act + (size_t)(t + 2) * 3 * H, *h1 = a - 3 * H + 2 * H, *h2 = a - 6 * H + 2 * H;
            const REAL g0 = gout[2 * t], g1 = gout[2 * t + 1];
            gp[obo] += g0; gp[obo + 1] += g1;
            REAL at[32] = {0}, nH[32];
            for (int j = 0; j < H; ++j) {
                gp[oWo + j] += g0 * a[2 * H + j]; gp[oWo + H + j] += g1 * a[2 * H + j];
                const REAL gh = gH[j] + g0 * Wo[j] + g1 * Wo[H + j];
                const REAL f = a[j], g = a[H + j];
                gM[j] += gh * (h1[j] - g) * f * ((REAL)1 - f);
                gM[H + j] += gh * ((REAL)1 - f) * g * ((REAL)1 - g);
                nH[j] = gh * f;
            }
            for (int r = 0; r < G; ++r) {
                GMs[(size_t)t * G + r] = gM[r];
                for (int k = 0; k < 6; ++k) gp[oWih + 6 * r + k] += gM[r] * (ft[8 * (t + 1) + k] - ft[8 * t + k]);
                for (int k = 0; k < H; ++k) { gp[oWhh + (size_t)H * r + k] += gM[r] * (h1[k] - h2[k]); at[k] += gM[r] * Whh[(size_t)H * r + k]; }
            }
            for (int k = 0; k < H; ++k) { gH[k] = nH[k] + at[k] + pend[k]; pend[k] = -at[k]; }
        }
        for (int r = 0; r < G; ++r) { gp[obih + r] += gM[r]; gp[obhh + r] += gM[r]; }
        if (gx)
            for (int t = 0; t < T; ++t) {
                REAL gf[8] = {0};
                for (int r = 0; r < G; ++r) {
                    const REAL d = GMs[(size_t)t * G + r] - GMs[(size_t)(t + 1) * G + r];
                    for (int k = 0; k < 6; ++k) gf[k] += d * Wih[6 * r + k];
                }
                features_bwd(CELL_DGRU, x, T, t, gf, gx);
            }
        free(GMs);
    }
    free(ft); free(act);
}

/* ================================================================ fake-quantised TRes-DeltaGRU (the W16A16 stage of bash_scripts/OpenDPDv2.sh:47-49)
 * quant/quant_envs.py:286-305 applied to backbones/deltagru_tcnskip.py: x2h / h2h / fc_out become INT_Linear (weight and input each fake-quantised,
 * quant_layers.py:70-82; 16-bit output quantiser on fc_out in eval only), the layer's add / mul / sigmoid / tanh modules become Quant_add / Quant_mult /
 * Quant_sigmoid / Quant_tanh (quant_ops.py:14-66).  The delta logic itself stays in float (deltagru_tcnskip.py:266-291):
 *   dx, dh thresholded as in the float cell (masks, counters on the UNquantised deltas);  mac_x = x2h(Q(dx)) + M, mac_h = h2h(Q(dh));  M and M_nh accumulate
 *   as before;  r = Qs(sigm(M_r)), z = Qs(sigm(M_z));  n = Qt(tanh(Qa(M_n + Qm(r M_nh))));  h = Qa(Qm(Qa(1 - z) n) + Qm(z h));  out = fc_out(Q(h)) + tcn(x).
 * Quantiser and STE as in seq_qgru_qat.  params (named_parameters order): x2h.W(3H,6) + 3 scales | h2h.W(3H,H) + 3 scales | add, mul, sigmoid, tanh scales |
 * fc_out.W(2,H) + 3 scales | tcn.0.W(3,2,3) tcn.2.W(2,3,1).   K packs bits_w | bits_a<<8 | eval<<16. */
static void seq_tres_qat(const Ctx *c, const REAL *x, const REAL *gout, REAL *out, REAL *gx, REAL *gp, int phase,
                         uint64_t *mask_x, uint64_t *mask_h, int64_t *stats) {
    const int H = c->H, T = c->T, F = 6, bw = c->K & 255, ba = (c->K >> 8) & 255, eval = (c->K >> 16) & 1;
    const REAL *p = c->params;
    const REAL *Wx = p, *sx = Wx + 3 * H * F, *Wh = sx + 3, *sh = Wh + 3 * H * H, *sop = sh + 3, *Wo = sop + 4, *so = Wo + 2 * H, *w0 = so + 3, *w2 = w0 + 18;
    const Quant qxw = mkq(sx[0], bw), qxa = mkq(sx[1], ba), qhw = mkq(sh[0], bw), qha = mkq(sh[1], ba), qadd = mkq(sop[0], ba), qmul = mkq(sop[1], ba),
                qsig = mkq(sop[2], ba), qtanh = mkq(sop[3], ba), qow = mkq(so[0], bw), qoa = mkq(so[1], ba), qoo = mkq(so[2], 16);
    /* saved per step: f(6) qdx(6) cdx(6) | per unit 16 values: qdh cdh r z n sr sz tn mn omz hs hq flags(packed) - - - | cv(5) */
    const int S = 18 + 16 * H + 5;
    static __thread REAL *sv = NULL; static __thread size_t sv_n = 0;
    static __thread uint64_t *mk = NULL; static __thread size_t mk_n = 0;
    if (sv_n < (size_t)T * S) { free(sv); sv = (REAL *)malloc((size_t)T * S * sizeof(REAL)); sv_n = (size_t)T * S; }
    if (mk_n < (size_t)2 * T) { free(mk); mk = (uint64_t *)malloc(sizeof(uint64_t) * 2 * T); mk_n = 2 * T; }
    REAL Wxq[192 * 6], Whq[192 * 64], Woq[128]; int cWx[192 * 6], cWh[192 * 64], cWo[128];
    for (int i = 0; i < 3 * H * F; ++i) Wxq[i] = qf(&qxw, Wx[i], &cWx[i]);
    for (int i = 0; i < 3 * H * H; ++i) Whq[i] = qf(&qhw, Wh[i], &cWh[i]);
    for (int i = 0; i < 2 * H; ++i) Woq[i] = qf(&qow, Wo[i], &cWo[i]);
    if (phase == 0) {
        REAL h[64] = {0}, hp[64] = {0}, xp[6] = {0}, M[192] = {0}, Mnh[64] = {0};
        int64_t zx = 0, zh = 0;
        for (int t = 0; t < T; ++t) {
            REAL *s = sv + (size_t)t * S, *f = s, *qdx = s + 6, *cdx = s + 12, *u = s + 18, *cv = u + 16 * H;
            features_fwd(CELL_TRES, x, T, t, f);
            uint64_t mx = 0, mh = 0;
            for (int k = 0; k < F; ++k) {
                REAL d = f[k] - xp[k], a = R_FABS(d);
                if (a < c->thx) d = 0;
                if (a >= c->thx) { xp[k] = f[k]; mx |= (uint64_t)1 << k; }
                zx += (d == 0);
                int in; qdx[k] = qf(&qxa, d, &in); cdx[k] = (REAL)in;
            }
            REAL qdh[64];
            for (int j = 0; j < H; ++j) {
                REAL d = h[j] - hp[j], a = R_FABS(d);
                if (a < c->thh) d = 0;
                if (a >= c->thh) { hp[j] = h[j]; mh |= (uint64_t)1 << j; }
                zh += (d == 0);
                int in; qdh[j] = qf(&qha, d, &in); u[16 * j] = qdh[j]; u[16 * j + 1] = (REAL)in;
            }
            mk[2 * t] = mx; mk[2 * t + 1] = mh;
            if (mask_x) mask_x[t] = mx;
            if (mask_h) mask_h[t] = mh;
            REAL hn[64], o0 = 0, o1 = 0;
            for (int j = 0; j < H; ++j) {
                REAL *uj = u + 16 * j;
                const REAL mxr = dotv(Wxq + (size_t)j * F, qdx, F) + M[j], mxz = dotv(Wxq + (size_t)(H + j) * F, qdx, F) + M[H + j],
                           mxn = dotv(Wxq + (size_t)(2 * H + j) * F, qdx, F) + M[2 * H + j];
                M[j] = mxr + dotv(Whq + (size_t)j * H, qdh, H);
                M[H + j] = mxz + dotv(Whq + (size_t)(H + j) * H, qdh, H);
                M[2 * H + j] = mxn;
                Mnh[j] = dotv(Whq + (size_t)(2 * H + j) * H, qdh, H) + Mnh[j];
                int c_r, c_z, c_m1, c_an, c_n, c_omz, c_m3, c_m2, c_h, c_oa;
                const REAL sr = sigm(M[j]), sz = sigm(M[H + j]);
                const REAL r = qf(&qsig, sr, &c_r), z = qf(&qsig, sz, &c_z);
                const REAL m1 = qf(&qmul, r * Mnh[j], &c_m1);
                const REAL an = qf(&qadd, M[2 * H + j] + m1, &c_an);
                const REAL tn = R_TANH(an), n = qf(&qtanh, tn, &c_n);
                const REAL omz = qf(&qadd, (REAL)1 + (-z), &c_omz);
                const REAL m3 = qf(&qmul, omz * n, &c_m3), m2 = qf(&qmul, z * h[j], &c_m2);
                hn[j] = qf(&qadd, m3 + m2, &c_h);
                const REAL hq = qf(&qoa, hn[j], &c_oa);
                uj[2] = r; uj[3] = z; uj[4] = n; uj[5] = sr; uj[6] = sz; uj[7] = tn; uj[8] = Mnh[j]; uj[9] = omz; uj[10] = hn[j]; uj[11] = hq;
                uj[12] = (REAL)(c_r | c_z << 1 | c_m1 << 2 | c_an << 3 | c_n << 4 | c_omz << 5 | c_m3 << 6 | c_m2 << 7 | c_h << 8 | c_oa << 9);
                o0 += Woq[j] * hq; o1 += Woq[H + j] * hq;
            }
            memcpy(h, hn, sizeof(REAL) * H);
            if (eval) { o0 = qf(&qoo, o0, NULL); o1 = qf(&qoo, o1, NULL); }
            for (int co = 0; co < 3; ++co) {
                REAL a = 0;
                for (int ci = 0; ci < 2; ++ci)
                    for (int k = 0; k < 3; ++k) { int tt = t + (k - 1) * 16; if (tt >= 0 && tt < T) a += w0[(co * 2 + ci) * 3 + k] * x[2 * tt + ci]; }
                cv[co] = a;
            }
            for (int o = 0; o < 2; ++o) { REAL a = 0; for (int ch = 0; ch < 3; ++ch) a += w2[o * 3 + ch] * hardswish(cv[ch]); cv[3 + o] = a; }
            out[2 * t] = o0 + hardswish(cv[3]); out[2 * t + 1] = o1 + hardswish(cv[4]);
        }
        if (stats) { stats[0] = zx; stats[1] = (int64_t)T * F; stats[2] = zh; stats[3] = (int64_t)T * H; }
        return;
    }
    REAL *gWx = gp, *gWh = gWx + 3 * H * F + 3, *gWo = gWh + 3 * H * H + 3 + 4, *gw0 = gWo + 2 * H + 3, *gw2 = gw0 + 18;
    REAL gH[64] = {0}, gM[192] = {0}, gMnh[64] = {0}, gxp[6] = {0}, ghp[64] = {0};
    for (int t = T - 1; t >= 0; --t) {
        REAL *s = sv + (size_t)t * S, *f = s, *qdx = s + 6, *cdx = s + 12, *u = s + 18, *cv = u + 16 * H;
        const REAL *up = t ? sv + (size_t)(t - 1) * S + 18 : NULL;
        (void)f;
        const REAL go0 = gout[2 * t], go1 = gout[2 * t + 1];
        {
            REAL g2[2] = {go0 * hardswish_grad(cv[3]), go1 * hardswish_grad(cv[4])}, ga1[3] = {0, 0, 0};
            for (int o = 0; o < 2; ++o)
                for (int ch = 0; ch < 3; ++ch) { gw2[o * 3 + ch] += g2[o] * hardswish(cv[ch]); ga1[ch] += w2[o * 3 + ch] * g2[o]; }
            for (int co = 0; co < 3; ++co) {
                const REAL gc1 = ga1[co] * hardswish_grad(cv[co]);
                for (int ci = 0; ci < 2; ++ci)
                    for (int k = 0; k < 3; ++k) {
                        int tt = t + (k - 1) * 16;
                        if (tt >= 0 && tt < T) { gw0[(co * 2 + ci) * 3 + k] += gc1 * x[2 * tt + ci]; if (gx) gx[2 * tt + ci] += w0[(co * 2 + ci) * 3 + k] * gc1; }
                    }
            }
        }
        REAL ghprev[64], qdh[64];
        for (int j = 0; j < H; ++j) {
            const REAL *uj = u + 16 * j;
            qdh[j] = uj[0];
            const int fl = (int)uj[12];
            const REAL r = uj[2], z = uj[3], n = uj[4], sr = uj[5], sz = uj[6], tn = uj[7], mn = uj[8], omz = uj[9], hq = uj[11];
            const REAL hpj = up ? up[16 * j + 10] : 0;
            if (cWo[j]) gWo[j] += go0 * hq;
            if (cWo[H + j]) gWo[H + j] += go1 * hq;
            REAL g = gH[j] + (((fl >> 9) & 1) ? Woq[j] * go0 + Woq[H + j] * go1 : 0);
            g = ((fl >> 8) & 1) ? g : 0;                                   /* h = Qa(m3 + m2) */
            const REAL gm2 = ((fl >> 7) & 1) ? g : 0, gm3 = ((fl >> 6) & 1) ? g : 0;
            REAL gz = gm2 * hpj;
            ghprev[j] = gm2 * z;
            const REAL gomz = ((fl >> 5) & 1) ? gm3 * n : 0;
            gz -= gomz;
            const REAL gn = ((fl >> 4) & 1) ? gm3 * omz : 0;
            const REAL gan = ((fl >> 3) & 1) ? gn * ((REAL)1 - tn * tn) : 0;
            const REAL gm1 = ((fl >> 2) & 1) ? gan : 0;
            const REAL gr = (fl & 1) ? gm1 * mn : 0;
            gM[j] += gr * sr * ((REAL)1 - sr);
            gM[H + j] += (((fl >> 1) & 1) ? gz : 0) * sz * ((REAL)1 - sz);
            gM[2 * H + j] += gan;
            gMnh[j] += gm1 * r;
        }
        REAL gdx[6] = {0}, gdh[64] = {0};
        for (int k = 0; k < 3 * H; ++k) {
            const REAL gk = gM[k], gk_h = (k < 2 * H) ? gM[k] : gMnh[k - 2 * H];
            for (int q = 0; q < F; ++q) { if (cWx[k * F + q]) gWx[k * F + q] += gk * qdx[q]; gdx[q] += Wxq[k * F + q] * gk; }
            for (int q = 0; q < H; ++q) { if (cWh[k * H + q]) gWh[k * H + q] += gk_h * qdh[q]; gdh[q] += Whq[k * H + q] * gk_h; }
        }
        for (int q = 0; q < F; ++q) if (cdx[q] == 0) gdx[q] = 0;
        for (int q = 0; q < H; ++q) if (u[16 * q + 1] == 0) gdh[q] = 0;
        const uint64_t mx = mk[2 * t], mh = mk[2 * t + 1];
        REAL gf[6];
        for (int k = 0; k < F; ++k) { if ((mx >> k) & 1) { gf[k] = gxp[k] + gdx[k]; gxp[k] = -gdx[k]; } else gf[k] = 0; }
        for (int j = 0; j < H; ++j) if ((mh >> j) & 1) { ghprev[j] += ghp[j] + gdh[j]; ghp[j] = -gdh[j]; }
        memcpy(gH, ghprev, sizeof(REAL) * H);
        if (gx) features_bwd(CELL_TRES, x, T, t, gf, gx);
    }
}

static size_t n_params(int cell, int H, int K) {
    switch (cell) {
    case CELL_GRU: return (size_t)3 * H * 2 + 3 * H * H + 6 * H + 2 * H + 2;
    case CELL_QGRU: case CELL_QGRU_AMP1: return (size_t)3 * H * 4 + 3 * H * H + 6 * H + 2 * H + 2;
    case CELL_DGRU: return (size_t)3 * H * 6 + 3 * H * H + 6 * H + 2 * (H + 6) + 2 + H * H + H;
    case CELL_LSTM: return (size_t)4 * H * 2 + 4 * H * H + 8 * H + 2 * H + 2;
    case CELL_DELTAGRU: return (size_t)3 * H * 6 + 3 * H * H + 6 * H + 2 * H + 2;
    case CELL_TRES: return (size_t)3 * H * 6 + 3 * H * H + 2 * H + 18 + 6;
    case CELL_PGJANET: return (size_t)3 * (H * (H + 1) + H) + 2 * (2 * H * H + H) + 2 * H + 2;
    case CELL_DVRJANET: return (size_t)K + 3 * H * H + 2 * H + H + 2 * (2 * H * H + H) + 2 * (H + 1);
    case CELL_GMP: return 495;
    case CELL_RVTDCNN: return (size_t)32 + 39 * H;
    case CELL_BOJANET: return (size_t)2 * H * H + 28 * H + 194;
    case CELL_TCNN: return (size_t)29 * H;
    case CELL_NEURALTX: return (size_t)27 * H + 14;
    case CELL_APNRRU: return (size_t)241 + 34 * (2 * H + 3) + 2 * H;
    case CELL_MCLDNN: return (size_t)190 * H + 589;
    case CELL_DELTAJANET: return (size_t)2 * H * H + 18 * H + 2;
    case CELL_TRES_QAT: return (size_t)3 * H * H + 20 * H + 37;
    case CELL_QGRU_QAT: case CELL_QGRU_AMP1_QAT: return (size_t)3 * H * 4 + 3 * H * H + 6 * H + 2 * H + 2 + 13;
    }
    return 0;
}

static void seq_dispatch(const Ctx *c, const REAL *x, const REAL *gout, REAL *out, REAL *gx, REAL *gp, int phase,
                         uint64_t *mx, uint64_t *mh, int64_t *st) {
    switch (c->cell) {
    case CELL_GRU: case CELL_DGRU: case CELL_QGRU: case CELL_QGRU_AMP1: seq_gru_family(c, x, gout, out, gx, gp, phase); break;
    case CELL_LSTM: seq_lstm(c, x, gout, out, gx, gp, phase); break;
    case CELL_DELTAGRU: case CELL_TRES: seq_delta(c, x, gout, out, gx, gp, phase, mx, mh, st); break;
    case CELL_PGJANET: seq_pgjanet(c, x, gout, out, gx, gp, phase); break;
    case CELL_DVRJANET: seq_dvrjanet(c, x, gout, out, gx, gp, phase); break;
    case CELL_GMP: seq_gmp(c, x, gout, out, gx, gp, phase); break;
    case CELL_RVTDCNN: seq_rvtdcnn(c, x, gout, out, gx, gp, phase); break;
    case CELL_BOJANET: seq_bojanet(c, x, gout, out, gx, gp, phase); break;
    case CELL_TCNN: case CELL_NEURALTX: seq_tcn(c, x, gout, out, gx, gp, phase); break;
    case CELL_APNRRU: seq_apnrru(c, x, gout, out, gx, gp, phase); break;
    case CELL_MCLDNN: seq_mcldnn(c, x, gout, out, gx, gp, phase); break;
    case CELL_DELTAJANET: seq_deltajanet(c, x, gout, out, gx, gp, phase); break;
    case CELL_TRES_QAT: seq_tres_qat(c, x, gout, out, gx, gp, phase, mx, mh, st); break;
    case CELL_QGRU_QAT: case CELL_QGRU_AMP1_QAT: seq_qgru_qat(c, x, gout, out, gx, gp, phase); break;
    }
}

/*
 * One full pass: forward (+ MSE vs target, nn.MSELoss mean over `loss_count` scalars, project.py:262-272)
 * and, when gx/gparams are requested, backward.  train_funcs.py:33-39 (fwd, criterion, backward).
 *   x,target,out,gx : (B,T,2) contiguous;  gout_in (B,T,2) used instead of the MSE gradient when target==NULL
 *   loss            : sum of squared errors / loss_count  (double)
 *   mask_x, mask_h  : (B,T) uint64 keep-bitfields (delta cells) or NULL;  stats: 4 x int64 accumulated (+=)
 * Returns 0, or -1 on bad arguments.
 */
int SUF(odpd_oracle_run)(int cell, int B, int T, int H, int K, double thx, double thh, const REAL *x, const REAL *target,
                         const REAL *gout_in, const REAL *params, REAL *out, double *loss, double loss_count, REAL *gx,
                         REAL *gparams, uint64_t *mask_x, uint64_t *mask_h, int64_t *stats, int nthreads) {
    if (B < 0 || T < 0 || H > 64 || (cell != CELL_GMP && H < 1)) return -1;
    const size_t P = n_params(cell, H, K);
    Ctx c = {cell, B, T, H, K, (REAL)thx, (REAL)thh, params, gx != NULL, gparams != NULL};
    const int bwd = (gx || gparams);
    if (nthreads < 1) nthreads = 1;
    double lsum = 0;
    REAL *gp_all = bwd ? (REAL *)calloc((size_t)nthreads * P, sizeof(REAL)) : NULL;
    int64_t st_all[4] = {0, 0, 0, 0};
    if (gx) memset(gx, 0, sizeof(REAL) * (size_t)B * T * 2);
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads) reduction(+ : lsum)
#endif
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        REAL *gtmp = (REAL *)malloc(sizeof(REAL) * (size_t)(T > 0 ? T : 1) * 2);
        int64_t st_loc[4] = {0, 0, 0, 0};
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int b = 0; b < B; ++b) {
            const REAL *xb = x + (size_t)b * T * 2;
            REAL *ob = out + (size_t)b * T * 2;
            int64_t st[4] = {0, 0, 0, 0};
            seq_dispatch(&c, xb, NULL, ob, NULL, NULL, 0, mask_x ? mask_x + (size_t)b * T : NULL,
                         mask_h ? mask_h + (size_t)b * T : NULL, st);
            for (int k = 0; k < 4; ++k) st_loc[k] += st[k];
            const REAL *g = NULL;
            if (target) {
                const REAL *yb = target + (size_t)b * T * 2;
                for (int i = 0; i < 2 * T; ++i) {
                    REAL d = ob[i] - yb[i];
                    lsum += (double)d * (double)d;
                    gtmp[i] = (REAL)((double)2 * (double)d / loss_count);
                }
                g = gtmp;
            } else if (gout_in) g = gout_in + (size_t)b * T * 2;
            if (bwd && g)
                seq_dispatch(&c, xb, g, ob, gx ? gx + (size_t)b * T * 2 : NULL, gp_all + (size_t)tid * P, 1, NULL, NULL, NULL);
        }
        free(gtmp);
#ifdef _OPENMP
#pragma omp critical
#endif
        for (int k = 0; k < 4; ++k) st_all[k] += st_loc[k];
    }
    if (loss) *loss = target ? lsum / loss_count : 0.0;
    if (gparams) {
        memset(gparams, 0, sizeof(REAL) * P);
        for (int t = 0; t < nthreads; ++t)
            for (size_t i = 0; i < P; ++i) gparams[i] += gp_all[(size_t)t * P + i];
    }
    if (stats) for (int k = 0; k < 4; ++k) stats[k] += st_all[k];
    free(gp_all);
    return 0;
}

#ifndef REAL_IS_DOUBLE
size_t odpd_oracle_n_params(int cell, int H, int K) { return n_params(cell, H, K); }
#endif
