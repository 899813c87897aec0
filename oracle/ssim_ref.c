/* ORACLE (test infrastructure, not product code).
 *
 * Independent fp64 restatement of the SSIM / PSNR / MSE arithmetic the reference obtains from
 * torchmetrics==0.11.4 (call sites: /root/reference/models/utils.py:38-47,
 * /root/reference/report.py:78-96,146,207-212; algorithm: SURVEY.md Appendix A).
 *
 * Deliberately a different formulation from oracle/torchmetrics_port.py (which mirrors the
 * upstream dataflow: reflect-pad -> 121-tap depthwise conv of 5 stacked planes in fp32):
 * here the 11-tap Gaussian is applied separably, in double precision, with reflect indexing
 * instead of a padded copy.  Agreement of the two pins the restatement (PARITY UNPINNED
 * against torchmetrics itself - it is not installed and the reference ships no vectors).
 *
 * Built by oracle/Makefile into oracle/_build/libssim_ref.so; loaded with ctypes by tests/,
 * bench.py's cpu_baseline leg and __graft_entry__.smoke() only.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TAPS 11
#define PAD 5

static void gaussian(double *g) {
    double s = 0.0;
    for (int i = 0; i < TAPS; ++i) {
        double d = (double)(i - PAD) / 1.5;
        g[i] = exp(-0.5 * d * d);
        s += g[i];
    }
    for (int i = 0; i < TAPS; ++i) g[i] /= s;
}

static inline int reflect(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

/* p,t: [n,h,w] float in [0,1].  Outputs (any may be NULL):
 *   ssim[n]      mean of the SSIM map over rows/cols PAD..dim-PAD-1
 *   sse[n]       sum of squared error per image
 *   full[n,h,w]  the full (reflect-padded) SSIM map, as return_full_image=True gives
 * returns 0, or -1 on bad sizes. */
int ssim_ref_f64(const float *p, const float *t, int n, int h, int w,
                 double *ssim, double *sse, double *full) {
    if (n < 0 || h <= 2 * PAD || w <= 2 * PAD) return -1;
    const double c1 = 1e-4, c2 = 9e-4;
    double g[TAPS];
    gaussian(g);
    size_t plane = (size_t)h * w;
    double *tmp = (double *)malloc(sizeof(double) * 5 * plane);
    if (!tmp) return -2;
    for (int im = 0; im < n; ++im) {
        const float *pp = p + im * plane, *tt = t + im * plane;
        double e = 0.0;
        /* horizontal pass of the five moments */
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
                for (int k = 0; k < TAPS; ++k) {
                    int xx = reflect(x + k - PAD, w);
                    double a = pp[(size_t)y * w + xx], b = tt[(size_t)y * w + xx];
                    a0 += g[k] * a; a1 += g[k] * b; a2 += g[k] * a * a; a3 += g[k] * b * b; a4 += g[k] * a * b;
                }
                size_t o = (size_t)y * w + x;
                tmp[o] = a0; tmp[plane + o] = a1; tmp[2 * plane + o] = a2; tmp[3 * plane + o] = a3; tmp[4 * plane + o] = a4;
                double d = (double)pp[o] - (double)tt[o];
                e += d * d;
            }
        double acc = 0.0;
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                double m[5] = {0, 0, 0, 0, 0};
                for (int k = 0; k < TAPS; ++k) {
                    int yy = reflect(y + k - PAD, h);
                    size_t o = (size_t)yy * w + x;
                    for (int q = 0; q < 5; ++q) m[q] += g[k] * tmp[q * plane + o];
                }
                double mpp = m[0] * m[0], mtt = m[1] * m[1], mpt = m[0] * m[1];
                double vp = m[2] - mpp, vt = m[3] - mtt, cv = m[4] - mpt;
                double s = ((2 * mpt + c1) * (2 * cv + c2)) / ((mpp + mtt + c1) * (vp + vt + c2));
                if (full) full[im * plane + (size_t)y * w + x] = s;
                if (y >= PAD && y < h - PAD && x >= PAD && x < w - PAD) acc += s;
            }
        if (ssim) ssim[im] = acc / ((double)(h - 2 * PAD) * (w - 2 * PAD));
        if (sse) sse[im] = e;
    }
    free(tmp);
    return 0;
}

/* Gradient of  L = ws * mean_b(ssim_b) + wp * psnr(batch-global)  w.r.t. p (de-normalised
 * domain), closed form of SURVEY.md Appendix A, fp64.  grad: [n,h,w]. */
int ssim_psnr_grad_ref_f64(const float *p, const float *t, int n, int h, int w,
                           double ws, double wp, double *grad) {
    if (n <= 0 || h <= 2 * PAD || w <= 2 * PAD) return -1;
    const double c1 = 1e-4, c2 = 9e-4;
    double g[TAPS];
    gaussian(g);
    int hv = h - 2 * PAD, wv = w - 2 * PAD;
    size_t plane = (size_t)h * w, vplane = (size_t)hv * wv;
    double *hm = (double *)malloc(sizeof(double) * 5 * (size_t)h * wv);
    double *co = (double *)malloc(sizeof(double) * 3 * vplane);
    double *ct = (double *)malloc(sizeof(double) * 3 * (size_t)h * wv);
    if (!hm || !co || !ct) return -2;
    double sse = 0.0;
    for (size_t i = 0; i < (size_t)n * plane; ++i) { double d = (double)p[i] - (double)t[i]; sse += d * d; }
    double norm = (double)n * hv * wv;
    for (int im = 0; im < n; ++im) {
        const float *pp = p + im * plane, *tt = t + im * plane;
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < wv; ++x) {
                double a[5] = {0, 0, 0, 0, 0};
                for (int k = 0; k < TAPS; ++k) {
                    double u = pp[(size_t)y * w + x + k], v = tt[(size_t)y * w + x + k];
                    a[0] += g[k] * u; a[1] += g[k] * v; a[2] += g[k] * u * u; a[3] += g[k] * v * v; a[4] += g[k] * u * v;
                }
                for (int q = 0; q < 5; ++q) hm[q * (size_t)h * wv + (size_t)y * wv + x] = a[q];
            }
        for (int y = 0; y < hv; ++y)
            for (int x = 0; x < wv; ++x) {
                double m[5] = {0, 0, 0, 0, 0};
                for (int k = 0; k < TAPS; ++k)
                    for (int q = 0; q < 5; ++q) m[q] += g[k] * hm[q * (size_t)h * wv + (size_t)(y + k) * wv + x];
                double mu1 = m[0], mu2 = m[1];
                double s11 = m[2] - mu1 * mu1, s22 = m[3] - mu2 * mu2, s12 = m[4] - mu1 * mu2;
                double A1 = 2 * mu1 * mu2 + c1, A2 = 2 * s12 + c2, B1 = mu1 * mu1 + mu2 * mu2 + c1, B2 = s11 + s22 + c2;
                double S = A1 * A2 / (B1 * B2);
                size_t o = (size_t)y * wv + x;
                co[o] = 2 * mu2 * (A2 - A1) / (B1 * B2) - 2 * mu1 * S * (B2 - B1) / (B1 * B2);
                co[vplane + o] = -S / B2;
                co[2 * vplane + o] = 2 * A1 / (B1 * B2);
            }
        /* adjoint: vertical then horizontal scatter == full correlation with the symmetric taps */
        memset(ct, 0, sizeof(double) * 3 * (size_t)h * wv);
        for (int q = 0; q < 3; ++q)
            for (int y = 0; y < hv; ++y)
                for (int x = 0; x < wv; ++x)
                    for (int k = 0; k < TAPS; ++k)
                        ct[q * (size_t)h * wv + (size_t)(y + k) * wv + x] += g[k] * co[q * vplane + (size_t)y * wv + x];
        for (int y = 0; y < h; ++y) {
            double row[3][512 + 2 * PAD];
            if (w > 512 + 2 * PAD) return -3;
            for (int q = 0; q < 3; ++q) for (int x = 0; x < w; ++x) row[q][x] = 0.0;
            for (int q = 0; q < 3; ++q)
                for (int x = 0; x < wv; ++x)
                    for (int k = 0; k < TAPS; ++k) row[q][x + k] += g[k] * ct[q * (size_t)h * wv + (size_t)y * wv + x];
            for (int x = 0; x < w; ++x) {
                size_t o = (size_t)y * w + x;
                double u = pp[o], v = tt[o];
                double gs = (row[0][x] + 2 * u * row[1][x] + v * row[2][x]) / norm;
                double gp = -(10.0 / log(10.0)) * 2.0 * (u - v) / sse;
                grad[im * plane + o] = ws * gs + wp * gp;
            }
        }
    }
    free(hm); free(co); free(ct);
    return 0;
}
