/* ============================================================================================
 * TEST INFRASTRUCTURE ONLY -- the product never links, imports or executes this file.
 *
 * Plain-C (C99, double precision, single thread) restatement of the coarse-to-fine variational
 * optical-flow path of ElijahHyndman/PAPTeam_OpticalFlow (Ce Liu's solver as wrapped by pyflow).
 * It is the CHECKER for the CUDA path: tests/ compare the GPU results with it, bench.py may time
 * it as a CPU baseline, and nothing else may call it.
 *
 * PARITY PINNED: tests/test_oracle_vs_ref.py checks every function here bit-for-bit against the
 * unmodified reference compiled from /root/reference (oracle/_ref, see oracle/Makefile), and
 * tests/test_oracle_golden.py checks it against fixtures generated from that reference
 * (tests/golden/make_golden.py).  The reference itself ships no tests or golden vectors.
 *
 * Citations: S/ = /root/reference/Code/Serial/src/.  All images are row-major HWC interleaved
 * doubles, pixel (y,x) channel k at [(y*W+x)*C+k]  (S/Image.h:36-461).
 * Compile with -ffp-contract=off (see oracle/Makefile) so no multiply-add is fused.
 * ========================================================================================== */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_LEVELS 64

static int clampi(int v, int n) { /* S/ImageProcessing.h:34 EnforceRange */
    if (v < 0) v = 0;
    if (v > n - 1) v = n - 1;
    return v;
}

static double* newz(size_t n) { return (double*)calloc(n ? n : 1, sizeof(double)); }

/* --------------------------------------------------------------------------------------------
 * 1-D correlation with replicate borders.  S/ImageProcessing.h:259-279 (h), :350-369 (v):
 *   dst(p) = sum_{l=-f..f} tap[l+f] * src(clamp(p+l)), l ascending, accumulator starts at 0.
 * ------------------------------------------------------------------------------------------ */
void oracle_filter_h(const double* src, double* dst, int w, int h, int c, const double* tap, int f) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            for (int k = 0; k < c; k++) {
                double acc = 0.0;
                for (int l = -f; l <= f; l++)
                    acc += src[((size_t)y * w + clampi(x + l, w)) * c + k] * tap[l + f];
                dst[((size_t)y * w + x) * c + k] = acc;
            }
}

void oracle_filter_v(const double* src, double* dst, int w, int h, int c, const double* tap, int f) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            for (int k = 0; k < c; k++) {
                double acc = 0.0;
                for (int l = -f; l <= f; l++)
                    acc += src[((size_t)clampi(y + l, h) * w + x) * c + k] * tap[l + f];
                dst[((size_t)y * w + x) * c + k] = acc;
            }
}

/* S/Image.h:1347-1356 imfilter_hv: horizontal pass into a temporary, then vertical pass. */
static void filter_hv(const double* src, double* dst, int w, int h, int c, const double* th, int fh,
                      const double* tv, int fv) {
    double* tmp = newz((size_t)w * h * c);
    oracle_filter_h(src, tmp, w, h, c, th, fh);
    oracle_filter_v(tmp, dst, w, h, c, tv, fv);
    free(tmp);
}

/* S/Image.h:1203-1225 GaussianSmoothing: taps exp(-i^2/(2 sigma^2)), normalised by their sum. */
void oracle_gaussian(const double* src, double* dst, int w, int h, int c, double sigma, int fsize) {
    double tap[2 * 32 + 1];
    double two_s2 = sigma * sigma * 2, sum = 0;
    for (int i = -fsize; i <= fsize; i++) {
        tap[i + fsize] = exp(-(double)(i * i) / two_s2);
        sum += tap[i + fsize];
    }
    for (int i = 0; i < 2 * fsize + 1; i++) tap[i] /= sum;
    filter_hv(src, dst, w, h, c, tap, fsize, tap, fsize);
}

/* --------------------------------------------------------------------------------------------
 * Bilinear sampler.  S/ImageProcessing.h:138-157: integer part by C truncation (toward zero),
 * fractional part clamped to [0,1], the four taps index-clamped, visiting order
 * (m,n)=(0,0),(0,1),(1,0),(1,1) with m the x offset, ACCUMULATING into `out` (caller zeroes it).
 * ------------------------------------------------------------------------------------------ */
static void bilinear_acc(const double* im, int w, int h, int c, double x, double y, double* out) {
    int xi = (int)x, yi = (int)y;
    double fx = x - xi, fy = y - yi;
    if (fx > 1) fx = 1;
    if (fx < 0) fx = 0;
    if (fy > 1) fy = 1;
    if (fy < 0) fy = 0;
    for (int m = 0; m <= 1; m++)
        for (int n = 0; n <= 1; n++) {
            int u = clampi(xi + m, w), v = clampi(yi + n, h);
            double s = fabs(1 - m - fx) * fabs(1 - n - fy);
            const double* p = im + ((size_t)v * w + u) * c;
            for (int k = 0; k < c; k++) out[k] += p[k] * s;
        }
}

/* S/ImageProcessing.h:214-232: resize by a ratio; dst size by double->int truncation
 * (S/Image.h:755-756); source coordinate (j+1)/ratio-1. */
void oracle_resize_ratio(const double* src, double* dst, int sw, int sh, int c, double ratio) {
    int dw = (int)((double)sw * ratio), dh = (int)((double)sh * ratio);
    memset(dst, 0, sizeof(double) * (size_t)dw * dh * c);
    for (int i = 0; i < dh; i++)
        for (int j = 0; j < dw; j++) {
            double x = (double)(j + 1) / ratio - 1, y = (double)(i + 1) / ratio - 1;
            bilinear_acc(src, sw, sh, c, x, y, dst + ((size_t)i * dw + j) * c);
        }
}

/* S/ImageProcessing.h:235-253: resize to explicit size with separate x / y ratios; then the
 * caller's Multiplywith(scale) (S/Image.h:1841-1850) folded in as a final multiply. */
void oracle_resize_to(const double* src, double* dst, int sw, int sh, int c, int dw, int dh,
                      double scale) {
    double rx = (double)dw / sw, ry = (double)dh / sh;
    memset(dst, 0, sizeof(double) * (size_t)dw * dh * c);
    for (int i = 0; i < dh; i++)
        for (int j = 0; j < dw; j++) {
            double x = (double)(j + 1) / rx - 1, y = (double)(i + 1) / ry - 1;
            double* o = dst + ((size_t)i * dw + j) * c;
            bilinear_acc(src, sw, sh, c, x, y, o);
            if (scale != 1.0)
                for (int k = 0; k < c; k++) o[k] *= scale;
        }
}

/* --------------------------------------------------------------------------------------------
 * Pyramid geometry and construction.  S/GaussianPyramid.cpp:47-108.
 * ------------------------------------------------------------------------------------------ */
double oracle_effective_ratio(double ratio) { /* :50-51 */
    return (ratio > 0.98 || ratio < 0.4) ? 0.75 : ratio;
}

int oracle_levels_from_min_width(int width, double ratio, int min_width) { /* :53 */
    ratio = oracle_effective_ratio(ratio);
    return (int)(log((double)min_width / width) / log(ratio));
}

/* Fills ws/hs[0..nlevels) and, per level, the source level, Gaussian half-width, sigma and the
 * resize ratio actually used.  Returns nlevels. */
int oracle_level_geometry(int w0, int h0, double ratio, int nlevels, int* ws, int* hs, int* src_lvl,
                          int* fsize, double* sigma, double* rate) {
    ratio = oracle_effective_ratio(ratio);
    double base_sigma = 1 / ratio - 1;              /* :58 */
    int n = (int)(log(0.25) / log(ratio));          /* :59 */
    double n_sigma = base_sigma * n;                /* :60 */
    ws[0] = w0; hs[0] = h0;
    if (src_lvl) src_lvl[0] = 0;
    if (fsize) fsize[0] = 0;
    if (sigma) sigma[0] = 0;
    if (rate) rate[0] = 1;
    for (int i = 1; i < nlevels; i++) {
        int s; double sg, r;
        if (i <= n) { s = 0; sg = base_sigma * i; r = pow(ratio, i); }                    /* :66-68 */
        else { s = i - n; sg = n_sigma; r = (double)pow(ratio, i) * w0 / ws[s]; }         /* :72-74 */
        ws[i] = (int)((double)ws[s] * r);
        hs[i] = (int)((double)hs[s] * r);
        if (src_lvl) src_lvl[i] = s;
        if (fsize) fsize[i] = (int)(sg * 3);        /* int fsize parameter: truncation */
        if (sigma) sigma[i] = sg;
        if (rate) rate[i] = r;
    }
    return nlevels;
}

/* Builds all levels into one caller-provided buffer laid out level after level (HWC each);
 * size needed = sum_k ws[k]*hs[k]*c (query with oracle_level_geometry). */
void oracle_pyramid(const double* im, int w0, int h0, int c, double ratio, int nlevels, double* out) {
    int ws[ORACLE_MAX_LEVELS], hs[ORACLE_MAX_LEVELS], sl[ORACLE_MAX_LEVELS], fs[ORACLE_MAX_LEVELS];
    double sg[ORACLE_MAX_LEVELS], rt[ORACLE_MAX_LEVELS];
    size_t off[ORACLE_MAX_LEVELS];
    oracle_level_geometry(w0, h0, ratio, nlevels, ws, hs, sl, fs, sg, rt);
    size_t o = 0;
    for (int i = 0; i < nlevels; i++) { off[i] = o; o += (size_t)ws[i] * hs[i] * c; }
    if (nlevels > 0) memcpy(out, im, sizeof(double) * (size_t)w0 * h0 * c);
    for (int i = 1; i < nlevels; i++) {
        int s = sl[i];
        double* blur = newz((size_t)ws[s] * hs[s] * c);
        oracle_gaussian(out + off[s], blur, ws[s], hs[s], c, sg[i], fs[i]);
        oracle_resize_ratio(blur, out + off[i], ws[s], hs[s], c, rt[i]);
        free(blur);
    }
}

/* --------------------------------------------------------------------------------------------
 * Features.  S/OpticalFlow.cpp:911-961; luma S/Image.h:1461-1480; 5-tap derivative
 * [1,-8,0,8,-1]/12 S/Image.h:987-993,1030-1036.  Returns the feature channel count.
 * swap_luma: S/Image.h:1475-1478 uses the B-first weights whenever colorType != RGB.
 * ------------------------------------------------------------------------------------------ */
static void deriv_taps(double* t) {
    static const double raw[5] = {1, -8, 0, 8, -1};
    for (int i = 0; i < 5; i++) t[i] = raw[i] / 12;
}

int oracle_im2feature(const double* im, double* feat, int w, int h, int c, int swap_luma) {
    size_t np = (size_t)w * h;
    double t5[5];
    deriv_taps(t5);
    if (c != 1 && c != 3) { /* :956-957 */
        memcpy(feat, im, sizeof(double) * np * c);
        return c;
    }
    double* g = newz(np);
    double* gx = newz(np);
    double* gy = newz(np);
    if (c == 1) memcpy(g, im, sizeof(double) * np);
    else
        for (size_t i = 0; i < np; i++) {
            const double* p = im + i * 3;
            g[i] = swap_luma ? (double)p[0] * .114 + p[1] * .587 + p[2] * .299
                             : (double)p[0] * .299 + p[1] * .587 + p[2] * .114;
        }
    oracle_filter_h(g, gx, w, h, 1, t5, 2);
    oracle_filter_v(g, gy, w, h, 1, t5, 2);
    int fc = (c == 1) ? 3 : 5;
    for (size_t i = 0; i < np; i++) {
        double* o = feat + i * fc;
        o[0] = g[i]; o[1] = gx[i]; o[2] = gy[i];
        if (c == 3) {
            o[3] = im[i * 3 + 1] - im[i * 3];
            o[4] = im[i * 3 + 1] - im[i * 3 + 2];
        }
    }
    free(g); free(gx); free(gy);
    return fc;
}

/* --------------------------------------------------------------------------------------------
 * getDxs.  S/OpticalFlow.cpp:80-122: smooth both images with [.02,.11,.74,.11,.02] (h then v),
 * blend 0.4/0.6, 5-tap derivatives of the blend, temporal difference of the smoothed images.
 * ------------------------------------------------------------------------------------------ */
void oracle_getdxs(double* imdx, double* imdy, double* imdt, const double* im1, const double* im2,
                   int w, int h, int c) {
    static const double g5[5] = {0.02, 0.11, 0.74, 0.11, 0.02};
    size_t n = (size_t)w * h * c;
    double t5[5];
    deriv_taps(t5);
    double* s1 = newz(n);
    double* s2 = newz(n);
    double* bl = newz(n);
    filter_hv(im1, s1, w, h, c, g5, 2, g5, 2);
    filter_hv(im2, s2, w, h, c, g5, 2, g5, 2);
    for (size_t i = 0; i < n; i++) {
        double a = s1[i] * 0.4;      /* Multiplywith(0.4)   :92 */
        bl[i] = a + s2[i] * 0.6;     /* Add(Im2, 0.6)       :93, S/Image.h:1905-1918 */
    }
    oracle_filter_h(bl, imdx, w, h, c, t5, 2);
    oracle_filter_v(bl, imdy, w, h, c, t5, 2);
    for (size_t i = 0; i < n; i++) imdt[i] = s2[i] - s1[i];
    free(s1); free(s2); free(bl);
}

/* --------------------------------------------------------------------------------------------
 * Bilinear warp with Im1 fallback outside the image.  S/OpticalFlow.cpp:154-159 ->
 * S/ImageProcessing.h:483-503.
 * ------------------------------------------------------------------------------------------ */
void oracle_warpfl(double* warp, const double* im1, const double* im2, const double* vx,
                   const double* vy, int w, int h, int c) {
    memset(warp, 0, sizeof(double) * (size_t)w * h * c);
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            size_t p = (size_t)i * w + j;
            double x = j + vx[p], y = i + vy[p];
            if (x < 0 || x > w - 1 || y < 0 || y > h - 1)
                for (int k = 0; k < c; k++) warp[p * c + k] = im1[p * c + k];
            else
                bilinear_acc(im2, w, h, c, x, y, warp + p * c);
        }
}

/* --------------------------------------------------------------------------------------------
 * Weighted Laplacian in the fork's fused form.  S/OpticalFlow.cpp:641-690.  The horizontal loop
 * stops at column W-2 and does its "+= flux from the left" inside that loop, so column W-1 never
 * receives a horizontal term; likewise row H-1 never receives a vertical term.
 * ------------------------------------------------------------------------------------------ */
void oracle_laplacian(double* out, const double* in, const double* wt, int w, int h) {
    size_t n = (size_t)w * h;
    double* flux = newz(n);
    memset(out, 0, sizeof(double) * n);
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w - 1; j++) {
            size_t p = (size_t)i * w + j;
            flux[p] = (in[p + 1] - in[p]) * wt[p];
            out[p] -= flux[p];
            if (j > 0) out[p] += flux[p - 1];
        }
    memset(flux, 0, sizeof(double) * n);
    for (int i = 0; i < h - 1; i++)
        for (int j = 0; j < w; j++) {
            size_t p = (size_t)i * w + j;
            flux[p] = (in[p + w] - in[p]) * wt[p];
            out[p] -= flux[p];
            if (i > 0) out[p] += flux[p - w];
        }
    free(flux);
}

/* --------------------------------------------------------------------------------------------
 * Alternative solver branches.  The reference selects them through two public
 * static members, OpticalFlow::interpolation {Bilinear, Bicubic} and OpticalFlow::noiseModel
 * {GMixture, Lap} (S/OpticalFlow.h:19-27, defaults Bilinear / Lap at S/OpticalFlow.cpp:33-34); the
 * oracle keeps the same process-global switch.  Enum values follow the reference's declaration order.
 * ------------------------------------------------------------------------------------------ */
enum { ORACLE_BILINEAR = 0, ORACLE_BICUBIC = 1 };
enum { ORACLE_GMIXTURE = 0, ORACLE_LAP = 1 };
static int g_interp = ORACLE_BILINEAR, g_noise = ORACLE_LAP;
/* GaussianMixture state, S/NoiseModel.h:17-24 */
static struct { double alpha[16], sigma[16], beta[16], sigma2[16], beta2[16]; } g_gm;

void oracle_set_variant(int interpolation, int noise_model) { g_interp = interpolation; g_noise = noise_model; }

static void gm_square(void) { /* S/NoiseModel.h:129-136 */
    for (int i = 0; i < 16; i++) { g_gm.sigma2[i] = g_gm.sigma[i] * g_gm.sigma[i]; g_gm.beta2[i] = g_gm.beta[i] * g_gm.beta[i]; }
}
static void gm_reset(void) { /* S/NoiseModel.h:98-108 */
    for (int i = 0; i < 16; i++) { g_gm.alpha[i] = 0.95; g_gm.sigma[i] = 0.05; g_gm.beta[i] = 0.5; }
    gm_square();
}
/* Quirk: S/NoiseModel.h:10-12 defines PI only #ifndef PI, and S/Image.h:14 has already pulled in
 * S/Stochastic.h:18-20 by then, so the value the reference's Gaussian() compiles with is 3.1415927. */
#define ORACLE_PI 3.1415927
static double gm_gaussian(double x, int i, int k) { /* S/NoiseModel.h:116-122 */
    if (i == 0) return exp(-x / (2 * g_gm.sigma2[k])) / (2 * ORACLE_PI * g_gm.sigma[k]);
    return exp(-x / (2 * g_gm.beta2[k])) / (2 * ORACLE_PI * g_gm.beta[k]);
}
void oracle_gm_get(double* alpha, double* sigma, double* beta, int c) {
    for (int k = 0; k < c; k++) { alpha[k] = g_gm.alpha[k]; sigma[k] = g_gm.sigma[k]; beta[k] = g_gm.beta[k]; }
}
void oracle_gm_reset(void) { gm_reset(); }

/* Three EM iterations of the two-component mixture on (Im1 - warpIm2)^2 per channel.
 * S/OpticalFlow.cpp:539-591.  Note the M step starts from para.reset() (:564), i.e. the sums of
 * sigma and beta start at 0.05 and 0.5, not at zero. */
void oracle_est_gaussian_mixture(const double* im1, const double* im2, int w, int h, int c) {
    const double prior = 0.9; /* default argument, S/OpticalFlow.h:44 */
    size_t n = (size_t)w * h;
    double* w1 = newz(n * c); double* w2 = newz(n * c);
    for (int count = 0; count < 3; count++) {
        double total1[16] = {0}, total2[16] = {0};
        for (size_t i = 0; i < n; i++)
            for (int k = 0; k < c; k++) {
                size_t q = i * c + k;
                double t = im1[q] - im2[q];
                t *= t;
                w1[q] = gm_gaussian(t, 0, k) * g_gm.alpha[k];
                w2[q] = gm_gaussian(t, 1, k) * (1 - g_gm.alpha[k]);
                t = w1[q] + w2[q];
                w1[q] /= t;
                w2[q] /= t;
                total1[k] += w1[q];
                total2[k] += w2[q];
            }
        gm_reset();
        for (size_t i = 0; i < n; i++)
            for (int k = 0; k < c; k++) {
                size_t q = i * c + k;
                double t = im1[q] - im2[q];
                t *= t;
                g_gm.sigma[k] += w1[q] * t;
                g_gm.beta[k] += w2[q] * t;
            }
        for (int k = 0; k < c; k++) {
            g_gm.alpha[k] = total1[k] / (total1[k] + total2[k]) * (1 - prior) + 0.95 * prior;
            g_gm.sigma[k] = sqrt(g_gm.sigma[k] / total1[k]);
            g_gm.beta[k] = sqrt(g_gm.beta[k] / total2[k]) * (1 - prior) + 0.3 * prior;
        }
        gm_square();
    }
    free(w1); free(w2);
}

/* Per-channel mean |Im1-warpIm2| over elements with 0<d<1e6, 0.001 if none.
 * S/OpticalFlow.cpp:594-639. */
void oracle_est_laplacian_noise(const double* im1, const double* im2, int w, int h, int c,
                                double* para) {
    double cnt[16] = {0};
    for (int k = 0; k < c; k++) para[k] = 0;
    for (size_t i = 0; i < (size_t)w * h; i++)
        for (int k = 0; k < c; k++) {
            double d = fabs(im1[i * c + k] - im2[i * c + k]);
            if (d > 0 && d < 1000000) { para[k] += d; cnt[k]++; }
        }
    for (int k = 0; k < c; k++) para[k] = (cnt[k] == 0) ? 0.001 : para[k] / cnt[k];
}

/* --------------------------------------------------------------------------------------------
 * One SOR relaxation of pixel p=(i,j).  S/OpticalFlow.cpp:463-504.
 * Weights: left phi(p-1), right phi(p), up phi(p-W), down phi(p); absent neighbours skipped.
 * du is updated first and dv then uses the NEW du.
 * ------------------------------------------------------------------------------------------ */
static void sor_pixel(int i, int j, int w, int h, double alpha, double omega, const double* phi,
                      const double* dxy, const double* dx2, const double* dy2, const double* bu,
                      const double* bv, double* du, double* dv) {
    size_t p = (size_t)i * w + j;
    double s1 = 0, s2 = 0, cf = 0, wt;
    if (j > 0)     { wt = phi[p - 1]; s1 += wt * du[p - 1]; s2 += wt * dv[p - 1]; cf += wt; }
    if (j < w - 1) { wt = phi[p];     s1 += wt * du[p + 1]; s2 += wt * dv[p + 1]; cf += wt; }
    if (i > 0)     { wt = phi[p - w]; s1 += wt * du[p - w]; s2 += wt * dv[p - w]; cf += wt; }
    if (i < h - 1) { wt = phi[p];     s1 += wt * du[p + w]; s2 += wt * dv[p + w]; cf += wt; }
    s1 *= -alpha;
    s2 *= -alpha;
    cf *= alpha;
    s1 += dxy[p] * dv[p];
    du[p] = (1 - omega) * du[p] + omega / (dx2[p] + alpha * 0.05 + cf) * (bu[p] - s1);
    s2 += dxy[p] * du[p];
    dv[p] = (1 - omega) * dv[p] + omega / (dy2[p] + alpha * 0.05 + cf) * (bv[p] - s2);
}

/* order 0: lexicographic Gauss-Seidel (the reference, S/OpticalFlow.cpp:458-505).
 * order 1: red-black (pixels with (i+j) even first, then odd) -- NOT the reference; provided so
 *          tests can separate ordering error from FP32 rounding error in the fast GPU mode. */
void oracle_sor_solve(double* du, double* dv, const double* phi, const double* dxy,
                      const double* dx2, const double* dy2, const double* bu, const double* bv,
                      int w, int h, double alpha, double omega, int nsor, int order) {
    for (int s = 0; s < nsor; s++) {
        if (order == 0) {
            for (int i = 0; i < h; i++)
                for (int j = 0; j < w; j++)
                    sor_pixel(i, j, w, h, alpha, omega, phi, dxy, dx2, dy2, bu, bv, du, dv);
        } else {
            for (int col = 0; col < 2; col++)
                for (int i = 0; i < h; i++)
                    for (int j = (i + col) & 1; j < w; j += 2)
                        sor_pixel(i, j, w, h, alpha, omega, phi, dxy, dx2, dy2, bu, bv, du, dv);
        }
    }
}

/* --------------------------------------------------------------------------------------------
 * Linear-system assembly for one inner fixed-point iteration.  S/OpticalFlow.cpp:295-448.
 * Outputs phi and the five collapsed planes; bu/bv are the right-hand sides after :444-448.
 * du/dv are the CURRENT increments (zero at the first inner iteration).
 * lap[k] < 1e-20 leaves psi of channel k at zero (:399-400).
 * ------------------------------------------------------------------------------------------ */
void oracle_assemble(const double* imdx, const double* imdy, const double* imdt, const double* u,
                     const double* v, const double* du, const double* dv, const double* lap,
                     int w, int h, int c, double alpha, double* phi, double* dxy, double* dx2,
                     double* dy2, double* bu, double* bv) {
    size_t n = (size_t)w * h;
    double eps = pow(0.001, 2); /* :261-262 */
    double* uu = newz(n); double* vv = newz(n);
    double* ux = newz(n); double* uy = newz(n); double* vx = newz(n); double* vy = newz(n);
    for (size_t i = 0; i < n; i++) { uu[i] = u[i] + du[i]; vv[i] = v[i] + dv[i]; } /* :297-303 */
    /* forward differences: last column zero (S/Image.h:975-986); last row never written and
     * therefore still zero from construction (S/Image.h:1017-1029, S/OpticalFlow.cpp:250-251) */
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w - 1; j++) {
            size_t p = (size_t)i * w + j;
            ux[p] = uu[p + 1] - uu[p];
            vx[p] = vv[p + 1] - vv[p];
        }
    for (int i = 0; i < h - 1; i++)
        for (int j = 0; j < w; j++) {
            size_t p = (size_t)i * w + j;
            uy[p] = uu[p + w] - uu[p];
            vy[p] = vv[p + w] - vv[p];
        }
    for (size_t i = 0; i < n; i++) { /* :325-331 */
        double t = ux[i] * ux[i] + uy[i] * uy[i] + vx[i] * vx[i] + vy[i] * vy[i];
        phi[i] = 0.5 / sqrt(t + eps);
    }
    double* dtdx = newz(n); double* dtdy = newz(n);
    for (size_t i = 0; i < n; i++) { /* :377-427 */
        double sxy = 0, sx2 = 0, sy2 = 0, stx = 0, sty = 0;
        for (int k = 0; k < c; k++) {
            size_t q = i * c + k;
            double psi = 0;
            if (g_noise == ORACLE_GMIXTURE) { /* :362-368, :388-394 */
                double t = imdt[q] + imdx[q] * du[i] + imdy[q] * dv[i];
                t *= t;
                double prob1 = gm_gaussian(t, 0, k) * g_gm.alpha[k];
                double prob2 = gm_gaussian(t, 1, k) * (1 - g_gm.alpha[k]);
                double prob11 = prob1 / (2 * g_gm.sigma2[k]);
                double prob22 = prob2 / (2 * g_gm.beta2[k]);
                psi = (prob11 + prob22) / (prob1 + prob2);
            } else if (!(lap[k] < 1E-20)) {
                double t = imdt[q] + imdx[q] * du[i] + imdy[q] * dv[i];
                t *= t;
                psi = 1 / (2 * sqrt(t + eps));
            }
            sxy += psi * imdx[q] * imdy[q];   /* Multiply(psi,a,b) = (psi*a)*b, S/Image.h:1762 */
            sx2 += psi * imdx[q] * imdx[q];
            sy2 += psi * imdy[q] * imdy[q];
            stx += psi * imdx[q] * imdt[q];
            sty += psi * imdy[q] * imdt[q];
        }
        if (c > 1) { /* collapse = mean, S/Image.h:1536-1543; single channel is a plain copy */
            sxy /= c; sx2 /= c; sy2 /= c; stx /= c; sty /= c;
        }
        dxy[i] = sxy; dx2[i] = sx2; dy2[i] = sy2; dtdx[i] = stx; dtdy[i] = sty;
    }
    double* l1 = newz(n); double* l2 = newz(n);
    oracle_laplacian(l1, u, phi, w, h); /* note: u, not u+du  (:437-438) */
    oracle_laplacian(l2, v, phi, w, h);
    for (size_t i = 0; i < n; i++) { /* :444-448 */
        bu[i] = -dtdx[i] - alpha * l1[i];
        bv[i] = -dtdy[i] - alpha * l2[i];
    }
    free(uu); free(vv); free(ux); free(uy); free(vx); free(vy);
    free(dtdx); free(dtdy); free(l1); free(l2);
}

/* --------------------------------------------------------------------------------------------
 * SmoothFlowSOR.  S/OpticalFlow.cpp:238-536.  u, v, warp are in/out; lap is the persistent
 * LapPara state (c entries used; :530 overwrites it every outer iteration).
 * ------------------------------------------------------------------------------------------ */
static void bicubic_warp_impl(double* out, const double* ref, const double* im2, const double* vx,
                              const double* vy, int w, int h, int c, int clamp);

void oracle_smoothflow_sor(const double* f1, const double* f2, double* warp, double* u, double* v,
                           double* lap, int w, int h, int c, double alpha, int n_outer,
                           int n_inner, int n_sor, int order) {
    size_t n = (size_t)w * h;
    double* imdx = newz(n * c); double* imdy = newz(n * c); double* imdt = newz(n * c);
    double* du = newz(n); double* dv = newz(n);
    double* phi = newz(n); double* dxy = newz(n); double* dx2 = newz(n); double* dy2 = newz(n);
    double* bu = newz(n); double* bv = newz(n);
    for (int it = 0; it < n_outer; it++) {
        oracle_getdxs(imdx, imdy, imdt, f1, warp, w, h, c);
        memset(du, 0, sizeof(double) * n);
        memset(dv, 0, sizeof(double) * n);
        for (int hh = 0; hh < n_inner; hh++) {
            oracle_assemble(imdx, imdy, imdt, u, v, du, dv, lap, w, h, c, alpha, phi, dxy, dx2,
                            dy2, bu, bv);
            memset(du, 0, sizeof(double) * n); /* :452-453 */
            memset(dv, 0, sizeof(double) * n);
            oracle_sor_solve(du, dv, phi, dxy, dx2, dy2, bu, bv, w, h, alpha, 1.8, n_sor, order);
        }
        for (size_t i = 0; i < n; i++) { u[i] += du[i]; v[i] += dv[i]; } /* :513-514 */
        if (g_interp == ORACLE_BILINEAR) oracle_warpfl(warp, f1, f2, u, v, w, h, c);   /* :515-516 */
        else bicubic_warp_impl(warp, f1, f2, u, v, w, h, c, 1);           /* :517-521 warpImageBicubicRef + threshold */
        if (g_noise == ORACLE_GMIXTURE) oracle_est_gaussian_mixture(f1, warp, w, h, c); /* :524-527 */
        else oracle_est_laplacian_noise(f1, warp, w, h, c, lap);          /* :528-530 */
    }
    free(imdx); free(imdy); free(imdt); free(du); free(dv);
    free(phi); free(dxy); free(dx2); free(dy2); free(bu); free(bv);
}

/* --------------------------------------------------------------------------------------------
 * Final output warp: Hermite-bicubic with Im1 fallback, then clamp to [0,1].
 * S/Image.h:2587-2595 (derivative images, taps [-.5,0,.5]), :2624-2701 (warp),
 * :2497-2530 (coefficients), :2031-2045 (threshold).
 *
 * The 16 polynomial coefficients a[i][j] (power of dx = i, of dy = j) are integer combinations of
 * the image value P and its derivative images X=dI/dx, Y=dI/dy, Z=d2I/dxdy at the four corners
 * (x0,y0) (x1,y0) (x0,y1) (x1,y1).  Source index = 4*quantity + corner with quantity P,X,Y,Z =
 * 0..3 and corner 00,10,01,11 = 0..3 (first digit is x).  Terms are summed left to right in the
 * table's order, which is the reference's evaluation order.
 * ------------------------------------------------------------------------------------------ */
enum { P00, P10, P01, P11, X00, X10, X01, X11, Y00, Y10, Y01, Y11, Z00, Z10, Z01, Z11 };
typedef struct { int n; signed char src[16]; signed char wt[16]; } bicubic_row;
static const bicubic_row BICUBIC[4][4] = {
    /* a[0][*] */
    {{1, {P00}, {1}},
     {1, {Y00}, {1}},
     {4, {P00, P01, Y00, Y01}, {-3, 3, -2, -1}},
     {4, {P00, P01, Y00, Y01}, {2, -2, 1, 1}}},
    /* a[1][*] */
    {{1, {X00}, {1}},
     {1, {Z00}, {1}},
     {4, {X00, X01, Z00, Z01}, {-3, 3, -2, -1}},
     {4, {X00, X01, Z00, Z01}, {2, -2, 1, 1}}},
    /* a[2][*] */
    {{4, {P00, P10, X00, X10}, {-3, 3, -2, -1}},
     {4, {Y00, Y10, Z00, Z10}, {-3, 3, -2, -1}},
     {16, {P00, P10, P01, P11, X00, X10, X01, X11, Y00, Y10, Y01, Y11, Z00, Z10, Z01, Z11},
          {9, -9, -9, 9, 6, 3, -6, -3, 6, -6, 3, -3, 4, 2, 2, 1}},
     {16, {P00, P10, P01, P11, X00, X10, X01, X11, Y00, Y10, Y01, Y11, Z00, Z10, Z01, Z11},
          {-6, 6, 6, -6, -4, -2, 4, 2, -3, 3, -3, 3, -2, -1, -2, -1}}},
    /* a[3][*] */
    {{4, {P00, P10, X00, X10}, {2, -2, 1, 1}},
     {4, {Y00, Y10, Z00, Z10}, {2, -2, 1, 1}},
     {16, {P00, P10, P01, P11, X00, X10, X01, X11, Y00, Y10, Y01, Y11, Z00, Z10, Z01, Z11},
          {-6, 6, 6, -6, -3, -3, 3, 3, -4, 4, -2, 2, -2, -2, -1, -1}},
     {16, {P00, P10, P01, P11, X00, X10, X01, X11, Y00, Y10, Y01, Y11, Z00, Z10, Z01, Z11},
          {4, -4, -4, 4, 2, 2, -2, -2, 2, -2, 2, -2, 1, 1, 1, 1}}},
};

static void bicubic_warp_impl(double* out, const double* ref, const double* im2, const double* vx,
                              const double* vy, int w, int h, int c, int clamp) {
    static const double d3[3] = {-0.5, 0, 0.5};
    size_t n = (size_t)w * h * c;
    double* ix = newz(n); double* iy = newz(n); double* ixy = newz(n);
    oracle_filter_h(im2, ix, w, h, c, d3, 1);
    oracle_filter_v(im2, iy, w, h, c, d3, 1);
    oracle_filter_v(ix, ixy, w, h, c, d3, 1);
    const double* q[4] = {im2, ix, iy, ixy};
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            size_t p = (size_t)i * w + j;
            double x = j + vx[p], y = i + vy[p];
            if (x < 0 || x > w - 1 || y < 0 || y > h - 1) {
                for (int k = 0; k < c; k++) out[p * c + k] = ref[p * c + k];
                continue;
            }
            int x0 = clampi((int)x, w), x1 = clampi((int)x + 1, w);
            int y0 = clampi((int)y, h), y1 = clampi((int)y + 1, h);
            double dx = x - x0, dy = y - y0;
            double dx2 = dx * dx, dy2 = dy * dy, dx3 = dx * dx2, dy3 = dy * dy2;
            size_t corner[4] = {(size_t)y0 * w + x0, (size_t)y0 * w + x1, (size_t)y1 * w + x0,
                                (size_t)y1 * w + x1};
            for (int k = 0; k < c; k++) {
                double s[16], a[4][4];
                for (int t = 0; t < 16; t++) s[t] = q[t >> 2][corner[t & 3] * c + k];
                for (int r = 0; r < 4; r++)
                    for (int m = 0; m < 4; m++) {
                        const bicubic_row* br = &BICUBIC[r][m];
                        double acc = br->wt[0] * s[br->src[0]];
                        for (int t = 1; t < br->n; t++) acc += br->wt[t] * s[br->src[t]];
                        a[r][m] = acc;
                    }
                double val = a[0][0] + a[0][1] * dy + a[0][2] * dy2 + a[0][3] * dy3 +
                             a[1][0] * dx + a[1][1] * dx * dy + a[1][2] * dx * dy2 + a[1][3] * dx * dy3 +
                             a[2][0] * dx2 + a[2][1] * dx2 * dy + a[2][2] * dx2 * dy2 + a[2][3] * dx2 * dy3 +
                             a[3][0] * dx3 + a[3][1] * dx3 * dy + a[3][2] * dx3 * dy2 + a[3][3] * dx3 * dy3;
                out[p * c + k] = val;
            }
        }
    /* threshold(): clamp to [0,1] for floating images; the fallback copies of Im1 are clamped as well */
    if (clamp)
        for (size_t i = 0; i < n; i++) {
            if (out[i] < 0) out[i] = 0;
            if (out[i] > 1) out[i] = 1;
        }
    free(ix); free(iy); free(ixy);
}

void oracle_bicubic_warp(double* out, const double* ref, const double* im2, const double* vx,
                         const double* vy, int w, int h, int c) {
    bicubic_warp_impl(out, ref, im2, vx, vy, w, h, c, 1);
}

/* Flow file encoding.
 * OpticalFlow::SaveOpticalFlow, S/OpticalFlow.cpp:993-1003:
 *   foo[i] = (__min(__max(flow[i], -200), 200) + 200) * 160   -- double arithmetic, then the implicit
 *   double -> unsigned short conversion (truncation; the value is within [0, 64000]).
 * __max(a,b) is ((a)>(b)?(a):(b)): a NaN sample compares false and becomes -200 -> 0.
 * OpticalFlow::LoadOpticalFlow, S/OpticalFlow.cpp:962-975: flow[i] = (double)foo[i] / 160 - 200. */
void oracle_flow_encode_u16(unsigned short* q, const double* flow, long n) {
    for (long i = 0; i < n; i++) {
        double f = flow[i];
        f = f > -200 ? f : -200;
        f = f < 200 ? f : 200;
        q[i] = (unsigned short)((f + 200) * 160);
    }
}

void oracle_flow_decode_u16(double* flow, const unsigned short* q, long n) {
    for (long i = 0; i < n; i++) flow[i] = (double)q[i] / 160 - 200;
}


/* --------------------------------------------------------------------------------------------
 * Whole solve.  Level loop S/OpticalFlow.cpp:735-846 with the parameters the fork hard-codes
 * (:747-751) exposed, and the level count either given (ConstructPyramidLevels) or derived from
 * min_width (ConstructPyramid, S/GaussianPyramid.cpp:53) when nlevels <= 0.
 * col_type follows the wrapper: 0 = RGB, 1 = GRAY (S/Image.h:97-101); it only changes the luma
 * weights of a 3-channel LEVEL-0 image (copyData keeps colorType, allocate() resets it to RGB).
 * order: see oracle_sor_solve.   Returns the number of levels used.
 * ------------------------------------------------------------------------------------------ */
int oracle_coarse2fine_flow(double* vx, double* vy, double* warp_out, const double* im1,
                            const double* im2, double alpha, double ratio, int min_width,
                            int nlevels, int n_outer, int n_inner, int n_sor, int col_type, int h,
                            int w, int c, int order) {
    if (nlevels <= 0) nlevels = oracle_levels_from_min_width(w, ratio, min_width);
    if (nlevels <= 0 || nlevels > ORACLE_MAX_LEVELS) return -1;
    int ws[ORACLE_MAX_LEVELS], hs[ORACLE_MAX_LEVELS];
    size_t off[ORACLE_MAX_LEVELS], total = 0;
    oracle_level_geometry(w, h, ratio, nlevels, ws, hs, 0, 0, 0, 0);
    for (int k = 0; k < nlevels; k++) { off[k] = total; total += (size_t)ws[k] * hs[k] * c; }
    double* p1 = newz(total); double* p2 = newz(total);
    oracle_pyramid(im1, w, h, c, ratio, nlevels, p1);
    oracle_pyramid(im2, w, h, c, ratio, nlevels, p2);
    int fc = (c == 1) ? 3 : (c == 3 ? 5 : c);
    double lap[16];
    for (int k = 0; k < 16; k++) lap[k] = 0.02; /* :773-775 */
    if (g_noise == ORACLE_GMIXTURE) gm_reset();  /* :769-770 */
    /* NB: the driver divides the flow by the ORIGINAL ratio argument (:810), while the pyramid
     * silently substitutes 0.75 for out-of-range ratios; the fork always passes 0.75. */
    size_t n0 = (size_t)w * h;
    double* f1 = newz(n0 * fc); double* f2 = newz(n0 * fc); double* wf = newz(n0 * fc);
    double* u = newz(n0); double* v = newz(n0); double* t = newz(n0);
    int pw = 0, ph = 0;
    for (int k = nlevels - 1; k >= 0; k--) {
        int lw = ws[k], lh = hs[k];
        size_t ln = (size_t)lw * lh;
        int swap = (k == 0 && col_type == 1);
        oracle_im2feature(p1 + off[k], f1, lw, lh, c, swap);
        oracle_im2feature(p2 + off[k], f2, lw, lh, c, swap);
        if (k == nlevels - 1) {
            memset(u, 0, sizeof(double) * ln);
            memset(v, 0, sizeof(double) * ln);
            memcpy(wf, f2, sizeof(double) * ln * fc);
        } else {
            oracle_resize_to(u, t, pw, ph, 1, lw, lh, 1 / ratio);
            memcpy(u, t, sizeof(double) * ln);
            oracle_resize_to(v, t, pw, ph, 1, lw, lh, 1 / ratio);
            memcpy(v, t, sizeof(double) * ln);
            if (g_interp == ORACLE_BILINEAR) oracle_warpfl(wf, f1, f2, u, v, lw, lh, fc); /* :812-813 */
            else bicubic_warp_impl(wf, f1, f2, u, v, lw, lh, fc, 0); /* :814-815: no threshold() here */
        }
        oracle_smoothflow_sor(f1, f2, wf, u, v, lap, lw, lh, fc, alpha, n_outer + k, n_inner,
                              n_sor + k * 3, order);
        pw = lw; ph = lh;
    }
    memcpy(vx, u, sizeof(double) * n0);
    memcpy(vy, v, sizeof(double) * n0);
    oracle_bicubic_warp(warp_out, im1, im2, u, v, w, h, c);
    free(p1); free(p2); free(f1); free(f2); free(wf); free(u); free(v); free(t);
    return nlevels;
}
