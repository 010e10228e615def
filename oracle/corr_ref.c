/*
 * ORACLE — test infrastructure only, never linked into or called by the product path.
 *
 * Plain-C restatement of the reference CUDA correlation package for every (pad, ks, md, s1, s2):
 *   forward   models/correlation_package/correlation_cuda_kernel.cu:41-114  (+ channels_first :15-39)
 *   backward  models/correlation_package/correlation_cuda_kernel.cu:116-300
 *   dims      models/correlation_package/correlation_cuda.cc:25-34
 * It keeps the reference's structure on purpose: zero-padded channel-last copies of both inputs,
 * one (n, y, x) site at a time, displacement channel tc = (tj+dr)*D + (ti+dr), divisor ks*ks*C,
 * and the truncating integer divisions of the backward windows.  Accumulation is in double so the
 * result is the "true" value both the reference and the B200 kernels are compared with.
 * Pinned by tests/golden (generated from the reference's own correlation_native, which equals this
 * for pad=md, ks=1, s1=s2=1) — see tests/test_oracle_golden.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int B, C, H, W, pad, ks, md, s1, s2;
    int kr, dr, D, oH, oW, pH, pW;
} geom_t;

static int make_geom(geom_t* g, int B, int C, int H, int W, int pad, int ks, int md, int s1, int s2) {
    g->B = B; g->C = C; g->H = H; g->W = W; g->pad = pad; g->ks = ks; g->md = md; g->s1 = s1; g->s2 = s2;
    g->kr = (ks - 1) / 2;
    g->dr = md / s2;
    g->D = 2 * g->dr + 1;
    g->pH = H + 2 * pad;
    g->pW = W + 2 * pad;
    int br = g->kr + md;
    if (g->pH - 2 * br <= 0 || g->pW - 2 * br <= 0) return -1;
    g->oH = (int)ceil((double)(g->pH - 2 * br) / (double)s1);
    g->oW = (int)ceil((double)(g->pW - 2 * br) / (double)s1);
    return 0;
}

int oracle_corr_dims(int H, int W, int pad, int ks, int md, int s1, int s2, int* D2, int* oH, int* oW) {
    geom_t g;
    if (make_geom(&g, 1, 1, H, W, pad, ks, md, s1, s2)) return -1;
    *D2 = g.D * g.D; *oH = g.oH; *oW = g.oW;
    return 0;
}

/* NCHW -> zero padded (B, pH, pW, C), the reference's rInput layout */
static double* to_padded_nhwc(const double* in, const geom_t* g) {
    size_t n = (size_t)g->B * g->pH * g->pW * g->C;
    double* r = (double*)calloc(n, sizeof(double));
    for (int b = 0; b < g->B; ++b)
        for (int c = 0; c < g->C; ++c)
            for (int y = 0; y < g->H; ++y)
                for (int x = 0; x < g->W; ++x)
                    r[(((size_t)b * g->pH + y + g->pad) * g->pW + x + g->pad) * g->C + c] =
                        in[(((size_t)b * g->C + c) * g->H + y) * g->W + x];
    return r;
}

static double at(const double* r, const geom_t* g, int b, int y, int x, int c) {
    if (y < 0 || y >= g->pH || x < 0 || x >= g->pW) return 0.0; /* reference would read out of bounds */
    return r[(((size_t)b * g->pH + y) * g->pW + x) * g->C + c];
}

int oracle_corr_fwd(const double* f1, const double* f2, double* out, int B, int C, int H, int W,
                    int pad, int ks, int md, int s1, int s2) {
    geom_t g;
    if (make_geom(&g, B, C, H, W, pad, ks, md, s1, s2)) return -1;
    double* r1 = to_padded_nhwc(f1, &g);
    double* r2 = to_padded_nhwc(f2, &g);
    double nelems = (double)(ks * ks * C);
    for (int n = 0; n < B; ++n)
        for (int by = 0; by < g.oH; ++by)
            for (int bx = 0; bx < g.oW; ++bx) {
                int y1 = by * s1 + md, x1 = bx * s1 + md;
                for (int tj = -g.dr; tj <= g.dr; ++tj)
                    for (int ti = -g.dr; ti <= g.dr; ++ti) {
                        int x2 = x1 + ti * s2, y2 = y1 + tj * s2;
                        double sum = 0.0;
                        for (int j = -g.kr; j <= g.kr; ++j)
                            for (int i = -g.kr; i <= g.kr; ++i)
                                for (int ch = 0; ch < C; ++ch)
                                    sum += at(r1, &g, n, y1 + j, x1 + i, ch) * at(r2, &g, n, y2 + j, x2 + i, ch);
                        int tc = (tj + g.dr) * g.D + (ti + g.dr);
                        out[(((size_t)n * g.D * g.D + tc) * g.oH + by) * g.oW + bx] = sum / nelems;
                    }
            }
    free(r1); free(r2);
    return 0;
}

int oracle_corr_bwd(const double* f1, const double* f2, const double* gout, double* g1, double* g2,
                    int B, int C, int H, int W, int pad, int ks, int md, int s1, int s2) {
    geom_t g;
    if (make_geom(&g, B, C, H, W, pad, ks, md, s1, s2)) return -1;
    double* r1 = to_padded_nhwc(f1, &g);
    double* r2 = to_padded_nhwc(f2, &g);
    double nelems = (double)(ks * ks * C);
    int nOut = g.D * g.D;
    for (int n = 0; n < B; ++n)
        for (int yu = 0; yu < H; ++yu)
            for (int xu = 0; xu < W; ++xu)
                for (int c = 0; c < C; ++c) {
                    int y = yu + pad, x = xu + pad;
                    double sum1 = 0.0, sum2 = 0.0;
                    for (int tc = 0; tc < nOut; ++tc) {
                        int i2 = (tc % g.D - g.dr) * s2;
                        int j2 = (tc / g.D - g.dr) * s2;
                        const double* go = gout + ((size_t)n * nOut + tc) * g.oH * g.oW;
                        /* input1 */
                        {
                            int xmin = (x - g.kr - md) / s1, ymin = (y - g.kr - md) / s1;
                            int xmax = (x + g.kr - md) / s1, ymax = (y + g.kr - md) / s1;
                            if (!(xmax < 0 || ymax < 0 || xmin >= g.oW || ymin >= g.oH) &&
                                !(xmin > xmax || ymin > ymax)) {
                                if (xmin < 0) xmin = 0;
                                if (xmax > g.oW - 1) xmax = g.oW - 1;
                                if (ymin < 0) ymin = 0;
                                if (ymax > g.oH - 1) ymax = g.oH - 1;
                                double v2 = at(r2, &g, n, y + j2, x + i2, c);
                                for (int j = ymin; j <= ymax; ++j)
                                    for (int i = xmin; i <= xmax; ++i) sum1 += go[(size_t)j * g.oW + i] * v2;
                            }
                        }
                        /* input2 */
                        {
                            int xmin = (x - g.kr - md - i2) / s1, ymin = (y - g.kr - md - j2) / s1;
                            int xmax = (x + g.kr - md - i2) / s1, ymax = (y + g.kr - md - j2) / s1;
                            if (!(xmax < 0 || ymax < 0 || xmin >= g.oW || ymin >= g.oH) &&
                                !(xmin > xmax || ymin > ymax)) {
                                if (xmin < 0) xmin = 0;
                                if (xmax > g.oW - 1) xmax = g.oW - 1;
                                if (ymin < 0) ymin = 0;
                                if (ymax > g.oH - 1) ymax = g.oH - 1;
                                double v1 = at(r1, &g, n, y - j2, x - i2, c);
                                for (int j = ymin; j <= ymax; ++j)
                                    for (int i = xmin; i <= xmax; ++i) sum2 += go[(size_t)j * g.oW + i] * v1;
                            }
                        }
                    }
                    size_t o = (((size_t)n * C + c) * H + yu) * W + xu;
                    if (g1) g1[o] = sum1 / nelems;
                    if (g2) g2[o] = sum2 / nelems;
                }
    free(r1); free(r2);
    return 0;
}
