/*
 * msg_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see msg_oracle.h).
 *
 * Restates, from their published algorithms, the OpenCV functions that the reference
 * (ShayHulud/opencv-msegment, Java) reaches through org.opencv.imgproc.Imgproc:
 *   - Imgproc.watershed            PictureService.java:909
 *   - Imgproc.connectedComponents  PictureService.java:441-442
 *   - colorByIndexes (Java loop)   PictureService.java:913-936
 * and the two Imgproc functions BASELINE.json's north_star puts on the path although the
 * reference has no call site for them:
 *   - Imgproc.pyrMeanShiftFiltering  (+ pyrDown / pyrUp / dilate helpers)
 *   - Imgproc.floodFill-style region growing == colour-predicate connected components
 * Third-party dependency restated: org.openpnp:opencv:3.4.2-1 (pom.xml:39-43), not vendored
 * under /root/reference.  Pinned against cv2 4.13.0 (tests/golden/gen_golden.py).
 *
 * Written for clarity, not speed: single thread, scalar.
 */
#include "msg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* cvRound: round-half-to-even (SSE cvtsd2si with default MXCSR) */
static inline int cv_round_d(double v) { return (int)lrint(v); }
static inline int cv_round_f(float v) { return (int)lrintf(v); }

static inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}

/* ------------------------------------------------------------------ pyramids (App. A.3) */

void orc_pyr_down_8uc3(const uint8_t* src, size_t sstep, int w, int h, uint8_t* dst, size_t dstep)
{
    static const int k[5] = {1, 4, 6, 4, 1};
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++)
            for (int c = 0; c < 3; c++) {
                int acc = 0;
                for (int a = -2; a <= 2; a++) {
                    int yy = reflect101(2 * y + a, h);
                    for (int b = -2; b <= 2; b++) {
                        int xx = reflect101(2 * x + b, w);
                        acc += k[a + 2] * k[b + 2] * src[(size_t)yy * sstep + 3 * xx + c];
                    }
                }
                dst[(size_t)y * dstep + 3 * x + c] = (uint8_t)((acc + 128) >> 8);
            }
}

/* one separable pyrUp axis: out[2i] = s[i-1] + 6 s[i] + s[i+1], out[2i+1] = 4 (s[i] + s[i+1]),
 * s[-1] := s[1] (reflect-101), s[n] := s[n-1] (replicate).  n==1: s[-1] := s[0]. */
static inline int up_tap(const int* s, int n, int o)
{
    int i = o >> 1;
    int sm = (i - 1 >= 0) ? s[i - 1] : s[n > 1 ? 1 : 0];
    int sp = (i + 1 < n) ? s[i + 1] : s[n - 1];
    return (o & 1) ? 4 * (s[i] + sp) : sm + 6 * s[i] + sp;
}

void orc_pyr_up_8uc3(const uint8_t* src, size_t sstep, int w, int h, uint8_t* dst, size_t dstep, int dw, int dh)
{
    int* col = (int*)malloc(sizeof(int) * (size_t)(h > w ? h : w));
    int* tmp = (int*)malloc(sizeof(int) * (size_t)dh * (size_t)w); /* vertical pass result, per channel */
    for (int c = 0; c < 3; c++) {
        for (int x = 0; x < w; x++) {
            for (int y = 0; y < h; y++) col[y] = src[(size_t)y * sstep + 3 * x + c];
            for (int oy = 0; oy < dh; oy++) tmp[(size_t)oy * w + x] = up_tap(col, h, oy);
        }
        for (int oy = 0; oy < dh; oy++) {
            const int* row = tmp + (size_t)oy * w;
            for (int ox = 0; ox < dw; ox++) {
                int v = up_tap(row, w, ox);
                dst[(size_t)oy * dstep + 3 * ox + c] = (uint8_t)((v + 32) >> 6);
            }
        }
    }
    free(col);
    free(tmp);
}

/* ------------------------------------------------------------------ mean shift (App. A.2) */

static inline int sq(int v) { return v * v; }

/* One level.  The plane S (w x h) is the window [xoff, xoff+w) x [yoff, yoff+h) of a level whose full size is
 * fullw x fullh: centres, clamping and the position sums use GLOBAL coordinates (the rounding of
 * sum * (1/count) depends on the absolute magnitude of the sums).  Whole image: xoff = yoff = 0, full = (w,h). */
static void meanshift_level(const uint8_t* S, size_t sstep, uint8_t* D, size_t dstep, int w, int h,
                            const uint8_t* mask, float sp, int isr2, int max_count, double eps,
                            orc_ms_counters* ct, int xoff, int yoff, int fullw, int fullh)
{
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            if (mask && !mask[(size_t)i * w + j]) continue;
            int x0 = j + xoff, y0 = i + yoff;
            int c0 = S[(size_t)i * sstep + 3 * j], c1 = S[(size_t)i * sstep + 3 * j + 1],
                c2 = S[(size_t)i * sstep + 3 * j + 2];
            if (ct) ct->pixels++;
            for (int iter = 0; iter < max_count; iter++) {
                int minx = cv_round_f((float)x0 - sp), miny = cv_round_f((float)y0 - sp);
                int maxx = cv_round_f((float)x0 + sp), maxy = cv_round_f((float)y0 + sp);
                if (minx < 0) minx = 0;
                if (miny < 0) miny = 0;
                if (maxx > fullw - 1) maxx = fullw - 1;
                if (maxy > fullh - 1) maxy = fullh - 1;
                /* rows / columns outside the stored window do not exist (ROI evaluation: such pixels lie in
                 * the caller's margin and their results are discarded) */
                if (minx < xoff) minx = xoff;
                if (miny < yoff) miny = yoff;
                if (maxx > xoff + w - 1) maxx = xoff + w - 1;
                if (maxy > yoff + h - 1) maxy = yoff + h - 1;
                int s0 = 0, s1 = 0, s2 = 0, count = 0;
                long long sx = 0, sy = 0;
                for (int y = miny; y <= maxy; y++) {
                    const uint8_t* p = S + (size_t)(y - yoff) * sstep + 3 * (size_t)(minx - xoff);
                    int row_count = 0;
                    for (int x = minx; x <= maxx; x++, p += 3) {
                        int t0 = p[0], t1 = p[1], t2 = p[2];
                        if (sq(t0 - c0) + sq(t1 - c1) + sq(t2 - c2) <= isr2) {
                            s0 += t0; s1 += t1; s2 += t2; sx += x; row_count++;
                        }
                    }
                    count += row_count;
                    sy += (long long)y * row_count;
                }
                if (ct) {
                    ct->iterations++;
                    if (maxx >= minx && maxy >= miny)
                        ct->window_tests += (uint64_t)(maxx - minx + 1) * (uint64_t)(maxy - miny + 1);
                    ct->hits += (uint64_t)count;
                }
                if (count == 0) break;
                double icount = 1.0 / count;
                int x1 = cv_round_d(sx * icount), y1 = cv_round_d(sy * icount);
                int n0 = cv_round_d(s0 * icount), n1 = cv_round_d(s1 * icount), n2 = cv_round_d(s2 * icount);
                int stop = (x0 == x1 && y0 == y1) ||
                           (double)(abs(x1 - x0) + abs(y1 - y0) + sq(n0 - c0) + sq(n1 - c1) + sq(n2 - c2)) <= eps;
                x0 = x1; y0 = y1; c0 = n0; c1 = n1; c2 = n2;
                if (ct) {
                    int ddx = abs(x0 - (j + xoff)), ddy = abs(y0 - (i + yoff));
                    uint64_t dr = (uint64_t)(ddx > ddy ? ddx : ddy);
                    if (dr > ct->max_drift) ct->max_drift = dr;
                }
                if (stop) break;
            }
            uint8_t* d = D + (size_t)i * dstep + 3 * j;
            d[0] = (uint8_t)c0; d[1] = (uint8_t)c1; d[2] = (uint8_t)c2;
        }
}

int orc_meanshift_filter(const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h,
                         double sp0, double sr, int max_level, int term_type, int max_count, double eps,
                         orc_ms_counters* ct)
{
    return orc_meanshift_filter_roi(src, sstep, dst, dstep, w, h, 0, 0, w, h, sp0, sr, max_level, term_type,
                                    max_count, eps, ct);
}

int orc_meanshift_filter_roi(const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h,
                             int xoff, int yoff, int full_w, int full_h,
                             double sp0, double sr, int max_level, int term_type, int max_count, double eps,
                             orc_ms_counters* ct)
{
    if (max_level < 0 || max_level > 8 || w <= 0 || h <= 0) return -1; /* OpenCV validates only maxLevel */
    if (xoff < 0 || yoff < 0 || xoff + w > full_w || yoff + h > full_h) return -1;
    if ((xoff | yoff) & ((1 << max_level) - 1)) return -1; /* pyramid phase must match the full image */
    if (!(term_type & ORC_TERM_COUNT)) max_count = 5;
    if (max_count < 1) max_count = 1;
    if (max_count > 100) max_count = 100;
    if (!(term_type & ORC_TERM_EPS)) eps = 1.0;
    if (eps < 0) eps = 0;
    int isr2 = cv_round_d(sr * sr);
    int isr22 = isr2 > 16 ? isr2 : 16;
    if (ct) memset(ct, 0, sizeof(*ct));

    uint8_t* S[9]; uint8_t* D[9]; int W[9], H[9], FW[9], FH[9];
    S[0] = (uint8_t*)src; D[0] = dst; W[0] = w; H[0] = h; FW[0] = full_w; FH[0] = full_h;
    size_t sst[9], dst_[9];
    sst[0] = sstep; dst_[0] = dstep;
    for (int l = 1; l <= max_level; l++) {
        W[l] = (W[l - 1] + 1) / 2; H[l] = (H[l - 1] + 1) / 2;
        FW[l] = (FW[l - 1] + 1) / 2; FH[l] = (FH[l - 1] + 1) / 2;
        sst[l] = dst_[l] = (size_t)3 * W[l];
        S[l] = (uint8_t*)malloc(sst[l] * H[l]);
        D[l] = (uint8_t*)malloc(dst_[l] * H[l]);
        orc_pyr_down_8uc3(S[l - 1], sst[l - 1], W[l - 1], H[l - 1], S[l], sst[l]);
    }
    uint8_t* mask = max_level > 0 ? (uint8_t*)malloc((size_t)w * h) : NULL;
    uint8_t* mtmp = max_level > 0 ? (uint8_t*)malloc((size_t)w * h) : NULL;

    for (int l = max_level; l >= 0; l--) {
        float sp = (float)(sp0 / (double)(1 << l));
        if (sp < 1.f) sp = 1.f;
        const uint8_t* m = NULL;
        int lw = W[l], lh = H[l];
        if (l < max_level) {
            int w1 = W[l + 1], h1 = H[l + 1];
            const uint8_t* P = D[l + 1]; size_t ps = dst_[l + 1];
            orc_pyr_up_8uc3(P, ps, w1, h1, D[l], dst_[l], lw, lh);
            memset(mtmp, 0, (size_t)lw * lh);
            for (int i = 1; i < h1 - 1; i++)
                for (int j = 1; j < w1 - 1; j++) {
                    const uint8_t* c = P + (size_t)i * ps + 3 * j;
                    int flag = 0;
                    for (int dy = -1; dy <= 1; dy++)
                        for (int dx = -1; dx <= 1; dx++) {
                            if (!dx && !dy) continue;
                            const uint8_t* n = P + (size_t)(i + dy) * ps + 3 * (j + dx);
                            if (sq(c[0] - n[0]) + sq(c[1] - n[1]) + sq(c[2] - n[2]) >= isr22) flag = 1;
                        }
                    int my = 2 * i + 1, mx = 2 * j - 1; /* empirical, exact vs cv2 */
                    if (my < lh && mx < lw) mtmp[(size_t)my * lw + mx] = (uint8_t)flag;
                }
            /* dilate 3x3, outside = absent */
            for (int y = 0; y < lh; y++)
                for (int x = 0; x < lw; x++) {
                    uint8_t v = 0;
                    for (int dy = -1; dy <= 1; dy++)
                        for (int dx = -1; dx <= 1; dx++) {
                            int yy = y + dy, xx = x + dx;
                            if (yy < 0 || yy >= lh || xx < 0 || xx >= lw) continue;
                            if (mtmp[(size_t)yy * lw + xx]) v = 1;
                        }
                    mask[(size_t)y * lw + x] = v;
                }
            m = mask;
        }
        meanshift_level(S[l], sst[l], D[l], dst_[l], lw, lh, m, sp, isr2, max_count, eps, ct, xoff >> l, yoff >> l, FW[l], FH[l]);
    }
    for (int l = 1; l <= max_level; l++) { free(S[l]); free(D[l]); }
    free(mask); free(mtmp);
    return 0;
}

/* ------------------------------------------------------------------ union-find labelling (App. A.4) */

static int32_t uf_find(int32_t* p, int32_t a)
{
    int32_t r = a;
    while (p[r] != r) r = p[r];
    while (p[a] != r) { int32_t n = p[a]; p[a] = r; a = n; }
    return r;
}
static void uf_union(int32_t* p, int32_t a, int32_t b)
{
    a = uf_find(p, a); b = uf_find(p, b);
    if (a < b) p[b] = a; else if (b < a) p[a] = b;
}

static inline int color_close(const uint8_t* a, const uint8_t* b, int d)
{
    return abs(a[0] - b[0]) <= d && abs(a[1] - b[1]) <= d && abs(a[2] - b[2]) <= d;
}

int32_t orc_relabel_canonical(int32_t* labels, size_t lstep, int w, int h)
{
    /* map arbitrary positive labels -> 1.. in raster order of first occurrence */
    int32_t maxl = 0;
    for (int y = 0; y < h; y++) {
        const int32_t* r = (const int32_t*)((const char*)labels + (size_t)y * lstep);
        for (int x = 0; x < w; x++) if (r[x] > maxl) maxl = r[x];
    }
    int32_t* map = (int32_t*)calloc((size_t)maxl + 1, sizeof(int32_t));
    int32_t n = 0;
    for (int y = 0; y < h; y++) {
        int32_t* r = (int32_t*)((char*)labels + (size_t)y * lstep);
        for (int x = 0; x < w; x++)
            if (r[x] > 0) {
                if (!map[r[x]]) map[r[x]] = ++n;
                r[x] = map[r[x]];
            }
    }
    free(map);
    return n;
}

int32_t orc_label_regions(const uint8_t* bgr, size_t step, int32_t* labels, size_t lstep, int w, int h, int d)
{
    return orc_label_regions_conn(bgr, step, labels, lstep, w, h, d, 4);
}

int32_t orc_label_regions_conn(const uint8_t* bgr, size_t step, int32_t* labels, size_t lstep, int w, int h, int d,
                               int connectivity)
{
    size_t n = (size_t)w * h;
    int32_t* p = (int32_t*)malloc(n * sizeof(int32_t));
    for (size_t i = 0; i < n; i++) p[i] = (int32_t)i;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* c = bgr + (size_t)y * step + 3 * x;
            if (x > 0 && color_close(c, c - 3, d)) uf_union(p, y * w + x, y * w + x - 1);
            if (y > 0 && color_close(c, c - step, d)) uf_union(p, y * w + x, (y - 1) * w + x);
            if (connectivity == 8 && y > 0) {   /* floodFill with flags & 8: diagonal neighbours too */
                if (x > 0 && color_close(c, c - step - 3, d)) uf_union(p, y * w + x, (y - 1) * w + x - 1);
                if (x < w - 1 && color_close(c, c - step + 3, d)) uf_union(p, y * w + x, (y - 1) * w + x + 1);
            }
        }
    for (int y = 0; y < h; y++) {
        int32_t* r = (int32_t*)((char*)labels + (size_t)y * lstep);
        for (int x = 0; x < w; x++) r[x] = uf_find(p, y * w + x) + 1;
    }
    free(p);
    return orc_relabel_canonical(labels, lstep, w, h);
}

int32_t orc_connected_components(const uint8_t* mask, size_t step, int32_t* labels, size_t lstep, int w, int h,
                                 int connectivity)
{
    size_t n = (size_t)w * h;
    int32_t* p = (int32_t*)malloc(n * sizeof(int32_t));
    for (size_t i = 0; i < n; i++) p[i] = (int32_t)i;
#define FG(yy, xx) (mask[(size_t)(yy) * step + (xx)] != 0)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            if (!FG(y, x)) continue;
            if (x > 0 && FG(y, x - 1)) uf_union(p, y * w + x, y * w + x - 1);
            if (y > 0 && FG(y - 1, x)) uf_union(p, y * w + x, (y - 1) * w + x);
            if (connectivity == 8 && y > 0) {
                if (x > 0 && FG(y - 1, x - 1)) uf_union(p, y * w + x, (y - 1) * w + x - 1);
                if (x < w - 1 && FG(y - 1, x + 1)) uf_union(p, y * w + x, (y - 1) * w + x + 1);
            }
        }
    for (int y = 0; y < h; y++) {
        int32_t* r = (int32_t*)((char*)labels + (size_t)y * lstep);
        for (int x = 0; x < w; x++) r[x] = FG(y, x) ? uf_find(p, y * w + x) + 1 : 0;
    }
#undef FG
    free(p);
    return orc_relabel_canonical(labels, lstep, w, h) + 1;
}

/* ------------------------------------------------------------------ region merge (self-defined spec)
 *
 * Regions = sets of pixels with equal positive label.  Per round:
 *   area[L], sum_c[L] (c = B,G,R);  mean_c[L] = floor((2*sum_c + area) / (2*area))   (0..255)
 *   dist2(L,Q) = sum_c (mean_c[L] - mean_c[Q])^2
 *   every *participating* region L selects, among the regions Q != L 4-adjacent to it, the one with the
 *   lexicographically smallest key (dist2(L,Q), Q); the selection is *accepted* if dist2 <= limit.
 *   All accepted selections are united simultaneously; a united component takes its smallest member label.
 * Phase A (color_dist > 0): every region participates, limit = color_dist^2; rounds until no selection is
 *   accepted (at most ORC_MERGE_MAX_ROUNDS).
 * Phase B (min_size > 0): regions with area < min_size participate, limit = +inf; rounds until no region
 *   participates-with-a-neighbour (at most ORC_MERGE_MAX_ROUNDS).
 * Finally labels are renumbered canonically (raster order of first pixel).
 */
#define ORC_MERGE_MAX_ROUNDS 64

static int merge_round(const uint8_t* bgr, size_t step, int32_t* labels, size_t lstep, int w, int h,
                       int32_t maxl, int64_t size_thr, int64_t dist_limit)
{
    size_t nl = (size_t)maxl + 1;
    int64_t* area = (int64_t*)calloc(nl, sizeof(int64_t));
    int64_t* sum = (int64_t*)calloc(nl * 3, sizeof(int64_t));
    int32_t* mean = (int32_t*)calloc(nl * 3, sizeof(int32_t));
    uint64_t* best = (uint64_t*)malloc(nl * sizeof(uint64_t));
    int32_t* par = (int32_t*)malloc(nl * sizeof(int32_t));
    for (size_t i = 0; i < nl; i++) { best[i] = ~(uint64_t)0; par[i] = (int32_t)i; }
#define LAB(yy, xx) (((const int32_t*)((const char*)labels + (size_t)(yy) * lstep))[xx])
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int32_t L = LAB(y, x);
            if (L <= 0) continue;
            const uint8_t* c = bgr + (size_t)y * step + 3 * x;
            area[L]++; sum[3 * L] += c[0]; sum[3 * L + 1] += c[1]; sum[3 * L + 2] += c[2];
        }
    for (size_t L = 1; L < nl; L++)
        if (area[L] > 0)
            for (int c = 0; c < 3; c++) mean[3 * L + c] = (int32_t)((2 * sum[3 * L + c] + area[L]) / (2 * area[L]));
    static const int dx4[4] = {-1, 1, 0, 0}, dy4[4] = {0, 0, -1, 1};
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int32_t L = LAB(y, x);
            if (L <= 0 || area[L] >= size_thr) continue;
            for (int k = 0; k < 4; k++) {
                int xx = x + dx4[k], yy = y + dy4[k];
                if (xx < 0 || yy < 0 || xx >= w || yy >= h) continue;
                int32_t Q = LAB(yy, xx);
                if (Q <= 0 || Q == L) continue;
                uint64_t d2 = (uint64_t)(sq(mean[3 * L] - mean[3 * Q]) + sq(mean[3 * L + 1] - mean[3 * Q + 1]) +
                                         sq(mean[3 * L + 2] - mean[3 * Q + 2]));
                uint64_t key = (d2 << 32) | (uint32_t)Q;
                if (key < best[L]) best[L] = key;
            }
        }
    int accepted = 0;
    for (size_t L = 1; L < nl; L++) {
        if (best[L] == ~(uint64_t)0) continue;
        int64_t d2 = (int64_t)(best[L] >> 32);
        if (d2 > dist_limit) continue;
        uf_union(par, (int32_t)L, (int32_t)(best[L] & 0xffffffffu));
        accepted++;
    }
    if (accepted)
        for (int y = 0; y < h; y++) {
            int32_t* r = (int32_t*)((char*)labels + (size_t)y * lstep);
            for (int x = 0; x < w; x++)
                if (r[x] > 0) r[x] = uf_find(par, r[x]);
        }
#undef LAB
    free(area); free(sum); free(mean); free(best); free(par);
    return accepted;
}

int32_t orc_merge_regions(const uint8_t* bgr, size_t step, int32_t* labels, size_t lstep, int w, int h,
                          int min_size, int color_dist)
{
    int32_t maxl = 0;
    for (int y = 0; y < h; y++) {
        const int32_t* r = (const int32_t*)((const char*)labels + (size_t)y * lstep);
        for (int x = 0; x < w; x++) if (r[x] > maxl) maxl = r[x];
    }
    const int64_t INF = (int64_t)1 << 40;
    if (color_dist > 0)
        for (int r = 0; r < ORC_MERGE_MAX_ROUNDS; r++)
            if (!merge_round(bgr, step, labels, lstep, w, h, maxl, INF, (int64_t)color_dist * color_dist)) break;
    if (min_size > 0)
        for (int r = 0; r < ORC_MERGE_MAX_ROUNDS; r++)
            if (!merge_round(bgr, step, labels, lstep, w, h, maxl, (int64_t)min_size, INF)) break;
    return orc_relabel_canonical(labels, lstep, w, h);
}

/* ------------------------------------------------------------------ render (PictureService.java:913-936) */

void orc_render_labels(const int32_t* labels, size_t lstep, uint8_t* dst, size_t dstep, int w, int h, int depth,
                       const uint8_t* colors_bgr)
{
    for (int y = 0; y < h; y++) {
        const int32_t* r = (const int32_t*)((const char*)labels + (size_t)y * lstep);
        uint8_t* d = dst + (size_t)y * dstep;
        for (int x = 0; x < w; x++) {
            int32_t L = r[x];
            if (L > 0 && L <= depth) {
                if (colors_bgr) { d[3 * x] = colors_bgr[3 * (L - 1)]; d[3 * x + 1] = colors_bgr[3 * (L - 1) + 1]; d[3 * x + 2] = colors_bgr[3 * (L - 1) + 2]; }
                else d[3 * x] = d[3 * x + 1] = d[3 * x + 2] = 255;
            } else d[3 * x] = d[3 * x + 1] = d[3 * x + 2] = 0;
        }
    }
}

/* ------------------------------------------------------------------ watershed (App. A.1) */

void orc_watershed(const uint8_t* bgr, size_t step, int32_t* markers, size_t mstep, int w, int h)
{
    enum { IN_QUEUE = -2, WSHED = -1 };
    size_t n = (size_t)w * h;
    int32_t* next = (int32_t*)malloc(n * sizeof(int32_t));
    int32_t head[256], tail[256];
    for (int i = 0; i < 256; i++) head[i] = tail[i] = -1;
#define M(yy, xx) (((int32_t*)((char*)markers + (size_t)(yy) * mstep))[xx])
#define PIX(yy, xx) (bgr + (size_t)(yy) * step + 3 * (xx))
#define PUSH(q, pos) do { next[pos] = -1; if (tail[q] < 0) head[q] = pos; else next[tail[q]] = pos; tail[q] = pos; } while (0)
    if (w < 1 || h < 1) { free(next); return; }
    for (int x = 0; x < w; x++) { M(0, x) = WSHED; M(h - 1, x) = WSHED; }
    static const int dxs[4] = {-1, 1, 0, 0}, dys[4] = {0, 0, -1, 1};
    for (int i = 1; i < h - 1; i++) {
        M(i, 0) = WSHED; M(i, w - 1) = WSHED;
        for (int j = 1; j < w - 1; j++) {
            if (M(i, j) < 0) M(i, j) = 0;
            if (M(i, j) != 0) continue;
            int idx = 256;
            for (int k = 0; k < 4; k++) {
                int yy = i + dys[k], xx = j + dxs[k];
                if (M(yy, xx) > 0) {
                    const uint8_t *a = PIX(i, j), *b = PIX(yy, xx);
                    int d0 = abs(a[0] - b[0]), d1 = abs(a[1] - b[1]), d2 = abs(a[2] - b[2]);
                    int t = d0 > d1 ? d0 : d1; if (d2 > t) t = d2;
                    if (t < idx) idx = t;
                }
            }
            if (idx < 256) { PUSH(idx, i * w + j); M(i, j) = IN_QUEUE; }
        }
    }
    int active = 0;
    while (active < 256 && head[active] < 0) active++;
    if (active == 256) { free(next); return; }
    for (;;) {
        if (head[active] < 0) {
            while (active < 256 && head[active] < 0) active++;
            if (active == 256) break;
        }
        int pos = head[active];
        head[active] = next[pos];
        if (head[active] < 0) tail[active] = -1;
        int i = pos / w, j = pos % w;
        int lab = 0;
        for (int k = 0; k < 4; k++) {
            int t = M(i + dys[k], j + dxs[k]);
            if (t > 0) { if (lab == 0) lab = t; else if (t != lab) lab = WSHED; }
        }
        M(i, j) = lab;
        if (lab == WSHED) continue;
        for (int k = 0; k < 4; k++) {
            int yy = i + dys[k], xx = j + dxs[k];
            if (M(yy, xx) == 0) {
                const uint8_t *a = PIX(i, j), *b = PIX(yy, xx);
                int d0 = abs(a[0] - b[0]), d1 = abs(a[1] - b[1]), d2 = abs(a[2] - b[2]);
                int t = d0 > d1 ? d0 : d1; if (d2 > t) t = d2;
                PUSH(t, yy * w + xx);
                if (t < active) active = t;
                M(yy, xx) = IN_QUEUE;
            }
        }
    }
#undef M
#undef PIX
#undef PUSH
    free(next);
}

/* ------------------------------------------------------------------ synthetic image */

static inline uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t synth_hash(uint64_t seedmix, uint32_t a, uint32_t b, uint32_t c)
{
    return splitmix64(seedmix ^ (((uint64_t)a << 40) | ((uint64_t)b << 16) | (uint64_t)c));
}

void orc_synth_bgr(uint8_t* dst, size_t step, int w, int h, uint64_t seed)
{
    uint64_t sm = splitmix64(seed);
    int ncx = (w + 63) / 64, ncy = (h + 63) / 64;
    int* sxs = (int*)malloc(sizeof(int) * (size_t)ncx * ncy);
    int* sys = (int*)malloc(sizeof(int) * (size_t)ncx * ncy);
    uint32_t* col = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)ncx * ncy);
    for (int cy = 0; cy < ncy; cy++)
        for (int cx = 0; cx < ncx; cx++) {
            sxs[cy * ncx + cx] = 64 * cx + (int)(synth_hash(sm, (uint32_t)cx, (uint32_t)cy, 0) % 64);
            sys[cy * ncx + cx] = 64 * cy + (int)(synth_hash(sm, (uint32_t)cx, (uint32_t)cy, 1) % 64);
            col[cy * ncx + cx] = (uint32_t)(synth_hash(sm, (uint32_t)cx, (uint32_t)cy, 2) & 0xFFFFFFu);
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int cx0 = x / 64, cy0 = y / 64;
            long best = -1; uint32_t bc = 0;
            for (int cy = cy0 - 1; cy <= cy0 + 1; cy++)
                for (int cx = cx0 - 1; cx <= cx0 + 1; cx++) {
                    if (cx < 0 || cy < 0 || cx >= ncx || cy >= ncy) continue;
                    long ddx = x - sxs[cy * ncx + cx], ddy = y - sys[cy * ncx + cx];
                    long d = ddx * ddx + ddy * ddy;
                    if (best < 0 || d < best) { best = d; bc = col[cy * ncx + cx]; }
                }
            for (int c = 0; c < 3; c++) {
                int base = (int)((bc >> (8 * c)) & 0xFF);
                int nz = (int)(synth_hash(sm, (uint32_t)x, (uint32_t)y, 16u + (uint32_t)c) % 13) +
                         (int)(synth_hash(sm, (uint32_t)x, (uint32_t)y, 32u + (uint32_t)c) % 13) - 12;
                int v = base + nz;
                dst[(size_t)y * step + 3 * x + c] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
            }
        }
    free(sxs); free(sys); free(col);
}

/* ------------------------------------------------------------------ f2 "next" row: the pre-filters the reference really calls
 * (PictureService.java:323-333 Laplacian sharpen chain, :405/:940 cvtColor BGR2GRAY, :408/:436 medianBlur).
 * All three have exact integer forms, pinned on cv2 by tests/golden/filters.npz. */

/* filter2D(src, lap, CV_32F, K) ; src.convertTo(CV_32F) ; subtract ; convertTo(CV_8U)  ==  saturate(src - sum_K taps * src),
 * correlation with the anchor at the kernel centre, BORDER_REFLECT_101; taps are small integers (the reference's
 * MatOfFloat(1,1,1,1,-8,1,1,1,1) as 9x1 -- the literal reading -- or 3x3 -- the intended one). */
void orc_laplacian_sharpen(const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h,
                           const int8_t* taps, int krows, int kcols)
{
    int ay = krows / 2, ax = kcols / 2;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            for (int c = 0; c < 3; c++) {
                int acc = 0;
                for (int a = 0; a < krows; a++) {
                    int yy = reflect101(y + a - ay, h);
                    for (int b = 0; b < kcols; b++) {
                        int xx = reflect101(x + b - ax, w);
                        acc += taps[a * kcols + b] * src[(size_t)yy * sstep + 3 * xx + c];
                    }
                }
                int v = src[(size_t)y * sstep + 3 * x + c] - acc;
                dst[(size_t)y * dstep + 3 * x + c] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
            }
}

/* cv::medianBlur on 8UC1, odd k, BORDER_REPLICATE: the (k*k/2)-th smallest of the window. */
void orc_median_blur_8uc1(const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h, int k)
{
    int r = k / 2, need = (k * k) / 2 + 1;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int hist[256];
            memset(hist, 0, sizeof(hist));
            for (int a = -r; a <= r; a++) {
                int yy = y + a; yy = yy < 0 ? 0 : (yy >= h ? h - 1 : yy);
                for (int b = -r; b <= r; b++) {
                    int xx = x + b; xx = xx < 0 ? 0 : (xx >= w ? w - 1 : xx);
                    hist[src[(size_t)yy * sstep + xx]]++;
                }
            }
            int acc = 0, v = 0;
            for (; v < 256; v++) { acc += hist[v]; if (acc >= need) break; }
            dst[(size_t)y * dstep + x] = (uint8_t)v;
        }
}

/* cvtColor(COLOR_BGR2GRAY) as cv2 4.13 computes it: (3735 B + 19235 G + 9798 R + 16384) >> 15
 * (OpenCV 3.4.2's scalar path used 14-bit coefficients: +-1 LSB on 0.26 % of colours). */
void orc_bgr2gray(const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* p = src + (size_t)y * sstep + 3 * x;
            dst[(size_t)y * dstep + x] = (uint8_t)((3735 * p[0] + 19235 * p[1] + 9798 * p[2] + 16384) >> 15);
        }
}

/* The same conversion with the coefficients of OpenCV 3.4.2's scalar path -- the version the reference binds
 * (pom.xml:39-43): B2Y = 1868, G2Y = 9617, R2Y = 4899, yuv_shift = 14 ([recalled]: that build is not
 * available offline, so this mode is pinned on the published constants only, not on golden vectors). */
void orc_bgr2gray_342(const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* p = src + (size_t)y * sstep + 3 * x;
            dst[(size_t)y * dstep + x] = (uint8_t)((1868 * p[0] + 9617 * p[1] + 4899 * p[2] + 8192) >> 14);
        }
}

/* ------------------------------------------------------------------------------------------------------------------
 * Shape-method seeds (PictureService.java:416-442): Canny, dilate, subtract.  Restated from the published OpenCV
 * algorithm (imgproc canny.cpp / morph.cpp semantics), pinned on cv2 4.13.0 by tests/golden/seeds.npz.
 * ------------------------------------------------------------------------------------------------------------------ */
static int clampi_(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void orc_canny(const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h, double low_d, double high_d)
{
    if (low_d > high_d) { double t = low_d; low_d = high_d; high_d = t; }
    const int low = (int)floor(low_d), high = (int)floor(high_d);
    const size_t n = (size_t)w * h;
    short* dx = (short*)malloc(n * sizeof(short));
    short* dy = (short*)malloc(n * sizeof(short));
    int* mag = (int*)calloc((size_t)(w + 2) * (h + 2), sizeof(int));       /* zero frame: magnitude outside the image */
    uint8_t* cls = (uint8_t*)calloc(n, 1);                                  /* 0 none, 1 candidate, 2 strong candidate */
    const size_t mp = (size_t)w + 2;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* r0 = src + (size_t)clampi_(y - 1, 0, h - 1) * sstep;
            const uint8_t* r1 = src + (size_t)y * sstep;
            const uint8_t* r2 = src + (size_t)clampi_(y + 1, 0, h - 1) * sstep;
            int xl = clampi_(x - 1, 0, w - 1), xr = clampi_(x + 1, 0, w - 1);
            int gx = (r0[xr] + 2 * r1[xr] + r2[xr]) - (r0[xl] + 2 * r1[xl] + r2[xl]);
            int gy = (r2[xl] + 2 * r2[x] + r2[xr]) - (r0[xl] + 2 * r0[x] + r0[xr]);
            dx[(size_t)y * w + x] = (short)gx;
            dy[(size_t)y * w + x] = (short)gy;
            mag[(size_t)(y + 1) * mp + x + 1] = abs(gx) + abs(gy);
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int* m = mag + (size_t)(y + 1) * mp + x + 1;
            int v = m[0];
            if (v <= low) continue;
            int xs = dx[(size_t)y * w + x], ys = dy[(size_t)y * w + x];
            int ax = abs(xs), ay = abs(ys) << 15;
            int tg22 = ax * 13573;
            int keep;
            if (ay < tg22) keep = v > m[-1] && v >= m[1];
            else {
                int tg67 = tg22 + (ax << 16);
                if (ay > tg67) keep = v > m[-(long)mp] && v >= m[mp];
                else {
                    int s = (xs ^ ys) < 0 ? -1 : 1;
                    keep = v > m[-(long)mp - s] && v > m[(long)mp + s];
                }
            }
            if (keep) cls[(size_t)y * w + x] = v > high ? 2 : 1;
        }
    /* hysteresis: flood from strong candidates over 8-connected candidates (order-independent result) */
    int* stack = (int*)malloc((n ? n : 1) * sizeof(int));
    size_t top = 0;
    for (int y = 0; y < h; y++) memset(dst + (size_t)y * dstep, 0, (size_t)w);
    for (size_t i = 0; i < n; i++)
        if (cls[i] == 2) { stack[top++] = (int)i; dst[(i / w) * dstep + i % w] = 255; }
    while (top) {
        int i = stack[--top], y = i / w, x = i % w;
        for (int yy = y - 1; yy <= y + 1; yy++)
            for (int xx = x - 1; xx <= x + 1; xx++) {
                if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                if (cls[(size_t)yy * w + xx] && !dst[(size_t)yy * dstep + xx]) {
                    dst[(size_t)yy * dstep + xx] = 255;
                    stack[top++] = yy * w + xx;
                }
            }
    }
    free(stack); free(cls); free(mag); free(dx); free(dy);
}

void orc_dilate_rect(const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h, int kw, int kh)
{
    const int ax = kw / 2, ay = kh / 2;
    uint8_t* out = (uint8_t*)malloc((size_t)w * h);                         /* src and dst may alias */
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int best = 0;
            for (int yy = y - ay; yy < y - ay + kh; yy++)
                for (int xx = x - ax; xx < x - ax + kw; xx++)
                    if (yy >= 0 && yy < h && xx >= 0 && xx < w && src[(size_t)yy * sstep + xx] > best) best = src[(size_t)yy * sstep + xx];
            out[(size_t)y * w + x] = (uint8_t)best;
        }
    for (int y = 0; y < h; y++) memcpy(dst + (size_t)y * dstep, out + (size_t)y * w, (size_t)w);
    free(out);
}

void orc_subtract_u8(const uint8_t* a, size_t astep, const uint8_t* b, size_t bstep, uint8_t* dst, size_t dstep, int w, int h)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int v = (int)a[(size_t)y * astep + x] - (int)b[(size_t)y * bstep + x];
            dst[(size_t)y * dstep + x] = (uint8_t)(v < 0 ? 0 : v);
        }
}

/* ------------------------------------------------------------------------------------------------------------------
 * Colour-method seeds (PictureService.java:338-366, :938-943, :1018-1023): Otsu threshold,
 * chamfer distance transform, min-max normalisation, peak threshold + dilate, contour labelling, background disc.
 * Restated from the published OpenCV algorithms (imgproc thresh.cpp / distransform.cpp / contours / drawing.cpp
 * semantics), pinned on cv2 4.13.0 by tests/golden/color_seeds.npz.
 * ------------------------------------------------------------------------------------------------------------------ */

/* cv::threshold(..., THRESH_OTSU) threshold selection on 8UC1 (getThreshVal_Otsu_8u): double arithmetic, sequential
 * over the 256 bins; the order of the operations is part of the specification. */
int orc_otsu_threshold(const uint8_t* src, size_t sstep, int w, int h)
{
    int hist[256];
    memset(hist, 0, sizeof(hist));
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) hist[src[(size_t)y * sstep + x]]++;
    double mu = 0, scale = 1. / ((double)w * h);
    for (int i = 0; i < 256; i++) mu += i * (double)hist[i];
    mu *= scale;
    double mu1 = 0, q1 = 0, max_sigma = 0, max_val = 0;
    for (int i = 0; i < 256; i++) {
        double p_i = hist[i] * scale;
        mu1 *= q1;
        q1 += p_i;
        double q2 = 1. - q1;
        double lo = q1 < q2 ? q1 : q2, hi = q1 > q2 ? q1 : q2;
        if (lo < 1.1920928955078125e-07 || hi > 1. - 1.1920928955078125e-07) continue;
        mu1 = (mu1 + i * p_i) / q1;
        double mu2 = (mu - q1 * mu1) / q2;
        double sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    return (int)max_val;
}

/* cv::threshold(src 8UC1, thresh, maxval, THRESH_BINARY): dst = src > thresh ? maxval : 0 */
void orc_threshold_binary_u8(const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h, int thresh, int maxval)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) dst[(size_t)y * dstep + x] = (uint8_t)(src[(size_t)y * sstep + x] > thresh ? maxval : 0);
}

/* cv::distanceTransform(src 8UC1, dst 32F, DIST_L2, 5): two-pass 5x5 chamfer with the metrics a = 1, b = 1.4,
 * c = 2.1969 accumulated in float32, frame of "infinite" (FLT_MAX) pixels.  This is the arithmetic of the IPP-backed cv2
 * 4.13 build the oracle is pinned on (OpenCV's own fallback uses 16.16 fixed point instead and differs by ~4e-6 relative),
 * established from its outputs:
 *   - rows are stored as float; the backward pass is the plain float recurrence;
 *   - in the FORWARD pass the running value of the in-row recurrence d[x] = min(t[x], d[x-1] + 1) is not rounded between
 *     the pixels of an aligned group of four columns: the candidates t[x] are exact float + float sums, the running value
 *     carries the exact sum through up to three additions of 1 and only the stored pixel is rounded; a group (x % 4 == 0)
 *     starts from the stored, rounded left neighbour, and the columns x >= ((w - 2) / 4) * 4 round at every pixel.
 *     (Exact sums fit a double.)  The difference to rounding everywhere is 1 ulp on pixels where a horizontal step crosses
 *     32, 64, 128 ... on an exact rounding tie -- rare on small images, common at 1080p.
 * cv2 == this function bit for bit on 1200 / 1200 random, blob, Otsu and sparse images up to 420 x 420 and on the Otsu masks
 * of the 1080p / 4K synthetic frames.  A source without zero pixels gives FLT_MAX everywhere, as cv2 does. */
void orc_distance_transform_l2_5(const uint8_t* src, size_t sstep, float* dst, size_t dstep, int w, int h)
{
    const double A = 1.0f, B = 1.4f, C = 2.1969f;
    const float INF = 3.402823466e+38f;
    const size_t tp = (size_t)w + 4;
    const int lim = ((w - 2) / 4) * 4;
    float* t = (float*)malloc(tp * ((size_t)h + 4) * sizeof(float));
    for (size_t i = 0; i < tp * ((size_t)h + 4); i++) t[i] = INF;
#define T_(y, x) t[(size_t)((y) + 2) * tp + (size_t)((x) + 2)]
#define MIN_(v) do { double q_ = (v); if (q_ < t0) t0 = q_; } while (0)
    for (int y = 0; y < h; y++) {
        double run = INF;                                  /* unrounded value of the pixel to the left */
        for (int x = 0; x < w; x++) {
            double t0;
            if (x % 4 == 0 || x >= lim) run = (double)T_(y, x - 1);
            if (!src[(size_t)y * sstep + x]) t0 = 0;
            else {
                t0 = (double)T_(y - 2, x - 1) + C;
                MIN_((double)T_(y - 2, x + 1) + C);
                MIN_((double)T_(y - 1, x - 2) + C);
                MIN_((double)T_(y - 1, x - 1) + B);
                MIN_((double)T_(y - 1, x) + A);
                MIN_((double)T_(y - 1, x + 1) + B);
                MIN_((double)T_(y - 1, x + 2) + C);
                MIN_(run + A);
            }
            run = t0;
            T_(y, x) = (float)t0;
        }
    }
    for (int y = h - 1; y >= 0; y--)
        for (int x = w - 1; x >= 0; x--) {
            float t0 = T_(y, x);
            if (t0 > (float)A) {
#undef MIN_
#define MIN_(v) do { float q_ = (v); if (q_ < t0) t0 = q_; } while (0)
                MIN_(T_(y + 2, x + 1) + (float)C);
                MIN_(T_(y + 2, x - 1) + (float)C);
                MIN_(T_(y + 1, x + 2) + (float)C);
                MIN_(T_(y + 1, x + 1) + (float)B);
                MIN_(T_(y + 1, x) + (float)A);
                MIN_(T_(y + 1, x - 1) + (float)B);
                MIN_(T_(y + 1, x - 2) + (float)C);
                MIN_(T_(y, x + 1) + (float)A);
                T_(y, x) = t0;
            }
            *(float*)((uint8_t*)dst + (size_t)y * dstep + (size_t)x * 4) = t0;
        }
#undef T_
#undef MIN_
    free(t);
}

/* The same call as a NON-IPP OpenCV build computes it (distransform.cpp, distanceTransform_5x5: what the openpnp 3.4.2 natives
 * the reference binds would run): 16.16 fixed point.  Metrics HV = cvRound(1 * 65536), DIAG = cvRound(1.4 * 65536),
 * LONG = cvRound(2.1969 * 65536); a 2-pixel border of DIST_MAX = INT_MAX >> 2 around an int plane; forward pass = min over the
 * eight forward mask cells, backward pass the same mirrored (skipped where the value is already <= HV: no candidate can be
 * smaller), output = (float)(t * (1.f / 65536)) with t clamped to DIST_MAX = UINT_MAX - LONG (cv2 4.13: a pixel no zero pixel
 * reaches -- only possible when the source has none -- reads 65533.805).  Integer min-plus: order independent.
 * Pinned on cv2 4.13.0 with cv2.ipp.setUseIPP(False) (tests/golden/gen_dt_fixed.py -> dt_fixed.npz). */
void orc_distance_transform_l2_5_fixed(const uint8_t* src, size_t sstep, float* dst, size_t dstep, int w, int h)
{
    const int HV = 65536, DIAG = 91750, LONG_ = 143976;      /* cvRound(1.4f * 65536) = 91750, cvRound(2.1969f * 65536) = 143976 */
    const long long DIST_MAX = 4294967295ll - LONG_;
    const float scale = 1.f / 65536;
    const size_t tp = (size_t)w + 4;
    long long* t = (long long*)malloc(tp * ((size_t)h + 4) * sizeof(long long));     /* 64-bit: sums cannot wrap */
    for (size_t i = 0; i < tp * ((size_t)h + 4); i++) t[i] = DIST_MAX;
#define T_(y, x) t[(size_t)((y) + 2) * tp + (size_t)((x) + 2)]
#define MIN_(v) do { long long q_ = (v); if (q_ < t0) t0 = q_; } while (0)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            long long t0;
            if (!src[(size_t)y * sstep + x]) t0 = 0;
            else {
                t0 = T_(y - 2, x - 1) + LONG_;
                MIN_(T_(y - 2, x + 1) + LONG_);
                MIN_(T_(y - 1, x - 2) + LONG_);
                MIN_(T_(y - 1, x - 1) + DIAG);
                MIN_(T_(y - 1, x) + HV);
                MIN_(T_(y - 1, x + 1) + DIAG);
                MIN_(T_(y - 1, x + 2) + LONG_);
                MIN_(T_(y, x - 1) + HV);
            }
            T_(y, x) = t0;
        }
    for (int y = h - 1; y >= 0; y--)
        for (int x = w - 1; x >= 0; x--) {
            long long t0 = T_(y, x);
            if (t0 > HV) {
                MIN_(T_(y + 2, x + 1) + LONG_);
                MIN_(T_(y + 2, x - 1) + LONG_);
                MIN_(T_(y + 1, x + 2) + LONG_);
                MIN_(T_(y + 1, x + 1) + DIAG);
                MIN_(T_(y + 1, x) + HV);
                MIN_(T_(y + 1, x - 1) + DIAG);
                MIN_(T_(y + 1, x - 2) + LONG_);
                MIN_(T_(y, x + 1) + HV);
                T_(y, x) = t0;
            }
            if (t0 > DIST_MAX) t0 = DIST_MAX;
            *(float*)((uint8_t*)dst + (size_t)y * dstep + (size_t)x * 4) = (float)((float)(unsigned int)t0 * scale);
        }
#undef T_
#undef MIN_
    free(t);
}

/* Core.normalize(src 32F, dst, 0, 1, NORM_MINMAX) (PictureService.java:1021): scale = 1 * (1 / (max - min)) in double,
 * shift = 0 - min * scale; dst = src * (float)scale + (float)shift in float; a constant image becomes all zero. */
void orc_normalize_minmax01_f32(const float* src, size_t sstep, float* dst, size_t dstep, int w, int h)
{
    double smin = INFINITY, smax = -INFINITY;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float v = *(const float*)((const uint8_t*)src + (size_t)y * sstep + (size_t)x * 4);
            if (v < smin) smin = v;
            if (v > smax) smax = v;
        }
    double scale = (1. - 0.) * (smax - smin > 2.220446049250313e-16 ? 1. / (smax - smin) : 0.);
    double shift = 0. - smin * scale;
    float a = (float)scale, b = (float)shift;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float v = *(const float*)((const uint8_t*)src + (size_t)y * sstep + (size_t)x * 4);
            *(float*)((uint8_t*)dst + (size_t)y * dstep + (size_t)x * 4) = fmaf(v, a, b);   /* cv2's vector path fuses */
        }
}

/* threshold(dist, .4, 1., THRESH_BINARY) + dilate(3x3 ones) + convertTo(CV_8U) (PictureService.java:348-356): 0 / 1 */
void orc_peaks_u8(const float* nrm, size_t sstep, uint8_t* dst, size_t dstep, int w, int h, double thresh)
{
    const float th = (float)thresh;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int on = 0;
            for (int yy = y - 1; yy <= y + 1; yy++)
                for (int xx = x - 1; xx <= x + 1; xx++)
                    if (yy >= 0 && yy < h && xx >= 0 && xx < w &&
                        *(const float*)((const uint8_t*)nrm + (size_t)yy * sstep + (size_t)xx * 4) > th) on = 1;
            dst[(size_t)y * dstep + x] = (uint8_t)on;
        }
}

/* findContours(mask, RETR_CCOMP, CHAIN_APPROX_NONE) followed by drawContours(markers, contours, i, i + 1, FILLED, 8,
 * hierarchy, INT_MAX) for i = 0 .. n-1 (PictureService.java:360-364), restated without contour tracing:
 *   - outer contours = 8-connected components of the non-zero pixels, holes = 4-connected components of the zero pixels
 *     that do not touch the image border; a hole belongs to the component of the pixel left of its first (raster) pixel;
 *   - contour index: components in REVERSE raster order of their first pixel, each followed by its holes in reverse
 *     raster order of their first pixel;
 *   - painting contour i (outer) covers the component's pixels; painting a hole covers the hole, everything enclosed
 *     by it, and the component's pixels 4-adjacent to the hole; later (larger) indices overwrite earlier ones.
 * Returns the number of contours. */
static int32_t cc_generic(const uint8_t* mask, size_t step, int w, int h, int want_nonzero, int conn8, int32_t* par)
{
    const size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; i++) par[i] = (int32_t)i;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            if ((mask[(size_t)y * step + x] != 0) != want_nonzero) continue;
            int32_t i = y * w + x;
            if (x > 0 && (mask[(size_t)y * step + x - 1] != 0) == want_nonzero) uf_union(par, i, i - 1);
            if (y > 0) {
                const uint8_t* up = mask + (size_t)(y - 1) * step;
                if ((up[x] != 0) == want_nonzero) uf_union(par, i, i - w);
                if (conn8 && x > 0 && (up[x - 1] != 0) == want_nonzero) uf_union(par, i, i - w - 1);
                if (conn8 && x + 1 < w && (up[x + 1] != 0) == want_nonzero) uf_union(par, i, i - w + 1);
            }
        }
    for (size_t i = 0; i < n; i++) par[i] = uf_find(par, (int32_t)i);
    return 0;
}

int32_t orc_contour_markers(const uint8_t* mask, size_t step, int32_t* markers, size_t mstep, int w, int h)
{
    const size_t n = (size_t)w * h;
    int32_t* fg = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));    /* root = first pixel of the component */
    int32_t* bg = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
    uint8_t* open = (uint8_t*)calloc(n ? n : 1, 1);                     /* per bg root: touches the image border */
    int32_t* nholes = (int32_t*)calloc(n ? n : 1, sizeof(int32_t));     /* per fg root */
    int32_t* idx = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));     /* per root (fg or hole): contour index */
    int32_t* seen = (int32_t*)calloc(n ? n : 1, sizeof(int32_t));
    cc_generic(mask, step, w, h, 1, 1, fg);
    cc_generic(mask, step, w, h, 0, 0, bg);
#define ISFG(i) (mask[(size_t)((i) / w) * step + (size_t)((i) % w)] != 0)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            if ((y == 0 || y == h - 1 || x == 0 || x == w - 1) && !ISFG(y * w + x)) open[bg[y * w + x]] = 1;
    int32_t total = 0;
    for (size_t i = 0; i < n; i++) {
        if (ISFG(i)) { if (fg[i] == (int32_t)i) total++; }
        else if (bg[i] == (int32_t)i && !open[i]) { nholes[fg[i - 1]]++; total++; }
    }
    int32_t acc = 0;
    for (size_t i = 0; i < n; i++)                       /* components in discovery order */
        if (ISFG(i) && fg[i] == (int32_t)i) { acc += 1 + nholes[i]; idx[i] = total - acc; }
    for (size_t i = 0; i < n; i++)                       /* holes in discovery order */
        if (!ISFG(i) && bg[i] == (int32_t)i && !open[i]) {
            int32_t c = fg[i - 1];
            idx[i] = idx[c] + nholes[c] - seen[c];
            seen[c]++;
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int32_t i = y * w + x, out = 0;
            /* walk outwards: the outermost enclosing hole wins */
            int32_t hole = -1, cur = i;
            int cur_fg = ISFG(i);
            for (;;) {
                if (cur_fg) {                              /* component -> the background region left of its first pixel */
                    int32_t r = fg[cur];
                    if (r % w == 0) break;
                    int32_t b = bg[r - 1];
                    if (open[b]) break;
                    cur = b; cur_fg = 0;
                } else {                                   /* hole (or open background) -> its component */
                    int32_t b = bg[cur];
                    if (open[b]) break;
                    hole = b;
                    cur = b - 1; cur_fg = 1;
                }
            }
            if (hole >= 0) out = idx[hole] + 1;
            else if (ISFG(i)) {
                int32_t best = idx[fg[i]];
                const int nx[4] = {x - 1, x + 1, x, x}, ny[4] = {y, y, y - 1, y + 1};
                for (int k = 0; k < 4; k++)
                    if (nx[k] >= 0 && nx[k] < w && ny[k] >= 0 && ny[k] < h && !ISFG(ny[k] * w + nx[k])) {
                        int32_t b = bg[ny[k] * w + nx[k]];
                        if (!open[b] && idx[b] > best) best = idx[b];
                    }
                out = best + 1;
            }
            *(int32_t*)((uint8_t*)markers + (size_t)y * mstep + (size_t)x * 4) = out;
        }
#undef ISFG
    free(seen); free(idx); free(nholes); free(open); free(bg); free(fg);
    return total;
}

/* Imgproc.circle(markers, centre, radius, value, FILLED) on 32SC1 (PictureService.java:366): OpenCV's midpoint circle,
 * horizontal spans clipped to the image. */
void orc_circle_filled_i32(int32_t* img, size_t step, int w, int h, int cx, int cy, int radius, int32_t value)
{
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
        int ys[4] = {cy - dy, cy + dy, cy - dx, cy + dx};
        int xa[4] = {cx - dx, cx - dx, cx - dy, cx - dy}, xb[4] = {cx + dx, cx + dx, cx + dy, cx + dy};
        for (int k = 0; k < 4; k++) {
            if (ys[k] < 0 || ys[k] >= h) continue;
            int a = xa[k] < 0 ? 0 : xa[k], b = xb[k] >= w ? w - 1 : xb[k];
            for (int x = a; x <= b; x++) *(int32_t*)((uint8_t*)img + (size_t)ys[k] * step + (size_t)x * 4) = value;
        }
        dy++;
        err += plus;
        plus += 2;
        int mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
}

/* The reference's first pipeline step (PictureService.java:309-318): pure white pixels become black. */
void orc_white_to_black(const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* p = src + (size_t)y * sstep + 3 * x;
            uint8_t* q = dst + (size_t)y * dstep + 3 * x;
            int white = p[0] == 255 && p[1] == 255 && p[2] == 255;
            q[0] = white ? 0 : p[0]; q[1] = white ? 0 : p[1]; q[2] = white ? 0 : p[2];
        }
}

/* cv::bilateralFilter on 8UC1 / 8UC3 (PictureService.java:490; f2 row), BORDER_REFLECT_101: the scalar OpenCV loop
 * (float accumulation in window order, colour weights (float)exp(i*i*coeff) by |delta| summed over channels, circular
 * window r <= radius).  cv2's SIMD path fuses the multiply-adds, hence the +-1 LSB tolerance against cv2. */
void orc_bilateral_filter(const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h, int cn, int d,
                          double sigma_color, double sigma_space)
{
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    double gc = -0.5 / (sigma_color * sigma_color), gs = -0.5 / (sigma_space * sigma_space);
    int radius = d <= 0 ? cv_round_d(sigma_space * 1.5) : d / 2;
    if (radius < 1) radius = 1;
    int dd = 2 * radius + 1, maxk = 0;
    float* cw = (float*)malloc(sizeof(float) * 256 * cn);
    float* sw = (float*)malloc(sizeof(float) * dd * dd);
    int* oy = (int*)malloc(sizeof(int) * dd * dd);
    int* ox = (int*)malloc(sizeof(int) * dd * dd);
    for (int i = 0; i < 256 * cn; i++) cw[i] = (float)exp(i * i * gc);
    for (int i = -radius; i <= radius; i++)
        for (int j = -radius; j <= radius; j++) {
            double r = sqrt((double)i * i + (double)j * j);
            if (r > radius) continue;
            sw[maxk] = (float)exp(r * r * gs);
            oy[maxk] = i; ox[maxk] = j; maxk++;
        }
    uint8_t* out = (uint8_t*)malloc((size_t)w * h * cn);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* p0 = src + (size_t)y * sstep + (size_t)x * cn;
            float sum[3] = {0, 0, 0}, wsum = 0;
            for (int k = 0; k < maxk; k++) {
                const uint8_t* p = src + (size_t)reflect101(y + oy[k], h) * sstep + (size_t)reflect101(x + ox[k], w) * cn;
                int diff = 0;
                for (int c = 0; c < cn; c++) diff += abs((int)p[c] - (int)p0[c]);
                float wgt = sw[k] * cw[diff];
                for (int c = 0; c < cn; c++) sum[c] += p[c] * wgt;
                wsum += wgt;
            }
            if (cn == 1) out[((size_t)y * w + x)] = (uint8_t)cv_round_f(sum[0] / wsum);
            else {
                float inv = 1.f / wsum;
                for (int c = 0; c < cn; c++) out[((size_t)y * w + x) * cn + c] = (uint8_t)cv_round_f(sum[c] * inv);
            }
        }
    for (int y = 0; y < h; y++) memcpy(dst + (size_t)y * dstep, out + (size_t)y * w * cn, (size_t)w * cn);
    free(out); free(ox); free(oy); free(sw); free(cw);
}
