/*
 * mvsv_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C, scalar restatement of the arithmetic that mvStereoVision3's hot
 * path performs through OpenCV.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this file's
 * library; the product (libmvsv.so) never links or calls it.
 *
 * The arithmetic itself lives in a third-party, un-vendored dependency of the
 * reference: OpenCV (un-pinned 3.x in the reference's Makefile:27-35; pinned
 * here to cv2 4.13.0, the build that ships in this image).  The reference has
 * no tests or golden vectors of its own, so this oracle
 * is pinned against outputs of cv2 4.13.0 on the same inputs:
 *   - live, in tests/test_oracle_vs_cv2.py (cv2 is importable in this image,
 *     here and on the GPU box), and
 *   - through committed fixtures tests/golden/*.npz made by
 *     tests/golden/make_golden.py (cv2 outputs, not oracle outputs).
 * Parity w.r.t. the reference's OWN tests is therefore "unpinned" (it has
 * none); parity w.r.t. the library the reference calls is pinned bit-exactly.
 *
 * Reference call sites each function follows (paths relative to the reference
 * root):
 *   orc_remap        src/Stereosystem.cpp:252-256  cv::remap(INTER_LINEAR, CV_32FC1 maps) + ROI crop
 *   orc_sgbm         src/disparity.cpp:6-10        StereoSGBM::compute, wiring src/disparity.cpp:83-95
 *   orc_bm           src/disparity.cpp:18-22       StereoBM::compute (configs/bm.yml)
 *   orc_median3      (inside StereoSGBM::compute)  medianBlur(disp, 3)
 *   orc_speckle      (inside compute)              filterSpeckles
 *   orc_mean         src/utility.cpp:265-285       Utility::calcMeanDisparity
 *   orc_reproject    src/utility.cpp:176-200,242-262  calcCoordinate / dmap2pcl
 *   orc_dmap_values  src/utility.cpp:224-240       calcDMapValues
 *   orc_tm           src/disparity.cpp:25-58       Disparity::tm (matchTemplate TM_CCORR_NORMED + minMaxLoc)
 *   orc_resize       src/Stereosystem.cpp:294-295  cv::resize(roi, dst, Size(0,0), factor, factor), CV_8UC1
 *   orc_rectify_maps src/Stereosystem.cpp:214-217  cv::initUndistortRectifyMap(K, D, R, P, size, CV_32FC1)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_MAX_COST 32767

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------------- */
/* remap: src/Stereosystem.cpp:252-253.  5-bit fixed-point bilinear, constant */
/* zero border.  maps are float (CV_32FC1), full-frame; dst is the cropped    */
/* ROI (src/Stereosystem.cpp:255-256).                                        */
/* ------------------------------------------------------------------------- */
static inline int sat_rint_i32(float v)
{
    double r = nearbyint((double)v); /* default FE_TONEAREST == ties-to-even */
    if (r >= 2147483647.0) return 2147483647;
    if (r <= -2147483648.0) return (-2147483647 - 1);
    return (int)r;
}
static inline int sat_i16(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

void orc_remap(const uint8_t* src, int H, int W, size_t sstride,
               const float* mapx, const float* mapy, size_t mstride_elems,
               int roi_x, int roi_y, int roi_w, int roi_h,
               uint8_t* dst, size_t dstride)
{
    for (int y = 0; y < roi_h; ++y) {
        for (int x = 0; x < roi_w; ++x) {
            size_t mi = (size_t)(y + roi_y) * mstride_elems + (size_t)(x + roi_x);
            int ix = sat_rint_i32(mapx[mi] * 32.0f);
            int iy = sat_rint_i32(mapy[mi] * 32.0f);
            int sx = sat_i16(ix >> 5), sy = sat_i16(iy >> 5);
            int fx = ix & 31, fy = iy & 31;
            int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
            int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
            int p00 = 0, p01 = 0, p10 = 0, p11 = 0;
            if (sy >= 0 && sy < H) {
                if (sx >= 0 && sx < W) p00 = src[(size_t)sy * sstride + sx];
                if (sx + 1 >= 0 && sx + 1 < W) p01 = src[(size_t)sy * sstride + sx + 1];
            }
            if (sy + 1 >= 0 && sy + 1 < H) {
                if (sx >= 0 && sx < W) p10 = src[(size_t)(sy + 1) * sstride + sx];
                if (sx + 1 >= 0 && sx + 1 < W) p11 = src[(size_t)(sy + 1) * sstride + sx + 1];
            }
            int v = (p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + 16384) >> 15;
            dst[(size_t)y * dstride + x] = (uint8_t)iclamp(v, 0, 255);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* StereoSGBM::compute: src/disparity.cpp:8                                   */
/* ------------------------------------------------------------------------- */
typedef struct {
    int minDisp, numDisp, blockSize, disp12MaxDiff, preFilterCap, uniquenessRatio;
    int speckleWindowSize, speckleRange, mode; /* field order == Disparity::sgbmParameters, inc/disparity.h:17-27 */
    int P1, P2;                                /* never set by the reference (src/disparity.cpp:83-90) -> 0 -> 2/5 */
} orc_sgbm_params;

typedef struct {
    int minD, D, bs, SW2, SH2, ftzero, uniq, d12, P1, P2, maxD, minX1, maxX1, W1, INV, mode;
} sgbm_norm;

static void sgbm_normalise(const orc_sgbm_params* p, int W, sgbm_norm* n)
{
    n->minD = p->minDisp;
    n->D = p->numDisp;
    n->bs = p->blockSize > 0 ? p->blockSize : 5;
    n->SW2 = n->SH2 = n->bs / 2;
    n->ftzero = imax(p->preFilterCap, 15) | 1;
    n->uniq = p->uniquenessRatio >= 0 ? p->uniquenessRatio : 10;
    n->d12 = p->disp12MaxDiff > 0 ? p->disp12MaxDiff : 1;
    n->P1 = p->P1 > 0 ? p->P1 : 2;
    n->P2 = imax(p->P2 > 0 ? p->P2 : 5, n->P1 + 1);
    n->maxD = n->minD + n->D;
    n->minX1 = imax(n->maxD, 0);
    n->maxX1 = W + imin(n->minD, 0);
    n->W1 = n->maxX1 - n->minX1;
    n->INV = (n->minD - 1) * 16;
    n->mode = p->mode == 1 ? 1 : 0; /* src/disparity.cpp:92-95 */
}

/* the two "channels" of one image row: clipped x-Sobel and raw intensity */
static void row_channels(const uint8_t* img, int H, int W, size_t stride, int y, int ftzero,
                         int* sob, int* raw)
{
    const uint8_t* r1 = img + (size_t)y * stride;
    const uint8_t* r0 = img + (size_t)imax(y - 1, 0) * stride;
    const uint8_t* r2 = img + (size_t)imin(y + 1, H - 1) * stride;
    for (int x = 1; x < W - 1; ++x) {
        int v = 2 * ((int)r1[x + 1] - (int)r1[x - 1]) + ((int)r0[x + 1] - (int)r0[x - 1]) +
                ((int)r2[x + 1] - (int)r2[x - 1]);
        sob[x] = iclamp(v, -ftzero, ftzero) + ftzero;
        raw[x] = r1[x];
    }
    sob[0] = sob[W - 1] = raw[0] = raw[W - 1] = ftzero;
}

static inline int lo_of(const int* a, int x, int W)
{
    int u = a[x];
    int l = x > 0 ? (u + a[x - 1]) / 2 : u;
    int r = x < W - 1 ? (u + a[x + 1]) / 2 : u;
    return imin(u, imin(l, r));
}
static inline int hi_of(const int* a, int x, int W)
{
    int u = a[x];
    int l = x > 0 ? (u + a[x - 1]) / 2 : u;
    int r = x < W - 1 ? (u + a[x + 1]) / 2 : u;
    return imax(u, imax(l, r));
}
static inline int bt(const int* a, const int* b, int x, int xr, int W)
{
    int u = a[x], v = b[xr];
    int c0 = imax(0, imax(u - hi_of(b, xr, W), lo_of(b, xr, W) - u));
    int c1 = imax(0, imax(v - hi_of(a, x, W), lo_of(a, x, W) - v));
    return imin(c0, c1);
}

/* one recurrence step. Lp == NULL means predecessor is outside the cost domain. */
static inline void path_step(const int16_t* Cp, const int16_t* Lp, int16_t* Ln, int D, int P1, int P2)
{
    if (!Lp) {
        for (int k = 0; k < D; ++k) Ln[k] = Cp[k];
        return;
    }
    int m = ORC_MAX_COST;
    for (int k = 0; k < D; ++k) m = imin(m, Lp[k]);
    for (int k = 0; k < D; ++k) {
        int a = Lp[k];
        int b = (k > 0 ? Lp[k - 1] : ORC_MAX_COST) + P1;
        int c = (k < D - 1 ? Lp[k + 1] : ORC_MAX_COST) + P1;
        int v = Cp[k] + imin(imin(a, b), imin(c, m + P2)) - m;
        Ln[k] = (int16_t)v;
    }
}

static void scan_dir(const int16_t* C, int16_t* S, int H, int W1, int D, int dx, int dy, int P1, int P2)
{
    /* predecessor of (x,y) is (x+dx, y+dy); L for the whole volume of this direction */
    size_t row = (size_t)W1 * D;
    int16_t* L = (int16_t*)malloc(sizeof(int16_t) * row * 2);
    int16_t* Lprev = L;
    int16_t* Lcur = L + row;
    int y0 = dy < 0 ? 0 : (dy > 0 ? H - 1 : 0), ystep = dy > 0 ? -1 : 1;
    for (int yi = 0, y = y0; yi < H; ++yi, y += ystep) {
        /* x order: predecessor first if it is in the same row */
        int x0 = dx > 0 ? W1 - 1 : 0, xstep = dx > 0 ? -1 : 1;
        for (int xi = 0, x = x0; xi < W1; ++xi, x += xstep) {
            int px = x + dx, py = y + dy;
            const int16_t* Lp = NULL;
            if (px >= 0 && px < W1 && py >= 0 && py < H)
                Lp = (dy == 0 ? Lcur : Lprev) + (size_t)px * D;
            int16_t* Ln = Lcur + (size_t)x * D;
            const int16_t* Cp = C + ((size_t)y * W1 + x) * D;
            path_step(Cp, Lp, Ln, D, P1, P2);
            int16_t* Sp = S + ((size_t)y * W1 + x) * D;
            for (int k = 0; k < D; ++k) {
                int s = Sp[k] + Ln[k];
                Sp[k] = (int16_t)(s > ORC_MAX_COST ? ORC_MAX_COST : s);
            }
        }
        int16_t* t = Lprev; Lprev = Lcur; Lcur = t;
    }
    free(L);
}

void orc_median3(const int16_t* in, int16_t* out, int H, int W);
void orc_speckle(int16_t* img, int H, int W, int newVal, int maxSize, int maxDiff);

/* Cost volume only: C[y][xi][k], int16 with wrap.  Cout must hold H*W1*D. Returns W1 (<=0: nothing). */
int orc_sgbm_cost(const uint8_t* left, const uint8_t* right, int H, int W, size_t lstride, size_t rstride,
                  const orc_sgbm_params* p, int16_t* Cout)
{
    sgbm_norm n;
    sgbm_normalise(p, W, &n);
    if (n.W1 <= 0) return n.W1;
    int D = n.D, W1 = n.W1;
    size_t row = (size_t)W1 * D;
    uint16_t* pix16 = (uint16_t*)malloc(sizeof(uint16_t) * row * H);
    int* sl = (int*)malloc(sizeof(int) * W * 4);
    int *sobL = sl, *rawL = sl + W, *sobR = sl + 2 * W, *rawR = sl + 3 * W;
    for (int y = 0; y < H; ++y) {
        row_channels(left, H, W, lstride, y, n.ftzero, sobL, rawL);
        row_channels(right, H, W, rstride, y, n.ftzero, sobR, rawR);
        for (int xi = 0; xi < W1; ++xi) {
            int x = xi + n.minX1;
            for (int k = 0; k < D; ++k) {
                int d = k + n.minD;
                int c = bt(sobL, sobR, x, x - d, W) + (bt(rawL, rawR, x, x - d, W) >> 2);
                pix16[(size_t)y * row + (size_t)xi * D + k] = (uint16_t)c;
            }
        }
    }
    /* separable clamped box sum, int16 wrap */
    uint16_t* hs = (uint16_t*)malloc(sizeof(uint16_t) * row * H);
    for (int y = 0; y < H; ++y)
        for (int xi = 0; xi < W1; ++xi)
            for (int k = 0; k < D; ++k) {
                unsigned s = 0;
                for (int dx = -n.SW2; dx <= n.SW2; ++dx)
                    s += pix16[(size_t)y * row + (size_t)iclamp(xi + dx, 0, W1 - 1) * D + k];
                hs[(size_t)y * row + (size_t)xi * D + k] = (uint16_t)s;
            }
    for (int y = 0; y < H; ++y)
        for (size_t i = 0; i < row; ++i) {
            unsigned s = 0;
            for (int dy = -n.SH2; dy <= n.SH2; ++dy) s += hs[(size_t)iclamp(y + dy, 0, H - 1) * row + i];
            Cout[(size_t)y * row + i] = (int16_t)(uint16_t)s;
        }
    free(hs); free(pix16); free(sl);
    return W1;
}

/*
 * Full SGBM.  disp: H x W int16 (dstride in elements).  Optional outputs (may be NULL):
 *   Cout, Sout : H*W1*D int16 volumes,  raw_out : H x W disparity before median/speckle.
 * Returns 0, or -1 on invalid parameters.
 */
int orc_sgbm(const uint8_t* left, const uint8_t* right, int H, int W, size_t lstride, size_t rstride,
             const orc_sgbm_params* p, int16_t* disp, size_t dstride,
             int16_t* Cout, int16_t* Sout, int16_t* raw_out)
{
    sgbm_norm n;
    sgbm_normalise(p, W, &n);
    if (n.D <= 0) return -1;
    int16_t* raw = (int16_t*)malloc(sizeof(int16_t) * (size_t)H * W);
    for (size_t i = 0; i < (size_t)H * W; ++i) raw[i] = (int16_t)n.INV;
    if (n.W1 > 0) {
        int D = n.D, W1 = n.W1;
        size_t vol = (size_t)H * W1 * D;
        int16_t* C = (int16_t*)malloc(sizeof(int16_t) * vol);
        int16_t* S = (int16_t*)calloc(vol, sizeof(int16_t));
        orc_sgbm_cost(left, right, H, W, lstride, rstride, p, C);
        /* (dx,dy) = offset of the predecessor */
        scan_dir(C, S, H, W1, D, -1, 0, n.P1, n.P2);
        scan_dir(C, S, H, W1, D, -1, -1, n.P1, n.P2);
        scan_dir(C, S, H, W1, D, 0, -1, n.P1, n.P2);
        scan_dir(C, S, H, W1, D, +1, -1, n.P1, n.P2);
        scan_dir(C, S, H, W1, D, +1, 0, n.P1, n.P2);
        if (n.mode == 1) {
            scan_dir(C, S, H, W1, D, -1, +1, n.P1, n.P2);
            scan_dir(C, S, H, W1, D, 0, +1, n.P1, n.P2);
            scan_dir(C, S, H, W1, D, +1, +1, n.P1, n.P2);
        }
        if (Cout) memcpy(Cout, C, sizeof(int16_t) * vol);
        if (Sout) memcpy(Sout, S, sizeof(int16_t) * vol);

        int* disp2 = (int*)malloc(sizeof(int) * W * 2);
        int* disp2cost = disp2 + W;
        for (int y = 0; y < H; ++y) {
            int16_t* dr = raw + (size_t)y * W;
            for (int x = 0; x < W; ++x) { disp2[x] = n.INV; disp2cost[x] = ORC_MAX_COST; }
            for (int xi = W1 - 1; xi >= 0; --xi) {
                const int16_t* Sp = S + ((size_t)y * W1 + xi) * D;
                int best = -1, minS = ORC_MAX_COST;
                for (int k = 0; k < D; ++k)
                    if (Sp[k] < minS) { minS = Sp[k]; best = k; }
                int k;
                for (k = 0; k < D; ++k)
                    if (Sp[k] * (100 - n.uniq) < minS * 100 && abs(k - best) > 1) break;
                if (k < D) continue;
                int v;
                if (best >= 0) {
                    int x2 = xi + n.minX1 - best - n.minD;
                    if (disp2cost[x2] > minS) { disp2cost[x2] = minS; disp2[x2] = best + n.minD; }
                }
                if (best > 0 && best < D - 1) {
                    int den = imax(Sp[best - 1] + Sp[best + 1] - 2 * Sp[best], 1);
                    v = 16 * best + ((Sp[best - 1] - Sp[best + 1]) * 16 + den) / (2 * den);
                } else
                    v = 16 * best;
                dr[xi + n.minX1] = (int16_t)(v + 16 * n.minD);
            }
            for (int x = n.minX1; x < n.maxX1; ++x) {
                int d1 = dr[x];
                if (d1 == n.INV) continue;
                int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
                int _x = x - _d, x_ = x - d_;
                if (0 <= _x && _x < W && disp2[_x] >= n.minD && abs(disp2[_x] - _d) > n.d12 &&
                    0 <= x_ && x_ < W && disp2[x_] >= n.minD && abs(disp2[x_] - d_) > n.d12)
                    dr[x] = (int16_t)n.INV;
            }
        }
        free(disp2); free(C); free(S);
    }
    if (raw_out) memcpy(raw_out, raw, sizeof(int16_t) * (size_t)H * W);
    int16_t* med = (int16_t*)malloc(sizeof(int16_t) * (size_t)H * W);
    orc_median3(raw, med, H, W);
    if (p->speckleWindowSize > 0) orc_speckle(med, H, W, n.INV, p->speckleWindowSize, 16 * p->speckleRange);
    for (int y = 0; y < H; ++y) memcpy(disp + (size_t)y * dstride, med + (size_t)y * W, sizeof(int16_t) * W);
    free(med); free(raw);
    return 0;
}

/* medianBlur(int16, 3), replicate border */
void orc_median3(const int16_t* in, int16_t* out, int H, int W)
{
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int16_t v[9];
            int n = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx)
                    v[n++] = in[(size_t)iclamp(y + dy, 0, H - 1) * W + iclamp(x + dx, 0, W - 1)];
            for (int i = 1; i < 9; ++i) { /* insertion sort */
                int16_t t = v[i]; int j = i - 1;
                while (j >= 0 && v[j] > t) { v[j + 1] = v[j]; --j; }
                v[j + 1] = t;
            }
            out[(size_t)y * W + x] = v[4];
        }
}

/* filterSpeckles: components of {px != newVal} under 4-neighbour |diff|<=maxDiff; size<=maxSize -> newVal */
void orc_speckle(int16_t* img, int H, int W, int newVal, int maxSize, int maxDiff)
{
    size_t N = (size_t)H * W;
    int* label = (int*)calloc(N, sizeof(int));
    int* stack = (int*)malloc(sizeof(int) * N);
    int* comp = (int*)malloc(sizeof(int) * N);
    int cur = 0;
    for (size_t s = 0; s < N; ++s) {
        if (img[s] == newVal || label[s]) continue;
        ++cur;
        int sp = 0, cnt = 0;
        stack[sp++] = (int)s; label[s] = cur;
        while (sp) {
            int p = stack[--sp];
            comp[cnt++] = p;
            int py = p / W, px = p % W, v = img[p];
            int nb[4], nn = 0;
            if (px > 0) nb[nn++] = p - 1;
            if (px < W - 1) nb[nn++] = p + 1;
            if (py > 0) nb[nn++] = p - W;
            if (py < H - 1) nb[nn++] = p + W;
            for (int i = 0; i < nn; ++i) {
                int q = nb[i];
                if (img[q] != newVal && !label[q] && abs((int)img[q] - v) <= maxDiff) { label[q] = cur; stack[sp++] = q; }
            }
        }
        if (cnt <= maxSize)
            for (int i = 0; i < cnt; ++i) comp[i] = -comp[i] - 1; /* mark */
        /* apply later: values must stay intact while other components are traced */
        for (int i = 0; i < cnt; ++i)
            if (comp[i] < 0) label[-comp[i] - 1] = -1;
    }
    for (size_t s = 0; s < N; ++s)
        if (label[s] == -1) img[s] = (int16_t)newVal;
    free(label); free(stack); free(comp);
}

/* ------------------------------------------------------------------------- */
/* StereoBM::compute (PREFILTER_XSOBEL): src/disparity.cpp:20, configs/bm.yml */
/* ------------------------------------------------------------------------- */
typedef struct {
    int minDisp, numDisp, blockSize, preFilterCap, textureThreshold, uniquenessRatio;
    int speckleWindowSize, speckleRange, disp12MaxDiff;
} orc_bm_params;

static void bm_prefilter_xsobel(const uint8_t* src, int H, int W, size_t stride, int cap, uint8_t* dst)
{
    /* rows in pairs; reflect-101 at top/bottom; odd H: last row constant cap */
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) dst[(size_t)y * W + x] = (uint8_t)cap;
    for (int y = 0; y + 1 < H; y += 2) {
        int r0 = y > 0 ? y - 1 : y + 1;
        int r3 = y < H - 2 ? y + 2 : y;
        const uint8_t *s0 = src + (size_t)r0 * stride, *s1 = src + (size_t)y * stride;
        const uint8_t *s2 = src + (size_t)(y + 1) * stride, *s3 = src + (size_t)r3 * stride;
        for (int x = 1; x < W - 1; ++x) {
            int d0 = s0[x + 1] - s0[x - 1], d1 = s1[x + 1] - s1[x - 1];
            int d2 = s2[x + 1] - s2[x - 1], d3 = s3[x + 1] - s3[x - 1];
            int v0 = d0 + 2 * d1 + d2, v1 = d1 + 2 * d2 + d3;
            dst[(size_t)y * W + x] = (uint8_t)(v0 < -cap ? 0 : (v0 > cap ? 2 * cap : v0 + cap));
            dst[(size_t)(y + 1) * W + x] = (uint8_t)(v1 < -cap ? 0 : (v1 > cap ? 2 * cap : v1 + cap));
        }
    }
}

int orc_bm(const uint8_t* left, const uint8_t* right, int H, int W, size_t lstride, size_t rstride,
           const orc_bm_params* p, int16_t* disp, size_t dstride, uint8_t* preL_out, uint8_t* preR_out)
{
    int D = p->numDisp, minD = p->minDisp, bs = p->blockSize, cap = p->preFilterCap;
    if (D <= 0 || D % 16 || bs < 5 || bs % 2 == 0 || cap < 1 || cap > 63) return -1;
    /* minDisparity != 0 is outside the contract: the reference never sets it for BM (no bm.yml loader exists,
       trgt/disparityTest.cpp builds the matcher from numDisp/blockSize only) and cv2 4.13 writes a stray
       pixel below the valid rectangle for minD > 0. */
    if (minD != 0) return -1;
    int w2 = bs / 2;
    int FILT = (minD - 1) * 16;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) disp[(size_t)y * dstride + x] = (int16_t)FILT;
    uint8_t* L = (uint8_t*)malloc((size_t)H * W);
    uint8_t* R = (uint8_t*)malloc((size_t)H * W);
    bm_prefilter_xsobel(left, H, W, lstride, cap, L);
    bm_prefilter_xsobel(right, H, W, rstride, cap, R);
    if (preL_out) memcpy(preL_out, L, (size_t)H * W);
    if (preR_out) memcpy(preR_out, R, (size_t)H * W);
    int lofs = imax(D - 1 + minD, 0), rofs = -imin(D - 1 + minD, 0);
    int width1 = W - rofs - D + 1;
    /* valid rectangle: X in [max(0,minD+D-1)+w2, W-w2), y in [w2, H-w2) */
    int x_lo = imax(0, minD + D - 1) + w2, x_hi = W - w2;
    int* sad = (int*)malloc(sizeof(int) * D);
    if (!(lofs >= W || rofs >= W || width1 < 1)) {
        for (int y = w2; y < H - w2; ++y) {
            for (int X = x_lo; X < x_hi; ++X) {
                int x = X - lofs; /* column in the width1 domain */
                if (x < 0 || x >= width1) continue;
                int tex = 0;
                for (int k = 0; k < D; ++k) sad[k] = 0;
                for (int dy = -w2; dy <= w2; ++dy) {
                    int cy = iclamp(y + dy, 0, H - 1);
                    for (int dx = -w2; dx <= w2; ++dx) {
                        int cl = iclamp(x + dx, -lofs, W - lofs - 1) + lofs;
                        int cr = iclamp(x + dx, -rofs, W - rofs - D) + rofs;
                        int lv = L[(size_t)cy * W + cl];
                        tex += abs(lv - cap);
                        const uint8_t* rp = R + (size_t)cy * W + cr;
                        for (int k = 0; k < D; ++k) sad[k] += abs(lv - (int)rp[k]);
                    }
                }
                if (tex < p->textureThreshold) continue;
                int mind = 0, minsad = sad[0];
                for (int k = 1; k < D; ++k)
                    if (sad[k] < minsad) { minsad = sad[k]; mind = k; }
                if (p->uniquenessRatio > 0) {
                    int thresh = minsad + (minsad * p->uniquenessRatio / 100);
                    int k;
                    for (k = 0; k < D; ++k)
                        if ((k < mind - 1 || k > mind + 1) && sad[k] <= thresh) break;
                    if (k < D) continue;
                }
                int pp = mind + 1 < D ? sad[mind + 1] : sad[D - 2];
                int nn = mind > 0 ? sad[mind - 1] : sad[1];
                int den = pp + nn - 2 * minsad + abs(pp - nn);
                int dd = ((D - 1 - mind + minD) * 256 + (den != 0 ? (pp - nn) * 256 / den : 0) + 15) >> 4;
                disp[(size_t)y * dstride + X] = (int16_t)dd;
            }
        }
    }
    free(sad); free(L); free(R);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* consumers: src/utility.cpp:176-285                                         */
/* ------------------------------------------------------------------------- */
float orc_mean(const int16_t* d, size_t stride, int x0, int y0, int w, int h)
{
    int total = 0, n = 0; /* int accumulator and truncating division: src/utility.cpp:267-283 */
    for (int y = y0; y < y0 + h; ++y)
        for (int x = x0; x < x0 + w; ++x) {
            int v = d[(size_t)y * stride + x];
            if (v > 1) { total += v; ++n; }
        }
    if (total == 0 || n == 0) return 0.0f;
    return (float)(total / abs(n));
}

/* xyz: H x W x 3 float; valid[y][x]=1 where value>0 (src/utility.cpp:249). Q row-major 4x4 float */
void orc_reproject(const int16_t* d, size_t stride, int H, int W, const float* Q, float* xyz, uint8_t* valid)
{
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float value = d[(size_t)y * stride + x];
            float* o = xyz + ((size_t)y * W + x) * 3;
            if (!(value > 0)) { o[0] = o[1] = o[2] = 0.f; if (valid) valid[(size_t)y * W + x] = 0; continue; }
            float dd = value / 16;
            float in[4] = {(float)x, (float)y, dd, 1.f}, c[4];
            for (int r = 0; r < 4; ++r) {
                /* cv::Mat_<float> product accumulates in double (gemm); keep float result */
                double acc = 0;
                for (int k = 0; k < 4; ++k) acc += (double)Q[r * 4 + k] * (double)in[k];
                c[r] = (float)acc;
            }
            float X = c[0] / c[3], Y = c[1] / c[3], Z = c[2] / c[3];
            float dist = Z / 1000;
            if (isinf(dist)) Z = 0;
            o[0] = X; o[1] = Y; o[2] = Z;
            if (valid) valid[(size_t)y * W + x] = 1;
        }
}

/* calcDMapValues: src/utility.cpp:224-240. out = {dValue, image_x, image_y} */
void orc_dmap_values(const float c[3], const float* Q, float out[3])
{
    float numerator = Q[2 * 4 + 3] - c[2] * Q[3 * 4 + 3];
    float denominator = c[2] * Q[3 * 4 + 2];
    float dv = numerator / denominator;
    out[1] = c[0] * (dv * Q[3 * 4 + 2] * Q[3 * 4 + 3]) + Q[0 * 4 + 3];
    out[2] = c[1] * (dv * Q[3 * 4 + 2] * Q[3 * 4 + 3]) + Q[1 * 4 + 3];
    out[0] = dv * 16;
}

/* ------------------------------------------------------------------------- */
/* initUndistortRectifyMap: src/Stereosystem.cpp:214-217 (K, 5 distortion      */
/* coefficients k1 k2 p1 p2 k3, rectifying rotation R, projection P; CV_32FC1 */
/* maps).  Double arithmetic, no fused multiply-add (build with               */
/* -ffp-contract=off), results rounded to float like the reference's maps.    */
/* ------------------------------------------------------------------------- */
static void inv3(const double* m, double* o)
{
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[3] * m[8] - m[5] * m[6], c02 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 - m[1] * c01 + m[2] * c02;
    const double d = 1.0 / det;
    o[0] = c00 * d;                          o[1] = (m[2] * m[7] - m[1] * m[8]) * d;  o[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    o[3] = (m[5] * m[6] - m[3] * m[8]) * d;  o[4] = (m[0] * m[8] - m[2] * m[6]) * d;  o[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    o[6] = c02 * d;                          o[7] = (m[1] * m[6] - m[0] * m[7]) * d;  o[8] = (m[0] * m[4] - m[1] * m[3]) * d;
}

void orc_rectify_maps(const double* K, const double* dist5, const double* R, const double* P /* 3x4 */, int W, int H,
                      float* mapx, float* mapy)
{
    double A[9], iR[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double acc = 0;
            for (int k = 0; k < 3; ++k) acc += P[r * 4 + k] * R[k * 3 + c];
            A[r * 3 + c] = acc;
        }
    inv3(A, iR);
    const double k1 = dist5[0], k2 = dist5[1], p1 = dist5[2], p2 = dist5[3], k3 = dist5[4];
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const double _x = j * iR[0] + i * iR[1] + iR[2], _y = j * iR[3] + i * iR[4] + iR[5], _w = j * iR[6] + i * iR[7] + iR[8];
            const double x = _x / _w, y = _y / _w;
            const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
            const double kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2;
            const double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2);
            const double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy;
            mapx[(size_t)i * W + j] = (float)(fx * xd + cx);
            mapy[(size_t)i * W + j] = (float)(fy * yd + cy);
        }
}

/* ------------------------------------------------------------------------- */
/* cv::resize(src, dst, Size(0,0), f, f), CV_8UC1, INTER_LINEAR (the default): */
/* src/Stereosystem.cpp:294-295.  dst must hold orc_resize_dim(w, f) x       */
/* orc_resize_dim(h, f) bytes.  Scale exactly 1/2 takes OpenCV's 2x2 area path. */
/* ------------------------------------------------------------------------- */
#include <float.h>
int orc_resize_dim(int n, double f) { return (int)lrint((double)n * f); }

void orc_resize(const uint8_t* src, int sw, int sh, double f, uint8_t* dst)
{
    const int dw = orc_resize_dim(sw, f), dh = orc_resize_dim(sh, f);
    const double scale = 1.0 / f;
    const long iscale = lrint(scale);
    if (fabs(scale - (double)iscale) < DBL_EPSILON && iscale == 2) {
        for (int dy = 0; dy < dh; ++dy)
            for (int dx = 0; dx < dw; ++dx) {
                const int x0 = 2 * dx, y0 = 2 * dy;
                int sum = 0, cnt = 0;
                for (int y = y0; y < y0 + 2 && y < sh; ++y)
                    for (int x = x0; x < x0 + 2 && x < sw; ++x) { sum += src[(size_t)y * sw + x]; ++cnt; }
                int v;
                if (cnt == 4) v = (sum + 2) >> 2;
                else v = cnt ? (int)lrintf((float)sum / (float)cnt) : 0;
                dst[(size_t)dy * dw + dx] = (uint8_t)v;
            }
        return;
    }
    int* xofs = (int*)malloc(sizeof(int) * (size_t)dw * 3);
    int *a0 = xofs + dw, *a1 = xofs + 2 * dw;
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale - 0.5);
        int sx = (int)floorf(fx);
        fx -= (float)sx;
        if (sx < 0) { sx = 0; fx = 0.f; }
        if (sx >= sw - 1) { sx = sw - 1; fx = 0.f; }
        xofs[dx] = sx;
        a0[dx] = (int)lrintf((1.f - fx) * 2048.f);
        a1[dx] = (int)lrintf(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale - 0.5);
        const int sy = (int)floorf(fy);
        fy -= (float)sy;
        const int b0 = (int)lrintf((1.f - fy) * 2048.f), b1 = (int)lrintf(fy * 2048.f);
        const int y0 = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy), y1 = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
        const uint8_t *r0 = src + (size_t)y0 * sw, *r1 = src + (size_t)y1 * sw;
        for (int dx = 0; dx < dw; ++dx) {
            const int sx = xofs[dx], x1 = sx + 1 < sw ? sx + 1 : sw - 1;
            const int s0 = r0[sx] * a0[dx] + r0[x1] * a1[dx], s1 = r1[sx] * a0[dx] + r1[x1] * a1[dx];
            int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
            dst[(size_t)dy * dw + dx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    free(xofs);
}

/* ------------------------------------------------------------------------- */
/* Disparity::tm: src/disparity.cpp:25-58.  For every pixel (i, j) with        */
/* i < rows-k, j < cols-k the k x k block of the left image at (j, i) is       */
/* matched against the right image's blocks at (j+x, i), x in [0, cols-j-k),   */
/* by normalised cross-correlation N / sqrt(A * B_x); the output byte is the   */
/* first x with the largest score.  A is constant per pixel, N >= 0, so the    */
/* ranking is N^2 / B_x, compared here as exact integers (cv::matchTemplate    */
/* evaluates it in floating point; see tests/test_tm.py for the tolerance).    */
/* ------------------------------------------------------------------------- */
void orc_tm(const uint8_t* left, const uint8_t* right, int W, int H, int k, uint8_t* out)
{
    memset(out, 0, (size_t)W * H);
    if (k < 1 || k >= W || k >= H) return;
    for (int i = 0; i < H - k; ++i)
        for (int j = 0; j < W - k; ++j) {
            uint64_t nb = 0, bb = 1;
            int bx = 0;
            for (int x = 0; x < W - j - k; ++x) {
                uint64_t n = 0, b = 0;
                for (int v = 0; v < k; ++v)
                    for (int u = 0; u < k; ++u) {
                        const uint64_t l = left[(size_t)(i + v) * W + j + u], r = right[(size_t)(i + v) * W + j + x + u];
                        n += l * r;
                        b += r * r;
                    }
                if ((unsigned __int128)(n * n) * bb > (unsigned __int128)(nb * nb) * b) { nb = n; bb = b; bx = x; }
            }
            out[(size_t)i * W + j] = (uint8_t)bx;
        }
}
