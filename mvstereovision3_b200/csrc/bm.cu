// StereoBM (PREFILTER_XSOBEL, minDisparity 0) on sm_100a -- replaces cv::StereoBM::compute behind
// Disparity::bm (reference src/disparity.cpp:18-22, parameters configs/bm.yml).  Semantics: SURVEY.md A.4.
//
//   pre[img][B][H][pitch] u8      : x-Sobel prefiltered images
//   col[B][H][width1][Dp] u16     : column sums  sum_dy |L[y+dy][x'+lofs] - R[y+dy][x'+k]|
//   tex[B][H][W] int              : window sums of |L - cap| (texture)
// The horizontal box sum of `col`, winner-take-all, texture/uniqueness tests and the sub-pixel step are
// fused in one row-marching kernel with the same 8-values-per-lane packed u16x2 layout as SGBM.
#include "mvsv_internal.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;

__global__ void k_bm_prefilter(const uint8_t* __restrict__ img0, const uint8_t* __restrict__ img1, size_t pitch, int W,
                               int H, int cap, uint8_t* __restrict__ o0, uint8_t* __restrict__ o1)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int f = blockIdx.z >> 1, im = blockIdx.z & 1;
    if (x >= W) return;
    const uint8_t* img = (im ? img1 : img0) + (size_t)f * H * pitch;
    uint8_t* out = (im ? o1 : o0) + ((size_t)f * H + y) * pitch + x;
    int res = cap;
    const bool lastOdd = (H & 1) && (y == H - 1);
    if (x > 0 && x < W - 1 && !lastOdd) {
        int ra, rb, rc;   // rows weighted 1, 2, 1
        if ((y & 1) == 0) { ra = y > 0 ? y - 1 : y + 1; rb = y; rc = y + 1; }
        else { const int yb = y - 1; ra = yb; rb = y; rc = (yb < H - 2) ? yb + 2 : yb; }
        const uint8_t *pa = img + (size_t)ra * pitch, *pb = img + (size_t)rb * pitch, *pc = img + (size_t)rc * pitch;
        const int v = ((int)pa[x + 1] - (int)pa[x - 1]) + 2 * ((int)pb[x + 1] - (int)pb[x - 1]) + ((int)pc[x + 1] - (int)pc[x - 1]);
        res = v < -cap ? 0 : (v > cap ? 2 * cap : v + cap);
    }
    *out = (uint8_t)res;
}

// texture: separable window sum of |L - cap|
__global__ void k_bm_tex_col(const uint8_t* __restrict__ preL, size_t pitch, int W, int H, int w2, int cap,
                             uint16_t* __restrict__ tc)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y + w2;
    if (x >= W || y >= H - w2) return;
    const uint8_t* p = preL + (size_t)blockIdx.z * H * pitch;
    int s = 0;
    for (int dy = -w2; dy <= w2; ++dy) s += abs((int)p[(size_t)(y + dy) * pitch + x] - cap);
    tc[((size_t)blockIdx.z * H + y) * W + x] = (uint16_t)s;
}
__global__ void k_bm_tex_row(const uint16_t* __restrict__ tc, int W, int H, int w2, int* __restrict__ tex)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x + w2, y = blockIdx.y + w2;
    if (x >= W - w2 || y >= H - w2) return;
    const size_t base = ((size_t)blockIdx.z * H + y) * W;
    int s = 0;
    for (int dx = -w2; dx <= w2; ++dx) s += tc[base + x + dx];
    tex[base + x] = s;
}

// column sums: lane (x', q) owns the eight disparity indices k = 8q..8q+7 of column x' and marches down the valid
// rows with a sliding sum.  The eight right-image bytes start at an arbitrary byte address, so they are cut out of
// three aligned words with PRMT (the selector is constant per thread); |L - R| is VABSDIFF4 on four bytes at once,
// widened to packed u16x2 for the running sums; one 128-bit store per row.
__device__ __forceinline__ void bm_ad8(const uint8_t* __restrict__ rowL, const unsigned* __restrict__ rowR, unsigned sel,
                                       unsigned (&e)[4])
{
    const unsigned lb = (unsigned)rowL[0] * 0x01010101u;
    const unsigned w0 = rowR[0], w1 = rowR[1], w2 = rowR[2];
    const unsigned dlo = __vabsdiffu4(lb, __byte_perm(w0, w1, sel));
    const unsigned dhi = __vabsdiffu4(lb, __byte_perm(w1, w2, sel));
    e[0] = __byte_perm(dlo, 0, 0x4140); e[1] = __byte_perm(dlo, 0, 0x4342);
    e[2] = __byte_perm(dhi, 0, 0x4140); e[3] = __byte_perm(dhi, 0, 0x4342);
}

template <int G>
__global__ void __launch_bounds__(128) k_bm_colsum(const uint8_t* __restrict__ preL, const uint8_t* __restrict__ preR,
                                                   size_t pitch, int H, int width1, int D, int lofs, int w2,
                                                   uint16_t* __restrict__ col)
{
    constexpr int Dp = 8 * G;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int xp = gtid / G, q = gtid % G;
    if (xp >= width1 || q * 8 >= D) return;
    const size_t fo = (size_t)blockIdx.y * H * pitch;
    const uint8_t* pl = preL + fo + xp + lofs;
    const size_t ra = fo + xp + q * 8;                       // byte address of R[.][x' + 8q]
    const unsigned* pr = reinterpret_cast<const unsigned*>(preR + (ra & ~(size_t)3));
    const unsigned sel = 0x3210u + 0x1111u * (unsigned)(ra & 3);
    const size_t pw = pitch / 4;                             // pitch is a multiple of 16 bytes
    const int bs = 2 * w2 + 1;
    unsigned acc[4] = {0, 0, 0, 0}, e[4];
    for (int y = 0; y < bs; ++y) {
        bm_ad8(pl + (size_t)y * pitch, pr + (size_t)y * pw, sel, e);
        acc[0] += e[0]; acc[1] += e[1]; acc[2] += e[2]; acc[3] += e[3];
    }
    uint16_t* out = col + ((size_t)blockIdx.y * H * width1 + xp) * Dp + q * 8;
    for (int y = w2; y < H - w2; ++y) {
        st128(out + (size_t)y * width1 * Dp, make_uint4(acc[0], acc[1], acc[2], acc[3]));
        if (y + 1 < H - w2) {
            unsigned f[4];
            bm_ad8(pl + (size_t)(y + 1 + w2) * pitch, pr + (size_t)(y + 1 + w2) * pw, sel, e);
            bm_ad8(pl + (size_t)(y - w2) * pitch, pr + (size_t)(y - w2) * pw, sel, f);
            acc[0] += e[0] - f[0]; acc[1] += e[1] - f[1]; acc[2] += e[2] - f[2]; acc[3] += e[3] - f[3];
        }
    }
}

__device__ __forceinline__ unsigned pick16b(const unsigned (&R)[4], int idx)
{
    const int w = (idx >> 1) & 3;
    unsigned r = w == 0 ? R[0] : w == 1 ? R[1] : w == 2 ? R[2] : R[3];
    return (idx & 1) ? (r >> 16) : (r & 0xffffu);
}

struct BmArgs {
    const uint16_t* col; const int* tex; int16_t* disp;
    int W, H, width1, D, Dp, lofs, w2, texThr, uniq, B;
};

template <int G>
__global__ void __launch_bounds__(128) k_bm_wta(BmArgs a)
{
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int vrows = a.H - 2 * a.w2;
    const long long nrows = (long long)a.B * vrows;
    long long r = gtid / G;
    const int q = (int)(gtid % G);
    const bool active = r < nrows;
    if (!active) r = nrows - 1;
    const int f = (int)(r / vrows), y = (int)(r % vrows) + a.w2;
    const size_t rowBase = ((size_t)f * a.H + y) * a.width1 * a.Dp + q * 8;
    const uint16_t* cp = a.col + rowBase;
    const int bs = 2 * a.w2 + 1;
    const unsigned kb = (unsigned)q * 8u;

    uint4 hs = make_uint4(0, 0, 0, 0);
    for (int j = 0; j < bs; ++j) {
        const uint4 v = ld128(cp + (size_t)j * a.Dp);
        hs.x += v.x; hs.y += v.y; hs.z += v.z; hs.w += v.w;
    }
    for (int xp = a.w2; xp < a.width1 - a.w2; ++xp) {
        unsigned Sf[4] = {hs.x, hs.y, hs.z, hs.w};
        unsigned keys[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const unsigned s = (j & 1) ? (Sf[j >> 1] >> 16) : (Sf[j >> 1] & 0xffffu);
            keys[j] = ((int)(kb + j) < a.D) ? ((s << 16) | (kb + j)) : 0xffffffffu;
        }
        unsigned key = min(min(min(keys[0], keys[1]), min(keys[2], keys[3])), min(min(keys[4], keys[5]), min(keys[6], keys[7])));
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(FULL, key, o, G));
        const int minsad = (int)(key >> 16), mind = (int)(key & 0xffffu);
        const int X = xp + a.lofs;
        bool reject = false;
        if (a.uniq > 0) {
            const int thresh = minsad + (minsad * a.uniq / 100);
            bool bad = false;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = (int)kb + j;
                const int s = (int)((j & 1) ? (Sf[j >> 1] >> 16) : (Sf[j >> 1] & 0xffffu));
                bad |= (k < a.D) && (k < mind - 1 || k > mind + 1) && (s <= thresh);
            }
            const unsigned b = __ballot_sync(FULL, bad);
            const unsigned gmask = (G == 32) ? FULL : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));
            reject = (b & gmask) != 0u;
        }
        const int ip = (mind + 1 < a.D) ? mind + 1 : a.D - 2;
        const int in = (mind > 0) ? mind - 1 : 1;
        unsigned vp = pick16b(Sf, ip & 7), vn = pick16b(Sf, in & 7);
        if (G > 1) {
            vp = __shfl_sync(FULL, vp, ip >> 3, G);
            vn = __shfl_sync(FULL, vn, in >> 3, G);
        }
        if (active && q == 0) {
            const size_t oi = ((size_t)f * a.H + y) * a.W + X;
            if (!reject && a.tex[oi] >= a.texThr) {
                const int p = (int)vp, n = (int)vn;
                const int den = p + n - 2 * minsad + abs(p - n);
                a.disp[oi] = (int16_t)((((a.D - 1 - mind) * 256) + (den != 0 ? (p - n) * 256 / den : 0) + 15) >> 4);
            }
        }
        if (xp + 1 < a.width1 - a.w2) {
            const uint4 nx = ld128(cp + (size_t)(xp + 1 + a.w2) * a.Dp);
            const uint4 od = ld128(cp + (size_t)(xp - a.w2) * a.Dp);
            hs.x += nx.x - od.x; hs.y += nx.y - od.y; hs.z += nx.z - od.z; hs.w += nx.w - od.w;
        }
    }
}

__global__ void k_fill16(int16_t* p, size_t n, int16_t v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

template <int G>
void launch_wta(mvsv_ctx* c, const BmArgs& a, int B)
{
    const long long threads = (long long)B * (c->H - 2 * a.w2) * G;
    k_bm_wta<G><<<(unsigned)((threads + 127) / 128), 128, 0, c->stream>>>(a);
}

}  // namespace

void launch_bm(mvsv_ctx* c, int B)
{
    const BmNorm& n = c->bm;
    const int W = c->W, H = c->H;
    const size_t npx = (size_t)B * W * H;
    { KernelTimer kt(c, KID_FILL); k_fill16<<<(unsigned)((npx + 255) / 256), 256, 0, c->stream>>>(c->disp, npx, (int16_t)n.FILT); }
    {
        dim3 blk(128), grd((W + 127) / 128, H, 2 * B);
        KernelTimer kt(c, KID_BM_PREFILTER);
        k_bm_prefilter<<<grd, blk, 0, c->stream>>>(c->rect[0], c->rect[1], c->pitch, W, H, n.cap, c->bm_pre[0], c->bm_pre[1]);
    }
    const bool any = !(n.lofs >= W || n.width1 < 1) && (H - 2 * n.w2 > 0) && (n.width1 - 2 * n.w2 > 0);
    if (!any) return;
    {
        dim3 blk(128), grd((W + 127) / 128, H - 2 * n.w2, B);
        { KernelTimer kt(c, KID_BM_TEX); k_bm_tex_col<<<grd, blk, 0, c->stream>>>(c->bm_pre[0], c->pitch, W, H, n.w2, n.cap, c->bm_tex); }
        dim3 grd2((W - 2 * n.w2 + 127) / 128, H - 2 * n.w2, B);
        { KernelTimer kt(c, KID_BM_TEX); k_bm_tex_row<<<grd2, blk, 0, c->stream>>>(c->bm_tex, W, H, n.w2, c->bm_tex2); }
    }
    {
        const long long threads = (long long)n.width1 * n.G;
        dim3 grd((unsigned)((threads + 127) / 128), B);
        KernelTimer kt(c, KID_BM_COLSUM);
        switch (n.G) {
            case 2: k_bm_colsum<2><<<grd, 128, 0, c->stream>>>(c->bm_pre[0], c->bm_pre[1], c->pitch, H, n.width1, n.D, n.lofs, n.w2, c->bm_col); break;
            case 4: k_bm_colsum<4><<<grd, 128, 0, c->stream>>>(c->bm_pre[0], c->bm_pre[1], c->pitch, H, n.width1, n.D, n.lofs, n.w2, c->bm_col); break;
            case 8: k_bm_colsum<8><<<grd, 128, 0, c->stream>>>(c->bm_pre[0], c->bm_pre[1], c->pitch, H, n.width1, n.D, n.lofs, n.w2, c->bm_col); break;
            case 16: k_bm_colsum<16><<<grd, 128, 0, c->stream>>>(c->bm_pre[0], c->bm_pre[1], c->pitch, H, n.width1, n.D, n.lofs, n.w2, c->bm_col); break;
            default: k_bm_colsum<32><<<grd, 128, 0, c->stream>>>(c->bm_pre[0], c->bm_pre[1], c->pitch, H, n.width1, n.D, n.lofs, n.w2, c->bm_col); break;
        }
    }
    BmArgs a;
    a.col = c->bm_col; a.tex = c->bm_tex2; a.disp = c->disp; a.W = W; a.H = H; a.width1 = n.width1; a.D = n.D; a.Dp = n.Dp;
    a.lofs = n.lofs; a.w2 = n.w2; a.texThr = n.tex; a.uniq = n.uniq; a.B = B;
    KernelTimer kt(c, KID_BM_WTA);
    switch (n.G) {
        case 2: launch_wta<2>(c, a, B); break;
        case 4: launch_wta<4>(c, a, B); break;
        case 8: launch_wta<8>(c, a, B); break;
        case 16: launch_wta<16>(c, a, B); break;
        default: launch_wta<32>(c, a, B); break;
    }
}
