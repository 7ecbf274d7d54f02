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

// Row-marching: a thread owns one column of BM_ROWS rows (an even count, so the row pairs of the reference's loop
// never straddle two CTAs) and keeps the horizontal differences of the previous rows in registers.
constexpr int BM_ROWS = 16;

__global__ void __launch_bounds__(128)
k_bm_prefilter(const uint8_t* __restrict__ img0, const uint8_t* __restrict__ img1, size_t pitch, int W, int H, int cap,
               uint8_t* __restrict__ o0, uint8_t* __restrict__ o1)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y0 = blockIdx.y * BM_ROWS, y1 = min(y0 + BM_ROWS, H);
    const int f = blockIdx.z >> 1, im = blockIdx.z & 1;
    if (x >= W) return;
    const uint8_t* img = (im ? img1 : img0) + (size_t)f * H * pitch;
    uint8_t* out = (im ? o1 : o0) + (size_t)f * H * pitch + x;
    const bool inner = x > 0 && x < W - 1;
    auto dx = [&](int r) -> int {                  // horizontal difference of row r (clamped into the image by the callers)
        const uint8_t* p = img + (size_t)r * pitch + x;
        return inner ? (int)p[1] - (int)p[-1] : 0;
    };
    auto clip = [&](int v) -> int { return v < -cap ? 0 : (v > cap ? 2 * cap : v + cap); };
    // rows are processed in pairs (y, y + 1), y even: out[y] = d[r0] + 2 d[y] + d[y+1] with r0 = y-1 (y+1 at the top),
    // out[y+1] = d[y] + 2 d[y+1] + d[r3] with r3 = y+2 (y at the bottom); an odd last row is the constant `cap`
    for (int y = y0; y < y1; y += 2) {
        if (y + 1 >= H) { out[(size_t)y * pitch] = (uint8_t)cap; break; }
        const int d0 = dx(y > 0 ? y - 1 : y + 1), d1 = dx(y), d2 = dx(y + 1), d3 = dx(y < H - 2 ? y + 2 : y);
        out[(size_t)y * pitch] = (uint8_t)(inner ? clip(d0 + 2 * d1 + d2) : cap);
        out[(size_t)(y + 1) * pitch] = (uint8_t)(inner ? clip(d1 + 2 * d2 + d3) : cap);
    }
}

// texture: separable window sum of |L - cap|; the column pass slides down BM_ROWS rows per thread
__global__ void __launch_bounds__(128)
k_bm_tex_col(const uint8_t* __restrict__ preL, size_t pitch, int W, int H, int w2, int cap, uint16_t* __restrict__ tc)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int ya = blockIdx.y * BM_ROWS + w2, yb = min(ya + BM_ROWS, H - w2);
    if (x >= W || ya >= yb) return;
    const uint8_t* p = preL + (size_t)blockIdx.z * H * pitch + x;
    int s = 0;
    for (int dy = -w2; dy <= w2; ++dy) s += abs((int)p[(size_t)(ya + dy) * pitch] - cap);
    uint16_t* o = tc + (size_t)blockIdx.z * H * W + x;
    for (int y = ya; y < yb; ++y) {
        o[(size_t)y * W] = (uint16_t)s;
        if (y + 1 < yb) s += abs((int)p[(size_t)(y + 1 + w2) * pitch] - cap) - abs((int)p[(size_t)(y - w2) * pitch] - cap);
    }
}
// row pass: a warp slides along a row segment; lanes own BM_SEG consecutive columns each
constexpr int BM_SEG = 8;
__global__ void __launch_bounds__(128)
k_bm_tex_row(const uint16_t* __restrict__ tc, int W, int H, int w2, int* __restrict__ tex)
{
    const int xa = (blockIdx.x * blockDim.x + threadIdx.x) * BM_SEG + w2, y = blockIdx.y + w2;
    if (xa >= W - w2 || y >= H - w2) return;
    const size_t base = ((size_t)blockIdx.z * H + y) * W;
    const int xb = min(xa + BM_SEG, W - w2);
    int s = 0;
    for (int dx = -w2; dx <= w2; ++dx) s += tc[base + xa + dx];
    for (int x = xa; x < xb; ++x) {
        tex[base + x] = s;
        if (x + 1 < xb) s += (int)tc[base + x + 1 + w2] - (int)tc[base + x - w2];
    }
}

// column sums: lane (x', q) owns the eight disparity indices k = 8q..8q+7 of column x' and marches down the valid
// rows with a sliding sum.  The eight right-image bytes start at an arbitrary byte address, so they are cut out of
// three aligned words with PRMT (the selector is constant per thread); |L - R| is VABSDIFF4 on four bytes at once,
// widened to packed u16x2 for the running sums; one 128-bit store per row.
__device__ __forceinline__ uint2 bm_ad8(const uint8_t* __restrict__ rowL, const unsigned* __restrict__ rowR, unsigned sel)
{
    const unsigned lb = (unsigned)rowL[0] * 0x01010101u;
    const unsigned w0 = rowR[0], w1 = rowR[1], w2 = rowR[2];
    return make_uint2(__vabsdiffu4(lb, __byte_perm(w0, w1, sel)), __vabsdiffu4(lb, __byte_perm(w1, w2, sel)));
}
// acc (packed u16x2) += the eight bytes of e, -= the eight bytes of f
__device__ __forceinline__ void bm_acc(unsigned (&acc)[4], const uint2& e, const uint2& f)
{
    acc[0] += __byte_perm(e.x, 0, 0x4140) - __byte_perm(f.x, 0, 0x4140); acc[1] += __byte_perm(e.x, 0, 0x4342) - __byte_perm(f.x, 0, 0x4342);
    acc[2] += __byte_perm(e.y, 0, 0x4140) - __byte_perm(f.y, 0, 0x4140); acc[3] += __byte_perm(e.y, 0, 0x4342) - __byte_perm(f.y, 0, 0x4342);
}

// The absolute differences of the last blockSize rows stay in a per-thread shared-memory ring (8 bytes per row), so
// the row that leaves the window is not evaluated a second time.
// Threads are numbered (column, octet) with exactly D / 8 octets per column -- not a power of two as in the row scans, so
// no lane idles at D = 80 -- and consecutive threads store consecutive 16-byte pieces of the volume.
__global__ void __launch_bounds__(128) k_bm_colsum(const uint8_t* __restrict__ preL, const uint8_t* __restrict__ preR,
                                                   size_t pitch, int H, int width1, int D, int lofs, int w2,
                                                   uint16_t* __restrict__ col)
{
    extern __shared__ uint2 bm_ring[];          // [bs][128]
    const int Dp = D;           // pixel stride of the column-sum volume: numDisp (a multiple of 16)
    const int nOct = D >> 3;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int xp = gtid / nOct, q = gtid - xp * nOct;
    if (xp >= width1) return;
    const size_t fo = (size_t)blockIdx.y * H * pitch;
    const uint8_t* pl = preL + fo + xp + lofs;
    const size_t ra = fo + xp + q * 8;                       // byte address of R[.][x' + 8q]
    const unsigned* pr = reinterpret_cast<const unsigned*>(preR + (ra & ~(size_t)3));
    const unsigned sel = 0x3210u + 0x1111u * (unsigned)(ra & 3);
    const size_t pw = pitch / 4;                             // pitch is a multiple of 16 bytes
    const int bs = 2 * w2 + 1;
    uint2* ring = bm_ring + threadIdx.x;
    unsigned acc[4] = {0, 0, 0, 0};
    for (int y = 0; y < bs; ++y) {
        const uint2 e = bm_ad8(pl + (size_t)y * pitch, pr + (size_t)y * pw, sel);
        ring[y * 128] = e;
        bm_acc(acc, e, make_uint2(0u, 0u));
    }
    uint16_t* out = col + ((size_t)blockIdx.y * H * width1 + xp) * Dp + q * 8;
    int slot = 0;                                            // ring slot of the oldest row (y - w2)
#pragma unroll 2
    for (int y = w2; y < H - w2; ++y) {
        st128(out + (size_t)y * width1 * Dp, make_uint4(acc[0], acc[1], acc[2], acc[3]));
        if (y + 1 < H - w2) {
            const uint2 e = bm_ad8(pl + (size_t)(y + 1 + w2) * pitch, pr + (size_t)(y + 1 + w2) * pw, sel);
            const uint2 f = ring[slot * 128];
            ring[slot * 128] = e;
            bm_acc(acc, e, f);
            if (++slot == bs) slot = 0;
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

// Lanes are numbered (row, octet) with exactly nOct = D / 8 octets per row -- 10 at bm.yml, so a warp holds three rows
// (30 lanes) instead of two rows in two power-of-two groups of 16 with 6 idle lanes each.  The minimum over a row's
// lanes is a shfl_down ladder bounded by the segment, then a broadcast from the segment's first lane.
__global__ void __launch_bounds__(128) k_bm_wta(BmArgs a)
{
    const int nOct = a.D >> 3, RW = 32 / nOct;                 // octets per row, rows per warp
    const int lane = threadIdx.x & 31, rw = lane / nOct, q = lane - rw * nOct;
    const int vrows = a.H - 2 * a.w2;
    const long long nrows = (long long)a.B * vrows;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long r = warp * RW + rw;
    const bool active = rw < RW && r < nrows;
    if (r >= nrows) r = nrows - 1;
    const int base = min(rw, RW - 1) * nOct;                   // first lane of this row's segment
    const int f = (int)(r / vrows), y = (int)(r % vrows) + a.w2;
    const size_t rowBase = ((size_t)f * a.H + y) * a.width1 * a.Dp + q * 8;
    const uint16_t* cp = a.col + (rw < RW ? rowBase : 0);
    const size_t outRow = ((size_t)f * a.H + y) * a.W;
    const int bs = 2 * a.w2 + 1;
    const unsigned kb = (unsigned)q * 8u;

    uint4 hs = make_uint4(0, 0, 0, 0);
    for (int j = 0; j < bs; ++j) {
        const uint4 v = ld128(cp + (size_t)j * a.Dp);
        hs.x += v.x; hs.y += v.y; hs.z += v.z; hs.w += v.w;
    }
    // The per-pixel epilogue (texture test, sub-pixel division, store) is identical on the lanes of a row, so it is
    // deferred: lane q keeps the winner of every nOct-th step and the lanes of a row finish nOct pixels at once.
    unsigned svKey = 0, svP = 0, svN = 0;
    int svX = -1, sc = 0;
    auto flush = [&]() {
        if (svX >= 0 && active) {
            const size_t oi = outRow + svX;
            if (a.tex[oi] >= a.texThr) {
                const int minsad = (int)(svKey >> 16), mind = (int)(svKey & 0xffffu);
                const int p = (int)svP, n = (int)svN;
                const int den = p + n - 2 * minsad + abs(p - n);
                a.disp[oi] = (int16_t)((((a.D - 1 - mind) * 256) + (den != 0 ? (p - n) * 256 / den : 0) + 15) >> 4);
            }
        }
        svX = -1;
    };
    const int xEnd = a.width1 - a.w2;
    for (int xp = a.w2; xp < xEnd; ++xp) {
        const unsigned Sf[4] = {hs.x, hs.y, hs.z, hs.w};
        // first argmin through (SAD << 16) | k keys: low halves by a shift-add, high halves by a mask-or, 32-bit min3
        unsigned key = __vimin3_u32((Sf[0] << 16) + kb, (Sf[0] & 0xffff0000u) | (kb + 1), (Sf[1] << 16) + (kb + 2));
        key = __vimin3_u32(key, (Sf[1] & 0xffff0000u) | (kb + 3), (Sf[2] << 16) + (kb + 4));
        key = __vimin3_u32(key, (Sf[2] & 0xffff0000u) | (kb + 5), (Sf[3] << 16) + (kb + 6));
        key = min(key, (Sf[3] & 0xffff0000u) | (kb + 7));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned other = __shfl_down_sync(FULL, key, o);
            if (q + o < nOct) key = min(key, other);
        }
        key = __shfl_sync(FULL, key, base);
        const int minsad = (int)(key >> 16), mind = (int)(key & 0xffffu);
        bool reject = false;
        if (a.uniq > 0) {
            const int thresh = minsad + (minsad * a.uniq / 100);
            bool bad = false;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = (int)kb + j;
                const int s = (int)((j & 1) ? (Sf[j >> 1] >> 16) : (Sf[j >> 1] & 0xffffu));
                bad |= (k < mind - 1 || k > mind + 1) && (s <= thresh);
            }
            const unsigned bal = __ballot_sync(FULL, bad && rw < RW);
            const unsigned gmask = (nOct == 32) ? FULL : (((1u << nOct) - 1u) << base);
            reject = (bal & gmask) != 0u;
        }
        const int ip = (mind + 1 < a.D) ? mind + 1 : a.D - 2;
        const int in = (mind > 0) ? mind - 1 : 1;
        unsigned vp = pick16b(Sf, ip & 7), vn = pick16b(Sf, in & 7);
        vp = __shfl_sync(FULL, vp, base + (ip >> 3));
        vn = __shfl_sync(FULL, vn, base + (in >> 3));
        if (sc == q) { svKey = key; svP = vp; svN = vn; svX = reject ? -1 : xp + a.lofs; }
        if (++sc == nOct) { flush(); sc = 0; }
        if (xp + 1 < xEnd) {
            const uint4 nx = ld128(cp + (size_t)(xp + 1 + a.w2) * a.Dp);
            const uint4 od = ld128(cp + (size_t)(xp - a.w2) * a.Dp);
            hs.x += nx.x - od.x; hs.y += nx.y - od.y; hs.z += nx.z - od.z; hs.w += nx.w - od.w;
        }
    }
    flush();
}

__global__ void k_fill16(int16_t* p, size_t n, int16_t v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

void launch_bm(mvsv_ctx* c, int B)
{
    const BmNorm& n = c->bm;
    const int W = c->W, H = c->H;
    const size_t npx = (size_t)B * W * H;
    { KernelTimer kt(c, KID_FILL); k_fill16<<<(unsigned)((npx + 255) / 256), 256, 0, c->stream>>>(c->disp, npx, (int16_t)n.FILT); }
    {
        dim3 blk(128), grd((W + 127) / 128, (H + BM_ROWS - 1) / BM_ROWS, 2 * B);
        KernelTimer kt(c, KID_BM_PREFILTER);
        k_bm_prefilter<<<grd, blk, 0, c->stream>>>(c->rect[0], c->rect[1], c->pitch, W, H, n.cap, c->bm_pre[0], c->bm_pre[1]);
    }
    const bool any = !(n.lofs >= W || n.width1 < 1) && (H - 2 * n.w2 > 0) && (n.width1 - 2 * n.w2 > 0);
    if (!any) return;
    {
        dim3 blk(128), grd((W + 127) / 128, (H - 2 * n.w2 + BM_ROWS - 1) / BM_ROWS, B);
        { KernelTimer kt(c, KID_BM_TEX); k_bm_tex_col<<<grd, blk, 0, c->stream>>>(c->bm_pre[0], c->pitch, W, H, n.w2, n.cap, c->bm_tex); }
        dim3 grd2((W - 2 * n.w2 + 128 * BM_SEG - 1) / (128 * BM_SEG), H - 2 * n.w2, B);
        { KernelTimer kt(c, KID_BM_TEX); k_bm_tex_row<<<grd2, blk, 0, c->stream>>>(c->bm_tex, W, H, n.w2, c->bm_tex2); }
    }
    {
        const long long threads = (long long)n.width1 * (n.D / 8);
        dim3 grd((unsigned)((threads + 127) / 128), B);
        // blockSize^2 * 2 * cap <= 65535 (contract) bounds blockSize by 181: the ring needs at most 181 KB
        const size_t ringBytes = (size_t)n.bs * 128 * sizeof(uint2);
        static bool configured[64] = {};        // per device: a process may drive several GPUs
        if (c->device < 0 || c->device >= 64 || !configured[c->device]) {
            cudaFuncSetAttribute(k_bm_colsum, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (c->device >= 0 && c->device < 64) configured[c->device] = true;
        }
        KernelTimer kt(c, KID_BM_COLSUM);
        k_bm_colsum<<<grd, 128, ringBytes, c->stream>>>(c->bm_pre[0], c->bm_pre[1], c->pitch, H, n.width1, n.D, n.lofs, n.w2, c->bm_col);
    }
    BmArgs a;
    a.col = c->bm_col; a.tex = c->bm_tex2; a.disp = c->disp; a.W = W; a.H = H; a.width1 = n.width1; a.D = n.D; a.Dp = n.Dp;
    a.lofs = n.lofs; a.w2 = n.w2; a.texThr = n.tex; a.uniq = n.uniq; a.B = B;
    KernelTimer kt(c, KID_BM_WTA);
    const int rowsPerWarp = 32 / (n.D / 8);
    const long long warps = ((long long)B * (H - 2 * n.w2) + rowsPerWarp - 1) / rowsPerWarp;
    k_bm_wta<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, c->stream>>>(a);
}
