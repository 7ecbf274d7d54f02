// StereoSGBM on sm_100a -- replaces cv::StereoSGBM::compute behind Disparity::sgbm
// (reference src/disparity.cpp:6-10).  Stage semantics: SURVEY.md Appendix A.2 (pinned against cv2 4.13).
//
// Data layout in HBM (all volumes int16, disparity innermost, padded to Dp = 8*G so that one pixel's
// disparities are G consecutive 16-byte vectors):
//   planes[img][6][B][H][pitch] u8  : prefilter channels a/lo/hi for the clipped x-Sobel and the raw image
//   VS[B][H][W1][Dp]                : vertical box sums of the Birchfield-Tomasi pixel cost
//   C [B][H][W1][Dp]                : block cost (horizontal box sum of VS)
//   S [B][H][W1][Dp]                : sum of path costs L_r
// Lane mapping everywhere: a pixel's D disparities are spread over G = Dp/8 adjacent lanes, 8 disparities
// (four packed u16x2 registers) per lane; the min over d is a width-G shuffle butterfly; all arithmetic is
// packed 16x2 (VIADD.16x2 / VIMNMX.U16x2 / VIMNMX3 / VIADDMNMX -- the DPX path on sm_100a).
#include "mvsv_internal.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int VS_THREADS = 256;

// ------------------------------------------------------------------------------------------------
// K2a: prefilter planes (A.2: sob/raw channels with ftzero borders, lo/hi half-sample bounds)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sob_at(const uint8_t* r0, const uint8_t* r1, const uint8_t* r2, int c, int W, int ftzero)
{
    if (c <= 0 || c >= W - 1) return ftzero;
    int v = 2 * ((int)r1[c + 1] - (int)r1[c - 1]) + ((int)r0[c + 1] - (int)r0[c - 1]) + ((int)r2[c + 1] - (int)r2[c - 1]);
    v = max(-ftzero, min(ftzero, v));
    return v + ftzero;
}
__device__ __forceinline__ int raw_at(const uint8_t* r1, int c, int W, int ftzero)
{
    return (c <= 0 || c >= W - 1) ? ftzero : (int)r1[c];
}

__global__ void k_sgbm_prefilter(const uint8_t* __restrict__ img0, const uint8_t* __restrict__ img1, size_t pitch,
                                 int W, int H, int ftzero, uint8_t* __restrict__ pl0, uint8_t* __restrict__ pl1,
                                 size_t planeStride)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    int f = blockIdx.z >> 1, im = blockIdx.z & 1;
    if (x >= W) return;
    const uint8_t* img = (im ? img1 : img0) + (size_t)f * H * pitch;
    uint8_t* pl = (im ? pl1 : pl0) + ((size_t)f * H + y) * pitch + x;
    const uint8_t* r1 = img + (size_t)y * pitch;
    const uint8_t* r0 = img + (size_t)max(y - 1, 0) * pitch;
    const uint8_t* r2 = img + (size_t)min(y + 1, H - 1) * pitch;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        int a = ch ? raw_at(r1, x, W, ftzero) : sob_at(r0, r1, r2, x, W, ftzero);
        int lo = a, hi = a;
        if (x > 0) {
            int l = ch ? raw_at(r1, x - 1, W, ftzero) : sob_at(r0, r1, r2, x - 1, W, ftzero);
            int m = (a + l) >> 1;
            lo = min(lo, m); hi = max(hi, m);
        }
        if (x < W - 1) {
            int r = ch ? raw_at(r1, x + 1, W, ftzero) : sob_at(r0, r1, r2, x + 1, W, ftzero);
            int m = (a + r) >> 1;
            lo = min(lo, m); hi = max(hi, m);
        }
        pl[(size_t)(ch * 3 + 0) * planeStride] = (uint8_t)a;
        pl[(size_t)(ch * 3 + 1) * planeStride] = (uint8_t)lo;
        pl[(size_t)(ch * 3 + 2) * planeStride] = (uint8_t)hi;
    }
}

// ------------------------------------------------------------------------------------------------
// K2b: Birchfield-Tomasi pixel cost + vertical box sum.  One CTA = PX = 256/G adjacent columns of one
// frame, marching down the rows in lock-step; per row the right-image channels of the columns the CTA can
// touch are staged in shared memory as eight element-shifted copies so that every lane fetches its eight
// consecutive disparities with one aligned 128-bit load per channel.
// ------------------------------------------------------------------------------------------------
struct VsArgs {
    const uint8_t* plL; const uint8_t* plR; size_t planeStride; size_t pitch;
    uint16_t* VS;
    int W, H, W1, D, Dp, minD, minX1, SH2, NEP, LEN, nepShift;
};

__device__ __forceinline__ unsigned bt_pair(unsigned v, unsigned v0, unsigned nv1, unsigned uu, unsigned nuu,
                                            unsigned uu1, unsigned uu0, unsigned kk)
{
    unsigned c0 = __vimax_s16x2_relu(__vadd2(uu, nv1), __vadd2(v0, nuu));   // max(0, u-v1, v0-u)
    unsigned c1 = __vmaxs2(v, uu1) - __vmins2(v, uu0) - kk;                  // max(0, v-u1, u0-v), no cross-half borrow
    return __vmins2(c0, c1);
}

template <int G>
__global__ void __launch_bounds__(VS_THREADS) k_sgbm_vsum(VsArgs a)
{
    constexpr int PX = VS_THREADS / G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t* sR = reinterpret_cast<uint16_t*>(smem_raw);                        // [6][8][LEN]
    unsigned* sL = reinterpret_cast<unsigned*>(sR + 6 * 8 * a.LEN);             // [PX][12]
    uint4* ring = reinterpret_cast<uint4*>(sL + PX * 12);                        // [bs][VS_THREADS]

    const int tid = threadIdx.x;
    const int p = tid / G, q = tid % G;
    const int f = blockIdx.y;
    const int xa = blockIdx.x * PX;
    const int xi = xa + p;
    const bool live = (xi < a.W1) && (q * 8 < a.D);
    const int bs = 2 * a.SH2 + 1;
    const int xr_max = xa + PX - 1 + a.minX1 - a.minD;       // entry e <-> right column xr_max - e
    const int e0 = (PX - 1 - p) + 8 * q;
    const int sh = (-e0) & 7;
    const int rbase = sh * a.LEN + e0 + sh;                   // + arr*8*LEN
    const size_t frameOff = (size_t)f * a.H * a.pitch;

    uint4 acc = make_uint4(0, 0, 0, 0);
    int slot = 0;
    const int steps = a.H + 2 * a.SH2;
    for (int t = 0; t < steps; ++t) {
        const int y = min(max(t - a.SH2, 0), a.H - 1);
        const size_t rowOff = frameOff + (size_t)y * a.pitch;
        __syncthreads();
        // ---- stage right-image channels: pairs of entries, 8 shifted copies, 32-bit stores
        const int npairs = 6 * (a.NEP >> 1);
        for (int idx = tid; idx < npairs; idx += VS_THREADS) {
            const int arr = idx >> (a.nepShift - 1);
            const int e = (idx & ((a.NEP >> 1) - 1)) << 1;
            const uint8_t* pl = a.plR + (size_t)arr * a.planeStride + rowOff;
            const int xr0 = xr_max - e;
            int vm1 = (xr0 + 1 >= 0 && xr0 + 1 < a.W) ? pl[xr0 + 1] : 0;
            int v0 = (xr0 >= 0 && xr0 < a.W) ? pl[xr0] : 0;
            int v1 = (xr0 - 1 >= 0 && xr0 - 1 < a.W) ? pl[xr0 - 1] : 0;
            if (arr == 2 || arr == 5) { vm1 = -vm1; v0 = -v0; v1 = -v1; }
            const unsigned we = ((unsigned)v0 & 0xffffu) | ((unsigned)v1 << 16);     // entries (e, e+1)
            const unsigned wo = ((unsigned)vm1 & 0xffffu) | ((unsigned)v0 << 16);    // entries (e-1, e)
            unsigned* dst = reinterpret_cast<unsigned*>(sR + (size_t)arr * 8 * a.LEN);
#pragma unroll
            for (int s = 0; s < 8; s += 2) {
                dst[(s * a.LEN + e + s) >> 1] = we;
                dst[((s + 1) * a.LEN + e + s) >> 1] = wo;      // copy s+1, positions (e+s, e+s+1) = entries (e-1, e)
            }
        }
        // ---- stage left-image scalars as packed words
        for (int sidx = tid; sidx < PX * 12; sidx += VS_THREADS) {
            const int pp = sidx / 12, w = sidx % 12;
            unsigned val = 0;
            if (w < 10 && xa + pp < a.W1) {
                const int ch = w / 5, k = w % 5;
                const uint8_t* pl = a.plL + (size_t)(ch * 3) * a.planeStride + rowOff + (xa + pp + a.minX1);
                const int u = pl[0], lo = pl[a.planeStride], hi = pl[2 * a.planeStride];
                val = k == 0 ? pk16(u) : k == 1 ? pk16(-u) : k == 2 ? pk16(hi) : k == 3 ? pk16(lo) : pk16(hi - lo);
            }
            sL[sidx] = val;
        }
        __syncthreads();
        if (live) {
            const uint4 A0 = ld128(sR + 0 * 8 * a.LEN + rbase), A1 = ld128(sR + 1 * 8 * a.LEN + rbase);
            const uint4 A2 = ld128(sR + 2 * 8 * a.LEN + rbase), A3 = ld128(sR + 3 * 8 * a.LEN + rbase);
            const uint4 A4 = ld128(sR + 4 * 8 * a.LEN + rbase), A5 = ld128(sR + 5 * 8 * a.LEN + rbase);
            const uint4 w0 = ld128(sL + p * 12), w1 = ld128(sL + p * 12 + 4), w2 = ld128(sL + p * 12 + 8);
            // words: 0 uu_s 1 nuu_s 2 uu1_s 3 uu0_s 4 kk_s 5 uu_r 6 nuu_r 7 uu1_r 8 uu0_r 9 kk_r
            uint4 pix;
#define MVSV_PIX(c)                                                                                   \
    {                                                                                                 \
        unsigned cs = bt_pair(A0.c, A1.c, A2.c, w0.x, w0.y, w0.z, w0.w, w1.x);                        \
        unsigned cr = bt_pair(A3.c, A4.c, A5.c, w1.y, w1.z, w1.w, w2.x, w2.y);                        \
        pix.c = cs + ((cr >> 2) & 0x3fff3fffu);                                                       \
    }
            MVSV_PIX(x) MVSV_PIX(y) MVSV_PIX(z) MVSV_PIX(w)
#undef MVSV_PIX
            uint4* rs = ring + (size_t)slot * VS_THREADS + tid;
            if (t >= bs) {
                const uint4 old = *rs;
                acc.x += pix.x - old.x; acc.y += pix.y - old.y; acc.z += pix.z - old.z; acc.w += pix.w - old.w;
            } else {
                acc.x += pix.x; acc.y += pix.y; acc.z += pix.z; acc.w += pix.w;
            }
            *rs = pix;
            if (t >= bs - 1) {
                const int yo = t - (bs - 1);
                st128(a.VS + (((size_t)f * a.H + yo) * a.W1 + xi) * a.Dp + q * 8, acc);
            }
        }
        if (++slot == bs) slot = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// K3: one step of the path recurrence (A.2 `step`) on 8 disparities per lane.
//   L[k] = C[k] + min(Lp[k], Lp[k-1]+P1, Lp[k+1]+P1, m+P2) - m ;  mm = packed min_k L[k]
// Off-domain predecessor == state (L = 0, mm = 0), which yields L = C.
// ------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void sgm_step(unsigned (&L)[4], unsigned& mm, const uint4& C, unsigned P1P1, unsigned P2P2,
                                         int q, bool padLane)
{
    unsigned up = MVSV_PK_MAX, dn = MVSV_PK_MAX;
    if (G > 1) {
        const unsigned u = __shfl_up_sync(FULL, L[3], 1, G);
        const unsigned d = __shfl_down_sync(FULL, L[0], 1, G);
        if (q != 0) up = u;
        if (q != G - 1) dn = d;
    }
    const unsigned X0 = __byte_perm(up, L[0], 0x5432);
    const unsigned X1 = __byte_perm(L[0], L[1], 0x5432);
    const unsigned X2 = __byte_perm(L[1], L[2], 0x5432);
    const unsigned X3 = __byte_perm(L[2], L[3], 0x5432);
    const unsigned X4 = __byte_perm(L[3], dn, 0x5432);
    const unsigned mP2 = __vadd2(mm, P2P2);
    unsigned n0 = __vminu2(__viaddmin_u16x2(__vminu2(X0, X1), P1P1, L[0]), mP2) + C.x - mm;
    unsigned n1 = __vminu2(__viaddmin_u16x2(__vminu2(X1, X2), P1P1, L[1]), mP2) + C.y - mm;
    unsigned n2 = __vminu2(__viaddmin_u16x2(__vminu2(X2, X3), P1P1, L[2]), mP2) + C.z - mm;
    unsigned n3 = __vminu2(__viaddmin_u16x2(__vminu2(X3, X4), P1P1, L[3]), mP2) + C.w - mm;
    if (padLane) { n0 = n1 = n2 = n3 = MVSV_PK_MAX; }
    unsigned m = __vminu2(__vimin3_u16x2(n0, n1, n2), n3);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) m = __vminu2(m, __shfl_xor_sync(FULL, m, o, G));
    mm = __vminu2(m, __byte_perm(m, 0, 0x1032));
    L[0] = n0; L[1] = n1; L[2] = n2; L[3] = n3;
}

__device__ __forceinline__ void reset_state(unsigned (&L)[4], unsigned& mm, bool padLane)
{
    const unsigned v = padLane ? MVSV_PK_MAX : 0u;
    L[0] = L[1] = L[2] = L[3] = v;
    mm = 0u;
}

struct AggArgs {
    const uint16_t* VS; uint16_t* C; uint16_t* S;
    int H, W, W1, D, Dp, SW2, B;
    unsigned P1P1, P2P2;
    // WTA
    int16_t* disp; int* d2;
    int minD, minX1, maxX1, INV, uniq, d12;
    int storeS;
};

// K3a: horizontal box sum (VS -> C) fused with the left-to-right path r=(-1,0).  Writes C and S = L.
template <int G>
__global__ void __launch_bounds__(128) k_sgbm_h1(AggArgs a)
{
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nrows = (long long)a.B * a.H;
    long long row = gtid / G;
    const int q = (int)(gtid % G);
    const bool active = row < nrows;
    if (!active) row = nrows - 1;
    const bool padLane = q * 8 >= a.D;
    const size_t rowBase = (size_t)row * a.W1 * a.Dp + q * 8;
    const uint16_t* vs = a.VS + rowBase;

    uint4 hs = make_uint4(0, 0, 0, 0);
    for (int j = -a.SW2; j <= a.SW2; ++j) {
        const uint4 v = ld128(vs + (size_t)min(max(j, 0), a.W1 - 1) * a.Dp);
        hs.x += v.x; hs.y += v.y; hs.z += v.z; hs.w += v.w;
    }
    unsigned L[4], mm;
    reset_state(L, mm, padLane);
    for (int xi = 0; xi < a.W1; ++xi) {
        const uint4 nx = ld128(vs + (size_t)min(xi + 1 + a.SW2, a.W1 - 1) * a.Dp);
        const uint4 od = ld128(vs + (size_t)max(xi - a.SW2, 0) * a.Dp);
        sgm_step<G>(L, mm, hs, a.P1P1, a.P2P2, q, padLane);
        if (active) {
            st128(a.C + rowBase + (size_t)xi * a.Dp, hs);
            st128(a.S + rowBase + (size_t)xi * a.Dp, make_uint4(L[0], L[1], L[2], L[3]));
        }
        hs.x += nx.x - od.x; hs.y += nx.y - od.y; hs.z += nx.z - od.z; hs.w += nx.w - od.w;
    }
}

// K3b: vertical / diagonal paths.  One lane group follows one path line through the frame: the vertical
// line of column g, or the diagonal that starts at column g and wraps around the cost domain (a wrap is
// exactly an off-domain predecessor, so the state is reset there).  No inter-thread communication.
//   dxs: x offset of the predecessor (-1, 0, +1);  bottomUp: predecessor row is y+1 instead of y-1.
template <int G>
__global__ void __launch_bounds__(128) k_sgbm_vdir(AggArgs a, int dxs, int bottomUp)
{
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long ncols = (long long)a.B * a.W1;
    long long col = gtid / G;
    const int q = (int)(gtid % G);
    const bool active = col < ncols;
    if (!active) col = ncols - 1;
    const int f = (int)(col / a.W1);
    int x = (int)(col % a.W1);
    const bool padLane = q * 8 >= a.D;
    const size_t frameBase = (size_t)f * a.H * a.W1 * a.Dp + q * 8;

    unsigned L[4], mm;
    reset_state(L, mm, padLane);
    for (int yi = 0; yi < a.H; ++yi) {
        const int y = bottomUp ? a.H - 1 - yi : yi;
        const size_t off = frameBase + ((size_t)y * a.W1 + x) * a.Dp;
        const uint4 Cc = ld128(a.C + off);
        uint4 Sc = ld128(a.S + off);
        if ((dxs < 0 && x == 0) || (dxs > 0 && x == a.W1 - 1)) reset_state(L, mm, padLane);
        sgm_step<G>(L, mm, Cc, a.P1P1, a.P2P2, q, padLane);
        Sc.x = __viaddmin_u16x2(Sc.x, L[0], MVSV_PK_MAX);
        Sc.y = __viaddmin_u16x2(Sc.y, L[1], MVSV_PK_MAX);
        Sc.z = __viaddmin_u16x2(Sc.z, L[2], MVSV_PK_MAX);
        Sc.w = __viaddmin_u16x2(Sc.w, L[3], MVSV_PK_MAX);
        if (active) st128(a.S + off, Sc);
        x -= dxs;
        if (x >= a.W1) x = 0;
        if (x < 0) x = a.W1 - 1;
    }
}

// K3c + K4: right-to-left path r=(+1,0) fused with winner-take-all, uniqueness, sub-pixel interpolation,
// the disp2 scatter (sequential in x per row, exactly the reference order) and the left-right check.
__device__ __forceinline__ unsigned pick16(const unsigned (&R)[4], int idx)
{
    const int w = (idx >> 1) & 3;
    unsigned r = w == 0 ? R[0] : w == 1 ? R[1] : w == 2 ? R[2] : R[3];
    return (idx & 1) ? (r >> 16) : (r & 0xffffu);
}

template <int G>
__device__ __forceinline__ bool group_any(bool v)
{
    const unsigned b = __ballot_sync(FULL, v);
    const unsigned gmask = (G == 32) ? FULL : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));
    return (b & gmask) != 0u;
}

template <int G>
__global__ void __launch_bounds__(128) k_sgbm_h2_wta(AggArgs a)
{
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nrows = (long long)a.B * a.H;
    long long row = gtid / G;
    const int q = (int)(gtid % G);
    const bool active = row < nrows;
    if (!active) row = nrows - 1;
    const bool padLane = q * 8 >= a.D;
    const size_t rowBase = (size_t)row * a.W1 * a.Dp + q * 8;
    int16_t* drow = a.disp + (size_t)row * a.W;
    int* d2row = a.d2 + (size_t)row * a.W;
    const int d2init = (MVSV_MAX_COST << 16) | (a.INV & 0xffff);
    if (active)
        for (int x = q; x < a.W; x += G) { drow[x] = (int16_t)a.INV; d2row[x] = d2init; }
    __syncwarp();

    unsigned L[4], mm;
    reset_state(L, mm, padLane);
    const int umul = 100 - a.uniq;
    for (int xi = a.W1 - 1; xi >= 0; --xi) {
        const size_t off = rowBase + (size_t)xi * a.Dp;
        const uint4 Cc = ld128(a.C + off);
        const uint4 Sc = ld128(a.S + off);
        sgm_step<G>(L, mm, Cc, a.P1P1, a.P2P2, q, padLane);
        unsigned Sf[4];
        Sf[0] = __viaddmin_u16x2(Sc.x, L[0], MVSV_PK_MAX);
        Sf[1] = __viaddmin_u16x2(Sc.y, L[1], MVSV_PK_MAX);
        Sf[2] = __viaddmin_u16x2(Sc.z, L[2], MVSV_PK_MAX);
        Sf[3] = __viaddmin_u16x2(Sc.w, L[3], MVSV_PK_MAX);
        if (padLane) Sf[0] = Sf[1] = Sf[2] = Sf[3] = MVSV_PK_MAX;
        if (a.storeS && active) st128(a.S + off, make_uint4(Sf[0], Sf[1], Sf[2], Sf[3]));
        // ---- first argmin via (S << 16 | k) keys
        const unsigned kb = (unsigned)q * 8u;
        unsigned key = min(min((Sf[0] << 16) | kb, (Sf[0] & 0xffff0000u) | (kb + 1)),
                           min((Sf[1] << 16) | (kb + 2), (Sf[1] & 0xffff0000u) | (kb + 3)));
        key = min(key, min(min((Sf[2] << 16) | (kb + 4), (Sf[2] & 0xffff0000u) | (kb + 5)),
                           min((Sf[3] << 16) | (kb + 6), (Sf[3] & 0xffff0000u) | (kb + 7))));
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(FULL, key, o, G));
        const int minS = (int)(key >> 16);
        const int best = (int)(key & 0xffffu);
        bool reject = (minS >= MVSV_MAX_COST);      // every S[d] saturated: best = -1, output stays INVALID
        if (a.uniq > 0) {
            bool bad = false;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = (int)kb + j;
                const int s = (int)pick16(Sf, j);
                bad |= (k < a.D) && (s * umul < minS * 100) && (abs(k - best) > 1);
            }
            reject |= group_any<G>(bad);
        }
        // ---- neighbours of the winner for the parabola
        int sm1 = 0, sp1 = 0;
        {
            const int im = max(best - 1, 0), ip = min(best + 1, a.D - 1);
            unsigned vm = pick16(Sf, im & 7), vp = pick16(Sf, ip & 7);
            if (G > 1) {
                vm = __shfl_sync(FULL, vm, im >> 3, G);
                vp = __shfl_sync(FULL, vp, ip >> 3, G);
            }
            sm1 = (int)vm; sp1 = (int)vp;
        }
        if (active && q == 0 && !reject) {
            const int x2 = xi + a.minX1 - best - a.minD;
            const int cur = d2row[x2];
            if ((cur >> 16) > minS) d2row[x2] = (minS << 16) | ((best + a.minD) & 0xffff);
            int v = 16 * best;
            if (best > 0 && best < a.D - 1) {
                const int den = max(sm1 + sp1 - 2 * minS, 1);
                v += ((sm1 - sp1) * 16 + den) / (2 * den);
            }
            drow[xi + a.minX1] = (int16_t)(v + 16 * a.minD);
        }
    }
    __syncwarp();
    if (active) {
        for (int x = a.minX1 + q; x < a.maxX1; x += G) {
            const int d1 = drow[x];
            if (d1 == a.INV) continue;
            const int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
            const int _x = x - _d, x_ = x - d_;
            if (_x < 0 || _x >= a.W || x_ < 0 || x_ >= a.W) continue;
            const int da = (int)(int16_t)(d2row[_x] & 0xffff), db = (int)(int16_t)(d2row[x_] & 0xffff);
            if (da >= a.minD && abs(da - _d) > a.d12 && db >= a.minD && abs(db - d_) > a.d12) drow[x] = (int16_t)a.INV;
        }
    }
}

__global__ void k_fill_i16(int16_t* p, size_t n, int16_t v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

template <int G>
void launch_sgbm_g(mvsv_ctx* c, int B)
{
    const SgbmNorm& n = c->sg;
    cudaStream_t st = c->stream;
    const size_t planeStride = (size_t)c->maxB * c->H * c->pitch;
    {
        dim3 blk(128), grd((c->W + 127) / 128, c->H, 2 * B);
        KernelTimer kt(c, KID_SGBM_PREFILTER);
        k_sgbm_prefilter<<<grd, blk, 0, st>>>(c->rect[0], c->rect[1], c->pitch, c->W, c->H, n.ftzero, c->planes[0],
                                              c->planes[1], planeStride);
    }
    {
        constexpr int PX = VS_THREADS / G;
        VsArgs a;
        a.plL = c->planes[0]; a.plR = c->planes[1]; a.planeStride = planeStride; a.pitch = c->pitch;
        a.VS = c->VS; a.W = c->W; a.H = c->H; a.W1 = n.W1; a.D = n.D; a.Dp = n.Dp; a.minD = n.minD; a.minX1 = n.minX1;
        a.SH2 = n.SH2;
        const int NE = PX - 1 + n.Dp;
        int nep = 16, sh = 4;
        while (nep < NE + 2) { nep <<= 1; ++sh; }
        a.NEP = nep; a.nepShift = sh; a.LEN = nep + 8;
        const size_t smem = (size_t)6 * 8 * a.LEN * 2 + (size_t)PX * 12 * 4 + (size_t)(2 * n.SH2 + 1) * VS_THREADS * 16;
        dim3 grd((n.W1 + PX - 1) / PX, B);
        KernelTimer kt(c, KID_SGBM_VSUM);
        k_sgbm_vsum<G><<<grd, VS_THREADS, smem, st>>>(a);
    }
    AggArgs a;
    a.VS = c->VS; a.C = c->C; a.S = c->S; a.H = c->H; a.W = c->W; a.W1 = n.W1; a.D = n.D; a.Dp = n.Dp; a.SW2 = n.SW2;
    a.B = B; a.P1P1 = ((unsigned)n.P1 & 0xffffu) * 0x10001u; a.P2P2 = ((unsigned)n.P2 & 0xffffu) * 0x10001u;
    a.disp = c->disp_raw; a.d2 = c->d2; a.minD = n.minD; a.minX1 = n.minX1; a.maxX1 = n.maxX1; a.INV = n.INV;
    a.uniq = n.uniq; a.d12 = n.d12;
    a.storeS = (c->debug_flags & 1) ? 1 : 0;
    const int TPB = 128;
    const long long rowThreads = (long long)B * c->H * G, colThreads = (long long)B * n.W1 * G;
    const unsigned rowBlocks = (unsigned)((rowThreads + TPB - 1) / TPB), colBlocks = (unsigned)((colThreads + TPB - 1) / TPB);
    { KernelTimer kt(c, KID_SGBM_H1); k_sgbm_h1<G><<<rowBlocks, TPB, 0, st>>>(a); }
    { KernelTimer kt(c, KID_SGBM_VDIR); k_sgbm_vdir<G><<<colBlocks, TPB, 0, st>>>(a, -1, 0); }
    { KernelTimer kt(c, KID_SGBM_VDIR); k_sgbm_vdir<G><<<colBlocks, TPB, 0, st>>>(a, 0, 0); }
    { KernelTimer kt(c, KID_SGBM_VDIR); k_sgbm_vdir<G><<<colBlocks, TPB, 0, st>>>(a, +1, 0); }
    if (n.mode == 1) {
        { KernelTimer kt(c, KID_SGBM_VDIR); k_sgbm_vdir<G><<<colBlocks, TPB, 0, st>>>(a, -1, 1); }
        { KernelTimer kt(c, KID_SGBM_VDIR); k_sgbm_vdir<G><<<colBlocks, TPB, 0, st>>>(a, 0, 1); }
        { KernelTimer kt(c, KID_SGBM_VDIR); k_sgbm_vdir<G><<<colBlocks, TPB, 0, st>>>(a, +1, 1); }
    }
    { KernelTimer kt(c, KID_SGBM_H2_WTA); k_sgbm_h2_wta<G><<<rowBlocks, TPB, 0, st>>>(a); }
}

template <int G>
cudaError_t cfg_vsum()
{
    return cudaFuncSetAttribute(k_sgbm_vsum<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

}  // namespace

cudaError_t sgbm_configure_kernels()
{
    cudaError_t e;
    if ((e = cfg_vsum<1>()) != cudaSuccess) return e;
    if ((e = cfg_vsum<2>()) != cudaSuccess) return e;
    if ((e = cfg_vsum<4>()) != cudaSuccess) return e;
    if ((e = cfg_vsum<8>()) != cudaSuccess) return e;
    if ((e = cfg_vsum<16>()) != cudaSuccess) return e;
    if ((e = cfg_vsum<32>()) != cudaSuccess) return e;
    return cudaSuccess;
}

void launch_sgbm(mvsv_ctx* c, int B)
{
    const SgbmNorm& n = c->sg;
    const size_t npx = (size_t)B * c->H * c->W;
    if (n.W1 <= 0) {
        KernelTimer kt(c, KID_FILL);
        k_fill_i16<<<(unsigned)((npx + 255) / 256), 256, 0, c->stream>>>(c->disp_raw, npx, (int16_t)n.INV);
    } else {
        switch (n.G) {
            case 1: launch_sgbm_g<1>(c, B); break;
            case 2: launch_sgbm_g<2>(c, B); break;
            case 4: launch_sgbm_g<4>(c, B); break;
            case 8: launch_sgbm_g<8>(c, B); break;
            case 16: launch_sgbm_g<16>(c, B); break;
            default: launch_sgbm_g<32>(c, B); break;
        }
    }
    launch_median(c, c->disp_raw, c->disp_med, B);
    cudaMemcpyAsync(c->disp, c->disp_med, npx * sizeof(int16_t), cudaMemcpyDeviceToDevice, c->stream);
    if (n.speckleWin > 0) launch_speckle(c, c->disp, B, n.INV, n.speckleWin, 16 * n.speckleRange);
}
