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

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr unsigned FULL = 0xffffffffu;
// compute threads of the cost kernel: 256, or 512 when a pixel needs >= 16 lanes (D > 64), so that the staging of
// the right-image entries (PX - 1 + Dp per row) is amortised over more columns
constexpr int vs_compute_threads(int G) { return G >= 16 ? 512 : 256; }

// ------------------------------------------------------------------------------------------------
// K2a: prefilter (A.2: sob/raw channels with ftzero borders, lo/hi half-sample bounds).
//   right image -> six u16 planes (a_s, lo_s, -hi_s, a_r, lo_r, -hi_r), stored REVERSED in x at index
//                  j = JOFF + W-1-x of a zero-padded row of RP elements, so that "disparity ascending" is
//                  "address ascending" and every CTA's first entry is 16-byte aligned;
//   left image  -> one 8-byte record per pixel (a_s, lo_s, hi_s, a_r, lo_r, hi_r, 0, 0).
// ------------------------------------------------------------------------------------------------
// Each warp covers 28 output columns (lanes 2..29; the outer two lanes on either side only feed neighbours) and
// marches down PF_ROWS rows, loading one byte per lane per row: with s(x) = r0[x] + 2 r1[x] + r2[x] the Sobel
// response is s(x+1) - s(x-1), so the horizontal taps come from shuffles and the vertical ones from registers.
constexpr int PF_ROWS = 16, PF_COLS = 28, PF_WARPS = 4;

__global__ void __launch_bounds__(PF_WARPS * 32)
k_sgbm_prefilter(const uint8_t* __restrict__ img0, const uint8_t* __restrict__ img1, size_t pitch,
                 int W, int H, int ftzero, uint2* __restrict__ recL, uint16_t* __restrict__ plR,
                 size_t planeStrideR, int RP, int JOFF)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = (blockIdx.x * PF_WARPS + warp) * PF_COLS + lane - 2;
    const int y0 = blockIdx.y * PF_ROWS, y1 = min(y0 + PF_ROWS, H);
    const int f = blockIdx.z >> 1, im = blockIdx.z & 1;
    const uint8_t* img = (im ? img1 : img0) + (size_t)f * H * pitch;
    const int xc = min(max(x, 0), W - 1);                       // out-of-image lanes load a valid byte, results unused
    const bool inimg = x >= 0 && x < W, border = x <= 0 || x >= W - 1;
    const bool writer = inimg && lane >= 2 && lane < 2 + PF_COLS;
    int p0 = img[(size_t)max(y0 - 1, 0) * pitch + xc], p1 = img[(size_t)y0 * pitch + xc];
    // both channels ride in one register (low half: clipped Sobel, high half: raw), so the half-sample bounds of the
    // two channels cost one set of packed operations; per-row addresses advance by constants
    const bool hasL = x > 0, hasR = x < W - 1;
    const unsigned ftz2 = (unsigned)ftzero * 0x10001u;
    uint2* rec = recL + ((size_t)f * H + y0) * W + xc;
    uint16_t* o = plR + ((size_t)f * H + y0) * RP + (JOFF + W - 1 - xc);
    // the row below is loaded one iteration ahead, so no iteration waits for its own load
    const uint8_t* nxt = img + (size_t)min(y0 + 1, H - 1) * pitch + xc;
    int pn = *nxt;
    for (int y = y0; y < y1; ++y) {
        const int p2 = pn;
        if (y + 2 < H) nxt += pitch;
        pn = *nxt;                                              // row min(y + 2, H - 1)
        const int sv = p0 + 2 * p1 + p2;
        const int sl = __shfl_up_sync(FULL, sv, 1), sr = __shfl_down_sync(FULL, sv, 1);
        const unsigned sob = (unsigned)(max(-ftzero, min(ftzero, sr - sl)) + ftzero);
        const unsigned A = border ? ftz2 : (sob | ((unsigned)p1 << 16));
        const unsigned l = __shfl_up_sync(FULL, A, 1), r = __shfl_down_sync(FULL, A, 1);
        // (a + neighbour) >> 1 per half: the sums stay below 2^9, so a plain add and a masked shift are exact
        const unsigned ml = hasL ? (((A + l) >> 1) & 0x7fff7fffu) : A;
        const unsigned mr = hasR ? (((A + r) >> 1) & 0x7fff7fffu) : A;
        const unsigned LO = __vminu2(A, __vminu2(ml, mr)), HI = __vmaxu2(A, __vmaxu2(ml, mr));
        if (writer) {
            if (im == 0) {
                // bytes: a_s lo_s hi_s a_r | lo_r hi_r 0 0
                const unsigned w0 = __byte_perm(__byte_perm(A, LO, 0x2040), HI, 0x3410);
                const unsigned w1 = __byte_perm(LO, HI, 0x4462) & 0xffffu;
                *rec = make_uint2(w0, w1);
            } else {
                o[0 * planeStrideR] = (uint16_t)(A & 0xffffu); o[1 * planeStrideR] = (uint16_t)(LO & 0xffffu);
                o[2 * planeStrideR] = (uint16_t)(0u - (HI & 0xffffu));
                o[3 * planeStrideR] = (uint16_t)(A >> 16); o[4 * planeStrideR] = (uint16_t)(LO >> 16);
                o[5 * planeStrideR] = (uint16_t)(0u - (HI >> 16));
            }
        }
        rec += W; o += RP;
        p0 = p1; p1 = p2;
    }
}

// ------------------------------------------------------------------------------------------------
// K2b: Birchfield-Tomasi pixel cost + vertical box sum.  One CTA = PX = 256/G adjacent columns of one frame,
// marching down the rows in lock-step.  Lane (p, q) evaluates column xa+p at disparities 8q..8q+7: the left
// pixel is a broadcast scalar, the right pixels are eight consecutive entries of the reversed planes.  Their
// start is only 2-byte aligned, so each row's entries are staged in shared memory as EIGHT element-shifted copies
// (built from aligned 128-bit global loads with funnel shifts); every lane then fetches a channel's eight values
// with one aligned LDS.128 from the copy matching its alignment.  Staging is double buffered (global loads for
// row t+2 in flight, copies of row t+1 written while row t is consumed): one __syncthreads per row.
// ------------------------------------------------------------------------------------------------
struct VsArgs {
    const uint2* recL; const uint16_t* plR; size_t planeStrideR;
    uint16_t* VS;
    int W, H, W1, D, Dp, minD, minX1, SH2, NV, LEN, RP, JOFF;
    int bandRows;               // output rows per blockIdx.z (== H: one band); a band restarts the vertical sum
};

__device__ __forceinline__ unsigned bt_pair(unsigned v, unsigned v0, unsigned nv1, unsigned uu, unsigned nuu,
                                            unsigned uu1, unsigned uu0, unsigned kk)
{
    unsigned c0 = __vimax_s16x2_relu(__vadd2(uu, nv1), __vadd2(v0, nuu));   // max(0, u-v1, v0-u)
    unsigned c1 = __vmaxs2(v, uu1) - __vmins2(v, uu0) - kk;                  // max(0, v-u1, u0-v), no cross-half borrow
    return __vmins2(c0, c1);
}

// cp.async (LDGSTS) helpers: asynchronous global -> shared copies, completion tracked per thread in groups
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
constexpr int VS_RD = 4;       // raw-ring depth of the cost kernel's producer (rows in flight + 1)

// geometry fixed by G: entries a CTA can touch, 8-entry vectors per channel, producer warps
template <int G> struct VsGeom {
    static constexpr int CT = vs_compute_threads(G);
    static constexpr int PX = CT / G;
    static constexpr int NE = PX - 1 + 8 * G;
    static constexpr int NV = (NE + 2 + 7) / 8;                // one spare entry pair for the odd copies
    static constexpr int ITEMS = 6 * NV;                       // (channel, vector) staging items per row
    static constexpr int NPW = (ITEMS + 95) / 96;              // producer warps: <= 3 items per producer lane
    static constexpr int NPT = 32 * NPW;                       // producer threads
    static constexpr int IPL = (ITEMS + NPT - 1) / NPT;        // items per producer lane
    static constexpr int RPL = (PX + NPT - 1) / NPT;           // left records per producer lane
    static constexpr int THREADS = CT + NPT;
};

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int G, bool R8, bool BANDS>
__global__ void __launch_bounds__(VsGeom<G>::THREADS) k_sgbm_vsum(VsArgs a)
{
    using GE = VsGeom<G>;
    constexpr int PX = GE::PX;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: sR[2 buf][6][8][LEN] u16 | sL[2 buf][PX][12] u32 | ring[bs][CT] (uint2 if the pixel cost
    // fits a byte -- 2*ftzero+63 <= 255 -- else uint4)
    uint16_t* sR = reinterpret_cast<uint16_t*>(smem_raw);
    unsigned* sL = reinterpret_cast<unsigned*>(sR + (size_t)2 * 6 * 8 * a.LEN);
    unsigned char* ring = reinterpret_cast<unsigned char*>(sL + 2 * PX * 12);

    const int tid = threadIdx.x;
    const int f = blockIdx.y;
    const int xa = blockIdx.x * PX;
    const int bs = 2 * a.SH2 + 1;
    // Row band [ya, yb) of this CTA (small batches are split into bands to fill the GPU; a band pays blockSize-1
    // warm-up rows): step t adds pixel row clamp(t - SH2) and, once bs rows are in, emits output row t - (bs-1)
    const int ya = BANDS ? blockIdx.z * a.bandRows : 0, yb = BANDS ? min(ya + a.bandRows, a.H) : a.H;
    const int tb = ya, steps = yb + 2 * a.SH2;
    const size_t rowsR = (size_t)f * a.H;
    // named barriers: 1 + buf = "stage buf is full", 3 + buf = "stage buf is free again"
    constexpr int NTH = GE::THREADS;

    constexpr int CT = GE::CT;
    if (tid >= CT) {
        // ===================== producer warps: stage row t while the compute warps consume row t-1 ===============
        // Global loads run VS_RD-1 rows ahead through cp.async into a private raw ring (every lane reads back only
        // what it copied), so the producer's critical path is LDS -> funnel shifts -> STS.
        const int pt = tid - CT;
        const int xr_max = xa + PX - 1 + a.minX1 - a.minD;       // entry e <-> right column xr_max - e
        const int j0 = a.JOFF + a.W - 1 - xr_max;                // reversed-plane index of entry 0 (multiple of 8)
        uint4* rawPQ = reinterpret_cast<uint4*>(ring + (size_t)bs * CT * (R8 ? 8 : 16));   // [VS_RD][IPL][2][NPT]
        uint2* rawL = reinterpret_cast<uint2*>(rawPQ + VS_RD * GE::IPL * 2 * GE::NPT);             // [VS_RD][RPL][NPT]
        size_t srcOff[GE::IPL]; int dstOff[GE::IPL]; bool itemOn[GE::IPL];
#pragma unroll
        for (int k = 0; k < GE::IPL; ++k) {
            const int item = pt + k * GE::NPT;
            itemOn[k] = item < GE::ITEMS;
            const int arr = itemOn[k] ? item / GE::NV : 0, m = itemOn[k] ? item % GE::NV : 0;
            srcOff[k] = (size_t)arr * a.planeStrideR + rowsR * a.RP + j0 + 8 * m;
            dstOff[k] = arr * 8 * a.LEN + 8 * m;
        }
        auto issue_loads = [&](int t) {
            if (t < steps) {
                const int y = min(max(t - a.SH2, 0), a.H - 1);
                const int rs = t % VS_RD;
#pragma unroll
                for (int k = 0; k < GE::IPL; ++k) {
                    if (itemOn[k]) {
                        const uint16_t* src = a.plR + srcOff[k] + (size_t)y * a.RP;
                        uint4* d = rawPQ + ((rs * GE::IPL + k) * 2) * GE::NPT + pt;
                        cp_async16(d, src - 8);
                        cp_async16(d + GE::NPT, src);
                    }
                }
#pragma unroll
                for (int k = 0; k < GE::RPL; ++k) {
                    const int px = pt + k * GE::NPT;
                    if (px < PX && xa + px < a.W1)
                        cp_async8(rawL + (rs * GE::RPL + k) * GE::NPT + pt, a.recL + (rowsR + y) * a.W + xa + px + a.minX1);
                }
            }
            cp_async_commit();
        };
        auto store_stage = [&](int buf, int t) {
            const int rs = t % VS_RD;
#pragma unroll
            for (int k = 0; k < GE::IPL; ++k) {
                if (itemOn[k]) {
                    uint16_t* dst = sR + (size_t)buf * 6 * 8 * a.LEN + dstOff[k];
                    const uint4* d = rawPQ + ((rs * GE::IPL + k) * 2) * GE::NPT + pt;
                    const uint4 p4 = d[0], q4 = d[GE::NPT];
                    const unsigned f0 = __funnelshift_r(p4.x, p4.y, 16), f1 = __funnelshift_r(p4.y, p4.z, 16);
                    const unsigned f2 = __funnelshift_r(p4.z, p4.w, 16), f3 = __funnelshift_r(p4.w, q4.x, 16);
                    const unsigned f4 = __funnelshift_r(q4.x, q4.y, 16), f5 = __funnelshift_r(q4.y, q4.z, 16);
                    const unsigned f6 = __funnelshift_r(q4.z, q4.w, 16);
                    // copy s holds entry e at position e+s: positions [8m, 8m+8) of copy s = entries [8m-s, 8m-s+8)
                    st128(dst + 0 * a.LEN, q4);
                    st128(dst + 1 * a.LEN, make_uint4(f3, f4, f5, f6));
                    st128(dst + 2 * a.LEN, make_uint4(p4.w, q4.x, q4.y, q4.z));
                    st128(dst + 3 * a.LEN, make_uint4(f2, f3, f4, f5));
                    st128(dst + 4 * a.LEN, make_uint4(p4.z, p4.w, q4.x, q4.y));
                    st128(dst + 5 * a.LEN, make_uint4(f1, f2, f3, f4));
                    st128(dst + 6 * a.LEN, make_uint4(p4.y, p4.z, p4.w, q4.x));
                    st128(dst + 7 * a.LEN, make_uint4(f0, f1, f2, f3));
                }
            }
#pragma unroll
            for (int k = 0; k < GE::RPL; ++k) {
                const int px = pt + k * GE::NPT;
                if (px < PX) {
                    // words: 0 uu_s 1 nuu_s 2 uu1_s 3 uu0_s 4 kk_s 5 uu_r 6 nuu_r 7 uu1_r 8 uu0_r 9 kk_r
                    const uint2 lr = (xa + px < a.W1) ? rawL[(rs * GE::RPL + k) * GE::NPT + pt] : make_uint2(0, 0);
                    const int us = lr.x & 0xff, los = (lr.x >> 8) & 0xff, his = (lr.x >> 16) & 0xff, ur = lr.x >> 24;
                    const int lor = lr.y & 0xff, hir = (lr.y >> 8) & 0xff;
                    unsigned* d = sL + (buf * PX + px) * 12;
                    st128(d, make_uint4(pk16(us), pk16(-us), pk16(his), pk16(los)));
                    st128(d + 4, make_uint4(pk16(his - los), pk16(ur), pk16(-ur), pk16(hir)));
                    st128(d + 8, make_uint4(pk16(lor), pk16(hir - lor), 0u, 0u));
                }
            }
        };
#pragma unroll
        for (int t0 = 0; t0 < VS_RD - 1; ++t0) issue_loads(tb + t0);
        for (int t = tb; t < steps; ++t) {
            const int buf = t & 1;
            issue_loads(t + VS_RD - 1);                        // into the raw slot consumed at step t-1
            cp_async_wait<VS_RD - 1>();                        // row t has landed
            if (t - tb >= 2) named_bar_sync(3 + buf, NTH);     // compute warps are done with row t-2 (same buffer)
            store_stage(buf, t);
            named_bar_arrive(1 + buf, NTH);
        }
        cp_async_wait<0>();
        return;
    }

    // ===================== compute warps ==========================================================================
    const int p = tid / G, q = tid % G;
    const int xi = xa + p;
    const bool live = (xi < a.W1) && (q * 8 < a.D);
    const int e0 = (PX - 1 - p) + 8 * q;
    const int sh = (-e0) & 7;
    const int rbase = sh * a.LEN + e0 + sh;                   // + (buf*6 + arr)*8*LEN
    uint4 acc = make_uint4(0, 0, 0, 0);
    int slot = 0;
    for (int t = tb; t < steps; ++t) {
        const int buf = t & 1;
        named_bar_sync(1 + buf, NTH);
        if (live) {
            const uint16_t* rb = sR + (size_t)buf * 6 * 8 * a.LEN + rbase;
            const uint4 A0 = ld128(rb + 0 * 8 * a.LEN), A1 = ld128(rb + 1 * 8 * a.LEN), A2 = ld128(rb + 2 * 8 * a.LEN);
            const uint4 A3 = ld128(rb + 3 * 8 * a.LEN), A4 = ld128(rb + 4 * 8 * a.LEN), A5 = ld128(rb + 5 * 8 * a.LEN);
            const unsigned* lw = sL + (buf * PX + p) * 12;
            const uint4 w0 = ld128(lw), w1 = ld128(lw + 4), w2 = ld128(lw + 8);
            uint4 pix;
#define MVSV_PIX(c)                                                                                   \
    {                                                                                                 \
        unsigned cs = bt_pair(A0.c, A1.c, A2.c, w0.x, w0.y, w0.z, w0.w, w1.x);                        \
        unsigned cr = bt_pair(A3.c, A4.c, A5.c, w1.y, w1.z, w1.w, w2.x, w2.y);                        \
        pix.c = cs + ((cr >> 2) & 0x3fff3fffu);                                                       \
    }
            MVSV_PIX(x) MVSV_PIX(y) MVSV_PIX(z) MVSV_PIX(w)
#undef MVSV_PIX
            uint4 old = make_uint4(0, 0, 0, 0);
            if (R8) {
                uint2* rs = reinterpret_cast<uint2*>(ring) + (size_t)slot * CT + tid;
                if (t - tb >= bs) {
                    const uint2 o = *rs;
                    old = make_uint4(__byte_perm(o.x, 0, 0x4140), __byte_perm(o.x, 0, 0x4342), __byte_perm(o.y, 0, 0x4140),
                                     __byte_perm(o.y, 0, 0x4342));
                }
                *rs = make_uint2(__byte_perm(pix.x, pix.y, 0x6420), __byte_perm(pix.z, pix.w, 0x6420));
            } else {
                uint4* rs = reinterpret_cast<uint4*>(ring) + (size_t)slot * CT + tid;
                if (t - tb >= bs) old = *rs;
                *rs = pix;
            }
            acc.x += pix.x - old.x; acc.y += pix.y - old.y; acc.z += pix.z - old.z; acc.w += pix.w - old.w;
            if (t - tb >= bs - 1) {
                const int yo = t - (bs - 1);
                st128(a.VS + (((size_t)f * a.H + yo) * a.W1 + xi) * a.Dp + q * 8, acc);
            }
        }
        // barrier instructions need a converged warp (`live` differs between lanes in edge CTAs / padded lanes)
        __syncwarp();
        if (t + 2 < steps) named_bar_arrive(3 + buf, NTH);       // stage buffer consumed: hand it back
        if (++slot == bs) slot = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// K3: one step of the path recurrence (A.2 `step`) on 8 disparities per lane.
//   L[k] = C[k] + min(Lp[k], Lp[k-1]+P1, Lp[k+1]+P1, m+P2) - m ;  mm = packed min_k L[k]
// Off-domain predecessor == state (L = 0, mm = 0), which yields L = C.
// PAD: numDisp is not 8*G, lanes with q*8 >= D hold the constant 0x7fff (the out-of-range neighbour value).
// ------------------------------------------------------------------------------------------------
template <int G, bool PAD>
__device__ __forceinline__ void sgm_step(unsigned (&L)[4], unsigned& mm, const uint4& C, unsigned P1P1, unsigned P2P2,
                                         int q, bool padLane)
{
    // min(min(Lp[k-1], Lp[k+1]) + P1, Lp[k], m + P2) == min3(Lp[k-1] + P1, Lp[k+1] + P1, min(Lp[k], m + P2)):
    // P1 is added once per register with a plain 32-bit add (no carry between the halves: L + P1 < 65536), which
    // issues at full rate, and one VIMNMX3 replaces a VIMNMX + VIADDMNMX pair on the half-rate DPX pipe.
    const unsigned P0 = L[0] + P1P1, P3 = L[3] + P1P1;
    unsigned up = 0xffffffffu, dn = 0xffffffffu;        // out-of-range neighbour: larger than any L + P1
    if (G > 1) {
        const unsigned u = __shfl_up_sync(FULL, P3, 1, G);
        const unsigned d = __shfl_down_sync(FULL, P0, 1, G);
        if (q != 0) up = u;
        if (q != G - 1) dn = d;
    }
    const unsigned Q1 = L[1] + P1P1, Q2 = L[2] + P1P1;
    const unsigned X0 = __byte_perm(up, P0, 0x5432);
    const unsigned X1 = __byte_perm(P0, Q1, 0x5432);
    const unsigned X2 = __byte_perm(Q1, Q2, 0x5432);
    const unsigned X3 = __byte_perm(Q2, P3, 0x5432);
    const unsigned X4 = __byte_perm(P3, dn, 0x5432);
    const unsigned mP2 = mm + P2P2;
    unsigned n0 = __vimin3_u16x2(X0, X1, __vminu2(L[0], mP2)) + C.x - mm;
    unsigned n1 = __vimin3_u16x2(X1, X2, __vminu2(L[1], mP2)) + C.y - mm;
    unsigned n2 = __vimin3_u16x2(X2, X3, __vminu2(L[2], mP2)) + C.z - mm;
    unsigned n3 = __vimin3_u16x2(X3, X4, __vminu2(L[3], mP2)) + C.w - mm;
    if (PAD && padLane) { n0 = n1 = n2 = n3 = MVSV_PK_MAX; }
    unsigned m = __vminu2(__vimin3_u16x2(n0, n1, n2), n3);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) m = __vminu2(m, __shfl_xor_sync(FULL, m, o, G));
    mm = __vminu2(m, __byte_perm(m, 0, 0x1032));
    L[0] = n0; L[1] = n1; L[2] = n2; L[3] = n3;
}

template <bool PAD>
__device__ __forceinline__ void reset_state(unsigned (&L)[4], unsigned& mm, bool padLane)
{
    const unsigned v = (PAD && padLane) ? MVSV_PK_MAX : 0u;
    L[0] = L[1] = L[2] = L[3] = v;
    mm = 0u;
}

__device__ __forceinline__ void ld_state(unsigned (&L)[4], const uint16_t* p)
{
    const uint4 v = ld128(p);
    L[0] = v.x; L[1] = v.y; L[2] = v.z; L[3] = v.w;
}
__device__ __forceinline__ void sat_acc(uint4& S, const unsigned (&L)[4])
{
    S.x = __viaddmin_u16x2(S.x, L[0], MVSV_PK_MAX); S.y = __viaddmin_u16x2(S.y, L[1], MVSV_PK_MAX);
    S.z = __viaddmin_u16x2(S.z, L[2], MVSV_PK_MAX); S.w = __viaddmin_u16x2(S.w, L[3], MVSV_PK_MAX);
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

// The row scans stream their operands through a per-lane shared-memory ring filled by cp.async (LDGSTS): the
// loads of the next PFD steps are in flight without holding registers.  Each lane only ever reads back the 16
// bytes it copied itself, so cp.async.wait_group is the only synchronisation needed.
constexpr int H1_PFD = 8;
constexpr int HPF = 2;         // k_sgbm_h2_wta is issue-bound: it keeps a cheap 2-step register prefetch instead


// K3a: horizontal box sum (VS -> C) fused with the left-to-right path r=(-1,0).  Writes C and S = L.
template <int G, bool PAD>
__global__ void __launch_bounds__(128) k_sgbm_h1(AggArgs a)
{
    constexpr int DP = 8 * G;                       // padded disparity count == a.Dp
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nrows = (long long)a.B * a.H;
    long long row = gtid / G;
    const int q = (int)(gtid % G);
    const bool active = row < nrows;
    if (!active) row = nrows - 1;
    const bool padLane = q * 8 >= a.D;
    const int W1 = a.W1, SW2 = a.SW2;
    const size_t rowBase = (size_t)row * W1 * DP + q * 8;
    const uint16_t* __restrict__ vs = a.VS + rowBase;
    uint16_t* __restrict__ cp = a.C + rowBase;
    uint16_t* __restrict__ sp = a.S + rowBase;

    // Stream element t = VS[clamp(t - SW2)], t >= 0.  The window of step xi is t in [xi, xi + bs): it gains
    // t = xi + bs and loses t = xi.  Ring slot of t is t mod R with R = bs + PFD, so the element fetched at step
    // xi (t = xi + bs + PFD) lands in the slot of the element that has just left.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* ring = reinterpret_cast<uint4*>(smem_raw) + threadIdx.x;
    const int bs = 2 * SW2 + 1, R = bs + H1_PFD;
    auto src = [&](int t) { return vs + min(max(t - SW2, 0), W1 - 1) * DP; };
    for (int t = 0; t < bs; ++t) cp_async16(ring + t * 128, src(t));
    cp_async_commit();
#pragma unroll
    for (int k = 0; k < H1_PFD; ++k) { cp_async16(ring + (bs + k) * 128, src(bs + k)); cp_async_commit(); }
    cp_async_wait<H1_PFD>();
    uint4 hs = make_uint4(0, 0, 0, 0);
    for (int t = 0; t < bs; ++t) {
        const uint4 v = ring[t * 128];
        hs.x += v.x; hs.y += v.y; hs.z += v.z; hs.w += v.w;
    }
    unsigned L[4], mm;
    reset_state<PAD>(L, mm, padLane);
    int sOut = 0, sIn = bs;                         // slots of t = xi and t = xi + bs
    for (int xi = 0; xi < W1; ++xi) {
        cp_async_wait<H1_PFD - 1>();                // the group issued H1_PFD steps ago (t = xi + bs) has landed
        const uint4 od = ring[sOut * 128];
        const uint4 nx = ring[sIn * 128];
        cp_async16(ring + sOut * 128, src(xi + bs + H1_PFD));
        cp_async_commit();
        if (++sOut == R) sOut = 0;
        if (++sIn == R) sIn = 0;
        sgm_step<G, PAD>(L, mm, hs, a.P1P1, a.P2P2, q, padLane);
        if (active) {
            st128(cp + xi * DP, hs);
            st128(sp + xi * DP, make_uint4(L[0], L[1], L[2], L[3]));
        }
        hs.x += nx.x - od.x; hs.y += nx.y - od.y; hs.z += nx.z - od.z; hs.w += nx.w - od.w;
    }
    cp_async_wait<0>();
}

// K3b: vertical / diagonal paths, one direction per launch (generic fallback when the fused sweep below does
// not fit in shared memory).  One lane group follows one path line through the frame: the vertical line of
// column g, or the diagonal that starts at column g and wraps around the cost domain (a wrap is exactly an
// off-domain predecessor, so the state is reset there).  No inter-thread communication.
//   dxs: x offset of the predecessor (-1, 0, +1);  bottomUp: predecessor row is y+1 instead of y-1.
template <int G, bool PAD>
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
    constexpr int DP = 8 * G;
    const size_t frameBase = (size_t)f * a.H * a.W1 * DP + q * 8;

    unsigned L[4], mm;
    reset_state<PAD>(L, mm, padLane);
    for (int yi = 0; yi < a.H; ++yi) {
        const int y = bottomUp ? a.H - 1 - yi : yi;
        const size_t off = frameBase + ((size_t)y * a.W1 + x) * DP;
        const uint4 Cc = ld128(a.C + off);
        uint4 Sc = ld128(a.S + off);
        if ((dxs < 0 && x == 0) || (dxs > 0 && x == a.W1 - 1)) reset_state<PAD>(L, mm, padLane);
        sgm_step<G, PAD>(L, mm, Cc, a.P1P1, a.P2P2, q, padLane);
        sat_acc(Sc, L);
        if (active) st128(a.S + off, Sc);
        x -= dxs;
        if (x >= a.W1) x = 0;
        if (x < 0) x = a.W1 - 1;
    }
}

// K3b': the three paths that come from the previous row -- r = (-1,dy), (0,dy), (+1,dy) with dy = -1 (top-down)
// or +1 (bottom-up, MODE_HH) -- fused: C and S are read once and S written once per sweep (6 B/cell instead of 18).
// One thread-block CLUSTER per frame; CTA r of the cluster owns the column strip [x0, x1).  The path state of the
// previous row lives in shared memory: the vertical path at slot lx, the diagonals at the skewed slots
// (lx -/+ y) mod M, so that a pixel's predecessor sits in the very slot the pixel overwrites (in place, no
// double buffering, no intra-row hazard).  Only the strip's border columns cross CTAs: they are written into the
// neighbour's halo through distributed shared memory, and one cluster barrier per row orders everything.
// The C/S vectors of the next work item (also across the row barrier) are prefetched into registers.
constexpr int TD_THREADS = 512;
constexpr int TD_SMEM_LIMIT = 200 * 1024;
// byte offset of the six halo mbarriers behind L | halo[3][2] | m | halom[3][2] (rounded up to 16)
__host__ __device__ inline size_t td_bar_offset(int Mmax, int Dp)
{
    return (((size_t)3 * Mmax * Dp * 2 + (size_t)6 * Dp * 2 + (size_t)3 * Mmax * 4 + 6 * 4) + 15) / 16 * 16;
}

struct TdArgs {
    const uint16_t* C; uint16_t* S;
    int H, W1, D, Dp, NC, Mmax, bottomUp;
    unsigned P1P1, P2P2;
};

// Cluster barrier split by memory semantics: only warps that wrote a neighbour's halo (distributed shared memory)
// arrive with .release (which costs a fence over their outstanding global stores); all other warps arrive
// .relaxed -- their shared-memory writes are CTA-local and ordered by the preceding __syncthreads().
__device__ __forceinline__ void cluster_arrive(bool release)
{
    if (release) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    else asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int G, bool PAD>
__global__ void __launch_bounds__(TD_THREADS) k_sgbm_td(TdArgs a)
{
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: L[3][Mmax][Dp] u16 | halo[2 parity][2 dir][Dp] u16 | m[3][Mmax] u32 | halom[2][2] u32
    uint16_t* Lb = reinterpret_cast<uint16_t*>(smem_raw);
    uint16_t* halo = Lb + (size_t)3 * a.Mmax * a.Dp;
    unsigned* mb = reinterpret_cast<unsigned*>(halo + 4 * a.Dp);
    unsigned* halom = mb + 3 * a.Mmax;

    const int r = (int)cluster.block_rank();
    const int f = blockIdx.y;
    const int x0 = (int)(((long long)a.W1 * r) / a.NC), x1 = (int)(((long long)a.W1 * (r + 1)) / a.NC);
    const int M = x1 - x0;
    constexpr int NG = TD_THREADS / G;
    const int g = threadIdx.x / G, q = threadIdx.x % G;
    const bool padLane = q * 8 >= a.D;
    uint16_t* haloR = (r + 1 < a.NC) ? cluster.map_shared_rank(halo, r + 1) : nullptr;   // CTA owning columns x1..
    uint16_t* haloL = (r > 0) ? cluster.map_shared_rank(halo, r - 1) : nullptr;
    unsigned* halomR = (r + 1 < a.NC) ? cluster.map_shared_rank(halom, r + 1) : nullptr;
    unsigned* halomL = (r > 0) ? cluster.map_shared_rank(halom, r - 1) : nullptr;
    const int iters = (M + NG - 1) / NG;
    constexpr int Dp = 8 * G;                       // == a.Dp
    const int Mmax = a.Mmax, W1 = a.W1;
    const int rowElems = W1 * Dp;
    // per-thread base pointers; all further offsets are 32-bit element counts
    const uint16_t* __restrict__ cbase = a.C + ((size_t)f * a.H + (a.bottomUp ? a.H - 1 : 0)) * rowElems + (size_t)x0 * Dp + q * 8;
    uint16_t* __restrict__ sbase = a.S + ((size_t)f * a.H + (a.bottomUp ? a.H - 1 : 0)) * rowElems + (size_t)x0 * Dp + q * 8;
    const int rowStep = a.bottomUp ? -rowElems : rowElems;
    uint16_t* const L0b = Lb + q * 8;                               // diagonal from x-1
    uint16_t* const L1b = Lb + Mmax * Dp + q * 8;                   // vertical
    uint16_t* const L2b = Lb + 2 * Mmax * Dp + q * 8;               // diagonal from x+1
    // warps holding the lane group of the strip's first / last column write the neighbours' halos
    const bool haloWarp = __any_sync(FULL, (g == 0 && haloL != nullptr) || (g == (M - 1) % NG && haloR != nullptr)) != 0;
    cluster.sync();     // every CTA of the cluster is resident before any remote shared-memory access

    // work item (yi, it): pixel lx = g + it*NG of row yi; its C/S are loaded one item ahead
    const int lxFirst = min(g, M - 1);
    uint4 Cn = ld128(cbase + lxFirst * Dp), Sn = ld128(sbase + lxFirst * Dp);
    int ymod = 0;       // yi mod M
    int rowOff = 0;     // yi * rowStep (fits 32 bit: H*W1*Dp < 2^31 is checked on the host)
    for (int yi = 0; yi < a.H; ++yi) {
        const int par = yi & 1;
        const bool firstRow = yi == 0;
        for (int it = 0; it < iters; ++it) {
            int lx = g + it * NG;
            const bool active = lx < M;
            if (!active) lx = M - 1;
            const int x = x0 + lx;
            const int off = rowOff + lx * Dp;
            const uint4 Cc = Cn;
            uint4 Sc = Sn;
            {
                const bool lastIt = it + 1 == iters;
                const int nlx = min(lastIt ? g : g + (it + 1) * NG, M - 1);
                const int noff = (lastIt ? rowOff + rowStep : rowOff) + nlx * Dp;
                if (!(lastIt && yi + 1 == a.H)) { Cn = ld128(cbase + noff); Sn = ld128(sbase + noff); }
            }
            unsigned L[4], mm;
            // ---- vertical path, slot lx
            {
                uint16_t* sl = L1b + lx * Dp;
                if (firstRow) reset_state<PAD>(L, mm, padLane);
                else { ld_state(L, sl); mm = mb[Mmax + lx]; }
                sgm_step<G, PAD>(L, mm, Cc, a.P1P1, a.P2P2, q, padLane);
                if (active) { st128(sl, make_uint4(L[0], L[1], L[2], L[3])); if (q == 0) mb[Mmax + lx] = mm; }
                sat_acc(Sc, L);
            }
            // ---- diagonal with predecessor (x-1, previous row): slot (lx - yi) mod M, halo from the left CTA
            {
                int s1 = lx - ymod; if (s1 < 0) s1 += M;
                uint16_t* sl = L0b + s1 * Dp;
                if (firstRow || x == 0) reset_state<PAD>(L, mm, padLane);
                else if (lx == 0) { ld_state(L, halo + (par * 2 + 0) * Dp + q * 8); mm = halom[par * 2 + 0]; }
                else { ld_state(L, sl); mm = mb[s1]; }
                sgm_step<G, PAD>(L, mm, Cc, a.P1P1, a.P2P2, q, padLane);
                if (active) {
                    st128(sl, make_uint4(L[0], L[1], L[2], L[3])); if (q == 0) mb[s1] = mm;
                    if (lx == M - 1 && haloR) {
                        st128(haloR + ((par ^ 1) * 2 + 0) * Dp + q * 8, make_uint4(L[0], L[1], L[2], L[3]));
                        if (q == 0) halomR[(par ^ 1) * 2 + 0] = mm;
                    }
                }
                sat_acc(Sc, L);
            }
            // ---- diagonal with predecessor (x+1, previous row): slot (lx + yi) mod M, halo from the right CTA
            {
                int s3 = lx + ymod; if (s3 >= M) s3 -= M;
                uint16_t* sl = L2b + s3 * Dp;
                if (firstRow || x == W1 - 1) reset_state<PAD>(L, mm, padLane);
                else if (lx == M - 1) { ld_state(L, halo + (par * 2 + 1) * Dp + q * 8); mm = halom[par * 2 + 1]; }
                else { ld_state(L, sl); mm = mb[2 * Mmax + s3]; }
                sgm_step<G, PAD>(L, mm, Cc, a.P1P1, a.P2P2, q, padLane);
                if (active) {
                    st128(sl, make_uint4(L[0], L[1], L[2], L[3])); if (q == 0) mb[2 * Mmax + s3] = mm;
                    if (lx == 0 && haloL) {
                        st128(haloL + ((par ^ 1) * 2 + 1) * Dp + q * 8, make_uint4(L[0], L[1], L[2], L[3]));
                        if (q == 0) halomL[(par ^ 1) * 2 + 1] = mm;
                    }
                }
                sat_acc(Sc, L);
            }
            if (active) st128(sbase + off, Sc);
        }
        if (++ymod == M) ymod = 0;
        rowOff += rowStep;
        __syncthreads();
        cluster_arrive(haloWarp);
        cluster_wait();
    }
}

// ------------------------------------------------------------------------------------------------
// 16 disparities per lane: lane q of a G2-lane group owns octet q (registers A) and octet q+G2 (registers B) of
// the pixel (Dp = 16*G2), so that both 128-bit accesses of a warp stay perfectly coalesced.  One recurrence step
// shares the neighbour shuffles, the min butterfly and all per-pixel bookkeeping between the two octets.
// ------------------------------------------------------------------------------------------------
template <int G2, bool PAD>
__device__ __forceinline__ void sgm_step2(unsigned (&A)[4], unsigned (&B)[4], unsigned& mm, const uint4& Ca, const uint4& Cb,
                                          unsigned P1P1, unsigned P2P2, int q, bool padA, bool padB)
{
    // see sgm_step: neighbours are taken from L + P1 (plain adds), one VIMNMX3 per register
    const unsigned pa0 = A[0] + P1P1, pa1 = A[1] + P1P1, pa2 = A[2] + P1P1, pa3 = A[3] + P1P1;
    const unsigned pb0 = B[0] + P1P1, pb1 = B[1] + P1P1, pb2 = B[2] + P1P1, pb3 = B[3] + P1P1;
    unsigned upA = 0xffffffffu, dnA, upB, dnB = 0xffffffffu;
    if (G2 > 1) {
        const int nxt = (q + 1) & (G2 - 1), prv = (q + G2 - 1) & (G2 - 1);
        const unsigned x = __shfl_sync(FULL, pb0, nxt, G2);     // bottom of octet (q+1)+G2; lane G2-1 gets octet G2
        const unsigned y = __shfl_sync(FULL, pa0, nxt, G2);     // bottom of octet q+1
        const unsigned u = __shfl_sync(FULL, pa3, prv, G2);     // top of octet q-1; lane 0 gets octet G2-1
        const unsigned w = __shfl_sync(FULL, pb3, prv, G2);     // top of octet q-1+G2
        dnA = (q == G2 - 1) ? x : y;
        if (q != G2 - 1) dnB = x;
        if (q != 0) upA = u;
        upB = (q == 0) ? u : w;
    } else {
        dnA = pb0; upB = pa3;
    }
    const unsigned mP2 = mm + P2P2;
    const unsigned XA0 = __byte_perm(upA, pa0, 0x5432), XA1 = __byte_perm(pa0, pa1, 0x5432);
    const unsigned XA2 = __byte_perm(pa1, pa2, 0x5432), XA3 = __byte_perm(pa2, pa3, 0x5432);
    const unsigned XA4 = __byte_perm(pa3, dnA, 0x5432);
    const unsigned XB0 = __byte_perm(upB, pb0, 0x5432), XB1 = __byte_perm(pb0, pb1, 0x5432);
    const unsigned XB2 = __byte_perm(pb1, pb2, 0x5432), XB3 = __byte_perm(pb2, pb3, 0x5432);
    const unsigned XB4 = __byte_perm(pb3, dnB, 0x5432);
    unsigned a0 = __vimin3_u16x2(XA0, XA1, __vminu2(A[0], mP2)) + Ca.x - mm;
    unsigned a1 = __vimin3_u16x2(XA1, XA2, __vminu2(A[1], mP2)) + Ca.y - mm;
    unsigned a2 = __vimin3_u16x2(XA2, XA3, __vminu2(A[2], mP2)) + Ca.z - mm;
    unsigned a3 = __vimin3_u16x2(XA3, XA4, __vminu2(A[3], mP2)) + Ca.w - mm;
    unsigned b0 = __vimin3_u16x2(XB0, XB1, __vminu2(B[0], mP2)) + Cb.x - mm;
    unsigned b1 = __vimin3_u16x2(XB1, XB2, __vminu2(B[1], mP2)) + Cb.y - mm;
    unsigned b2 = __vimin3_u16x2(XB2, XB3, __vminu2(B[2], mP2)) + Cb.z - mm;
    unsigned b3 = __vimin3_u16x2(XB3, XB4, __vminu2(B[3], mP2)) + Cb.w - mm;
    if (PAD && padA) { a0 = a1 = a2 = a3 = MVSV_PK_MAX; }
    if (PAD && padB) { b0 = b1 = b2 = b3 = MVSV_PK_MAX; }
    unsigned m = __vimin3_u16x2(__vimin3_u16x2(a0, a1, a2), __vimin3_u16x2(a3, b0, b1), __vimin3_u16x2(b2, b3, b3));
#pragma unroll
    for (int o = G2 / 2; o > 0; o >>= 1) m = __vminu2(m, __shfl_xor_sync(FULL, m, o, G2));
    mm = __vminu2(m, __byte_perm(m, 0, 0x1032));
    A[0] = a0; A[1] = a1; A[2] = a2; A[3] = a3;
    B[0] = b0; B[1] = b1; B[2] = b2; B[3] = b3;
}

template <bool PAD>
__device__ __forceinline__ void reset_state2(unsigned (&A)[4], unsigned (&B)[4], unsigned& mm, bool padA, bool padB)
{
    const unsigned va = (PAD && padA) ? MVSV_PK_MAX : 0u, vb = (PAD && padB) ? MVSV_PK_MAX : 0u;
    A[0] = A[1] = A[2] = A[3] = va;
    B[0] = B[1] = B[2] = B[3] = vb;
    mm = 0u;
}

// ---- point-to-point halo hand-off between the CTAs of a cluster: asynchronous remote stores (st.async) that
// complete a transaction count on an mbarrier in the RECEIVER's shared memory.  Only the lane group that needs a
// halo ever waits; there is no cluster-wide barrier (and no release fence over outstanding global stores) in the
// row loop.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned cta)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void st_async_v4(unsigned raddr, const uint4& v, unsigned rbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_async_b32(unsigned raddr, unsigned v, unsigned rbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(v), "r"(rbar) : "memory");
}

// K3b'' -- the fused previous-row sweep (see k_sgbm_td above for the scheme) with 16 disparities per lane.
constexpr int TD2_THREADS = 512;      // launch bound; the launcher picks 256 or 512 threads per CTA

template <int G2, bool PAD>
__global__ void __launch_bounds__(TD2_THREADS) k_sgbm_td2(TdArgs a)
{
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: L[3][Mmax][Dp] u16 | halo[3 slots][2 dir][Dp] u16 | m[3][Mmax] u32 | halom[3][2] u32 | bars[3][2] u64
    // Halo slots are triple buffered (row y reads slot y % 3, a sender in row y fills slot (y+1) % 3): a neighbour
    // can be at most one row ahead, so the slot it fills is never the one still being read.
    constexpr int Dp = 16 * G2;                     // == a.Dp
    constexpr int OB = 8 * G2;                      // element offset of the lane's second octet
    // Bank-conflict swizzle of the state slots: a quarter-warp phase (8 lanes) covers 8/G2 pixels that each touch
    // one half (16*G2 bytes) of their slot; slots of 32*G2 bytes repeat every 4/G2 pixels in the 128-byte bank
    // space, so the two halves are exchanged for every other group of 4/G2 slots.  Keyed on the slot index, hence
    // the same for the writer and the reader of a slot.
    constexpr int SWS = (G2 == 4) ? 0 : (G2 == 2) ? 1 : (G2 == 1) ? 2 : -1;
    auto offA = [&](int slot) { return (SWS >= 0 && ((slot >> (SWS < 0 ? 0 : SWS)) & 1)) ? OB : 0; };
    uint16_t* Lb = reinterpret_cast<uint16_t*>(smem_raw);
    uint16_t* halo = Lb + (size_t)3 * a.Mmax * Dp;
    unsigned* mb = reinterpret_cast<unsigned*>(halo + 6 * Dp);
    unsigned* halom = mb + 3 * a.Mmax;
    // bars[slot][dir]: completes when the halo of that slot/direction has fully arrived (16-byte aligned region)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + td_bar_offset(a.Mmax, Dp));

    const int r = (int)cluster.block_rank();
    const int f = blockIdx.y;
    const int x0 = (int)(((long long)a.W1 * r) / a.NC), x1 = (int)(((long long)a.W1 * (r + 1)) / a.NC);
    const int M = x1 - x0;
    const int NG = blockDim.x / G2;
    const int g = threadIdx.x / G2, q = threadIdx.x % G2;
    const bool padA = q * 8 >= a.D, padB = (q + G2) * 8 >= a.D;
    const bool hasL = r > 0, hasR = r + 1 < a.NC;
    // shared::cluster addresses of the neighbours' halo / halom / mbarrier arrays (same layout in every CTA)
    const unsigned rHalo = hasR ? mapa_u32(smem_u32(halo), r + 1) : 0u, lHalo = hasL ? mapa_u32(smem_u32(halo), r - 1) : 0u;
    const unsigned rHalom = hasR ? mapa_u32(smem_u32(halom), r + 1) : 0u, lHalom = hasL ? mapa_u32(smem_u32(halom), r - 1) : 0u;
    const unsigned rBars = hasR ? mapa_u32(smem_u32(bars), r + 1) : 0u, lBars = hasL ? mapa_u32(smem_u32(bars), r - 1) : 0u;
    const unsigned haloBytes = Dp * 2 + 4;          // one pixel's path costs + its packed minimum
    const int iters = (M + NG - 1) / NG;
    const int Mmax = a.Mmax, W1 = a.W1;
    const int rowElems = W1 * Dp;
    const uint16_t* __restrict__ cbase = a.C + ((size_t)f * a.H + (a.bottomUp ? a.H - 1 : 0)) * rowElems + (size_t)x0 * Dp + q * 8;
    uint16_t* __restrict__ sbase = a.S + ((size_t)f * a.H + (a.bottomUp ? a.H - 1 : 0)) * rowElems + (size_t)x0 * Dp + q * 8;
    const int rowStep = a.bottomUp ? -rowElems : rowElems;
    uint16_t* const L0b = Lb + q * 8;                               // diagonal from x-1
    uint16_t* const L1b = Lb + Mmax * Dp + q * 8;                   // vertical
    uint16_t* const L2b = Lb + 2 * Mmax * Dp + q * 8;               // diagonal from x+1
    if (!PAD) {
        // "No predecessor" is the state (L = 0, m = 0).  Zero every slot once (first row) and both halos (the
        // frame's left / right border never receives a neighbour's write), so the row loop needs no reset tests.
        const int nwords = (3 * Mmax * Dp + 6 * Dp) / 2 + 3 * Mmax + 6;     // L, halo (u16 pairs), m, halom
        unsigned* wz = reinterpret_cast<unsigned*>(smem_raw);
        for (int i = threadIdx.x; i < nwords; i += blockDim.x) wz[i] = 0u;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(smem_u32(bars + i), 1);
        // arm the first phase of every halo that will be received: slot p, direction d = bars[p*2 + d]
        for (int pp = 0; pp < 3; ++pp) {
            if (hasL) mbar_expect_tx(smem_u32(bars + pp * 2 + 0), haloBytes);
            if (hasR) mbar_expect_tx(smem_u32(bars + pp * 2 + 1), haloBytes);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();     // every CTA of the cluster is resident, zeroed and armed before any remote access
    const int lxFirst = min(g, M - 1);
    uint4 CnA = ld128(cbase + lxFirst * Dp), CnB = ld128(cbase + lxFirst * Dp + OB);
    uint4 SnA = ld128(sbase + lxFirst * Dp), SnB = ld128(sbase + lxFirst * Dp + OB);
    int ymod = 0, rowOff = 0;
    for (int yi = 0; yi < a.H; ++yi) {
        const int par = yi % 3, parNext = (yi + 1) % 3;      // halo slot read in this row / filled for the next row
        const bool firstRow = yi == 0;
        for (int it = 0; it < iters; ++it) {
            int lx = g + it * NG;
            const bool active = lx < M;
            if (!active) lx = M - 1;
            const int x = x0 + lx;
            const int off = rowOff + lx * Dp;
            const uint4 CcA = CnA, CcB = CnB;
            uint4 ScA = SnA, ScB = SnB;
            {
                const bool lastIt = it + 1 == iters;
                const int nlx = min(lastIt ? g : g + (it + 1) * NG, M - 1);
                const int noff = (lastIt ? rowOff + rowStep : rowOff) + nlx * Dp;
                if (!(lastIt && yi + 1 == a.H)) {
                    CnA = ld128(cbase + noff); CnB = ld128(cbase + noff + OB);
                    SnA = ld128(sbase + noff); SnB = ld128(sbase + noff + OB);
                }
            }
            unsigned A[4], B[4], mm;
            // ---- vertical path, slot lx
            {
                uint16_t* sl = L1b + lx * Dp;
                const int oa = offA(lx), ob = OB - oa;
                ld_state(A, sl + oa); ld_state(B, sl + ob); mm = mb[Mmax + lx];
                if (PAD && firstRow) reset_state2<PAD>(A, B, mm, padA, padB);
                sgm_step2<G2, PAD>(A, B, mm, CcA, CcB, a.P1P1, a.P2P2, q, padA, padB);
                if (active) {
                    st128(sl + oa, make_uint4(A[0], A[1], A[2], A[3])); st128(sl + ob, make_uint4(B[0], B[1], B[2], B[3]));
                    if (q == 0) mb[Mmax + lx] = mm;
                }
                sat_acc(ScA, A); sat_acc(ScB, B);
            }
            // ---- diagonal with predecessor (x-1, previous row): slot (lx - yi) mod M, halo from the left CTA
            {
                int s1 = lx - ymod; if (s1 < 0) s1 += M;
                uint16_t* sl = L0b + s1 * Dp;
                const bool fromHalo = lx == 0;
                const int oa = offA(s1), ob = OB - oa;
                if (active && fromHalo && hasL && yi > 0) {
                    // halo slot `par` was filled by the left CTA during its row yi-1: use n of this slot
                    const int n = (yi - 1) / 3;
                    mbar_wait(smem_u32(bars + par * 2 + 0), n & 1);
                }
                const uint16_t* src = fromHalo ? halo + (par * 2 + 0) * Dp + q * 8 : sl;
                ld_state(A, src + (fromHalo ? 0 : oa)); ld_state(B, src + (fromHalo ? OB : ob)); mm = fromHalo ? halom[par * 2 + 0] : mb[s1];
                if (active && fromHalo && hasL && yi > 0 && q == 0 && yi + 3 < a.H)
                    mbar_expect_tx(smem_u32(bars + par * 2 + 0), haloBytes);                                // arm the next use
                if (PAD && (firstRow || x == 0)) reset_state2<PAD>(A, B, mm, padA, padB);
                sgm_step2<G2, PAD>(A, B, mm, CcA, CcB, a.P1P1, a.P2P2, q, padA, padB);
                if (active) {
                    st128(sl + oa, make_uint4(A[0], A[1], A[2], A[3])); st128(sl + ob, make_uint4(B[0], B[1], B[2], B[3]));
                    if (q == 0) mb[s1] = mm;
                    if (lx == M - 1 && hasR && yi + 1 < a.H) {
                        const int hs = parNext * 2 + 0;
                        const unsigned h = rHalo + (hs * Dp + q * 8) * 2, rb = rBars + hs * 8;
                        st_async_v4(h, make_uint4(A[0], A[1], A[2], A[3]), rb);
                        st_async_v4(h + OB * 2, make_uint4(B[0], B[1], B[2], B[3]), rb);
                        if (q == 0) st_async_b32(rHalom + hs * 4, mm, rb);
                    }
                }
                sat_acc(ScA, A); sat_acc(ScB, B);
            }
            // ---- diagonal with predecessor (x+1, previous row): slot (lx + yi) mod M, halo from the right CTA
            {
                int s3 = lx + ymod; if (s3 >= M) s3 -= M;
                uint16_t* sl = L2b + s3 * Dp;
                const bool fromHalo = lx == M - 1;
                const int oa = offA(s3), ob = OB - oa;
                if (active && fromHalo && hasR && yi > 0) {
                    const int n = (yi - 1) / 3;
                    mbar_wait(smem_u32(bars + par * 2 + 1), n & 1);
                }
                const uint16_t* src = fromHalo ? halo + (par * 2 + 1) * Dp + q * 8 : sl;
                ld_state(A, src + (fromHalo ? 0 : oa)); ld_state(B, src + (fromHalo ? OB : ob)); mm = fromHalo ? halom[par * 2 + 1] : mb[2 * Mmax + s3];
                if (active && fromHalo && hasR && yi > 0 && q == 0 && yi + 3 < a.H)
                    mbar_expect_tx(smem_u32(bars + par * 2 + 1), haloBytes);
                if (PAD && (firstRow || x == W1 - 1)) reset_state2<PAD>(A, B, mm, padA, padB);
                sgm_step2<G2, PAD>(A, B, mm, CcA, CcB, a.P1P1, a.P2P2, q, padA, padB);
                if (active) {
                    st128(sl + oa, make_uint4(A[0], A[1], A[2], A[3])); st128(sl + ob, make_uint4(B[0], B[1], B[2], B[3]));
                    if (q == 0) mb[2 * Mmax + s3] = mm;
                    if (lx == 0 && hasL && yi + 1 < a.H) {
                        const int hs = parNext * 2 + 1;
                        const unsigned h = lHalo + (hs * Dp + q * 8) * 2, lb = lBars + hs * 8;
                        st_async_v4(h, make_uint4(A[0], A[1], A[2], A[3]), lb);
                        st_async_v4(h + OB * 2, make_uint4(B[0], B[1], B[2], B[3]), lb);
                        if (q == 0) st_async_b32(lHalom + hs * 4, mm, lb);
                    }
                }
                sat_acc(ScA, A); sat_acc(ScB, B);
            }
            if (active) { st128(sbase + off, ScA); st128(sbase + off + OB, ScB); }
        }
        if (++ymod == M) ymod = 0;
        rowOff += rowStep;
        __syncthreads();        // the row's slot updates are visible CTA-wide; neighbours are paced by the mbarriers
    }
    cluster.sync();             // no CTA may exit while a neighbour could still address its shared memory
}

// K3c + K4: right-to-left path r=(+1,0) fused with winner-take-all, uniqueness, sub-pixel interpolation,
// the disp2 scatter (sequential in x per row, exactly the reference order) and the left-right check.
__device__ __forceinline__ unsigned pick16(const unsigned (&R)[4], int idx)
{
    const unsigned lo = (idx & 2) ? R[1] : R[0], hi = (idx & 2) ? R[3] : R[2];
    const unsigned r = (idx & 4) ? hi : lo;
    return __byte_perm(r, 0, (idx & 1) ? 0x4432 : 0x4410);
}

template <int G>
__device__ __forceinline__ bool group_any(bool v)
{
    const unsigned b = __ballot_sync(FULL, v);
    const unsigned gmask = (G == 32) ? FULL : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));
    return (b & gmask) != 0u;
}

// trunc(n / d) for |n| < 2^22, 0 < d < 2^20: float reciprocal estimate plus an exact two-sided correction
__device__ __forceinline__ int div_trunc_small(int n, int d)
{
    const int an = abs(n);
    int qq = __float2int_rz(__fdividef((float)an, (float)d));
    const int rem = an - qq * d;
    qq += (rem >= d) - (rem < 0);
    return n < 0 ? -qq : qq;
}

template <int G, bool PAD>
__global__ void __launch_bounds__(128) k_sgbm_h2_wta(AggArgs a)
{
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nrows = (long long)a.B * a.H;
    long long row = gtid / G;
    const int q = (int)(gtid % G);
    const bool active = row < nrows;
    if (!active) row = nrows - 1;
    const bool padLane = q * 8 >= a.D;
    constexpr int Dp = 8 * G;                       // == a.Dp
    const int W1 = a.W1;
    const size_t rowBase = (size_t)row * W1 * Dp + q * 8;
    const uint16_t* __restrict__ cp = a.C + rowBase;
    uint16_t* __restrict__ sp = a.S + rowBase;
    int16_t* __restrict__ drow = a.disp + (size_t)row * a.W;
    int* __restrict__ d2row = a.d2 + (size_t)row * a.W;
    // disp2 entry of right-image column j: (minS << 16) | (W1-1-xi) of the best left pixel xi that maps to j.  The
    // reference scans x downwards and replaces an entry only by a strictly smaller cost, i.e. it keeps the lowest
    // cost and among equals the largest x: exactly the minimum of this key, so the scatter is an order-free
    // atomicMin and the matched disparity is recovered as xi + minX1 - j.
    constexpr unsigned D2_EMPTY = 0xffffffffu;
    if (active)
        for (int x = q; x < a.W; x += G) { drow[x] = (int16_t)a.INV; d2row[x] = (int)D2_EMPTY; }
    __syncwarp();

    unsigned L[4], mm;
    reset_state<PAD>(L, mm, padLane);
    const int umul = 100 - a.uniq;
    const unsigned kb = (unsigned)q * 8u;
    // The per-pixel epilogue (parabola, division, scatter, store) is identical on all G lanes of a pixel, so it is
    // deferred: lane q keeps the winner of every G-th step and the G lanes finish G pixels at once.
    unsigned svKey = 0, svM = 0, svP = 0;
    int svX = -1, sc = 0;
    auto flush = [&]() {
        if (svX >= 0) {
            const int minS = (int)(svKey >> 16), best = (int)(svKey & 0xffffu);
            const int sm1 = (int)svM, sp1 = (int)svP;
            int v = 16 * best;
            if (best > 0 && best < a.D - 1) {
                const int den = max(sm1 + sp1 - 2 * minS, 1);
                v += div_trunc_small((sm1 - sp1) * 16 + den, 2 * den);
            }
            if (active) {
                atomicMin(reinterpret_cast<unsigned*>(d2row) + (svX + a.minX1 - best - a.minD),
                          ((unsigned)minS << 16) | (unsigned)(W1 - 1 - svX));
                drow[svX + a.minX1] = (int16_t)(v + 16 * a.minD);
            }
        }
        svX = -1;
    };
    auto step = [&](const uint4& Cq, const uint4& Sq, int xi) {
        sgm_step<G, PAD>(L, mm, Cq, a.P1P1, a.P2P2, q, padLane);
        unsigned Sf[4];
        Sf[0] = __viaddmin_u16x2(Sq.x, L[0], MVSV_PK_MAX);
        Sf[1] = __viaddmin_u16x2(Sq.y, L[1], MVSV_PK_MAX);
        Sf[2] = __viaddmin_u16x2(Sq.z, L[2], MVSV_PK_MAX);
        Sf[3] = __viaddmin_u16x2(Sq.w, L[3], MVSV_PK_MAX);
        if (PAD && padLane) Sf[0] = Sf[1] = Sf[2] = Sf[3] = MVSV_PK_MAX;
        if (a.storeS && active) st128(sp + xi * Dp, make_uint4(Sf[0], Sf[1], Sf[2], Sf[3]));
        // ---- first argmin via (S << 16 | k) keys
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
                const int kk = (int)kb + j;
                const int s = (int)((j & 1) ? (Sf[j >> 1] >> 16) : (Sf[j >> 1] & 0xffffu));
                bad |= (!PAD || kk < a.D) && (s * umul < minS * 100) && (abs(kk - best) > 1);
            }
            reject |= group_any<G>(bad);
        }
        // ---- neighbours of the winner for the parabola
        const int im = max(best - 1, 0), ip = min(best + 1, a.D - 1);
        unsigned vm = pick16(Sf, im & 7), vp = pick16(Sf, ip & 7);
        if (G > 1) {
            vm = __shfl_sync(FULL, vm, im >> 3, G);
            vp = __shfl_sync(FULL, vp, ip >> 3, G);
        }
        if (sc == q) { svKey = key; svM = vm; svP = vp; svX = reject ? -1 : xi; }
        if (++sc == G) { flush(); sc = 0; }
    };
    // operands of two steps are held one pair of steps ahead, in two register sets used alternately
    uint4 CA[HPF], SA[HPF], CB[HPF], SB[HPF];
    auto fetch = [&](uint4 (&Cd)[HPF], uint4 (&Sd)[HPF], int xfirst) {
#pragma unroll
        for (int k = 0; k < HPF; ++k) {
            const int xn = max(xfirst - k, 0);
            Cd[k] = ld128(cp + xn * Dp); Sd[k] = ld128(sp + xn * Dp);
        }
    };
    auto run = [&](const uint4 (&Cs)[HPF], const uint4 (&Ss)[HPF], int xfirst) {
#pragma unroll
        for (int k = 0; k < HPF; ++k)
            if (xfirst - k >= 0) step(Cs[k], Ss[k], xfirst - k);
    };
    fetch(CA, SA, W1 - 1);
    for (int x0 = W1 - 1; x0 >= 0; x0 -= 2 * HPF) {
        fetch(CB, SB, x0 - HPF);
        run(CA, SA, x0);
        fetch(CA, SA, x0 - 2 * HPF);
        run(CB, SB, x0 - HPF);
    }
    flush();
    __syncwarp();
    if (active) {
        for (int x = a.minX1 + q; x < a.maxX1; x += G) {
            const int d1 = drow[x];
            if (d1 == a.INV) continue;
            const int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
            const int _x = x - _d, x_ = x - d_;
            if (_x < 0 || _x >= a.W || x_ < 0 || x_ >= a.W) continue;
            // the entries were produced by atomics (performed in L2): read them past L1
            const unsigned ea = (unsigned)__ldcg(d2row + _x), eb = (unsigned)__ldcg(d2row + x_);
            // an untouched entry reads as the reference's initial value INVALID_DISP_SCALED = (minD-1)*16, which
            // passes the `>= minD` test below once minD >= 2 (kept: it is what cv::StereoSGBM computes)
            const int da = ea == D2_EMPTY ? a.INV : (W1 - 1 - (int)(ea & 0xffffu)) + a.minX1 - _x;
            const int db = eb == D2_EMPTY ? a.INV : (W1 - 1 - (int)(eb & 0xffffu)) + a.minX1 - x_;
            if (da >= a.minD && abs(da - _d) > a.d12 && db >= a.minD && abs(db - d_) > a.d12) drow[x] = (int16_t)a.INV;
        }
    }
}

__global__ void k_fill_i16(int16_t* p, size_t n, int16_t v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// threads per CTA of the 16-disparity sweep: 512 when the strip gives every lane group >= 2 pixels per row
inline int td2_threads(int Mmax, int G2)
{
    // The row barrier waits for the lane groups with the most pixels, so pick the CTA size (<= 512 threads,
    // whole warps) whose groups all get (almost) the same number k of pixels: largest size with >= 93 % of the
    // group-iterations doing useful work; e.g. 344 columns x 4 lanes -> 480 threads (120 groups x 3 = 360 slots).
    int best = (Mmax * G2 >= 1024) ? 512 : 256;
    double bestEff = 0.0;
    {
        const int groups = best / G2, k = (Mmax + groups - 1) / groups;
        bestEff = (double)Mmax / ((double)groups * k);
    }
    if (bestEff >= 0.93) return best;
    for (int k = 1; k <= 16; ++k) {
        const int groups = (Mmax + k - 1) / k;
        int threads = (groups * G2 + 31) / 32 * 32;
        if (threads > 512 || threads < 128) continue;
        const double eff = (double)Mmax / ((double)(threads / G2) * k);
        if (eff >= 0.93) return threads;          // smallest k = most threads first
    }
    return best;
}

inline size_t td_smem_bytes(int Mmax, int Dp)
{
    return td_bar_offset(Mmax, Dp) + 6 * 8;
}

template <int G, bool PAD>
void launch_sgbm_g(mvsv_ctx* c, int B)
{
    const SgbmNorm& n = c->sg;
    cudaStream_t st = c->stream;
    const size_t planeStrideR = (size_t)c->maxB * c->H * c->vsRP;
    {
        dim3 blk(PF_WARPS * 32), grd((c->W + PF_WARPS * PF_COLS - 1) / (PF_WARPS * PF_COLS), (c->H + PF_ROWS - 1) / PF_ROWS, 2 * B);
        KernelTimer kt(c, KID_SGBM_PREFILTER);
        k_sgbm_prefilter<<<grd, blk, 0, st>>>(c->rect[0], c->rect[1], c->pitch, c->W, c->H, n.ftzero, c->recL, c->plR,
                                              planeStrideR, c->vsRP, c->vsJOFF);
    }
    {
        constexpr int PX = VsGeom<G>::PX;
        VsArgs a;
        a.recL = c->recL; a.plR = c->plR; a.planeStrideR = planeStrideR;
        a.VS = c->VS; a.W = c->W; a.H = c->H; a.W1 = n.W1; a.D = n.D; a.Dp = n.Dp; a.minD = n.minD; a.minX1 = n.minX1;
        a.SH2 = n.SH2; a.NV = c->vsNV; a.LEN = 8 * c->vsNV + 8; a.RP = c->vsRP; a.JOFF = c->vsJOFF;
        const bool r8 = 2 * n.ftzero + 63 <= 255;
        using GE = VsGeom<G>;
        const size_t smem = (size_t)2 * 6 * 8 * a.LEN * 2 + (size_t)2 * PX * 12 * 4 +
                            (size_t)(2 * n.SH2 + 1) * GE::CT * (r8 ? 8 : 16) +
                            (size_t)VS_RD * GE::IPL * 2 * GE::NPT * 16 + (size_t)VS_RD * GE::RPL * GE::NPT * 8;
        // small batches: split the rows into bands until the grid covers the SMs (each band repeats bs-1 rows)
        const int gx = (n.W1 + PX - 1) / PX, bs = 2 * n.SH2 + 1;
        int bands = 1;
        while (bands < 8 && (long long)gx * B * (bands * 2) <= c->num_sms && c->H / (bands * 2) >= 4 * bs) bands *= 2;
        a.bandRows = (c->H + bands - 1) / bands;
        dim3 grd(gx, B, (c->H + a.bandRows - 1) / a.bandRows);
        KernelTimer kt(c, KID_SGBM_VSUM);
        if (bands > 1) {
            if (r8) k_sgbm_vsum<G, true, true><<<grd, VsGeom<G>::THREADS, smem, st>>>(a);
            else k_sgbm_vsum<G, false, true><<<grd, VsGeom<G>::THREADS, smem, st>>>(a);
        } else {
            if (r8) k_sgbm_vsum<G, true, false><<<grd, VsGeom<G>::THREADS, smem, st>>>(a);
            else k_sgbm_vsum<G, false, false><<<grd, VsGeom<G>::THREADS, smem, st>>>(a);
        }
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
    { KernelTimer kt(c, KID_SGBM_H1); k_sgbm_h1<G, PAD><<<rowBlocks, TPB, (size_t)(2 * n.SW2 + 1 + H1_PFD) * TPB * 16, st>>>(a); }
    // throughput: the smallest cluster that fits (clusters of 2 pack the SMs exactly); small batches: more CTAs per
    // frame as long as all clusters are still resident at once -- up to the non-portable size 16 (single pair: sweep
    // 1.12 -> 0.62 ms at D = 128, 0.67 -> 0.51 ms at cfg 2)
    int nc = c->td_nc;
    if (nc > 0 && !((c->debug_flags >> 8) & 0xff)) {
        auto cap = [&](int n) { int i = 0; while ((1 << i) < n) ++i; return c->td_nc_cap[i]; };
        while (nc < 16 && ((unsigned)(nc << 1) & c->td_nc_mask) && B <= cap(nc << 1)) nc <<= 1;
    }
    auto vdirs = [&](int bottomUp) {
        if (nc > 0) {
            TdArgs t;
            t.C = c->C; t.S = c->S; t.H = c->H; t.W1 = n.W1; t.D = n.D; t.Dp = n.Dp; t.NC = nc; t.Mmax = (n.W1 + nc - 1) / nc;
            t.bottomUp = bottomUp; t.P1P1 = a.P1P1; t.P2P2 = a.P2P2;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(nc, B, 1); cfg.blockDim = dim3(TD_THREADS, 1, 1);
            cfg.dynamicSmemBytes = td_smem_bytes(t.Mmax, n.Dp); cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            KernelTimer kt(c, KID_SGBM_TD);
            if constexpr (G >= 2) {
                cfg.blockDim = dim3(td2_threads(t.Mmax, G / 2), 1, 1);
                cudaLaunchKernelEx(&cfg, k_sgbm_td2<G / 2, PAD>, t);
            } else {
                cudaLaunchKernelEx(&cfg, k_sgbm_td<G, PAD>, t);
            }
        } else {
            { KernelTimer kt(c, KID_SGBM_VDIR); k_sgbm_vdir<G, PAD><<<colBlocks, TPB, 0, st>>>(a, -1, bottomUp); }
            { KernelTimer kt(c, KID_SGBM_VDIR); k_sgbm_vdir<G, PAD><<<colBlocks, TPB, 0, st>>>(a, 0, bottomUp); }
            { KernelTimer kt(c, KID_SGBM_VDIR); k_sgbm_vdir<G, PAD><<<colBlocks, TPB, 0, st>>>(a, +1, bottomUp); }
        }
    };
    vdirs(0);
    if (n.mode == 1) vdirs(1);
    { KernelTimer kt(c, KID_SGBM_H2_WTA); k_sgbm_h2_wta<G, PAD><<<rowBlocks, TPB, 0, st>>>(a); }
}

template <int G>
cudaError_t cfg_vsum()
{
    cudaError_t e = cudaFuncSetAttribute(k_sgbm_vsum<G, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_sgbm_vsum<G, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_sgbm_vsum<G, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_sgbm_vsum<G, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_sgbm_h1<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_sgbm_h1<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_sgbm_td<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TD_SMEM_LIMIT);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_sgbm_td<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TD_SMEM_LIMIT);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_sgbm_td<G, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_sgbm_td<G, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    if constexpr (G >= 2) {
        e = cudaFuncSetAttribute(k_sgbm_td2<G / 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TD_SMEM_LIMIT);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_sgbm_td2<G / 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TD_SMEM_LIMIT);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_sgbm_td2<G / 2, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_sgbm_td2<G / 2, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <int G>
int td_max_clusters(int nc, size_t smem, int Mmax)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nc, 1, 1); cfg.blockDim = dim3(G >= 2 ? td2_threads(Mmax, G / 2) : TD_THREADS, 1, 1); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e;
    if constexpr (G >= 2) e = cudaOccupancyMaxActiveClusters(&n, k_sgbm_td2<G / 2, false>, &cfg);
    else e = cudaOccupancyMaxActiveClusters(&n, k_sgbm_td<G, false>, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
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

// Cluster size for the fused previous-row sweep: the smallest power of two whose column strip fits in shared
// memory (clusters of 2 pack the 148 SMs exactly; clusters of 4 strand 16 of them, measured).
// 0 = does not fit (or clusters unavailable): fall back to the three independent k_sgbm_vdir passes.
int sgbm_choose_td_cluster(mvsv_ctx* c)
{
    const SgbmNorm& n = c->sg;
    c->td_nc_mask = 0;
    for (int& v : c->td_nc_cap) v = 0;
    if (n.W1 <= 0) return 0;
    const int forced = (int)((c->debug_flags >> 8) & 0xff);
    if (forced == 0xff) return 0;
    int smallest = 0;
    // clusters of 16 (non-portable) fit only a handful at a time: for throughput they were measured slower than the
    // independent passes (95 % of HBM peak), so size 16 never becomes the default -- it is only recorded as available
    // for the small-batch case (or forced by the test hook)
    for (int nc = 1, idx = 0; nc <= 16; nc <<= 1, ++idx) {
        if (forced && nc != forced) continue;
        if (nc > n.W1) break;
        const int Mmax = (n.W1 + nc - 1) / nc;
        const size_t smem = td_smem_bytes(Mmax, n.Dp);
        if (smem > (size_t)TD_SMEM_LIMIT) continue;
        int ok = 0;
        switch (n.G) {
            case 1: ok = td_max_clusters<1>(nc, smem, Mmax); break;
            case 2: ok = td_max_clusters<2>(nc, smem, Mmax); break;
            case 4: ok = td_max_clusters<4>(nc, smem, Mmax); break;
            case 8: ok = td_max_clusters<8>(nc, smem, Mmax); break;
            case 16: ok = td_max_clusters<16>(nc, smem, Mmax); break;
            default: ok = td_max_clusters<32>(nc, smem, Mmax); break;
        }
        if (ok > 0) {
            c->td_nc_mask |= (unsigned)nc;
            c->td_nc_cap[idx] = ok;
            if (!smallest && (forced || nc <= 8)) smallest = nc;
        }
    }
    return smallest;
}

// Geometry of the reversed right-image planes for the cost kernel (see k_sgbm_prefilter / k_sgbm_vsum).
void sgbm_plane_geometry(const SgbmNorm& n, int W, int* NV, int* RP, int* JOFF)
{
    const int PX = vs_compute_threads(n.G) / n.G;
    const int NE = PX - 1 + n.Dp;                    // entries a CTA can touch
    const int nv = (NE + 2 + 7) / 8;                 // == VsGeom<G>::NV
    const int K = W - PX - n.minX1 + n.minD;         // j0 = JOFF + K - xa must be a multiple of 8 (xa is)
    const int joff = PX + 8 + (((8 - ((PX + 8 + K) % 8)) % 8 + 8) % 8);
    *NV = nv; *JOFF = joff; *RP = (joff + W + nv * 8 + 16 + 7) / 8 * 8;
}

void launch_sgbm(mvsv_ctx* c, int B)
{
    const SgbmNorm& n = c->sg;
    const size_t npx = (size_t)B * c->H * c->W;
    if (n.W1 <= 0) {
        KernelTimer kt(c, KID_FILL);
        k_fill_i16<<<(unsigned)((npx + 255) / 256), 256, 0, c->stream>>>(c->disp_raw, npx, (int16_t)n.INV);
    } else {
        const bool pad = n.D != n.Dp;
        switch (n.G) {
            case 1: if (pad) launch_sgbm_g<1, true>(c, B); else launch_sgbm_g<1, false>(c, B); break;
            case 2: if (pad) launch_sgbm_g<2, true>(c, B); else launch_sgbm_g<2, false>(c, B); break;
            case 4: if (pad) launch_sgbm_g<4, true>(c, B); else launch_sgbm_g<4, false>(c, B); break;
            case 8: if (pad) launch_sgbm_g<8, true>(c, B); else launch_sgbm_g<8, false>(c, B); break;
            case 16: if (pad) launch_sgbm_g<16, true>(c, B); else launch_sgbm_g<16, false>(c, B); break;
            default: if (pad) launch_sgbm_g<32, true>(c, B); else launch_sgbm_g<32, false>(c, B); break;
        }
    }
    if (n.speckleWin > 0) {
        launch_median(c, c->disp_raw, c->disp_med, B);
        launch_speckle(c, c->disp_med, c->disp, B, n.INV, n.speckleWin, 16 * n.speckleRange);
    } else {
        launch_median(c, c->disp_raw, c->disp, B);
    }
}
