// SOR kernels (S/OpticalFlow.cpp:451-505).
//
//   k_sor_wavefront   lexicographic Gauss-Seidel, the reference's exact update order, run as a
//                     pipelined anti-diagonal wavefront in one cooperative launch (parity mode).
//   k_sor_rb_tile     red-black SOR with several sweeps fused per launch (temporal blocking),
//                     coefficients and du/dv resident in registers (fast mode).
//   k_sor_rb_half     one red-black half-sweep straight from global memory (cross-check only).
//
// Per pixel p=(i,j) the reference computes
//     s1 = sum_nbr w du_nbr, s2 = sum_nbr w dv_nbr            w: left phi(p-1), right phi(p),
//     s1 *= -alpha; s2 *= -alpha; s1 += dxy dv                   up phi(p-W), down phi(p)
//     du = (1-omega) du + omega/(dx2 + .05 alpha + alpha sum_nbr w) * (bu - s1)
//     s2 += dxy du(new);  dv likewise with dy2, bv
// The omega/(...) factors iu, iv are loop invariant and come precomputed from k_assemble.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "tma.cuh"

namespace pf {
namespace cg = cooperative_groups;

template <typename T>
struct SorArgs {
    const T *phi, *dxy, *iu, *iv, *bu, *bv;
    T *du, *dv;              // in/out (wavefront, half-sweep) or output (tile kernel)
    const T *du_in, *dv_in;  // tile kernel input (ping-pong: other CTAs read halos of the input)
    int w, h, pitch;
    T alpha, omega;
};

// ------------------------------------------------------------------------------------------------
// Parity mode.  Pixel (i,j) of sweep s needs (i,j-1),(i-1,j) of sweep s and (i,j+1),(i+1,j) of sweep
// s-1, so anti-diagonal d=i+j of sweep s can run at step t = d + 2s: all nsor sweeps are pipelined
// through (W+H-1) + 2(nsor-1) grid-synchronised steps instead of nsor*(W+H-1).  At any step the
// written diagonals have the parity of t and the read neighbours the opposite parity, so the
// in-place update is race free and reproduces the lexicographic order exactly.
// du/dv are read with ld.global.cg (L2) because other SMs wrote them in the previous step.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_sor_wavefront(SorArgs<T> a, int nsor) {
    cg::grid_group grid = cg::this_grid();
    const int W = a.w, H = a.h, P = a.pitch;
    const int nd = W + H - 1, L = min(W, H);
    const int nsteps = nd + 2 * (nsor - 1);
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const T one_m = (T)1 - a.omega;
    const T nalpha = -a.alpha;
    for (int t = 0; t < nsteps; t++) {
        int s_hi = min(nsor - 1, t >> 1);
        int over = t - (nd - 1);
        int s_lo = over <= 0 ? 0 : (over + 1) >> 1;
        long long work = (long long)(s_hi - s_lo + 1) * L;
        for (long long idx = tid; idx < work; idx += nthreads) {
            int s = s_lo + (int)(idx / L), m = (int)(idx % L);
            int d = t - 2 * s;
            int i = max(0, d - (W - 1)) + m;
            if (i > min(H - 1, d)) continue;
            int j = d - i;
            size_t o = (size_t)i * P + j;
            T s1 = 0, s2 = 0, wt;
            if (j > 0)     { wt = a.phi[o - 1]; s1 += wt * __ldcg(a.du + o - 1); s2 += wt * __ldcg(a.dv + o - 1); }
            if (j < W - 1) { wt = a.phi[o];     s1 += wt * __ldcg(a.du + o + 1); s2 += wt * __ldcg(a.dv + o + 1); }
            if (i > 0)     { wt = a.phi[o - P]; s1 += wt * __ldcg(a.du + o - P); s2 += wt * __ldcg(a.dv + o - P); }
            if (i < H - 1) { wt = a.phi[o];     s1 += wt * __ldcg(a.du + o + P); s2 += wt * __ldcg(a.dv + o + P); }
            s1 *= nalpha;
            s2 *= nalpha;
            T du = __ldcg(a.du + o), dv = __ldcg(a.dv + o), dxy = a.dxy[o];
            s1 += dxy * dv;
            du = one_m * du + a.iu[o] * (a.bu[o] - s1);
            s2 += dxy * du;
            dv = one_m * dv + a.iv[o] * (a.bv[o] - s2);
            a.du[o] = du;
            a.dv[o] = dv;
        }
        grid.sync();
    }
}

// ------------------------------------------------------------------------------------------------
// Cross-check: one colour of one red-black sweep, in place, one thread per updated pixel.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_sor_rb_half(SorArgs<T> a, int colour) {
    int y = blockIdx.y;
    int x = 2 * (blockIdx.x * blockDim.x + threadIdx.x) + ((y + colour) & 1);
    if (x >= a.w) return;
    const int W = a.w, H = a.h, P = a.pitch;
    size_t o = (size_t)y * P + x;
    T s1 = 0, s2 = 0, wt;
    if (x > 0)     { wt = a.phi[o - 1]; s1 += wt * a.du[o - 1]; s2 += wt * a.dv[o - 1]; }
    if (x < W - 1) { wt = a.phi[o];     s1 += wt * a.du[o + 1]; s2 += wt * a.dv[o + 1]; }
    if (y > 0)     { wt = a.phi[o - P]; s1 += wt * a.du[o - P]; s2 += wt * a.dv[o - P]; }
    if (y < H - 1) { wt = a.phi[o];     s1 += wt * a.du[o + P]; s2 += wt * a.dv[o + P]; }
    s1 *= -a.alpha;
    s2 *= -a.alpha;
    T du = a.du[o], dv = a.dv[o], dxy = a.dxy[o];
    s1 += dxy * dv;
    du = ((T)1 - a.omega) * du + a.iu[o] * (a.bu[o] - s1);
    s2 += dxy * du;
    dv = ((T)1 - a.omega) * dv + a.iv[o] * (a.bv[o] - s2);
    a.du[o] = du;
    a.dv[o] = dv;
}

// ------------------------------------------------------------------------------------------------
// Fast mode: temporally blocked red-black SOR, register resident.
//
// A CTA owns a region of 64 x (NW*R) pixels: warp `wp` owns rows [wp*R, wp*R+R), lane `l` owns the
// column pair {2l, 2l+1}.  All eight planes of the thread's 2 x R patch (alpha*phi, dxy, iu, iv, bu,
// bv, du, dv) are loaded ONCE with 8-byte (16-byte in FP64) coalesced vector loads and stay in
// registers for `nsw` fused sweeps; region origins and R are even, so the pixel updated in row r of
// colour c is the compile-time column p = (r + c) & 1 and every lane does identical work.
// Horizontal neighbours in the other lane travel by warp shuffle, vertical neighbours across warps
// through one row pair of shared memory per warp and one __syncthreads per half-sweep.
// Pixels outside the image are held at exactly zero with zero coefficients, so the reference's
// "skip the missing neighbour" rule (S/OpticalFlow.cpp:468-495) needs no branches.
// Regions overlap: a pixel at distance k from a region edge that is not an image edge is wrong after
// k half-sweeps, so only the window 2*nsw inside such edges is written back, to a second buffer
// (other CTAs still read the old values of their halos).
// Algorithmic traffic per pixel-sweep drops from 10 words to about (8/eff + 2)/nsw, eff = window/region.
// ------------------------------------------------------------------------------------------------
template <typename T> struct Vec2;
template <> struct Vec2<float>  { typedef float2 type; };
template <> struct Vec2<double> { typedef double2 type; };

constexpr int kSorRegionW = 64;

template <typename T, int R, int NW>
__global__ void __launch_bounds__(NW * 32, 1) k_sor_rb_tile(SorArgs<T> a, int nsw, int step_x, int step_y) {
    static_assert(R % 2 == 0, "R must be even so that pixel colour is a compile-time function of (r,p)");
    typedef typename Vec2<T>::type V2;
    constexpr int RH = NW * R;
    __shared__ T ex[2][NW][2][kSorRegionW];  // [du|dv][warp][top|bottom][x]

    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int W = a.w, H = a.h, P = a.pitch;
    const int HL = 2 * nsw;
    const int rx0 = blockIdx.x * step_x, ry0 = blockIdx.y * step_y;  // even by construction
    const int xa = rx0 + 2 * lane, ya = ry0 + wp * R;

    T w[R][2], dxy[R][2], iu[R][2], iv[R][2], bu[R][2], bv[R][2], du[R][2], dv[R][2];
    T wl[R], wu[2];

    auto ld2 = [&](const T* base, int y, T& v0, T& v1) {
        if (y < H && xa < W) {
            V2 t = *reinterpret_cast<const V2*>(base + (size_t)y * P + xa);
            v0 = t.x;
            v1 = (xa + 1 < W) ? t.y : (T)0;
        } else {
            v0 = 0;
            v1 = 0;
        }
    };
#pragma unroll
    for (int r = 0; r < R; r++) {
        int y = ya + r;
        ld2(a.phi, y, w[r][0], w[r][1]);
        ld2(a.dxy, y, dxy[r][0], dxy[r][1]);
        ld2(a.iu, y, iu[r][0], iu[r][1]);
        ld2(a.iv, y, iv[r][0], iv[r][1]);
        ld2(a.bu, y, bu[r][0], bu[r][1]);
        ld2(a.bv, y, bv[r][0], bv[r][1]);
        if (a.du_in) {                       // nullptr: the solve starts from du = dv = 0
            ld2(a.du_in, y, du[r][0], du[r][1]);
            ld2(a.dv_in, y, dv[r][0], dv[r][1]);
        } else {
            du[r][0] = du[r][1] = dv[r][0] = dv[r][1] = 0;
        }
    }
    // weight of the row above the patch (phi at y-1) -- zero outside the image
    {
        int y = ya - 1;
        if (y >= 0) ld2(a.phi, y, wu[0], wu[1]);
        else { wu[0] = 0; wu[1] = 0; }
        wu[0] *= a.alpha;
        wu[1] *= a.alpha;
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        w[r][0] *= a.alpha;
        w[r][1] *= a.alpha;
        // weight towards the left lane's odd column, phi(xa-1, y): zero for lane 0 when rx0 == 0
        T t = __shfl_up_sync(0xffffffffu, w[r][1], 1);
        if (lane == 0) {
            int y = ya + r;
            t = (xa > 0 && y < H) ? a.alpha * a.phi[(size_t)y * P + xa - 1] : (T)0;
        }
        wl[r] = t;
    }
    const T one_m = (T)1 - a.omega;

    auto publish = [&]() {
        *reinterpret_cast<V2*>(&ex[0][wp][0][2 * lane]) = V2{du[0][0], du[0][1]};
        *reinterpret_cast<V2*>(&ex[1][wp][0][2 * lane]) = V2{dv[0][0], dv[0][1]};
        *reinterpret_cast<V2*>(&ex[0][wp][1][2 * lane]) = V2{du[R - 1][0], du[R - 1][1]};
        *reinterpret_cast<V2*>(&ex[1][wp][1][2 * lane]) = V2{dv[R - 1][0], dv[R - 1][1]};
    };
    publish();
    __syncthreads();

    for (int s = 0; s < nsw; s++) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
            // vertical neighbours owned by other warps (zero at the region's top / bottom edge)
            const int p_top = c & 1, p_bot = (R - 1 + c) & 1;
            T up_du = 0, up_dv = 0, dn_du = 0, dn_dv = 0;
            if (wp > 0) {
                up_du = ex[0][wp - 1][1][2 * lane + p_top];
                up_dv = ex[1][wp - 1][1][2 * lane + p_top];
            }
            if (wp < NW - 1) {
                dn_du = ex[0][wp + 1][0][2 * lane + p_bot];
                dn_dv = ex[1][wp + 1][0][2 * lane + p_bot];
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int p = (r + c) & 1;
                // horizontal neighbours: one is the thread's own other column, one lives in a lane
                T lw, ldu, ldv, rdu, rdv;
                if (p == 1) {
                    lw = w[r][0]; ldu = du[r][0]; ldv = dv[r][0];
                    rdu = __shfl_down_sync(0xffffffffu, du[r][0], 1);
                    rdv = __shfl_down_sync(0xffffffffu, dv[r][0], 1);
                    if (lane == 31) { rdu = 0; rdv = 0; }
                } else {
                    lw = wl[r];
                    ldu = __shfl_up_sync(0xffffffffu, du[r][1], 1);
                    ldv = __shfl_up_sync(0xffffffffu, dv[r][1], 1);
                    if (lane == 0) { ldu = 0; ldv = 0; }
                    rdu = du[r][1]; rdv = dv[r][1];
                }
                T uw, udu, udv, ddu, ddv;
                if (r > 0) { uw = w[r - 1][p]; udu = du[r - 1][p]; udv = dv[r - 1][p]; }
                else       { uw = wu[p];       udu = up_du;        udv = up_dv; }
                if (r < R - 1) { ddu = du[r + 1][p]; ddv = dv[r + 1][p]; }
                else           { ddu = dn_du;        ddv = dn_dv; }
                const T cw = w[r][p];
                T s1 = bu[r][p] + lw * ldu + cw * rdu + uw * udu + cw * ddu;
                T s2 = bv[r][p] + lw * ldv + cw * rdv + uw * udv + cw * ddv;
                s1 -= dxy[r][p] * dv[r][p];
                T nu = one_m * du[r][p] + iu[r][p] * s1;
                s2 -= dxy[r][p] * nu;
                T nv = one_m * dv[r][p] + iv[r][p] * s2;
                du[r][p] = nu;
                dv[r][p] = nv;
            }
            __syncthreads();   // everyone has consumed the previous exchange rows
            publish();
            __syncthreads();
        }
    }

    // write back the window that is still exact
    const int ox_lo = blockIdx.x > 0 ? rx0 + HL : 0;
    const int ox_hi = (rx0 + kSorRegionW >= W) ? W : rx0 + kSorRegionW - HL;
    const int oy_lo = blockIdx.y > 0 ? ry0 + HL : 0;
    const int oy_hi = (ry0 + RH >= H) ? H : ry0 + RH - HL;
#pragma unroll
    for (int r = 0; r < R; r++) {
        int y = ya + r;
        if (y < oy_lo || y >= oy_hi) continue;
        bool v0 = xa >= ox_lo && xa < ox_hi, v1 = xa + 1 >= ox_lo && xa + 1 < ox_hi;
        size_t o = (size_t)y * P + xa;
        if (v0 && v1) {
            *reinterpret_cast<V2*>(a.du + o) = V2{du[r][0], du[r][1]};
            *reinterpret_cast<V2*>(a.dv + o) = V2{dv[r][0], dv[r][1]};
        } else if (v0) {
            a.du[o] = du[r][0];
            a.dv[o] = dv[r][0];
        } else if (v1) {
            a.du[o + 1] = du[r][1];
            a.dv[o + 1] = dv[r][1];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fast mode, production variant: the same register-resident red-black update as k_sor_rb_tile, run
// as a PERSISTENT kernel (one CTA per SM, tiles dealt round-robin) whose next tile is staged by TMA
// while the current one is being swept.
//
// Per tile eight 2-D bulk tensor copies (cp.async.bulk.tensor -> SASS UTMALDG), one per warp, fill a
// shared-memory stage -- phi as a 72 x (RH+1) box at (rx0-4, ry0-1) so that the weights of the
// left / upper neighbours outside the region come along, the other seven planes as 64 x RH boxes --
// and thread 0 arms one mbarrier with the byte count.  TMA zero-fills everything outside the
// plane (negative coordinates, row padding, rows past the image), which is exactly the "phantom
// pixel" convention of the update.  Threads wait on the mbarrier, pull their 2 x R patch from shared
// memory into registers with conflict-free 8-byte loads, release the stage with one __syncthreads,
// the next tile's copies are issued immediately and overlap the 2*nsw half-sweeps and the
// write-back.  HBM/L2 reads, register compute and stores of different tiles therefore overlap on
// every SM instead of alternating in lock-step across the chip.
// ------------------------------------------------------------------------------------------------
struct SorMaps {
    CUtensorMap phi, dxy, iu, iv, bu, bv, du, dv;   // du/dv: the INPUT buffers of this pass
};

// Row-band split over several GPUs: output rows [up_lo, up_hi) are ALSO stored straight into the
// upper neighbour's output planes and rows [dn_lo, dn_hi) into the lower neighbour's (peer-mapped
// pointers, posted stores over NVLink), which is the halo the neighbour's next pass reads.  Empty
// ranges / null pointers on a single GPU.
//
// Ordering between the passes of neighbouring bands without the host or stream events: every CTA of a pass bumps a
// counter in each neighbour's memory once its stores (own planes and pushed halo rows) are out (release, system
// scope); the CTAs of the neighbour's NEXT pass spin on that counter (acquire) until all CTAs of this pass have
// arrived, before they issue their first tile load.  wait[0] / signal[0] belong to the upper neighbour, [1] to the
// lower one; null pointers = no such neighbour / first or last pass of a solve (those are ordered by events).
template <typename T>
struct SorPeer {
    T *up_du = nullptr, *up_dv = nullptr, *dn_du = nullptr, *dn_dv = nullptr;
    int up_lo = 0, up_hi = 0, dn_lo = 0, dn_hi = 0;
    unsigned int* wait_flag[2] = {nullptr, nullptr};     // counters in THIS device's memory
    unsigned int wait_count[2] = {0, 0};                 // CTAs of the neighbour's previous pass
    unsigned int* signal_flag[2] = {nullptr, nullptr};   // counters in the neighbours' memory (peer mapped)
    unsigned int* error_word = nullptr;                  // set (this device's memory) when a wait ran out of time
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_sys_add(unsigned int* p, unsigned int v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename T, int R, int NW>
struct SorStage {
    static constexpr int RH = NW * R;
    static constexpr int PHW = 72, PHH = RH + 1;   // phi box starts at rx0-4: TMA needs 16-byte aligned inner coordinates
    alignas(128) T phi[PHH][PHW];
    alignas(128) T pl[7][RH][kSorRegionW];   // dxy, iu, iv, bu, bv, du, dv (each plane a multiple of 128 B)
};

#ifndef PF_SOR_MINB
#define PF_SOR_MINB 1
#endif
// The stage is handed back to TMA in PF_SOR_GROUPS groups of planes ({phi,dxy} {iu,iv} {bu,bv} {du,dv} for 4), each with
// its own mbarrier: as soon as every thread has pulled a group into registers the same planes of the NEXT tile are
// requested, so the loads of a tile start while the rest of the previous one is still being unpacked and the unpacking of
// a tile starts while its last planes are still in flight.  With one group (round 1) the whole stage was released at
// once: the L2 -> SM stream (131 KB per tile, ~3.7 us per wave of 148 tiles) idled during the 0.75 us of unpacking and
// 0.8 us of it stayed exposed after the sweeps.  1, 2 or 4.
#ifndef PF_SOR_GROUPS
#define PF_SOR_GROUPS 1
#endif
// developer build (tools/build_variant.sh stats -DPF_SOR_STATS=1): CTAs 0 and 100 print where the cycles of their tile
// visits went (waiting for the stage, stage -> registers, sweeps, write-back)
#ifndef PF_SOR_STATS
#define PF_SOR_STATS 0
#endif
// developer ablation builds (tools/sor_ablation.sh): bit 0 = no sweeps, bit 1 = no write-back, bit 2 = no TMA loads
#ifndef PF_SORX
#define PF_SORX 0
#endif
#ifdef PF_SOR_MAXREG
template <typename T, int R, int NW>
__global__ void __maxnreg__(sizeof(T) == 4 ? PF_SOR_MAXREG : 128)
k_sor_rb_tma(
#else
template <typename T, int R, int NW>
__global__ void __launch_bounds__(NW * 32, PF_SOR_MINB)
k_sor_rb_tma(
#endif
const __grid_constant__ SorMaps maps, T* __restrict__ du_out, T* __restrict__ dv_out, int W, int H, int P,
             T alpha, T omega, int nsw, int has_input, int ntx, int nty, int step_x, int step_y, int ty0, SorPeer<T> peer) {
    static_assert(R % 2 == 0, "R must be even so that pixel colour is a compile-time function of (r,p)");
    typedef typename Vec2<T>::type V2;
    typedef SorStage<T, R, NW> Stage;
    constexpr int RH = NW * R;
    // TMA destinations must be 128-byte aligned: declared on the dynamic segment (an integer round-up
    // would demote every stage access from LDS to a generic LD)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage& st = *reinterpret_cast<Stage*>(smem_raw);
    // exchange rows between warps, double buffered by half-sweep parity so that ONE barrier per
    // half-sweep suffices: [buffer][du|dv][warp][top|bottom][x]
    typedef T ExBuf[2][NW][2][kSorRegionW];
    ExBuf* ex = reinterpret_cast<ExBuf*>(smem_raw + sizeof(Stage));
    constexpr int NG = PF_SOR_GROUPS, PPG = 8 / NG;
    static_assert(NG == 1 || NG == 2 || NG == 4, "planes per group must divide 8 and keep du, dv together");
    __shared__ __align__(8) uint64_t full_bar[NG];

    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int HL = 2 * nsw;
    const int ntiles = ntx * nty;
    const T one_m = (T)1 - omega;

    // plane p (0 phi, 1 dxy, 2 iu, 3 iv, 4 bu, 5 bv, 6 du, 7 dv) travels in group p / PPG.  The copies of a group are
    // issued by DIFFERENT warps (lane 0 of warp p issues plane p) and the byte count is armed by thread 0: issued by one
    // thread, the eight UTMALDG (each a uniform-datapath sequence of ~60 cycles) kept warp 0, and with it the first
    // half-sweep barrier of the whole CTA, ~500 cycles behind on every tile (tools/sor_stats.py).  A copy may complete
    // before the count is armed: the transaction count goes negative and the phase cannot complete without the arrival.
    // (NW < 8, the small single-region variants: warp w issues planes w, w + NW, ...)
    auto issue_group = [&](int tile, int g) {
        if (PF_SORX & 4) return;
        const int p_lo = g * PPG, p_hi = min(p_lo + PPG, has_input ? 8 : 6);
        if (p_lo >= p_hi) return;                           // the du/dv group of a pass that starts from zero
        if (tid == 0) {
            uint32_t bytes = (uint32_t)(sizeof(T) * RH * kSorRegionW) * (uint32_t)(p_hi - p_lo);
            if (p_lo == 0) bytes += (uint32_t)(sizeof(T) * (Stage::PHH * Stage::PHW - RH * kSorRegionW));
            mbar_expect_tx(&full_bar[g], bytes);
        }
        if (lane == 0) {
            const int tx = tile % ntx, ty = tile / ntx + ty0;   // ty0: first tile row of this launch (row-band split)
            const int rx0 = tx * step_x, ry0 = ty * step_y;
            for (int p = p_lo + (wp + NW - p_lo % NW) % NW; p < p_hi; p += NW) {   // the planes p of the group with p % NW == wp
                if (p == 0) tma_load_2d(&st.phi[0][0], &maps.phi, rx0 - 4, ry0 - 1, &full_bar[g]);
                else tma_load_2d(&st.pl[p - 1][0][0], (&maps.dxy) + (p - 1), rx0, ry0, &full_bar[g]);
            }
        }
    };
    auto issue = [&](int tile) {
        for (int g = 0; g < NG; g++) issue_group(tile, g);
    };

    pdl_trigger();
    if (tid == 0) {
        for (int g = 0; g < NG; g++) mbar_init(&full_bar[g], 1);
        mbar_fence_init();
    }
    pdl_wait();   // everything above overlaps the tail of the previous kernel of the stream (launch_chain)
    // row-band split: the neighbours' previous pass must be complete (its halo rows are in our input planes, and
    // it no longer reads the planes this pass pushes into) before anything of this pass touches memory.  A neighbour
    // that is merely busy (another process on its GPU, time slicing, a debugger) is an ordinary scheduling delay:
    // the wait is bounded by wall-clock time (20 s), not by a spin count, and running out of it is reported through
    // an error word the host checks after the solve (PF_ECUDA) -- never by a trap, which would poison the context
    // and with it every pooled plan of the process.
    __shared__ int peer_lost;
    if (tid == 0) peer_lost = 0;
    if (peer.wait_flag[0] || peer.wait_flag[1]) {
        if (tid == 0) {
            unsigned long long t_start = 0;
            for (int side = 0; side < 2; side++) {
                if (!peer.wait_flag[side]) continue;
                unsigned int spins = 0;
                while (ld_acquire_sys(peer.wait_flag[side]) < peer.wait_count[side]) {
                    __nanosleep(64);
                    if ((++spins & 0xfffu) == 0) {
                        unsigned long long now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (!t_start) t_start = now;
                        if (now - t_start > 20000000000ull) {
                            peer_lost = 1;
                            if (peer.error_word) atomicExch(peer.error_word, 1u + (unsigned)side);
                            break;
                        }
                    }
                }
            }
            asm volatile("fence.proxy.async.global;" ::: "memory");   // the tile loads below go through the async proxy
        }
    }
    __syncthreads();
    if (peer.wait_flag[0] || peer.wait_flag[1])      // every issuing thread orders the acquire above before its tile loads
        asm volatile("fence.proxy.async.global;" ::: "memory");
    int tile = peer_lost ? ntiles : blockIdx.x;      // a lost neighbour: do nothing, still signal below
    if (tile < ntiles) issue(tile);
    uint32_t parity = 0;
    long long sc_wait = 0, sc_unpack = 0, sc_sweep = 0, sc_wb = 0, sc_t0 = 0, sc_first = 0;
    int sc_n = 0;
    if (PF_SOR_STATS) sc_t0 = clock64();

    for (; tile < ntiles; tile += gridDim.x) {
        const int tx = tile % ntx, ty = tile / ntx + ty0;   // ty0: first tile row of this launch (row-band split)
        const int rx0 = tx * step_x, ry0 = ty * step_y;   // even by construction
        const int xa = rx0 + 2 * lane, ya = ry0 + wp * R;

        T w[R][2], dxy[R][2], iu[R][2], iv[R][2], bu[R][2], bv[R][2], du[R][2], dv[R][2];
        T wl[R], wr[R], wu[2];   // wl / wr: weights towards the neighbouring lanes' columns, zero at the region edge
        // stage -> registers, one plane (compile-time index after unrolling)
        auto unpack = [&](const int p) {
            if (p == 0) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int row = wp * R + r;
                    V2 t = *reinterpret_cast<const V2*>(&st.phi[row + 1][2 * lane + 4]);
                    w[r][0] = t.x * alpha; w[r][1] = t.y * alpha;
                    wl[r] = lane == 0 ? (T)0 : st.phi[row + 1][2 * lane + 3] * alpha;
                    wr[r] = lane == 31 ? (T)0 : w[r][1];
                }
                V2 t = *reinterpret_cast<const V2*>(&st.phi[wp * R][2 * lane + 4]);
                wu[0] = t.x * alpha; wu[1] = t.y * alpha;
                return;
            }
            T (*dst)[2] = p == 1 ? dxy : p == 2 ? iu : p == 3 ? iv : p == 4 ? bu : p == 5 ? bv : p == 6 ? du : dv;
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (p >= 6 && !has_input) { dst[r][0] = dst[r][1] = 0; continue; }
                V2 t = *reinterpret_cast<const V2*>(&st.pl[p - 1][wp * R + r][2 * lane]);
                dst[r][0] = t.x; dst[r][1] = t.y;
            }
        };
        auto publish = [&](int b) {
            *reinterpret_cast<V2*>(&ex[b][0][wp][0][2 * lane]) = V2{du[0][0], du[0][1]};
            *reinterpret_cast<V2*>(&ex[b][1][wp][0][2 * lane]) = V2{dv[0][0], dv[0][1]};
            *reinterpret_cast<V2*>(&ex[b][0][wp][1][2 * lane]) = V2{du[R - 1][0], du[R - 1][1]};
            *reinterpret_cast<V2*>(&ex[b][1][wp][1][2 * lane]) = V2{dv[R - 1][0], dv[R - 1][1]};
        };
        int buf = 0;
        const int next = tile + gridDim.x;
        long long sc_a = 0, sc_b = 0;
        if (PF_SOR_STATS) sc_a = clock64();
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const bool in_flight = has_input || g * PPG < 6;
            if (in_flight && !(PF_SORX & 4)) mbar_wait(&full_bar[g], parity);
            if (PF_SOR_STATS && g == 0) { sc_b = clock64(); sc_wait += sc_b - sc_a; if (!sc_n) sc_first = sc_b - sc_a; }
#pragma unroll
            for (int p = g * PPG; p < (g + 1) * PPG; p++) unpack(p);
            if (g == NG - 1) publish(0);
            __syncthreads();   // group consumed by everyone (last group: exchange rows published)
            if (next < ntiles) issue_group(next, g);   // overlaps the rest of the unpacking and the sweeps below
        }
        parity ^= 1;
        if (PF_SOR_STATS) { sc_a = clock64(); sc_unpack += sc_a - sc_b; }

        // one pixel of row r (compile-time after unrolling), column p = (r + c) & 1
        auto update = [&](const int r, const int c, T up_du, T up_dv, T dn_du, T dn_dv) {
            const int p = (r + c) & 1;
            // horizontal neighbours: one is the thread's own other column, one lives in the next lane.  The lane at
            // the region edge gets its own value back from the shuffle and multiplies it by a zero weight.
            T lw, rw, ldu, ldv, rdu, rdv;
            if (p == 1) {
                lw = w[r][0]; ldu = du[r][0]; ldv = dv[r][0];
                rw = wr[r];
                rdu = __shfl_down_sync(0xffffffffu, du[r][0], 1);
                rdv = __shfl_down_sync(0xffffffffu, dv[r][0], 1);
            } else {
                lw = wl[r];
                ldu = __shfl_up_sync(0xffffffffu, du[r][1], 1);
                ldv = __shfl_up_sync(0xffffffffu, dv[r][1], 1);
                rw = w[r][0];
                rdu = du[r][1]; rdv = dv[r][1];
            }
            T uw, udu, udv, ddu, ddv;
            if (r > 0) { uw = w[r - 1][p]; udu = du[r - 1][p]; udv = dv[r - 1][p]; }
            else       { uw = wu[p];       udu = up_du;        udv = up_dv; }
            if (r < R - 1) { ddu = du[r + 1][p]; ddv = dv[r + 1][p]; }
            else           { ddu = dn_du;        ddv = dn_dv; }
            const T cw = w[r][p];
            T s1 = bu[r][p] + lw * ldu + rw * rdu + uw * udu + cw * ddu;
            T s2 = bv[r][p] + lw * ldv + rw * rdv + uw * udv + cw * ddv;
            s1 -= dxy[r][p] * dv[r][p];
            T nu = one_m * du[r][p] + iu[r][p] * s1;
            s2 -= dxy[r][p] * nu;
            T nv = one_m * dv[r][p] + iv[r][p] * s2;
            du[r][p] = nu;
            dv[r][p] = nv;
        };

        for (int s = 0; s < ((PF_SORX & 1) ? 0 : nsw); s++) {
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int p_top = c & 1, p_bot = (R - 1 + c) & 1;
                T up_du = 0, up_dv = 0, dn_du = 0, dn_dv = 0;
                if (wp > 0) {
                    up_du = ex[buf][0][wp - 1][1][2 * lane + p_top];
                    up_dv = ex[buf][1][wp - 1][1][2 * lane + p_top];
                }
                if (wp < NW - 1) {
                    dn_du = ex[buf][0][wp + 1][0][2 * lane + p_bot];
                    dn_dv = ex[buf][1][wp + 1][0][2 * lane + p_bot];
                }
#pragma unroll
                for (int r = 0; r < R; r++) update(r, c, up_du, up_dv, dn_du, dn_dv);
                buf ^= 1;
                publish(buf);      // readers of the other buffer are at most one barrier behind
                __syncthreads();   // (pairwise 64-thread named barriers and a split-phase mbarrier both measured slower)
            }
        }

        if (PF_SOR_STATS) { sc_b = clock64(); sc_sweep += sc_b - sc_a; }
        const int ox_lo = tx > 0 ? rx0 + HL : 0;
        const int ox_hi = (rx0 + kSorRegionW >= W) ? W : rx0 + kSorRegionW - HL;
        const int oy_lo = ty > 0 ? ry0 + HL : 0;
        const int oy_hi = (ry0 + RH >= H) ? H : ry0 + RH - HL;
        // write back the window that is still exact.  Column validity does not depend on the row and the
        // valid rows of a warp are one interval, so the common case is R predicated 8-byte store pairs
        // off two running pointers (a column pair is split only in the last column of an odd-width level).
        {
            const bool v0 = xa >= ox_lo && xa < ox_hi, v1 = xa + 1 >= ox_lo && xa + 1 < ox_hi;
            const int r_lo = oy_lo - ya;
            const unsigned r_n = (unsigned)max(oy_hi - oy_lo, 0);
            char* const pu = reinterpret_cast<char*>(du_out) + ((size_t)ya * P + xa) * sizeof(T);
            const ptrdiff_t to_dv = reinterpret_cast<char*>(dv_out) - reinterpret_cast<char*>(du_out);
            const unsigned pitch_b = (unsigned)P * (unsigned)sizeof(T);
            if (PF_SORX & 2) {
                // ablation: keep the values alive without storing them
                T acc = 0;
#pragma unroll
                for (int r = 0; r < R; r++) acc += du[r][0] + du[r][1] + dv[r][0] + dv[r][1];
                if (acc == (T)123456789) du_out[0] = acc;
            } else if (v0 && v1) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    char* q = pu + (size_t)((unsigned)r * pitch_b);
                    if ((unsigned)(r - r_lo) < r_n) {
                        *reinterpret_cast<V2*>(q) = V2{du[r][0], du[r][1]};
                        *reinterpret_cast<V2*>(q + to_dv) = V2{dv[r][0], dv[r][1]};
                    }
                }
            } else if (v0 || v1) {
                const int k = v1 ? 1 : 0;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    T* q = reinterpret_cast<T*>(pu + (size_t)((unsigned)r * pitch_b)) + k;
                    if ((unsigned)(r - r_lo) < r_n) {
                        *q = v1 ? du[r][1] : du[r][0];
                        *reinterpret_cast<T*>(reinterpret_cast<char*>(q) + to_dv) = v1 ? dv[r][1] : dv[r][0];
                    }
                }
            }
        }
        // row-band split over several GPUs (kernel-uniform branch, not taken on a single GPU): the same
        // values go straight into the planes of the neighbour(s) whose next pass reads the row; a row can
        // be wanted by both neighbours when bands are thin
        if (peer.up_hi > peer.up_lo || peer.dn_hi > peer.dn_lo) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int y = ya + r;
                if (y < oy_lo || y >= oy_hi) continue;
                const bool v0 = xa >= ox_lo && xa < ox_hi, v1 = xa + 1 >= ox_lo && xa + 1 < ox_hi;
                const size_t o = (size_t)y * P + xa;
#pragma unroll
                for (int side = 0; side < 2; side++) {
                    const bool want = side ? (y >= peer.dn_lo && y < peer.dn_hi) : (y >= peer.up_lo && y < peer.up_hi);
                    if (!want) continue;
                    T* pu = side ? peer.dn_du : peer.up_du;
                    T* pv = side ? peer.dn_dv : peer.up_dv;
                    if (v0 && v1) {
                        *reinterpret_cast<V2*>(pu + o) = V2{du[r][0], du[r][1]};
                        *reinterpret_cast<V2*>(pv + o) = V2{dv[r][0], dv[r][1]};
                    } else if (v0) {
                        pu[o] = du[r][0]; pv[o] = dv[r][0];
                    } else if (v1) {
                        pu[o + 1] = du[r][1]; pv[o + 1] = dv[r][1];
                    }
                }
            }
        }
        if (PF_SOR_STATS) { sc_wb += clock64() - sc_b; sc_n++; }
    }
    if (PF_SOR_STATS && tid == 0 && (blockIdx.x == 0 || blockIdx.x == 100) && sc_n > 0)
        printf("SORSTAT cta %d %dx%d nsw %d tiles %d | total %lld cycles: wait %lld (first %lld) unpack %lld sweeps %lld write-back+rest %lld\n", blockIdx.x,
               W, H, nsw, sc_n, clock64() - sc_t0, sc_wait, sc_first, sc_unpack, sc_sweep, sc_wb);
    // row-band split: this CTA is done reading and writing -- tell the neighbours (one arrival per CTA)
    if (peer.signal_flag[0] || peer.signal_flag[1]) {
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            if (peer.signal_flag[0]) red_release_sys_add(peer.signal_flag[0], 1u);
            if (peer.signal_flag[1]) red_release_sys_add(peer.signal_flag[1], 1u);
        }
    }
}

// Tiling of an image dimension of size n by regions of size `region` whose exact window shrinks by
// `halo` on every side that is not an image edge.  Region origins are step*k.
struct SorTiling {
    int ntiles, step;
};
inline SorTiling sor_tiling(int n, int region, int halo) {
    if (n <= region) return {1, region};
    int step = region - 2 * halo;
    if (step < 2) return {0, 0};
    return {ceil_div(n - region, step) + 1, step};
}

}  // namespace pf
