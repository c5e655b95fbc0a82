// k_fused_tma: getDxs + phi + linear-system assembly of one (inner) fixed-point iteration in one
// kernel, with the per-channel input tiles staged by TMA and double buffered.
//
// Same arithmetic, same term order and same replicate-border semantics as k_fused_assemble
// (kernels.cuh) -- the FP64 instantiation stays bit-identical to the reference -- but organised for
// throughput:
//   * the warped-feature tile (72 x (TY+8), halo 4) and the smoothed-Im1 tile (72 x (TY+4), halo 2)
//     of channel c+2 are fetched by cp.async.bulk.tensor (3-D tensor maps x, y, channel) into the
//     stage that channel c has just released, so global latency hides behind the shared-memory
//     stencil stages of channels c and c+1;
//   * TMA zero-fills outside the image; tiles that touch the image border run a short fix-up that
//     copies the nearest in-image entry over every out-of-image entry after each stage, which IS the
//     reference's clamp rule (S/ImageProcessing.h:259-279, 350-369).  All stencil stages are
//     therefore clamp-free straight-line code for every tile;
//   * thread t owns column t%64 and row segment t/64 in every stage, vertical 5-tap windows slide
//     through registers, all loop trip counts are compile-time constants.
#pragma once
#include "kernels.cuh"
#include "tma.cuh"
#include <type_traits>

namespace pf {

struct FusedMaps {
    CUtensorMap wf, s1;   // (W, H, C) planar tensors: warped Im2 features, smoothed Im1 features
    CUtensorMap u, v;     // (W, H) flow planes, box 72 x (TY+2)
};

template <typename T, int TY>
struct FusedSmem {
    static constexpr int TX = 64, RW = 72, RH = TY + 8, SH = TY + 4, HW = 68, BW = 68;
    alignas(128) T raw[2][RH][RW];   // TMA destinations
    alignas(128) T s1[2][SH][RW];
    alignas(16) T hs[RH][HW];
    T bl[SH][BW];
    T dt[TY][TX];
};

#ifndef PF_FUSED_MINB
#define PF_FUSED_MINB 4   // 4 CTAs per SM (<= 64 registers, 32 B of spill) measured 13 % faster than 3 CTAs at 80 registers
#endif
template <typename T, int TY, int SEG>
__global__ void __launch_bounds__(64 * SEG, sizeof(T) == 4 ? PF_FUSED_MINB : 1)
k_fused_tma(const __grid_constant__ FusedMaps maps, FusedArgs<T> a) {
    typedef FusedSmem<T, TY> Smem;
    constexpr int TX = 64, NT = TX * SEG, NWARP = NT / 32;
    constexpr int RW = Smem::RW, RHt = Smem::RH, SH = Smem::SH, HW = Smem::HW, BW = Smem::BW;
    constexpr int PPT = TY / SEG;                    // centre rows per thread
    constexpr int BROWS = (SH + SEG - 1) / SEG;      // blend rows per thread
    constexpr int UW = TX + 2, UH = TY + 2, PW = TX + 1, PH = TY + 1;
    static_assert(TY % SEG == 0, "tile height must split into SEG segments");
    static_assert(4 * (TY + 4) <= 64 * SEG, "one thread per entry of the four halo columns of the blend tile");
    static_assert(2 * UW * UH <= RHt * HW + SH * BW, "u/v tiles alias hs + bl");

    // the alignment is declared on the dynamic segment itself (TMA destinations need 128 B); going
    // through an integer round-up would make the compiler lose the shared address space and emit
    // generic LD/ST instead of LDS/STS for every tile access
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_dyn);
    __shared__ __align__(8) uint64_t full_bar[2], uv_bar;
    __shared__ T sm_phi[PW * PH];

    const int W = a.w, H = a.h;
    const int x0 = blockIdx.x * TX, y0 = (blockIdx.y + a.ty0) * TY;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col = tid & (TX - 1), seg = tid / TX;
    const int C = a.wf.c;
    const bool border = x0 < 4 || y0 < 4 || x0 + TX + 4 > W || y0 + TY + 4 > H;
    constexpr uint32_t kStageBytes = (uint32_t)(sizeof(T) * (RHt * RW + SH * RW));

    auto issue = [&](int c) {
        const int s = c & 1;
        mbar_expect_tx(&full_bar[s], kStageBytes);
        tma_load_3d(&sm.raw[s][0][0], &maps.wf, x0 - 4, y0 - 4, c, &full_bar[s]);
        tma_load_3d(&sm.s1[s][0][0], &maps.s1, x0 - 4, y0 - 2, c, &full_bar[s]);
    };
    // The u / v tiles of the Laplacian stage (halo 1; fetched as 72 x (TY+2) boxes at (x0-4, y0-1) for
    // TMA's 16-byte coordinate rule) go into the input stage that the last-but-one channel releases,
    // so their latency hides behind the last channel.  Only when no increment has to be added (du ==
    // nullptr, i.e. every first inner iteration); otherwise the tiles are built with plain loads.
    const bool uv_staged = a.du == nullptr;
    const int su = C & 1;
    static_assert(UH * RW <= RHt * RW && UH * RW <= SH * RW, "u / v boxes fit the released input stage");
    auto issue_uv = [&]() {
        mbar_expect_tx(&uv_bar, (uint32_t)(2 * sizeof(T) * UH * RW));
        tma_load_2d(&sm.raw[su][0][0], &maps.u, x0 - 4, y0 - 1, &uv_bar);
        tma_load_2d(&sm.s1[su][0][0], &maps.v, x0 - 4, y0 - 1, &uv_bar);
    };
    pdl_trigger();
    if (tid == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        mbar_init(&uv_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    pdl_wait();   // nothing above touches global memory
    if (tid == 0) {
        issue(0);
        if (C > 1) issue(1);
        else if (uv_staged) issue_uv();   // a single channel never touches stage 1
    }

    // replicate-border fix-up of a tile whose entry (r, c) sits at image coordinate (oy + r, ox + c)
    auto fixup = [&](T* tile, int rows, int cols, int stride, int ox, int oy) {
        const int c_lo = max(0, -ox), c_hi = min(cols - 1, W - 1 - ox);
        const int r_lo = max(0, -oy), r_hi = min(rows - 1, H - 1 - oy);
        for (int r = warp; r < rows; r += NWARP) {
            const int rr = min(max(r, r_lo), r_hi);
            for (int c = lane; c < cols; c += 32) {
                const int cc = min(max(c, c_lo), c_hi);
                if (rr != r || cc != c) tile[r * stride + c] = tile[rr * stride + cc];
            }
        }
    };

    const int PX = x0 + col;
    T sxy[PPT], sx2[PPT], sy2[PPT], stx[PPT], sty[PPT];
#pragma unroll
    for (int k = 0; k < PPT; k++) sxy[k] = sx2[k] = sy2[k] = stx[k] = sty[k] = 0;
    const T g0 = a.g5.v[0], g1 = a.g5.v[1], g2 = a.g5.v[2], g3 = a.g5.v[3], g4 = a.g5.v[4];
    const T d0 = a.d5.v[0], d1 = a.d5.v[1], d3 = a.d5.v[3], d4 = a.d5.v[4];   // centre tap is 0

    for (int c = 0; c < C; c++) {
        const int s = c & 1;
        const bool active = !(a.lap && a.lap[c] < 1e-20);   // S/OpticalFlow.cpp:399-400
        mbar_wait(&full_bar[s], (uint32_t)((c >> 1) & 1));
        T* raw = &sm.raw[s][0][0];
        T* s1t = &sm.s1[s][0][0];
        if (border) {
            fixup(raw, RHt, RW, RW, x0 - 4, y0 - 4);
            fixup(s1t, SH, RW, RW, x0 - 4, y0 - 2);
            __syncthreads();
        }
        // ---- horizontal smoothing: hs(ry, hx) for image columns x0-2+hx; hx sits at raw column hx+2, so
        //      outputs 4q..4q+3 need raw columns 4q..4q+7: two 16-byte loads feed four outputs --------
        {
            constexpr int NQ = HW / 4;
            static_assert(HW % 4 == 0 && RW % 4 == 0 && 4 * (NQ - 1) + 8 <= RW, "quads stay inside a raw row, rows stay 16-byte aligned");
            for (int task = tid; task < RHt * NQ; task += NT) {
                const int ry = task / NQ, q = task - ry * NQ;
                T lo[4], hi[4], out[4];
                ld4(raw + ry * RW + 4 * q, lo);
                ld4(raw + ry * RW + 4 * q + 4, hi);
                const T win[8] = {lo[0], lo[1], lo[2], lo[3], hi[0], hi[1], hi[2], hi[3]};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    T acc = 0;
                    acc += win[j] * g0; acc += win[j + 1] * g1; acc += win[j + 2] * g2; acc += win[j + 3] * g3; acc += win[j + 4] * g4;
                    out[j] = acc;
                }
                st4(&sm.hs[ry][4 * q], out);
            }
        }
        __syncthreads();
        if (border) {
            fixup(&sm.hs[0][0], RHt, HW, HW, x0 - 2, y0 - 4);
            __syncthreads();
        }
        // ---- vertical smoothing (sliding window), blend with smoothed Im1, temporal difference ----
        auto vstage = [&](int bx, int sg) {
            const int by0 = sg * BROWS;
            if (by0 >= SH) return;
            const T* hcol = &sm.hs[0][0] + bx;
            T w0 = hcol[(by0 + 0) * HW], w1 = hcol[(by0 + 1) * HW], w2 = hcol[(by0 + 2) * HW], w3 = hcol[(by0 + 3) * HW];
#pragma unroll
            for (int j = 0; j < BROWS; j++) {
                const int by = by0 + j;
                if (by >= SH) break;
                T w4 = hcol[(by + 4) * HW];
                T acc = 0;
                acc += w0 * g0; acc += w1 * g1; acc += w2 * g2; acc += w3 * g3; acc += w4 * g4;
                w0 = w1; w1 = w2; w2 = w3; w3 = w4;
                const T s1v = s1t[by * RW + bx + 2];
                const T t = s1v * (T)0.4;
                sm.bl[by][bx] = t + acc * (T)0.6;
                const int ccx = bx - 2, ccy = by - 2;
                if (ccx >= 0 && ccx < TX && ccy >= 0 && ccy < TY) sm.dt[ccy][ccx] = acc - s1v;
            }
        };
        vstage(col, seg);
        // the four halo columns right of the tile: one output per thread (the first 4 * SH threads, spread over several
        // warps) instead of a second sliding pass by half a warp, which every other warp of the CTA waited for
        if (tid < 4 * SH) {
            const int bx = TX + (tid & 3), by = tid >> 2;
            const T* hcol = &sm.hs[0][0] + bx;
            T acc = 0;
            acc += hcol[(by + 0) * HW] * g0; acc += hcol[(by + 1) * HW] * g1; acc += hcol[(by + 2) * HW] * g2;
            acc += hcol[(by + 3) * HW] * g3; acc += hcol[(by + 4) * HW] * g4;
            const T s1v = s1t[by * RW + bx + 2];
            const T t = s1v * (T)0.4;
            sm.bl[by][bx] = t + acc * (T)0.6;
            const int ccx = bx - 2, ccy = by - 2;
            if (ccx < TX && ccy >= 0 && ccy < TY) sm.dt[ccy][ccx] = acc - s1v;
        }
        __syncthreads();
        // both input stages of this channel are free: prefetch channel c+2 into them
        if (tid == 0) {
            if (c + 2 < C) issue(c + 2);
            else if (c + 2 == C && uv_staged) issue_uv();
        }
        if (border) {
            fixup(&sm.bl[0][0], SH, BW, BW, x0 - 2, y0 - 2);
            __syncthreads();
        }
        // ---- derivatives of the blend at the centre pixels, psi-weighted products ----------------
        {
            const int cy0 = seg * PPT;
            const T* bcol = &sm.bl[0][0] + (col + 2);
            T v0 = bcol[(cy0 + 0) * BW], v1 = bcol[(cy0 + 1) * BW], v2 = bcol[(cy0 + 2) * BW], v3 = bcol[(cy0 + 3) * BW];
#pragma unroll
            for (int k = 0; k < PPT; k++) {
                const int cy = cy0 + k;
                T v4 = bcol[(cy + 4) * BW];
                const T* brow = &sm.bl[cy + 2][col];
                T ix = 0, iy = 0;
                ix += brow[0] * d0; ix += brow[1] * d1; ix += brow[3] * d3; ix += brow[4] * d4;
                iy += v0 * d0; iy += v1 * d1; iy += v3 * d3; iy += v4 * d4;
                v0 = v1; v1 = v2; v2 = v3; v3 = v4;
                const T it = sm.dt[cy][col];
                T psi = 0;
                if (active) {
                    T t = it;
                    if (a.du) {
                        const int PY = min(y0 + cy, H - 1), PXc = min(PX, W - 1);
                        t = it + ix * a.du[(size_t)PY * a.pitch + PXc] + iy * a.dv[(size_t)PY * a.pitch + PXc];
                    }
                    psi = psi_of(t * t, a.eps);
                }
                const T px = psi * ix, py = psi * iy;
                sxy[k] += px * iy; sx2[k] += px * ix; sy2[k] += py * iy; stx[k] += px * it; sty[k] += py * it;
            }
        }
        // no barrier here: the next channel's first stage writes hs, which nobody reads any more,
        // and its barrier separates this stage's bl/dt reads from the next vertical stage's writes
    }
    __syncthreads();

    // ---- u+du, v+dv tiles (halo 1) -> phi on the tile plus its left/up halo; tiles reloaded with
    //      plain u, v when du is present: the Laplacian acts on u (S/OpticalFlow.cpp:437-438) ------
    T* tphi = sm_phi;
    auto tail = [&](T* tu, T* tv, auto stride_c) {
    constexpr int US = decltype(stride_c)::value;      // row stride of the u / v tiles; entry (uy, ux) is pixel (y0-1+uy, x0-1+ux)
    auto load_uv = [&](bool with_increment) {
        for (int uy = warp; uy < UH; uy += NWARP) {
            const size_t ro = (size_t)clampi(y0 - 1 + uy, H) * a.pitch;
            for (int ux = lane; ux < UW; ux += 32) {
                const size_t o = ro + clampi(x0 - 1 + ux, W);
                T uv = a.u[o], vv = a.v[o];
                if (with_increment) { uv += a.du[o]; vv += a.dv[o]; }
                tu[uy * US + ux] = uv;
                tv[uy * US + ux] = vv;
            }
        }
    };
    if (uv_staged) {
        mbar_wait(&uv_bar, 0);
        if (border) {                                   // replicate clamp over TMA's zero fill
            fixup(tu - 3, UH, RW, RW, x0 - 4, y0 - 1);
            fixup(tv - 3, UH, RW, RW, x0 - 4, y0 - 1);
            __syncthreads();
        }
    } else {
        load_uv(true);
        __syncthreads();
    }
    for (int py = warp; py < PH; py += NWARP) {
        const int Y = y0 - 1 + py;
        for (int px = lane; px < PW; px += 32) {
            const int X = x0 - 1 + px;
            T val = 0;
            if (X >= 0 && X < W && Y >= 0 && Y < H) {
                const int ui = py * US + px;
                const T u0 = tu[ui], v0 = tv[ui];
                T ux = 0, uy = 0, vx = 0, vy = 0;
                if (X < W - 1) { ux = tu[ui + 1] - u0; vx = tv[ui + 1] - v0; }
                if (Y < H - 1) { uy = tu[ui + US] - u0; vy = tv[ui + US] - v0; }
                const T t = ux * ux + uy * uy + vx * vx + vy * vy;
                val = phi_of(t, a.eps);
            }
            tphi[py * PW + px] = val;
        }
    }
    __syncthreads();
    if (!uv_staged) {
        load_uv(false);
        __syncthreads();
    }
    // ---- Laplacian (fork quirk F3), right-hand sides, inverse diagonals ---------------------------
    if (PX >= W) return;
    const T inv_c = (T)1 / (T)C;
#pragma unroll
    for (int k = 0; k < PPT; k++) {
        const int cy = seg * PPT + k, Y = y0 + cy, X = PX;
        if (Y >= H) continue;
        const int pi = (cy + 1) * PW + (col + 1), ui = (cy + 1) * US + (col + 1);
        const T ph = tphi[pi];
        const bool xr = X < W - 1, xl = X > 0, yd = Y < H - 1, yu = Y > 0;
        T lu = 0, lv = 0, cf = 0;
        if (xr) {
            lu -= (tu[ui + 1] - tu[ui]) * ph;
            lv -= (tv[ui + 1] - tv[ui]) * ph;
            if (xl) {
                lu += (tu[ui] - tu[ui - 1]) * tphi[pi - 1];
                lv += (tv[ui] - tv[ui - 1]) * tphi[pi - 1];
            }
        }
        if (yd) {
            lu -= (tu[ui + US] - tu[ui]) * ph;
            lv -= (tv[ui + US] - tv[ui]) * ph;
            if (yu) {
                lu += (tu[ui] - tu[ui - US]) * tphi[pi - PW];
                lv += (tv[ui] - tv[ui - US]) * tphi[pi - PW];
            }
        }
        if (xl) cf += tphi[pi - 1];
        if (xr) cf += ph;
        if (yu) cf += tphi[pi - PW];
        if (yd) cf += ph;
        cf *= a.alpha;
        T a_xy = sxy[k], a_x2 = sx2[k], a_y2 = sy2[k], a_tx = stx[k], a_ty = sty[k];
        if (C > 1) {
            a_xy = mean_of(a_xy, C, inv_c); a_x2 = mean_of(a_x2, C, inv_c); a_y2 = mean_of(a_y2, C, inv_c);
            a_tx = mean_of(a_tx, C, inv_c); a_ty = mean_of(a_ty, C, inv_c);
        }
        const T reg = a.alpha * (T)0.05;
        const size_t o = (size_t)Y * a.pitch + X;
        a.phi[o] = ph;
        a.dxy[o] = a_xy;
        a.iu[o] = ratio_of(a.omega, a_x2 + reg + cf);
        a.iv[o] = ratio_of(a.omega, a_y2 + reg + cf);
        a.bu[o] = -a_tx - a.alpha * lu;
        a.bv[o] = -a_ty - a.alpha * lv;
    }
    };
    if (uv_staged) tail(&sm.raw[su][0][0] + 3, &sm.s1[su][0][0] + 3, std::integral_constant<int, RW>{});
    else tail(&sm.hs[0][0], &sm.hs[0][0] + UW * UH, std::integral_constant<int, UW>{});
}

}  // namespace pf
