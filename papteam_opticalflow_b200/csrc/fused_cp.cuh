// k_fused_cp: the fused getDxs + assembly kernel of fused_tma.cuh with the feature channels processed IN PARALLEL.
//
// On the coarse pyramid levels one pair alone leaves the GPU almost empty, and an assembly launch is as long as the chain
// of ONE tile: five channels, each three shared-memory stencil stages separated by barriers, one after the other
// (~14 us per launch from the 455-px level down, 165 launches per 1920-wide pair).  Here a CTA of C x 64*SEG threads
// gives every channel its own thread group and its own stage buffers: all 2 C + 2 input tiles are requested at once (one
// bulk tensor copy per warp), the groups walk h-smoothing -> v-smoothing + blend -> derivatives in lock step, leave
// {psi Ix, psi Iy, Ix, Iy, It} of their channel in shared memory, and group 0 adds them up in channel order with the
// expressions of k_fused_tma (same contraction, same rounding) before the common tail (phi, fork-quirk Laplacian,
// right-hand sides, inverse diagonals).  Same arithmetic per pixel as k_fused_tma: the results are bit-identical
// (tests: every latency-tuned solve against the throughput-tuned one).  Used by latency-tuned plans on the levels whose
// 64x16 tiling leaves most SMs idle (Plan::small_tiles); more halo work per pixel, which is why nothing else uses it.
//
// WARP = true folds the flow update + bilinear warp (k_update_warp, S/OpticalFlow.cpp:513-516 -> S/ImageProcessing.h:483-503)
// into the head of the kernel: all threads compute the sampling geometry of the 72 x (TY+8) halo tile once (flow u + du of
// the PREVIOUS solve, k_update_warp's arithmetic) into shared memory, every channel group gathers its channel's warped tile
// from it (entries outside the image take the nearest in-image pixel's value = the replicate rule), group 0 stores the
// updated flow of the tile's own pixels to a second pair of planes.  One launch and one dependent chain less per outer
// iteration (~6 us of ~25 on the coarse levels); the warped features are never written to memory.
#pragma once
#include "fused_tma.cuh"

namespace pf {

template <typename T, int TY>
struct alignas(128) FusedCpChannel {
    static constexpr int TX = 64, RW = 72, RH = TY + 8, SH = TY + 4, HW = 68, BW = 68;
    alignas(128) T raw[RH][RW];   // TMA destination: warped features, halo 4
    alignas(128) T s1[SH][RW];    // TMA destination: smoothed Im1 features, halo 2
    alignas(16) T hs[RH][HW];
    T bl[SH][BW];
    T dt[TY][TX];
    T out[5][TY][TX];             // psi Ix, psi Iy, Ix, Iy, It of this channel at the centre pixels
};

struct alignas(16) WarpGeom {     // sampling geometry of one entry of the halo tile
    int o00, o01, o10, o11;       // offsets inside a feature plane of Im2; o00 < 0: outside the image, -(o00 + 1) = Im1 offset
    float w00, w01, w10, w11;     // (double in the FP64 instantiation is not needed: the folded form is FP32 only)
};

template <typename T, int TY>
inline size_t fused_cp_smem_bytes(int channels, bool warp = false) {
    typedef FusedCpChannel<T, TY> Ch;
    return sizeof(Ch) * (size_t)channels + 2 * round_up(sizeof(T) * (TY + 2) * Ch::RW, 128) + round_up(sizeof(T) * 65 * (TY + 1), 128) +
           (warp ? sizeof(WarpGeom) * Ch::RH * Ch::RW : 0) + 128;
}

template <typename T, int TY, int SEG, bool WARP = false>
__global__ void __launch_bounds__(1024, 1)
k_fused_cp(const __grid_constant__ FusedMaps maps, FusedArgs<T> a) {
    typedef FusedCpChannel<T, TY> Ch;
    constexpr int TX = 64, NT = TX * SEG, NWG = NT / 32;      // threads / warps of one channel group
    constexpr int RW = Ch::RW, RHt = Ch::RH, SH = Ch::SH, HW = Ch::HW, BW = Ch::BW;
    constexpr int PPT = TY / SEG;
    constexpr int BROWS = (SH + SEG - 1) / SEG;
    constexpr int UW = TX + 2, UH = TY + 2, PW = TX + 1, PH = TY + 1;
    static_assert(TY % SEG == 0, "tile height must split into SEG segments");
    static_assert(4 * SH <= NT, "one thread per entry of the four halo columns of the blend tile");
    static_assert(2 * UW * UH <= RHt * HW + SH * BW, "u/v tiles alias hs + bl of channel 0");
    constexpr size_t kUvBytes = (sizeof(T) * UH * RW + 127) / 128 * 128;

    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar;
    const int C = a.wf.c;                                     // blockDim.x == NT * C
    Ch* const chs = reinterpret_cast<Ch*>(smem_dyn);
    T* const tu_s = reinterpret_cast<T*>(smem_dyn + sizeof(Ch) * (size_t)C);          // [UH][RW], TMA destination
    T* const tv_s = reinterpret_cast<T*>(smem_dyn + sizeof(Ch) * (size_t)C + kUvBytes);
    T* const tphi = reinterpret_cast<T*>(smem_dyn + sizeof(Ch) * (size_t)C + 2 * kUvBytes);   // [PH][PW]
    constexpr size_t kPhiBytes = (sizeof(T) * PW * PH + 127) / 128 * 128;
    WarpGeom* const geom = reinterpret_cast<WarpGeom*>(smem_dyn + sizeof(Ch) * (size_t)C + 2 * kUvBytes + kPhiBytes);   // WARP only

    const int W = a.w, H = a.h;
    const int x0 = blockIdx.x * TX, y0 = (blockIdx.y + a.ty0) * TY;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int g = tid / NT, lt = tid - g * NT;                 // channel group, thread within the group
    const int gw = lt >> 5;                                    // warp within the group
    const int col = lt & (TX - 1), seg = lt / TX;
    const bool border = x0 < 4 || y0 < 4 || x0 + TX + 4 > W || y0 + TY + 4 > H;
    const bool uv_staged = !WARP && a.du == nullptr;   // WARP: the tail's flow tiles are u + wdu, built with plain loads
    Ch& ch = chs[g];

    if (tid == 0) {
        mbar_init(&full_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    // all input tiles at once: warp k issues copy k (2 per channel, then u and v); thread 0 arms the byte count
    if (tid == 0) {
        uint32_t bytes = (uint32_t)(sizeof(T) * ((WARP ? 0 : RHt * RW) + SH * RW)) * (uint32_t)C;
        if (uv_staged) bytes += (uint32_t)(2 * sizeof(T) * UH * RW);
        mbar_expect_tx(&full_bar, bytes);
    }
    if (lane == 0) {
        for (int k = warp; k < 2 * C + 2; k += nwarp) {
            if (k < 2 * C) {
                const int c = k >> 1;
                if (k & 1) tma_load_3d(&chs[c].s1[0][0], &maps.s1, x0 - 4, y0 - 2, c, &full_bar);
                else if (!WARP) tma_load_3d(&chs[c].raw[0][0], &maps.wf, x0 - 4, y0 - 4, c, &full_bar);
            } else if (uv_staged) {
                if (k == 2 * C) tma_load_2d(tu_s, &maps.u, x0 - 4, y0 - 1, &full_bar);
                else tma_load_2d(tv_s, &maps.v, x0 - 4, y0 - 1, &full_bar);
            }
        }
    }

    // replicate-border fix-up of a tile whose entry (r, c) sits at image coordinate (oy + r, ox + c), by the warps wi, wi + nw, ...
    auto fixup = [&](T* tile, int rows, int cols, int stride, int ox, int oy, int wi, int nw) {
        const int c_lo = max(0, -ox), c_hi = min(cols - 1, W - 1 - ox);
        const int r_lo = max(0, -oy), r_hi = min(rows - 1, H - 1 - oy);
        for (int r = wi; r < rows; r += nw) {
            const int rr = min(max(r, r_lo), r_hi);
            for (int c = lane; c < cols; c += 32) {
                const int cc = min(max(c, c_lo), c_hi);
                if (rr != r || cc != c) tile[r * stride + c] = tile[rr * stride + cc];
            }
        }
    };

    const int PX = x0 + col;
    const T g0 = a.g5.v[0], g1 = a.g5.v[1], g2 = a.g5.v[2], g3 = a.g5.v[3], g4 = a.g5.v[4];
    const T d0 = a.d5.v[0], d1 = a.d5.v[1], d3 = a.d5.v[3], d4 = a.d5.v[4];   // centre tap is 0
    const bool active = !(a.lap && a.lap[g] < 1e-20);   // S/OpticalFlow.cpp:399-400

    T* raw = &ch.raw[0][0];
    T* s1t = &ch.s1[0][0];
    if constexpr (WARP) {
        // ---- flow update + sampling geometry of every entry of the halo tile, once for all channels (k_update_warp's
        //      arithmetic); entries outside the image take their nearest in-image pixel (the replicate rule) ----------------
        constexpr int NE = RHt * RW;
        // (two entries per thread and trip, their four flow loads issued together: one memory round trip per trip)
        for (int e0 = tid; e0 < NE; e0 += 2 * blockDim.x) {
            int ee[2]; size_t of[2]; int Yq[2], Xq[2]; bool own[2]; T uu[2], vv[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                ee[q] = e0 + q * blockDim.x;
                const int e = min(ee[q], NE - 1);
                const int ry = e / RW, rx = e - ry * RW;
                const int Yr = y0 - 4 + ry, Xr = x0 - 4 + rx;
                Yq[q] = clampi(Yr, H); Xq[q] = clampi(Xr, W);
                own[q] = Yr == Yq[q] && Xr == Xq[q] && ry >= 4 && ry < 4 + TY && rx >= 4 && rx < 4 + TX;   // the tile's own pixels
                of[q] = (size_t)Yq[q] * a.pitch + Xq[q];
                uu[q] = a.u[of[q]]; vv[q] = a.v[of[q]];
            }
            if (a.wdu) {
#pragma unroll
                for (int q = 0; q < 2; q++) { uu[q] += a.wdu[of[q]]; vv[q] += a.wdv[of[q]]; }
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {
                if (ee[q] >= NE) continue;
                if (a.wdu && own[q]) {
                    a.uo[of[q]] = uu[q];
                    a.vo[of[q]] = vv[q];
                }
                {   // the tail's flow tiles (halo 1) are rows 3 .. 3 + UH of this tile: same origin and width as the u / v boxes
                    const int ry = ee[q] / RW, rx = ee[q] - ry * RW;
                    if (ry >= 3 && ry < 3 + UH) { tu_s[(ry - 3) * RW + rx] = uu[q]; tv_s[(ry - 3) * RW + rx] = vv[q]; }
                }
                const int X = Xq[q], Y = Yq[q];
                WarpGeom gm;
                T fx, fy;
                const SamplePos px = sample_pos(X, uu[q], W, fx), py = sample_pos(Y, vv[q], H, fy);
                if (px.out || py.out) {
                    gm.o00 = -(Y * a.f1.pitch + X) - 1;
                    gm.o01 = gm.o10 = gm.o11 = 0;
                    gm.w00 = gm.w01 = gm.w10 = gm.w11 = 0;
                } else {
                    const int xa = clampi(px.i, W), xb = clampi(px.i + 1, W), ya = clampi(py.i, H), yb = clampi(py.i + 1, H);
                    const T ax0 = fabs((T)1 - fx), ax1 = fabs((T)0 - fx), ay0 = fabs((T)1 - fy), ay1 = fabs((T)0 - fy);
                    gm.w00 = ax0 * ay0; gm.w01 = ax0 * ay1; gm.w10 = ax1 * ay0; gm.w11 = ax1 * ay1;
                    gm.o00 = ya * a.f2.pitch + xa; gm.o01 = yb * a.f2.pitch + xa;
                    gm.o10 = ya * a.f2.pitch + xb; gm.o11 = yb * a.f2.pitch + xb;
                }
                // (two 16-byte stores: a 32-byte struct store would be split into conflicting 4-byte ones)
                reinterpret_cast<int4*>(&geom[ee[q]])[0] = make_int4(gm.o00, gm.o01, gm.o10, gm.o11);
                reinterpret_cast<float4*>(&geom[ee[q]])[1] = make_float4(gm.w00, gm.w01, gm.w10, gm.w11);
            }
        }
        __syncthreads();
        // ---- every group gathers its channel's warped tile ---------------------------------------------------------------
        {
            const T* const p2 = a.f2.ch(g);
            const T* const p1 = a.f1.ch(g);
            constexpr int EPT = (NE + NT - 1) / NT;
            T v00[EPT], v01[EPT], v10[EPT], v11[EPT];
#pragma unroll
            for (int i = 0; i < EPT; i++) {
                const int e = lt + i * NT;
                v00[i] = v01[i] = v10[i] = v11[i] = 0;
                if (e < NE) {
                    const int4 o = reinterpret_cast<const int4*>(&geom[e])[0];
                    if (o.x >= 0) { v00[i] = p2[o.x]; v01[i] = p2[o.y]; v10[i] = p2[o.z]; v11[i] = p2[o.w]; }
                    else v00[i] = p1[-(o.x + 1)];
                }
            }
#pragma unroll
            for (int i = 0; i < EPT; i++) {
                const int e = lt + i * NT;
                if (e < NE) {
                    const float4 wq = reinterpret_cast<const float4*>(&geom[e])[1];
                    T acc;
                    if (geom[e].o00 >= 0) {
                        acc = 0;
                        acc += v00[i] * (T)wq.x;
                        acc += v01[i] * (T)wq.y;
                        acc += v10[i] * (T)wq.z;
                        acc += v11[i] * (T)wq.w;
                    } else {
                        acc = v00[i];          // Im1 fallback outside the image
                    }
                    raw[e] = acc;
                }
            }
        }
        mbar_wait(&full_bar, 0);
        if (border) fixup(s1t, SH, RW, RW, x0 - 4, y0 - 2, gw, NWG);
        __syncthreads();
    } else {
        mbar_wait(&full_bar, 0);
        if (border) {
            fixup(raw, RHt, RW, RW, x0 - 4, y0 - 4, gw, NWG);
            fixup(s1t, SH, RW, RW, x0 - 4, y0 - 2, gw, NWG);
            __syncthreads();
        }
    }
    // ---- horizontal smoothing (k_fused_tma's stage, one channel per group) ---------------------------------------
    {
        constexpr int NQ = HW / 4;
        static_assert(HW % 4 == 0 && RW % 4 == 0 && 4 * (NQ - 1) + 8 <= RW, "quads stay inside a raw row, rows stay 16-byte aligned");
        for (int task = lt; task < RHt * NQ; task += NT) {
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
            st4(&ch.hs[ry][4 * q], out);
        }
    }
    __syncthreads();
    if (border) {
        fixup(&ch.hs[0][0], RHt, HW, HW, x0 - 2, y0 - 4, gw, NWG);
        __syncthreads();
    }
    // ---- vertical smoothing (sliding window), blend with smoothed Im1, temporal difference ------------------------
    {
        const int by0 = seg * BROWS;
        if (by0 < SH) {
            const int bx = col;
            const T* hcol = &ch.hs[0][0] + bx;
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
                ch.bl[by][bx] = t + acc * (T)0.6;
                const int ccx = bx - 2, ccy = by - 2;
                if (ccx >= 0 && ccx < TX && ccy >= 0 && ccy < TY) ch.dt[ccy][ccx] = acc - s1v;
            }
        }
        if (lt < 4 * SH) {       // the four halo columns right of the tile, one entry per thread
            const int bx = TX + (lt & 3), by = lt >> 2;
            const T* hcol = &ch.hs[0][0] + bx;
            T acc = 0;
            acc += hcol[(by + 0) * HW] * g0; acc += hcol[(by + 1) * HW] * g1; acc += hcol[(by + 2) * HW] * g2;
            acc += hcol[(by + 3) * HW] * g3; acc += hcol[(by + 4) * HW] * g4;
            const T s1v = s1t[by * RW + bx + 2];
            const T t = s1v * (T)0.4;
            ch.bl[by][bx] = t + acc * (T)0.6;
            const int ccx = bx - 2, ccy = by - 2;
            if (ccx < TX && ccy >= 0 && ccy < TY) ch.dt[ccy][ccx] = acc - s1v;
        }
    }
    __syncthreads();
    if (border) {
        fixup(&ch.bl[0][0], SH, BW, BW, x0 - 2, y0 - 2, gw, NWG);
        __syncthreads();
    }
    // ---- derivatives of the blend at the centre pixels; this channel's factors of the psi-weighted products ----------
    {
        const int cy0 = seg * PPT;
        const T* bcol = &ch.bl[0][0] + (col + 2);
        T v0 = bcol[(cy0 + 0) * BW], v1 = bcol[(cy0 + 1) * BW], v2 = bcol[(cy0 + 2) * BW], v3 = bcol[(cy0 + 3) * BW];
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            const int cy = cy0 + k;
            T v4 = bcol[(cy + 4) * BW];
            const T* brow = &ch.bl[cy + 2][col];
            T ix = 0, iy = 0;
            ix += brow[0] * d0; ix += brow[1] * d1; ix += brow[3] * d3; ix += brow[4] * d4;
            iy += v0 * d0; iy += v1 * d1; iy += v3 * d3; iy += v4 * d4;
            v0 = v1; v1 = v2; v2 = v3; v3 = v4;
            const T it = ch.dt[cy][col];
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
            ch.out[0][cy][col] = px; ch.out[1][cy][col] = py; ch.out[2][cy][col] = ix; ch.out[3][cy][col] = iy; ch.out[4][cy][col] = it;
        }
    }
    __syncthreads();
    // ---- sums over the channels, in channel order, with k_fused_tma's expressions (group 0 owns the centre pixels) ----
    T sxy[PPT], sx2[PPT], sy2[PPT], stx[PPT], sty[PPT];
#pragma unroll
    for (int k = 0; k < PPT; k++) sxy[k] = sx2[k] = sy2[k] = stx[k] = sty[k] = 0;
    if (g == 0) {
        for (int c = 0; c < C; c++) {
#pragma unroll
            for (int k = 0; k < PPT; k++) {
                const int cy = seg * PPT + k;
                const T px = chs[c].out[0][cy][col], py = chs[c].out[1][cy][col], ix = chs[c].out[2][cy][col],
                        iy = chs[c].out[3][cy][col], it = chs[c].out[4][cy][col];
                sxy[k] += px * iy; sx2[k] += px * ix; sy2[k] += py * iy; stx[k] += px * it; sty[k] += py * it;
            }
        }
    }
    __syncthreads();     // hs / bl of channel 0 are free for the u / v tiles of the non-staged path

    // ---- tail: u, v tiles (halo 1) -> phi on the tile plus its left / up halo -> Laplacian, right-hand sides ----------
    auto tail = [&](T* tu, T* tv, auto stride_c) {
    constexpr int US = decltype(stride_c)::value;      // row stride of the u / v tiles; entry (uy, ux) is pixel (y0-1+uy, x0-1+ux)
    auto load_uv = [&](bool with_increment) {
        for (int uy = warp; uy < UH; uy += nwarp) {
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
    if (WARP) {
        // the new flow u + wdu on the tile and its halo (phi AND the Laplacian act on it: it is this iteration's u) was left
        // in the u / v tile buffers by the geometry pass at the head of the kernel, replicate-clamped already
    } else if (uv_staged) {
        if (border) {                                   // replicate clamp over TMA's zero fill
            fixup(tu - 3, UH, RW, RW, x0 - 4, y0 - 1, warp, nwarp);
            fixup(tv - 3, UH, RW, RW, x0 - 4, y0 - 1, warp, nwarp);
            __syncthreads();
        }
    } else {
        load_uv(true);
        __syncthreads();
    }
    for (int py = warp; py < PH; py += nwarp) {
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
    if (!WARP && !uv_staged) {
        load_uv(false);
        __syncthreads();
    }
    if (g != 0 || PX >= W) return;
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
    if (WARP || uv_staged) tail(tu_s + 3, tv_s + 3, std::integral_constant<int, RW>{});
    else tail(&chs[0].hs[0][0], &chs[0].hs[0][0] + UW * UH, std::integral_constant<int, UW>{});
}

}  // namespace pf
