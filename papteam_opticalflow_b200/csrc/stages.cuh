// Single-stage entry points: each runs the SAME kernels the Plan uses on caller-supplied HWC
// float64 host buffers, so tests can check every stage in isolation
// (the reference exposes these stages as public statics, S/OpticalFlow.h:28-56).
#pragma once
#include "solver.cuh"

namespace pf {

// Owning planar device image with HWC-double upload/download.
template <typename T>
struct DevImg {
    Img<T> img;
    DevImg() {}
    DevImg(int w, int h, int c) { alloc(w, h, c); }
    DevImg(const DevImg&) = delete;
    DevImg& operator=(const DevImg&) = delete;
    ~DevImg() { if (img.p) cudaFree(img.p); }
    void alloc(int w, int h, int c) {
        img.w = w; img.h = h; img.c = c;
        img.pitch = pitch_for(w);
        img.plane = plane_for(w, h);
        PF_CUDA(cudaMalloc(&img.p, img.elems() * sizeof(T)));
        PF_CUDA(cudaMemset(img.p, 0, img.elems() * sizeof(T)));
    }
    void upload(const double* hwc) {
        size_t n = (size_t)img.w * img.h * img.c;
        double* d = nullptr;
        PF_CUDA(cudaMalloc(&d, n * sizeof(double)));
        cudaError_t e = cudaMemcpy(d, hwc, n * sizeof(double), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            k_import_hwc<T><<<dim3(ceil_div(img.w, 128), img.h), 128>>>(d, img);
            e = cudaDeviceSynchronize();
        }
        cudaFree(d);
        PF_CUDA(e);
    }
    void download(double* hwc) const {
        size_t n = (size_t)img.w * img.h * img.c;
        double* d = nullptr;
        PF_CUDA(cudaMalloc(&d, n * sizeof(double)));
        k_export_hwc<T><<<dim3(ceil_div(img.w, 128), img.h), 128>>>(img, d);
        cudaError_t e = cudaMemcpy(hwc, d, n * sizeof(double), cudaMemcpyDeviceToHost);
        cudaFree(d);
        PF_CUDA(e);
    }
};


// download of a non-owning view
template <typename T>
struct DevImgView {
    static void download(const Img<T>& img, double* hwc) {
        size_t n = (size_t)img.w * img.h * img.c;
        double* d = nullptr;
        PF_CUDA(cudaMalloc(&d, n * sizeof(double)));
        k_export_hwc<T><<<dim3(ceil_div(img.w, 128), img.h), 128>>>(img, d);
        cudaError_t e = cudaMemcpy(hwc, d, n * sizeof(double), cudaMemcpyDeviceToHost);
        cudaFree(d);
        PF_CUDA(e);
    }
};

// iu = omega/(dx2 + .05 alpha + alpha sum_nbr phi), iv likewise: the same expression k_assemble
// evaluates, for stage tests that start from dx2/dy2 planes.
template <typename T>
__global__ void k_sor_inverse(const T* __restrict__ phi, const T* __restrict__ dx2, const T* __restrict__ dy2,
                              T* __restrict__ iu, T* __restrict__ iv, int w, int h, int pitch, T alpha, T omega) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    size_t o = (size_t)y * pitch + x;
    T cf = 0;
    if (x > 0) cf += phi[o - 1];
    if (x < w - 1) cf += phi[o];
    if (y > 0) cf += phi[o - pitch];
    if (y < h - 1) cf += phi[o];
    cf *= alpha;
    T reg = alpha * (T)0.05;
    iu[o] = omega / (dx2[o] + reg + cf);
    iv[o] = omega / (dy2[o] + reg + cf);
}

template <typename T>
struct Stages {
    static dim3 g2(int w, int h, int z = 1) { return dim3(ceil_div(w, 128), h, z); }
    static Taps<T> d5() { const double r[5] = {1.0 / 12, -8.0 / 12, 0.0 / 12, 8.0 / 12, -1.0 / 12}; return make_taps<T>(r, 2); }
    static Taps<T> g5() { const double r[5] = {0.02, 0.11, 0.74, 0.11, 0.02}; return make_taps<T>(r, 2); }
    static Taps<T> d3() { const double r[3] = {-0.5, 0, 0.5}; return make_taps<T>(r, 1); }

    static void pyramid(double* out, const double* im, int h, int w, int c, double ratio, int levels) {
        std::vector<Level> geo = level_geometry(w, h, ratio, levels);
        std::vector<std::unique_ptr<DevImg<T>>> pyr;
        for (auto& g : geo) pyr.emplace_back(new DevImg<T>(g.w, g.h, c));
        pyr[0]->upload(im);
        DevImg<T> tmp(w, h, c), blur(w, h, c);
        size_t off = 0;
        for (int i = 0; i < levels; i++) {
            if (i > 0) {
                const Level& g = geo[i];
                Img<T> src = pyr[g.src]->img, b = src;
                if (g.half > 0) {
                    Taps<T> gt = make_taps<T>(g.taps.data(), g.half);
                    Img<T> t = tmp.img; t.w = src.w; t.h = src.h; t.pitch = src.pitch; t.plane = src.plane;
                    b = blur.img; b.w = src.w; b.h = src.h; b.pitch = src.pitch; b.plane = src.plane;
                    k_filter_h<T><<<g2(src.w, src.h, c), 128>>>(src, t, gt);
                    k_filter_v<T><<<g2(src.w, src.h, c), 128>>>(t, b, gt);
                }
                k_resize<T><<<g2(g.w, g.h), 128>>>(b, pyr[i]->img, g.rate, g.rate, (T)1, 0);
            }
            PF_CUDA(cudaDeviceSynchronize());
            pyr[i]->download(out + off);
            off += (size_t)geo[i].w * geo[i].h * c;
        }
    }

    static int im2feature(double* feat, const double* im, int h, int w, int c, int swap) {
        int fc = c == 1 ? 3 : (c == 3 ? 5 : c);
        DevImg<T> a(w, h, c), f(w, h, fc);
        a.upload(im);
        if (c == 1 || c == 3) k_im2feature<T><<<g2(w, h), 128>>>(a.img, f.img, d5(), swap);
        else k_copy<T><<<g2(w, h, c), 128>>>(a.img, f.img);
        PF_CUDA(cudaDeviceSynchronize());
        f.download(feat);
        return fc;
    }

    static void getdxs(double* odx, double* ody, double* odt, const double* im1, const double* im2, int h, int w, int c) {
        DevImg<T> a(w, h, c), b(w, h, c), t(w, h, c), s1(w, h, c), s2(w, h, c), bl(w, h, c), dx(w, h, c), dy(w, h, c), dt(w, h, c);
        a.upload(im1);
        b.upload(im2);
        k_filter_h<T><<<g2(w, h, c), 128>>>(a.img, t.img, g5());
        k_filter_v<T><<<g2(w, h, c), 128>>>(t.img, s1.img, g5());
        k_filter_h<T><<<g2(w, h, c), 128>>>(b.img, t.img, g5());
        k_filter_v<T><<<g2(w, h, c), 128>>>(t.img, s2.img, g5());
        k_blend_dt<T><<<g2(w, h, c), 128>>>(s1.img, s2.img, bl.img, dt.img);
        k_filter_h<T><<<g2(w, h, c), 128>>>(bl.img, dx.img, d5());
        k_filter_v<T><<<g2(w, h, c), 128>>>(bl.img, dy.img, d5());
        PF_CUDA(cudaDeviceSynchronize());
        dx.download(odx);
        dy.download(ody);
        dt.download(odt);
    }

    static void warpfl(double* out, const double* im1, const double* im2, const double* vx, const double* vy, int h, int w, int c) {
        DevImg<T> a(w, h, c), b(w, h, c), o(w, h, c), u(w, h, 1), v(w, h, 1);
        a.upload(im1); b.upload(im2); u.upload(vx); v.upload(vy);
        k_update_warp<T><<<warp_grid(w, h), 128>>>(a.img, b.img, o.img, u.img.p, v.img.p, nullptr, nullptr, u.img.pitch);
        PF_CUDA(cudaDeviceSynchronize());
        o.download(out);
    }

    static void resize_to(double* dst, const double* src, int h, int w, int c, int dh, int dw, double scale) {
        DevImg<T> a(w, h, c), o(dw, dh, c);
        a.upload(src);
        k_resize<T><<<g2(dw, dh), 128>>>(a.img, o.img, (double)dw / w, (double)dh / h, (T)scale, scale != 1.0);
        PF_CUDA(cudaDeviceSynchronize());
        o.download(dst);
    }

    static void bicubic(double* out, const double* ref, const double* im2, const double* vx, const double* vy, int h, int w, int c) {
        DevImg<T> r(w, h, c), b(w, h, c), ix(w, h, c), iy(w, h, c), ixy(w, h, c), u(w, h, 1), v(w, h, 1);
        r.upload(ref); b.upload(im2); u.upload(vx); v.upload(vy);
        BicubicTable tab = make_bicubic_table(), *dtab = nullptr;
        double* dout = nullptr;
        size_t n = (size_t)w * h * c;
        PF_CUDA(cudaMalloc(&dtab, sizeof(tab)));
        PF_CUDA(cudaMalloc(&dout, n * sizeof(double)));
        PF_CUDA(cudaMemcpy(dtab, &tab, sizeof(tab), cudaMemcpyHostToDevice));
        k_filter_h<T><<<g2(w, h, c), 128>>>(b.img, ix.img, d3());
        k_filter_v<T><<<g2(w, h, c), 128>>>(b.img, iy.img, d3());
        k_filter_v<T><<<g2(w, h, c), 128>>>(ix.img, ixy.img, d3());
        BicubicOut<T> bo;
        bo.hwc = dout; bo.clamp = 1;
        k_bicubic_warp<T><<<g2(w, h), 128>>>(r.img, b.img, ix.img, iy.img, ixy.img, u.img.p, v.img.p, u.img.pitch, dtab, bo);
        cudaError_t e = cudaMemcpy(out, dout, n * sizeof(double), cudaMemcpyDeviceToHost);
        cudaFree(dtab);
        cudaFree(dout);
        PF_CUDA(e);
    }

    static void assemble(double* ophi, double* odxy, double* odx2, double* ody2, double* obu, double* obv,
                         const double* imdx, const double* imdy, const double* imdt, const double* u,
                         const double* v, const double* du, const double* dv, const double* lap,
                         double alpha, int h, int w, int c) {
        DevImg<T> dx(w, h, c), dy(w, h, c), dt(w, h, c), U(w, h, 1), V(w, h, 1), DU(w, h, 1), DV(w, h, 1);
        DevImg<T> phi(w, h, 1), dxy(w, h, 1), iu(w, h, 1), iv(w, h, 1), bu(w, h, 1), bv(w, h, 1), dx2(w, h, 1), dy2(w, h, 1);
        dx.upload(imdx); dy.upload(imdy); dt.upload(imdt); U.upload(u); V.upload(v);
        if (du) DU.upload(du);
        if (dv) DV.upload(dv);
        double* dlap = nullptr;
        if (lap) {
            PF_CUDA(cudaMalloc(&dlap, 64 * sizeof(double)));
            PF_CUDA(cudaMemcpy(dlap, lap, c * sizeof(double), cudaMemcpyHostToDevice));
        }
        T eps = (T)std::pow(0.001, 2);
        k_phi<T><<<g2(w, h), 128>>>(U.img.p, V.img.p, du ? DU.img.p : nullptr, dv ? DV.img.p : nullptr, phi.img.p, w, h, U.img.pitch, eps);
        AssembleArgs<T> a;
        a.imdx = dx.img; a.imdy = dy.img; a.imdt = dt.img;
        a.u = U.img.p; a.v = V.img.p; a.du = du ? DU.img.p : nullptr; a.dv = dv ? DV.img.p : nullptr;
        a.phi = phi.img.p; a.lap = dlap; a.gm = nullptr;
        a.dxy = dxy.img.p; a.iu = iu.img.p; a.iv = iv.img.p; a.bu = bu.img.p; a.bv = bv.img.p;
        a.dx2 = dx2.img.p; a.dy2 = dy2.img.p;
        a.w = w; a.h = h; a.pitch = U.img.pitch;
        a.alpha = (T)alpha; a.omega = (T)1.8; a.eps = eps;
        k_assemble<T><<<g2(w, h), 128>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (dlap) cudaFree(dlap);
        PF_CUDA(e);
        phi.download(ophi); dxy.download(odxy); dx2.download(odx2); dy2.download(ody2);
        bu.download(obu); bv.download(obv);
    }
    // SOR alone, taking dx2/dy2 like the reference's loop rather than the precomputed inverses
    static void sor(double* odu, double* odv, const double* phi, const double* dxy, const double* dx2,
                    const double* dy2, const double* bu, const double* bv, double alpha, int nsor,
                    int h, int w, int mode, int device, int repeats, double* ms_per_solve, double* launches) {
        DevImg<T> PHI(w, h, 1), DXY(w, h, 1), DX2(w, h, 1), DY2(w, h, 1), BU(w, h, 1), BV(w, h, 1);
        DevImg<T> IU(w, h, 1), IV(w, h, 1), DU(w, h, 1), DV(w, h, 1), DU2(w, h, 1), DV2(w, h, 1);
        PHI.upload(phi); DXY.upload(dxy); DX2.upload(dx2); DY2.upload(dy2); BU.upload(bu); BV.upload(bv);
        k_sor_inverse<T><<<g2(w, h), 128>>>(PHI.img.p, DX2.img.p, DY2.img.p, IU.img.p, IV.img.p, w, h, PHI.img.pitch, (T)alpha, (T)1.8);
        PF_CUDA(cudaDeviceSynchronize());
        cudaStream_t st;
        PF_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        cudaEvent_t e0, e1;
        PF_CUDA(cudaEventCreate(&e0));
        PF_CUDA(cudaEventCreate(&e1));
        SorRunner<T> run;
        run.init(mode, device, st);
        SorArgs<T> a;
        a.phi = PHI.img.p; a.dxy = DXY.img.p; a.iu = IU.img.p; a.iv = IV.img.p; a.bu = BU.img.p; a.bv = BV.img.p;
        a.du = a.dv = nullptr; a.du_in = a.dv_in = nullptr;
        a.w = w; a.h = h; a.pitch = PHI.img.pitch; a.alpha = (T)alpha; a.omega = (T)1.8;
        T *du = DU.img.p, *dv = DV.img.p, *du2 = DU2.img.p, *dv2 = DV2.img.p;
        int n = 0;
        try {
            n = run.run(a, du, dv, du2, dv2, nsor);       // warm-up (and the result for repeats<=1)
            PF_CUDA(cudaEventRecord(e0, st));
            for (int i = 1; i < repeats; i++) run.run(a, du, dv, du2, dv2, nsor);
            PF_CUDA(cudaEventRecord(e1, st));
            PF_CUDA(cudaStreamSynchronize(st));
            PF_CHECK_LAUNCH();
            run.check_lex();
        } catch (...) {
            cudaStreamDestroy(st); cudaEventDestroy(e0); cudaEventDestroy(e1);
            throw;
        }
        float ms = 0;
        if (repeats > 1) PF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms_per_solve) *ms_per_solve = repeats > 1 ? ms / (repeats - 1) : 0;
        if (launches) *launches = n;
        cudaStreamDestroy(st); cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (odu) {
            Img<T> r = DU.img; r.p = du; DevImgView<T>::download(r, odu);
            r.p = dv; DevImgView<T>::download(r, odv);
        }
    }
};

}  // namespace pf
