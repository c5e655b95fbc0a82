// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or shipped with the product.
//
// C-ABI harness around the UNMODIFIED reference solver.  This file is compiled together with
// the reference's own sources *where they lie* under /root/reference (see oracle/Makefile);
// nothing from the reference is copied into this repository.  The resulting shared objects land
// in oracle/_ref/ (git-ignored) and are used (a) to pin the plain-C oracle restatement in
// oracle/pyflow_oracle.c, (b) to generate tests/golden/ fixtures, (c) as the "reference" CPU
// baseline that bench.py times next to the GPU path.
//
// Two flavours are built from this one file:
//   libpyflow_ref_serial.so    <- Code/Serial/src   (deterministic: THE correctness oracle)
//   libpyflow_ref_parallel.so  <- Code/Parallel/src (-DPAP_PARALLEL; racy for nCores>1, timing only)
//
// Entry points mirror the reference's own interfaces:
//   ref_coarse2fine_flow_levels  = Coarse2FineFlowWrapper (S/Coarse2FineFlowWrapper.cpp:14-51)
//   ref_coarse2fine_flow         = upstream-shaped driver (alpha, ratio, minWidth, nOuter, nInner,
//                                  nSOR, colType) re-driving the level loop of
//                                  S/OpticalFlow.cpp:784-842 through the reference's PUBLIC STATIC
//                                  stage functions (S/OpticalFlow.h:28-56), because the fork hard-codes
//                                  its parameters (S/OpticalFlow.cpp:747-751).
//   ref_stage_*                  = the individual public statics, for per-stage parity dumps.
#include "Image.h"
#include "OpticalFlow.h"
#include "GaussianPyramid.h"
#include <cstring>
#include <map>
#include <string>
#include <cstdio>

#ifdef PAP_PARALLEL
extern int GLOBAL_nThreads;
#endif

namespace {

void load(DImage& dst, const double* src, int h, int w, int c) {
    dst.allocate(w, h, c);
    memcpy(dst.pData, src, sizeof(double) * (size_t)h * w * c);
}
void store(double* dst, const DImage& src) {
    memcpy(dst, src.pData, sizeof(double) * (size_t)src.nelements());
}

void smooth_flow_sor(const DImage& I1, const DImage& I2, DImage& warp, DImage& u, DImage& v,
                     double alpha, int nOuter, int nInner, int nSOR) {
#ifdef PAP_PARALLEL
    OpticalFlow::SmoothFlowSOR(I1, I2, warp, u, v, alpha, nOuter, nInner, nSOR, GLOBAL_nThreads);
#else
    OpticalFlow::SmoothFlowSOR(I1, I2, warp, u, v, alpha, nOuter, nInner, nSOR);
#endif
}

void init_noise(int image_channels) {
    // S/OpticalFlow.cpp:768-776
    switch (OpticalFlow::noiseModel) {
        case OpticalFlow::GMixture:
            OpticalFlow::GMPara.reset(image_channels + 2);
            break;
        case OpticalFlow::Lap:
            OpticalFlow::LapPara.allocate(image_channels + 2);
            for (int i = 0; i < OpticalFlow::LapPara.dim(); i++) OpticalFlow::LapPara[i] = 0.02;
            break;
    }
}

}  // namespace

extern "C" {

// The alternative solver branches are selected in the reference by two PUBLIC
// static members (S/OpticalFlow.h:19-27; defaults Bilinear / Lap at S/OpticalFlow.cpp:33-34): the harness
// assigns them, the reference's sources stay untouched.  interpolation: 0 Bilinear, 1 Bicubic;
// noise_model: 0 GMixture, 1 Lap (the reference's enum order).
void ref_set_variant(int interpolation, int noise_model) {
    OpticalFlow::interpolation = interpolation ? OpticalFlow::Bicubic : OpticalFlow::Bilinear;
    OpticalFlow::noiseModel = noise_model ? OpticalFlow::Lap : OpticalFlow::GMixture;
}
// current mixture parameters (after a solve: the state left by the last estGaussianMixture call)
void ref_gm_get(double* alpha, double* sigma, double* beta, int c) {
    for (int k = 0; k < c && k < OpticalFlow::GMPara.nChannels; k++) {
        alpha[k] = OpticalFlow::GMPara.alpha[k];
        sigma[k] = OpticalFlow::GMPara.sigma[k];
        beta[k] = OpticalFlow::GMPara.beta[k];
    }
}

// Fills `timing_out` (if non-NULL, capacity `cap`) with "key=value\n" lines of the timing map.
int ref_coarse2fine_flow_levels(double* vx, double* vy, double* warpI2, const double* Im1,
                                const double* Im2, int pyramidLevels, int nCores, int h, int w,
                                int c, char* timing_out, int cap) {
    DImage I1, I2, VX, VY, W2;
    load(I1, Im1, h, w, c);
    load(I2, Im2, h, w, c);
    I1.setColorType(0);
    I2.setColorType(0);
    std::map<std::string, std::string> timing;
#ifdef PAP_PARALLEL
    OpticalFlow::Coarse2FineFlow(&timing, VX, VY, W2, I1, I2, pyramidLevels, nCores);
#else
    (void)nCores;
    OpticalFlow::Coarse2FineFlow(&timing, VX, VY, W2, I1, I2, pyramidLevels);
#endif
    store(vx, VX);
    store(vy, VY);
    store(warpI2, W2);
    if (timing_out && cap > 0) {
        std::string s;
        for (auto& kv : timing) s += kv.first + "=" + kv.second + "\n";
        strncpy(timing_out, s.c_str(), cap - 1);
        timing_out[cap - 1] = 0;
    }
    return 0;
}

// Pyramid geometry + data from the real GaussianPyramid.  Call once with data==NULL to get sizes.
// use_min_width!=0 -> ConstructPyramid(image, ratio, minWidth) ; else ConstructPyramidLevels(levels)
int ref_stage_pyramid(const double* im, int h, int w, int c, double ratio, int use_min_width,
                      int min_width_or_levels, int* widths, int* heights, double* data) {
    DImage I;
    load(I, im, h, w, c);
    GaussianPyramid P;
    if (use_min_width) P.ConstructPyramid(I, ratio, min_width_or_levels);
    else               P.ConstructPyramidLevels(I, ratio, min_width_or_levels);
    size_t off = 0;
    for (int k = 0; k < P.nlevels(); k++) {
        if (widths)  widths[k]  = P.Image(k).width();
        if (heights) heights[k] = P.Image(k).height();
        if (data) { store(data + off, P.Image(k)); off += P.Image(k).nelements(); }
    }
    return P.nlevels();
}

#ifndef PAP_PARALLEL
// ---- upstream-shaped, fully parameterised driver over the reference's public statics ----------
int ref_coarse2fine_flow(double* vx, double* vy, double* warpI2, const double* Im1,
                         const double* Im2, double alpha, double ratio, int minWidth, int nOuter,
                         int nInner, int nSOR, int colType, int h, int w, int c) {
    DImage I1, I2, VX, VY, W2;
    load(I1, Im1, h, w, c);
    load(I2, Im2, h, w, c);
    I1.setColorType(colType);
    I2.setColorType(colType);
    GaussianPyramid P1, P2;
    P1.ConstructPyramid(I1, ratio, minWidth);
    P2.ConstructPyramid(I2, ratio, minWidth);
    // NB: ConstructPyramid substitutes 0.75 for an out-of-range ratio only in its local copy; the
    // upstream driver keeps dividing the flow by the caller's ratio, and so does this harness.
    init_noise(I1.nchannels());
    DImage F1, F2, WF;
    for (int k = P1.nlevels() - 1; k >= 0; k--) {
        int lw = P1.Image(k).width(), lh = P1.Image(k).height();
        OpticalFlow::im2feature(F1, P1.Image(k));
        OpticalFlow::im2feature(F2, P2.Image(k));
        if (k == P1.nlevels() - 1) {
            VX.allocate(lw, lh);
            VY.allocate(lw, lh);
            WF.copyData(F2);
        } else {
            VX.imresize(lw, lh);
            VX.Multiplywith(1 / ratio);
            VY.imresize(lw, lh);
            VY.Multiplywith(1 / ratio);
            if (OpticalFlow::interpolation == OpticalFlow::Bilinear) OpticalFlow::warpFL(WF, F1, F2, VX, VY);
            else F2.warpImageBicubicRef(F1, WF, VX, VY);   // S/OpticalFlow.cpp:812-815
        }
        smooth_flow_sor(F1, F2, WF, VX, VY, alpha, nOuter + k, nInner, nSOR + k * 3);
    }
    I2.warpImageBicubicRef(I1, W2, VX, VY);
    W2.threshold();
    store(vx, VX);
    store(vy, VY);
    store(warpI2, W2);
    return 0;
}

// ---- individual stages (all arrays HWC double, caller-allocated) -------------------------------
int ref_stage_im2feature(double* feat, const double* im, int h, int w, int c) {
    DImage I, F;
    load(I, im, h, w, c);
    OpticalFlow::im2feature(F, I);
    store(feat, F);
    return F.nchannels();
}

void ref_stage_getdxs(double* imdx, double* imdy, double* imdt, const double* im1,
                      const double* im2, int h, int w, int c) {
    DImage A, B, dx, dy, dt;
    load(A, im1, h, w, c);
    load(B, im2, h, w, c);
    OpticalFlow::getDxs(dx, dy, dt, A, B);
    store(imdx, dx);
    store(imdy, dy);
    store(imdt, dt);
}

void ref_stage_warpfl(double* warp, const double* im1, const double* im2, const double* vx,
                      const double* vy, int h, int w, int c) {
    DImage A, B, U, V, W;
    load(A, im1, h, w, c);
    load(B, im2, h, w, c);
    load(U, vx, h, w, 1);
    load(V, vy, h, w, 1);
    OpticalFlow::warpFL(W, A, B, U, V);
    store(warp, W);
}

void ref_stage_laplacian(double* out, const double* in, const double* weight, int h, int w) {
    DImage I, Wt, O;
    load(I, in, h, w, 1);
    load(Wt, weight, h, w, 1);
    OpticalFlow::Laplacian(O, I, Wt);
    store(out, O);
}

// resize to explicit destination size (flow upsampling path, S/Image.h:778-783) then scale
void ref_stage_resize_to(double* dst, const double* src, int h, int w, int c, int dh, int dw,
                         double scale) {
    DImage I;
    load(I, src, h, w, c);
    I.imresize(dw, dh);
    if (scale != 1.0) I.Multiplywith(scale);
    store(dst, I);
}

void ref_stage_gaussian(double* dst, const double* src, int h, int w, int c, double sigma,
                        int fsize) {
    DImage I, O;
    load(I, src, h, w, c);
    I.GaussianSmoothing(O, sigma, fsize);
    store(dst, O);
}

void ref_stage_bicubic(double* out, const double* ref, const double* im2, const double* vx,
                       const double* vy, int h, int w, int c) {
    DImage R, B, U, V, O;
    load(R, ref, h, w, c);
    load(B, im2, h, w, c);
    load(U, vx, h, w, 1);
    load(V, vy, h, w, 1);
    B.warpImageBicubicRef(R, O, U, V);
    O.threshold();
    store(out, O);
}

// SmoothFlowSOR on caller-supplied feature images; u,v,warp are in/out.  lap_init_channels>0
// re-initialises LapPara as the driver does (pass the *image* channel count, e.g. 3 for RGB).
void ref_stage_smoothflow_sor(const double* f1, const double* f2, double* warp, double* u,
                              double* v, double alpha, int nOuter, int nInner, int nSOR, int h,
                              int w, int c, int lap_init_channels) {
    DImage A, B, Wp, U, V;
    load(A, f1, h, w, c);
    load(B, f2, h, w, c);
    load(Wp, warp, h, w, c);
    load(U, u, h, w, 1);
    load(V, v, h, w, 1);
    if (lap_init_channels > 0) init_noise(lap_init_channels);
    smooth_flow_sor(A, B, Wp, U, V, alpha, nOuter, nInner, nSOR);
    store(warp, Wp);
    store(u, U);
    store(v, V);
}

// Flow file format: the reference's own writer / reader, file in, file out.
// flow is (h, w, 2) interleaved (u, v).  Returns 1 on success.
int ref_save_optical_flow(const double* flow, int h, int w, const char* path) {
    DImage F;
    load(F, flow, h, w, 2);
    return OpticalFlow::SaveOpticalFlow(F, path) ? 1 : 0;
}

int ref_load_optical_flow(double* flow, int h, int w, const char* path) {
    DImage F;
    if (!OpticalFlow::LoadOpticalFlow(path, F)) return 0;
    if (F.height() != h || F.width() != w || F.nchannels() != 2) return 0;
    store(flow, F);
    return 1;
}
#endif  // !PAP_PARALLEL

}  // extern "C"
