// Elementwise / stencil / gather kernels of the solver, templated on the arithmetic type
// (float for the fast mode, double for the parity mode).  In the double instantiation every
// expression keeps the reference's operation order (the translation unit is compiled with
// -fmad=false), so results agree with the reference at the bit level; see SURVEY.md Appendix A.
// Citations: S/ = /root/reference/Code/Serial/src/.
#pragma once
#include "common.cuh"

namespace pf {

__device__ __forceinline__ int clampi(int v, int n) { return min(max(v, 0), n - 1); }

// ---------------------------------------------------------------------------------------------
// Boundary conversion: HWC float64 (the numpy buffer, P/Coarse2FineFlowWrapper.cpp:23-26) <-> planar T
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_import_hwc(const double* __restrict__ src, Img<T> dst) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dst.w) return;
    const double* s = src + ((size_t)y * dst.w + x) * dst.c;
    for (int k = 0; k < dst.c; k++) dst.ch(k)[(size_t)y * dst.pitch + x] = (T)s[k];
}

template <typename T>
__global__ void k_export_hwc(Img<T> src, double* __restrict__ dst) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= src.w) return;
    double* d = dst + ((size_t)y * src.w + x) * src.c;
    for (int k = 0; k < src.c; k++) d[k] = (double)src.ch(k)[(size_t)y * src.pitch + x];
}

// Compact boundary for sequences ("next" rows f1/f2 of SURVEY.md 8f): uint8 HWC frames in -- the
// conversion the reference driver does on the host, astype(float)/255. (Par/OpticalFlowCalculation.py:
// 70-71), evaluated in double so the pixel values are bit-identical -- and the flow out as
// interleaved float32 (u, v), the driver's `flow = concat(u, v)` (:76).
template <typename T>
__global__ void k_import_u8(const unsigned char* __restrict__ src, Img<T> dst) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dst.w) return;
    const unsigned char* s = src + ((size_t)y * dst.w + x) * dst.c;
    for (int k = 0; k < dst.c; k++) dst.ch(k)[(size_t)y * dst.pitch + x] = (T)((double)s[k] / 255.0);
}

template <typename T>
__global__ void k_export_flow_f32(const T* __restrict__ u, const T* __restrict__ v, int pitch, int w, float2* __restrict__ out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    out[(size_t)y * w + x] = make_float2((float)u[(size_t)y * pitch + x], (float)v[(size_t)y * pitch + x]);
}

// The reference's 16-bit flow encoding (OpticalFlow::SaveOpticalFlow, S/OpticalFlow.cpp:993-1003):
// q = (unsigned short)((min(max(f, -200), 200) + 200) * 160), evaluated in double, (u, v) interleaved.
__device__ __forceinline__ unsigned short flow_to_u16(double f) {
    f = fmin(fmax(f, -200.0), 200.0);
    return (unsigned short)((f + 200.0) * 160.0);
}
template <typename T>
__global__ void k_export_flow_u16(const T* __restrict__ u, const T* __restrict__ v, int pitch, int w, ushort2* __restrict__ out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    out[(size_t)y * w + x] = make_ushort2(flow_to_u16((double)u[(size_t)y * pitch + x]), flow_to_u16((double)v[(size_t)y * pitch + x]));
}

// ---------------------------------------------------------------------------------------------
// Flow visualisation of the reference driver, generateOutputFlowImageFile,
// Par/OpticalFlowCalculation.py:143-162: (mag, ang) = cv2.cartToPolar(u, v); H = ang*180/pi/2,
// S = 255, V = cv2.normalize(mag, 0..255, NORM_MINMAX), both stored to uint8 by truncation;
// BGR = cv2.cvtColor(hsv, COLOR_HSV2BGR).  OpenCV's arithmetic for these calls (float32 magnitude with
// one FMA, fastAtan32f's degree polynomial with FMAs, float64 min-max scaling, float32 HSV sector
// formula whose 8-bit results are truncated) is restated operation by operation with explicitly
// rounded intrinsics, so nvcc can neither contract nor reorder anything (the tests hold a numpy
// restatement of the same operations that is pinned against cv2 4.13).
// Two kernels: a min/max reduction of the magnitude (non-negative floats order like their bit
// patterns, so atomicMin/atomicMax on the bits), then the per-pixel conversion.
// Inputs are addressed as base[y*pitch + x*stride]: planar solver planes (stride 1) or an
// interleaved (u, v) host layout (stride 2, v = u + 1).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float flow_mag(float x, float y) { return __fsqrt_rn(__fmaf_rn(x, x, __fmul_rn(y, y))); }

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float k = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, k), p3 = __fmul_rn(-0.3258083974640975f, k);
    const float p5 = __fmul_rn(0.1555786518463281f, k), p7 = __fmul_rn(-0.04432655554792128f, k);
    const float ax = fabsf(x), ay = fabsf(y);
    const float c = __fdiv_rn(fminf(ax, ay), __fadd_rn(fmaxf(ax, ay), 2.220446049250313e-16f));
    const float c2 = __fmul_rn(c, c);
    float a = __fmaf_rn(c2, p7, p5);
    a = __fmaf_rn(a, c2, p3);
    a = __fmaf_rn(a, c2, p1);
    a = __fmul_rn(a, c);
    if (!(ax >= ay)) a = __fsub_rn(90.f, a);
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

static __global__ void k_fill_f64(double* p, double v, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
static __global__ void k_minmax_init(unsigned int* mm) {
    mm[0] = 0x7f800000u;   // +inf
    mm[1] = 0u;
}

template <typename T>
__global__ void k_flow_mag_minmax(const T* __restrict__ u, const T* __restrict__ v, int pitch, int stride, int w, int h,
                                  unsigned int* __restrict__ mm) {
    unsigned int lo = 0x7f800000u, hi = 0u;
    for (int y = blockIdx.y; y < h; y += gridDim.y)
        for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < w; x += gridDim.x * blockDim.x) {
            const size_t o = (size_t)y * pitch + (size_t)x * stride;
            const unsigned int b = __float_as_uint(flow_mag((float)u[o], (float)v[o]));
            lo = min(lo, b);
            hi = max(hi, b);
        }
    for (int d = 16; d > 0; d >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mm, lo);
        atomicMax(mm + 1, hi);
    }
}

template <typename T>
__global__ void k_flow_to_bgr(const T* __restrict__ u, const T* __restrict__ v, int pitch, int stride, int w,
                              const unsigned int* __restrict__ mm, unsigned char* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const size_t o = (size_t)y * pitch + (size_t)x * stride;
    const float fx = (float)u[o], fy = (float)v[o];
    // hue byte: ang * 180 / pi / 2 in double, truncated
    const float rad = __fmul_rn(fast_atan2_deg(fy, fx), (float)(3.14159265358979323846 / 180.0));
    const double hd = __ddiv_rn(__ddiv_rn(__dmul_rn((double)rad, 180.0), 3.14159265358979323846), 2.0);
    const int H = (int)hd & 255;
    // value byte: min-max scaling to 0..255 in double, truncated
    const double smin = (double)__uint_as_float(mm[0]), smax = (double)__uint_as_float(mm[1]);
    const double range = __dsub_rn(smax, smin);
    const double scale = __dmul_rn(255.0, range > 2.220446049250313e-16 ? __ddiv_rn(1.0, range) : 0.0);
    const double shift = __dsub_rn(0.0, __dmul_rn(smin, scale));
    const int V = (int)__dadd_rn(__dmul_rn((double)flow_mag(fx, fy), scale), shift) & 255;
    // 8-bit HSV -> BGR, hue range 180, S = 255
    float hh = __fmul_rn((float)H, 6.0f / 180.0f);
    const float s = __fmul_rn(255.f, 1.0f / 255.0f), val = __fmul_rn((float)V, 1.0f / 255.0f);
    if (hh >= 6.f) hh = __fsub_rn(hh, 6.f);
    const int sector = (int)floorf(hh);
    const float f = __fsub_rn(hh, (float)sector);
    const float t0 = val, t1 = __fmul_rn(val, __fsub_rn(1.f, s));
    const float t2 = __fmul_rn(val, __fsub_rn(1.f, __fmul_rn(s, f)));
    const float t3 = __fmul_rn(val, __fsub_rn(1.f, __fmul_rn(s, __fsub_rn(1.f, f))));
    float b, g, r;
    switch (sector) {
        case 0: b = t1; g = t3; r = t0; break;
        case 1: b = t1; g = t0; r = t2; break;
        case 2: b = t3; g = t0; r = t1; break;
        case 3: b = t0; g = t2; r = t1; break;
        case 4: b = t0; g = t1; r = t3; break;
        default: b = t2; g = t1; r = t0; break;
    }
    unsigned char* q = out + ((size_t)y * w + x) * 3;
    q[0] = (unsigned char)(int)__fmul_rn(b, 255.f);
    q[1] = (unsigned char)(int)__fmul_rn(g, 255.f);
    q[2] = (unsigned char)(int)__fmul_rn(r, 255.f);
}

template <typename T>
__global__ void k_fill(Img<T> img, T value) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, k = blockIdx.z;
    if (x >= img.w) return;
    img.ch(k)[(size_t)y * img.pitch + x] = value;
}

template <typename T>
__global__ void k_copy(Img<T> src, Img<T> dst) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, k = blockIdx.z;
    if (x >= src.w) return;
    dst.ch(k)[(size_t)y * dst.pitch + x] = src.ch(k)[(size_t)y * src.pitch + x];
}

// ---------------------------------------------------------------------------------------------
// 1-D correlation with replicate borders (S/ImageProcessing.h:259-279, :350-369):
//   dst(p) = sum_{l=-f..f} tap[l+f] * src(clamp(p+l)), ascending l, accumulator starts at 0.
// One thread per output element; channels on blockIdx.z.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_filter_h(Img<T> src, Img<T> dst, Taps<T> t) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, k = blockIdx.z;
    if (x >= src.w) return;
    const T* s = src.ch(k) + (size_t)y * src.pitch;
    T acc = 0;
    for (int l = -t.half; l <= t.half; l++) acc += s[clampi(x + l, src.w)] * t.v[l + t.half];
    dst.ch(k)[(size_t)y * dst.pitch + x] = acc;
}

template <typename T>
__global__ void k_filter_v(Img<T> src, Img<T> dst, Taps<T> t) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, k = blockIdx.z;
    if (x >= src.w) return;
    const T* s = src.ch(k);
    T acc = 0;
    for (int l = -t.half; l <= t.half; l++)
        acc += s[(size_t)clampi(y + l, src.h) * src.pitch + x] * t.v[l + t.half];
    dst.ch(k)[(size_t)y * dst.pitch + x] = acc;
}

// ---------------------------------------------------------------------------------------------
// Separable filter, both passes in one kernel (Image::imfilter_hv, S/Image.h:1347-1356): a 64 x 16
// output tile per CTA, the raw tile (halo fh in x, fv in y) staged in shared memory with replicated
// borders, the horizontal pass written to a second shared tile, the vertical pass to global.
// Replicating raw rows commutes with the horizontal pass, and the horizontal pass is only evaluated
// at in-image columns, so this is exactly vfiltering(hfiltering(src)) with the reference's clamping.
// A half-width of 0 with tap 1.0 makes a pass the identity (used for the one-directional filters).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_filter_hv(Img<T> src, Img<T> dst, Taps<T> th, Taps<T> tv) {
    constexpr int TX = 64, TY = 16, MAXH = kMaxHalf;
    __shared__ T raw[(TY + 2 * MAXH) * (TX + 2 * MAXH)];
    __shared__ T hs[(TY + 2 * MAXH) * TX];
    __shared__ T tap_h[2 * MAXH + 1], tap_v[2 * MAXH + 1];   // dynamic tap index: keep them out of local memory
    if (threadIdx.x < 2 * MAXH + 1) {
        tap_h[threadIdx.x] = th.v[threadIdx.x];
        tap_v[threadIdx.x] = tv.v[threadIdx.x];
    }
    const int fh = th.half, fv = tv.half;
    const int RWd = TX + 2 * fh, RHt = TY + 2 * fv;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, k = blockIdx.z;
    const int W = src.w, H = src.h;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T* sp = src.ch(k);
    for (int ry = warp; ry < RHt; ry += 8) {
        const T* row = sp + (size_t)clampi(y0 - fv + ry, H) * src.pitch;
        for (int rx = lane; rx < RWd; rx += 32) raw[ry * RWd + rx] = row[clampi(x0 - fh + rx, W)];
    }
    __syncthreads();
    for (int ry = warp; ry < RHt; ry += 8) {
        const T* r = raw + ry * RWd;
        for (int cx = lane; cx < TX; cx += 32) {
            T acc = 0;
            for (int l = 0; l <= 2 * fh; l++) acc += r[cx + l] * tap_h[l];
            hs[ry * TX + cx] = acc;
        }
    }
    __syncthreads();
    const int cx = threadIdx.x & (TX - 1), seg = threadIdx.x / TX;     // 4 row segments of 4 rows
    const int X = x0 + cx;
    if (X >= W) return;
#pragma unroll
    for (int j = 0; j < TY / 4; j++) {
        const int cy = seg * (TY / 4) + j, Y = y0 + cy;
        if (Y >= H) break;
        T acc = 0;
        for (int l = 0; l <= 2 * fv; l++) acc += hs[(cy + l) * TX + cx] * tap_v[l];
        dst.ch(k)[(size_t)Y * dst.pitch + X] = acc;
    }
}

// Same operation with COMPILE-TIME half-widths (the pyramid's Gaussians have half-widths 1..4, the
// 5x5 smoothing 2, the bicubic gradients 1/0): taps come straight from the constant bank, the
// horizontal pass computes four adjacent outputs from 16-byte shared-memory loads, the vertical pass
// slides a register window down four rows.  Same term order as the generic kernel (ascending tap
// index, accumulator starting at 0), so both produce identical bits; ~6x fewer instructions.
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const double* p, double (&v)[4]) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(double* p, const double (&v)[4]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}

template <typename T, int FH, int FV>
__global__ void __launch_bounds__(256) k_filter_hv_t(Img<T> src, Img<T> dst, const Taps<T> th, const Taps<T> tv) {
    constexpr int TX = 64, TY = 16;
    constexpr int NL = (4 + 2 * FH + 3) / 4;             // 4-element groups feeding four adjacent outputs
    constexpr int RW = 60 + 4 * NL;                      // raw row stride: last group of the last column quad stays in the row
    constexpr int RHt = TY + 2 * FV;
    static_assert(RW >= TX + 2 * FH && RW % 4 == 0, "raw row holds the halo and keeps 16-byte alignment");
    __shared__ __align__(16) T raw[RHt * RW];
    __shared__ __align__(16) T hs[RHt * TX];
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, k = blockIdx.z;
    const int W = src.w, H = src.h;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T* sp = src.ch(k);
    for (int ry = warp; ry < RHt; ry += 8) {
        const T* row = sp + (size_t)clampi(y0 - FV + ry, H) * src.pitch;
        for (int rx = lane; rx < RW; rx += 32) raw[ry * RW + rx] = row[clampi(x0 - FH + rx, W)];
    }
    __syncthreads();
    for (int task = threadIdx.x; task < RHt * (TX / 4); task += 256) {
        const int ry = task / (TX / 4), cq = task % (TX / 4);
        T win[4 * NL];
#pragma unroll
        for (int g = 0; g < NL; g++) {
            T q[4];
            ld4(&raw[ry * RW + 4 * cq + 4 * g], q);
#pragma unroll
            for (int e = 0; e < 4; e++) win[4 * g + e] = q[e];
        }
        T out[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            T acc = 0;
#pragma unroll
            for (int l = 0; l <= 2 * FH; l++) acc += win[j + l] * th.v[l];
            out[j] = acc;
        }
        st4(&hs[ry * TX + 4 * cq], out);
    }
    __syncthreads();
    const int cx = threadIdx.x & (TX - 1), seg = threadIdx.x / TX;     // 4 row segments of 4 rows
    const int X = x0 + cx;
    if (X >= W) return;
    T win[4 + 2 * FV];
#pragma unroll
    for (int i = 0; i < 4 + 2 * FV; i++) win[i] = hs[(seg * 4 + i) * TX + cx];
    T* dp = dst.ch(k) + X;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int Y = y0 + seg * 4 + j;
        T acc = 0;
#pragma unroll
        for (int l = 0; l <= 2 * FV; l++) acc += win[j + l] * tv.v[l];
        if (Y < H) dp[(size_t)Y * dst.pitch] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// Bilinear sampler (S/ImageProcessing.h:138-157).  Coordinates stay in double in BOTH modes
// (x = j + u at j ~ 3840 has only 2.4e-4 px resolution in FP32): the integer
// part comes from C truncation toward zero, the fraction is clamped to [0,1], taps are
// index-clamped and visited in the order (m,n) = (0,0),(0,1),(1,0),(1,1), m being the x offset.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Bilin {
    int x0, x1, y0, y1;
    T w00, w01, w10, w11;  // w[m][n]
    __device__ __forceinline__ Bilin(double x, double y, int w, int h) {
        int xi = (int)x, yi = (int)y;
        double fx = x - xi, fy = y - yi;
        fx = fmax(fmin(fx, 1.0), 0.0);
        fy = fmax(fmin(fy, 1.0), 0.0);
        x0 = clampi(xi, w); x1 = clampi(xi + 1, w);
        y0 = clampi(yi, h); y1 = clampi(yi + 1, h);
        T ax0 = (T)fabs(1 - fx), ax1 = (T)fabs(0 - fx);   // |1-m-dx|, m = 0,1
        T ay0 = (T)fabs(1 - fy), ay1 = (T)fabs(0 - fy);
        w00 = ax0 * ay0; w01 = ax0 * ay1; w10 = ax1 * ay0; w11 = ax1 * ay1;
    }
    __device__ __forceinline__ T sample(const T* __restrict__ p, int pitch) const {
        T acc = 0;
        acc += p[(size_t)y0 * pitch + x0] * w00;
        acc += p[(size_t)y1 * pitch + x0] * w01;
        acc += p[(size_t)y0 * pitch + x1] * w10;
        acc += p[(size_t)y1 * pitch + x1] * w11;
        return acc;
    }
};

// Bilinear resize (S/ImageProcessing.h:214-253): source coordinate (j+1)/r - 1 with per-axis
// ratios rx, ry; `scale` folds the caller's Multiplywith (flow upsampling, S/OpticalFlow.cpp:809-812).
template <typename T>
__global__ void k_resize(Img<T> src, Img<T> dst, double rx, double ry, T scale, int apply_scale) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dst.w) return;
    double sx = (double)(x + 1) / rx - 1, sy = (double)(y + 1) / ry - 1;
    Bilin<T> b(sx, sy, src.w, src.h);
    for (int k = 0; k < src.c; k++) {
        T v = b.sample(src.ch(k), src.pitch);
        if (apply_scale) v *= scale;
        dst.ch(k)[(size_t)y * dst.pitch + x] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// im2feature (S/OpticalFlow.cpp:911-961): RGB -> [gray, d/dx gray, d/dy gray, G-R, G-B];
// gray -> [I, Ix, Iy].  Luma (S/Image.h:1461-1480) is evaluated left to right; the derivative is the
// 5-tap [1,-8,0,8,-1]/12 correlation with replicate borders (S/Image.h:987-993, 1030-1036), computed
// here by re-evaluating the luma at the eight neighbours (same arithmetic, no intermediate plane).
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T luma_at(const Img<T>& im, int x, int y, int swap) {
    size_t o = (size_t)y * im.pitch + x;
    if (im.c == 1) return im.ch(0)[o];
    T r = im.ch(0)[o], g = im.ch(1)[o], b = im.ch(2)[o];
    return swap ? r * (T).114 + g * (T).587 + b * (T).299 : r * (T).299 + g * (T).587 + b * (T).114;
}

template <typename T>
__global__ void k_im2feature(Img<T> im, Img<T> feat, Taps<T> d5, int swap) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= im.w) return;
    size_t o = (size_t)y * feat.pitch + x;
    T gx = 0, gy = 0;
    for (int l = -2; l <= 2; l++) gx += luma_at(im, clampi(x + l, im.w), y, swap) * d5.v[l + 2];
    for (int l = -2; l <= 2; l++) gy += luma_at(im, x, clampi(y + l, im.h), swap) * d5.v[l + 2];
    feat.ch(0)[o] = luma_at(im, x, y, swap);
    feat.ch(1)[o] = gx;
    feat.ch(2)[o] = gy;
    if (im.c == 3) {
        size_t oi = (size_t)y * im.pitch + x;
        T r = im.ch(0)[oi], g = im.ch(1)[oi], b = im.ch(2)[oi];
        feat.ch(3)[o] = g - r;
        feat.ch(4)[o] = g - b;
    }
}

// ---------------------------------------------------------------------------------------------
// Bilinear warp with Im1 fallback outside the image (S/ImageProcessing.h:483-503), optionally
// preceded by the flow update u += du, v += dv of S/OpticalFlow.cpp:513-514 (du == nullptr: none).
// ---------------------------------------------------------------------------------------------
// Sampling position of pixel j displaced by flow value f, split into integer part and fraction
// exactly as the reference's double arithmetic would (x = j + f; xx = (int)x; dx = x - xx for the
// in-image case x >= 0, where truncation is floor).  The FP32 overload needs no FP64 instruction:
// floor(j + f) = j + floor(f) and f - floor(f) is exact in single precision, so the tap indices, the
// weights and the out-of-image decision are identical to evaluating j + (double)f.
struct SamplePos {
    int i;        // integer part
    bool out;     // x < 0 || x > n-1
};
__device__ __forceinline__ SamplePos sample_pos(int j, double f, int n, double& frac) {
    double x = (double)j + f;
    SamplePos p;
    p.out = x < 0 || x > n - 1;
    p.i = (int)x;
    frac = fmax(fmin(x - p.i, 1.0), 0.0);
    return p;
}
__device__ __forceinline__ SamplePos sample_pos(int j, float f, int n, float& frac) {
    float fl = floorf(f);
    frac = f - fl;
    SamplePos p;
    p.i = j + (int)fl;
    p.out = p.i < 0 || p.i > n - 1 || (p.i == n - 1 && frac > 0.f);
    return p;
}

// One thread handles kWarpPix pixels 32 columns apart (2 measured best, with 1/3/4 within 10 %): the flow loads of all of them are issued
// first, then per channel all 4*kWarpPix gathers, so each thread keeps 16+ independent loads in
// flight (the kernel is latency bound otherwise: two dependent memory round trips per pixel).
#ifndef PF_WARP_PIX
#define PF_WARP_PIX 2
#endif
constexpr int kWarpPix = PF_WARP_PIX;
#ifndef PF_WARP_ROWS
#define PF_WARP_ROWS 1
#endif
// launch grid of k_update_warp (128 threads)
inline dim3 warp_grid(int w, int h) {
#if PF_WARP_ROWS
    return dim3((unsigned)((w + 32 * kWarpPix - 1) / (32 * kWarpPix)), (unsigned)((h + 3) / 4));
#else
    return dim3((unsigned)((w + 128 * kWarpPix - 1) / (128 * kWarpPix)), (unsigned)h);
#endif
}

template <typename T>
__global__ void __launch_bounds__(128) k_update_warp(Img<T> im1, Img<T> im2, Img<T> warp, T* __restrict__ u,
                              T* __restrict__ v, const T* __restrict__ du,
                              const T* __restrict__ dv, int fpitch, int warp_lo = 0, int warp_hi = 0x7fffffff) {
    // rows outside [warp_lo, warp_hi) only get their flow update (row-band split over several GPUs: a device warps the
    // rows its own band's assembly reads; the flow itself stays complete everywhere, it accumulates across iterations)
    // the four warps of a CTA take the same columns of four consecutive rows: their bilinear taps
    // share image rows (row y+1 of one warp is row y of the next), which L1 then serves
    pdl_trigger();
    pdl_wait();
    const int W = im1.w, H = im1.h;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#if PF_WARP_ROWS
    const int y = blockIdx.y * 4 + wrp;
    if (y >= H) return;
    const int xb = blockIdx.x * (32 * kWarpPix) + lane;   // first pixel of this thread
#else
    const int y = blockIdx.y;
    const int xb = (blockIdx.x * (blockDim.x >> 5) + wrp) * (32 * kWarpPix) + lane;   // first pixel of this thread
#endif
    T uu[kWarpPix], vv[kWarpPix];
#pragma unroll
    for (int i = 0; i < kWarpPix; i++) {
        const int x = xb + 32 * i;
        uu[i] = 0; vv[i] = 0;
        if (x < W) {
            const size_t of = (size_t)y * fpitch + x;
            uu[i] = u[of]; vv[i] = v[of];
            if (du) { uu[i] += du[of]; vv[i] += dv[of]; }
        }
    }
    // element offsets inside one plane fit 32 bits (a 3840x2160 plane is 8.3 M elements)
    int o00[kWarpPix], o01[kWarpPix], o10[kWarpPix], o11[kWarpPix];
    T w00[kWarpPix], w01[kWarpPix], w10[kWarpPix], w11[kWarpPix];
    bool inside[kWarpPix];
#pragma unroll
    for (int i = 0; i < kWarpPix; i++) {
        const int x = xb + 32 * i;
        inside[i] = false;
        o00[i] = o01[i] = o10[i] = o11[i] = 0;
        w00[i] = w01[i] = w10[i] = w11[i] = 0;
        if (x >= W) continue;
        if (du) {
            const size_t of = (size_t)y * fpitch + x;
            u[of] = uu[i];
            v[of] = vv[i];
        }
        if (y < warp_lo || y >= warp_hi) continue;   // (warp-uniform: a warp works on one row)
        T fx, fy;
        const SamplePos px = sample_pos(x, uu[i], W, fx), py = sample_pos(y, vv[i], H, fy);
        if (px.out || py.out) continue;
        inside[i] = true;
        const int x0 = clampi(px.i, W), x1 = clampi(px.i + 1, W), y0 = clampi(py.i, H), y1 = clampi(py.i + 1, H);
        const T ax0 = fabs((T)1 - fx), ax1 = fabs((T)0 - fx), ay0 = fabs((T)1 - fy), ay1 = fabs((T)0 - fy);
        w00[i] = ax0 * ay0; w01[i] = ax0 * ay1; w10[i] = ax1 * ay0; w11[i] = ax1 * ay1;
        o00[i] = y0 * im2.pitch + x0; o01[i] = y1 * im2.pitch + x0;
        o10[i] = y0 * im2.pitch + x1; o11[i] = y1 * im2.pitch + x1;
    }
    if (y < warp_lo || y >= warp_hi) return;
    const int orow1 = y * im1.pitch, orow_w = y * warp.pitch;
    // the gathers of channel k+1 are issued before channel k is reduced and stored (two register
    // buffers), so a thread has up to 8*kWarpPix loads in flight and the memory round trips of
    // consecutive channels overlap instead of adding up
    auto gather = [&](int k, T (&a)[4][kWarpPix]) {
        const T* p = im2.ch(k);
        const T* q = im1.ch(k);
#pragma unroll
        for (int i = 0; i < kWarpPix; i++) {
            const int x = xb + 32 * i;
            if (inside[i]) { a[0][i] = p[o00[i]]; a[1][i] = p[o01[i]]; a[2][i] = p[o10[i]]; a[3][i] = p[o11[i]]; }
            else if (x < W) { a[0][i] = q[orow1 + x]; a[1][i] = a[2][i] = a[3][i] = 0; }
            else { a[0][i] = a[1][i] = a[2][i] = a[3][i] = 0; }
        }
    };
    auto reduce = [&](int k, const T (&a)[4][kWarpPix]) {
        T* o = warp.ch(k) + orow_w;
#pragma unroll
        for (int i = 0; i < kWarpPix; i++) {
            const int x = xb + 32 * i;
            if (x >= W) continue;
            T acc;
            if (inside[i]) {
                acc = 0;
                acc += a[0][i] * w00[i];
                acc += a[1][i] * w01[i];
                acc += a[2][i] * w10[i];
                acc += a[3][i] * w11[i];
            } else {
                acc = a[0][i];          // Im1 fallback outside the image
            }
            o[x] = acc;
        }
    };
    const int C = im1.c;
    T A[4][kWarpPix], B[4][kWarpPix];
    gather(0, A);
    int k = 0;
    for (; k + 1 < C; k += 2) {
        gather(k + 1, B);
        reduce(k, A);
        if (k + 2 < C) gather(k + 2, A);
        reduce(k + 1, B);
    }
    if (k < C) reduce(k, A);
}

// ---------------------------------------------------------------------------------------------
// getDxs pieces (S/OpticalFlow.cpp:80-122): blend = Im1s*0.4 + Im2s*0.6 ; imdt = Im2s - Im1s
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_blend_dt(Img<T> s1, Img<T> s2, Img<T> blend, Img<T> dt) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, k = blockIdx.z;
    if (x >= s1.w) return;
    size_t o = (size_t)y * s1.pitch + x;
    T a = s1.ch(k)[o], b = s2.ch(k)[o];
    T t = a * (T)0.4;
    blend.ch(k)[o] = t + b * (T)0.6;
    dt.ch(k)[o] = b - a;
}

// ---------------------------------------------------------------------------------------------
// phi (S/OpticalFlow.cpp:295-331): forward differences of uu = u + du (last column / last row
// zero, S/Image.h:969-986, 1013-1029), phi = 0.5 / sqrt(ux^2 + uy^2 + vx^2 + vy^2 + eps)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_phi(const T* __restrict__ u, const T* __restrict__ v, const T* __restrict__ du,
                      const T* __restrict__ dv, T* __restrict__ phi, int w, int h, int pitch,
                      T eps) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    size_t o = (size_t)y * pitch + x;
    auto U = [&](size_t i) { return du ? u[i] + du[i] : u[i]; };
    auto V = [&](size_t i) { return dv ? v[i] + dv[i] : v[i]; };
    T u0 = U(o), v0 = V(o);
    T ux = 0, uy = 0, vx = 0, vy = 0;
    if (x < w - 1) { ux = U(o + 1) - u0; vx = V(o + 1) - v0; }
    if (y < h - 1) { uy = U(o + pitch) - u0; vy = V(o + pitch) - v0; }
    T t = ux * ux + uy * uy + vx * vx + vy * vy;
    phi[o] = (T)0.5 / sqrt(t + eps);
}

// ---------------------------------------------------------------------------------------------
// Linear-system assembly (S/OpticalFlow.cpp:333-448 + the loop-invariant part of :463-504):
//   psi_k  = 1 / (2 sqrt((It + Ix du + Iy dv)^2 + eps))        unless lap[k] < 1e-20  (:399-402)
//   dxy    = mean_k (psi Ix) Iy, dx2, dy2, dtdx, dtdy likewise  (:414-427, S/Image.h:1536-1543)
//   bu     = -dtdx - alpha * L(u),  bv = -dtdy - alpha * L(v)   (:444-448) with the fork's fused
//            Laplacian that drops the inflow of the last column / last row (:641-690, quirk F3)
//   iu     = omega / (dx2 + alpha*0.05 + alpha * sum_nbr phi)   (:496-501), iv likewise with dy2
// iu/iv are exactly the `omega/(...)` sub-expression the reference evaluates first in :501/:504,
// so precomputing them changes no rounding.  dx2/dy2 are written only when requested (stage tests).
// ---------------------------------------------------------------------------------------------
template <typename T>
struct AssembleArgs {
    Img<T> imdx, imdy, imdt;
    const T *u, *v, *du, *dv, *phi;
    const double* lap;  // per-channel Laplacian noise scale (device), may be nullptr
    const double* gm;   // Gaussian-mixture parameters (GmState layout, device): noiseModel == GMixture, else nullptr
    T *dxy, *iu, *iv, *bu, *bv, *dx2, *dy2;
    int w, h, pitch;
    T alpha, omega, eps;
};

// Gaussian-mixture noise model (S/NoiseModel.h:16-24): per channel alpha, sigma, beta and the two squares, in that
// order, 16 doubles each.  The mixture arithmetic stays in double in every mode (alternative branch, not a hot path).
constexpr int kGmStride = 16;
enum { GM_ALPHA = 0, GM_SIGMA = 1, GM_BETA = 2, GM_SIGMA2 = 3, GM_BETA2 = 4, GM_FIELDS = 5 };
// Quirk: the reference's Gaussian() is compiled with PI = 3.1415927 -- S/NoiseModel.h:10-12 defines its own value
// only #ifndef PI, and S/Image.h:14 has already included S/Stochastic.h:18-20.
#define PF_GM_PI 3.1415927
__device__ __forceinline__ double gm_gaussian(const double* gm, double x, int i, int k) {   // S/NoiseModel.h:116-122
    if (i == 0) return exp(-x / (2 * gm[GM_SIGMA2 * kGmStride + k])) / (2 * PF_GM_PI * gm[GM_SIGMA * kGmStride + k]);
    return exp(-x / (2 * gm[GM_BETA2 * kGmStride + k])) / (2 * PF_GM_PI * gm[GM_BETA * kGmStride + k]);
}

template <typename T>
__device__ __forceinline__ T laplacian_at(const T* __restrict__ in, const T* __restrict__ phi,
                                          size_t o, int x, int y, int w, int h, int pitch) {
    T out = 0;
    if (x < w - 1) {
        out -= (in[o + 1] - in[o]) * phi[o];
        if (x > 0) out += (in[o] - in[o - 1]) * phi[o - 1];
    }
    if (y < h - 1) {
        out -= (in[o + pitch] - in[o]) * phi[o];
        if (y > 0) out += (in[o] - in[o - pitch]) * phi[o - pitch];
    }
    return out;
}

template <typename T>
__global__ void k_assemble(AssembleArgs<T> a) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.w) return;
    size_t o = (size_t)y * a.pitch + x;
    T du = a.du ? a.du[o] : (T)0, dv = a.dv ? a.dv[o] : (T)0;
    T sxy = 0, sx2 = 0, sy2 = 0, stx = 0, sty = 0;
    int C = a.imdx.c;
    for (int k = 0; k < C; k++) {
        size_t oc = (size_t)y * a.imdx.pitch + x;
        T ix = a.imdx.ch(k)[oc], iy = a.imdy.ch(k)[oc], it = a.imdt.ch(k)[oc];
        T psi = 0;
        if (a.gm) {   // S/OpticalFlow.cpp:389-396
            double t = (double)it + (double)ix * (double)du + (double)iy * (double)dv;
            t *= t;
            double prob1 = gm_gaussian(a.gm, t, 0, k) * a.gm[GM_ALPHA * kGmStride + k];
            double prob2 = gm_gaussian(a.gm, t, 1, k) * (1 - a.gm[GM_ALPHA * kGmStride + k]);
            double prob11 = prob1 / (2 * a.gm[GM_SIGMA2 * kGmStride + k]);
            double prob22 = prob2 / (2 * a.gm[GM_BETA2 * kGmStride + k]);
            psi = (T)((prob11 + prob22) / (prob1 + prob2));
        } else if (!(a.lap && a.lap[k] < 1e-20)) {
            T t = it + ix * du + iy * dv;
            t *= t;
            psi = (T)1 / ((T)2 * sqrt(t + a.eps));
        }
        T px = psi * ix, py = psi * iy;
        sxy += px * iy;
        sx2 += px * ix;
        sy2 += py * iy;
        stx += px * it;
        sty += py * it;
    }
    if (C > 1) {
        T n = (T)C;
        sxy /= n; sx2 /= n; sy2 /= n; stx /= n; sty /= n;
    }
    T lu = laplacian_at(a.u, a.phi, o, x, y, a.w, a.h, a.pitch);
    T lv = laplacian_at(a.v, a.phi, o, x, y, a.w, a.h, a.pitch);
    T cf = 0;
    if (x > 0) cf += a.phi[o - 1];
    if (x < a.w - 1) cf += a.phi[o];
    if (y > 0) cf += a.phi[o - a.pitch];
    if (y < a.h - 1) cf += a.phi[o];
    cf *= a.alpha;
    T reg = a.alpha * (T)0.05;
    a.dxy[o] = sxy;
    a.iu[o] = a.omega / (sx2 + reg + cf);
    a.iv[o] = a.omega / (sy2 + reg + cf);
    a.bu[o] = -stx - a.alpha * lu;
    a.bv[o] = -sty - a.alpha * lv;
    if (a.dx2) { a.dx2[o] = sx2; a.dy2[o] = sy2; }
}

// ---------------------------------------------------------------------------------------------
// estLaplacianNoise (S/OpticalFlow.cpp:594-639): per-channel mean of d = |Im1 - warpIm2| over
// 0 < d < 1e6, 0.001 if none.  Its only consumer is the `< 1e-20` guard above, so the summation
// order (block tree + atomics here) is immaterial.  acc = [sum_0..sum_C-1, cnt_0..cnt_C-1].
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_noise_accum(Img<T> a, Img<T> b, double* __restrict__ acc) {
    int k = blockIdx.z;
    double s = 0, n = 0;
    for (int y = blockIdx.y; y < a.h; y += gridDim.y)
        for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < a.w; x += gridDim.x * blockDim.x) {
            size_t o = (size_t)y * a.pitch + x;
            double d = fabs((double)a.ch(k)[o] - (double)b.ch(k)[o]);
            if (d > 0 && d < 1000000) { s += d; n += 1; }
        }
    for (int off = 16; off; off >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, off);
        n += __shfl_down_sync(0xffffffffu, n, off);
    }
    if ((threadIdx.x & 31) == 0 && n > 0) {
        atomicAdd(acc + k, s);
        atomicAdd(acc + a.c + k, n);
    }
}

static __global__ void k_noise_final(double* acc, double* lap, int c) {
    int k = threadIdx.x;
    if (k >= c) return;
    lap[k] = acc[c + k] == 0 ? 0.001 : acc[k] / acc[c + k];
    acc[k] = 0;
    acc[c + k] = 0;
}

// ---------------------------------------------------------------------------------------------
// estGaussianMixture (S/OpticalFlow.cpp:539-591): three EM steps of the two-component mixture on
// t = (Im1 - warpIm2)^2 per channel.  One step = k_gm_accum (responsibilities with the current parameters and the
// four sums per channel: total1, total2, sum w1 t, sum w2 t -- the reference's two loops read the same weights) +
// k_gm_update (the M step; it starts from para.reset(), :564, so the sums of sigma / beta start at 0.05 / 0.5).
// acc = [total1[16], total2[16], s1[16], s2[16]].  Summation order differs from the reference's (rounding level).
// ---------------------------------------------------------------------------------------------
static __global__ void k_gm_reset(double* gm) {   // GaussianMixture::reset, S/NoiseModel.h:98-108
    int k = threadIdx.x;
    if (k >= kGmStride) return;
    gm[GM_ALPHA * kGmStride + k] = 0.95;
    gm[GM_SIGMA * kGmStride + k] = 0.05;
    gm[GM_BETA * kGmStride + k] = 0.5;
    gm[GM_SIGMA2 * kGmStride + k] = 0.05 * 0.05;
    gm[GM_BETA2 * kGmStride + k] = 0.5 * 0.5;
}

template <typename T>
__global__ void k_gm_accum(Img<T> a, Img<T> b, const double* __restrict__ gm, double* __restrict__ acc) {
    int k = blockIdx.z;
    double t1 = 0, t2 = 0, s1 = 0, s2 = 0;
    const double alpha = gm[GM_ALPHA * kGmStride + k];
    for (int y = blockIdx.y; y < a.h; y += gridDim.y)
        for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < a.w; x += gridDim.x * blockDim.x) {
            size_t o = (size_t)y * a.pitch + x;
            double t = (double)a.ch(k)[o] - (double)b.ch(k)[o];
            t *= t;
            double w1 = gm_gaussian(gm, t, 0, k) * alpha;
            double w2 = gm_gaussian(gm, t, 1, k) * (1 - alpha);
            double n = w1 + w2;
            w1 /= n;
            w2 /= n;
            t1 += w1; t2 += w2;
            s1 += w1 * t; s2 += w2 * t;
        }
    for (int off = 16; off; off >>= 1) {
        t1 += __shfl_down_sync(0xffffffffu, t1, off);
        t2 += __shfl_down_sync(0xffffffffu, t2, off);
        s1 += __shfl_down_sync(0xffffffffu, s1, off);
        s2 += __shfl_down_sync(0xffffffffu, s2, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(acc + 0 * kGmStride + k, t1);
        atomicAdd(acc + 1 * kGmStride + k, t2);
        atomicAdd(acc + 2 * kGmStride + k, s1);
        atomicAdd(acc + 3 * kGmStride + k, s2);
    }
}

static __global__ void k_gm_update(double* gm, double* acc, int c) {
    int k = threadIdx.x;
    if (k >= c) return;
    const double prior = 0.9;   // default argument, S/OpticalFlow.h:44
    double t1 = acc[k], t2 = acc[kGmStride + k];
    double sg = 0.05 + acc[2 * kGmStride + k], be = 0.5 + acc[3 * kGmStride + k];
    double alpha = t1 / (t1 + t2) * (1 - prior) + 0.95 * prior;
    sg = sqrt(sg / t1);
    be = sqrt(be / t2) * (1 - prior) + 0.3 * prior;
    gm[GM_ALPHA * kGmStride + k] = alpha;
    gm[GM_SIGMA * kGmStride + k] = sg;
    gm[GM_BETA * kGmStride + k] = be;
    gm[GM_SIGMA2 * kGmStride + k] = sg * sg;
    gm[GM_BETA2 * kGmStride + k] = be * be;
    for (int j = 0; j < 4; j++) acc[j * kGmStride + k] = 0;
}

// ---------------------------------------------------------------------------------------------
// Final output warp (S/Image.h:2624-2701 with coefficients :2497-2530, then threshold :2031-2045):
// Hermite-bicubic from I, Ix, Iy, Ixy at the four clamped corners, Im1 fallback outside the image,
// clamp to [0,1], written straight into the HWC float64 output buffer.
// The 16 coefficients are integer combinations of the 16 corner samples; sources are indexed
// 4*quantity + corner (quantity I,Ix,Iy,Ixy; corner 00,10,01,11, x digit first) and summed in the
// listed order, which is the reference's left-to-right evaluation order.
// ---------------------------------------------------------------------------------------------
struct BicubicTable {
    signed char n[16];
    signed char src[16][16];
    signed char wt[16][16];
};

// Destination of the bicubic warp: the boundary's HWC float64 buffer (final im2W, always clamped:
// S/OpticalFlow.cpp:841-842) or a planar image of the solver (the Bicubic inner warp of
// S/OpticalFlow.cpp:517-521 with threshold(), and the level-start warp of :814-815 without it).
template <typename T>
struct BicubicOut {
    double* hwc;      // non-null: HWC float64, clamped
    Img<T> planar;    // used when hwc == nullptr
    int clamp;        // planar output only
    __device__ __forceinline__ void put(int x, int y, int w, int C, int k, T val) const {
        if (hwc) {
            hwc[((size_t)y * w + x) * C + k] = (double)min(max(val, (T)0), (T)1);
        } else {
            if (clamp) val = min(max(val, (T)0), (T)1);
            planar.ch(k)[(size_t)y * planar.pitch + x] = val;
        }
    }
};

template <typename T>
__global__ void k_bicubic_warp(Img<T> ref, Img<T> im, Img<T> ix, Img<T> iy, Img<T> ixy,
                               const T* __restrict__ u, const T* __restrict__ v, int fpitch,
                               const BicubicTable* __restrict__ tab, BicubicOut<T> out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= im.w) return;
    int w = im.w, h = im.h, C = im.c;
    size_t of = (size_t)y * fpitch + x;
    double sx = (double)x + (double)u[of], sy = (double)y + (double)v[of];
    if (sx < 0 || sx > w - 1 || sy < 0 || sy > h - 1) {
        for (int k = 0; k < C; k++) out.put(x, y, w, C, k, ref.ch(k)[(size_t)y * ref.pitch + x]);
        return;
    }
    int x0 = clampi((int)sx, w), x1 = clampi((int)sx + 1, w);
    int y0 = clampi((int)sy, h), y1 = clampi((int)sy + 1, h);
    T dx = (T)(sx - x0), dy = (T)(sy - y0);
    T dx2 = dx * dx, dy2 = dy * dy, dx3 = dx * dx2, dy3 = dy * dy2;
    size_t corner[4] = {(size_t)y0 * im.pitch + x0, (size_t)y0 * im.pitch + x1,
                        (size_t)y1 * im.pitch + x0, (size_t)y1 * im.pitch + x1};
    if (sizeof(T) == 4) {
        // FP32 fast mode: the same bicubic Hermite patch in separable basis form,
        //   f = sum_{a,b in {0,1}} I_ab hv_a(dx) hv_b(dy) + Ix_ab hd_a(dx) hv_b(dy) + Iy_ab hv_a(dx) hd_b(dy) + Ixy_ab hd_a(dx) hd_b(dy)
        // with hv_0 = 2t^3-3t^2+1, hv_1 = 3t^2-2t^3, hd_0 = t^3-2t^2+t, hd_1 = t^3-t^2: algebraically the
        // polynomial the 16 coefficients of S/Image.h:2497-2530 describe (16 FMAs per channel instead of ~150).
        const T vx0 = 2 * dx3 - 3 * dx2 + 1, vx1 = 3 * dx2 - 2 * dx3, gx0 = dx3 - 2 * dx2 + dx, gx1 = dx3 - dx2;
        const T vy0 = 2 * dy3 - 3 * dy2 + 1, vy1 = 3 * dy2 - 2 * dy3, gy0 = dy3 - 2 * dy2 + dy, gy1 = dy3 - dy2;
        // corner index: 0 = (x0,y0), 1 = (x1,y0), 2 = (x0,y1), 3 = (x1,y1)
        const T bp[4] = {vx0 * vy0, vx1 * vy0, vx0 * vy1, vx1 * vy1};
        const T bx[4] = {gx0 * vy0, gx1 * vy0, gx0 * vy1, gx1 * vy1};
        const T by[4] = {vx0 * gy0, vx1 * gy0, vx0 * gy1, vx1 * gy1};
        const T bz[4] = {gx0 * gy0, gx1 * gy0, gx0 * gy1, gx1 * gy1};
        for (int k = 0; k < C; k++) {
            const T *pi = im.ch(k), *px = ix.ch(k), *py = iy.ch(k), *pz = ixy.ch(k);
            T val = 0;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                val += pi[corner[t]] * bp[t];
                val += px[corner[t]] * bx[t];
                val += py[corner[t]] * by[t];
                val += pz[corner[t]] * bz[t];
            }
            out.put(x, y, w, C, k, val);
        }
        return;
    }
    for (int k = 0; k < C; k++) {
        T s[16];
        const T* q[4] = {im.ch(k), ix.ch(k), iy.ch(k), ixy.ch(k)};
#pragma unroll
        for (int t = 0; t < 16; t++) s[t] = q[t >> 2][corner[t & 3]];
        T a[16];
#pragma unroll
        for (int r = 0; r < 16; r++) {
            T acc = (T)tab->wt[r][0] * s[tab->src[r][0]];
            for (int t = 1; t < tab->n[r]; t++) acc += (T)tab->wt[r][t] * s[tab->src[r][t]];
            a[r] = acc;
        }
        // a[4*i + j] multiplies dx^i dy^j; evaluation order of S/Image.h:2688-2691
        T val = a[0] + a[1] * dy + a[2] * dy2 + a[3] * dy3 +
                a[4] * dx + a[5] * dx * dy + a[6] * dx * dy2 + a[7] * dx * dy3 +
                a[8] * dx2 + a[9] * dx2 * dy + a[10] * dx2 * dy2 + a[11] * dx2 * dy3 +
                a[12] * dx3 + a[13] * dx3 * dy + a[14] * dx3 * dy2 + a[15] * dx3 * dy3;
        out.put(x, y, w, C, k, val);
    }
}

// u += du, v += dv (S/OpticalFlow.cpp:513-514) on its own: the Bicubic inner warp does not go through k_update_warp
template <typename T>
__global__ void k_add_flow(T* __restrict__ u, T* __restrict__ v, const T* __restrict__ du, const T* __restrict__ dv,
                           int w, int pitch) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    size_t o = (size_t)y * pitch + x;
    u[o] += du[o];
    v[o] += dv[o];
}

// Host-side definition of the table (shared by both instantiations).
inline BicubicTable make_bicubic_table() {
    enum { P00, P10, P01, P11, X00, X10, X01, X11, Y00, Y10, Y01, Y11, Z00, Z10, Z01, Z11 };
    BicubicTable t;
    memset(&t, 0, sizeof(t));
    auto set = [&](int r, std::initializer_list<int> src, std::initializer_list<int> wt) {
        t.n[r] = (signed char)src.size();
        int i = 0;
        for (int s : src) t.src[r][i++] = (signed char)s;
        i = 0;
        for (int w : wt) t.wt[r][i++] = (signed char)w;
    };
    const std::initializer_list<int> all = {P00, P10, P01, P11, X00, X10, X01, X11,
                                            Y00, Y10, Y01, Y11, Z00, Z10, Z01, Z11};
    // row index = 4*i + j for a[i][j]  (S/Image.h:2499-2529)
    set(0, {P00}, {1});
    set(1, {Y00}, {1});
    set(2, {P00, P01, Y00, Y01}, {-3, 3, -2, -1});
    set(3, {P00, P01, Y00, Y01}, {2, -2, 1, 1});
    set(4, {X00}, {1});
    set(5, {Z00}, {1});
    set(6, {X00, X01, Z00, Z01}, {-3, 3, -2, -1});
    set(7, {X00, X01, Z00, Z01}, {2, -2, 1, 1});
    set(8, {P00, P10, X00, X10}, {-3, 3, -2, -1});
    set(9, {Y00, Y10, Z00, Z10}, {-3, 3, -2, -1});
    set(10, all, {9, -9, -9, 9, 6, 3, -6, -3, 6, -6, 3, -3, 4, 2, 2, 1});
    set(11, all, {-6, 6, 6, -6, -4, -2, 4, 2, -3, 3, -3, 3, -2, -1, -2, -1});
    set(12, {P00, P10, X00, X10}, {2, -2, 1, 1});
    set(13, {Y00, Y10, Z00, Z10}, {2, -2, 1, 1});
    set(14, all, {-6, 6, 6, -6, -3, -3, 3, 3, -4, 4, -2, 2, -2, -2, -1, -1});
    set(15, all, {4, -4, -4, 4, 2, 2, -2, -2, 2, -2, 2, -2, 1, 1, 1, 1});
    return t;
}


// =============================================================================================
// Fused outer-iteration front end: getDxs + phi + linear-system assembly in ONE kernel.
//
// Replaces nine launches (2x smooth h/v of the warped image, blend/dt, 2x derivative, phi, assemble)
// and their ~90 words/pixel of plane traffic by a single pass that reads the warped features (with a
// 4-pixel halo), the per-level smoothed Im1 (2-pixel halo), u and v (1-pixel halo) and writes the six
// coefficient planes: ~21 words/pixel.  One CTA owns a TX x TY output tile and walks the feature
// channels; per channel the raw tile is staged in shared memory, smoothed horizontally then
// vertically ([.02,.11,.74,.11,.02], S/OpticalFlow.cpp:84-90), blended .4/.6 with the smoothed Im1
// (:91-93), differentiated with [1,-8,0,8,-1]/12 (:95-96) and folded into the five psi-weighted sums
// (:377-427) held in registers.  Every shared-memory read goes through the CLAMPED image coordinate,
// which is exactly the replicate-border rule of S/ImageProcessing.h:259-279/350-369, and every sum
// keeps the reference's term order, so the FP64 instantiation is bit-identical to the unfused path.
// =============================================================================================
template <typename T>
struct FusedArgs {
    Img<T> s1, wf;                 // smoothed Im1 features (per level), warped Im2 features
    const T *u, *v, *du, *dv;      // du/dv: current increments (nullptr = zero, first inner iteration)
    const double* lap;             // per-channel noise scale guard (nullptr = always on)
    T *phi, *dxy, *iu, *iv, *bu, *bv;
    int w, h, pitch;               // pitch of the scalar planes
    T alpha, omega, eps;
    Taps<T> g5, d5;
    int ty0 = 0;                   // first tile row of this launch (k_fused_tma; a row band of the level, multigpu.cuh)
    // k_fused_cp<..., WARP = true> only: the warp of S/OpticalFlow.cpp:513-516 is computed inside the assembly kernel from
    // the level's features f1, f2 and the flow u + wdu, v + wdv (wdu == nullptr: the flow is u, v as they are); the
    // updated flow of the tile's pixels is stored to uo, vo (u, v stay intact: other CTAs read their halos from them)
    Img<T> f1, f2;
    const T *wdu = nullptr, *wdv = nullptr;
    T *uo = nullptr, *vo = nullptr;
};

// psi = 1 / (2 sqrt(t + eps)).  FP64 keeps the reference's expression; FP32 uses the hardware
// reciprocal square root (2 ulp), far below the single-precision noise of the products it scales.
__device__ __forceinline__ double psi_of(double t, double eps) { return 1.0 / (2.0 * sqrt(t + eps)); }
// t + eps >= 1e-6 is always a normal number, so the flush-to-zero form (one MUFU.RSQ, no denormal
// pre/post-scaling) returns exactly what rsqrtf() would
__device__ __forceinline__ float rsqrt_normal(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float psi_of(float t, float eps) { return 0.5f * rsqrt_normal(t + eps); }
// the same split for phi = 0.5/sqrt(t+eps), the channel mean and omega/denominator
__device__ __forceinline__ double phi_of(double t, double eps) { return 0.5 / sqrt(t + eps); }
__device__ __forceinline__ float phi_of(float t, float eps) { return 0.5f * rsqrt_normal(t + eps); }
__device__ __forceinline__ double mean_of(double s, int c, double) { return s / (double)c; }
__device__ __forceinline__ float mean_of(float s, int, float inv_c) { return s * inv_c; }
__device__ __forceinline__ double ratio_of(double a, double b) { return a / b; }
__device__ __forceinline__ float ratio_of(float a, float b) { return __fdividef(a, b); }

// Thread layout (256 threads, 8 warps): tile rows are dealt to warps and columns to lanes when
// staging (coalesced, conflict free, no div/mod); for the vertical filters, the derivative stage
// and the final stage thread t owns column t%64 and the contiguous row segment t/64, so the
// vertical 5-tap windows slide through registers (1.4 shared loads per output instead of 5).
// INTERIOR tiles (tile + 4-pixel halo inside the image) skip every clamp and border test.
template <typename T, int TX, int TY, int SEG, bool INTERIOR>
__device__ __forceinline__ void fused_assemble_body(const FusedArgs<T>& a, T* sm_raw, T* sm_bl, T* sm_dt) {
    constexpr int NT = TX * SEG, NWARP = NT / 32;
    constexpr int PPT = TY / SEG;                     // centre pixels per thread (rows of its segment)
    constexpr int RW = TX + 8, RHt = TY + 8;          // raw tile   (halo 4)
    constexpr int HW = TX + 4;                        // h-smoothed (halo 2 in x, 4 in y)
    constexpr int BW = TX + 4, BH = TY + 4;           // blend tile (halo 2)
    constexpr int BSEG = (BH + SEG - 1) / SEG;        // blend rows per segment (last one may be short)
    constexpr int UW = TX + 2, UH = TY + 2;           // u/v tiles  (halo 1)
    constexpr int PW = TX + 1, PH = TY + 1;           // phi tile   (halo 1 left/up)
    static_assert(TX == 64, "thread layout assumes 64-wide tiles");
    static_assert(TY % SEG == 0, "tile height must split into SEG segments");
    T* raw = sm_raw;
    T* hs = sm_raw + RHt * RW;
    T* bl = sm_bl;

    const int W = a.w, H = a.h;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col = tid & (TX - 1), seg = tid / TX;
    const int C = a.wf.c;
    auto cx_ = [&](int X) { return INTERIOR ? X : clampi(X, W); };
    auto cy_ = [&](int Y) { return INTERIOR ? Y : clampi(Y, H); };

    // centre pixels of this thread: (x0+col, y0 + seg*PPT + k)
    const int PX = x0 + col;
    const bool col_ok = INTERIOR || PX < W;
    T sxy[PPT], sx2[PPT], sy2[PPT], stx[PPT], sty[PPT], cdu[PPT], cdv[PPT];
#pragma unroll
    for (int k = 0; k < PPT; k++) {
        sxy[k] = sx2[k] = sy2[k] = stx[k] = sty[k] = 0;
        cdu[k] = cdv[k] = 0;
        int PY = y0 + seg * PPT + k;
        if (a.du && col_ok && (INTERIOR || PY < H)) {
            cdu[k] = a.du[(size_t)PY * a.pitch + PX];
            cdv[k] = a.dv[(size_t)PY * a.pitch + PX];
        }
    }
    const T g0 = a.g5.v[0], g1 = a.g5.v[1], g2 = a.g5.v[2], g3 = a.g5.v[3], g4 = a.g5.v[4];
    const T d0 = a.d5.v[0], d1 = a.d5.v[1], d3 = a.d5.v[3], d4 = a.d5.v[4];   // centre tap is 0

    for (int c = 0; c < C; c++) {
        const T* wfc = a.wf.ch(c);
        const T* s1c = a.s1.ch(c);
        const bool active = !(a.lap && a.lap[c] < 1e-20);   // S/OpticalFlow.cpp:399-400
        // 1. stage the raw tile: warp per row, lanes across the row
        for (int ry = warp; ry < RHt; ry += NWARP) {
            const T* row = wfc + (size_t)cy_(y0 - 4 + ry) * a.wf.pitch;
            for (int rx = lane; rx < RW; rx += 32) raw[ry * RW + rx] = row[cx_(x0 - 4 + rx)];
        }
        __syncthreads();
        // 2. horizontal smoothing for columns x0-2 .. x0+TX+1 of every staged row
        for (int ry = warp; ry < RHt; ry += NWARP) {
            const T* r = raw + ry * RW;
            for (int hx = lane; hx < HW; hx += 32) {
                T acc = 0;
                if (INTERIOR) {
                    acc += r[hx] * g0; acc += r[hx + 1] * g1; acc += r[hx + 2] * g2;
                    acc += r[hx + 3] * g3; acc += r[hx + 4] * g4;
                } else {
                    int X = clampi(x0 - 2 + hx, W), o = x0 - 4;
                    acc += r[clampi(X - 2, W) - o] * g0; acc += r[clampi(X - 1, W) - o] * g1;
                    acc += r[X - o] * g2;
                    acc += r[clampi(X + 1, W) - o] * g3; acc += r[clampi(X + 2, W) - o] * g4;
                }
                hs[ry * HW + hx] = acc;
            }
        }
        __syncthreads();
        // 3. vertical smoothing down a column segment (sliding window), blend with the smoothed Im1,
        //    temporal difference at the centre.  Columns 64..67 are done by the first 16 threads.
        auto vsmooth = [&](int bx, int sg) {
            const int X = cx_(x0 - 2 + bx);
            const int by0 = sg * BSEG;
            if (by0 >= BH) return;
            if (INTERIOR) {
                const T* hcol = hs + bx;
                T w0 = hcol[(by0 + 0) * HW], w1 = hcol[(by0 + 1) * HW], w2 = hcol[(by0 + 2) * HW], w3 = hcol[(by0 + 3) * HW];
#pragma unroll
                for (int j = 0; j < BSEG; j++) {
                    const int by = by0 + j;
                    if (by >= BH) break;
                    T w4 = hcol[(by + 4) * HW];
                    T acc = 0;
                    acc += w0 * g0; acc += w1 * g1; acc += w2 * g2; acc += w3 * g3; acc += w4 * g4;
                    w0 = w1; w1 = w2; w2 = w3; w3 = w4;
                    T s1v = s1c[(size_t)(y0 - 2 + by) * a.s1.pitch + X];
                    T t = s1v * (T)0.4;
                    bl[by * BW + bx] = t + acc * (T)0.6;
                    int ccx = bx - 2, ccy = by - 2;
                    if (ccx >= 0 && ccx < TX && ccy >= 0 && ccy < TY) sm_dt[ccy * TX + ccx] = acc - s1v;
                }
            } else {
#pragma unroll
                for (int j = 0; j < BSEG; j++) {
                    const int by = by0 + j;
                    if (by >= BH) break;
                    const int Y = clampi(y0 - 2 + by, H), o = y0 - 4;
                    T acc = 0;
                    acc += hs[(clampi(Y - 2, H) - o) * HW + bx] * g0;
                    acc += hs[(clampi(Y - 1, H) - o) * HW + bx] * g1;
                    acc += hs[(Y - o) * HW + bx] * g2;
                    acc += hs[(clampi(Y + 1, H) - o) * HW + bx] * g3;
                    acc += hs[(clampi(Y + 2, H) - o) * HW + bx] * g4;
                    T s1v = s1c[(size_t)Y * a.s1.pitch + X];
                    T t = s1v * (T)0.4;
                    bl[by * BW + bx] = t + acc * (T)0.6;
                    int ccx = bx - 2, ccy = by - 2;
                    if (ccx >= 0 && ccx < TX && ccy >= 0 && ccy < TY) sm_dt[ccy * TX + ccx] = acc - s1v;
                }
            }
        };
        vsmooth(col, seg);
        if (tid < 4 * SEG) vsmooth(TX + (tid & 3), tid >> 2);
        __syncthreads();
        // 4. derivatives of the blend at the centre pixels (vertical window slides), psi products
        if (col_ok) {
            const int cy0 = seg * PPT;
            if (INTERIOR) {
                const T* bcol = bl + (col + 2);
                T v0 = bcol[(cy0 + 0) * BW], v1 = bcol[(cy0 + 1) * BW], v2 = bcol[(cy0 + 2) * BW], v3 = bcol[(cy0 + 3) * BW];
#pragma unroll
                for (int k = 0; k < PPT; k++) {
                    const int cy = cy0 + k;
                    T v4 = bcol[(cy + 4) * BW];
                    const T* brow = bl + (cy + 2) * BW + col;
                    T ix = 0, iy = 0;
                    ix += brow[0] * d0; ix += brow[1] * d1; ix += brow[3] * d3; ix += brow[4] * d4;
                    iy += v0 * d0; iy += v1 * d1; iy += v3 * d3; iy += v4 * d4;
                    v0 = v1; v1 = v2; v2 = v3; v3 = v4;
                    T it = sm_dt[cy * TX + col];
                    T psi = 0;
                    if (active) {
                        T t = it + ix * cdu[k] + iy * cdv[k];
                        psi = psi_of(t * t, a.eps);
                    }
                    T px = psi * ix, py = psi * iy;
                    sxy[k] += px * iy; sx2[k] += px * ix; sy2[k] += py * iy; stx[k] += px * it; sty[k] += py * it;
                }
            } else {
#pragma unroll
                for (int k = 0; k < PPT; k++) {
                    const int cy = cy0 + k, Y = y0 + cy;
                    if (Y >= H) continue;
                    const int ox = x0 - 2, oy = y0 - 2;
                    const T* brow = bl + (cy + 2) * BW;
                    T ix = 0, iy = 0;
                    ix += brow[clampi(PX - 2, W) - ox] * d0; ix += brow[clampi(PX - 1, W) - ox] * d1;
                    ix += brow[clampi(PX + 1, W) - ox] * d3; ix += brow[clampi(PX + 2, W) - ox] * d4;
                    iy += bl[(clampi(Y - 2, H) - oy) * BW + col + 2] * d0; iy += bl[(clampi(Y - 1, H) - oy) * BW + col + 2] * d1;
                    iy += bl[(clampi(Y + 1, H) - oy) * BW + col + 2] * d3; iy += bl[(clampi(Y + 2, H) - oy) * BW + col + 2] * d4;
                    T it = sm_dt[cy * TX + col];
                    T psi = 0;
                    if (active) {
                        T t = it + ix * cdu[k] + iy * cdv[k];
                        psi = psi_of(t * t, a.eps);
                    }
                    T px = psi * ix, py = psi * iy;
                    sxy[k] += px * iy; sx2[k] += px * ix; sy2[k] += py * iy; stx[k] += px * it; sty[k] += py * it;
                }
            }
        }
        __syncthreads();
    }

    // 5. u+du, v+dv tiles with a 1-pixel halo -> phi on the tile plus its left/up halo; then the
    //    same two tiles are reloaded with plain u, v for the Laplacian (which acts on u, not u+du:
    //    S/OpticalFlow.cpp:437-438).  With du == nullptr both are the same and nothing is reloaded.
    T* tu = sm_raw;
    T* tv = tu + UW * UH;
    T* tphi = sm_bl;
    auto load_uv = [&](bool with_increment) {
        for (int uy = warp; uy < UH; uy += NWARP) {
            const size_t ro = (size_t)cy_(y0 - 1 + uy) * a.pitch;
            for (int ux = lane; ux < UW; ux += 32) {
                size_t o = ro + cx_(x0 - 1 + ux);
                T uv = a.u[o], vv = a.v[o];
                if (with_increment) { uv += a.du[o]; vv += a.dv[o]; }
                tu[uy * UW + ux] = uv;
                tv[uy * UW + ux] = vv;
            }
        }
    };
    load_uv(a.du != nullptr);
    __syncthreads();
    for (int py = warp; py < PH; py += NWARP) {
        const int Y = y0 - 1 + py;
        for (int px = lane; px < PW; px += 32) {
            const int X = x0 - 1 + px;
            T val = 0;
            if (INTERIOR || (X >= 0 && X < W && Y >= 0 && Y < H)) {
                int ui = py * UW + px;
                T u0 = tu[ui], v0 = tv[ui];
                T ux = 0, uy = 0, vx = 0, vy = 0;
                if (INTERIOR || X < W - 1) { ux = tu[ui + 1] - u0; vx = tv[ui + 1] - v0; }
                if (INTERIOR || Y < H - 1) { uy = tu[ui + UW] - u0; vy = tv[ui + UW] - v0; }
                T t = ux * ux + uy * uy + vx * vx + vy * vy;
                val = (T)0.5 / sqrt(t + a.eps);
            }
            tphi[py * PW + px] = val;
        }
    }
    __syncthreads();
    if (a.du) {
        load_uv(false);
        __syncthreads();
    }
    // 6. Laplacian (fork quirk F3), right-hand sides, inverse diagonals
    if (!col_ok) return;
#pragma unroll
    for (int k = 0; k < PPT; k++) {
        const int cy = seg * PPT + k, Y = y0 + cy, X = PX;
        if (!INTERIOR && Y >= H) continue;
        const int pi = (cy + 1) * PW + (col + 1), ui = (cy + 1) * UW + (col + 1);
        const T ph = tphi[pi];
        const bool xr = INTERIOR || X < W - 1, xl = INTERIOR || X > 0, yd = INTERIOR || Y < H - 1, yu = INTERIOR || Y > 0;
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
            lu -= (tu[ui + UW] - tu[ui]) * ph;
            lv -= (tv[ui + UW] - tv[ui]) * ph;
            if (yu) {
                lu += (tu[ui] - tu[ui - UW]) * tphi[pi - PW];
                lv += (tv[ui] - tv[ui - UW]) * tphi[pi - PW];
            }
        }
        if (xl) cf += tphi[pi - 1];
        if (xr) cf += ph;
        if (yu) cf += tphi[pi - PW];
        if (yd) cf += ph;
        cf *= a.alpha;
        T a_xy = sxy[k], a_x2 = sx2[k], a_y2 = sy2[k], a_tx = stx[k], a_ty = sty[k];
        if (C > 1) {
            T n = (T)C;
            a_xy /= n; a_x2 /= n; a_y2 /= n; a_tx /= n; a_ty /= n;
        }
        T reg = a.alpha * (T)0.05;
        size_t o = (size_t)Y * a.pitch + X;
        a.phi[o] = ph;
        a.dxy[o] = a_xy;
        a.iu[o] = a.omega / (a_x2 + reg + cf);
        a.iv[o] = a.omega / (a_y2 + reg + cf);
        a.bu[o] = -a_tx - a.alpha * lu;
        a.bv[o] = -a_ty - a.alpha * lv;
    }
}

template <typename T, int TX, int TY, int SEG>
__global__ void __launch_bounds__(TX * SEG) k_fused_assemble(FusedArgs<T> a) {
    __shared__ T sm_raw[(TY + 8) * (TX + 8) + (TY + 8) * (TX + 4)];
    __shared__ T sm_bl[(TY + 4) * (TX + 4)];
    __shared__ T sm_dt[TY * TX];
    static_assert(2 * (TX + 2) * (TY + 2) <= (TY + 8) * (TX + 8) + (TY + 8) * (TX + 4), "u/v tiles alias raw + hs");
    static_assert((TX + 1) * (TY + 1) <= (TY + 4) * (TX + 4), "phi tile aliases the blend tile");
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const bool interior = x0 >= 4 && y0 >= 4 && x0 + TX + 4 <= a.w && y0 + TY + 4 <= a.h;
    if (interior) fused_assemble_body<T, TX, TY, SEG, true>(a, sm_raw, sm_bl, sm_dt);
    else          fused_assemble_body<T, TX, TY, SEG, false>(a, sm_raw, sm_bl, sm_dt);
}

}  // namespace pf
