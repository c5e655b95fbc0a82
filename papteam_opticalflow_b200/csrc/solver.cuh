// Plan<T>: device arena + launch sequence of one coarse-to-fine solve
// (level driver S/OpticalFlow.cpp:735-846, SmoothFlowSOR :238-536), instantiated for float
// (fast mode) and double (parity mode).  Host code only orchestrates: all arithmetic is in
// kernels.cuh / sor.cuh.
#pragma once
#include <cstring>
#include <chrono>
#include <algorithm>
#include <cmath>
#include <memory>
#include <tuple>

#include "common.cuh"
#include "geometry.hpp"
#include "kernels.cuh"
#include "sor.cuh"
#include "sor_lex.cuh"
#include "fused_tma.cuh"
#include "fused_cp.cuh"
#include "staging.hpp"

namespace pf {

// ---- abstract plan the C ABI talks to -----------------------------------------------------------
struct PlanBase {
    virtual ~PlanBase() {}
    virtual int levels() const = 0;
    virtual void upload(const double* im1, const double* im2) = 0;
    virtual void solve(int repeats, double* ms_total) = 0;
    virtual void download(double* vx, double* vy, double* warp) = 0;
    virtual void execute(double* vx, double* vy, double* warp, const double* im1,
                         const double* im2, double* timings) = 0;
    virtual void profile(double* timings, double* counters) = 0;
    virtual int level_timings(double* out, int max_levels) const = 0;
    virtual int mixture_params(double* alpha, double* sigma, double* beta, int n) = 0;   // channels written
    virtual void solve_async(int repeats) = 0;   // enqueue only
    virtual void seq_first(const unsigned char* frame) = 0;
    virtual void seq_next(const unsigned char* frame, void* out, int format) = 0;   // format: PF_SEQ_*
    virtual cudaStream_t stream() const = 0;
    virtual int device() const = 0;
    virtual void set_blocking_wait(int block) = 0;   // 1 / 0, or -1 = the process default (common.cuh)
};

// ---- simple bump allocator over one cudaMalloc ---------------------------------------------------
class Arena {
  public:
    ~Arena() { release(); }
    void reserve(size_t bytes) {
        release();
        PF_CUDA(cudaMalloc(&base_, bytes));
        cap_ = bytes;
        off_ = 0;
    }
    void release() {
        if (base_) cudaFree(base_);
        base_ = nullptr;
        cap_ = off_ = 0;
    }
    template <typename U>
    U* take(size_t n) {
        size_t bytes = round_up(n * sizeof(U), 256);
        if (off_ + bytes > cap_) throw Error(PF_ENOMEM, "arena overflow");
        U* p = reinterpret_cast<U*>(static_cast<char*>(base_) + off_);
        off_ += bytes;
        return p;
    }
    static size_t need(size_t n, size_t elem) { return round_up(n * elem, 256); }

  private:
    void* base_ = nullptr;
    size_t cap_ = 0, off_ = 0;
};

struct Span {
    int phase, level;
    cudaEvent_t a, b;
};

template <typename T>
Taps<T> make_taps(const double* v, int half) {
    if (half > kMaxHalf) throw Error(PF_EUNSUPPORTED, "Gaussian half-width > 8 not supported");
    Taps<T> t;
    t.half = half;
    for (int i = 0; i < 2 * kMaxHalf + 1; i++) t.v[i] = 0;
    for (int i = 0; i < 2 * half + 1; i++) t.v[i] = (T)v[i];
    return t;
}

// Sweeps fused per launch of the tile kernel: the whole solve if the image fits one region, otherwise
// the value minimising a cost model fitted on B200 (tools/sor_sweep.py, tools/ablation.py): a tile
// visit costs ~a cycles (stage -> registers, write-back, exposed TMA) plus ~b per half-sweep, a pass
// costs `launch` cycles on top.  Two fits:
//   throughput (default): SM time = tiles/148 x visit cost, small launch term -- what counts when many
//       pairs are in flight and other streams fill the idle SMs of a small level (+6 % pairs/s);
//   latency (PF_TUNE_LATENCY plans: the one-shot entry points; PF_SOR_TUNE=latency forces it): whole waves and the
//       full launch/ramp cost of every pass -- it spreads small levels over all SMs with large halos and is ~5 %
//       faster for ONE pair alone (19.5 vs 20.5 ms at 1920x1080), 5 % slower with 16 pairs in flight.
// PF_SOR_MODEL=a:b:launch:kind (kind 1 = whole waves, 0 = at least one wave, 2 = SM time) overrides.
inline int choose_fused_sweeps(int w, int h, int nsor, int region_h, int forced, int tune = PF_TUNE_THROUGHPUT) {
    if (w <= kSorRegionW && h <= region_h) return nsor;
    if (forced > 0) return std::min(forced, nsor);
    if (const char* e = getenv("PF_SOR_TUNE")) {   // forces one fit for every plan
        if (!strcmp(e, "latency")) tune = PF_TUNE_LATENCY;
        else if (!strcmp(e, "throughput")) tune = PF_TUNE_THROUGHPUT;
    }
    double ma, mb, ml, mkind;
    if (tune == PF_TUNE_LATENCY) { ma = 5500; mb = 380; ml = 7200; mkind = 1; }
    else { ma = 12000; mb = 380; ml = 1500; mkind = 2; }   // refitted at the end of round 1: 127.0 -> 128.5 pairs/s vs a = 8000
    if (const char* e = getenv("PF_SOR_MODEL")) sscanf(e, "%lf:%lf:%lf:%lf", &ma, &mb, &ml, &mkind);
    double best = 1e300;
    int best_t = 1;
    // 64-px regions shrink by 4 t per pass: t <= 15.  The latency fit may use all of it (a coarse level of <= 148 tiles is one
    // wave whatever the step, and every pass saved is ~3.5 us: 16.26 -> 15.93 ms per 1920-wide pair with the cap at 15 instead of
    // 12); the throughput fit never chooses more than 8.
    int max_fuse = tune == PF_TUNE_LATENCY ? 15 : 12;
    if (const char* e = getenv("PF_SOR_MAXFUSE")) max_fuse = std::max(1, atoi(e));
    for (int t = 1; t <= std::min(nsor, max_fuse); t++) {
        SorTiling tx = sor_tiling(w, kSorRegionW, 2 * t), ty = sor_tiling(h, region_h, 2 * t);
        if (tx.ntiles == 0 || ty.ntiles == 0) break;
        const double ctas = (double)tx.ntiles * ty.ntiles;
        const double waves = mkind == 1 ? std::ceil(ctas / 148.0) : mkind == 0 ? std::max(1.0, ctas / 148.0) : ctas / 148.0;
        const double passes = std::ceil((double)nsor / t);
        const double cost = passes * (waves * (ma + 2.0 * t * mb) + ml);
        if (cost < best) { best = cost; best_t = t; }
    }
    return best_t;
}


// ---- SOR dispatch shared by the Plan and the single-stage entry points ---------------------------
// Solves from du = dv = 0 and leaves the result in du/dv (the ping-pong pointers may be swapped).
template <typename T>
struct SorRunner {
    static constexpr bool kF64 = sizeof(T) == 8;
    // tile-kernel shape: FP32 64x64 regions (R=4, 16 warps), FP64 64x32 (R=2)
#ifndef PF_SOR_R
#define PF_SOR_R 8
#define PF_SOR_NW 8
#endif
    static constexpr int kR = kF64 ? 2 : PF_SOR_R;
    static constexpr int kNW = kF64 ? 16 : PF_SOR_NW;
    static constexpr int kRegionH = kR * kNW;
    bool lex = false, hybrid = false, simple_rb = false, use_tma = true, small_regions = false, small_four_warps = true;
    int forced_fuse = 0, coop_max_blocks = 1, sms = 148, ctas_per_sm = 1;
    int tune = PF_TUNE_THROUGHPUT;
    bool pdl = false;    // launch the tile kernel with programmatic stream serialisation (common.cuh; latency-tuned plans)
    int lex_from = -1;   // experiment (PF_LEX_FROM=k): pyramid levels >= k use the lexicographic kernel in every mode
    bool lex_band = true;          // k_sor_lex (band-march) instead of the grid-synchronised k_sor_wavefront (PF_LEX_IMPL=coop)
    // sweeps (= compute warps) per CTA of k_sor_lex.  More sweeps per CTA mean fewer hand-offs between sweep groups (each costs
    // ~18 steps of latency): FP32 16 (19 warps x 96 registers fill the register file; measured 4 / 8 / 16: 107 / 83 / 71 ms per
    // 1920-wide pair in fp32_wavefront, 42.8 / 35.3 / 33.8 ms in fp32_hybrid), FP64 8 (158 registers, 175 KB of rings).
#ifndef PF_LEX_NS
#define PF_LEX_NS 16
#endif
    static constexpr int kLexNS = kF64 ? 8 : PF_LEX_NS;
    static constexpr size_t kLexFlagWords = 1u << 16;
    int* lex_flags = nullptr;      // ticket + abort + progress words of k_sor_lex, zeroed before every launch
    int* lex_err = nullptr;        // sticky error word
    cudaStream_t st = nullptr;

    SorRunner() = default;
    ~SorRunner() {
        if (lex_flags) cudaFree(lex_flags);
    }
    SorRunner(const SorRunner&) = delete;
    SorRunner& operator=(const SorRunner&) = delete;

    // stage + double-buffered exchange rows + alignment slack
    static size_t sor_smem_bytes() { return sizeof(SorStage<T, kR, kNW>) + sizeof(T) * 2 * 2 * kNW * 2 * kSorRegionW + 128; }

    void init(int mode, int device, cudaStream_t stream, int tuning = PF_TUNE_THROUGHPUT) {
        lex = mode_is_lex(mode);
        hybrid = mode == PF_MODE_FP32_HYBRID;
        st = stream;
        tune = tuning;
        const char* e = getenv("PF_SOR_FUSE");
        forced_fuse = e ? atoi(e) : 0;
        e = getenv("PF_SOR_SIMPLE");
        simple_rb = e && atoi(e);
        e = getenv("PF_SOR_TMA");
        use_tma = !(e && !atoi(e));
        e = getenv("PF_LEX_FROM");
        lex_from = e ? atoi(e) : -1;
        PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        if (!lex && !simple_rb && use_tma) {   // (the hybrid mode needs both kernels)
            size_t bytes = sor_smem_bytes();
            PF_CUDA(cudaFuncSetAttribute(k_sor_rb_tma<T, kR, kNW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            int per_sm = 1;
            PF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sor_rb_tma<T, kR, kNW>, kNW * 32, bytes));
            ctas_per_sm = std::max(1, per_sm);
            e = getenv("PF_SOR_SMALL");
            small_regions = !kF64 && !(e && !atoi(e));
            if constexpr (!kF64) {
                if (small_regions) { init_small<2, 16>(); init_small<4, 16>(); init_small<8, 4>(); init_small<10, 4>(); }
                e = getenv("PF_SOR_SMALL_WARPS");      // 4 (default): the four-warp variants below; 16: round 2's R = 2 / 4, 16 warps
                small_four_warps = !(e && atoi(e) == 16);
            }
        }
        e = getenv("PF_LEX_IMPL");
        lex_band = !(e && !strcmp(e, "coop"));
        if ((lex || hybrid || lex_from >= 0) && lex_band) {
            PF_CUDA(cudaFuncSetAttribute(k_sor_lex<T, kLexNS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)LexCfg<T, kLexNS>::smem_bytes()));
            PF_CUDA(cudaMalloc(&lex_flags, (kLexFlagWords + 64) * sizeof(int)));
            PF_CUDA(cudaMemset(lex_flags, 0, (kLexFlagWords + 64) * sizeof(int)));
            lex_err = lex_flags + kLexFlagWords;
        }
        if ((lex || hybrid || lex_from >= 0) && !lex_band) {
            int coop = 0, per_sm = 0;
            PF_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
            if (!coop) throw Error(PF_EUNSUPPORTED, "device lacks cooperative launch");
            PF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sor_wavefront<T>, 256, 0));
            coop_max_blocks = std::max(1, per_sm * sms);
        }
    }

    // ---- levels that fit ONE region: the whole solve is one launch of one CTA, bound by the chain of half-sweeps.  With
    //      R rows per thread a half-sweep is R dependent pixel updates; regions of 64 x 32 (R = 2) and 64 x 64 (R = 4) with
    //      16 warps cut that chain to a quarter / half of the 64 x 64, R = 8, 8-warp region the large levels use (which
    //      trades the longer chain for fewer, register-richer warps).  Same update, same bits.
    template <int R, int NW>
    static size_t small_smem_bytes() { return sizeof(SorStage<T, R, NW>) + sizeof(T) * 2 * 2 * NW * 2 * kSorRegionW + 128; }
    template <int R, int NW>
    void init_small() {
        PF_CUDA(cudaFuncSetAttribute(k_sor_rb_tma<T, R, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem_bytes<R, NW>()));
    }
    template <int R, int NW>
    void launch_single_region(const SorArgs<T>& a, T* du_out, T* dv_out, int nsor) {
        constexpr int RH = R * NW;
        SorMaps m;
        m.phi = make_plane_map(a.phi, a.w, a.h, a.pitch, 72, RH + 1);
        m.dxy = make_plane_map(a.dxy, a.w, a.h, a.pitch, kSorRegionW, RH);
        m.iu = make_plane_map(a.iu, a.w, a.h, a.pitch, kSorRegionW, RH);
        m.iv = make_plane_map(a.iv, a.w, a.h, a.pitch, kSorRegionW, RH);
        m.bu = make_plane_map(a.bu, a.w, a.h, a.pitch, kSorRegionW, RH);
        m.bv = make_plane_map(a.bv, a.w, a.h, a.pitch, kSorRegionW, RH);
        m.du = make_plane_map(du_out, a.w, a.h, a.pitch, kSorRegionW, RH);   // not read: the solve starts from zero
        m.dv = make_plane_map(dv_out, a.w, a.h, a.pitch, kSorRegionW, RH);
        launch_chain(pdl, k_sor_rb_tma<T, R, NW>, dim3(1), dim3(NW * 32), small_smem_bytes<R, NW>(), st, m, du_out, dv_out, a.w, a.h, a.pitch,
                     a.alpha, a.omega, nsor, 0, 1, 1, kSorRegionW, RH, 0, SorPeer<T>());
    }

    // one launch of the tile kernel: sweeps fused, whether it reads du/dv, and its tiling
    struct SorPass {
        int nsw;
        bool has_input;
        SorTiling tx, ty;
        // exact output rows of tile row t, and the rows its region reads
        int out_lo(int t) const { return t > 0 ? t * ty.step + 2 * nsw : 0; }
        int out_hi(int t, int h) const { return (t * ty.step + kRegionH >= h) ? h : t * ty.step + kRegionH - 2 * nsw; }
        int in_lo(int t) const { return std::max(0, t * ty.step - 1); }
        int in_hi(int t, int h) const { return std::min(h, t * ty.step + kRegionH); }
    };

    std::vector<SorPass> schedule(int w, int h, int nsor) const {
        std::vector<SorPass> v;
        int fuse = choose_fused_sweeps(w, h, nsor, kRegionH, forced_fuse, tune);
        for (int done = 0; done < nsor;) {
            SorPass ps;
            ps.nsw = std::min(fuse, nsor - done);
            ps.has_input = done > 0;
            ps.tx = sor_tiling(w, kSorRegionW, 2 * ps.nsw);
            ps.ty = sor_tiling(h, kRegionH, 2 * ps.nsw);
            if (ps.tx.ntiles == 0 || ps.ty.ntiles == 0) throw Error(PF_EINVAL, "SOR tiling failed");
            v.push_back(ps);
            done += ps.nsw;
        }
        return v;
    }

    // developer aid (PF_LEX_STATS=1): one synchronous launch of the instrumented kernel, per-CTA clocks to stderr
    void lex_stats(LexArgs<T> la) {
        const int n = la.NI * la.NK;
        long long* d = nullptr;
        PF_CUDA(cudaMalloc(&d, (size_t)n * 16 * sizeof(long long)));
        PF_CUDA(cudaMemset(d, 0, (size_t)n * 16 * sizeof(long long)));
        la.stats = d;
        PF_CUDA(cudaFuncSetAttribute(k_sor_lex<T, kLexNS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)LexCfg<T, kLexNS>::smem_bytes()));
        k_sor_lex<T, kLexNS, true><<<n, LexCfg<T, kLexNS>::THREADS, LexCfg<T, kLexNS>::smem_bytes(), st>>>(la);
        PF_CUDA(cudaStreamSynchronize(st));
        std::vector<long long> h((size_t)n * 16);
        PF_CUDA(cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(d);
        long long t0 = 0, t1 = 0;
        for (int i = 0; i < n; i++)
            if (h[i * 16 + 2]) { t0 = t0 ? std::min(t0, h[i * 16 + 2]) : h[i * 16 + 2]; t1 = std::max(t1, h[i * 16 + 3]); }
        fprintf(stderr, "k_sor_lex %dx%d nsor=%d NI=%d NK=%d: %lld ns first start to last end\n", la.w, la.h, la.nsor, la.NI, la.NK, t1 - t0);
        for (int i = 0; i < n; i++) {
            const long long* o = &h[i * 16];
            if (!o[2]) continue;
            fprintf(stderr, "  I=%lld K=%lld start %7lld ns end %7lld ns | cycles %8lld wait_plane %8lld wait_loader %8lld steps %lld -> %lld cyc/step busy\n", o[0], o[1],
                    o[2] - t0, o[3] - t0, o[7], o[4], o[5], o[6], (o[7] - o[4] - o[5]) / std::max(1ll, o[6]));
            fprintf(stderr, "      passes step 0/64/128/... at us:");
            for (int q = 0; q < 8; q++) if (o[8 + q]) fprintf(stderr, " %.1f", (o[8 + q] - t0) * 1e-3);
            fprintf(stderr, "\n");
        }
    }

    // after the stream has been synchronised: did a hand-off of k_sor_lex time out?
    void check_lex() {
        if (!lex_err) return;
        int e = 0;
        PF_CUDA(cudaMemcpy(&e, lex_err, sizeof(int), cudaMemcpyDeviceToHost));
        if (e) {
            cudaMemset(lex_err, 0, sizeof(int));
            throw Error(PF_ECUDA, "lexicographic SOR: a hand-off between bands timed out");
        }
    }

    // CTAs of a launch over `nrows` tile rows of pass `ps` (persistent kernel: at most one wave)
    int grid_for(const SorPass& ps, int nrows) const { return std::min(ps.tx.ntiles * nrows, sms * ctas_per_sm); }

    // tile rows [ty_begin, ty_end) of one pass: reads du/dv (if has_input), writes du2/dv2
    void launch_pass(SorArgs<T> a, const SorPass& ps, T* du, T* dv, T* du2, T* dv2, int ty_begin, int ty_end,
                     const SorPeer<T>& peer = SorPeer<T>()) {
        const int w = a.w, h = a.h, nrows = ty_end - ty_begin;
        if (nrows <= 0) return;
        a.du_in = ps.has_input ? du : nullptr; a.dv_in = ps.has_input ? dv : nullptr; a.du = du2; a.dv = dv2;
        if (use_tma) {
            // persistent, TMA-staged variant: one CTA per SM walks the tiles round-robin
            SorMaps m;
            m.phi = make_plane_map(a.phi, w, h, a.pitch, 72, kRegionH + 1);
            m.dxy = make_plane_map(a.dxy, w, h, a.pitch, kSorRegionW, kRegionH);
            m.iu = make_plane_map(a.iu, w, h, a.pitch, kSorRegionW, kRegionH);
            m.iv = make_plane_map(a.iv, w, h, a.pitch, kSorRegionW, kRegionH);
            m.bu = make_plane_map(a.bu, w, h, a.pitch, kSorRegionW, kRegionH);
            m.bv = make_plane_map(a.bv, w, h, a.pitch, kSorRegionW, kRegionH);
            m.du = make_plane_map(ps.has_input ? du : du2, w, h, a.pitch, kSorRegionW, kRegionH);
            m.dv = make_plane_map(ps.has_input ? dv : dv2, w, h, a.pitch, kSorRegionW, kRegionH);
            size_t smem = sor_smem_bytes();
            launch_chain(pdl, k_sor_rb_tma<T, kR, kNW>, dim3(grid_for(ps, nrows)), dim3(kNW * 32), smem, st,
                         m, du2, dv2, w, h, a.pitch, a.alpha, a.omega, ps.nsw, ps.has_input ? 1 : 0, ps.tx.ntiles, nrows,
                         ps.tx.step, ps.ty.step, ty_begin, peer);
        } else {
            if (ty_begin != 0 || ty_end != ps.ty.ntiles) throw Error(PF_EUNSUPPORTED, "row-band split needs the TMA SOR kernel");
            k_sor_rb_tile<T, kR, kNW><<<dim3(ps.tx.ntiles, ps.ty.ntiles), kNW * 32, 0, st>>>(a, ps.nsw, ps.tx.step, ps.ty.step);
        }
    }

    // returns the number of kernel launches
    int run(SorArgs<T> a, T*& du, T*& dv, T*& du2, T*& dv2, int nsor, int level = -1) {
        const int w = a.w, h = a.h;
        const bool lex = this->lex || (hybrid && w <= PF_HYBRID_MAX_WIDTH) || (lex_from >= 0 && level >= lex_from);
        size_t bytes = plane_for(w, h) * sizeof(T);
        int launches = 0;
        if (lex || simple_rb || nsor == 0) {
            PF_CUDA(cudaMemsetAsync(du, 0, bytes, st));
            PF_CUDA(cudaMemsetAsync(dv, 0, bytes, st));
        }
        if (nsor == 0) return 0;
        if (lex && lex_band) {
            LexArgs<T> la;
            la.phi = a.phi; la.dxy = a.dxy; la.iu = a.iu; la.iv = a.iv; la.bu = a.bu; la.bv = a.bv;
            la.du = du; la.dv = dv; la.w = w; la.h = h; la.pitch = a.pitch; la.alpha = a.alpha; la.omega = a.omega;
            la.nsor = nsor;
            lex_grid<kLexNS>(h, nsor, la.NI, la.NK);
            const size_t words = kLexFlagHeader + (size_t)la.NI * nsor;
            if (words > kLexFlagWords) throw Error(PF_EUNSUPPORTED, "level too tall for the lexicographic SOR flag buffer");
            la.flags = lex_flags; la.err = lex_err;
            PF_CUDA(cudaMemsetAsync(lex_flags, 0, words * sizeof(int), st));
            la.stats = nullptr;
            la.pub_every = 1; la.opt = 0;
            if (const char* pe = getenv("PF_LEX_PUB")) la.pub_every = std::max(1, atoi(pe));
            if (const char* oe = getenv("PF_LEX_OPT")) la.opt = atoi(oe);
            if (const char* se = getenv("PF_LEX_STATS")) {
                if (atoi(se)) { lex_stats(la); return 1; }
            }
            k_sor_lex<T, kLexNS><<<la.NI * la.NK, LexCfg<T, kLexNS>::THREADS, LexCfg<T, kLexNS>::smem_bytes(), st>>>(la);
            return 1;
        }
        if (lex) {
            a.du = du; a.dv = dv; a.du_in = nullptr; a.dv_in = nullptr;
            long long work = (long long)nsor * std::min(w, h);
            int blocks = (int)std::min<long long>(coop_max_blocks, std::max<long long>(1, (work + 255) / 256));
            void* args[] = {&a, &nsor};
            PF_CUDA(cudaLaunchCooperativeKernel((void*)k_sor_wavefront<T>, dim3(blocks), dim3(256), args, 0, st));
            return 1;
        }
        if (simple_rb) {
            a.du = du; a.dv = dv; a.du_in = nullptr; a.dv_in = nullptr;
            for (int s = 0; s < nsor; s++)
                for (int c = 0; c < 2; c++) {
                    k_sor_rb_half<T><<<dim3(ceil_div(ceil_div(w, 2) + 1, 128), h), 128, 0, st>>>(a, c);
                    launches++;
                }
            return launches;
        }
        if (small_regions && use_tma && w <= kSorRegionW && h <= 64) {
            if constexpr (!kF64) {
                // four warps, one per scheduler, R = 8 / 10 rows each: a half-sweep is R independent updates issued back to
                // back by ONE warp per scheduler plus a four-warp barrier, instead of sixteen warps each doing two updates
                // between 512-thread barriers
                if (small_four_warps && h <= 32) launch_single_region<8, 4>(a, du2, dv2, nsor);
                else if (small_four_warps && h <= 40) launch_single_region<10, 4>(a, du2, dv2, nsor);
                else if (h <= 32) launch_single_region<2, 16>(a, du2, dv2, nsor);
                else launch_single_region<4, 16>(a, du2, dv2, nsor);
                std::swap(du, du2);
                std::swap(dv, dv2);
                return 1;
            }
        }
        for (const SorPass& ps : schedule(w, h, nsor)) {
            launch_pass(a, ps, du, dv, du2, dv2, 0, ps.ty.ntiles);
            launches++;
            std::swap(du, du2);
            std::swap(dv, dv2);
        }
        return launches;
    }
};

template <typename T>
class Plan : public PlanBase {
  public:
    static constexpr bool kF64 = sizeof(T) == 8;

    explicit Plan(const Params& p) : P(p) {
        lex_ = mode_is_lex(P.mode);
        PF_CUDA(cudaSetDevice(P.device));
        nlev_ = P.levels > 0 ? P.levels : levels_from_min_width(P.w, P.ratio, P.min_width);
        if (nlev_ < 1 || nlev_ > 64) throw Error(PF_EINVAL, "pyramid level count out of range (got " + std::to_string(nlev_) + ")");
        geo_ = level_geometry(P.w, P.h, P.ratio, nlev_);
        for (auto& g : geo_)
            if (g.w < 1 || g.h < 1) throw Error(PF_EINVAL, "pyramid level collapsed to zero size");
        fc_ = P.c == 1 ? 3 : (P.c == 3 ? 5 : P.c);
        if (fc_ > 16) throw Error(PF_EUNSUPPORTED, "more than 16 channels");
        bicubic_ = P.interp == PF_INTERP_BICUBIC;
        gmix_ = P.noise == PF_NOISE_GMIXTURE;
        const char* e = getenv("PF_NO_GRAPH");
        {
            const char* li = getenv("PF_LEX_IMPL");
            const bool coop = li && !strcmp(li, "coop");    // cooperative launches are not captured
            use_graph_ = !(e && atoi(e)) && !((lex_ || P.mode == PF_MODE_FP32_HYBRID || getenv("PF_LEX_FROM")) && coop);
        }
        e = getenv("PF_UNFUSED");
        fused_ = !(e && atoi(e)) && P.noise != PF_NOISE_GMIXTURE;   // the mixture weight lives in the stage-by-stage path
        e = getenv("PF_FUSED_TMA");
        fused_tma_ = !(e && !atoi(e));
        if (fused_tma_)
        {
            PF_CUDA(cudaFuncSetAttribute(k_fused_tma<T, kFTY, kFSEG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(FusedSmem<T, kFTY>) + 128)));
            PF_CUDA(cudaFuncSetAttribute(k_fused_tma<T, kFTYs, kFSEGs>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(FusedSmem<T, kFTYs>) + 128)));
            if constexpr (!kF64) {
                // channel-parallel form for the coarse levels of latency-tuned plans (fused_cp.cuh): C groups of 64 * kCpSeg threads
                e = getenv("PF_FUSED_CP");
                fused_cp_ = !(e && !atoi(e)) && 64 * kCpSeg * fc_ <= 1024;
                if (fused_cp_) {
                    PF_CUDA(cudaFuncSetAttribute(k_fused_cp<T, kFTYs, kCpSeg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)fused_cp_smem_bytes<T, kFTYs>(fc_)));
                    PF_CUDA(cudaFuncSetAttribute(k_fused_cp<T, kFTYs, kCpSeg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)fused_cp_smem_bytes<T, kFTYs>(fc_, true)));
                    // update + warp folded into that kernel (latency-tuned plans, default solver branches, nInner = 1)
                    e = getenv("PF_FUSED_WARP");
                    fold_ = !(e && !atoi(e)) && P.tune == PF_TUNE_LATENCY && P.n_inner == 1 && !bicubic_ && !gmix_ && fused_;
                }
            }
        }
        try {
            PF_CUDA(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
            for (auto& ev : ev_) PF_CUDA(cudaEventCreate(&ev));
            // measured (tools/env_ab.py PF_BRANCHES 0 1): 15.86 -> 15.73 ms for a latency-tuned plan, 19.6 -> 20.0 ms for one
            // pair alone through a throughput-tuned plan (the cross-stream edges cost what the overlap gains; with 16 pairs
            // in flight no difference): latency-tuned plans only
            e = getenv("PF_BRANCHES");
            branches_ = e ? atoi(e) != 0 : P.tune == PF_TUNE_LATENCY;
            if (branches_) {
                PF_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
                for (int i = 0; i < kAux; i++) {
                    PF_CUDA(cudaStreamCreateWithFlags(&aux_[i], cudaStreamNonBlocking));
                    PF_CUDA(cudaEventCreateWithFlags(&ev_join_[i], cudaEventDisableTiming));
                }
            }
            allocate();
            sor_.init(P.mode, P.device, st_, P.tune);
            set_pdl(P.tune == PF_TUNE_LATENCY);
        } catch (...) {
            release();       // the destructor does not run for a half-built object (PF_ENOMEM is the realistic cause)
            throw;
        }
    }

    ~Plan() override { release(); }

  private:
    void release() {
        cudaSetDevice(P.device);
        if (gexec_) cudaGraphExecDestroy(gexec_);
        for (auto& gp : gexec_seq_) for (auto& g : gp) if (g) cudaGraphExecDestroy(g);
        if (graph_) cudaGraphDestroy(graph_);
        for (auto& s : spans_) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
        for (auto& ev : ev_) if (ev) cudaEventDestroy(ev);
        if (ev_block_) cudaEventDestroy(ev_block_);
        ev_block_ = nullptr;
        if (ev_fork_) cudaEventDestroy(ev_fork_);
        ev_fork_ = nullptr;
        for (int i = 0; i < kAux; i++) {
            if (ev_join_[i]) cudaEventDestroy(ev_join_[i]);
            if (aux_[i]) cudaStreamDestroy(aux_[i]);
            ev_join_[i] = nullptr; aux_[i] = nullptr;
        }
        if (st_) cudaStreamDestroy(st_);
        gexec_ = nullptr; graph_ = nullptr; st_ = nullptr;
        for (auto& gp : gexec_seq_) for (auto& g : gp) g = nullptr;
        for (auto& ev : ev_) ev = nullptr;
        spans_.clear();
        arena_.release();
    }

  public:

    int levels() const override { return nlev_; }

    // Host buffers are the caller's numpy arrays (pageable) or pinned memory.  Pageable ones of a useful size go
    // through the multi-threaded stager (staging.hpp), pinned ones straight to the copy engine.
    static constexpr size_t kStageMinBytes = 8u << 20;

    void upload(const double* im1, const double* im2) override {
        PF_CUDA(cudaSetDevice(P.device));
        size_t n = (size_t)P.h * P.w * P.c * sizeof(double);
        HostStager* hs = (2 * n >= kStageMinBytes && host_is_pageable(im1) && host_is_pageable(im2)) ? HostStager::for_device(P.device) : nullptr;
        if (hs) {
            hs->to_device({{d_in1_, const_cast<double*>(im1), n}, {d_in2_, const_cast<double*>(im2), n}}, st_);
            return;
        }
        PF_CUDA(cudaMemcpyAsync(d_in1_, im1, n, cudaMemcpyHostToDevice, st_));
        PF_CUDA(cudaMemcpyAsync(d_in2_, im2, n, cudaMemcpyHostToDevice, st_));
    }

    // enqueue (pinned destinations) or perform (pageable destinations) the output copies; an output the caller passed
    // as NULL is not copied at all (the reference driver never reads warpI2, Par/OpticalFlowCalculation.py:74-76,
    // and it is 60 % of the device->host bytes)
    void download_outputs(double* vx, double* vy, double* warp) {
        const size_t n = (size_t)P.h * P.w * sizeof(double);
        std::vector<CopyJob> jobs;
        if (vx) jobs.push_back({d_vx_, vx, n});
        if (vy) jobs.push_back({d_vy_, vy, n});
        if (warp) jobs.push_back({d_warp_, warp, n * P.c});
        size_t total = 0;
        bool pageable = !jobs.empty();
        for (const CopyJob& j : jobs) {
            total += j.bytes;
            pageable = pageable && host_is_pageable(j.host);
        }
        HostStager* hs = (pageable && total >= kStageMinBytes) ? HostStager::for_device(P.device) : nullptr;
        if (hs) {
            hs->to_host(jobs, st_);
            return;
        }
        for (const CopyJob& j : jobs) PF_CUDA(cudaMemcpyAsync(j.host, j.dev, j.bytes, cudaMemcpyDeviceToHost, st_));
    }

    void download(double* vx, double* vy, double* warp) override {
        PF_CUDA(cudaSetDevice(P.device));
        download_outputs(vx, vy, warp);
        sync_stream();
    }

    // device-only solve, `repeats` times back to back, timed with events on the launching stream
    void solve(int repeats, double* ms_total) override {
        PF_CUDA(cudaSetDevice(P.device));
        PF_CUDA(cudaEventRecord(ev_[0], st_));
        for (int i = 0; i < repeats; i++) run_solve();
        PF_CUDA(cudaEventRecord(ev_[1], st_));
        sync_stream();
        sor_.check_lex();
        float ms = 0;
        PF_CUDA(cudaEventElapsedTime(&ms, ev_[0], ev_[1]));
        if (ms_total) *ms_total = ms;
    }

    void execute(double* vx, double* vy, double* warp, const double* im1, const double* im2,
                 double* timings) override {
        PF_CUDA(cudaSetDevice(P.device));
        const auto h0 = std::chrono::steady_clock::now();
        PF_CUDA(cudaEventRecord(ev_[0], st_));
        upload(im1, im2);
        PF_CUDA(cudaEventRecord(ev_[1], st_));
        const auto h1 = std::chrono::steady_clock::now();
        run_solve();
        const auto h2 = std::chrono::steady_clock::now();
        PF_CUDA(cudaEventRecord(ev_[2], st_));
        download_outputs(vx, vy, warp);
        PF_CUDA(cudaEventRecord(ev_[3], st_));
        sync_stream();
        sor_.check_lex();
        if (timings) {
            for (int i = 0; i < PF_NUM_TIMINGS; i++) timings[i] = 0;
            float a = 0, b = 0, c = 0, d = 0;
            PF_CUDA(cudaEventElapsedTime(&a, ev_[0], ev_[3]));
            PF_CUDA(cudaEventElapsedTime(&b, ev_[0], ev_[1]));
            PF_CUDA(cudaEventElapsedTime(&c, ev_[1], ev_[2]));
            PF_CUDA(cudaEventElapsedTime(&d, ev_[2], ev_[3]));
            timings[PF_T_TOTAL] = a;
            timings[PF_T_H2D] = b;
            timings[PF_T_SOLVE] = c;
            timings[PF_T_D2H] = d;
            // host-side wall clock of the enqueue calls and of the whole call (diagnostics, ms)
            timings[13] = std::chrono::duration<double, std::milli>(h1 - h0).count();
            timings[14] = std::chrono::duration<double, std::milli>(h2 - h1).count();
            timings[15] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
        }
    }

    // eager solve with an event pair around every phase
    void profile(double* timings, double* counters) override {
        PF_CUDA(cudaSetDevice(P.device));
        profiling_ = true;
        span_used_ = 0;
        launches_ = sor_launches_ = sor_launches_l0_ = 0;
        PF_CUDA(cudaEventRecord(ev_[0], st_));
        enqueue_solve();
        set_phase(-1, 0);
        PF_CUDA(cudaEventRecord(ev_[1], st_));
        sync_stream();
        profiling_ = false;
        for (int i = 0; i < PF_NUM_TIMINGS; i++) timings[i] = 0;
        double sor_l0 = 0;
        level_ms_.assign((size_t)nlev_ * PF_NUM_TIMINGS, 0.0);
        for (size_t i = 0; i < span_used_; i++) {
            float ms = 0;
            PF_CUDA(cudaEventElapsedTime(&ms, spans_[i].a, spans_[i].b));
            timings[spans_[i].phase] += ms;
            level_ms_[(size_t)spans_[i].level * PF_NUM_TIMINGS + spans_[i].phase] += ms;
            if (spans_[i].phase == PF_T_PHASE5_SOR && spans_[i].level == 0) sor_l0 += ms;
        }
        float tot = 0;
        PF_CUDA(cudaEventElapsedTime(&tot, ev_[0], ev_[1]));
        timings[PF_T_SOLVE] = tot;
        timings[PF_T_TOTAL] = tot;
        if (counters) {
            double ps = 0;
            for (int k = 0; k < nlev_; k++)
                ps += (double)geo_[k].w * geo_[k].h * (P.n_outer + k) * P.n_inner * (P.n_sor + 3 * k);
            counters[0] = (double)launches_;
            counters[1] = (double)sor_launches_;
            counters[2] = ps;
            counters[3] = sor_l0;
            counters[4] = (double)sor_launches_l0_;
            counters[5] = (double)geo_[0].w * geo_[0].h * P.n_outer * P.n_inner * P.n_sor;
            counters[6] = 0;
            counters[7] = 0;
        }
    }

    // Gaussian-mixture parameters left by the last solve (GMPara after the last estGaussianMixture call)
    int mixture_params(double* alpha, double* sigma, double* beta, int n) override {
        if (!gmix_) throw Error(PF_EINVAL, "plan was not created with the Gaussian-mixture noise model");
        PF_CUDA(cudaSetDevice(P.device));
        double host[GM_FIELDS * kGmStride];
        sync_stream();
        PF_CUDA(cudaMemcpy(host, d_gm_, sizeof(host), cudaMemcpyDeviceToHost));
        n = std::min(n, fc_);
        for (int k = 0; k < n; k++) {
            alpha[k] = host[GM_ALPHA * kGmStride + k];
            sigma[k] = host[GM_SIGMA * kGmStride + k];
            beta[k] = host[GM_BETA * kGmStride + k];
        }
        return n;
    }

    void solve_async(int repeats) override {
        PF_CUDA(cudaSetDevice(P.device));
        for (int i = 0; i < repeats; i++) run_solve();
    }
    cudaStream_t stream() const override { return st_; }
    int device() const override { return P.device; }
    void set_blocking_wait(int block) override { block_wait_ = block < 0 ? default_block_waits() : block != 0; }
    void sync_stream() {
        if (!block_wait_) { PF_CUDA(cudaStreamSynchronize(st_)); return; }
        if (!ev_block_) PF_CUDA(cudaEventCreateWithFlags(&ev_block_, cudaEventBlockingSync | cudaEventDisableTiming));
        PF_CUDA(cudaEventRecord(ev_block_, st_));
        PF_CUDA(cudaEventSynchronize(ev_block_));
    }

    // per-level phase times of the last profile() call: out[level][PF_NUM_TIMINGS]
    int level_timings(double* out, int max_levels) const override {
        int n = std::min(max_levels, nlev_);
        if (level_ms_.empty()) return 0;
        for (int k = 0; k < n; k++)
            for (int j = 0; j < PF_NUM_TIMINGS; j++) out[k * PF_NUM_TIMINGS + j] = level_ms_[(size_t)k * PF_NUM_TIMINGS + j];
        return n;
    }

  private:
    // ------------------------------------------------------------------------------------------
    Img<T> view(T* base, int w, int h, int c) const {
        Img<T> v;
        v.p = base; v.w = w; v.h = h; v.c = c;
        v.pitch = pitch_for(w);
        v.plane = plane_for(w, h);
        return v;
    }

    void allocate() {
        const size_t pl0 = plane_for(P.w, P.h);
        const size_t in_elems = (size_t)P.h * P.w * P.c;
        size_t total = 0;
        total += 3 * Arena::need(in_elems, sizeof(double));              // in1, in2, warp out
        total += 2 * Arena::need((size_t)P.h * P.w, sizeof(double));     // vx, vy out
        for (int k = 0; k < nlev_; k++) total += 2 * Arena::need(plane_for(geo_[k].w, geo_[k].h) * P.c, sizeof(T));
        total += 2 * Arena::need(pl0 * P.c, sizeof(T));                  // blur tmp, blur
        total += 3 * Arena::need(pl0 * P.c, sizeof(T));                  // bicubic ix, iy, ixy
        total += 10 * Arena::need(pl0 * fc_, sizeof(T));                 // f1 f2 wf s1 s2 tmp blend dx dy dt
        total += 16 * Arena::need(pl0, sizeof(T));                       // scalar planes
        if (bicubic_) total += 3 * Arena::need(pl0 * fc_, sizeof(T));    // gradients of the Im2 features (Bicubic inner warp)
        total += Arena::need(64, sizeof(double)) * 3 + Arena::need(GM_FIELDS * kGmStride, sizeof(double)) + Arena::need(1, sizeof(BicubicTable));
        arena_.reserve(total + 4096);
        d_in1_ = arena_.take<double>(in_elems);
        d_in2_ = arena_.take<double>(in_elems);
        d_warp_ = arena_.take<double>(in_elems);
        d_vx_ = arena_.take<double>((size_t)P.h * P.w);
        d_vy_ = arena_.take<double>((size_t)P.h * P.w);
        pyr1_.resize(nlev_);
        pyr2_.resize(nlev_);
        for (int k = 0; k < nlev_; k++) {
            size_t n = plane_for(geo_[k].w, geo_[k].h) * P.c;
            pyr1_[k] = view(arena_.take<T>(n), geo_[k].w, geo_[k].h, P.c);
            pyr2_[k] = view(arena_.take<T>(n), geo_[k].w, geo_[k].h, P.c);
        }
        b_tmp_ = arena_.take<T>(pl0 * P.c);
        b_out_ = arena_.take<T>(pl0 * P.c);
        b_ix_ = arena_.take<T>(pl0 * P.c);
        b_iy_ = arena_.take<T>(pl0 * P.c);
        b_ixy_ = arena_.take<T>(pl0 * P.c);
        T** fcb[] = {&f1_, &f2_, &wf_, &s1_, &s2_, &tmp_, &blend_, &imdx_, &imdy_, &imdt_};
        for (T** b : fcb) *b = arena_.take<T>(pl0 * fc_);
        T** sc[] = {&u_, &v_, &u2_, &v2_, &du_, &dv_, &du2_, &dv2_, &phi_, &dxy_, &iu_, &iv_, &bu_, &bv_, &dx2_, &dy2_};
        for (T** b : sc) *b = arena_.take<T>(pl0);
        if (bicubic_) {
            g_ix_ = arena_.take<T>(pl0 * fc_);
            g_iy_ = arena_.take<T>(pl0 * fc_);
            g_ixy_ = arena_.take<T>(pl0 * fc_);
        }
        d_lap_ = arena_.take<double>(64);
        d_acc_ = arena_.take<double>(64);
        d_gm_ = arena_.take<double>(GM_FIELDS * kGmStride);
        d_tab_ = arena_.take<BicubicTable>(1);
        d_mm_ = reinterpret_cast<unsigned int*>(arena_.take<double>(64));   // magnitude min / max of the flow visualisation
        BicubicTable tab = make_bicubic_table();
        PF_CUDA(cudaMemcpy(d_tab_, &tab, sizeof(tab), cudaMemcpyHostToDevice));
        PF_CUDA(cudaMemset(d_acc_, 0, 64 * sizeof(double)));
    }

    // ------------------------------------------------------------------------------------------
    void set_phase(int phase, int level) {
        if (!profiling_) return;
        if (span_used_ == spans_.size()) {
            Span s{0, 0, nullptr, nullptr};
            PF_CUDA(cudaEventCreate(&s.a));
            PF_CUDA(cudaEventCreate(&s.b));
            spans_.push_back(s);
        }
        if (open_) {
            PF_CUDA(cudaEventRecord(spans_[span_used_ - 1].b, st_));
            open_ = false;
        }
        if (phase >= 0) {
            spans_[span_used_].phase = phase;
            spans_[span_used_].level = level;
            PF_CUDA(cudaEventRecord(spans_[span_used_].a, st_));
            span_used_++;
            open_ = true;
        }
    }

    dim3 grid2(int w, int h, int z = 1) const { return dim3(ceil_div(w, 128), h, z); }

    void filter_h(const Img<T>& s, const Img<T>& d, const Taps<T>& t) {
        k_filter_h<T><<<grid2(s.w, s.h, s.c), 128, 0, st_>>>(s, d, t);
        launches_++;
    }
    void filter_v(const Img<T>& s, const Img<T>& d, const Taps<T>& t) {
        k_filter_v<T><<<grid2(s.w, s.h, s.c), 128, 0, st_>>>(s, d, t);
        launches_++;
    }

    // h then v pass in one kernel (no intermediate plane)
    void filter_hv(const Img<T>& s, const Img<T>& d, const Taps<T>& th, const Taps<T>& tv, cudaStream_t st = nullptr) {
        const dim3 grid(ceil_div(s.w, 64), ceil_div(s.h, 16), s.c);
        if (!st) st = st_;
#define PF_HV(FH, FV) k_filter_hv_t<T, FH, FV><<<grid, 256, 0, st>>>(s, d, th, tv)
        switch (th.half * 16 + tv.half) {      // half-widths the pipeline uses get straight-line kernels
            case 0x11: PF_HV(1, 1); break;
            case 0x22: PF_HV(2, 2); break;
            case 0x33: PF_HV(3, 3); break;
            case 0x44: PF_HV(4, 4); break;
            case 0x10: PF_HV(1, 0); break;
            case 0x01: PF_HV(0, 1); break;
            default: k_filter_hv<T><<<grid, 256, 0, st>>>(s, d, th, tv);
        }
#undef PF_HV
        launches_++;
    }

    void run_solve() {
        if (!use_graph_) {
            enqueue_solve();
            return;
        }
        if (!gexec_) {
            PF_CUDA(cudaStreamBeginCapture(st_, cudaStreamCaptureModeThreadLocal));
            try {
                enqueue_solve();
            } catch (...) {
                cudaGraph_t g = nullptr;
                cudaStreamEndCapture(st_, &g);
                if (g) cudaGraphDestroy(g);
                throw;
            }
            PF_CUDA(cudaStreamEndCapture(st_, &graph_));
            PF_CUDA(cudaGraphInstantiate(&gexec_, graph_, 0));
        }
        PF_CUDA(cudaGraphLaunch(gexec_, st_));
    }

    // ---- SOR dispatch: leaves the result in du_/dv_ (pointers may be swapped) ------------------
    void run_sor(int w, int h, int pitch, int nsor, int level) {
        SorArgs<T> a;
        a.phi = phi_; a.dxy = dxy_; a.iu = iu_; a.iv = iv_; a.bu = bu_; a.bv = bv_;
        a.du = nullptr; a.dv = nullptr; a.du_in = nullptr; a.dv_in = nullptr;
        a.w = w; a.h = h; a.pitch = pitch;
        a.alpha = (T)P.alpha; a.omega = (T)1.8;
        int n = sor_.run(a, du_, dv_, du2_, dv2_, nsor, level);
        launches_ += n;
        sor_launches_ += n;
        if (level == 0) sor_launches_l0_ += n;
    }

    // ---- the whole solve on st_ ------------------------------------------------------------------
    // ---- the whole solve on st_, as phases so that several plans (one per GPU) can be driven in
    //      lock step by MultiPlan (multigpu.cuh) ---------------------------------------------------
    struct Ctx {
        Taps<T> d5, g5, d3;
        T eps;
        T *du_e, *dv_e, *du2_e, *dv2_e, *u_e, *v_e, *u2_e, *v2_e;   // pointer roles on entry
        int pw, ph;                                                   // previous (coarser) level size
        int w, h, pitch;                                              // current level
        int it = 0;                                                   // outer iteration of the level being enqueued
        bool folded = false;                                          // this level: update + warp inside the assembly kernel
        Img<T> f1, f2, wf, s1, s2, tmp, blend, imdx, imdy, imdt;
        Img<T> gix, giy, gixy;                                        // Bicubic inner warp: derivative images of f2
        FusedMaps fmaps;
    };
    Ctx cx_;

  public:
    // programmatic dependent launch of the chain kernels (common.cuh): latency-tuned plans; never the plans of a MultiPlan
    void set_pdl(bool on) { pdl_ = on; sor_.pdl = on; }
    void set_fold(bool on) { fold_ = fold_ && on; }   // MultiPlan drives the phases itself: never folded there
    int n_outer_at(int k) const { return P.n_outer + k; }
    int n_sor_at(int k) const { return P.n_sor + 3 * k; }
    int level_w(int k) const { return geo_[k].w; }
    int fused_tile_rows(int k) const { return small_tiles(geo_[k].w, geo_[k].h) ? kFTYs : kFTY; }   // tile height of k_fused_tma at level k
    int level_h(int k) const { return geo_[k].h; }

    // filter taps, eps, pointer roles on entry
    void ph_ctx() {
        const double d5raw[5] = {1.0 / 12, -8.0 / 12, 0.0 / 12, 8.0 / 12, -1.0 / 12};
        const double g5raw[5] = {0.02, 0.11, 0.74, 0.11, 0.02};
        const double d3raw[3] = {-0.5, 0, 0.5};
        Ctx& c = cx_;
        c.d5 = make_taps<T>(d5raw, 2); c.g5 = make_taps<T>(g5raw, 2); c.d3 = make_taps<T>(d3raw, 1);
        c.eps = (T)std::pow(0.001, 2);
        c.du_e = du_; c.dv_e = dv_; c.du2_e = du2_; c.dv2_e = dv2_;
        c.u_e = u_; c.v_e = v_; c.u2_e = u2_; c.v2_e = v2_;
        c.pw = c.ph = 0;
    }

    // Gaussian pyramid of one frame from its level 0 (S/GaussianPyramid.cpp:79-108)
    // `st`: the stream the chain is enqueued on (nullptr = st_); `tmp`: its own blur buffer (the two frames' pyramids run on
    // two streams side by side, see fork() / join())
    void ph_pyramid(int side, cudaStream_t st = nullptr, T* tmp = nullptr) {
        auto& pyr = side ? pyr2_ : pyr1_;
        if (!st) st = st_;
        if (!tmp) tmp = b_out_;
        for (int i = 1; i < nlev_; i++) {
            const Level& g = geo_[i];
            const Img<T>& src = pyr[g.src];
            Img<T> blurred = src;
            if (g.half > 0) {   // half-width 0 is the identity filter (level 1, quirk Q2)
                Taps<T> gt = make_taps<T>(g.taps.data(), g.half);
                blurred = view(tmp, src.w, src.h, src.c);
                filter_hv(src, blurred, gt, gt, st);
            }
            k_resize<T><<<grid2(g.w, g.h), 128, 0, st>>>(blurred, pyr[i], g.rate, g.rate, (T)1, 0);
            launches_++;
        }
    }

    // ---- independent kernels of one pair side by side: n auxiliary streams branch off st_ and are joined again.  Inside a
    //      captured graph these become parallel branches; the arithmetic is untouched.  (PF_BRANCHES=0: one chain.) ----
    int fork(int n) {
        if (!branches_) return 0;
        n = std::min(n, kAux);
        PF_CUDA(cudaEventRecord(ev_fork_, st_));
        for (int i = 0; i < n; i++) PF_CUDA(cudaStreamWaitEvent(aux_[i], ev_fork_, 0));
        return n;
    }
    void join(int n) {
        for (int i = 0; i < n; i++) {
            PF_CUDA(cudaEventRecord(ev_join_[i], aux_[i]));
            PF_CUDA(cudaStreamWaitEvent(st_, ev_join_[i], 0));
        }
    }
    cudaStream_t branch(int i, int forked) const { return i < forked ? aux_[i] : st_; }

    void ph_lap_init() {
        if (gmix_) {   // S/OpticalFlow.cpp:769-770
            k_gm_reset<<<1, 32, 0, st_>>>(d_gm_);
            launches_++;
            return;
        }
        if (kF64 && lex_) {
            k_fill_f64<<<1, 64, 0, st_>>>(d_lap_, 0.02, 64);   // S/OpticalFlow.cpp:773-775 (a kernel, not a copy from the stack: the solve is graph captured)
            launches_++;
        }
    }

    // -- Construction: import + both pyramids --
    void ph_begin() {
        ph_ctx();
        set_phase(PF_T_CONSTRUCTION, 0);
        const int nb = fork(1);                 // frame 2's import + pyramid next to frame 1's
        k_import_hwc<T><<<grid2(P.w, P.h), 128, 0, st_>>>(d_in1_, pyr1_[0]);
        k_import_hwc<T><<<grid2(P.w, P.h), 128, 0, branch(0, nb)>>>(d_in2_, pyr2_[0]);
        launches_ += 2;
        ph_pyramid(0);
        ph_pyramid(1, branch(0, nb), nb ? b_tmp_ : b_out_);
        join(nb);
        ph_lap_init();
    }

    // -- Allocation: features, flow upsampling, warp (S/OpticalFlow.cpp:790-818); smoothed Im1 --
    void ph_level(int k) {
        Ctx& c = cx_;
        const int w = geo_[k].w, h = geo_[k].h, pitch = pitch_for(w);
        c.w = w; c.h = h; c.pitch = pitch;
        const size_t plane_bytes = plane_for(w, h) * sizeof(T);
        set_phase(PF_T_ALLOCATION, k);
        c.f1 = view(f1_, w, h, fc_); c.f2 = view(f2_, w, h, fc_); c.wf = view(wf_, w, h, fc_);
        c.s1 = view(s1_, w, h, fc_); c.s2 = view(s2_, w, h, fc_); c.tmp = view(tmp_, w, h, fc_);
        c.blend = view(blend_, w, h, fc_); c.imdx = view(imdx_, w, h, fc_);
        c.imdy = view(imdy_, w, h, fc_); c.imdt = view(imdt_, w, h, fc_);
        // the two feature images and the two flow upsamplings are independent of one another: four branches
        const int nb = fork(3);
        if (P.c == 1 || P.c == 3) {
            int swap = (k == 0 && P.col_type == 1) ? 1 : 0;
            k_im2feature<T><<<grid2(w, h), 128, 0, st_>>>(pyr1_[k], c.f1, c.d5, swap);
            k_im2feature<T><<<grid2(w, h), 128, 0, branch(0, nb)>>>(pyr2_[k], c.f2, c.d5, swap);
        } else {
            k_copy<T><<<grid2(w, h, fc_), 128, 0, st_>>>(pyr1_[k], c.f1);
            k_copy<T><<<grid2(w, h, fc_), 128, 0, branch(0, nb)>>>(pyr2_[k], c.f2);
        }
        launches_ += 2;
        if (k != nlev_ - 1) {
            Img<T> su = view(u_, c.pw, c.ph, 1), sv = view(v_, c.pw, c.ph, 1);
            Img<T> du = view(u2_, w, h, 1), dv = view(v2_, w, h, 1);
            double rx = (double)w / c.pw, ry = (double)h / c.ph;
            T scale = (T)(1 / P.ratio);
            k_resize<T><<<grid2(w, h), 128, 0, branch(1, nb)>>>(su, du, rx, ry, scale, 1);
            k_resize<T><<<grid2(w, h), 128, 0, branch(2, nb)>>>(sv, dv, rx, ry, scale, 1);
            std::swap(u_, u2_);
            std::swap(v_, v2_);
            launches_ += 2;
        }
        join(nb);
        c.folded = fold_ && fused_tma_ && cp_level(w, h);
        const int nb2 = fork(1);               // the smoothed Im1 features (needs f1 only) next to the level-start warp
        if (bicubic_) {
            // warpImageBicubicRef differentiates the image it warps on every call (S/Image.h:2587-2595); the
            // Im2 features are constant within a level, so the three derivative images are computed once
            const double one[1] = {1.0};
            const Taps<T> id = make_taps<T>(one, 0);
            c.gix = view(g_ix_, w, h, fc_); c.giy = view(g_iy_, w, h, fc_); c.gixy = view(g_ixy_, w, h, fc_);
            filter_hv(c.f2, c.gix, c.d3, id);
            filter_hv(c.f2, c.giy, id, c.d3);
            filter_hv(c.f2, c.gixy, c.d3, c.d3);
        }
        if (k == nlev_ - 1) {
            PF_CUDA(cudaMemsetAsync(u_, 0, plane_bytes, st_));
            PF_CUDA(cudaMemsetAsync(v_, 0, plane_bytes, st_));
            if (!c.folded) {
                k_copy<T><<<grid2(w, h, fc_), 128, 0, st_>>>(c.f2, c.wf);
                launches_++;
            }
        } else if (!c.folded) {
            if (bicubic_) {
                bicubic_inner(k, 0);   // S/OpticalFlow.cpp:814-815: no threshold() at the level start
            } else {
                launch_chain(pdl_, k_update_warp<T>, warp_grid(w, h), dim3(128), 0, st_, c.f1, c.f2, c.wf, u_, v_, (const T*)nullptr, (const T*)nullptr, pitch, 0, 0x7fffffff);
                launches_++;
            }
        }
        // Im1 is constant within a level: its smoothed copy is computed once instead of every
        // outer iteration (S/OpticalFlow.cpp:89 recomputes it; same arithmetic, same result)
        set_phase(PF_T_PHASE1_GENERATE, k);
        filter_hv(c.f1, c.s1, c.g5, c.g5, branch(0, nb2));
        join(nb2);
        if (fused_ && fused_tma_) {
            const int ty = small_tiles(w, h) ? kFTYs : kFTY;
            c.fmaps.wf = make_image_map(c.wf.p, w, h, fc_, c.wf.pitch, c.wf.plane, 72, ty + 8);
            c.fmaps.s1 = make_image_map(c.s1.p, w, h, fc_, c.s1.pitch, c.s1.plane, 72, ty + 4);
            c.fmaps.u = make_plane_map(u_, w, h, pitch, 72, ty + 2);   // u_ / v_ are updated in place within a level
            c.fmaps.v = make_plane_map(v_, w, h, pitch, 72, ty + 2);
        }
        c.pw = w;
        c.ph = h;
    }

    // -- Phase1 (un-fused path only): getDxs (S/OpticalFlow.cpp:80-122), one kernel per reference step --
    void ph_getdxs(int k) {
        if (fused_) return;
        Ctx& c = cx_;
        set_phase(PF_T_PHASE1_GENERATE, k);
        filter_h(c.wf, c.tmp, c.g5);
        filter_v(c.tmp, c.s2, c.g5);
        k_blend_dt<T><<<grid2(c.w, c.h, fc_), 128, 0, st_>>>(c.s1, c.s2, c.blend, c.imdt);
        launches_++;
        filter_h(c.blend, c.imdx, c.d5);
        filter_v(c.blend, c.imdy, c.d5);
    }

    // -- Phases 1-4 of inner iteration hh: the linear system of S/OpticalFlow.cpp:295-448 --
    // rows [row_lo, row_hi) only (row-band split, multigpu.cuh: the rows this device's SOR band reads); the fused TMA
    // kernel computes whole tile rows, the other paths always the whole level
    void ph_assemble(int k, int hh, int row_lo = 0, int row_hi = -1) {
        Ctx& c = cx_;
        const int w = c.w, h = c.h, pitch = c.pitch;
        if (row_hi < 0 || row_hi > h) row_hi = h;
        row_lo = std::max(0, std::min(row_lo, row_hi));
        const T* cdu = hh > 0 ? du_ : nullptr;
        const T* cdv = hh > 0 ? dv_ : nullptr;
        if (fused_) {
            set_phase(PF_T_PHASE4_SYSTEM, k);
            FusedArgs<T> fa;
            fa.s1 = c.s1; fa.wf = c.wf;
            fa.u = u_; fa.v = v_; fa.du = cdu; fa.dv = cdv;
            fa.lap = (kF64 && lex_) ? d_lap_ : nullptr;
            fa.phi = phi_; fa.dxy = dxy_; fa.iu = iu_; fa.iv = iv_; fa.bu = bu_; fa.bv = bv_;
            fa.w = w; fa.h = h; fa.pitch = pitch;
            fa.alpha = (T)P.alpha; fa.omega = (T)1.8; fa.eps = c.eps;
            fa.g5 = c.g5; fa.d5 = c.d5;
            if (fused_tma_) {
                if (c.folded) {
                    if constexpr (!kF64) {
                        // the flow of this iteration is u + du of the previous solve (none before the first); it lands in u2 / v2
                        const bool inc = c.it > 0;
                        fa.f1 = c.f1; fa.f2 = c.f2;
                        fa.wdu = inc ? du_ : nullptr; fa.wdv = inc ? dv_ : nullptr;
                        fa.uo = u2_; fa.vo = v2_;
                        launch_chain(pdl_, k_fused_cp<T, kFTYs, kCpSeg, true>, dim3(ceil_div(w, 64), ceil_div(h, kFTYs)), dim3(64 * kCpSeg * fc_),
                                     fused_cp_smem_bytes<T, kFTYs>(fc_, true), st_, c.fmaps, fa);
                        if (inc) { std::swap(u_, u2_); std::swap(v_, v2_); }
                    }
                } else if (cp_level(w, h)) {
                    if constexpr (!kF64) {
                        fa.ty0 = row_lo / kFTYs;
                        launch_chain(pdl_, k_fused_cp<T, kFTYs, kCpSeg>, dim3(ceil_div(w, 64), ceil_div(row_hi, kFTYs) - fa.ty0), dim3(64 * kCpSeg * fc_),
                                     fused_cp_smem_bytes<T, kFTYs>(fc_), st_, c.fmaps, fa);
                    }
                } else if (small_tiles(w, h)) {
                    size_t smem = sizeof(FusedSmem<T, kFTYs>) + 128;
                    fa.ty0 = row_lo / kFTYs;
                    launch_chain(pdl_, k_fused_tma<T, kFTYs, kFSEGs>, dim3(ceil_div(w, 64), ceil_div(row_hi, kFTYs) - fa.ty0), dim3(64 * kFSEGs), smem, st_, c.fmaps, fa);
                } else {
                    size_t smem = sizeof(FusedSmem<T, kFTY>) + 128;
                    fa.ty0 = row_lo / kFTY;
                    launch_chain(pdl_, k_fused_tma<T, kFTY, kFSEG>, dim3(ceil_div(w, 64), ceil_div(row_hi, kFTY) - fa.ty0), dim3(64 * kFSEG), smem, st_, c.fmaps, fa);
                }
            } else {
                k_fused_assemble<T, kFTX, kFTY, kFSEG><<<dim3(ceil_div(w, kFTX), ceil_div(h, kFTY)), kFTX * kFSEG, 0, st_>>>(fa);
            }
            launches_++;
        } else {
            set_phase(PF_T_PHASE2_DERIVS, k);
            k_phi<T><<<grid2(w, h), 128, 0, st_>>>(u_, v_, cdu, cdv, phi_, w, h, pitch, c.eps);
            launches_++;
            set_phase(PF_T_PHASE4_SYSTEM, k);
            AssembleArgs<T> a;
            a.imdx = c.imdx; a.imdy = c.imdy; a.imdt = c.imdt;
            a.u = u_; a.v = v_; a.du = cdu; a.dv = cdv; a.phi = phi_;
            a.lap = (kF64 && lex_) ? d_lap_ : nullptr;
            a.gm = gmix_ ? d_gm_ : nullptr;
            a.dxy = dxy_; a.iu = iu_; a.iv = iv_; a.bu = bu_; a.bv = bv_;
            a.dx2 = nullptr; a.dy2 = nullptr;
            a.w = w; a.h = h; a.pitch = pitch;
            a.alpha = (T)P.alpha; a.omega = (T)1.8; a.eps = c.eps;
            k_assemble<T><<<grid2(w, h), 128, 0, st_>>>(a);
            launches_++;
        }
    }

    // Bicubic warp of the Im2 features into the warped-feature image of this level (interpolation == Bicubic)
    void bicubic_inner(int k, int clamp) {
        (void)k;
        Ctx& c = cx_;
        BicubicOut<T> bo;
        bo.hwc = nullptr; bo.planar = c.wf; bo.clamp = clamp;
        k_bicubic_warp<T><<<grid2(c.w, c.h), 128, 0, st_>>>(c.f1, c.f2, c.gix, c.giy, c.gixy, u_, v_, c.pitch, d_tab_, bo);
        launches_++;
    }

    // -- Phase5: SOR on this device alone --
    void ph_sor(int k) {
        set_phase(PF_T_PHASE5_SOR, k);
        run_sor(cx_.w, cx_.h, cx_.pitch, n_sor_at(k), k);
    }

    // -- Phase6: update + warp (+ noise estimate in the parity mode) --
    // warp_lo / warp_hi: rows whose warped features this device will read again (row-band split, multigpu.cuh); the
    // flow update always covers the whole level
    void ph_update(int k, int warp_lo = 0, int warp_hi = 0x7fffffff) {
        Ctx& c = cx_;
        set_phase(PF_T_PHASE6_UPDATE, k);
        if (c.folded) {
            // the next assembly updates and warps; after the last iteration of the level only the flow update is left
            if (c.it + 1 == n_outer_at(k)) {
                k_add_flow<T><<<grid2(c.w, c.h), 128, 0, st_>>>(u_, v_, du_, dv_, c.w, c.pitch);
                launches_++;
            }
            return;
        }
        if (bicubic_ || gmix_ || (kF64 && lex_)) { warp_lo = 0; warp_hi = 0x7fffffff; }   // these read the whole warped image
        if (bicubic_) {
            k_add_flow<T><<<grid2(c.w, c.h), 128, 0, st_>>>(u_, v_, du_, dv_, c.w, c.pitch);
            launches_++;
            bicubic_inner(k, 1);   // S/OpticalFlow.cpp:517-521: warpImageBicubicRef + threshold()
        } else {
            launch_chain(pdl_, k_update_warp<T>, warp_grid(c.w, c.h), dim3(128), 0, st_, c.f1, c.f2, c.wf, u_, v_, (const T*)du_, (const T*)dv_, c.pitch, warp_lo, warp_hi);
            launches_++;
        }
        if (gmix_) {   // S/OpticalFlow.cpp:524-527
            for (int it = 0; it < 3; it++) {
                k_gm_accum<T><<<dim3(std::min(8, ceil_div(c.w, 128)), std::min(c.h, 64), fc_), 128, 0, st_>>>(c.f1, c.wf, d_gm_, d_acc_);
                k_gm_update<<<1, 32, 0, st_>>>(d_gm_, d_acc_, fc_);
                launches_ += 2;
            }
        } else if (kF64 && lex_) {
            k_noise_accum<T><<<dim3(std::min(8, ceil_div(c.w, 128)), std::min(c.h, 64), fc_), 128, 0, st_>>>(c.f1, c.wf, d_acc_);
            k_noise_final<<<1, 32, 0, st_>>>(d_acc_, d_lap_, fc_);
            launches_ += 2;
        }
    }

    // -- PostProcessing: bicubic warp of the original frame + clamp; export the flow --
    void ph_end() {
        Ctx& c = cx_;
        set_phase(PF_T_POST, 0);
        const Img<T>& im1 = pyr1_[0];
        const Img<T>& im2 = pyr2_[0];
        Img<T> ix = view(b_ix_, P.w, P.h, P.c), iy = view(b_iy_, P.w, P.h, P.c), ixy = view(b_ixy_, P.w, P.h, P.c);
        const double one[1] = {1.0};
        const Taps<T> id = make_taps<T>(one, 0);
        filter_hv(im2, ix, c.d3, id);
        filter_hv(im2, iy, id, c.d3);
        filter_hv(im2, ixy, c.d3, c.d3);
        BicubicOut<T> bo;
        bo.hwc = d_warp_; bo.clamp = 1;
        k_bicubic_warp<T><<<grid2(P.w, P.h), 128, 0, st_>>>(im1, im2, ix, iy, ixy, u_, v_, pitch_for(P.w), d_tab_, bo);
        Img<T> uo = view(u_, P.w, P.h, 1), vo = view(v_, P.w, P.h, 1);
        k_export_hwc<T><<<grid2(P.w, P.h), 128, 0, st_>>>(uo, d_vx_);
        k_export_hwc<T><<<grid2(P.w, P.h), 128, 0, st_>>>(vo, d_vy_);
        launches_ += 3;
        ph_restore();
    }

    // restore the pointer roles so that a captured graph and a later eager run agree
    void ph_restore() {
        Ctx& c = cx_;
        PF_CHECK_LAUNCH();
        du_ = c.du_e; dv_ = c.dv_e; du2_ = c.du2_e; dv2_ = c.dv2_e;
        u_ = c.u_e; v_ = c.v_e; u2_ = c.u2_e; v2_ = c.v2_e;
    }

    void ph_levels() {
        for (int k = nlev_ - 1; k >= 0; k--) {
            ph_level(k);
            for (int it = 0; it < n_outer_at(k); it++) {
                cx_.it = it;
                ph_getdxs(k);
                for (int hh = 0; hh < P.n_inner; hh++) {
                    ph_assemble(k, hh);
                    ph_sor(k);
                }
                ph_update(k);
            }
        }
    }

    // ---- sequence mode: consecutive pairs share a frame, so the pyramid
    //      of frame t+1 built for pair t is pair t+1's "Im1" pyramid; frames arrive as uint8, the flow
    //      leaves as interleaved float32.  Same arithmetic as the pairwise path (bit-identical flow).
    void seq_first(const unsigned char* frame) override {
        PF_CUDA(cudaSetDevice(P.device));
        size_t n = (size_t)P.h * P.w * P.c;
        PF_CUDA(cudaMemcpyAsync(d_in1_, frame, n, cudaMemcpyHostToDevice, st_));
        k_import_u8<T><<<grid2(P.w, P.h), 128, 0, st_>>>(reinterpret_cast<const unsigned char*>(d_in1_), pyr2_[0]);
        ph_pyramid(1);
        PF_CHECK_LAUNCH();
    }

    void seq_next(const unsigned char* frame, void* out, int format) override {
        if (format < 0 || format >= kSeqFormats) throw Error(PF_EINVAL, "unknown sequence output format");
        PF_CUDA(cudaSetDevice(P.device));
        std::swap(pyr1_, pyr2_);            // the newer frame of the previous pair becomes Im1
        seq_parity_ ^= 1;
        size_t n = (size_t)P.h * P.w * P.c;
        PF_CUDA(cudaMemcpyAsync(d_in1_, frame, n, cudaMemcpyHostToDevice, st_));
        if (use_graph_) {
            cudaGraphExec_t& ge = gexec_seq_[seq_parity_][format];
            if (!ge) {
                cudaGraph_t g = nullptr;
                PF_CUDA(cudaStreamBeginCapture(st_, cudaStreamCaptureModeThreadLocal));
                try {
                    enqueue_seq_pair(format);
                } catch (...) {
                    cudaStreamEndCapture(st_, &g);
                    if (g) cudaGraphDestroy(g);
                    throw;
                }
                PF_CUDA(cudaStreamEndCapture(st_, &g));
                PF_CUDA(cudaGraphInstantiate(&ge, g, 0));
                cudaGraphDestroy(g);
            }
            PF_CUDA(cudaGraphLaunch(ge, st_));
        } else {
            enqueue_seq_pair(format);
        }
        PF_CUDA(cudaMemcpyAsync(out, d_vx_, (size_t)P.h * P.w * seq_bytes_per_pixel(format), cudaMemcpyDeviceToHost, st_));
        sync_stream();
    }

  private:
    static size_t seq_bytes_per_pixel(int format) {
        return format == PF_SEQ_FLOW_U16 ? sizeof(ushort2) : format == PF_SEQ_FLOW_BGR8 ? 3 : sizeof(float2);
    }
    void enqueue_seq_pair(int format) {
        ph_ctx();
        k_import_u8<T><<<grid2(P.w, P.h), 128, 0, st_>>>(reinterpret_cast<const unsigned char*>(d_in1_), pyr2_[0]);
        launches_++;
        ph_pyramid(1);
        ph_lap_init();
        ph_levels();
        if (format == PF_SEQ_FLOW_BGR8) {
            const int pitch = pitch_for(P.w);
            k_minmax_init<<<1, 1, 0, st_>>>(d_mm_);
            k_flow_mag_minmax<T><<<dim3(ceil_div(P.w, 256), std::min(P.h, 256)), 256, 0, st_>>>(u_, v_, pitch, 1, P.w, P.h, d_mm_);
            k_flow_to_bgr<T><<<grid2(P.w, P.h), 128, 0, st_>>>(u_, v_, pitch, 1, P.w, d_mm_, reinterpret_cast<unsigned char*>(d_vx_));
            launches_ += 2;
        } else if (format == PF_SEQ_FLOW_U16)
            k_export_flow_u16<T><<<grid2(P.w, P.h), 128, 0, st_>>>(u_, v_, pitch_for(P.w), P.w, reinterpret_cast<ushort2*>(d_vx_));
        else
            k_export_flow_f32<T><<<grid2(P.w, P.h), 128, 0, st_>>>(u_, v_, pitch_for(P.w), P.w, reinterpret_cast<float2*>(d_vx_));
        launches_++;
        ph_restore();
    }

  public:

    // buffers the cooperative (row-band) SOR of MultiPlan works on
    struct SorView {
        SorArgs<T> args;
        T **du, **dv, **du2, **dv2;
        SorRunner<T>* runner;
    };
    SorView sor_view() {
        SorView v;
        v.args.phi = phi_; v.args.dxy = dxy_; v.args.iu = iu_; v.args.iv = iv_; v.args.bu = bu_; v.args.bv = bv_;
        v.args.du = nullptr; v.args.dv = nullptr; v.args.du_in = nullptr; v.args.dv_in = nullptr;
        v.args.w = cx_.w; v.args.h = cx_.h; v.args.pitch = cx_.pitch;
        v.args.alpha = (T)P.alpha; v.args.omega = (T)1.8;
        v.du = &du_; v.dv = &dv_; v.du2 = &du2_; v.dv2 = &dv2_;
        v.runner = &sor_;
        return v;
    }

  private:
    void enqueue_solve() {
        ph_begin();
        ph_levels();
        ph_end();
    }

  public:
    Params P;

  private:
#ifndef PF_FUSED_TY
#define PF_FUSED_TY 16
#define PF_FUSED_SEG 4
#endif
    static constexpr int kFTX = 64, kFTY = kF64 ? 16 : PF_FUSED_TY, kFSEG = kF64 ? 8 : PF_FUSED_SEG;   // k_fused_* tile, 64*SEG threads
    // Latency-tuned plans: levels whose 64x16 tiling leaves most SMs idle use 64x4 tiles instead.  A tile takes ~28 us
    // whatever the level (five channels through three stencil stages, one after the other), so the coarse levels are
    // bound by the latency of ONE tile; a 64x4 tile has half the rows to smooth (4+8 instead of 16+8) and a quarter of
    // the centre rows.  More halo work in total, which is why throughput-tuned plans keep 64x16.  Same arithmetic per pixel.
#ifndef PF_FUSED_TYS
#define PF_FUSED_TYS 4
#define PF_FUSED_SEGS 4
#endif
    static constexpr int kFTYs = kF64 ? 8 : PF_FUSED_TYS, kFSEGs = kF64 ? 8 : PF_FUSED_SEGS;
    // the channel-parallel kernel holds two CTAs of C x 128 threads per SM: worth it while all its tiles are resident at once
    // (the 455-px level of the 1920 pyramid has 512 of them: 0.40 ms per level with it, 0.28 ms without)
    bool cp_level(int w, int h) const { return fused_cp_ && small_tiles(w, h) && ceil_div(w, 64) * ceil_div(h, kFTYs) <= 2 * 148; }
    bool small_tiles(int w, int h) const {
        if (const char* e = getenv("PF_FUSED_SMALL")) return atoi(e) != 0 && ceil_div(w, 64) * ceil_div(h, kFTY) <= 148;
        return P.tune == PF_TUNE_LATENCY && ceil_div(w, 64) * ceil_div(h, kFTY) <= 148;
    }
    bool lex_ = false, use_graph_ = true, fused_ = true, fused_tma_ = true, profiling_ = false, open_ = false;
    bool bicubic_ = false, gmix_ = false;   // alternative solver branches
    bool pdl_ = false;
    static constexpr int kAux = 3;
    cudaStream_t aux_[kAux] = {nullptr, nullptr, nullptr};   // branches of one pair's launch sequence (fork / join)
    cudaEvent_t ev_fork_ = nullptr, ev_join_[kAux] = {nullptr, nullptr, nullptr};
    bool branches_ = true;
    bool fold_ = false;                      // k_fused_cp<..., WARP>: update + warp inside the assembly of the small-tile levels
    bool fused_cp_ = false;                  // channel-parallel assembly kernel on the small-tile levels (fused_cp.cuh)
    static constexpr int kCpSeg = 2;
    int nlev_ = 0, fc_ = 0;
    SorRunner<T> sor_;
    std::vector<Level> geo_;
    Arena arena_;
    cudaStream_t st_ = nullptr;
    cudaEvent_t ev_[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_block_ = nullptr;                 // blocking-sync event of sync_stream()
    bool block_wait_ = default_block_waits();
    cudaGraph_t graph_ = nullptr;
    cudaGraphExec_t gexec_ = nullptr;
    static constexpr int kSeqFormats = 3;
    cudaGraphExec_t gexec_seq_[2][kSeqFormats] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    unsigned int* d_mm_ = nullptr;
    int seq_parity_ = 0;
    std::vector<Span> spans_;
    std::vector<double> level_ms_;
    size_t span_used_ = 0;
    long long launches_ = 0, sor_launches_ = 0, sor_launches_l0_ = 0;
    double *d_in1_ = nullptr, *d_in2_ = nullptr, *d_warp_ = nullptr, *d_vx_ = nullptr, *d_vy_ = nullptr;
    double* d_gm_ = nullptr;   // Gaussian-mixture parameters (noiseModel == GMixture)
    std::vector<Img<T>> pyr1_, pyr2_;
    T *b_tmp_ = nullptr, *b_out_ = nullptr, *b_ix_ = nullptr, *b_iy_ = nullptr, *b_ixy_ = nullptr;
    T *g_ix_ = nullptr, *g_iy_ = nullptr, *g_ixy_ = nullptr;
    T *f1_ = nullptr, *f2_ = nullptr, *wf_ = nullptr, *s1_ = nullptr, *s2_ = nullptr, *tmp_ = nullptr;
    T *blend_ = nullptr, *imdx_ = nullptr, *imdy_ = nullptr, *imdt_ = nullptr;
    T *u_ = nullptr, *v_ = nullptr, *u2_ = nullptr, *v2_ = nullptr, *du_ = nullptr, *dv_ = nullptr;
    T *du2_ = nullptr, *dv2_ = nullptr, *phi_ = nullptr, *dxy_ = nullptr, *iu_ = nullptr, *iv_ = nullptr;
    T *bu_ = nullptr, *bv_ = nullptr, *dx2_ = nullptr, *dy2_ = nullptr;
    double *d_lap_ = nullptr, *d_acc_ = nullptr;
    BicubicTable* d_tab_ = nullptr;
};

}  // namespace pf
