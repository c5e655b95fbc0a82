// Shared host/device definitions for the B200 optical-flow solver.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <stdexcept>
#include <vector>
#include <utility>

#include "../../include/pyflow_b200.h"

namespace pf {

// ---- error plumbing: CUDA failures become exceptions, the C ABI turns them into PF_ECUDA -----
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define PF_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            throw pf::Error(_e == cudaErrorMemoryAllocation ? PF_ENOMEM : PF_ECUDA,           \
                            std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +       \
                                __FILE__ + ":" + std::to_string(__LINE__) + ")");             \
    } while (0)

#define PF_CHECK_LAUNCH() PF_CUDA(cudaGetLastError())

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---- planar image view --------------------------------------------------------------------------
// HBM layout: channel-planar (SoA), rows padded to a pitch of 32 elements (128 B in FP32, 256 B in
// FP64) so every row starts on a cache-line boundary and float4/double2 accesses stay aligned.
// The reference's HWC interleaving (S/Image.h:36-461) exists only at the import/export kernels.
template <typename T>
struct Img {
    T* p = nullptr;
    int w = 0, h = 0, c = 0;
    int pitch = 0;       // elements per row
    size_t plane = 0;    // elements per channel plane
    __host__ __device__ T* ch(int k) const { return p + (size_t)k * plane; }
    __host__ __device__ size_t elems() const { return plane * (size_t)c; }
};

constexpr int kPitchAlign = 32;
inline int pitch_for(int w) { return (int)round_up((size_t)w, kPitchAlign); }
inline size_t plane_for(int w, int h) { return round_up((size_t)pitch_for(w) * h, 64); }

// up to 2*8+1 filter taps passed by value in kernel parameters
constexpr int kMaxHalf = 8;
template <typename T>
struct Taps {
    T v[2 * kMaxHalf + 1];
    int half;
};

// solver parameters after the boundary has normalised both call shapes
// Alternative solver branches.  In the reference they are two process-global public
// statics, OpticalFlow::interpolation / OpticalFlow::noiseModel (S/OpticalFlow.h:19-27, defaults Bilinear / Lap
// at S/OpticalFlow.cpp:33-34); here the same process-global switch (pf_set_solver_variant) is sampled whenever
// solver parameters are put together, so a plan keeps the variant it was created under.
struct SolverVariant {
    int interp = PF_INTERP_BILINEAR;
    int noise = PF_NOISE_LAP;
};
inline SolverVariant& solver_variant() {
    static SolverVariant v;
    return v;
}

struct Params {
    int h, w, c;
    double alpha, ratio;
    int min_width, levels;   // levels > 0 wins
    int n_outer, n_inner, n_sor;
    int col_type;
    int mode, device;
    int interp = solver_variant().interp;   // PF_INTERP_*: warp inside the pyramid loop
    int noise = solver_variant().noise;     // PF_NOISE_*: data-term weight model
    int tune = PF_TUNE_THROUGHPUT;          // PF_TUNE_*: what the launch schedule of the SOR solves is fitted to (same results)
};
inline bool same_solver(const Params& a, const Params& b) {
    return a.h == b.h && a.w == b.w && a.c == b.c && a.alpha == b.alpha && a.ratio == b.ratio && a.min_width == b.min_width &&
           a.levels == b.levels && a.n_outer == b.n_outer && a.n_inner == b.n_inner && a.n_sor == b.n_sor &&
           a.col_type == b.col_type && a.mode == b.mode && a.interp == b.interp && a.noise == b.noise && a.tune == b.tune;
}

// How host threads wait for the GPU.  cudaStreamSynchronize spins (one context per process => the driver's heuristic
// chooses spinning); a node runs one process per GPU with several waiting host threads each (batch workers, stager
// threads), and 8 ranks x 8 spinning workers on 32 cores starve the threads that have copies to issue.  Blocking waits
// (an event created with cudaEventBlockingSync) cost a few tens of microseconds of wake-up latency per wait.
//   PF_WAIT=spin | block | auto (default): auto blocks when several processes share the node (LOCAL_WORLD_SIZE > 1);
//   batch / sequence workers always block.
inline bool default_block_waits() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PF_WAIT");
        if (e && !strcmp(e, "spin")) v = 0;
        else if (e && !strcmp(e, "block")) v = 1;
        else {
            const char* lw = getenv("LOCAL_WORLD_SIZE");
            v = (lw && atoi(lw) > 1) ? 1 : 0;
        }
    }
    return v != 0;
}
// wait for everything enqueued on `st` so far without burning a core
inline void stream_wait_blocking(cudaStream_t st) {
    cudaEvent_t ev = nullptr;
    PF_CUDA(cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ev, st);
    if (e == cudaSuccess) e = cudaEventSynchronize(ev);
    cudaEventDestroy(ev);
    PF_CUDA(e);
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// One pair alone is a chain of ~1 450 dependent kernels, most of them shorter than 20 us, and every kernel boundary costs
// the drain of the previous grid plus the launch of the next one.  Kernels that start with pdl_trigger() / pdl_wait() and
// are launched through launch_chain() with `pdl` set let the NEXT kernel of the stream be launched, scheduled onto the SMs
// the previous grid has already left and run its prologue (barrier initialisation, index arithmetic) while the previous
// grid is still finishing; pdl_wait() returns when the previous grid has completed and its memory is visible, so every
// global access stays after it.  Launched the ordinary way the two instructions do nothing.  Only latency-tuned plans would
// use it: with many pairs in flight a waiting grid would hold SMs that another pair's kernels could use.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// MEASURED AND OFF BY DEFAULT (PF_PDL=1 enables it): graph replay of a 1920x1080 pair 17.59 ms without, 17.92 ms with it;
// 480x270: 6.32 / 6.43 ms (tools/env_ab.py PF_PDL 0 1, identical results).  Inside a captured graph a kernel boundary already costs
// well under a microsecond, and an early-launched grid only adds CTAs that wait on the SMs.
inline bool pdl_allowed() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PF_PDL");
        v = (e && atoi(e)) ? 1 : 0;
    }
    return v != 0;
}

template <typename... KArgs, typename... Args>
inline void launch_chain(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl && pdl_allowed()) ? 1 : 0;
    PF_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}

inline bool mode_is_fp64(int mode) { return mode == PF_MODE_FP64_WAVEFRONT || mode == PF_MODE_FP64_REDBLACK; }
inline bool mode_is_lex(int mode) { return mode == PF_MODE_FP64_WAVEFRONT || mode == PF_MODE_FP32_WAVEFRONT; }
// does a pyramid level of this width run its SOR solves in the reference's lexicographic order?
inline bool mode_lex_at(int mode, int level_width) {
    return mode_is_lex(mode) || (mode == PF_MODE_FP32_HYBRID && level_width <= PF_HYBRID_MAX_WIDTH);
}

}  // namespace pf
