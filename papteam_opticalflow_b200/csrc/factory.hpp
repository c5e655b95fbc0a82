// Per-precision factories; each is defined in its own translation unit so that the FP64
// instantiation can be compiled with -fmad=false (bit-level agreement with the reference) while the
// FP32 instantiation keeps fused multiply-adds.
#pragma once
#include "common.cuh"

namespace pf {
struct PlanBase;

struct StageCalls {
    void (*pyramid)(double*, const double*, int, int, int, double, int);
    int (*im2feature)(double*, const double*, int, int, int, int);
    void (*getdxs)(double*, double*, double*, const double*, const double*, int, int, int);
    void (*warpfl)(double*, const double*, const double*, const double*, const double*, int, int, int);
    void (*resize_to)(double*, const double*, int, int, int, int, int, double);
    void (*bicubic)(double*, const double*, const double*, const double*, const double*, int, int, int);
    void (*assemble)(double*, double*, double*, double*, double*, double*, const double*, const double*,
                     const double*, const double*, const double*, const double*, const double*,
                     const double*, double, int, int, int);
    void (*sor)(double*, double*, const double*, const double*, const double*, const double*,
                const double*, const double*, double, int, int, int, int, int, int, double*, double*);
};

// one pair over several GPUs (row-band split of the SOR solve); FP32 red-black only.
// stats (4 doubles, may be NULL): solve ms, halo bytes pulled, gather bytes pulled, solves that were split
void multigpu_flow_f32(double* vx, double* vy, double* warp, const double* im1, const double* im2, const Params& p,
                       const int* devices, int ndev, long long split_min_pixels, double* stats);
PlanBase* make_plan_f32(const Params& p);
PlanBase* make_plan_f64(const Params& p);
const StageCalls& stages_f32();
const StageCalls& stages_f64();
}  // namespace pf
