// float instantiation of the solver (fast mode).
#include "factory.hpp"
#include "stages.cuh"
#include "multigpu.cuh"
#include <mutex>

namespace pf {
void multigpu_flow_f32(double* vx, double* vy, double* warp, const double* im1, const double* im2, const Params& p,
                       const int* devices, int ndev, long long split_min_pixels, double* stats) {
    // one cached MultiPlan (arena per device) reused while the request keeps the same shape
    static std::mutex mu;
    static std::unique_ptr<MultiPlan> cached;
    std::lock_guard<std::mutex> lock(mu);
    if (!cached || !cached->matches(p, devices, ndev, split_min_pixels)) {
        cached.reset();
        cached.reset(new MultiPlan(p, devices, ndev, split_min_pixels));
    }
    try {
        cached->execute(vx, vy, warp, im1, im2, stats);
    } catch (...) {
        cached.reset();      // never reuse a plan whose solve failed (sticky context errors, lost neighbours)
        throw;
    }
}
PlanBase* make_plan_f32(const Params& p) { return new Plan<float>(p); }
const StageCalls& stages_f32() {
    typedef Stages<float> S;
    static const StageCalls c = {S::pyramid, S::im2feature, S::getdxs, S::warpfl, S::resize_to,
                                 S::bicubic, S::assemble, S::sor};
    return c;
}
}  // namespace pf
