// float instantiation of the solver (fast mode).
#include "factory.hpp"
#include "stages.cuh"

namespace pf {
PlanBase* make_plan_f32(const Params& p) { return new Plan<float>(p); }
const StageCalls& stages_f32() {
    typedef Stages<float> S;
    static const StageCalls c = {S::pyramid, S::im2feature, S::getdxs, S::warpfl, S::resize_to,
                                 S::bicubic, S::assemble, S::sor};
    return c;
}
}  // namespace pf
