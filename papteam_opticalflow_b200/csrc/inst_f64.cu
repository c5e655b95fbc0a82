// double instantiation of the solver (compiled with -fmad=false: parity mode).
#include "factory.hpp"
#include "stages.cuh"

namespace pf {
PlanBase* make_plan_f64(const Params& p) { return new Plan<double>(p); }
const StageCalls& stages_f64() {
    typedef Stages<double> S;
    static const StageCalls c = {S::pyramid, S::im2feature, S::getdxs, S::warpfl, S::resize_to,
                                 S::bicubic, S::assemble, S::sor};
    return c;
}
}  // namespace pf
