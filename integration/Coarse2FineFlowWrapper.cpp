// Native-level integration: the body of the reference's wrapper
// P/Coarse2FineFlowWrapper.cpp:14-51 replaced by a shim over the C ABI of libpyflow_b200.so.  The reference's
// Cython module Par/pyflow.pyx, its extern declaration Par/coarse2Fine.pxd:8-12 and the wrapper header
// P/Coarse2FineFlowWrapper.h:12-15 are compiled UNCHANGED, where they lie, against this one file
// (integration/Makefile); nothing of the reference's solver (OpticalFlow.cpp, GaussianPyramid.cpp, ...) is linked.
//
// Errors: the reference lets none cross this boundary (it prints and returns, S/Image.h:1749-1753) and the untouched
// .pxd declares no `except +`, so a C++ exception here would terminate the interpreter.  The shim therefore reports a
// failure the reference's way -- a line on stderr, zeroed outputs -- plus an "error" entry in the returned timing map.
// Arithmetic is chosen by PYFLOW_B200_MODE (fp32_redblack default, fp64_wavefront = the reference to 1e-6), the GPU by
// PYFLOW_B200_DEVICE.
#include "Coarse2FineFlowWrapper.h"   // the reference's own header (-I<reference>/Code/Parallel/src)
#include "pyflow_b200.h"              // this repository: include/

#include <cstdlib>
#include <cstring>
#include <string>

map<string, string> Coarse2FineFlowWrapper(double* vx, double* vy, double* warpI2, const double* Im1, const double* Im2,
                                           int pyramidLevels, int nCores, int h, int w, int c) {
    int mode = PF_MODE_FP32_REDBLACK;
    if (const char* e = getenv("PYFLOW_B200_MODE")) {
        if (!strcmp(e, "fp64_wavefront")) mode = PF_MODE_FP64_WAVEFRONT;
        else if (!strcmp(e, "fp64_redblack")) mode = PF_MODE_FP64_REDBLACK;
        else if (!strcmp(e, "fp32_wavefront")) mode = PF_MODE_FP32_WAVEFRONT;
    }
    const char* d = getenv("PYFLOW_B200_DEVICE");
    double t[PF_NUM_TIMINGS] = {0};
    map<string, string> timing;
    const int rc = pf_coarse2fine_flow_levels(vx, vy, warpI2, Im1, Im2, pyramidLevels, nCores, h, w, c, mode, d ? atoi(d) : 0, t);
    if (rc != PF_OK) {
        cerr << "pyflow_b200: " << pf_last_error() << endl;
        timing["error"] = pf_last_error();
    }
    // the keys of the reference's timing map (S/OpticalFlow.cpp:850-860; the Parallel build fills only the first,
    // P/OpticalFlow.cpp:939), seconds as decimal strings
    static const char* const keys[] = {"Total C++ Execution", "Construction", "Allocation", "Phase1_Generate", "Phase2_Derivatives",
                                       "Phase3_PsiData", "Phase4_LinearSystem", "Phase5_SOR", "Phase6_Update", "PostProcessing"};
    for (int i = 0; i < 10; i++) timing[keys[i]] = to_string(t[i] / 1000.0);
    return timing;
}
