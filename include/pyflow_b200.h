/* pyflow_b200 -- C ABI of the B200-native coarse-to-fine variational optical-flow solver.
 *
 * This is the drop-in boundary for the hot path of ElijahHyndman/PAPTeam_OpticalFlow: every
 * function here replaces (or is a finer-grained view of) one reference interface, cited per
 * entry as <tree>/<file>:<lines> with S/ = Code/Serial/src, P/ = Code/Parallel/src,
 * Par/ = Code/Parallel.  Plain pointers and sizes only; all image arguments are HOST pointers to
 * row-major HWC interleaved float64, exactly the buffers the reference's wrapper receives from
 * numpy (P/Coarse2FineFlowWrapper.cpp:14-51).  The library copies in and out and retains nothing.
 *
 * Every function returns 0 on success or a negative PF_E* code; pf_last_error() describes the
 * failure of the calling thread.  There is NO CPU fallback: without a usable CUDA device every
 * compute entry point fails with PF_ENODEVICE.
 */
#ifndef PYFLOW_B200_H
#define PYFLOW_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------ */
#define PF_OK            0
#define PF_EINVAL       -1   /* bad argument (NULL pointer, non-positive size, unknown mode ...) */
#define PF_ENODEVICE    -2   /* no CUDA device / device index out of range */
#define PF_ECUDA        -3   /* a CUDA runtime call or kernel failed (message has the details) */
#define PF_ENOMEM       -4   /* host or device allocation failed */
#define PF_EUNSUPPORTED -5   /* valid request this build does not implement */

/* ---- solver modes (north_star: two required modes, plus the two cross combinations that the
 *      tests use to separate ordering error from rounding error) ---------------------------- */
#define PF_MODE_FP64_WAVEFRONT 0 /* FP64, lexicographic Gauss-Seidel as a pipelined anti-diagonal
                                    wavefront: matches the reference to <=1e-6 (bit-level in practice) */
#define PF_MODE_FP32_REDBLACK  1 /* FP32, red-black SOR with fused sweeps: the fast mode */
#define PF_MODE_FP64_REDBLACK  2 /* FP64 arithmetic, red-black ordering */
#define PF_MODE_FP32_WAVEFRONT 3 /* FP32 arithmetic, lexicographic ordering */
#define PF_MODE_FP32_HYBRID    4 /* the fast mode with the reference's lexicographic order on the coarse pyramid levels
                                    (width <= PF_HYBRID_MAX_WIDTH) and red-black above: an ordering difference on a coarse,
                                    under-iterated level is multiplied by 1/ratio per level on the way up, and where the flow
                                    leaves the frame (weak data term) that alone moves single pixels by more than 0.5 px
                                    (BASELINE config 5); with this mode they stay within 0.2 px */
#define PF_HYBRID_MAX_WIDTH    400

/* ---- alternative solver branches ----------------------------------------
 * The reference selects them through two process-global PUBLIC STATIC members,
 *   OpticalFlow::interpolation {Bilinear, Bicubic}   S/OpticalFlow.h:19-20, default Bilinear S/OpticalFlow.cpp:33
 *   OpticalFlow::noiseModel    {GMixture, Lap}       S/OpticalFlow.h:21-27, default Lap      S/OpticalFlow.cpp:34
 * (a user switches them by assigning the statics before the call).  The values below follow the
 * reference's enum order.  Bicubic: the warp inside the pyramid loop is warpImageBicubicRef, followed
 * by threshold() inside SmoothFlowSOR (S/OpticalFlow.cpp:517-521, 814-815).  GMixture: the data-term
 * weight is the two-component Gaussian mixture of S/OpticalFlow.cpp:359-368 / 389-396 with its
 * parameters re-estimated by three EM steps after every outer iteration (estGaussianMixture, :539-591). */
#define PF_INTERP_BILINEAR 0
#define PF_INTERP_BICUBIC  1
#define PF_NOISE_GMIXTURE  0
#define PF_NOISE_LAP       1

/* ---- schedule tuning: how many red-black sweeps one launch of the SOR kernel fuses is chosen per pyramid level from a
 *      cost model with two fits (the result is bit-identical either way: the red-black update does not depend on the
 *      tiling).  THROUGHPUT minimises SM time -- right when many pairs are in flight (batches, sequences, explicit
 *      plans run concurrently); LATENCY minimises the time of one pair alone -- used by the one-shot entry points.
 *      Environment PF_SOR_TUNE=latency|throughput forces one fit everywhere. */
#define PF_TUNE_THROUGHPUT 0
#define PF_TUNE_LATENCY    1

/* ---- timing report: the reference's timing-map keys (S/OpticalFlow.cpp:850-860), in
 *      milliseconds of GPU time from CUDA events, plus the transfer legs ---------------------- */
enum {
    PF_T_TOTAL = 0,        /* "Total C++ Execution": H2D + solve + D2H                      */
    PF_T_CONSTRUCTION,     /* "Construction": both Gaussian pyramids                         */
    PF_T_ALLOCATION,       /* "Allocation": im2feature, flow upsampling, per-level warp      */
    PF_T_PHASE1_GENERATE,  /* "Phase1_Generate": getDxs                                      */
    PF_T_PHASE2_DERIVS,    /* "Phase2_Derivatives": flow derivatives / phi                   */
    PF_T_PHASE3_PSI,       /* "Phase3_PsiData" (fused into PHASE4 on the GPU; reported 0)    */
    PF_T_PHASE4_SYSTEM,    /* "Phase4_LinearSystem": psi, collapsed products, Laplacian, rhs */
    PF_T_PHASE5_SOR,       /* "Phase5_SOR": the SOR sweeps                                   */
    PF_T_PHASE6_UPDATE,    /* "Phase6_Update": u+=du, warp, noise estimate                   */
    PF_T_POST,             /* "PostProcessing": bicubic warp + clamp                         */
    PF_T_H2D,              /* host->device copy of the two frames                            */
    PF_T_D2H,              /* device->host copy of vx, vy, warpI2                            */
    PF_T_SOLVE,            /* device-only solve (no transfers)                               */
    PF_NUM_TIMINGS = 16
};

typedef struct pf_plan pf_plan;

/* ---- library / device ---------------------------------------------------------------------- */
const char* pf_last_error(void);
const char* pf_version(void);
int  pf_device_count(void);                       /* number of CUDA devices, 0 if none */
void* pf_host_alloc(size_t bytes);                /* pinned host memory (NULL on failure) */
void  pf_host_free(void* p);

/* Process-global solver variant, like the reference's statics (see PF_INTERP_* / PF_NOISE_* above).
 * It is sampled when a plan is created or a one-shot / batch / sequence call starts; existing plans keep
 * the variant they were created under, and the plan pool is keyed by it.  Host-only, usable without a GPU. */
int pf_set_solver_variant(int interpolation, int noise_model);
int pf_get_solver_variant(int* interpolation, int* noise_model);

/* ---- pyramid geometry: host-only double arithmetic, usable without a GPU --------------------
 * S/GaussianPyramid.cpp:50-53 (level count from minWidth) and :89-106 + S/Image.h:755-756
 * (level sizes by double->int truncation). */
int pf_pyramid_levels(int width, double ratio, int minWidth);
int pf_level_geometry(int w, int h, double ratio, int levels, int* widths, int* heights);

/* ---- one-shot solves --------------------------------------------------------------------------
 * Upstream pyflow call shape named by north_star (parameter list survives at Par/pyflow.pyx:36-41;
 * declared at S/OpticalFlow.h:49-50).  timings may be NULL or point to PF_NUM_TIMINGS doubles. */
int pf_coarse2fine_flow(double* vx, double* vy, double* warpI2, const double* im1,
                        const double* im2, double alpha, double ratio, int minWidth,
                        int nOuterFPIterations, int nInnerFPIterations, int nSORIterations,
                        int colType, int h, int w, int c, int mode, int device, double* timings);

/* The fork's call shape: replaces Coarse2FineFlowWrapper (P/Coarse2FineFlowWrapper.h:12-15,
 * P/Coarse2FineFlowWrapper.cpp:14-51) with its hard-coded alpha=0.012 ratio=0.75 7/1/30
 * (S/OpticalFlow.cpp:747-751) and colType=0.  nCores is accepted and ignored. */
int pf_coarse2fine_flow_levels(double* vx, double* vy, double* warpI2, const double* im1,
                               const double* im2, int pyramidLevels, int nCores, int h, int w,
                               int c, int mode, int device, double* timings);

/* The one-shot and batch entry points keep up to 16 idle plans (device arena + CUDA graph) in a
 * pool keyed by every parameter, so repeated calls pay neither allocation nor graph capture;
 * pf_pool_clear destroys the idle ones and returns how many it freed. */
int pf_pool_clear(void);

/* ---- plans: device arena + captured CUDA graph for one (h, w, c, parameters, mode, device) ---
 * levels > 0 selects the fork's explicit level count, otherwise minWidth decides. */
int pf_plan_create(pf_plan** plan, int h, int w, int c, double alpha, double ratio, int minWidth,
                   int levels, int nOuterFPIterations, int nInnerFPIterations, int nSORIterations,
                   int colType, int mode, int device);
/* pf_plan_create uses PF_TUNE_THROUGHPUT; this one takes the tuning explicitly. */
int pf_plan_create_tuned(pf_plan** plan, int h, int w, int c, double alpha, double ratio, int minWidth,
                         int levels, int nOuterFPIterations, int nInnerFPIterations, int nSORIterations,
                         int colType, int mode, int device, int tuning);
int pf_plan_destroy(pf_plan* plan);
int pf_plan_levels(const pf_plan* plan);
/* H2D + solve + D2H with host buffers (pageable or pinned).  Any of vx / vy / warpI2 may be NULL: that output is
 * not copied back (the reference driver never reads warpI2, Par/OpticalFlowCalculation.py:74-76; it is 60 % of the
 * device->host bytes).  The same holds for pf_plan_download and, per pair or per array, for pf_batch_flow. */
int pf_plan_execute(pf_plan* plan, double* vx, double* vy, double* warpI2, const double* im1,
                    const double* im2, double* timings);
/* The same three legs separately, for resident-input measurement. */
int pf_plan_upload(pf_plan* plan, const double* im1, const double* im2);
int pf_plan_solve(pf_plan* plan, int repeats, double* ms_total);
int pf_plan_download(pf_plan* plan, double* vx, double* vy, double* warpI2);
/* One eager (un-graphed) solve with CUDA events around every phase; fills PF_NUM_TIMINGS doubles.
 * counters (may be NULL, 8 doubles): [0] kernel launches per solve, [1] SOR launches,
 * [2] SOR pixel-sweeps, [3] SOR ms at level 0, [4] SOR launches at level 0,
 * [5] pixel-sweeps at level 0. */
int pf_plan_profile(pf_plan* plan, double* timings, double* counters);
/* `repeats` solves on each of `nplans` plans of ONE device, all streams running concurrently
 * (pairs are independent): total GPU milliseconds between a start event every stream waits on and a
 * stop event that waits on every stream.  This is how a batch of resident pairs is timed. */
int pf_multi_solve(pf_plan* const* plans, int nplans, int repeats, double* ms_total);
/* Per-level phase times (ms) of the last pf_plan_profile call: out[level][PF_NUM_TIMINGS];
 * returns the number of levels written. */
int pf_plan_level_timings(const pf_plan* plan, double* out, int max_levels);
/* Gaussian-mixture parameters (alpha, sigma, beta per feature channel) a PF_NOISE_GMIXTURE plan holds after its last
 * solve: the state of OpticalFlow::GMPara (S/OpticalFlow.h:25) after the last estGaussianMixture call.  Returns the
 * number of channels written (<= n). */
int pf_plan_mixture_params(pf_plan* plan, double* alpha, double* sigma, double* beta, int n);

/* ---- batches: N independent frame pairs sharded over devices, no collective ---
 * pair p runs on devices[p % ndevices]; im1/im2/vx/vy/warpI2 are arrays of N host pointers. */
int pf_batch_flow(int npairs, double* const* vx, double* const* vy, double* const* warpI2,
                  const double* const* im1, const double* const* im2, double alpha, double ratio,
                  int minWidth, int levels, int nOuterFPIterations, int nInnerFPIterations,
                  int nSORIterations, int colType, int h, int w, int c, int mode,
                  const int* devices, int ndevices, double* seconds);

/* Event-timed legs of the LAST pf_batch_flow call of this process, 8 doubles: [0] pairs, [1] worker threads,
 * [2] wall seconds, mean per pair in ms of [3] H2D, [4] solve, [5] D2H (CUDA events on the pair's stream),
 * [6] host wall time of the whole pf_plan_execute call, [7] 0. */
int pf_batch_last_stats(double* out);

/* ---- sequences (SURVEY.md 8f "next" rows f1 + f2): nframes uint8 HWC frames (as PIL decodes
 * them, Par/OpticalFlowCalculation.py:66-67) -> nframes-1 flows of the consecutive pairs (t, t+1)
 * (pairing rule Par/InputCreation/TestImagePairGenerator.py:151-171) as interleaved float32 (u, v).
 * The /255 conversion runs on the device in double (bit-identical pixel values), every frame's
 * pyramid is built once and reused as the next pair's first image, and only uint8 in / float32
 * out cross PCIe.  Flows equal the pairwise entry points' (u, v) cast to float32.  Contiguous
 * chunks of the sequence are spread over devices x PF_BATCH_STREAMS workers; no collective. */
int pf_sequence_flow_u8(int nframes, const unsigned char* const* frames, float* const* flows,
                        double alpha, double ratio, int minWidth, int levels, int nOuterFPIterations,
                        int nInnerFPIterations, int nSORIterations, int colType, int h, int w, int c,
                        int mode, const int* devices, int ndevices, double* seconds);

/* Output formats of a sequence (bytes per pixel crossing PCIe): */
#define PF_SEQ_FLOW_F32 0 /* interleaved float32 (u, v), 8 B */
#define PF_SEQ_FLOW_U16 1 /* the reference's 16-bit flow encoding, interleaved (u, v), 4 B */
#define PF_SEQ_FLOW_BGR8 2 /* the reference driver's HSV flow visualisation as a BGR image, 3 B */

/* the same sequence with every flow leaving the device in the reference's
 * on-disk encoding, q = (unsigned short)((min(max(f,-200),200)+200)*160) evaluated in double
 * (OpticalFlow::SaveOpticalFlow, S/OpticalFlow.cpp:993-1003; decoded by LoadOpticalFlow :962-975 as
 * q/160-200): flows[t] receives h*w*2 unsigned shorts.  Half the D2H bytes of the float32 form. */
int pf_sequence_flow_u8_u16(int nframes, const unsigned char* const* frames, unsigned short* const* flows,
                            double alpha, double ratio, int minWidth, int levels, int nOuterFPIterations,
                            int nInnerFPIterations, int nSORIterations, int colType, int h, int w, int c,
                            int mode, const int* devices, int ndevices, double* seconds);

/* the same sequence with every flow leaving the device as the image the reference
 * driver writes for it (generateOutputFlowImageFile, Par/OpticalFlowCalculation.py:143-162:
 * cv2.cartToPolar -> hue = ang*180/pi/2, value = cv2.normalize(mag, 0..255, NORM_MINMAX), saturation 255
 * -> cv2.cvtColor(HSV2BGR)), computed on the device from the float32 flow: images[t] receives h*w*3 bytes
 * (BGR, what cv2.imwrite takes).  uint8 frames in, uint8 images out: 3/8 of the float32 D2H bytes. */
int pf_sequence_flow_u8_bgr(int nframes, const unsigned char* const* frames, unsigned char* const* images,
                            double alpha, double ratio, int minWidth, int levels, int nOuterFPIterations,
                            int nInnerFPIterations, int nSORIterations, int colType, int h, int w, int c,
                            int mode, const int* devices, int ndevices, double* seconds);

/* The visualisation alone, for callers of the pairwise entry points: flow is h*w interleaved float32
 * (u, v) in host memory, bgr receives h*w*3 bytes. */
int pf_flow_to_bgr(const float* flow, unsigned char* bgr, int h, int w, int device);

/* ---- ONE large pair over several GPUs (BASELINE config 5): every device runs the
 * cheap stages redundantly, the SOR solve is split into row bands whose halo rows -- and, after the
 * last pass, whole bands -- are stored by the SOR kernel itself into the neighbours' planes (peer-mapped
 * pointers over NVLink); passes are ordered by device-side counters in peer memory when every band has its
 * own GPU, by stream events otherwise; rows a non-adjacent band needs are fetched by a copy kernel.  No
 * collective, and the launch sequence of all devices is ONE CUDA graph.
 * FP32 red-black mode only; the result is bit-identical to the single-GPU fast mode.  Levels with
 * fewer than split_min_pixels pixels (<0: default 2000000) are solved redundantly.  devices may
 * repeat an index (bands then share one GPU -- used by the single-GPU tests of the exchange logic).
 * stats (may be NULL, 8 doubles): solve ms, halo bytes exchanged, gather bytes exchanged, split solves,
 * 1 if the solve was one multi-device graph replay, split solves ordered by device-side flags, 0, 0. */
int pf_multigpu_flow(double* vx, double* vy, double* warpI2, const double* im1, const double* im2,
                     double alpha, double ratio, int minWidth, int levels, int nOuterFPIterations,
                     int nInnerFPIterations, int nSORIterations, int colType, int h, int w, int c,
                     const int* devices, int ndevices, long long split_min_pixels, double* stats);

/* ---- single stages: the reference's public static functions (S/OpticalFlow.h:28-56) and the
 *      Image/ImageProcessing primitives they use, one call each, for per-stage parity tests.
 *      Same HWC float64 host buffers; `mode` selects the arithmetic (FP64 or FP32). ------------ */
/* GaussianPyramid::ConstructPyramidLevels (S/GaussianPyramid.cpp:79-108); out = levels concatenated */
int pf_stage_pyramid(double* out, const double* im, int h, int w, int c, double ratio, int levels,
                     int mode, int device);
/* OpticalFlow::im2feature (S/OpticalFlow.cpp:911-961); returns feature channel count (>0) */
int pf_stage_im2feature(double* feat, const double* im, int h, int w, int c, int swap_luma,
                        int mode, int device);
/* OpticalFlow::getDxs (S/OpticalFlow.cpp:80-122) */
int pf_stage_getdxs(double* imdx, double* imdy, double* imdt, const double* im1, const double* im2,
                    int h, int w, int c, int mode, int device);
/* OpticalFlow::warpFL (S/OpticalFlow.cpp:154-159 -> S/ImageProcessing.h:483-503) */
int pf_stage_warpfl(double* warp, const double* im1, const double* im2, const double* vx,
                    const double* vy, int h, int w, int c, int mode, int device);
/* Image::imresize(dstW,dstH) + Multiplywith (S/Image.h:778-783,1841-1850): flow upsampling */
int pf_stage_resize_to(double* dst, const double* src, int h, int w, int c, int dh, int dw,
                       double scale, int mode, int device);
/* Image::warpImageBicubicRef + threshold (S/Image.h:2587-2701, 2031-2045) */
int pf_stage_bicubic(double* out, const double* ref, const double* im2, const double* vx,
                     const double* vy, int h, int w, int c, int mode, int device);
/* Linear-system assembly of one inner iteration (S/OpticalFlow.cpp:295-448); lap has c entries;
 * du/dv may be NULL (zero).  Outputs are (h,w) planes. */
int pf_stage_assemble(double* phi, double* dxy, double* dx2, double* dy2, double* bu, double* bv,
                      const double* imdx, const double* imdy, const double* imdt, const double* u,
                      const double* v, const double* du, const double* dv, const double* lap,
                      double alpha, int h, int w, int c, int mode, int device);
/* The SOR sweeps alone (S/OpticalFlow.cpp:451-505) from du=dv=0; ordering from `mode`. */
int pf_stage_sor(double* du, double* dv, const double* phi, const double* dxy, const double* dx2,
                 const double* dy2, const double* bu, const double* bv, double alpha, int nsor,
                 int h, int w, int mode, int device);

/* ---- micro-benchmark of the SOR kernel on resident synthetic coefficients (bench.py roofline) */
int pf_bench_sor(int h, int w, int nsor, int repeats, int mode, int device, double* ms_per_solve,
                 double* launches_per_solve);

#ifdef __cplusplus
}
#endif
#endif /* PYFLOW_B200_H */
