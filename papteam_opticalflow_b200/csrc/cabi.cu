// extern "C" surface declared in include/pyflow_b200.h.  Argument validation, error translation
// and mode dispatch only; the work is in Plan<T> (solver.cuh) and Stages<T> (stages.cuh).
#include <atomic>
#include <cstdlib>
#include <chrono>
#include <memory>
#include <mutex>
#include <thread>

#include "factory.hpp"
#include "geometry.hpp"
#include "solver.cuh"

using namespace pf;

struct pf_plan {
    std::unique_ptr<PlanBase> impl;
};

namespace {

thread_local std::string g_err;
int (*g_pool_purge)() = nullptr;   // set below to pf_pool_clear

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

template <typename F>
int guarded(F&& f) {
    try {
        g_err.clear();
        return f();
    } catch (const Error& e) {
        cudaGetLastError();  // clear the sticky-less error state
        // a failed kernel leaves the context unusable (sticky error): idle pooled plans would be handed out again
        // and again -- drop them so that later calls rebuild (or fail cleanly at plan creation)
        if (e.code == PF_ECUDA && g_pool_purge) g_pool_purge();
        return fail(e.code, e.what());
    } catch (const std::bad_alloc&) {
        return fail(PF_ENOMEM, "host allocation failed");
    } catch (const std::exception& e) {
        return fail(PF_ECUDA, e.what());
    }
}

int check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(PF_ENODEVICE, std::string("no usable CUDA device (") +
                                      (e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)) +
                                      "); this library has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(PF_ENODEVICE, "device index " + std::to_string(device) + " out of range [0," + std::to_string(n) + ")");
    return PF_OK;
}

int check_mode(int mode) {
    if (mode < 0 || mode > 4) return fail(PF_EINVAL, "unknown mode " + std::to_string(mode));
    return PF_OK;
}

int check_image(int h, int w, int c) {
    if (h < 1 || w < 1 || c < 1) return fail(PF_EINVAL, "image dimensions must be positive");
    if (c > 16) return fail(PF_EUNSUPPORTED, "more than 16 channels");
    if ((long long)h * w > (1LL << 28)) return fail(PF_EUNSUPPORTED, "image too large");
    return PF_OK;
}

PlanBase* make_plan(const Params& p) { return mode_is_fp64(p.mode) ? make_plan_f64(p) : make_plan_f32(p); }
const StageCalls& stages(int mode) { return mode_is_fp64(mode) ? stages_f64() : stages_f32(); }

int stage_prologue(int h, int w, int c, int mode, int device) {
    int r;
    if ((r = check_image(h, w, c)) || (r = check_mode(mode)) || (r = check_device(device))) return r;
    PF_CUDA(cudaSetDevice(device));
    return PF_OK;
}


// ---- plan pool: one-shot and batch entry points reuse arenas + captured graphs across calls ------
struct PoolEntry {
    Params key;
    pf_plan* plan;
    bool busy;
    unsigned long long stamp;
};
std::mutex g_pool_mu;
std::vector<PoolEntry> g_pool;
unsigned long long g_pool_clock = 0;
const size_t kPoolMax = 16;

std::mutex g_batch_stats_mu;
double g_batch_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // of the last pf_batch_flow call (pf_batch_last_stats)

bool same_params(const Params& a, const Params& b) { return same_solver(a, b) && a.device == b.device; }

void pool_release(pf_plan* pl) {
    std::lock_guard<std::mutex> g(g_pool_mu);
    for (auto& e : g_pool)
        if (e.plan == pl) {
            pl->impl->set_blocking_wait(-1);   // back to the process default (batch / sequence workers switch to blocking waits)
            e.busy = false;
            e.stamp = ++g_pool_clock;
            return;
        }
}

}  // namespace

extern "C" int pf_plan_create_tuned(pf_plan** plan, int h, int w, int c, double alpha, double ratio, int minWidth,
                                    int levels, int nOuter, int nInner, int nSOR, int colType, int mode, int device, int tuning);
extern "C" int pf_plan_destroy(pf_plan* plan);

namespace {

// returns an idle pooled plan for these parameters, creating one if needed (rc != 0 on failure)
pf_plan* pool_acquire(const Params& p, int& rc) {
    rc = PF_OK;
    std::vector<pf_plan*> evict;
    {
        std::lock_guard<std::mutex> g(g_pool_mu);
        for (auto& e : g_pool)
            if (!e.busy && same_params(e.key, p)) {
                e.busy = true;
                return e.plan;
            }
        while (g_pool.size() >= kPoolMax) {     // drop the least recently used idle plan
            size_t best = g_pool.size();
            for (size_t i = 0; i < g_pool.size(); i++)
                if (!g_pool[i].busy && (best == g_pool.size() || g_pool[i].stamp < g_pool[best].stamp)) best = i;
            if (best == g_pool.size()) break;
            evict.push_back(g_pool[best].plan);
            g_pool.erase(g_pool.begin() + (long)best);
        }
    }
    for (pf_plan* e : evict) pf_plan_destroy(e);
    pf_plan* pl = nullptr;
    rc = pf_plan_create_tuned(&pl, p.h, p.w, p.c, p.alpha, p.ratio, p.min_width, p.levels, p.n_outer, p.n_inner, p.n_sor,
                              p.col_type, p.mode, p.device, p.tune);
    if (rc) return nullptr;
    std::lock_guard<std::mutex> g(g_pool_mu);
    g_pool.push_back(PoolEntry{p, pl, true, ++g_pool_clock});
    return pl;
}

}  // namespace

// Concurrent pairs run on one CUDA stream each.  The driver maps streams onto CUDA_DEVICE_MAX_CONNECTIONS
// hardware work queues and streams that share a queue serialise (measured on B200, driver 580: the batch
// entry point loses 13 % with the variable unset).  It is read when the CUDA context is created, so it is
// set -- without overriding the user's choice -- when this library is loaded; a host process that created
// its context earlier must export it itself.
extern "C" int pf_pool_clear(void);
__attribute__((constructor)) static void pf_default_connections() {
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    g_pool_purge = pf_pool_clear;
}

extern "C" {

const char* pf_last_error(void) { return g_err.c_str(); }
const char* pf_version(void) { return "pyflow_b200 0.1 (sm_100a)"; }

int pf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void* pf_host_alloc(size_t bytes) {
    void* p = nullptr;
    // PF_HOST_ALLOC_WC=1 (experiment): write-combined pages -- not snooped during DMA, very slow to READ from the CPU, so
    // only for buffers the host fills and the GPU reads
    static const bool wc = [] { const char* e = getenv("PF_HOST_ALLOC_WC"); return e && atoi(e) != 0; }();
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable | (wc ? cudaHostAllocWriteCombined : 0)) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void pf_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int pf_set_solver_variant(int interpolation, int noise_model) {
    if (interpolation != PF_INTERP_BILINEAR && interpolation != PF_INTERP_BICUBIC)
        return fail(PF_EINVAL, "interpolation must be PF_INTERP_BILINEAR or PF_INTERP_BICUBIC");
    if (noise_model != PF_NOISE_GMIXTURE && noise_model != PF_NOISE_LAP)
        return fail(PF_EINVAL, "noise_model must be PF_NOISE_GMIXTURE or PF_NOISE_LAP");
    solver_variant().interp = interpolation;
    solver_variant().noise = noise_model;
    return PF_OK;
}

int pf_get_solver_variant(int* interpolation, int* noise_model) {
    if (interpolation) *interpolation = solver_variant().interp;
    if (noise_model) *noise_model = solver_variant().noise;
    return PF_OK;
}

int pf_pyramid_levels(int width, double ratio, int minWidth) {
    if (width < 1 || minWidth < 1) return fail(PF_EINVAL, "width and minWidth must be positive");
    return levels_from_min_width(width, ratio, minWidth);
}

int pf_level_geometry(int w, int h, double ratio, int levels, int* widths, int* heights) {
    if (w < 1 || h < 1 || levels < 1 || levels > 64 || !widths || !heights) return fail(PF_EINVAL, "bad geometry request");
    std::vector<Level> g = level_geometry(w, h, ratio, levels);
    for (int i = 0; i < levels; i++) {
        widths[i] = g[(size_t)i].w;
        heights[i] = g[(size_t)i].h;
    }
    return PF_OK;
}

int pf_plan_create(pf_plan** plan, int h, int w, int c, double alpha, double ratio, int minWidth,
                   int levels, int nOuter, int nInner, int nSOR, int colType, int mode, int device) {
    return pf_plan_create_tuned(plan, h, w, c, alpha, ratio, minWidth, levels, nOuter, nInner, nSOR, colType, mode, device,
                                PF_TUNE_THROUGHPUT);
}

int pf_plan_create_tuned(pf_plan** plan, int h, int w, int c, double alpha, double ratio, int minWidth,
                         int levels, int nOuter, int nInner, int nSOR, int colType, int mode, int device, int tuning) {
    return guarded([&]() -> int {
        if (!plan) return fail(PF_EINVAL, "plan is NULL");
        *plan = nullptr;
        if (tuning != PF_TUNE_THROUGHPUT && tuning != PF_TUNE_LATENCY) return fail(PF_EINVAL, "tuning must be PF_TUNE_THROUGHPUT or PF_TUNE_LATENCY");
        int r;
        if ((r = check_image(h, w, c)) || (r = check_mode(mode))) return r;
        if (nOuter < 0 || nInner < 0 || nSOR < 0) return fail(PF_EINVAL, "iteration counts must be non-negative");
        if (levels <= 0 && minWidth < 1) return fail(PF_EINVAL, "minWidth must be positive");
        if (!(alpha == alpha) || !(ratio == ratio) || ratio <= 0) return fail(PF_EINVAL, "alpha/ratio invalid");
        if ((r = check_device(device))) return r;
        Params p{h, w, c, alpha, ratio, minWidth, levels, nOuter, nInner, nSOR, colType, mode, device};
        p.tune = tuning;
        std::unique_ptr<pf_plan> pl(new pf_plan);
        pl->impl.reset(make_plan(p));
        *plan = pl.release();
        return PF_OK;
    });
}

int pf_plan_destroy(pf_plan* plan) {
    return guarded([&]() -> int {
        delete plan;
        return PF_OK;
    });
}

int pf_plan_levels(const pf_plan* plan) { return plan ? plan->impl->levels() : PF_EINVAL; }

int pf_plan_execute(pf_plan* plan, double* vx, double* vy, double* warpI2, const double* im1,
                    const double* im2, double* timings) {
    return guarded([&]() -> int {
        if (!plan || !im1 || !im2) return fail(PF_EINVAL, "NULL argument");
        plan->impl->execute(vx, vy, warpI2, im1, im2, timings);   // NULL outputs are not copied back
        return PF_OK;
    });
}

int pf_plan_upload(pf_plan* plan, const double* im1, const double* im2) {
    return guarded([&]() -> int {
        if (!plan || !im1 || !im2) return fail(PF_EINVAL, "NULL argument");
        plan->impl->upload(im1, im2);
        return PF_OK;
    });
}

int pf_plan_solve(pf_plan* plan, int repeats, double* ms_total) {
    return guarded([&]() -> int {
        if (!plan || repeats < 1) return fail(PF_EINVAL, "bad argument");
        plan->impl->solve(repeats, ms_total);
        return PF_OK;
    });
}

int pf_plan_download(pf_plan* plan, double* vx, double* vy, double* warpI2) {
    return guarded([&]() -> int {
        if (!plan) return fail(PF_EINVAL, "NULL argument");
        plan->impl->download(vx, vy, warpI2);   // NULL outputs are not copied back
        return PF_OK;
    });
}

int pf_plan_profile(pf_plan* plan, double* timings, double* counters) {
    return guarded([&]() -> int {
        if (!plan || !timings) return fail(PF_EINVAL, "NULL argument");
        plan->impl->profile(timings, counters);
        return PF_OK;
    });
}

int pf_multi_solve(pf_plan* const* plans, int nplans, int repeats, double* ms_total) {
    return guarded([&]() -> int {
        if (!plans || nplans < 1 || repeats < 1) return fail(PF_EINVAL, "bad argument");
        for (int i = 0; i < nplans; i++) {
            if (!plans[i]) return fail(PF_EINVAL, "NULL plan");
            if (plans[i]->impl->device() != plans[0]->impl->device()) return fail(PF_EINVAL, "plans must share one device");
        }
        PF_CUDA(cudaSetDevice(plans[0]->impl->device()));
        cudaStream_t s0 = plans[0]->impl->stream();
        cudaEvent_t start, stop;
        std::vector<cudaEvent_t> done((size_t)nplans);
        PF_CUDA(cudaEventCreate(&start));
        PF_CUDA(cudaEventCreate(&stop));
        for (auto& e : done) PF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        auto cleanup = [&]() {
            cudaEventDestroy(start); cudaEventDestroy(stop);
            for (auto& e : done) cudaEventDestroy(e);
        };
        try {
            PF_CUDA(cudaDeviceSynchronize());
            // the start event is recorded on plan 0's stream; every other stream waits for it, and
            // plan 0's stream waits for every other stream before the stop event is recorded
            PF_CUDA(cudaEventRecord(start, s0));
            for (int i = 1; i < nplans; i++) PF_CUDA(cudaStreamWaitEvent(plans[i]->impl->stream(), start, 0));
            for (int r = 0; r < repeats; r++)
                for (int i = 0; i < nplans; i++) plans[i]->impl->solve_async(1);
            for (int i = 1; i < nplans; i++) {
                PF_CUDA(cudaEventRecord(done[(size_t)i], plans[i]->impl->stream()));
                PF_CUDA(cudaStreamWaitEvent(s0, done[(size_t)i], 0));
            }
            PF_CUDA(cudaEventRecord(stop, s0));
            PF_CUDA(cudaStreamSynchronize(s0));
            float ms = 0;
            PF_CUDA(cudaEventElapsedTime(&ms, start, stop));
            if (ms_total) *ms_total = ms;
        } catch (...) {
            cleanup();
            throw;
        }
        cleanup();
        return PF_OK;
    });
}

int pf_plan_level_timings(const pf_plan* plan, double* out, int max_levels) {
    if (!plan || !out || max_levels < 1) return fail(PF_EINVAL, "bad argument");
    return plan->impl->level_timings(out, max_levels);
}

int pf_plan_mixture_params(pf_plan* plan, double* alpha, double* sigma, double* beta, int n) {
    if (!plan || !alpha || !sigma || !beta || n < 1) return fail(PF_EINVAL, "bad argument");
    return guarded([&]() -> int { return plan->impl->mixture_params(alpha, sigma, beta, n); });
}

int pf_coarse2fine_flow(double* vx, double* vy, double* warpI2, const double* im1,
                        const double* im2, double alpha, double ratio, int minWidth, int nOuter,
                        int nInner, int nSOR, int colType, int h, int w, int c, int mode, int device,
                        double* timings) {
    int r = PF_OK;
    Params p{h, w, c, alpha, ratio, minWidth, 0, nOuter, nInner, nSOR, colType, mode, device};
    p.tune = PF_TUNE_LATENCY;   // one pair at a time
    pf_plan* pl = pool_acquire(p, r);
    if (r) return r;
    r = pf_plan_execute(pl, vx, vy, warpI2, im1, im2, timings);
    pool_release(pl);
    return r;
}

int pf_coarse2fine_flow_levels(double* vx, double* vy, double* warpI2, const double* im1,
                               const double* im2, int pyramidLevels, int nCores, int h, int w, int c,
                               int mode, int device, double* timings) {
    (void)nCores;
    if (pyramidLevels < 1) return fail(PF_EINVAL, "pyramidLevels must be >= 1");
    // hard-coded solver constants of the fork: S/OpticalFlow.cpp:747-751; colType 0: wrapper :22
    int r = PF_OK;
    Params p{h, w, c, 0.012, 0.75, 20, pyramidLevels, 7, 1, 30, 0, mode, device};
    p.tune = PF_TUNE_LATENCY;   // one pair at a time
    pf_plan* pl = pool_acquire(p, r);
    if (r) return r;
    r = pf_plan_execute(pl, vx, vy, warpI2, im1, im2, timings);
    pool_release(pl);
    return r;
}

int pf_batch_last_stats(double* out) {
    if (!out) return fail(PF_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(g_batch_stats_mu);
    for (int i = 0; i < 8; i++) out[i] = g_batch_stats[i];
    return PF_OK;
}

int pf_pool_clear(void) {
    std::vector<pf_plan*> all;
    {
        std::lock_guard<std::mutex> g(g_pool_mu);
        for (size_t i = 0; i < g_pool.size();)
            if (!g_pool[i].busy) {
                all.push_back(g_pool[i].plan);
                g_pool.erase(g_pool.begin() + (long)i);
            } else {
                i++;
            }
    }
    for (pf_plan* p : all) pf_plan_destroy(p);
    return (int)all.size();
}

int pf_batch_flow(int npairs, double* const* vx, double* const* vy, double* const* warpI2,
                  const double* const* im1, const double* const* im2, double alpha, double ratio,
                  int minWidth, int levels, int nOuter, int nInner, int nSOR, int colType, int h,
                  int w, int c, int mode, const int* devices, int ndevices, double* seconds) {
    if (npairs < 0 || ndevices < 1 || !devices) return fail(PF_EINVAL, "bad batch arguments");
    if (npairs > 0 && (!im1 || !im2)) return fail(PF_EINVAL, "NULL argument");
    for (int d = 0; d < ndevices; d++) {
        int r = check_device(devices[d]);
        if (r) return r;
    }
    auto t0 = std::chrono::steady_clock::now();
    // PF_BATCH_STREAMS workers (plan + stream each) per device pull pairs from one queue, so the
    // H2D / solve / D2H legs of different pairs overlap and the latency-bound coarse levels of one
    // pair hide behind the bandwidth-bound fine levels of another.
    const char* env = getenv("PF_BATCH_STREAMS");
    int per_dev = env ? atoi(env) : 8;
    if (per_dev < 1) per_dev = 1;
    if (mode_is_lex(mode)) per_dev = 1;   // cooperative launches do not overlap usefully
    const int nworkers = std::min(std::max(npairs, 1), ndevices * per_dev);
    std::vector<std::thread> workers;
    std::vector<int> status((size_t)nworkers, PF_OK);
    std::vector<std::string> messages((size_t)nworkers);
    std::vector<std::atomic<int>> next((size_t)ndevices);
    for (auto& n : next) n.store(0);
    // PF_BATCH_TRACE=1: mean CUDA-event time of the three legs of a pair, printed to stderr (diagnostic)
    const bool trace = getenv("PF_BATCH_TRACE") && atoi(getenv("PF_BATCH_TRACE"));
    std::mutex trace_mu;
    double trace_sum[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int wk = 0; wk < nworkers; wk++) {
        workers.emplace_back([&, wk]() {
            const int d = wk % ndevices;
            int r = PF_OK;
            Params pp{h, w, c, alpha, ratio, minWidth, levels, nOuter, nInner, nSOR, colType, mode, devices[d]};
            pf_plan* pl = pool_acquire(pp, r);
            if (pl) pl->impl->set_blocking_wait(1);   // many waiting host threads per process: never spin
            // pair p belongs to device p % ndevices; workers of one device share its queue
            while (r == PF_OK) {
                int k = next[(size_t)d].fetch_add(1);
                int p = d + k * ndevices;
                if (p >= npairs) break;
                double t[PF_NUM_TIMINGS];
                // an output array that is NULL, or a NULL entry in it, means "do not copy this output back"
                r = pf_plan_execute(pl, vx ? vx[p] : nullptr, vy ? vy[p] : nullptr, warpI2 ? warpI2[p] : nullptr, im1[p], im2[p], t);
                if (r == PF_OK) {
                    std::lock_guard<std::mutex> lk(trace_mu);
                    trace_sum[0] += t[PF_T_H2D]; trace_sum[1] += t[PF_T_SOLVE]; trace_sum[2] += t[PF_T_D2H]; trace_sum[3] += 1;
                    trace_sum[4] += t[13]; trace_sum[5] += t[14]; trace_sum[6] += t[15];
                }
            }
            if (r) messages[(size_t)wk] = g_err;
            if (pl) pool_release(pl);
            status[(size_t)wk] = r;
        });
    }
    for (auto& t : workers) t.join();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (seconds) *seconds = secs;
    {
        std::lock_guard<std::mutex> lk(g_batch_stats_mu);
        const double n = trace_sum[3] > 0 ? trace_sum[3] : 1;
        g_batch_stats[0] = trace_sum[3]; g_batch_stats[1] = (double)nworkers; g_batch_stats[2] = secs;
        g_batch_stats[3] = trace_sum[0] / n; g_batch_stats[4] = trace_sum[1] / n; g_batch_stats[5] = trace_sum[2] / n;
        g_batch_stats[6] = trace_sum[6] / n; g_batch_stats[7] = 0;
    }
    if (trace && trace_sum[3] > 0)
        fprintf(stderr, "[pf_batch_flow] %d pairs, %d workers, %.3f s: mean per pair H2D %.2f ms, solve %.2f ms, D2H %.2f ms; host: copy enqueue %.2f ms, "
                "graph launch %.2f ms, whole call %.2f ms\n", npairs, nworkers, secs, trace_sum[0] / trace_sum[3], trace_sum[1] / trace_sum[3],
                trace_sum[2] / trace_sum[3], trace_sum[4] / trace_sum[3], trace_sum[5] / trace_sum[3], trace_sum[6] / trace_sum[3]);
    for (int wk = 0; wk < nworkers; wk++)
        if (status[(size_t)wk]) return fail(status[(size_t)wk], messages[(size_t)wk]);
    return PF_OK;
}

}  // extern "C"

static int sequence_flow(int format, int nframes, const unsigned char* const* frames, void* const* flows, double alpha, double ratio,
                         int minWidth, int levels, int nOuter, int nInner, int nSOR, int colType, int h, int w, int c, int mode,
                         const int* devices, int ndevices, double* seconds) {
    if (nframes < 0 || ndevices < 1 || !devices) return fail(PF_EINVAL, "bad sequence arguments");
    const int npairs = nframes > 0 ? nframes - 1 : 0;
    if (npairs > 0 && (!frames || !flows)) return fail(PF_EINVAL, "NULL argument");
    int r;
    if ((r = check_image(h, w, c)) || (r = check_mode(mode))) return r;
    for (int d = 0; d < ndevices; d++)
        if ((r = check_device(devices[d]))) return r;
    auto t0 = std::chrono::steady_clock::now();
    // contiguous chunks of the pair list per worker, so that inside a chunk every frame's pyramid is
    // built once and reused as the next pair's first image
    const char* env = getenv("PF_BATCH_STREAMS");
    int per_dev = env ? atoi(env) : 8;
    if (per_dev < 1 || mode_is_lex(mode)) per_dev = 1;
    const int nworkers = std::max(1, std::min(npairs, ndevices * per_dev));
    std::vector<std::thread> workers;
    std::vector<int> status((size_t)nworkers, PF_OK);
    std::vector<std::string> messages((size_t)nworkers);
    for (int wk = 0; wk < nworkers && npairs > 0; wk++) {
        workers.emplace_back([&, wk]() {
            const int p0 = (int)((long long)wk * npairs / nworkers), p1 = (int)((long long)(wk + 1) * npairs / nworkers);
            if (p1 <= p0) return;
            int rc = PF_OK;
            Params pp{h, w, c, alpha, ratio, minWidth, levels, nOuter, nInner, nSOR, colType, mode, devices[wk % ndevices]};
            pf_plan* pl = pool_acquire(pp, rc);
            if (pl) pl->impl->set_blocking_wait(1);
            if (rc == PF_OK) {
                rc = guarded([&]() -> int {
                    pl->impl->seq_first(frames[p0]);
                    for (int p = p0; p < p1; p++) pl->impl->seq_next(frames[p + 1], flows[p], format);
                    return PF_OK;
                });
            }
            if (rc) messages[(size_t)wk] = g_err;
            if (pl) pool_release(pl);
            status[(size_t)wk] = rc;
        });
    }
    for (auto& t : workers) t.join();
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int wk = 0; wk < nworkers; wk++)
        if (status[(size_t)wk]) return fail(status[(size_t)wk], messages[(size_t)wk]);
    return PF_OK;
}

int pf_sequence_flow_u8(int nframes, const unsigned char* const* frames, float* const* flows, double alpha, double ratio,
                        int minWidth, int levels, int nOuter, int nInner, int nSOR, int colType, int h, int w, int c, int mode,
                        const int* devices, int ndevices, double* seconds) {
    return sequence_flow(PF_SEQ_FLOW_F32, nframes, frames, reinterpret_cast<void* const*>(flows), alpha, ratio, minWidth, levels, nOuter,
                         nInner, nSOR, colType, h, w, c, mode, devices, ndevices, seconds);
}

int pf_sequence_flow_u8_bgr(int nframes, const unsigned char* const* frames, unsigned char* const* images, double alpha, double ratio,
                            int minWidth, int levels, int nOuter, int nInner, int nSOR, int colType, int h, int w, int c, int mode,
                            const int* devices, int ndevices, double* seconds) {
    return sequence_flow(PF_SEQ_FLOW_BGR8, nframes, frames, reinterpret_cast<void* const*>(images), alpha, ratio, minWidth, levels, nOuter,
                         nInner, nSOR, colType, h, w, c, mode, devices, ndevices, seconds);
}

int pf_flow_to_bgr(const float* flow, unsigned char* bgr, int h, int w, int device) {
    if (!flow || !bgr) return fail(PF_EINVAL, "NULL argument");
    int r;
    if ((r = check_image(h, w, 1)) || (r = check_device(device))) return r;
    return guarded([&]() -> int {
        PF_CUDA(cudaSetDevice(device));
        const size_t n = (size_t)h * w;
        float* d_flow = nullptr;
        unsigned char* d_out = nullptr;
        unsigned int* d_mm = nullptr;
        cudaStream_t st = nullptr;
        auto cleanup = [&]() {
            if (d_flow) cudaFree(d_flow);
            if (d_out) cudaFree(d_out);
            if (d_mm) cudaFree(d_mm);
            if (st) cudaStreamDestroy(st);
        };
        try {
            PF_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            PF_CUDA(cudaMalloc(&d_flow, n * 2 * sizeof(float)));
            PF_CUDA(cudaMalloc(&d_out, n * 3));
            PF_CUDA(cudaMalloc(&d_mm, 2 * sizeof(unsigned int)));
            PF_CUDA(cudaMemcpyAsync(d_flow, flow, n * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
            k_minmax_init<<<1, 1, 0, st>>>(d_mm);
            k_flow_mag_minmax<float><<<dim3(ceil_div(w, 256), std::min(h, 256)), 256, 0, st>>>(d_flow, d_flow + 1, 2 * w, 2, w, h, d_mm);
            k_flow_to_bgr<float><<<dim3(ceil_div(w, 128), h), 128, 0, st>>>(d_flow, d_flow + 1, 2 * w, 2, w, d_mm, d_out);
            PF_CHECK_LAUNCH();
            PF_CUDA(cudaMemcpyAsync(bgr, d_out, n * 3, cudaMemcpyDeviceToHost, st));
            PF_CUDA(cudaStreamSynchronize(st));
        } catch (...) {
            cleanup();
            throw;
        }
        cleanup();
        return PF_OK;
    });
}

int pf_sequence_flow_u8_u16(int nframes, const unsigned char* const* frames, unsigned short* const* flows, double alpha, double ratio,
                            int minWidth, int levels, int nOuter, int nInner, int nSOR, int colType, int h, int w, int c, int mode,
                            const int* devices, int ndevices, double* seconds) {
    return sequence_flow(PF_SEQ_FLOW_U16, nframes, frames, reinterpret_cast<void* const*>(flows), alpha, ratio, minWidth, levels, nOuter,
                         nInner, nSOR, colType, h, w, c, mode, devices, ndevices, seconds);
}

extern "C" {

int pf_multigpu_flow(double* vx, double* vy, double* warpI2, const double* im1, const double* im2, double alpha,
                     double ratio, int minWidth, int levels, int nOuter, int nInner, int nSOR, int colType, int h, int w,
                     int c, const int* devices, int ndevices, long long split_min_pixels, double* stats) {
    return guarded([&]() -> int {
        if (!vx || !vy || !warpI2 || !im1 || !im2 || !devices || ndevices < 1 || ndevices > 16) return fail(PF_EINVAL, "bad argument");
        int r;
        if ((r = check_image(h, w, c))) return r;
        for (int d = 0; d < ndevices; d++)
            if ((r = check_device(devices[d]))) return r;
        if (nOuter < 0 || nInner < 0 || nSOR < 0) return fail(PF_EINVAL, "iteration counts must be non-negative");
        Params p{h, w, c, alpha, ratio, minWidth, levels, nOuter, nInner, nSOR, colType, PF_MODE_FP32_REDBLACK, devices[0]};
        p.tune = PF_TUNE_LATENCY;   // one pair over several GPUs
        multigpu_flow_f32(vx, vy, warpI2, im1, im2, p, devices, ndevices, split_min_pixels < 0 ? 2000000 : split_min_pixels, stats);
        return PF_OK;
    });
}

// ---- single stages ---------------------------------------------------------------------------------
int pf_stage_pyramid(double* out, const double* im, int h, int w, int c, double ratio, int levels, int mode, int device) {
    return guarded([&]() -> int {
        if (!out || !im || levels < 1 || levels > 64) return fail(PF_EINVAL, "bad argument");
        int r = stage_prologue(h, w, c, mode, device);
        if (r) return r;
        stages(mode).pyramid(out, im, h, w, c, ratio, levels);
        return PF_OK;
    });
}

int pf_stage_im2feature(double* feat, const double* im, int h, int w, int c, int swap_luma, int mode, int device) {
    return guarded([&]() -> int {
        if (!feat || !im) return fail(PF_EINVAL, "NULL argument");
        int r = stage_prologue(h, w, c, mode, device);
        if (r) return r;
        return stages(mode).im2feature(feat, im, h, w, c, swap_luma);
    });
}

int pf_stage_getdxs(double* imdx, double* imdy, double* imdt, const double* im1, const double* im2, int h, int w, int c, int mode, int device) {
    return guarded([&]() -> int {
        if (!imdx || !imdy || !imdt || !im1 || !im2) return fail(PF_EINVAL, "NULL argument");
        int r = stage_prologue(h, w, c, mode, device);
        if (r) return r;
        stages(mode).getdxs(imdx, imdy, imdt, im1, im2, h, w, c);
        return PF_OK;
    });
}

int pf_stage_warpfl(double* warp, const double* im1, const double* im2, const double* vx, const double* vy, int h, int w, int c, int mode, int device) {
    return guarded([&]() -> int {
        if (!warp || !im1 || !im2 || !vx || !vy) return fail(PF_EINVAL, "NULL argument");
        int r = stage_prologue(h, w, c, mode, device);
        if (r) return r;
        stages(mode).warpfl(warp, im1, im2, vx, vy, h, w, c);
        return PF_OK;
    });
}

int pf_stage_resize_to(double* dst, const double* src, int h, int w, int c, int dh, int dw, double scale, int mode, int device) {
    return guarded([&]() -> int {
        if (!dst || !src || dh < 1 || dw < 1) return fail(PF_EINVAL, "bad argument");
        int r = stage_prologue(h, w, c, mode, device);
        if (r) return r;
        stages(mode).resize_to(dst, src, h, w, c, dh, dw, scale);
        return PF_OK;
    });
}

int pf_stage_bicubic(double* out, const double* ref, const double* im2, const double* vx, const double* vy, int h, int w, int c, int mode, int device) {
    return guarded([&]() -> int {
        if (!out || !ref || !im2 || !vx || !vy) return fail(PF_EINVAL, "NULL argument");
        int r = stage_prologue(h, w, c, mode, device);
        if (r) return r;
        stages(mode).bicubic(out, ref, im2, vx, vy, h, w, c);
        return PF_OK;
    });
}

int pf_stage_assemble(double* phi, double* dxy, double* dx2, double* dy2, double* bu, double* bv,
                      const double* imdx, const double* imdy, const double* imdt, const double* u,
                      const double* v, const double* du, const double* dv, const double* lap,
                      double alpha, int h, int w, int c, int mode, int device) {
    return guarded([&]() -> int {
        if (!phi || !dxy || !dx2 || !dy2 || !bu || !bv || !imdx || !imdy || !imdt || !u || !v) return fail(PF_EINVAL, "NULL argument");
        int r = stage_prologue(h, w, c, mode, device);
        if (r) return r;
        stages(mode).assemble(phi, dxy, dx2, dy2, bu, bv, imdx, imdy, imdt, u, v, du, dv, lap, alpha, h, w, c);
        return PF_OK;
    });
}

int pf_stage_sor(double* du, double* dv, const double* phi, const double* dxy, const double* dx2,
                 const double* dy2, const double* bu, const double* bv, double alpha, int nsor, int h,
                 int w, int mode, int device) {
    return guarded([&]() -> int {
        if (!du || !dv || !phi || !dxy || !dx2 || !dy2 || !bu || !bv || nsor < 0) return fail(PF_EINVAL, "bad argument");
        int r = stage_prologue(h, w, 1, mode, device);
        if (r) return r;
        stages(mode).sor(du, dv, phi, dxy, dx2, dy2, bu, bv, alpha, nsor, h, w, mode, device, 1, nullptr, nullptr);
        return PF_OK;
    });
}

int pf_bench_sor(int h, int w, int nsor, int repeats, int mode, int device, double* ms_per_solve, double* launches_per_solve) {
    return guarded([&]() -> int {
        if (nsor < 1 || repeats < 2) return fail(PF_EINVAL, "nsor >= 1 and repeats >= 2 required");
        int r = stage_prologue(h, w, 1, mode, device);
        if (r) return r;
        // synthetic but well-conditioned coefficients (positive weights, diagonally dominant system)
        size_t n = (size_t)h * w;
        std::vector<double> phi(n), dxy(n), dx2(n), dy2(n), bu(n), bv(n);
        unsigned s = 12345u;
        auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / 16777216.0; };
        for (size_t i = 0; i < n; i++) {
            phi[i] = 0.5 + 20 * rnd(); dxy[i] = 0.2 * (rnd() - 0.5); dx2[i] = 0.1 + rnd(); dy2[i] = 0.1 + rnd();
            bu[i] = rnd() - 0.5; bv[i] = rnd() - 0.5;
        }
        stages(mode).sor(nullptr, nullptr, phi.data(), dxy.data(), dx2.data(), dy2.data(), bu.data(), bv.data(),
                         0.012, nsor, h, w, mode, device, repeats, ms_per_solve, launches_per_solve);
        return PF_OK;
    });
}

}  // extern "C"
