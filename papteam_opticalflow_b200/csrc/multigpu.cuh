// MultiPlan: ONE frame pair solved by several GPUs (BASELINE config 5, SURVEY.md 8e second row).
//
// Every device holds a full replica of the plan and runs the cheap stages (pyramid, features,
// assembly, warp) redundantly, so those stages need no communication and stay bit-identical to the
// single-GPU path.  The SOR solve -- ~80 % of the time of a 3840x2160 pair with 60 sweeps -- is split
// into contiguous ROW BANDS of tile rows: in each fused-sweep pass device g runs the tile kernel only
// on its band, then pulls the 2*nsw halo rows its next pass will read from the neighbours that own
// them with peer-to-peer copies over NVLink (cudaMemcpyPeerAsync between peer-enabled devices, ordered
// by CUDA events; no host synchronisation, no NCCL: this is a nearest-neighbour halo exchange, not a
// collective).  After the last pass every device gathers the other bands so that the replicated
// update + warp stage sees the complete du, dv.  The red-black update is independent of the tiling, so
// the result equals the single-GPU result bit for bit.
// Between the passes of one solve the bands are ordered by DEVICE-SIDE FLAGS when every band has its own
// GPU: the CTAs of a pass bump counters in the neighbours' memory when they are done and the neighbours'
// next pass spins on them (SorPeer, sor.cuh) -- a cross-device stream-event edge costs about as much as a
// whole pass of a 4K level (~60 us), a flag in peer memory a few microseconds.  Stream events still open
// every split solve and order its final gather.  Bands that share a GPU (tests) keep the event path: a
// spinning persistent kernel would occupy the SMs its producer needs.
// Levels below `split_min_pixels` are solved redundantly on every device (no exchange at all).
// The lexicographic parity mode does not band-split (band b would wait for band b-1 in every sweep):
// replicas only.
#pragma once
#include <chrono>
#include "solver.cuh"

namespace pf {

// grid-stride 16-byte copy; `src` may be a peer-mapped pointer (loads travel over NVLink)
static __global__ void k_copy_peer(float4* __restrict__ dst, const float4* __restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

class MultiPlan {
  public:
    MultiPlan(const Params& p, const int* devices, int ndev, long long split_min_pixels)
        : P(p), split_min_(split_min_pixels) {
        const char* e = getenv("PF_MULTI_PULL");
        push_ = !(e && atoi(e));
        if (mode_is_lex(p.mode) || mode_is_fp64(p.mode)) throw Error(PF_EUNSUPPORTED, "row-band split needs the fp32_redblack mode");
        for (int g = 0; g < ndev; g++) devs_.push_back(devices[g]);
        for (int g = 0; g < ndev; g++) {
            PF_CUDA(cudaSetDevice(devs_[g]));
            for (int h = 0; h < ndev; h++) {
                if (devs_[h] == devs_[g]) continue;
                int can = 0;
                PF_CUDA(cudaDeviceCanAccessPeer(&can, devs_[g], devs_[h]));
                if (can) {
                    cudaError_t e = cudaDeviceEnablePeerAccess(devs_[h], 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) PF_CUDA(e);
                    cudaGetLastError();
                }
            }
            Params q = p;
            q.device = devs_[g];
            plans_.emplace_back(new Plan<float>(q));
            plans_.back()->set_pdl(false);      // flag-ordered passes and cross-device edges: ordinary launches only
            plans_.back()->set_fold(false);     // the phases are driven from here, per band
            cudaEvent_t e1, e2;
            PF_CUDA(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
            PF_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
            ev_pass_.push_back(e1);
            ev_gather_.push_back(e2);
        }
        PF_CUDA(cudaSetDevice(devs_[0]));
        PF_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
        // device-side flags need one GPU per band and peer access between neighbouring bands.  PF_MULTI_FLAGS_SHARED=1
        // (tests on a single GPU) also allows them between bands that share a device, for the split solves whose passes
        // leave room for the spinning CTAs of a waiting pass next to the CTAs of the pass they wait for (checked per solve)
        e = getenv("PF_MULTI_FLAGS");
        flags_ok_ = push_ && !(e && !atoi(e)) && ndev > 1;
        const char* es = getenv("PF_MULTI_FLAGS_SHARED");
        const bool allow_shared = es && atoi(es) != 0;
        for (int g = 0; g < ndev && flags_ok_; g++) {
            for (int h = 0; h < g; h++)
                if (devs_[h] == devs_[g]) {
                    if (allow_shared) flags_shared_ = true;
                    else flags_ok_ = false;
                }
            if (g + 1 < ndev && flags_ok_ && devs_[g] != devs_[g + 1]) {
                int a = 0, b = 0;
                PF_CUDA(cudaDeviceCanAccessPeer(&a, devs_[g], devs_[g + 1]));
                PF_CUDA(cudaDeviceCanAccessPeer(&b, devs_[g + 1], devs_[g]));
                if (!a || !b) flags_ok_ = false;
            }
        }
        if (flags_ok_) {
            flags_.assign((size_t)ndev, nullptr);
            for (int g = 0; g < ndev; g++) {
                PF_CUDA(cudaSetDevice(devs_[g]));
                PF_CUDA(cudaMalloc(&flags_[(size_t)g], (kFlagSlots * 2 + 1) * sizeof(unsigned int)));   // + the error word
            }
        }
    }
    ~MultiPlan() {
        if (gexec_) cudaGraphExecDestroy(gexec_);
        if (graph_) cudaGraphDestroy(graph_);
        if (ev_fork_) cudaEventDestroy(ev_fork_);
        for (size_t g = 0; g < flags_.size(); g++) {
            cudaSetDevice(devs_[g]);
            if (flags_[g]) cudaFree(flags_[g]);
        }
        for (size_t g = 0; g < devs_.size(); g++) {
            cudaSetDevice(devs_[g]);
            cudaEventDestroy(ev_pass_[g]);
            cudaEventDestroy(ev_gather_[g]);
        }
    }
    bool matches(const Params& p, const int* devices, int ndev, long long split_min) const {
        if ((int)devs_.size() != ndev || split_min != split_min_) return false;
        for (int g = 0; g < ndev; g++)
            if (devs_[g] != devices[g]) return false;
        return same_solver(P, p);
    }

    // returns milliseconds of the solve (host clock around device-synchronised region)
    double execute(double* vx, double* vy, double* warp, const double* im1, const double* im2, double* stats) {
        const int G = (int)devs_.size();
        for (int g = 0; g < G; g++) plans_[g]->upload(im1, im2);
        sync_all();
        auto t0 = std::chrono::steady_clock::now();
        run_solve();
        sync_all();
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        for (size_t g = 0; g < flags_.size(); g++) {      // a pass that ran out of time waiting for a neighbour's counters
            unsigned int lost = 0;
            on((int)g);
            PF_CUDA(cudaMemcpy(&lost, flags_[g] + 2 * kFlagSlots, sizeof(lost), cudaMemcpyDeviceToHost));
            if (lost) {
                flags_ok_ = false;                         // this plan falls back to stream events from now on
                if (gexec_) { cudaGraphExecDestroy(gexec_); gexec_ = nullptr; }
                if (graph_) { cudaGraphDestroy(graph_); graph_ = nullptr; }
                throw Error(PF_ECUDA, "row-band split: device " + std::to_string(devs_[g]) + " waited 20 s for the " +
                                          (lost == 1 ? "upper" : "lower") + " neighbour's pass counters (is that GPU busy with other work?); "
                                          "the result was discarded, later calls order the passes with stream events");
            }
        }
        plans_[0]->download(vx, vy, warp);
        if (stats) {
            stats[0] = ms;
            stats[1] = (double)halo_bytes_;
            stats[2] = (double)gather_bytes_;
            stats[3] = (double)split_solves_;
            stats[4] = gexec_ ? 1.0 : 0.0;       // the whole multi-device launch sequence was one CUDA graph replay
            stats[5] = (double)flag_solves_;     // split solves whose passes were ordered by device-side flags
        }
        if (getenv("PF_MULTI_TRACE"))
            fprintf(stderr, "pyflow_b200 multigpu [%s]: %.3f ms, %d split solves (%d ordered by device-side flags, %d flag slots), halo %.1f MB, gather %.1f MB\n",
                    gexec_ ? "one multi-device graph" : "eager", ms, split_solves_, flag_solves_, flag_slot_, halo_bytes_ / 1e6, gather_bytes_ / 1e6);
        return ms;
    }

  private:
    void on(int g) { PF_CUDA(cudaSetDevice(devs_[g])); }

    // The launch sequence of all devices (kernels, peer copies, cross-device event edges) is captured
    // once into ONE multi-device CUDA graph: device 0's stream is the origin, the other streams fork
    // from it and join back, so a replay needs no host thread in the loop (an eager solve issues
    // ~10^4 API calls from one thread, which is what bounds a 2-GPU solve otherwise).
    void run_solve() {
        const int G = (int)devs_.size();
        const char* e = getenv("PF_NO_GRAPH");
        if (e && atoi(e)) { solve(); return; }
        cudaStream_t s0 = plans_[0]->stream();
        // ThreadLocal like Plan::run_solve: another thread's cudaMalloc / cudaHostAlloc (pool_acquire, the stager, ...)
        // must neither fail nor invalidate this capture.  A failed capture is retried once before the plan settles for
        // eager launches (reported through stats[4] = 0 and PF_MULTI_TRACE).
        for (int attempt = 0; attempt < 2 && !gexec_ && !graph_failed_; attempt++) {
            on(0);
            cudaError_t st = cudaStreamBeginCapture(s0, cudaStreamCaptureModeThreadLocal);
            std::string why;
            if (st == cudaSuccess) {
                try {
                    PF_CUDA(cudaEventRecord(ev_fork_, s0));
                    for (int g = 1; g < G; g++) { on(g); PF_CUDA(cudaStreamWaitEvent(plans_[g]->stream(), ev_fork_, 0)); }
                    solve();
                    for (int g = 1; g < G; g++) {
                        on(g);
                        PF_CUDA(cudaEventRecord(ev_pass_[g], plans_[g]->stream()));
                        on(0);
                        PF_CUDA(cudaStreamWaitEvent(s0, ev_pass_[g], 0));
                    }
                    on(0);
                    PF_CUDA(cudaStreamEndCapture(s0, &graph_));
                    PF_CUDA(cudaGraphInstantiate(&gexec_, graph_, 0));
                } catch (const std::exception& err) {
                    why = err.what();
                    cudaGraph_t g = nullptr;
                    on(0);
                    cudaStreamEndCapture(s0, &g);
                    if (g) cudaGraphDestroy(g);
                    if (graph_) { cudaGraphDestroy(graph_); graph_ = nullptr; }
                    cudaGetLastError();
                    gexec_ = nullptr;
                }
            } else {
                why = std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(st);
                cudaGetLastError();
            }
            if (!gexec_) {
                if (getenv("PF_MULTI_TRACE")) fprintf(stderr, "pyflow_b200 multigpu: graph capture failed (attempt %d): %s\n", attempt + 1, why.c_str());
                if (attempt == 1) graph_failed_ = true;
            }
        }
        if (gexec_) {
            on(0);
            PF_CUDA(cudaGraphLaunch(gexec_, s0));
        } else {
            solve();
        }
    }
    void sync_all() {
        for (size_t g = 0; g < devs_.size(); g++) {
            on((int)g);
            PF_CUDA(cudaStreamSynchronize(plans_[g]->stream()));
        }
    }

    void solve() {
        const int G = (int)devs_.size();
        halo_bytes_ = gather_bytes_ = 0;
        split_solves_ = 0;
        flag_slot_ = 0;
        flag_solves_ = 0;
        if (flags_ok_)   // counters start at zero in every solve (the first split solve's opening barrier orders this)
            for (int g = 0; g < G; g++) {
                on(g);
                PF_CUDA(cudaMemsetAsync(flags_[(size_t)g], 0, (kFlagSlots * 2 + 1) * sizeof(unsigned int), plans_[g]->stream()));
            }
        for (int g = 0; g < G; g++) { on(g); plans_[g]->ph_begin(); }
        const int nlev = plans_[0]->levels();
        for (int k = nlev - 1; k >= 0; k--) {
            for (int g = 0; g < G; g++) { on(g); plans_[g]->ph_level(k); }
            for (int it = 0; it < plans_[0]->n_outer_at(k); it++) {
                for (int g = 0; g < G; g++) { on(g); plans_[g]->ph_getdxs(k); }
                for (int hh = 0; hh < P.n_inner; hh++) {
                    for (int g = 0; g < G; g++) {
                        on(g);
                        int lo = 0, hi = -1;
                        sor_rows(k, g, lo, hi);          // split solve: only the rows this device's band reads
                        plans_[g]->ph_assemble(k, hh, lo, hi);
                    }
                    sor(k);
                }
                const bool last_it = it + 1 == plans_[0]->n_outer_at(k);
                for (int g = 0; g < G; g++) {
                    on(g);
                    int lo = 0, hi = -1;
                    if (!last_it) sor_rows(k, g, lo, hi);     // (after the last iteration nobody reads the warped features again:
                    if (hi < 0) plans_[g]->ph_update(k);      //  the next level warps from scratch; kept whole for simplicity)
                    else {
                        // the assembly of rows [lo, hi) works on whole tile rows and reads the warped features 4 rows beyond them
                        const int ty = plans_[g]->fused_tile_rows(k);
                        plans_[g]->ph_update(k, lo / ty * ty - 4, (hi + ty - 1) / ty * ty + 4);
                    }
                }
            }
        }
        for (int g = 0; g < G; g++) { on(g); plans_[g]->ph_end(); }
    }

    // rows [r0, r1) of plane `src` on device hs -> same rows of plane `dst` on device hd, on hd's stream.
    // Between different GPUs a copy KERNEL on the destination reads the peer-mapped source over NVLink:
    // cudaMemcpyPeerAsync is not permitted while a stream is capturing, and the whole multi-device launch
    // sequence has to stay one CUDA graph (an eager 2-GPU solve is bound by the host issuing ~10^4 launches).
    void pull_rows(int hd, float* dst, int hs, const float* src, int r0, int r1, int pitch, long long& counter) {
        if (r1 <= r0) return;
        size_t off = (size_t)r0 * pitch, bytes = (size_t)(r1 - r0) * pitch * sizeof(float);
        if (devs_[hd] == devs_[hs]) {
            PF_CUDA(cudaMemcpyAsync(dst + off, src + off, bytes, cudaMemcpyDeviceToDevice, plans_[hd]->stream()));
        } else {
            const size_t n4 = bytes / sizeof(float4);   // rows are padded to 128-byte multiples
            const int blocks = (int)std::min<size_t>((n4 + 255) / 256, 148 * 8);
            k_copy_peer<<<blocks, 256, 0, plans_[hd]->stream()>>>(reinterpret_cast<float4*>(dst + off),
                                                                   reinterpret_cast<const float4*>(src + off), n4);
        }
        counter += (long long)bytes;
    }

    // will the SOR solve of level k be split into row bands?  (the one rule sor() and sor_rows() share)
    bool sor_is_split(int k, const std::vector<SorRunner<float>::SorPass>& sched, const SorRunner<float>& r) const {
        const int G = (int)devs_.size();
        const int w = plans_[0]->level_w(k), h = plans_[0]->level_h(k), nsor = plans_[0]->n_sor_at(k);
        bool split = G > 1 && (long long)w * h >= split_min_ && r.use_tma && !r.simple_rb && nsor > 0;
        for (auto& ps : sched)
            if (ps.ty.ntiles < G) split = false;          // every device needs at least one tile row
        return split;
    }
    // Rows of the coefficient planes device g's band reads in ANY pass of the split solve of level k (the union of its
    // tile rows' regions, plus the phi row above): the only rows its assembly has to produce.  The whole level when the
    // solve is not split (lo = 0, hi = -1).  PF_MULTI_ASM_SPLIT=0 keeps the redundant whole-level assembly.
    void sor_rows(int k, int g, int& lo, int& hi) {
        typedef SorRunner<float> Runner;
        lo = 0; hi = -1;
        static const bool on_ = [] { const char* e = getenv("PF_MULTI_ASM_SPLIT"); return !(e && !atoi(e)); }();
        if (!on_) return;
        const int G = (int)devs_.size();
        const int w = plans_[0]->level_w(k), h = plans_[0]->level_h(k), nsor = plans_[0]->n_sor_at(k);
        Plan<float>::SorView v = plans_[g]->sor_view();
        std::vector<Runner::SorPass> sched = v.runner->schedule(w, h, nsor);
        if (!sor_is_split(k, sched, *v.runner)) return;
        int a = h, b = 0;
        for (auto& ps : sched) {
            const int tb = (int)((long long)g * ps.ty.ntiles / G), te = (int)((long long)(g + 1) * ps.ty.ntiles / G);
            a = std::min(a, ps.in_lo(tb));
            b = std::max(b, ps.in_hi(te - 1, h));
        }
        lo = a; hi = b;
    }

    void sor(int k) {
        typedef SorRunner<float> Runner;
        const int G = (int)devs_.size();
        const int w = plans_[0]->level_w(k), h = plans_[0]->level_h(k), nsor = plans_[0]->n_sor_at(k);
        std::vector<Plan<float>::SorView> v;
        for (int g = 0; g < G; g++) v.push_back(plans_[g]->sor_view());
        std::vector<Runner::SorPass> sched = v[0].runner->schedule(w, h, nsor);
        const bool split = sor_is_split(k, sched, *v[0].runner);
        if (!split) {
            for (int g = 0; g < G; g++) { on(g); plans_[g]->ph_sor(k); }
            return;
        }
        split_solves_++;
        const int pitch = v[0].args.pitch;
        // Every device must have finished all earlier work on its du/dv buffers (previous solve, its
        // gather, a coarser level solved locally) before a neighbour's first pass starts storing halo
        // rows into them: one event per device, awaited by all others, opens each split solve.
        for (int g = 0; g < G; g++) {
            on(g);
            PF_CUDA(cudaEventRecord(ev_gather_[g], plans_[g]->stream()));
        }
        for (int g = 0; g < G; g++) {
            on(g);
            for (int o = 0; o < G; o++)
                if (o != g) PF_CUDA(cudaStreamWaitEvent(plans_[g]->stream(), ev_gather_[o], 0));
        }
        auto band = [&](const Runner::SorPass& ps, int g, int& tb, int& te) {
            tb = (int)((long long)g * ps.ty.ntiles / G);
            te = (int)((long long)(g + 1) * ps.ty.ntiles / G);
        };
        struct Range { int lo, hi; };
        auto cut = [](Range a, Range b) { return Range{std::max(a.lo, b.lo), std::min(a.hi, b.hi)}; };
        // device-side flags between the passes of this solve: possible when every row a band needs from another band
        // comes from an ADJACENT one (then the producing kernel pushes exactly those rows) and counters are left
        bool use_flags = flags_ok_ && sched.size() > 1 && flag_slot_ + (int)sched.size() <= kFlagSlots;
        if (use_flags && flags_shared_) {
            // bands on one device: at most one pass per band is resident at a time (a band's passes are ordered by its
            // stream), so a waiting pass cannot starve the pass it waits for as long as one pass of EVERY band fits the SMs
            for (size_t p = 0; p < sched.size() && use_flags; p++) {
                int ctas = 0;
                for (int g = 0; g < G; g++) {
                    int tb, te;
                    band(sched[p], g, tb, te);
                    ctas += v[g].runner->grid_for(sched[p], te - tb);
                }
                if (ctas > v[0].runner->sms * v[0].runner->ctas_per_sm) use_flags = false;
            }
        }
        for (size_t p = 0; p + 1 < sched.size() && use_flags; p++)
            for (int g = 0; g < G && use_flags; g++) {
                int tb, te;
                band(sched[p + 1], g, tb, te);
                const Range need_g{sched[p + 1].in_lo(tb), sched[p + 1].in_hi(te - 1, h)};
                for (int o = 0; o < G; o++) {
                    if (o == g || o == g - 1 || o == g + 1) continue;
                    band(sched[p], o, tb, te);
                    const Range r = cut(need_g, Range{sched[p].out_lo(tb), sched[p].out_hi(te - 1, h)});
                    if (r.hi > r.lo) use_flags = false;
                }
            }
        if (use_flags) flag_solves_++;
        for (size_t p = 0; p < sched.size(); p++) {
            const Runner::SorPass& ps = sched[p];
            const bool last = p + 1 == sched.size();
            // rows each device produces in this pass, rows it must hold before its next step (its next
            // pass's input rows, or everything after the last pass), and this pass's buffers
            std::vector<Range> own(G), need(G);
            std::vector<float*> in_u(G), in_v(G), out_u(G), out_v(G);
            for (int g = 0; g < G; g++) {
                int tb, te;
                band(ps, g, tb, te);
                own[g] = Range{ps.out_lo(tb), ps.out_hi(te - 1, h)};
                need[g] = Range{0, h};
                if (!last) {
                    band(sched[p + 1], g, tb, te);
                    need[g] = Range{sched[p + 1].in_lo(tb), sched[p + 1].in_hi(te - 1, h)};
                }
                in_u[g] = *v[g].du; in_v[g] = *v[g].dv; out_u[g] = *v[g].du2; out_v[g] = *v[g].dv2;
            }
            // launch: halo rows wanted by the ADJACENT bands are pushed by the kernel itself
            std::vector<Range> pushed_up(G, Range{0, 0}), pushed_dn(G, Range{0, 0});
            for (int g = 0; g < G; g++) {
                on(g);
                int tb, te;
                band(ps, g, tb, te);
                SorPeer<float> peer;
                // (after the last pass the neighbours need the WHOLE band: the kernel's own stores replace the gather
                // between adjacent bands -- with two GPUs no separate copy is left)
                if (push_) {
                    if (g > 0) {
                        Range r = cut(need[g - 1], own[g]);
                        if (r.hi > r.lo) { peer.up_du = out_u[g - 1]; peer.up_dv = out_v[g - 1]; peer.up_lo = r.lo; peer.up_hi = r.hi; pushed_up[g] = r; }
                    }
                    if (g + 1 < G) {
                        Range r = cut(need[g + 1], own[g]);
                        if (r.hi > r.lo) { peer.dn_du = out_u[g + 1]; peer.dn_dv = out_v[g + 1]; peer.dn_lo = r.lo; peer.dn_hi = r.hi; pushed_dn[g] = r; }
                    }
                    (last ? gather_bytes_ : halo_bytes_) += 2LL * ((peer.up_hi - peer.up_lo) + (peer.dn_hi - peer.dn_lo)) * w * (long long)sizeof(float);
                }
                if (use_flags) {
                    const int slot = flag_slot_ + (int)p;
                    peer.error_word = flags_[(size_t)g] + 2 * kFlagSlots;
                    if (p > 0) {          // the neighbours' previous pass: counters of slot - 1 in OUR memory
                        int nb, ne;
                        if (g > 0) {
                            band(sched[p - 1], g - 1, nb, ne);
                            peer.wait_flag[0] = flags_[(size_t)g] + 2 * (slot - 1) + 0;
                            peer.wait_count[0] = (unsigned)v[g - 1].runner->grid_for(sched[p - 1], ne - nb);
                        }
                        if (g + 1 < G) {
                            band(sched[p - 1], g + 1, nb, ne);
                            peer.wait_flag[1] = flags_[(size_t)g] + 2 * (slot - 1) + 1;
                            peer.wait_count[1] = (unsigned)v[g + 1].runner->grid_for(sched[p - 1], ne - nb);
                        }
                    }
                    if (!last) {          // our arrival: the upper neighbour's "from below" counter, the lower one's "from above"
                        if (g > 0) peer.signal_flag[0] = flags_[(size_t)g - 1] + 2 * slot + 1;
                        if (g + 1 < G) peer.signal_flag[1] = flags_[(size_t)g + 1] + 2 * slot + 0;
                    }
                }
                v[g].runner->launch_pass(v[g].args, ps, in_u[g], in_v[g], out_u[g], out_v[g], tb, te, peer);
                PF_CUDA(cudaEventRecord(ev_pass_[g], plans_[g]->stream()));
            }
            for (int g = 0; g < G; g++) {          // the result of this pass is now in du/dv
                std::swap(*v[g].du, *v[g].du2);
                std::swap(*v[g].dv, *v[g].dv2);
            }
            if (use_flags && !last) continue;   // the next pass waits on the neighbours' counters inside the kernel
            // whatever a device still misses is pulled from its owner; pushes only need the event
            for (int g = 0; g < G; g++) {
                on(g);
                // a device's next pass stores halo rows into its neighbours' buffers, so it must not
                // start before they have finished this pass -- even if it needs no rows from them
                if (g > 0) PF_CUDA(cudaStreamWaitEvent(plans_[g]->stream(), ev_pass_[g - 1], 0));
                if (g + 1 < G) PF_CUDA(cudaStreamWaitEvent(plans_[g]->stream(), ev_pass_[g + 1], 0));
                for (int o = 0; o < G; o++) {
                    if (o == g) continue;
                    Range r = cut(need[g], own[o]);
                    if (r.hi <= r.lo) continue;
                    PF_CUDA(cudaStreamWaitEvent(plans_[g]->stream(), ev_pass_[o], 0));
                    const Range& ps_r = (o == g + 1) ? pushed_up[o] : (o == g - 1 ? pushed_dn[o] : Range{0, 0});
                    if (ps_r.hi > ps_r.lo && ps_r.lo == r.lo && ps_r.hi == r.hi) continue;   // already delivered by o's kernel
                    long long& ctr = last ? gather_bytes_ : halo_bytes_;
                    pull_rows(g, out_u[g], o, out_u[o], r.lo, r.hi, pitch, ctr);
                    pull_rows(g, out_v[g], o, out_v[o], r.lo, r.hi, pitch, ctr);
                }
            }
        }
        if (use_flags) flag_slot_ += (int)sched.size();
    }

    static constexpr int kFlagSlots = 1 << 14;   // passes of all split solves of one pair (two counters each)
    Params P;
    long long split_min_;
    std::vector<int> devs_;
    std::vector<std::unique_ptr<Plan<float>>> plans_;
    std::vector<cudaEvent_t> ev_pass_, ev_gather_;
    long long halo_bytes_ = 0, gather_bytes_ = 0;
    int split_solves_ = 0;
    cudaGraph_t graph_ = nullptr;
    cudaGraphExec_t gexec_ = nullptr;
    cudaEvent_t ev_fork_ = nullptr;
    bool graph_failed_ = false;
    bool push_ = true;     // halo rows stored by the producing kernel into the neighbour (PF_MULTI_PULL=1: copy-engine pulls)
    bool flags_ok_ = false;                 // one GPU per band with peer access: passes ordered by device-side counters
    bool flags_shared_ = false;             // PF_MULTI_FLAGS_SHARED=1 and some bands share a device
    std::vector<unsigned int*> flags_;      // per device: kFlagSlots x {from above, from below}
    int flag_slot_ = 0, flag_solves_ = 0;
};

}  // namespace pf
