// k_sor_lex: the reference's lexicographic Gauss-Seidel/SOR order (S/OpticalFlow.cpp:451-505), exactly, as a
// band-march over a time-skewed iteration space -- the fast replacement for the grid-synchronised wavefront.
//
// Pixel (i,j) of sweep s needs (i,j-1),(i-1,j) of sweep s and (i,j+1),(i+1,j),(i,j) of sweep s-1, so it can run at
// wavefront time t = i + j + 2s.  The iteration space is cut so that every dependence points the same way:
//
//   band I of sweep s  = image rows [32 I - s, 32 I + 32 - s)          (the band slides up one row per sweep)
//   CTA (I, K)         = band I of the NS consecutive sweeps s = K NS + k, k = 0..NS-1: one warp per sweep,
//                        lane l = row 32 I - s + l, marching along the columns
//   CTA step tau       : warp k, lane l updates column j = tau - l - k
//
// Inside a CTA every dependence is met by the previous step (left: own register, up: lane l-1 by shuffle, right /
// down / own old value: warp k-1, through a shared-memory ring that is updated in place), so the NS compute warps
// run in lock step with one named barrier per step.  Between CTAs all dependences point to (I-1, K) and (I, K-1)
// (and (I-1, K-1)): no cycle, whatever the grain of the hand-off.  They travel through the du/dv planes themselves:
// a warp stores a result to the plane when the CTA will not touch that pixel again (lane 31, whose row leaves the
// band with the next sweep, and every lane of the group's last sweep); lane 0 of every warp reads its upper
// neighbour (sweep s) and its own / right-hand old value (sweep s-1) from the plane, written there by band I-1; the
// first warp of a group reads all its old values from the plane, written by group K-1.  Because a pixel's version
// s is consumed by exactly the updates that precede its version s+1 in the dependence order, ONE copy per pixel --
// the plane, in place -- is enough.
//
// Hand-off: per (band, sweep) a progress word P = completed CTA steps - k in global memory.  A helper warp
// ("comm") publishes the CTA's progress (one gpu-scope fence per publication, off the compute warps' path) and
// polls the (at most NS + 2) words this CTA depends on, turning them into the highest CTA step whose plane reads
// are safe.  Plane reads are issued PD steps ahead of their use and bypass L1.  A second helper warp ("loader")
// streams the six read-only coefficient planes of the band into a 64-column shared ring with cp.async (zero fill
// outside the image), ahead of the march.  CTAs draw tickets in wavefront order (I + K), so a CTA only ever waits
// for CTAs with smaller tickets: no co-residency assumption, no deadlock.
//
// Critical path: W + 31 + NS steps of march + ~(32 + hand-off) per band + ~(NS + hand-off) per sweep group, against
// the (W + H + 2 nsor) grid-wide barriers of k_sor_wavefront.
#pragma once
#ifndef PF_LEX_F32_FOLD
#define PF_LEX_F32_FOLD 0
#endif
#include <type_traits>
#include "common.cuh"
#include "sor.cuh"

namespace pf {

constexpr int kLexInf = 0x3fffffff;
constexpr int kLexFlagHeader = 4;   // words: [0] ticket counter, [1] abort, [2..3] spare; then P[band][sweep]

template <typename T>
struct LexArgs {
    const T *phi, *dxy, *iu, *iv, *bu, *bv;
    T *du, *dv;          // in/out, zero on entry
    int w, h, pitch;
    T alpha, omega;
    int nsor, NI, NK;
    int* flags;          // zeroed before every launch
    int* err;            // sticky: set when a hand-off timed out (never by a healthy run); checked by the host
    int pub_every;       // publish progress every this many steps (and at the end)
    int opt;             // developer knobs (PF_LEX_OPT): bit 0 = __threadfence instead of a release store, bit 1 = no fence at all
                         // (timing experiment only, results undefined), bit 2 = plain bar.sync per step
    long long* stats;    // developer build of the kernel only (PF_LEX_STATS=1): per CTA {I, K, start, end, wait plane, wait loader}
};

template <typename T, int NS>
struct LexCfg {
    static constexpr int RH = 32 + NS;        // local rows: row 0 = global row 32 I - K NS - NS
    static constexpr int RING = 64;           // columns resident per row
    static constexpr int RS = RING + 4;       // row stride: the skewed accesses (row l, column c - l) hit distinct banks
    static constexpr int PD = 4;              // plane reads are issued this many steps before they are used
    static constexpr int CW = 8;              // columns per loader chunk
    static constexpr int VEC = 16 / (int)sizeof(T);
    static constexpr int RHD = NS + 33;       // du/dv ring rows: one more, the row below the band (old value, group K-1)
    static constexpr int FB = 8;              // diagonals per fetcher batch (loads in flight together: the batch costs one L2 round trip)
    static constexpr int FAHEAD = 24;         // the fetcher may run this many steps ahead of the march (ring slots must be dead)
    static constexpr int THREADS = (NS + 3) * 32;
    static constexpr size_t smem_bytes() { return 2 * sizeof(T) * (3 * RH + RHD) * RS; }
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// named barrier over `nthreads` threads that also ORs a flag across them (uniform "stop" decision)
__device__ __forceinline__ int bar_red_or_named(int id, int nthreads, int v) {
    int r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.s32 q, %3, 0;\n\t"
        "bar.red.or.pred p, %1, %2, q;\n\t"
        "selp.s32 %0, 1, 0, p;\n\t}"
        : "=r"(r)
        : "r"(id), "r"(nthreads), "r"(v)
        : "memory");
    return r;
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {   // MEMBAR.ALL.GPU + store, no L1 invalidation
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_gpu(int* p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename T, int NS, bool STATS = false>
__global__ void __launch_bounds__((NS + 3) * 32, 1) k_sor_lex(LexArgs<T> a) {
    typedef LexCfg<T, NS> Cfg;
    constexpr int RH = Cfg::RH, RHD = Cfg::RHD, RS = Cfg::RS, RING = Cfg::RING, PD = Cfg::PD;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef typename Vec2<T>::type V2;
    // coefficient rings as PAIRS, so that a pixel costs three 8/16-byte reads instead of six scalar ones:
    // [0] = {phi, dxy}, [1] = {iu, iv}, [2] = {bu, bv}; and the {du, dv} ring, updated in place
    V2(*C)[RH][RS] = reinterpret_cast<V2(*)[RH][RS]>(smem_raw);
    V2(*D)[RS] = reinterpret_cast<V2(*)[RS]>(smem_raw + sizeof(V2) * 3 * RH * RS);
    __shared__ volatile int s_ctr, s_loaded, s_avail, s_abort, s_fetched;
    __shared__ int s_I, s_K;

    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int W = a.w, H = a.h, P = a.pitch;

    if (tid == 0) {
        // tickets in wavefront order T = I + K (then by I): every CTA this one waits for holds a smaller ticket
        int t = atomicAdd(a.flags, 1);
        int I = 0, K = 0;
        for (int d = 0; d < a.NI + a.NK - 1; d++) {
            const int lo = max(0, d - (a.NK - 1)), hi = min(d, a.NI - 1), n = hi - lo + 1;
            if (t < n) { I = lo + t; K = d - I; break; }
            t -= n;
        }
        s_I = I; s_K = K;
        s_ctr = 0; s_loaded = 0; s_avail = -1; s_abort = 0; s_fetched = 0;
    }
    __syncthreads();
    const int I = s_I, K = s_K;
    const int s_first = K * NS;
    const int nk = min(NS, a.nsor - s_first);
    const int rb = I * 32 - s_first - NS;          // global row of local row 0
    const int nsteps = W + 30 + nk;
    int* const myflag = a.flags + kLexFlagHeader + (size_t)I * a.nsor + s_first;

    // rows of the band over all its sweeps: [32 I - s_first - nk + 1, 32 I - s_first + 31]
    if (I * 32 - s_first + 31 < 0 || I * 32 - s_first - nk + 1 >= H) {
        if (wp == NS && lane < nk) st_relaxed_gpu(myflag + lane, kLexInf);   // nothing to do and nothing stored
        return;
    }

    long long st_t0 = 0, st_g0 = 0, st_wa = 0, st_wl = 0;
    if (STATS && tid == 0) {
        st_t0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(st_g0));
    }
    if (wp < NS) {
        // ------------------------------------------------------------------------------------ compute warps
        // Shared memory and registers only.  Every neighbour value comes from the du/dv ring: the lower and right old
        // values were put there by warp k-1 (or by the fetcher warp for warp 0 and for lane 0), the upper new value
        // by this warp's lane l-1 in the previous step (fetcher for lane 0).  The step body is branch free and
        // software pipelined: the coefficients of step tau+1 are read during step tau, so that after the step barrier
        // only the ring reads, the update chain and the ring write remain on the critical path.  Waits (fetcher,
        // loader) are taken once per PD steps.
        const int k = wp;
        const bool live = k < nk;
        const int s = s_first + k;
        const int li = NS - k + lane;              // local row of this lane's pixels
        const int gi = rb + li;                    // = 32 I - s + lane
        const bool rowok = live && gi >= 0 && gi < H;
        const bool zero_old = s == 0;              // the solve starts from du = dv = 0
        const bool has_dn = rowok && gi + 1 < H && !zero_old;
        const bool has_up = rowok && gi > 0;
        const bool st_plane = (k == nk - 1 || lane == 31) && !(a.opt & 8);   // opt bit 3: timing experiment without plane stores
        const T one_m = (T)1 - a.omega, nalpha = -a.alpha;
        const size_t row_o = (size_t)(rowok ? gi : 0) * P;
        constexpr int CP = RH * RS;   // stride between the coefficient pair arrays
        const V2* const Crow = &C[0][li][0];
        V2* const Drow = &D[li][0];
        T* const ps_du = a.du + row_o;
        T* const ps_dv = a.dv + row_o;
        // every lane of the warp has a row, an upper and a lower neighbour, and an old value: with the columns inside the
        // image too, a step needs no predicate at all (the fast body below)
        const bool full_rows = __all_sync(0xffffffffu, rowok && gi > 0 && gi + 1 < H && !zero_old);

        int fetched = 0, loaded = 0;
        bool ok = true;
        // diagonals < need_f are in the ring, coefficient columns <= need_c are resident
        auto wait_for = [&](int need_f, int need_c) {
            long long w0 = 0;
            if (fetched < need_f) {
                if (STATS && tid == 0) w0 = clock64();
                unsigned spins = 0;
                while ((fetched = s_fetched) < need_f) {
                    if (s_abort) { ok = false; break; }
                    if (++spins > 2) __nanosleep(64);   // waiting warps must not take issue slots from the helper they wait for
                }
                if (STATS && tid == 0) st_wa += clock64() - w0;
            }
            if (loaded <= need_c) {
                if (STATS && tid == 0) w0 = clock64();
                unsigned spins = 0;
                while ((loaded = s_loaded) <= need_c) {
                    if (s_abort) { ok = false; break; }
                    if (++spins > 2) __nanosleep(64);   // waiting warps must not take issue slots from the helper they wait for
                }
                if (STATS && tid == 0) st_wl += clock64() - w0;
            }
            // The helpers publish their data before their counters (fence on their side); on this side shared-memory reads
            // of one thread are served in order, so a compiler barrier is all that is needed -- a MEMBAR here would also
            // wait for this thread's plane stores in flight (an L2 round trip per block of steps).
            asm volatile("" ::: "memory");
        };
        const int nsteps4 = (nsteps + PD - 1) / PD * PD;   // the padding steps touch nothing (every column is past W)
        wait_for(1, 0);
        T cw_prev = 0, ru_prev = 0, rv_prev = 0;   // phi, du, dv of this row's previous column (zero before column 0)
        T oc_du = 0, oc_dv = 0;                    // old value of the current pixel = last step's right value
        if (k == 0 && lane == 0 && rowok && !zero_old) {   // the one thread whose column 0 has no previous step
            const V2 t = Drow[0];
            oc_du = t.x; oc_dv = t.y;
        }
        // coefficients of the current step (column j); garbage while the lane is outside the image, never used then
        int j = -lane - k;
        V2 n_a, n_b, n_c;   // {phi, dxy}, {iu, iv}, {bu, bv}
        T n_wu;
        {
            const int c = j & (RING - 1);
            n_a = Crow[c]; n_b = Crow[CP + c]; n_c = Crow[2 * CP + c]; n_wu = Crow[c - RS].x;
        }
        // the update of one pixel, in the reference's operation order (missing neighbours contribute +0)
        auto update = [&](T cw, T wl, T wu, V2 r, V2 u, V2 d, T& nu, T& nv) {
            if constexpr (sizeof(T) == 4 && PF_LEX_F32_FOLD) {
                // FP32 experiment (-DPF_LEX_F32_FOLD=1; off): everything that does not need this step's ring reads is
                // folded beforehand, leaving four dependent operations for du and one more for dv (each dependent FMA
                // costs ~8 cycles with two lock-stepped warps per scheduler): 393 instead of 415 cycles per step.  Not the
                // default because the coarse levels of config 5 sit next to a bifurcation: with this association 5 of
                // 8.3 M pixels end 1.2 px from the reference (tools/hybrid_repro.py), with the reference's own operation
                // order -- the code below, FP32 rounding being the only difference -- every pixel stays within 0.2 px.
                const T au = one_m * oc_du + n_b.x * (n_c.x - n_a.y * oc_dv);   // no ring value in here
                const T av = one_m * oc_dv + n_b.y * n_c.y;
                const T ku = n_b.x * a.alpha, kv = n_b.y * a.alpha;
                const T l1 = wl * ru_prev, l2 = wl * rv_prev;
                const T t1 = fmaf(wu, u.x, fmaf(cw, r.x + d.x, l1));
                const T t2 = fmaf(wu, u.y, fmaf(cw, r.y + d.y, l2));
                nu = fmaf(ku, t1, au);
                nv = fmaf(-(n_b.y * n_a.y), nu, fmaf(kv, t2, av));
                return;
            }
            T s1 = 0, s2 = 0;
            s1 += wl * ru_prev; s2 += wl * rv_prev;
            s1 += cw * r.x;     s2 += cw * r.y;
            s1 += wu * u.x;     s2 += wu * u.y;
            s1 += cw * d.x;     s2 += cw * d.y;
            s1 *= nalpha;
            s2 *= nalpha;
            s1 += n_a.y * oc_dv;
            nu = one_m * oc_du + n_b.x * (n_c.x - s1);
            s2 += n_a.y * nu;
            nv = one_m * oc_dv + n_b.y * (n_c.y - s2);
        };
        bool stop = false;
        for (int tau0 = 0; !stop && tau0 < nsteps4; tau0 += PD) {
            // this block consumes the diagonals tau0 .. tau0+PD-1; its last step preloads column tau0 + PD
            wait_for(min(tau0 + PD, nsteps), min(tau0 + PD, W - 1));
            if (STATS && tid == 0 && (tau0 & 63) == 0 && tau0 < 512) {   // timeline: when the march passes steps 0, 64, 128, ...
                long long g;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
                a.stats[16 * (size_t)(K * a.NI + I) + 8 + (tau0 >> 6)] = g;
            }
            // all 32 lanes inside the image with a right-hand neighbour for the whole block (lane 31 is the last to enter,
            // lane 0 the first to reach the last column)?
            const bool fast = full_rows && tau0 - 31 - k >= 0 && tau0 + PD - k < W;
            if (fast) {
#pragma unroll
                for (int q = 0; q < PD; q++) {
                    const int tau = tau0 + q;
                    const int c = j & (RING - 1), c1 = (j + 1) & (RING - 1);
                    const V2 r = Drow[c1], d = Drow[RS + c], u = Drow[c - RS];
                    T nu, nv;
                    update(n_a.x, cw_prev, n_wu, r, u, d, nu, nv);
                    Drow[c] = V2{nu, nv};
                    if (st_plane) { ps_du[j] = nu; ps_dv[j] = nv; }
                    cw_prev = n_a.x;
                    ru_prev = nu; rv_prev = nv;
                    oc_du = r.x; oc_dv = r.y;
                    n_a = Crow[c1]; n_b = Crow[CP + c1]; n_c = Crow[2 * CP + c1]; n_wu = Crow[c1 - RS].x;
                    j++;
                    if (q == PD - 1) {
                        stop = bar_red_or_named(1, NS * 32, ok ? 0 : 1) != 0;   // an abort seen by anyone stops all
                        // progress once per block: a store after every barrier costs ~50 cycles per step (tools/bar_micro.cu)
                        if (tid == 0) s_ctr = (stop || tau + 1 >= nsteps) ? nsteps : tau + 1;
                    } else {
                        asm volatile("bar.sync 1, %0;" ::"r"(NS * 32) : "memory");
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < PD; q++) {
                    const int tau = tau0 + q;
                    const bool act = rowok && (unsigned)j < (unsigned)W;
                    const bool rvalid = rowok && !zero_old && (unsigned)(j + 1) < (unsigned)W;
                    const int c = j & (RING - 1), c1 = (j + 1) & (RING - 1);
                    V2 r = Drow[c1], d = Drow[RS + c], u = Drow[c - RS];
                    if (!rvalid) r = V2{0, 0};
                    if (!(has_dn && act)) d = V2{0, 0};
                    if (!(has_up && act)) u = V2{0, 0};
                    const T cw = act ? n_a.x : (T)0;
                    T nu, nv;
                    update(cw, cw_prev, n_wu, r, u, d, nu, nv);
                    if (act) Drow[c] = V2{nu, nv};
                    if (act && st_plane) { ps_du[j] = nu; ps_dv[j] = nv; }
                    nu = act ? nu : (T)0;
                    nv = act ? nv : (T)0;
                    cw_prev = cw;
                    ru_prev = nu; rv_prev = nv;
                    oc_du = r.x; oc_dv = r.y;
                    n_a = Crow[c1]; n_b = Crow[CP + c1]; n_c = Crow[2 * CP + c1]; n_wu = Crow[c1 - RS].x;
                    j++;
                    if (q == PD - 1) {
                        stop = bar_red_or_named(1, NS * 32, ok ? 0 : 1) != 0;
                        if (tid == 0) s_ctr = (stop || tau + 1 >= nsteps) ? nsteps : tau + 1;
                    } else {
                        asm volatile("bar.sync 1, %0;" ::"r"(NS * 32) : "memory");
                    }
                }
            }
        }
        if (STATS && tid == 0) {
            long long* o = a.stats + 16 * (size_t)(K * a.NI + I);
            long long g1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
            o[0] = I; o[1] = K; o[2] = st_g0; o[3] = g1; o[4] = st_wa; o[5] = st_wl; o[6] = nsteps; o[7] = clock64() - st_t0;
        }
        return;
    }

    if (wp == NS) {
        // ------------------------------------------------------------------------------------ comm warp
        // lanes 0..NS: P of band I-1, sweeps s_first-1 .. s_first+NS-1;  lane NS+1: P of this band's sweep s_first-1
        const int* dep = nullptr;
        if (lane <= NS && I > 0 && s_first - 1 + lane >= 0 && lane <= nk)
            dep = a.flags + kLexFlagHeader + (size_t)(I - 1) * a.nsor + (s_first - 1 + lane);
        if (lane == NS + 1 && K > 0) dep = a.flags + kLexFlagHeader + (size_t)I * a.nsor + (s_first - 1);
        int published = 0;
        long long t0 = clock64();
        for (;;) {
            int bound = kLexInf;
            if (dep) {
                const int f = ld_relaxed_gpu(dep);
                if (lane == NS + 1) bound = f - 1;                       // group K-1: P >= tau + 1
                else {
                    if (lane >= 1 && lane - 1 < nk) bound = min(bound, f + (lane - 1) - 32);   // upper neighbour of warp lane-1
                    if (lane < nk) bound = min(bound, f + lane - 33);                          // right / own old value of warp lane
                }
            }
            int ab = (lane == 31) ? ld_relaxed_gpu(a.flags + 1) : 0;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                bound = min(bound, __shfl_xor_sync(0xffffffffu, bound, o));
                ab |= __shfl_xor_sync(0xffffffffu, ab, o);
            }
            const int ctr = s_ctr;
            if (clock64() - t0 > 8000000000ll) ab = 1;                   // ~4 s: a lost dependence must not hang the GPU
            if (ab) {
                if (lane == 0) { st_relaxed_gpu(a.flags + 1, 1); st_relaxed_gpu(a.err, 1); s_abort = 1; }
                break;
            }
            if (lane == 0) s_avail = bound;
            if (ctr >= published + a.pub_every || (ctr >= nsteps && ctr > published)) {
                // release: the compute warps' plane stores (ordered before s_ctr by their barrier) before the progress word
                const int v = ctr >= nsteps ? kLexInf : ctr - lane;
                if (a.opt & 1) { __threadfence(); if (lane < nk) st_relaxed_gpu(myflag + lane, v); }
                else if (a.opt & 2) { if (lane < nk) st_relaxed_gpu(myflag + lane, v); }
                else if (lane < nk) st_release_gpu(myflag + lane, v);
                published = ctr;
                t0 = clock64();
            }
            if (ctr >= nsteps) break;
        }
        return;
    }

    if (wp == NS + 2) {
        // ------------------------------------------------------------------------------------ fetcher warp
        // Diagonal tau = everything step tau of the march reads that another CTA produced, copied from the planes (L2)
        // into the du/dv ring:
        //   task t <= NS  : local row t (the row above / the top row of warp NS-1-t / NS-t), column tau - (NS-1-t),
        //                   written by band I-1 (sweep s_first + NS-1-t); unused rows of a short group are skipped
        //   task NS+1+m   : local row NS+1+m (m = 0..31, the band of warp 0 shifted down by one), column tau - m,
        //                   written by group K-1 -- the old values of the group's first sweep
        // FB diagonals per batch (loads in flight together), at most FAHEAD steps ahead of the march so that the ring
        // slots overwritten are dead, never beyond what the comm warp has seen published.
        constexpr int FB = Cfg::FB, NT = NS + 33;
        int row[2], off[2];          // this lane's two tasks: ring row (or -1) and column offset against the diagonal
        size_t grow[2];
#pragma unroll
        for (int h2 = 0; h2 < 2; h2++) {
            const int t = h2 * 32 + lane;
            int r = -1, o = 0;
            if (t <= NS) {
                if (I > 0 && t >= NS - nk && s_first + NS - 1 - t >= 0) { r = t; o = NS - 1 - t; }
            } else if (t < NT) {
                if (K > 0) { r = t; o = t - NS - 1; }
            }
            const int g = rb + (r >= 0 ? r : 0);
            if (r >= 0 && (g < 0 || g >= H)) r = -1;
            row[h2] = r; off[h2] = o; grow[h2] = (size_t)(g < 0 ? 0 : g) * P;
        }
        // the old value of pixel (top row of warp 0, column 0): no step reads it as a right-hand value
        if (lane == 0 && K > 0 && I > 0 && rb + NS >= 0 && rb + NS < H) {
            unsigned spins = 0;
            while (s_avail < 0) {
                if (s_abort) return;
                if (++spins > 1) __nanosleep(100);
            }
            __threadfence_block();
            D[NS][0] = V2{__ldcg(a.du + (size_t)(rb + NS) * P), __ldcg(a.dv + (size_t)(rb + NS) * P)};
        }
        __syncwarp();
        for (int t0 = 0; t0 < nsteps;) {
            // batch = the diagonals that are safe right now, at most FB: one diagonal at a time while this CTA follows
            // its producers closely (latency), full batches when it is behind (throughput)
            int n = 0;
            unsigned spins = 0;
            for (;;) {
                const int av = min(min(s_avail, s_ctr + Cfg::FAHEAD), nsteps - 1);
                n = __shfl_sync(0xffffffffu, min(FB, av - t0 + 1), 0);   // one view of the counters for the whole warp
                if (n > 0) break;
                if (s_abort) return;
                if (++spins > 1) __nanosleep(40);
            }
            const int last = t0 + n - 1;
            __threadfence_block();
            T vu[FB][2], vv[FB][2];
#pragma unroll
            for (int f = 0; f < FB; f++)
#pragma unroll
                for (int h2 = 0; h2 < 2; h2++) {
                    const int x = t0 + f - off[h2];
                    const bool on = row[h2] >= 0 && f < n && (unsigned)x < (unsigned)W;
                    vu[f][h2] = 0; vv[f][h2] = 0;
                    if (on) vu[f][h2] = __ldcg(a.du + grow[h2] + x);
                    if (on) vv[f][h2] = __ldcg(a.dv + grow[h2] + x);
                }
#pragma unroll
            for (int f = 0; f < FB; f++)
#pragma unroll
                for (int h2 = 0; h2 < 2; h2++) {
                    const int x = t0 + f - off[h2];
                    const bool on = row[h2] >= 0 && f < n && (unsigned)x < (unsigned)W;
                    if (on) D[row[h2]][x & (RING - 1)] = V2{vu[f][h2], vv[f][h2]};
                }
            __syncwarp();
            __threadfence_block();
            if (lane == 0) s_fetched = last + 1;
            t0 = last + 1;
        }
        return;
    }

    // ---------------------------------------------------------------------------------------- loader warp
    // The six coefficient planes of the band, 8 columns at a time, ahead of the march: 16-byte loads per plane and row
    // piece, interleaved pairwise in registers ({phi, dxy}, {iu, iv}, {bu, bv}) and stored as pairs; rows outside the image
    // are zero ("phantom pixels": the missing upper neighbour of row 0 gets weight 0 without a test).
    {
        constexpr int CW = Cfg::CW, VEC = Cfg::VEC, GPC = CW / VEC, RPI = 32 / GPC;   // pieces per row and chunk, rows per pass
        constexpr int NIT = (RH + RPI - 1) / RPI;
        typedef typename std::conditional<sizeof(T) == 4, float4, double2>::type LV;   // 16 bytes
        const T* planes[6] = {a.phi, a.dxy, a.iu, a.iv, a.bu, a.bv};
        const int nchunks = (W + CW - 1) / CW;
        const int g = lane % GPC, rsub = lane / GPC;
        for (int ch = 0; ch < nchunks; ch++) {
            // columns [CW ch - RING, ...) must be dead: the last reader is warp NS-1, lane 31
            const int need = CW * ch + CW - RING + 31 + NS;
            unsigned spins = 0;
            while (s_ctr < need) {
                if (s_abort) return;
                if (++spins > 1) __nanosleep(100);
            }
            const int c0 = CW * ch, cs = c0 & (RING - 1);
            LV v[NIT][6];
#pragma unroll
            for (int it = 0; it < NIT; it++) {
                const int r = it * RPI + rsub, gr = rb + r;
                const bool in = r < RH && gr >= 0 && gr < H;
#pragma unroll
                for (int p = 0; p < 6; p++) {
                    if (in) v[it][p] = __ldg(reinterpret_cast<const LV*>(planes[p] + (size_t)gr * P + c0 + g * VEC));
                    else v[it][p] = LV{};
                }
            }
#pragma unroll
            for (int it = 0; it < NIT; it++) {
                const int r = it * RPI + rsub;
                if (r < RH) {
#pragma unroll
                    for (int pp = 0; pp < 3; pp++) {
                        const T* x = reinterpret_cast<const T*>(&v[it][2 * pp]);
                        const T* y = reinterpret_cast<const T*>(&v[it][2 * pp + 1]);
                        V2* dst = &C[pp][r][cs + g * VEC];
#pragma unroll
                        for (int e = 0; e < VEC; e++) dst[e] = V2{x[e], y[e]};
                    }
                }
            }
            __syncwarp();
            __threadfence_block();
            if (lane == 0) s_loaded = ch + 1 < nchunks ? CW * (ch + 1) : kLexInf;
        }
    }
}

// bands and sweep groups of a level
template <int NS>
inline void lex_grid(int h, int nsor, int& NI, int& NK) {
    NI = (h + nsor - 2) / 32 + 1;
    NK = ceil_div(nsor, NS);
}

}  // namespace pf
