// Micro-benchmark: cycles per lock-step iteration of  LDS -> dependent FFMA chain -> STS -> named barrier,
// with and without helper warps that poll shared memory (the structure of k_sor_lex's march).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, int chain) {
    __shared__ float2 D[44][68];
    __shared__ volatile int s_ctr;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    if (tid == 0) s_ctr = 0;
    for (int i = tid; i < 44 * 68; i += blockDim.x) (&D[0][0])[i] = make_float2(0.001f * i, 0.002f * i);
    __syncthreads();
    if (wp < 8) {
        const int li = 8 - wp + lane;
        float2 acc = make_float2(0.f, 0.f);
        int j = -lane - wp;
        long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            const int c = j & 63, c1 = (j + 1) & 63;
            float2 r = D[li][c1], d = D[li + 1][c], u = D[li - 1][c];
            float s1 = acc.x, s2 = acc.y;
            for (int q = 0; q < chain; q++) { s1 = fmaf(s1, 0.5f, r.x + d.x * (q + 1)); s2 = fmaf(s2, 0.5f, u.y + s1); }
            acc = make_float2(s1, s2);
            D[li][c] = acc;
            j++;
            if (MODE & 1) asm volatile("bar.sync 1, 256;" ::: "memory");
            else if (MODE & 4) __syncwarp();
            if ((MODE & 8) && tid == 0) s_ctr = it + 1;
        }
        long long t1 = clock64();
        if (tid == 0) cyc[0] = t1 - t0;
        out[tid] = acc.x + acc.y;
        if (tid == 0) s_ctr = iters;
    } else if (MODE & 2) {
        // helper warps: poll a counter with nanosleep, like the loader / fetcher / comm warps
        unsigned spins = 0;
        while (s_ctr < iters) { if (++spins > 16) __nanosleep(40); }
    }
}
template <int MODE>
void run(const char* name, int threads, int chain) {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
    const int iters = 2000;
    k<MODE><<<1, threads>>>(out, cyc, iters, chain);
    k<MODE><<<1, threads>>>(out, cyc, iters, chain);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-64s chain %2d: %6.1f cycles/iteration (%s)\n", name, chain, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int chain : {1, 6, 12}) {
        run<0>("8 warps, no barrier", 256, chain);
        run<4>("8 warps, __syncwarp only", 256, chain);
        run<1>("8 warps, bar.sync 1,256", 256, chain);
        run<1 | 8>("8 warps, bar.sync + counter store", 256, chain);
        run<1 | 2 | 8>("8 warps + 3 polling helper warps, bar.sync + counter", 352, chain);
    }
    return 0;
}
