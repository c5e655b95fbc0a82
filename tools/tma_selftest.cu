// Developer self-test for the TMA plumbing in csrc/tma.cuh (run on the GPU box).
// usage: tma_selftest <mode> <box_w> <box_h> <x> <y>   mode 2: descriptor in param, 3: in global
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../papteam_opticalflow_b200/csrc/tma.cuh"
using namespace pf;

struct Maps { CUtensorMap a; };

__global__ void k(const __grid_constant__ Maps m, const CUtensorMap* gm, float* out, int* flag, int mode, int bw, int bh, int x, int y) {
    extern __shared__ unsigned char raw[];
    float* tile = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(raw) + 127) & ~(uintptr_t)127);
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (uint32_t)(bw * bh * sizeof(float)));
        tma_load_2d(tile, mode == 2 ? &m.a : gm, x, y, &bar);
    }
    // bounded wait that reports instead of trapping
    uint32_t done = 0, spins = 0;
    while (!done && spins < (1u << 22)) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(&bar)), "r"(0u) : "memory");
        spins++;
    }
    if (threadIdx.x == 0) *flag = done ? 1 : -1;
    __syncthreads();
    if (done) for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char** argv) {
    int mode = argc > 1 ? atoi(argv[1]) : 2, BW = argc > 2 ? atoi(argv[2]) : 68, BH = argc > 3 ? atoi(argv[3]) : 33;
    int x = argc > 4 ? atoi(argv[4]) : -2, y = argc > 5 ? atoi(argv[5]) : -1;
    const int W = 100, H = 50, P = 128;
    std::vector<float> h((size_t)P * H);
    for (int i = 0; i < H; i++) for (int j = 0; j < P; j++) h[(size_t)i * P + j] = j < W ? i * 1000.f + j : -7.f;
    float *d, *dout; int* dflag;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&dout, BW * BH * 4); cudaMalloc(&dflag, 4);
    cudaMemset(dflag, 0, 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    Maps m;
    try { m.a = make_plane_map(d, W, H, P, BW, BH); } catch (const Error& e) { printf("encode failed: %s\n", e.what()); return 2; }
    CUtensorMap* gm; cudaMalloc(&gm, sizeof(CUtensorMap)); cudaMemcpy(gm, &m.a, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    k<<<1, 128, 100000>>>(m, gm, dout, dflag, mode, BW, BH, x, y);
    cudaError_t e = cudaDeviceSynchronize();
    int flag = 0;
    if (e == cudaSuccess) cudaMemcpy(&flag, dflag, 4, cudaMemcpyDeviceToHost);
    printf("mode %d box %dx%d at (%d,%d): %s, barrier %s\n", mode, BW, BH, x, y, cudaGetErrorString(e),
           flag == 1 ? "completed" : flag == -1 ? "TIMED OUT" : "n/a");
    if (e != cudaSuccess || flag != 1) return 1;
    std::vector<float> o(BW * BH);
    cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < BH; i++) for (int j = 0; j < BW; j++) {
        int Y = i + y, X = j + x;
        float want = (X >= 0 && X < W && Y >= 0 && Y < H) ? Y * 1000.f + X : 0.f;
        if (o[i * BW + j] != want) { if (bad < 3) printf("  mismatch at (%d,%d): %f vs %f\n", i, j, o[i * BW + j], want); bad++; }
    }
    printf("  box check: %d mismatches\n", bad);
    return bad != 0;
}
