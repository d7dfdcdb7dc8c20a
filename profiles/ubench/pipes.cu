// pipes.cu -- micro-benchmark of the integer pipes the chamfer scan leans on (B200, sm_100a).
// Measures warp-instructions per clock per SM for streams of independent chains of one opcode or a mix,
// and two formulations of the 7-candidate (min,+) stencil.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define NCH 8
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, int iters, uint32_t one, uint32_t c0, uint32_t c1)
{
    uint32_t x[NCH], y[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) { x[j] = threadIdx.x * 7 + j + c0; y[j] = threadIdx.x * 13 + j * 3 + c1; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            if (MODE == 0) { x[j] = __viaddmin_u32(x[j], c0, y[j]); }                       // VIADDMNMX
            if (MODE == 1) { x[j] = x[j] * one + c0; }                                      // IMAD
            if (MODE == 2) { x[j] = __viaddmin_u32(x[j], c0, y[j]); y[j] = y[j] * one + c1; } // 1:1 mix
            if (MODE == 3) { x[j] = __vimin3_u32(x[j] ^ c0, y[j], c1); }                    // (LOP3 +) VIMNMX3
            if (MODE == 4) { x[j] = (x[j] & c0) ^ y[j]; }                                   // LOP3
            if (MODE == 5) { x[j] = __viaddmin_u32(x[j], c0, y[j]); y[j] = (y[j] & c1) ^ x[j]; } // VIADDMNMX + LOP3
            if (MODE == 6) { x[j] = __vimin3_u32(x[j], y[j], c1 + j); y[j] += one; }        // VIMNMX3 + IADD
            if (MODE == 7) { x[j] = __vimin3_u32(x[j], y[j], c1 + j); y[j] = y[j] * one + c0; }  // VIMNMX3 + IMAD
            if (MODE == 8) { x[j] = min(x[j] + c0, y[j]); y[j] = y[j] * one + c1; y[j] = y[j] * one + c0; } // 1 ALU : 2 FMA
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < NCH; ++j) s += x[j] ^ y[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the forward stencil of the scan on a register row of P pixels, two formulations
template <int P, int MODE>
__global__ void __launch_bounds__(128) stencil(uint32_t* out, int iters, uint32_t one, uint32_t seed)
{
    uint32_t A[P + 4], B[P + 4];
#pragma unroll
    for (int i = 0; i < P + 4; ++i) { A[i] = (threadIdx.x * 31 + i * 7 + seed) << 21; B[i] = (threadIdx.x * 17 + i * 5 + seed) << 21; }
    const uint32_t K3 = 3u << 21, K2 = 2u << 21, K1 = 1u << 21, O = 1u << 17;
    for (int it = 0; it < iters; ++it) {
        uint32_t c[P];
#pragma unroll
        for (int i = 0; i < P; ++i) {
            if (MODE == 0) {       // 1 IMAD + 6 VIADDMNMX
                uint32_t m = B[i + 1] * one + K3;
                m = __viaddmin_u32(B[i + 3], K3 + 2 * O, m);
                m = __viaddmin_u32(A[i], K3 + 4 * O, m);
                m = __viaddmin_u32(A[i + 1], K2 + 6 * O, m);
                m = __viaddmin_u32(A[i + 2], K1 + 8 * O, m);
                m = __viaddmin_u32(A[i + 3], K2 + 10 * O, m);
                m = __viaddmin_u32(A[i + 4], K3 + 12 * O, m);
                c[i] = m;
            } else if (MODE == 1) {  // 7 IMAD + 3 VIMNMX3
                const uint32_t t0 = B[i + 1] * one + K3, t1 = B[i + 3] * one + (K3 + 2 * O), t2 = A[i] * one + (K3 + 4 * O);
                const uint32_t t3 = A[i + 1] * one + (K2 + 6 * O), t4 = A[i + 2] * one + (K1 + 8 * O);
                const uint32_t t5 = A[i + 3] * one + (K2 + 10 * O), t6 = A[i + 4] * one + (K3 + 12 * O);
                c[i] = __vimin3_u32(__vimin3_u32(t0, t1, t2), __vimin3_u32(t3, t4, t5), t6);
            } else if (MODE == 2) {  // 3 IMAD + 4 VIADDMNMX + 1 VIMNMX3 ... mixed
                const uint32_t t0 = B[i + 1] * one + K3, t2 = A[i] * one + (K3 + 4 * O), t4 = A[i + 2] * one + (K1 + 8 * O);
                const uint32_t m0 = __viaddmin_u32(B[i + 3], K3 + 2 * O, t0);
                const uint32_t m1 = __viaddmin_u32(A[i + 1], K2 + 6 * O, t2);
                const uint32_t m2 = __viaddmin_u32(A[i + 3], K2 + 10 * O, t4);
                c[i] = __viaddmin_u32(A[i + 4], K3 + 12 * O, __vimin3_u32(m0, m1, m2));
            } else {                 // shared adds: rows pre-incremented once per row (+1,+2,+3 copies) then min3 only
                // a[i]+1, a[i]+2, a[i]+3 computed once per pixel of the row (3 IMAD) serve 5 consumers; orders differ per
                // consumer so this is NOT exact -- throughput probe only
                const uint32_t t0 = B[i + 1] * one + K3, t1 = B[i + 3] * one + K3;
                const uint32_t p1 = A[i + 2] * one + K1, p2 = A[i + 1] * one + K2, p3 = A[i] * one + K3;
                c[i] = __vimin3_u32(__vimin3_u32(t0, t1, p3), __vimin3_u32(p2, p1, A[i + 3] + K2), A[i + 4] + K3);
            }
        }
#pragma unroll
        for (int i = 0; i < P; ++i) { B[i + 2] = A[i + 2]; A[i + 2] = c[i]; }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < P + 4; ++i) s += A[i] ^ B[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main()
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t* out; cudaMalloc(&out, 64 << 20);
    const int iters = 4096, grid = sms * 8, block = 256;      // 64 warps/SM
    const char* names[] = {"VIADDMNMX", "IMAD", "VIADDMNMX+IMAD 1:1", "LOP3+VIMNMX3", "LOP3", "VIADDMNMX+LOP3", "VIMNMX3+IADD",
                           "VIMNMX3+IMAD", "1 VIADDMNMX : 2 IMAD"};
    const int per[] = {1, 1, 2, 2, 1, 2, 2, 2, 3};
#define RUN(M) { float ms = timeit([&] { k<M><<<grid, block>>>(out, iters, 1u, 3u << 21, 5u << 17); }); \
    double wi = (double)grid * (block / 32) * iters * NCH * per[M]; \
    printf("%-24s %8.3f ms  %6.2f warp-instr/clk/SM (at %d MHz nominal)\n", names[M], ms, wi / (ms * 1e-3) / (khz * 1e3) / sms, khz / 1000); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
    const char* sn[] = {"stencil 1 IMAD + 6 VIADDMNMX", "stencil 7 IMAD + 3 VIMNMX3", "stencil 3 IMAD + 4 VIADDMNMX + 1 VIMNMX3", "probe 5 IMAD + 3 VIMNMX3 + 2 IADD"};
#define RUNS(M, W) { const int g2 = sms * W; float ms = timeit([&] { stencil<20, M><<<g2, 128>>>(out, 2048, 1u, 7u); }); \
    double px = (double)g2 * 128 * 2048 * 20; \
    printf("%-44s warps/SM %2d  %8.3f ms  %6.2f px/clk/SM\n", sn[M], W * 4, ms, px / (ms * 1e-3) / (khz * 1e3) / sms); }
    RUNS(0, 2) RUNS(1, 2) RUNS(2, 2) RUNS(3, 2)
    RUNS(0, 4) RUNS(1, 4) RUNS(2, 4) RUNS(3, 4)
    RUNS(0, 1) RUNS(1, 1) RUNS(2, 1) RUNS(3, 1)
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
