#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint64_t pol_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__global__ void k_ld_hint(const float* in, float* out) {
    uint64_t p = pol_last(); float v;
    asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(in + threadIdx.x), "l"(p));
    out[threadIdx.x] = v;
}
__global__ void k_st_hint(float* out) {
    uint64_t p = pol_last();
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(out + threadIdx.x), "f"(1.0f), "l"(p) : "memory");
}
__global__ void k_st_hint_v4(uint32_t* out) {
    uint64_t p = pol_last();
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(out + threadIdx.x * 4), "r"(1u), "r"(2u), "r"(3u), "r"(4u), "l"(p) : "memory");
}
__global__ void k_cpasync_hint(const uint32_t* in, uint32_t* out) {
    __shared__ __align__(16) uint32_t sm[32 * 4];
    uint64_t p = pol_first();
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm + threadIdx.x * 4);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(sa), "l"(in + threadIdx.x * 4), "l"(p) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    out[threadIdx.x] = sm[threadIdx.x * 4];
}
__global__ void k_discard(float* buf) {
    buf[threadIdx.x] = 2.0f;
    __syncwarp();
    if (threadIdx.x == 0) asm volatile("discard.global.L2 [%0], 128;" :: "l"(buf) : "memory");
}
__global__ void k_ld_evl(const float* in, float* out) {
    float v;
    asm volatile("ld.global.L1::evict_last.f32 %0, [%1];" : "=f"(v) : "l"(in + threadIdx.x));
    out[threadIdx.x] = v;
}
#define RUN(name, ...) do { name<<<1, 32>>>(__VA_ARGS__); cudaError_t e = cudaDeviceSynchronize(); printf("%-16s %s\n", #name, cudaGetErrorString(e)); if (e != cudaSuccess) return 1; } while (0)
int main(int argc, char** argv) {
    float* a; float* b; cudaMalloc(&a, 4096); cudaMalloc(&b, 4096); cudaMemset(a, 0, 4096);
    int which = argc > 1 ? atoi(argv[1]) : 0;
    if (which == 0) RUN(k_ld_hint, a, b);
    if (which == 1) RUN(k_st_hint, b);
    if (which == 2) RUN(k_st_hint_v4, (uint32_t*)b);
    if (which == 3) RUN(k_cpasync_hint, (const uint32_t*)a, (uint32_t*)b);
    if (which == 4) RUN(k_discard, b);
    if (which == 5) RUN(k_ld_evl, a, b);
    return 0;
}
