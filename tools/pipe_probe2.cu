// Second issue-rate probe: which cheap integer instructions run beside VIMNMX3.U16x2 without taking its pipe?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define NCHAIN 8
__device__ __forceinline__ uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("{.reg .b32 t; max.u16x2 t, %1, %2; max.u16x2 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t shr8(uint32_t a, uint32_t b) { uint32_t r; asm volatile("shf.r.clamp.b32 %0, %1, %2, 8;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t iadd3(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("{.reg .b32 t; add.u32 t, %1, %2; add.u32 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, 0x7351;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t imadshl(uint32_t a) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, 256;" : "=r"(r) : "r"(a)); return r; }
__device__ __forceinline__ uint32_t vsub2(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
template <int MODE>
__global__ void __launch_bounds__(256) probe(uint32_t* out, const uint32_t* in, int iters, long long* cyc) {
    uint32_t a[NCHAIN], d[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) { a[i] = in[(threadIdx.x + i) & 255]; d[i] = in[(threadIdx.x * 3 + i) & 255]; }
    const uint32_t c = in[8 + (threadIdx.x & 7)];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) {
            const int j = (i + 1) & (NCHAIN - 1), k = (i + 3) & (NCHAIN - 1);
            if (MODE == 0) { a[i] = vmax3(a[i], a[j], c); d[i] = lop3(d[i], d[j], c); }
            if (MODE == 1) { a[i] = vmax3(a[i], a[j], c); d[i] = shr8(d[i], d[j]); }
            if (MODE == 2) { a[i] = vmax3(a[i], a[j], c); d[i] = iadd3(d[i], d[j], c); }
            if (MODE == 3) { a[i] = vmax3(a[i], a[j], c); d[i] = lop3(d[i], d[j], c); d[k] = shr8(d[k], d[i]); }
            if (MODE == 4) { a[i] = vmax3(a[i], a[j], c); d[i] = imadshl(d[i]) + 0; }
            if (MODE == 5) { d[i] = lop3(d[i], d[j], c); }
            if (MODE == 6) { d[i] = shr8(d[i], d[j]); }
            if (MODE == 7) { a[i] = vmax3(a[i], a[j], c); d[i] = imadshl(d[i]); d[k] = lop3(d[k], d[i], c); }
            if (MODE == 8) { a[i] = vmax3(a[i], a[j], c); d[i] = vsub2(d[i], d[j]); }
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
    for (int i = 0; i < NCHAIN; ++i) r ^= a[i] ^ d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
static const char* NAMES[] = {"VIMNMX3 + LOP3", "VIMNMX3 + SHF", "VIMNMX3 + IADD3", "VIMNMX3 + LOP3 + SHF", "VIMNMX3 + IMAD.SHL", "LOP3", "SHF", "VIMNMX3 + IMAD.SHL + LOP3", "VIMNMX3 + IADD(sub)"};
static const int NINS[] = {2, 2, 2, 3, 2, 1, 1, 3, 2};
template <int MODE> static void run(uint32_t* out, const uint32_t* in, long long* cyc, int sms) {
    const int iters = 4096, grid = sms * 2;
    probe<MODE><<<grid, 256>>>(out, in, 64, cyc);
    cudaDeviceSynchronize();
    probe<MODE><<<grid, 256>>>(out, in, iters, cyc);
    cudaDeviceSynchronize();
    long long* h = new long long[grid];
    cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    delete[] h;
    // 4 warps per sub-partition; cycles per chain step of one sub-partition's four warps (VIMNMX3 alone: 8.0)
    printf("{\"mode\": %d, \"name\": \"%s\", \"instr_per_step\": %d, \"cycles_per_step_per_smsp\": %.3f, \"err\": \"%s\"}\n", MODE, NAMES[MODE], NINS[MODE],
           (double)mx / ((double)iters * NCHAIN), cudaGetErrorString(cudaGetLastError()));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    uint32_t *out, *in; long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 2 * 256); cudaMalloc(&in, 1024); cudaMalloc(&cyc, sizeof(long long) * sms * 2);
    uint32_t h[256]; for (int i = 0; i < 256; ++i) h[i] = 0x01230045u * (i + 1);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>(out, in, cyc, sms); run<1>(out, in, cyc, sms); run<2>(out, in, cyc, sms); run<3>(out, in, cyc, sms); run<4>(out, in, cyc, sms);
    run<5>(out, in, cyc, sms); run<6>(out, in, cyc, sms); run<7>(out, in, cyc, sms); run<8>(out, in, cyc, sms);
    return 0;
}
