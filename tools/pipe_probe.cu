// Issue-rate probe for the instructions the ellipse-morphology kernel is made of (sm_100a).
// Every mode runs 8 independent dependency chains per thread, 8 warps per SM sub-partition, and reports
// warp-instructions per clock per SM from clock64() deltas.  Which pairs of instructions overlap (different pipes)
// and which serialise (same pipe) decides how the morphology walk is laid out.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define NCHAIN 8

__device__ __forceinline__ uint32_t vmax2(uint32_t a, uint32_t b) { uint32_t r; asm volatile("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("{.reg .b32 t; max.u16x2 t, %1, %2; max.u16x2 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t hmin2(uint32_t a, uint32_t b) { uint32_t r; asm volatile("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, 0x7351;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t lop(uint32_t a, uint32_t b) { uint32_t r; asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t lds(uint32_t addr) { uint32_t r; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr)); return r; }
__device__ __forceinline__ uint32_t hmul2(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t shl8(uint32_t a) { return a << 8; }

template <int MODE>
__global__ void __launch_bounds__(256) probe(uint32_t* out, const uint32_t* in, int iters, long long* cyc) {
    __shared__ uint32_t sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = in[i & 255];
    __syncthreads();
    uint32_t a[NCHAIN], d[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) { a[i] = in[(threadIdx.x + i) & 255]; d[i] = in[(threadIdx.x * 3 + i) & 255]; }
    const uint32_t b = in[threadIdx.x & 7], c = in[8 + (threadIdx.x & 7)];
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm) + 4 * threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) {
            const int j = (i + 1) & (NCHAIN - 1), k = (i + 3) & (NCHAIN - 1);     // operands from neighbouring chains: nothing folds
            if (MODE == 0) a[i] = vmax2(a[i], a[j]);
            if (MODE == 1) a[i] = vmax3(a[i], a[j], d[k]);
            if (MODE == 2) a[i] = hmin2(a[i], a[j]);
            if (MODE == 3) { a[i] = vmax3(a[i], a[j], c); d[i] = hmin2(d[i], d[j]); }
            if (MODE == 4) { a[i] = vmax3(a[i], a[j], c); d[i] = imad(d[i], b, d[j]); }
            if (MODE == 5) d[i] = imad(d[i], b, d[j]);
            if (MODE == 6) { a[i] = vmax3(a[i], a[j], c); d[i] = imad(lds(sbase + 1024 * i), b, d[i]); }
            if (MODE == 7) d[i] = imad(lds(sbase + 1024 * i), b, d[i]);
            if (MODE == 8) a[i] = prmt(a[i], a[j]);
            if (MODE == 9) a[i] = lop(a[i], a[j]) + 1;
            if (MODE == 10) { a[i] = vmax3(a[i], a[j], c); d[i] = prmt(d[i], d[j]); }
            if (MODE == 11) { a[i] = vmax2(a[i], a[j]); d[i] = hmin2(d[i], d[j]); }
            if (MODE == 12) { a[i] = vmax2(a[i], a[j]); d[i] = hmin2(d[i], d[j]); d[k] = hmin2(d[k], a[i]); }
            if (MODE == 13) { a[i] = vmax3(a[i], shl8(lds(sbase + 1024 * i)), lds(sbase + 1024 * i + 512)); }   // LDS x2 + SHL + VIMNMX3
            if (MODE == 14) { a[i] = vmax3(a[i], a[j], c); d[i] = hmin2(d[i], d[j]); d[k] = imad(d[k], b, c); }
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
    for (int i = 0; i < NCHAIN; ++i) r ^= a[i] ^ d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

struct Mode { const char* name; int instr_per_chain_step; };
static const Mode MODES[] = {
    {"VIMNMX.U16x2 (2 inputs)", 1}, {"VIMNMX3.U16x2", 1}, {"HMNMX2", 1}, {"VIMNMX3 + HMNMX2", 2}, {"VIMNMX3 + IMAD", 2},
    {"IMAD", 1}, {"VIMNMX3 + LDS + IMAD", 3}, {"LDS + IMAD", 2}, {"PRMT", 1}, {"LOP3 + IADD", 2}, {"VIMNMX3 + PRMT", 2},
    {"VIMNMX + HMNMX2", 2}, {"VIMNMX + 2 HMNMX2", 3}, {"2 LDS + SHL + VIMNMX3", 4}, {"VIMNMX3 + HMNMX2 + IMAD", 3}};

template <int MODE> static void run(uint32_t* out, const uint32_t* in, long long* cyc, int sms, int ctas_per_sm) {
    const int iters = 4096, grid = sms * ctas_per_sm;
    probe<MODE><<<grid, 256>>>(out, in, 64, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<grid, 256>>>(out, in, iters, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[grid];
    cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0; double mean = 0;
    for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mean += (double)h[i] / grid; }
    delete[] h;
    const double winstr_per_sm = (double)iters * NCHAIN * MODES[MODE].instr_per_chain_step * 8 * ctas_per_sm;   // 8 warps per CTA
    printf("{\"mode\": %d, \"name\": \"%s\", \"ctas_per_sm\": %d, \"warp_instr_per_clk_per_sm\": %.3f, \"cycles_mean\": %.0f, "
           "\"cycles_max\": %lld, \"ms\": %.4f, \"err\": \"%s\"}\n", MODE, MODES[MODE].name, ctas_per_sm, winstr_per_sm / mean, mean, mx, ms,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    uint32_t *out, *in; long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 8 * 256); cudaMalloc(&in, 1024); cudaMalloc(&cyc, sizeof(long long) * sms * 8);
    uint32_t h[256]; for (int i = 0; i < 256; ++i) h[i] = 0x01230045u * (i + 1) & 0x3BFF3BFFu;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    for (int c = 2; c <= 4; c += 2) {
        run<0>(out, in, cyc, sms, c); run<1>(out, in, cyc, sms, c); run<2>(out, in, cyc, sms, c); run<3>(out, in, cyc, sms, c);
        run<4>(out, in, cyc, sms, c); run<5>(out, in, cyc, sms, c); run<6>(out, in, cyc, sms, c); run<7>(out, in, cyc, sms, c);
        run<8>(out, in, cyc, sms, c); run<9>(out, in, cyc, sms, c); run<10>(out, in, cyc, sms, c); run<11>(out, in, cyc, sms, c);
        run<12>(out, in, cyc, sms, c); run<13>(out, in, cyc, sms, c); run<14>(out, in, cyc, sms, c);
    }
    return 0;
}
