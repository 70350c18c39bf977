// Pipe-rate micro-benchmark: how many IMAD, IMAD.WIDE, DFMA and IADD3 a B200 SM issues per clock,
// alone and mixed.  Decides whether a 52-bit-limb FP64 multiplier can run beside the INT32 one.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/pipebench tools/pipebench.cu
#include <cstdio>
#include <cstdint>

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a0 = threadIdx.x + seed, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3;
    uint32_t b0 = a0 ^ 0x9e3779b9u, b1 = a1 ^ 0x85ebca6bu, b2 = a2 ^ 0xc2b2ae35u, b3 = a3 ^ 0x27d4eb2fu;
    unsigned long long w0 = a0, w1 = a1, w2 = a2, w3 = a3;
    double d0 = a0 * 1e-3, d1 = a1 * 1e-3, d2 = a2 * 1e-3, d3 = a3 * 1e-3, e = 1.0000001, f = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE == 0 || MODE == 4 || MODE == 5) {  // IMAD (32-bit lo)
                a0 = a0 * b0 + b1; a1 = a1 * b1 + b2; a2 = a2 * b2 + b3; a3 = a3 * b3 + b0;
            }
            if (MODE == 1) {  // IMAD.WIDE
                w0 = (unsigned long long)(uint32_t)w0 * b0 + w1; w1 = (unsigned long long)(uint32_t)w1 * b1 + w2;
                w2 = (unsigned long long)(uint32_t)w2 * b2 + w3; w3 = (unsigned long long)(uint32_t)w3 * b3 + w0;
            }
            if (MODE == 2 || MODE == 4) {  // DFMA
                d0 = fma(d0, e, f); d1 = fma(d1, e, f); d2 = fma(d2, e, f); d3 = fma(d3, e, f);
            }
            if (MODE == 3 || MODE == 5) {  // LOP3/IADD3 on the ALU pipe
                b0 = (b0 ^ a1) + 0x1234567u; b1 = (b1 ^ a2) + 0x2345678u; b2 = (b2 ^ a3) + 0x3456789u; b3 = (b3 ^ a0) + 0x456789au;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + b0 + b1 + b2 + b3 + (uint32_t)(w0 + w1 + w2 + w3) +
                                                 (uint32_t)(d0 + d1 + d2 + d3);
}

template <int MODE>
static void run(const char* name, int ops_per_inner, uint32_t* out, int sms) {
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<sms * 4, 256>>>(out, iters, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<sms * 4, 256>>>(out, iters, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)sms * 4 * 256 * iters * 16 * ops_per_inner;
    printf("{\"pipe\": \"%s\", \"ms\": %.3f, \"gops_per_s\": %.1f, \"per_sm_per_clk_at_1965MHz\": %.1f}\n", name, ms, ops / ms / 1e6,
           ops / ms / 1e6 / sms / 1.965);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    uint32_t* out;
    cudaMalloc(&out, 4 * prop.multiProcessorCount * 4 * 256);
    run<0>("imad_lo", 4, out, prop.multiProcessorCount);
    run<1>("imad_wide", 4, out, prop.multiProcessorCount);
    run<2>("dfma", 4, out, prop.multiProcessorCount);
    run<3>("alu_lop_iadd(2 per op)", 4, out, prop.multiProcessorCount);
    run<4>("imad_lo+dfma (8 per inner)", 8, out, prop.multiProcessorCount);
    run<5>("imad_lo+alu (8 per inner)", 8, out, prop.multiProcessorCount);
    return 0;
}
