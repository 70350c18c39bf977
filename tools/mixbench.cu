// Instruction-mix ceiling benchmark: how fast can one B200 SM retire the INSTRUCTION MIX of a 256-bit fixed-operand
// multiply, for (A) the mix the shipped multiplier compiles to and (B, C) the mixes a multiplier that moves its
// high-half product (B) or all three products (C) to 52-bit-limb FP64 arithmetic would have?  No arithmetic meaning:
// the kernels issue the same NUMBER of IMAD.WIDE / IMAD / DFMA / ALU instructions per "multiply" in independent
// dependency chains, so the result is an upper bound for a real multiplier with that mix.  Mix A must land near the
// measured rate of the real multiplier (88 G/s) for the method to mean anything.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mixbench tools/mixbench.cu
#include <cstdint>
#include <cstdio>

template <int NW, int NI, int ND, int NA>
__global__ void __launch_bounds__(256) mix(uint32_t* out, int iters, uint32_t seed) {
    uint32_t x[8], y[8], z[8];
    uint32_t wl[8], wh[8];  // even-aligned pairs: ptxas fuses each mad.lo.cc / madc.hi pair into one IMAD.WIDE.U32(.X)
    double d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        x[i] = threadIdx.x * (2 * i + 3) + seed;
        y[i] = x[i] ^ 0x9e3779b9u;
        wl[i] = x[i];
        wh[i] = x[i] * 7 + 1;
        z[i] = x[i] + 17;
        d[i] = 1.0 + x[i] * 1e-9;
    }
    const double e = 1.0000001, f = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < (NW + 7) / 8; k++)
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (k * 8 + i < NW)  // the multiplier operand depends on the chain: nothing to strength-reduce
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                                 : "+r"(wl[i]), "+r"(wh[i]) : "r"(wh[(i + 1) & 7]), "r"(y[(i + 1) & 7]));
#pragma unroll
        for (int k = 0; k < (NI + 7) / 8; k++)
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (k * 8 + i < NI) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(y[(i + 3) & 7]));
#pragma unroll
        for (int k = 0; k < (ND + 7) / 8; k++)
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (k * 8 + i < ND) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e), "d"(f));
#pragma unroll
        for (int k = 0; k < (NA + 7) / 8; k++)
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (k * 8 + i < NA) {  // carry-chain adds (IADD3 / IADD3.X cannot move to the multiplier pipe) and logic ops
                    if ((k & 1) == 0 && k * 8 + i + 8 < NA)
                        asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(y[i]), "+r"(z[i]) : "r"(x[(i + 5) & 7]), "r"(x[(i + 6) & 7]));
                    else if ((k & 1) == 0)
                        asm volatile("xor.b32 %0, %0, %1;" : "+r"(y[i]) : "r"(x[(i + 2) & 7]));
                }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc += x[i] + y[i] + z[i] + wl[i] + wh[i] + (uint32_t)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int NW, int NI, int ND, int NA>
static void run(const char* name, uint32_t* out, int sms, int bps) {
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    mix<NW, NI, ND, NA><<<sms * bps, 256>>>(out, iters, 1);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        mix<NW, NI, ND, NA><<<sms * bps, 256>>>(out, iters, 1);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double muls = (double)sms * bps * 256 * iters;
    printf("{\"mix\": \"%s\", \"imad_wide\": %d, \"imad\": %d, \"fp64\": %d, \"alu\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, "
           "\"g_mix_per_s\": %.2f, \"clk_per_mix_per_sm\": %.2f}\n",
           name, NW, NI, ND, NA, bps, best, muls / best / 1e6, 1.965e6 * best * sms / muls);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    uint32_t* out;
    cudaMalloc(&out, 4 * sms * 8 * 256);
    for (int bps : {2, 4, 8}) {
        // A: SASS histogram of the shipped mul_pre loop body (BLS12-381): 78 IMAD.WIDE, 28 IMAD, 67 ALU
        run<78, 28, 0, 67>("A shipped mul_pre", out, sms, bps);
        // B: high-half product on FP64 (15 x 2 DFMA + 4 DFMA + 15 DADD + 5 conversions), low halves stay on IMAD;
        //    ALU: +68 (64-bit integer accumulation of 34 mantissas) +35 (limb conversions) on top of the 67
        run<35, 28, 54, 170>("B hi-half on FP64", out, sms, bps);
        // C: all three products on FP64 (hi 34 + lo 2 x ~34 DFMA + DADDs), accumulation and carries on the ALU
        run<0, 15, 150, 260>("C all products on FP64", out, sms, bps);
        // reference points: pure pipes
        run<64, 0, 0, 0>("64 IMAD.WIDE", out, sms, bps);
        run<0, 64, 0, 0>("64 IMAD", out, sms, bps);
        run<0, 0, 64, 0>("64 DFMA", out, sms, bps);
        run<0, 0, 0, 64>("64 ALU", out, sms, bps);
    }
    return 0;
}
