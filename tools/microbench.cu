// Micro-benchmarks that set the expectations for every kernel:
// 256-bit Montgomery multiplications per second and Blake2s compressions per second on one B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o build/microbench tools/microbench.cu
#include <cstdio>
#include <cstring>
#include <vector>
#include "../hodor_b200/csrc/field.cuh"
#include "../hodor_b200/csrc/merkle.cuh"
using namespace hodor;

// REGS: modulus in vector registers (IMAD.WIDE everywhere) vs immediates; SPLIT: mont_split.cuh
template <class F, int ILP, bool REGS, bool SPLIT = false>
__global__ void __launch_bounds__(256) mul_kernel(const Fe* in, Fe* out, int iters, uint32_t zero) {
    const Field<F> fld(REGS ? (threadIdx.x & zero) : 0u);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    Fe x[ILP];
    Fe y = in[(tid + 1) & 1023];
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = in[(tid + 7 * j) & 1023];
    for (int k = 0; k < iters; k++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = SPLIT ? fld.mul_split(x[j], y) : fld.mul_evenodd(x[j], y);
    }
    Fe acc = x[0];
#pragma unroll
    for (int j = 1; j < ILP; j++) acc = fld.add(acc, x[j]);
    out[tid] = acc;
}

// fixed-operand multiplier (Field::mul_pre): x <- x * w with (w, wq) precomputed
template <class F, int ILP, uint32_t GUARD>
__global__ void __launch_bounds__(256) mul_pre_kernel(const Fe* in, Fe* out, int iters, uint32_t zero) {
    const Field<F> fld(0u);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    Fe x[ILP];
    Fe w, wq;
    fld.make_pre(in[(tid + 1) & 1023], w, wq);
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = in[(tid + 7 * j) & 1023];
    for (int k = 0; k < iters; k++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = fld.template mul_pre<GUARD>(x[j], w, wq);
    }
    Fe acc = x[0];
#pragma unroll
    for (int j = 1; j < ILP; j++) acc = fld.add(acc, x[j]);
    out[tid] = acc;
}
// same chain through the Montgomery multiplier: must give identical bits
template <class F>
__global__ void __launch_bounds__(256) mul_ref_kernel(const Fe* in, Fe* out, int iters, uint32_t zero) {
    const Field<F> fld(0u);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    Fe x = in[tid & 1023];
    const Fe y = in[(tid + 1) & 1023];
    for (int k = 0; k < iters; k++) x = fld.mul_evenodd(x, y);
    out[tid] = x;
}

template <class F>
__global__ void __launch_bounds__(256) bfly_kernel(const Fe* in, Fe* out, int iters, uint32_t zero) {
    const Field<F> fld(threadIdx.x & zero);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    Fe x = in[tid & 1023], y = in[(tid + 1) & 1023], w = in[(tid + 2) & 1023];
    for (int k = 0; k < iters; k++) {
        const Fe s = fld.add(x, y), d = fld.sub(x, y);
        x = s;
        y = fld.mul(d, w);
    }
    out[tid] = fld.add(x, y);
}

__global__ void __launch_bounds__(256) b2s_kernel(const uint4* in, uint4* out, int iters, const __grid_constant__ B2sState key) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    Digest a = ld_digest(in, tid & 1023), b = ld_digest(in, (tid + 1) & 1023);
    for (int k = 0; k < iters; k++) a = hash_node64(key, a, b);
    st_digest(out, tid, a);
}
__global__ void __launch_bounds__(256) b2s_leaf_kernel(const uint4* in, uint4* out, int iters, const __grid_constant__ B2sState key) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    Digest a = ld_digest(in, tid & 1023);
    for (int k = 0; k < iters; k++) a = hash_leaf32(key, a.w);
    st_digest(out, tid, a);
}

template <class K, class... A>
static double time_ms(K kern, dim3 grid, dim3 block, A... args) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<<<grid, block>>>(args...);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        kern<<<grid, block>>>(args...);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return best;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    const int sms = prop.multiProcessorCount;
    std::vector<Fe> h(1024);
    for (int i = 0; i < 1024; i++)
        for (int j = 0; j < 8; j++) h[i].v[j] = (uint32_t)(0x9E3779B9u * (i * 8 + j + 1)) & (j == 7 ? 0x3fffffffu : 0xffffffffu);
    Fe *d_in, *d_out;
    cudaMalloc(&d_in, sizeof(Fe) * 1024);
    cudaMalloc(&d_out, sizeof(Fe) * sms * 16 * 256);
    cudaMemcpy(d_in, h.data(), sizeof(Fe) * 1024, cudaMemcpyHostToDevice);
    const int iters = 2000;
    {   // parity of mul_pre (normal guard, and guard forced low so the out-of-line fix-up runs half the time)
        const int nthr = sms * 256;
        std::vector<Fe> r0(nthr), r1(nthr), r2(nthr);
        mul_ref_kernel<BlsFr><<<sms, 256>>>(d_in, d_out, 300, 0u);
        cudaMemcpy(r0.data(), d_out, sizeof(Fe) * nthr, cudaMemcpyDeviceToHost);
        mul_pre_kernel<BlsFr, 1, 0xfffffff2u><<<sms, 256>>>(d_in, d_out, 300, 0u);
        cudaMemcpy(r1.data(), d_out, sizeof(Fe) * nthr, cudaMemcpyDeviceToHost);
        mul_pre_kernel<BlsFr, 1, 0x80000000u><<<sms, 256>>>(d_in, d_out, 300, 0u);
        cudaMemcpy(r2.data(), d_out, sizeof(Fe) * nthr, cudaMemcpyDeviceToHost);
        int bad1 = 0, bad2 = 0;
        for (int i = 0; i < nthr; i++) {
            bad1 += memcmp(&r0[i], &r1[i], sizeof(Fe)) != 0;
            bad2 += memcmp(&r0[i], &r2[i], sizeof(Fe)) != 0;
        }
        printf("{\"check\": \"mul_pre_vs_mont\", \"threads\": %d, \"chain\": 300, \"mismatch\": %d, \"mismatch_forced_fixup\": %d, \"err\": \"%s\"}\n",
               nthr, bad1, bad2, cudaGetErrorString(cudaGetLastError()));
    }
    for (int bps : {1, 2, 4, 8}) {
        dim3 grid(sms * bps), block(256);
        const double threads = (double)sms * bps * 256;
#define RUN_MUL(F, ILP, REGS, NAME)                                                                          \
    {                                                                                                        \
        double ms = time_ms(mul_kernel<F, ILP, REGS>, grid, block, (const Fe*)d_in, d_out, iters, 0u);      \
        printf("{\"bench\": \"%s\", \"ilp\": %d, \"modulus_in_regs\": %d, \"blocks_per_sm\": %d, \"gmul_per_s\": %.2f}\n", NAME, \
               ILP, (int)REGS, bps, threads * iters * ILP / ms / 1e6);                                      \
    }
        RUN_MUL(BlsFr, 1, false, "mont_mul_bls")
#define RUN_PRE(F, ILP, NAME)                                                                                 \
    {                                                                                                        \
        double ms = time_ms(mul_pre_kernel<F, ILP, 0xfffffff2u>, grid, block, (const Fe*)d_in, d_out, iters, 0u); \
        printf("{\"bench\": \"%s\", \"ilp\": %d, \"blocks_per_sm\": %d, \"gmul_per_s\": %.2f}\n", NAME, ILP, bps, \
               threads * iters * ILP / ms / 1e6);                                                            \
    }
        RUN_PRE(BlsFr, 1, "mul_pre_bls")
        RUN_PRE(BlsFr, 2, "mul_pre_bls")
        RUN_PRE(Bn254Fr, 1, "mul_pre_bn254")
        RUN_PRE(Stark252, 1, "mul_pre_stark")
        {
            double ms = time_ms(mul_kernel<BlsFr, 1, false, true>, grid, block, (const Fe*)d_in, d_out, iters, 0u);
            printf("{\"bench\": \"mont_mul_bls_split\", \"blocks_per_sm\": %d, \"gmul_per_s\": %.2f}\n", bps, threads * iters / ms / 1e6);
            ms = time_ms(mul_kernel<Bn254Fr, 1, false, true>, grid, block, (const Fe*)d_in, d_out, iters, 0u);
            printf("{\"bench\": \"mont_mul_bn254_split\", \"blocks_per_sm\": %d, \"gmul_per_s\": %.2f}\n", bps, threads * iters / ms / 1e6);
        }
        RUN_MUL(BlsFr, 1, true, "mont_mul_bls")
        RUN_MUL(BlsFr, 2, true, "mont_mul_bls")
        RUN_MUL(BlsFr, 4, true, "mont_mul_bls")
        RUN_MUL(Bn254Fr, 2, true, "mont_mul_bn254")
        RUN_MUL(Bn254Fr, 2, false, "mont_mul_bn254")
        {
            double ms = time_ms(bfly_kernel<BlsFr>, grid, block, (const Fe*)d_in, d_out, iters, 0u);
            printf("{\"bench\": \"butterfly_bls\", \"blocks_per_sm\": %d, \"gbfly_per_s\": %.2f}\n", bps, threads * iters / ms / 1e6);
        }
        {
            B2sState key = b2s_keyed_state();
            double ms = time_ms(b2s_kernel, grid, block, (const uint4*)d_in, (uint4*)d_out, 500, key);
            printf("{\"bench\": \"blake2s_node\", \"blocks_per_sm\": %d, \"gcompress_per_s\": %.2f}\n", bps, threads * 500 / ms / 1e6);
            ms = time_ms(b2s_leaf_kernel, grid, block, (const uint4*)d_in, (uint4*)d_out, 500, key);
            printf("{\"bench\": \"blake2s_leaf\", \"blocks_per_sm\": %d, \"gcompress_per_s\": %.2f}\n", bps, threads * 500 / ms / 1e6);
        }
    }
    return 0;
}
