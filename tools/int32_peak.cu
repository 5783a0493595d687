// int32_peak.cu — micro-benchmark of the B200 integer issue rate (not part of the product).
// MEASURED_PEAKS.json has no INT32 figure, so the ALU-bound roofline of
// the (64,5) shape needs one: dependency-free streams of IADD3/LOP3 (ALU pipe), IMAD (FMA pipe)
// and a 1:1 mix, 8 independent accumulators per thread, full occupancy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/int32_peak tools/int32_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(int* out, int iters, int a, int b) {
    int r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE == 0) {  // ALU pipe: add + logic
                r0 = (r0 + a) ^ b; r1 = (r1 + a) ^ b; r2 = (r2 + a) ^ b; r3 = (r3 + a) ^ b;
                r4 = (r4 + a) ^ b; r5 = (r5 + a) ^ b; r6 = (r6 + a) ^ b; r7 = (r7 + a) ^ b;
            } else if (MODE == 1) {  // FMA pipe: IMAD
                r0 = r0 * a + b; r1 = r1 * a + b; r2 = r2 * a + b; r3 = r3 * a + b;
                r4 = r4 * a + b; r5 = r5 * a + b; r6 = r6 * a + b; r7 = r7 * a + b;
            } else {  // mix: one IMAD + one LOP3 per accumulator
                r0 = (r0 * a + b) ^ a; r1 = (r1 * a + b) ^ a; r2 = (r2 * a + b) ^ a; r3 = (r3 * a + b) ^ a;
                r4 = (r4 * a + b) ^ a; r5 = (r5 * a + b) ^ a; r6 = (r6 * a + b) ^ a; r7 = (r7 * a + b) ^ a;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7;
}

template <int MODE>
double run(const char* name, int ops_per_acc) {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * 8, iters = 4096; int* out; cudaMalloc(&out, (size_t)blocks * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 64, 3, 5); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, iters, 3, 5); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double ops = (double)blocks * 256 * iters * 16 * 8 * ops_per_acc;
    double rate = ops / (best * 1e-3);
    printf("  \"%s\": %.4e,\n", name, rate);
    cudaFree(out);
    return rate;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"unit\": \"int32 lane-ops/s\",\n", p.name, p.multiProcessorCount);
    double a = run<0>("alu_pipe_add_xor", 2);
    double f = run<1>("fma_pipe_imad", 1);
    double m = run<2>("mix_imad_lop3", 2);
    printf("  \"int32_peak_ops\": %.4e,\n  \"how\": \"best of 5, 8 independent accumulators x 16x unroll, 8 CTAs of 256 threads per SM; peak = max(alu, mix)\"\n}\n", a > m ? a : m);
    (void)f;
    return 0;
}
