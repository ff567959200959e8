// Dependent-chain latencies (SM cycles per op, one warp) of the operations the ICP loop's reduction is built from.  nvcc -arch=sm_100a -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
template <int OP> __global__ void k(double* out, long long* cyc, double a, double b, int lanes)
{
    double x = a + threadIdx.x; long long q = (long long)threadIdx.x + 3; float f = (float)a + threadIdx.x; unsigned u = threadIdx.x;
    __shared__ double sm[64]; sm[threadIdx.x & 63] = a; __syncthreads();
    const long long t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = __dadd_rn(x, b);
        if (OP == 1) x = __fma_rn(x, b, a);
        if (OP == 2) x = __dmul_rn(x, b);
        if (OP == 3) { q = __double2ll_rn(x); x = (double)q * b; }                 // F2I + I2F + DMUL
        if (OP == 4) q = q + (q >> 3) + 1;                                          // 64-bit integer adds
        if (OP == 5) f = __fmaf_rn(f, 1.0001f, 0.5f);
        if (OP == 6) x = __shfl_xor_sync(0xffffffffu, x, 1) + b;                   // 64-bit shuffle + DADD
        if (OP == 7) q = __shfl_xor_sync(0xffffffffu, q, 1) + 1;                   // 64-bit shuffle + IADD
        if (OP == 8) x = floor(x) * b;
        if (OP == 9) { u = atomicAdd((unsigned*)&sm[0] + (u & 1), 1u); }            // shared atomic, dependent
        if (OP == 10) x = sm[(int)x & 63] + b;                                      // LDS + DADD dependent
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x + (double)q + f + u;
}
int main()
{
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
    const char* names[] = {"DADD", "DFMA", "DMUL", "F2I.S64.F64 + I2F.F64.S64 + DMUL", "IADD64 x2 + SHR", "FFMA", "SHFL64 + DADD", "SHFL64 + IADD64", "floor + DMUL", "ATOMS.ADD u32 (dependent)", "LDS.64 + DADD (dependent address)"};
    for (int threads : {32, 512}) {
        printf("%d threads per block, 1 block\n", threads);
        for (int op = 0; op <= 10; op++) {
            long long c = 0;
            for (int rep = 0; rep < 2; rep++) {
                switch (op) {
                    case 0: k<0><<<1, threads>>>(out, cyc, 1.0, 1e-9, 32); break; case 1: k<1><<<1, threads>>>(out, cyc, 1.0, 0.999, 32); break;
                    case 2: k<2><<<1, threads>>>(out, cyc, 1.0, 0.9999, 32); break; case 3: k<3><<<1, threads>>>(out, cyc, 1000.0, 1.0001, 32); break;
                    case 4: k<4><<<1, threads>>>(out, cyc, 1.0, 1.0, 32); break; case 5: k<5><<<1, threads>>>(out, cyc, 1.0, 1.0, 32); break;
                    case 6: k<6><<<1, threads>>>(out, cyc, 1.0, 1e-9, 32); break; case 7: k<7><<<1, threads>>>(out, cyc, 1.0, 1.0, 32); break;
                    case 8: k<8><<<1, threads>>>(out, cyc, 1000.5, 1.0001, 32); break; case 9: k<9><<<1, threads>>>(out, cyc, 1.0, 1.0, 32); break;
                    case 10: k<10><<<1, threads>>>(out, cyc, 1.0, 1.0, 32); break;
                }
                cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            }
            printf("  %-40s %6.1f cycles per iteration\n", names[op], (double)c / N);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
