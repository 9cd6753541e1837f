// Development micro-benchmark: issue rate of the integer instructions the front-end kernels lean on (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 512
template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed * (threadIdx.x + i + 1);
    uint32_t b = seed ^ 0x01020304u, c = seed + 7;
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) a[i] = a[i] * b + c;                                   // IMAD
            if (OP == 1) a[i] = __vimin3_s16x2(a[i], b, c);                       // VIMNMX3.S16x2
            if (OP == 2) a[i] = __byte_perm(a[i], b, 0x5140);                     // PRMT
            if (OP == 3) a[i] = __funnelshift_r(a[i], b, 8);                      // SHF
            if (OP == 4) a[i] = __dp4a(a[i], b, c);                               // IDP.4A
            if (OP == 5) a[i] = __dp2a_lo(a[i], b, c);                            // IDP.2A
            if (OP == 6) a[i] = (a[i] & b) ^ c;                                   // LOP3
            if (OP == 7) a[i] = a[i] + b + c;                                     // IADD3
            if (OP == 8) a[i] = __vmins2(a[i], b);                                // VIMNMX.S16x2
            if (OP == 9) a[i] = min(a[i], b);                                     // VIMNMX.U32
            if (OP == 10) a[i] = __vabsdiffu4(a[i], b);                           // VABSDIFF4
            if (OP == 11) a[i] = __popc(a[i]) + b;                                // POPC
            if (OP == 12) { a[i] = a[i] * b + c; a[i] = __vimin3_s16x2(a[i], b, c); }   // IMAD + VIMNMX3 pair (dual pipe?)
            if (OP == 13) { a[i] = __dp4a(a[i], b, c); a[i] = __vimin3_s16x2(a[i], b, c); }   // IDP + VIMNMX3
            if (OP == 14) { a[i] = __dp4a(a[i], b, c); a[i] = a[i] * b + c; }      // IDP + IMAD
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i];
    if (s == 0x12345678u) out[threadIdx.x] = s;
}
template <int OP> void run(const char* name, int per) {
    uint32_t* d; cudaMalloc(&d, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8;
    k<OP><<<blocks, 256>>>(d, 3); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<OP><<<blocks, 256>>>(d, 5); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double winst = (double)blocks * 8 * ITER * 8 * per;   // warp instructions
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %.3f ms  %.2f warp-inst/clk/SM (at %d MHz nominal)\n", name, ms, winst / cycles / 148, clk / 1000);
    cudaFree(d);
}
int main() {
    run<0>("IMAD", 1); run<1>("VIMNMX3.S16x2", 1); run<2>("PRMT", 1); run<3>("SHF", 1); run<4>("IDP.4A", 1); run<5>("IDP.2A", 1);
    run<6>("LOP3", 1); run<7>("IADD3", 1); run<8>("VIMNMX.S16x2", 1); run<9>("VIMNMX.U32", 1); run<10>("VABSDIFF4", 1); run<11>("POPC+IADD", 2);
    run<12>("IMAD+VIMNMX3", 2); run<13>("IDP4A+VIMNMX3", 2); run<14>("IDP4A+IMAD", 2);
    return 0;
}
