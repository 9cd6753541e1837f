// Development micro-benchmark: issue rate of the integer instructions the front-end kernels lean on (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#define ITER 4096
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
            if (OP == 15) { __half2 h = __hmin2(*reinterpret_cast<__half2*>(&a[i]), *reinterpret_cast<__half2*>(&b)); a[i] = *reinterpret_cast<uint32_t*>(&h); }   // HMNMX2
            if (OP == 16) { __half2 h = __hmin2(*reinterpret_cast<__half2*>(&a[i]), *reinterpret_cast<__half2*>(&b)); a[i] = __vimin3_s16x2(*reinterpret_cast<uint32_t*>(&h), b, c); }
            if (OP == 17) a[i] = __float_as_uint(fminf(__uint_as_float(a[i]), __uint_as_float(b)));   // FMNMX
            if (OP == 18) a[i] = __umulhi(a[i], b) + c;                            // IMAD.HI
            if (OP == 19) { __half2 h = __hfma2(*reinterpret_cast<__half2*>(&a[i]), *reinterpret_cast<__half2*>(&b), *reinterpret_cast<__half2*>(&c)); a[i] = *reinterpret_cast<uint32_t*>(&h); }   // HFMA2
            if (OP == 20) { __half2 h = __hfma2(*reinterpret_cast<__half2*>(&a[i]), *reinterpret_cast<__half2*>(&b), *reinterpret_cast<__half2*>(&c)); a[i] = __vimin3_s16x2(*reinterpret_cast<uint32_t*>(&h), b, c); }
            if (OP == 21) { a[i] = __viaddmin_s16x2(a[i], b, c); }                 // VIADDMNMX.S16x2
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
    run<15>("HMNMX2", 1); run<16>("HMNMX2+VIMNMX3", 2); run<17>("FMNMX", 1); run<18>("IMAD.HI+IADD", 2); run<19>("HFMA2", 1); run<20>("HFMA2+VIMNMX3", 2); run<21>("VIADDMNMX.S16x2", 1);
    return 0;
}
