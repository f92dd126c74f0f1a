// Integer pipe micro-benchmark for sm_100a: warp-instruction throughput per SM sub-partition of the ops the
// Myers sweep is made of, alone and mixed.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
#define UNROLL 8

template <int OP>
__global__ void k(uint32_t* out, uint32_t two, uint32_t seed) {
    uint32_t a[UNROLL], b = seed + threadIdx.x, c = seed ^ 0x9e3779b9u;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) a[i] = seed * (i + 1) + threadIdx.x;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 2) asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(a[i]) : "r"(b));
            if (OP == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(two), "r"(b));
            if (OP == 4) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(two));
            if (OP == 5) { uint64_t r; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a[i]), "r"(two)); a[i] = (uint32_t)r ^ (uint32_t)(r >> 32); }
            if (OP == 6) {  // mixed: LOP3 + IMAD alternating (independent chains)
                if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
                else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(two), "r"(b));
            }
            if (OP == 7) {  // mixed: LOP3 + IMAD.HI
                if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
                else asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(two));
            }
            if (OP == 8) asm volatile("prmt.b32 %0, %0, %1, 0x4441;" : "+r"(a[i]) : "r"(b));
            if (OP == 9) asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
            if (OP == 10) {  // 3 LOP3 : 1 IMAD
                if ((i & 3) == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(two), "r"(b));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            }
            if (OP == 11) asm volatile("min.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, uint32_t* d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 4, threads = 512;  // 16 warps per SMSP... 4 blocks x 16 warps = 64 warps/SM
    k<OP><<<blocks, threads>>>(d, 2u, 12345u);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(d, 2u, 12345u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double warpInstr = (double)blocks * (threads / 32) * ITER * UNROLL;
    double smspCycles = ms * 1e-3 * clk * 1e3 * 148 * 4;
    printf("%-28s %8.3f ms  %.3f warp-instr/cycle/SMSP (clock %d kHz)\n", name, ms, warpInstr / smspCycles, clk);
}

int main() {
    uint32_t* d; cudaMalloc(&d, 148 * 4 * 512 * 4);
    run<0>("LOP3", d); run<1>("IADD", d); run<2>("SHF.L.W", d); run<3>("IMAD (mad.lo, UR mult)", d);
    run<4>("IMAD.HI", d); run<5>("IMAD.WIDE+xor", d); run<6>("LOP3+IMAD 1:1", d); run<7>("LOP3+IMAD.HI 1:1", d);
    run<8>("PRMT", d); run<9>("POPC", d); run<10>("LOP3+IMAD 3:1", d); run<11>("IMNMX", d);
    return 0;
}
