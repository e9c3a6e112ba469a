// pipe_probe.cu -- issue rate of the integer instructions the threshold kernel is built from (B200, sm_100a).
// Each kernel runs N_IT iterations of 8 independent chains of one instruction per thread; reports cycles per
// warp-instruction per SM sub-partition (4 SMSPs per SM), with 16 warps per SMSP worth of CTAs resident.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N_IT 4096

template <int OP>
__global__ void __launch_bounds__(256) k_probe(uint32_t *out, uint32_t a0, uint32_t b0, long long *cyc)
{
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = a0 + threadIdx.x * (i + 1);
    uint32_t b = b0 + threadIdx.x, c = b0 * 3 + 1;
    __shared__ uint32_t sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = i * b0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));                         // IADD3
            if (OP == 1) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(b), "r"(c));          // IMAD
            if (OP == 2) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(b), "r"(c));        // IDP.4A
            if (OP == 3) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(b), "r"(c));     // IDP.2A
            if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));            // PRMT
            if (OP == 5) asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(r[i]) : "r"(b));               // SHF
            if (OP == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(b), "r"(c));      // LOP3
            if (OP == 7) {                                                                                   // alternate IADD3 / IMAD
                if (i & 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
                else asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(b), "r"(c));
            }
            if (OP == 8) {                                                                                   // alternate IADD3 / IDP.2A
                if (i & 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
                else asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(b), "r"(c));
            }
            if (OP == 9) {                                                                                   // alternate IMAD / IDP.2A
                if (i & 1) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(b), "r"(c));
                else asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(b), "r"(c));
            }
            if (OP == 10) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(&sm[(threadIdx.x + i * 32) & 2047]))); r[i] ^= v; }          // LDS.32 + LOP
            if (OP == 11) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(&sm[((threadIdx.x + i * 32) * 4) & 2047]))); r[i] ^= v.x ^ v.w; }   // LDS.128 + LOP3
            if (OP == 12) r[i] ^= __ballot_sync(0xffffffffu, (int)r[i] < 0);
            if (OP == 13) asm volatile("{.reg .pred p; setp.lt.s32 p, %0, 0; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "+r"(r[i]));          // ISETP + VOTE
            if (OP == 14) asm volatile("mad.lo.u32 %0, %1, 1, %0;" : "+r"(r[i]) : "r"(b));                  // IMAD by immediate 1 (does ptxas keep it on the fma pipe?)
            if (OP == 15) asm volatile("{.reg .f32 f; add.f32 f, %1, %2; mov.b32 %0, f;}" : "=r"(r[i]) : "f"(__uint_as_float(r[i])), "f"(__uint_as_float(b)));  // FADD
            if (OP == 16) asm volatile("bfe.u32 %0, %0, 8, 8;" : "+r"(r[i]));                                // BFE (SGXT/LOP/SHF?)
            if (OP == 17) asm volatile("vsub4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));  // video SAD-like
            if (OP == 18) asm volatile("{.reg .pred p; setp.lt.s32 p, %0, 0; @p or.b32 %0, %0, 16;}" : "+r"(r[i]));   // ISETP + predicated LOP
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= r[i];
    out[blockIdx.x * 256 + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(const char *name, uint32_t *d_out, long long *d_cyc, int sms)
{
    const int ctas = sms * 4;                     // 4 CTAs x 8 warps = 32 warps per SM = 8 per SMSP
    k_probe<OP><<<ctas, 256>>>(d_out, 3, 5, d_cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_probe<OP><<<ctas, 256>>>(d_out, 3, 5, d_cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[8];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    // warp-instructions per SMSP = 8 warps x N_IT x 8
    const double wi = 8.0 * N_IT * 8;
    printf("%-28s %8.3f ms   %6.2f cycles per warp-instruction per SMSP (clock64 of CTA 0: %lld)  err=%s\n", name, ms, (double)h[0] / wi, h[0],
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *d_out; long long *d_cyc;
    cudaMalloc(&d_out, (size_t)sms * 4 * 256 * 4);
    cudaMalloc(&d_cyc, (size_t)sms * 4 * 8);
    printf("SMs %d\n", sms);
    run<0>("IADD3", d_out, d_cyc, sms);
    run<1>("IMAD", d_out, d_cyc, sms);
    run<14>("IMAD x1 imm", d_out, d_cyc, sms);
    run<2>("IDP.4A", d_out, d_cyc, sms);
    run<3>("IDP.2A", d_out, d_cyc, sms);
    run<4>("PRMT", d_out, d_cyc, sms);
    run<5>("SHF", d_out, d_cyc, sms);
    run<6>("LOP3", d_out, d_cyc, sms);
    run<16>("BFE", d_out, d_cyc, sms);
    run<15>("FADD", d_out, d_cyc, sms);
    run<7>("IADD3 / IMAD alternating", d_out, d_cyc, sms);
    run<8>("IADD3 / IDP.2A alternating", d_out, d_cyc, sms);
    run<9>("IMAD / IDP.2A alternating", d_out, d_cyc, sms);
    run<10>("LDS.32 + LOP", d_out, d_cyc, sms);
    run<11>("LDS.128 + LOP3", d_out, d_cyc, sms);
    run<12>("VOTE", d_out, d_cyc, sms);
    run<13>("ISETP + VOTE", d_out, d_cyc, sms);
    run<18>("ISETP + @p LOP", d_out, d_cyc, sms);
    run<17>("vsub4.add", d_out, d_cyc, sms);
    return 0;
}
