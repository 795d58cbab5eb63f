// int_peak.cu -- integer-ALU issue-rate microbenchmark for B200 (sm_100a).
//
// MEASURED_PEAKS.json has HBM and bf16 figures but no integer figure; K1 (mean shift) is bound by the
// INT32 ALU, so this tool measures the sustained per-SM issue rates of the instructions K1 is made of
// (IADD3, IMAD, IDP4A, VABSDIFF4, LOP3, PRMT, ISETP + predicated add, LDS) and of the K1 inner-loop mix.
// Output: one JSON object on stdout (bench.py stores it next to the roofline it feeds).
//
// Method: 8 independent dependency chains per thread, fully unrolled bodies of inline PTX (checked with
// cuobjdump -sass to map 1:1 to the intended SASS), 512 threads/CTA, 2 CTAs/SM (32 warps/SM, 64 registers each), timed with
// CUDA events over all SMs; rate = lane-ops / (elapsed * SMs * sm_clock) where sm_clock is measured in the
// same launch from clock64() vs globaltimer.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define CHAINS 8
#define UNROLL 16          // ops per chain per loop iteration
#define ITERS 4096

enum Op { OP_IADD3, OP_IMAD, OP_IDP4A, OP_VABSDIFF4, OP_LOP3, OP_PRMT, OP_SETP_PADD, OP_MIX_IADD_IMAD, OP_LDS, OP_K1MIX, OP_COUNT };
static const char* op_names[OP_COUNT] = {"iadd3", "imad", "idp4a", "vabsdiff4", "lop3", "prmt", "isetp_plus_pred_iadd",
                                         "iadd3_imad_interleaved", "lds32", "k1_inner_mix"};
// lane-ops counted per "op" of the unrolled body (e.g. the K1 mix is one window test = 9 instructions)
// two dependent PTX adds fuse into one IADD3 (checked in SASS): 0.5 instruction per counted add
static const double op_instr[OP_COUNT] = {0.5, 1, 1, 1, 1, 1, 2, 2, 1, 9};

template <int OP>
__global__ void __launch_bounds__(512, 2) rate_kernel(uint32_t* out, uint32_t seed, long long* cycles, unsigned long long* ns)
{
    __shared__ uint32_t sm[1024 + 64];
    sm[threadIdx.x] = seed * threadIdx.x + 12345u;
    sm[threadIdx.x + 512] = seed * threadIdx.x + 999u;
    if (threadIdx.x < 64) sm[1024 + threadIdx.x] = threadIdx.x;
    __syncthreads();
    uint32_t a[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) a[k] = seed + k * 77u + threadIdx.x;
    uint32_t b = seed | 1u, c = seed * 3u + 7u;
    uint32_t acc1 = 0, acc2 = 0, acc3 = 0;
    const uint32_t* lp = sm + (threadIdx.x & 31);
    unsigned long long t0 = 0;
    long long c0 = 0;
    if (threadIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        c0 = clock64();
    }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int k = 0; k < CHAINS; k++) {
                if (OP == OP_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(b));
                if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
                if (OP == OP_IDP4A) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(b), "r"(c));
                if (OP == OP_VABSDIFF4) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(0));
                if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b), "r"(c));
                if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x4240;" : "+r"(a[k]) : "r"(b));
                if (OP == OP_SETP_PADD)
                    asm volatile("{ .reg .pred p; setp.le.s32 p, %0, %1; @p add.u32 %0, %0, %2; }" : "+r"(a[k]) : "r"(b), "r"(c));
                if (OP == OP_MIX_IADD_IMAD) {
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(b));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc1) : "r"(b), "r"(a[k]));
                }
                if (OP == OP_LDS) {
                    uint32_t v;
                    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(lp + ((u * CHAINS + k) & 31) * 32)));
                    a[k] ^= v;   // one LOP per load (counted as 1 op: the LDS)
                }
                if (OP == OP_K1MIX) {
                    // one K1 window test as the compiler emits it: LDS, VABSDIFF4, IDP4A, ISETP, 2 PRMT, 3 predicated adds
                    asm volatile(
                        "{ .reg .pred p; .reg .b32 t, e, d, lo, hi;\n"
                        "  ld.volatile.shared.u32 t, [%3];\n"
                        "  vabsdiff4.u32.u32.u32 e, t, %4, %5;\n"
                        "  dp4a.u32.u32 d, e, e, %5;\n"
                        "  setp.le.s32 p, d, %6;\n"
                        "  prmt.b32 lo, t, %5, 0x4240;\n"
                        "  prmt.b32 hi, t, %5, 0x4341;\n"
                        "  @p add.u32 %0, %0, lo;\n"
                        "  @p add.u32 %1, %1, hi;\n"
                        "  @p add.u32 %2, %2, %7; }"
                        : "+r"(a[k]), "+r"(a[(k + 4) & 7]), "+r"(acc3)
                        : "r"((uint32_t)__cvta_generic_to_shared(lp + ((u * CHAINS + k) & 31) * 32)), "r"(c), "r"(0), "r"(b), "r"(u));
                }
            }
        }
    }
    uint32_t r = acc1 ^ acc2 ^ acc3;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) r ^= a[k];
    if (threadIdx.x == 0) {
        long long c1 = clock64();
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        cycles[blockIdx.x] = c1 - c0;
        ns[blockIdx.x] = t1 - t0;
    }
    if (r == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = r;   // keep the chains alive
}

template <int OP>
static void run(int sms, uint32_t* d_out, long long* d_cyc, unsigned long long* d_ns, bool last)
{
    int blocks = sms * 2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    rate_kernel<OP><<<blocks, 512>>>(d_out, 17u, d_cyc, d_ns);   // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    double mhz = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        rate_kernel<OP><<<blocks, 512>>>(d_out, 17u + rep, d_cyc, d_ns);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) {
            best = ms;
            long long cyc[8];
            unsigned long long ns[8];
            cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
            cudaMemcpy(ns, d_ns, sizeof(ns), cudaMemcpyDeviceToHost);
            mhz = ns[0] ? (double)cyc[0] / (double)ns[0] * 1e3 : 0;
        }
    }
    double ops = (double)blocks * 512.0 * ITERS * UNROLL * CHAINS;   // "ops" (see op_instr for instructions per op)
    double gops = ops / (best * 1e-3) / 1e9;
    double per_clk_sm = mhz > 0 ? ops / (best * 1e-3) / (mhz * 1e6) / sms : 0;
    printf("  \"%s\": {\"ms\": %.4f, \"gops\": %.1f, \"ginstr\": %.1f, \"sm_mhz\": %.0f, \"ops_per_clk_sm\": %.2f, \"instr_per_clk_sm\": %.2f}%s\n",
           op_names[OP], best, gops, gops * op_instr[OP], mhz, per_clk_sm, per_clk_sm * op_instr[OP], last ? "" : ",");
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

int main()
{
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
    int sms = p.multiProcessorCount;
    uint32_t* d_out;
    long long* d_cyc;
    unsigned long long* d_ns;
    cudaMalloc(&d_out, (size_t)sms * 2 * 512 * 4);
    cudaMalloc(&d_cyc, sizeof(long long) * sms * 2);
    cudaMalloc(&d_ns, sizeof(unsigned long long) * sms * 2);
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"chains\": %d,\n", p.name, sms, CHAINS);
    run<OP_IADD3>(sms, d_out, d_cyc, d_ns, false);
    run<OP_IMAD>(sms, d_out, d_cyc, d_ns, false);
    run<OP_IDP4A>(sms, d_out, d_cyc, d_ns, false);
    run<OP_VABSDIFF4>(sms, d_out, d_cyc, d_ns, false);
    run<OP_LOP3>(sms, d_out, d_cyc, d_ns, false);
    run<OP_PRMT>(sms, d_out, d_cyc, d_ns, false);
    run<OP_SETP_PADD>(sms, d_out, d_cyc, d_ns, false);
    run<OP_MIX_IADD_IMAD>(sms, d_out, d_cyc, d_ns, false);
    run<OP_LDS>(sms, d_out, d_cyc, d_ns, false);
    run<OP_K1MIX>(sms, d_out, d_cyc, d_ns, true);
    printf("}\n");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
    return 0;
}
