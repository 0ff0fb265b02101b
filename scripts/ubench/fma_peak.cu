// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/fma_peak scripts/ubench/fma_peak.cu
// run on a B200: scripts/ubench/fma_peak > profiles/r02_fma_peaks.json
// Pipe THROUGHPUT on sm_100a, all SMs, ILP 8 per thread, 16 warps per SM (4 per scheduler), CUDA events:
//   DFMA / DADD / DADD.RM / DMUL (FP64 pipe), FFMA (FMA pipe), F2F.F64.F16, F2F.F32.F64 + F2F.F64.F32 (conversion unit),
//   MUFU.RCP64H, 32-bit SHFL.BFLY, HADD2, IMAD.MOV, and two mixes in the Gauss-Newton kernel's proportions.
// These are the denominators of bench.py's roofline fractions ("builder must confirm with an FMA microbenchmark").
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int OP>
__global__ void __launch_bounds__(128) tput(double* out, double seed, int n)
{
    double d0 = seed + threadIdx.x, d1 = d0 * 0.5, d2 = d0 * 0.25, d3 = d0 * 0.125, d4 = d0 + 1, d5 = d0 + 2, d6 = d0 + 3, d7 = d0 + 4;
    float f0 = (float)d0, f1 = (float)d1, f2 = (float)d2, f3 = (float)d3, f4 = (float)d4, f5 = (float)d5, f6 = (float)d6, f7 = (float)d7;
    unsigned u0 = threadIdx.x | 0x3c003c00u, u1 = u0 + 1, u2 = u0 + 2, u3 = u0 + 3, u4 = u0 + 4, u5 = u0 + 5, u6 = u0 + 6, u7 = u0 + 7;
    const double db = 1.0000001, dc = 1e-9;
    const float fb = 1.0000001f, fc = 1e-9f;
    for (int i = 0; i < n; ++i) {
        if (OP == 0) {
#define X(k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d##k) : "d"(db), "d"(dc));
            REP8(X) REP8(X) REP8(X) REP8(X) REP8(X) REP8(X) REP8(X) REP8(X)
#undef X
        } else if (OP == 1) {
#define X(k) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d##k) : "d"(dc));
            REP8(X) REP8(X) REP8(X) REP8(X) REP8(X) REP8(X) REP8(X) REP8(X)
#undef X
        } else if (OP == 2) {
#define X(k) asm volatile("add.rm.f64 %0, %0, %1;" : "+d"(d##k) : "d"(dc));
            REP8(X) REP8(X) REP8(X) REP8(X) REP8(X) REP8(X) REP8(X) REP8(X)
#undef X
        } else if (OP == 3) {
#define X(k) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d##k) : "d"(db));
            REP8(X) REP8(X) REP8(X) REP8(X) REP8(X) REP8(X) REP8(X) REP8(X)
#undef X
        } else if (OP == 4) {
#define X(k) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f##k) : "f"(fb), "f"(fc));
            REP8(X) REP8(X) REP8(X) REP8(X) REP8(X) REP8(X) REP8(X) REP8(X)
#undef X
        } else if (OP == 5) {          // F2F.F64.F16 (independent: the source is an integer register that a cheap IADD advances)
#define X(k) { double t; asm volatile("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f64.f16 %0, lo; }" : "=d"(t) : "r"(u##k)); d##k += 0; u##k ^= __double2hiint(t) & 1; }
            REP8(X) REP8(X)
#undef X
        } else if (OP == 6) {          // F2F.F32.F64 then F2F.F64.F32 (two conversions per X)
#define X(k) { float t; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(t) : "d"(d##k)); asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d##k) : "f"(t)); }
            REP8(X) REP8(X)
#undef X
        } else if (OP == 7) {          // MUFU.RCP64H
#define X(k) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(d##k));
            REP8(X) REP8(X)
#undef X
        } else if (OP == 8) {          // SHFL.BFLY (32-bit)
#define X(k) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(u##k));
            REP8(X) REP8(X)
#undef X
        } else if (OP == 9) {          // HADD2
#define X(k) asm volatile("sub.rn.f16x2 %0, %0, %1;" : "+r"(u##k) : "r"(0x00010001u));
            REP8(X) REP8(X)
#undef X
        } else if (OP == 10) {         // mix of the Gauss-Newton sample body: 40 FP64 : 12 F2F.F64.F16 (here 16 : 5, rounded up)
#define X(k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d##k) : "d"(db), "d"(dc));
            REP8(X) REP8(X)
#undef X
#define X(k) { double t; asm volatile("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f64.f16 %0, lo; }" : "=d"(t) : "r"(u##k)); u##k ^= __double2hiint(t) & 1; }
            X(0) X(1) X(2) X(3) X(4)
#undef X
        } else if (OP == 11) {         // 16 FP64 : 16 integer moves (does the ALU pipe issue beside the FP64 pipe?)
#define X(k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d##k) : "d"(db), "d"(dc));
            REP8(X) REP8(X)
#undef X
#define X(k) asm volatile("xor.b32 %0, %0, %1;" : "+r"(u##k) : "r"(i));
            REP8(X) REP8(X)
#undef X
        } else if (OP == 12) {         // integer ALU alone
#define X(k) asm volatile("xor.b32 %0, %0, %1;" : "+r"(u##k) : "r"(i));
            REP8(X) REP8(X)
#undef X
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 + (double)(u0 ^ u1 ^ u2 ^ u3 ^ u4 ^ u5 ^ u6 ^ u7);
}

template <int OP>
static double run(double* out, int sms, int n, int ctasPerSm)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tput<OP><<<sms * ctasPerSm, 128>>>(out, 1.0, n / 8);
    cudaEventRecord(e0);
    tput<OP><<<sms * ctasPerSm, 128>>>(out, 1.0, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = pr.multiProcessorCount, n = 20000, cps = 4;
    double* out; cudaMalloc(&out, sizeof(double) * sms * cps * 128);
    const char* names[13] = {"dfma", "dadd", "dadd_rm", "dmul", "ffma", "f2f_f64_f16", "f2f_f32_f64_roundtrip", "mufu_rcp64h", "shfl_bfly_b32", "hadd2",
                             "mix_16dfma_5f2f", "mix_16dfma_16alu", "alu_xor"};
    const double per_iter[13] = {64, 64, 64, 64, 64, 16, 32, 16, 16, 16, 21, 32, 16};
    double ms[13];
    ms[0] = run<0>(out, sms, n, cps); ms[1] = run<1>(out, sms, n, cps); ms[2] = run<2>(out, sms, n, cps); ms[3] = run<3>(out, sms, n, cps);
    ms[4] = run<4>(out, sms, n, cps); ms[5] = run<5>(out, sms, n, cps); ms[6] = run<6>(out, sms, n, cps); ms[7] = run<7>(out, sms, n, cps);
    ms[8] = run<8>(out, sms, n, cps); ms[9] = run<9>(out, sms, n, cps); ms[10] = run<10>(out, sms, n, cps); ms[11] = run<11>(out, sms, n, cps);
    ms[12] = run<12>(out, sms, n, cps);
    const double warps = (double)sms * cps * 4;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_attr_mhz\": %.0f, \"warps_per_sm\": %d, \"ilp\": 8,\n \"how\": \"scripts/ubench/fma_peak.cu: n=%d loop iterations of 64 (FP64 / FP32 arithmetic), 16 (the others) or the stated mix of independent inline-PTX instructions per thread, CUDA events, second launch timed\",\n \"ops\": {\n",
           pr.name, sms, clk_khz / 1e3, cps * 4, n);
    for (int k = 0; k < 13; ++k) {
        const double winst = warps * n * per_iter[k];                 // warp instructions of the named kind(s)
        const double per_sm_clk = winst / (ms[k] * 1e-3) / sms / (clk_khz * 1e3);   // warp instructions per SM per clock (at the attribute clock)
        printf("  \"%s\": {\"ms\": %.3f, \"warp_inst_per_sm_per_clk\": %.3f, \"lane_ops_per_sm_per_clk\": %.1f}%s\n", names[k], ms[k], per_sm_clk, per_sm_clk * 32, k < 12 ? "," : "");
    }
    const double dfma_tf = warps * n * 64 * 32 * 2 / (ms[0] * 1e-3) / 1e12, ffma_tf = warps * n * 64 * 32 * 2 / (ms[4] * 1e-3) / 1e12;
    printf(" },\n \"fp64_fma_tflops\": %.2f, \"fp32_fma_tflops\": %.2f\n}\n", dfma_tf, ffma_tf);
    return 0;
}
