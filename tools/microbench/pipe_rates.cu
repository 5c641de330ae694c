// Issue-rate microbenchmark of the instructions of the policy epilogue on sm_100a (B200): cycles per warp instruction per SM
// sub-partition for FFMA, FFMA.SAT, FMUL, FMNMX, FHFMA (fma.rn.f32.f16), F2FP (cvt.rn.f16x2.f32), MUFU.EX2, MUFU.RCP alone and in the
// pairs that matter (do two instruction kinds share a pipe: time(A+B) = time(A) + time(B), or overlap: max).
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 256, CH = 8;
#define OP_FFMA(x)    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(c1), "f"(c2));
#define OP_FFMASAT(x) asm volatile("fma.rn.ftz.sat.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(c1), "f"(c2));
#define OP_FADD(x)    asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(c2));
#define OP_FMUL(x)    asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(c1));
#define OP_FMNMX(x)   asm volatile("min.NaN.f32 %0, %0, %1;" : "+f"(x) : "f"(c2));
#define OP_EX2(x)     asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x));
#define OP_RCP(x)     asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x) :: "memory");
#define OP_TANH(x)    asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x));
#define OP_F2FP(x)    asm volatile("{ .reg .b32 t; cvt.rn.f16x2.f32 t, %0, %1; mov.b32 %0, t; }" : "+f"(x) : "f"(c1));
#define OP_FHFMA(x)   asm volatile("{ .reg .f16 h; mov.b16 h, 0x3C00; fma.rn.f32.f16 %0, h, h, %0; }" : "+f"(x));
#define OP_IADD(x)    asm volatile("{ .reg .b32 t; mov.b32 t, %0; add.s32 t, t, 12345; mov.b32 %0, t; }" : "+f"(x));
#define OP_HFMA2(x)   asm volatile("{ .reg .b32 t; mov.b32 t, %0; fma.rn.f16x2 t, t, t, t; mov.b32 %0, t; }" : "+f"(x));
#define OP_EX2H2(x)   asm volatile("{ .reg .b32 t; mov.b32 t, %0; ex2.approx.f16x2 t, t; mov.b32 %0, t; }" : "+f"(x));
#define OP_TANHH2(x)  asm volatile("{ .reg .b32 t; mov.b32 t, %0; tanh.approx.f16x2 t, t; mov.b32 %0, t; }" : "+f"(x));

#define KERNEL(name, BODY)                                                                                   \
    __global__ void __launch_bounds__(1024, 1) k_##name(float* out, long long* cyc, float c1, float c2) {      \
        float v[CH];                                                                                         \
        for (int j = 0; j < CH; ++j) v[j] = threadIdx.x * 1e-3f + j;                                         \
        __syncthreads();                                                                                     \
        long long t0 = clock64();                                                                            \
        _Pragma("unroll 2") for (int i = 0; i < ITER; ++i) {                                                                   \
            _Pragma("unroll") for (int j = 0; j < CH; ++j) { BODY(v[j]) }                                    \
        }                                                                                                    \
        __syncthreads();                                                                                     \
        long long t1 = clock64();                                                                                   \
        float s = 0;                                                                                         \
        for (int j = 0; j < CH; ++j) s += v[j];                                                              \
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;                                                      \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                                     \
    }
#define B2(A, B) A B
#define BODY_FFMA(x) OP_FFMA(x)
#define BODY_FFMASAT(x) OP_FFMASAT(x)
#define BODY_FMUL(x) OP_FMUL(x)
#define BODY_FMNMX(x) OP_FMNMX(x)
#define BODY_EX2(x) OP_EX2(x)
#define BODY_RCP(x) OP_RCP(x)
#define BODY_TANH(x) OP_TANH(x)
#define BODY_F2FP(x) OP_F2FP(x)
#define BODY_FHFMA(x) OP_FHFMA(x)
#define BODY_IADD(x) OP_IADD(x)
#define BODY_HFMA2(x) OP_HFMA2(x)
#define BODY_EX2H2(x) OP_EX2H2(x)
#define BODY_TANHH2(x) OP_TANHH2(x)
#define BODY_EX2_F2FP(x) OP_EX2(x) OP_F2FP(x)
#define BODY_EX2_FHFMA(x) OP_EX2(x) OP_FHFMA(x)
#define BODY_EX2_FFMA(x) OP_EX2(x) OP_FFMA(x)
#define BODY_F2FP_FHFMA(x) OP_F2FP(x) OP_FHFMA(x)
#define BODY_F2FP_FFMA(x) OP_F2FP(x) OP_FFMA(x)
#define BODY_F2FP_FMNMX(x) OP_F2FP(x) OP_FMNMX(x)
#define BODY_FHFMA_FFMA(x) OP_FHFMA(x) OP_FFMA(x)
#define BODY_FFMA_FMNMX(x) OP_FFMA(x) OP_FMNMX(x)
#define BODY_EX2_7FFMA(x) OP_EX2(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x)
#define BODY_EPI(x) OP_EX2(x) OP_FFMASAT(x) OP_FMUL(x) OP_FMUL(x) OP_FFMA(x) OP_F2FP(x) OP_FHFMA(x) OP_RCP(x) OP_FMUL(x) OP_FFMA(x) OP_EX2(x) OP_FFMASAT(x) OP_FMUL(x) OP_FHFMA(x) OP_EX2(x) OP_FFMASAT(x) OP_FMUL(x) OP_FFMA(x) OP_F2FP(x) OP_FHFMA(x) OP_EX2(x) OP_FFMASAT(x) OP_FMUL(x) OP_FFMA(x) OP_FHFMA(x) OP_FMUL(x) OP_FMUL(x) OP_FMUL(x)
#define BODY_OLD3(x) OP_FMNMX(x) OP_EX2(x) OP_FADD(x) OP_FMUL(x) OP_FFMA(x) OP_FMUL(x) OP_FFMA(x) OP_FMUL(x)
#define BODY_NEW2(x) OP_EX2(x) OP_FFMASAT(x) OP_FMUL(x) OP_FFMA(x) OP_FMUL(x) OP_FFMA(x) OP_FMUL(x)
#define BODY_EX2_3FFMA(x) OP_EX2(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x)
#define BODY_EX2_5FFMA(x) OP_EX2(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x)
#define BODY_EX2_9FFMA(x) OP_EX2(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x)
#define BODY_EX2_11FFMA(x) OP_EX2(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x)
#define BODY_EX2_4FFMA_3ALU(x) OP_EX2(x) OP_FFMA(x) OP_FMNMX(x) OP_FFMA(x) OP_FMNMX(x) OP_FFMA(x) OP_FMNMX(x) OP_FFMA(x)
#define BODY_3FFMA_FMNMX(x) OP_FFMA(x) OP_FFMA(x) OP_FFMA(x) OP_FMNMX(x)
KERNEL(OLD3, BODY_OLD3) KERNEL(NEW2, BODY_NEW2) KERNEL(EX2_3FFMA, BODY_EX2_3FFMA) KERNEL(EX2_5FFMA, BODY_EX2_5FFMA) KERNEL(EX2_9FFMA, BODY_EX2_9FFMA)
KERNEL(EX2_11FFMA, BODY_EX2_11FFMA) KERNEL(EX2_4FFMA_3ALU, BODY_EX2_4FFMA_3ALU) KERNEL(3FFMA_FMNMX, BODY_3FFMA_FMNMX)
KERNEL(FFMA, BODY_FFMA) KERNEL(FFMASAT, BODY_FFMASAT) KERNEL(FMUL, BODY_FMUL) KERNEL(FMNMX, BODY_FMNMX) KERNEL(EX2, BODY_EX2)
KERNEL(RCP, BODY_RCP) KERNEL(TANH, BODY_TANH) KERNEL(F2FP, BODY_F2FP) KERNEL(FHFMA, BODY_FHFMA) KERNEL(IADD, BODY_IADD) KERNEL(HFMA2, BODY_HFMA2)
KERNEL(EX2H2, BODY_EX2H2) KERNEL(TANHH2, BODY_TANHH2)
KERNEL(EX2_F2FP, BODY_EX2_F2FP) KERNEL(EX2_FHFMA, BODY_EX2_FHFMA) KERNEL(EX2_FFMA, BODY_EX2_FFMA) KERNEL(F2FP_FHFMA, BODY_F2FP_FHFMA)
KERNEL(F2FP_FFMA, BODY_F2FP_FFMA) KERNEL(F2FP_FMNMX, BODY_F2FP_FMNMX) KERNEL(FHFMA_FFMA, BODY_FHFMA_FFMA) KERNEL(FFMA_FMNMX, BODY_FFMA_FMNMX)
KERNEL(EX2_7FFMA, BODY_EX2_7FFMA) KERNEL(EPI, BODY_EPI)

template <typename K>
void run(const char* name, K kern, int ops_per_body, int warps) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int rep = 0; rep < 2; ++rep) kern<<<148, warps * 32>>>(out, cyc, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double winst_per_smsp = (double)ITER * CH * ops_per_body * warps / 4.0;
    printf("%-14s warps/SM %2d  cycles %9.0f  -> %6.3f cycles per warp-instruction per sub-partition (%5.3f inst/clk)  err=%d\n", name, warps, avg,
           avg / winst_per_smsp, winst_per_smsp / avg, (int)cudaGetLastError());
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int warps : {4, 16, 24}) {
        run("FFMA", k_FFMA, 1, warps); run("FFMA.SAT", k_FFMASAT, 1, warps); run("FMUL", k_FMUL, 1, warps); run("FMNMX", k_FMNMX, 1, warps);
        run("IADD", k_IADD, 1, warps); run("HFMA2", k_HFMA2, 1, warps);
        run("FHFMA", k_FHFMA, 1, warps); run("F2FP", k_F2FP, 1, warps); run("MUFU.EX2", k_EX2, 1, warps); run("MUFU.RCP", k_RCP, 1, warps);
        run("MUFU.TANH", k_TANH, 1, warps); run("EX2.F16x2", k_EX2H2, 1, warps); run("TANH.F16x2", k_TANHH2, 1, warps);
        run("EX2+F2FP", k_EX2_F2FP, 2, warps); run("EX2+FHFMA", k_EX2_FHFMA, 2, warps); run("EX2+FFMA", k_EX2_FFMA, 2, warps);
        run("F2FP+FHFMA", k_F2FP_FHFMA, 2, warps); run("F2FP+FFMA", k_F2FP_FFMA, 2, warps); run("F2FP+FMNMX", k_F2FP_FMNMX, 2, warps);
        run("FHFMA+FFMA", k_FHFMA_FFMA, 2, warps); run("FFMA+FMNMX", k_FFMA_FMNMX, 2, warps); run("EX2+7FFMA", k_EX2_7FFMA, 8, warps);
        run("EPI mix(28)", k_EPI, 28, warps);
        run("EX2+3FFMA", k_EX2_3FFMA, 4, warps); run("EX2+5FFMA", k_EX2_5FFMA, 6, warps); run("EX2+9FFMA", k_EX2_9FFMA, 10, warps);
        run("EX2+11FFMA", k_EX2_11FFMA, 12, warps); run("EX2+4FFMA+3ALU", k_EX2_4FFMA_3ALU, 8, warps); run("3FFMA+FMNMX", k_3FFMA_FMNMX, 4, warps);
        run("old MIN,EX2,FADD+5", k_OLD3, 8, warps); run("new EX2,FFMA.SAT+5", k_NEW2, 7, warps);
    }
    return 0;
}
