// diag.cu -- measured integer-pipe peak of the device the library runs on.
//
// The sketch kernels are bound by the INT32 ALU pipe (LOP3 / SHF / IADD3 / ISETP), for which the
// driver-written MEASURED_PEAKS.json has no entry.  sw_measure_int_peak times dependent chains of those
// instructions (8 independent chains per thread, enough warps to fill every scheduler) with CUDA
// events and reports lane-operations per second; bench.py uses the result as the denominator of the
// sketch kernel's roofline fraction, next to the per-op figures.
#include <cuda_runtime.h>

#include "device.h"

namespace sw {
namespace {

constexpr int kChains = 8;
constexpr int kUnroll = 16;

template <int OP>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t* out, int iters, uint32_t c)
{
    uint32_t a[kChains];
#pragma unroll
    for (int j = 0; j < kChains; ++j) a[j] = threadIdx.x * 2654435761u + j * 40503u + blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
            for (int j = 0; j < kChains; ++j) {
                const uint32_t b = a[(j + 1) & (kChains - 1)];
                if (OP == 0 || (OP == 3 && (u % 3) == 0))
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b), "r"(c));
                else if (OP == 1 || (OP == 3 && (u % 3) == 1))
                    asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(c));
                else
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[j]) : "r"(b));
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < kChains; ++j) r ^= a[j];
    if (r == 0x12345u) out[0] = r;   // keeps the chains alive; practically never true
}

template <int OP>
double run_one(uint32_t* d_out, int grid, int iters, cudaStream_t s)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int_peak_kernel<OP><<<grid, 256, 0, s>>>(d_out, iters / 8, 13u);   // warm-up
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, s);
        int_peak_kernel<OP><<<grid, 256, 0, s>>>(d_out, iters, 13u);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)grid * 256.0 * iters * kUnroll * kChains;
        best = std::max(best, ops / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

}  // namespace

void measure_int_peak(double out[4], cudaStream_t s)
{
    DevBuf<uint32_t> d_out(4, s, true);
    const int grid = sm_count() * 8;
    const int iters = 4096;
    out[0] = run_one<0>(d_out.p, grid, iters, s);
    out[1] = run_one<1>(d_out.p, grid, iters, s);
    out[2] = run_one<2>(d_out.p, grid, iters, s);
    out[3] = run_one<3>(d_out.p, grid, iters, s);
    SW_CUDA(cudaGetLastError());
}

}  // namespace sw
