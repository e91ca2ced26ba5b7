// FP64 peak probes for the K2 roofline ("FP64 peak: must be measured"): a DFMA loop on the FP64
// pipe and a DMMA loop (mma.sync.m8n8k4.f64, the instruction k2_contract_mma issues).  Debug / measurement entry
// point agnf_debug_fp64_peak; not on any product path.
#include <cstdint>
#include <cuda_runtime.h>

namespace agnf {

namespace {

constexpr int PEAK_THREADS = 256;
constexpr int PEAK_CHAINS = 8;

__global__ void __launch_bounds__(PEAK_THREADS) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x[PEAK_CHAINS];
#pragma unroll
  for (int c = 0; c < PEAK_CHAINS; ++c) x[c] = (double)(threadIdx.x + c) * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < PEAK_CHAINS; ++c) x[c] = fma(x[c], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < PEAK_CHAINS; ++c) s += x[c];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;     // never true: keeps the chains alive
}

__global__ void __launch_bounds__(PEAK_THREADS) dmma_peak_kernel(double* out, int iters, double a, double b) {
  double d[PEAK_CHAINS][2];
#pragma unroll
  for (int c = 0; c < PEAK_CHAINS; ++c) d[c][0] = d[c][1] = (double)(threadIdx.x + c) * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < PEAK_CHAINS; ++c)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
                   : "+d"(d[c][0]), "+d"(d[c][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < PEAK_CHAINS; ++c) s += d[c][0] + d[c][1];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

// kind 0: DFMA, 1: DMMA.  Returns the launch's flop count; the caller times it with events.
cudaError_t launch_fp64_peak(int kind, double* scratch, int blocks, int iters, double* flops, cudaStream_t stream) {
  const double threads = (double)blocks * PEAK_THREADS;
  if (kind == 0) {
    dfma_peak_kernel<<<blocks, PEAK_THREADS, 0, stream>>>(scratch, iters, 0.999999, 1e-9);
    *flops = threads * iters * PEAK_CHAINS * 2.0;
  } else {
    dmma_peak_kernel<<<blocks, PEAK_THREADS, 0, stream>>>(scratch, iters, 0.5, 1e-9);
    *flops = (threads / 32.0) * iters * PEAK_CHAINS * (8.0 * 8.0 * 4.0 * 2.0);
  }
  return cudaGetLastError();
}

}  // namespace agnf
