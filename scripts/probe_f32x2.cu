#include <cuda_runtime.h>
#include <cstdio>
// throughput probe: scalar FADD / FFMA chains vs packed f32x2
template <int MODE>
__global__ void probe(float *out, int iters)
{
  float2 a[8], b = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) { a[i].x = a[i].x + b.x; a[i].y = a[i].y + b.y; }
      else if (MODE == 1) a[i] = __fadd2_rn(a[i], b);
      else if (MODE == 2) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
      else a[i] = __ffma2_rn(a[i], b, c);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> float run(float *d, int iters)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<148 * 4, 256>>>(d, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); probe<MODE><<<148 * 4, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
  float *d; cudaMalloc(&d, 148 * 4 * 256 * 4);
  const int it = 20000;
  const double ops = 148.0 * 4 * 256 * it * 16;   // scalar-equivalent ops
  float t;
  t = run<0>(d, it); printf("FADD  scalar: %.3f ms  %.2f Tlane-op/s\n", t, ops / t / 1e9);
  t = run<1>(d, it); printf("FADD2 packed: %.3f ms  %.2f Tlane-op/s\n", t, ops / t / 1e9);
  t = run<2>(d, it); printf("FFMA  scalar: %.3f ms  %.2f Tlane-op/s\n", t, ops / t / 1e9);
  t = run<3>(d, it); printf("FFMA2 packed: %.3f ms  %.2f Tlane-op/s\n", t, ops / t / 1e9);
  return 0;
}
