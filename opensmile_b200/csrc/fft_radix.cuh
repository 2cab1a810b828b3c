// fft_radix.cuh -- register-level radix-2/4/8/16 DFT butterflies and the factorisation /
// digit-reversal helpers of the in-place decimation-in-frequency FFT used by the fused LLD
// kernel.  __host__ __device__ so that tests/native can check them against a naive DFT on the
// CPU without a GPU.
#pragma once
#include <cuda_runtime.h>

namespace osm {

// ------------------------------------------------------------------------------------------
// complex helpers
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 w)
{
  return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}
// multiply by -i
__host__ __device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }

// forward DFTs (kernel e^{-2 pi i nk/R}), in place.  X[q] ends up in v[Dft<R>::out(q)].
template <int R> struct Dft;

template <> struct Dft<2> {
  static __host__ __device__ __forceinline__ void run(float2 (&v)[2])
  {
    float2 t = v[0];
    v[0] = cadd(t, v[1]);
    v[1] = csub(t, v[1]);
  }
  static __host__ __device__ constexpr int out(int q) { return q; }
};

__host__ __device__ __forceinline__ void dft4(float2 &a, float2 &b, float2 &c, float2 &d)
{
  float2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
  a = cadd(t0, t2);
  c = csub(t0, t2);
  b = make_float2(t1.x + t3.y, t1.y - t3.x);   // t1 - i t3
  d = make_float2(t1.x - t3.y, t1.y + t3.x);   // t1 + i t3
}

template <> struct Dft<4> {
  static __host__ __device__ __forceinline__ void run(float2 (&v)[4]) { dft4(v[0], v[1], v[2], v[3]); }
  static __host__ __device__ constexpr int out(int q) { return q; }
};

// 8 = 2 x 4 : n = 4 n1 + n2, k = k1 + 2 k2
template <> struct Dft<8> {
  static __host__ __device__ __forceinline__ void run(float2 (&v)[8])
  {
    const float c = 0.70710678118654752440f;
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) {
      float2 t = v[n2];
      v[n2] = cadd(t, v[n2 + 4]);
      v[n2 + 4] = csub(t, v[n2 + 4]);
    }
    // twiddle W8^{n2} on the k1 = 1 outputs
    { float2 a = v[5]; v[5] = make_float2(c * (a.x + a.y), c * (a.y - a.x)); }   // (1 - i)/sqrt2
    v[6] = cmul_mi(v[6]);                                                        // -i
    { float2 a = v[7]; v[7] = make_float2(c * (a.y - a.x), -c * (a.x + a.y)); }  // (-1 - i)/sqrt2
    dft4(v[0], v[1], v[2], v[3]);
    dft4(v[4], v[5], v[6], v[7]);
  }
  static __host__ __device__ constexpr int out(int q) { return 4 * (q % 2) + q / 2; }
};

// 16 = 4 x 4 : n = 4 n1 + n2, k = k1 + 4 k2
template <> struct Dft<16> {
  static __host__ __device__ __forceinline__ void run(float2 (&v)[16])
  {
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) dft4(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);
    // v[n2 + 4 k1] *= W16^{n2 k1}
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;  // cos, sin(pi/8)
    const float c2 = 0.70710678118654752440f;
    v[5] = cmul(v[5], make_float2(c1, -s1));                                  // W^1
    { float2 a = v[6]; v[6] = make_float2(c2 * (a.x + a.y), c2 * (a.y - a.x)); }   // W^2
    v[7] = cmul(v[7], make_float2(s1, -c1));                                  // W^3
    { float2 a = v[9]; v[9] = make_float2(c2 * (a.x + a.y), c2 * (a.y - a.x)); }   // W^2
    v[10] = cmul_mi(v[10]);                                                   // W^4
    { float2 a = v[11]; v[11] = make_float2(c2 * (a.y - a.x), -c2 * (a.x + a.y)); }  // W^6
    v[13] = cmul(v[13], make_float2(s1, -c1));                                // W^3
    { float2 a = v[14]; v[14] = make_float2(c2 * (a.y - a.x), -c2 * (a.x + a.y)); }  // W^6
    v[15] = cmul(v[15], make_float2(-c1, s1));                                // W^9
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  }
  static __host__ __device__ constexpr int out(int q) { return 4 * (q % 4) + q / 4; }
};

// ------------------------------------------------------------------------------------------
// factorisation of the complex FFT size M = N/2
// ------------------------------------------------------------------------------------------
template <int M> struct Fact;
template <> struct Fact<256>  { static constexpr int NS = 2, R0 = 16, R1 = 16, R2 = 1; };
template <> struct Fact<512>  { static constexpr int NS = 3, R0 = 8,  R1 = 8,  R2 = 8; };
template <> struct Fact<1024> { static constexpr int NS = 3, R0 = 16, R1 = 16, R2 = 4; };
template <> struct Fact<2048> { static constexpr int NS = 3, R0 = 16, R1 = 16, R2 = 8; };

// position of X[k] after the in-place DIF passes (digit reversal)
template <int M>
__host__ __host__ __device__ __forceinline__ int fft_pos(int k)
{
  using Fc = Fact<M>;
  const int q1 = k % Fc::R0;
  const int k2 = k / Fc::R0;
  if (Fc::NS == 2) return q1 * (M / Fc::R0) + k2;
  const int q2 = k2 % Fc::R1;
  const int q3 = k2 / Fc::R1;
  return q1 * (M / Fc::R0) + q2 * (M / (Fc::R0 * Fc::R1)) + q3;
}


}  // namespace osm
