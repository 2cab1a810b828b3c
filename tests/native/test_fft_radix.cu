// Host-side check of the register butterflies / DIF index algebra used by the fused kernel
// (opensmile_b200/csrc/fft_radix.cuh) against a naive O(n^2) DFT.  Runs on the CPU.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../opensmile_b200/csrc/fft_radix.cuh"

using namespace osm;

static double maxerr = 0;

template <int R>
static void check_dft()
{
  float2 v[R];
  std::vector<double> re(R), im(R);
  for (int i = 0; i < R; i++) { re[i] = drand48() - 0.5; im[i] = drand48() - 0.5; v[i] = make_float2((float)re[i], (float)im[i]); }
  Dft<R>::run(v);
  for (int q = 0; q < R; q++) {
    double xr = 0, xi = 0;
    for (int n = 0; n < R; n++) {
      const double a = -2.0 * M_PI * n * q / R;
      xr += re[n] * cos(a) - im[n] * sin(a);
      xi += re[n] * sin(a) + im[n] * cos(a);
    }
    const float2 g = v[Dft<R>::out(q)];
    maxerr = fmax(maxerr, fmax(fabs(g.x - xr), fabs(g.y - xi)));
  }
}

template <int M, int R, int MS, bool LAST>
static void stage(std::vector<float2> &Z)
{
  constexpr int stride = MS / R;
  for (int t = 0; t < M / R; t++) {
    const int blk = t / stride, j = t % stride, base = blk * MS + j;
    float2 v[R];
    for (int r = 0; r < R; r++) v[r] = Z[base + stride * r];
    Dft<R>::run(v);
    for (int q = 0; q < R; q++) {
      float2 x = v[Dft<R>::out(q)];
      if (!LAST) {
        const double a = -2.0 * M_PI * (double)j * q / MS;
        x = cmul(x, make_float2((float)cos(a), (float)sin(a)));
      }
      Z[base + stride * q] = x;
    }
  }
}

template <int M>
static void check_fft()
{
  using Fc = Fact<M>;
  std::vector<float2> Z(M);
  std::vector<double> re(M), im(M);
  for (int i = 0; i < M; i++) { re[i] = drand48() - 0.5; im[i] = drand48() - 0.5; Z[i] = make_float2((float)re[i], (float)im[i]); }
  stage<M, Fc::R0, M, false>(Z);
  if constexpr (Fc::NS == 2) {
    stage<M, Fc::R1, M / Fc::R0, true>(Z);
  } else {
    stage<M, Fc::R1, M / Fc::R0, false>(Z);
    stage<M, Fc::R2, M / (Fc::R0 * Fc::R1), true>(Z);
  }
  for (int k = 0; k < M; k += 7) {
    double xr = 0, xi = 0;
    for (int n = 0; n < M; n++) {
      const double a = -2.0 * M_PI * (double)n * k / M;
      xr += re[n] * cos(a) - im[n] * sin(a);
      xi += re[n] * sin(a) + im[n] * cos(a);
    }
    const float2 g = Z[fft_pos<M>(k)];
    maxerr = fmax(maxerr, fmax(fabs(g.x - xr), fabs(g.y - xi)) / sqrt((double)M));
  }
}

int main()
{
  srand48(1);
  check_dft<2>(); check_dft<4>(); check_dft<8>(); check_dft<16>();
  printf("butterflies max err %.3g\n", maxerr);
  if (maxerr > 2e-6) return 1;
  maxerr = 0;
  check_fft<256>(); check_fft<512>(); check_fft<1024>(); check_fft<2048>();
  printf("fft max err / sqrt(M) %.3g\n", maxerr);
  return maxerr > 2e-6 ? 1 : 0;
}
