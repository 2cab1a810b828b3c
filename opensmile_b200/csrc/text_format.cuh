// text_format.cuh -- the number formats of the reference's text sinks for the device (sinks.cu) and, compiled for the host, for
// the CPU check against printf (tests/native/fmt_device_check.cpp):
//   cCsvSink / cArffSink print a value with "%.0f" when it is integer valued and with "%e" otherwise (iocore/csvSink.cpp:216-233).
// printf converts the exact binary value and rounds the decimal expansion half-to-even.  fmt_value() does the same for every finite
// float: "%e" needs 7 significant digits of v = x * 10^p (p = 6 - floor(log10 |x|)); the product is formed in double with the exact
// table 10^0 .. 10^22 (one rounding, <= 2^-53 relative, i.e. < 2e-9 absolute on a 7-digit integer), so the digits are certain unless
// the product lies within 1e-7 of a rounding boundary k + 0.5 -- then an FMA decides whether the product was exact (a true tie, as
// for dyadic values like 2^-11 = 4.8828125e-04, rounds to even) and anything else is reported as `uncertain`: the caller lets the
// host format that row (probability ~1e-7 per value).  |x| >= 1e15 on the integer path and non-finite values are also left to the
// host.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define OSM_TF_HD __host__ __device__ __forceinline__
#else
#define OSM_TF_HD inline
#endif

namespace osm {
namespace tf {

constexpr int kMaxValueChars = 17;      // "-d.dddddde-dd" = 13, "-" + 15 digits = 16

OSM_TF_HD double pow10_exact(int p)     // 10^p, exact for 0 <= p <= 22
{
  const double t[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20,
                        1e21, 1e22};
  return t[p];
}

// writes the characters of v to dst (at most kMaxValueChars), returns their number, or -1 when the host has to format this value
// alwaysE: cArffSink prints every value with "%e" (iocore/arffSink.cpp:300-312), cCsvSink only the non-integer ones
OSM_TF_HD int fmt_value(float v, char *dst, bool alwaysE = false)
{
  if (!(fabsf(v) <= 3.402823466e+38f)) return -1;                       // nan / inf
  int n = 0;
  if (signbit(v)) dst[n++] = '-';
  const float a = fabsf(v);
  if (alwaysE) {
    if (a == 0.0f) { const char z[] = "0.000000e+00"; for (int i = 0; i < 12; i++) dst[n++] = z[i]; return n; }
    if (a >= 1e7f) return -1;                                            // integers beyond the 7 printed digits: exact big-number rounding, left to the host
  }
  if (!alwaysE && a == floorf(a)) {                                      // "%.0f"
    if (a >= 1e15f) return -1;
    unsigned long long u = (unsigned long long)a;
    char tmp[16];
    int k = 0;
    do { tmp[k++] = (char)('0' + (int)(u % 10ull)); u /= 10ull; } while (u != 0ull);
    while (k > 0) dst[n++] = tmp[--k];
    return n;
  }
  // "%e": a is not an integer (a < 2^23) or, with alwaysE, below 1e7: p = 6 - E10 >= 0
  const double x = (double)a;
  int E10 = (int)floor(log10(x));
  double s;
  bool exactTable;
  for (int guard = 0; guard < 3; guard++) {
    const int p = 6 - E10;
    if (p <= 22) { s = x * pow10_exact(p); exactTable = true; }
    else if (p <= 44) { s = (x * 1e22) * pow10_exact(p - 22); exactTable = false; }
    else { s = ((x * 1e22) * 1e22) * pow10_exact(p - 44 > 22 ? 22 : p - 44); exactTable = false; if (p - 44 > 22) return -1; }
    if (s >= 1e7) E10++;
    else if (s < 1e6) E10--;
    else break;
    if (guard == 2) return -1;
  }
  double fl = floor(s);
  const double frac = s - fl;                                            // exact
  unsigned long long q = (unsigned long long)fl;
  if (fabs(frac - 0.5) < 1e-7) {
    if (frac != 0.5 || !exactTable) return -1;
    const int p = 6 - E10;
    if (fma(x, pow10_exact(p), -s) != 0.0) return -1;                    // the product was rounded: not a certain tie
    if (q & 1ull) q++;                                                   // true tie: to even
  } else if (frac > 0.5) q++;
  if (q >= 10000000ull) { q = 1000000ull; E10++; }                       // 9.9999995 -> 1.000000e+01
  char d[7];
  for (int i = 6; i >= 0; i--) { d[i] = (char)('0' + (int)(q % 10ull)); q /= 10ull; }
  dst[n++] = d[0]; dst[n++] = '.';
  for (int i = 1; i < 7; i++) dst[n++] = d[i];
  dst[n++] = 'e';
  int e = E10;
  if (e < 0) { dst[n++] = '-'; e = -e; } else dst[n++] = '+';
  dst[n++] = (char)('0' + e / 10); dst[n++] = (char)('0' + e % 10);      // |E10| <= 45 for floats
  return n;
}

}  // namespace tf
}  // namespace osm
