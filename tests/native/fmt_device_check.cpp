// Host check of the device text formatter (opensmile_b200/csrc/text_format.cuh) against printf, the reference's formatter
// (iocore/csvSink.cpp:216-233: "%.0f" for integer-valued values, "%e" otherwise).  Prints: checked mismatches uncertain
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../opensmile_b200/csrc/text_format.cuh"

static long bad = 0, unc = 0, tot = 0;
static void check(float v)
{
  char a[64], b[64];
  const int n = osm::tf::fmt_value(v, a);
  tot++;
  if (n < 0) { if (fabsf(v) < 1e15f) { unc++; if (unc <= 5) fprintf(stderr, "uncertain %a = %.9e\n", v, (double)v); } return; }
  a[n] = 0;
  if (v == floorf(v)) snprintf(b, sizeof b, "%.0f", (double)v); else snprintf(b, sizeof b, "%e", (double)v);
  if (strcmp(a, b) != 0) { if (bad < 10) fprintf(stderr, "mismatch %a: got %s want %s\n", v, a, b); bad++; }
  // cArffSink's mode: "%e" for every value (integers included, below 1e7; the rest is the host's)
  const int m = osm::tf::fmt_value(v, a, true);
  if (m < 0) { if (fabsf(v) < 1e7f) { unc++; if (unc <= 5) fprintf(stderr, "uncertain (always %%e) %a = %.9e\n", v, (double)v); } return; }
  a[m] = 0;
  snprintf(b, sizeof b, "%e", (double)v);
  if (strcmp(a, b) != 0) { if (bad < 10) fprintf(stderr, "mismatch (always %%e) %a: got %s want %s\n", v, a, b); bad++; }
}

int main(int argc, char **argv)
{
  const long N = argc > 1 ? atol(argv[1]) : 20000000;
  unsigned long long st = 88172645463325252ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
  for (long i = 0; i < N; i++) {                      // random bit patterns (all exponents, both signs)
    const uint32_t u = (uint32_t)rnd();
    float v; memcpy(&v, &u, 4);
    if (v != v || fabsf(v) > 3.4e38f) continue;
    check(v);
  }
  for (long i = 0; i < N / 4; i++) {                  // values of LLD magnitude, 7-digit decimals and their float neighbours
    const double d = (double)(rnd() % 20000000ull) / 1e6 * ((rnd() & 1) ? 1.0 : -1.0) * pow(10.0, (int)(rnd() % 14) - 9);
    float v = (float)d;
    check(v); check(nextafterf(v, 1e30f)); check(nextafterf(v, -1e30f));
  }
  for (int e = -149; e <= 30; e++)                    // dyadic values: exact ties of the 7-digit rounding
    for (int m = 1; m < 4096; m += 2) { check(ldexpf((float)m, e)); check(-ldexpf((float)m, e)); }
  for (int e = -44; e <= 6; e++)                      // around the powers of ten
    for (int k = -3; k <= 3; k++) { float v = (float)pow(10.0, e); for (int j = 0; j < (k < 0 ? -k : k); j++) v = nextafterf(v, k < 0 ? 0.0f : 1e30f); check(v); }
  for (long i = -70000; i <= 70000; i++) { check((float)i); check((float)i * 0.5f); check((float)i / 32767.0f); }
  check(0.0f); check(-0.0f); check(1e14f); check(-9.9999995e6f); check(8388607.5f); check(1.17549435e-38f); check(1.4e-45f);
  printf("%ld %ld %ld\n", tot, bad, unc);
  return bad != 0;
}
