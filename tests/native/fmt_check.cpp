// byte comparison of the buffered formatter against fprintf over many values
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <random>
int main() {
  std::mt19937_64 g(1);
  char a[64], b[64];
  long bad = 0, n = 0;
  auto check = [&](float v) {
    n++;
    if (!std::isfinite(v)) return;
    if (v == floorf(v)) {
      if (fabsf(v) >= 1e15f) return;
      auto r = std::to_chars(a, a + 48, v, std::chars_format::fixed, 0); *r.ptr = 0;
      snprintf(b, sizeof b, "%.0f", (double)v);
    } else {
      auto r = std::to_chars(a, a + 48, v, std::chars_format::scientific, 6); *r.ptr = 0;
      snprintf(b, sizeof b, "%e", (double)v);
    }
    if (strcmp(a, b)) { if (bad < 5) printf("MISMATCH %a: '%s' vs '%s'\n", v, a, b); bad++; }
  };
  for (long i = 0; i < 20000000; i++) { uint32_t u = (uint32_t)g(); float v; memcpy(&v, &u, 4); check(v); }        // all bit patterns
  std::normal_distribution<float> nd(0.f, 1.f);
  for (long i = 0; i < 5000000; i++) { check(nd(g)); check(nd(g) * 1e-3f); check(nd(g) * 5000.f); check(floorf(nd(g) * 1000.f)); }
  const float sp[] = {0.f, -0.f, 1.f, -1.f, 0.5f, 1e-45f, 1.17549435e-38f, 3.4028235e38f, 9.9999995e14f, 999999.94f, 9.9999994e-5f, 0.99999994f, 9.999995f, 99999.95f};
  for (float v : sp) check(v);
  printf("values %ld mismatches %ld\n", n, bad);
  return bad != 0;
}
