// fft_ref_order_host.cpp -- host build of opensmile_b200/csrc/fft_ref_order.cuh + the product's table builder (test
// infrastructure).  The CUDA kernel hands the work items of a phase to threads; here they are loops (in a scrambled order,
// to show that the result does not depend on the order inside a phase).  tests/test_fft_ref_order_cpu.py compares the output
// bit for bit with the reference's own rdft (oracle/_ref/libfftsg.so).
//   g++ -O2 -ffp-contract=off -shared -fPIC -o fft_ref_order_host.so fft_ref_order_host.cpp ../../opensmile_b200/csrc/tables.cpp
#include "../../opensmile_b200/csrc/fft_ref_order.cuh"
#include <vector>

namespace osm { void build_ref_fft_tables(std::vector<float> &wc); }

extern "C" {

// x[512] (real input) -> out[512] packed like the reference packs it (a[0] = Re X0, a[1] = Re X256, a[2k], a[2k+1] = X_k)
void roh_rdft512(const float *x, float *out, int scramble)
{
  using namespace osm::ro;
  static std::vector<float> wc;
  if (wc.empty()) osm::build_ref_fft_tables(wc);
  std::vector<float> buf(4 * kPlane, 0.0f);
  Planes a{buf.data(), buf.data() + kPlane}, b{buf.data() + 2 * kPlane, buf.data() + 3 * kPlane};
  for (int c = 0; c < kC; c++) a.st(c, mk(x[2 * c], x[2 * c + 1]));
  auto order = [&](int n, int i) { return scramble ? (int)(((long long)i * 37 + 11) % n) : i; };   // 37 is coprime to 64, 16, 129
  for (int i = 0; i < kItemsA; i++) phase_a(a, wc.data(), order(kItemsA, i));
  for (int i = 0; i < kItemsB; i++) phase_b(a, wc.data(), order(kItemsB, i));
  for (int i = 0; i < kItemsC; i++) phase_c(a, wc.data(), order(kItemsC, i));
  for (int i = 0; i < kItemsD; i++) phase_d(a, b, wc.data() + kNw, order(kItemsD, i));
  for (int c = 0; c < kC; c++) { const cf z = b.ld(c); out[2 * c] = z.r; out[2 * c + 1] = z.i; }
}

void roh_tables(float *wc) {
  std::vector<float> t;
  osm::build_ref_fft_tables(t);
  for (size_t i = 0; i < t.size(); i++) wc[i] = t[i];
}

}
