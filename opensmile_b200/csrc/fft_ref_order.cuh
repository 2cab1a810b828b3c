// fft_ref_order.cuh -- a 512-point real FFT whose every float rounding equals the reference's transform
// (cTransformFFT -> rdft, Ooura's split-radix package as vendored in /root/reference/src/dspcore/fftsg.c), written for
// the formant branch of the GeMAPS graphs: order-11 LPC in float amplifies a 2e-7 difference of the spectrum into
// 1e-3 .. 1e-2 of the formant frequencies, so that ONE consumer needs the reference's rounding sequence, not just
// its values.  Everything else in this library uses the faster register-radix FFT (fft_radix.cuh).
//
// What is reproduced (citations relative to /root/reference/src/dspcore/fftsg.c):
//   * the table of twiddle factors w[] / c[] (makewt :660-718, makect :741-757) -- built on the host by
//     tables.cpp:build_ref_fft_tables with the same float / double expression types;
//   * the decomposition of the 256-point complex transform behind rdft(512): one radix-4 pass over the whole array
//     whose odd twiddles are interpolated from the even table entries (cftf1st :1801-2005), one radix-4 pass per
//     quarter in two flavours -- plain (cftmdl1 :2441-2548) and "rotated input" (cftmdl2 :2551-2682) -- and sixteen
//     16-point transforms in two flavours (cftf161 :2706-2862, cftf162 :2865-3045), in the order cftleaf lists them for
//     n = 512 (:2383-2407); bit reversal (bitrv2 :913) and the real-transform post pass (rftfsub :3241-3263, rdft :347-349).
// What is NOT taken over: the data layout and the work decomposition.  Here a transform is a set of independent WORK
// ITEMS per phase (one radix-4 butterfly, one 16-point leaf, one conjugate bin pair); the kernel hands the items of a
// phase to the threads of a CTA (item index fastest, so neighbouring lanes touch neighbouring elements), with a barrier
// between phases.  Real and imaginary parts live in two planes, padded by one float per 16 so that both the
// unit-stride butterflies and the stride-16 leaves are free of bank conflicts.
//
// Rounding discipline: only complex add / sub, the four twiddle product forms below and negation occur, each written
// exactly once with explicit single roundings (no FMA contraction; host build: -ffp-contract=off).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define OSM_RO_HD __host__ __device__ __forceinline__
#else
#define OSM_RO_HD inline
#endif

namespace osm {
namespace ro {

constexpr int kN = 512;                 // real transform length
constexpr int kC = kN / 2;              // complex points
constexpr int kPlane = kC + kC / 16;    // floats per padded plane
constexpr int kNw = kN / 4, kNc = kN / 4;   // table sizes (floats): w[kNw] followed by c[kNc]

OSM_RO_HD int phys(int c) { return c + (c >> 4); }

#if defined(__CUDA_ARCH__)
OSM_RO_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
OSM_RO_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
OSM_RO_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
#else
OSM_RO_HD float fadd(float a, float b) { return a + b; }
OSM_RO_HD float fsub(float a, float b) { return a - b; }
OSM_RO_HD float fmul(float a, float b) { return a * b; }
#endif

struct cf { float r, i; };
OSM_RO_HD cf mk(float r, float i) { cf z; z.r = r; z.i = i; return z; }
OSM_RO_HD cf add(cf a, cf b) { return mk(fadd(a.r, b.r), fadd(a.i, b.i)); }
OSM_RO_HD cf sub(cf a, cf b) { return mk(fsub(a.r, b.r), fsub(a.i, b.i)); }
OSM_RO_HD cf addj(cf a, cf b) { return mk(fsub(a.r, b.i), fadd(a.i, b.r)); }      // a + i b
OSM_RO_HD cf subj(cf a, cf b) { return mk(fadd(a.r, b.i), fsub(a.i, b.r)); }      // a - i b
// x (A + iB) and x (A - iB) with the products rounded before the sum
OSM_RO_HD cf mulp(float A, float B, cf x) { return mk(fsub(fmul(A, x.r), fmul(B, x.i)), fadd(fmul(A, x.i), fmul(B, x.r))); }
OSM_RO_HD cf muln(float A, float B, cf x) { return mk(fadd(fmul(A, x.r), fmul(B, x.i)), fsub(fmul(A, x.i), fmul(B, x.r))); }
// x w (1 + i) and x w (1 - i): the sum is rounded before the product
OSM_RO_HD cf rotp(float w, cf x) { return mk(fmul(w, fsub(x.r, x.i)), fmul(w, fadd(x.i, x.r))); }
OSM_RO_HD cf rotn(float w, cf x) { return mk(fmul(w, fadd(x.r, x.i)), fmul(w, fsub(x.i, x.r))); }

struct Planes {                          // one frame: padded real / imaginary planes
  float *re, *im;
  OSM_RO_HD cf ld(int c) const { const int p = phys(c); return mk(re[p], im[p]); }
  OSM_RO_HD void st(int c, cf z) const { const int p = phys(c); re[p] = z.r; im[p] = z.i; }
};

// ---- radix-4 butterflies ---------------------------------------------------------------------------------------
// plain flavour: sums / differences of (a0, a2) and (a1, a3), outputs 2 and 3 carry the twiddles
struct Quad { cf s02, d02, s13, d13; };
OSM_RO_HD Quad quad_plain(cf a0, cf a1, cf a2, cf a3) { Quad q; q.s02 = add(a0, a2); q.d02 = sub(a0, a2); q.s13 = add(a1, a3); q.d13 = sub(a1, a3); return q; }

// first pass over the whole array (cftf1st): butterfly q of M = n/8 on elements q, q+M, q+2M, q+3M.
// Twiddles of even q are table entries; odd q interpolate their two even neighbours (csc1, csc3 = w[2], w[3]);
// q > M/2 mirror M-q with real and imaginary coefficient swapped; q = 0 and q = M/2 are the multiplication-free cases.
OSM_RO_HD void first_pass_item(const Planes &a, const float *w, int q)
{
  constexpr int M = kN / 8, H = M / 2;
  const cf a0 = a.ld(q), a1 = a.ld(q + M), a2 = a.ld(q + 2 * M), a3 = a.ld(q + 3 * M);
  const Quad u = quad_plain(a0, a1, a2, a3);
  a.st(q, add(u.s02, u.s13));
  a.st(q + M, sub(u.s02, u.s13));
  const cf t = addj(u.d02, u.d13), v = subj(u.d02, u.d13);
  if (q == 0) { a.st(2 * M, t); a.st(3 * M, v); return; }
  const float wn4r = w[1];
  if (q == H) {
    a.st(q + 2 * M, rotp(wn4r, t));
    a.st(q + 3 * M, rotn(-wn4r, v));
    return;
  }
  const int qq = q < H ? q : M - q;
  float c1r, c1i, c3r, c3i;
  if ((qq & 1) == 0) {
    const int k = 2 * qq;
    c1r = w[k]; c1i = w[k + 1]; c3r = w[k + 2]; c3i = w[k + 3];
  } else {
    const int k = 2 * (qq + 1);
    float p1r = 1.0f, p1i = 0.0f, p3r = 1.0f, p3i = 0.0f;
    if (k > 4) { p1r = w[k - 4]; p1i = w[k - 3]; p3r = w[k - 2]; p3i = w[k - 1]; }
    float n1r, n1i, n3r, n3i;
    if (qq == H - 1) { n1r = wn4r; n1i = wn4r; n3r = -wn4r; n3i = -wn4r; }
    else { n1r = w[k]; n1i = w[k + 1]; n3r = w[k + 2]; n3i = w[k + 3]; }
    const float csc1 = w[2], csc3 = w[3];
    c1r = fmul(csc1, fadd(p1r, n1r)); c1i = fmul(csc1, fadd(p1i, n1i));
    c3r = fmul(csc3, fadd(p3r, n3r)); c3i = fmul(csc3, fadd(p3i, n3i));
  }
  if (q < H) { a.st(q + 2 * M, mulp(c1r, c1i, t)); a.st(q + 3 * M, muln(c3r, c3i, v)); }
  else       { a.st(q + 2 * M, mulp(c1i, c1r, t)); a.st(q + 3 * M, muln(c3i, c3r, v)); }
}

// second pass, plain flavour (cftmdl1) on a block of 4M elements starting at `base`; w = table of this level
OSM_RO_HD void mid_plain_item(const Planes &a, int base, int M, const float *w, int q)
{
  const int H = M / 2;
  const cf a0 = a.ld(base + q), a1 = a.ld(base + q + M), a2 = a.ld(base + q + 2 * M), a3 = a.ld(base + q + 3 * M);
  const Quad u = quad_plain(a0, a1, a2, a3);
  a.st(base + q, add(u.s02, u.s13));
  a.st(base + q + M, sub(u.s02, u.s13));
  const cf t = addj(u.d02, u.d13), v = subj(u.d02, u.d13);
  if (q == 0) { a.st(base + 2 * M, t); a.st(base + 3 * M, v); return; }
  if (q == H) { a.st(base + q + 2 * M, rotp(w[1], t)); a.st(base + q + 3 * M, rotn(-w[1], v)); return; }
  const int k = 4 * (q < H ? q : M - q);
  if (q < H) { a.st(base + q + 2 * M, mulp(w[k], w[k + 1], t)); a.st(base + q + 3 * M, muln(w[k + 2], w[k + 3], v)); }
  else       { a.st(base + q + 2 * M, mulp(w[k + 1], w[k], t)); a.st(base + q + 3 * M, muln(w[k + 3], w[k + 2], v)); }
}

// second pass, rotated flavour (cftmdl2): inputs combine as a0 +- i a2, a1 +- i a3 and all four outputs are twiddled
OSM_RO_HD void mid_rot_item(const Planes &a, int base, int M, const float *w, int q)
{
  const int H = M / 2;
  const cf a0 = a.ld(base + q), a1 = a.ld(base + q + M), a2 = a.ld(base + q + 2 * M), a3 = a.ld(base + q + 3 * M);
  const cf x0 = addj(a0, a2), x1 = subj(a0, a2), x2 = addj(a1, a3), x3 = subj(a1, a3);
  if (q == 0) {
    const cf y = rotp(w[1], x2);
    a.st(base, add(x0, y));
    a.st(base + M, sub(x0, y));
    const cf z = rotp(w[1], x3);
    a.st(base + 2 * M, addj(x1, z));
    a.st(base + 3 * M, subj(x1, z));
    return;
  }
  if (q == H) {
    const float kr = w[2 * M], ki = w[2 * M + 1];
    const cf y0 = mulp(kr, ki, x0), y2 = mulp(ki, kr, x2);
    a.st(base + q, add(y0, y2));
    a.st(base + q + M, sub(y0, y2));
    const cf z0 = mulp(ki, kr, x1), z2 = mulp(kr, ki, x3);
    a.st(base + q + 2 * M, sub(z0, z2));
    a.st(base + q + 3 * M, add(z0, z2));
    return;
  }
  const int qq = q < H ? q : M - q;
  const int k = 4 * qq, kr = 4 * M - 4 * qq;
  const float k1r = w[k], k1i = w[k + 1], k3r = w[k + 2], k3i = w[k + 3];
  const float d1i = w[kr], d1r = w[kr + 1], d3i = w[kr + 2], d3r = w[kr + 3];
  cf y0, y2, z0, z2;
  if (q < H) { y0 = mulp(k1r, k1i, x0); y2 = mulp(d1r, d1i, x2); z0 = muln(k3r, k3i, x1); z2 = muln(d3r, d3i, x3); }
  else       { y0 = mulp(d1i, d1r, x0); y2 = mulp(k1i, k1r, x2); z0 = muln(d3i, d3r, x1); z2 = muln(k3i, k3r, x3); }
  a.st(base + q, add(y0, y2));
  a.st(base + q + M, sub(y0, y2));
  a.st(base + q + 2 * M, add(z0, z2));
  a.st(base + q + 3 * M, sub(z0, z2));
}

// ---- 16-point leaves: four radix-4 butterflies over stride 4, then four over stride 1 ----------------------------
// plain flavour (cftf161); w = table of this level: w[1] = cos(pi/4), (w[2], w[3]) = e^{i pi/8}
OSM_RO_HD void leaf_plain(const Planes &a, int base, const float *w)
{
  const float h = w[1], er = w[2], ei = w[3];
  cf y[16];
  {  // column 0: no twiddles
    const Quad u = quad_plain(a.ld(base + 0), a.ld(base + 4), a.ld(base + 8), a.ld(base + 12));
    y[0] = add(u.s02, u.s13); y[4] = sub(u.s02, u.s13); y[8] = addj(u.d02, u.d13); y[12] = subj(u.d02, u.d13);
  }
  {  // column 1: e^{i pi/8} and its mirror
    const Quad u = quad_plain(a.ld(base + 1), a.ld(base + 5), a.ld(base + 9), a.ld(base + 13));
    y[1] = add(u.s02, u.s13); y[5] = sub(u.s02, u.s13);
    y[9] = mulp(er, ei, addj(u.d02, u.d13)); y[13] = mulp(ei, er, subj(u.d02, u.d13));
  }
  {  // column 2: (1 +- i) / sqrt 2
    const Quad u = quad_plain(a.ld(base + 2), a.ld(base + 6), a.ld(base + 10), a.ld(base + 14));
    y[2] = add(u.s02, u.s13); y[6] = sub(u.s02, u.s13);
    y[10] = rotp(h, addj(u.d02, u.d13)); y[14] = rotn(h, subj(u.d02, u.d13));
  }
  {  // column 3
    const Quad u = quad_plain(a.ld(base + 3), a.ld(base + 7), a.ld(base + 11), a.ld(base + 15));
    y[3] = add(u.s02, u.s13); y[7] = sub(u.s02, u.s13);
    y[11] = mulp(ei, er, addj(u.d02, u.d13)); y[15] = mulp(er, ei, subj(u.d02, u.d13));
  }
  {  // rows, written back in the package's order of evaluation (the values do not depend on it)
    const cf p0 = sub(y[12], y[14]), p1 = add(y[12], y[14]), p2 = sub(y[13], y[15]), p3 = add(y[13], y[15]);
    a.st(base + 12, add(p0, p2)); a.st(base + 13, sub(p0, p2)); a.st(base + 14, addj(p1, p3)); a.st(base + 15, subj(p1, p3));
  }
  {
    const cf p0 = add(y[8], y[10]), p1 = sub(y[8], y[10]), p2 = add(y[9], y[11]), p3 = sub(y[9], y[11]);
    a.st(base + 8, add(p0, p2)); a.st(base + 9, sub(p0, p2)); a.st(base + 10, addj(p1, p3)); a.st(base + 11, subj(p1, p3));
  }
  {
    const cf p2 = rotp(h, addj(y[5], y[7])), p3 = rotp(h, subj(y[5], y[7]));
    const cf p0 = addj(y[4], y[6]), p1 = subj(y[4], y[6]);
    a.st(base + 4, add(p0, p2)); a.st(base + 5, sub(p0, p2)); a.st(base + 6, addj(p1, p3)); a.st(base + 7, subj(p1, p3));
  }
  {
    const cf p0 = add(y[0], y[2]), p1 = sub(y[0], y[2]), p2 = add(y[1], y[3]), p3 = sub(y[1], y[3]);
    a.st(base + 0, add(p0, p2)); a.st(base + 1, sub(p0, p2)); a.st(base + 2, addj(p1, p3)); a.st(base + 3, subj(p1, p3));
  }
}

// rotated flavour (cftf162); w = table of this level: w[1] = cos(pi/4), (w[4], w[5]) = e^{i pi/16}, (w[6], -w[7]) = e^{i 3pi/16},
// (w[8], w[9]) = e^{i pi/8}
OSM_RO_HD void leaf_rot(const Planes &a, int base, const float *w)
{
  const float h = w[1], k1r = w[4], k1i = w[5], k3r = w[6], k3i = -w[7], k2r = w[8], k2i = w[9];
  cf y[16];
  {
    const cf a0 = a.ld(base + 0), a8 = a.ld(base + 8), a4 = a.ld(base + 4), a12 = a.ld(base + 12);
    cf x1 = addj(a0, a8), x2 = rotp(h, addj(a4, a12));
    y[0] = add(x1, x2); y[4] = sub(x1, x2);
    x1 = subj(a0, a8); x2 = rotp(h, subj(a4, a12));
    y[8] = addj(x1, x2); y[12] = subj(x1, x2);
  }
  {
    const cf a1 = a.ld(base + 1), a9 = a.ld(base + 9), a5 = a.ld(base + 5), a13 = a.ld(base + 13);
    cf x1 = mulp(k1r, k1i, addj(a1, a9)), x2 = mulp(k3i, k3r, addj(a5, a13));
    y[1] = add(x1, x2); y[5] = sub(x1, x2);
    x1 = mulp(k3r, k3i, subj(a1, a9)); x2 = muln(k1r, k1i, subj(a5, a13));
    y[9] = sub(x1, x2); y[13] = add(x1, x2);
  }
  {
    const cf a2 = a.ld(base + 2), a10 = a.ld(base + 10), a6 = a.ld(base + 6), a14 = a.ld(base + 14);
    cf x1 = mulp(k2r, k2i, addj(a2, a10)), x2 = mulp(k2i, k2r, addj(a6, a14));
    y[2] = add(x1, x2); y[6] = sub(x1, x2);
    x1 = mulp(k2i, k2r, subj(a2, a10)); x2 = mulp(k2r, k2i, subj(a6, a14));
    y[10] = sub(x1, x2); y[14] = add(x1, x2);
  }
  {
    const cf a3 = a.ld(base + 3), a11 = a.ld(base + 11), a7 = a.ld(base + 7), a15 = a.ld(base + 15);
    cf x1 = mulp(k3r, k3i, addj(a3, a11)), x2 = mulp(k1i, k1r, addj(a7, a15));
    y[3] = add(x1, x2); y[7] = sub(x1, x2);
    x1 = muln(k1i, k1r, subj(a3, a11)); x2 = mulp(k3i, k3r, subj(a7, a15));
    y[11] = add(x1, x2); y[15] = sub(x1, x2);
  }
  {
    const cf p1 = add(y[0], y[2]), p2 = add(y[1], y[3]);
    a.st(base + 0, add(p1, p2)); a.st(base + 1, sub(p1, p2));
    const cf q1 = sub(y[0], y[2]), q2 = sub(y[1], y[3]);
    a.st(base + 2, addj(q1, q2)); a.st(base + 3, subj(q1, q2));
  }
  {
    cf p1 = addj(y[4], y[6]), p2 = rotp(h, addj(y[5], y[7]));
    a.st(base + 4, add(p1, p2)); a.st(base + 5, sub(p1, p2));
    p1 = subj(y[4], y[6]); p2 = rotp(h, subj(y[5], y[7]));
    a.st(base + 6, addj(p1, p2)); a.st(base + 7, subj(p1, p2));
  }
  {
    const cf p1 = add(y[8], y[10]), p2 = sub(y[9], y[11]);
    a.st(base + 8, add(p1, p2)); a.st(base + 9, sub(p1, p2));
    const cf q1 = sub(y[8], y[10]), q2 = add(y[9], y[11]);
    a.st(base + 10, addj(q1, q2)); a.st(base + 11, subj(q1, q2));
  }
  {
    cf p1 = addj(y[12], y[14]), p2 = rotp(h, subj(y[13], y[15]));
    a.st(base + 12, add(p1, p2)); a.st(base + 13, sub(p1, p2));
    p1 = subj(y[12], y[14]); p2 = rotp(h, addj(y[13], y[15]));
    a.st(base + 14, addj(p1, p2)); a.st(base + 15, subj(p1, p2));
  }
}

// ---- phases of the 512-point transform ---------------------------------------------------------------------------
constexpr int kItemsA = kN / 8;        // 64 butterflies of the first pass
constexpr int kItemsB = kN / 8;        // 4 quarters x 16 butterflies
constexpr int kItemsC = kC / 16;       // 16 leaves
constexpr int kItemsD = kC / 2 + 1;    // bin 0 (DC / Nyquist) and the conjugate pairs 1 .. 128 (128 = the self-paired bin)

OSM_RO_HD void phase_a(const Planes &a, const float *w, int item) { first_pass_item(a, w + kNw - kN / 4, item); }

// cftleaf(512, isplt = 1): quarters 0, 2, 3 plain, quarter 1 rotated
OSM_RO_HD void phase_b(const Planes &a, const float *w, int item)
{
  const int blk = item >> 4, q = item & 15;
  if (blk == 1) mid_rot_item(a, blk * 64, 16, w + kNw - 128, q);
  else mid_plain_item(a, blk * 64, 16, w + kNw - 64, q);
}

OSM_RO_HD void phase_c(const Planes &a, const float *w, int item)
{
  // leaves 1, 5, 7, 9, 13 are of the rotated flavour
  const unsigned rotMask = (1u << 1) | (1u << 5) | (1u << 7) | (1u << 9) | (1u << 13);
  if ((rotMask >> item) & 1u) leaf_rot(a, item * 16, w + kNw - 32);
  else leaf_plain(a, item * 16, w + kNw - 8);
}

OSM_RO_HD int bitrev8(int x)
{
  x = ((x & 0x0f) << 4) | ((x & 0xf0) >> 4);
  x = ((x & 0x33) << 2) | ((x & 0xcc) >> 2);
  x = ((x & 0x55) << 1) | ((x & 0xaa) >> 1);
  return x;
}

// bit reversal + real post pass, out of place: src = transform output in bit-reversed order, dst = bins in natural
// order, packed like the reference packs them: bin 0 = (Re X0, Re X256), bin p = X_p.  c = w + kNw.
OSM_RO_HD void phase_d(const Planes &src, const Planes &dst, const float *c, int item)
{
  if (item == 0) {
    const cf z = src.ld(0);
    dst.st(0, mk(fadd(z.r, z.i), fsub(z.r, z.i)));
    return;
  }
  const int p = item, k = kC - p;
  const cf zp = src.ld(bitrev8(p));
  if (p == k) { dst.st(p, zp); return; }
  const cf zk = src.ld(bitrev8(k));
  const float wkr = fsub(0.5f, c[kNc - p]), wki = c[p];
  const cf x = mk(fsub(zp.r, zk.r), fadd(zp.i, zk.i));
  const cf y = mulp(wkr, wki, x);
  dst.st(p, mk(fsub(zp.r, y.r), fsub(zp.i, y.i)));
  dst.st(k, mk(fadd(zk.r, y.r), fsub(zk.i, y.i)));
}

}  // namespace ro
}  // namespace osm
