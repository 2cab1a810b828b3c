// lld_fast.cu -- the 512-point mono MFCC instance of the fused per-frame kernel (sm_100a).
//
// Same contract, shared-memory layout, tables, chunk / tile geometry and results as lld_kernel<256,32,256,2,VEC2,MFCC>
// (kernels.cu); launch_lld() selects it when the pass is: N = 512, one channel, cMfcc on the power spectrum, no
// magnitude dump, no window offset, frameStep and frameSize multiples of 8.  What differs is the instruction stream:
//
//   stage    one code path (mono, 8 samples per thread, 16-byte loads from the bulk-copy landing zone)
//   pass 1   radix-16 butterflies straight from the sample tile; rows of the butterfly that only ever see zero padding
//            (frameSize <= 416: rows 13..15) are not loaded and their additions are pruned
//   pass 2   the second radix-16 pass, the real-FFT split and re^2 + im^2 are ONE register-resident step: butterfly t
//            produces the bins k = t (mod 16) and the split pairs bin k with bin M - k = -t (mod 16), so a warp that owns
//            butterflies t and 16 - t holds both halves of 16 pairs in registers.  The transformed tile is never written
//            back and never re-read (one Z round trip and one barrier less than lld_kernel).  The two self-paired
//            butterflies (t = 0: k and 256 - k both = 0 mod 16; t = 8) go to warp 0, which reorders its registers into the
//            same (a_q, b_{15-q}) pairing so that every warp executes the same code.
//   mel+DCT  visit list read as float4 (two bins per load); a warp adds its finished log band values straight into partial
//            DCT sums in registers, a short second step adds the eight warps' partial sums: the DCT is balanced like the
//            band work (13 coefficients do not divide over 8 warps) and the band level never goes through shared memory
//
// Reference rows as in kernels.cu (SURVEY.md 8a-1 ... a-8, a-13, a-15).
#include "lld_common.cuh"

// A/B switch: 1 = a warp adds its finished log band values straight into partial DCT sums (balanced, no band level in
// shared memory) -- 4 % faster, but the partial sums reorder the reference's sequential m = 0..25 accumulation, which moves
// 0.3 % of the delta-delta values past 1e-5 of their column's scale (measured, profiles/r02_v3_*).  Default: reference order.
#ifndef OSM_FAST_FUSED_DCT
#define OSM_FAST_FUSED_DCT 0
#endif
// further A/B switches (scripts/ab_lld512.py): unroll factors of the two passes' butterfly loops, of the mel group loop and of the
// DCT loop.  The defaults are what was measured fastest.
#ifndef OSM_FAST_P1_UNROLL
#define OSM_FAST_P1_UNROLL 1
#endif
#ifndef OSM_FAST_P2_UNROLL
#define OSM_FAST_P2_UNROLL 2
#endif
#ifndef OSM_FAST_MEL_UNROLL
#define OSM_FAST_MEL_UNROLL 1
#endif
#ifndef OSM_FAST_DCT_UNROLL
#define OSM_FAST_DCT_UNROLL 4
#endif
#ifndef OSM_FAST_STAGE_UNROLL
#define OSM_FAST_STAGE_UNROLL 1
#endif
#ifndef OSM_FAST_EMIT_K13
#define OSM_FAST_EMIT_K13 0
#endif
#define OSM_PRAGMA_(x) _Pragma(#x)
#define OSM_UNROLL(n) OSM_PRAGMA_(unroll n)

namespace osm {
namespace {

constexpr int kM = 256, kF = 32, kNT = 256, kNW = 8;
constexpr int kKMax = 16;             // static outputs per frame the in-register DCT accumulates (nStat <= 16)
constexpr int kPartOff = 9216;        // float offset of the DCT partial sums inside the FFT tile: behind P (257 x 32 floats)

// t1 -+ i t3 helper of the radix-4 butterfly whose fourth input is zero: d == 0 on entry
__device__ __forceinline__ void dft4_d0(float2 &a, float2 &b, float2 &c, float2 &d)
{
  const float2 t0 = cadd(a, c), t1 = csub(a, c), bb = b;
  a = cadd(t0, bb);
  c = csub(t0, bb);
  b = make_float2(t1.x + bb.y, t1.y - bb.x);   // t1 - i b
  d = make_float2(t1.x - bb.y, t1.y + bb.x);   // t1 + i b
}

// Dft<16>::run with rows NZR..15 known to be zero (NZR = 13 or 16)
template <int NZR>
__device__ __forceinline__ void dft16_first(float2 (&v)[16])
{
  dft4(v[0], v[4], v[8], v[12]);
  if (NZR <= 13) {
    dft4_d0(v[1], v[5], v[9], v[13]);
    dft4_d0(v[2], v[6], v[10], v[14]);
    dft4_d0(v[3], v[7], v[11], v[15]);
  } else {
    dft4(v[1], v[5], v[9], v[13]);
    dft4(v[2], v[6], v[10], v[14]);
    dft4(v[3], v[7], v[11], v[15]);
  }
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
  const float c2 = 0.70710678118654752440f;
  v[5] = cmul(v[5], make_float2(c1, -s1));
  { float2 a = v[6]; v[6] = make_float2(c2 * (a.x + a.y), c2 * (a.y - a.x)); }
  v[7] = cmul(v[7], make_float2(s1, -c1));
  { float2 a = v[9]; v[9] = make_float2(c2 * (a.x + a.y), c2 * (a.y - a.x)); }
  v[10] = cmul_mi(v[10]);
  { float2 a = v[11]; v[11] = make_float2(c2 * (a.y - a.x), -c2 * (a.x + a.y)); }
  v[13] = cmul(v[13], make_float2(s1, -c1));
  { float2 a = v[14]; v[14] = make_float2(c2 * (a.y - a.x), -c2 * (a.x + a.y)); }
  v[15] = cmul(v[15], make_float2(-c1, s1));
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}

// real-FFT split of one pair: a = Z[k], b = Z[M-k], w = exp(-2 pi i k / N)  ->  4 |X[k]|^2, 4 |X[M-k]|^2
// (same statements as lld_kernel's split; the squares are accumulated with one FMA)
__device__ __forceinline__ void split_pair(float2 a, float2 b, float2 w, float &pk, float &pm)
{
  const float2 e2 = make_float2(a.x + b.x, a.y - b.y);
  const float2 o2 = make_float2(a.x - b.x, a.y + b.y);
  const float2 t2 = cmul(o2, w);
  const float xr = e2.x + t2.y, xi = e2.y - t2.x;
  const float yr = e2.x - t2.y, yi = e2.y + t2.x;
  pk = __fmaf_rn(xr, xr, __fmul_rn(xi, xi));
  pm = __fmaf_rn(yr, yr, __fmul_rn(yi, yi));
}

template <int NZR>
__global__ void __launch_bounds__(kNT, 2) lld512_kernel(const LldParams p)
{
  constexpr int M = kM, F = kF, NT = kNT, NW = kNW;
  using D16 = Dft<16>;

  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ ChunkCtx sCx[2];
  const SmemLayout L = make_layout(p, M, F);
  float2 *Z = reinterpret_cast<float2 *>(smem + L.zbuf);
  float *P = reinterpret_cast<float *>(smem + L.zbuf);
  float *samp = reinterpret_cast<float *>(smem + L.samp);
  float *raw = reinterpret_cast<float *>(smem + L.raw);
  unsigned char *rawPcm = smem + L.rawPcm;
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + L.mbar);
  float4 *sWinLut = reinterpret_cast<float4 *>(smem + L.winLut);
  float2 *sTw = reinterpret_cast<float2 *>(smem + L.tw);
  float2 *sSplit = reinterpret_cast<float2 *>(smem + L.splitTw);
  float2 *sMelCoef = reinterpret_cast<float2 *>(smem + L.melCoef);
  int *sMelRange = reinterpret_cast<int *>(smem + L.melRange);
  float *sDct = reinterpret_cast<float *>(smem + L.dctCos);
  float *sLift = reinterpret_cast<float *>(smem + L.dctLift);
  float *ring = reinterpret_cast<float *>(smem + L.ring);
  float *Dbuf = reinterpret_cast<float *>(smem + L.zbuf);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, f = tid & 31;

  if (tid == 0) mbar_init(mbar, 1);
  for (int i = tid; i < M; i += NT) sWinLut[i] = p.winLut[i];
  for (int i = tid; i < p.twCount; i += NT) sTw[i] = p.twiddles[i];
  for (int i = tid; i < M / 2 + 1; i += NT) sSplit[i] = p.splitTw[i];
  for (int i = tid; i < p.melVCount; i += NT) sMelCoef[i] = p.melVisit[i];
  for (int i = tid; i < p.nBands + 2; i += NT) { sMelRange[i] = p.melRange[i]; sMelRange[p.nBands + 2 + i] = p.melVB[i]; }
  for (int i = tid; i < p.nBands * kKMax; i += NT) {            // transposed, zero padded: sDct[band][kKMax]
    const int m = i / kKMax, c = i - m * kKMax;
    sDct[i] = (c < p.nStat) ? p.dctCos[c * p.dctStride + m] : 0.f;
  }
  for (int i = tid; i < p.nStat; i += NT) sLift[i] = p.dctLift[i];
  for (int i = tid; i < L.sampFloats; i += NT) samp[i] = 0.f;
  __syncthreads();

  const int hop = p.frameStep;
  const int S = hop + p.sPad;
  uint32_t phase = 0;

  // The chunk context is CTA-uniform: it lives in shared memory (current / next chunk, alternating) instead of a dozen
  // registers per thread; thread 0 fills the next slot when it prefetches that chunk's first tile.
  int chunk = blockIdx.x;
  if (chunk >= p.nChunks) return;
  int cpar = 0;
  if (tid == 0) {
    sCx[0] = load_chunk<F>(p, chunk);
    const TileGeom g0 = tile_geom<F>(p, sCx[0], 0);
    mbar_expect_tx(mbar, g0.bytes);
    bulk_g2s(rawPcm, g0.src, g0.bytes, mbar);
  }
  __syncthreads();
  int j = 0;
  int emitted = sCx[0].a;

  // pass 2: butterflies of this warp; bins of the pair slots (see the file header).  Slots 0..7 hold k = wl + 16 q
  // (and M - k = 256 - wl - 16 q), slots 8..15 hold k = 256 - wh - 16 q (and M - k = wh + 16 q): warps 1..7 have
  // wl = wh = warp; warp 0 has wl = 8 (butterfly 8) and wh = 0 (butterfly 0), so every address is base + constant
  const int tA = (warp == 0) ? 0 : warp, tB = (warp == 0) ? 8 : 16 - warp;
  const int wl = (warp == 0) ? 8 : warp, wh = (warp == 0) ? 0 : warp;
  const int melBs = p.melSplit[warp], melBe = p.melSplit[warp + 1];

  while (chunk < p.nChunks) {
    const ChunkCtx &cx = sCx[cpar];

    // ================= stage: PCM (landing zone) -> float -> pre-emphasis -> sample tile =================
    mbar_wait(mbar, phase);
    phase ^= 1;
    {
      const TileGeom tg = tile_geom<F>(p, cx, j);
      const int count = tg.count;
      const int16_t *rp = reinterpret_cast<const int16_t *>(rawPcm + tg.mis) + tg.lead;
      const bool aligned = (tg.mis == 0);
      const bool hasLead = tg.lead > 0;
      const float ks = p.preDe ? p.preK : -p.preK;
      OSM_UNROLL(OSM_FAST_STAGE_UNROLL)
      for (int i = tid * 8; i < count; i += NT * 8) {
        int wds[4];
        if (aligned) {
          const int4 w4 = *reinterpret_cast<const int4 *>(rp + i);
          wds[0] = w4.x; wds[1] = w4.y; wds[2] = w4.z; wds[3] = w4.w;
        } else {
          const unsigned short *up = reinterpret_cast<const unsigned short *>(rp + i);
#pragma unroll
          for (int jj = 0; jj < 4; jj++) wds[jj] = (int)((unsigned)up[2 * jj] | ((unsigned)up[2 * jj + 1] << 16));
        }
        float x[8], y[8];
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
          x[2 * jj] = div32767((float)(short)(wds[jj] & 0xffff));
          x[2 * jj + 1] = div32767((float)(wds[jj] >> 16));
        }
        if (p.preemph) {
          // vectorPreemphasis.cpp:96-104 : x[n] -/+ k * x[n-1], two roundings
          float xprev = 0.f;
          if (i > 0 || hasLead) xprev = div32767((float)rp[i - 1]);
#pragma unroll
          for (int jj = 0; jj < 8; jj++) y[jj] = __fadd_rn(x[jj], __fmul_rn(ks, (jj == 0) ? xprev : x[jj - 1]));
        } else {
#pragma unroll
          for (int jj = 0; jj < 8; jj++) y[jj] = x[jj];
        }
        const int q = (int)__umulhi((unsigned)i, p.hopMagic);      // i / hop
        float *dst = samp + i + q * p.sPad;
        if (i == q * hop && q < F) raw[q] = x[0];                  // first sample of frame q, not pre-emphasised
#pragma unroll
        for (int jj = 0; jj < 8; jj += 2) *reinterpret_cast<float2 *>(dst + jj) = make_float2(y[jj], y[jj + 1]);
      }
    }
    __syncthreads();
    if (tid == 0) {
      if (j + 1 < cx.nT) {
        const TileGeom gn = tile_geom<F>(p, cx, j + 1);
        mbar_expect_tx(mbar, gn.bytes);
        bulk_g2s(rawPcm, gn.src, gn.bytes, mbar);
      } else if (chunk + (int)gridDim.x < p.nChunks) {
        const ChunkCtx cn = load_chunk<F>(p, chunk + gridDim.x);
        sCx[cpar ^ 1] = cn;                           // read by everyone after the barriers of this tile
        const TileGeom gn = tile_geom<F>(p, cn, 0);
        mbar_expect_tx(mbar, gn.bytes);
        bulk_g2s(rawPcm, gn.src, gn.bytes, mbar);
      }
    }

    // ================= FFT pass 1: window, radix 16, twiddles -> Z =================
    {
      const float *sampF = samp + f * S;
      const float2 *tw0 = sTw + p.twOff[0];
      OSM_UNROLL(OSM_FAST_P1_UNROLL)
      for (int t = warp; t < 16; t += NW) {
        float2 v[16];
#pragma unroll
        for (int r = 0; r < 16; r++) {
          if (r < NZR) {
            const float4 wl = sWinLut[t + 16 * r];   // (w[2e], w[2e+1], offset, #valid); padding: weight 0, offset 0
            const float2 x = *reinterpret_cast<const float2 *>(sampF + __float_as_int(wl.z));
            v[r] = make_float2(__fmul_rn(x.x, wl.x), __fmul_rn(x.y, wl.y));   // windower.cpp:226
          } else {
            v[r] = make_float2(0.f, 0.f);
          }
        }
        if (t == 0 && p.preemph)      // first sample of the frame, vectorPreemphasis.cpp:94
          v[0].x = __fmul_rn(__fmul_rn(p.oneMinusK, raw[f]), sWinLut[0].x);
        dft16_first<NZR>(v);
        const float2 *twj = tw0 + t * 16;
#pragma unroll
        for (int q = 1; q < 16; q++) v[D16::out(q)] = cmul(v[D16::out(q)], twj[q]);
        float2 *zp = Z + t * F + f;
#pragma unroll
        for (int q = 0; q < 16; q++) zp[(16 * q) * F] = v[D16::out(q)];
      }
    }
    __syncthreads();

    // ================= FFT pass 2 + real-FFT split + power, in registers =================
    {
      float2 A[16], v[16];
      OSM_UNROLL(OSM_FAST_P2_UNROLL)
      for (int h = 0; h < 2; h++) {
        const float2 *zp = Z + ((h ? tB : tA) * 16) * F + f;
#pragma unroll
        for (int r = 0; r < 16; r++) v[r] = zp[r * F];
        D16::run(v);
        if (h == 0) {
#pragma unroll
          for (int r = 0; r < 16; r++) A[r] = v[r];
        }
      }
      // X[tA + 16 q] = A[out(q)] =: a_q ; X[tB + 16 q] = v[out(q)] =: b_q ; slot q pairs a_q with b_{15-q}
      const float2 x0 = A[D16::out(0)];
      if (warp == 0) {
        // a = butterfly 0 (X[16 q]), b = butterfly 8 (X[8 + 16 q]):
        //   slots 0..7 : k = 8 + 16 q   -> (b_q, b_{15-q})                         : a'_q = b_q
        //   slots 8..15: k = 256 - 16 q -> Z[k] = a_{16-q} = b'_{15-q}, Z[M-k] = a_q : b'_j = a_{j+1}, j = 0..7
        float2 na[8], nb[8];
#pragma unroll
        for (int q = 0; q < 8; q++) { na[q] = v[D16::out(q)]; nb[q] = A[D16::out(q + 1)]; }
#pragma unroll
        for (int q = 0; q < 8; q++) { A[D16::out(q)] = na[q]; v[D16::out(q)] = nb[q]; }
      }
      float pk[16], pm[16];
      {
        const float2 *swl = sSplit + wl, *swh = sSplit + (256 - wh);
#pragma unroll
        for (int q = 0; q < 16; q++) {
          const float2 aq = A[D16::out(q)], bq = v[D16::out(15 - q)];
          if (q < 8) split_pair(aq, bq, swl[16 * q], pk[q], pm[q]);        // aq = Z[k], bq = Z[M-k], k = wl + 16 q
          else       split_pair(bq, aq, swh[-16 * q], pk[q], pm[q]);       // bq = Z[k], aq = Z[M-k], k = 256 - wh - 16 q
        }
      }
      __syncthreads();   // every warp has read its part of Z before P (aliasing Z) is written
      {
        float *Plo = P + wl * F + f, *Plm = P + (256 - wl) * F + f;
        float *Phi = P + (256 - wh) * F + f, *Phm = P + wh * F + f;
#pragma unroll
        for (int q = 0; q < 8; q++) { Plo[(16 * q) * F] = pk[q]; Plm[(-16 * q) * F] = pm[q]; }
#pragma unroll
        for (int q = 8; q < 16; q++) {
          Phi[(-16 * q) * F] = pk[q];
          if (q > 8 || warp != 0) Phm[(16 * q) * F] = pm[q];               // warp 0, slot 8: k = 128 = M - k
        }
      }
      if (warp == 0) {
        // k = 0: a = b = Z[0], w = 1 -> 2 X[0] = 2 (re + im), 2 X[M] = 2 (re - im), both real
        const float xr = 2.0f * (x0.x + x0.y), yr = 2.0f * (x0.x - x0.y);
        P[f] = __fmul_rn(xr, xr);
        P[M * F + f] = __fmul_rn(yr, yr);
      }
    }
    __syncthreads();

#if OSM_FAST_FUSED_DCT
    // ================= mel filterbank (melspec.cpp:543-569) + log (mfcc.cpp:239-243) + DCT-II partial sums =================
    // Every warp owns a contiguous band group (cost-balanced on the host).  A finished log band value goes straight into
    // the warp's partial DCT sums (mfcc.cpp:251-272: sum_m log[m] cos(...), here grouped by warp: bands ascending inside a
    // warp, warps ascending in the final sum), so the DCT work is balanced like the band work and the band level never
    // touches shared memory.
    {
      float acc[kKMax];
#pragma unroll
      for (int c = 0; c < kKMax; c++) acc[c] = 0.f;
      if (melBs < melBe) {
        const int *sVB = sMelRange + p.nBands + 2;
        float cur = 0.f;
        for (int r = melBs; r <= melBe; r++) {
          float nxt = 0.f;
          const float *pp = P + sMelRange[r] * F + f;
          const int v0 = sVB[r];
          const float4 *cp = reinterpret_cast<const float4 *>(sMelCoef + v0);
          OSM_UNROLL(OSM_FAST_MEL_UNROLL)
          for (int q = (sVB[r + 1] - v0) >> 2; q > 0; q--, pp += 4 * F, cp += 2) {
            const float p0 = pp[0], p1 = pp[F], p2 = pp[2 * F], p3 = pp[3 * F];
            const float4 wa = cp[0], wb = cp[1];
            cur = __fmaf_rn(p0, wa.x, cur); nxt = __fmaf_rn(p0, wa.y, nxt);
            cur = __fmaf_rn(p1, wa.z, cur); nxt = __fmaf_rn(p1, wa.w, nxt);
            cur = __fmaf_rn(p2, wb.x, cur); nxt = __fmaf_rn(p2, wb.y, nxt);
            cur = __fmaf_rn(p3, wb.z, cur); nxt = __fmaf_rn(p3, wb.w, nxt);
          }
          if (r > melBs) {
            float mval = __fmul_rn(cur, p.melScale);
            if (p.doLog) mval = (mval < p.melfloor) ? p.logMelfloor : logf(mval);
            const float4 *dt = reinterpret_cast<const float4 *>(sDct + (r - 1) * kKMax);
#pragma unroll
            for (int c4 = 0; c4 < kKMax / 4; c4++) {
              const float4 d = dt[c4];
              acc[4 * c4 + 0] = __fmaf_rn(mval, d.x, acc[4 * c4 + 0]);
              acc[4 * c4 + 1] = __fmaf_rn(mval, d.y, acc[4 * c4 + 1]);
              acc[4 * c4 + 2] = __fmaf_rn(mval, d.z, acc[4 * c4 + 2]);
              acc[4 * c4 + 3] = __fmaf_rn(mval, d.w, acc[4 * c4 + 3]);
            }
          }
          cur = nxt;
        }
      }
      float *part = P + kPartOff + warp * (kKMax * F) + f;      // [warp][kKMax][F], behind the power spectrum
#pragma unroll
      for (int c = 0; c < kKMax; c++) part[c * F] = acc[c];
    }
    __syncthreads();

    // ================= DCT-II: sum of the warps' partial sums, lifter (mfcc.cpp:268-272) =================
    const int ringBase = (j & 1) * F;
    for (int c = warp; c < p.nStat; c += NW) {
      const float *pp = P + kPartOff + c * F + f;
      float a = pp[0];
#pragma unroll
      for (int w = 1; w < NW; w++) a = __fadd_rn(a, pp[w * (kKMax * F)]);
      ring[c * (2 * F) + ringBase + f] = __fmul_rn(a, sLift[c]);
    }
    __syncthreads();

#else
    // ================= mel filterbank (melspec.cpp:543-569) + log (mfcc.cpp:239-243) =================
    // the band level lives behind the power spectrum (P + kPartOff), not in the sample tile
    float *melS = P + kPartOff;
    if (melBs < melBe) {
      const int *sVB = sMelRange + p.nBands + 2;
      float cur = 0.f;
      for (int r = melBs; r <= melBe; r++) {
        float nxt = 0.f;
        const float *pp = P + sMelRange[r] * F + f;
        const int v0 = sVB[r];
        const float4 *cp = reinterpret_cast<const float4 *>(sMelCoef + v0);
        OSM_UNROLL(OSM_FAST_MEL_UNROLL)
        for (int q = (sVB[r + 1] - v0) >> 2; q > 0; q--, pp += 4 * F, cp += 2) {
          const float p0 = pp[0], p1 = pp[F], p2 = pp[2 * F], p3 = pp[3 * F];
          const float4 wa = cp[0], wb = cp[1];
          cur = __fmaf_rn(p0, wa.x, cur); nxt = __fmaf_rn(p0, wa.y, nxt);
          cur = __fmaf_rn(p1, wa.z, cur); nxt = __fmaf_rn(p1, wa.w, nxt);
          cur = __fmaf_rn(p2, wb.x, cur); nxt = __fmaf_rn(p2, wb.y, nxt);
          cur = __fmaf_rn(p3, wb.z, cur); nxt = __fmaf_rn(p3, wb.w, nxt);
        }
        if (r > melBs) {
          float mval = __fmul_rn(cur, p.melScale);
          if (p.doLog) mval = (mval < p.melfloor) ? p.logMelfloor : logf(mval);
          melS[(r - 1) * F + f] = mval;
        }
        cur = nxt;
      }
    }
    __syncthreads();

    // ================= DCT-II + lifter (mfcc.cpp:251-272), the reference's m = 0 .. nBands-1 accumulation order =================
    // table transposed and zero padded, sDct[band][kKMax]: warp w evaluates coefficients 2w and 2w+1 together (one 8-byte
    // table read and one band value feed two dot products); the padding columns are zero
    const int ringBase = (j & 1) * F;
    {
      const int i = 2 * warp;
      if (i < p.nStat) {
        const float2 *cc = reinterpret_cast<const float2 *>(sDct + i);
        const float *lp = melS + f;
        float a0 = 0.f, a1 = 0.f;
        OSM_UNROLL(OSM_FAST_DCT_UNROLL)
        for (int m = 0; m < p.nBands; m++, lp += F, cc += kKMax / 2) {
          const float l0 = lp[0];
          const float2 c = cc[0];
          a0 = __fmaf_rn(l0, c.x, a0); a1 = __fmaf_rn(l0, c.y, a1);
        }
        ring[i * (2 * F) + ringBase + f] = __fmul_rn(a0, sLift[i]);
        if (i + 1 < p.nStat) ring[(i + 1) * (2 * F) + ringBase + f] = __fmul_rn(a1, sLift[i + 1]);
      }
    }
    __syncthreads();

#endif
    // ================= store (same statements as lld_kernel) =================
    const int tfs = cx.s0 + j * F;                    // first static frame of this tile
    if (!p.fused) {
      const int tot = min(F, cx.sEnd - tfs) * p.nStat;
      for (int idx = tid; idx < tot; idx += NT) {
        const int ff = idx / p.nStat, c = idx - ff * p.nStat;
        p.out[(cx.row0 + tfs + ff) * p.outStride + p.outCol + c] = ring[c * (2 * F) + ringBase + ff];
      }
    } else {
      const int K = p.nStat, W1 = p.fW1, W2 = p.fW2, H = W1 + W2;
      const int T = cx.T;
      const int r0 = emitted;
      const int r1 = (j + 1 == cx.nT) ? cx.b : min(tfs + F - H, cx.b);
      const int T1 = T + W1, c01 = max(T - W1, 0), c02 = max(c01 - W2, 0);
      const float norm1 = p.fNorm1, norm2 = p.fNorm2;
      const int d0 = max(r0 - W2, 0), d1 = min(r1 + W2, T1);
      const int dRows = F + 24;
      float *outS = Dbuf + ((K * dRows + 3) & ~3);
      const int K3 = 3 * K;
      const int nr = r1 - r0;
      const bool interior1 = (d0 >= W1) && (d1 + W1 <= T);
      const bool interior2 = (r0 >= W2) && (r1 <= c02);
      if (interior1 && interior2 && W1 == 2 && W2 == 2 && nr == F) {
#if OSM_FAST_EMIT_K13
        if (K == 13) emit_interior<F, NT, 13>(ring, Dbuf, outS, K, dRows, d0 - cx.s0, r0 - cx.s0, norm1, p.fRcp1, norm2, p.fRcp2, tid);
        else
#endif
        emit_interior<F, NT, 0>(ring, Dbuf, outS, K, dRows, d0 - cx.s0, r0 - cx.s0, norm1, p.fRcp1, norm2, p.fRcp2, tid);
      } else {
        emit_edge<F, NW>(ring, Dbuf, outS, K, W1, W2, T, T1, c01, c02, cx.s0, r0, r1, d0, d1, dRows, norm1, p.fRcp1, norm2, p.fRcp2, warp, f);
      }
      {
        float *o = p.out + (cx.row0 + r0) * (long long)K3;
        const int n = nr * K3;
        for (int i = tid; i < n; i += NT) o[i] = outS[i];
      }
      emitted = r1;
    }

    j++;
    if (j == cx.nT) {
      chunk += gridDim.x;
      j = 0;
      cpar ^= 1;
      if (chunk < p.nChunks) emitted = sCx[cpar].a;
    }
  }
}

template <int NZR>
cudaError_t launch_fast_t(const LldParams &p, int numSMs, cudaStream_t st, LldLaunchInfo *info)
{
  const size_t smem = (size_t)make_layout(p, kM, kF).total;
  auto kern = lld512_kernel<NZR>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kNT, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorLaunchOutOfResources;
  int grid = numSMs * occ;
  if (grid > p.nChunks) grid = p.nChunks;
  if (grid < 1) grid = 1;
  if (info) { info->grid = grid; info->block = kNT; info->smem = smem; }
  kern<<<grid, kNT, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

bool lld_fast_applies(const LldParams &p, int nfft)
{
  return nfft == 512 && !p.narrow && p.opKind == 0 && p.magOut == nullptr && p.melUsePower && p.nChan == 1 &&
         p.frameStep % 8 == 0 && p.frameSize % 8 == 0 && p.frameSize <= 512 && !p.hasWinOffset && p.nStat <= kKMax &&
         ((p.frameStep + p.sPad) % 2) == 0;
}

cudaError_t launch_lld_fast(const LldParams &p, int numSMs, cudaStream_t st, LldLaunchInfo *info)
{
  if (p.frameSize <= 416) return launch_fast_t<13>(p, numSMs, st, info);
  return launch_fast_t<16>(p, numSMs, st, info);
}

}  // namespace osm
