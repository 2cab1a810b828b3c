// kernels.cu -- fused per-frame LLD kernels for sm_100a.
//
// Design: one persistent CTA processes "tiles" of F consecutive frames of one
// utterance.  Inside a tile every thread keeps the mapping  lane -> frame  for ALL phases:
//
//   stage   PCM (int16, HBM, coalesced 16-byte loads) -> float -> pre-emphasis -> smem
//   FFT     real FFT as an M = N/2 point complex FFT, in-place decimation-in-frequency with
//           register radix-8/16 butterflies; the data tile lives in shared memory as
//           Z[element][frame], so every warp-wide access is conflict free and every table
//           (window, twiddles, mel weights, DCT) is warp-uniform (broadcast)
//   split   real-FFT post-processing + |X|^2  -> P[bin][frame]
//   mel     two-tap triangular filterbank, sequential in the bin index exactly like the
//           reference's loop (lldcore/melspec.cpp:543-553) -> bit-faithful summation order
//   dct     log, DCT-II, lifter (lldcore/mfcc.cpp:238-273), again in the reference's order
//   store   rows of the output level
//
// The temporal regression stages (cDeltaRegression / cContourSmoother) run in a second,
// memory-bound kernel (post_kernel) with the reference's edge / phantom-frame semantics.
//
// Arithmetic that the reference performs in a fixed float order (conversion, pre-emphasis,
// window, power, mel, log, DCT, lifter, delta) uses explicit non-fused __fmul_rn/__fadd_rn so
// that, given identical inputs, results are bit-identical to the x86-64 reference build
// (which has no FMA contraction).  Only the FFT itself uses FMA freely.
#include <cstdio>

#include "fft_radix.cuh"
#include "kernels.cuh"

namespace osm {

// ------------------------------------------------------------------------------------------
// shared memory layout (identical computation on host and device)
// ------------------------------------------------------------------------------------------
struct SmemLayout {
  int zbuf, samp, raw, winPairs, sampLut, tw, splitTw, melCoef, melRange, dctCos, dctLift, melS, mfccS;
  int total;
  int sampFloats;
};

__host__ __device__ inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

__host__ __device__ inline SmemLayout make_layout(const LldParams &p, int M, int F)
{
  SmemLayout L;
  int o = 0;
  L.zbuf = o; o += M * F * 8;
  const int S = p.frameStep + p.sPad;
  L.sampFloats = align_up((F - 1) * S + p.frameSize + ((p.frameSize - 1) / p.frameStep) * p.sPad + 2, 4);
  L.samp = o; o += L.sampFloats * 4;
  L.raw = o; o += F * 4;
  o = align_up(o, 16);
  L.winPairs = o; o += M * 8;
  L.sampLut = o; o += M * 4;
  o = align_up(o, 16);
  L.tw = o; o += p.twCount * 8;
  L.splitTw = o; o += (M / 2 + 1) * 8;
  o = align_up(o, 16);
  L.melCoef = o; o += (M + 1) * 4;
  L.melRange = o; o += (p.nBands + 2) * 4;
  o = align_up(o, 16);
  L.dctCos = o; o += p.nMfcc * p.nBands * 4;
  L.dctLift = o; o += p.nMfcc * 4;
  o = align_up(o, 16);
  L.melS = o; o += p.nBands * F * 4;
  L.mfccS = o; o += p.nMfcc * F * 4;
  L.total = align_up(o, 16);
  return L;
}

// ------------------------------------------------------------------------------------------
// PCM conversion, smileutil/smileUtil.c:2520-2534 : ((sum_c (float)x_c) / nChan) / 32767
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float pcm_to_float(const int16_t *s, int nChan)
{
  float tmp = (float)s[0];
  for (int c = 1; c < nChan; c++) tmp = __fadd_rn(tmp, (float)s[c]);
  return __fdiv_rn(__fdiv_rn(tmp, (float)nChan), 32767.0f);
}
__device__ __forceinline__ float pcm1_to_float(int v) { return __fdiv_rn((float)v, 32767.0f); }

// ------------------------------------------------------------------------------------------
// one in-place DIF stage.  Virtual warp vw (of NVW) handles butterflies t = vw, vw+NVW, ...
// ------------------------------------------------------------------------------------------
template <int M, int F, int NVW, int R, int MS, bool FIRST, bool LAST>
__device__ __forceinline__ void fft_stage(float2 *__restrict__ Z, const float *__restrict__ sampF,
                                          const float *__restrict__ raw,
                                          const float2 *__restrict__ winPairs,
                                          const int *__restrict__ lut, const float2 *__restrict__ tw,
                                          const LldParams &p, int vw, int f)
{
  constexpr int stride = MS / R;
  for (int t = vw; t < M / R; t += NVW) {
    const int blk = t / stride, j = t % stride;
    const int base = blk * MS + j;
    float2 v[R];
    if (FIRST) {
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int e = base + stride * r;
        const int n = 2 * e;
        float2 x = make_float2(0.f, 0.f);
        if (n < p.frameSize) {          // warp-uniform
          const int off = lut[e];
          x.x = sampF[off];
          if (n + 1 < p.frameSize) x.y = sampF[off + 1];
          if (e == 0 && p.preemph) x.x = __fmul_rn(p.oneMinusK, raw[f]);   // vectorPreemphasis.cpp:94
          const float2 w = winPairs[e];
          // windower.cpp:226 : src * (float)w + (float)offset (two roundings)
          x.x = __fmul_rn(x.x, w.x);
          x.y = __fmul_rn(x.y, w.y);
          if (p.hasWinOffset) {
            x.x = __fadd_rn(x.x, p.winOffset);
            if (n + 1 < p.frameSize) x.y = __fadd_rn(x.y, p.winOffset);
          }
        }
        v[r] = x;
      }
    } else {
#pragma unroll
      for (int r = 0; r < R; r++) v[r] = Z[(base + stride * r) * F + f];
    }
    Dft<R>::run(v);
    if (!LAST) {
      const float2 *twj = tw + j * R;
#pragma unroll
      for (int q = 1; q < R; q++) v[Dft<R>::out(q)] = cmul(v[Dft<R>::out(q)], twj[q]);
    }
#pragma unroll
    for (int q = 0; q < R; q++) Z[(base + stride * q) * F + f] = v[Dft<R>::out(q)];
  }
}

// ------------------------------------------------------------------------------------------
// the fused kernel
// ------------------------------------------------------------------------------------------
template <int M, int F, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) lld_kernel(const LldParams p)
{
  constexpr int NW = NT / 32, G = 32 / F, NVW = NW * G;
  constexpr int NBINS = M + 1;
  constexpr int NPAIR = M / 2 + 1;                 // pairs (k, M-k), k = 0..M/2
  constexpr int PAIRS_PER_VW = (NPAIR + NVW - 1) / NVW;
  using Fc = Fact<M>;

  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = make_layout(p, M, F);
  float2 *Z = reinterpret_cast<float2 *>(smem + L.zbuf);
  float *P = reinterpret_cast<float *>(smem + L.zbuf);   // aliases Z (used after the split)
  float *samp = reinterpret_cast<float *>(smem + L.samp);
  float *raw = reinterpret_cast<float *>(smem + L.raw);
  float2 *sWin = reinterpret_cast<float2 *>(smem + L.winPairs);
  int *sLut = reinterpret_cast<int *>(smem + L.sampLut);
  float2 *sTw = reinterpret_cast<float2 *>(smem + L.tw);
  float2 *sSplit = reinterpret_cast<float2 *>(smem + L.splitTw);
  float *sMelCoef = reinterpret_cast<float *>(smem + L.melCoef);
  int *sMelRange = reinterpret_cast<int *>(smem + L.melRange);
  float *sDct = reinterpret_cast<float *>(smem + L.dctCos);
  float *sLift = reinterpret_cast<float *>(smem + L.dctLift);
  float *melS = reinterpret_cast<float *>(smem + L.melS);
  float *mfccS = reinterpret_cast<float *>(smem + L.mfccS);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int f = lane & (F - 1);
  const int vw = warp * G + lane / F;

  // ---- constant tables -> smem (once per CTA) ----
  for (int i = tid; i < M; i += NT) { sWin[i] = p.winPairs[i]; sLut[i] = p.sampLut[i]; }
  for (int i = tid; i < p.twCount; i += NT) sTw[i] = p.twiddles[i];
  for (int i = tid; i < NPAIR; i += NT) sSplit[i] = p.splitTw[i];
  for (int i = tid; i < NBINS; i += NT) sMelCoef[i] = p.melCoef[i];
  for (int i = tid; i < p.nBands + 2; i += NT) sMelRange[i] = p.melRange[i];
  for (int i = tid; i < p.nMfcc * p.nBands; i += NT) sDct[i] = p.dctCos[i];
  for (int i = tid; i < p.nMfcc; i += NT) sLift[i] = p.dctLift[i];
  __syncthreads();

  const int hop = p.frameStep, size = p.frameSize, nChan = p.nChan;
  const int S = hop + p.sPad;

  for (int tile = blockIdx.x; tile < p.nTiles; tile += gridDim.x) {
    const TileRef tr = p.tiles[tile];
    const long long uo = p.uttOff[tr.utt];
    const long long Ls = p.uttOff[tr.utt + 1] - uo;
    const int T = (int)((Ls - size) / hop + 1);
    const int nf = min(F, T - tr.f0);
    const long long s0 = (long long)tr.f0 * hop;
    const int count = (nf - 1) * hop + size;
    const int16_t *src = p.pcm + (uo + s0) * nChan;

    // ================= stage: PCM -> float -> pre-emphasis -> smem =================
    {
      const bool vecOk = (nChan == 1) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
      for (int c = tid; c * 8 < count; c += NT) {
        const int i = c * 8;
        float x[8];
        float xprev = 0.f;
        const int nvalid = min(8, count - i);
        if (vecOk && nvalid == 8) {
          const int4 raw4 = __ldg(reinterpret_cast<const int4 *>(src + i));
          const int wds[4] = {raw4.x, raw4.y, raw4.z, raw4.w};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            x[2 * j] = pcm1_to_float((int)(short)(wds[j] & 0xffff));
            x[2 * j + 1] = pcm1_to_float(wds[j] >> 16);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; j++) x[j] = (j < nvalid) ? pcm_to_float(src + (long long)(i + j) * nChan, nChan) : 0.f;
        }
        if (p.preemph && (s0 + i) > 0) xprev = pcm_to_float(src + (long long)(i - 1) * nChan, nChan);
        int q = i / hop;
        int r = i - q * hop;
        float *dst = samp + i + q * p.sPad;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (j < nvalid) {
            float y = x[j];
            if (p.preemph) {
              // vectorPreemphasis.cpp:96-104 : x[n] -/+ k * x[n-1], two roundings
              const float kx = __fmul_rn(p.preK, (j == 0) ? xprev : x[j - 1]);
              y = p.preDe ? __fadd_rn(x[j], kx) : __fsub_rn(x[j], kx);
            }
            if (r == 0) {
              if (q < F) raw[q] = x[j];
            }
            dst[j] = y;
            r++;
            if (r == hop) { r = 0; q++; dst += p.sPad; }
          }
        }
      }
    }
    __syncthreads();

    // ================= FFT =================
    {
      const float *sampF = samp + f * S;
      fft_stage<M, F, NVW, Fc::R0, M, true, false>(Z, sampF, raw, sWin, sLut, sTw + p.twOff[0], p, vw, f);
      __syncthreads();
      if constexpr (Fc::NS == 2) {
        fft_stage<M, F, NVW, Fc::R1, M / Fc::R0, false, true>(Z, nullptr, nullptr, nullptr, nullptr, nullptr, p, vw, f);
      } else {
        fft_stage<M, F, NVW, Fc::R1, M / Fc::R0, false, false>(Z, nullptr, nullptr, nullptr, nullptr, sTw + p.twOff[1], p, vw, f);
        __syncthreads();
        fft_stage<M, F, NVW, Fc::R2, M / (Fc::R0 * Fc::R1), false, true>(Z, nullptr, nullptr, nullptr, nullptr, nullptr, p, vw, f);
      }
      __syncthreads();
    }

    // ================= real-FFT split + power spectrum =================
    // X[k] = E - i W^k O,  X[M-k] = conj(E + i W^k O),  E = (Z[k]+conj(Z[M-k]))/2, O = (Z[k]-conj(Z[M-k]))/2
    {
      float pk[PAIRS_PER_VW], pm[PAIRS_PER_VW];
#pragma unroll
      for (int i = 0; i < PAIRS_PER_VW; i++) {
        const int k = vw + i * NVW;
        pk[i] = 0.f; pm[i] = 0.f;
        if (k < NPAIR) {
          const float2 a = Z[fft_pos<M>(k) * F + f];
          const float2 b = Z[fft_pos<M>((M - k) & (M - 1)) * F + f];   // Z[M] == Z[0]
          const float2 w = sSplit[k];
          const float2 e2 = make_float2(a.x + b.x, a.y - b.y);         // 2E
          const float2 o2 = make_float2(a.x - b.x, a.y + b.y);         // 2O
          const float2 t2 = cmul(o2, w);                               // 2 W^k O
          // 2 X[k] = e2 - i t2 ; 2 conj(X[M-k]) = e2 + i t2
          const float xr = 0.5f * (e2.x + t2.y), xi = 0.5f * (e2.y - t2.x);
          const float yr = 0.5f * (e2.x - t2.y), yi = 0.5f * (e2.y + t2.x);
          // fftmagphase.cpp:215-221 computes sqrt(re*re+im*im), melspec.cpp:524 squares it
          // again; we keep re*re+im*im (<= 1.5 ulp apart, below the FFT's own noise floor)
          if (p.melUsePower) {
            pk[i] = __fadd_rn(__fmul_rn(xr, xr), __fmul_rn(xi, xi));
            pm[i] = __fadd_rn(__fmul_rn(yr, yr), __fmul_rn(yi, yi));
          } else {
            pk[i] = __fsqrt_rn(__fadd_rn(__fmul_rn(xr, xr), __fmul_rn(xi, xi)));
            pm[i] = __fsqrt_rn(__fadd_rn(__fmul_rn(yr, yr), __fmul_rn(yi, yi)));
          }
        }
      }
      __syncthreads();   // all Z reads done before P (aliasing Z) is written
#pragma unroll
      for (int i = 0; i < PAIRS_PER_VW; i++) {
        const int k = vw + i * NVW;
        if (k < NPAIR) {
          P[k * F + f] = pk[i];
          if (k != M - k) P[(M - k) * F + f] = pm[i];
        }
      }
    }
    __syncthreads();

    // ================= mel filterbank (melspec.cpp:543-569) + log (mfcc.cpp:239-243) =================
    {
      const int bs = p.melSplit[vw], be = p.melSplit[vw + 1];
      if (bs < be) {
        float cur = 0.f, nxt = 0.f;
        for (int r = bs; r <= be; r++) {
          const int n0 = sMelRange[r], n1 = sMelRange[r + 1];
          for (int n = n0; n < n1; n++) {
            const float pw = P[n * F + f];
            const float a = __fmul_rn(pw, sMelCoef[n]);    // (float)((double)p*(double)w) == fl(p*w)
            if (r > bs) cur = __fadd_rn(cur, a);
            nxt = __fadd_rn(nxt, __fsub_rn(pw, a));
          }
          if (r > bs) {
            float mval = __fmul_rn(cur, p.melScale);
            if (p.doLog) mval = (mval < p.melfloor) ? p.logMelfloor : logf(mval);
            melS[(r - 1) * F + f] = mval;
          }
          cur = nxt;
          nxt = 0.f;
        }
      }
    }
    __syncthreads();

    // ================= DCT-II + lifter (mfcc.cpp:251-272) =================
    for (int i = vw; i < p.nMfcc; i += NVW) {
      const float *ct = sDct + i * p.nBands;
      float acc = 0.f;
      for (int m = 0; m < p.nBands; m++) acc = __fadd_rn(acc, __fmul_rn(melS[m * F + f], ct[m]));
      mfccS[i * F + f] = __fmul_rn(acc, sLift[i]);
    }
    __syncthreads();

    // ================= store =================
    {
      const long long row0 = p.rowOff[tr.utt] + tr.f0;
      const int tot = nf * p.nMfcc;
      for (int idx = tid; idx < tot; idx += NT) {
        const int ff = idx / p.nMfcc, c = idx - ff * p.nMfcc;
        p.out[(row0 + ff) * p.outStride + p.outCol + c] = mfccS[c * F + ff];
      }
    }
    // no barrier needed here: the next tile's staging only writes samp/raw, which no thread
    // reads after the FFT's first stage; the barrier after staging orders everything else.
  }
}

// ------------------------------------------------------------------------------------------
// temporal post-processing
// ------------------------------------------------------------------------------------------
struct PostCtx {
  const float *base;     // static rows of this utterance
  int stride;
  int T;                 // static frames
};

// Tick-order model of chained window processors (cWindowProcessor with blocksize=1, components
// ticking in data-flow order; core/windowProcessor.cpp:85-119,167-230, core/componentManager.cpp:
// 1233-1262).  For the level produced by stage s:  final_s = final_{s-1} + W_s frames in total,
// c0_s = max(c0_{s-1} - W_s, 0) of them produced before EOI is raised (c0_0 = final_0 = T).
// When the consumer computes frame t >= c0_s its input level holds
//     navail = min(c0_{s-1} + (t - c0_s) + 1, final_{s-1})
// frames.  Matrix reads (core/dataMemoryLevel.cpp:1651-1738): window start >= 0: rows >= navail
// replicate row navail-1; window start < 0: rows < 0 replicate row 0 and rows >= navail come
// from the zero-initialised, not yet written level buffer (0.0).  For c0_{s-1} >= W_s this is the
// plain "clamp to [0, final-1]" rule; the rest only triggers for utterances shorter than the
// summed window lengths and is reproduced because the reference does it.
template <int LVL>
__device__ float post_eval(const PostCtx &cx, const PostGroup &g, int t, int c);

template <>
__device__ __forceinline__ float post_eval<0>(const PostCtx &cx, const PostGroup &g, int t, int c)
{
  return cx.base[(long long)t * cx.stride + c];
}

template <int LVL>
__device__ __forceinline__ float post_read(const PostCtx &cx, const PostGroup &g, int t, int W, int navail, int i, int c)
{
  if (t - W < 0) {
    if (i < 0) return post_eval<LVL - 1>(cx, g, 0, c);
    if (i >= navail) return 0.f;
    return post_eval<LVL - 1>(cx, g, i, c);
  }
  return post_eval<LVL - 1>(cx, g, min(i, navail - 1), c);
}

template <int LVL>
__device__ float post_eval(const PostCtx &cx, const PostGroup &g, int t, int c)
{
  int Tprev = cx.T, n0 = cx.T;          // input level: final frame count, frames before EOI
#pragma unroll
  for (int i = 0; i < LVL - 1; i++) { Tprev += g.win[i]; n0 = max(n0 - g.win[i], 0); }
  const int W = g.win[LVL - 1];
  const int c0 = max(n0 - W, 0);
  const int navail = (t < c0) ? Tprev : min(n0 + (t - c0) + 1, Tprev);
  if (g.kind[LVL - 1] == 0) {
    // deltaRegression.cpp:139-146 : num = sum_i i*(x[t+i]-x[t-i]) ; y = num / (2 sum i^2)
    float norm = 0.f;
    for (int i = 1; i <= W; i++) norm = __fadd_rn(norm, __fmul_rn((float)i, (float)i));
    norm = __fmul_rn(norm, 2.0f);
    float num = 0.f;
    for (int i = 1; i <= W; i++) {
      const float later = post_read<LVL>(cx, g, t, W, navail, t + i, c);
      const float prior = post_read<LVL>(cx, g, t, W, navail, t - i, c);
      num = __fadd_rn(num, __fmul_rn((float)i, __fsub_rn(later, prior)));
    }
    return __fdiv_rn(num, norm);
  } else {
    // contourSmoother.cpp:84-117 : y = x[n]; y += x[n-w]; y += x[n+w]; y /= smaWin
    const int noZero = g.flags[LVL - 1];
    const float x0 = post_read<LVL>(cx, g, t, W, navail, t, c);
    if (noZero && x0 == 0.f) return 0.f;
    float y = x0;
    int cnt = 1;
    for (int w = 1; w <= W; w++) {
      const float a = post_read<LVL>(cx, g, t, W, navail, t - w, c);
      const float b = post_read<LVL>(cx, g, t, W, navail, t + w, c);
      if (!noZero || a != 0.f) { y = __fadd_rn(y, a); cnt++; }
      if (!noZero || b != 0.f) { y = __fadd_rn(y, b); cnt++; }
    }
    return __fdiv_rn(y, noZero ? (float)cnt : (float)(2 * W + 1));
  }
}

__global__ void __launch_bounds__(256) post_kernel(const PostParams p)
{
  // one thread per (output row, column of a staged group); rows are found by binary search
  // over the per-utterance row offsets.
  int colsTotal = 0;
  for (int g = 0; g < p.nGroups; g++) colsTotal += p.groups[g].n;
  const long long total = p.totalRows * colsTotal;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / colsTotal;
    int col = (int)(idx - row * colsTotal);
    int gi = 0;
    while (col >= p.groups[gi].n) { col -= p.groups[gi].n; gi++; }
    const PostGroup &g = p.groups[gi];
    // utterance of this row
    int lo = 0, hi = p.nUtt;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (p.rowOff[mid] <= row) lo = mid; else hi = mid;
    }
    const int u = lo;
    const int t = (int)(row - p.rowOff[u]);
    const long long Ls = p.uttOff[u + 1] - p.uttOff[u];
    PostCtx cx;
    cx.T = (int)((Ls - p.frameSize) / p.frameStep + 1);
    cx.stride = p.statStride;
    cx.base = p.stat + p.statOff[u] * (long long)p.statStride + g.srcCol;
    float v;
    switch (g.nStages) {
      case 1: v = post_eval<1>(cx, g, t, col); break;
      case 2: v = post_eval<2>(cx, g, t, col); break;
      case 3: v = post_eval<3>(cx, g, t, col); break;
      default: v = post_eval<0>(cx, g, min(t, cx.T - 1), col); break;
    }
    p.out[row * p.outStride + g.outCol + col] = v;
  }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
int lld_tile_frames(int nfft) { return nfft == 2048 ? 16 : 32; }
int lld_virtual_warps(int nfft) { return nfft == 512 ? 8 : (nfft == 1024 ? 16 : 32); }
bool lld_supported_fft(int nfft) { return nfft == 512 || nfft == 1024 || nfft == 2048; }

size_t lld_smem_bytes(const LldParams &p, int nfft)
{
  return (size_t)make_layout(p, nfft / 2, lld_tile_frames(nfft)).total;
}

template <int M, int F, int NT, int MINB>
static cudaError_t launch_t(const LldParams &p, int numSMs, cudaStream_t st, LldLaunchInfo *info)
{
  const size_t smem = (size_t)make_layout(p, M, F).total;
  auto kern = lld_kernel<M, F, NT, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorLaunchOutOfResources;
  int grid = numSMs * occ;
  if (grid > p.nTiles) grid = p.nTiles;
  if (grid < 1) grid = 1;
  if (info) { info->grid = grid; info->block = NT; info->smem = smem; }
  kern<<<grid, NT, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_lld(const LldParams &p, int nfft, int numSMs, cudaStream_t st, LldLaunchInfo *info)
{
  switch (nfft) {
    case 512:  return launch_t<256, 32, 256, 2>(p, numSMs, st, info);
    case 1024: return launch_t<512, 32, 512, 1>(p, numSMs, st, info);
    case 2048: return launch_t<1024, 16, 512, 1>(p, numSMs, st, info);
    default:   return cudaErrorInvalidValue;
  }
}

cudaError_t launch_post(const PostParams &p, cudaStream_t st)
{
  int colsTotal = 0;
  for (int g = 0; g < p.nGroups; g++) colsTotal += p.groups[g].n;
  const long long total = p.totalRows * colsTotal;
  if (total <= 0) return cudaSuccess;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  post_kernel<<<(int)blocks, 256, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace osm
