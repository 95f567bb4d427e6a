// gl_conv.cuh -- PSF convolution fused with ss x ss mean pooling, the pixel likelihood, and the
// adjoint of all three (reference: src/gigalens/tf/simulator.py:142-156, tf/model.py:91-101).
//
// conv2d(SAME, stride 1, K x K) followed by avg_pool(ss) is ONE stride-ss correlation with the
// folded kernel Keff = box_ss * K of side E = K + ss - 1 (3.7x fewer MACs at K=25, ss=2).  Written
// in polyphase form that is ss^2 stride-1 correlations with A x A taps, A = ceil(E/ss):
//
//   out[oy][ox] = sum_{py,px} sum_{a,b<A} W[py][px][a][b] * I_{py,px}[oy+a][ox+b],
//   I_{py,px}[r][c] = S[ss*r + py - pad][ss*c + px - pad]   (zero outside the image),
//
// and the adjoint is, per phase, the same stride-1 correlation of the zero-padded dL/d(out) with
// the flipped taps.
//
// HBM layout of the supersampled image: PHASE-MAJOR, ss[b][py'][px'][n][n] with
// S[ss*ri + py'][ss*cj + px'] at ((py'*ss + px')*n + ri)*n + cj.  The ray-tracing kernels are
// oblivious (their coordinate tables are permuted once at plan creation), and both conv kernels move
// contiguous rows: the forward kernel streams one phase sub-image at a time through a two-stage
// cp.async pipeline (load of phase k+1 overlaps the FMAs of phase k; 2 x one phase tile of shared
// memory instead of ss^2 tiles => 4 CTAs/SM instead of 2), the adjoint writes each phase contiguously.  Both directions share corr_rows(): each thread owns an RY x RX register tile
// of outputs, walks the RY+A-1 input rows of its strip once (RX+A-1 floats per row, LDS.128) and
// applies each row to every output row it reaches; tap weights are broadcast LDS.128.  In the
// steady state ~90% of issued instructions are FFMA (FP32-FMA bound; north_star forbids tensor
// cores here), and HBM sees each ss pixel once per direction.
#pragma once
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)
#include <cuda_runtime.h>

#define GLC_RX 4
#define GLC_RY 6
#define GLC_MAX_COL_ITERS 10   // staged tile rows are at most 320 floats wide

// 4-byte asynchronous global->shared copy (LDGSTS); src_bytes = 0 zero-fills (used for the SAME padding).
__device__ __forceinline__ void glc_cp_async4(float* smem_dst, const float* gsrc, int src_bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sa), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void glc_cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

struct GlConvGeom {
  int n;        // image side (pooled)
  int hs;       // n * ss
  int ss;
  int A;        // taps per phase (template value actually used; weights zero-padded up to it)
  int pad;      // SAME padding of the K x K kernel: (K-1)/2
  int ntx, nty; // thread tiles per CTA tile in x / y
  int tpr;      // threads per thread-row (ntx rounded up so that LDS.128 quarter-warps do not straddle rows)
  int tw, th;   // CTA tile size in outputs: ntx*RX, nty*RY
  int tiles_x, tiles_y;
  int in_rows, in_pitch;   // smem input tile: rows = th + A - 1, pitch = roundup4(tw + A - 1)
  int wpitch;   // roundup4(A)
  int tma_pitch;     // forward, TMA staging: row pitch of the staged tile (in_pitch + 4)
  int out_stride;    // adjoint, TMA staging: floats of the output staging tile (tw*th rounded up to 32)
  int band_rows;     // adjoint, TMA staging: > 0 = every warp owns band_rows complete tile rows and stores them itself
  int band_stride;   //   floats between the warps' staging bands (band_rows*tw rounded up to 32)
  int phase_stride;  // forward only: floats between phase sub-images (in_rows*in_pitch padded so that the
                     // de-interleaving stores of one warp fall into distinct banks)
  int rc0;      // adjoint only: first padded-phase row/column (pad / ss)
};

// ---- register-tile correlation with packed fp32 FMA (FFMA2, sm_100+) ---------------------------
// acc2[rp][c] holds the output pair (row 2rp, row 2rp+1) of column c.  One input row of the thread's
// strip (RX+A-1 floats, LDS.128) sits `row` below the strip origin and reaches output row r through
// tap row a = row - r.  For tap column b the two taps of a pair are (w[row-2rp][b], w[row-2rp-1][b]):
// consecutive entries of the per-column REVERSED, zero-padded tap array
//     U[b][j] = w[A-1-(j-(RY-1))][b]   (0 outside),   j0(row) = A + RY - 2 - row,
// so the pair for rp is (U[b][j0+2rp], U[b][j0+2rp+1]) -- one broadcast LDS.64.  Two copies of U
// (shifted by one float) keep that load 8-byte aligned for both parities of j0.  The FMA is
//     FFMA2 acc2, strip[c+b].F32 (scalar broadcast), taps.F32x2, acc2
// i.e. two FMAs per issue slot; out-of-range taps are zeros, and whole pairs that a ramp-up /
// ramp-down row cannot reach are skipped with a compile-time mask.
#define GLC_RP (GLC_RY / 2)
__host__ __device__ constexpr int glc_ulen(int A) { return (A + 2 * GLC_RY - 2 + 3) & ~3; }   // floats per U copy
// taps table: [phase][b][copy(2)][ulen]

// Where the tap pairs come from.
//   GlTapsSmem : the packed table staged in shared memory; a pair is one broadcast LDS.64 (one LSU wavefront per 4 FFMA2 --
//                with the strip loads that put the shared-memory pipe at ~90 % of the FMA pipe's cycles, the co-limiter ncu showed).
//   GlTapsConst: the same table passed BY VALUE as a __grid_constant__ kernel parameter (constant bank 0).  The index is warp-uniform,
//                so the compiler loads a pair with LDCU.64 into a uniform register pair and FFMA2 takes that pair as its
//                operand (`FFMA2 R, R.F32, UR.F32x2, R`): tap traffic leaves the LSU entirely and needs no shared memory.
//                Used when the table fits the parameter space (NF * 4 <= GLC_CONST_TAP_BYTES).
#define GLC_CONST_TAP_BYTES (24 * 1024)
template <int NF> struct alignas(16) GlTapsC { float w[NF > 0 ? NF : 4]; };
struct GlTapsSmem {
  const float* u;
  __device__ __forceinline__ float2 pair(int off) const { return *reinterpret_cast<const float2*>(u + off); }
};
template <int NF> struct GlTapsConst {
  const GlTapsC<NF>& T;
  int base;
  __device__ __forceinline__ float2 pair(int off) const { return *reinterpret_cast<const float2*>(&T.w[base + off]); }
};

// LW = floats per strip load: 4 (LDS.128, strip origin 16-byte aligned) or 2 (LDS.64, 8-byte aligned origin -- the
// TMA-staged tile, whose first column is rounded down to the unit's 16-byte coordinate granularity).
template <int A, int MASK, int LW, class Taps>
__device__ __forceinline__ void corr_row2(const float* __restrict__ inrow, const Taps& taps, int uoff /* offset of U[0][copy][j0e] */,
                                          float2 (&acc2)[GLC_RP][GLC_RX]) {
  constexpr int SL = (GLC_RX + A - 1 + 3) & ~3;
  constexpr int UL = glc_ulen(A);
  float strip[SL];
  if constexpr (LW == 4) {
    const float4* src = reinterpret_cast<const float4*>(inrow);
#pragma unroll
    for (int v = 0; v < SL / 4; ++v) {
      float4 t = src[v];
      strip[4 * v] = t.x; strip[4 * v + 1] = t.y; strip[4 * v + 2] = t.z; strip[4 * v + 3] = t.w;
    }
  } else {
    const float2* src = reinterpret_cast<const float2*>(inrow);
#pragma unroll
    for (int v = 0; v < SL / 2; ++v) {
      float2 t = src[v];
      strip[2 * v] = t.x; strip[2 * v + 1] = t.y;
    }
  }
#pragma unroll
  for (int b = 0; b < A; ++b) {
#pragma unroll
    for (int rp = 0; rp < GLC_RP; ++rp) {
      if ((MASK >> rp) & 1) {
        const float2 w2 = taps.pair(uoff + b * 2 * UL + 2 * rp);
#pragma unroll
        for (int c = 0; c < GLC_RX; ++c) {
          const float sv = strip[c + b];
          acc2[rp][c] = __ffma2_rn(make_float2(sv, sv), w2, acc2[rp][c]);
        }
      }
    }
  }
}

// the pairs a ramp-up / ramp-down row reaches form a contiguous range [LO, HI]: one instance per range, picked by a uniform if-chain
template <int A, int LW, class Taps, int LO = 0, int HI = 0>
__device__ __forceinline__ void corr_row2_ramp(int lo_p, int hi_p, const float* __restrict__ inrow, const Taps& taps, int uoff,
                                               float2 (&acc2)[GLC_RP][GLC_RX]) {
  if (lo_p == LO && hi_p == HI) corr_row2<A, ((1 << (HI + 1)) - 1) & ~((1 << LO) - 1), LW>(inrow, taps, uoff, acc2);
  else if constexpr (HI + 1 < GLC_RP) corr_row2_ramp<A, LW, Taps, LO, HI + 1>(lo_p, hi_p, inrow, taps, uoff, acc2);
  else if constexpr (LO + 1 < GLC_RP) corr_row2_ramp<A, LW, Taps, LO + 1, LO + 1>(lo_p, hi_p, inrow, taps, uoff, acc2);
}

// acc2 += correlation of the thread's strip (origin `in`) with the taps of one phase (`taps` addresses that phase's table).
template <int A, int LW, class Taps>
__device__ __forceinline__ void corr_rows2_t(const float* __restrict__ in, int pitch, const Taps& taps,
                                           float2 (&acc2)[GLC_RP][GLC_RX]) {
  constexpr int UL = glc_ulen(A);
  constexpr int ROWS = GLC_RY + A - 1;
#pragma unroll 1
  for (int row = 0; row < ROWS; ++row) {
    const int j0 = A + GLC_RY - 2 - row;
    // copy 1 holds U shifted left by one float, so an odd j0 reads copy 1 at index j0 - 1 (even)
    const int uoff = (j0 & 1) ? (UL + (j0 - 1)) : j0;
    const int lo_r = row - (A - 1) > 0 ? row - (A - 1) : 0;
    const int hi_r = row < GLC_RY - 1 ? row : GLC_RY - 1;
    const int lo_p = lo_r >> 1, hi_p = hi_r >> 1;
    const float* inrow = in + row * pitch;
    if (lo_p == 0 && hi_p == GLC_RP - 1) corr_row2<A, (1 << GLC_RP) - 1, LW>(inrow, taps, uoff, acc2);
    else corr_row2_ramp<A, LW, Taps>(lo_p, hi_p, inrow, taps, uoff, acc2);
  }
}
// shared-memory table: the signature the cp.async kernels use
template <int A, int LW = 4>
__device__ __forceinline__ void corr_rows2(const float* __restrict__ in, int pitch, const float* __restrict__ u,
                                           float2 (&acc2)[GLC_RP][GLC_RX]) {
  corr_rows2_t<A, LW>(in, pitch, GlTapsSmem{u}, acc2);
}

// ---- column-pair scheme ------------------------------------------------------------------------------------------
// acc[r][cp] holds the output pair (column 2cp, column 2cp+1) of row r.  For tap (a, b) the pair needs the two ADJACENT inputs
// (strip[2cp+b], strip[2cp+b+1]) and ONE tap:  FFMA2 acc, w[a][b].F32 (scalar broadcast), (s_k, s_k+1).F32x2, acc.  Against the
// row-pair scheme above: (i) no zero taps -- a row pair spends A+1 FMAs on A taps because its two rows are one tap apart, 7.7 % of
// the FMA-pipe work at A = 13; (ii) the taps of one (strip row, output row) are one contiguous tap row, read with LDS.128
// (A/4 wavefronts per 2A FFMA2 instead of one LDS.64 per 4); (iii) ramp rows need no specialised instances: output rows outside
// [row-A+1, row] are skipped by a uniform test, so the loop body is one block of code per output row.
// Even input pairs come straight from the LDS.64 / LDS.128 strip loads; odd pairs are re-packed once per strip row.
#define GLC_CP (GLC_RX / 2)
__host__ __device__ constexpr int glc_wpitch(int A) { return (A + 3) & ~3; }   // floats per tap row; table: [phase][A][wpitch]
// `zero` is a run-time 0 (a kernel parameter): the odd pairs are formed as x ^ zero, two LOP3 on the integer pipe, so that they
// are COMPUTED values with registers of their own.  A plain register re-pack is a copy to ptxas, which coalesces it with the strip
// registers and then re-aligns the pair with two MOVs in front of almost every FFMA2 that uses it (one MOV per FFMA2 overall).
template <int A, int LW>
__device__ __forceinline__ void corr_rows_cp(const float* __restrict__ in, int pitch, const float* __restrict__ w_ph,
                                             float2 (&acc)[GLC_RY][GLC_CP], unsigned zero) {
  constexpr int SL = (GLC_RX + A - 1 + 3) & ~3;
  constexpr int NE = SL / 2;          // even pairs (s[2v], s[2v+1])
  constexpr int WP = glc_wpitch(A);
  constexpr int ROWS = GLC_RY + A - 1;
#pragma unroll 1
  for (int row = 0; row < ROWS; ++row) {
    const float* inrow = in + row * pitch;
    // input pairs as opaque 64-bit registers: even pairs (s[2v], s[2v+1]) straight from the strip loads, odd pairs
    // (s[2v+1], s[2v+2]) packed ONCE per strip row -- left to itself ptxas re-packs an odd pair with two MOVs in front of almost
    // every FFMA2 that uses it (one MOV per FFMA2 overall: the kernel became issue-bound)
    unsigned long long se[NE], so[NE];
    if constexpr (LW == 4) {
      const ulonglong2* src = reinterpret_cast<const ulonglong2*>(inrow);
#pragma unroll
      for (int v = 0; v < SL / 4; ++v) { const ulonglong2 t = src[v]; se[2 * v] = t.x; se[2 * v + 1] = t.y; }
    } else {
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(inrow);
#pragma unroll
      for (int v = 0; v < NE; ++v) se[v] = src[v];
    }
#pragma unroll
    for (int v = 0; v + 1 < NE; ++v)
      asm("{ .reg .b32 a0, a1, b0, b1; mov.b64 {a0, a1}, %1; mov.b64 {b0, b1}, %2; xor.b32 a1, a1, %3; xor.b32 b0, b0, %3; mov.b64 %0, {a1, b0}; }"
          : "=l"(so[v]) : "l"(se[v]), "l"(se[v + 1]), "r"(zero));
    so[NE - 1] = se[NE - 1];   // never used (k <= RX + A - 3 < SL - 1); keeps the array fully defined
    auto apply = [&](int r, int a) {
      const float4* wv = reinterpret_cast<const float4*>(w_ph + a * WP);
      float w[WP];
#pragma unroll
      for (int v = 0; v < WP / 4; ++v) { const float4 t = wv[v]; w[4 * v] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w; }
#pragma unroll
      for (int b = 0; b < A; ++b) {
#pragma unroll
        for (int cp = 0; cp < GLC_CP; ++cp) {
          const int k = 2 * cp + b;
          const unsigned long long sp = (k & 1) ? so[k >> 1] : se[k >> 1];
          asm("{ .reg .b64 ww, cc; mov.b64 ww, {%2, %2}; mov.b64 cc, {%0, %1}; fma.rn.f32x2 cc, ww, %3, cc; mov.b64 {%0, %1}, cc; }"
              : "+f"(acc[r][cp].x), "+f"(acc[r][cp].y) : "f"(w[b]), "l"(sp));
        }
      }
    };
    if (row >= GLC_RY - 1 && row <= A - 1) {
      // every output row of the tile is reached: one straight-line block (the odd pairs are packed once, the tap loads of the six
      // output rows are scheduled ahead of the FFMA2 stream)
#pragma unroll
      for (int r = 0; r < GLC_RY; ++r) apply(r, row - r);
    } else {
#pragma unroll
      for (int r = 0; r < GLC_RY; ++r) {
        const int a = row - r;
        if (a >= 0 && a < A) apply(r, a);   // warp-uniform
      }
    }
  }
}

struct GlLikeArgs {
  const float* observed;    // [n*n]
  const float* error_map;   // [n*n] or null
  const unsigned char* mask;  // [n*n] or null
  float bg2;                // background_rms^2
  float inv_exp;            // 1 / exp_time
  int enabled;
};

// Epilogue shared by the forward kernels: x scale, image store, chi^2 / normalisation partial sums of this tile
// and dL/d(image) (tf/model.py:91-101).
__device__ __forceinline__ void glc_fwd_epilogue(const GlConvGeom& g, const float2 (&acc2)[GLC_RP][GLC_RX], bool active, int b, int tile,
                                                 int ntiles, int oy0, int ox0, int ty, int tx, float scale, float* __restrict__ img,
                                                 const GlLikeArgs& like, float* __restrict__ part, float* __restrict__ gimg,
                                                 float (*s_red)[8]) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  float chi2 = 0.f, norm = 0.f;
  if (active) {
#pragma unroll
    for (int r = 0; r < GLC_RY; ++r) {
      const int oy = oy0 + ty * GLC_RY + r;
#pragma unroll
      for (int c = 0; c < GLC_RX; ++c) {
        const int ox = ox0 + tx * GLC_RX + c;
        if (oy < g.n && ox < g.n) {
          const float v = ((r & 1) ? acc2[r >> 1][c].y : acc2[r >> 1][c].x) * scale;
          const size_t o = (size_t)oy * g.n + ox;
          if (img) img[(size_t)b * g.n * g.n + o] = v;
          if (like.enabled) {
            const float m = like.mask ? (like.mask[o] ? 1.f : 0.f) : 1.f;
            const float res = v - like.observed[o];
            // var = err^2 with err = sqrt(bg^2 + I/t) (tf/model.py:95-98) or the fixed error map; a negative variance
            // gives NaN like the reference.  1/err is one MUFU rsqrt plus a Newton step (full fp32 accuracy) and the
            // log is lg2 * ln2 (abs. error 2^-22): the IEEE sqrt / four divisions / logf of a literal transcription
            // were a sixth of this kernel's instructions.
            float var, dvar;
            if (like.error_map) { const float e = like.error_map[o]; var = e * e; dvar = 0.f; }
            else { var = fmaf(v, like.inv_exp, like.bg2); dvar = like.inv_exp; }
            float r;
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(var));
            r = r * fmaf(-0.5f * var * r, r, 1.5f);
            const float ivar = r * r;
            const float q = res * r;
            chi2 += q * q * m;
            norm += __logf(6.283185307179586f * var) * m;
            if (gimg) {
              // d(-0.5*(chi2+norm))/dI = -0.5 m (2 res/var - res^2 dvar/var^2 + dvar/var)
              const float gi = -0.5f * m * ivar * fmaf(-(res * res) * ivar, dvar, fmaf(2.f, res, dvar));
              gimg[(size_t)b * g.n * g.n + o] = gi;
            }
          }
        }
      }
    }
  }
  if (like.enabled && part) {
    for (int off = 16; off > 0; off >>= 1) {
      chi2 += __shfl_xor_sync(0xffffffffu, chi2, off);
      norm += __shfl_xor_sync(0xffffffffu, norm, off);
    }
    const int warp = tid >> 5, lane = tid & 31, nw = (nthr + 31) >> 5;
    if (lane == 0) { s_red[0][warp] = chi2; s_red[1][warp] = norm; }
    __syncthreads();
    if (tid == 0) {
      float c2 = 0.f, nm = 0.f;
      for (int w2 = 0; w2 < nw; ++w2) { c2 += s_red[0][w2]; nm += s_red[1][w2]; }
      float* p = part + ((size_t)b * ntiles + tile) * 2;
      p[0] = c2; p[1] = nm;
    }
  }
}

// Forward: ss image -> pooled image (x scale), optional likelihood partial sums and dL/d(image).
//   grid = tiles_x * tiles_y * n_images (tile fastest), block = ntx*nty threads (rounded up to a warp multiple)
//   part [bs][tiles][2] = (chi2, normalization) partial sums of this tile
template <int A>
__global__ void __maxnreg__(96) k_conv_fwd(GlConvGeom g, const float* __restrict__ ss_img, const float* __restrict__ wts,
                                                  float scale, float* __restrict__ img, GlLikeArgs like,
                                                  float* __restrict__ part, float* __restrict__ gimg) {
  extern __shared__ __align__(128) float glc_smem_f[];
  float* smem = glc_smem_f;
  const int nph = g.ss * g.ss;
  const int tile_size = g.in_rows * g.in_pitch;
  float* s_in = smem;                              // [2][in_rows][in_pitch]  double-buffered phase tile
  float* s_w = smem + 2 * tile_size;               // [nph][A][2][ulen] packed tap table
  __shared__ float s_red[2][8];

  const int ntiles = g.tiles_x * g.tiles_y;
  const int b = blockIdx.x / ntiles, tile = blockIdx.x - b * ntiles;
  const int oy0 = (tile / g.tiles_x) * g.th, ox0 = (tile % g.tiles_x) * g.tw;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;

  constexpr int UTAB = A * 2 * glc_ulen(A);         // floats per phase
  for (int i = tid; i < nph * UTAB; i += nthr) s_w[i] = wts[i];

  const float* src_b = ss_img + (size_t)b * nph * g.n * g.n;
  // Stage the tile of image-phase (py', px') = (q / ss, q % ss): padded-phase row r of the conv holds
  // image row ri = r - sy of that phase (sy = (py' + pad) / ss), zero outside the image.
  auto issue_phase = [&](int q, float* buf) {
    const int pyi = q / g.ss, pxi = q - pyi * g.ss;
    const int sy = (pyi + g.pad) / g.ss, sx = (pxi + g.pad) / g.ss;
    const float* src = src_b + (size_t)q * g.n * g.n;
    const unsigned buf_u32 = (unsigned)__cvta_generic_to_shared(buf);
    if (((g.n | ox0 | sx) & 1) == 0) {
      // 8-byte copies: n, ox0 and the phase shift are even, so element pairs are aligned on both sides
      // and lie entirely inside or entirely outside the image.  Per-lane column bookkeeping is hoisted.
      const int npair = g.in_pitch >> 1;
      const int cj0 = ox0 - sx + 2 * lane;          // image column of this lane's first pair
      for (int li = warp; li < g.in_rows; li += nw) {
        const int ri = oy0 + li - sy;
        const bool row_ok = ri >= 0 && ri < g.n;
        const float* srow = src + (size_t)min(max(ri, 0), g.n - 1) * g.n;
        const unsigned drow = buf_u32 + (unsigned)(li * g.in_pitch) * 4u + 8u * lane;
        for (int pj = lane, cj = cj0, k = 0; pj < npair; pj += 32, cj += 64, ++k) {
          const bool ok = row_ok && cj >= 0 && cj < g.n;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(drow + 256u * k), "l"(srow + min(max(cj, 0), g.n - 2)),
                       "r"(ok ? 8u : 0u) : "memory");
        }
      }
    } else {
      for (int li = warp; li < g.in_rows; li += nw) {
        const int ri = oy0 + li - sy;
        const bool row_ok = ri >= 0 && ri < g.n;
        const float* srow = src + (size_t)min(max(ri, 0), g.n - 1) * g.n;
        const unsigned drow = buf_u32 + (unsigned)(li * g.in_pitch) * 4u;
        for (int lj = lane; lj < g.in_pitch; lj += 32) {
          const int cj = ox0 + lj - sx;
          const bool ok = row_ok && cj >= 0 && cj < g.n;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(drow + 4u * lj), "l"(srow + min(max(cj, 0), g.n - 1)),
                       "r"(ok ? 4u : 0u) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int ty = tid / g.tpr, tx = tid - ty * g.tpr;
  const bool active = ty < g.nty && tx < g.ntx;
  float2 acc2[GLC_RP][GLC_RX];
#pragma unroll
  for (int r = 0; r < GLC_RP; ++r)
#pragma unroll
    for (int c = 0; c < GLC_RX; ++c) acc2[r][c] = make_float2(0.f, 0.f);
  const int origin = ty * GLC_RY * g.in_pitch + tx * GLC_RX;

  issue_phase(0, s_in);
  for (int q = 0; q < nph; ++q) {
    if (q + 1 < nph) {
      issue_phase(q + 1, s_in + ((q + 1) & 1) * tile_size);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (active) {
      // conv phase of this image phase: (py, px) = ((py' + pad) % ss, (px' + pad) % ss)
      const int pyi = q / g.ss, pxi = q - pyi * g.ss;
      const int ph = ((pyi + g.pad) % g.ss) * g.ss + (pxi + g.pad) % g.ss;
      corr_rows2<A>(s_in + (q & 1) * tile_size + origin, g.in_pitch, s_w + ph * UTAB, acc2);
    }
    __syncthreads();
  }

  glc_fwd_epilogue(g, acc2, active, b, tile, ntiles, oy0, ox0, ty, tx, scale, img, like, part, gimg, s_red);
}

// ---- TMA (cp.async.bulk.tensor) + mbarrier plumbing of the forward kernel ---------------------------
__device__ __forceinline__ unsigned glc_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void glc_mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(glc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void glc_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(glc_smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait (a broken pipeline must not hang the GPU): false on time-out.
__device__ __forceinline__ bool glc_mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = glc_smem_u32(bar);
  for (int spin = 0; spin < (1 << 22); ++spin) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void glc_tma_load_3d(void* dst, const CUtensorMap* tmap, int c0, int c1, int c2, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(glc_smem_u32(dst)), "l"(tmap), "r"(glc_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void glc_bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(glc_smem_u32(dst)), "l"(src), "r"(bytes), "r"(glc_smem_u32(bar)) : "memory");
}

// Forward kernel with TMA staging.  The supersampled buffer is a 3-D tensor (n, n, n_images * ss^2) of phase
// sub-images; one elected thread asks the TMA unit for the (in_pitch x in_rows) halo tile of a phase at signed
// coordinates (ox0 - sx, oy0 - sy, b * ss^2 + q) -- the SAME zero padding is the unit's out-of-bounds fill --
// and for the tap table (1-D bulk copy).  Completion is an mbarrier transaction count, so the other threads
// issue nothing but FFMA2 / LDS until they wait on the barrier of the phase they need: the per-element
// address arithmetic and bounds tests of k_conv_fwd's cp.async loader (a quarter of its instructions) are gone.
// The unit wants the innermost coordinate on a 16-byte boundary (measured: anything else faults), so the tile
// starts at the column rounded down to a multiple of 4 and is 4 columns wider (pitch tma_pitch); the thread
// strips then start r = 0 or 2 floats into it and are read with LDS.64.
// Needs n % 4 == 0 (16-byte global strides), even shifts and a tile of at most 256 x 256; otherwise k_conv_fwd runs.
// NFC > 0: the tap table (NFC floats) arrives as the by-value parameter `ctaps` and is read through the uniform datapath
// (GlTapsConst); NFC == 0: it is staged in shared memory from `wts` as before.
template <int A, int NFC = 0>
__global__ void __maxnreg__(96) k_conv_fwd_tma(const __grid_constant__ CUtensorMap tmap, GlConvGeom g, const float* __restrict__ wts,
                                               float scale, float* __restrict__ img, GlLikeArgs like,
                                               float* __restrict__ part, float* __restrict__ gimg,
                                               const __grid_constant__ GlTapsC<NFC> ctaps) {
  extern __shared__ __align__(128) float glc_smem_ft[];
  float* smem = glc_smem_ft;
  const int nph = g.ss * g.ss;
  const int tstride = g.phase_stride;              // floats per staged tile, a multiple of 32 (128-byte TMA destinations)
  float* s_in = smem;                              // [2][tstride]  double-buffered phase tile
  float* s_w = smem + 2 * tstride;                 // [nph][A][2][ulen] packed tap table
  __shared__ float s_red[2][8];
  __shared__ __align__(8) unsigned long long s_bar[5];   // tile buffer 0 / 1 full, tap table full, tile buffer 0 / 1 drained

  const int ntiles = g.tiles_x * g.tiles_y;
  const int b = blockIdx.x / ntiles, tile = blockIdx.x - b * ntiles;
  const int oy0 = (tile / g.tiles_x) * g.th, ox0 = (tile % g.tiles_x) * g.tw;
  const int tid = threadIdx.x;
  constexpr int UTAB = A * 2 * glc_ulen(A);         // floats per phase
  const unsigned tile_bytes = (unsigned)(g.in_rows * g.tma_pitch) * 4u;

  if (tid == 0) {
    glc_mbar_init(&s_bar[0], 1); glc_mbar_init(&s_bar[1], 1); glc_mbar_init(&s_bar[2], 1);
    glc_mbar_init(&s_bar[3], (int)(blockDim.x >> 5)); glc_mbar_init(&s_bar[4], (int)(blockDim.x >> 5));   // one arrival per warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // image phase (py', px') = (q / ss, q % ss): padded-phase row r holds image row r - sy, sy = (py' + pad) / ss
  auto issue_phase = [&](int q) {
    const int pyi = q / g.ss, pxi = q - pyi * g.ss;
    const int sy = (pyi + g.pad) / g.ss, sx = (pxi + g.pad) / g.ss;
    glc_mbar_expect_tx(&s_bar[q & 1], tile_bytes);
    glc_tma_load_3d(s_in + (q & 1) * tstride, &tmap, (ox0 - sx) & ~3, oy0 - sy, b * nph + q, &s_bar[q & 1]);
  };
  if (tid == 0) {
    if constexpr (NFC == 0) {
      glc_mbar_expect_tx(&s_bar[2], (unsigned)(nph * UTAB) * 4u);
      glc_bulk_load(s_w, wts, (unsigned)(nph * UTAB) * 4u, &s_bar[2]);
    }
    issue_phase(0);
    if (nph > 1) issue_phase(1);
  }

  const int ty = tid / g.tpr, tx = tid - ty * g.tpr;
  const bool active = ty < g.nty && tx < g.ntx;
  float2 acc2[GLC_RP][GLC_RX];
#pragma unroll
  for (int r = 0; r < GLC_RP; ++r)
#pragma unroll
    for (int c = 0; c < GLC_RX; ++c) acc2[r][c] = make_float2(0.f, 0.f);
  const int origin = ty * GLC_RY * g.tma_pitch + tx * GLC_RX;

  bool ok = true;
  if constexpr (NFC == 0) ok = glc_mbar_wait(&s_bar[2], 0);
  for (int q = 0; q < nph; ++q) {
    ok = glc_mbar_wait(&s_bar[q & 1], (unsigned)(q >> 1) & 1u) && ok;
    if (active) {
      // conv phase of this image phase: (py, px) = ((py' + pad) % ss, (px' + pad) % ss)
      const int pyi = q / g.ss, pxi = q - pyi * g.ss;
      const int ph = ((pyi + g.pad) % g.ss) * g.ss + (pxi + g.pad) % g.ss;
      const int r = (ox0 - (pxi + g.pad) / g.ss) & 3;      // columns between the aligned tile start and the first one needed
      if constexpr (NFC > 0) corr_rows2_t<A, 2>(s_in + (q & 1) * tstride + origin + r, g.tma_pitch, GlTapsConst<NFC>{ctaps, ph * UTAB}, acc2);
      else corr_rows2<A, 2>(s_in + (q & 1) * tstride + origin + r, g.tma_pitch, s_w + ph * UTAB, acc2);
    }
    if (q + 2 < nph) {
      // Buffer hand-back without a CTA barrier: every warp reports that it is done with this buffer and moves on to
      // the next phase (already staged); only thread 0 waits for the five reports before it re-arms the buffer.
      __syncwarp();
      if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(glc_smem_u32(&s_bar[3 + (q & 1)])) : "memory");
      if (tid == 0) {
        ok = glc_mbar_wait(&s_bar[3 + (q & 1)], (unsigned)(q >> 1) & 1u) && ok;
        issue_phase(q + 2);
      }
    }
  }
  if (!ok) {                              // barrier time-out: poison the outputs instead of returning stale data
#pragma unroll
    for (int r = 0; r < GLC_RP; ++r)
#pragma unroll
      for (int c = 0; c < GLC_RX; ++c) acc2[r][c] = make_float2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
  }
  glc_fwd_epilogue(g, acc2, active, b, tile, ntiles, oy0, ox0, ty, tx, scale, img, like, part, gimg, s_red);
}

// Adjoint: dL/d(image) [bs][n][n] -> dL/d(ss image) [bs][hs][hs]  (x scale).
//   Work space is the padded-phase grid (r, c): ss pixel (i, j) = (ss*r + py - pad, ss*c + px - pad).
//   grid = tiles * n_images over r,c in [rc0, rc0 + nr), nr = number of padded-phase rows touching the image.
//   wts here are the flipped taps: wflip[ph][a'][b'] = W[ph][A-1-a'][A-1-b'].
template <int A>
__global__ void __launch_bounds__(256, 2) k_conv_bwd(GlConvGeom g, const float* __restrict__ gimg, const float* __restrict__ wts,
                                                  float scale, const unsigned char* __restrict__ ss_mask,
                                                  float* __restrict__ gss, const int* __restrict__ img_list) {
  (void)ss_mask;
  extern __shared__ __align__(128) float glc_smem_b[];
  float* smem = glc_smem_b;
  const int nph = g.ss * g.ss;
  float* s_in = smem;                                   // [in_rows][in_pitch]  zero-padded dL/d(image)
  float* s_w = smem + g.in_rows * g.in_pitch;           // [nph][A][2][ulen] packed (flipped) tap table
  const int ntiles = g.tiles_x * g.tiles_y;
  int b = blockIdx.x / ntiles;
  const int tile = blockIdx.x - b * ntiles;
  if (img_list) {                                       // only the listed images: [0] = count, [1..] = image indices
    if (b >= img_list[0]) return;
    b = img_list[1 + b];
  }
  const int r0 = g.rc0 + (tile / g.tiles_x) * g.th, c0 = g.rc0 + (tile % g.tiles_x) * g.tw;
  const int tid = threadIdx.x, nthr = blockDim.x;

  constexpr int UTAB = A * 2 * glc_ulen(A);
  for (int i = tid; i < nph * UTAB; i += nthr) s_w[i] = wts[i];
  {
    const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
    const float* src = gimg + (size_t)b * g.n * g.n;
    for (int li = warp; li < g.in_rows; li += nw) {
      const int oy = r0 + li - (A - 1);
      const bool row_ok = oy >= 0 && oy < g.n;
      const float* srow = src + (size_t)(row_ok ? oy : 0) * g.n + (c0 - (A - 1));
      float* drow = s_in + li * g.in_pitch;
      for (int lj = lane; lj < g.in_pitch; lj += 32) {
        const int ox = c0 + lj - (A - 1);
        const bool ok = row_ok && ox >= 0 && ox < g.n;
        glc_cp_async4(drow + lj, ok ? srow + lj : src, ok ? 4 : 0);
      }
    }
    glc_cp_async_wait_all();
  }
  __syncthreads();
  const int ty = tid / g.tpr, tx = tid - ty * g.tpr;
  if (ty >= g.nty || tx >= g.ntx) return;
  const int origin = ty * GLC_RY * g.in_pitch + tx * GLC_RX;
  float* dst = gss + (size_t)b * g.hs * g.hs;
  for (int ph = 0; ph < nph; ++ph) {
    const int py = ph / g.ss, px = ph - py * g.ss;
    float2 acc2[GLC_RP][GLC_RX];
#pragma unroll
    for (int r = 0; r < GLC_RP; ++r)
#pragma unroll
      for (int c = 0; c < GLC_RX; ++c) acc2[r][c] = make_float2(0.f, 0.f);
    corr_rows2<A>(s_in + origin, g.in_pitch, s_w + ph * UTAB, acc2);
    // ss pixel (i, j) = (ss*r + py - pad, ss*c + px - pad) lives in image phase (i mod ss, j mod ss) at
    // sub-image pixel (i div ss, j div ss): both are constant offsets per conv phase, so the phase-major
    // destination is contiguous in c and needs no per-element division.
    const int fy = py - g.pad, fx = px - g.pad;
    const int dy = (fy >= 0) ? fy / g.ss : -((-fy + g.ss - 1) / g.ss);   // floor division
    const int dx = (fx >= 0) ? fx / g.ss : -((-fx + g.ss - 1) / g.ss);
    const int q = (fy - dy * g.ss) * g.ss + (fx - dx * g.ss);
    float* dq = dst + (size_t)q * g.n * g.n;
#pragma unroll
    for (int r = 0; r < GLC_RY; ++r) {
      const int ri = r0 + ty * GLC_RY + r + dy;
      if (ri < 0 || ri >= g.n) continue;
#pragma unroll
      for (int c = 0; c < GLC_RX; ++c) {
        const int cj = c0 + tx * GLC_RX + c + dx;
        if (cj >= 0 && cj < g.n) dq[(size_t)ri * g.n + cj] = ((r & 1) ? acc2[r >> 1][c].y : acc2[r >> 1][c].x) * scale;
      }
    }
  }
}

// Adjoint with TMA staging on both sides: the zero-padded dL/d(image) halo tile comes in through one
// cp.async.bulk.tensor load (out-of-bounds fill = the zero padding; same 16-byte coordinate rule and LDS.64
// strips as k_conv_fwd_tma), and every phase result leaves through a shared-memory staging tile and one
// cp.async.bulk.tensor store, which clips the tile against the image bounds by itself -- the per-element
// address arithmetic and range tests of k_conv_bwd's loader and writer (a quarter of its instructions) are gone.
// Needs n % 4 == 0, an even input shift and output tile origins on 16-byte boundaries; otherwise k_conv_bwd runs.
template <int A, int NFC = 0>
__global__ void __launch_bounds__(256, 2) k_conv_bwd_tma(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                                                       GlConvGeom g, const float* __restrict__ wts, float scale,
                                                       const __grid_constant__ GlTapsC<NFC> ctaps, const int* __restrict__ img_list) {
  extern __shared__ __align__(128) float glc_smem_bt[];
  float* smem = glc_smem_bt;
  const int nph = g.ss * g.ss;
  float* s_in = smem;                                   // [in_rows][tma_pitch]  zero-padded dL/d(image)
  float* s_out = smem + g.phase_stride;                 // [th][tw]  one phase of the result, staged for the TMA store
  float* s_w = s_out + g.out_stride;                    // [nph][A][2][ulen] packed (flipped) tap table
  __shared__ __align__(8) unsigned long long s_bar[1];
  const int ntiles = g.tiles_x * g.tiles_y;
  int b = blockIdx.x / ntiles;
  const int tile = blockIdx.x - b * ntiles;
  if (img_list) {                                       // only the listed images: [0] = count, [1..] = image indices
    if (b >= img_list[0]) return;
    b = img_list[1 + b];
  }
  const int r0 = g.rc0 + (tile / g.tiles_x) * g.th, c0 = g.rc0 + (tile % g.tiles_x) * g.tw;
  const int tid = threadIdx.x;
  constexpr int UTAB = A * 2 * glc_ulen(A);

  if (tid == 0) {
    glc_mbar_init(&s_bar[0], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned wbytes = NFC == 0 ? (unsigned)(nph * UTAB) * 4u : 0u, tbytes = (unsigned)(g.in_rows * g.tma_pitch) * 4u;
    glc_mbar_expect_tx(&s_bar[0], wbytes + tbytes);
    if constexpr (NFC == 0) glc_bulk_load(s_w, wts, wbytes, &s_bar[0]);
    glc_tma_load_3d(s_in, &tm_in, (c0 - (A - 1)) & ~3, r0 - (A - 1), b, &s_bar[0]);
  }
  const int ty = tid / g.tpr, tx = tid - ty * g.tpr;
  const bool active = ty < g.nty && tx < g.ntx;
  const int origin = ty * GLC_RY * g.tma_pitch + tx * GLC_RX + ((c0 - (A - 1)) & 3);
  const bool ok = glc_mbar_wait(&s_bar[0], 0);
  if (!ok) scale = __int_as_float(0x7fc00000);          // barrier time-out: poison the result instead of returning stale data
  // Output path.  When a warp's 32 threads cover complete rows of the thread grid (band_rows > 0) the warp owns a
  // band of full tile rows: it stages and stores that band by itself (its lane 0 issues the TMA store and waits on its
  // own bulk group), so after the input barrier the warps never meet again -- two CTA barriers per phase cost 21 % of
  // the kernel's stall samples.  Otherwise the CTA stages the whole tile and thread 0 stores it.
  const int warp = tid >> 5, lane = tid & 31;
  const bool wstore = g.band_rows > 0;
  float* s_band = wstore ? s_out + warp * g.band_stride : s_out;
  const int ty_loc = wstore ? ty - warp * (32 / g.tpr) : ty;      // thread row inside the staging area it writes to
  for (int ph = 0; ph < nph; ++ph) {
    float2 acc2[GLC_RP][GLC_RX];
#pragma unroll
    for (int r = 0; r < GLC_RP; ++r)
#pragma unroll
      for (int c = 0; c < GLC_RX; ++c) acc2[r][c] = make_float2(0.f, 0.f);
    if (active) {
      if constexpr (NFC > 0) corr_rows2_t<A, 2>(s_in + origin, g.tma_pitch, GlTapsConst<NFC>{ctaps, ph * UTAB}, acc2);
      else corr_rows2<A, 2>(s_in + origin, g.tma_pitch, s_w + ph * UTAB, acc2);
    }
    if (ph > 0) {                                       // the previous phase's store must have read the staging tile
      if (wstore) { if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); __syncwarp(); }
      else { if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); __syncthreads(); }
    }
    if (active) {
      float4* drow = reinterpret_cast<float4*>(s_band + (ty_loc * GLC_RY) * g.tw + tx * GLC_RX);
      const int rpitch = g.tw >> 2;
#pragma unroll
      for (int rp = 0; rp < GLC_RP; ++rp) {
        drow[(2 * rp) * rpitch] = make_float4(acc2[rp][0].x * scale, acc2[rp][1].x * scale, acc2[rp][2].x * scale, acc2[rp][3].x * scale);
        drow[(2 * rp + 1) * rpitch] = make_float4(acc2[rp][0].y * scale, acc2[rp][1].y * scale, acc2[rp][2].y * scale, acc2[rp][3].y * scale);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the TMA unit
    if (wstore) __syncwarp(); else __syncthreads();
    if (wstore ? (lane == 0 && warp * (32 / g.tpr) < g.nty) : (tid == 0)) {
      // ss pixel (i, j) = (ss*r + py - pad, ss*c + px - pad) lives in image phase (i mod ss, j mod ss) at sub-image
      // pixel (i div ss, j div ss): constant offsets per conv phase, so the tile lands as one box of phase image q
      const int py = ph / g.ss, px = ph - py * g.ss;
      const int fy = py - g.pad, fx = px - g.pad;
      const int dy = (fy >= 0) ? fy / g.ss : -((-fy + g.ss - 1) / g.ss);   // floor division
      const int dx = (fx >= 0) ? fx / g.ss : -((-fx + g.ss - 1) / g.ss);
      const int q = (fy - dy * g.ss) * g.ss + (fx - dx * g.ss);
      const int yb = wstore ? warp * g.band_rows : 0;
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                   ::"l"(&tm_out), "r"(glc_smem_u32(s_band)), "r"(c0 + dx), "r"(r0 + dy + yb), "r"(b * nph + q) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (wstore ? lane == 0 : tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the last store's read
}
