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

template <int A, int MASK>
__device__ __forceinline__ void corr_row2(const float* __restrict__ inrow, const float* __restrict__ urow /* &U[0][copy][j0e] */,
                                          float2 (&acc2)[GLC_RP][GLC_RX]) {
  constexpr int SL = (GLC_RX + A - 1 + 3) & ~3;
  constexpr int UL = glc_ulen(A);
  float strip[SL];
  const float4* src = reinterpret_cast<const float4*>(inrow);
#pragma unroll
  for (int v = 0; v < SL / 4; ++v) {
    float4 t = src[v];
    strip[4 * v] = t.x; strip[4 * v + 1] = t.y; strip[4 * v + 2] = t.z; strip[4 * v + 3] = t.w;
  }
#pragma unroll
  for (int b = 0; b < A; ++b) {
    const float2* u2 = reinterpret_cast<const float2*>(urow + b * 2 * UL);
#pragma unroll
    for (int rp = 0; rp < GLC_RP; ++rp) {
      if ((MASK >> rp) & 1) {
        const float2 w2 = u2[rp];
#pragma unroll
        for (int c = 0; c < GLC_RX; ++c) {
          const float sv = strip[c + b];
          acc2[rp][c] = __ffma2_rn(make_float2(sv, sv), w2, acc2[rp][c]);
        }
      }
    }
  }
}

// acc2 += correlation of the thread's strip (origin `in`) with the taps of one phase (table `u`).
template <int A>
__device__ __forceinline__ void corr_rows2(const float* __restrict__ in, int pitch, const float* __restrict__ u,
                                           float2 (&acc2)[GLC_RP][GLC_RX]) {
  constexpr int UL = glc_ulen(A);
  constexpr int ROWS = GLC_RY + A - 1;
#pragma unroll 1
  for (int row = 0; row < ROWS; ++row) {
    const int j0 = A + GLC_RY - 2 - row;
    // copy 1 holds U shifted left by one float, so an odd j0 reads copy 1 at index j0 - 1 (even)
    const float* up = (j0 & 1) ? (u + UL + (j0 - 1)) : (u + j0);
    const int lo_r = row - (A - 1) > 0 ? row - (A - 1) : 0;
    const int hi_r = row < GLC_RY - 1 ? row : GLC_RY - 1;
    const int lo_p = lo_r >> 1, hi_p = hi_r >> 1;
    const float* inrow = in + row * pitch;
    if (lo_p == 0 && hi_p == GLC_RP - 1) corr_row2<A, (1 << GLC_RP) - 1>(inrow, up, acc2);
    else {
      int mask = 0;
      for (int q = lo_p; q <= hi_p; ++q) mask |= 1 << q;
      switch (mask) {
        case 1: corr_row2<A, 1>(inrow, up, acc2); break;
        case 2: corr_row2<A, 2>(inrow, up, acc2); break;
        case 3: corr_row2<A, 3>(inrow, up, acc2); break;
        case 4: corr_row2<A, 4>(inrow, up, acc2); break;
        case 6: corr_row2<A, 6>(inrow, up, acc2); break;
        default: corr_row2<A, 7>(inrow, up, acc2); break;
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

// Forward: ss image -> pooled image (x scale), optional likelihood partial sums and dL/d(image).
//   grid = tiles_x * tiles_y * n_images (tile fastest), block = ntx*nty threads (rounded up to a warp multiple)
//   part [bs][tiles][2] = (chi2, normalization) partial sums of this tile
template <int A>
__global__ void __maxnreg__(96) k_conv_fwd(GlConvGeom g, const float* __restrict__ ss_img, const float* __restrict__ wts,
                                                  float scale, float* __restrict__ img, GlLikeArgs like,
                                                  float* __restrict__ part, float* __restrict__ gimg) {
  extern __shared__ __align__(16) float smem[];
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
            float var, dvar;
            if (like.error_map) { const float e = like.error_map[o]; var = e * e; dvar = 0.f; }
            else {
              // err = sqrt(bg^2 + I/t), var = err^2 (tf/model.py:95-98); negative => NaN like the reference
              const float e = sqrtf(like.bg2 + v * like.inv_exp); var = e * e; dvar = like.inv_exp;
            }
            const float q = res / sqrtf(var);
            chi2 += q * q * m;
            norm += logf(6.283185307179586f * var) * m;
            if (gimg) {
              // d(-0.5*(chi2+norm))/dI
              const float gi = -0.5f * m * (2.f * res / var - (res * res) / (var * var) * dvar + dvar / var);
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

// Adjoint: dL/d(image) [bs][n][n] -> dL/d(ss image) [bs][hs][hs]  (x scale).
//   Work space is the padded-phase grid (r, c): ss pixel (i, j) = (ss*r + py - pad, ss*c + px - pad).
//   grid = tiles * n_images over r,c in [rc0, rc0 + nr), nr = number of padded-phase rows touching the image.
//   wts here are the flipped taps: wflip[ph][a'][b'] = W[ph][A-1-a'][A-1-b'].
template <int A>
__global__ void __launch_bounds__(256, 2) k_conv_bwd(GlConvGeom g, const float* __restrict__ gimg, const float* __restrict__ wts,
                                                  float scale, const unsigned char* __restrict__ ss_mask,
                                                  float* __restrict__ gss) {
  (void)ss_mask;
  extern __shared__ __align__(16) float smem[];
  const int nph = g.ss * g.ss;
  float* s_in = smem;                                   // [in_rows][in_pitch]  zero-padded dL/d(image)
  float* s_w = smem + g.in_rows * g.in_pitch;           // [nph][A][2][ulen] packed (flipped) tap table
  const int ntiles = g.tiles_x * g.tiles_y;
  const int b = blockIdx.x / ntiles, tile = blockIdx.x - b * ntiles;
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
