// Explicit data rows: the general case of build_A_data_matrix (SLR:1301-1654) -- tilt/psi/dy != 0 and/or trilinear
// interpolation -- where a ray is no longer one in-plane gather per slice, so the matrix-free projector kernels do not
// apply.  The rows are BUILT ON THE GPU with the reference's float64 operation sequence (one thread per image pixel
// (k, j) and symmetry copy, the numba loop of SLR:1403-1557) into a CSR + its transpose, and the solvers apply them
// with k_fwd_csr / k_adj_csc.  Inside the batch these rows occupy "pseudo views": chunks of rows_per_view padded rows
// marked as tie views, so that every other kernel (symmetry rows, LSMR/TRF state machines, norms, score) is unchanged
// -- the fast projector kernels skip them, the adjoint kernels add BD::vtie.
#pragma once
#include "hb2_tie.cuh"

struct ExpGeo {
  int D2, L2, L3, L3P, linear;
  double s, dy;
  double Ayx[9];  // transpose of R.from_euler("yx", (tilt, psi)).as_matrix(): apply(inverse=True) multiplies by it
};
struct ExpCopy {
  double A[9];    // transpose of R.from_euler("z", angle).as_matrix() (from scipy, incl. its M22 = 1 or 1 - 2^-53)
  double zshift;  // h * rise_pixel
};

// scipy's Rotation.apply on one point: out[r] = fma(A[r][2], z, fma(A[r][1], y, A[r][0] * x)) (verified against
// scipy 1.18 with exact rationals, tests/test_host_cpu.py)
__device__ __forceinline__ void rot_apply(const double* A, double x, double y, double z, double& ox, double& oy, double& oz) {
  ox = __fma_rn(A[2], z, __fma_rn(A[1], y, __dmul_rn(A[0], x)));
  oy = __fma_rn(A[5], z, __fma_rn(A[4], y, __dmul_rn(A[3], x)));
  oz = __fma_rn(A[8], z, __fma_rn(A[7], y, __dmul_rn(A[6], x)));
}

// One thread per potential row (copy, k, j).  FILL = 0: cnt[t] = number of matrix entries of the row (0: the row does
// not exist, SLR:1547/1496).  FILL = 1: writes the entries at ent_off[t] (col = internal voxel index p*L3P + z,
// weight, row index) and the row's b / pixel id.
template <int FILL>
__global__ void __launch_bounds__(128) k_exp_rows(ExpGeo G, int ncopies, const ExpCopy* __restrict__ copies,
                                                  const double* __restrict__ xt, const double* __restrict__ zt,
                                                  const int* __restrict__ rank, const float* __restrict__ pix,
                                                  int* __restrict__ cnt, const int* __restrict__ ent_off,
                                                  const int* __restrict__ row_idx, int* __restrict__ col,
                                                  float* __restrict__ w, int* __restrict__ erow, int* __restrict__ rptr,
                                                  float* __restrict__ rb, int* __restrict__ rpid) {
  const long long R = (long long)G.L2 * G.D2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)ncopies * R) return;
  const int cp = (int)(t / R);
  const int k = (int)((t % R) / G.D2), j = (int)(t % G.D2);
  if (FILL && cnt[t] == 0) return;
  const ExpCopy& C = copies[cp];
  const int D2 = G.D2, c0 = D2 / 2, L3 = G.L3;
  double y0 = (double)(j - c0);
  if (G.s != 1.0) y0 = __dmul_rn(y0, G.s);
  y0 = __dsub_rn(y0, G.dy);
  int n = 0;
  int off = 0, row = 0;
  if (FILL) {
    off = ent_off[t]; row = row_idx[t];
    rptr[row] = off;
    rb[row] = pix[(size_t)j * G.L2 + k];
    rpid[row] = k * D2 + j;
  }
  for (int i = 0; i < D2; ++i) {
    const double x0 = xt[(size_t)k * D2 + i], z0 = zt[(size_t)k * D2 + i];
    double x1, y1, z1, x2, y2, z2;
    rot_apply(G.Ayx, x0, y0, z0, x1, y1, z1);
    rot_apply(C.A, x1, y1, z1, x2, y2, z2);
    z2 = __dsub_rn(z2, C.zshift);
    const double X = __dadd_rn(x2, (double)c0), Y = __dadd_rn(y2, (double)c0), Z = __dadd_rn(z2, (double)(L3 / 2));
    if (!G.linear) {
      const double zr = rint(Z), yr = rint(Y), xr = rint(X);
      if (!(zr >= 0.0 && zr <= (double)(L3 - 1) && yr >= 0.0 && yr <= (double)(D2 - 1) && xr >= 0.0 && xr <= (double)(D2 - 1)))
        continue;
      const int p = rank[(int)yr * D2 + (int)xr];
      if (p < 0) continue;
      if (FILL) { col[off + n] = p * G.L3P + (int)zr; w[off + n] = 1.f; erow[off + n] = row; }
      ++n;
    } else {
      // int() truncates toward zero (SLR:1418-1420); all eight corners must be inside the grid and the mask
      if (!(Z > -1.0 && Z < (double)L3 && Y > -1.0 && Y < (double)D2 && X > -1.0 && X < (double)D2)) continue;
      const int zi = (int)Z, yi = (int)Y, xi = (int)X;
      if (zi + 1 > L3 - 1 || yi + 1 > D2 - 1 || xi + 1 > D2 - 1) continue;
      const int p00 = rank[yi * D2 + xi], p01 = rank[yi * D2 + xi + 1];
      const int p10 = rank[(yi + 1) * D2 + xi], p11 = rank[(yi + 1) * D2 + xi + 1];
      if (p00 < 0 || p01 < 0 || p10 < 0 || p11 < 0) continue;
      if (FILL) {
        const double zf = __dsub_rn(Z, (double)zi), yf = __dsub_rn(Y, (double)yi), xf = __dsub_rn(X, (double)xi);
        const double mz = __dsub_rn(1.0, zf), my = __dsub_rn(1.0, yf), mx = __dsub_rn(1.0, xf);
        const int o = off + n;
        const int L3P = G.L3P;
        col[o + 0] = p00 * L3P + zi;     w[o + 0] = (float)__dmul_rn(__dmul_rn(mz, my), mx);
        col[o + 1] = p01 * L3P + zi;     w[o + 1] = (float)__dmul_rn(__dmul_rn(mz, my), xf);
        col[o + 2] = p10 * L3P + zi;     w[o + 2] = (float)__dmul_rn(__dmul_rn(mz, yf), mx);
        col[o + 3] = p11 * L3P + zi;     w[o + 3] = (float)__dmul_rn(__dmul_rn(mz, yf), xf);
        col[o + 4] = p00 * L3P + zi + 1; w[o + 4] = (float)__dmul_rn(__dmul_rn(zf, my), mx);
        col[o + 5] = p01 * L3P + zi + 1; w[o + 5] = (float)__dmul_rn(__dmul_rn(zf, my), xf);
        col[o + 6] = p10 * L3P + zi + 1; w[o + 6] = (float)__dmul_rn(__dmul_rn(zf, yf), mx);
        col[o + 7] = p11 * L3P + zi + 1; w[o + 7] = (float)__dmul_rn(__dmul_rn(zf, yf), xf);
#pragma unroll
        for (int e = 0; e < 8; ++e) erow[o + e] = row;
      }
      n += 8;
    }
  }
  if (!FILL) cnt[t] = n;
}

__global__ void k_exp_flags(long long n, const int* __restrict__ cnt, int* __restrict__ flag) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) flag[t] = cnt[t] > 0;
}
// half sets: potential row t = (copy, k, j) survives iff its pixel id k*D2 + j is set in the mask
__global__ void k_exp_maskrows(long long n, long long R, int D2, const uint8_t* __restrict__ mask, int* __restrict__ cnt) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long long kj = t % R;  // = k*D2 + j: the potential rows of a copy are laid out [k][j]
  if (!mask[kj]) cnt[t] = 0;
}
__global__ void k_iota(int n, int* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = e;
}
__global__ void k_exp_colcount(int nnz, const int* __restrict__ col, int* __restrict__ cc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nnz) atomicAdd(&cc[col[e]], 1);
}
__global__ void k_exp_gather(int nnz, const int* __restrict__ order, const int* __restrict__ erow, const float* __restrict__ w,
                             int* __restrict__ crow, float* __restrict__ cw) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nnz) { crow[e] = erow[order[e]]; cw[e] = w[order[e]]; }
}

// Forward: rows <- A_explicit src.  One CTA per pseudo view (rows_per_view consecutive padded rows), one warp per row.
// Epilogues and per-view partials as k_fwd_tie.
template <typename T, bool TRF>
__global__ void __launch_bounds__(HB2_BLOCK) k_fwd_csr(BD B, TD Tt, const T* __restrict__ src, T* __restrict__ rows, int mode) {
  const int view = B.tie_views[blockIdx.x];
  const int c = B.view_cand[view];
  __shared__ float red[HB2_BLOCK / 32];
  const int ppv = B.fwd_ppv;       // partial-sum slots per view = gridDim.y sub-CTAs per pseudo view
  const int sub = blockIdx.y;
  const bool act = tie_active<TRF>(B, Tt, c, mode, false);
  if (!act) {
    if (!TRF && threadIdx.x == 0) {
      if (mode == MODE_LSMR) B.part_u[view * ppv + sub] = 0.f;
      if (mode == MODE_SCORE) { B.part_s[3 * (view * ppv + sub)] = 0.f; B.part_s[3 * (view * ppv + sub) + 1] = 0.f; B.part_s[3 * (view * ppv + sub) + 2] = 0.f; }
    }
    return;
  }
  const T* __restrict__ vsrc = src + (size_t)c * B.npad;
  const long long vo = B.view_uoff[view] - B.cand_uoff[c];  // first explicit row of this pseudo view
  T* urow = rows + B.view_uoff[view];
  const float* brow = B.b + B.view_uoff[view];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float alpha = 0.f, inv_beta = 0.f;
  if (!TRF) { alpha = B.st[c].alpha; inv_beta = B.st[c].inv_beta; }
  float ss = 0.f, s_pb = 0.f, s_bb = 0.f;
  const int nrow = (int)min((long long)B.rows_per_view, (long long)B.exp_m - vo);
  for (int r = sub * (HB2_BLOCK / 32) + warp; r < nrow; r += ppv * (HB2_BLOCK / 32)) {
    const int e0 = B.exp_ptr[vo + r], e1 = B.exp_ptr[vo + r + 1];
    T acc = (T)0;
    int e = e0 + lane;
    for (; e + 96 < e1; e += 128) {  // four independent (column, weight, gather) chains in flight per lane
      const int c0 = B.exp_col[e], c1 = B.exp_col[e + 32], c2 = B.exp_col[e + 64], c3 = B.exp_col[e + 96];
      const float w0 = B.exp_w[e], w1 = B.exp_w[e + 32], w2 = B.exp_w[e + 64], w3 = B.exp_w[e + 96];
      const T v0 = vsrc[c0], v1 = vsrc[c1], v2 = vsrc[c2], v3 = vsrc[c3];
      acc += (T)w0 * v0; acc += (T)w1 * v1; acc += (T)w2 * v2; acc += (T)w3 * v3;
    }
    for (; e < e1; e += 32) acc += (T)B.exp_w[e] * vsrc[B.exp_col[e]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      if (TRF) {
        urow[r] = acc;
      } else if (mode == MODE_LSMR) {
        const float un = fadd_(fmul_(fmul_((float)urow[r], inv_beta), -alpha), (float)acc);
        urow[r] = (T)un;
        ss += un * un;
      } else if (mode == MODE_PLAIN) {
        urow[r] = acc;
      } else if (vo + r < B.exp_m_data) {  // the score covers the data rows only (SLR:500-525)
        const float pred = B.clip_pred ? fmaxf((float)acc, 0.f) : (float)acc;
        const float bv = brow[r];
        ss += pred * pred; s_pb += pred * bv; s_bb += bv * bv;
      }
    }
  }
  if (!TRF) {
    if (mode == MODE_LSMR) {
      const float tot = block_sum(ss, red);
      if (threadIdx.x == 0) B.part_u[view * ppv + sub] = tot;
    } else if (mode == MODE_SCORE) {
      const float t0 = block_sum(s_pb, red), t1 = block_sum(ss, red), t2 = block_sum(s_bb, red);
      if (threadIdx.x == 0) {
        B.part_s[3 * (view * ppv + sub)] = t0;
        B.part_s[3 * (view * ppv + sub) + 1] = t1;
        B.part_s[3 * (view * ppv + sub) + 2] = t2;
      }
    }
  }
}

// Adjoint: vt[c][g] = sum over the transpose list of voxel g of w * rows[row] (* inv_beta in the LSMR modes, like
// k_adj_tie); the adjoint kernels add vt to the symmetry part.  One warp per voxel entry g = p*L3P + z (fixed
// lane-strided order + xor tree: deterministic).
template <typename T, bool TRF>
__global__ void __launch_bounds__(HB2_BLOCK) k_adj_csc(BD B, TD Tt, const T* __restrict__ rows, T* __restrict__ vt, int mode) {
  const int c = blockIdx.y;
  if (B.cand_tie_count[c] == 0) return;
  if (!tie_active<TRF>(B, Tt, c, mode, true)) return;
  const int g = blockIdx.x * (HB2_BLOCK / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;  // one warp per voxel
  if (g >= B.npad) return;
  T ib = (T)1;
  if (!TRF && mode != MODE_PLAIN) ib = (T)B.st[c].inv_beta;
  const T* __restrict__ ub = rows + B.cand_uoff[c];
  T acc = (T)0;
  const int e0 = B.exp_cptr[g], e1 = B.exp_cptr[g + 1];
  for (int e = e0 + lane; e < e1; e += 32) {
    const T uv = ub[B.exp_crow[e]];
    const T val = TRF ? uv : (T)fmaf((float)uv, (float)ib, 0.f);
    acc += (T)B.exp_cw[e] * val;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) vt[(size_t)c * B.npad + g] = acc;
}

// ===========================================================================
// Trilinear symmetry rows (SLR:910-1138 + the pair loop 1221-1287).  Per ordered pair of symmetry operations and per
// mask voxel (np.nonzero order): the two images of the voxel, int() truncation corners, all 16 corners inside the grid
// and the mask, |dz|, |dy|, |dx| >= 3 between the corner origins, first-seen-wins on the ROUNDED image pair (the
// reference's pair_ids set), then 8 + 8 trilinear weights -- with the reference's corner-110 weight xf*yf*(1-xf)
// (SLR:1089, 1125) reproduced.  Rounds run one pair at a time on the host (early stop SLR:1286).
// ===========================================================================
struct LsymPair { double mi[5]; double zi; double mj[5]; double zj; };  // M00, M01, M10, M11, M22 of scipy + rise*h
struct LsymSetup {
  int n, ndisk, D3, L3, L3P;
  const int* rank_sym;            // [D3*D3] internal disk rank or -1
  const short2* disk_yx_sym;      // [ndisk] reference order -> (y, x)
  unsigned long long* tab_key;
  unsigned long long* tab_seq;
  unsigned long long tab_cap;
  int* tmp_a;                     // rounded image i (internal voxel index or -1), -2: no row
  int* tmp_b;
  int* flag;
  int* pos;
  int* overflow;
};

struct LsymImg { double X, Y, Z; int xi, yi, zi; bool ok; };
__device__ __forceinline__ LsymImg lsym_image(const double* M, double zs, int xc, int yc, int zc, int D3, int L3,
                                              const int* __restrict__ rank) {
  LsymImg r;
  const double x = (double)xc, y = (double)yc, z = (double)zc;
  // Rotation.apply(inverse=False): out = fma(M[r][2], z, fma(M[r][1], y, M[r][0]*x)); M02 = M12 = M20 = M21 = 0
  r.X = __dadd_rn(__fma_rn(M[1], y, __dmul_rn(M[0], x)), (double)(D3 / 2));
  r.Y = __dadd_rn(__fma_rn(M[3], y, __dmul_rn(M[2], x)), (double)(D3 / 2));
  r.Z = __dadd_rn(__dadd_rn(__dmul_rn(M[4], z), (double)(L3 / 2)), zs);
  r.ok = false;
  if (!(r.Z > -1.0 && r.Z < (double)L3 && r.Y > -1.0 && r.Y < (double)D3 && r.X > -1.0 && r.X < (double)D3)) return r;
  r.zi = (int)r.Z; r.yi = (int)r.Y; r.xi = (int)r.X;
  if (r.zi + 1 > L3 - 1 || r.yi + 1 > D3 - 1 || r.xi + 1 > D3 - 1) return r;
  if (rank[r.yi * D3 + r.xi] < 0 || rank[r.yi * D3 + r.xi + 1] < 0 || rank[(r.yi + 1) * D3 + r.xi] < 0 ||
      rank[(r.yi + 1) * D3 + r.xi + 1] < 0)
    return r;
  r.ok = true;
  return r;
}
// mask_nonzero_indices_matrix[round(Z), round(Y), round(X)] with numpy's negative-index wrap (Z in (-1, -0.5) rounds to
// -1 = the last slice); -1 when the rounded voxel is outside the mask
__device__ __forceinline__ int lsym_rounded(const LsymImg& g, int D3, int L3, int L3P, const int* __restrict__ rank) {
  int zr = (int)rint(g.Z), yr = (int)rint(g.Y), xr = (int)rint(g.X);
  if (zr < 0) zr += L3;
  if (yr < 0) yr += D3;
  if (xr < 0) xr += D3;
  const int p = rank[yr * D3 + xr];
  return p < 0 ? -1 : p * L3P + zr;
}
__device__ __forceinline__ bool lsym_eval(const LsymSetup& Q, const LsymPair& P, int g, LsymImg& a, LsymImg& b) {
  const int z = g / Q.ndisk, p = g % Q.ndisk;
  const short2 yx = Q.disk_yx_sym[p];
  const int xc = yx.y - Q.D3 / 2, yc = yx.x - Q.D3 / 2, zc = z - Q.L3 / 2;
  a = lsym_image(P.mi, P.zi, xc, yc, zc, Q.D3, Q.L3, Q.rank_sym);
  if (!a.ok) return false;
  b = lsym_image(P.mj, P.zj, xc, yc, zc, Q.D3, Q.L3, Q.rank_sym);
  if (!b.ok) return false;
  if (abs(a.zi - b.zi) < 3 || abs(a.yi - b.yi) < 3 || abs(a.xi - b.xi) < 3) return false;
  return true;
}
__device__ __forceinline__ unsigned long long lsym_key(int ir, int jr) {
  const unsigned lo = (unsigned)(min(ir, jr) + 1), hi = (unsigned)(max(ir, jr) + 1);
  return ((unsigned long long)lo << 32) | hi;
}

__global__ void k_lsym_insert(LsymSetup Q, const LsymPair* __restrict__ pairs, int rnd) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Q.n) return;
  LsymImg a, b;
  if (!lsym_eval(Q, pairs[rnd], g, a, b)) { Q.tmp_a[g] = -2; Q.tmp_b[g] = -2; return; }
  const int ir = lsym_rounded(a, Q.D3, Q.L3, Q.L3P, Q.rank_sym), jr = lsym_rounded(b, Q.D3, Q.L3, Q.L3P, Q.rank_sym);
  Q.tmp_a[g] = ir; Q.tmp_b[g] = jr;
  const unsigned long long key = lsym_key(ir, jr);
  const unsigned long long seq = (unsigned long long)rnd * (unsigned long long)Q.n + (unsigned long long)g;
  unsigned long long slot = hash64(key) % Q.tab_cap;
  for (unsigned long long probe = 0; probe < Q.tab_cap; ++probe) {
    const unsigned long long old = atomicCAS(&Q.tab_key[slot], HB2_EMPTY_KEY, key);
    if (old == HB2_EMPTY_KEY || old == key) { atomicMin(&Q.tab_seq[slot], seq); return; }
    slot = slot + 1 == Q.tab_cap ? 0 : slot + 1;
  }
  atomicExch(Q.overflow, 1);
}
__global__ void k_lsym_check(LsymSetup Q, int rnd) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > Q.n) return;
  int keep = 0;
  if (g < Q.n && Q.tmp_a[g] != -2) {
    const unsigned long long key = lsym_key(Q.tmp_a[g], Q.tmp_b[g]);
    const unsigned long long seq = (unsigned long long)rnd * (unsigned long long)Q.n + (unsigned long long)g;
    unsigned long long slot = hash64(key) % Q.tab_cap;
    for (unsigned long long probe = 0; probe < Q.tab_cap; ++probe) {
      const unsigned long long kk = Q.tab_key[slot];
      if (kk == key) { keep = Q.tab_seq[slot] == seq; break; }
      if (kk == HB2_EMPTY_KEY) break;
      slot = slot + 1 == Q.tab_cap ? 0 : slot + 1;
    }
  }
  Q.flag[g] = keep;  // flag[n] = 0: the scan's last entry is the round's row count
}
__device__ __forceinline__ void lsym_weights(const LsymImg& g, double sgn, int L3P, int D3, const int* __restrict__ rank,
                                             int* __restrict__ col, float* __restrict__ w) {
  const double zf = __dsub_rn(g.Z, (double)g.zi), yf = __dsub_rn(g.Y, (double)g.yi), xf = __dsub_rn(g.X, (double)g.xi);
  const double mz = __dsub_rn(1.0, zf), my = __dsub_rn(1.0, yf), mx = __dsub_rn(1.0, xf);
  const int p00 = rank[g.yi * D3 + g.xi], p01 = rank[g.yi * D3 + g.xi + 1];
  const int p10 = rank[(g.yi + 1) * D3 + g.xi], p11 = rank[(g.yi + 1) * D3 + g.xi + 1];
  const double a = sgn * mz, c = sgn * zf, e = sgn * xf;  // unary minus binds first in the reference's expressions
  col[0] = p00 * L3P + g.zi;     w[0] = (float)__dmul_rn(__dmul_rn(a, my), mx);
  col[1] = p01 * L3P + g.zi;     w[1] = (float)__dmul_rn(__dmul_rn(a, my), xf);
  col[2] = p10 * L3P + g.zi;     w[2] = (float)__dmul_rn(__dmul_rn(a, yf), mx);
  col[3] = p11 * L3P + g.zi;     w[3] = (float)__dmul_rn(__dmul_rn(a, yf), xf);
  col[4] = p00 * L3P + g.zi + 1; w[4] = (float)__dmul_rn(__dmul_rn(c, my), mx);
  col[5] = p01 * L3P + g.zi + 1; w[5] = (float)__dmul_rn(__dmul_rn(c, my), xf);
  col[6] = p10 * L3P + g.zi + 1; w[6] = (float)__dmul_rn(__dmul_rn(e, yf), mx);   // SLR:1089/1125: xf*yf*(1-xf)
  col[7] = p11 * L3P + g.zi + 1; w[7] = (float)__dmul_rn(__dmul_rn(e, yf), zf);   // xf*yf*zf
}
__global__ void k_lsym_emit(LsymSetup Q, const LsymPair* __restrict__ pairs, int rnd, int row0, int cap_rows,
                            int* __restrict__ col, float* __restrict__ w) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Q.n || !Q.flag[g]) return;
  const int row = row0 + Q.pos[g];
  if (row >= cap_rows) { atomicExch(Q.overflow, 2); return; }
  LsymImg a, b;
  lsym_eval(Q, pairs[rnd], g, a, b);
  lsym_weights(a, 1.0, Q.L3P, Q.D3, Q.rank_sym, col + (size_t)row * 16, w + (size_t)row * 16);
  lsym_weights(b, -1.0, Q.L3P, Q.D3, Q.rank_sym, col + (size_t)row * 16 + 8, w + (size_t)row * 16 + 8);
}
__global__ void k_lsym_ptr(int m0, int nnz0, int ms, int* __restrict__ ptr, int* __restrict__ erow) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t <= ms) ptr[m0 + t] = nnz0 + 16 * t;
  if (t < ms * 16) erow[nnz0 + t] = m0 + t / 16;
}

// ===========================================================================
// Merging duplicate entries of the explicit rows.  A trilinear row touches every voxel around the ray several times
// (consecutive samples share corners): the reference accumulates them per voxel (row_tmp dict, SLR:1478-1496), here
// the unmerged row has ~1.8x the entries.  One CTA per row: bitonic sort of (column, original position) in shared
// memory -- unique 64-bit keys, so the order and therefore the float32 sums are deterministic -- then one entry per run.
// pass 0 sorts the row in place and counts its distinct columns, pass 1 (after a scan of the counts) writes them.
// ===========================================================================
#define HB2_MERGE_THREADS 256
template <int PASS>
__global__ void __launch_bounds__(HB2_MERGE_THREADS) k_exp_merge(int m, int maxe, const int* __restrict__ ptr, int* __restrict__ col,
                                                                  float* __restrict__ w, int* __restrict__ ucnt,
                                                                  const int* __restrict__ nptr, int* __restrict__ ncol,
                                                                  float* __restrict__ nw, int* __restrict__ nerow) {
  extern __shared__ __align__(16) unsigned char msm[];
  const int r = blockIdx.x;
  if (r >= m) return;
  const int e0 = ptr[r], n = ptr[r + 1] - e0;
  if (PASS == 0) {
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(msm);
    float* ws = reinterpret_cast<float*>(msm + (size_t)maxe * sizeof(unsigned long long));
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += HB2_MERGE_THREADS) {
      keys[i] = i < n ? (((unsigned long long)(unsigned)col[e0 + i] << 32) | (unsigned)i) : ~0ull;
      if (i < n) ws[i] = w[e0 + i];
    }
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < np2; i += HB2_MERGE_THREADS) {
          const int l = i ^ j;
          if (l > i) {
            const unsigned long long a = keys[i], b = keys[l];
            const bool up = (i & k) == 0;
            if ((a > b) == up) { keys[i] = b; keys[l] = a; }
          }
        }
        __syncthreads();
      }
    // write the row back sorted; count the distinct columns
    int heads = 0;
    for (int i = threadIdx.x; i < n; i += HB2_MERGE_THREADS) {
      const unsigned long long kq = keys[i];
      const int c = (int)(kq >> 32);
      col[e0 + i] = c;
      w[e0 + i] = ws[(unsigned)(kq & 0xFFFFFFFFull)];
      heads += (i == 0 || (int)(keys[i - 1] >> 32) != c) ? 1 : 0;
    }
    __shared__ int s_red[HB2_MERGE_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) heads += __shfl_xor_sync(0xffffffffu, heads, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = heads;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int q = 0; q < HB2_MERGE_THREADS / 32; ++q) tot += s_red[q];
      ucnt[r] = tot;
    }
  } else {
    // the row is sorted by (column, original position): thread i owning the head of a run sums it in that order
    int* rank = reinterpret_cast<int*>(msm);  // exclusive count of heads before position i
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int o0 = nptr[r];
    for (int base = 0; base < n; base += HB2_MERGE_THREADS) {
      const int i = base + threadIdx.x;
      const bool head = i < n && (i == 0 || col[e0 + i - 1] != col[e0 + i]);
      // block-wide exclusive scan of the head flags (ballot per warp + warp totals)
      const unsigned bal = __ballot_sync(0xffffffffu, head);
      const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
      __shared__ int s_wtot[HB2_MERGE_THREADS / 32];
      if (lane == 0) s_wtot[wid] = __popc(bal);
      __syncthreads();
      int off = s_base;
      for (int q = 0; q < wid; ++q) off += s_wtot[q];
      off += __popc(bal & ((1u << lane) - 1u));
      if (head) {
        const int c = col[e0 + i];
        float s = w[e0 + i];
        for (int q = i + 1; q < n && col[e0 + q] == c; ++q) s += w[e0 + q];
        ncol[o0 + off] = c; nw[o0 + off] = s; nerow[o0 + off] = r;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int tot = 0;
        for (int q = 0; q < HB2_MERGE_THREADS / 32; ++q) tot += s_wtot[q];
        s_base += tot;
      }
      __syncthreads();
    }
    (void)rank;
  }
}
