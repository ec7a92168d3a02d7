// Matrix-free trilinear data rows (build_A_data_matrix with interpolation "linear", SLR:1403-1510, in the grid-search
// case tilt = psi = dy = 0, scale2d_to_3d = 1).  The reference's row of (symmetry copy, image column k, ray j) is
//     sum_i  [(1 - zf) x[zi] + zf x[zi + 1]]  (x)  [(1 - yf)(1 - xf), (1 - yf) xf, yf (1 - xf), yf xf] at (yi, xi)
// over the depth samples i whose 8 corners lie inside the mask.  With tilt = psi = dy = 0 the slice pair (zi, zf)
// depends on the column k only and the in-plane footprint on (view angle, j, i) only, so the row factors into
//     row(k, j) = a_k P[zi_k][j] + b_k P[zi_k + 1][j],     P[z][j] = sum_e w_e x[z][p_e]
// with the MERGED in-plane footprint of the ray (p_e, w_e): a voxel met by several consecutive samples carries the
// float64 sum of their (1 - yf)(1 - xf)-type products, accumulated in sample order like the reference's row dict
// (SLR:1478-1496).  The footprints depend on the view angle only and are shared by all candidates of a batch, like the
// nearest-neighbour maps: the explicit matrix (192 M entries per candidate at 256 x 256) never exists.
//
//   * maps: k_bil_T builds, per (map, voxel), the <= KB rays that touch the voxel and their weights (voxel-driven, the
//     reference's float64 coordinate arithmetic: bil_sample == k_exp_rows); the forward lists are their transpose,
//     sorted by voxel rank per ray (deterministic).
//   * EXACT maps: where a sample coordinate is integer-valued (angle 0 / 90 / 180 / 270; integer h * rise) the int()
//     truncation follows the last-bit noise of the reference's coordinate tables (SLR:1712-1719), which depends on
//     (image column k, sample i).  Bilinear weights are continuous across that flip, but the 8-corner validity test is
//     not, so such columns get their own map built from the table rows of THAT column (hb2_bilinear_map.xrow / zrow);
//     the host turns them into single-column views.
//   * views: slot t of a view holds the rows of one image column; canonical form
//         t = 0 : row = a_0 P[0] + b_0 P[1]          (columns with Z in (-1, 0): int() truncates toward zero, zf < 0)
//         t >= 1: row = a_t P[t - 1] + b_t P[t]
//     so every slice index is static in the kernels.
//   * trilinear symmetry rows (k_lsym_*, hb2_explicit.cuh) stay explicit: 16 (column, weight) entries per row, per
//     candidate, plus their transpose lists.
// Inside the batch all of these are "pseudo views" (BD::view_tie >= 0): the nearest-neighbour projector kernels skip
// them, k_fwd_bil / k_fwd_lsym produce their rows and partial sums, k_adj_bil / k_adj_lsym their contribution to
// A^T u (BD::vtie / vtie64), which the adjoint kernels add -- LSMR / TRF state machines, norms and score are unchanged.
#pragma once
#include "hb2_explicit.cuh"

struct BilMap {  // == hb2_bilinear_map (include/helicon_b200.h)
  double m00, m01, m10, m11, m22, zshift;
  int32_t xrow, zrow;
};

struct BilS { int xi, yi; double xf, yf, X, Y; bool ok, nearx, neary; };

// One depth sample of ray j exactly as k_exp_rows computes it (rot_yx = identity, dy = 0, s = 1).
__device__ __forceinline__ BilS bil_sample(const BilMap& M, const double* __restrict__ xrows, const double* __restrict__ zrows,
                                           int D2, int L3, const int* __restrict__ rank, int j, int i) {
  BilS r;
  r.ok = false; r.nearx = r.neary = false; r.xi = r.yi = 0; r.xf = r.yf = 0.0;
  const int c0 = D2 / 2;
  const double x0 = M.xrow >= 0 ? xrows[(size_t)M.xrow * D2 + i] : -(double)(i - c0);
  const double y0 = (double)(j - c0);
  const double X = __dadd_rn(__fma_rn(M.m10, y0, __dmul_rn(M.m00, x0)), (double)c0);
  const double Y = __dadd_rn(__fma_rn(M.m11, y0, __dmul_rn(M.m01, x0)), (double)c0);
  r.X = X; r.Y = Y;
  if (!(Y > -1.0 && Y < (double)D2 && X > -1.0 && X < (double)D2)) return r;
  r.nearx = fabs(X - rint(X)) < 1e-9; r.neary = fabs(Y - rint(Y)) < 1e-9;
  if (M.zrow >= 0) {  // per-sample slice validity from the reference's z table (integer h * rise)
    const double z2 = __dmul_rn(M.m22, zrows[(size_t)M.zrow * D2 + i]);
    const double Z = __dadd_rn(__dsub_rn(z2, M.zshift), (double)(L3 / 2));
    if (!(Z > -1.0 && Z < (double)L3)) return r;
    if ((int)Z + 1 > L3 - 1) return r;
  }
  const int xi = (int)X, yi = (int)Y;
  if (yi + 1 > D2 - 1 || xi + 1 > D2 - 1) return r;
  if (rank[yi * D2 + xi] < 0 || rank[yi * D2 + xi + 1] < 0 || rank[(yi + 1) * D2 + xi] < 0 || rank[(yi + 1) * D2 + xi + 1] < 0)
    return r;
  r.xi = xi; r.yi = yi;
  r.xf = __dsub_rn(X, (double)xi); r.yf = __dsub_rn(Y, (double)yi);
  r.ok = true;
  return r;
}

// rayvalid[m][j] = the ray has a valid sample (SLR:1496 has_projection_data); tie[m] = samples with a coordinate within
// 1e-9 of an integer WHOSE VALIDITY DEPENDS ON THE int() DECISION (the four corners of one choice lie inside the mask,
// those of the other do not): there the reference follows the last-bit noise of its coordinate tables and the host
// replaces the view by exact per-column maps.  Everywhere else the flip is invisible -- bilinear weights are continuous
// across an integer coordinate (e.g. the sample at the rotation centre, integer-valued at every angle).
__global__ void k_bil_rayvalid(int nM, int D2, int L3, const BilMap* __restrict__ maps, const double* __restrict__ xrows,
                               const double* __restrict__ zrows, const int* __restrict__ rank, uint8_t* __restrict__ rayvalid,
                               int* __restrict__ tie) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nM * D2 * D2) return;
  const int i = (int)(t % D2), j = (int)((t / D2) % D2), m = (int)(t / ((long long)D2 * D2));
  const BilS s = bil_sample(maps[m], xrows, zrows, D2, L3, rank, j, i);
  if (s.ok) rayvalid[(size_t)m * D2 + j] = 1;  // benign race: all writers store 1
  if (s.nearx || s.neary) {
    const int xr = (int)rint(s.X), yr = (int)rint(s.Y);
    const int xa0 = s.nearx ? max(xr - 1, 0) : (int)s.X, xa1 = s.nearx ? max(xr, 0) : (int)s.X;
    const int ya0 = s.neary ? max(yr - 1, 0) : (int)s.Y, ya1 = s.neary ? max(yr, 0) : (int)s.Y;
    bool any = false, all = true;
    for (int yy = ya0; yy <= ya1; ++yy)
      for (int xx = xa0; xx <= xa1; ++xx) {
        const bool v = yy + 1 <= D2 - 1 && xx + 1 <= D2 - 1 && rank[yy * D2 + xx] >= 0 && rank[yy * D2 + xx + 1] >= 0 &&
                       rank[(yy + 1) * D2 + xx] >= 0 && rank[(yy + 1) * D2 + xx + 1] >= 0;
        any = any || v; all = all && v;
      }
    if (any && !all) atomicAdd(&tie[m], 1);
  }
}

// Transposed maps, voxel-driven.  PASS 0: kmax = max rays per (map, voxel), fcount[m][j / 2] += 1 per ray PAIR the voxel
// belongs to (the forward lists are kept per pair of adjacent rays: a voxel met by both rays is gathered once);
// PASS 1: Tj / Tw[(m*KB + k)*apitch + aslot[p]] (0xFFFF = empty).  Rays ascending, samples ascending inside a ray.
template <int PASS>
__global__ void k_bil_T(int nM, int D2, int L3, int ndisk, int apitch, int KB, const BilMap* __restrict__ maps,
                        const double* __restrict__ xrows, const double* __restrict__ zrows, const int* __restrict__ rank,
                        const short2* __restrict__ disk_yx, const int* __restrict__ aslot, uint16_t* __restrict__ Tj,
                        float* __restrict__ Tw, int* __restrict__ kmax, int* __restrict__ fcount) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nM * ndisk) return;
  const int p = (int)(t % ndisk), m = (int)(t / ndisk);
  const BilMap M = maps[m];
  const int c0 = D2 / 2;
  const short2 yx = disk_yx[p];
  const int vx = yx.y, vy = yx.x;
  const double dx = (double)(vx - c0), dy = (double)(vy - c0);
  // inverse of (X - c0, Y - c0) = (m00 x0 + m10 y0, m01 x0 + m11 y0): the transpose (a rotation)
  const double x0 = M.m00 * dx + M.m01 * dy, y0 = M.m10 * dx + M.m11 * dy;
  const double is = c0 - x0, js = c0 + y0, rad = 1.4143 + 0.02;
  const int i0 = max(0, (int)ceil(is - rad)), i1 = min(D2 - 1, (int)floor(is + rad));
  const int j0 = max(0, (int)ceil(js - rad)), j1 = min(D2 - 1, (int)floor(js + rad));
  int cnt = 0, last_pair = -1;
  const int NP = (D2 + 1) / 2;
  for (int j = j0; j <= j1; ++j) {
    double w = 0.0;
    bool hit = false;
    for (int i = i0; i <= i1; ++i) {
      const BilS s = bil_sample(M, xrows, zrows, D2, L3, rank, j, i);
      if (!s.ok) continue;
      const int cx = vx - s.xi, cy = vy - s.yi;
      if (cx < 0 || cx > 1 || cy < 0 || cy > 1) continue;
      const double wy = cy ? s.yf : __dsub_rn(1.0, s.yf), wx = cx ? s.xf : __dsub_rn(1.0, s.xf);
      w = __dadd_rn(w, __dmul_rn(wy, wx));
      hit = true;
    }
    if (hit && w != 0.0) {
      if (PASS == 0) {
        if ((j >> 1) != last_pair) { atomicAdd(&fcount[(size_t)m * NP + (j >> 1)], 1); last_pair = j >> 1; }
      } else if (cnt < KB) {
        Tj[((size_t)m * KB + cnt) * apitch + aslot[p]] = (uint16_t)j;
        Tw[((size_t)m * KB + cnt) * apitch + aslot[p]] = (float)w;
      }
      ++cnt;
    }
  }
  if (PASS == 0) {
    if (cnt > 0) atomicMax(kmax, cnt);
  } else {
    for (int k = cnt; k < KB; ++k) Tj[((size_t)m * KB + k) * apitch + aslot[p]] = 0xFFFFu;
  }
}

// content hash of every map (transposed entries + ray validity), order-independent: the host merges exact per-column
// maps that came out identical (the table noise has one sign per side of the image centre) into one multi-column view
__global__ void k_bil_hash(int nM, int D2, int apitch, int KB, const uint16_t* __restrict__ Tj, const float* __restrict__ Tw,
                           const uint8_t* __restrict__ rayvalid, unsigned long long* __restrict__ out) {
  const int m = blockIdx.y;
  const long long per = (long long)KB * apitch;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long h = 0;
  if (e < per) {
    const unsigned j = Tj[(size_t)m * per + e];
    if (j != 0xFFFFu)
      h = hash64(((unsigned long long)e << 32) ^ ((unsigned long long)j << 48) ^ (unsigned long long)__float_as_uint(Tw[(size_t)m * per + e]));
  }
  if (e < D2 && rayvalid[(size_t)m * D2 + e]) h += hash64(0x9E3779B97F4A7C15ull + (unsigned long long)e);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
  if ((threadIdx.x & 31) == 0 && h) atomicAdd(&out[m], h);
}

// forward lists, one per (map, pair of adjacent rays 2P, 2P + 1): entry (voxel p, weight for ray 2P, weight for ray 2P + 1)
// at fptr[m*NP + P] + cursor; sorted by p afterwards
__global__ void k_bil_F_fill(int nM, int D2, int ndisk, int apitch, int KB, const int* __restrict__ aslot,
                             const uint16_t* __restrict__ Tj, const float* __restrict__ Tw, const int* __restrict__ fptr,
                             int* __restrict__ cursor, unsigned* __restrict__ key, float2* __restrict__ val) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nM * ndisk) return;
  const int p = (int)(t % ndisk), m = (int)(t / ndisk);
  const int NP = (D2 + 1) / 2;
  int pair = -1;
  float w0 = 0.f, w1 = 0.f;
  for (int k = 0; k <= KB; ++k) {  // the entries of a voxel are rays in ascending order
    unsigned j = 0xFFFFu;
    float w = 0.f;
    if (k < KB) {
      const size_t mi = ((size_t)m * KB + k) * apitch + aslot[p];
      j = Tj[mi]; w = Tw[mi];
    }
    const int pj = j == 0xFFFFu ? -2 : (int)(j >> 1);
    if (pj != pair) {
      if (pair >= 0) {
        const int e = fptr[(size_t)m * NP + pair] + atomicAdd(&cursor[(size_t)m * NP + pair], 1);
        key[e] = (unsigned)p; val[e] = make_float2(w0, w1);
      }
      pair = pj; w0 = w1 = 0.f;
      if (pj == -2) break;
    }
    if (j & 1u) w1 = w; else w0 = w;
  }
}
template <typename IdxT>
__global__ void k_bil_pack(long long n, const unsigned* __restrict__ key, IdxT* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = (IdxT)key[e];
}

// right-hand side of the bilinear views (the pseudo views were zeroed by k_build_rhs) + max(b) per candidate
__global__ void k_bil_rhs(BD B, const float* __restrict__ pix, int nviews, float* __restrict__ bmax_bits) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nviews * B.rows_per_view) return;
  const int view = (int)(t / B.rows_per_view), r = (int)(t % B.rows_per_view);
  const int map = B.bil_view_map[view];
  const int zm = r % B.ZMP, j = r / B.ZMP;
  const int c = B.view_cand[view];
  bool rowok = false;
  float val = 0.f;
  if (map >= 0) {
    const int k = B.bil_colk[(size_t)view * B.ZMP + zm];
    rowok = k >= 0 && B.bil_rayvalid[(size_t)map * B.D2 + j];
    if (rowok) {
      const uint8_t* pm = cand_mask(B, c);
      if (pm && !pm[(size_t)k * B.D2 + j]) rowok = false;
    }
    if (rowok) val = pix[(size_t)j * B.L2 + k];
    B.b[B.view_uoff[view] + r] = val;
  }
  int iv = (int)0x80000000;
  if (rowok) { iv = __float_as_int(val); iv = iv >= 0 ? iv : iv ^ 0x7fffffff; }
  const unsigned act = __activemask();
  const unsigned peers = __match_any_sync(act, c);
  const int mx = __reduce_max_sync(peers, iv);
  if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31) && mx != (int)0x80000000) atomicMax((int*)bmax_bits + c, mx);
}

template <typename T>
__device__ __forceinline__ void ld4(const T* __restrict__ p, T& a, T& b, T& c, T& d);
template <>
__device__ __forceinline__ void ld4<float>(const float* __restrict__ p, float& a, float& b, float& c, float& d) {
  const float4 q = *reinterpret_cast<const float4*>(p);
  a = q.x; b = q.y; c = q.z; d = q.w;
}
template <>
__device__ __forceinline__ void ld4<double>(const double* __restrict__ p, double& a, double& b, double& c, double& d) {
  const double2 q0 = *reinterpret_cast<const double2*>(p), q1 = *reinterpret_cast<const double2*>(p + 2);
  a = q0.x; b = q0.y; c = q1.x; d = q1.y;
}

// Forward of the bilinear views.  Grid (pseudo views, fwd_ppv); one warp per PAIR of adjacent rays (2P, 2P + 1): their
// footprints overlap in ~1.4 of ~2.4 voxels per column, so the pair's merged list (voxel, w_even, w_odd) has ~27 % fewer
// gathers than two lists.  Lane = (slice quad q, entry group g): lanes q*G .. q*G + G - 1 walk the list G entries at a
// time, one 128-bit gather of 4 slices feeds both rays; a shuffle tree over g leaves P[ray][4q .. 4q+3]; lane (ray, t)
// then blends and finishes column slot t.  Quads whose slices no used slot of the view needs (single-column views of the
// exact maps) issue no gathers.  Halton duplicates of a copy (identical rows, SLR:1559-1571) are served by the first
// copy's CTAs in the float32 modes (BD::view_dups).
template <typename IdxT, int Q, typename T, bool TRF>
__global__ void __launch_bounds__(HB2_BLOCK) k_fwd_bil(BD B, TD Tt, const T* __restrict__ src, T* __restrict__ rows, int mode) {
  const int view = B.tie_views[blockIdx.x];
  const int map = B.bil_view_map[view];
  if (map < 0) return;  // trilinear symmetry rows: k_fwd_lsym
  if (!TRF && B.view_dupof[view] >= 0) return;  // duplicate of an earlier view: written by that view's CTAs
  const int c = B.view_cand[view];
  __shared__ float red[HB2_BLOCK / 32];
  __shared__ T s_P[HB2_BLOCK / 32][2][16];
  __shared__ T s_a[16], s_b[16];
  __shared__ int s_colk[16];
  __shared__ unsigned s_qmask;
  const int ppv = B.fwd_ppv, sub = blockIdx.y;
  int dupv[HB2_MAXDUP];
#pragma unroll
  for (int d = 0; d < HB2_MAXDUP; ++d) dupv[d] = TRF ? -1 : B.view_dups[view * HB2_MAXDUP + d];
  if (!tie_active<TRF>(B, Tt, c, mode, false)) {
    if (!TRF && threadIdx.x == 0) {
#pragma unroll
      for (int d = -1; d < HB2_MAXDUP; ++d) {
        const int vw = d < 0 ? view : dupv[d < 0 ? 0 : d];
        if (vw < 0) continue;
        if (mode == MODE_LSMR) B.part_u[vw * ppv + sub] = 0.f;
        if (mode == MODE_SCORE) { B.part_s[3 * (vw * ppv + sub)] = 0.f; B.part_s[3 * (vw * ppv + sub) + 1] = 0.f; B.part_s[3 * (vw * ppv + sub) + 2] = 0.f; }
      }
    }
    return;
  }
  constexpr int L3P = 4 * Q, G = 32 / Q;
  constexpr int P2 = G >= 32 ? 32 : (G >= 16 ? 16 : (G >= 8 ? 8 : 4));
  const int D2 = B.D2, ZMP = B.ZMP, NP = (D2 + 1) / 2;
  if (threadIdx.x < 16) {
    const bool in = (int)threadIdx.x < ZMP;
    const int ck = in ? B.bil_colk[(size_t)view * ZMP + threadIdx.x] : -1;
    s_colk[threadIdx.x] = ck;
    s_a[threadIdx.x] = in ? (T)B.bil_ab[((size_t)view * ZMP + threadIdx.x) * 2] : (T)0;
    s_b[threadIdx.x] = in ? (T)B.bil_ab[((size_t)view * ZMP + threadIdx.x) * 2 + 1] : (T)0;
    // slices the used slots read: t = 0 -> {0, 1}, t >= 1 -> {t - 1, t}
    const int t = threadIdx.x;
    unsigned need = 0;
    if (ck >= 0) need = t == 0 ? 3u : (3u << (t - 1));
    need |= __shfl_xor_sync(0xffffu, need, 8); need |= __shfl_xor_sync(0xffffu, need, 4);
    need |= __shfl_xor_sync(0xffffu, need, 2); need |= __shfl_xor_sync(0xffffu, need, 1);
    if (t == 0) {
      unsigned qm = 0;
      for (int qq = 0; qq < Q; ++qq) qm |= ((need >> (4 * qq)) & 15u) ? (1u << qq) : 0u;
      s_qmask = qm;
    }
  }
  __syncthreads();
  const IdxT* __restrict__ Fp = (const IdxT*)B.bilF_p;
  const float2* __restrict__ Fw = B.bilF_w;
  const int* __restrict__ ptr = B.bilF_ptr + (size_t)map * NP;
  const uint8_t* __restrict__ rv = B.bil_rayvalid + (size_t)map * D2;
  const T* __restrict__ vsrc = src + (size_t)c * B.npad;
  T* urow = rows + B.view_uoff[view];
  const float* brow = B.b + B.view_uoff[view];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane / G, g = lane % G;
  const bool lane_on = q < Q && ((s_qmask >> q) & 1u);
  float alpha = 0.f, inv_beta = 0.f;
  if (!TRF) { alpha = B.st[c].alpha; inv_beta = B.st[c].inv_beta; }
  float ss = 0.f, s_pb = 0.f, s_bb = 0.f;
  const uint8_t* __restrict__ pm = cand_mask(B, c);
  for (int pr = sub * (HB2_BLOCK / 32) + warp; pr < NP; pr += ppv * (HB2_BLOCK / 32)) {
    const int j0 = 2 * pr;
    const bool v0 = rv[j0] != 0, v1 = j0 + 1 < D2 && rv[j0 + 1] != 0;
    if (!v0 && !v1) continue;  // warp-uniform
    const int e0 = ptr[pr], e1 = ptr[pr + 1];
    T a0 = (T)0, a1 = (T)0, a2 = (T)0, a3 = (T)0, b0 = (T)0, b1 = (T)0, b2 = (T)0, b3 = (T)0;
    if (lane_on) {
      int e = e0 + g;
      for (; e + 3 * G < e1; e += 4 * G) {  // four independent (rank, weights, gather) chains in flight
        const size_t p0 = Fp[e], p1 = Fp[e + G], p2 = Fp[e + 2 * G], p3 = Fp[e + 3 * G];
        const float2 w0 = Fw[e], w1 = Fw[e + G], w2 = Fw[e + 2 * G], w3 = Fw[e + 3 * G];
        T x0, x1, x2, x3, y0, y1, y2, y3, z0, z1, z2, z3, t0, t1, t2, t3;
        ld4<T>(vsrc + p0 * L3P + 4 * q, x0, x1, x2, x3);
        ld4<T>(vsrc + p1 * L3P + 4 * q, y0, y1, y2, y3);
        ld4<T>(vsrc + p2 * L3P + 4 * q, z0, z1, z2, z3);
        ld4<T>(vsrc + p3 * L3P + 4 * q, t0, t1, t2, t3);
        a0 += (T)w0.x * x0; a1 += (T)w0.x * x1; a2 += (T)w0.x * x2; a3 += (T)w0.x * x3;
        b0 += (T)w0.y * x0; b1 += (T)w0.y * x1; b2 += (T)w0.y * x2; b3 += (T)w0.y * x3;
        a0 += (T)w1.x * y0; a1 += (T)w1.x * y1; a2 += (T)w1.x * y2; a3 += (T)w1.x * y3;
        b0 += (T)w1.y * y0; b1 += (T)w1.y * y1; b2 += (T)w1.y * y2; b3 += (T)w1.y * y3;
        a0 += (T)w2.x * z0; a1 += (T)w2.x * z1; a2 += (T)w2.x * z2; a3 += (T)w2.x * z3;
        b0 += (T)w2.y * z0; b1 += (T)w2.y * z1; b2 += (T)w2.y * z2; b3 += (T)w2.y * z3;
        a0 += (T)w3.x * t0; a1 += (T)w3.x * t1; a2 += (T)w3.x * t2; a3 += (T)w3.x * t3;
        b0 += (T)w3.y * t0; b1 += (T)w3.y * t1; b2 += (T)w3.y * t2; b3 += (T)w3.y * t3;
      }
      for (; e < e1; e += G) {
        const size_t p0 = Fp[e];
        const float2 w0 = Fw[e];
        T x0, x1, x2, x3;
        ld4<T>(vsrc + p0 * L3P + 4 * q, x0, x1, x2, x3);
        a0 += (T)w0.x * x0; a1 += (T)w0.x * x1; a2 += (T)w0.x * x2; a3 += (T)w0.x * x3;
        b0 += (T)w0.y * x0; b1 += (T)w0.y * x1; b2 += (T)w0.y * x2; b3 += (T)w0.y * x3;
      }
    }
    // sum over the entry groups of a quad: lanes q*G + g, g < G (G = 32, 16, 10, 8); fixed tree -> deterministic
    if (G > P2) {
      const T c0 = __shfl_down_sync(0xffffffffu, a0, P2), c1 = __shfl_down_sync(0xffffffffu, a1, P2);
      const T c2 = __shfl_down_sync(0xffffffffu, a2, P2), c3 = __shfl_down_sync(0xffffffffu, a3, P2);
      const T d0 = __shfl_down_sync(0xffffffffu, b0, P2), d1 = __shfl_down_sync(0xffffffffu, b1, P2);
      const T d2 = __shfl_down_sync(0xffffffffu, b2, P2), d3 = __shfl_down_sync(0xffffffffu, b3, P2);
      if (q < Q && g + P2 < G) { a0 += c0; a1 += c1; a2 += c2; a3 += c3; b0 += d0; b1 += d1; b2 += d2; b3 += d3; }
    }
#pragma unroll
    for (int o = P2 / 2; o > 0; o >>= 1) {
      a0 += __shfl_down_sync(0xffffffffu, a0, o); a1 += __shfl_down_sync(0xffffffffu, a1, o);
      a2 += __shfl_down_sync(0xffffffffu, a2, o); a3 += __shfl_down_sync(0xffffffffu, a3, o);
      b0 += __shfl_down_sync(0xffffffffu, b0, o); b1 += __shfl_down_sync(0xffffffffu, b1, o);
      b2 += __shfl_down_sync(0xffffffffu, b2, o); b3 += __shfl_down_sync(0xffffffffu, b3, o);
    }
    if (q < Q && g == 0) {
      s_P[warp][0][4 * q] = a0; s_P[warp][0][4 * q + 1] = a1; s_P[warp][0][4 * q + 2] = a2; s_P[warp][0][4 * q + 3] = a3;
      s_P[warp][1][4 * q] = b0; s_P[warp][1][4 * q + 1] = b1; s_P[warp][1][4 * q + 2] = b2; s_P[warp][1][4 * q + 3] = b3;
    }
    __syncwarp();
    {  // lane = (ray of the pair, column slot)
      const int rr = lane >> 4, t = lane & 15;
      const int j = j0 + rr;
      const bool rok = rr == 0 ? v0 : v1;
      if (rok && t < ZMP && s_colk[t] >= 0 && !(pm && !pm[(size_t)s_colk[t] * D2 + j])) {
        const T lo = t == 0 ? s_P[warp][rr][0] : s_P[warp][rr][t - 1];
        const T hi = t == 0 ? s_P[warp][rr][1] : s_P[warp][rr][t];
        const T acc = s_a[t] * lo + s_b[t] * hi;
        const size_t ri = (size_t)j * ZMP + t;
        if (TRF) {
          urow[ri] = acc;
        } else if (mode == MODE_LSMR) {
          const float un = fadd_(fmul_(fmul_((float)urow[ri], inv_beta), -alpha), (float)acc);
          urow[ri] = (T)un;
#pragma unroll
          for (int d = 0; d < HB2_MAXDUP; ++d)
            if (dupv[d] >= 0) (rows + B.view_uoff[dupv[d]])[ri] = (T)un;  // identical row of the duplicate view
          ss += un * un;
        } else if (mode == MODE_PLAIN) {
          urow[ri] = acc;
#pragma unroll
          for (int d = 0; d < HB2_MAXDUP; ++d)
            if (dupv[d] >= 0) (rows + B.view_uoff[dupv[d]])[ri] = acc;
        } else {
          const float pred = B.clip_pred ? fmaxf((float)acc, 0.f) : (float)acc;
          const float bv = brow[ri];
          ss += pred * pred; s_pb += pred * bv; s_bb += bv * bv;
        }
      }
    }
    __syncwarp();
  }
  if (!TRF) {  // partials of the view and of the duplicates it serves (identical rows -> identical partial sums)
    if (mode == MODE_LSMR) {
      const float tot = block_sum(ss, red);
      if (threadIdx.x == 0) {
        B.part_u[view * ppv + sub] = tot;
#pragma unroll
        for (int d = 0; d < HB2_MAXDUP; ++d)
          if (dupv[d] >= 0) B.part_u[dupv[d] * ppv + sub] = tot;
      }
    } else if (mode == MODE_SCORE) {
      const float t0 = block_sum(s_pb, red), t1 = block_sum(ss, red), t2 = block_sum(s_bb, red);
      if (threadIdx.x == 0) {
#pragma unroll
        for (int d = -1; d < HB2_MAXDUP; ++d) {
          const int vw = d < 0 ? view : dupv[d < 0 ? 0 : d];
          if (vw < 0) continue;
          B.part_s[3 * (vw * ppv + sub)] = t0;
          B.part_s[3 * (vw * ppv + sub) + 1] = t1;
          B.part_s[3 * (vw * ppv + sub) + 2] = t2;
        }
      }
    }
  }
}

// Forward of the trilinear symmetry rows: one thread per row (16 entries, 128-bit loads of columns and weights).
// The rows of candidate c occupy its pseudo views after the bilinear ones; never scored.
template <typename T, bool TRF>
__global__ void __launch_bounds__(HB2_BLOCK) k_fwd_lsym(BD B, TD Tt, const T* __restrict__ src, T* __restrict__ rows, int mode) {
  const int view = B.tie_views[blockIdx.x];
  if (B.bil_view_map[view] >= 0) return;
  const int c = B.view_cand[view];
  __shared__ float red[HB2_BLOCK / 32];
  const int ppv = B.fwd_ppv, sub = blockIdx.y;
  const bool act = tie_active<TRF>(B, Tt, c, mode, false);
  if (!act || (!TRF && mode == MODE_SCORE)) {
    if (!TRF && threadIdx.x == 0) {
      if (mode == MODE_LSMR) B.part_u[view * ppv + sub] = 0.f;
      if (mode == MODE_SCORE) { B.part_s[3 * (view * ppv + sub)] = 0.f; B.part_s[3 * (view * ppv + sub) + 1] = 0.f; B.part_s[3 * (view * ppv + sub) + 2] = 0.f; }
    }
    return;
  }
  const int first = B.cand_view_begin[c] + B.bil_cand_nview[c];  // first symmetry pseudo view of the candidate
  const long long r_lo = (long long)(view - first) * B.rows_per_view;
  const int nrow = (int)max(0ll, min((long long)B.rows_per_view, (long long)B.ls_m[c] - r_lo));
  const T* __restrict__ vsrc = src + (size_t)c * B.npad;
  T* urow = rows + B.view_uoff[view];
  const int4* __restrict__ ent4 = reinterpret_cast<const int4*>(B.ls_ent + B.ls_eoff[c] + r_lo * 16);
  float alpha = 0.f, inv_beta = 0.f;
  if (!TRF) { alpha = B.st[c].alpha; inv_beta = B.st[c].inv_beta; }
  float ss = 0.f;
  // 8 lanes per row: lane l of a warp loads the l-th 16-byte piece (two entries) of 4 consecutive rows -- one coalesced
  // 512-byte request per warp instruction -- and a 3-step shuffle tree adds the 8 partial sums of a row (fixed order)
  const int sub8 = threadIdx.x & 7;
  constexpr int RPB = HB2_BLOCK / 8;  // rows per CTA pass
  for (int base = sub * RPB; base < nrow; base += ppv * RPB) {  // block-uniform trip count: the shuffles see full warps
    const int r = base + (int)(threadIdx.x >> 3);
    const bool rok = r < nrow;
    T acc = (T)0;
    if (rok) {
      const int4 cc = ent4[(size_t)r * 8 + sub8];  // two (column, weight) entries
      acc = (T)__int_as_float(cc.y) * vsrc[cc.x];
      acc += (T)__int_as_float(cc.w) * vsrc[cc.z];
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (rok && sub8 == 0) {
      if (TRF || mode == MODE_PLAIN) {
        urow[r] = acc;
      } else {
        const float un = fadd_(fmul_(fmul_((float)urow[r], inv_beta), -alpha), (float)acc);
        urow[r] = (T)un;
        ss += un * un;
      }
    }
  }
  if (!TRF && mode == MODE_LSMR) {
    const float tot = block_sum(ss, red);
    if (threadIdx.x == 0) B.part_u[view * ppv + sub] = tot;
  }
}

// Adjoint of the bilinear views, step 1: un-blend the rows of every (view, ray) into slice space,
//     ub[view][j][z] = scale * (b_z r[z] + a_{z+1} r[z+1])   (+ the slot-0 terms a_0 r[0] on slice 0, b_0 r[0] on slice 1),
// scale = 1/beta in the LSMR modes (x the multiplicity of a view that serves Halton duplicates; the duplicates are
// skipped -- float32 solver modes only).  One thread per ray; 2 x rows traffic, ~2 us per candidate-pass.
template <int Q, typename T, bool TRF>
__global__ void __launch_bounds__(HB2_BLOCK) k_bil_unblend(BD B, TD Tt, const T* __restrict__ rows, T* __restrict__ ub, int mode) {
  const int view = B.tie_views[blockIdx.x];
  const int map = B.bil_view_map[view];
  if (map < 0) return;
  const int c = B.view_cand[view];
  if (!tie_active<TRF>(B, Tt, c, mode, true)) return;
  const bool dedupe = !TRF && mode != MODE_PLAIN;
  if (dedupe && B.view_dupof[view] >= 0) return;
  constexpr int L3P = 4 * Q;
  __shared__ T s_a[20], s_b[20];
  const int ZMP = B.ZMP, D2 = B.D2;
  T scale = (T)1;
  if (!TRF && mode != MODE_PLAIN) scale = (T)B.st[c].inv_beta;
  if (dedupe) scale *= (T)B.view_mult[view];
  if (threadIdx.x < 20) {
    const bool in = (int)threadIdx.x < ZMP;
    s_a[threadIdx.x] = in ? scale * (T)B.bil_ab[((size_t)view * ZMP + threadIdx.x) * 2] : (T)0;
    s_b[threadIdx.x] = in ? scale * (T)B.bil_ab[((size_t)view * ZMP + threadIdx.x) * 2 + 1] : (T)0;
  }
  __syncthreads();
  const uint8_t* __restrict__ rv = B.bil_rayvalid + (size_t)map * D2;
  const T* __restrict__ src = rows + B.view_uoff[view];
  T* dst = ub + B.view_uoff[view];
  for (int j = threadIdx.x; j < D2; j += HB2_BLOCK) {
    if (!rv[j]) continue;  // rays without data appear in no transposed map
    T r[L3P + 1], o[L3P];
#pragma unroll
    for (int z4 = 0; z4 < L3P; z4 += 4) ld4<T>(src + (size_t)j * ZMP + z4, r[z4], r[z4 + 1], r[z4 + 2], r[z4 + 3]);
    r[L3P] = (T)0;
#pragma unroll
    for (int z = 0; z < L3P; ++z) o[z] = s_b[z] * r[z] + s_a[z + 1] * r[z + 1];
    o[0] = s_a[0] * r[0] + s_a[1] * r[1];  // slot 0 blends slices (0, 1): a_0 r[0] -> slice 0, b_0 r[0] -> slice 1
    o[1] += s_b[0] * r[0];
#pragma unroll
    for (int z4 = 0; z4 < L3P; z4 += 4) {
      if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(dst + (size_t)j * ZMP + z4) = make_float4(o[z4], o[z4 + 1], o[z4 + 2], o[z4 + 3]);
      } else {
        *reinterpret_cast<double2*>(dst + (size_t)j * ZMP + z4) = make_double2(o[z4], o[z4 + 1]);
        *reinterpret_cast<double2*>(dst + (size_t)j * ZMP + z4 + 2) = make_double2(o[z4 + 2], o[z4 + 3]);
      }
    }
  }
}

// Adjoint of the bilinear views, step 2: weighted gather of the un-blended rows.  Lane = (in-plane voxel p, slice quad
// q), 32 / Q voxels per warp; grid (ceil(ndisk / voxels per CTA), nc):
//     vt[c][p*L3P + 4q ..] = sum over the candidate's views and the <= KB (ray, weight) entries of voxel p of w * ub[view][ray][4q ..]
// The entries of a (view, voxel) are loaded together (independent 128-bit row loads in flight).
template <int Q, typename T, bool TRF>
__global__ void __launch_bounds__(HB2_BLOCK) k_adj_bil(BD B, TD Tt, const T* __restrict__ ub, T* __restrict__ vt, int mode) {
  const int c = blockIdx.y;
  if (B.cand_tie_count[c] == 0) return;
  if (!tie_active<TRF>(B, Tt, c, mode, true)) return;
  constexpr int L3P = 4 * Q, NV = 64, VPW = 32 / Q, KMAX = 4;
  __shared__ int s_map[NV];
  __shared__ long long s_uoff[NV];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vl = lane / Q, q = lane % Q;
  const int p = (blockIdx.x * (HB2_BLOCK / 32) + warp) * VPW + vl;
  const bool on = vl < VPW && p < B.ndisk;
  const int ZMP = B.ZMP, KB = B.bil_KB;
  const bool dedupe = !TRF && mode != MODE_PLAIN;
  T acc0 = (T)0, acc1 = (T)0, acc2 = (T)0, acc3 = (T)0;
  const int slot = on ? B.aslot[p] : 0;
  const int vb = B.cand_view_begin[c], nvw = B.bil_cand_nview[c];
  const size_t kstride = (size_t)B.apitch;
  for (int v0 = 0; v0 < nvw; v0 += NV) {
    const int nv = min(NV, nvw - v0);
    __syncthreads();
    if (threadIdx.x < nv) {
      const int view = vb + v0 + threadIdx.x;
      s_map[threadIdx.x] = (dedupe && B.view_dupof[view] >= 0) ? -1 : B.bil_view_map[view];
      s_uoff[threadIdx.x] = B.view_uoff[view];
    }
    __syncthreads();
    for (int v = 0; v < nv; ++v) {  // lanes without a voxel stay in the loop (warp votes below) with slot 0 and discard
      const int map = s_map[v];
      if (map < 0) continue;  // block-uniform
      const T* __restrict__ ubv = ub + s_uoff[v] + 4 * q;
      const uint16_t* __restrict__ tj = B.bilT_j + (size_t)map * KB * kstride + slot;
      const float* __restrict__ tw = B.bilT_w + (size_t)map * KB * kstride + slot;
      unsigned jj[KMAX];
      float ww[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        jj[k] = 0xFFFFu; ww[k] = 0.f;
        if (k < KB) { jj[k] = tj[(size_t)k * kstride]; ww[k] = tw[(size_t)k * kstride]; }
      }
      // entries are packed from k = 0: slots no voxel of the warp uses are skipped by a uniform branch
      bool use[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) use[k] = k < KB && (k < 2 || __any_sync(0xffffffffu, jj[k] != 0xFFFFu));
      T r0[KMAX], r1[KMAX], r2[KMAX], r3[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (use[k]) {
          r0[k] = r1[k] = r2[k] = r3[k] = (T)0;
          if (jj[k] != 0xFFFFu) ld4<T>(ubv + (size_t)jj[k] * ZMP, r0[k], r1[k], r2[k], r3[k]);
        }
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (use[k]) { const T w = (T)ww[k]; acc0 += w * r0[k]; acc1 += w * r1[k]; acc2 += w * r2[k]; acc3 += w * r3[k]; }
    }
  }
  if (!on) return;
  T* dst = vt + (size_t)c * B.npad + (size_t)p * L3P + 4 * q;
  const int z = 4 * q;
  dst[0] = z < B.L3 ? acc0 : (T)0; dst[1] = z + 1 < B.L3 ? acc1 : (T)0;
  dst[2] = z + 2 < B.L3 ? acc2 : (T)0; dst[3] = z + 3 < B.L3 ? acc3 : (T)0;
}

// Adjoint of the bilinear views, step 2, TILE path (the default when the views of a candidate fit the tables): one CTA =
// one 16 x 16 voxel tile (<= 256 voxels, one consumer thread each, all slices in registers), a producer warp streams, per
// stage of HB2_BILT_SV views, the tile's runs of the transposed maps (KB x 512 bytes of rays + KB x 1 KB of weights) and
// the contiguous window of un-blended rows the tile's voxels touch in that view ([jlo, jlo + nr) x L3P) into a ring of
// shared-memory stages by cp.async.bulk (TMA, full / empty mbarriers) -- the same structure as k_adj_tile, with weights.
// The inner loop reads shared memory only.  Addition order: views, then k (= k_adj_bil).
#define HB2_BILT_SV 2
#define HB2_BILT_NS 4
#define HB2_BILT_THREADS (HB2_BLOCK + 32)
#define HB2_BILT_MAXV 512
template <int NQT, int KT, typename T, bool TRF>
__global__ void __launch_bounds__(HB2_BILT_THREADS) k_adj_bil_tile(BD B, TD Tt, const T* __restrict__ ub, T* __restrict__ vt, int mode) {
  extern __shared__ __align__(128) unsigned char dsm[];
  const int c = blockIdx.y, tile = blockIdx.x;
  if (B.cand_tie_count[c] == 0) return;
  if (!tie_active<TRF>(B, Tt, c, mode, true)) return;
  __shared__ unsigned long long full_bar[HB2_BILT_NS], empty_bar[HB2_BILT_NS];
  __shared__ int s_map[HB2_BILT_MAXV];
  __shared__ uint16_t s_jlo[HB2_BILT_MAXV], s_nr[HB2_BILT_MAXV];
  constexpr int L3P = 4 * NQT;
  const int KB = KT > 0 ? KT : B.bil_KB;  // KT = 3 (the usual maximum of rays per voxel): the entry loop is unrolled
  const int ndisk_t = B.tile_begin[tile + 1] - B.tile_begin[tile];
  const int p = B.tile_begin[tile] + threadIdx.x;
  const bool producer = threadIdx.x >= HB2_BLOCK;
  const bool live = threadIdx.x < ndisk_t;
  const int vb = B.cand_view_begin[c], nv = B.bil_cand_nview[c];
  const int rmax = B.bil_rmax;
  const bool dedupe = !TRF && mode != MODE_PLAIN;
  // dynamic shared memory per (stage, view slot): [KB][256] rays (uint16), [KB][256] weights (float), [rmax*L3P] rows
  const size_t slot_bytes = (size_t)KB * HB2_BLOCK * (sizeof(uint16_t) + sizeof(float)) + (size_t)rmax * L3P * sizeof(T);
  for (int e = threadIdx.x; e < nv; e += HB2_BILT_THREADS) {
    const int view = vb + e;
    const int map = B.bil_view_map[view];
    const bool skip = dedupe && B.view_dupof[view] >= 0;
    s_map[e] = map;
    s_jlo[e] = skip ? (uint16_t)0xFFFFu : B.bil_tile_jlo[(size_t)map * B.ntile + tile];
    s_nr[e] = skip ? (uint16_t)0 : B.bil_tile_nr[(size_t)map * B.ntile + tile];
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < HB2_BILT_NS; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], HB2_BLOCK / 32); }
  }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncthreads();
  const int nstage = (nv + HB2_BILT_SV - 1) / HB2_BILT_SV;
  T acc[L3P];
#pragma unroll
  for (int i = 0; i < L3P; ++i) acc[i] = (T)0;
  if (producer) {
    const int w = threadIdx.x - HB2_BLOCK;
    for (int st = 0; st < nstage; ++st) {
      const int buf = st % HB2_BILT_NS;
      if (st >= HB2_BILT_NS) mbar_wait(&empty_bar[buf], (unsigned)(((st / HB2_BILT_NS) - 1) & 1));
      const int v = st * HB2_BILT_SV + w;
      const bool has = w < HB2_BILT_SV && v < nv && s_nr[min(v, nv - 1)] != 0;
      const unsigned nr = has ? s_nr[v] : 0;
      unsigned tot = has ? (unsigned)KB * HB2_BLOCK * (unsigned)(sizeof(uint16_t) + sizeof(float)) + nr * L3P * (unsigned)sizeof(T) : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      if (w == 0) mbar_expect_tx(&full_bar[buf], tot);
      __syncwarp();
      if (has) {
        unsigned char* sb = dsm + (size_t)(buf * HB2_BILT_SV + w) * slot_bytes;
        const size_t mrow = (size_t)s_map[v] * KB * B.apitch + (size_t)tile * HB2_BLOCK;
        for (int k = 0; k < KB; ++k) {
          bulk_g2s(sb + (size_t)k * HB2_BLOCK * sizeof(uint16_t), B.bilT_j + mrow + (size_t)k * B.apitch,
                   HB2_BLOCK * (unsigned)sizeof(uint16_t), &full_bar[buf]);
          bulk_g2s(sb + (size_t)KB * HB2_BLOCK * sizeof(uint16_t) + (size_t)k * HB2_BLOCK * sizeof(float),
                   B.bilT_w + mrow + (size_t)k * B.apitch, HB2_BLOCK * (unsigned)sizeof(float), &full_bar[buf]);
        }
        bulk_g2s(sb + (size_t)KB * HB2_BLOCK * (sizeof(uint16_t) + sizeof(float)),
                 ub + B.view_uoff[vb + v] + (size_t)s_jlo[v] * L3P, nr * L3P * (unsigned)sizeof(T), &full_bar[buf]);
      }
    }
  } else {
    for (int st = 0; st < nstage; ++st) {
      const int buf = st % HB2_BILT_NS;
      mbar_wait(&full_bar[buf], (unsigned)((st / HB2_BILT_NS) & 1));
      if (live) {
        const int nvs = min(HB2_BILT_SV, nv - st * HB2_BILT_SV);
#pragma unroll
        for (int w = 0; w < HB2_BILT_SV; ++w) {
          if (w >= nvs || s_nr[st * HB2_BILT_SV + w] == 0) continue;
          const int jl = (int)s_jlo[st * HB2_BILT_SV + w];
          const unsigned char* sb = dsm + (size_t)(buf * HB2_BILT_SV + w) * slot_bytes;
          const uint16_t* mp = reinterpret_cast<const uint16_t*>(sb) + threadIdx.x;
          const float* wp = reinterpret_cast<const float*>(sb + (size_t)KB * HB2_BLOCK * sizeof(uint16_t)) + threadIdx.x;
          const T* uw = reinterpret_cast<const T*>(sb + (size_t)KB * HB2_BLOCK * (sizeof(uint16_t) + sizeof(float)));
#pragma unroll
          for (int k = 0; k < (KT > 0 ? KT : 4); ++k) {
            if (KT == 0 && k >= KB) break;
            const unsigned j = mp[(size_t)k * HB2_BLOCK];
            if (j == 0xFFFFu) break;  // entries are packed from k = 0
            const T wk = (T)wp[(size_t)k * HB2_BLOCK];
            const T* r = uw + ((int)j - jl) * L3P;
#pragma unroll
            for (int z4 = 0; z4 < L3P; z4 += 4) {
              T x0, x1, x2, x3;
              ld4<T>(r + z4, x0, x1, x2, x3);
              acc[z4] += wk * x0; acc[z4 + 1] += wk * x1; acc[z4 + 2] += wk * x2; acc[z4 + 3] += wk * x3;
            }
          }
        }
      }
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty_bar[buf]);
    }
  }
  if (live && !producer) {
    T* dst = vt + (size_t)c * B.npad + (size_t)p * L3P;
#pragma unroll
    for (int z = 0; z < L3P; ++z) dst[z] = z < B.L3 ? acc[z] : (T)0;
  }
}

// Adjoint of the trilinear symmetry rows, ADDED to vt (run after k_adj_bil): one thread per voxel entry g = p*L3P + z
// walking its transpose list (~6 entries, sorted by row: deterministic).
template <typename T, bool TRF>
__global__ void __launch_bounds__(HB2_BLOCK) k_adj_lsym(BD B, TD Tt, const T* __restrict__ rows, T* __restrict__ vt, int mode) {
  const int c = blockIdx.y;
  if (B.cand_tie_count[c] == 0) return;
  if (B.ls_m[c] == 0) return;
  if (!tie_active<TRF>(B, Tt, c, mode, true)) return;
  const int g = blockIdx.x * HB2_BLOCK + threadIdx.x;
  if (g >= B.npad) return;
  T ib = (T)1;
  if (!TRF && mode != MODE_PLAIN) ib = (T)B.st[c].inv_beta;
  const T* __restrict__ ub = rows + B.cand_uoff[c] + (long long)B.bil_cand_nview[c] * B.rows_per_view;
  const int* __restrict__ cp = B.ls_cptr + (size_t)c * (B.npad + 1);
  const int2* __restrict__ ce = B.ls_cent + B.ls_ceoff[c];
  const int e0 = cp[g], e1 = cp[g + 1];
  if (e1 <= e0) return;
  T acc = (T)0;
  for (int e = e0; e < e1; ++e) {
    const int2 en = ce[e];
    acc += (T)__int_as_float(en.y) * ub[en.x];
  }
  vt[(size_t)c * B.npad + g] += acc * ib;
}

// trilinear symmetry rows of one candidate: pack (column, weight) pairs; transpose helpers
__global__ void k_ls_pack(long long n, const int* __restrict__ col, const float* __restrict__ w, int2* __restrict__ ent) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) ent[e] = make_int2(col[e], __float_as_int(w[e]));
}
__global__ void k_ls_keys(long long n, const int2* __restrict__ ent, int* __restrict__ key, int* __restrict__ id, int* __restrict__ cc) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int c = ent[e].x;
  key[e] = c; id[e] = (int)e;
  atomicAdd(&cc[c], 1);
}
__global__ void k_ls_gather(long long n, const int* __restrict__ order, const int2* __restrict__ ent, int2* __restrict__ cent) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) { const int o = order[e]; cent[e] = make_int2(o >> 4, ent[o].y); }
}
