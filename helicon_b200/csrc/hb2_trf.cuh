// Bounded branch of the solve: scipy.optimize.lsq_linear(method='trf', lsq_solver='lsmr')
// as the reference calls it (SLR:246-270): reflective trust-region iterations in
// float64 (scipy/optimize/_lsq/trf_linear.py:147-249, common.py) with an inner
// float64 LSMR on [A D; diag(sqrt(diag_h))].  Batched: every kernel covers all
// candidates and consults per-candidate activity flags; data-dependent decisions
// are taken by one-thread-per-candidate scalar kernels.
#pragma once
#include "hb2_kernels.cuh"

#define TRF_NRED 8  // reduction slots per CTA

struct Lsmr64 {  // lsmr.py:197-480, all float64
  double alpha, beta, normb;
  double zetabar, alphabar, rho, rhobar, cbar, sbar;
  double betadd, betad, rhodold, tautildeold, thetatilde, zeta, d;
  double normA2, maxrbar, minrbar, normA, normr, normar, normx, condA, rhotemp;
  double cf_hbar, cf_x, cf_h, inv_alpha, inv_beta;
  int minrbar_inf, itn, istop, active, skip_adj;
};

struct TrfState {
  double lb, ub, tol;
  double cost, g_norm, theta, lsmr_tol, p_dot_g, p_stride, r_stride_u, r_stride_l, r_stride, ag_stride;
  double cost_change;
  double coef_p, coef_r, coef_ag;  // step = D*(coef_p*PH + coef_r*RH + coef_ag*(-GH))
  double acc[TRF_NRED];            // last reduction (sums 0..5, min 6, max 7)
  double m0, m1, m2, r_value, p_value, ag_n0, ag_n1, ag_min, step_g;
  int active;      // outer loop still running
  int needed;      // bounded branch taken at all
  int status, nit, stop_next, full_step;
  Lsmr64 in;
  double ag_value;
  double trace[24][8];  // per outer iteration: cost, g_norm, inner itn, kind(0 full,1 p,2 r,3 ag), p/r/ag value, cost_change
};

struct TD {  // device buffers of the bounded branch
  TrfState* st;
  double *G, *D, *DH, *PH, *RH, *STEP, *V, *H, *HBAR, *XIN, *W, *UN;  // [nc][npad]
  double *R, *UM, *Y, *Y2;                                          // rows layout of u (u_total)
  double* red;      // [nc][red_slots][TRF_NRED]
  int red_slots;
  int* nactive;
  int* ninner;
};

HB2_HD void sym_ortho64_(double a, double b, double& c, double& s, double& r) {
  if (b == 0.0) { c = a > 0 ? 1.0 : (a < 0 ? -1.0 : 0.0); s = 0.0; r = fabs(a); }
  else if (a == 0.0) { c = 0.0; s = b > 0 ? 1.0 : -1.0; r = fabs(b); }
  else if (fabs(b) > fabs(a)) { double tau = a / b; s = (b > 0 ? 1.0 : -1.0) / sqrt(1 + tau * tau); c = s * tau; r = b / s; }
  else { double tau = b / a; c = (a > 0 ? 1.0 : -1.0) / sqrt(1 + tau * tau); s = c * tau; r = a / c; }
}
HB2_HD void lsmr64_init_(Lsmr64& S, double alpha, double beta) {
  S.alpha = alpha; S.beta = beta; S.normb = beta;
  S.zetabar = alpha * beta; S.alphabar = alpha; S.rho = 1; S.rhobar = 1; S.cbar = 1; S.sbar = 0;
  S.betadd = beta; S.betad = 0; S.rhodold = 1; S.tautildeold = 0; S.thetatilde = 0; S.zeta = 0; S.d = 0;
  S.normA2 = alpha * alpha; S.maxrbar = 0; S.minrbar = 0; S.minrbar_inf = 1; S.normA = sqrt(S.normA2);
  S.condA = 1; S.normx = 0; S.normr = beta; S.normar = alpha * beta; S.itn = 0; S.istop = 0; S.skip_adj = 0;
  S.cf_hbar = S.cf_x = S.cf_h = 0; S.rhotemp = 0;
  S.inv_alpha = alpha > 0 ? 1.0 / alpha : 1.0;
  S.inv_beta = beta > 0 ? 1.0 / beta : 0.0;
  S.active = (S.normar != 0.0 && beta != 0.0) ? 1 : 0;
}
HB2_HD void lsmr64_rotate_(Lsmr64& S, double alpha, double beta) {
  S.itn += 1; S.alpha = alpha; S.beta = beta;
  double chat = S.alphabar > 0 ? 1.0 : (S.alphabar < 0 ? -1.0 : 0.0), alphahat = fabs(S.alphabar);
  double rhoold = S.rho, c, s;
  sym_ortho64_(alphahat, beta, c, s, S.rho);
  double thetanew = s * alpha;
  S.alphabar = c * alpha;
  double rhobarold = S.rhobar, zetaold = S.zeta, thetabar = S.sbar * S.rho;
  S.rhotemp = S.cbar * S.rho;
  double cb, sb, rb;
  sym_ortho64_(S.cbar * S.rho, thetanew, cb, sb, rb);
  S.cbar = cb; S.sbar = sb; S.rhobar = rb;
  S.zeta = S.cbar * S.zetabar; S.zetabar = -S.sbar * S.zetabar;
  S.cf_hbar = -(thetabar * S.rho / (rhoold * rhobarold));
  S.cf_x = S.zeta / (S.rho * S.rhobar);
  S.cf_h = -(thetanew / S.rho);
  double betaacute = chat * S.betadd, betahat = c * betaacute;
  S.betadd = -s * betaacute;
  double thetatildeold = S.thetatilde, ct, st, rt;
  sym_ortho64_(S.rhodold, thetabar, ct, st, rt);
  S.thetatilde = st * S.rhobar; S.rhodold = ct * S.rhobar;
  S.betad = -st * S.betad + ct * betahat;
  S.tautildeold = (zetaold - thetatildeold * S.tautildeold) / rt;
  double taud = (S.zeta - S.thetatilde * S.tautildeold) / S.rhodold;
  S.normr = sqrt(S.d + (S.betad - taud) * (S.betad - taud) + S.betadd * S.betadd);
  S.normA2 += beta * beta; S.normA = sqrt(S.normA2); S.normA2 += alpha * alpha;
  S.maxrbar = fmax(S.maxrbar, rhobarold);
  if (S.itn > 1) { S.minrbar = S.minrbar_inf ? rhobarold : fmin(S.minrbar, rhobarold); S.minrbar_inf = 0; }
  double mn = S.minrbar_inf ? S.rhotemp : fmin(S.minrbar, S.rhotemp);
  S.condA = fmax(S.maxrbar, S.rhotemp) / mn;
  S.normar = fabs(S.zetabar);
  S.inv_alpha = alpha > 0 ? 1.0 / alpha : 1.0;
}
HB2_HD int lsmr64_test_(Lsmr64& S, double normx, double atol, double btol, double conlim, int maxiter) {
  S.normx = normx;
  double test1 = S.normr / S.normb;
  double test2 = (S.normA * S.normr) != 0 ? S.normar / (S.normA * S.normr) : INFINITY;
  double test3 = 1 / S.condA;
  double t1 = test1 / (1 + S.normA * normx / S.normb);
  double rtol = btol + atol * S.normA * normx / S.normb;
  double ctol = conlim > 0 ? 1 / conlim : 0;
  int istop = 0;
  if (S.itn >= maxiter) istop = 7;
  if (1 + test3 <= 1) istop = 6;
  if (1 + test2 <= 1) istop = 5;
  if (1 + t1 <= 1) istop = 4;
  if (test3 <= ctol) istop = 3;
  if (test2 <= atol) istop = 2;
  if (test1 <= rtol) istop = 1;
  S.istop = istop;
  return istop;
}

// gate: 0 outer-active, 1 inner-active, 2 outer-active and reflective part needed
__device__ __forceinline__ bool trf_gate(const TrfState& S, int gate) {
  if (!S.active) return false;
  if (gate == 1) return S.in.active != 0;
  if (gate == 2) return S.full_step == 0;
  return true;
}

// ---------------------------------------------------------------------------
// float64 operators (plain): rows <- A w ; w <- A^T rows.  Same maps as the
// float32 kernels; gathers in double.
// ---------------------------------------------------------------------------
template <typename IdxT>
__global__ void __launch_bounds__(HB2_BLOCK) k_fwd64_data(BD B, TD T, const double* __restrict__ src, double* __restrict__ dst,
                                                         int inner) {
  const int ntiles = (B.D2 + HB2_TILE_RAYS - 1) / HB2_TILE_RAYS;
  const int view = blockIdx.x / ntiles, tile = blockIdx.x % ntiles;
  const int c = B.view_cand[view];
  if (B.view_tie && B.view_tie[view] >= 0) return;  // tie views: k_fwd_tie<double>
  const TrfState& S = T.st[c];
  if (!trf_gate(S, inner)) return;
  __shared__ int s_colk[HB2_MAX_ZMC];
  for (int e = threadIdx.x; e < B.ZMC; e += HB2_BLOCK) s_colk[e] = B.colk[B.view_colbegin[view] + e];
  __syncthreads();
  const int a = B.view_angle[view];
  const int D2 = B.D2, L3 = B.L3, L3P = B.L3P, MC = B.MC, ZMP = B.ZMP;
  const IdxT* __restrict__ fm = (const IdxT*)B.fmap + (size_t)a * D2 * D2;
  const double* __restrict__ vsrc = src + (size_t)c * B.npad;
  double* urow = dst + B.view_uoff[view];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane & 3, sg = lane >> 2;
  const uint8_t* __restrict__ pm = cand_mask(B, c);
  for (int r = warp; r < HB2_TILE_RAYS; r += HB2_BLOCK / 32) {
    const int j = tile * HB2_TILE_RAYS + r;
    if (j >= D2) break;
    if (!B.rayvalid[a * D2 + j]) continue;
    const IdxT* __restrict__ fj = fm + (size_t)j * D2;
    for (int z0 = 0; z0 < L3P; z0 += 16) {
      // lane (sg, q): the 4 consecutive slices z0 + 4q ... + 3 of every 8th sample, two 128-bit loads per sample and
      // four samples in flight (same lane layout as the float32 kernel)
      double acc[4] = {0, 0, 0, 0};
      const int zq = z0 + 4 * q;
      if (zq < L3P) {
        const double* __restrict__ vb = vsrc + zq;
        for (int i0 = sg; i0 < D2; i0 += 32) {
          IdxT ids[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) ids[w] = (i0 + 8 * w < D2) ? fj[i0 + 8 * w] : Sent<IdxT>::v;
          double2 ta[4], tb[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            if (ids[w] != Sent<IdxT>::v) {
              const double2* vp = reinterpret_cast<const double2*>(vb + (size_t)ids[w] * L3P);
              ta[w] = __ldg(vp); tb[w] = __ldg(vp + 1);
            } else {
              ta[w] = make_double2(0.0, 0.0); tb[w] = make_double2(0.0, 0.0);
            }
          }
#pragma unroll
          for (int w = 0; w < 4; ++w)
            if (ids[w] != Sent<IdxT>::v) { acc[0] += ta[w].x; acc[1] += ta[w].y; acc[2] += tb[w].x; acc[3] += tb[w].y; }
        }
      }
#pragma unroll
      for (int o = 4; o < 32; o <<= 1)
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
      if (sg < 4) {
        const int z = zq + sg;
        if (z < L3) {
          const double sum = sg == 0 ? acc[0] : (sg == 1 ? acc[1] : (sg == 2 ? acc[2] : acc[3]));
          for (int mc = 0; mc < MC; ++mc) {
            const int zm = z * MC + mc;
            if (s_colk[zm] >= 0 && !(pm && !pm[(size_t)s_colk[zm] * D2 + j])) urow[(size_t)j * ZMP + zm] = sum;
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(HB2_BLOCK) k_fwd64_sym(BD B, TD T, const double* __restrict__ src, double* __restrict__ dst,
                                                        int inner) {
  const int c = blockIdx.y;
  const TrfState& S = T.st[c];
  if (!trf_gate(S, inner)) return;
  const int m = B.cand_msym[c];
  const double* __restrict__ vsrc = src + (size_t)c * B.npad;
  double* us = dst + B.cand_uoff[c] + B.cand_mdata[c];
  const int* __restrict__ sa = B.sym_a + B.cand_symoff[c];
  const int* __restrict__ sb = B.sym_b + B.cand_symoff[c];
  for (int r = blockIdx.x * HB2_BLOCK * 4 + threadIdx.x, qq = 0; qq < 4; ++qq, r += HB2_BLOCK)
    if (r < m) us[r] = __ldg(vsrc + sa[r]) - __ldg(vsrc + sb[r]);
}

// same lane layout as k_adj: one thread = (voxel p, slice quad q)
__global__ void __launch_bounds__(HB2_BLOCK) k_adj64(BD B, TD T, const double* __restrict__ rows, double* __restrict__ dst,
                                                    int inner) {
  const int c = blockIdx.y;
  const TrfState& S = T.st[c];
  if (!trf_gate(S, inner) || (inner == 1 && S.in.skip_adj)) return;
  const int L3 = B.L3, L3P = B.L3P, ZMP = B.ZMP, ndisk = B.ndisk, K = B.K, MC = B.MC;
  const int NQ = L3P >> 2;
  const int t = blockIdx.x * HB2_BLOCK + threadIdx.x;
  const int p = t / NQ, q = t - p * NQ;
  if (p >= ndisk) return;
  const int slot = B.aslot[p];
  const int z0 = 4 * q;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const int vb = B.cand_view_begin[c], nv = B.cand_view_count[c];
  for (int vi = 0; vi < nv; ++vi) {
    const int view = vb + vi;
    if (B.view_tie && B.view_tie[view] >= 0) continue;  // tie views: k_adj_tie<double> (BD::vtie64)
    const int a = __ldg(B.view_angle + view);
    const double* __restrict__ ub = rows + __ldg(B.view_uoff + view) + z0 * MC;
    const uint16_t* __restrict__ am = B.amap + (size_t)a * K * B.apitch + slot;
    for (int k = 0; k < K; ++k) {
      const uint16_t j = am[(size_t)k * B.apitch];
      if (j != 0xFFFFu) {
        const double* __restrict__ uj = ub + (size_t)j * ZMP;
        if (MC == 1) {
          const double2 r0 = __ldg(reinterpret_cast<const double2*>(uj));
          const double2 r1 = __ldg(reinterpret_cast<const double2*>(uj) + 1);
          acc[0] += r0.x; acc[1] += r0.y; acc[2] += r1.x; acc[3] += r1.y;
        } else {
#pragma unroll
          for (int zz = 0; zz < 4; ++zz)
            if (z0 + zz < L3)
              for (int mc = 0; mc < MC; ++mc) acc[zz] += __ldg(uj + zz * MC + mc);
        }
      }
    }
  }
  const int* __restrict__ ptr = B.csc_ptr + (size_t)c * (B.npad + 1);
  const int* __restrict__ ent = B.csc_ent + B.cand_cscoff[c];
  const double* __restrict__ us = rows + B.cand_uoff[c] + B.cand_mdata[c];
  double* vdst = dst + (size_t)c * B.npad + (size_t)p * L3P + z0;
  if (B.vtie64 && B.cand_tie_count[c] > 0) {
    const double* vt = B.vtie64 + (size_t)c * B.npad + (size_t)p * L3P + z0;
#pragma unroll
    for (int zz = 0; zz < 4; ++zz) acc[zz] += vt[zz];
  }
#pragma unroll
  for (int zz = 0; zz < 4; ++zz) {
    double a2 = acc[zz];
    if (z0 + zz < L3) {
      const int g = p * L3P + z0 + zz;
      for (int e = ptr[g]; e < ptr[g + 1]; ++e) {
        int en = ent[e];
        double val = __ldg(us + (en & 0x7fffffff));
        a2 += en < 0 ? -val : val;
      }
    }
    vdst[zz] = a2;
  }
}

// ---------------------------------------------------------------------------
// reductions: each CTA writes TRF_NRED partials to a fixed slot (sums 0..5, min
// 6, max 7); k_trf_reduce combines a candidate's slots in a fixed order into
// TrfState::acc (deterministic).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double block_minmax_d(double x, bool is_min, double* shd) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double y = __shfl_xor_sync(0xffffffffu, x, o);
    x = is_min ? fmin(x, y) : fmax(x, y);
  }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) shd[w] = x;
  __syncthreads();
  if (w == 0) {
    x = l < HB2_BLOCK / 32 ? shd[l] : (is_min ? INFINITY : -INFINITY);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double y = __shfl_xor_sync(0xffffffffu, x, o);
      x = is_min ? fmin(x, y) : fmax(x, y);
    }
  }
  __syncthreads();
  return x;
}
__device__ __forceinline__ void trf_store_partials(TD T, int c, int slot, double (&v)[TRF_NRED], double* shd) {
#pragma unroll
  for (int k = 0; k < TRF_NRED; ++k) {
    double x = k < 6 ? block_sum_d(v[k], shd) : block_minmax_d(v[k], k == 6, shd);
    if (threadIdx.x == 0) T.red[((size_t)c * T.red_slots + slot) * TRF_NRED + k] = x;
  }
}
__global__ void __launch_bounds__(HB2_BLOCK) k_trf_reduce(BD B, TD T, int nslots, int add, int gate, int start) {
  const int c = blockIdx.x;
  __shared__ double shd[HB2_BLOCK / 32];
  TrfState& S = T.st[c];
  if (!(start ? (S.needed != 0) : trf_gate(S, gate))) return;
  double v[TRF_NRED];
#pragma unroll
  for (int k = 0; k < TRF_NRED; ++k) v[k] = k < 6 ? 0.0 : (k == 6 ? INFINITY : -INFINITY);
  for (int s2 = threadIdx.x; s2 < nslots; s2 += HB2_BLOCK) {
    const double* p = T.red + ((size_t)c * T.red_slots + s2) * TRF_NRED;
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] += p[k];
    v[6] = fmin(v[6], p[6]); v[7] = fmax(v[7], p[7]);
  }
#pragma unroll
  for (int k = 0; k < TRF_NRED; ++k) {
    double x = k < 6 ? block_sum_d(v[k], shd) : block_minmax_d(v[k], k == 6, shd);
    if (threadIdx.x == 0) {
      if (add) S.acc[k] = k < 6 ? S.acc[k] + x : (k == 6 ? fmin(S.acc[k], x) : fmax(S.acc[k], x));
      else S.acc[k] = x;
    }
  }
}

__device__ __forceinline__ double step_to_bound(double x, double s, double lb, double ub) {
  if (s == 0.0) return INFINITY;  // common.py:372-399, one component
  return fmax((lb - x) / s, (ub - x) / s);
}

enum {
  EW_START = 0,   // reduction: min/max of x_lsq (in_bounds test, lsq_linear.py:328)
  EW_PREP,        // Coleman-Li scaling from (x, g): D = sqrt(v), DH = g*dv ; reduction: max |g v|   (trf_linear.py:183-199)
  EW_INNER_INIT,  // v~ = D*W (W = A^T r, u_n = 0) ; reduction: sum v~^2
  EW_INNER_UN,    // u~_n = sqrt(DH)*V - alpha*(u~_n*inv_beta) ; reduction: sum u~_n^2
  EW_INNER_V,     // v~ = (D*W + sqrt(DH)*u~_n)*inv_beta - beta*V ; reduction: sum v~^2
  EW_INNER_UPD,   // V = v~/alpha ; hbar, xin, h ; W = D*V ; reduction: sum xin^2
  EW_STEP1,       // PH = -XIN ; reductions: in_bounds(x + D*PH), min step_size_to_bound(x,p), p.g
  EW_STEP2,       // RH = reflected PH, PH *= p_stride ; W = D*RH ; reduction: min step_size_to_bound(x+p, r)
  EW_LOAD_PH,     // W = D*PH
  EW_DOTS,        // n-space dots of build_quadratic_1d / evaluate_quadratic
  EW_AG,          // W = D*(-GH) ; reductions: dots, min step_size_to_bound(x, ag)
  EW_MKSTEP,      // STEP = D*(cp*PH + cr*RH - cag*GH) ; W = STEP ; reduction: STEP.G
  EW_XUPD,        // X = make_strictly_feasible(X + STEP, rstep=0) ; W = X
  EW_FINAL        // xs = float32(X)   (SLR:270)
};

__global__ void __launch_bounds__(HB2_BLOCK) k_trf_ew(BD B, TD T, int op) {
  const int c = blockIdx.y;
  __shared__ double shd[HB2_BLOCK / 32];
  const TrfState& S = T.st[c];
  bool act;
  if (op == EW_START || op == EW_FINAL) act = S.needed != 0;
  else if (op == EW_INNER_UPD) act = S.active && (S.in.active || S.in.itn == 0);
  else if (op == EW_INNER_UN || op == EW_INNER_V) act = trf_gate(S, 1);
  else if (op == EW_STEP2 || op == EW_LOAD_PH || op == EW_DOTS || op == EW_AG) act = trf_gate(S, 2);
  else act = trf_gate(S, 0);
  if (op == EW_INNER_V && S.in.skip_adj) act = false;
  if (!act) return;
  const size_t o = (size_t)c * B.npad;
  double* X = B.x + o;
  double red[TRF_NRED] = {0, 0, 0, 0, 0, 0, INFINITY, -INFINITY};
  const double lb = S.lb, ub = S.ub;
  for (int i = blockIdx.x * HB2_BLOCK * 4 + threadIdx.x, qq = 0; qq < 4; ++qq, i += HB2_BLOCK) {
    if (i >= B.npad || (i % B.L3P) >= B.L3) continue;  // padded slices are not unknowns
    const size_t e = o + i;
    switch (op) {
      case EW_START: {
        double y = X[i];
        red[6] = fmin(red[6], y); red[7] = fmax(red[7], y);
      } break;
      case EW_PREP: {
        double x = X[i], g = T.G[e], v = 1.0, dv = 0.0;
        if (g < 0) { v = ub - x; dv = -1.0; }
        if (g > 0) { v = x - lb; dv = 1.0; }
        T.D[e] = sqrt(v); T.DH[e] = g * dv;
        red[7] = fmax(red[7], fabs(g * v));
      } break;
      case EW_INNER_INIT: {
        double vt = T.D[e] * T.W[e];
        T.UN[e] = 0.0; T.V[e] = vt; red[0] += vt * vt;
      } break;
      case EW_INNER_UN: {
        double un = sqrt(T.DH[e]) * T.V[e] - S.in.alpha * (T.UN[e] * S.in.inv_beta);
        T.UN[e] = un; red[0] += un * un;
      } break;
      case EW_INNER_V: {
        double vt = (T.D[e] * T.W[e] + sqrt(T.DH[e]) * T.UN[e]) * S.in.inv_beta - S.in.beta * T.V[e];
        T.V[e] = vt; red[0] += vt * vt;
      } break;
      case EW_INNER_UPD: {
        double vn = T.V[e];
        if (S.in.itn == 0 || !S.in.skip_adj) { vn *= S.in.inv_alpha; T.V[e] = vn; }
        if (S.in.itn == 0) { T.H[e] = vn; T.HBAR[e] = 0.0; T.XIN[e] = 0.0; }
        else {
          double ho = T.H[e];
          double hb = T.HBAR[e] * S.in.cf_hbar + ho;
          T.HBAR[e] = hb;
          double xn = T.XIN[e] + S.in.cf_x * hb;
          T.XIN[e] = xn;
          T.H[e] = ho * S.in.cf_h + vn;
          red[0] += xn * xn;
        }
        T.W[e] = T.D[e] * vn;
      } break;
      case EW_STEP1: {
        double ph = -T.XIN[e];
        T.PH[e] = ph;
        double x = X[i], p = T.D[e] * ph, xp = x + p;
        red[6] = fmin(red[6], fmin(xp - lb, ub - xp));
        red[5] += p * T.G[e];
        red[7] = fmax(red[7], -step_to_bound(x, p, lb, ub));
      } break;
      case EW_STEP2: {
        double ph = T.PH[e], x = X[i], d = T.D[e];
        double p = d * ph;
        double st = step_to_bound(x, p, lb, ub);
        double rh = (st == S.p_stride && p != 0.0) ? -ph : ph;  // hits -> reflect (trf_linear.py:97-99)
        T.RH[e] = rh;
        T.PH[e] = ph * S.p_stride;
        double xb = x + p * S.p_stride, r = d * rh;
        red[7] = fmax(red[7], -step_to_bound(xb, r, lb, ub));
        T.W[e] = r;
      } break;
      case EW_LOAD_PH: T.W[e] = T.D[e] * T.PH[e]; break;
      case EW_DOTS: {
        double ph = T.PH[e], rh = T.RH[e], dh = T.DH[e], gh = T.D[e] * T.G[e];
        red[0] += rh * dh * rh; red[1] += gh * rh; red[2] += ph * dh * rh; red[3] += gh * ph; red[4] += ph * dh * ph;
      } break;
      case EW_AG: {
        double d = T.D[e], g = T.G[e], gh = d * g, agh = -gh, ag = d * agh;
        red[0] += agh * T.DH[e] * agh; red[1] += gh * agh;
        red[7] = fmax(red[7], -step_to_bound(X[i], ag, lb, ub));
        T.W[e] = ag;
      } break;
      case EW_MKSTEP: {
        double d = T.D[e];
        double s2 = d * (S.coef_p * T.PH[e] + S.coef_r * T.RH[e] - S.coef_ag * (d * T.G[e]));
        T.STEP[e] = s2; T.W[e] = s2; red[0] += s2 * T.G[e];
      } break;
      case EW_XUPD: {
        double x = X[i];
        if (S.cost_change >= 0) {  // else: scipy's backtracking() hands back the OLD x (trf_linear.py:69-88)
          x = x + T.STEP[e];
          if (x <= lb) x = nextafter(lb, ub);
          if (x >= ub) x = nextafter(ub, lb);
          if (x < lb || x > ub) x = 0.5 * (lb + ub);
          X[i] = x;
        }
        T.W[e] = x;
      } break;
      case EW_FINAL: B.xs[e] = (float)X[i]; break;
    }
  }
  if (op == EW_LOAD_PH || op == EW_XUPD || op == EW_FINAL) return;
  trf_store_partials(T, c, blockIdx.x, red, shd);
}

// x <- make_strictly_feasible(reflective_transformation(x_lsq), rstep=0.1)  (trf_linear.py:150-151)
__global__ void __launch_bounds__(HB2_BLOCK) k_trf_start_apply(BD B, TD T) {
  const int c = blockIdx.y;
  const TrfState& S = T.st[c];
  if (!S.active) return;
  double* X = B.x + (size_t)c * B.npad;
  const double lb = S.lb, ub = S.ub, d = ub - lb;
  for (int i = blockIdx.x * HB2_BLOCK * 4 + threadIdx.x, qq = 0; qq < 4; ++qq, i += HB2_BLOCK) {
    if (i >= B.npad || (i % B.L3P) >= B.L3) continue;
    double y = X[i];
    double t = fmod(y - lb, 2 * d);
    if (t < 0) t += 2 * d;  // np.remainder takes the sign of the divisor (common.py:535)
    double x = lb + fmin(t, 2 * d - t);
    // make_strictly_feasible(rstep=0.1): common.py:440-464 with find_active_constraints (:401-437)
    double ld = x - lb, udist = ub - x;
    double lth = 0.1 * fmax(1.0, fabs(lb)), uth = 0.1 * fmax(1.0, fabs(ub));
    bool la = ld <= fmin(udist, lth), ua = udist <= fmin(ld, uth);
    if (la) x = lb + 0.1 * fmax(1.0, fabs(lb));
    if (ua) x = ub - 0.1 * fmax(1.0, fabs(ub));
    if (x < lb || x > ub) x = 0.5 * (lb + ub);
    X[i] = x;
    T.W[(size_t)c * B.npad + i] = x;
  }
}

// m-space kernels (rows of candidate c live at [cand_uoff, + mdata + msym)) ---------
enum { MW_RESID = 0, MW_INNER_U, MW_DOTS3, MW_DOT_YY };
__global__ void __launch_bounds__(HB2_BLOCK) k_trf_mw(BD B, TD T, int op) {
  const int c = blockIdx.y;
  __shared__ double shd[HB2_BLOCK / 32];
  const TrfState& S = T.st[c];
  bool act = op == MW_INNER_U ? trf_gate(S, 1) : (op == MW_DOTS3 ? trf_gate(S, 2) : trf_gate(S, 0));
  if (!act) return;
  const long long m = (long long)B.cand_mdata[c] + B.cand_msym[c];
  const size_t o = (size_t)B.cand_uoff[c];
  double red[TRF_NRED] = {0, 0, 0, 0, 0, 0, INFINITY, -INFINITY};
  for (long long r = (long long)blockIdx.x * HB2_BLOCK * 4 + threadIdx.x, qq = 0; qq < 4; ++qq, r += HB2_BLOCK) {
    if (r >= m) continue;
    const size_t e = o + r;
    switch (op) {
      case MW_RESID: {  // R = A x - b ; u~_m = R (right-hand side of the next inner solve) ; sum R^2
        double bb = r < B.cand_mdata[c] ? (double)B.b[e] : 0.0;
        double res = T.Y[e] - bb;
        T.R[e] = res; T.UM[e] = res; red[0] += res * res;
      } break;
      case MW_INNER_U: {
        double un = T.Y[e] - S.in.alpha * (T.UM[e] * S.in.inv_beta);
        T.UM[e] = un; red[0] += un * un;
      } break;
      case MW_DOTS3: {  // Y = A_h r_h, Y2 = A_h p_h
        double v1 = T.Y[e], u1 = T.Y2[e];
        red[0] += v1 * v1; red[1] += u1 * v1; red[2] += u1 * u1;
      } break;
      case MW_DOT_YY: { double v = T.Y[e]; red[0] += v * v; } break;
    }
  }
  trf_store_partials(T, c, blockIdx.x, red, shd);
}

// scalar decisions: one thread per candidate, reads TrfState::acc ---------------------
enum {
  SC_START = 0, SC_RESID0, SC_PREP, SC_INNER_INIT, SC_INNER_BETA, SC_INNER_ROT, SC_INNER_TEST, SC_STEP1, SC_STEP2,
  SC_SAVE_M, SC_QUAD, SC_SAVE_AG, SC_AG, SC_SAVE_SG, SC_CC, SC_RESID
};
__device__ __forceinline__ void min_quadratic_1d(double a, double b, double lo, double hi, double c, double& t, double& y) {
  // common.py:302-322
  double tt[3] = {lo, hi, 0};
  int n = 2;
  if (a != 0) {
    double ext = -0.5 * b / a;
    if (lo < ext && ext < hi) tt[n++] = ext;
  }
  t = tt[0]; y = tt[0] * (a * tt[0] + b) + c;
  for (int k = 1; k < n; ++k) {
    double yy = tt[k] * (a * tt[k] + b) + c;
    if (yy < y) { y = yy; t = tt[k]; }  // np.argmin keeps the first minimum
  }
}
__global__ void k_trf_scal(BD B, TD T, int op, double eps, int lsmr_maxiter, int max_iter) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= B.nc) return;
  TrfState& S = T.st[c];
  if (op == SC_START) { if (!S.needed) return; }
  else if (op == SC_INNER_BETA || op == SC_INNER_ROT || op == SC_INNER_TEST) { if (!trf_gate(S, 1)) return; }
  else if (op == SC_STEP2 || op == SC_SAVE_M || op == SC_QUAD || op == SC_SAVE_AG || op == SC_AG) { if (!trf_gate(S, 2)) return; }
  else if (!trf_gate(S, 0)) return;
  const double* r = S.acc;
  switch (op) {
    case SC_START: {
      bool inb = r[6] >= S.lb && r[7] <= S.ub;  // unconstrained solution already feasible (lsq_linear.py:328)
      S.active = inb ? 0 : 1; S.status = inb ? 3 : 0; S.nit = 0; S.stop_next = 0; S.full_step = 0;
      if (!inb) atomicAdd(T.nactive, 1);
    } break;
    case SC_RESID0: S.cost = 0.5 * r[0]; break;
    case SC_PREP: {
      S.g_norm = r[7];
      bool stop = S.stop_next != 0;
      if (S.g_norm < S.tol) { S.status = 1; stop = true; }
      if (!stop && S.nit >= max_iter) stop = true;
      if (stop) { S.active = 0; if (S.nit < max_iter) S.nit += 1; atomicSub(T.nactive, 1); break; }  // nit = iteration + 1
      S.nit += 1;
      if (S.nit <= 24) { S.trace[S.nit - 1][0] = S.cost; S.trace[S.nit - 1][1] = S.g_norm; }
      double eta = 1e-2 * fmin(0.5, S.g_norm);
      S.lsmr_tol = fmax(eps, fmin(0.1, eta * S.g_norm));
      S.theta = 1 - fmin(0.005, S.g_norm);
      S.in.beta = sqrt(2 * S.cost);
      S.in.inv_beta = S.in.beta > 0 ? 1.0 / S.in.beta : 0.0;
      S.in.itn = 0; S.in.active = 1; S.in.skip_adj = 0; S.in.alpha = 0; S.full_step = 0;
    } break;
    case SC_INNER_INIT: {  // V holds D*(A^T r) (not yet divided by beta)
      double nv = sqrt(r[0]);
      double beta = S.in.beta;
      lsmr64_init_(S.in, nv * (beta > 0 ? 1.0 / beta : 0.0), beta);
      S.in.inv_alpha = nv > 0 ? 1.0 / nv : 1.0;
      if (S.in.active) atomicAdd(T.ninner, 1);
    } break;
    case SC_INNER_BETA: {
      S.in.beta = sqrt(r[0]);
      if (S.in.beta > 0) { S.in.inv_beta = 1.0 / S.in.beta; S.in.skip_adj = 0; } else S.in.skip_adj = 1;
    } break;
    case SC_INNER_ROT: lsmr64_rotate_(S.in, S.in.skip_adj ? S.in.alpha : sqrt(r[0]), S.in.beta); break;
    case SC_INNER_TEST: {
      int istop = lsmr64_test_(S.in, sqrt(r[0]), S.lsmr_tol, S.lsmr_tol, 1e8, lsmr_maxiter);
      if (istop > 0) { S.in.active = 0; atomicSub(T.ninner, 1); }
    } break;
    case SC_STEP1: {
      S.p_dot_g = r[5];
      if (S.p_dot_g > 0) { S.status = -1; S.stop_next = 1; }
      S.full_step = r[6] >= 0.0 ? 1 : 0;
      S.p_stride = -r[7];
      if (S.full_step) { S.coef_p = 1.0; S.coef_r = 0.0; S.coef_ag = 0.0; }
    } break;
    case SC_STEP2: {
      double rsu = -r[7];
      S.r_stride_l = (1 - S.theta) * rsu;
      S.r_stride_u = rsu * S.theta;
    } break;
    case SC_SAVE_M: S.m0 = r[0]; S.m1 = r[1]; S.m2 = r[2]; break;
    case SC_QUAD: {  // build_quadratic_1d(A_h, g_h, r_h, s0=p_h, diag=diag_h), common.py:251-299
      double a = 0.5 * (S.m0 + r[0]);
      double b = r[1] + S.m1 + r[2];
      double cc = 0.5 * S.m2 + r[3] + 0.5 * r[4];
      if (S.r_stride_u > 0) min_quadratic_1d(a, b, S.r_stride_l, S.r_stride_u, cc, S.r_stride, S.r_value);
      else { S.r_value = INFINITY; S.r_stride = 0; }
      double th = S.theta;  // p_h *= theta ; evaluate_quadratic(A_h, g_h, p_h, diag_h)
      S.p_value = 0.5 * (th * th * S.m2 + th * th * r[4]) + th * r[3];
    } break;
    case SC_SAVE_AG: S.ag_n0 = r[0]; S.ag_n1 = r[1]; S.ag_min = -r[7]; break;
    case SC_AG: {
      double a = 0.5 * (r[0] + S.ag_n0), b = S.ag_n1;
      double ag_value;
      min_quadratic_1d(a, b, 0.0, S.ag_min * S.theta, 0.0, S.ag_stride, ag_value);
      S.ag_value = ag_value;
      if (S.p_value < S.r_value && S.p_value < ag_value) { S.coef_p = S.theta; S.coef_r = 0; S.coef_ag = 0; }
      else if (S.r_value < S.p_value && S.r_value < ag_value) { S.coef_p = 1.0; S.coef_r = S.r_stride; S.coef_ag = 0; }
      else { S.coef_p = 0; S.coef_r = 0; S.coef_ag = S.ag_stride; }
    } break;
    case SC_SAVE_SG: S.step_g = r[0]; break;
    case SC_CC: {
      S.cost_change = -(0.5 * r[0] + S.step_g);  // -evaluate_quadratic(A, g, step)
      if (S.nit <= 24) {
        double* t = S.trace[S.nit - 1];
        t[2] = S.in.itn;
        t[3] = S.full_step ? 0 : (S.coef_ag != 0 ? 3 : (S.coef_r != 0 ? 2 : 1));
        t[4] = S.p_value; t[5] = S.r_value; t[6] = S.ag_value; t[7] = S.cost_change;
      }
    } break;
    case SC_RESID: {
      if (S.cost_change < S.tol * S.cost) { S.status = 2; S.stop_next = 1; }
      S.cost = 0.5 * r[0];
    } break;
  }
}
