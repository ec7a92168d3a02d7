// Shared definitions of the helicon_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define HB2_HD __host__ __device__ __forceinline__

// ---------------------------------------------------------------------------
// IEEE helpers without FMA contraction.  The reference executes numpy scalar
// arithmetic (one rounding per operation); nvcc would fuse a*b+c.
// ---------------------------------------------------------------------------
HB2_HD float fmul_(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
HB2_HD float fadd_(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
HB2_HD float fdiv_(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b;
  return r;
#endif
}
HB2_HD double dmul_(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}
HB2_HD double dadd_(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}
HB2_HD float fsqrt64_(float t) {  // float(math.sqrt(t)): sqrt in double, then weak-scalar cast to float32
  return (float)sqrt((double)t);
}
HB2_HD float fsign_(float a) { return a > 0.f ? 1.f : (a < 0.f ? -1.f : 0.f); }

// ---------------------------------------------------------------------------
// LSMR scalar state: scipy/sparse/linalg/_isolve/lsmr.py:197-480 as executed
// for float32 A and b.  One per candidate.
// ---------------------------------------------------------------------------
struct LsmrState {
  float alpha, beta, normb;
  float zetabar, alphabar, rho, rhobar, cbar, sbar;
  float betadd, betad, rhodold, tautildeold, thetatilde, zeta, d;
  float normA2, maxrbar, minrbar;
  float rhotemp, condA;
  float normar;
  float cf_hbar, cf_x, cf_h;  // coefficients of the h/hbar/x update
  float inv_alpha, inv_beta;
  float test1, test2;
  int minrbar_inf;  // minrbar still at its 1e100 initial value
  int itn, istop, active;
  int skip_adj;     // beta == 0: lsmr.py:336 skips the v update
  double normA, normr, normx;
};

// scipy lsqr.py:62-94 (_sym_ortho) on float32 operands
HB2_HD void sym_ortho_(float a, float b, float& c, float& s, float& r) {
  if (b == 0.f) {
    c = fsign_(a); s = 0.f; r = fabsf(a);
  } else if (a == 0.f) {
    c = 0.f; s = fsign_(b); r = fabsf(b);
  } else if (fabsf(b) > fabsf(a)) {
    float tau = fdiv_(a, b);
    s = fdiv_(fsign_(b), fsqrt64_(fadd_(1.f, fmul_(tau, tau))));
    c = fmul_(s, tau);
    r = fdiv_(b, s);
  } else {
    float tau = fdiv_(b, a);
    c = fdiv_(fsign_(a), fsqrt64_(fadd_(1.f, fmul_(tau, tau))));
    s = fmul_(c, tau);
    r = fdiv_(a, c);
  }
}

// lsmr.py:239-300: state after u = b/beta, v = A^T u / alpha
HB2_HD void lsmr_init_(LsmrState& S, float alpha, float beta) {
  S.alpha = alpha; S.beta = beta; S.normb = beta;
  S.zetabar = fmul_(alpha, beta);
  S.alphabar = alpha;
  S.rho = 1.f; S.rhobar = 1.f; S.cbar = 1.f; S.sbar = 0.f;
  S.betadd = beta; S.betad = 0.f; S.rhodold = 1.f; S.tautildeold = 0.f; S.thetatilde = 0.f; S.zeta = 0.f; S.d = 0.f;
  S.normA2 = fmul_(alpha, alpha);
  S.maxrbar = 0.f; S.minrbar = 0.f; S.minrbar_inf = 1;
  S.normA = sqrt((double)S.normA2);
  S.condA = 1.f; S.normx = 0.0;
  S.normr = (double)beta;
  S.normar = fmul_(alpha, beta);
  S.itn = 0; S.istop = 0; S.skip_adj = 0;
  S.cf_hbar = S.cf_x = S.cf_h = 0.f;
  S.test1 = S.test2 = 0.f;
  S.rhotemp = 0.f;
  S.inv_alpha = alpha > 0.f ? fdiv_(1.f, alpha) : 1.f;
  S.inv_beta = beta > 0.f ? fdiv_(1.f, beta) : 0.f;
  // lsmr.py:303-311: normar == 0 or normb == 0 -> return x = 0 immediately
  S.active = (S.normar != 0.f && beta != 0.f) ? 1 : 0;
}

// lsmr.py:341-414: everything between the bidiagonalisation step and norm(x).
// alpha, beta are the NEW alpha_{k+1}, beta_{k+1}.
HB2_HD void lsmr_rotate_(LsmrState& S, float alpha, float beta) {
  S.itn += 1;
  S.alpha = alpha; S.beta = beta;
  float chat = fsign_(S.alphabar), alphahat = fabsf(S.alphabar);  // _sym_ortho(alphabar, damp=0)
  float rhoold = S.rho;
  float c, s;
  sym_ortho_(alphahat, beta, c, s, S.rho);
  float thetanew = fmul_(s, alpha);
  S.alphabar = fmul_(c, alpha);
  float rhobarold = S.rhobar, zetaold = S.zeta;
  float thetabar = fmul_(S.sbar, S.rho);
  S.rhotemp = fmul_(S.cbar, S.rho);
  float cb, sb, rb;
  sym_ortho_(fmul_(S.cbar, S.rho), thetanew, cb, sb, rb);
  S.cbar = cb; S.sbar = sb; S.rhobar = rb;
  S.zeta = fmul_(S.cbar, S.zetabar);
  S.zetabar = fmul_(-S.sbar, S.zetabar);
  S.cf_hbar = -fdiv_(fmul_(thetabar, S.rho), fmul_(rhoold, rhobarold));
  S.cf_x = fdiv_(S.zeta, fmul_(S.rho, S.rhobar));
  S.cf_h = -fdiv_(thetanew, S.rho);
  // estimate of ||r||
  float betaacute = fmul_(chat, S.betadd);
  float betacheck = 0.f;  // -shat*betadd with shat = 0
  float betahat = fmul_(c, betaacute);
  S.betadd = fmul_(-s, betaacute);
  float thetatildeold = S.thetatilde;
  float ct, st, rt;
  sym_ortho_(S.rhodold, thetabar, ct, st, rt);
  S.thetatilde = fmul_(st, S.rhobar);
  S.rhodold = fmul_(ct, S.rhobar);
  S.betad = fadd_(fmul_(-st, S.betad), fmul_(ct, betahat));
  S.tautildeold = fdiv_(fadd_(zetaold, -fmul_(thetatildeold, S.tautildeold)), rt);
  float taud = fdiv_(fadd_(S.zeta, -fmul_(S.thetatilde, S.tautildeold)), S.rhodold);
  S.d = fadd_(S.d, fmul_(betacheck, betacheck));
  float dd = fadd_(S.betad, -taud);
  S.normr = sqrt((double)fadd_(fadd_(S.d, fmul_(dd, dd)), fmul_(S.betadd, S.betadd)));
  // estimate of ||A||
  S.normA2 = fadd_(S.normA2, fmul_(beta, beta));
  S.normA = sqrt((double)S.normA2);
  S.normA2 = fadd_(S.normA2, fmul_(alpha, alpha));
  // estimate of cond(A)
  S.maxrbar = fmaxf(S.maxrbar, rhobarold);
  if (S.itn > 1) {
    S.minrbar = S.minrbar_inf ? rhobarold : fminf(S.minrbar, rhobarold);
    S.minrbar_inf = 0;
  }
  float mx = fmaxf(S.maxrbar, S.rhotemp);
  float mn = S.minrbar_inf ? S.rhotemp : fminf(S.minrbar, S.rhotemp);
  S.condA = fdiv_(mx, mn);
  S.normar = fabsf(S.zetabar);
  S.inv_alpha = alpha > 0.f ? fdiv_(1.f, alpha) : 1.f;
}

// lsmr.py:416-457: stopping tests once norm(x) of the updated x is known.
HB2_HD int lsmr_test_(LsmrState& S, double normx, double atol, double btol, double conlim, int maxiter) {
  S.normx = normx;
  float test1 = fdiv_((float)S.normr, S.normb);
  double prod = dmul_(S.normA, S.normr);
  float test2 = prod != 0.0 ? fdiv_(S.normar, (float)prod) : INFINITY;
  float test3 = fdiv_(1.f, S.condA);
  double nAx = dmul_(S.normA, normx) / (double)S.normb;
  double t1 = (double)test1 / dadd_(1.0, nAx);
  double rtol = dadd_(btol, dmul_(dmul_(atol, S.normA), normx) / (double)S.normb);
  double ctol = conlim > 0 ? 1.0 / conlim : 0.0;
  S.test1 = test1; S.test2 = test2;
  int istop = 0;
  if (S.itn >= maxiter) istop = 7;
  if (fadd_(1.f, test3) <= 1.f) istop = 6;
  if (fadd_(1.f, test2) <= 1.f) istop = 5;
  if (dadd_(1.0, t1) <= 1.0) istop = 4;
  if (test3 <= (float)ctol) istop = 3;
  if (test2 <= (float)atol) istop = 2;
  if ((double)test1 <= rtol) istop = 1;
  S.istop = istop;
  return istop;
}
