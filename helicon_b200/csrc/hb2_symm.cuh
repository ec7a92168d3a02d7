// Post-solve display products of pipeline.process_one_task:
// helicon.apply_helical_symmetry (lib/transforms.py:58-165) and the projections /
// z-section sums of pipeline.py:435-447, with the reference's float64 operation order
// (no FMA contraction) and float32 accumulation, so that results are bit-identical.
#pragma once
#include "hb2_common.cuh"

struct SymmP {
  int nz0, ny0, nx0;   // asymmetric unit
  int nz, ny, nx;      // working grid (max of unit and output per axis)
  int oz, oy, ox;      // crop offset of the output inside the working grid
  int nz1, ny1, nx1;   // output
  double apix, new_apix;
  int csym;
  int zs0, zs1;        // slab of output slices summed into the z-section
};

// one thread per output voxel; entries of a slice arrive in the reference's h order
__global__ void k_symm_volume(SymmP P, const float* __restrict__ data, const long long* __restrict__ k_begin,
                              const int* __restrict__ ent_h, const int* __restrict__ ent_floor,
                              const int* __restrict__ ent_ceil, const double* __restrict__ ent_wk,
                              const double* __restrict__ mats, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)P.nz1 * P.ny1 * P.nx1;
  if (t >= total) return;
  const int io = (int)(t % P.nx1), jo = (int)((t / P.nx1) % P.ny1), ko = (int)(t / ((long long)P.nx1 * P.ny1));
  const int j = jo + P.oy, i = io + P.ox;
  const double a = (double)(j - P.ny / 2);
  const double b = __dadd_rn((double)i, -__ddiv_rn((double)P.nx, 2.0));  // i - nx / 2 (true division)
  const double cj = (double)(P.ny0 / 2), ci0 = (double)(P.nx0 / 2);
  float acc = 0.f;
  int cnt = 0;
  const size_t plane = (size_t)P.ny0 * P.nx0;
  for (long long e = k_begin[ko]; e < k_begin[ko + 1]; ++e) {
    const double wk = ent_wk[e], wk1 = __dadd_rn(1.0, -wk);
    const float* __restrict__ df = data + (size_t)ent_floor[e] * plane;
    const float* __restrict__ dc = data + (size_t)ent_ceil[e] * plane;
    const double* __restrict__ mh = mats + (size_t)ent_h[e] * P.csym * 4;
    for (int c = 0; c < P.csym; ++c) {
      const double m00 = mh[4 * c], m01 = mh[4 * c + 1], m10 = mh[4 * c + 2], m11 = mh[4 * c + 3];
      const double j2 = __dadd_rn(__ddiv_rn(__dmul_rn(__dadd_rn(__dmul_rn(m00, a), __dmul_rn(m01, b)), P.new_apix), P.apix), cj);
      const double i2 = __dadd_rn(__ddiv_rn(__dmul_rn(__dadd_rn(__dmul_rn(m10, a), __dmul_rn(m11, b)), P.new_apix), P.apix), ci0);
      const double jf = floor(j2), if_ = floor(i2);
      if (jf < 0.0 || jf >= (double)(P.ny0 - 1)) continue;
      if (if_ < 0.0 || if_ >= (double)(P.nx0 - 1)) continue;
      const int jfl = (int)jf, jce = (int)ceil(j2), ifl = (int)if_, ice = (int)ceil(i2);
      const double wj = __dadd_rn(j2, -jf), wi = __dadd_rn(i2, -if_);
      const double wj1 = __dadd_rn(1.0, -wj), wi1 = __dadd_rn(1.0, -wi);
      const double d000 = df[(size_t)jfl * P.nx0 + ifl], d001 = df[(size_t)jfl * P.nx0 + ice];
      const double d010 = df[(size_t)jce * P.nx0 + ifl], d011 = df[(size_t)jce * P.nx0 + ice];
      const double d100 = dc[(size_t)jfl * P.nx0 + ifl], d101 = dc[(size_t)jfl * P.nx0 + ice];
      const double d110 = dc[(size_t)jce * P.nx0 + ifl], d111 = dc[(size_t)jce * P.nx0 + ice];
      double v = __dmul_rn(__dmul_rn(__dmul_rn(wk1, wj1), wi1), d000);
      v = __dadd_rn(v, __dmul_rn(__dmul_rn(__dmul_rn(wk1, wj1), wi), d001));
      v = __dadd_rn(v, __dmul_rn(__dmul_rn(__dmul_rn(wk1, wj), wi1), d010));
      v = __dadd_rn(v, __dmul_rn(__dmul_rn(__dmul_rn(wk1, wj), wi), d011));
      v = __dadd_rn(v, __dmul_rn(__dmul_rn(__dmul_rn(wk, wj1), wi1), d100));
      v = __dadd_rn(v, __dmul_rn(__dmul_rn(__dmul_rn(wk, wj1), wi), d101));
      v = __dadd_rn(v, __dmul_rn(__dmul_rn(__dmul_rn(wk, wj), wi1), d110));
      v = __dadd_rn(v, __dmul_rn(__dmul_rn(__dmul_rn(wk, wj), wi), d111));
      acc = (float)__dadd_rn((double)acc, v);  // float32 store of the float64 sum (data_work[k,j,i] += ...)
      ++cnt;
    }
  }
  out[t] = cnt > 0 ? __fdiv_rn(acc, (float)cnt) : acc;
}

// numpy's float32 pairwise summation of a contiguous run (loops_utils.h.src): 0 + pairwise_sum(a, n)
__device__ float np_pairwise_f32(const float* __restrict__ a, int n) {
  if (n < 8) {
    float r = 0.f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
    return r;
  }
  if (n <= 128) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __fadd_rn(np_pairwise_f32(a, n2), np_pairwise_f32(a + n2, n - n2));
}

// xsum[k][j] = np.sum(vol, axis=2) (pairwise); ysum[k][i] = np.sum(vol, axis=1) and zsum[j][i] = np.sum(vol[zs0:zs1],
// axis=0) (sequential along the reduced axis, as numpy iterates them)
__global__ void k_symm_project(SymmP P, const float* __restrict__ vol, float* __restrict__ xsum, float* __restrict__ ysum,
                               float* __restrict__ zsum) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nkj = (long long)P.nz1 * P.ny1, nki = (long long)P.nz1 * P.nx1, nji = (long long)P.ny1 * P.nx1;
  if (t < nkj) {
    xsum[t] = np_pairwise_f32(vol + t * P.nx1, P.nx1);
  } else if (t < nkj + nki) {
    const long long q = t - nkj;
    const int k = (int)(q / P.nx1), i = (int)(q % P.nx1);
    const float* v = vol + (size_t)k * P.ny1 * P.nx1 + i;
    float r = v[0];
    for (int j = 1; j < P.ny1; ++j) r = __fadd_rn(r, v[(size_t)j * P.nx1]);
    ysum[q] = r;
  } else if (t < nkj + nki + nji) {
    const long long q = t - nkj - nki;
    float r = 0.f;
    if (P.zs1 > P.zs0) {
      r = vol[(size_t)P.zs0 * nji + q];
      for (int k = P.zs0 + 1; k < P.zs1; ++k) r = __fadd_rn(r, vol[(size_t)k * nji + q]);
    }
    zsum[q] = r;
  }
}
