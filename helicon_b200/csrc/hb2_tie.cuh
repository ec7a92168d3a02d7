// Tie views: symmetry copies whose column -> slice rounding is a half-integer tie.  Every sample of such a
// view picks its slice from the reference's noisy z table (host: planner.reference_z_table), so a row (column, ray)
// gathers from TWO neighbouring slices -- outside the one-slice-per-row structure of the fast projector kernels.  These
// are the (rare: 0-2 views of ~73, only for "round" rise values) exact kernels for them: the fast kernels skip tie views
// (BD::view_tie), k_fwd_tie produces their rows, k_adj_tie their contribution to A^T u (BD::vtie / vtie64), which the
// adjoint kernels add before the symmetry rows.
#pragma once
#include "hb2_trf.cuh"

#define HB2_TIE_MAXZMC 16

// activity of candidate c for (mode | gate): f32 kernels use the LSMR modes, f64 kernels the TRF gates
template <bool TRF>
__device__ __forceinline__ bool tie_active(const BD& B, const TD& T, int c, int mode, bool adjoint) {
  if (TRF) {
    const TrfState& S = T.st[c];
    if (!trf_gate(S, mode)) return false;
    return !(adjoint && mode == 1 && S.in.skip_adj);
  }
  const LsmrState& S = B.st[c];
  if (!adjoint) return mode == MODE_LSMR ? (S.active != 0) : (B.only_cand < 0 || B.only_cand == c);
  return (mode == MODE_LSMR) ? (S.active != 0 && !S.skip_adj)
                             : (mode == MODE_INIT ? (S.beta > 0.f) : (B.only_cand < 0 || B.only_cand == c));
}

// bit t of upmask[(tv*G + g)*D2 + i] = up[(tv*TS + g*ZMC + t)*D2 + i]: one 16-bit load per sample instead of ZMC byte loads
__global__ void k_tie_pack(int n_tie, int TS, int ZMC, int D2, const unsigned char* __restrict__ up, uint16_t* __restrict__ mask) {
  const int G = TS / ZMC;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_tie * G * D2) return;
  const int i = (int)(t % D2), g = (int)((t / D2) % G), tv = (int)(t / ((long long)D2 * G));
  unsigned m = 0;
  for (int q = 0; q < ZMC; ++q) m |= (unsigned)(up[((size_t)tv * TS + (size_t)g * ZMC + q) * D2 + i] & 1u) << q;
  mask[t] = (uint16_t)m;
}

// per tie view slot: which column slots are used and whether they sit on consecutive slices from zb = -1 or 0
__global__ void k_tie_info(int n_tie_views, const int* __restrict__ tie_views, int ZMC, int tie_TS, const int* __restrict__ view_tie,
                           const int* __restrict__ view_tie_slot0, const int* __restrict__ view_colbegin,
                           const int* __restrict__ colk, const signed char* __restrict__ tie_zlo, int* __restrict__ info) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_tie_views) return;
  const int view = tie_views[e];
  const int tv = view_tie[view], s0 = view_tie_slot0[view];
  const int* ck = colk + view_colbegin[view];
  const signed char* zl = tie_zlo + (size_t)tv * tie_TS + s0;
  unsigned used = 0;
  for (int t = 0; t < ZMC && t < 16; ++t) used |= (ck[t] >= 0 ? 1u : 0u) << t;
  const int zb = zl[0];
  bool ok = (zb == -1 || zb == 0) && s0 % ZMC == 0 && tie_TS % ZMC == 0 && ZMC <= 16;
  for (int t = 0; t < ZMC && ok; ++t)
    if (ck[t] >= 0 && zl[t] != zb + t) ok = false;
  info[view] = (int)(used | (ok ? (unsigned)(zb + 2) << 16 : 0u));
}

// Grid (tie view slots, ray groups): a CTA takes the rays j = blockIdx.y, blockIdx.y + gridDim.y, ... of one tie view slot
// (ZMC column slots of one tie view), one warp per ray, lanes over the depth samples; gridDim.y = the view's partial-sum
// slots (BD::fwd_ppv).  (One CTA per view slot cost 2.2 ms per pass at 512 x 512 -- a handful of CTAs walking 512 x 512
// samples each -- and dominated small batches, profiles/r1_summary.md section 7.)
//   f32 (TRF = false): src = v or xs, rows = u, LSMR / PLAIN / SCORE epilogue of k_fwd_data incl. the view's partials;
//   f64 (TRF = true) : rows <- A w plain.
template <typename IdxT, typename T, bool TRF>
__global__ void __launch_bounds__(HB2_BLOCK, 6) k_fwd_tie(BD B, TD Tt, const T* __restrict__ src, T* __restrict__ rows, int mode) {
  const int view = B.tie_views[blockIdx.x];
  const int c = B.view_cand[view];
  __shared__ float red[HB2_BLOCK / 32];
  __shared__ int s_colk[HB2_TIE_MAXZMC], s_zlo[HB2_TIE_MAXZMC];
  const int ppv = B.fwd_ppv, sub = blockIdx.y;
  const bool act = tie_active<TRF>(B, Tt, c, mode, false);
  if (!act) {
    if (!TRF && threadIdx.x == 0) {
      if (mode == MODE_LSMR) B.part_u[view * ppv + sub] = 0.f;
      if (mode == MODE_SCORE) { B.part_s[3 * (view * ppv + sub)] = 0.f; B.part_s[3 * (view * ppv + sub) + 1] = 0.f; B.part_s[3 * (view * ppv + sub) + 2] = 0.f; }
    }
    return;
  }
  const int D2 = B.D2, L3 = B.L3, L3P = B.L3P, ZMC = B.ZMC, ZMP = B.ZMP;
  const int tv = B.view_tie[view], s0 = B.view_tie_slot0[view], a = B.view_angle[view];
  if (threadIdx.x < ZMC) {
    s_colk[threadIdx.x] = B.colk[B.view_colbegin[view] + threadIdx.x];
    s_zlo[threadIdx.x] = B.tie_zlo[(size_t)tv * B.tie_TS + s0 + threadIdx.x];
  }
  __syncthreads();
  const IdxT* __restrict__ fm = (const IdxT*)B.fmap + (size_t)a * D2 * D2;
  const T* __restrict__ vsrc = src + (size_t)c * B.npad;
  const unsigned char* __restrict__ up = B.tie_up + ((size_t)tv * B.tie_TS + s0) * D2;
  const unsigned char* __restrict__ rv = B.tie_rowvalid + ((size_t)tv * B.tie_TS + s0) * D2;
  // fast path: used slots on consecutive slices starting at zb = -1 or 0 (CTA-uniform), packed masks available
  int fast_zb = 99;
  if (B.tie_upmask && s0 % ZMC == 0 && B.tie_TS % ZMC == 0) {
    const int zb = s_zlo[0];
    bool ok = zb == -1 || zb == 0;
    for (int t = 0; t < ZMC && ok; ++t)
      if (s_colk[t] >= 0 && s_zlo[t] != zb + t) ok = false;
    if (ok) fast_zb = zb;
  }
  T* urow = rows + B.view_uoff[view];
  const float* brow = B.b + B.view_uoff[view];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float alpha = 0.f, inv_beta = 0.f;
  if (!TRF) { alpha = B.st[c].alpha; inv_beta = B.st[c].inv_beta; }
  float ss = 0.f, s_pb = 0.f, s_bb = 0.f;
  const uint8_t* __restrict__ pm = cand_mask(B, c);
  for (int j = sub * (HB2_BLOCK / 32) + warp; j < D2; j += ppv * (HB2_BLOCK / 32)) {
    T acc[HB2_TIE_MAXZMC];
#pragma unroll
    for (int t = 0; t < HB2_TIE_MAXZMC; ++t) acc[t] = (T)0;
    const IdxT* __restrict__ fj = fm + (size_t)j * D2;
    if (fast_zb != 99) {
      // consecutive column slots sit on consecutive slices (zlo[t] = zb + t, zb = -1 or 0): the voxel record comes in
      // with 128-bit loads and slot t takes slice zb + t or zb + t + 1 by its bit of the packed mask
      const uint16_t* __restrict__ um = B.tie_upmask + ((size_t)tv * (B.tie_TS / ZMC) + s0 / ZMC) * D2;
      for (int i = lane; i < D2; i += 32) {
        const IdxT id = fj[i];
        if (id == Sent<IdxT>::v) continue;
        const T* __restrict__ vb = vsrc + (size_t)id * L3P;
        const unsigned m = um[i];
        T rec[HB2_TIE_MAXZMC];
#pragma unroll
        for (int z4 = 0; z4 < HB2_TIE_MAXZMC; z4 += 4) {
          if (z4 < L3P) {
            if constexpr (sizeof(T) == 4) {
              const float4 q4 = *reinterpret_cast<const float4*>(vb + z4);
              rec[z4] = q4.x; rec[z4 + 1] = q4.y; rec[z4 + 2] = q4.z; rec[z4 + 3] = q4.w;
            } else {
              const double2 a2 = *reinterpret_cast<const double2*>(vb + z4), b2 = *reinterpret_cast<const double2*>(vb + z4 + 2);
              rec[z4] = a2.x; rec[z4 + 1] = a2.y; rec[z4 + 2] = b2.x; rec[z4 + 3] = b2.y;
            }
          } else {
            rec[z4] = rec[z4 + 1] = rec[z4 + 2] = rec[z4 + 3] = (T)0;
          }
        }
#pragma unroll
        for (int t = 0; t < HB2_TIE_MAXZMC; ++t) {
          if (t < ZMC && s_colk[t] >= 0) {
            // slice index zb + t + bit, zb in {-1, 0}: static register picks
            const bool hi = (m >> t) & 1u;
            T lo_v, hi_v;
            if (fast_zb == 0) { lo_v = t < L3 ? rec[t] : (T)0; hi_v = (t + 1 < L3 && t + 1 < HB2_TIE_MAXZMC) ? rec[(t + 1) % HB2_TIE_MAXZMC] : (T)0; }
            else { lo_v = (t >= 1 && t - 1 < L3) ? rec[(t + HB2_TIE_MAXZMC - 1) % HB2_TIE_MAXZMC] : (T)0; hi_v = t < L3 ? rec[t] : (T)0; }
            acc[t] += hi ? hi_v : lo_v;
          }
        }
      }
    } else
    for (int i = lane; i < D2; i += 32) {
      const IdxT id = fj[i];
      if (id == Sent<IdxT>::v) continue;
      const T* __restrict__ vb = vsrc + (size_t)id * L3P;
#pragma unroll
      for (int t = 0; t < HB2_TIE_MAXZMC; ++t) {
        if (t < ZMC && s_colk[t] >= 0) {
          const int z = s_zlo[t] + (int)up[(size_t)t * D2 + i];
          if ((unsigned)z < (unsigned)L3) acc[t] += vb[z];
        }
      }
    }
#pragma unroll
    for (int t = 0; t < HB2_TIE_MAXZMC; ++t) {
      if (t < ZMC && s_colk[t] >= 0) {  // warp-uniform: unused column slots (11 of 12 in a view's second slot group) cost nothing
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
      }
    }
    // lane t finishes column slot t
#pragma unroll
    for (int t = 0; t < HB2_TIE_MAXZMC; ++t) {
      if (lane == t && t < ZMC && s_colk[t] >= 0 && rv[(size_t)t * D2 + j] && !(pm && !pm[(size_t)s_colk[t] * D2 + j])) {
        const size_t ri = (size_t)j * ZMP + t;
        if (TRF) {
          urow[ri] = acc[t];
        } else if (mode == MODE_LSMR) {
          const float un = fadd_(fmul_(fmul_((float)urow[ri], inv_beta), -alpha), (float)acc[t]);
          urow[ri] = (T)un;
          ss += un * un;
        } else if (mode == MODE_PLAIN) {
          urow[ri] = acc[t];
        } else {
          const float pred = B.clip_pred ? fmaxf((float)acc[t], 0.f) : (float)acc[t];
          const float bv = brow[ri];
          ss += pred * pred; s_pb += pred * bv; s_bb += bv * bv;
        }
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

// Adjoint of the tie views: one thread per in-plane voxel, all slices; grid (ceil(ndisk/256), nc).
// vt[c][p*L3P + z] = sum over the candidate's tie view slots, samples (j, i) landing in p and column slots t with
// zlo[t] + up[t][i] == z of rows[view][j][t] * ib.  Candidates without tie views write nothing (the consumers test
// cand_tie_count).
template <typename T, bool TRF>
__global__ void __launch_bounds__(HB2_BLOCK) k_adj_tie(BD B, TD Tt, const T* __restrict__ rows, T* __restrict__ vt, int mode) {
  const int c = blockIdx.y;
  const int nt = B.cand_tie_count[c];
  if (nt == 0) return;
  if (!tie_active<TRF>(B, Tt, c, mode, true)) return;
  const int p = blockIdx.x * HB2_BLOCK + threadIdx.x;
  if (p >= B.ndisk) return;
  const int D2 = B.D2, L3 = B.L3, L3P = B.L3P, ZMC = B.ZMC, ZMP = B.ZMP, K = B.K;
  T ib = (T)1;
  if (!TRF && mode != MODE_PLAIN) ib = (T)B.st[c].inv_beta;
  T acc[HB2_TIE_MAXZMC];
#pragma unroll
  for (int z = 0; z < HB2_TIE_MAXZMC; ++z) acc[z] = (T)0;
  const int slot = B.aslot[p];
  for (int e = 0; e < nt; ++e) {
    const int view = B.tie_views[B.cand_tie_begin[c] + e];
    const int a = B.view_angle[view], tv = B.view_tie[view], s0 = B.view_tie_slot0[view];
    const int* __restrict__ ck = B.colk + B.view_colbegin[view];
    const signed char* __restrict__ zl = B.tie_zlo + (size_t)tv * B.tie_TS + s0;
    const unsigned char* __restrict__ up = B.tie_up + ((size_t)tv * B.tie_TS + s0) * D2;
    const T* __restrict__ ub = rows + B.view_uoff[view];
    // fast path (see k_fwd_tie): used slots on consecutive slices from zb = -1 or 0, packed masks; the per-slot facts
    // come precomputed (BD::tie_info), not re-derived by every voxel
    const int inf = B.tie_info ? B.tie_info[view] : 0;
    const unsigned used = (unsigned)inf & 0xFFFFu;
    const int fast_zb = (B.tie_upmask && (inf >> 16)) ? (inf >> 16) - 2 : 99;
    const uint16_t* __restrict__ um = B.tie_upmask ? B.tie_upmask + ((size_t)tv * (B.tie_TS / ZMC) + s0 / ZMC) * D2 : nullptr;
    for (int k = 0; k < K; ++k) {
      const size_t mi = ((size_t)a * K + k) * B.apitch + slot;
      const unsigned j = B.amap[mi];
      if (j == 0xFFFFu) continue;
      const unsigned i = B.amap_i[mi];
      const T* __restrict__ uj = ub + (size_t)j * ZMP;
      if (fast_zb != 99) {
        const unsigned m = um[i];
        const unsigned lo_m = used & ~m, hi_m = used & m;  // slots whose sample falls into the lower / the upper slice
        if (fast_zb == 0) {  // slot t -> slice t (lower) or t + 1 (upper)
#pragma unroll
          for (int t = 0; t < HB2_TIE_MAXZMC; ++t) {
            if (t < ZMC) {
              const T val = TRF ? uj[t] : (T)fmaf((float)uj[t], (float)ib, 0.f);
              if (((lo_m >> t) & 1u) && t < L3) acc[t] += val;
              if (t + 1 < HB2_TIE_MAXZMC) { if (((hi_m >> t) & 1u) && t + 1 < L3) acc[(t + 1) % HB2_TIE_MAXZMC] += val; }
            }
          }
        } else {             // slot t -> slice t - 1 (lower) or t (upper)
#pragma unroll
          for (int t = 0; t < HB2_TIE_MAXZMC; ++t) {
            if (t < ZMC) {
              const T val = TRF ? uj[t] : (T)fmaf((float)uj[t], (float)ib, 0.f);
              if (((hi_m >> t) & 1u) && t < L3) acc[t] += val;
              if (t >= 1) { if (((lo_m >> t) & 1u) && t - 1 < L3) acc[(t + HB2_TIE_MAXZMC - 1) % HB2_TIE_MAXZMC] += val; }
            }
          }
        }
        continue;
      }
      for (int t = 0; t < ZMC; ++t) {
        if (ck[t] < 0) continue;
        const int z = (int)zl[t] + (int)up[(size_t)t * D2 + i];
        if ((unsigned)z < (unsigned)L3) {
          const T val = TRF ? uj[t] : (T)fmaf((float)uj[t], (float)ib, 0.f);
          // select the accumulator without dynamic register indexing
#pragma unroll
          for (int zz = 0; zz < HB2_TIE_MAXZMC; ++zz)
            if (zz == z) acc[zz] += val;
        }
      }
    }
  }
  T* dst = vt + (size_t)c * B.npad + (size_t)p * L3P;
#pragma unroll
  for (int z = 0; z < HB2_TIE_MAXZMC; ++z)
    if (z < L3P) dst[z] = z < L3 ? acc[z] : (T)0;
}
