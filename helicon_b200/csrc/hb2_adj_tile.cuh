// Tile adjoint (TMA-staged), float32 for LSMR and float64 for the bounded branch.  Included after hb2_trf.cuh /
// hb2_tie.cuh because the float64 instantiation consults the TRF gates.
#pragma once
#include "hb2_tie.cuh"

// ---------------------------------------------------------------------------
// Adjoint, tile path (MC == 1, K <= 2, L3P <= 16): one CTA = one voxel tile
// (8 x 32 in-plane patch, <= 256 voxels, one thread each, all slices in
// registers).  profiles/r1_summary.md: with global gathers every view costs a
// dependent L2 round trip (map entry -> row) and the kernel is bound by
// loads-in-flight / latency.  Here both operands of a stage of HB2_ADJT_SV
// views are brought to shared memory by the TMA engine (cp.async.bulk, mbarrier
// completion, two stages in flight): the tile's 512-byte runs of the adjoint
// map, and -- because a compact tile is crossed by a short contiguous range of
// rays in every view -- one contiguous window [jlo, jlo+nr) x L3P of the view's
// rows.  The inner loop then touches shared memory only.
// Addition order: views, then k, then symmetry rows (= k_adj / k_adj_pq).
// ---------------------------------------------------------------------------
#define HB2_ADJT_SV 4      // views per stage
#define HB2_ADJT_NS 4      // stages in flight
#define HB2_ADJT_THREADS (HB2_BLOCK + 32)  // 8 consumer warps (one thread per voxel) + 1 producer warp
#define HB2_ADJT_MAXV 256
// T = float : the LSMR adjoint (modes of k_adj: LSMR / INIT / PLAIN) on B.u -> B.v / B.xs
// T = double: the bounded branch's plain adjoint dst <- A^T rows (gate semantics of k_adj64)
// CHUNKED (more than 16 slices): grid.z = chunks of 4*NQT = 16 slices; a CTA handles one chunk of its tile's voxels.  The
// rows of a view are then no longer one contiguous window: every ray row contributes one 64-byte piece (its slices of
// the chunk), copied by its own bulk copy -- the producer warp spreads the rows of a stage over its 32 lanes.
template <int NQT, int KT, typename T, bool TRF, bool CHUNKED = false>
__global__ void __launch_bounds__(HB2_ADJT_THREADS) k_adj_tile(BD B, TD Tt, const T* __restrict__ rows, T* __restrict__ dst_all, int mode) {
  extern __shared__ __align__(128) unsigned char dsm[];
  const int c = blockIdx.y, tile = blockIdx.x;
  __shared__ float red[HB2_BLOCK / 32];
  __shared__ unsigned long long full_bar[HB2_ADJT_NS], empty_bar[HB2_ADJT_NS];
  __shared__ int s_ang[HB2_ADJT_MAXV];
  __shared__ uint16_t s_jlo[HB2_ADJT_MAXV], s_nr[HB2_ADJT_MAXV];
  __shared__ float s_w[HB2_ADJT_MAXV];  // multiplicity of the view (Halton duplicates are skipped, their first copy counts twice)
  const LsmrState& S = B.st[c];
  const bool act = tie_active<TRF>(B, Tt, c, mode, true);
  const int pi = c * B.part_v_per_cand + (CHUNKED ? (int)blockIdx.z * B.ntile : 0) + blockIdx.x;
  if (!act) {
    if (!TRF && threadIdx.x == 0 && mode != MODE_PLAIN) B.part_v[pi] = 0.f;
    return;
  }
  constexpr int L3P = 4 * NQT;  // slices handled by this CTA (= all of them unless CHUNKED)
  const int zc0 = CHUNKED ? (int)blockIdx.z * L3P : 0;                 // first slice of the chunk
  const int cw = CHUNKED ? min(L3P, B.L3P - zc0) : L3P;                // slices of the chunk (multiple of 4)
  const int vstride = CHUNKED ? B.L3P : L3P;                           // voxel / ray-row stride in global memory
  const int ndisk_t = B.tile_begin[tile + 1] - B.tile_begin[tile];
  const int p = B.tile_begin[tile] + threadIdx.x;
  const bool producer = threadIdx.x >= HB2_BLOCK;
  const bool live = threadIdx.x < ndisk_t;
  const unsigned rpv = (unsigned)B.rows_per_view;
  const int vb = B.cand_view_begin[c], nv = B.cand_view_count[c];
  const int rmax = B.rmax;
  // dynamic shared memory: [NS][SV][KT][256] map entries, then [NS][SV][rmax*L3P] row windows
  uint16_t* s_map = reinterpret_cast<uint16_t*>(dsm);
  T* s_u = reinterpret_cast<T*>(dsm + (size_t)HB2_ADJT_NS * HB2_ADJT_SV * KT * HB2_BLOCK * sizeof(uint16_t));
  const int ustride = rmax * L3P;
  for (int e = threadIdx.x; e < nv; e += HB2_ADJT_THREADS) {
    const int a = B.view_angle[vb + e];
    // tie views are handled by k_adj_tie (BD::vtie); duplicates (solver modes only) by the weight of their first copy
    const bool dedupe_adj = !TRF && mode != MODE_PLAIN;  // the float64 operators keep duplicate rows apart
    const bool tie = (B.view_tie && B.view_tie[vb + e] >= 0) || (dedupe_adj && B.view_dupof[vb + e] >= 0);
    s_w[e] = dedupe_adj ? (float)B.view_mult[vb + e] : 1.f;
    s_ang[e] = a;
    s_jlo[e] = tie ? (uint16_t)0xFFFFu : B.tile_jlo[(size_t)a * B.ntile + tile];  // 0xFFFF: skip the view
    s_nr[e] = tie ? (uint16_t)0xFFFFu : B.tile_nr[(size_t)a * B.ntile + tile];
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < HB2_ADJT_NS; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], HB2_BLOCK / 32); }
  }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncthreads();
  const int nstage = (nv + HB2_ADJT_SV - 1) / HB2_ADJT_SV;
  const T ib = (TRF || mode == MODE_PLAIN) ? (T)1 : (T)S.inv_beta;
  const float beta = S.beta;
  T acc[4 * NQT];
#pragma unroll
  for (int i = 0; i < 4 * NQT; ++i) acc[i] = (T)0;
  if (producer) {
    // producer warp: lane w of a stage loads view st*SV + w (K map runs of 512 bytes + the row window)
    const T* __restrict__ ucand = rows + B.cand_uoff[c];
    const uint16_t* __restrict__ amt = B.amap + (size_t)tile * HB2_BLOCK;
    const int w = threadIdx.x - HB2_BLOCK;
    for (int st = 0; st < nstage; ++st) {
      const int buf = st % HB2_ADJT_NS;
      if (st >= HB2_ADJT_NS) mbar_wait(&empty_bar[buf], (unsigned)(((st / HB2_ADJT_NS) - 1) & 1));
      const int v = st * HB2_ADJT_SV + w;
      const bool has = w < HB2_ADJT_SV && v < nv && s_nr[min(v, nv - 1)] != 0xFFFFu;
      const unsigned nr = has ? s_nr[v] : 0;
      unsigned tot = has ? KT * HB2_BLOCK * (unsigned)sizeof(uint16_t) + nr * (unsigned)cw * (unsigned)sizeof(T) : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      if (w == 0) mbar_expect_tx(&full_bar[buf], tot);
      __syncwarp();
      if (has) {
        const size_t arow = (size_t)s_ang[v] * KT * B.apitch;
#pragma unroll
        for (int k = 0; k < KT; ++k)
          bulk_g2s(s_map + ((size_t)(buf * HB2_ADJT_SV + w) * KT + k) * HB2_BLOCK, amt + arow + (size_t)k * B.apitch,
                   HB2_BLOCK * (unsigned)sizeof(uint16_t), &full_bar[buf]);
        if (!CHUNKED && nr)
          bulk_g2s(s_u + (size_t)(buf * HB2_ADJT_SV + w) * ustride, ucand + ((size_t)v * rpv + (size_t)s_jlo[v] * L3P),
                   nr * L3P * (unsigned)sizeof(T), &full_bar[buf]);
      }
      if (CHUNKED) {  // one piece per ray row: rows of the stage's views spread over the lanes
#pragma unroll
        for (int ws = 0; ws < HB2_ADJT_SV; ++ws) {
          const int vs = st * HB2_ADJT_SV + ws;
          if (vs >= nv || s_nr[vs] == 0xFFFFu) continue;
          const unsigned nrs = s_nr[vs];
          const T* src = ucand + ((size_t)vs * rpv + (size_t)s_jlo[vs] * vstride + zc0);
          T* dstw = s_u + (size_t)(buf * HB2_ADJT_SV + ws) * ustride;
          for (unsigned r = (unsigned)w; r < nrs; r += 32u)
            bulk_g2s(dstw + (size_t)r * L3P, src + (size_t)r * vstride, (unsigned)cw * (unsigned)sizeof(T), &full_bar[buf]);
        }
      }
    }
  } else {
    for (int st = 0; st < nstage; ++st) {
      const int buf = st % HB2_ADJT_NS;
      mbar_wait(&full_bar[buf], (unsigned)((st / HB2_ADJT_NS) & 1));
      if (live) {
        const int nvs = min(HB2_ADJT_SV, nv - st * HB2_ADJT_SV);
#pragma unroll
        for (int w = 0; w < HB2_ADJT_SV; ++w) {
          const int jl = w < nvs ? (int)s_jlo[st * HB2_ADJT_SV + w] : 0xFFFF;
          if (jl != 0xFFFF) {
            const uint16_t* mp = s_map + ((size_t)(buf * HB2_ADJT_SV + w) * KT) * HB2_BLOCK + threadIdx.x;
            const T* uw = s_u + (size_t)(buf * HB2_ADJT_SV + w) * ustride;
            const T ibw = ib * (T)s_w[st * HB2_ADJT_SV + w];
#pragma unroll
            for (int k = 0; k < KT; ++k) {
              const unsigned j = mp[(size_t)k * HB2_BLOCK];
              if (j != 0xFFFFu) {
                const T* r = uw + ((int)j - jl) * L3P;
                if constexpr (sizeof(T) == 4) {
#pragma unroll
                  for (int q4 = 0; q4 < NQT; ++q4) {
                    const float4 t = reinterpret_cast<const float4*>(r)[q4];
                    acc[4 * q4 + 0] = fmaf(t.x, ibw, acc[4 * q4 + 0]); acc[4 * q4 + 1] = fmaf(t.y, ibw, acc[4 * q4 + 1]);
                    acc[4 * q4 + 2] = fmaf(t.z, ibw, acc[4 * q4 + 2]); acc[4 * q4 + 3] = fmaf(t.w, ibw, acc[4 * q4 + 3]);
                  }
                } else {
#pragma unroll
                  for (int q2 = 0; q2 < 2 * NQT; ++q2) {
                    const double2 t = reinterpret_cast<const double2*>(r)[q2];
                    acc[2 * q2 + 0] += t.x; acc[2 * q2 + 1] += t.y;  // ibw == 1 on the float64 path
                  }
                }
              }
            }
          }
        }
      }
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty_bar[buf]);  // this warp is done with the stage's buffers
    }
  }
  float ss = 0.f;
  if (live && !producer) {
    const int g0 = p * vstride + zc0;
    const T* __restrict__ us = rows + B.cand_uoff[c] + B.cand_mdata[c];
    const int* __restrict__ ell = B.ell + (size_t)c * HB2_ELL_W * B.npad + g0;
    T* vdst = (TRF ? dst_all : (T*)(mode == MODE_PLAIN ? (void*)B.xs : (void*)B.v)) + (size_t)c * B.npad + g0;
    const T* vtie = TRF ? (const T*)(const void*)B.vtie64 : (const T*)(const void*)B.vtie;
    if (vtie && B.cand_tie_count[c] > 0) {  // rows of the tie views (k_adj_tie)
      const T* vt = vtie + (size_t)c * B.npad + g0;
#pragma unroll
      for (int z = 0; z < 4 * NQT; ++z)
        if (z < cw) acc[z] += vt[z];
    }
#pragma unroll
    for (int q4 = 0; q4 < NQT; ++q4) {
      if (4 * q4 >= cw) break;  // partial last chunk
      int ev[HB2_ELL_W][4];
#pragma unroll
      for (int w = 0; w < HB2_ELL_W; ++w) {
        const int4 e4 = __ldg(reinterpret_cast<const int4*>(ell + (size_t)w * B.npad) + q4);
        ev[w][0] = e4.x; ev[w][1] = e4.y; ev[w][2] = e4.z; ev[w][3] = e4.w;
      }
      T old[4] = {(T)0, (T)0, (T)0, (T)0};
      if constexpr (sizeof(T) == 4) {
        if (!TRF && mode == MODE_LSMR) {
          const float4 o = *reinterpret_cast<const float4*>(vdst + 4 * q4);
          old[0] = o.x; old[1] = o.y; old[2] = o.z; old[3] = o.w;
        }
      }
      T vals[HB2_ELL_W][4];
#pragma unroll
      for (int w = 0; w < HB2_ELL_W; ++w)
#pragma unroll
        for (int tz = 0; tz < 4; ++tz) {
          const bool ok = ev[w][tz] < 0 || ev[w][tz] < HB2_ELL_OVERFLOW;  // a real entry (either sign)
          vals[w][tz] = ok ? __ldg(us + (ev[w][tz] & 0x7fffffff)) : (T)0;
        }
      T vn[4];
#pragma unroll
      for (int tz = 0; tz < 4; ++tz) {
        T s2 = acc[4 * q4 + tz];
        if (ev[HB2_ELL_W - 1][tz] == HB2_ELL_OVERFLOW) {  // long list: walk the CSR copy
          const int* __restrict__ ptr = B.csc_ptr + (size_t)c * (B.npad + 1) + g0 + 4 * q4 + tz;
          const int* __restrict__ ent = B.csc_ent + B.cand_cscoff[c];
          for (int e = ptr[0]; e < ptr[1]; ++e) {
            const int x = ent[e];
            const T val = __ldg(us + (x & 0x7fffffff));
            if constexpr (sizeof(T) == 4) s2 = fmaf(x < 0 ? -val : val, ib, s2);
            else s2 += x < 0 ? -val : val;
          }
        } else {
#pragma unroll
          for (int w = 0; w < HB2_ELL_W; ++w)
            if (ev[w][tz] != HB2_ELL_NONE) {
              if constexpr (sizeof(T) == 4) s2 = fmaf(ev[w][tz] < 0 ? -vals[w][tz] : vals[w][tz], ib, s2);
              else s2 += ev[w][tz] < 0 ? -vals[w][tz] : vals[w][tz];
            }
        }
        if constexpr (sizeof(T) == 4) {
          vn[tz] = (!TRF && mode == MODE_LSMR) ? fadd_(fmul_(old[tz], -beta), s2) : s2;
          ss += vn[tz] * vn[tz];
        } else {
          vn[tz] = s2;
        }
      }
      if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(vdst + 4 * q4) = make_float4(vn[0], vn[1], vn[2], vn[3]);
      } else {
        reinterpret_cast<double2*>(vdst + 4 * q4)[0] = make_double2(vn[0], vn[1]);
        reinterpret_cast<double2*>(vdst + 4 * q4)[1] = make_double2(vn[2], vn[3]);
      }
    }
  }
  if (!TRF && mode != MODE_PLAIN) {  // 9 warps: the producer warp only joins the barrier
    ss = warp_sum(ss);
    if (!producer && (threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < HB2_BLOCK / 32; ++w) tot += red[w];
      B.part_v[pi] = tot;
    }
  }
}

