// Forward projector, band path: float32 for LSMR / plain / score, float64 for the bounded branch (plain, TRF gates).
// Included after hb2_tie.cuh (tie_active, TD).
#pragma once
#include "hb2_tie.cuh"

// ===========================================================================
// The gather kernel k_fwd_data re-reads v from L2 for every view (157 MB per
// candidate-pass at cfg2, ~9 TB/s at the L2 port: that port, not HBM, bounds
// it).  Here a CTA keeps one BAND of the candidate's voxels -- a contiguous run
// of the band-column-major voxel order inside one 16-row band of the disk,
// <= ~196 KB of v -- in shared memory (TMA bulk copies) and applies ALL views of
// the candidate to it: every ray of every view that crosses the band contributes
// one partial sum (its samples inside the band); k_fwd_band_reduce adds a ray's
// partials over the bands it crosses, in band order, and applies the row
// epilogue.  v leaves L2 once per pass.
// Lanes: a warp item is RPW = 32/NL adjacent rays of one view x NL 16-byte
// chunks of a voxel record (4 float32 / 2 float64 slices each); a lane owns one
// (ray, chunk), walks the ray's samples inside the band and keeps its sums in
// registers -- no cross-lane reduction.  All lanes of the warp sit on the SAME
// depth sample at any time: adjacent rays are then one voxel row apart (view
// angles below 45 degrees), rows of a column are adjacent records and a column
// is 16 records, so the 8 lanes of a quarter-warp read 8 distinct 16-byte bank
// groups (tile_order in hb2_api.cu).  Map entries arrive 8 at a time: the NL
// lanes of a ray fetch NL consecutive 128-bit chunks of its map row and pass
// them around with shuffles; "rank inside the band" is the only test per sample.
// ===========================================================================
#ifndef HB2_FWDB_THREADS
#define HB2_FWDB_THREADS 768
#endif
#define HB2_FWDB_MAXV 256
#define HB2_MAX_BANDS 96

// sample range of every ray inside every band (setup): one thread per (angle, band, ray)
template <typename IdxT>
__global__ void k_band_segs(int nA, int nband, int D2, const int* __restrict__ band_begin, const IdxT* __restrict__ fmap,
                            ushort2* __restrict__ seg) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nA * nband * D2) return;
  const int j = (int)(t % D2), b = (int)((t / D2) % nband), a = (int)(t / ((long long)D2 * nband));
  const unsigned bb = (unsigned)band_begin[b], bn = (unsigned)band_begin[b + 1] - bb;
  const IdxT* __restrict__ fj = fmap + ((size_t)a * D2 + j) * D2;
  int lo = D2, hi = 0;
  for (int i = 0; i < D2; ++i) {
    const unsigned rel = (unsigned)fj[i] - bb;
    if (fj[i] != Sent<IdxT>::v && rel < bn) { lo = min(lo, i); hi = i + 1; }
  }
  seg[t] = hi > lo ? make_ushort2((unsigned short)lo, (unsigned short)hi) : make_ushort2(0, 0);
}
// rays crossing every band (setup): one CTA per (band, angle)
__global__ void __launch_bounds__(HB2_BLOCK) k_band_rng(int nband, int D2, const ushort2* __restrict__ seg,
                                                       ushort2* __restrict__ rng) {
  const int b = blockIdx.x, a = blockIdx.y;
  __shared__ int s_lo, s_hi;
  if (threadIdx.x == 0) { s_lo = D2; s_hi = 0; }
  __syncthreads();
  const ushort2* sg = seg + ((size_t)a * nband + b) * D2;
  for (int j = threadIdx.x; j < D2; j += HB2_BLOCK)
    if (sg[j].y > sg[j].x) { atomicMin(&s_lo, j); atomicMax(&s_hi, j + 1); }
  __syncthreads();
  if (threadIdx.x == 0) rng[(size_t)a * nband + b] = s_hi > s_lo ? make_ushort2((unsigned short)s_lo, (unsigned short)s_hi) : make_ushort2(0, 0);
}

template <typename T> struct Vec16;
template <> struct Vec16<float> {
  typedef float4 type;
  static constexpr int N = 4;
  static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ float4 lds(unsigned a) { return lds128(a); }
  static __device__ __forceinline__ void add(float4& s, const float4& t) { s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
};
template <> struct Vec16<double> {
  typedef double2 type;
  static constexpr int N = 2;
  static __device__ __forceinline__ double2 zero() { return make_double2(0.0, 0.0); }
  static __device__ __forceinline__ double2 lds(unsigned a) {
    double2 r;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(r.x), "=d"(r.y) : "r"(a));
    return r;
  }
  static __device__ __forceinline__ void add(double2& s, const double2& t) { s.x += t.x; s.y += t.y; }
};

// NL: 16-byte chunks per voxel record = L3P * sizeof(T) / 16.  TRF: float64 operator of the bounded branch (src / gate).
template <int NL, typename T, bool TRF>
__global__ void __launch_bounds__(HB2_FWDB_THREADS, 1) k_fwd_band(BD B, TD Tt, const T* __restrict__ src_all, int mode) {
  typedef Vec16<T> V;
  typedef typename V::type vec_t;
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ unsigned long long bar;
  __shared__ int s_pref[HB2_FWDB_MAXV + 1];
  __shared__ unsigned short s_jlo[HB2_FWDB_MAXV], s_cnt[HB2_FWDB_MAXV];
  __shared__ int s_ang[HB2_FWDB_MAXV];
  __shared__ int s_next;
  const BandTab& Bt = TRF ? B.bt64 : B.bt32;
  const int c = blockIdx.y, b = blockIdx.x;
  if (!tie_active<TRF>(B, Tt, c, mode, false)) return;
  constexpr int L3P = NL * V::N;
  constexpr int RPW = 32 / NL;                       // rays per warp item
  constexpr unsigned REC = L3P * (unsigned)sizeof(T);
  const int D2 = B.D2, NB = Bt.nband;
  const unsigned bb = (unsigned)Bt.band_begin[b], bn = (unsigned)Bt.band_begin[b + 1] - bb;
  const int vb = B.cand_view_begin[c], nv = B.cand_view_count[c];
  const T* src = TRF ? src_all : (const T*)(const void*)(mode == MODE_LSMR ? B.v : B.xs);
  const unsigned char* __restrict__ vsrc = reinterpret_cast<const unsigned char*>(src + (size_t)c * B.npad + (size_t)bb * L3P);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); s_next = HB2_FWDB_THREADS / 32; }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  for (int e = threadIdx.x; e < nv; e += HB2_FWDB_THREADS) {
    const int view = vb + e;
    const int a = B.view_angle[view];
    ushort2 r = Bt.rng[(size_t)a * NB + b];
    // tie views: k_fwd_tie; duplicates of an earlier view: served by that view (the epilogue copies the rows)
    if ((B.view_tie && B.view_tie[view] >= 0) || B.view_dupof[view] >= 0) r = make_ushort2(0, 0);
    s_ang[e] = a; s_jlo[e] = r.x; s_cnt[e] = (unsigned short)(r.y - r.x); s_pref[e + 1] = ((int)r.y - (int)r.x + RPW - 1) / RPW;
  }
  __syncthreads();
  if (threadIdx.x < 32) {  // warp 0: TMA bulk load of the band (32 KB pieces), then the prefix over views
    const unsigned total = bn * REC;
    if (threadIdx.x == 0) mbar_expect_tx(&bar, total);
    __syncwarp();
    const unsigned piece = 32768u;
    for (unsigned off = threadIdx.x * piece; off < total; off += 32u * piece)
      bulk_g2s(dsm + off, vsrc + off, min(piece, total - off), &bar);
    if (threadIdx.x == 0) {
      s_pref[0] = 0;
      for (int e = 0; e < nv; ++e) s_pref[e + 1] += s_pref[e];
    }
  }
  __syncthreads();
  const int total_items = s_pref[nv];
  mbar_wait(&bar, 0u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rl = lane / NL, q = lane - rl * NL;
  const bool lane_on = rl < RPW;
  const int lane0 = lane_on ? lane - q : lane;        // first lane of this ray's chunk group
  T* __restrict__ part = reinterpret_cast<T*>(Bt.part);
  const unsigned tile_s = smem_u32(dsm) + 16u * (unsigned)q;
  const uint16_t* __restrict__ fmap = (const uint16_t*)B.fmap;
  const uint4 none = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
  int vcur = 0;
  // items are handed out through a shared counter (their lengths differ with the chord of the band)
  for (int it = warp; it < total_items;) {
    while (it >= s_pref[vcur + 1]) ++vcur;
    const int a = s_ang[vcur];
    const int jlo = (int)s_jlo[vcur];
    const int j = jlo + RPW * (it - s_pref[vcur]) + rl;
    const bool ray_on = lane_on && j < jlo + (int)s_cnt[vcur];
    ushort2 sg = make_ushort2(0, 0);
    if (ray_on) sg = __ldg(Bt.seg + ((size_t)a * NB + b) * D2 + j);
    const int lo = sg.x, hi = sg.y;
    const bool has = hi > lo;
    // common origin of the warp: in step g every lane is on depth samples ib + 8 g ... ib + 8 g + 7
    const int ib = __reduce_min_sync(0xffffffffu, has ? (lo & ~7) : 0x7fffffff);
    const int gmax = __reduce_max_sync(0xffffffffu, has ? ((hi - ib + 7) >> 3) : 0);
    const int g_lo = has ? ((lo & ~7) - ib) >> 3 : 0x7fffffff;  // first / past-last 8-sample chunk holding samples of this ray
    const int g_hi = has ? ((hi - ib + 7) >> 3) : 0;
    const uint4* __restrict__ fp = reinterpret_cast<const uint4*>(fmap + ((size_t)a * D2 + (ray_on ? j : 0)) * D2 + (has ? ib : 0));
    vec_t acc = V::zero();
    uint4 nxt = none;
    if (q >= g_lo && q < g_hi) nxt = __ldg(fp + q);
    for (int g0 = 0; g0 < gmax; g0 += NL) {
      const uint4 mine = nxt;
      nxt = none;
      if (g0 + NL + q >= g_lo && g0 + NL + q < g_hi) nxt = __ldg(fp + g0 + NL + q);
#pragma unroll
      for (int sgi = 0; sgi < NL; ++sgi) {
        uint4 pk;
        if (NL == 1) pk = mine;
        else {
          pk.x = __shfl_sync(0xffffffffu, mine.x, lane0 + sgi); pk.y = __shfl_sync(0xffffffffu, mine.y, lane0 + sgi);
          pk.z = __shfl_sync(0xffffffffu, mine.z, lane0 + sgi); pk.w = __shfl_sync(0xffffffffu, mine.w, lane0 + sgi);
        }
        if (g0 + sgi >= gmax) break;  // warp-uniform
        if (!lane_on) pk = none;      // spare lanes (32 is not a multiple of NL) must not gather
        const unsigned wds[4] = {pk.x, pk.y, pk.z, pk.w};
        vec_t tv[8];
        bool ok[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const unsigned id = (e & 1) ? (wds[e >> 1] >> 16) : (wds[e >> 1] & 0xFFFFu);
          const unsigned rel = id - bb;
          ok[e] = rel < bn;
          tv[e] = V::zero();
          if (ok[e]) tv[e] = V::lds(tile_s + rel * REC);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (ok[e]) V::add(acc, tv[e]);
      }
    }
    if (ray_on) {
      const long long prow = Bt.view_poff[vb + vcur] + Bt.band_off[(size_t)a * NB + b] + (j - jlo);
      *reinterpret_cast<vec_t*>(part + (size_t)prow * L3P + V::N * q) = acc;
    }
    int nx = 0;
    if (lane == 0) nx = atomicAdd(&s_next, 1);
    it = __shfl_sync(0xffffffffu, nx, 0);
  }
}

// Row epilogue of the band path: a ray's partials are added over the bands it crosses, in band order, then
//   float32: the LSMR row update / plain store / score accumulation of k_fwd_data,
//   float64: the plain store of k_fwd64_data (dst <- A src),
// follows (duplicate views get the rows and partial sums of their first copy; half-set masks drop rows).  One CTA per
// (view of the candidate, candidate); a pure stream over the partials (8 independent 128-bit loads in flight per thread).
// (Tried: the last band CTA of a candidate doing this for the whole candidate while the partials are L2-resident --
// one SM cannot keep enough loads in flight, 140-270 us per candidate, forward 15.2 -> 18.5 us per candidate-pass.)
template <int NL, typename T, bool TRF>
__global__ void __launch_bounds__(HB2_BLOCK) k_fwd_band_reduce(BD B, TD Tt, T* __restrict__ dst_all, int mode) {
  typedef Vec16<T> V;
  typedef typename V::type vec_t;
  const BandTab& Bt = TRF ? B.bt64 : B.bt32;
  const int c = blockIdx.y, vi = blockIdx.x;
  __shared__ float red[HB2_BLOCK / 32];
  __shared__ ushort2 s_rng[HB2_MAX_BANDS];
  __shared__ int s_off[HB2_MAX_BANDS];
  __shared__ int s_colk[16];
  const int nv = B.cand_view_count[c];
  if (vi >= nv) return;
  const int view = B.cand_view_begin[c] + vi;
  if (B.view_tie && B.view_tie[view] >= 0) return;  // tie views: k_fwd_tie (rows and partials)
  if (B.view_dupof[view] >= 0) return;              // duplicate of an earlier view: written by that view's CTA
  int dupv[HB2_MAXDUP];
#pragma unroll
  for (int d = 0; d < HB2_MAXDUP; ++d) dupv[d] = B.view_dups[view * HB2_MAXDUP + d];
  if (!tie_active<TRF>(B, Tt, c, mode, false)) {
    if (!TRF && threadIdx.x == 0) {
#pragma unroll
      for (int d = -1; d < HB2_MAXDUP; ++d) {
        const int vw = d < 0 ? view : dupv[d < 0 ? 0 : d];
        if (vw < 0) continue;
        if (mode == MODE_LSMR) B.part_u[vw] = 0.f;
        if (mode == MODE_SCORE) { B.part_s[3 * vw] = 0.f; B.part_s[3 * vw + 1] = 0.f; B.part_s[3 * vw + 2] = 0.f; }
      }
    }
    return;
  }
  constexpr int L3P = NL * V::N;
  const int D2 = B.D2, NB = Bt.nband, L3 = B.L3;
  const int a = B.view_angle[view];
  for (int e = threadIdx.x; e < NB; e += HB2_BLOCK) { s_rng[e] = Bt.rng[(size_t)a * NB + e]; s_off[e] = Bt.band_off[(size_t)a * NB + e]; }
  if (threadIdx.x < 16) s_colk[threadIdx.x] = threadIdx.x < L3 ? B.colk[B.view_colbegin[view] + threadIdx.x] : -1;
  __syncthreads();
  const LsmrState& S = B.st[c];
  const float alpha = TRF ? 0.f : S.alpha, inv_beta = TRF ? 0.f : S.inv_beta;
  const T* __restrict__ part = reinterpret_cast<const T*>(Bt.part) + (size_t)Bt.view_poff[view] * L3P;
  T* urow = (TRF ? dst_all : (T*)(void*)B.u) + B.view_uoff[view];
  const float* brow = B.b + B.view_uoff[view];
  const uint8_t* __restrict__ pm = cand_mask(B, c);
  float ss = 0.f, s_pb = 0.f, s_bb = 0.f;
  for (int item = threadIdx.x; item < D2 * NL; item += HB2_BLOCK) {
    const int j = item / NL, q = item - j * NL;
    if (!B.rayvalid[a * D2 + j]) continue;  // no projection data: the padded rows stay 0 (SLR:1547)
    vec_t sum = V::zero();
    for (int b0 = 0; b0 < NB; b0 += 8) {  // 8 independent loads in flight, added in band order
      vec_t t[8];
      bool on[8];
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        on[w] = false;
        ushort2 r = make_ushort2(0, 0);
        if (b0 + w < NB) { r = s_rng[b0 + w]; on[w] = j >= (int)r.x && j < (int)r.y; }
        t[w] = V::zero();
        if (on[w]) t[w] = __ldcs(reinterpret_cast<const vec_t*>(part + ((size_t)s_off[b0 + w] + (j - (int)r.x)) * L3P + V::N * q));
      }
#pragma unroll
      for (int w = 0; w < 8; ++w)
        if (on[w]) V::add(sum, t[w]);
    }
    const size_t r0 = (size_t)j * L3P + V::N * q;
    bool keep[V::N];
#pragma unroll
    for (int tz = 0; tz < V::N; ++tz) {
      const int k = s_colk[V::N * q + tz];
      keep[tz] = k >= 0 && !(pm && !pm[(size_t)k * D2 + j]);
    }
    if constexpr (TRF) {
      const double sv[2] = {sum.x, sum.y};
      double2 uo = *reinterpret_cast<const double2*>(urow + r0);
      if (keep[0]) uo.x = sv[0];
      if (keep[1]) uo.y = sv[1];
      *reinterpret_cast<double2*>(urow + r0) = uo;
#pragma unroll
      for (int d = 0; d < HB2_MAXDUP; ++d)
        if (dupv[d] >= 0) *reinterpret_cast<double2*>(dst_all + B.view_uoff[dupv[d]] + r0) = uo;
    } else {
      const float sv[4] = {sum.x, sum.y, sum.z, sum.w};
      if (mode == MODE_LSMR || mode == MODE_PLAIN) {
        const float4 uo = *reinterpret_cast<const float4*>(urow + r0);
        float un[4] = {uo.x, uo.y, uo.z, uo.w};
#pragma unroll
        for (int tz = 0; tz < 4; ++tz) {
          if (!keep[tz]) continue;
          if (mode == MODE_LSMR) {
            un[tz] = fadd_(fmul_(fmul_(un[tz], inv_beta), -alpha), sv[tz]);
            ss += un[tz] * un[tz];
          } else {
            un[tz] = sv[tz];
          }
        }
        const float4 uw = make_float4(un[0], un[1], un[2], un[3]);
        *reinterpret_cast<float4*>(urow + r0) = uw;
#pragma unroll
        for (int d = 0; d < HB2_MAXDUP; ++d)
          if (dupv[d] >= 0) *reinterpret_cast<float4*>(B.u + B.view_uoff[dupv[d]] + r0) = uw;  // identical rows of the duplicate
      } else {
        const float4 bv4 = *reinterpret_cast<const float4*>(brow + r0);
        const float bv[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
#pragma unroll
        for (int tz = 0; tz < 4; ++tz) {
          if (!keep[tz]) continue;
          const float pred = B.clip_pred ? fmaxf(sv[tz], 0.f) : sv[tz];
          ss += pred * pred; s_pb += pred * bv[tz]; s_bb += bv[tz] * bv[tz];
        }
      }
    }
  }
  if constexpr (!TRF) {
    if (mode == MODE_LSMR) {
      float tot = block_sum(ss, red);
      if (threadIdx.x == 0) {
        B.part_u[view] = tot;
#pragma unroll
        for (int d = 0; d < HB2_MAXDUP; ++d)
          if (dupv[d] >= 0) B.part_u[dupv[d]] = tot;
      }
    } else if (mode == MODE_SCORE) {
      float t0 = block_sum(s_pb, red), t1 = block_sum(ss, red), t2 = block_sum(s_bb, red);
      if (threadIdx.x == 0) {
#pragma unroll
        for (int d = -1; d < HB2_MAXDUP; ++d) {
          const int vw = d < 0 ? view : dupv[d < 0 ? 0 : d];
          if (vw < 0) continue;
          B.part_s[3 * vw] = t0; B.part_s[3 * vw + 1] = t1; B.part_s[3 * vw + 2] = t2;
        }
      }
    }
  }
}
