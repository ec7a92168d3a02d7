// CUDA kernels of the denovo3D hot path (sm_100a).  See DESIGN.md for the
// data layout; reference line numbers are SLR = solver_linear_regression.py.
#pragma once
#include "hb2_common.cuh"

#define HB2_MAX_ZMC 1024   // L3*MC columns per view kept in shared memory
#define HB2_TILE_RAYS 32   // rays per CTA of the forward projector
#define HB2_BLOCK 256
#define HB2_FWD_U 8      // samples in flight per lane of the forward projector
#define HB2_ELL_W 4      // fixed-width part of the symmetry transpose lists
#define HB2_MAXDUP 3     // duplicates of one view served by its first copy
#define HB2_ELL_NONE 0x7FFFFFFF
#define HB2_ELL_OVERFLOW 0x7FFFFFFE  // in plane W-1: the voxel has more than W entries, use the CSR list

enum { MODE_LSMR = 0, MODE_PLAIN = 1, MODE_SCORE = 2, MODE_INIT = 3 };

// Tables of the forward band path for one element type (the float64 vectors need bands of half the voxels).
struct BandTab {
  int nband;                   // 0: not available
  const int* band_begin;       // [nband+1] first internal rank of every band (a contiguous run of the band-column-major order)
  const ushort2* seg;          // [nA][nband][D2] sample range [ilo, ihi) of ray j inside the band (ilo == ihi: none)
  const ushort2* rng;          // [nA][nband] rays [jlo, jhi) crossing the band
  const int* band_off;         // [nA][nband] partial-row offset of the band inside a view of that angle (rays before it)
  const long long* view_poff;  // [nviews] first partial row of the view
  void* part;                  // partial ray sums: [partial row][L3P] of the element type
};

// Batch descriptor passed by value to every kernel.
struct BD {
  // geometry (batch-uniform).  Voxel space is stored z-fastest: g = p*L3P + z (p = disk rank, L3P = L3 rounded
  // up to 4) so that one in-plane index lookup serves all slices with 128-bit loads; data rows likewise:
  // row = view_uoff + j*ZMP + zm with zm = z*MC + mc (ZMP = ZMC rounded up to 4).
  int D2, L2, D3, ndisk, L3, L3P, MC, ZMC, ZMP, n, nrp, npad, rows_per_view;
  int nA, K, nc;
  double s;
  // in-plane maps
  const void* fmap;            // [nA][D2][D2] disk rank or SENT
  const uint8_t* rayvalid;     // [nA][D2]
  const ushort2* rayrange;     // [nA][D2] depth samples [ilo, ihi) of ray j that land inside the disk
  const uint16_t* amap;        // [nA][K][apitch] ray j of the k-th sample landing in voxel p, or 0xFFFF
  int apitch;                  // row pitch of amap = ntile*256: every voxel tile owns a 512-byte aligned run of 256 slots
  const int* aslot;            // [ndisk] slot of voxel p in an amap row (tile*256 + rank inside the tile)
  const int* tile_begin;       // [ntile+1] first internal rank of every voxel tile (tile-major voxel order)
  int ntile;
  const uint16_t* tile_jlo;    // [nA][ntile] first ray of the view that crosses the tile
  const uint16_t* tile_nr;     // [nA][ntile] number of consecutive rays crossing it (0 = none)
  int rmax;                    // max of tile_nr
  // views (flat over the batch)
  const int* view_cand;
  const int* view_angle;
  const int* view_colbegin;
  const long long* view_uoff;  // absolute offset of the view's padded rows in u
  const int* colk;
  // candidates
  const int* cand_view_begin;
  const int* cand_view_count;
  const long long* cand_uoff;    // start of the candidate's rows in u
  const int* cand_mdata;         // padded data rows
  const long long* cand_symoff;  // offset into sym_a/sym_b
  const long long* cand_cscoff;  // offset into csc_ent
  int* cand_msym;                // symmetry rows (device-updated during setup)
  // vectors
  float* u;        // rows: [data padded | symmetry] per candidate
  float* b;        // same layout, data part only meaningful
  float* v;        // [nc][npad]
  float* h;        // [nc][npad]
  float* xs;       // [nc][npad] float32 copy of x (score / operator tests)
  double* x;       // [nc][npad]
  double* hbar;    // [nc][npad]
  // symmetry operator
  const int* sym_a;
  const int* sym_b;
  const int* csc_ptr;   // [nc][n+1]
  const int* csc_ent;   // row | sign<<31
  const int* ell;       // [nc][HB2_ELL_W][npad]: first entries of every voxel's transpose list (z-fastest like v)
  // solver state
  LsmrState* st;
  float* part_u;   // per-CTA partial sums of squares
  float* part_us;  // symmetry rows part
  float* part_v;
  double* part_x;
  float* part_s;   // score partials: 3 per CTA (dot, pp, bb)
  int part_u_n, part_us_per_cand, part_v_per_cand, part_x_per_cand;
  // Halton-duplicated views (SLR:1559-1571): identical rows, computed once
  const int* view_dupof;          // [nviews] view index of the first copy, or -1
  const int* view_mult;           // [nviews] 1 + number of duplicates (0 for a duplicate)
  const int* view_dups;           // [nviews][HB2_MAXDUP] view indices of the duplicates of a first copy, -1 padded
  // tie views (exact per-sample slices); null / 0 when the batch has none
  const int* view_tie;            // [nviews] tie index or -1
  const int* view_tie_slot0;      // [nviews]
  const signed char* tie_zlo;     // [n_tie][TS]
  const unsigned char* tie_up;    // [n_tie][TS][D2]
  const uint16_t* tie_upmask;     // [n_tie][TS/ZMC][D2]: bit t = tie_up of column slot g*ZMC + t (built at create)
  const int* tie_info;            // [nviews]: used-slot mask (low 16 bits) | (zb + 2) << 16 when the used slots sit on
                                  // consecutive slices from zb = -1 or 0 (fast path of the tie kernels), else high half 0
  const unsigned char* tie_rowvalid;  // [n_tie][TS][D2]
  int tie_TS, n_tie_views;
  const int* tie_views;           // [n_tie_views] view slots that are tie views, grouped by candidate
  const int* cand_tie_begin;      // [nc] range into tie_views
  const int* cand_tie_count;      // [nc]
  const uint16_t* amap_i;         // [nA][K][apitch] depth sample i of the entry in amap (built only with ties)
  float* vtie;                    // [nc][npad] adjoint contribution of the tie views (f32 path)
  double* vtie64;                 // same for the float64 operators of the bounded branch
  // forward band path (hb2_fwd_band.cuh): tables for the float32 (LSMR) and the float64 (bounded branch) operators
  int fwd_band;                // 1: the float32 forward uses the band path
  BandTab bt32, bt64;          // bt64.nband == 0: float64 forward stays on the gather kernel
  int fwd_ppv;                 // norm/score partials per view (1 on the band path, ceil(D2/32) otherwise)
  int adj_fast;    // MC == 1, K <= 2, row offsets fit 32 bits: k_adj_pq
  int adj_tile;    // additionally L3P <= 16, <= 256 views per candidate, windows fit shared memory: k_adj_tile
  int only_cand;   // MODE_PLAIN: restrict to one candidate (-1 all)
  int clip_pred;
  // half-set solves (fsc_test, SLR:175-203, 441-482): a candidate keeps only the data rows whose image pixel
  // pid = k*D2 + j is set in its mask; the other rows are absent (u, b stay 0 like every other padded row)
  const uint8_t* pixmask;       // [n_masks][L2*D2] or null
  const int* cand_pixmask;      // [nc] mask index or -1
  // explicit data rows (hb2_explicit.cuh): CSR of the rows (col = internal voxel index) and its transpose
  int exp_m;                    // number of explicit rows (0: matrix-free batch)
  int exp_m_data;               // the first exp_m_data of them are data rows (scored); the rest trilinear symmetry rows
  const int* exp_ptr;           // [exp_m + 1]
  const int* exp_col;
  const float* exp_w;
  const int* exp_cptr;          // [npad + 1]
  const int* exp_crow;
  const float* exp_cw;
  // matrix-free trilinear rows (hb2_bilinear.cuh): bilinear in-plane footprints per map + per-column slice blends
  int bil;                      // 1: every data view of the batch is a bilinear view (all views are pseudo views)
  int bil_KB;                   // (ray, weight) slots per (map, voxel) of the transposed maps
  const void* bilF_p;           // forward lists (one per pair of adjacent rays): voxel rank per entry (uint16 while the disk has < 65535 voxels, else uint32)
  const float2* bilF_w;         //                weights per entry: for ray 2P (.x) and ray 2P + 1 (.y) of the pair
  const int* bilF_ptr;          // [nM*NP + 1], NP = ceil(D2 / 2): entries of the ray pair (map, P)
  const uint16_t* bilT_j;       // [nM][KB][apitch] transposed maps: ray of the k-th entry of voxel p, 0xFFFF = none
  const float* bilT_w;          // [nM][KB][apitch]
  const uint8_t* bil_rayvalid;  // [nM][D2]
  const int* bil_view_map;      // [nviews] map of a bilinear view, -1 for the pseudo views of the trilinear symmetry rows
  const int* bil_colk;          // [nviews][ZMP] image column of slot t or -1
  const double* bil_ab;         // [nviews][ZMP][2] blend of slot t: a P[t-1] + b P[t] (t = 0: a P[0] + b P[1])
  const int* bil_cand_nview;    // [nc] bilinear views of the candidate (first in its view range)
  const uint16_t* bil_tile_jlo; // [nM][ntile] first ray of the map that touches the voxel tile (tile adjoint)
  const uint16_t* bil_tile_nr;  // [nM][ntile] number of consecutive rays touching it
  int bil_rmax;                 // max of bil_tile_nr
  int bil_adj_tile;             // 1: k_adj_bil_tile (else the gather kernel k_adj_bil)
  float* bil_ub;                // rows layout of u: the un-blended (slice-space) rows the adjoint gathers (k_bil_unblend)
  double* bil_ub64;             // same for the float64 operators of the bounded branch
  const int2* ls_ent;           // trilinear symmetry rows: 16 (internal voxel index, float weight bits) entries per row
  const long long* ls_eoff;     // [nc] first entry of candidate c relative to ls_ent (the candidates' arrays are separate allocations)
  const int* ls_m;              // [nc] rows
  const int* ls_cptr;           // [nc][npad + 1] transpose lists of candidate c: entry offsets relative to ls_ceoff[c]
  const int2* ls_cent;          // (row inside the candidate's symmetry block, float weight bits), sorted by voxel
  const long long* ls_ceoff;    // [nc]
};

// pixel mask of candidate c (null: all rows kept)
__device__ __forceinline__ const uint8_t* cand_mask(const BD& B, int c) {
  if (!B.pixmask) return nullptr;
  const int m = B.cand_pixmask[c];
  return m < 0 ? nullptr : B.pixmask + (size_t)m * B.L2 * B.D2;
}

template <typename T>
struct Sent;
template <>
struct Sent<uint16_t> { static constexpr uint16_t v = 0xFFFFu; };
template <>
struct Sent<uint32_t> { static constexpr uint32_t v = 0xFFFFFFFFu; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block reduction (blockDim.x == HB2_BLOCK), result valid in thread 0
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = l < (HB2_BLOCK / 32) ? sh[l] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_sum_d(double v, double* sh) {
  v = warp_sum_d(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = l < (HB2_BLOCK / 32) ? sh[l] : 0.0;
    r = warp_sum_d(r);
  }
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------
// TMA bulk copies (cp.async.bulk) + mbarrier helpers (sm_90+/sm_100a PTX)
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(unsigned saddr) {
  float4 r;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr));
  return r;
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ===========================================================================
// setup: in-plane sample -> voxel maps (replaces SLR:1514-1557 + 1614-1620)
// ===========================================================================
// One thread per (angle, ray j, depth i).  Coordinates follow the reference's
// float64 operation sequence: back_project (SLR:1712-1719, noise-free nominal
// column), Rotation.apply(inverse=True) = fma(M10,y,M00*x) / fma(M11,y,M01*x)
// (the order scipy executes, verified bit-exact in tests), + D2//2, rint.
template <typename IdxT>
__global__ void k_build_fmap(int nA, int D2, double s, const double* __restrict__ cs, const int* __restrict__ rank,
                             IdxT* __restrict__ fmap, uint8_t* __restrict__ rayvalid, int* __restrict__ tie) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)nA * D2 * D2;
  if (t >= total) return;
  int i = (int)(t % D2);
  int j = (int)((t / D2) % D2);
  int a = (int)(t / ((long long)D2 * D2));
  const int c0 = D2 / 2;
  double C = cs[2 * a], S = cs[2 * a + 1];
  // x' = fma(-1, Zd, e*Xl) with Xl = 0 (nominal column): -(i - c0); y' = j - c0
  double x0 = -(double)(i - c0), y0 = (double)(j - c0);
  if (s != 1.0) { x0 = __dmul_rn(x0, s); y0 = __dmul_rn(y0, s); }
  double X = __dadd_rn(__fma_rn(S, y0, __dmul_rn(C, x0)), (double)c0);
  double Y = __dadd_rn(__fma_rn(C, y0, __dmul_rn(-S, x0)), (double)c0);
  double xr = rint(X), yr = rint(Y);
  IdxT out = Sent<IdxT>::v;
  bool inb = xr >= 0.0 && xr <= (double)(D2 - 1) && yr >= 0.0 && yr <= (double)(D2 - 1);
  if (inb) {
    int r = rank[(int)yr * D2 + (int)xr];
    if (r >= 0) {
      out = (IdxT)r;
      rayvalid[a * D2 + j] = 1;  // benign race: all writers store 1
    }
  }
  fmap[t] = out;
  // tie detector: a rounding decision within 1e-9 of a boundary
  // may follow the reference's last-bit coordinate noise instead of ours.
  if (X > -1.0 && X < (double)D2 && Y > -1.0 && Y < (double)D2) {
    double fx = fabs(fabs(X - floor(X)) - 0.5), fy = fabs(fabs(Y - floor(Y)) - 0.5);
    if (fx < 1e-9 || fy < 1e-9) atomicAdd(&tie[a], 1);
  }
}

// Exact maps for in-plane tie views: same arithmetic as k_build_fmap, but the depth coordinate x0 of
// sample i comes from the reference's own coordinate table row of ONE image column (host: planner.reference_xz_tables;
// the table's last-bit noise depends on (column, i) only and decides the half-integer roundings at 30/60/... degrees).
template <typename IdxT>
__global__ void k_build_fmap_exact(int nE, int D2, double s, const double* __restrict__ cs, const double* __restrict__ x0tab,
                                   const int* __restrict__ rank, IdxT* __restrict__ fmap, uint8_t* __restrict__ rayvalid) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)nE * D2 * D2;
  if (t >= total) return;
  int i = (int)(t % D2);
  int j = (int)((t / D2) % D2);
  int a = (int)(t / ((long long)D2 * D2));
  const int c0 = D2 / 2;
  double C = cs[2 * a], S = cs[2 * a + 1];
  double x0 = x0tab[(size_t)a * D2 + i], y0 = (double)(j - c0);
  if (s != 1.0) y0 = __dmul_rn(y0, s);
  double X = __dadd_rn(__fma_rn(S, y0, __dmul_rn(C, x0)), (double)c0);
  double Y = __dadd_rn(__fma_rn(C, y0, __dmul_rn(-S, x0)), (double)c0);
  double xr = rint(X), yr = rint(Y);
  IdxT out = Sent<IdxT>::v;
  if (xr >= 0.0 && xr <= (double)(D2 - 1) && yr >= 0.0 && yr <= (double)(D2 - 1)) {
    int r = rank[(int)yr * D2 + (int)xr];
    if (r >= 0) {
      out = (IdxT)r;
      rayvalid[a * D2 + j] = 1;
    }
  }
  fmap[t] = out;
}

// Adjoint map, voxel-driven: for voxel p and angle a list the rays j of all
// samples (j,i) with fmap[a][j][i] == p, j-major then i (deterministic).
// pass 0: count only (max multiplicity -> *kmax); pass 1: fill amap[a][k][p].
template <typename IdxT>
__global__ void k_build_amap(int nA, int D2, int ndisk, int apitch, double s, int K, int pass, const double* __restrict__ cs,
                             const short2* __restrict__ disk_yx, const int* __restrict__ aslot, const IdxT* __restrict__ fmap,
                             uint16_t* __restrict__ amap, int* __restrict__ kmax, uint16_t* __restrict__ amap_i) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nA * ndisk) return;
  int p = (int)(t % ndisk), a = (int)(t / ndisk);
  const int c0 = D2 / 2;
  double C = cs[2 * a], S = cs[2 * a + 1];
  short2 yx = disk_yx[p];
  double dx = (double)(yx.y - c0), dy = (double)(yx.x - c0);  // yx.x = row y, yx.y = column x
  // inverse rotation: x0 = C*dx - S*dy, y0 = S*dx + C*dy ; i = c0 - x0/s ; j = c0 + y0/s
  double x0 = C * dx - S * dy, y0 = S * dx + C * dy;
  double is = c0 - x0 / s, js = c0 + y0 / s;
  double rad = 0.7072 / s + 0.02;
  int i0 = max(0, (int)ceil(is - rad)), i1 = min(D2 - 1, (int)floor(is + rad));
  int j0 = max(0, (int)ceil(js - rad)), j1 = min(D2 - 1, (int)floor(js + rad));
  const IdxT* fm = fmap + (size_t)a * D2 * D2;
  int cnt = 0;
  for (int j = j0; j <= j1; ++j)
    for (int i = i0; i <= i1; ++i)
      if (fm[(size_t)j * D2 + i] == (IdxT)p) {
        if (pass == 1 && cnt < K) {
          amap[((size_t)a * K + cnt) * apitch + aslot[p]] = (uint16_t)j;
          if (amap_i) amap_i[((size_t)a * K + cnt) * apitch + aslot[p]] = (uint16_t)i;
        }
        ++cnt;
      }
  if (pass == 0) {
    if (cnt > 0) atomicMax(kmax, cnt);
  } else {
    for (int k = cnt; k < K; ++k) amap[((size_t)a * K + k) * apitch + aslot[p]] = 0xFFFFu;
  }
}

// total samples with a hit per angle (consistency check of the adjoint map).  One CTA covers 256 consecutive
// entries of ONE angle (grid.y = angle), counts with a ballot and issues a single atomic.
template <typename IdxT>
__global__ void __launch_bounds__(HB2_BLOCK) k_count_hits(int nA, int D2, const IdxT* __restrict__ fmap,
                                                         unsigned long long* __restrict__ hits) {
  const int a = blockIdx.y;
  const long long per = (long long)D2 * D2;
  const long long e = (long long)blockIdx.x * HB2_BLOCK + threadIdx.x;
  const bool hit = e < per && fmap[(size_t)a * per + e] != Sent<IdxT>::v;
  const int n = __syncthreads_count(hit);
  if (threadIdx.x == 0 && n) atomicAdd(&hits[a], (unsigned long long)n);
}
__global__ void __launch_bounds__(HB2_BLOCK) k_count_amap(int nA, int K, int apitch, const uint16_t* __restrict__ amap,
                                                         unsigned long long* __restrict__ hits) {
  const int a = blockIdx.y;
  const long long per = (long long)K * apitch;
  const long long e = (long long)blockIdx.x * HB2_BLOCK + threadIdx.x;
  const bool hit = e < per && amap[(size_t)a * per + e] != 0xFFFFu;
  const int n = __syncthreads_count(hit);
  if (threadIdx.x == 0 && n) atomicAdd(&hits[a], (unsigned long long)n);
}

// right-hand side in padded layout: b[view][z][mc][j] = pix[j][k] when the
// column exists and the ray has data (SLR:1548), else 0.  Also max(b) per
// candidate (upper bound of the positive constraint, SLR:248).
__global__ void k_build_rhs(BD B, const float* __restrict__ pix, int nviews, float* __restrict__ bmax_bits) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)nviews * B.rows_per_view;
  if (t >= total) return;  // rows_per_view is a multiple of 4, blocks of 256: a warp may still straddle two views
  int view = (int)(t / B.rows_per_view);
  int r = (int)(t % B.rows_per_view);
  int zm = r % B.ZMP, j = r / B.ZMP;
  int k = zm < B.ZMC ? B.colk[B.view_colbegin[view] + zm] : -1;
  int a = B.view_angle[view];
  float val = 0.f;
  bool rowok = k >= 0 && B.rayvalid[a * B.D2 + j];
  if (rowok && B.view_tie && B.view_tie[view] >= 0)  // tie view: the row exists iff a sample of ITS slices hits
    rowok = B.tie_rowvalid[((size_t)B.view_tie[view] * B.tie_TS + B.view_tie_slot0[view] + zm) * B.D2 + j] != 0;
  if (rowok) {
    const uint8_t* pm = cand_mask(B, B.view_cand[view]);
    if (pm && !pm[(size_t)k * B.D2 + j]) rowok = false;  // row of the other half set
  }
  if (rowok) {
    val = pix[(size_t)j * B.L2 + k];
  }
  B.b[B.view_uoff[view] + r] = val;
  // max(b) per candidate: float max via the ordered-int trick, one atomic per group of lanes of the same candidate
  // (one atomic per row serialised 224 k atomics per candidate on a single address)
  {
    const int c = B.view_cand[view];
    int iv = (int)0x80000000;
    if (rowok) { iv = __float_as_int(val); iv = iv >= 0 ? iv : iv ^ 0x7fffffff; }
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, c);  // lanes of this warp that belong to the same candidate
    const int m = __reduce_max_sync(peers, iv);
    if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31) && m != (int)0x80000000) atomicMax((int*)bmax_bits + c, m);
  }
}

// ===========================================================================
// symmetry rows (replaces SLR:1142-1218, 1221-1287)
// ===========================================================================
struct SymSetup {
  const double* pairs;          // [npairs][6] ci,si,zi,cj,sj,zj
  const int* pair_begin;
  const int* pair_count;
  const long long* min_pairs;
  const long long* tab_off;     // hash table offset per candidate
  const long long* tab_cap;
  unsigned long long* tab_key;
  unsigned long long* tab_seq;
  int* tmp_a;                   // [nc][npad]
  int* tmp_b;
  int* flag;                    // [nc][npad]
  int* pos;                     // exclusive scan of flag
  int* done;                    // [nc]
  int* ndone;                   // [1]
  int* sym_a_w;
  int* sym_b_w;
  const int* rank_sym;          // [D3*D3]
  const short2* disk_yx_sym;    // [ndisk]
  const int* int2ref;           // [ndisk] internal disk rank -> reference rank
  int* sym_g_w;                 // per kept row: reference index g of the voxel that generated it (export order)
  const long long* symcap;
  int* overflow;
};

#define HB2_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ unsigned long long hash64(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}

// image of voxel (zc,yc,xc centred) under one pair member, SLR:1225-1243:
//   X = fma(-S,y,C*x) + D3//2 ; Y = fma(C,y,S*x) + D3//2 ; Z = (z + L3//2) + rise*h ; rint each.
__device__ __forceinline__ int sym_image(double C, double S, double zs, int xc, int yc, int zc, int D3, int L3,
                                         int L3P, const int* __restrict__ rank) {
  double x = (double)xc, y = (double)yc;
  double X = __dadd_rn(__fma_rn(-S, y, __dmul_rn(C, x)), (double)(D3 / 2));
  double Y = __dadd_rn(__fma_rn(C, y, __dmul_rn(S, x)), (double)(D3 / 2));
  double Z = __dadd_rn(__dadd_rn((double)zc, (double)(L3 / 2)), zs);
  double xr = rint(X), yr = rint(Y), zr = rint(Z);
  if (!(zr >= 0.0 && zr <= (double)(L3 - 1) && yr >= 0.0 && yr <= (double)(D3 - 1) && xr >= 0.0 &&
        xr <= (double)(D3 - 1)))
    return -1;
  int r = rank[(int)yr * D3 + (int)xr];
  if (r < 0) return -1;
  return r * L3P + (int)zr;  // internal (z-fastest) voxel index
}

// round `rnd`: every still-open candidate processes its rnd-th pair.
// phase 0: compute (a,b) per voxel, insert unordered key with min sequence number.
__global__ void k_sym_insert(BD B, SymSetup Q, int rnd) {
  int c = blockIdx.y;
  if (Q.done[c]) return;
  // Threads (and with them the kept rows of a pair block) enumerate the voxels in INTERNAL order, z fastest: the rows
  // generated by one voxel column are consecutive and their images a, b are runs of consecutive floats of v, so the
  // gathers of k_fwd_sym and the row reads of the adjoint lists are (nearly) sequential.  The reference's enumeration
  // index g (z slowest, mask order) still decides which duplicate survives (seq) and is kept per row for the exports.
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B.n) return;
  const double* pr = Q.pairs + (size_t)(Q.pair_begin[c] + rnd) * 6;
  int z = t % B.L3, p = Q.int2ref[t / B.L3];
  int g = z * B.ndisk + p;
  short2 yx = Q.disk_yx_sym[p];
  int xc = yx.y - B.D3 / 2, yc = yx.x - B.D3 / 2, zc = z - B.L3 / 2;
  int a = sym_image(pr[0], pr[1], pr[2], xc, yc, zc, B.D3, B.L3, B.L3P, Q.rank_sym);
  int b = sym_image(pr[3], pr[4], pr[5], xc, yc, zc, B.D3, B.L3, B.L3P, Q.rank_sym);
  size_t ti = (size_t)c * B.nrp + t;
  if (a < 0 || b < 0) {
    Q.tmp_a[ti] = -1; Q.tmp_b[ti] = -1;
    return;
  }
  Q.tmp_a[ti] = a; Q.tmp_b[ti] = b;
  unsigned lo = (unsigned)min(a, b), hi = (unsigned)max(a, b);
  unsigned long long key = ((unsigned long long)lo << 32) | hi;
  unsigned long long seq = (unsigned long long)rnd * (unsigned long long)B.n + (unsigned long long)g;
  unsigned long long cap = (unsigned long long)Q.tab_cap[c];
  unsigned long long* keys = Q.tab_key + Q.tab_off[c];
  unsigned long long* seqs = Q.tab_seq + Q.tab_off[c];
  unsigned long long slot = hash64(key) % cap;
  for (unsigned long long probe = 0; probe < cap; ++probe) {
    unsigned long long old = atomicCAS(&keys[slot], HB2_EMPTY_KEY, key);
    if (old == HB2_EMPTY_KEY || old == key) {
      atomicMin(&seqs[slot], seq);
      return;
    }
    slot = slot + 1 == cap ? 0 : slot + 1;
  }
  atomicExch(Q.overflow, 1);
}

// phase 1: a candidate row survives iff its sequence number is the table minimum
// (first-seen-wins over pairs in list order, then voxels in mask order, SLR:1197-1202).
__global__ void k_sym_check(BD B, SymSetup Q, int rnd) {
  int c = blockIdx.y;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B.nrp) return;
  size_t ti = (size_t)c * B.nrp + t;
  int keep = 0;
  if (!Q.done[c] && t < B.n) {
    const int g = (t % B.L3) * B.ndisk + Q.int2ref[t / B.L3];
    int a = Q.tmp_a[ti], b = Q.tmp_b[ti];
    if (a >= 0) {
      unsigned lo = (unsigned)min(a, b), hi = (unsigned)max(a, b);
      unsigned long long key = ((unsigned long long)lo << 32) | hi;
      unsigned long long seq = (unsigned long long)rnd * (unsigned long long)B.n + (unsigned long long)g;
      unsigned long long cap = (unsigned long long)Q.tab_cap[c];
      const unsigned long long* keys = Q.tab_key + Q.tab_off[c];
      const unsigned long long* seqs = Q.tab_seq + Q.tab_off[c];
      unsigned long long slot = hash64(key) % cap;
      for (unsigned long long probe = 0; probe < cap; ++probe) {
        unsigned long long kk = keys[slot];
        if (kk == key) { keep = seqs[slot] == seq; break; }
        if (kk == HB2_EMPTY_KEY) break;
        slot = slot + 1 == cap ? 0 : slot + 1;
      }
    }
  }
  Q.flag[ti] = keep;
}

// phase 2: ordered compaction (row order = voxel order inside the pair block)
__global__ void k_sym_compact(BD B, SymSetup Q) {
  int c = blockIdx.y;
  if (Q.done[c]) return;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B.n) return;
  size_t ti = (size_t)c * B.nrp + t;
  if (!Q.flag[ti]) return;
  long long row = (long long)B.cand_msym[c] + (Q.pos[ti] - Q.pos[(size_t)c * B.nrp]);
  if (row >= Q.symcap[c]) { atomicExch(Q.overflow, 2); return; }
  Q.sym_a_w[B.cand_symoff[c] + row] = Q.tmp_a[ti];
  Q.sym_b_w[B.cand_symoff[c] + row] = Q.tmp_b[ti];
  Q.sym_g_w[B.cand_symoff[c] + row] = (t % B.L3) * B.ndisk + Q.int2ref[t / B.L3];
}

// phase 3: row count + early stop (SLR:1286) per candidate
__global__ void k_sym_finalize(BD B, SymSetup Q, int rnd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= B.nc || Q.done[c]) return;
  size_t last = (size_t)c * B.nrp + (B.nrp - 1);
  int cnt = Q.pos[last] + Q.flag[last] - Q.pos[(size_t)c * B.nrp];
  int tot = B.cand_msym[c] + cnt;
  B.cand_msym[c] = tot;
  if ((long long)tot >= Q.min_pairs[c] || rnd + 1 >= Q.pair_count[c]) {
    Q.done[c] = 1;
    atomicAdd(Q.ndone, 1);
  }
}

// transpose lists of the symmetry rows: per voxel the incident rows (+ sign)
__global__ void k_csc_count(BD B, int* __restrict__ cnt) {
  int c = blockIdx.y;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B.cand_msym[c]) return;
  int a = B.sym_a[B.cand_symoff[c] + r], b = B.sym_b[B.cand_symoff[c] + r];
  atomicAdd(&cnt[(size_t)c * (B.npad + 1) + a], 1);
  atomicAdd(&cnt[(size_t)c * (B.npad + 1) + b], 1);
}
__global__ void k_csc_rebase2(BD B, const int* __restrict__ scan, int* __restrict__ ptr) {
  int c = blockIdx.y;
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > B.npad) return;
  ptr[(size_t)c * (B.npad + 1) + g] = scan[(size_t)c * (B.npad + 1) + g] - scan[(size_t)c * (B.npad + 1)];
}
__global__ void k_csc_fill(BD B, int* __restrict__ cursor, int* __restrict__ ent) {
  int c = blockIdx.y;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B.cand_msym[c]) return;
  int a = B.sym_a[B.cand_symoff[c] + r], b = B.sym_b[B.cand_symoff[c] + r];
  const int* ptr = B.csc_ptr + (size_t)c * (B.npad + 1);
  int sa = atomicAdd(&cursor[(size_t)c * (B.npad + 1) + a], 1);
  ent[B.cand_cscoff[c] + ptr[a] + sa] = r;
  int sb = atomicAdd(&cursor[(size_t)c * (B.npad + 1) + b], 1);
  ent[B.cand_cscoff[c] + ptr[b] + sb] = r | (int)0x80000000;
}
__global__ void k_csc_sort(BD B, int* __restrict__ ent) {  // deterministic order: by row index
  int c = blockIdx.y;
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= B.npad) return;
  const int* ptr = B.csc_ptr + (size_t)c * (B.npad + 1);
  int e0 = ptr[g], e1 = ptr[g + 1];
  int* e = ent + B.cand_cscoff[c];
  for (int i = e0 + 1; i < e1; ++i) {
    int key = e[i];
    unsigned kr = ((unsigned)key & 0x7fffffffu) * 2u + ((unsigned)key >> 31);
    int j = i - 1;
    while (j >= e0) {
      unsigned jr = ((unsigned)e[j] & 0x7fffffffu) * 2u + ((unsigned)e[j] >> 31);
      if (jr <= kr) break;
      e[j + 1] = e[j];
      --j;
    }
    e[j + 1] = key;
  }
}

// fixed-width copy of the (sorted) transpose lists: ell[c][w][g] = w-th entry of voxel g, HB2_ELL_NONE beyond the
// list; voxels with more than HB2_ELL_W entries get HB2_ELL_OVERFLOW in the last plane (the adjoint then walks the
// CSR list).  Measured on cfg-like geometries: 2-3 entries per voxel, > 4 for < 0.01 % of the voxels.
__global__ void k_ell_fill(BD B, int* __restrict__ ell) {
  int c = blockIdx.y;
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= B.npad) return;
  const int* ptr = B.csc_ptr + (size_t)c * (B.npad + 1);
  const int* e = B.csc_ent + B.cand_cscoff[c];
  const int e0 = ptr[g], cnt = ptr[g + 1] - e0;
  int* dst = ell + (size_t)c * HB2_ELL_W * B.npad + g;
#pragma unroll
  for (int w = 0; w < HB2_ELL_W; ++w) {
    int val = HB2_ELL_NONE;
    if (cnt <= HB2_ELL_W) { if (w < cnt) val = e[e0 + w]; }
    else if (w == HB2_ELL_W - 1) val = HB2_ELL_OVERFLOW;
    dst[(size_t)w * B.npad] = val;
  }
}

// ===========================================================================
// forward projector: data rows.  One CTA = 32 consecutive rays of one view, one
// warp per ray.  Lanes are split 8 sample groups x 4 slice quads: per step the
// warp reads 8 map entries and, for each, one 128-bit load per quad gathers 4
// slices of the voxel (z-fastest layout), so one index lookup serves up to 16
// slices.  The 8 sample groups are combined with 3 shuffle steps.
// MODE_LSMR : u~ <- A v - alpha * (u~ * inv_beta)       (lsmr.py:331-332)
// MODE_PLAIN: u  <- A xs
// MODE_SCORE: accumulate <pred,b>, <pred,pred>, <b,b>   (SLR:500-525)
// ===========================================================================
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Sample range of every ray: [ilo, ihi) bounds the depth samples of ray j that land inside the disk (the map row is
// SENT outside).  One warp per (angle, ray).
template <typename IdxT>
__global__ void __launch_bounds__(HB2_BLOCK) k_ray_range(int nA, int D2, const IdxT* __restrict__ fmap, ushort2* __restrict__ rr) {
  const long long w = ((long long)blockIdx.x * HB2_BLOCK + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)nA * D2) return;
  const IdxT* __restrict__ fj = fmap + (size_t)w * D2;
  int lo = D2, hi = 0;
  for (int i = lane; i < D2; i += 32)
    if (fj[i] != Sent<IdxT>::v) { lo = min(lo, i); hi = max(hi, i + 1); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) rr[w] = make_ushort2((unsigned short)min(lo, hi), (unsigned short)hi);
}

// Lanes of a warp = SG sample groups x NQL slice quads (lane = sg * NQL + q): the NQL lanes of a sample read the
// 16 * NQL contiguous bytes of its voxel record, consecutive sample groups read consecutive samples of the ray, i.e.
// mostly neighbouring voxels of a tile row.  NQL = L3P / 4 for L3P <= 16, so no lane idles on a slice quad that does
// not exist (L3P = 12: 10 x 3 lanes busy instead of 8 x 3 of 8 x 4; profiles/r2_summary.md); larger L3P run in
// passes of 4 quads.
template <typename IdxT, int NQL>
__global__ void __launch_bounds__(HB2_BLOCK) k_fwd_data(BD B, int mode) {
  constexpr int SG = 32 / NQL;      // sample groups: 32, 16, 10, 8
  constexpr int NACT = SG * NQL;    // busy lanes
  constexpr int UN = NQL >= 3 ? HB2_FWD_U : (NQL == 2 ? 6 : 4);  // samples in flight per lane
  const int ntiles = (B.D2 + HB2_TILE_RAYS - 1) / HB2_TILE_RAYS;
  const int view = blockIdx.x / ntiles, tile = blockIdx.x % ntiles;
  const int c = B.view_cand[view];
  if (B.view_tie && B.view_tie[view] >= 0) return;  // tie views: k_fwd_tie (rows and partials)
  if (B.view_dupof[view] >= 0) return;              // duplicate of an earlier view: served by that view's CTAs
  __shared__ float red[HB2_BLOCK / 32];
  __shared__ int s_colk[HB2_MAX_ZMC];
  int dupv[HB2_MAXDUP];
#pragma unroll
  for (int d = 0; d < HB2_MAXDUP; ++d) dupv[d] = B.view_dups[view * HB2_MAXDUP + d];
  const LsmrState& S = B.st[c];
  bool act = mode == MODE_LSMR ? (S.active != 0) : (B.only_cand < 0 || B.only_cand == c);
  if (!act) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int d = -1; d < HB2_MAXDUP; ++d) {
        const int vw = d < 0 ? view : dupv[d < 0 ? 0 : d];
        if (vw < 0) continue;
        const int pb = vw * ntiles + tile;
        if (mode == MODE_LSMR) B.part_u[pb] = 0.f;
        if (mode == MODE_SCORE) { B.part_s[3 * pb] = 0.f; B.part_s[3 * pb + 1] = 0.f; B.part_s[3 * pb + 2] = 0.f; }
      }
    }
    return;
  }
  for (int e = threadIdx.x; e < B.ZMC; e += HB2_BLOCK) s_colk[e] = B.colk[B.view_colbegin[view] + e];
  __syncthreads();
  const float alpha = S.alpha, inv_beta = S.inv_beta;
  const int a = B.view_angle[view];
  const int D2 = B.D2, L3 = B.L3, L3P = B.L3P, MC = B.MC, ZMP = B.ZMP;
  const IdxT* __restrict__ fm = (const IdxT*)B.fmap + (size_t)a * D2 * D2;
  const float* __restrict__ vsrc = (mode == MODE_LSMR ? B.v : B.xs) + (size_t)c * B.npad;
  float* urow = B.u + B.view_uoff[view];
  const float* brow = B.b + B.view_uoff[view];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sg = lane / NQL, q = lane - sg * NQL;
  const bool lane_on = lane < NACT;
  const uint8_t* __restrict__ pm = cand_mask(B, c);
  float ss = 0.f, s_pb = 0.f, s_bb = 0.f;
  for (int r = warp; r < HB2_TILE_RAYS; r += HB2_BLOCK / 32) {
    const int j = tile * HB2_TILE_RAYS + r;
    if (j >= D2) break;
    if (!B.rayvalid[a * D2 + j]) continue;  // no projection data: the padded rows stay 0 (SLR:1547)
    const IdxT* __restrict__ fj = fm + (size_t)j * D2;
    const ushort2 rng = B.rayrange[(size_t)a * D2 + j];
    const int ilo = rng.x, ihi = rng.y;
    for (int z0 = 0; z0 < L3P; z0 += 4 * NQL) {
      const int zq = z0 + 4 * q;  // this lane's 4 slices
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lane_on && zq < L3P) {
        const float* __restrict__ vb = vsrc + zq;
        // UN map entries, then their gathers, are issued before the first add (independent loads in flight);
        // the additions keep the sample order.
        for (int i0 = ilo + sg; i0 < ihi; i0 += SG * UN) {
          IdxT ids[UN];
#pragma unroll
          for (int w = 0; w < UN; ++w) ids[w] = (i0 + SG * w < ihi) ? fj[i0 + SG * w] : Sent<IdxT>::v;
          float4 tt[UN];
#pragma unroll
          for (int w = 0; w < UN; ++w)
            tt[w] = ids[w] != Sent<IdxT>::v ? ldg4(vb + (size_t)ids[w] * L3P) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int w = 0; w < UN; ++w)
            if (ids[w] != Sent<IdxT>::v) { acc.x += tt[w].x; acc.y += tt[w].y; acc.z += tt[w].z; acc.w += tt[w].w; }
        }
      }
      // fold the sample groups (fixed tree: deterministic): after the loop lane q (sg = 0) holds quad q's 4 sums
#pragma unroll
      for (int s2 = (SG > 16 ? 16 : (SG > 8 ? 8 : (SG > 4 ? 4 : 2))); s2 >= 1; s2 >>= 1) {
        const float4 o = make_float4(__shfl_down_sync(0xffffffffu, acc.x, s2 * NQL), __shfl_down_sync(0xffffffffu, acc.y, s2 * NQL),
                                     __shfl_down_sync(0xffffffffu, acc.z, s2 * NQL), __shfl_down_sync(0xffffffffu, acc.w, s2 * NQL));
        if (sg < s2 && sg + s2 < SG && lane_on) { acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
      }
      // lane t < 4 * NQL finishes slice z0 + t: component t & 3 of lane t >> 2
      const int src_lane = (lane >> 2) < NQL ? (lane >> 2) : 0;
      const float c0 = __shfl_sync(0xffffffffu, acc.x, src_lane), c1 = __shfl_sync(0xffffffffu, acc.y, src_lane);
      const float c2 = __shfl_sync(0xffffffffu, acc.z, src_lane), c3 = __shfl_sync(0xffffffffu, acc.w, src_lane);
      if (lane < 4 * NQL) {
        const int z = z0 + lane;
        if (z < L3) {
          const int t = lane & 3;
          const float sum = t == 0 ? c0 : (t == 1 ? c1 : (t == 2 ? c2 : c3));
          for (int mc = 0; mc < MC; ++mc) {
            const int zm = z * MC + mc;
            if (s_colk[zm] < 0) continue;
            if (pm && !pm[(size_t)s_colk[zm] * D2 + j]) continue;
            const size_t ri = (size_t)j * ZMP + zm;
            if (mode == MODE_LSMR) {
              float un = fadd_(fmul_(fmul_(urow[ri], inv_beta), -alpha), sum);
              urow[ri] = un;
#pragma unroll
              for (int d = 0; d < HB2_MAXDUP; ++d)
                if (dupv[d] >= 0) (B.u + B.view_uoff[dupv[d]])[ri] = un;  // identical row of the duplicate view
              ss += un * un;
            } else if (mode == MODE_PLAIN) {
              urow[ri] = sum;
#pragma unroll
              for (int d = 0; d < HB2_MAXDUP; ++d)
                if (dupv[d] >= 0) (B.u + B.view_uoff[dupv[d]])[ri] = sum;
            } else {
              float pred = B.clip_pred ? fmaxf(sum, 0.f) : sum;
              float bv = brow[ri];
              ss += pred * pred; s_pb += pred * bv; s_bb += bv * bv;
            }
          }
        }
      }
    }
  }
  // partials of the view and of the duplicates it serves (identical rows -> identical partial sums)
  if (mode == MODE_LSMR) {
    float tot = block_sum(ss, red);
    if (threadIdx.x == 0) {
      B.part_u[blockIdx.x] = tot;
#pragma unroll
      for (int d = 0; d < HB2_MAXDUP; ++d)
        if (dupv[d] >= 0) B.part_u[dupv[d] * ntiles + tile] = tot;
    }
  } else if (mode == MODE_SCORE) {
    float t0 = block_sum(s_pb, red), t1 = block_sum(ss, red), t2 = block_sum(s_bb, red);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int d = -1; d < HB2_MAXDUP; ++d) {
        const int vw = d < 0 ? view : dupv[d < 0 ? 0 : d];
        if (vw < 0) continue;
        const int pb = vw * ntiles + tile;
        B.part_s[3 * pb] = t0; B.part_s[3 * pb + 1] = t1; B.part_s[3 * pb + 2] = t2;
      }
    }
  }
}

// forward, symmetry rows: u~[r] <- (v[a]-v[b]) - alpha*(u~[r]*inv_beta)
__global__ void __launch_bounds__(HB2_BLOCK) k_fwd_sym(BD B, int mode) {
  const int c = blockIdx.y;
  __shared__ float red[HB2_BLOCK / 32];
  const LsmrState& S = B.st[c];
  bool act = mode == MODE_LSMR ? (S.active != 0) : (B.only_cand < 0 || B.only_cand == c);
  const int pi = c * B.part_us_per_cand + blockIdx.x;
  if (!act) {
    if (threadIdx.x == 0 && mode == MODE_LSMR) B.part_us[pi] = 0.f;
    return;
  }
  const int m = B.cand_msym[c];
  const float alpha = S.alpha, inv_beta = S.inv_beta;
  const float* __restrict__ vsrc = (mode == MODE_LSMR ? B.v : B.xs) + (size_t)c * B.npad;
  float* us = B.u + B.cand_uoff[c] + B.cand_mdata[c];
  const int* __restrict__ sa = B.sym_a + B.cand_symoff[c];
  const int* __restrict__ sb = B.sym_b + B.cand_symoff[c];
  float ss = 0.f;
  for (int r = blockIdx.x * HB2_BLOCK * 4 + threadIdx.x, q = 0; q < 4; ++q, r += HB2_BLOCK) {
    if (r < m) {
      float d = fadd_(__ldg(vsrc + sa[r]), -__ldg(vsrc + sb[r]));
      if (mode == MODE_LSMR) {
        float un = fadd_(fmul_(fmul_(us[r], inv_beta), -alpha), d);
        us[r] = un;
        ss += un * un;
      } else {
        us[r] = d;
      }
    }
  }
  if (mode == MODE_LSMR) {
    float tot = block_sum(ss, red);
    if (threadIdx.x == 0) B.part_us[pi] = tot;
  }
}

// ===========================================================================
// adjoint: voxel-driven gather.  One thread = one in-plane voxel p and one slice
// quad q (4 slices); the NQ = L3P/4 lanes of a voxel are adjacent, so a warp
// covers ~32/NQ consecutive voxels of a disk row.  Per view the lanes read the
// adjoint map (the rays whose samples land in p) and gather their 16 bytes of
// the ray's row: neighbouring voxels map to the same or the neighbouring ray,
// whose rows are contiguous ([view][j][z]), so one warp-wide 128-bit gather
// touches few 128-byte lines.  Then the symmetry rows incident to each voxel.
// MODE_LSMR : v~ <- A^T (u~*inv_beta) - beta * v        (lsmr.py:336-338)
// MODE_INIT : v~ <- A^T (u~*inv_beta)                    (lsmr.py:251)
// MODE_PLAIN: xs <- A^T u
// ===========================================================================
#define HB2_ADJ_VIEWS 128  // views staged in shared memory per pass
template <int KT, int MCT>
__global__ void __launch_bounds__(HB2_BLOCK) k_adj(BD B, int mode) {
  const int c = blockIdx.y;
  __shared__ float red[HB2_BLOCK / 32];
  __shared__ int s_ang[HB2_ADJ_VIEWS];
  __shared__ long long s_uoff[HB2_ADJ_VIEWS];
  __shared__ float s_w[HB2_ADJ_VIEWS];
  const LsmrState& S = B.st[c];
  bool act = (mode == MODE_LSMR) ? (S.active != 0 && !S.skip_adj)
                                 : (mode == MODE_INIT ? (S.beta > 0.f) : (B.only_cand < 0 || B.only_cand == c));
  const int pi = c * B.part_v_per_cand + blockIdx.x;
  if (!act) {
    if (threadIdx.x == 0 && mode != MODE_PLAIN) B.part_v[pi] = 0.f;
    return;
  }
  const int L3 = B.L3, L3P = B.L3P, ZMP = B.ZMP, ndisk = B.ndisk;
  const int K = KT > 0 ? KT : B.K, MC = MCT > 0 ? MCT : B.MC;
  const int NQ = L3P >> 2;
  const int t = blockIdx.x * HB2_BLOCK + threadIdx.x;
  const int p = t / NQ, q = t - p * NQ;
  const bool live = p < ndisk;
  const int slot = live ? B.aslot[p] : 0;
  const int z0 = 4 * q;
  const float ib = mode == MODE_PLAIN ? 1.f : S.inv_beta;
  // Halton duplicates carry identical rows inside the solver (u starts as b and is updated row-wise), so the
  // first copy can stand for them with its multiplicity; an arbitrary row vector (MODE_PLAIN, the API's
  // apply_adjoint) must honour every row on its own.
  const bool dedupe_adj = mode != MODE_PLAIN;
  const float beta = S.beta;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  const int vb = B.cand_view_begin[c], nv = B.cand_view_count[c];
  for (int v0 = 0; v0 < nv; v0 += HB2_ADJ_VIEWS) {
    const int nvc = min(HB2_ADJ_VIEWS, nv - v0);
    __syncthreads();
    for (int e = threadIdx.x; e < nvc; e += HB2_BLOCK) {
      const bool skip = (B.view_tie && B.view_tie[vb + v0 + e] >= 0) || (dedupe_adj && B.view_dupof[vb + v0 + e] >= 0);
      s_ang[e] = skip ? -1 : B.view_angle[vb + v0 + e];
      s_uoff[e] = B.view_uoff[vb + v0 + e];
      s_w[e] = dedupe_adj ? (float)B.view_mult[vb + v0 + e] : 1.f;
    }
    __syncthreads();
    if (!live) continue;
    for (int vi = 0; vi < nvc; ++vi) {
      if (s_ang[vi] < 0) continue;  // tie view (k_adj_tie) or duplicate (weight of its first copy)
      const float ibw = ib * s_w[vi];
      const float* __restrict__ ub = B.u + s_uoff[vi] + z0 * MC;
      const uint16_t* __restrict__ am = B.amap + (size_t)s_ang[vi] * K * B.apitch + slot;
      for (int k = 0; k < K; ++k) {
        const uint16_t j = am[(size_t)k * B.apitch];
        if (j != 0xFFFFu) {
          const float* __restrict__ uj = ub + (size_t)j * ZMP;
          if (MCT == 1) {
            const float4 r4 = ldg4(uj);
            acc0 = fmaf(r4.x, ibw, acc0); acc1 = fmaf(r4.y, ibw, acc1);
            acc2 = fmaf(r4.z, ibw, acc2); acc3 = fmaf(r4.w, ibw, acc3);
          } else {
            for (int mc = 0; mc < MC; ++mc) {
              if (z0 + 0 < L3) acc0 = fmaf(__ldg(uj + 0 * MC + mc), ibw, acc0);
              if (z0 + 1 < L3) acc1 = fmaf(__ldg(uj + 1 * MC + mc), ibw, acc1);
              if (z0 + 2 < L3) acc2 = fmaf(__ldg(uj + 2 * MC + mc), ibw, acc2);
              if (z0 + 3 < L3) acc3 = fmaf(__ldg(uj + 3 * MC + mc), ibw, acc3);
            }
          }
        }
      }
    }
  }
  float ss = 0.f;
  if (live) {
    // symmetry rows incident to (p, z0..z0+3)
    const int* __restrict__ ptr = B.csc_ptr + (size_t)c * (B.npad + 1);
    const int* __restrict__ ent = B.csc_ent + B.cand_cscoff[c];
    const float* __restrict__ us = B.u + B.cand_uoff[c] + B.cand_mdata[c];
    float* vdst = (mode == MODE_PLAIN ? B.xs : B.v) + (size_t)c * B.npad + (size_t)p * L3P + z0;
    const int g0 = p * L3P + z0;
    float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mode == MODE_LSMR) old = *reinterpret_cast<const float4*>(vdst);
    if (B.vtie && B.cand_tie_count[c] > 0) {  // rows of the tie views (k_adj_tie)
      const float4 t = *reinterpret_cast<const float4*>(B.vtie + (size_t)c * B.npad + g0);
      acc0 += t.x; acc1 += t.y; acc2 += t.z; acc3 += t.w;
    }
    float vn[4] = {acc0, acc1, acc2, acc3};
    const int nz = min(4, L3 - z0);  // real slices of this quad (padded slices stay 0)
    int e = nz > 0 ? ptr[g0] : 0;
#pragma unroll
    for (int tz = 0; tz < 4; ++tz) {
      float a2 = vn[tz];
      if (tz < nz) {
        const int e1 = ptr[g0 + tz + 1];
        for (; e < e1; ++e) {
          const int en = ent[e];
          const float val = __ldg(us + (en & 0x7fffffff));
          a2 = fmaf(en < 0 ? -val : val, ib, a2);
        }
      }
      const float o = tz == 0 ? old.x : (tz == 1 ? old.y : (tz == 2 ? old.z : old.w));
      vn[tz] = mode == MODE_LSMR ? fadd_(fmul_(o, -beta), a2) : a2;
      ss += vn[tz] * vn[tz];
    }
    *reinterpret_cast<float4*>(vdst) = make_float4(vn[0], vn[1], vn[2], vn[3]);
  }
  if (mode != MODE_PLAIN) {
    float tot = block_sum(ss, red);
    if (threadIdx.x == 0) B.part_v[pi] = tot;
  }
}

// ---------------------------------------------------------------------------
// Adjoint, fast path (MC == 1, K <= 2).  One thread = one in-plane voxel p and one
// slice quad q; the NQ = L3P/4 lanes of a voxel are adjacent, so every warp-wide
// 128-bit gather covers whole 4*NQ-float rows of ~32/NQ neighbouring voxels, which
// map to the same or neighbouring rays (contiguous rows): ~5 L1 wavefronts per
// gather instead of ~16 for three strided gathers (profiles/r1_summary.md: the
// adjoint is bound by the L1 wavefront rate, 66 % of it these row gathers and
// 27 % the 4-byte walks of the per-voxel symmetry lists).  The symmetry part
// reads the fixed-width list copy (BD::ell) with 128-bit loads laid out like v.
// All offsets are 32-bit (view rows of a candidate are contiguous in u).
// Addition order: views, then k, then symmetry rows -- as k_adj.
// ---------------------------------------------------------------------------
#define HB2_ADJ_VU 4
template <int KT>
__global__ void __launch_bounds__(HB2_BLOCK) k_adj_pq(BD B, int mode) {
  const int c = blockIdx.y;
  __shared__ float red[HB2_BLOCK / 32];
  __shared__ unsigned s_aoff[HB2_ADJ_VIEWS];
  __shared__ float s_w[HB2_ADJ_VIEWS];
  const LsmrState& S = B.st[c];
  bool act = (mode == MODE_LSMR) ? (S.active != 0 && !S.skip_adj)
                                 : (mode == MODE_INIT ? (S.beta > 0.f) : (B.only_cand < 0 || B.only_cand == c));
  const int pi = c * B.part_v_per_cand + blockIdx.x;
  if (!act) {
    if (threadIdx.x == 0 && mode != MODE_PLAIN) B.part_v[pi] = 0.f;
    return;
  }
  const int L3P = B.L3P, ndisk = B.ndisk;
  const int NQ = L3P >> 2;
  const unsigned rpv = (unsigned)B.rows_per_view, kstride = (unsigned)KT * (unsigned)B.apitch;
  const int t = blockIdx.x * HB2_BLOCK + threadIdx.x;
  const int p = t / NQ, q = t - p * NQ, z0 = 4 * q;
  const bool live = p < ndisk;
  const float ib = mode == MODE_PLAIN ? 1.f : S.inv_beta;
  // Halton duplicates carry identical rows inside the solver (u starts as b and is updated row-wise), so the
  // first copy can stand for them with its multiplicity; an arbitrary row vector (MODE_PLAIN, the API's
  // apply_adjoint) must honour every row on its own.
  const bool dedupe_adj = mode != MODE_PLAIN;
  const float beta = S.beta;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  const int vb = B.cand_view_begin[c], nv = B.cand_view_count[c];
  const float* __restrict__ ucand = B.u + B.cand_uoff[c] + z0;
  const uint16_t* __restrict__ am0 = B.amap + (live ? B.aslot[p] : 0);
  const uint16_t* __restrict__ am1 = am0 + B.apitch;
  for (int v0 = 0; v0 < nv; v0 += HB2_ADJ_VIEWS) {
    const int nvc = min(HB2_ADJ_VIEWS, nv - v0);
    __syncthreads();
    for (int e = threadIdx.x; e < HB2_ADJ_VIEWS; e += HB2_BLOCK) {
      const int vw = vb + v0 + min(e, nvc - 1);  // tail entries repeat the last view
      const bool skip = (B.view_tie && B.view_tie[vw] >= 0) || (dedupe_adj && B.view_dupof[vw] >= 0);  // k_adj_tie / weight of the first copy
      s_aoff[e] = skip ? 0xFFFFFFFFu : (unsigned)B.view_angle[vw] * kstride;
      s_w[e] = dedupe_adj ? (float)B.view_mult[vw] : 1.f;
    }
    __syncthreads();
    if (!live) continue;
    for (int vi = 0; vi < nvc; vi += HB2_ADJ_VU) {
      unsigned j0[HB2_ADJ_VU], j1[HB2_ADJ_VU];
#pragma unroll
      for (int w = 0; w < HB2_ADJ_VU; ++w) {
        const unsigned ao = s_aoff[vi + w];
        const bool skip = ao == 0xFFFFFFFFu || vi + w >= nvc;  // tie view (k_adj_tie) or past the end
        j0[w] = am0[skip ? 0u : ao];
        j1[w] = KT == 2 ? (unsigned)am1[skip ? 0u : ao] : 0xFFFFu;
        if (skip) { j0[w] = 0xFFFFu; j1[w] = 0xFFFFu; }
      }
      float4 r0[HB2_ADJ_VU];
#pragma unroll
      for (int w = 0; w < HB2_ADJ_VU; ++w) {
        const unsigned uo = (unsigned)(v0 + vi + w) * rpv;
        r0[w] = j0[w] != 0xFFFFu ? ldg4(ucand + (uo + j0[w] * (unsigned)L3P)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int w = 0; w < HB2_ADJ_VU; ++w) {
        const float ibw = ib * s_w[vi + w];
        if (j0[w] != 0xFFFFu) {
          acc0 = fmaf(r0[w].x, ibw, acc0); acc1 = fmaf(r0[w].y, ibw, acc1);
          acc2 = fmaf(r0[w].z, ibw, acc2); acc3 = fmaf(r0[w].w, ibw, acc3);
        }
        if (KT == 2 && j1[w] != 0xFFFFu) {  // second sample of the same view in this voxel (rare)
          const float4 r1 = ldg4(ucand + ((unsigned)(v0 + vi + w) * rpv + j1[w] * (unsigned)L3P));
          acc0 = fmaf(r1.x, ibw, acc0); acc1 = fmaf(r1.y, ibw, acc1);
          acc2 = fmaf(r1.z, ibw, acc2); acc3 = fmaf(r1.w, ibw, acc3);
        }
      }
    }
  }
  float ss = 0.f;
  if (live) {
    const int g0 = p * L3P + z0;
    const float* __restrict__ us = B.u + B.cand_uoff[c] + B.cand_mdata[c];
    const int* __restrict__ ell = B.ell + (size_t)c * HB2_ELL_W * B.npad + g0;
    float* vdst = (mode == MODE_PLAIN ? B.xs : B.v) + (size_t)c * B.npad + g0;
    if (B.vtie && B.cand_tie_count[c] > 0) {  // rows of the tie views (k_adj_tie)
      const float4 t = *reinterpret_cast<const float4*>(B.vtie + (size_t)c * B.npad + g0);
      acc0 += t.x; acc1 += t.y; acc2 += t.z; acc3 += t.w;
    }
    int ev[HB2_ELL_W][4];
#pragma unroll
    for (int w = 0; w < HB2_ELL_W; ++w) {
      const int4 e4 = __ldg(reinterpret_cast<const int4*>(ell + (size_t)w * B.npad));
      ev[w][0] = e4.x; ev[w][1] = e4.y; ev[w][2] = e4.z; ev[w][3] = e4.w;
    }
    float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mode == MODE_LSMR) old = *reinterpret_cast<const float4*>(vdst);
    float vals[HB2_ELL_W][4];
#pragma unroll
    for (int w = 0; w < HB2_ELL_W; ++w)
#pragma unroll
      for (int tz = 0; tz < 4; ++tz) {
        const bool ok = ev[w][tz] < 0 || ev[w][tz] < HB2_ELL_OVERFLOW;  // a real entry (either sign)
        vals[w][tz] = ok ? __ldg(us + (ev[w][tz] & 0x7fffffff)) : 0.f;
      }
    const float a2[4] = {acc0, acc1, acc2, acc3};
    float vn[4];
#pragma unroll
    for (int tz = 0; tz < 4; ++tz) {
      float s2 = a2[tz];
      if (ev[HB2_ELL_W - 1][tz] == HB2_ELL_OVERFLOW) {  // long list: walk the CSR copy
        const int* __restrict__ ptr = B.csc_ptr + (size_t)c * (B.npad + 1) + g0 + tz;
        const int* __restrict__ ent = B.csc_ent + B.cand_cscoff[c];
        for (int e = ptr[0]; e < ptr[1]; ++e) {
          const int x = ent[e];
          const float val = __ldg(us + (x & 0x7fffffff));
          s2 = fmaf(x < 0 ? -val : val, ib, s2);
        }
      } else {
#pragma unroll
        for (int w = 0; w < HB2_ELL_W; ++w)
          if (ev[w][tz] != HB2_ELL_NONE) s2 = fmaf(ev[w][tz] < 0 ? -vals[w][tz] : vals[w][tz], ib, s2);
      }
      const float o = tz == 0 ? old.x : (tz == 1 ? old.y : (tz == 2 ? old.z : old.w));
      vn[tz] = mode == MODE_LSMR ? fadd_(fmul_(o, -beta), s2) : s2;
      ss += vn[tz] * vn[tz];
    }
    *reinterpret_cast<float4*>(vdst) = make_float4(vn[0], vn[1], vn[2], vn[3]);
  }
  if (mode != MODE_PLAIN) {
    float tot = block_sum(ss, red);
    if (threadIdx.x == 0) B.part_v[pi] = tot;
  }
}

// rays of view angle a crossing voxel tile t: [jlo, jlo + nr).  One CTA per (tile, angle).
__global__ void __launch_bounds__(HB2_BLOCK) k_tile_rays(int K, int apitch, int ntile, const uint16_t* __restrict__ amap,
                                                        uint16_t* __restrict__ jlo, uint16_t* __restrict__ nr,
                                                        int* __restrict__ rmax) {
  const int t = blockIdx.x, a = blockIdx.y;
  __shared__ int s_lo[HB2_BLOCK / 32], s_hi[HB2_BLOCK / 32];
  int lo = 0x7fffffff, hi = -1;
  for (int k = 0; k < K; ++k) {
    const unsigned j = amap[((size_t)a * K + k) * apitch + (size_t)t * HB2_BLOCK + threadIdx.x];
    if (j != 0xFFFFu) { lo = min(lo, (int)j); hi = max(hi, (int)j); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < HB2_BLOCK / 32; ++w) { lo = min(lo, s_lo[w]); hi = max(hi, s_hi[w]); }
    const int n = hi >= 0 ? hi - lo + 1 : 0;
    jlo[(size_t)a * ntile + t] = (uint16_t)(n ? lo : 0);
    nr[(size_t)a * ntile + t] = (uint16_t)n;
    if (n) atomicMax(rmax, n);
  }
}

// ===========================================================================
// vector update (lsmr.py:364-368) fused with v normalisation and ||x||^2
// ===========================================================================
__global__ void __launch_bounds__(HB2_BLOCK) k_update(BD B, int mode) {
  // npad is a multiple of 4: every thread moves 4 consecutive elements with 128-bit accesses
  // (v, h: one float4 each; x, hbar: two double2 each) -- the kernel is a pure HBM stream (44 n bytes);
  // profiles/r1_summary.md: 4.5 -> 6.6 TB/s against scalar accesses.
  const int c = blockIdx.y;
  __shared__ double redd[HB2_BLOCK / 32];
  const LsmrState& S = B.st[c];
  const int pi = c * B.part_x_per_cand + blockIdx.x;
  if (!S.active) {
    if (threadIdx.x == 0) B.part_x[pi] = 0.0;
    return;
  }
  const float ia = S.inv_alpha;
  const double cfhb = (double)S.cf_hbar, cfx = (double)S.cf_x;
  const float cfh = S.cf_h;
  const bool normalise = !(mode == MODE_LSMR && S.skip_adj);
  float4* v4 = reinterpret_cast<float4*>(B.v + (size_t)c * B.npad);
  float4* h4 = reinterpret_cast<float4*>(B.h + (size_t)c * B.npad);
  double2* x2 = reinterpret_cast<double2*>(B.x + (size_t)c * B.npad);
  double2* hb2 = reinterpret_cast<double2*>(B.hbar + (size_t)c * B.npad);
  const int n4 = B.npad >> 2;
  double sx = 0.0;
  const int i = blockIdx.x * HB2_BLOCK + threadIdx.x;
  if (i < n4) {
    float4 vq = v4[i];
    float vn[4] = {vq.x, vq.y, vq.z, vq.w};
    if (normalise) {
#pragma unroll
      for (int k = 0; k < 4; ++k) vn[k] = fmul_(vn[k], ia);
      v4[i] = make_float4(vn[0], vn[1], vn[2], vn[3]);
    }
    if (mode == MODE_INIT) {
      h4[i] = make_float4(vn[0], vn[1], vn[2], vn[3]);
      hb2[2 * i] = make_double2(0.0, 0.0); hb2[2 * i + 1] = make_double2(0.0, 0.0);
      x2[2 * i] = make_double2(0.0, 0.0); x2[2 * i + 1] = make_double2(0.0, 0.0);
    } else {
      const float4 hq = h4[i];
      const float ho[4] = {hq.x, hq.y, hq.z, hq.w};
      const double2 b0 = hb2[2 * i], b1 = hb2[2 * i + 1];
      const double2 x0 = x2[2 * i], x1 = x2[2 * i + 1];
      double hb[4] = {b0.x, b0.y, b1.x, b1.y};
      double xn[4] = {x0.x, x0.y, x1.x, x1.y};
      float hn[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        hb[k] = dadd_(dmul_(hb[k], cfhb), (double)ho[k]);
        xn[k] = dadd_(xn[k], dmul_(cfx, hb[k]));
        hn[k] = fadd_(fmul_(ho[k], cfh), vn[k]);
        sx += xn[k] * xn[k];
      }
      hb2[2 * i] = make_double2(hb[0], hb[1]); hb2[2 * i + 1] = make_double2(hb[2], hb[3]);
      x2[2 * i] = make_double2(xn[0], xn[1]); x2[2 * i + 1] = make_double2(xn[2], xn[3]);
      h4[i] = make_float4(hn[0], hn[1], hn[2], hn[3]);
    }
  }
  double tot = block_sum_d(sx, redd);
  if (threadIdx.x == 0) B.part_x[pi] = tot;
}

// x (float64) -> xs (float32), SLR:270
__global__ void k_x_to_f32(BD B) {
  const int c = blockIdx.y;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B.npad) B.xs[(size_t)c * B.npad + i] = (float)B.x[(size_t)c * B.npad + i];
}

// ===========================================================================
// scalar kernels: one CTA per candidate reduces the per-CTA partials in a
// fixed order (deterministic) and advances the LSMR recurrences.
// ===========================================================================
__device__ __forceinline__ double reduce_partials_f(const float* p, int n, double* shd) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += HB2_BLOCK) s += (double)p[i];
  return block_sum_d(s, shd);
}

// ---------------------------------------------------------------------------
// numpy.linalg.norm of a float32 vector AS THE REFERENCE EXECUTES IT.  scipy's LSMR calls numpy.linalg.norm(u) /
// norm(v) on float32 vectors (lsmr.py:239-340); numpy computes sqrt(dot(x, x)) with OpenBLAS sdot, whose x86 AVX-512
// kernel (OpenBLAS 0.3.30 SkylakeX, fitted bit for bit against numpy in this container, oracle/blas_sdot.py) keeps
// 64 float32 accumulators: accumulator l adds elements l, l + 64, l + 128, ... one after the other with a fused
// multiply-add, then folds 64 -> 32 -> 8 -> 4 -> 1.  Sequentially accumulating ~n/64 positive terms in float32 loses
// the low bits of every late term: the result is BIASED low (measured: -7.6e-6 relative at 6 M elements, -2.2e-5 at
// 12 M), and LSMR amplifies a 1e-5 error of alpha/beta to ~6e-3 in x within ten iterations.  An exactly rounded norm
// is therefore NOT what the reference computes at the BASELINE sizes.  What matters is the chain structure (64
// accumulators, sequential float32 FMA, chain length n/64), not which element sits in which accumulator (measured
// with the oracle: the same chains over another element order move x by 1e-5...3e-5, an exact norm by 5.8e-3).
// One CTA of 64 threads per candidate: thread l owns accumulator l and consumes 4 consecutive floats per step; zero
// padding adds nothing.  All candidates of the batch run concurrently, so the sequential chain costs latency, not
// throughput; the kernel is a pure HBM stream (TMA bulk copies into a shared-memory ring).
// ---------------------------------------------------------------------------
#define HB2_CHAIN_STAGES 6
#define HB2_CHAIN_STAGE_F4 1024   // float4 per stage = 16 KB = 16 steps of the 64 accumulators
#define HB2_CHAIN_SMEM (HB2_CHAIN_STAGES * HB2_CHAIN_STAGE_F4 * 16)
enum { CHAIN_U = 0, CHAIN_V = 1, CHAIN_B = 2 };
template <int WHICH>
__global__ void __launch_bounds__(64) k_chain_sumsq(BD B, float* __restrict__ out, int mode) {
  const int c = blockIdx.x;
  if (mode == MODE_LSMR && !B.st[c].active) return;
  const float* p;
  long long n;
  if (WHICH == CHAIN_U) { p = B.u + B.cand_uoff[c]; n = (long long)B.cand_mdata[c] + ((B.cand_msym[c] + 3) & ~3); }
  else if (WHICH == CHAIN_V) { p = B.v + (size_t)c * B.npad; n = B.npad; }
  else { p = B.b + B.cand_uoff[c]; n = B.cand_mdata[c]; }
  const float4* __restrict__ p4 = reinterpret_cast<const float4*>(p);
  const long long n4 = n >> 2;  // every vector here is a multiple of 4 floats and 16-byte aligned
  // The vector streams through a ring of shared-memory stages filled by TMA bulk copies (one elected thread issues
  // them, mbarrier transaction counts signal arrival): 5 x 16 KB in flight per CTA keep HBM busy although only two
  // warps per candidate consume (the chains are sequential by definition).
  extern __shared__ __align__(128) unsigned char chain_smem[];  // HB2_CHAIN_STAGES x 16 KB (dynamic: > 48 KB)
  float4 (*ring)[HB2_CHAIN_STAGE_F4] = reinterpret_cast<float4 (*)[HB2_CHAIN_STAGE_F4]>(chain_smem);
  __shared__ unsigned long long full_bar[HB2_CHAIN_STAGES];
  __shared__ float a[64];
  const int lane = threadIdx.x;
  const long long nstage = (n4 + HB2_CHAIN_STAGE_F4 - 1) / HB2_CHAIN_STAGE_F4;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < HB2_CHAIN_STAGES; ++i) mbar_init(&full_bar[i], 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncthreads();
  auto issue = [&](long long st) {
    const int buf = (int)(st % HB2_CHAIN_STAGES);
    const long long lo = st * HB2_CHAIN_STAGE_F4;
    const unsigned bytes = (unsigned)(min((long long)HB2_CHAIN_STAGE_F4, n4 - lo) * 16);
    mbar_expect_tx(&full_bar[buf], bytes);
    bulk_g2s(&ring[buf][0], p4 + lo, bytes, &full_bar[buf]);
  };
  if (lane == 0)
    for (long long st = 0; st < min((long long)HB2_CHAIN_STAGES - 1, nstage); ++st) issue(st);
  float acc = 0.f;
  for (long long st = 0; st < nstage; ++st) {
    const int buf = (int)(st % HB2_CHAIN_STAGES);
    // refill the stage consumed in the previous round (every thread passed the barrier at its end)
    if (lane == 0 && st + HB2_CHAIN_STAGES - 1 < nstage) issue(st + HB2_CHAIN_STAGES - 1);
    mbar_wait(&full_bar[buf], (unsigned)((st / HB2_CHAIN_STAGES) & 1));
    const int cnt = (int)min((long long)HB2_CHAIN_STAGE_F4, n4 - st * HB2_CHAIN_STAGE_F4);
    const float4* __restrict__ rb = ring[buf];
    // All 16 float4 of this lane are requested before the first FMA: the chain (4 cycles per FMA, sequential by
    // definition) starts when the first load lands and never waits for shared memory again inside the stage.  Slots
    // past the end of the vector read as zero: fma(0, 0, acc) leaves acc unchanged.
    static_assert(HB2_CHAIN_STAGE_F4 == 64 * 16, "16 float4 per lane and stage");
    float4 t[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int i = lane + 64 * k;
      t[k] = i < cnt ? rb[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      acc = __fmaf_rn(t[k].x, t[k].x, acc); acc = __fmaf_rn(t[k].y, t[k].y, acc);
      acc = __fmaf_rn(t[k].z, t[k].z, acc); acc = __fmaf_rn(t[k].w, t[k].w, acc);
    }
    __syncthreads();  // the stage may be overwritten from the next round on
  }
  // fold like the sdot kernel: 4 vectors of 16 lanes -> 4 x 8 -> 8 -> 4 -> 1
  a[lane] = acc;
  __syncthreads();
  if (lane == 0) {
    float a8[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      float q[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) q[k] = fadd_(a[16 * k + l], a[16 * k + l + 8]);
      a8[l] = fadd_(fadd_(fadd_(q[0], q[1]), q[2]), q[3]);
    }
    float h[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) h[l] = fadd_(a8[l], a8[l + 4]);
    out[c] = fadd_(fadd_(h[0], h[1]), fadd_(h[2], h[3]));
  }
}

// phase 0 (init): normb = ||b||
__global__ void __launch_bounds__(HB2_BLOCK) k_scal_normb(BD B, const float* __restrict__ chain) {
  const int c = blockIdx.x;
  __shared__ double shd[HB2_BLOCK / 32];
  const float* bb = B.b + B.cand_uoff[c];
  double s = 0.0;
  if (!chain) {
    for (int i = threadIdx.x; i < B.cand_mdata[c]; i += HB2_BLOCK) { double t = bb[i]; s += t * t; }
    s = block_sum_d(s, shd);
  }
  if (threadIdx.x == 0) {
    LsmrState& S = B.st[c];
    float beta = __fsqrt_rn(chain ? chain[c] : (float)s);
    S.beta = beta; S.normb = beta;
    S.inv_beta = beta > 0.f ? fdiv_(1.f, beta) : 0.f;
    S.alpha = 0.f; S.active = 1; S.skip_adj = 0; S.istop = 0; S.itn = 0;
  }
}
// phase 0b: alpha = ||A^T u||, initialise recurrences
__global__ void __launch_bounds__(HB2_BLOCK) k_scal_init(BD B, int* nactive, const float* __restrict__ chain) {
  const int c = blockIdx.x;
  __shared__ double shd[HB2_BLOCK / 32];
  double s = chain ? 0.0 : reduce_partials_f(B.part_v + (size_t)c * B.part_v_per_cand, B.part_v_per_cand, shd);
  if (threadIdx.x == 0) {
    LsmrState& S = B.st[c];
    float alpha = __fsqrt_rn(chain ? chain[c] : (float)s);
    lsmr_init_(S, alpha, S.beta);
    if (S.active) atomicAdd(nactive, 1);
  }
}
// beta = ||u~||
__global__ void __launch_bounds__(HB2_BLOCK) k_scal_beta(BD B, const float* __restrict__ chain) {
  const int c = blockIdx.x;
  __shared__ double shd[HB2_BLOCK / 32];
  LsmrState& S = B.st[c];
  if (!S.active) return;
  const int ntiles = B.fwd_ppv;
  double s = 0.0;
  if (!chain) {
    s = reduce_partials_f(B.part_u + (size_t)B.cand_view_begin[c] * ntiles, B.cand_view_count[c] * ntiles, shd);
    s += reduce_partials_f(B.part_us + (size_t)c * B.part_us_per_cand, B.part_us_per_cand, shd);
  }
  if (threadIdx.x == 0) {
    float beta = __fsqrt_rn(chain ? chain[c] : (float)s);
    S.beta = beta;
    if (beta > 0.f) { S.inv_beta = fdiv_(1.f, beta); S.skip_adj = 0; }
    else { S.skip_adj = 1; }
  }
}
// alpha = ||v~||, rotations, update coefficients
__global__ void __launch_bounds__(HB2_BLOCK) k_scal_rot(BD B, const float* __restrict__ chain) {
  const int c = blockIdx.x;
  __shared__ double shd[HB2_BLOCK / 32];
  LsmrState& S = B.st[c];
  if (!S.active) return;
  double s = chain ? 0.0 : reduce_partials_f(B.part_v + (size_t)c * B.part_v_per_cand, B.part_v_per_cand, shd);
  if (threadIdx.x == 0) {
    float alpha = S.skip_adj ? S.alpha : __fsqrt_rn(chain ? chain[c] : (float)s);
    lsmr_rotate_(S, alpha, S.beta);
  }
}
// ||x||, stopping tests
__global__ void __launch_bounds__(HB2_BLOCK) k_scal_test(BD B, double atol, double btol, double conlim, int maxiter,
                                                         int fixed_iters, int* nactive) {
  const int c = blockIdx.x;
  __shared__ double shd[HB2_BLOCK / 32];
  LsmrState& S = B.st[c];
  if (!S.active) return;
  double s = 0.0;
  for (int i = threadIdx.x; i < B.part_x_per_cand; i += HB2_BLOCK) s += B.part_x[(size_t)c * B.part_x_per_cand + i];
  s = block_sum_d(s, shd);
  if (threadIdx.x == 0) {
    int istop = lsmr_test_(S, sqrt(s), atol, btol, conlim, maxiter);
    if (fixed_iters > 0) { istop = S.itn >= fixed_iters ? 7 : 0; S.istop = istop; }
    if (istop > 0) { S.active = 0; atomicSub(nactive, 1); }
  }
}
// cosine score from the projector's partials (lib/analysis.py:802-821)
__global__ void __launch_bounds__(HB2_BLOCK) k_scal_score(BD B, float* __restrict__ score) {
  const int c = blockIdx.x;
  __shared__ double shd[HB2_BLOCK / 32];
  const int ntiles = B.fwd_ppv;
  const int n = B.cand_view_count[c] * ntiles;
  const float* p = B.part_s + (size_t)B.cand_view_begin[c] * ntiles * 3;
  double d = 0, pp = 0, bb = 0;
  for (int i = threadIdx.x; i < n; i += HB2_BLOCK) { d += p[3 * i]; pp += p[3 * i + 1]; bb += p[3 * i + 2]; }
  d = block_sum_d(d, shd); pp = block_sum_d(pp, shd); bb = block_sum_d(bb, shd);
  if (threadIdx.x == 0) {
    float na = __fsqrt_rn((float)pp), nb = __fsqrt_rn((float)bb);
    float norm = fmul_(na, nb);
    score[c] = norm == 0.f ? 0.f : fdiv_((float)d, norm);
  }
}
