// Fused backward kernel: one pass over the arrival points produces grad_u, grad_v AND grad_field.
//
// Work unit ("task") = one warp = (plane, band of destination rows [ra, rb), strip of destination
// columns [ja, jb)).  The warp sweeps the arrival rows that can reach its tile in ascending order;
// per row it visits its own columns ("core" chunks: full work incl. grad_u / grad_v) and `hx` halo
// columns either side (trajectory + adjoint only).  Adjoint contributions go into a warp-private
// ring of `ring` destination rows in shared memory; a destination row is written to grad_field
// (coalesced) as soon as no later arrival row can reach it.
//
// Determinism without atomics: only this warp ever touches its ring, rows/chunks/tap phases are
// visited in a fixed order, and within one 32-lane step two lanes can only clash when their
// departure cells coincide.  Clashes are detected with one-byte tags (plain stores: who wins only
// signals that a clash exists), then every clashing cell is resolved by a ballot on the exact key
// and summed into its lowest lane in ascending lane order, so each add into the ring has a single
// writer and a data-defined order.
//
// Contract: |floor(iy) - row| <= rr and the longitudinal reach of a band <= its hx.  Every core
// point checks it; a violation marks the plane in `plane_flag`, and the host always enqueues the
// general (two-kernel) path behind the sweep, which recomputes exactly the flagged planes.
#pragma once
#include "sl_device.cuh"

namespace psl {

constexpr int kSweepWarps = 4;
constexpr int kMaxBands = 48;

constexpr int kTagRows = 4;                   // clash tags cover (slot mod kTagRows, column)

// Longitudinal reach (in cells, rounded up to 16) of an arrival row: the halo a strip needs on
// either side.  |dlon| <= asin(sin(delta) / cos(|lat| + delta)) for a great-circle step delta.
// Evaluated identically by every task (and by the host plan, which only uses it for balancing).
struct ReachModel { float sin_delta, cos_delta, inv_dlam; int extra, max_halo; };
__host__ __device__ __forceinline__ int halo_cells(const ReachModel& m, float sin_lat, float cos_lat) {
  const float c = cos_lat * m.cos_delta - fabsf(sin_lat) * m.sin_delta;   // cos(|lat| + delta)
  if (!(c > m.sin_delta)) return 1 << 20;
  const int n = (int)ceilf(asinf(m.sin_delta / c) * m.inv_dlam) + m.extra;
  return (n + 15) & ~15;
}

struct SweepPlan {
  int nbands, nstrips, wc, rr, ring, pitch;   // pitch = wc + 2 * (NT - 1)
  int planes;
  ReachModel reach;
  short ra[kMaxBands], rb[kMaxBands];         // band k owns destination rows [ra[k], rb[k]); bands are
                                              // sorted by decreasing cost
  unsigned char* plane_flag;                  // [planes] set to 1 on a contract violation
  int out0, outN;                             // rows held by the output tensors (global first row, count)
};

template <int NT>
__device__ __forceinline__ void resolve_clashes(int key, int tkey, unsigned char* tag, int lane,
                                                float (&c)[NT * NT], bool& writer) {
  // tkey: tag slot (hashed rows; a false alias only costs one pass of the loop below)
  if (key >= 0) tag[tkey] = (unsigned char)lane;
  __syncwarp();
  const bool lost = (key >= 0) && (tag[tkey] != (unsigned char)lane);
  unsigned pending = __ballot_sync(0xffffffffu, lost);
  writer = key >= 0;
  while (pending) {  // uniform loop: one iteration per clashing cell
    const int j = __ffs(pending) - 1;
    const int kj = __shfl_sync(0xffffffffu, key, j);
    const unsigned group = __ballot_sync(0xffffffffu, key == kj);
    pending &= ~group;
    const int leader = __ffs(group) - 1;
    unsigned rest = group & (group - 1);
    while (rest) {
      const int src = __ffs(rest) - 1;
      rest &= rest - 1;
#pragma unroll
      for (int t = 0; t < NT * NT; ++t) {
        const float o = __shfl_sync(0xffffffffu, c[t], src);
        if (lane == leader) c[t] += o;
      }
    }
    if (lane != leader && ((group >> lane) & 1u)) writer = false;
  }
}

template <bool EXACT, int INTERP>
__global__ void __launch_bounds__(kSweepWarps * 32) sl_bwd_sweep_kernel(const Params P, const SweepPlan S) {
  constexpr int NT = Stencil<INTERP>::NT, OMIN = Stencil<INTERP>::OMIN;
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1-D grid, most expensive bands first: task -> (band, plane, strip)
  const int task = blockIdx.x * kSweepWarps + warp;
  const int per_band = S.planes * S.nstrips;
  if (task >= S.nbands * per_band) return;
  const int band = task / per_band, rem = task - band * per_band;
  const int pl = rem / S.nstrips, strip = rem - pl * S.nstrips;
  const int b = pl / P.V, c = pl - b * P.V;
  const int ra = S.ra[band], rb = S.rb[band];
  const int ja = strip * S.wc, jb = min(ja + S.wc, P.W), wc = jb - ja;
  const int ring = S.ring, pitch = S.pitch, rr = S.rr;
  const int cells = ring * pitch;
  float* acc = smem + (size_t)warp * (cells + (kTagRows * pitch + 3) / 4);
  unsigned char* tag = reinterpret_cast<unsigned char*>(acc + cells);
  for (int i = lane; i < cells; i += 32) acc[i] = 0.0f;
  __syncwarp();

  const float* up = plane_ptr_bc(P.u, P.u_sB, b, c, P.arrN, P.W);
  const float* vp = plane_ptr_bc(P.v, P.v_sB, b, c, P.arrN, P.W);
  const float* gp = plane_ptr_bc(P.gout, P.gout_sB, b, c, P.arrN, P.W);
  const float* f = plane_ptr_bc(P.field, P.field_sB, b, c, P.fldN, P.W);
  float mean0 = 0.0f, mean1 = 0.0f;
  if (P.pole_fix) { mean0 = __ldg(P.fmean + 2 * pl); mean1 = __ldg(P.fmean + 2 * pl + 1); }
  const bool want_uv = P.gu != nullptr;
  const int ncore = (wc + 31) >> 5;
  bool violated = false;

  // destination row i_done(y) = y - rr + OMIN is complete after arrival row y; it sits in slot `head`
  const int y_first = ra - (ring - 1) + rr - OMIN, y_last = rb - 1 + rr - OMIN;
  int head = 0;
  for (int y = y_first; y <= y_last; ++y) {
    if (y >= P.arr0 && y < P.arr0 + P.arrN) {
      const float sp = __ldg(P.sin_lat + y), cp = __ldg(P.cos_lat + y);
      const int rowoff = (y - P.arr0) * P.W;
      const bool core_row = (y >= ra) && (y < rb);
      int hx = halo_cells(S.reach, sp, cp);           // <= max_halo for every row a band may visit
      hx = min(hx, S.reach.max_halo + 16);            // (host/device rounding may differ by one step)
      const int nhalo = (2 * hx + 31) >> 5;
      for (int ch = 0; ch < ncore + nhalo; ++ch) {
        // column of this lane: core chunks cover [ja, jb), halo chunks cover [ja-hx, ja) then [jb, jb+hx)
        int x;
        bool live, core = false, check = false;
        if (ch < ncore) {
          x = ja + (ch << 5) + lane;
          live = x < jb;
          core = live && core_row;
          check = live;          // every arrival point this task sees in its own columns is checked
        } else {
          const int h = ((ch - ncore) << 5) + lane;
          live = h < 2 * hx;
          x = h < hx ? ja - hx + h : jb + (h - hx);
        }
        int xw = x;
        if (xw < 0) xw += P.W; else if (xw >= P.W) xw -= P.W;
        live = live && ((unsigned)xw < (unsigned)P.W);
        float cc[NT * NT];
#pragma unroll
        for (int t = 0; t < NT * NT; ++t) cc[t] = 0.0f;
        int key = -1, slot0 = 0, cidx = 0;
        if (live) {
          const float uu = __ldg(up + rowoff + xw), vv = __ldg(vp + rowoff + xw);
          const float g = __ldg(gp + rowoff + xw);
          Traj t;
          trajectory<EXACT>(P, uu, vv, sp, cp, __ldg(P.lon + xw), t);
          const float fx = floorf(t.ix), fy = floorf(t.iy);
          const float tx = __fsub_rn(t.ix, fx), ty = __fsub_rn(t.iy, fy);
          const int x0 = (int)fx + OMIN;                 // padded column of tap 0
          const int cls = (int)fy - (y + P.p);           // row class
          int dx = x0 - P.p - xw;                        // longitudinal cell displacement of tap 0
          if (dx < -P.halfW) dx += P.W; else if (dx >= P.halfW) dx -= P.W;
          float wx[NT], wy[NT], dwx[NT], dwy[NT];
          axis_weights<INTERP, true>(tx, wx, dwx);
          axis_weights<INTERP, true>(ty, wy, dwy);
          const bool in_ring = (unsigned)(cls + rr) <= (unsigned)(2 * rr);
          if (check && (!in_ring || dx < -hx || dx > hx - NT + 1)) violated = true;
          if (core) {
            if (want_uv) {
              float val, ddx, ddy;
              stencil_eval<INTERP, true>(P, f, t, mean0, mean1, val, ddx, ddy);
              float ou, ov;
              velocity_grads(P, t, sp, cp, g * ddx, g * ddy, ou, ov);
              const long long o = ((long long)pl * S.outN + (y - S.out0)) * P.W + xw;
              __stcs(P.gu + o, ou);
              __stcs(P.gv + o, ov);
            }
          }
          cidx = (x - ja) + dx + (NT - 1);               // ring column of tap 0
          if (in_ring && (unsigned)cidx <= (unsigned)(wc + NT - 2)) {
            slot0 = head + cls + rr;
            if (slot0 >= ring) slot0 -= ring;
            key = slot0 * pitch + cidx;
#pragma unroll
            for (int a = 0; a < NT; ++a)
#pragma unroll
              for (int bb = 0; bb < NT; ++bb)
                cc[a * NT + bb] = g * __fmul_rn(wy[a], wx[bb]);
            // padding_mode="zeros": only the last tap column can fall off the padded plane (x0 >= 0)
            if (x0 + NT - 1 >= P.Wp) {
#pragma unroll
              for (int a = 0; a < NT; ++a)
#pragma unroll
                for (int bb = 0; bb < NT; ++bb)
                  if (x0 + bb >= P.Wp) cc[a * NT + bb] = 0.0f;
            }
          }
        }
        bool writer;
        resolve_clashes<NT>(key, (slot0 & (kTagRows - 1)) * pitch + cidx, tag, lane, cc, writer);
#pragma unroll
        for (int a = 0; a < NT; ++a) {
          int sl = slot0 + a;
          if (sl >= ring) sl -= ring;
#pragma unroll
          for (int bb = 0; bb < NT; ++bb) {
            if (writer) acc[sl * pitch + cidx + bb] += cc[a * NT + bb];
            __syncwarp();
          }
        }
      }
    }
    // retire destination row i = y - rr + OMIN
    const int i = y - rr + OMIN;
    float* row = acc + head * pitch;
    if (i >= ra && i < rb) {
      float* orow = P.gfield + ((long long)pl * S.outN + (i - S.out0)) * P.W + ja;
      for (int k = lane; k < wc; k += 32) orow[k] = row[k + NT - 1];
    }
    __syncwarp();
    for (int k = lane; k < pitch; k += 32) row[k] = 0.0f;
    __syncwarp();
    head = head + 1 == ring ? 0 : head + 1;
  }
  if (__any_sync(0xffffffffu, violated) && lane == 0) S.plane_flag[pl] = 1;
}

}  // namespace psl
