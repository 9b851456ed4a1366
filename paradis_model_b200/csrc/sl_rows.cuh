// Fused backward, round-2 design: one persistent, warp-specialised CTA per SM sweeps whole latitude
// circles.  One pass over the arrival points produces grad_u, grad_v AND grad_field, at every latitude
// (mid-latitudes and polar caps alike), deterministic and without atomics.
//
// Roles inside a CTA (one CTA per SM, launched once, walks a contiguous range of (plane, row) work):
//
//   loader     1 warp, one elected lane: streams the u, v, grad_out rows of the next arrival rows into
//              shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx), kRowStages
//              rows ahead of the producers.
//   producers  nP warps: per arrival point the departure trajectory, the stencil gather of `field`
//              with its derivative (grad_u, grad_v -> global, coalesced) and a 16-byte RECORD
//              {frac x, frac y, grad_out, (row class | departure column)} into a shared-memory row of
//              records.  Embarrassingly parallel: 4 x 32 consecutive columns per warp step, no scatter.
//   consumers  nC warps: each OWNS a strip of destination columns of the CTA-wide ring of `ring`
//              destination rows (full longitude circle, shared memory).  It scans the records of the
//              arrival columns that can reach its strip ([ja - hx, jb + hx), hx = longitudinal reach of
//              the row, the whole circle near the poles), and adds the stencil weights of the taps that
//              fall into ITS columns.  A destination row leaves the ring (coalesced float4 store to
//              grad_field) once no later arrival row can reach it.
//
// No trajectory is ever recomputed for a halo: a point is computed once per CTA and its record is read by
// the (one or two) consumers whose strips it touches.  The GeoCyclic cap rows fold back onto rows 1..p /
// H-1-p..H-2 (180 degrees shifted) when they are scattered, so the same kernel serves the polar caps.
//
// Determinism without atomics: a ring cell is only ever written by the consumer that owns its column;
// that consumer visits arrival rows, scan steps and tap phases in a fixed order; lanes of one step that
// hit the same departure cell are detected with one-byte tags and summed into the lowest lane in lane
// order (resolve_clashes), so every add has a single writer and a data-defined order.
//
// Pipelines (all mbarrier based, no __syncthreads after start-up):
//   stage_full / stage_free [kRowStages]   loader -> producers   (TMA transaction count / nP arrivals)
//   rec_full   / rec_free   [kRowRecs]     producers -> consumers (nP arrivals / nC arrivals)
//
// Contract: |floor(iy) - row| <= rr and |departure column - arrival column| <= hx(row).  Every point is
// checked by its producer; a violation marks the plane in `plane_flag` and the host always enqueues the
// general (two-kernel) path behind this kernel, which recomputes exactly the flagged planes.
#pragma once
#include "sl_device.cuh"

namespace psl {

constexpr int kRowStages = 2;      // u, v, grad_out rows in flight (TMA)
constexpr int kRowRecs = 3;        // record rows in flight between producers and consumers
constexpr int kStepSub = 4;        // 32-column sub-blocks per producer step
constexpr int kTagBytes = 1024;    // clash tags per consumer warp (hashed, power of two)
#ifndef PSL_ROWS_WARPS
#define PSL_ROWS_WARPS 24          // loader + producers + consumers; 24 warps = 768 threads -> 80 registers
#endif
constexpr int kRowsWarps = PSL_ROWS_WARPS;

struct RowsPlan {
  int planes, rr, ring;            // ring = 2 rr + NT destination rows
  int nP, nC;                      // producer / consumer warps (warp 0 is the loader)
  int wc;                          // consumer strip width (multiple of 4)
  int nsteps;                      // producer steps per arrival row = ceil(W / 128)
  int total_rows;                  // planes * ownN
  int min_seg;                     // CTA boundaries closer than this to a plane boundary snap onto it
  const int* __restrict__ hx_tab;  // [H] longitudinal reach (cells) of an arrival row; >= W: whole circle
  unsigned char* plane_flag;       // [planes] set to 1 on a contract violation
  int out0, outN;                  // rows held by the output tensors
  unsigned off_stage, off_rec, off_tag, off_bar;   // byte offsets in dynamic shared memory (ring at 0)
};

// ---- mbarrier / bulk-copy wrappers (PTX; SASS: SYNCS.*, UBLKCP) ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(b)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- work partition: CTA k of n owns the rows [rows_bound(k), rows_bound(k + 1)) of the concatenated
// (plane, own row) space; boundaries close to a plane boundary snap onto it (a segment costs ring - 1
// extra arrival rows of trajectory work).
__device__ __forceinline__ int rows_bound(const RowsPlan& S, int ownN, int k, int n) {
  const long long t = (long long)k * S.total_rows / n;
  int pl = (int)(t / ownN), r = (int)(t - (long long)pl * ownN);
  if (r < S.min_seg) r = 0;
  else if (ownN - r < S.min_seg) { r = 0; ++pl; }
  return pl * ownN + r;
}

struct RowsSeg { int pl, ra, rb, y_first, y_last; };   // destination rows [ra, rb) of plane pl (global rows)

template <int INTERP>
__device__ __forceinline__ bool rows_next_seg(const Params& P, const RowsPlan& S, int& g, int g1, RowsSeg& s) {
  constexpr int OMIN = Stencil<INTERP>::OMIN;
  if (g >= g1) return false;
  s.pl = g / P.ownN;
  const int r = g - s.pl * P.ownN;
  const int n = min(P.ownN - r, g1 - g);
  s.ra = P.own0 + r; s.rb = s.ra + n;
  // destination row i = y - rr + OMIN is complete after arrival row y
  s.y_first = s.ra - (S.ring - 1) + S.rr - OMIN;
  s.y_last = s.rb - 1 + S.rr - OMIN;
  g += n;
  return true;
}

#ifndef PSL_HAVE_RESOLVE_CLASHES
template <int NT>
__device__ __forceinline__ void resolve_clashes(int key, int tkey, unsigned char* tag, int lane,
                                                float (&c)[NT * NT], bool& writer) {
  // tkey: tag slot (hashed; a false alias only costs one pass of the loop below)
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
#endif

// Record key: bits [31:16] row class + rr (0 .. 2 rr), bit 15: the last tap column lies outside the padded
// plane (padding_mode="zeros"), bits [14:0] unpadded column of tap 0 wrapped into [0, W).  < 0: no contribution.
__device__ __forceinline__ int rows_key(int t, int col0, bool last_out) { return (t << 16) | (last_out ? 0x8000 : 0) | col0; }

// ---- consumer: one 32-record step, rows whose stencils stay inside rows 0 .. H-1 (no cap fold) -----
template <int INTERP>
__device__ __forceinline__ void rows_scatter_plain(const float4 rec, bool inr, float* ring_base, unsigned char* tag,
                                                   int W, int ring, int head, int ja, int wc, int lane) {
  constexpr int NT = Stencil<INTERP>::NT;
  const int key = inr ? __float_as_int(rec.w) : -1;
  const int t = key >> 16, cx = key & 0x7fff;
  int col[NT];
  bool own[NT], any_own = false;
#pragma unroll
  for (int b = 0; b < NT; ++b) {
    int c = cx + b;
    if (c >= W) c -= W;
    col[b] = c;
    own[b] = (unsigned)(c - ja) < (unsigned)wc;
    any_own = any_own || own[b];
  }
  if ((key & 0x8000) != 0) own[NT - 1] = false;
  const bool act = key >= 0 && any_own;
  if (!__any_sync(0xffffffffu, act)) return;
  int s0 = head + t;
  if (s0 >= ring) s0 -= ring;
  float wx[NT], wy[NT], d0[NT], d1[NT], cc[NT * NT];
  axis_weights<INTERP, false>(rec.x, wx, d0);
  axis_weights<INTERP, false>(rec.y, wy, d1);
  const float g = act ? rec.z : 0.0f;
#pragma unroll
  for (int a = 0; a < NT; ++a) {
    const float gw = __fmul_rn(g, wy[a]);
#pragma unroll
    for (int b = 0; b < NT; ++b) cc[a * NT + b] = __fmul_rn(gw, wx[b]);
  }
  bool writer;
  const int k = act ? s0 * W + cx : -1;
  resolve_clashes<NT>(k, ((s0 & 3) * 257 + cx) & (kTagBytes - 1), tag, lane, cc, writer);
#pragma unroll
  for (int a = 0; a < NT; ++a) {
    int sl = s0 + a;
    if (sl >= ring) sl -= ring;
    float* row = ring_base + sl * W;
#pragma unroll
    for (int b = 0; b < NT; ++b) {
      if (writer && own[b]) row[col[b]] += cc[a * NT + b];
      __syncwarp();
    }
  }
}

// ---- consumer: one 32-record step of a row near a pole: tap rows outside 0 .. H-1 fold back through the
// GeoCyclic map (model/padding.py:26-37: reflect about the first / last row excluding it, 180 degrees
// shifted).  Two points with different departure cells can then meet in one tap phase, but only when one
// of the taps is folded and the other is not, so every tap phase is split in two (plain taps, then folded
// taps); within each half distinct departure cells still mean distinct ring cells.
template <int INTERP>
__device__ __noinline__ void rows_scatter_fold(int W, int H, const float4 rec, bool inr, float* ring_base,
                                               unsigned char* tag, int ring, int head, int y, int rr, int ja, int wc,
                                               int lane) {
  constexpr int NT = Stencil<INTERP>::NT, OMIN = Stencil<INTERP>::OMIN;
  constexpr int pad = INTERP;                    // padding width = interpolation id (advection.py:22-24)
  const int halfW = W >> 1;
  const int key = inr ? __float_as_int(rec.w) : -1;
  const int t = key >> 16, cx = key & 0x7fff;
  const bool last_out = (key & 0x8000) != 0;
  const int i0 = y + (t - rr) + OMIN;             // unpadded row of tap row 0
  const int i_ret = y - rr + OMIN;                // row in slot `head`
  float wx[NT], wy[NT], d0[NT], d1[NT], cc[NT * NT];
  axis_weights<INTERP, false>(rec.x, wx, d0);
  axis_weights<INTERP, false>(rec.y, wy, d1);
  const float g = key >= 0 ? rec.z : 0.0f;
#pragma unroll
  for (int a = 0; a < NT; ++a) {
    const float gw = __fmul_rn(g, wy[a]);
#pragma unroll
    for (int b = 0; b < NT; ++b) cc[a * NT + b] = __fmul_rn(gw, wx[b]);
  }
  bool writer;
  const int k = key >= 0 ? (t * W + cx) : -1;     // same arrival row: (class, column) identifies the departure cell
  resolve_clashes<NT>(k, ((t & 3) * 257 + cx) & (kTagBytes - 1), tag, lane, cc, writer);
#pragma unroll
  for (int a = 0; a < NT; ++a) {
    int i = i0 + a;
    bool ok = writer && i >= -pad && i <= H - 1 + pad;      // outside the padded plane: zeros
    bool fold = false;
    if (i < 0) { i = -i; fold = true; }
    else if (i >= H) { i = 2 * (H - 1) - i; fold = true; }
    int sl = head + (i - i_ret);
    ok = ok && (unsigned)(i - i_ret) < (unsigned)ring;
    if (sl >= ring) sl -= ring;
    float* row = ring_base + sl * W;
    const bool any_fold = __any_sync(0xffffffffu, ok && fold);
#pragma unroll
    for (int b = 0; b < NT; ++b) {
      int c = cx + b;
      if (c >= W) c -= W;
      if (fold) { c -= halfW; if (c < 0) c += W; }
      const bool mine = ok && (unsigned)(c - ja) < (unsigned)wc && !(last_out && b == NT - 1);
      if (mine && !fold) row[c] += cc[a * NT + b];
      __syncwarp();
      if (any_fold) {
        if (mine && fold) row[c] += cc[a * NT + b];
        __syncwarp();
      }
    }
  }
}

// ---- producer: one arrival point -> record (+ grad_u, grad_v for the rows this segment owns) -----
template <bool EXACT, int INTERP, bool PEER, bool SMALL>
__device__ __forceinline__ float4 rows_point(const Params& P, const float* __restrict__ f, int pl, float mean0,
                                             float mean1, float sp, float cp, int y, int x, int rr, int hx, bool core,
                                             float uu, float vv, float g, float lonp, bool& violated, float& ou,
                                             float& ov) {
  constexpr int NT = Stencil<INTERP>::NT, OMIN = Stencil<INTERP>::OMIN;
  Traj t;
  trajectory<EXACT, SMALL>(P, uu, vv, sp, cp, lonp, t);
  const float fx = floorf(t.ix), fy = floorf(t.iy);
  const float tx = __fsub_rn(t.ix, fx), ty = __fsub_rn(t.iy, fy);
  const int x0 = (int)fx + OMIN;                 // padded column of tap 0
  const int cls = (int)fy - (y + P.p);           // row class
  int dx = x0 - P.p - x;                         // longitudinal cell displacement of tap 0
  if (dx < -P.halfW) dx += P.W; else if (dx >= P.halfW) dx -= P.W;
  const bool in_ring = (unsigned)(cls + rr) <= (unsigned)(2 * rr);
  if (!in_ring || dx < -hx || dx > hx - NT + 1) violated = true;
  if (core) {
    float val, ddx, ddy;
    stencil_eval<INTERP, true, PEER>(P, f, pl, t, mean0, mean1, val, ddx, ddy);
    velocity_grads<EXACT>(P, t, sp, cp, g * ddx, g * ddy, ou, ov);
  }
  int col0 = x + dx;
  if (col0 < 0) col0 += P.W; else if (col0 >= P.W) col0 -= P.W;
  const bool valid = in_ring && (unsigned)col0 < (unsigned)P.W;   // false for non-finite coordinates
  const int key = valid ? rows_key(cls + rr, col0, x0 + NT - 1 >= P.Wp) : -1;
  return make_float4(tx, ty, g, __int_as_float(key));
}

template <bool EXACT, int INTERP, bool PEER>
__global__ void __launch_bounds__(kRowsWarps * 32, 1) sl_bwd_rows_kernel(const Params P, const RowsPlan S) {
  constexpr int NT = Stencil<INTERP>::NT, OMIN = Stencil<INTERP>::OMIN;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = P.W, ring = S.ring, rr = S.rr;
  float* ring_base = reinterpret_cast<float*>(smem_raw);
  float* stage_base = reinterpret_cast<float*>(smem_raw + S.off_stage);      // [kRowStages][3][W]
  float4* rec_base = reinterpret_cast<float4*>(smem_raw + S.off_rec);        // [kRowRecs][W]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + S.off_bar);
  uint64_t* stage_full = bars;
  uint64_t* stage_free = bars + kRowStages;
  uint64_t* rec_full = bars + 2 * kRowStages;
  uint64_t* rec_free = bars + 2 * kRowStages + kRowRecs;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRowStages; ++i) { mbar_init(&stage_full[i], 1); mbar_init(&stage_free[i], S.nP); }
    for (int i = 0; i < kRowRecs; ++i) { mbar_init(&rec_full[i], S.nP); mbar_init(&rec_free[i], S.nC); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  int g = rows_bound(S, P.ownN, blockIdx.x, gridDim.x);
  const int g1 = rows_bound(S, P.ownN, blockIdx.x + 1, gridDim.x);
  const int arr_lo = P.arr0, arr_hi = P.arr0 + P.arrN;
  RowsSeg sg;

  if (warp == 0) {
    // ===================================== loader =====================================
    if (lane != 0) return;
    int n = 0;
    const uint32_t row_bytes = (uint32_t)W * 4u;
    while (rows_next_seg<INTERP>(P, S, g, g1, sg)) {
      const float* up = plane_ptr(P.u, P.u_sB, P.V, P.uvgN, W, sg.pl);
      const float* vp = plane_ptr(P.v, P.v_sB, P.V, P.uvgN, W, sg.pl);
      const float* gp = plane_ptr(P.gout, P.gout_sB, P.V, P.uvgN, W, sg.pl);
      for (int y = max(sg.y_first, arr_lo); y <= min(sg.y_last, arr_hi - 1); ++y, ++n) {
        const int st = n % kRowStages;
        mbar_wait(&stage_free[st], ((n / kRowStages) & 1) ^ 1);
        float* dst = stage_base + (size_t)st * 3 * W;
        mbar_arrive_expect_tx(&stage_full[st], 3 * row_bytes);
        bulk_g2s(dst, arr_row<PEER>(P, up, 0, sg.pl, y), row_bytes, &stage_full[st]);
        bulk_g2s(dst + W, arr_row<PEER>(P, vp, 1, sg.pl, y), row_bytes, &stage_full[st]);
        bulk_g2s(dst + 2 * W, arr_row<PEER>(P, gp, 2, sg.pl, y), row_bytes, &stage_full[st]);
      }
    }
    return;
  }

  if (warp <= S.nP) {
    // ==================================== producers ====================================
    const int p = warp - 1;
    int n = 0;
    int next = p;                     // global step index (row sequence number * nsteps + step) this warp does next
    bool violated = false;
    int flagged_pl = -1;
    while (rows_next_seg<INTERP>(P, S, g, g1, sg)) {
      const int pl = sg.pl;
      const int b = pl / P.V, c = pl - b * P.V;
      const float* f = plane_ptr_bc(P.field, P.field_sB, b, c, P.fldN, W);
      float* gu_pl = P.gu ? P.gu + (long long)pl * S.outN * W : nullptr;
      float* gv_pl = P.gu ? P.gv + (long long)pl * S.outN * W : nullptr;
      float mean0 = 0.0f, mean1 = 0.0f, gm0 = 0.0f, gm1 = 0.0f;
      if (P.pole_fix) {
        mean0 = __ldg(P.fmean + 2 * pl); mean1 = __ldg(P.fmean + 2 * pl + 1);
        gm0 = __ldg(P.gmean + 2 * pl); gm1 = __ldg(P.gmean + 2 * pl + 1);
      }
      for (int y = max(sg.y_first, arr_lo); y <= min(sg.y_last, arr_hi - 1); ++y, ++n) {
        const int row_end = (n + 1) * S.nsteps;
        const int st = n % kRowStages, rs = n % kRowRecs;
        // Every producer warp takes part in every row of both pipelines (it arrives on stage_free and rec_full
        // whether or not one of the row's steps is its own): a parity wait can only tell the current phase of an
        // mbarrier from the previous one, so no waiter may fall -- or run -- more than one phase away from it.
        mbar_wait(&stage_full[st], (n / kRowStages) & 1);
        mbar_wait(&rec_free[rs], ((n / kRowRecs) & 1) ^ 1);
        if (next >= row_end) {
          if (lane == 0) { mbar_arrive(&stage_free[st]); mbar_arrive(&rec_full[rs]); }
          continue;
        }
        const float* su = stage_base + (size_t)st * 3 * W;
        float4* recs = rec_base + (size_t)rs * W;
        const float sp = __ldg(P.sin_lat + y), cp = __ldg(P.cos_lat + y);
        const int hx = __ldg(S.hx_tab + y);
        const bool core = (y >= sg.ra) && (y < sg.rb) && gu_pl != nullptr;
        const bool pole_row = P.pole_fix && (y == 0 || y == P.H - 1);
        const float gpole = y == 0 ? gm0 : gm1;
        for (; next < row_end; next += S.nP) {
          const int xb = (next - n * S.nsteps) * (32 * kStepSub);
          float uu[kStepSub], vv[kStepSub], gg[kStepSub], ll[kStepSub];
#pragma unroll
          for (int j = 0; j < kStepSub; ++j) {
            const int x = min(xb + 32 * j + lane, W - 1);   // lanes past the row end redo its last point (not stored)
            uu[j] = su[x];
            vv[j] = su[W + x];
            gg[j] = pole_row ? gpole : su[2 * W + x];       // adjoint of the output pole mean (advection.py:169)
            ll[j] = __ldg(P.lon + x);
          }
          const bool last_step = next + S.nP >= row_end;     // this warp's last step in this row
          __syncwarp();
          if (last_step && lane == 0) mbar_arrive(&stage_free[st]);
          float4 rec[kStepSub];
          float ou[kStepSub], ov[kStepSub];
#pragma unroll
          for (int j = 0; j < kStepSub; ++j)
            rec[j] = rows_point<EXACT, INTERP, PEER, false>(P, f, pl, mean0, mean1, sp, cp, y, min(xb + 32 * j + lane, W - 1), rr, hx,
                                                            core, uu[j], vv[j], gg[j], ll[j], violated, ou[j], ov[j]);
#pragma unroll
          for (int j = 0; j < kStepSub; ++j) {
            const int x = xb + 32 * j + lane;
            if (x < W) {
              recs[x] = rec[j];
              if (core) {
                __stcs(gu_pl + (long long)(y - S.out0) * W + x, ou[j]);
                __stcs(gv_pl + (long long)(y - S.out0) * W + x, ov[j]);
              }
            }
          }
          if (last_step) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&rec_full[rs]);
          }
        }
      }
      if (__any_sync(0xffffffffu, violated) && flagged_pl != pl) {
        if (lane == 0) S.plane_flag[pl] = 1;
        flagged_pl = pl;
      }
      violated = false;
    }
    return;
  }

  if (warp <= S.nP + S.nC) {
    // ==================================== consumers ====================================
    const int cidx = warp - 1 - S.nP;
    const int ja = cidx * S.wc, jb = min(ja + S.wc, W), wc = jb - ja;
    unsigned char* tag = smem_raw + S.off_tag + cidx * kTagBytes;
    int n = 0;
    while (rows_next_seg<INTERP>(P, S, g, g1, sg)) {
      float* gf_pl = P.gfield + (long long)sg.pl * S.outN * W;
      for (int r = 0; r < ring; ++r)
        for (int k = 4 * lane; k < wc; k += 128)
          *reinterpret_cast<float4*>(ring_base + r * W + ja + k) = make_float4(0.f, 0.f, 0.f, 0.f);
      __syncwarp();
      int head = 0;
      for (int y = sg.y_first; y <= sg.y_last; ++y) {
        if (y >= arr_lo && y < arr_hi) {
          const int rs = n % kRowRecs;
          mbar_wait(&rec_full[rs], (n / kRowRecs) & 1);
          const float4* recs = rec_base + (size_t)rs * W;
          const int hx = __ldg(S.hx_tab + y);
          int start = ja - hx, len = wc + 2 * hx;
          if (hx >= W || len >= W) { start = 0; len = W; }
          if (start < 0) start += W;
          const bool fold_row = (y + OMIN - rr < 0) || (y + OMIN + rr + NT - 1 >= P.H);
          for (int s = 0; s < len; s += 32) {
            const bool inr = s + lane < len;
            int idx = start + s + lane;
            if (idx >= W) idx -= W;
            const float4 rec = inr ? recs[idx] : make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
            if (fold_row) rows_scatter_fold<INTERP>(W, P.H, rec, inr, ring_base, tag, ring, head, y, rr, ja, wc, lane);
            else rows_scatter_plain<INTERP>(rec, inr, ring_base, tag, W, ring, head, ja, wc, lane);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&rec_free[rs]);
          ++n;
        }
        // retire destination row i = y - rr + OMIN (slot `head`)
        const int i = y - rr + OMIN;
        float* row = ring_base + head * W + ja;
        if (i >= sg.ra && i < sg.rb) {
          float* orow = gf_pl + (long long)(i - S.out0) * W + ja;
          for (int k = 4 * lane; k < wc; k += 128)
            __stcs(reinterpret_cast<float4*>(orow + k), *reinterpret_cast<const float4*>(row + k));
        }
        __syncwarp();
        for (int k = 4 * lane; k < wc; k += 128) *reinterpret_cast<float4*>(row + k) = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        head = head + 1 == ring ? 0 : head + 1;
      }
    }
  }
}

}  // namespace psl
