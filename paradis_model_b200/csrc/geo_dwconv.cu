// Depthwise k x k convolution with the GeoCyclic padding applied on the fly (sm_100a).
//
// Replaces, in one kernel, `GeoCyclicPadding((k-1)/2)` followed by the depthwise `nn.Conv2d(C, C, k,
// groups=C)` of the reference's SepConv (model/blocks.py:92-116) and of the static encoder
// (model/paradis.py:186-190): the padded tensor [B, C, H+k-1, W+k-1] is never written or read.
// SURVEY section 8(f) rank 3.  Algorithmic traffic: read x (4 B) + write y (4 B) per element instead of
// 16 B for pad-then-convolve.
//
//   forward      y[i,j]  = bias + sum_{a,b} w[a,b] * xpad[i+a, j+b]
//   grad input   gx      = P^T ( full correlation of gy with w ): longitude is periodic, so the direct part
//                          is the same tiled kernel with the flipped filter, zero rows outside the mesh and
//                          circular columns; the two polar caps fold a few rows back (fixed order, no atomics)
//   grad weight  gw[a,b] = sum_{n,i,j} gy[i,j] * xpad[i+a, j+b]: per-tile partial sums, then a fixed-order
//                          reduction over tiles and batch (deterministic)
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/paradis_sl.h"

void psl_set_error(const char* msg);   // paradis_sl.cu: the calling thread's paradis_last_error() string

namespace {

constexpr int TW = 32, TH = 32, RPT = 4;   // tile 32 x 32 outputs, 256 threads, 4 vertically adjacent outputs each

// source of padded cell (i, j) (unpadded coordinates, may be outside [0,H) x [0,W)); model/padding.py:26-37
__device__ __forceinline__ float geo_load(const float* __restrict__ x, int i, int j, int H, int W, int P) {
  if (i < -P || i >= H + P || j < -P || j >= W + P) return 0.0f;
  int shift = 0;
  if (i < 0) { i = -i; shift = W / 2; }
  else if (i >= H) { i = 2 * (H - 1) - i; shift = W / 2; }
  j -= shift;
  if (j < 0) j += W; else if (j >= W) j -= W;
  return __ldg(x + (long long)i * W + j);
}

// zero rows outside the mesh, circular columns (the adjoint's view of grad_y)
__device__ __forceinline__ float zc_load(const float* __restrict__ g, int i, int j, int H, int W) {
  if (i < 0 || i >= H) return 0.0f;
  if (j < 0) j += W; else if (j >= W) j -= W;
  if (j < 0 || j >= W) return 0.0f;
  return __ldg(g + (long long)i * W + j);
}

// MODE 0: forward (GeoCyclic source, filter as is, + bias).  MODE 1: direct part of grad input
// (zero-row / circular-column source, flipped filter).
template <int K, int MODE>
__global__ void __launch_bounds__(256) geo_dwconv_tile_kernel(const float* __restrict__ src, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ dst,
                                                              int C, int H, int W) {
  constexpr int P = (K - 1) / 2, SW = TW + K - 1, SH = TH + K - 1;
  __shared__ float tile[SH][SW + 1];
  const int plane = blockIdx.z, c = plane % C;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const float* s = src + (long long)plane * H * W;
  for (int idx = threadIdx.x; idx < SH * SW; idx += 256) {
    const int r = idx / SW, q = idx - r * SW;
    tile[r][q] = MODE == 0 ? geo_load(s, y0 + r - P, x0 + q - P, H, W, P) : zc_load(s, y0 + r - P, x0 + q - P, H, W);
  }
  float wk[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) wk[t] = __ldg(w + c * K * K + (MODE == 0 ? t : K * K - 1 - t));
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[RPT];
  const float b0 = (MODE == 0 && bias) ? __ldg(bias + c) : 0.0f;
#pragma unroll
  for (int r = 0; r < RPT; ++r) acc[r] = b0;
#pragma unroll
  for (int b = 0; b < K; ++b) {
    float v[RPT + K - 1];
#pragma unroll
    for (int m = 0; m < RPT + K - 1; ++m) v[m] = tile[ty * RPT + m][tx + b];
#pragma unroll
    for (int a = 0; a < K; ++a)
#pragma unroll
      for (int r = 0; r < RPT; ++r) acc[r] = fmaf(wk[a * K + b], v[r + a], acc[r]);
  }
  const int j = x0 + tx;
  if (j < W) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int i = y0 + ty * RPT + r;
      if (i < H) dst[((long long)plane * H + i) * W + j] = acc[r];
    }
  }
}

// Polar-cap part of grad input: rows 1..P receive the fold of padded rows P-i (north), rows H-1-P..H-2 the fold
// of padded rows 2(H-1)-i+P (south), both shifted by W/2.  Runs after the direct kernel (single writer per cell).
template <int K>
__global__ void geo_dwconv_bwd_caps_kernel(const float* __restrict__ gy, const float* __restrict__ w,
                                           float* __restrict__ gx, int C, int H, int W) {
  constexpr int P = (K - 1) / 2;
  const int plane = blockIdx.z, c = plane % C;
  const int which = blockIdx.y / P, k = blockIdx.y % P;            // which: 0 north, 1 south
  const int i = which == 0 ? 1 + k : H - 2 - k;                    // source row
  if (i < 1 || i > H - 2) return;
  if (which == 1 && i <= P && H - 2 - k <= P) { /* tiny meshes: both folds hit the row, handled below alike */ }
  const float* g = gy + (long long)plane * H * W;
  const float* wc = w + c * K * K;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < W; j += gridDim.x * blockDim.x) {
    float s = 0.0f;
    const int R = which == 0 ? P - i : 2 * (H - 1) - i + P;       // padded row folding onto row i
    for (int a = 0; a < K; ++a) {
      const int gi = R - a;                                        // output row that used padded row R with tap a
      if (gi < 0 || gi >= H) continue;
      for (int b = 0; b < K; ++b) {
        int gj = j + W / 2 + P - b;                                // padded column C = j + W/2 + P (mod W), output col C - b
        gj %= W;
        if (gj < 0) gj += W;
        s = fmaf(__ldg(wc + a * K + b), __ldg(g + (long long)gi * W + gj), s);
      }
    }
    gx[((long long)plane * H + i) * W + j] += s;
  }
}

// grad weight / grad bias: per-tile partial sums, [plane][tile][K*K+1]
template <int K>
__global__ void __launch_bounds__(256) geo_dwconv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                               float* __restrict__ partial, int C, int H, int W) {
  constexpr int P = (K - 1) / 2, SW = TW + K - 1, SH = TH + K - 1, NV = K * K + 1;
  __shared__ float tile[SH][SW + 1];
  __shared__ float red[8][NV];
  const int plane = blockIdx.z;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const float* s = x + (long long)plane * H * W;
  for (int idx = threadIdx.x; idx < SH * SW; idx += 256) {
    const int r = idx / SW, q = idx - r * SW;
    tile[r][q] = geo_load(s, y0 + r - P, x0 + q - P, H, W, P);
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, lane = tx;
  float g[RPT];
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int i = y0 + ty * RPT + r, j = x0 + tx;
    g[r] = (i < H && j < W) ? __ldg(gy + ((long long)plane * H + i) * W + j) : 0.0f;
  }
  float acc[NV];
#pragma unroll
  for (int t = 0; t < NV; ++t) acc[t] = 0.0f;
#pragma unroll
  for (int b = 0; b < K; ++b) {
    float v[RPT + K - 1];
#pragma unroll
    for (int m = 0; m < RPT + K - 1; ++m) v[m] = tile[ty * RPT + m][tx + b];
#pragma unroll
    for (int a = 0; a < K; ++a)
#pragma unroll
      for (int r = 0; r < RPT; ++r) acc[a * K + b] = fmaf(g[r], v[r + a], acc[a * K + b]);
  }
#pragma unroll
  for (int r = 0; r < RPT; ++r) acc[K * K] += g[r];
  // fixed-order reduction: lanes (xor tree), then the 8 warps in order
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    float vsum = acc[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
    if (lane == 0) red[ty][t] = vsum;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    float vsum = 0.0f;
    for (int wq = 0; wq < 8; ++wq) vsum += red[wq][threadIdx.x];
    const int tile_id = blockIdx.y * gridDim.x + blockIdx.x, ntiles = gridDim.x * gridDim.y;
    partial[((long long)plane * ntiles + tile_id) * NV + threadIdx.x] = vsum;
  }
}

// gw[c][t] = sum over batch n and tiles (ascending) of partial[n*C+c][tile][t]
__global__ void geo_dwconv_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ gw,
                                               float* __restrict__ gbias, int B, int C, int ntiles, int NV) {
  const int c = blockIdx.x, t = threadIdx.x;
  if (t >= NV) return;
  float s = 0.0f;
  for (int n = 0; n < B; ++n) {
    const float* p = partial + ((long long)(n * C + c) * ntiles) * NV + t;
    for (int k = 0; k < ntiles; ++k) s += p[(long long)k * NV];
  }
  if (t < NV - 1) gw[c * (NV - 1) + t] = s;
  else if (gbias) gbias[c] = s;
}

int err(int code, const char* msg) {
  psl_set_error(msg);
  return code;
}

int check(const void* a, const void* b, int B, int C, int H, int W, int k) {
  if (!a || !b) return err(PARADIS_ERR_NULL_POINTER, "geocyclic_dwconv: NULL tensor pointer");
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return err(PARADIS_ERR_BAD_SHAPE, "geocyclic_dwconv: non-positive dimension");
  if (W % 2) return err(PARADIS_ERR_ODD_WIDTH, "Number of longitude points must be even");
  if (k != 3 && k != 5 && k != 7) return err(PARADIS_ERR_BAD_INTERP, "geocyclic_dwconv: kernel size must be 3, 5 or 7");
  // H >= 2P + 2: the north and the south cap never fold onto the same row (the cap kernel of the backward adds one
  // fold per block, see geo_dwconv_bwd_caps_kernel)
  if (H < k + 1 || W < k - 1) return err(PARADIS_ERR_BAD_SHAPE, "geocyclic_dwconv: mesh too small for the kernel size (H >= k + 1, W >= k - 1)");
  if ((long long)B * C > 65535) return err(PARADIS_ERR_BAD_SHAPE, "geocyclic_dwconv: B*C exceeds 65535 planes per call");
  return PARADIS_OK;
}

template <int K>
int run(int what, const float* a, const float* w, const float* bias, float* out, float* out2, float* ws, int B, int C,
        int H, int W, cudaStream_t st) {
  dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, B * C);
  if (what == 0) geo_dwconv_tile_kernel<K, 0><<<grid, 256, 0, st>>>(a, w, bias, out, C, H, W);
  else if (what == 1) {
    geo_dwconv_tile_kernel<K, 1><<<grid, 256, 0, st>>>(a, w, nullptr, out, C, H, W);
    dim3 cg((W + 255) / 256, 2 * ((K - 1) / 2), B * C);
    geo_dwconv_bwd_caps_kernel<K><<<cg, 256, 0, st>>>(a, w, out, C, H, W);
  } else {
    geo_dwconv_wgrad_kernel<K><<<grid, 256, 0, st>>>(a, w /* = gy */, ws, C, H, W);
    geo_dwconv_wgrad_reduce_kernel<<<C, 64, 0, st>>>(ws, out, out2, B, C, grid.x * grid.y, K * K + 1);
  }
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? PARADIS_OK : err(PARADIS_ERR_CUDA, cudaGetErrorString(e));
}

int dispatch(int what, const float* a, const float* w, const float* bias, float* out, float* out2, float* ws, int B,
             int C, int H, int W, int k, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (k == 3) return run<3>(what, a, w, bias, out, out2, ws, B, C, H, W, st);
  if (k == 5) return run<5>(what, a, w, bias, out, out2, ws, B, C, H, W, st);
  return run<7>(what, a, w, bias, out, out2, ws, B, C, H, W, st);
}


// PhysicalDownsample (model/blocks.py:57-71): GeoCyclic pad 2 + AvgPool2d(5, stride) = the 5x5 box mean of the padded
// field at every stride-th point.  One thread per OUTPUT point: only Ho x Wo values are computed and written (the
// round-1 version ran the full-resolution depthwise kernel and discarded all but 1 / stride^2 of it).
__global__ void geo_avgpool5_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int Ho, int Wo,
                                    int stride) {
  const long long pl = blockIdx.z;
  const int io = blockIdx.y, jo = blockIdx.x * blockDim.x + threadIdx.x;
  if (jo >= Wo) return;
  const float* xp = x + pl * (long long)H * W;
  const int halfW = W >> 1;
  float s = 0.0f;
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    int i = io * stride + a - 2, shift = 0;         // unpadded row of the tap
    if (i < 0) { i = -i; shift = halfW; }
    else if (i >= H) { i = 2 * (H - 1) - i; shift = halfW; }
    const float* row = xp + (long long)i * W;
    float r = 0.0f;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
      int j = jo * stride + b - 2 - shift;
      if (j < 0) j += W;
      if (j < 0) j += W;
      if (j >= W) j -= W;
      r += __ldg(row + j);
    }
    s += r;
  }
  y[(pl * Ho + io) * Wo + jo] = s * (1.0f / 25.0f);
}

}  // namespace

extern "C" int paradis_geocyclic_dwconv_fwd(const float* x, const float* weight, const float* bias, float* y, int B,
                                            int C, int H, int W, int k, void* stream) {
  if (int rc = check(x, y, B, C, H, W, k)) return rc;
  if (!weight) return err(PARADIS_ERR_NULL_POINTER, "geocyclic_dwconv: weight is NULL");
  return dispatch(0, x, weight, bias, y, nullptr, nullptr, B, C, H, W, k, stream);
}

extern "C" int paradis_geocyclic_dwconv_bwd_input(const float* gy, const float* weight, float* gx, int B, int C, int H,
                                                  int W, int k, void* stream) {
  if (int rc = check(gy, gx, B, C, H, W, k)) return rc;
  if (!weight) return err(PARADIS_ERR_NULL_POINTER, "geocyclic_dwconv: weight is NULL");
  return dispatch(1, gy, weight, nullptr, gx, nullptr, nullptr, B, C, H, W, k, stream);
}

extern "C" size_t paradis_geocyclic_dwconv_wgrad_workspace(int B, int C, int H, int W, int k) {
  const size_t tiles = (size_t)((W + TW - 1) / TW) * ((H + TH - 1) / TH);
  return (size_t)B * C * tiles * (k * k + 1) * sizeof(float);
}

extern "C" int paradis_geocyclic_dwconv_bwd_weight(const float* x, const float* gy, float* gweight, float* gbias,
                                                   int B, int C, int H, int W, int k, void* workspace,
                                                   size_t workspace_bytes, void* stream) {
  if (int rc = check(x, gy, B, C, H, W, k)) return rc;
  if (!gweight) return err(PARADIS_ERR_NULL_POINTER, "geocyclic_dwconv: grad weight is NULL");
  if (!workspace || workspace_bytes < paradis_geocyclic_dwconv_wgrad_workspace(B, C, H, W, k))
    return err(PARADIS_ERR_WORKSPACE, "geocyclic_dwconv: weight-gradient workspace too small");
  return dispatch(2, x, gy, nullptr, gweight, gbias, (float*)workspace, B, C, H, W, k, stream);
}

extern "C" int paradis_geocyclic_avgpool5_fwd(const float* x, float* y, int B, int C, int H, int W, int stride, void* stream) {
  if (int rc = check(x, y, B, C, H, W, 5)) return rc;
  if (stride < 1) return err(PARADIS_ERR_BAD_SHAPE, "geocyclic_avgpool5: stride must be >= 1");
  const int Ho = (H + 4 - 5) / stride + 1, Wo = (W + 4 - 5) / stride + 1;     // AvgPool2d(5, stride) of the padded plane
  dim3 grid((Wo + 127) / 128, Ho, B * C);
  geo_avgpool5_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, y, H, W, Ho, Wo, stride);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? PARADIS_OK : err(PARADIS_ERR_CUDA, cudaGetErrorString(e));
}
