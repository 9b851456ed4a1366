// Test harness (host only): runs the row sweep's work partition (csrc/sl_partition.cuh) for the shapes given on stdin and
// prints, per shape, the CTA bounds as JSON.  Built and run by tests/test_partition.py; no GPU involved.
// input lines: interp H W planes own0 ownN arr0 arrN cfl cut grid consumers
#include "../paradis_model_b200/csrc/sl_device.cuh"
// helpers the kernel templates of the headers refer to; they live in paradis_sl.cu and are never called here
__device__ const float* plane_ptr(const float* base, long long sB, int V, int rows, int W, int pl);
__device__ const float* plane_ptr_bc(const float* base, long long sB, int b, int c, int rows, int W);
#include "../paradis_model_b200/csrc/sl_partition.cuh"
#include <cstdio>
using namespace psl;

template <int INTERP>
static void run(int H, int W, int planes, int own0, int ownN, int arr0, int arrN, float cfl, int cut, int grid, int nC) {
  constexpr int NT = Stencil<INTERP>::NT;
  const int rr = (int)ceil(cfl);
  Params P;
  memset(&P, 0, sizeof(P));
  P.H = H; P.W = W; P.B = 1; P.V = planes; P.own0 = own0; P.ownN = ownN; P.arr0 = arr0; P.arrN = arrN;
  P.min_lat = -1.5707963f; P.d_lat = 3.1415926f;
  const double dphi = 3.14159265358979 / (H - 1), dlam = 6.28318530717959 / W, delta = cfl * dphi;
  RowsPlan S;
  memset(&S, 0, sizeof(S));
  S.planes = planes; S.rr = rr; S.ring = 2 * rr + NT; S.cut = cut; S.GR = rr + NT;
  int wc = ((W + nC - 1) / nC + 3) & ~3;
  if (wc < 32) wc = 32;
  S.total_rows = planes * ownN;
  ReachModel reach;
  reach.sin_delta = (float)sin(delta); reach.cos_delta = (float)cos(delta);
  reach.inv_dlam = (float)(1.0 / dlam); reach.extra = NT + 2; reach.max_halo = 0;
  rows_partition<INTERP>(P, S, reach, wc, planes, grid);
  // per-row cost of the model, for the caller's balance check
  printf("{\"cut\": %d, \"ring\": %d, \"GR\": %d, \"wc\": %d, \"bound\": [", S.cut, S.ring, S.GR, wc);
  for (int c = 0; c <= grid; ++c) printf("%d%s", S.bound[c], c < grid ? ", " : "");
  printf("], \"hx\": [");
  for (int y = 0; y < H; ++y) {
    const double lat = -1.5707963 + y * dphi;
    const int hx = halo_cells(reach, (float)sin(lat), (float)cos(lat));
    printf("%d%s", hx < W ? hx : W, y < H - 1 ? ", " : "");
  }
  printf("]}\n");
}

int main() {
  int interp, H, W, planes, own0, ownN, arr0, arrN, cut, grid, nC;
  float cfl;
  while (scanf("%d %d %d %d %d %d %d %d %f %d %d %d", &interp, &H, &W, &planes, &own0, &ownN, &arr0, &arrN, &cfl, &cut,
               &grid, &nC) == 12) {
    if (interp == 1) run<1>(H, W, planes, own0, ownN, arr0, arrN, cfl, cut, grid, nC);
    else run<2>(H, W, planes, own0, ownN, arr0, arrN, cfl, cut, grid, nC);
  }
  return 0;
}
