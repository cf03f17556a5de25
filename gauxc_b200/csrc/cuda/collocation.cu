// K_A: Gaussian basis collocation (and gradient) for shell-batched tiles.
//
// Replaces collocation_device_shell_to_task_kernel_* / masked_combined kernels
// (src/xc_integrator/local_work_driver/device/cuda/kernels/collocation_device.cu:280-443)
// and matches the host semantics of gau2grid gg_collocation[_deriv1]
// (local_work_driver/host/reference/gau2grid_collocation.cxx:25-116): every listed shell
// is evaluated at every point of the task (no per-point screening), CCA ordering,
// cartesian xx,xy,xz,yy,yz,zz; spherical m=-l..l (p pure: y,z,x).
//
// One CTA = one tile of TP points, thread = point; the tile's shells are staged in shared
// memory once and broadcast.  Output matrices are [mu][TP], point index fastest (XOR-swizzled
// inside 128-byte lines, rows padded to a multiple of 16), so every store is a fully coalesced
// 1 KB row.  Bound: FP64 exp/ALU; HBM traffic = the write of
// 8*k*nbe bytes per point (k = 1 LDA, 4 GGA).
#include "kernels.cuh"

namespace gxb {

namespace {

// element (row, i) of a tile matrix lives at column i ^ ((row & 3) << 2) (device_plan.hpp)
template <bool GRAD>
__device__ __forceinline__ void store(double* __restrict__ B, size_t ms, int row, int i, bool ok,
                                      double v, double gx, double gy, double gz) {
  const size_t o = (size_t)row * TP + swz(row, i);
  B[o] = ok ? v : 0.;
  if (GRAD) {
    B[o + ms] = ok ? gx : 0.;
    B[o + 2 * ms] = ok ? gy : 0.;
    B[o + 3 * ms] = ok ? gz : 0.;
  }
}

template <bool GRAD>
__global__ void __launch_bounds__(TP) collocation_kernel(PlanView pv,
                                                          const DevTile* __restrict__ tiles,
                                                          double* __restrict__ ws) {
  const DevTile tile = tiles[blockIdx.x];
  const DevTask task = pv.tasks[tile.task];
  const int i = threadIdx.x;
  if (i >= tile_width(tile.npts)) return;  // columns beyond the tile width are never read
  const bool ok = i < tile.npts;
  const int ip = tile.pt_off + (ok ? i : 0);
  const double px = pv.px[ip], py = pv.py[ip], pz = pv.pz[ip];
  double* __restrict__ B = ws + tile.ws_off;
  const int nbp = pad16(task.nbe);
  const size_t ms = (size_t)nbp * TP;
  // zero pad rows: they are K rows of the X = B P contraction
  for (int r = task.nbe; r < nbp; ++r) store<GRAD>(B, ms, r, i, false, 0., 0., 0., 0.);

  const double sqrt3 = 1.7320508075688772935;

  for (int s = 0; s < task.nshells; ++s) {
    const DevShell sh = pv.shells[pv.task_shells[task.shell_off + s]];
    const int bf = pv.task_shell_bf[task.shell_off + s];
    const double x = px - sh.x, y = py - sh.y, z = pz - sh.z;
    const double r2 = x * x + y * y + z * z;
    double S0 = 0., S1 = 0.;
    const double* __restrict__ al = pv.prim_alpha + sh.prim_off;
    const double* __restrict__ co = pv.prim_coeff + sh.prim_off;
    for (int k = 0; k < sh.nprim; ++k) {
      const double a = __ldg(al + k);
      const double e = __ldg(co + k) * exp(-a * r2);
      S0 += e;
      if (GRAD) S1 += -2. * a * e;
    }
    const double S1x = S1 * x, S1y = S1 * y, S1z = S1 * z;

    if (sh.l == 0) {
      store<GRAD>(B, ms, bf, i, ok, S0, S1x, S1y, S1z);
    } else if (sh.l == 1) {
      // d/da (f S0) = (d f/da) S0 + f a S1
      const double vx = x * S0, vy = y * S0, vz = z * S0;
      const int rx = sh.pure ? bf + 2 : bf, ry = sh.pure ? bf : bf + 1,
                rz = sh.pure ? bf + 1 : bf + 2;
      store<GRAD>(B, ms, rx, i, ok, vx, S0 + x * S1x, x * S1y, x * S1z);
      store<GRAD>(B, ms, ry, i, ok, vy, y * S1x, S0 + y * S1y, y * S1z);
      store<GRAD>(B, ms, rz, i, ok, vz, z * S1x, z * S1y, S0 + z * S1z);
    } else if (sh.l == 2) {
      // cartesian monomials and gradients
      const double xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
      double v[6], gx[6], gy[6], gz[6];
      v[0] = xx * S0; v[1] = xy * S0; v[2] = xz * S0; v[3] = yy * S0; v[4] = yz * S0; v[5] = zz * S0;
      if (GRAD) {
        const double x2 = 2. * x * S0, y2 = 2. * y * S0, z2 = 2. * z * S0;
        const double xS = x * S0, yS = y * S0, zS = z * S0;
        gx[0] = x2 + xx * S1x; gy[0] = xx * S1y;      gz[0] = xx * S1z;
        gx[1] = yS + xy * S1x; gy[1] = xS + xy * S1y; gz[1] = xy * S1z;
        gx[2] = zS + xz * S1x; gy[2] = xz * S1y;      gz[2] = xS + xz * S1z;
        gx[3] = yy * S1x;      gy[3] = y2 + yy * S1y; gz[3] = yy * S1z;
        gx[4] = yz * S1x;      gy[4] = zS + yz * S1y; gz[4] = yS + yz * S1z;
        gx[5] = zz * S1x;      gy[5] = zz * S1y;      gz[5] = z2 + zz * S1z;
      }
      if (!sh.pure) {
#pragma unroll
        for (int c = 0; c < 6; ++c) store<GRAD>(B, ms, bf + c, i, ok, v[c], gx[c], gy[c], gz[c]);
      } else {
        // m=-2: sqrt3 xy; -1: sqrt3 yz; 0: zz - (xx+yy)/2; 1: sqrt3 xz; 2: sqrt3/2 (xx-yy)
        store<GRAD>(B, ms, bf + 0, i, ok, sqrt3 * v[1], sqrt3 * gx[1], sqrt3 * gy[1], sqrt3 * gz[1]);
        store<GRAD>(B, ms, bf + 1, i, ok, sqrt3 * v[4], sqrt3 * gx[4], sqrt3 * gy[4], sqrt3 * gz[4]);
        store<GRAD>(B, ms, bf + 2, i, ok, v[5] - 0.5 * (v[0] + v[3]), gx[5] - 0.5 * (gx[0] + gx[3]),
                    gy[5] - 0.5 * (gy[0] + gy[3]), gz[5] - 0.5 * (gz[0] + gz[3]));
        store<GRAD>(B, ms, bf + 3, i, ok, sqrt3 * v[2], sqrt3 * gx[2], sqrt3 * gy[2], sqrt3 * gz[2]);
        const double h = 0.5 * sqrt3;
        store<GRAD>(B, ms, bf + 4, i, ok, h * (v[0] - v[3]), h * (gx[0] - gx[3]),
                    h * (gy[0] - gy[3]), h * (gz[0] - gz[3]));
      }
    } else {
      // generic l = 3,4: loop over cartesian monomials; pure shells use the solid-harmonic
      // tables below.
      const int l = sh.l;
      double xp[5], yp[5], zp[5];
      xp[0] = yp[0] = zp[0] = 1.;
      for (int k = 1; k <= 4; ++k) { xp[k] = xp[k - 1] * x; yp[k] = yp[k - 1] * y; zp[k] = zp[k - 1] * z; }
      // cartesian value/gradient of monomial (a,b,c)
      auto mono = [&](int a, int b, int c, double& vv, double& dx, double& dy, double& dz) {
        const double f = xp[a] * yp[b] * zp[c];
        vv = f * S0;
        if (GRAD) {
          dx = (a ? a * xp[a - 1] * yp[b] * zp[c] * S0 : 0.) + f * S1x;
          dy = (b ? b * xp[a] * yp[b - 1] * zp[c] * S0 : 0.) + f * S1y;
          dz = (c ? c * xp[a] * yp[b] * zp[c - 1] * S0 : 0.) + f * S1z;
        }
      };
      if (!sh.pure) {
        int c = 0;
        for (int a = l; a >= 0; --a)
          for (int b = l - a; b >= 0; --b, ++c) {
            double vv, dx = 0, dy = 0, dz = 0;
            mono(a, b, l - a - b, vv, dx, dy, dz);
            store<GRAD>(B, ms, bf + c, i, ok, vv, dx, dy, dz);
          }
      } else {
        // sparse real solid harmonics, rows m=-l..l: terms {a,b,c,coef}
        struct T { signed char a, b, c; double f; };
        // l = 3
        const double s10 = 0.79056941504209483300, s15 = 3.8729833462074168852,
                     s6 = 0.61237243569579452455, s15h = 1.9364916731037084426;
        const T f3[7][3] = {
            {{2, 1, 0, 3 * s10}, {0, 3, 0, -s10}, {0, 0, 0, 0}},
            {{1, 1, 1, s15}, {0, 0, 0, 0}, {0, 0, 0, 0}},
            {{0, 1, 2, 4 * s6}, {2, 1, 0, -s6}, {0, 3, 0, -s6}},
            {{0, 0, 3, 1.0}, {2, 0, 1, -1.5}, {0, 2, 1, -1.5}},
            {{1, 0, 2, 4 * s6}, {3, 0, 0, -s6}, {1, 2, 0, -s6}},
            {{2, 0, 1, s15h}, {0, 2, 1, -s15h}, {0, 0, 0, 0}},
            {{3, 0, 0, s10}, {1, 2, 0, -3 * s10}, {0, 0, 0, 0}}};
        // l = 4
        const double s35h = 2.9580398915498080213, s70q = 2.0916500663351888699,
                     s5h = 1.1180339887498948482, s10q = 0.79056941504209483300,
                     s5q = 0.55901699437494742410, s35e = 0.73950997288745200532;
        const T g4[9][6] = {
            {{3, 1, 0, s35h}, {1, 3, 0, -s35h}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}},
            {{2, 1, 1, 3 * s70q}, {0, 3, 1, -s70q}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}},
            {{1, 1, 2, 6 * s5h}, {3, 1, 0, -s5h}, {1, 3, 0, -s5h}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}},
            {{0, 1, 3, 4 * s10q}, {2, 1, 1, -3 * s10q}, {0, 3, 1, -3 * s10q}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}},
            {{0, 0, 4, 1.0}, {2, 0, 2, -3.0}, {0, 2, 2, -3.0}, {4, 0, 0, 0.375}, {2, 2, 0, 0.75}, {0, 4, 0, 0.375}},
            {{1, 0, 3, 4 * s10q}, {3, 0, 1, -3 * s10q}, {1, 2, 1, -3 * s10q}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}},
            {{2, 0, 2, 6 * s5q}, {0, 2, 2, -6 * s5q}, {4, 0, 0, -s5q}, {0, 4, 0, s5q}, {0, 0, 0, 0}, {0, 0, 0, 0}},
            {{3, 0, 1, s70q}, {1, 2, 1, -3 * s70q}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}},
            {{4, 0, 0, s35e}, {2, 2, 0, -6 * s35e}, {0, 4, 0, s35e}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}}};
        const int nrow = 2 * l + 1;
        for (int m = 0; m < nrow; ++m) {
          double vv = 0, dx = 0, dy = 0, dz = 0;
          const int nt = (l == 3) ? 3 : 6;
          for (int t = 0; t < nt; ++t) {
            const T tt = (l == 3) ? f3[m][t] : g4[m][t];
            if (tt.f == 0.) continue;
            double v1, d1 = 0, d2 = 0, d3 = 0;
            mono(tt.a, tt.b, tt.c, v1, d1, d2, d3);
            vv += tt.f * v1; dx += tt.f * d1; dy += tt.f * d2; dz += tt.f * d3;
          }
          store<GRAD>(B, ms, bf + m, i, ok, vv, dx, dy, dz);
        }
      }
    }
  }
}

}  // namespace

void launch_collocation(const PlanView& pv, const DevTile* tiles, int ntiles, double* ws,
                        bool gradient, cudaStream_t s) {
  if (ntiles <= 0) return;
  if (gradient) collocation_kernel<true><<<ntiles, TP, 0, s>>>(pv, tiles, ws);
  else collocation_kernel<false><<<ntiles, TP, 0, s>>>(pv, tiles, ws);
}

}  // namespace gxb
