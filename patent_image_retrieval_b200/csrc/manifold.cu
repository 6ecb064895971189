// Small manifold kernels of the train_hyp step (SURVEY 8f-4 and the per-pair loops of 8a-5): everything here is
// row-local arithmetic over <= tens of thousands of label / figure rows -- latency-sized, one warp per row or pair,
// values and analytic gradients on the device so that a training step never leaves it.
//
//   rowpair_dist         d_t = pmath.dist(x[ia_t], y[ib_t])                    /root/reference/src/models.py:712-719,
//                        (the per-pair Python loops, batched)                  824-829; src/train.py:1036,1433-1443
//   hmi_pairs            _hmi_insideness / _hmi_disjointedness + relu margin   src/models.py:550-604, 630-674
//   dist0_reg            relu(lo - dist0) + relu(dist0 - hi), mean             src/models.py:606-628
//   radam_ball_step      geoopt RiemannianAdam on a ManifoldParameter          src/train.py:1362 (optimizer.step())
//
// Forward kernels write per-item values (and a loss sum); backward kernels recompute the item and scatter
// g * d(value)/d(row) into the gradient rows with atomicAdd (label rows are shared between pairs).
// fp32 rows, fp64 accumulation of the reductions (the rows are short; accuracy over speed).
#include <math.h>

#include "common.cuh"

namespace {

constexpr int MF_WARPS = 4;
constexpr double MIN_NORM = 1e-15;

__device__ __forceinline__ double row_dot(const float* __restrict__ a, const float* __restrict__ b, int d, int lane) {
  double s = 0.0;
  for (int j = lane; j < d; j += 32) s += (double)a[j] * (double)b[j];
  return warp_sum(s);
}

// geoopt project(): rows with ||x|| > (1 - eps) / sqrt(c) are scaled back onto that sphere.  Returns the scale (1 when
// the row is inside) for a row of squared norm xsq.
__device__ __forceinline__ double project_scale(double xsq, double c, double eps = 4e-3) {
  const double maxnorm = (1.0 - eps) / sqrt(c);
  const double n = fmax(sqrt(xsq), MIN_NORM);
  return n > maxnorm ? maxnorm / n : 1.0;
}

// ------------------------------------------------------------------------------------------------ row-pair distance
// one warp per pair: s = |x - y|^2 from explicit differences, d = arccosh(1 + 2 c s / (alpha beta)) / sqrt(c)
__global__ void __launch_bounds__(MF_WARPS * 32)
rowpair_dist_kernel(const float* __restrict__ x, const float* __restrict__ y, const int64_t* __restrict__ ia,
                    const int64_t* __restrict__ ib, int64_t n_pairs, int d, float c, float* __restrict__ out,
                    const float* __restrict__ grad_out, float* __restrict__ gx, float* __restrict__ gy) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * MF_WARPS + (threadIdx.x >> 5);
  if (t >= n_pairs) return;
  const float* a = x + ia[t] * d;
  const float* p = y + ib[t] * d;
  double s = 0.0, aa = 0.0, pp = 0.0;
  for (int j = lane; j < d; j += 32) {
    const double av = a[j], pv = p[j], e = av - pv;
    s += e * e; aa += av * av; pp += pv * pv;
  }
  s = warp_sum(s); aa = warp_sum(aa); pp = warp_sum(pp);
  const double cc = c, al = 1.0 - cc * aa, be = 1.0 - cc * pp;
  const double t0 = 2.0 * cc * s / (al * be);
  const double q = t0 * (t0 + 2.0);                      // z^2 - 1
  if (out != nullptr && lane == 0) out[t] = (float)(log1p(t0 + sqrt(q)) / sqrt(cc));
  if (grad_out != nullptr) {
    // dd/da = w [(1 + c s / alpha) a - p],  w = 4 sqrt(c) / (alpha beta sqrt(z^2 - 1))   
    const double w = (double)grad_out[t] * 4.0 * sqrt(cc) / (al * be * sqrt(fmax(q, 1e-30)));
    const double fa = w * (1.0 + cc * s / al), fp = w * (1.0 + cc * s / be);
    for (int j = lane; j < d; j += 32) {
      const double av = a[j], pv = p[j];
      if (gx != nullptr) atomicAdd(gx + ia[t] * d + j, (float)(fa * av - w * pv));
      if (gy != nullptr) atomicAdd(gy + ib[t] * d + j, (float)(fp * pv - w * av));
    }
  }
}

// ------------------------------------------------------------------------------------------------ HMI pair losses
// Per pair (a, b) of label rows (after projx): with n = |x|, r(n) = (1 - c n^2) / (2 sqrt(c) n), centre(x) = x phi(n),
// phi(n) = 1 + r sqrt(c) / n = 1 - c/2 + 1 / (2 n^2):
//   insideness     v = (r_b - r_a) - |centre_a - centre_b|         (mode 0)
//   disjointedness v = |centre_a - centre_b| - (r_a + r_b)         (mode 1)
// loss = mean_t relu(margin - v_t).  Backward: dr/dn = -(1/n^2 + c) / (2 sqrt(c)), dphi/dn = -1/n^3.
__global__ void __launch_bounds__(MF_WARPS * 32)
hmi_pairs_kernel(const float* __restrict__ emb, const int64_t* __restrict__ pairs, int64_t n_pairs, int d, float c,
                 int mode, float margin, float proj_eps, float* __restrict__ values, double* __restrict__ loss_sum,
                 const float* __restrict__ grad_scale, float* __restrict__ grad_emb) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * MF_WARPS + (threadIdx.x >> 5);
  if (t >= n_pairs) return;
  const int64_t ia = pairs[2 * t], ib = pairs[2 * t + 1];
  const float* a = emb + ia * d;
  const float* b = emb + ib * d;
  const double cc = c, sc = sqrt(cc);
  double aa = row_dot(a, a, d, lane), bb = row_dot(b, b, d, lane);
  const double pa = project_scale(aa, cc, proj_eps), pb = project_scale(bb, cc, proj_eps);      // projx
  const double na = fmax(sqrt(aa) * pa, MIN_NORM), nb = fmax(sqrt(bb) * pb, MIN_NORM);
  const double ra = (1.0 - cc * na * na) / (2.0 * sc * na), rb = (1.0 - cc * nb * nb) / (2.0 * sc * nb);
  const double pha = 1.0 + ra * sc / na, phb = 1.0 + rb * sc / nb;
  double cd2 = 0.0;
  for (int j = lane; j < d; j += 32) {
    const double e = pha * pa * a[j] - phb * pb * b[j];
    cd2 += e * e;
  }
  const double cd = sqrt(warp_sum(cd2));
  const double v = mode == 0 ? (rb - ra) - cd : cd - (ra + rb);
  const double hinge = (double)margin - v;
  if (values != nullptr && lane == 0) values[t] = (float)v;
  if (loss_sum != nullptr && lane == 0 && hinge > 0.0) atomicAdd(loss_sum, hinge);
  if (grad_emb != nullptr && hinge > 0.0) {
    // d loss / d v = -gs / n_pairs on the active side of the hinge
    const double g = -(double)(grad_scale != nullptr ? *grad_scale : 1.f) / (double)n_pairs;
    // v = sa r_a + sb r_b + sd cd
    const double sa = mode == 0 ? -1.0 : -1.0, sb = mode == 0 ? 1.0 : -1.0, sd = mode == 0 ? -1.0 : 1.0;
    const double dra = -(1.0 / (na * na) + cc) / (2.0 * sc), drb = -(1.0 / (nb * nb) + cc) / (2.0 * sc);
    const double dpa = -1.0 / (na * na * na), dpb = -1.0 / (nb * nb * nb);
    // u = (centre_a - centre_b) / cd; a.u and b.u (projected points)
    double au = 0.0, bu = 0.0;
    const double inv_cd = cd > 0.0 ? 1.0 / cd : 0.0;
    for (int j = lane; j < d; j += 32) {
      const double xa = pa * a[j], xb = pb * b[j];
      const double u = (pha * xa - phb * xb) * inv_cd;
      au += xa * u; bu += xb * u;
    }
    au = warp_sum(au); bu = warp_sum(bu);
    // gradient w.r.t. the PROJECTED points, then through projx (clipped rows: (maxnorm/|x|)(I - x^ x^T))
    double ga_dot = 0.0, gb_dot = 0.0;     // <grad_projected, x^> for the clip Jacobian
    for (int pass = 0; pass < 2; ++pass) {
      for (int j = lane; j < d; j += 32) {
        const double xa = pa * a[j], xb = pb * b[j];
        const double u = (pha * xa - phb * xb) * inv_cd;
        const double gpa = g * (sa * dra * xa / na + sd * (pha * u + dpa / na * xa * au));
        const double gpb = g * (sb * drb * xb / nb - sd * (phb * u + dpb / nb * xb * bu));
        if (pass == 0) {
          ga_dot += gpa * xa / na; gb_dot += gpb * xb / nb;
        } else {
          const double oa = pa < 1.0 ? pa * (gpa - xa / na * ga_dot) : gpa;
          const double ob = pb < 1.0 ? pb * (gpb - xb / nb * gb_dot) : gpb;
          atomicAdd(grad_emb + ia * d + j, (float)oa);
          atomicAdd(grad_emb + ib * d + j, (float)ob);
        }
      }
      if (pass == 0) { ga_dot = warp_sum(ga_dot); gb_dot = warp_sum(gb_dot); }
    }
  }
}

// ------------------------------------------------------------------------------------------------ dist0 regulariser
// per row: d0 = 2 artanh(sqrt(c) |x|) / sqrt(c) (clamped like geoopt), value = relu(lo - d0) + relu(d0 - hi);
// loss = mean over rows.  lo < 0: no lower hinge (the figure-embedding regulariser).
__global__ void __launch_bounds__(MF_WARPS * 32)
dist0_reg_kernel(const float* __restrict__ x, int64_t n, int d, float c, float lo, float hi, double* __restrict__ loss_sum,
                 const float* __restrict__ grad_scale, float* __restrict__ grad_x) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * MF_WARPS + (threadIdx.x >> 5);
  if (i >= n) return;
  const float* r = x + i * d;
  const double cc = c, sc = sqrt(cc);
  const double nn = fmax(sqrt(row_dot(r, r, d, lane)), MIN_NORM);
  const double arg = fmin(fmax(sc * nn, -1.0 + 1e-7), 1.0 - 1e-7);           // geoopt artanh clamp (fp32 eps)
  const double d0 = fmax(2.0 * atanh(arg) / sc, MIN_NORM);
  double val = 0.0, sgn = 0.0;
  if (lo >= 0.f && (double)lo - d0 > 0.0) { val += (double)lo - d0; sgn -= 1.0; }
  if (d0 - (double)hi > 0.0) { val += d0 - (double)hi; sgn += 1.0; }
  if (loss_sum != nullptr && lane == 0 && val > 0.0) atomicAdd(loss_sum, val);
  if (grad_x != nullptr && sgn != 0.0 && sc * nn < 1.0 - 1e-7) {
    // d d0 / dx = 2 / (1 - c |x|^2) * x / |x|
    const double g = (double)(grad_scale != nullptr ? *grad_scale : 1.f) / (double)n * sgn * 2.0 /
                     ((1.0 - cc * nn * nn) * nn);
    for (int j = lane; j < d; j += 32) grad_x[i * d + j] += (float)(g * r[j]);
  }
}

// ------------------------------------------------------------------------------------------------ RiemannianAdam
// One step of geoopt.optim.RiemannianAdam for a [n, d] ManifoldParameter on the Poincare ball (row-wise manifold):
//   g   = (grad + wd x) / lambda_x^2                                  egrad2rgrad
//   m   = b1 m + (1 - b1) g ;  v = b2 v + (1 - b2) lambda_x^2 <g, g>   (v: one value per row, stored broadcast)
//   dir = (m / bc1) / (sqrt(v / bc2) + eps)
//   x'  = project(x - lr dir) ;  m' = gyr[x', -x] m * lambda_x / lambda_x'   (retraction + parallel transport)
// exp_avg_sq keeps geoopt's [n, d] shape (every column of a row holds the row's value).
__global__ void __launch_bounds__(MF_WARPS * 32)
radam_ball_kernel(float* __restrict__ x, const float* __restrict__ grad, float* __restrict__ exp_avg,
                  float* __restrict__ exp_avg_sq, int64_t n, int d, float c, float lr, float b1, float b2, float eps,
                  float wd, float bc1, float bc2) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * MF_WARPS + (threadIdx.x >> 5);
  if (i >= n) return;
  float* xr = x + i * d;
  const float* gr = grad + i * d;
  float* mr = exp_avg + i * d;
  float* vr = exp_avg_sq + i * d;
  const double cc = c, k = -cc;
  const double x2 = row_dot(xr, xr, d, lane);
  const double lam = 2.0 / fmax(1.0 - cc * x2, MIN_NORM);
  const double inv_l2 = 1.0 / (lam * lam);
  double gg = 0.0;
  for (int j = lane; j < d; j += 32) {
    const double g = ((double)gr[j] + (double)wd * xr[j]) * inv_l2;
    gg += g * g;
  }
  gg = warp_sum(gg);
  const double v_new = (double)b2 * vr[0] + (1.0 - (double)b2) * lam * lam * gg;      // row value (all columns equal)
  const double denom = sqrt(v_new / (double)bc2) + (double)eps;
  // first pass: new moment (before transport), step, new point's norm and the dots the gyration needs
  double y2 = 0.0, xy = 0.0, xm = 0.0, ym = 0.0;
  __syncwarp();
  for (int j = lane; j < d; j += 32) {
    const double g = ((double)gr[j] + (double)wd * xr[j]) * inv_l2;
    const double m = (double)b1 * mr[j] + (1.0 - (double)b1) * g;
    const double yv = (double)xr[j] - (double)lr * (m / (double)bc1) / denom;
    y2 += yv * yv;
  }
  y2 = warp_sum(y2);
  const double ps = project_scale(y2, cc);
  y2 *= ps * ps;
  for (int j = lane; j < d; j += 32) {
    const double g = ((double)gr[j] + (double)wd * xr[j]) * inv_l2;
    const double m = (double)b1 * mr[j] + (1.0 - (double)b1) * g;
    const double yv = ps * ((double)xr[j] - (double)lr * (m / (double)bc1) / denom);
    xy += yv * (-(double)xr[j]);       // a = y (new point), b = -x
    xm += (-(double)xr[j]) * m;        // <b, u>
    ym += yv * m;                      // <a, u>
  }
  xy = warp_sum(xy); xm = warp_sum(xm); ym = warp_sum(ym);
  // gyration(a = y, b = -x, u = m): u + 2 (A a + B b) / D
  const double a2 = y2, b2v = x2, ab = xy, au = ym, bu = xm, K2 = k * k;
  const double A = -K2 * au * b2v - k * bu + 2.0 * K2 * ab * bu;
  const double B = -K2 * bu * a2 + k * au;
  const double Dn = fmax(1.0 - 2.0 * k * ab + K2 * a2 * b2v, MIN_NORM);
  const double lam_y = 2.0 / fmax(1.0 - cc * y2, MIN_NORM);
  const double ratio = lam / lam_y;
  for (int j = lane; j < d; j += 32) {
    const double xo = xr[j];
    const double g = ((double)gr[j] + (double)wd * xo) * inv_l2;
    const double m = (double)b1 * mr[j] + (1.0 - (double)b1) * g;
    const double yv = ps * (xo - (double)lr * (m / (double)bc1) / denom);
    const double mt = (m + 2.0 * (A * yv + B * (-xo)) / Dn) * ratio;
    xr[j] = (float)yv;
    mr[j] = (float)mt;
    vr[j] = (float)v_new;
  }
}

}  // namespace

int hypret_launch_rowpair_dist(const float* x, const float* y, const int64_t* ia, const int64_t* ib, int64_t n_pairs,
                               int d, float c, float* out, const float* grad_out, float* gx, float* gy,
                               cudaStream_t stream) {
  if (n_pairs == 0) return HYPRET_OK;
  rowpair_dist_kernel<<<(unsigned)((n_pairs + MF_WARPS - 1) / MF_WARPS), MF_WARPS * 32, 0, stream>>>(
      x, y, ia, ib, n_pairs, d, c, out, grad_out, gx, gy);
  return (int)cudaGetLastError();
}

int hypret_launch_hmi_pairs(const float* emb, const int64_t* pairs, int64_t n_pairs, int d, float c, int mode,
                            float margin, float proj_eps, float* values, double* loss_sum, const float* grad_scale,
                            float* grad_emb, cudaStream_t stream) {
  if (n_pairs == 0) return HYPRET_OK;
  hmi_pairs_kernel<<<(unsigned)((n_pairs + MF_WARPS - 1) / MF_WARPS), MF_WARPS * 32, 0, stream>>>(
      emb, pairs, n_pairs, d, c, mode, margin, proj_eps, values, loss_sum, grad_scale, grad_emb);
  return (int)cudaGetLastError();
}

int hypret_launch_dist0_reg(const float* x, int64_t n, int d, float c, float lo, float hi, double* loss_sum,
                            const float* grad_scale, float* grad_x, cudaStream_t stream) {
  if (n == 0) return HYPRET_OK;
  dist0_reg_kernel<<<(unsigned)((n + MF_WARPS - 1) / MF_WARPS), MF_WARPS * 32, 0, stream>>>(
      x, n, d, c, lo, hi, loss_sum, grad_scale, grad_x);
  return (int)cudaGetLastError();
}

int hypret_launch_radam_ball(float* x, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int d, float c,
                             float lr, float b1, float b2, float eps, float wd, float bc1, float bc2,
                             cudaStream_t stream) {
  if (n == 0) return HYPRET_OK;
  radam_ball_kernel<<<(unsigned)((n + MF_WARPS - 1) / MF_WARPS), MF_WARPS * 32, 0, stream>>>(
      x, grad, exp_avg, exp_avg_sq, n, d, c, lr, b1, b2, eps, wd, bc1, bc2);
  return (int)cudaGetLastError();
}
