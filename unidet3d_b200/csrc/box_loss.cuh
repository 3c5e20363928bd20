// DIoU box losses of the criterion as templates over the scalar type, for their GRADIENTS:
//   reference: unidet3d/axis_aligned_iou_loss.py:14-53 (axis-aligned DIoU on corner boxes, criterion.py:180-198),
//              unidet3d/rotated_iou_loss.py:14-82 (rotated DIoU: mmcv box2corners + oriented_box_intersection_2d under
//              torch.autograd).
// Instantiated with `Dual<F, N>` (forward-mode derivative w.r.t. the N parameters of the predicted box) the same code
// that evaluates the loss also yields its gradient: the intersection polygon's vertices are edge-edge intersections
// or corners, the shoelace area is a smooth function of them, and every branch (which vertices exist, their angular
// order, min / max / clamp selections) is taken on the VALUES -- exactly what autograd does through mmcv's
// sort_vertices.  A few hundred matched pairs per (head, scene): latency-bound scalar work, no tensor-core shape.
//
// The header is self-contained (math.h only) and compiles as plain C++ too: the CPU tests build it with g++ and check
// the derivatives against finite differences in double and against torch.autograd (tests/test_box_loss_host.py).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define UD3D_BL_HD __host__ __device__ __forceinline__
#define UD3D_BL_HDN __host__ __device__
#else
#define UD3D_BL_HD inline
#define UD3D_BL_HDN inline
#endif

namespace ud3d {
namespace bl {

template <class F, int N>
struct Dual {
  F v;
  F d[N];
  UD3D_BL_HD Dual() {}
  UD3D_BL_HD Dual(F x) : v(x) {
    for (int i = 0; i < N; ++i) d[i] = F(0);
  }
};

// ---- value access / arithmetic for plain scalars and duals
UD3D_BL_HD float val(float x) { return x; }
UD3D_BL_HD double val(double x) { return x; }
template <class F, int N> UD3D_BL_HD F val(const Dual<F, N>& x) { return x.v; }

template <class F, int N> UD3D_BL_HD Dual<F, N> operator+(const Dual<F, N>& a, const Dual<F, N>& b) {
  Dual<F, N> r; r.v = a.v + b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <class F, int N> UD3D_BL_HD Dual<F, N> operator-(const Dual<F, N>& a, const Dual<F, N>& b) {
  Dual<F, N> r; r.v = a.v - b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <class F, int N> UD3D_BL_HD Dual<F, N> operator-(const Dual<F, N>& a) {
  Dual<F, N> r; r.v = -a.v;
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}
template <class F, int N> UD3D_BL_HD Dual<F, N> operator*(const Dual<F, N>& a, const Dual<F, N>& b) {
  Dual<F, N> r; r.v = a.v * b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <class F, int N> UD3D_BL_HD Dual<F, N> operator/(const Dual<F, N>& a, const Dual<F, N>& b) {
  Dual<F, N> r; r.v = a.v / b.v;
  const F inv = F(1) / b.v;
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
// mixed with the base scalar (constants, GT box parameters)
template <class F, int N> UD3D_BL_HD Dual<F, N> operator+(const Dual<F, N>& a, F b) { Dual<F, N> r = a; r.v += b; return r; }
template <class F, int N> UD3D_BL_HD Dual<F, N> operator+(F b, const Dual<F, N>& a) { Dual<F, N> r = a; r.v += b; return r; }
template <class F, int N> UD3D_BL_HD Dual<F, N> operator-(const Dual<F, N>& a, F b) { Dual<F, N> r = a; r.v -= b; return r; }
template <class F, int N> UD3D_BL_HD Dual<F, N> operator-(F b, const Dual<F, N>& a) { Dual<F, N> r = -a; r.v += b; return r; }
template <class F, int N> UD3D_BL_HD Dual<F, N> operator*(const Dual<F, N>& a, F b) {
  Dual<F, N> r; r.v = a.v * b;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b;
  return r;
}
template <class F, int N> UD3D_BL_HD Dual<F, N> operator*(F b, const Dual<F, N>& a) { return a * b; }
template <class F, int N> UD3D_BL_HD Dual<F, N> operator/(const Dual<F, N>& a, F b) { return a * (F(1) / b); }

UD3D_BL_HD float s_sin(float x) { return sinf(x); }
UD3D_BL_HD float s_cos(float x) { return cosf(x); }
UD3D_BL_HD double s_sin(double x) { return sin(x); }
UD3D_BL_HD double s_cos(double x) { return cos(x); }
template <class F, int N> UD3D_BL_HD Dual<F, N> s_sin(const Dual<F, N>& a) {
  Dual<F, N> r; r.v = s_sin(a.v);
  const F c = s_cos(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i];
  return r;
}
template <class F, int N> UD3D_BL_HD Dual<F, N> s_cos(const Dual<F, N>& a) {
  Dual<F, N> r; r.v = s_cos(a.v);
  const F s = -s_sin(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i];
  return r;
}
UD3D_BL_HD float s_exp(float x) { return expf(x); }
UD3D_BL_HD double s_exp(double x) { return exp(x); }
template <class F, int N> UD3D_BL_HD Dual<F, N> s_exp(const Dual<F, N>& a) {
  Dual<F, N> r; r.v = s_exp(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = r.v * a.d[i];
  return r;
}
// sqrt(x^2 + y^2) and atan2(y, x) of two scalars; at the origin the derivative is taken as 0 (torch.autograd gives NaN there)
UD3D_BL_HD float s_hypot(float x, float y) { return sqrtf(x * x + y * y); }
UD3D_BL_HD double s_hypot(double x, double y) { return sqrt(x * x + y * y); }
template <class F, int N> UD3D_BL_HD Dual<F, N> s_hypot(const Dual<F, N>& x, const Dual<F, N>& y) {
  Dual<F, N> r; r.v = s_hypot(x.v, y.v);
  const F inv = r.v > F(0) ? F(1) / r.v : F(0);
  for (int i = 0; i < N; ++i) r.d[i] = (x.v * x.d[i] + y.v * y.d[i]) * inv;
  return r;
}
UD3D_BL_HD float s_atan2(float y, float x) { return atan2f(y, x); }
UD3D_BL_HD double s_atan2(double y, double x) { return atan2(y, x); }
template <class F, int N> UD3D_BL_HD Dual<F, N> s_atan2(const Dual<F, N>& y, const Dual<F, N>& x) {
  Dual<F, N> r; r.v = s_atan2(y.v, x.v);
  const F n2 = x.v * x.v + y.v * y.v;
  const F inv = n2 > F(0) ? F(1) / n2 : F(0);
  for (int i = 0; i < N; ++i) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) * inv;
  return r;
}
// selections on the values (torch.min / torch.max / clamp / abs propagate the gradient of the selected operand)
template <class S> UD3D_BL_HD S s_min(const S& a, const S& b) { return val(b) < val(a) ? b : a; }
template <class S> UD3D_BL_HD S s_max(const S& a, const S& b) { return val(b) > val(a) ? b : a; }
template <class S> UD3D_BL_HD S s_abs(const S& a) { return val(a) < 0 ? -a : a; }

template <class S> struct base_of { typedef S type; };
template <class F, int N> struct base_of<Dual<F, N> > { typedef F type; };

template <class S>
struct Pt {
  S x, y;
};

template <class S> UD3D_BL_HD S cross3(const Pt<S>& p1, const Pt<S>& p2, const Pt<S>& p0) {
  return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

// (x, y, w, h, alpha) of a 7-parameter box -> 4 corners (+ the first repeated); mmcv box2corners / iou3d box corners
template <class S>
UD3D_BL_HD void corners_of(const S* box, Pt<S>* c) {
  typedef typename base_of<S>::type F;
  const S hx = box[3] * F(0.5), hy = box[4] * F(0.5);
  const S ac = s_cos(box[6]), as = s_sin(box[6]);
  const F sx[4] = {F(-1), F(1), F(1), F(-1)}, sy[4] = {F(-1), F(-1), F(1), F(1)};
  for (int k = 0; k < 4; ++k) {
    const S rx = hx * sx[k], ry = hy * sy[k];
    c[k].x = rx * ac - ry * as + box[0];
    c[k].y = rx * as + ry * ac + box[1];
  }
  c[4] = c[0];
}

template <class S>
UD3D_BL_HD bool in_box2d(const S* box, const Pt<S>& p, typename base_of<S>::type margin) {
  typedef typename base_of<S>::type F;
  const F cx = val(box[0]), cy = val(box[1]), a = -val(box[6]);
  const F ac = s_cos(a), as = s_sin(a);
  const F rx = (val(p.x) - cx) * ac + (val(p.y) - cy) * (-as);
  const F ry = (val(p.x) - cx) * as + (val(p.y) - cy) * ac;
  return fabs(rx) < val(box[3]) / 2 + margin && fabs(ry) < val(box[4]) / 2 + margin;
}

template <class S>
UD3D_BL_HD bool seg_intersection(const Pt<S>& p1, const Pt<S>& p0, const Pt<S>& q1, const Pt<S>& q0, Pt<S>& ans) {
  typedef typename base_of<S>::type F;
  const F p0x = val(p0.x), p0y = val(p0.y), p1x = val(p1.x), p1y = val(p1.y);
  const F q0x = val(q0.x), q0y = val(q0.y), q1x = val(q1.x), q1y = val(q1.y);
  const bool rect = fmin(p0x, p1x) <= fmax(q0x, q1x) && fmin(q0x, q1x) <= fmax(p0x, p1x) && fmin(p0y, p1y) <= fmax(q0y, q1y) &&
                    fmin(q0y, q1y) <= fmax(p0y, p1y);
  if (!rect) return false;
  const S s1 = cross3(q0, p1, p0);
  const S s2 = cross3(p1, q1, p0);
  const S s3 = cross3(p0, q1, q0);
  const S s4 = cross3(q1, p1, q0);
  if (!(val(s1) * val(s2) > 0 && val(s3) * val(s4) > 0)) return false;
  const S s5 = cross3(q1, p1, p0);
  if (fabs(val(s5) - val(s1)) > F(1e-8)) {
    const S den = s5 - s1;
    ans.x = (s5 * q0.x - s1 * q1.x) / den;
    ans.y = (s5 * q0.y - s1 * q1.y) / den;
  } else {
    const S a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    const S a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    const S D = a0 * b1 - a1 * b0;
    ans.x = (b0 * c1 - b1 * c0) / D;
    ans.y = (a1 * c0 - a0 * c1) / D;
  }
  return true;
}

// BEV intersection area of two rotated rectangles given as 7-parameter boxes (same vertex collection, angular bubble
// sort and shoelace as box_overlap_rot in boxes.cuh, which is the float-only version the matcher and the NMS use)
template <class S>
UD3D_BL_HDN S overlap_rot(const S* a, const S* b, typename base_of<S>::type margin) {
  typedef typename base_of<S>::type F;
  Pt<S> ca[5], cb[5];
  corners_of(a, ca);
  corners_of(b, cb);
  Pt<S> pts[24];
  F ang[24];
  F cxs = 0, cys = 0;
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      Pt<S> ans;
      if (seg_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], ans)) {
        cxs += val(ans.x);
        cys += val(ans.y);
        pts[cnt++] = ans;
      }
    }
  for (int k = 0; k < 4; ++k) {
    if (in_box2d(a, cb[k], margin)) {
      cxs += val(cb[k].x);
      cys += val(cb[k].y);
      pts[cnt++] = cb[k];
    }
    if (in_box2d(b, ca[k], margin)) {
      cxs += val(ca[k].x);
      cys += val(ca[k].y);
      pts[cnt++] = ca[k];
    }
  }
  if (cnt == 0) return S(F(0));
  cxs /= cnt;
  cys /= cnt;
  for (int i = 0; i < cnt; ++i) ang[i] = atan2(val(pts[i].y) - cys, val(pts[i].x) - cxs);
  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (ang[i] > ang[i + 1]) {
        const Pt<S> tp = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = tp;
        const F ta = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = ta;
      }
  S area = S(F(0));
  for (int k = 0; k < cnt - 1; ++k) {
    const S ax = pts[k].x - pts[0].x, ay = pts[k].y - pts[0].y;
    const S bx = pts[k + 1].x - pts[0].x, by = pts[k + 1].y - pts[0].y;
    area = area + (ax * by - ay * bx);
  }
  return s_abs(area) * F(0.5);
}

// rotated DIoU loss of one pair, boxes (x, y, z, w, h, l, alpha): 1 - (IoU3D - r2 / c2), rotated_iou_loss.py:14-82
template <class S>
UD3D_BL_HDN S diou_rotated(const S* p, const S* t) {
  typedef typename base_of<S>::type F;
  const S inter = overlap_rot(p, t, F(1e-6));
  const S zmax1 = p[2] + p[5] * F(0.5), zmin1 = p[2] - p[5] * F(0.5);
  const S zmax2 = t[2] + t[5] * F(0.5), zmin2 = t[2] - t[5] * F(0.5);
  const S z_ov = s_max(s_min(zmax1, zmax2) - s_max(zmin1, zmin2), S(F(0)));
  const S inter3 = inter * z_ov;
  const S union3 = p[3] * p[4] * p[5] + t[3] * t[4] * t[5] - inter3;
  Pt<S> c1[5], c2[5];
  corners_of(p, c1);
  corners_of(t, c2);
  S x_max = c1[0].x, x_min = c1[0].x, y_max = c1[0].y, y_min = c1[0].y;
  for (int k = 0; k < 4; ++k) {
    if (k) {
      x_max = s_max(x_max, c1[k].x); x_min = s_min(x_min, c1[k].x);
      y_max = s_max(y_max, c1[k].y); y_min = s_min(y_min, c1[k].y);
    }
    x_max = s_max(x_max, c2[k].x); x_min = s_min(x_min, c2[k].x);
    y_max = s_max(y_max, c2[k].y); y_min = s_min(y_min, c2[k].y);
  }
  const S z_max = s_max(zmax1, zmax2), z_min = s_min(zmin1, zmin2);
  // r2 over (x, y, w) of the BEV boxes, as the reference writes it (rotated_iou_loss.py:24-25,61)
  const S dx = p[0] - t[0], dy = p[1] - t[1], dw = p[3] - t[3];
  const S r2 = dx * dx + dy * dy + dw * dw;
  const S ex = x_min - x_max, ey = y_min - y_max, ez = z_min - z_max;
  const S cc = ex * ex + ey * ey + ez * ez;
  return F(1) - (inter3 / union3 - r2 / cc);
}

// axis-aligned DIoU loss of one pair, boxes (centre, size) -> corners (criterion.py:180-198), IoU of mmdet3d's
// AxisAlignedBboxOverlaps3D (eps 1e-6 on the union), centre-distance penalty of axis_aligned_iou_loss.py:41-52
template <class S>
UD3D_BL_HDN S diou_aligned(const S* p, const S* t) {
  typedef typename base_of<S>::type F;
  S p1[3], p2[3], t1[3], t2[3];
  for (int a = 0; a < 3; ++a) {
    p1[a] = p[a] - p[a + 3] * F(0.5); p2[a] = p[a] + p[a + 3] * F(0.5);
    t1[a] = t[a] - t[a + 3] * F(0.5); t2[a] = t[a] + t[a + 3] * F(0.5);
  }
  const S a1 = (p2[0] - p1[0]) * (p2[1] - p1[1]) * (p2[2] - p1[2]);
  const S a2 = (t2[0] - t1[0]) * (t2[1] - t1[1]) * (t2[2] - t1[2]);
  S ov = S(F(1));
  for (int a = 0; a < 3; ++a) ov = ov * s_max(s_min(p2[a], t2[a]) - s_max(p1[a], t1[a]), S(F(0)));
  const S uni = s_max(a1 + a2 - ov, S(F(1e-6)));
  S r2 = S(F(0)), c2 = S(F(0));
  for (int a = 0; a < 3; ++a) {
    const S dc = (p1[a] + p2[a]) * F(0.5) - (t1[a] + t2[a]) * F(0.5);
    r2 = r2 + dc * dc;
    const S e = s_min(p1[a], t1[a]) - s_max(p2[a], t2[a]);
    c2 = c2 + e * e;
  }
  return F(1) - ov / uni + r2 / c2;
}

// loss and gradient w.r.t. the predicted box of one matched pair; dim = 6 (axis-aligned) or 7 (rotated)
template <class F>
UD3D_BL_HDN F pair_loss_grad(const F* pred, const F* target, int dim, F* grad) {
  typedef Dual<F, 7> D;
  D p[7], t[7];
  for (int i = 0; i < dim; ++i) {
    p[i] = D(pred[i]);
    p[i].d[i] = F(1);
    t[i] = D(target[i]);
  }
  const D l = dim == 7 ? diou_rotated(p, t) : diou_aligned(p, t);
  for (int i = 0; i < dim; ++i) grad[i] = l.d[i];
  return l.v;
}

// PredBBox's exp + _bbox_pred_to_bbox (unidet3d/encoder.py:109-111,241-283): raw[8] (6 log-distances to the faces + the two
// angle logits), query centre -> box[6] (centre, size) or box[7] (x, y, z, w, l, h, alpha).  Same expressions as
// bbox_decode_kernel (encoder.cu), which produces the forward value.
template <class S, class F>
UD3D_BL_HDN void bbox_decode(const S* raw, const F* center, bool with_angle, S* box) {
  S e[6];
  for (int j = 0; j < 6; ++j) e[j] = s_exp(raw[j]);
  box[0] = center[0] + (e[1] - e[0]) * F(0.5);
  box[1] = center[1] + (e[3] - e[2]) * F(0.5);
  box[2] = center[2] + (e[5] - e[4]) * F(0.5);
  if (!with_angle) {
    box[3] = e[0] + e[1];
    box[4] = e[2] + e[3];
    box[5] = e[4] + e[5];
    return;
  }
  const S scale = e[0] + e[1] + e[2] + e[3];
  const S q = s_exp(s_hypot(raw[6], raw[7]));
  const S w = scale / (q + F(1));
  box[3] = w;
  box[4] = w * q;
  box[5] = e[5] + e[4];
  box[6] = s_atan2(raw[6], raw[7]) * F(0.5);
}

// d_raw[8] = J^T d_box for the decode above (forward-mode over the 8 raw values)
template <class F>
UD3D_BL_HDN void bbox_decode_backward(const F* raw, bool with_angle, const F* d_box, F* d_raw) {
  typedef Dual<F, 8> D;
  D r[8], box[7];
  for (int i = 0; i < 8; ++i) {
    r[i] = D(raw[i]);
    r[i].d[i] = F(1);
  }
  const F zero[3] = {F(0), F(0), F(0)};
  bbox_decode(r, zero, with_angle, box);
  const int dim = with_angle ? 7 : 6;
  for (int i = 0; i < 8; ++i) {
    F s = F(0);
    for (int j = 0; j < dim; ++j) s += d_box[j] * box[j].d[i];
    d_raw[i] = s;
  }
}

}  // namespace bl
}  // namespace ud3d
