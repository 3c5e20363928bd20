// Rotated-rectangle intersection (BEV) shared by the NMS kernels (post.cu) and the training-side criterion
// (criterion.cu).  Restates mmcv's iou3d `box_overlap` (vertex collection + angular sort + shoelace); `margin` is its
// corner-in-box tolerance: 1e-2 in mmcv's NMS kernels (over-estimates the area by up to ~1 %), 1e-6 for the exact
// intersection of mmcv's diff_iou_rotated that the rotated DIoU loss is built on.
#pragma once
#include "common.cuh"

namespace ud3d {

struct P2 {
  float x, y;
};
__device__ __forceinline__ float cross3(P2 p1, P2 p2, P2 p0) { return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y); }
__device__ __forceinline__ int check_rect_cross(P2 p1, P2 p2, P2 q1, P2 q2) {
  return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
         fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}
__device__ __forceinline__ int check_in_box2d(const float* box, P2 p, float MARGIN) {
  float cx = box[0], cy = box[1];
  float ac = cosf(-box[6]), as = sinf(-box[6]);
  float rx = (p.x - cx) * ac + (p.y - cy) * (-as);
  float ry = (p.x - cx) * as + (p.y - cy) * ac;
  return (fabsf(rx) < box[3] / 2 + MARGIN && fabsf(ry) < box[4] / 2 + MARGIN);
}
__device__ __forceinline__ int seg_intersection(P2 p1, P2 p0, P2 q1, P2 q0, P2& ans) {
  const float EPS = 1e-8f;
  if (!check_rect_cross(p0, p1, q0, q1)) return 0;
  float s1 = cross3(q0, p1, p0);
  float s2 = cross3(p1, q1, p0);
  float s3 = cross3(p0, q1, q0);
  float s4 = cross3(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
  float s5 = cross3(q1, p1, p0);
  if (fabsf(s5 - s1) > EPS) {
    ans.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
    ans.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
  } else {
    float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    float D = a0 * b1 - a1 * b0;
    ans.x = (b0 * c1 - b1 * c0) / D;
    ans.y = (a1 * c0 - a0 * c1) / D;
  }
  return 1;
}
__device__ __forceinline__ void box_corners(const float* box, P2* c) {
  float hx = box[3] / 2, hy = box[4] / 2;
  float x1 = box[0] - hx, y1 = box[1] - hy, x2 = box[0] + hx, y2 = box[1] + hy;
  float ac = cosf(box[6]), as = sinf(box[6]);
  P2 raw[4] = {{x1, y1}, {x2, y1}, {x2, y2}, {x1, y2}};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float nx = (raw[k].x - box[0]) * ac + (raw[k].y - box[1]) * (-as) + box[0];
    float ny = (raw[k].x - box[0]) * as + (raw[k].y - box[1]) * ac + box[1];
    c[k].x = nx;
    c[k].y = ny;
  }
  c[4] = c[0];
}
static __device__ float box_overlap_rot(const float* a, const float* b, float margin = 1e-2f) {
  P2 ca[5], cb[5];
  box_corners(a, ca);
  box_corners(b, cb);
  P2 pts[24];
  float ang[24];
  P2 center = {0.f, 0.f};
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      P2 ans;
      if (seg_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], ans)) {
        center.x += ans.x;
        center.y += ans.y;
        pts[cnt++] = ans;
      }
    }
  for (int k = 0; k < 4; ++k) {
    if (check_in_box2d(a, cb[k], margin)) {
      center.x += cb[k].x;
      center.y += cb[k].y;
      pts[cnt++] = cb[k];
    }
    if (check_in_box2d(b, ca[k], margin)) {
      center.x += ca[k].x;
      center.y += ca[k].y;
      pts[cnt++] = ca[k];
    }
  }
  if (cnt == 0) return 0.f;
  center.x /= cnt;
  center.y /= cnt;
  for (int i = 0; i < cnt; ++i) ang[i] = atan2f(pts[i].y - center.y, pts[i].x - center.x);
  // bubble sort ascending by angle (stable), as in the reference kernel
  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (ang[i] > ang[i + 1]) {
        P2 tp = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = tp;
        float ta = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = ta;
      }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k) {
    float ax = pts[k].x - pts[0].x, ay = pts[k].y - pts[0].y;
    float bx = pts[k + 1].x - pts[0].x, by = pts[k + 1].y - pts[0].y;
    area += ax * by - ay * bx;
  }
  return fabsf(area) / 2.0f;
}

}  // namespace ud3d
