// Test-only host build of unidet3d_b200/csrc/box_loss.cuh (the header the GPU criterion-gradient kernel instantiates):
// lets the CPU tests check the dual-number derivatives without a GPU.  Not part of the product library.
#include "../../unidet3d_b200/csrc/box_loss.cuh"

using namespace ud3d::bl;

extern "C" {

void bl_pair_loss_grad_f32(const float* pred, const float* tgt, int n, int dim, float* loss, float* grad) {
  for (int i = 0; i < n; ++i) loss[i] = pair_loss_grad<float>(pred + (long)i * dim, tgt + (long)i * dim, dim, grad + (long)i * dim);
}

void bl_pair_loss_grad_f64(const double* pred, const double* tgt, int n, int dim, double* loss, double* grad) {
  for (int i = 0; i < n; ++i) loss[i] = pair_loss_grad<double>(pred + (long)i * dim, tgt + (long)i * dim, dim, grad + (long)i * dim);
}

// the same templates instantiated with a plain double: the function the finite differences are taken of
void bl_pair_loss_f64(const double* pred, const double* tgt, int n, int dim, double* loss) {
  for (int i = 0; i < n; ++i) {
    const double* p = pred + (long)i * dim;
    const double* t = tgt + (long)i * dim;
    loss[i] = dim == 7 ? diou_rotated<double>(p, t) : diou_aligned<double>(p, t);
  }
}

void bl_bbox_decode_backward_f32(const float* raw, int n, int with_angle, const float* d_box, float* d_raw) {
  const int dim = with_angle ? 7 : 6;
  for (int i = 0; i < n; ++i) bbox_decode_backward<float>(raw + (long)i * 8, with_angle != 0, d_box + (long)i * dim, d_raw + (long)i * 8);
}

void bl_bbox_decode_f32(const float* raw, const float* centers, int n, int with_angle, float* box) {
  const int dim = with_angle ? 7 : 6;
  for (int i = 0; i < n; ++i) bbox_decode<float, float>(raw + (long)i * 8, centers + (long)i * 3, with_angle != 0, box + (long)i * dim);
}
}
