// Stage plan of the superpoint transformer encoder: input_proj, the num_layers x (self-attention + LayerNorm, FFN +
// LayerNorm) blocks and the final head of UniDet3DEncoder.forward (reference unidet3d/encoder.py:203-239, heads
// :165-201) issued by ONE C call -- 4 + 7 num_layers launches through ud3d_gemm_fwd / ud3d_attention_fwd_tc /
// ud3d_layernorm_split with the operand-form dataflow of unidet3d_b200/encoder.py:forward_packed (every GEMM input is
// written once, in tensor-core tile form, by its producer's epilogue).  Host code only: it exists to take ~1 ms of
// per-batch launch marshalling off the Python thread (see unet_plan.cu).
#include "common.cuh"

using namespace ud3d;

namespace {

struct Bump {
  uint8_t* base;
  size_t top = 0;
  float* take(size_t rows, size_t cols) {
    float* p = (float*)(base + top);
    top += (rows * cols * 4 + 255) & ~(size_t)255;
    return p;
  }
};

struct Bufs {
  float *t_s, *H[2], *Hs[2], *qkv_s, *A_s, *Z, *F_s, *nq_s, *h_s, *scratch;
};

Bufs carve(Bump& b, const ud3d_encoder_plan* p, size_t n) {
  Bufs f;
  const size_t d = (size_t)p->d_model;
  f.t_s = b.take(n, d);
  f.H[0] = b.take(n, d); f.H[1] = b.take(n, d);
  f.Hs[0] = b.take(n, d); f.Hs[1] = b.take(n, d);
  f.qkv_s = b.take(n, 3 * d);
  f.A_s = b.take(n, d);
  f.Z = b.take(n, d);
  f.F_s = b.take(n, (size_t)p->hidden);
  f.nq_s = b.take(n, d);
  f.h_s = b.take(n, d);
  f.scratch = b.take(n, 3 * d > (size_t)p->hidden ? 3 * d : (size_t)p->hidden);   // `out` of the GEMMs whose fp32 result is not stored
  return f;
}

// out = act(in @ W^T + bias) (+ residual); `in` fp32 (in_split 0) or operand form (1); fp32 result to `raw` unless NULL,
// operand-form copy (no affine, no ReLU) to `split` unless NULL
int linear(const ud3d_linear& L, const float* in, int in_split, int c_in, int c_out, int n, int act, const float* residual, float* raw,
           float* split, float* scratch, void* stream) {
  ud3d_gemm_args a = {};
  a.in = in; a.ld_in = c_in; a.c_in = c_in;
  a.K = 1; a.n_out = n;
  a.w_packed = L.w; a.bias = L.bias; a.act = act;
  a.out = raw ? raw : scratch; a.ld_out = c_out; a.c_out = c_out;
  a.no_raw = raw ? 0 : 1;
  a.residual = residual; a.ld_res = c_out;
  a.in_split = in_split;
  if (split) {
    a.out_act[0] = split; a.ld_act[0] = c_out;
    a.act_norelu = 1;
  }
  return ud3d_gemm_fwd(&a, stream);
}

}  // namespace

extern "C" {

size_t ud3d_encoder_workspace_bytes(const ud3d_encoder_plan* plan, int n) {
  if (!plan || n <= 0) return 256;
  Bump b{nullptr};
  carve(b, plan, (size_t)n);
  return b.top + 256;
}

int ud3d_encoder_forward(const ud3d_encoder_plan* plan, const float* x, int n, const int32_t* cu_seqlens, int B, int max_T,
                         float* logits, float* raw_boxes, float* h_out, void* ws, size_t ws_bytes, void* stream) {
  UD3D_CHECK_ARG(plan && x && cu_seqlens && logits && raw_boxes && ws, "ud3d_encoder_forward: NULL argument");
  UD3D_CHECK_ARG(n >= 0 && B > 0 && max_T >= 0, "ud3d_encoder_forward: bad sizes");
  UD3D_CHECK_ARG(plan->num_layers >= 0 && plan->num_layers <= UD3D_ENCODER_MAX_LAYERS, "ud3d_encoder_forward: at most %d layers",
                 UD3D_ENCODER_MAX_LAYERS);
  UD3D_CHECK_ARG((plan->d_model == 128 || plan->d_model == 256) && plan->d_model == 32 * plan->num_heads && plan->hidden % 32 == 0 &&
                     plan->in_channels > 0 && plan->n_union > 0,
                 "ud3d_encoder_forward: needs d_model in {128, 256} = 32 x heads and hidden %% 32 == 0");
  UD3D_CHECK_ARG(((uintptr_t)ws & 255) == 0, "ud3d_encoder_forward: workspace must be 256-byte aligned");
  if (ws_bytes < ud3d_encoder_workspace_bytes(plan, n)) {
    set_error("ud3d_encoder_forward: workspace too small");
    return UD3D_EWORKSPACE;
  }
  if (n == 0) return UD3D_OK;
  Bump bump{(uint8_t*)ws};
  Bufs f = carve(bump, plan, (size_t)n);
  const int d = plan->d_model, hid = plan->hidden;
  int rc;
  // input_proj: Linear(in, d) + ReLU + Linear(d, d)   (encoder.py:138-140)
  if ((rc = linear(plan->ip0, x, 0, plan->in_channels, d, n, 1, nullptr, nullptr, f.t_s, f.scratch, stream))) return rc;
  int cur = 0;
  if ((rc = linear(plan->ip2, f.t_s, 1, d, d, n, 0, nullptr, f.H[cur], f.Hs[cur], nullptr, stream))) return rc;
  for (int l = 0; l < plan->num_layers; ++l) {
    const ud3d_encoder_layer& L = plan->layer[l];
    // SelfAttentionLayer (encoder.py:24-41): in_proj -> attention -> out_proj + residual -> LayerNorm
    if ((rc = linear(L.qkv, f.Hs[cur], 1, d, 3 * d, n, 0, nullptr, nullptr, f.qkv_s, f.scratch, stream))) return rc;
    if ((rc = ud3d_attention_fwd_tc(f.qkv_s, cu_seqlens, B, max_T, n, plan->num_heads, f.A_s, stream))) return rc;
    if ((rc = linear(L.out, f.A_s, 1, d, d, n, 0, f.H[cur], f.Z, nullptr, nullptr, stream))) return rc;
    if ((rc = ud3d_layernorm_split(f.Z, nullptr, L.n1_gamma, L.n1_beta, f.H[cur ^ 1], f.Hs[cur ^ 1], n, d, L.n1_eps, stream))) return rc;
    cur ^= 1;
    // FFN (encoder.py:63-80): Linear + activation + Linear + residual -> LayerNorm
    if ((rc = linear(L.f1, f.Hs[cur], 1, d, hid, n, plan->activation, nullptr, nullptr, f.F_s, f.scratch, stream))) return rc;
    if ((rc = linear(L.f2, f.F_s, 1, hid, d, n, 0, f.H[cur], f.Z, nullptr, nullptr, stream))) return rc;
    const bool last = l == plan->num_layers - 1;
    float* dst = (last && h_out) ? h_out : f.H[cur ^ 1];
    if ((rc = ud3d_layernorm_split(f.Z, nullptr, L.n2_gamma, L.n2_beta, dst, f.Hs[cur ^ 1], n, d, L.n2_eps, stream))) return rc;
    cur ^= 1;
    if (last && h_out) f.H[cur] = h_out;
  }
  if (plan->num_layers == 0 && h_out) UD3D_CUDA(cudaMemcpyAsync(h_out, f.H[cur], (size_t)n * d * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  // head of the last layer (encoder.py:165-201): out_norm -> cls MLP (union of classes), box Linear
  if ((rc = ud3d_layernorm_split(f.H[cur], nullptr, plan->on_gamma, plan->on_beta, nullptr, f.nq_s, n, d, plan->on_eps, stream))) return rc;
  if ((rc = linear(plan->c0, f.nq_s, 1, d, d, n, 1, nullptr, nullptr, f.h_s, f.scratch, stream))) return rc;
  if ((rc = linear(plan->c2, f.h_s, 1, d, plan->n_union, n, 0, nullptr, logits, nullptr, nullptr, stream))) return rc;
  return linear(plan->bb, f.nq_s, 1, d, 8, n, 0, nullptr, raw_boxes, nullptr, nullptr, stream);
}

}  // extern "C"
