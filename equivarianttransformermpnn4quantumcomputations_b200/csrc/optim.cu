// Optimizer-side step of the reference training loops as multi-tensor kernels (train_oc20v2_parallel.py:95-126,177-186:
// loss.backward(); clip_grad_norm_(params, grad_clip); AdamW.step(); EMA.update()).  The model's ~800 parameter tensors
// (82.5 M floats for the OC20 config) are described by ONE device table; a launch covers all of them in chunks of
// OPT_CHUNK elements, so the whole update is three launches instead of one multi_tensor_apply pass per operation:
//
//   grad_sqnorm (2 launches): per-chunk sums of squares -> deterministic ordered reduction -> total L2 norm and the
//                             clip coefficient  min(1, max_norm / (norm + 1e-6))          (torch.nn.utils.clip_grad_norm_)
//   adamw_ema   (1 launch)  : g *= clip;  p *= 1 - lr wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//                             p -= (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps)             (torch.optim.AdamW, decoupled decay)
//                             shadow = (1 - d) p + d shadow                               (ExponentialMovingAverage.update)
//   HBM-bound: 5 reads + 4 writes of 4 bytes per parameter (36 B) in one pass, against 16 + 12 + 12 B in three.
#include "common.cuh"

namespace {

constexpr int OPT_CHUNK = 16384;       // elements per CTA
constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS)
grad_sqnorm_partial_kernel(const eqv2_opt_tensor* __restrict__ T, const int* __restrict__ chunk_tensor,
                           const int* __restrict__ chunk_index, float* __restrict__ partial) {
  __shared__ float red[OPT_THREADS / 32];
  const eqv2_opt_tensor t = T[chunk_tensor[blockIdx.x]];
  const long long beg = (long long)chunk_index[blockIdx.x] * OPT_CHUNK;
  const long long end = beg + OPT_CHUNK < t.n ? beg + OPT_CHUNK : t.n;
  float s = 0.f;
  if (t.g != nullptr)
    for (long long i = beg + threadIdx.x; i < end; i += OPT_THREADS) {
      const float g = t.g[i];
      s = fmaf(g, g, s);
    }
  s = eqv2_warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < OPT_THREADS / 32; ++w) a += red[w];
    partial[blockIdx.x] = a;
  }
}

// one CTA: ordered (thread-strided, then tree in fixed order) sum of the partials in double precision
__global__ void __launch_bounds__(OPT_THREADS)
grad_sqnorm_final_kernel(const float* __restrict__ partial, int n, float max_norm, float* __restrict__ out /*[2]*/) {
  __shared__ double red[OPT_THREADS];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += OPT_THREADS) a += (double)partial[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = OPT_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(red[0]);
    out[0] = norm;
    float c = 1.0f;
    if (max_norm > 0.f) c = fminf(1.0f, max_norm / (norm + 1e-6f));
    out[1] = c;
  }
}

__global__ void __launch_bounds__(OPT_THREADS)
adamw_ema_kernel(const eqv2_opt_tensor* __restrict__ T, const int* __restrict__ chunk_tensor,
                 const int* __restrict__ chunk_index, const float* __restrict__ clip /*[2] or null*/, float b1, float b2,
                 float eps, int step, float ema_decay) {
  const eqv2_opt_tensor t = T[chunk_tensor[blockIdx.x]];
  if (t.g == nullptr) {       // parameter without gradient this step: untouched by AdamW, like torch -- but the
    if (t.ema != nullptr) {   // reference's EMA.update still averages every parameter that requires grad
      const long long b0 = (long long)chunk_index[blockIdx.x] * OPT_CHUNK;
      const long long e0 = b0 + OPT_CHUNK < t.n ? b0 + OPT_CHUNK : t.n;
      for (long long i = b0 + threadIdx.x; i < e0; i += OPT_THREADS)
        t.ema[i] = (1.0f - ema_decay) * t.p[i] + ema_decay * t.ema[i];
    }
    return;
  }
  // torch keeps a step count PER PARAMETER (a parameter without gradient does not advance): this tensor has taken
  // step - lag updates including this one.  Bias corrections in double, as torch computes them on the host.
  const double nstep = (double)(step - t.lag);
  const float bc1 = (float)(1.0 - pow((double)b1, nstep));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, nstep));
  const long long beg = (long long)chunk_index[blockIdx.x] * OPT_CHUNK;
  const long long end = beg + OPT_CHUNK < t.n ? beg + OPT_CHUNK : t.n;
  const float c = clip != nullptr ? clip[1] : 1.0f;
  const float decay = 1.0f - t.lr * t.wd;
  const float step_size = t.lr / bc1;
  for (long long i = beg + threadIdx.x; i < end; i += OPT_THREADS) {
    const float g = t.g[i] * c;
    float p = t.p[i] * decay;
    const float m = b1 * t.m[i] + (1.0f - b1) * g;          // torch: exp_avg.lerp_(grad, 1 - beta1)
    const float v = b2 * t.v[i] + (1.0f - b2) * g * g;      //        exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p -= step_size * (m / denom);
    t.p[i] = p;
    t.m[i] = m;
    t.v[i] = v;
    if (t.ema != nullptr) t.ema[i] = (1.0f - ema_decay) * p + ema_decay * t.ema[i];
  }
}

}  // namespace

extern "C" int eqv2_opt_chunk_elems(void) { return OPT_CHUNK; }

extern "C" int eqv2_grad_sqnorm(const eqv2_opt_tensor* tensors, const int* chunk_tensor, const int* chunk_index, int nchunks,
                                float max_norm, float* partial, float* out, void* stream) {
  if (nchunks == 0) return 0;
  EQV2_REQUIRE(tensors && chunk_tensor && chunk_index && partial && out, "eqv2_grad_sqnorm: null pointer");
  EQV2_LAUNCH(grad_sqnorm_partial_kernel, dim3((unsigned)nchunks), dim3(OPT_THREADS), 0, stream, tensors, chunk_tensor,
              chunk_index, partial);
  EQV2_CHECK_LAUNCH("eqv2_grad_sqnorm (partial)");
  EQV2_LAUNCH(grad_sqnorm_final_kernel, dim3(1), dim3(OPT_THREADS), 0, stream, partial, nchunks, max_norm, out);
  EQV2_CHECK_LAUNCH("eqv2_grad_sqnorm (final)");
  return 0;
}

extern "C" int eqv2_adamw_ema_step(const eqv2_opt_tensor* tensors, const int* chunk_tensor, const int* chunk_index,
                                   int nchunks, const float* clip, float beta1, float beta2, float eps, int step,
                                   float ema_decay, void* stream) {
  if (nchunks == 0) return 0;
  EQV2_REQUIRE(tensors && chunk_tensor && chunk_index && step >= 1, "eqv2_adamw_ema_step: bad arguments");
  EQV2_LAUNCH(adamw_ema_kernel, dim3((unsigned)nchunks), dim3(OPT_THREADS), 0, stream, tensors, chunk_tensor, chunk_index,
              clip, beta1, beta2, eps, step, ema_decay);
  EQV2_CHECK_LAUNCH("eqv2_adamw_ema_step");
  return 0;
}
