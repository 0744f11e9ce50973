// Elementwise legacy-Keras optimizers over a flat fp32 buffer (movierec/model.py:197-204; SURVEY
// App. A-4):  Adam  m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g^2 ; p -= lr_t * m / (sqrt(v) + eps)
// with lr_t = lr*sqrt(1-b2^t)/(1-b1^t) computed on the host;  SGD  p -= lr*g.
// `l2` adds the regulariser gradient 2*l2*p (embedding tables with layers_l2reg[0] != 0).
// HBM-bound: 28 bytes per element (read p,g,m,v; write p,m,v); 128-bit accesses, grid-stride.
#include "launchers.h"

namespace mr {

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float lr_t, float b1, float b2,
                                      float eps) {
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  p = p - lr_t * m / (sqrtf(v) + eps);
}

template <bool ADAM>
__global__ void __launch_bounds__(256) optimizer_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                             float* __restrict__ m, float* __restrict__ v,
                                                             int64_t n, float lr_t, float b1, float b2, float eps,
                                                             float l2, bool vec) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  const float c2 = 2.f * l2;
  if (vec) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += nth) {
      float4 pv = reinterpret_cast<float4*>(p)[i];
      float4 gv = reinterpret_cast<const float4*>(g)[i];
      if (l2 != 0.f) {
        gv.x += c2 * pv.x; gv.y += c2 * pv.y; gv.z += c2 * pv.z; gv.w += c2 * pv.w;
      }
      if (ADAM) {
        float4 mv = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        adam1(pv.x, gv.x, mv.x, vv.x, lr_t, b1, b2, eps);
        adam1(pv.y, gv.y, mv.y, vv.y, lr_t, b1, b2, eps);
        adam1(pv.z, gv.z, mv.z, vv.z, lr_t, b1, b2, eps);
        adam1(pv.w, gv.w, mv.w, vv.w, lr_t, b1, b2, eps);
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(v)[i] = vv;
      } else {
        pv.x -= lr_t * gv.x; pv.y -= lr_t * gv.y; pv.z -= lr_t * gv.z; pv.w -= lr_t * gv.w;
      }
      reinterpret_cast<float4*>(p)[i] = pv;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += nth) {
      float gi = g[i] + (l2 != 0.f ? c2 * p[i] : 0.f);
      if (ADAM) adam1(p[i], gi, m[i], v[i], lr_t, b1, b2, eps);
      else p[i] -= lr_t * gi;
    }
  } else {
    for (int64_t i = tid; i < n; i += nth) {
      float gi = g[i] + (l2 != 0.f ? c2 * p[i] : 0.f);
      if (ADAM) adam1(p[i], gi, m[i], v[i], lr_t, b1, b2, eps);
      else p[i] -= lr_t * gi;
    }
  }
}

int launch_optimizer_flat(float* p, const float* g, float* m, float* v, int64_t n, int optimizer, float lr_t,
                          float beta_1, float beta_2, float epsilon, float l2, cudaStream_t st) {
  if (n == 0) return MR_OK;
  uintptr_t a = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g);
  if (optimizer == MR_OPT_ADAM) a |= reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v);
  const bool vec = (a & 15) == 0;
  int64_t blocks = ((vec ? n / 4 : n) + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (optimizer == MR_OPT_ADAM)
    optimizer_flat_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, lr_t, beta_1, beta_2, epsilon, l2, vec);
  else
    optimizer_flat_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, lr_t, beta_1, beta_2, epsilon, l2, vec);
  MR_LAUNCH_CHECK("optimizer_flat_kernel");
  return MR_OK;
}

}  // namespace mr
