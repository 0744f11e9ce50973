// The reference's DEFAULT tower (trainer.py DEFAULT_PARAMS: layers 64-32-16-8, here with GMF 8) as a projected,
// grouped train step on CUDA cores.  The layer widths are below every tensor-core tile, and the whole per-row network
// after the first layer is 32 -> 16 -> 8 -> 1: small enough for ONE THREAD per group of rows with everything in
// registers.  model.py:154-188 (Embedding + concat + Dense/ReLU + sigmoid head), :213-215 (BCE), backward of the same.
//
//   z1[r] = E_item[i_r] . W1i + (E_user[u_r] . W1u + b1) = Pi[i_r] + Pu[u_r]          (first layer, linear before ReLU)
//
//   small_rows_gemm_kernel        Pi = E_item . W1i, Pu = E_user . W1u + b1 over the TABLES (and, backward,
//                                 dE = S . W1^T on the per-item / per-user sums S of dZ1)
//   small_tower_train_kernel      thread = group (user row loaded once), loop over its rows: gather Pi row, ReLU,
//                                 32->16->8 forward from shared-memory weights, head + GMF, BCE, full backward to dZ1,
//                                 staged rows [dZ1 | GMF row gradient] per row (items) and per group (users).
//                                 Weight gradients: the warp stages its 32 rows' activations / pre-activation
//                                 gradients in shared memory and every lane owns a slice of the accumulators
//                                 (lane k: dW2[k][0..15], ...), so each outer product is computed once, in registers,
//                                 in a fixed order; CTAs write their sums to their row of the partial buffer.
//   small_table_wgrad_kernel      dW1 = E^T . S (and db1 = colsum(Su)) over the tables, per-CTA partial rows.
// No atomics on floats; results are bit-identical run to run.
#include "launchers.h"

namespace mr {
namespace st {

constexpr int L1 = 32, L2 = 16, L3 = 8, F = 8;
constexpr int SW = L1 + F;          // staged row: [dZ1 | GMF row gradient]
constexpr int kWarps = 4, kThreads = kWarps * 32;
constexpr int H1S = 33, DS = 28, H2S = 17, HDS = 17;  // row strides of the per-warp staging tiles (floats)
constexpr int kWeightFloats = 688;                    // W2 512, b2 16, W3 128, b3 8, w_out 16, b_out 1, padded to 16 B
constexpr int kWarpFloats = 32 * (H1S + DS + H2S + HDS);
constexpr int kSlots = 22;                            // per-lane accumulators: dW2 row 16, dW3 part 4, d w_out, bias sums

struct Params {
  const float *Pi, *Pu, *gmf_u, *gmf_i;
  const float *W2, *b2, *W3, *b3, *w_out, *b_out;
  const int32_t *users, *items;
  const float* labels;
  int64_t G;
  int num_users, num_items;
  float inv_batch;
  float *probs, *stage_i, *stage_u;
  float* partial;  // this launch's rows of the dense partial buffer (row = CTA), offsets of the blocks below
  int64_t stride;
  int off_W2, off_b2, off_W3, off_b3, off_wout, off_bout;
  float* loss_partial;
  int32_t* flags;
};

template <int GROUP>
__global__ void __launch_bounds__(kThreads, 4) small_tower_train_kernel(const Params p) {
  extern __shared__ __align__(16) float smem[];
  float* W2_s = smem;
  float* b2_s = W2_s + L1 * L2;
  float* W3_s = b2_s + L2;
  float* b3_s = W3_s + L2 * L3;
  float* wo_s = b3_s + L3;
  float* bo_s = wo_s + F + L3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* h1_s = smem + kWeightFloats + warp * kWarpFloats;
  float* d_s = h1_s + 32 * H1S;   // per row: dZ2 [0,16), dZ3 [16,24), dz [24]
  float* h2_s = d_s + 32 * DS;
  float* hd_s = h2_s + 32 * H2S;  // per row: GMF products [0,8), h3 [8,16)
  for (int i = tid; i < L1 * L2; i += kThreads) W2_s[i] = p.W2[i];
  for (int i = tid; i < L2 * L3; i += kThreads) W3_s[i] = p.W3[i];
  if (tid < L2) b2_s[tid] = p.b2[tid];
  if (tid < L3) b3_s[tid] = p.b3[tid];
  if (tid < F + L3) wo_s[tid] = p.w_out[tid];
  if (tid == 0) bo_s[0] = p.b_out[0];
  __syncthreads();

  float acc2[L2], acc3[4], dwo = 0.f, misc = 0.f, loss = 0.f;
#pragma unroll
  for (int j = 0; j < L2; ++j) acc2[j] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) acc3[j] = 0.f;
  const int k3 = lane & 15, q3 = 4 + (lane >> 4), mi = lane < 24 ? lane : 24;
  bool any_bad = false;

  for (int64_t g0 = ((int64_t)blockIdx.x * kWarps + warp) * 32; g0 < p.G; g0 += (int64_t)gridDim.x * kThreads) {
    const int64_t g = g0 + lane;
    const bool live = g < p.G;
    float gu[F], gug[F], gsum[L1];
    bool bad_u = false;
    const float4* pu4 = reinterpret_cast<const float4*>(p.Pu);  // the group's user row: re-read per row (L1 hits)
    if (live) {
      int u = __ldg(p.users + g * GROUP);
      bad_u = (unsigned)u >= (unsigned)p.num_users;
      if (bad_u) u = 0;
      pu4 = reinterpret_cast<const float4*>(p.Pu + (size_t)u * L1);
      const float4* gu4 = reinterpret_cast<const float4*>(p.gmf_u + (size_t)u * F);
#pragma unroll
      for (int c = 0; c < F / 4; ++c) {
        const float4 a = __ldg(gu4 + c);
        gu[4 * c] = a.x; gu[4 * c + 1] = a.y; gu[4 * c + 2] = a.z; gu[4 * c + 3] = a.w;
      }
    } else {  // idle lane of the last iteration: its staged rows contribute zeros
#pragma unroll
      for (int k = 0; k < L1; ++k) h1_s[lane * H1S + k] = 0.f;
#pragma unroll
      for (int k = 0; k < DS; ++k) d_s[lane * DS + k] = 0.f;
#pragma unroll
      for (int k = 0; k < L2; ++k) h2_s[lane * H2S + k] = 0.f;
#pragma unroll
      for (int k = 0; k < F + L3; ++k) hd_s[lane * HDS + k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < L1; ++k) gsum[k] = 0.f;
#pragma unroll
    for (int k = 0; k < F; ++k) gug[k] = 0.f;

#pragma unroll 1
    for (int j = 0; j < GROUP; ++j) {
      if (live) {
        const int64_t row = g * GROUP + j;
        int it = __ldg(p.items + row);
        const bool bad = bad_u || (unsigned)it >= (unsigned)p.num_items;
        if (bad) it = 0;
        any_bad |= bad;
        const float y = __ldg(p.labels + row);
        const float4* pi4 = reinterpret_cast<const float4*>(p.Pi + (size_t)it * L1);
        const float4* gi4 = reinterpret_cast<const float4*>(p.gmf_i + (size_t)it * F);
        float gi[F];
#pragma unroll
        for (int c = 0; c < F / 4; ++c) {
          const float4 a = __ldg(gi4 + c);
          gi[4 * c] = a.x; gi[4 * c + 1] = a.y; gi[4 * c + 2] = a.z; gi[4 * c + 3] = a.w;
        }
        // ---- forward ----
        float z2[L2];
#pragma unroll
        for (int o = 0; o < L2; ++o) z2[o] = b2_s[o];
        uint32_t m1 = 0;
#pragma unroll
        for (int c = 0; c < L1 / 4; ++c) {
          const float4 a = __ldg(pi4 + c), b = __ldg(pu4 + c);
          const float hv[4] = {fmaxf(a.x + b.x, 0.f), fmaxf(a.y + b.y, 0.f), fmaxf(a.z + b.z, 0.f), fmaxf(a.w + b.w, 0.f)};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k = 4 * c + q;
            h1_s[lane * H1S + k] = hv[q];
            if (hv[q] > 0.f) m1 |= 1u << k;
            const float4* w = reinterpret_cast<const float4*>(W2_s + k * L2);
#pragma unroll
            for (int o4 = 0; o4 < L2 / 4; ++o4) {
              const float4 wv = w[o4];
              z2[4 * o4] = fmaf(hv[q], wv.x, z2[4 * o4]);
              z2[4 * o4 + 1] = fmaf(hv[q], wv.y, z2[4 * o4 + 1]);
              z2[4 * o4 + 2] = fmaf(hv[q], wv.z, z2[4 * o4 + 2]);
              z2[4 * o4 + 3] = fmaf(hv[q], wv.w, z2[4 * o4 + 3]);
            }
          }
        }
        uint32_t m2 = 0;
        float z3[L3];
#pragma unroll
        for (int o = 0; o < L3; ++o) z3[o] = b3_s[o];
#pragma unroll
        for (int k = 0; k < L2; ++k) {
          const float h = fmaxf(z2[k], 0.f);
          h2_s[lane * H2S + k] = h;
          if (h > 0.f) m2 |= 1u << k;
          const float4* w = reinterpret_cast<const float4*>(W3_s + k * L3);
          const float4 w0 = w[0], w1 = w[1];
          z3[0] = fmaf(h, w0.x, z3[0]); z3[1] = fmaf(h, w0.y, z3[1]); z3[2] = fmaf(h, w0.z, z3[2]); z3[3] = fmaf(h, w0.w, z3[3]);
          z3[4] = fmaf(h, w1.x, z3[4]); z3[5] = fmaf(h, w1.y, z3[5]); z3[6] = fmaf(h, w1.z, z3[6]); z3[7] = fmaf(h, w1.w, z3[7]);
        }
        uint32_t m3 = 0;
        float s = bo_s[0];
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const float gp = gu[f] * gi[f];
          hd_s[lane * HDS + f] = gp;
          s = fmaf(wo_s[f], gp, s);
        }
#pragma unroll
        for (int o = 0; o < L3; ++o) {
          const float h = fmaxf(z3[o], 0.f);
          hd_s[lane * HDS + F + o] = h;
          if (h > 0.f) m3 |= 1u << o;
          s = fmaf(wo_s[F + o], h, s);
        }
        const float pr = sigmoidf_stable(s);
        const float dz = bad ? 0.f : (pr - y) * p.inv_batch;
        if (!bad) loss += bce_logits(s, y);
        if (p.probs != nullptr) p.probs[row] = bad ? nanf("") : pr;
        // ---- backward ----
        float dz3[L3];
#pragma unroll
        for (int o = 0; o < L3; ++o) dz3[o] = ((m3 >> o) & 1u) ? dz * wo_s[F + o] : 0.f;
        float* si = p.stage_i + (size_t)row * SW;
        {
          float gq[F];
#pragma unroll
          for (int f = 0; f < F; ++f) {
            const float gv = dz * wo_s[f];
            gug[f] = fmaf(gv, gi[f], gug[f]);
            gq[f] = gv * gu[f];
          }
          reinterpret_cast<float4*>(si + L1)[0] = make_float4(gq[0], gq[1], gq[2], gq[3]);
          reinterpret_cast<float4*>(si + L1)[1] = make_float4(gq[4], gq[5], gq[6], gq[7]);
        }
        float dz2[L2];
#pragma unroll
        for (int k = 0; k < L2; ++k) {
          const float4* w = reinterpret_cast<const float4*>(W3_s + k * L3);
          const float4 w0 = w[0], w1 = w[1];
          float v = dz3[0] * w0.x;
          v = fmaf(dz3[1], w0.y, v); v = fmaf(dz3[2], w0.z, v); v = fmaf(dz3[3], w0.w, v);
          v = fmaf(dz3[4], w1.x, v); v = fmaf(dz3[5], w1.y, v); v = fmaf(dz3[6], w1.z, v); v = fmaf(dz3[7], w1.w, v);
          dz2[k] = ((m2 >> k) & 1u) ? v : 0.f;
        }
        float4* dq = reinterpret_cast<float4*>(d_s + lane * DS);
        dq[0] = make_float4(dz2[0], dz2[1], dz2[2], dz2[3]);
        dq[1] = make_float4(dz2[4], dz2[5], dz2[6], dz2[7]);
        dq[2] = make_float4(dz2[8], dz2[9], dz2[10], dz2[11]);
        dq[3] = make_float4(dz2[12], dz2[13], dz2[14], dz2[15]);
        dq[4] = make_float4(dz3[0], dz3[1], dz3[2], dz3[3]);
        dq[5] = make_float4(dz3[4], dz3[5], dz3[6], dz3[7]);
        d_s[lane * DS + 24] = dz;
#pragma unroll
        for (int c = 0; c < L1 / 4; ++c) {
          float o4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k = 4 * c + q;
            const float4* w = reinterpret_cast<const float4*>(W2_s + k * L2);
            float v = 0.f;
#pragma unroll
            for (int o = 0; o < L2 / 4; ++o) {
              const float4 wv = w[o];
              v = fmaf(dz2[4 * o], wv.x, v); v = fmaf(dz2[4 * o + 1], wv.y, v);
              v = fmaf(dz2[4 * o + 2], wv.z, v); v = fmaf(dz2[4 * o + 3], wv.w, v);
            }
            v = ((m1 >> k) & 1u) ? v : 0.f;
            gsum[k] += v;
            o4[q] = v;
          }
          reinterpret_cast<float4*>(si)[c] = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
      }
      __syncwarp();
      // ---- weight gradients of the warp's 32 rows: lane = one slice of the accumulators ----
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const float a = h1_s[r * H1S + lane];
        const float4* dr = reinterpret_cast<const float4*>(d_s + r * DS);
#pragma unroll
        for (int o4 = 0; o4 < L2 / 4; ++o4) {
          const float4 dv = dr[o4];
          acc2[4 * o4] = fmaf(a, dv.x, acc2[4 * o4]);
          acc2[4 * o4 + 1] = fmaf(a, dv.y, acc2[4 * o4 + 1]);
          acc2[4 * o4 + 2] = fmaf(a, dv.z, acc2[4 * o4 + 2]);
          acc2[4 * o4 + 3] = fmaf(a, dv.w, acc2[4 * o4 + 3]);
        }
        const float b = h2_s[r * H2S + k3];
        const float4 t3 = dr[q3];
        acc3[0] = fmaf(b, t3.x, acc3[0]); acc3[1] = fmaf(b, t3.y, acc3[1]);
        acc3[2] = fmaf(b, t3.z, acc3[2]); acc3[3] = fmaf(b, t3.w, acc3[3]);
        dwo = fmaf(d_s[r * DS + 24], hd_s[r * HDS + k3], dwo);
        misc += d_s[r * DS + mi];
      }
      __syncwarp();
    }
    if (live) {
      float4* su = reinterpret_cast<float4*>(p.stage_u + (size_t)g * SW);
#pragma unroll
      for (int c = 0; c < L1 / 4; ++c) su[c] = make_float4(gsum[4 * c], gsum[4 * c + 1], gsum[4 * c + 2], gsum[4 * c + 3]);
      su[L1 / 4] = make_float4(gug[0], gug[1], gug[2], gug[3]);
      su[L1 / 4 + 1] = make_float4(gug[4], gug[5], gug[6], gug[7]);
    }
  }
  if (any_bad) atomicOr(p.flags, 1);

  // ---- per-CTA sums, fixed order over the warps, into this CTA's row of the partial buffer ----
  __syncthreads();
  float* red = smem + kWeightFloats;  // [warp][slot][lane]
  {
    float* rw = red + warp * kSlots * 32;
#pragma unroll
    for (int j = 0; j < L2; ++j) rw[j * 32 + lane] = acc2[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) rw[(16 + j) * 32 + lane] = acc3[j];
    rw[20 * 32 + lane] = dwo;
    rw[21 * 32 + lane] = misc;
  }
  loss = warp_sum(loss);
  __shared__ float loss_s[kWarps];
  if (lane == 0) loss_s[warp] = loss;
  __syncthreads();
  auto total = [&](int slot, int ln) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) v += red[(w * kSlots + slot) * 32 + ln];
    return v;
  };
  float* out = p.partial + (size_t)blockIdx.x * p.stride;
  for (int i = tid; i < L1 * L2; i += kThreads) out[p.off_W2 + i] = total(i % L2, i / L2);
  for (int i = tid; i < L2 * L3; i += kThreads) {
    const int k = i / L3, j = i % L3;
    out[p.off_W3 + i] = total(16 + (j & 3), k + 16 * (j >> 2));
  }
  if (tid < F + L3) out[p.off_wout + tid] = total(20, tid);
  if (tid < L2) out[p.off_b2 + tid] = total(21, tid);
  if (tid < L3) out[p.off_b3 + tid] = total(21, 16 + tid);
  if (tid == 0) {
    out[p.off_bout] = total(21, 24);
    float l = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) l += loss_s[w];
    p.loss_partial[blockIdx.x] = l;
  }
}

// out[r][j] = (bias ? bias[j] : 0) + sum_k A[r][k] * Wv(k, j), Wv(k, j) = transpose ? W[j * ldw + k] : W[k * ldw + j];
// K, N multiples of 4, K * N <= 4096.  256 threads = 256 / (N / 4) rows per CTA, four outputs per thread.
__global__ void __launch_bounds__(256) small_rows_gemm_kernel(const float* __restrict__ A, int64_t rows, int K,
                                                              const float* __restrict__ W, int ldw, int N, int transpose,
                                                              const float* __restrict__ bias, float* __restrict__ out) {
  __shared__ __align__(16) float W_s[4096];
  for (int i = threadIdx.x; i < K * N; i += 256) {
    const int k = i / N, j = i % N;
    W_s[i] = transpose ? W[(size_t)j * ldw + k] : W[(size_t)k * ldw + j];
  }
  __syncthreads();
  const int nq = N >> 2, rpc = 256 / nq;
  const int rl = threadIdx.x / nq, jq = threadIdx.x % nq;
  if (rl >= rpc) return;
  for (int64_t r = (int64_t)blockIdx.x * rpc + rl; r < rows; r += (int64_t)gridDim.x * rpc) {
    float4 acc = bias != nullptr ? *reinterpret_cast<const float4*>(bias + 4 * jq) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* a4 = reinterpret_cast<const float4*>(A + (size_t)r * K);
    for (int k4 = 0; k4 < (K >> 2); ++k4) {
      const float4 a = __ldg(a4 + k4);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(W_s + (4 * k4 + q) * N + 4 * jq);
        acc.x = fmaf(av[q], w.x, acc.x); acc.y = fmaf(av[q], w.y, acc.y);
        acc.z = fmaf(av[q], w.z, acc.z); acc.w = fmaf(av[q], w.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(out + (size_t)r * N + 4 * jq) = acc;
  }
}

// dW[i][j] = sum_r E[r][i] * S[r][j] (i < 32, j < 32), db[j] = sum_r S[r][j]: CTA = chunks of 128 rows staged in
// shared memory, thread = (i, four j); the CTA's sum goes to its row of the partial buffer.
constexpr int kWgRows = 128;
__global__ void __launch_bounds__(256) small_table_wgrad_kernel(const float* __restrict__ E, const float* __restrict__ S,
                                                                int64_t rows, float* __restrict__ dw_partial,
                                                                float* __restrict__ db_partial, int64_t stride) {
  __shared__ __align__(16) float E_s[kWgRows * 33];
  __shared__ __align__(16) float S_s[kWgRows * 32];
  const int i = threadIdx.x >> 3, jq = threadIdx.x & 7;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float db = 0.f;
  for (int64_t r0 = (int64_t)blockIdx.x * kWgRows; r0 < rows; r0 += (int64_t)gridDim.x * kWgRows) {
    const int n = (int)(rows - r0 < kWgRows ? rows - r0 : kWgRows);
    __syncthreads();
    for (int t = threadIdx.x; t < kWgRows * 32; t += 256) {
      const int r = t >> 5, c = t & 31;
      E_s[r * 33 + c] = r < n ? __ldg(E + (size_t)(r0 + r) * 32 + c) : 0.f;
      S_s[t] = r < n ? __ldg(S + (size_t)(r0 + r) * 32 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kWgRows; ++r) {
      const float e = E_s[r * 33 + i];
      const float4 s = *reinterpret_cast<const float4*>(S_s + r * 32 + 4 * jq);
      acc.x = fmaf(e, s.x, acc.x); acc.y = fmaf(e, s.y, acc.y); acc.z = fmaf(e, s.z, acc.z); acc.w = fmaf(e, s.w, acc.w);
    }
    if (db_partial != nullptr && threadIdx.x < 32)
      for (int r = 0; r < kWgRows; ++r) db += S_s[r * 32 + threadIdx.x];
  }
  float* dw = dw_partial + (size_t)blockIdx.x * stride;
  *reinterpret_cast<float4*>(dw + i * 32 + 4 * jq) = acc;
  if (db_partial != nullptr && threadIdx.x < 32) db_partial[(size_t)blockIdx.x * stride + threadIdx.x] = db;
}

}  // namespace st

bool small_tower_supported(const MrModel& m, int group) {
  return m.n_layers == 4 && m.L[0] == 2 * st::L1 && m.L[1] == st::L1 && m.L[2] == st::L2 && m.L[3] == st::L3 &&
         m.mf_dim == st::F && group == 5;
}

int launch_small_tower_train(const SmallTowerArgs& a, cudaStream_t stream, int* grid_out) {
  const MrModel& m = *a.model;
  st::Params p{};
  p.Pi = a.Pi; p.Pu = a.Pu; p.gmf_u = m.user_gmf; p.gmf_i = m.item_gmf;
  p.W2 = m.W[2]; p.b2 = m.b[2]; p.W3 = m.W[3]; p.b3 = m.b[3]; p.w_out = m.w_out; p.b_out = m.b_out;
  p.users = a.users; p.items = a.items; p.labels = a.labels;
  p.G = a.B / 5;
  p.num_users = m.num_users; p.num_items = m.num_items;
  p.inv_batch = a.inv_batch;
  p.probs = a.probs; p.stage_i = a.stage_i; p.stage_u = a.stage_u;
  p.partial = a.dense_partial; p.stride = a.dense_stride;
  p.off_W2 = (int)(m.W[2] - m.dense); p.off_b2 = (int)(m.b[2] - m.dense);
  p.off_W3 = (int)(m.W[3] - m.dense); p.off_b3 = (int)(m.b[3] - m.dense);
  p.off_wout = (int)(m.w_out - m.dense); p.off_bout = (int)(m.b_out - m.dense);
  p.loss_partial = a.loss_partial; p.flags = a.flags;
  const size_t smem = (size_t)(st::kWeightFloats + st::kWarps * st::kWarpFloats) * sizeof(float);
  auto kern = st::small_tower_train_kernel<5>;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    MR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int occ = 0;
  MR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, st::kThreads, smem));
  if (occ < 1) occ = 1;
  int64_t grid = (p.G + st::kThreads - 1) / st::kThreads;
  const int64_t cap = (int64_t)sm_count() * occ < a.max_ctas ? (int64_t)sm_count() * occ : a.max_ctas;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, st::kThreads, smem, stream>>>(p);
  MR_LAUNCH_CHECK("small_tower_train_kernel");
  *grid_out = (int)grid;
  return MR_OK;
}

int launch_small_rows_gemm(const float* A, int64_t rows, int K, const float* W, int ldw, int N, bool transpose,
                           const float* bias, float* out, cudaStream_t stream) {
  if (rows == 0) return MR_OK;
  if ((K & 3) || (N & 3) || K * N > 4096 || N > 256) {
    set_error("small_rows_gemm: K=%d N=%d not supported", K, N);
    return MR_ERR_INVALID;
  }
  const int rpc = 256 / (N >> 2);
  int64_t grid = (rows + rpc - 1) / rpc;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  st::small_rows_gemm_kernel<<<(unsigned)grid, 256, 0, stream>>>(A, rows, K, W, ldw, N, transpose ? 1 : 0, bias, out);
  MR_LAUNCH_CHECK("small_rows_gemm_kernel");
  return MR_OK;
}

int launch_small_table_wgrad(const float* E, const float* S, int64_t rows, float* dw_partial, float* db_partial,
                             int64_t stride, int max_ctas, cudaStream_t stream, int* grid_out) {
  int64_t grid = (rows + st::kWgRows - 1) / st::kWgRows;
  if (grid > max_ctas) grid = max_ctas;
  if (grid < 1) grid = 1;
  st::small_table_wgrad_kernel<<<(unsigned)grid, 256, 0, stream>>>(E, S, rows, dw_partial, db_partial, stride);
  MR_LAUNCH_CHECK("small_table_wgrad_kernel");
  *grid_out = (int)grid;
  return MR_OK;
}

}  // namespace mr
