// Backward-pass kernels for the training step (BASELINE config 5; reference train.py:234-262 runs
// loss.backward() through the same modules).  The dense contractions of the backward pass (dgrad = dY.W and
// wgrad = dY^T.X) reuse the tcgen05 GEMM on transposed bf16 operands; what lives here are the HBM-bound pieces
// autograd would otherwise run as ATen kernels: operand transposition, LayerNorm backward, GELU forward/backward,
// the backward of the small attentions, and the bias-gradient column sum.  All fp32 in / fp32 out.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace afft {

// dst[c, r] = src[r, c], bf16 -> bf16 (src [rows, cols] pitch lds, dst [cols, rows] pitch ldd)
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, long long lds, int rows, int cols,
                                      __nv_bfloat16* __restrict__ dst, long long ldd) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? src[r * lds + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows) dst[c * ldd + r] = tile[threadIdx.x][j];
  }
}

// fp32 [rows, cols] -> bf16 row-major copy (pitch ldh) AND bf16 transpose [cols, rows] (pitch ldt) in one pass over the
// source, optionally with the column sums (bias gradient: colsum[c] += sum_r src[r, c]).  A Linear's backward needs its
// dy in both orientations (dgrad contracts over N, wgrad over the rows) and the bias gradient: three passes over the
// fp32 tensor become one.  32 x 32 tiles through shared memory; grid (cols / 32, rows / 32), block (32, 8).
// OP 0: v = src.  OP 1: v = gelu(src) (the MLP's activation feeding the second Linear, both operand orientations from one
// pass over the pre-activation).  OP 2: v = aux * gelu'(src) (the activation's backward feeding the first Linear's
// dgrad/wgrad operands and its bias gradient; aux = the gradient w.r.t. the activation's output, pitch lda).
__device__ __forceinline__ float gelu_fwd_exact(float x, int kind);
__device__ __forceinline__ float gelu_grad_exact(float x, int kind);
template <int OP>
__global__ void __launch_bounds__(256) convert_dual_kernel(const float* __restrict__ src, long long lds, int rows, int cols,
                                                           __nv_bfloat16* __restrict__ hi, long long ldh,
                                                           __nv_bfloat16* __restrict__ tr, long long ldt,
                                                           float* __restrict__ colsum, const float* __restrict__ aux,
                                                           long long lda, int kind) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  float csum = 0.f;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    float v = (r < rows && c < cols) ? src[r * lds + c] : 0.f;
    if (OP == 1) v = gelu_fwd_exact(v, kind);
    if (OP == 2) v = (r < rows && c < cols) ? aux[r * lda + c] * gelu_grad_exact(v, kind) : 0.f;
    tile[j][threadIdx.x] = v;
    csum += v;
    if (hi != nullptr && r < rows && c < cols) hi[r * ldh + c] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  if (tr != nullptr) {
    for (int j = threadIdx.y; j < 32; j += 8) {
      const int c = c0 + j, r = r0 + threadIdx.x;
      if (c < cols && r < rows) tr[c * ldt + r] = __float2bfloat16_rn(tile[threadIdx.x][j]);
    }
  }
  if (colsum != nullptr) {  // 8 row lanes per column -> one atomic per column and block
    __syncthreads();
    tile[threadIdx.y][threadIdx.x] = csum;
    __syncthreads();
    if (threadIdx.y == 0 && c0 + threadIdx.x < cols) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += tile[i][threadIdx.x];
      atomicAdd(colsum + c0 + threadIdx.x, t);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  xhat = (x - mean) * rstd (recomputed), g = dy * gamma:
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat));  dgamma += sum_rows dy * xhat;  dbeta += sum_rows dy
// One warp per row (grid-stride), per-lane partial dgamma/dbeta in registers, one atomicAdd per element per block.
// ------------------------------------------------------------------------------------------------
struct LayerNormBwdArgs {
  const float* x;
  long long ldx;
  const float* gamma;  // may be nullptr (no affine)
  float eps;
  const float* dy;
  long long lddy;
  int rows, dim;
  float* dx;
  long long lddx;
  float* dgamma;  // accumulated (+=); may be nullptr
  float* dbeta;
};

template <int NV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const LayerNormBwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const float inv_d = 1.0f / static_cast<float>(a.dim);
  float4 pg[NV], pb[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) pg[i] = pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int row = gw; row < a.rows; row += nw) {
    const float4* xr = reinterpret_cast<const float4*>(a.x + row * a.ldx);
    const float4* dyr = reinterpret_cast<const float4*>(a.dy + row * a.lddy);
    float4 v[NV], d[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] = xr[lane + 32 * i];
      d[i] = dyr[lane + 32 * i];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + a.eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xhat
      pg[i].x += d[i].x * v[i].x; pg[i].y += d[i].y * v[i].y; pg[i].z += d[i].z * v[i].z; pg[i].w += d[i].w * v[i].w;
      pb[i].x += d[i].x; pb[i].y += d[i].y; pb[i].z += d[i].z; pb[i].w += d[i].w;
      if (a.gamma != nullptr) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma) + lane + 32 * i);
        d[i].x *= g.x; d[i].y *= g.y; d[i].z *= g.z; d[i].w *= g.w;  // g = dy * gamma
      }
      sg += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      sgx += (d[i].x * v[i].x + d[i].y * v[i].y) + (d[i].z * v[i].z + d[i].w * v[i].w);
    }
    const float mg = warp_sum(sg) * inv_d, mgx = warp_sum(sgx) * inv_d;
    float4* dxr = reinterpret_cast<float4*>(a.dx + row * a.lddx);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = rstd * (d[i].x - mg - v[i].x * mgx);
      o.y = rstd * (d[i].y - mg - v[i].y * mgx);
      o.z = rstd * (d[i].z - mg - v[i].z * mgx);
      o.w = rstd * (d[i].w - mg - v[i].w * mgx);
      dxr[lane + 32 * i] = o;
    }
  }
  if (a.dgamma != nullptr) {
    // Block-level reduction in shared memory first (8 warps -> 1), then ONE global atomic per element and block.
    // (Per-warp global atomics: 1280 warps x 2048 atomics on the same 2048 addresses took 102 us for 5 MB of data.)
    __shared__ float red[2 * NV * 128];
    for (int i = threadIdx.x; i < 2 * NV * 128; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 4;
      atomicAdd(red + c + 0, pg[i].x); atomicAdd(red + c + 1, pg[i].y);
      atomicAdd(red + c + 2, pg[i].z); atomicAdd(red + c + 3, pg[i].w);
      atomicAdd(red + NV * 128 + c + 0, pb[i].x); atomicAdd(red + NV * 128 + c + 1, pb[i].y);
      atomicAdd(red + NV * 128 + c + 2, pb[i].z); atomicAdd(red + NV * 128 + c + 3, pb[i].w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NV * 128; i += blockDim.x) {
      atomicAdd(a.dgamma + i, red[i]);
      atomicAdd(a.dbeta + i, red[NV * 128 + i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GELU forward / backward (elementwise, fp32).  kind 1: erf (nn.GELU), kind 2: tanh (HF gelu_new).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_fwd_exact(float x, int kind) {
  if (kind == 1) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}
__device__ __forceinline__ float gelu_grad_exact(float x, int kind) {
  if (kind == 1) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
    return cdf + x * pdf;
  }
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  const float t = tanhf(u);
  const float du = 0.7978845608028654f * (1.0f + 3.0f * 0.044715f * x * x);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
}
// 16-byte accesses when n is a multiple of 4 and the pointers are 16-byte aligned (every activation tensor of the path)
__global__ void gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int kind) {
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0) {
    for (long long i = tid; i < n / 4; i += nth) {
      const float4 v = reinterpret_cast<const float4*>(x)[i];
      reinterpret_cast<float4*>(y)[i] = make_float4(gelu_fwd_exact(v.x, kind), gelu_fwd_exact(v.y, kind),
                                                    gelu_fwd_exact(v.z, kind), gelu_fwd_exact(v.w, kind));
    }
    return;
  }
  for (long long i = tid; i < n; i += nth) y[i] = gelu_fwd_exact(x[i], kind);
}
__global__ void gelu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                long long n, int kind) {
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15u) == 0) {
    for (long long i = tid; i < n / 4; i += nth) {
      const float4 v = reinterpret_cast<const float4*>(x)[i];
      const float4 d = reinterpret_cast<const float4*>(dy)[i];
      reinterpret_cast<float4*>(dx)[i] = make_float4(d.x * gelu_grad_exact(v.x, kind), d.y * gelu_grad_exact(v.y, kind),
                                                     d.z * gelu_grad_exact(v.z, kind), d.w * gelu_grad_exact(v.w, kind));
    }
    return;
  }
  for (long long i = tid; i < n; i += nth) dx[i] = dy[i] * gelu_grad_exact(x[i], kind);
}

// out[c] += sum_r x[r, c]   (bias gradients).  Block = 32 columns x 8 row lanes: every warp reads 128 contiguous bytes
// of a row, the 8 row lanes (and gridDim.y blocks) stride over the rows, partial sums meet in shared memory and one
// atomic per column and block reaches global memory.  (The first version gave one thread a whole column stripe: at
// 2048 x 3806 it took 76 us for 31 MB.)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long long ld, int rows, int cols, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < cols)
    for (int r = blockIdx.y * 8 + ry; r < rows; r += gridDim.y * 8) s += x[r * ld + c];
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][cx];
    atomicAdd(out + c, t);
  }
}

// ------------------------------------------------------------------------------------------------
// SGD with (Nesterov) momentum and weight decay over a flat parameter buffer - torch.optim.SGD's update
// (reference expts/01_SA-Fuser_ek100_train.txt:48-52: nesterov, momentum 0.9; train.py builds torch.optim.SGD):
//   g += wd * p;  buf = momentum * buf + g;  g = nesterov ? g + momentum * buf : buf;  p -= lr * g
// (a zero-initialised buffer reproduces torch's "buf = g" first step).  The same pass writes the bf16 image of the new
// parameters: it is the GEMM operand of the next step, so no per-step fp32 -> bf16 conversion of the 388 M weights.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgd_nesterov_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                           __nv_bfloat16* __restrict__ p16, long long n, float lr, float momentum,
                                                           float wd, int nesterov) {
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = tid; i < n / 4; i += nth) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float pe[4] = {pv.x, pv.y, pv.z, pv.w};
    const float ge[4] = {gv.x, gv.y, gv.z, gv.w};
    float me[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gg = fmaf(wd, pe[e], ge[e]);
      me[e] = fmaf(momentum, me[e], gg);
      pe[e] -= lr * (nesterov ? fmaf(momentum, me[e], gg) : me[e]);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pe[0], pe[1], pe[2], pe[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(me[0], me[1], me[2], me[3]);
    if (p16 != nullptr) reinterpret_cast<uint2*>(p16)[i] = make_uint2(pack_bf16x2(pe[0], pe[1]), pack_bf16x2(pe[2], pe[3]));
  }
  for (long long i = (n / 4) * 4 + tid; i < n; i += nth) {
    const float gg = fmaf(wd, p[i], g[i]);
    const float mm = fmaf(momentum, m[i], gg);
    m[i] = mm;
    p[i] -= lr * (nesterov ? fmaf(momentum, mm, gg) : mm);
    if (p16 != nullptr) p16[i] = __float2bfloat16_rn(p[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// Backward of the small attentions (SA-Fuser 5x5, GPT-2 18x18 causal, ...): one CTA per (sequence, head).
//   dV_j = sum_i P_ij dO_i;  dP_ij = dO_i . V_j;  dS_ij = P_ij (dP_ij - sum_j P_ij dP_ij) * scale;
//   dQ_i = sum_j dS_ij K_j;  dK_j = sum_i dS_ij Q_i.       Masked entries have P = 0, hence dS = 0.
// With attention-probability dropout (factors M_ij of the forward): dV uses P_ij M_ij and dP_ij = M_ij (dO_i . V_j).
// q/k/v fp32 at qkv[(seq*L + i)*ld + {0, D, 2D} + h*HD + d]; probs fp32 [n_seq, H, L, L]; dO fp32 [rows, D];
// dqkv fp32 with the layout of qkv.
// ------------------------------------------------------------------------------------------------
struct AttentionBwdArgs {
  const float* qkv;
  long long ld;
  const float* probs;
  const float* d_out;
  long long ldo;
  float* dqkv;
  int n_seq, L, H, HD;
  float scale;
  const float* drop;  // attention-probability dropout factors of the forward ([n_seq, H, L, L]; 0 or 1 / (1 - p)) or nullptr
};

__global__ void __launch_bounds__(256) attention_bwd_kernel(const AttentionBwdArgs a) {
  extern __shared__ float smem_ab[];
  const int L = a.L, HD = a.HD;
  const long long D = static_cast<long long>(a.H) * HD;
  float* sq = smem_ab;
  float* sk = sq + L * HD;
  float* sv = sk + L * HD;
  float* sdo = sv + L * HD;
  float* sp = sdo + L * HD;   // [L][L]
  float* sds = sp + L * L;    // [L][L]
  const int seq = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const float* base = a.qkv + static_cast<long long>(seq) * L * a.ld + h * HD;
  // 16-byte loads, all four operands of a position in flight together (HD, ld, ldo are multiples of 4; the first version
  // issued 4 scalar loads per element and took 244 us per launch at 128 clips)
  for (int i = threadIdx.x; i < L * (HD / 4); i += blockDim.x) {
    const int r = i / (HD / 4), d = (i % (HD / 4)) * 4;
    const float4 q4 = *reinterpret_cast<const float4*>(base + r * a.ld + d);
    const float4 k4 = *reinterpret_cast<const float4*>(base + r * a.ld + D + d);
    const float4 v4 = *reinterpret_cast<const float4*>(base + r * a.ld + 2 * D + d);
    const float4 o4 = *reinterpret_cast<const float4*>(a.d_out + (static_cast<long long>(seq) * L + r) * a.ldo + h * HD + d);
    *reinterpret_cast<float4*>(sq + r * HD + d) = q4;
    *reinterpret_cast<float4*>(sk + r * HD + d) = k4;
    *reinterpret_cast<float4*>(sv + r * HD + d) = v4;
    *reinterpret_cast<float4*>(sdo + r * HD + d) = o4;
  }
  const float* pg = a.probs + (static_cast<long long>(seq) * a.H + h) * L * L;
  const float* dm = a.drop != nullptr ? a.drop + (static_cast<long long>(seq) * a.H + h) * L * L : nullptr;
  for (int i = threadIdx.x; i < L * L; i += blockDim.x) sp[i] = pg[i];
  __syncthreads();
  // dP (stored in sds)
  for (int ij = threadIdx.x; ij < L * L; ij += blockDim.x) {
    const int i = ij / L, j = ij % L;
    float acc = 0.f;
    const float m = dm != nullptr ? dm[ij] : 1.0f;
    if (sp[ij] != 0.f && m != 0.f) {
      const float* o = sdo + i * HD;
      const float* v = sv + j * HD;
      for (int d = 0; d < HD; ++d) acc = fmaf(o[d], v[d], acc);
    }
    sds[ij] = acc * m;
  }
  __syncthreads();
  // dS = P * (dP - rowsum(P * dP)) * scale, one thread per row
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    float rs = 0.f;
    for (int j = 0; j < L; ++j) rs = fmaf(sp[i * L + j], sds[i * L + j], rs);
    for (int j = 0; j < L; ++j) sds[i * L + j] = sp[i * L + j] * (sds[i * L + j] - rs) * a.scale;
  }
  __syncthreads();
  if (dm != nullptr) {  // dV below needs the dropped probabilities P_ij M_ij; dS no longer needs P
    for (int i = threadIdx.x; i < L * L; i += blockDim.x) sp[i] *= dm[i];
    __syncthreads();
  }
  float* dbase = a.dqkv + static_cast<long long>(seq) * L * a.ld + h * HD;
  for (int id = threadIdx.x; id < L * (HD / 4); id += blockDim.x) {
    const int r = id / (HD / 4), d = (id % (HD / 4)) * 4;
    float4 dq = make_float4(0.f, 0.f, 0.f, 0.f), dk = dq, dv = dq;
    for (int j = 0; j < L; ++j) {
      const float s_rj = sds[r * L + j], s_jr = sds[j * L + r], p_jr = sp[j * L + r];
      const float4 k4 = *reinterpret_cast<const float4*>(sk + j * HD + d);
      const float4 q4 = *reinterpret_cast<const float4*>(sq + j * HD + d);
      const float4 o4 = *reinterpret_cast<const float4*>(sdo + j * HD + d);
      dq.x = fmaf(s_rj, k4.x, dq.x); dq.y = fmaf(s_rj, k4.y, dq.y); dq.z = fmaf(s_rj, k4.z, dq.z); dq.w = fmaf(s_rj, k4.w, dq.w);  // dQ_r = sum_j dS_rj K_j
      dk.x = fmaf(s_jr, q4.x, dk.x); dk.y = fmaf(s_jr, q4.y, dk.y); dk.z = fmaf(s_jr, q4.z, dk.z); dk.w = fmaf(s_jr, q4.w, dk.w);  // dK_r = sum_i dS_ir Q_i
      dv.x = fmaf(p_jr, o4.x, dv.x); dv.y = fmaf(p_jr, o4.y, dv.y); dv.z = fmaf(p_jr, o4.z, dv.z); dv.w = fmaf(p_jr, o4.w, dv.w);  // dV_r = sum_i P_ir dO_i
    }
    *reinterpret_cast<float4*>(dbase + r * a.ld + d) = dq;
    *reinterpret_cast<float4*>(dbase + r * a.ld + D + d) = dk;
    *reinterpret_cast<float4*>(dbase + r * a.ld + 2 * D + d) = dv;
  }
}

}  // namespace afft
