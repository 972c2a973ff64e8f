// DSSM training tail, fused: L2-normalise the two tower outputs, gather the in-batch negatives, InfoNCE loss, and the
// complete backward down to the RAW tower outputs — two launches for what the reference does with ~25 small kernels.
//
//   reference: recall/DSSM/model.py:51-73  forward: negatives = item_emb[randperm(B)] x negative_sample_rate,
//                                          F.normalize(user), F.normalize(item), F.normalize(negatives)
//              recall/DSSM/model.py:92-110 infoNCE_loss: logits = [u.p, u.n_1 .. u.n_J] / temperature, cross entropy with
//                                          target 0 per sample, times the mask (label[:, 1]), mean over the batch
//
//   kernel A (one warp per SAMPLE b): norms, the 1 + J logits z, softmax, loss_b, dL/dz (of the MEAN loss), the whole
//       gradient of the user row (it only depends on sample b):  dU_b = (v - u^ (u^ . v)) / |u|,  v = sum_k dz_k x^_k / T,
//       u^ . v = sum_k dz_k z_k;  it also inverts the permutations: inv_j[perm_j[b]] = b.
//   kernel B (one warp per ITEM i): item i is the positive of sample i and, for every j, the negative of exactly ONE
//       sample inv_j[i] (perm_j is a permutation), so its gradient is a GATHER — deterministic, no float atomics:
//       w = sum_s dz_s(b_s) u^_{b_s} / T,  c = sum_s dz_s(b_s) z_s(b_s),  dI_i = (w - i^ c) / |i|.
// F.normalize's clamp (norm >= 1e-12) and a max-subtracted log-sum-exp are kept.  HBM-bound, a few rows per warp.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace nrx {

static constexpr int kNceMaxNeg = 7;     // negatives per sample
static constexpr int kNceMaxK = 8;       // elements per lane: d <= 256

struct NcePerms { const long long* p[kNceMaxNeg]; int n; };

__device__ __forceinline__ float nce_rnorm(float ss) { return 1.f / fmaxf(sqrtf(ss), 1e-12f); }

__global__ void __launch_bounds__(256)
infonce_sample_kernel(const float* __restrict__ U, long long uld, const float* __restrict__ I, long long ild, long long B, int d,
                      const __grid_constant__ NcePerms P, const float* __restrict__ mask, long long mstride, float inv_t,
                      float* __restrict__ loss, float* __restrict__ z_out, float* __restrict__ dz_out, int* __restrict__ inv,
                      float* __restrict__ gU, long long guld) {
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int J = P.n;
  float u[kNceMaxK], x[kNceMaxNeg + 1][kNceMaxK];   // x[0] = positive (item b), x[1 + j] = negative j
  long long nb[kNceMaxNeg + 1];
  nb[0] = b;
  for (int j = 0; j < J; ++j) nb[1 + j] = __ldg(P.p[j] + b);
  float uu = 0.f, xx[kNceMaxNeg + 1], ux[kNceMaxNeg + 1];
#pragma unroll
  for (int s = 0; s <= kNceMaxNeg; ++s) { xx[s] = 0.f; ux[s] = 0.f; }
#pragma unroll
  for (int k = 0; k < kNceMaxK; ++k) {
    const int c = lane + 32 * k;
    u[k] = c < d ? __ldg(U + b * uld + c) : 0.f;
    uu = fmaf(u[k], u[k], uu);
#pragma unroll
    for (int s = 0; s <= kNceMaxNeg; ++s) {
      if (s <= J) {
        const long long r = nb[s];
        x[s][k] = (c < d && r >= 0 && r < B) ? __ldg(I + r * ild + c) : 0.f;
        xx[s] = fmaf(x[s][k], x[s][k], xx[s]);
        ux[s] = fmaf(u[k], x[s][k], ux[s]);
      }
    }
  }
  uu = warp_sum(uu);
#pragma unroll
  for (int s = 0; s <= kNceMaxNeg; ++s)
    if (s <= J) { xx[s] = warp_sum(xx[s]); ux[s] = warp_sum(ux[s]); }
  const float ru = nce_rnorm(uu);
  float rx[kNceMaxNeg + 1], z[kNceMaxNeg + 1], dz[kNceMaxNeg + 1];
  float zmax = -FLT_MAX;
#pragma unroll
  for (int s = 0; s <= kNceMaxNeg; ++s)
    if (s <= J) { rx[s] = nce_rnorm(xx[s]); z[s] = ux[s] * ru * rx[s] * inv_t; zmax = fmaxf(zmax, z[s]); }
  float se = 0.f;
#pragma unroll
  for (int s = 0; s <= kNceMaxNeg; ++s)
    if (s <= J) se += expf(z[s] - zmax);
  const float lse = zmax + logf(se);
  const float mk = mask ? __ldg(mask + b * mstride) : 1.f;
  const float scale = mk / (float)B;
  float cu = 0.f;   // u^ . v = sum_k dz_k z_k
#pragma unroll
  for (int s = 0; s <= kNceMaxNeg; ++s)
    if (s <= J) { dz[s] = (expf(z[s] - lse) - (s == 0 ? 1.f : 0.f)) * scale; cu = fmaf(dz[s], z[s], cu); }
  if (lane == 0) {
    loss[b] = (lse - z[0]) * mk;
    for (int s = 0; s <= J; ++s) { z_out[b * (J + 1) + s] = z[s]; dz_out[b * (J + 1) + s] = dz[s]; }
    for (int j = 0; j < J; ++j) {
      const long long r = nb[1 + j];
      if (r >= 0 && r < B) inv[(long long)j * B + r] = (int)b;
    }
  }
  if (gU != nullptr) {
#pragma unroll
    for (int k = 0; k < kNceMaxK; ++k) {
      const int c = lane + 32 * k;
      if (c >= d) continue;
      float v = 0.f;
#pragma unroll
      for (int s = 0; s <= kNceMaxNeg; ++s)
        if (s <= J) v = fmaf(dz[s] * rx[s] * inv_t, x[s][k], v);
      gU[b * guld + c] = ru * (v - u[k] * ru * cu);
    }
  }
}

__global__ void __launch_bounds__(256)
infonce_item_kernel(const float* __restrict__ U, long long uld, const float* __restrict__ I, long long ild, long long B, int d,
                    const __grid_constant__ NcePerms P, float inv_t, const float* __restrict__ z, const float* __restrict__ dz,
                    const int* __restrict__ inv, float* __restrict__ gI, long long gild, int* __restrict__ status) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= B) return;
  const int J = P.n;
  float it[kNceMaxK], w[kNceMaxK];
  float ii = 0.f;
#pragma unroll
  for (int k = 0; k < kNceMaxK; ++k) {
    const int c = lane + 32 * k;
    it[k] = c < d ? __ldg(I + i * ild + c) : 0.f;
    ii = fmaf(it[k], it[k], ii);
    w[k] = 0.f;
  }
  const float ri = nce_rnorm(warp_sum(ii));
  float ctot = 0.f;
  for (int s = 0; s <= J; ++s) {
    long long b = i;
    if (s > 0) {
      b = inv[(long long)(s - 1) * B + i];
      // perm_j must be a permutation (torch.randperm): the sample that drew item i as its j-th negative exists and is unique
      if (b < 0 || b >= B || __ldg(P.p[s - 1] + b) != i) { if (lane == 0 && status) atomicOr(status, 1); continue; }
    }
    float ub[kNceMaxK], uu = 0.f;
#pragma unroll
    for (int k = 0; k < kNceMaxK; ++k) {
      const int c = lane + 32 * k;
      ub[k] = c < d ? __ldg(U + b * uld + c) : 0.f;
      uu = fmaf(ub[k], ub[k], uu);
    }
    const float ru = nce_rnorm(warp_sum(uu));
    const float g = __ldg(dz + b * (J + 1) + s);
    ctot = fmaf(g, __ldg(z + b * (J + 1) + s), ctot);
    const float a = g * inv_t * ru;
#pragma unroll
    for (int k = 0; k < kNceMaxK; ++k) w[k] = fmaf(a, ub[k], w[k]);
  }
#pragma unroll
  for (int k = 0; k < kNceMaxK; ++k) {
    const int c = lane + 32 * k;
    if (c < d) gI[i * gild + c] = ri * (w[k] - it[k] * ri * ctot);
  }
}

}  // namespace nrx

extern "C" size_t nrx_dssm_infonce_workspace_bytes(int64_t B, int n_neg) {
  if (B < 0 || n_neg < 0 || n_neg > nrx::kNceMaxNeg) return 0;
  const size_t zb = ((size_t)B * (n_neg + 1) * sizeof(float) + 255) & ~(size_t)255;
  const size_t ib = ((size_t)B * (n_neg > 0 ? n_neg : 1) * sizeof(int) + 255) & ~(size_t)255;
  return 2 * zb + ib + 256;
}

extern "C" int nrx_dssm_infonce(const float* user, int64_t u_ld, const float* item, int64_t i_ld, int64_t B, int d,
                                const int64_t* const* h_perms, int n_neg, const float* mask, int64_t mask_stride,
                                float temperature, float* loss_per_sample, float* grad_user, int64_t gu_ld, float* grad_item,
                                int64_t gi_ld, int32_t* status, void* ws, size_t ws_bytes, nrx_stream_t stream) {
  using namespace nrx;
  NRX_REQUIRE(B >= 0 && d >= 1 && d <= 32 * kNceMaxK, NRX_EUNSUPPORTED, "infonce: d=%d outside [1,%d]", d, 32 * kNceMaxK);
  NRX_REQUIRE(n_neg >= 0 && n_neg <= kNceMaxNeg, NRX_EUNSUPPORTED, "infonce: %d negatives per sample (max %d)", n_neg, kNceMaxNeg);
  NRX_REQUIRE(temperature > 0.f, NRX_EINVAL, "infonce: temperature must be positive");
  if (B == 0) return NRX_OK;
  NRX_REQUIRE(user && item && loss_per_sample && u_ld >= d && i_ld >= d, NRX_EINVAL, "infonce: null / bad argument");
  NRX_REQUIRE((!grad_user || gu_ld >= d) && (!grad_item || gi_ld >= d), NRX_EINVAL, "infonce: gradient leading dimension < d");
  NRX_REQUIRE(n_neg == 0 || h_perms != nullptr, NRX_EINVAL, "infonce: null permutation list");
  const size_t need = nrx_dssm_infonce_workspace_bytes(B, n_neg);
  NRX_REQUIRE(ws && ws_bytes >= need, NRX_EWORKSPACE, "workspace %zu < %zu", ws_bytes, need);
  NcePerms P;
  memset(&P, 0, sizeof(P));
  P.n = n_neg;
  for (int j = 0; j < n_neg; ++j) {
    NRX_REQUIRE(h_perms[j] != nullptr, NRX_EINVAL, "infonce: null permutation %d", j);
    P.p[j] = (const long long*)h_perms[j];
  }
  const size_t zb = ((size_t)B * (n_neg + 1) * sizeof(float) + 255) & ~(size_t)255;
  float* z = (float*)ws;
  float* dz = (float*)((char*)ws + zb);
  int* inv = (int*)((char*)ws + 2 * zb);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_neg > 0) {   // -1 = "nobody drew this item": caught in kernel B when a list is not a permutation
    cudaError_t e = cudaMemsetAsync(inv, 0xff, (size_t)B * n_neg * sizeof(int), st);
    NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "memset: %s", cudaGetErrorString(e));
  }
  const unsigned blocks = (unsigned)((B + 7) / 8);
  infonce_sample_kernel<<<blocks, 256, 0, st>>>(user, u_ld, item, i_ld, B, d, P, mask, mask_stride, 1.f / temperature,
                                                loss_per_sample, z, dz, inv, grad_user, gu_ld);
  int rc = check_launch("infonce_sample");
  if (rc != NRX_OK) return rc;
  if (grad_item != nullptr) {
    infonce_item_kernel<<<blocks, 256, 0, st>>>(user, u_ld, item, i_ld, B, d, P, 1.f / temperature, z, dz, inv, grad_item, gi_ld,
                                                status);
    rc = check_launch("infonce_item");
  }
  return rc;
}
