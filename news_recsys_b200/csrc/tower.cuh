// Shared geometry of the K4 tower kernels (tower.cu: backward + one-tile forward, tower_fwd.cu: pipelined forward).
#pragma once
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace nrx {

static constexpr int kRows = 128;        // rows per tile == UMMA M
static constexpr int kFwdThreads = 512;  // 16 warps: TMEM lane quadrant = warp % 4, column slice = warp / 4
static constexpr int kColSplit = kFwdThreads / 128;
static constexpr int kBiasStride = 256;   // floats of shared memory per layer bias
static constexpr int kMaxTiny = 4;

struct TowerK {
  int n_layers, n_mma, tiny;
  int K[NRX_MAX_LAYERS], N[NRX_MAX_LAYERS], Kp[NRX_MAX_LAYERS], Np[NRX_MAX_LAYERS];
  const float* w[NRX_MAX_LAYERS];
  const float* bias[NRX_MAX_LAYERS];
  unsigned w_off[NRX_MAX_LAYERS], wt_off[NRX_MAX_LAYERS];  // byte offsets inside the W / W^T image blocks
  unsigned w_bytes, wt_bytes;
  long long act_off[NRX_MAX_LAYERS];  // ws byte offset of the image of layer l's INPUT (width Kp[l])
  long long dz_off[NRX_MAX_LAYERS];   // ws byte offset of the image of dL/dz_l (width Np[l])
  long long wpack_off, wtpack_off, part_off;
  long long part_layer_off[NRX_MAX_LAYERS];  // float offset of layer l inside one CTA's partial block
  long long part_stride;                     // floats per CTA
  int act;
  float slope;
  int max_kp;     // widest A operand (activation buffer width)
  int tmem_cols;
  int wide0;      // first layer wider than the one-tile kernels' 240 columns (pipelined forward / split backward only)
  long long n_tiles;
  size_t total_bytes;
};

int make_tower(const NrxTower* t, long long B, int training, TowerK* k);

__device__ __forceinline__ float act_fwd(float z, float slope) { return z > 0.f ? z : z * slope; }
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ---- pipelined forward (tower_fwd.cu) ----
// true if the warp-specialised kernel handles this tower (every MMA layer <= 128 outputs, weights + >= 2 input stages fit)
bool tower_fwd3_eligible(const TowerK& k);
// x (fp32, may be null when the a_0 image is already in ws) -> a_0 image -> pipelined forward.
int tower_fwd3_launch(const TowerK& k, const float* x, long long ldx, long long B, float* y, long long ldy, uint8_t* ws,
                      int training, const NrxTowerHead* head, cudaStream_t st);


// ---- pipelined dX chain (tower_bwd_dx.cu) ----
bool tower_dx3_eligible(const TowerK& k, bool need_gx);
int tower_dx3_launch(const TowerK& k, long long B, const float* gy, long long ldgy, float* gx, long long ldgx, int accumulate_gx,
                     uint8_t* ws, cudaStream_t st);

}  // namespace nrx
