// pf_device.cuh — device-side parameter block and the deterministic fp32 math shared by the kernels.
//
// Arithmetic contract: every fp32 operation is a separately rounded IEEE
// operation. The explicit __f*_rn intrinsics below are never contracted into FMA by nvcc, so the
// contract holds regardless of -fmad. No libdevice transcendentals on the agent path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pherofield.h"

#define PF_PI_F 3.14159274101257324f
#define PF_TWO_PI_F 6.28318548202514648f
#define PF_HALF_PI_F 1.57079637050628662f
#define PF_QUARTER_PI_F 0.785398185253143311f

struct PFDev {
  // geometry
  int H, W, C, row0, row1, rows;
  long long pitch;        // floats per row
  long long chan_stride;  // floats per channel plane = (rows + 2) * pitch
  float ox, oy, inv_cs, xmin, xmax, ymin, ymax;
  // stencil
  int stencil;
  float keep[PF_MAX_CHANNELS];
  float dk[PF_MAX_CHANNELS];  // D (5-point) or D/6 (9-point)
  float del_thr;
  // deposit
  float i_explore, i_high, i_homing;
  int dep_red;  // all intensities >= 1e-30: "cell += sum" may be a reduction at L2 (see pf_deposit_owner.cuh)
  int dep_exact;  // all intensities are integers in [1, 32768]: a cell's total is exact whatever the order of the
                  // additions, so the deposits riding on the stencil need no grouping pass (pf_deposit_tile.cuh)
  unsigned high_every;
  float move_thr;
  int dep_ch[4], sen_ch[4];
  // sense
  int sense_mode;
  int drow[5], dcol[5];
  float sense_dist;
  // decide
  float kp, kd, cruise, vmax, bl, br, home_radius, home_lo, home_hi, nest_x, nest_y;
  int n_food;
  float food_x[PF_MAX_FOOD], food_y[PF_MAX_FOOD];
  float food_r2;
  int fsm;
  // move
  float dt, wheel_r, axle;
};

// Peer-store halo (multi-GPU row strips over NVLink P2P): where this strip's first / last interior
// row is written, the same value is also stored into the neighbour strip's ghost row, so no halo
// message is needed. up = the neighbour above's bottom ghost row, dn = the neighbour below's top
// ghost row (channel 0 of the buffer being written; *_cs = that neighbour's channel stride).
struct PeerRows {
  float* up;
  long long up_cs;
  float* dn;
  long long dn_cs;
};

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

// Cody-Waite reduction by pi/2 + Cephes minimax polynomials; mirrors oracle/pf_oracle.c po_sincosf
__device__ __forceinline__ void pf_sincos(float t, float& s_out, float& c_out) {
  float k = rintf(fmul(t, 0.636619747f));
  float r = fsub(t, fmul(k, 1.5703125f));
  r = fsub(r, fmul(k, 4.83751296997070312e-4f));
  r = fsub(r, fmul(k, 7.54978995489188e-8f));
  float z = fmul(r, r);
  float sp = fmul(-1.9515295891e-4f, z);
  sp = fadd(sp, 8.3321608736e-3f);
  sp = fmul(sp, z);
  sp = fsub(sp, 1.6666654611e-1f);
  sp = fmul(sp, z);
  sp = fmul(sp, r);
  float s = fadd(sp, r);
  float cp = fmul(2.443315711809948e-5f, z);
  cp = fsub(cp, 1.388731625493765e-3f);
  cp = fmul(cp, z);
  cp = fadd(cp, 4.166664568298827e-2f);
  cp = fmul(cp, z);
  cp = fmul(cp, z);
  cp = fsub(cp, fmul(0.5f, z));
  float c = fadd(cp, 1.0f);
  int q = ((int)k) & 3;
  if (q == 0) { s_out = s; c_out = c; }
  else if (q == 1) { s_out = c; c_out = -s; }
  else if (q == 2) { s_out = -s; c_out = -c; }
  else { s_out = -c; c_out = s; }
}

__device__ __forceinline__ float pf_atan2(float y, float x) {
  float ax = fabsf(x), ay = fabsf(y);
  if (ax == 0.0f && ay == 0.0f) return 0.0f;
  bool swap = ay > ax;
  float t = swap ? fdiv(ax, ay) : fdiv(ay, ax);
  float a0 = 0.0f;
  if (t > 0.414213568f) {
    t = fdiv(fsub(t, 1.0f), fadd(t, 1.0f));
    a0 = PF_QUARTER_PI_F;
  }
  float z = fmul(t, t);
  float p = fmul(8.05374449538e-2f, z);
  p = fsub(p, 1.38776856032e-1f);
  p = fmul(p, z);
  p = fadd(p, 1.99777106478e-1f);
  p = fmul(p, z);
  p = fsub(p, 3.33329491539e-1f);
  p = fmul(p, z);
  p = fmul(p, t);
  p = fadd(p, t);
  float a = fadd(a0, p);
  if (swap) a = fsub(PF_HALF_PI_F, a);
  if (x < 0.0f) a = fsub(PF_PI_F, a);
  if (y < 0.0f) a = -a;
  return a;
}

__device__ __forceinline__ float pf_wrap(float a) {
  if (a > PF_PI_F) a = fsub(a, PF_TWO_PI_F);
  if (a < -PF_PI_F) a = fadd(a, PF_TWO_PI_F);
  return a;
}

// PIDController.update (reference sup:199-226) with old_e = 0
__device__ __forceinline__ float pf_pid_turn(const PFDev& p, float cx, float cy, float gx, float gy) {
  float a = pf_atan2(fsub(gy, cy), fsub(gx, cx));
  float e = pf_wrap(fsub(a, PF_HALF_PI_F));
  return pf_wrap(fadd(fmul(p.kp, e), fmul(p.kd, e)));
}

// cell of a world position: truncation toward zero (Python int(), reference algo:351), clamped
// into the arena; returns whether the unclamped index lies inside (upper wall inclusive)
__device__ __forceinline__ bool pf_cell(const PFDev& p, float x, float y, int& row, int& col) {
  float fr = fmul(fsub(x, p.ox), p.inv_cs);
  float fc = fmul(fsub(y, p.oy), p.inv_cs);
  bool inside = (fr >= 0.0f) && (fc >= 0.0f) && (fr <= (float)p.H) && (fc <= (float)p.W);
  int r = __float2int_rz(fr), q = __float2int_rz(fc);
  if (!(fr >= 0.0f)) r = 0;
  if (!(fc >= 0.0f)) q = 0;
  if (fr >= (float)p.H) r = p.H - 1;
  if (fc >= (float)p.W) q = p.W - 1;
  row = r;
  col = q;
  return inside;
}

__device__ __forceinline__ long long pf_addr(const PFDev& p, int ch, int grow, int col) {
  return (long long)ch * p.chan_stride + (long long)(grow - p.row0 + 1) * p.pitch + col;
}
