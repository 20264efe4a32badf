// pf_stencil.cuh — K1: fused evaporate + diffuse, fp32, double-buffered field.
//
//   g  = c + Dk * L(c)      L = 5-point (N+S+W+E-4c) or 9-point ((4*cross + diag - 20c), Dk = D/6)
//   f' = keep * g           (evaporation, reference algo:354-356)
//
// Op order is the oracle's (oracle/pf_oracle.c po_step_field), each op rounded separately.
// HBM-bound: 4 B read + 4 B written per cell-update. Two variants:
//   k_stencil_rows  — register sliding window: a warp owns 128 columns and marches down a block of
//                     rows; every cell is loaded from global exactly once per CTA as part of a
//                     coalesced 512 B warp request (LDG.128), W/E neighbours come from warp
//                     shuffles, the two warp-edge scalars from L1/L2; loads are issued PF rows ahead.
//   k_stencil_tma   — (pf_stencil_tma.cuh) TMA box loads into an mbarrier-pipelined shared-memory
//                     ring, same register sliding window on the consumer side.
#pragma once
#include "pf_device.cuh"

struct StencilArgs {
  const float* cur;
  float* nxt;
  int W;            // columns
  int rows;         // local rows of the strip
  long long pitch;  // floats
  long long chan_stride;
  int r_begin, r_end;  // local rows to compute
  int row_lo, row_hi;  // clamp for row reads: reflect edge -> 0 / rows-1, ghost -> -1 / rows
  int rows_per_cta;
  float keep[PF_MAX_CHANNELS];
  float dk[PF_MAX_CHANNELS];
  float del_thr;
  PeerRows peer;  // neighbour ghost rows of the NEXT buffer (null when not attached)
  // next tick's deposits, applied to the output rows (pf_deposit_tile.cuh; TMA variant only)
  const uint32_t* dep_flag;   // != 0: lists are not usable this tick
  const uint32_t* dep_count;  // [tile] entries
  const unsigned char* dep_blk;  // [tile] start[] / col[] / sum[] block
  int dep_ty0, dep_ny;           // tile row of this launch's first CTA row; tile rows per channel
  const int* dep_dense;          // [tile] side buffer slot of a crowded tile, < 0 = none (pf_deposit_tile.cuh)
  float* dep_side;               // [slot][64 * 1024] per-cell totals: added to the output rows, then cleared
};

// per-channel constant without dynamic indexing of the kernel parameter block (which would make
// the compiler copy the whole struct to local memory)
__device__ __forceinline__ float pick4(const float (&v)[PF_MAX_CHANNELS], int ch) {
  return ch == 0 ? v[0] : (ch == 1 ? v[1] : (ch == 2 ? v[2] : v[3]));
}

template <int ST>
__device__ __forceinline__ float stencil_cell(float n, float c, float s, float w, float e, float nw,
                                              float ne, float sw, float se, float keep, float dk,
                                              float thr) {
  float v;
  if (ST == 0) {
    v = fmul(keep, c);
  } else {
    float cross = fadd(fadd(n, s), fadd(w, e));
    float lap;
    if (ST == 9) {
      float diag = fadd(fadd(nw, ne), fadd(sw, se));
      lap = fsub(fadd(fmul(4.0f, cross), diag), fmul(20.0f, c));
    } else {
      lap = fsub(cross, fmul(4.0f, c));
    }
    v = fmul(keep, fadd(c, fmul(dk, lap)));
  }
  if (thr > 0.0f && v <= thr) v = 0.0f;
  return v;
}

template <int ST>
__device__ __forceinline__ float4 stencil_vec(const float4& n, const float4& c, const float4& s,
                                              float cw, float ce, float nw, float ne, float sw,
                                              float se, float keep, float dk, float thr) {
  float4 o;
  o.x = stencil_cell<ST>(n.x, c.x, s.x, cw, c.y, nw, n.y, sw, s.y, keep, dk, thr);
  o.y = stencil_cell<ST>(n.y, c.y, s.y, c.x, c.z, n.x, n.z, s.x, s.z, keep, dk, thr);
  o.z = stencil_cell<ST>(n.z, c.z, s.z, c.y, c.w, n.y, n.w, s.y, s.w, keep, dk, thr);
  o.w = stencil_cell<ST>(n.w, c.w, s.w, c.z, ce, n.z, ne, s.z, se, keep, dk, thr);
  return o;
}

// one queued row: the lane's 4 columns plus, on the warp-edge lanes, the scalar just outside
struct RowQ {
  float4 v;
  float edge;  // lane 0: column col4-1 ; lane 31: column col4+4 ; unused elsewhere
};

__device__ __forceinline__ RowQ load_row(const float* __restrict__ rowp, int col4, int W, int lane) {
  RowQ q;
  q.v = __ldg(reinterpret_cast<const float4*>(rowp + col4));
  q.edge = 0.0f;
  if (lane == 0 && col4 > 0) q.edge = __ldg(rowp + col4 - 1);
  if (lane == 31 && col4 + 4 < W) q.edge = __ldg(rowp + col4 + 4);
  return q;
}

// west/east neighbours of a row vector: shuffles inside the warp, queued scalars at its edges,
// duplicated edge value at the arena wall (reflect)
__device__ __forceinline__ void row_we(const RowQ& q, int col4, int W, int lane, float& w, float& e) {
  w = __shfl_up_sync(0xffffffffu, q.v.w, 1);
  e = __shfl_down_sync(0xffffffffu, q.v.x, 1);
  if (lane == 0) w = q.edge;
  if (lane == 31) e = q.edge;
  if (col4 == 0) w = q.v.x;
  if (col4 + 4 >= W) e = q.v.w;
}

constexpr int kStencilThreads = 256;
constexpr int kStencilPF = 4;  // rows of loads in flight per thread

template <int ST, bool PEER>
__global__ void __launch_bounds__(kStencilThreads)
k_stencil_rows(StencilArgs a) {
  const int lane = threadIdx.x & 31;
  const int ch = blockIdx.z;
  int col4 = (blockIdx.x * kStencilThreads + threadIdx.x) * 4;
  const bool live = col4 < a.W;
  if (!live) col4 = a.W - 4;  // keep the lane in the shuffles; its stores are predicated off
  const int r0 = a.r_begin + blockIdx.y * a.rows_per_cta;
  const int r1 = min(a.r_end, r0 + a.rows_per_cta);
  if (r0 >= r1) return;
  const float keep = pick4(a.keep, ch), dk = pick4(a.dk, ch), thr = a.del_thr;
  const float* __restrict__ base = a.cur + ch * a.chan_stride + a.pitch;  // local row 0
  float* __restrict__ obase = a.nxt + ch * a.chan_stride + a.pitch;
  float* const peer_first = (PEER && r0 == 0 && a.peer.up && live) ? a.peer.up + ch * a.peer.up_cs + col4 : nullptr;
  float* const peer_last = (PEER && r1 == a.rows && a.peer.dn && live) ? a.peer.dn + ch * a.peer.dn_cs + col4 : nullptr;
  const int last_row = a.rows - 1;

  auto rowp = [&](int r) -> const float* {
    r = max(a.row_lo, min(a.row_hi, r));
    return base + (long long)r * a.pitch;
  };

  RowQ qn = load_row(rowp(r0 - 1), col4, a.W, lane);
  RowQ qc = load_row(rowp(r0), col4, a.W, lane);
  RowQ q[kStencilPF];
#pragma unroll
  for (int i = 0; i < kStencilPF; ++i) q[i] = load_row(rowp(min(r0 + 1 + i, r1)), col4, a.W, lane);

  float nw, ne, cw, ce;
  row_we(qn, col4, a.W, lane, nw, ne);
  row_we(qc, col4, a.W, lane, cw, ce);

  for (int r = r0; r < r1; r += kStencilPF) {
#pragma unroll
    for (int i = 0; i < kStencilPF; ++i) {
      const int rr = r + i;
      if (rr < r1) {  // warp-uniform
        RowQ qs = q[i];
        q[i] = load_row(rowp(min(rr + 1 + kStencilPF, r1)), col4, a.W, lane);
        float sw, se;
        row_we(qs, col4, a.W, lane, sw, se);
        float4 o = stencil_vec<ST>(qn.v, qc.v, qs.v, cw, ce, nw, ne, sw, se, keep, dk, thr);
        if (live) *reinterpret_cast<float4*>(obase + (long long)rr * a.pitch + col4) = o;
        if (PEER) {
          if (peer_first && rr == 0) *reinterpret_cast<float4*>(peer_first) = o;
          if (peer_last && rr == last_row) *reinterpret_cast<float4*>(peer_last) = o;
        }
        qn = qc;
        qc = qs;
        nw = cw; ne = ce;
        cw = sw; ce = se;
      }
    }
  }
}

// generic fallback for widths that are not a multiple of 4 (e.g. the 497 x 495 heat-map of
// reference graph.py): one thread per cell, neighbours through L1/L2
template <int ST>
__global__ void __launch_bounds__(256)
k_stencil_scalar(StencilArgs a) {
  const int ch = blockIdx.z;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = a.r_begin + blockIdx.y;
  if (col >= a.W || r >= a.r_end) return;
  const float* __restrict__ base = a.cur + ch * a.chan_stride + a.pitch;
  const int rn = max(a.row_lo, r - 1), rs = min(a.row_hi, r + 1);
  const int jw = max(0, col - 1), je = min(a.W - 1, col + 1);
  const float* pc = base + (long long)r * a.pitch;
  const float* pn = base + (long long)rn * a.pitch;
  const float* ps = base + (long long)rs * a.pitch;
  float v = stencil_cell<ST>(pn[col], pc[col], ps[col], pc[jw], pc[je], pn[jw], pn[je], ps[jw],
                             ps[je], pick4(a.keep, ch), pick4(a.dk, ch), a.del_thr);
  a.nxt[ch * a.chan_stride + a.pitch + (long long)r * a.pitch + col] = v;
  if (r == 0 && a.peer.up) a.peer.up[ch * a.peer.up_cs + col] = v;
  if (r == a.rows - 1 && a.peer.dn) a.peer.dn[ch * a.peer.dn_cs + col] = v;
}
