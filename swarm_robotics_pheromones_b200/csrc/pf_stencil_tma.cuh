// pf_stencil_tma.cuh — K1 variant 2: TMA bulk copies into an mbarrier-pipelined shared-memory ring.
//
// One CTA owns a strip of kTmaCols columns and streams the rows r0-1 .. r1 of its row block through
// a ring of kStages stages x kStageRows rows. A single producer thread issues one
// cp.async.bulk.shared.global (SASS UBLKCP) per row segment — (kTmaCols + 8) floats, 16 B aligned,
// including 4 halo columns on each side — and arms the stage's `full` mbarrier with the byte count.
// Eight consumer warps (128 columns each) wait on `full`, pull their float4 (LDS.128, conflict-free)
// and the warp-edge scalar into registers, release the stage through the `empty` mbarrier, and run
// the same register sliding window as k_stencil_rows; results go straight to HBM with STG.128.
// Bytes in flight per SM are set by the ring depth, not by registers or occupancy.
#pragma once
#include "pf_stencil.cuh"

constexpr int kTmaWarps = 8;                  // consumer warps
constexpr int kTmaCols = kTmaWarps * 128;     // columns per CTA
constexpr int kTmaRowFloats = kTmaCols + 8;   // + 4 halo floats each side
constexpr int kTmaThreads = kTmaWarps * 32 + 32;  // + producer warp

constexpr int kDepRows = 64;    // rows per CTA when deposits ride along (= kTileRows)
constexpr int kDepEnt = 1408;   // = kTileEnt
constexpr int kDepPtrBytes = 48;           // uint16 start[] of the per-stage entry groups (= kTilePtrBytes)
constexpr int kDepColBytes = kDepEnt * 2;  // uint16 row << 10 | column
constexpr int kDepBlkBytes = kDepPtrBytes + kDepColBytes + kDepEnt * 4;

struct TmaStencilState {
  bool attr_set[3][9];
  char err[256];
};

static inline void tma_stencil_init(TmaStencilState* s) { memset(s, 0, sizeof(*s)); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// PEER = true only for launches that contain the strip's first / last row while neighbours are attached
// (the sharded driver launches those two rows on their own), so the streaming instantiation carries
// no peer logic at all.
// DEP = true: the CTA is one deposit tile (kDepRows x kTmaCols of one channel, whole-field launch);
// its (cell, sum) list arrives by two more bulk copies and the producer warp adds them to the rows
// the consumers have just stored (reductions performed by the L2).
template <int ST, int kStageRows, int kStages, bool PEER, bool DEP>
__global__ void __launch_bounds__(kTmaThreads, DEP ? 3 : 1)
k_stencil_tma(StencilArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)kStages * kStageRows * kTmaRowFloats);
  uint64_t* empty = full + kStages;
  static_assert(!DEP || kStageRows == 4, "deposit entries are grouped by stages of 4 rows (pf_deposit_tile.cuh)");
  uint64_t* dep_bar = empty + kStages;  // the tile's deposit lists have landed
  uint64_t* fin = dep_bar + 1;          // every consumer warp has issued its last store
  uint32_t* progress = reinterpret_cast<uint32_t*>(fin + 1);  // last iteration whose `empty` wait the producer passed
  // 16 B aligned: ring rows are 4128 B, 2 * kStages + 3 slots of 8 B
  unsigned char* dblk =
      smem_raw + (((size_t)kStages * kStageRows * kTmaRowFloats * 4 + (2 * kStages + 3) * 8 + 15) & ~(size_t)15);
  const uint16_t* dstart = reinterpret_cast<const uint16_t*>(dblk);
  const uint16_t* dcol = reinterpret_cast<const uint16_t*>(dblk + kDepPtrBytes);
  const float* dsum = reinterpret_cast<const float*>(dblk + kDepPtrBytes + kDepColBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ch = blockIdx.z;
  const int c0 = blockIdx.x * kTmaCols;
  const int r0 = a.r_begin + blockIdx.y * a.rows_per_cta;
  const int r1 = min(a.r_end, r0 + a.rows_per_cta);
  if (r0 >= r1) return;  // uniform per CTA
  const int nstream = r1 - r0 + 2;  // rows r0-1 .. r1
  const int niter = (nstream + kStageRows - 1) / kStageRows;

  // columns copied per row: [g0, g1), placed so that global column c0-4 maps to ring offset 0
  const int g0 = max(c0 - 4, 0);
  const int g1 = (int)min((long long)c0 + kTmaCols + 4, a.pitch);
  const uint32_t row_bytes = (uint32_t)(g1 - g0) * 4u;
  const int dst_off = g0 - (c0 - 4);

  uint32_t ndep = 0u;
  long long tile = 0;
  float* side = nullptr;      // crowded tile: per-cell totals to add to the output (uniform per CTA)
  bool side_add = false;
  if (DEP) {
    tile = ((long long)ch * a.dep_ny + a.dep_ty0 + blockIdx.y) * gridDim.x + blockIdx.x;
    const bool usable = *a.dep_flag == 0u;
    if (usable) ndep = a.dep_count[tile];  // uniform per CTA
    const int slot = a.dep_dense ? a.dep_dense[tile] : -1;
    if (slot >= 0) side = a.dep_side + (long long)slot * (kDepRows * kTmaCols);
    side_add = usable;   // a tick that fell back only clears the buffer: the regular deposit kernels redo it
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kTmaWarps);
    }
    if (DEP) {
      mbar_init(dep_bar, 1);
      mbar_init(fin, kTmaWarps);
      *progress = 0u;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const float* __restrict__ base = a.cur + ch * a.chan_stride + a.pitch;

  if (warp == kTmaWarps) {
    // ===== producer: one elected lane issues the copies =====
    // DEP: the other 31 lanes apply the tile's deposits, kStages + 1 stages behind the copies being
    // issued: once the consumers have released stage it - kStages (the producer's wait), every one of
    // them has issued its stores for the output rows of stage it - kStages - 1 (program order before
    // its arrive; arrive = release, wait = acquire), so a reduction sent to those cells now lands
    // on the stored value, in L2, microseconds after the store: nxt[cell] = stencil value + sum.
    float* const otile = a.nxt + ch * a.chan_stride + a.pitch + (long long)r0 * a.pitch + c0;
    if (lane == 0) {
      for (int it = 0; it < niter; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (uint32_t)(it / kStages) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        if (DEP && ndep)  // a counter, not the barrier's parity: the deposit lanes may lag by any amount
          asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(progress)), "r"((uint32_t)it) : "memory");
        const int first = it * kStageRows;
        const int cnt = min(kStageRows, nstream - first);
        mbar_expect_tx(&full[s], row_bytes * (uint32_t)cnt);
        for (int k = 0; k < cnt; ++k) {
          int r = r0 - 1 + first + k;
          r = max(a.row_lo, min(a.row_hi, r));
          bulk_g2s(ring + ((size_t)s * kStageRows + k) * kTmaRowFloats + dst_off,
                   base + (long long)r * a.pitch + g0, row_bytes, &full[s]);
        }
        if (DEP && it == 0 && ndep) {  // the tile's deposit list, behind the first stage
          const uint32_t col_bytes = (ndep * 2u + 15u) & ~15u, sum_bytes = (ndep * 4u + 15u) & ~15u;
          const unsigned char* src = a.dep_blk + tile * kDepBlkBytes;
          mbar_expect_tx(dep_bar, (uint32_t)kDepPtrBytes + col_bytes + sum_bytes);
          bulk_g2s(dblk, src, (uint32_t)kDepPtrBytes + col_bytes, dep_bar);  // start[] and cell[] are adjacent
          bulk_g2s(dblk + kDepPtrBytes + kDepColBytes, src + kDepPtrBytes + kDepColBytes, sum_bytes, dep_bar);
        }
      }
    } else if (DEP && ndep) {
      // lanes 1..31 (diverged from the producer lane for good; they only meet it at mbarriers)
      auto dep_stages = [&](int jb, int je) {  // the entries of stages [jb, je): output rows [4 jb - 2, 4 je - 2)
        const int e1 = dstart[je];
        for (int e = dstart[jb] + lane - 1; e < e1; e += 31) {
          const uint32_t dc = dcol[e];  // row << 10 | column, both within the tile
          atomicAdd(otile + (long long)(dc >> 10) * a.pitch + (dc & 1023u), dsum[e]);
        }
      };
      mbar_wait(dep_bar, 0u);
      for (int it = kStages + 1; it < niter; ++it) {
        while (true) {  // stage it - kStages released by every consumer warp?
          uint32_t seen;
          asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(progress)) : "memory");
          if ((int)seen >= it) break;
          __nanosleep(200);
        }
        dep_stages(it - kStages - 1, it - kStages);
      }
      mbar_wait(fin, 0u);  // the last kStages + 1 stages, after the consumers' last stores
      dep_stages(max(niter - kStages - 1, 0), niter);
    }
    return;
  }

  // ===== consumers =====
  const int t = threadIdx.x;  // 0..255
  int col4 = c0 + 4 * t;
  const bool live = col4 < a.W;
  const float keep = pick4(a.keep, ch), dk = pick4(a.dk, ch), thr = a.del_thr;
  float* __restrict__ obase = a.nxt + ch * a.chan_stride + a.pitch;
  const int my_off = 4 + 4 * t;  // ring offset of this thread's float4 within a row
  // fused halo (NVLink P2P): only the CTAs that own the strip's first / last row have a peer target;
  // everything is hoisted into registers so the row loop of all other CTAs is unchanged
  float* const peer_first = (PEER && r0 == 0 && a.peer.up && live) ? a.peer.up + ch * a.peer.up_cs + col4 : nullptr;
  float* const peer_last = (PEER && r1 == a.rows && a.peer.dn && live) ? a.peer.dn + ch * a.peer.dn_cs + col4 : nullptr;
  const int last_row = a.rows - 1;

  RowQ qn, qc;
  qn.v = make_float4(0.f, 0.f, 0.f, 0.f); qn.edge = 0.f;
  qc = qn;
  float nw = 0.f, ne = 0.f, cw = 0.f, ce = 0.f;

  for (int it = 0; it < niter; ++it) {
    const int s = it % kStages;
    const uint32_t ph = (uint32_t)(it / kStages) & 1u;
    mbar_wait(&full[s], ph);
    const int first = it * kStageRows;
    const int cnt = min(kStageRows, nstream - first);
    RowQ rows[kStageRows];
#pragma unroll
    for (int k = 0; k < kStageRows; ++k) {
      if (k < cnt) {
        const float* rp = ring + ((size_t)s * kStageRows + k) * kTmaRowFloats;
        rows[k].v = *reinterpret_cast<const float4*>(rp + my_off);
        rows[k].edge = 0.f;
        if (lane == 0) rows[k].edge = rp[my_off - 1];
        if (lane == 31) rows[k].edge = rp[my_off + 4];
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
#pragma unroll
    for (int k = 0; k < kStageRows; ++k) {
      if (k < cnt) {
        const int si = first + k;  // stream index; stream row = r0 - 1 + si
        float sw, se;
        row_we(rows[k], live ? col4 : a.W - 4, a.W, lane, sw, se);
        if (si >= 2) {
          const int rr = r0 + si - 2;
          float4 o = stencil_vec<ST>(qn.v, qc.v, rows[k].v, cw, ce, nw, ne, sw, se, keep, dk, thr);
          if (DEP && side && live) {  // crowded tile: next tick's totals of this row segment, touched cells only
            float4* sp = reinterpret_cast<float4*>(side + (long long)(rr - r0) * kTmaCols + 4 * t);
            const float4 sv = __ldcg(sp);
            if (sv.x != 0.f || sv.y != 0.f || sv.z != 0.f || sv.w != 0.f) {
              __stcg(sp, make_float4(0.f, 0.f, 0.f, 0.f));
              if (side_add) {
                if (sv.x != 0.f) o.x = fadd(o.x, sv.x);
                if (sv.y != 0.f) o.y = fadd(o.y, sv.y);
                if (sv.z != 0.f) o.z = fadd(o.z, sv.z);
                if (sv.w != 0.f) o.w = fadd(o.w, sv.w);
              }
            }
          }
          if (live) *reinterpret_cast<float4*>(obase + (long long)rr * a.pitch + col4) = o;
          if (PEER) {
            if (peer_first && rr == 0) *reinterpret_cast<float4*>(peer_first) = o;
            if (peer_last && rr == last_row) *reinterpret_cast<float4*>(peer_last) = o;
          }
        }
        qn = qc; qc = rows[k];
        nw = cw; ne = ce;
        cw = sw; ce = se;
      }
    }
  }
  if (DEP && ndep) {
    __syncwarp();
    if (lane == 0) mbar_arrive(fin);
  }
}

template <int ST, int R, int S, bool PEER, bool DEP = false>
static int tma_launch_one(TmaStencilState* st, int slot, const StencilArgs& a, dim3 grid, cudaStream_t s) {
  size_t smem = (size_t)S * R * kTmaRowFloats * sizeof(float) + (2 * S + 3) * sizeof(uint64_t);
  if (DEP) smem = ((smem + 15) & ~(size_t)15) + kDepBlkBytes;
  const int sti = ST == 0 ? 0 : (ST == 5 ? 1 : 2);
  if (!st->attr_set[sti][slot]) {
    cudaError_t e = cudaFuncSetAttribute(k_stencil_tma<ST, R, S, PEER, DEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      snprintf(st->err, sizeof(st->err), "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return PF_ECUDA;
    }
    st->attr_set[sti][slot] = true;
    if (getenv("PF_DEBUG_OCC")) {
      int occ = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stencil_tma<ST, R, S, PEER, DEP>, kTmaThreads, smem);
      fprintf(stderr, "[pherofield] k_stencil_tma<%d,%d,%d,%d,%d>: %zu B smem, %d CTAs/SM\n", ST, R, S, (int)PEER,
              (int)DEP, smem, occ);
    }
  }
  k_stencil_tma<ST, R, S, PEER, DEP><<<grid, kTmaThreads, smem, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(st->err, sizeof(st->err), "launch: %s", cudaGetErrorString(e));
    return PF_ECUDA;
  }
  return PF_OK;
}

// rows_per_cta <= 0 selects the default. Returns PF_ESTATE when the shape is not supported
// (the caller then uses the register sliding-window kernel).
static int tma_stencil_launch(TmaStencilState* st, int shape, const pf_config& cfg, const PFDev& d, StencilArgs a,
                              int rows_per_cta, cudaStream_t s) {
  (void)cfg;
  if (d.W % 4 != 0 || d.W < 128) return PF_ESTATE;
  const int nrows = a.r_end - a.r_begin;
  int rpc = rows_per_cta > 0 ? rows_per_cta : (d.stencil == 9 ? 64 : 128);
  {  // small fields: keep at least ~2 CTAs per SM in flight
    const long long xb = (d.W + kTmaCols - 1) / kTmaCols;
    while (rpc > 8 && xb * d.C * ((nrows + rpc - 1) / rpc) < 296) rpc /= 2;
  }
  if (rpc > nrows) rpc = nrows;
  a.rows_per_cta = rpc;
  dim3 grid((d.W + kTmaCols - 1) / kTmaCols, (nrows + rpc - 1) / rpc, d.C);
  const bool peer = (a.peer.up && a.r_begin == 0) || (a.peer.dn && a.r_end == a.rows);
#define PF_TMA_DISPATCH(R, S, SLOT)                                                           \
  do {                                                                                        \
    if (peer) {                                                                               \
      if (d.stencil == 0) return tma_launch_one<0, R, S, true>(st, 4 + SLOT, a, grid, s);     \
      if (d.stencil == 5) return tma_launch_one<5, R, S, true>(st, 4 + SLOT, a, grid, s);     \
      return tma_launch_one<9, R, S, true>(st, 4 + SLOT, a, grid, s);                         \
    }                                                                                         \
    if (d.stencil == 0) return tma_launch_one<0, R, S, false>(st, SLOT, a, grid, s);          \
    if (d.stencil == 5) return tma_launch_one<5, R, S, false>(st, SLOT, a, grid, s);          \
    return tma_launch_one<9, R, S, false>(st, SLOT, a, grid, s);                              \
  } while (0)
  // ring shapes: rows per stage x stages (x 4128 B per row)
  if (shape == 1) PF_TMA_DISPATCH(8, 3, 1);   //  99 KB: 2 CTAs/SM
  if (shape == 2) PF_TMA_DISPATCH(2, 8, 2);   //  66 KB: 3 CTAs/SM, finer stages
  if (shape == 3) PF_TMA_DISPATCH(4, 6, 3);   //  99 KB: 2 CTAs/SM, deeper
  PF_TMA_DISPATCH(4, 4, 0);                   //  66 KB: 3 CTAs/SM
#undef PF_TMA_DISPATCH
}

// Launch over the tile rows [ty0, ty1) with the next tick's deposits riding along: CTA = deposit tile
// (kDepRows rows x kTmaCols columns x one channel), ring shape 4 x 4. The caller has checked
// W % 4 == 0, W >= 128, stencil != 0. with_peer: the launch contains the strip's first / last row and
// neighbours are attached (pf_strip_tick steps the whole strip in this one launch: the tiles of the edge bands
// simply have no passengers).
static int tma_stencil_launch_dep(TmaStencilState* st, const PFDev& d, StencilArgs a, int ty0, int ty1, bool with_peer,
                                  cudaStream_t s) {
  a.r_begin = ty0 * kDepRows;
  a.r_end = ty1 * kDepRows < d.rows ? ty1 * kDepRows : d.rows;
  a.rows_per_cta = kDepRows;
  a.dep_ty0 = ty0;
  a.dep_ny = (d.rows + kDepRows - 1) / kDepRows;
  const bool peer = with_peer && ((a.peer.up && a.r_begin == 0) || (a.peer.dn && a.r_end == a.rows));
  if (!peer) {
    a.peer.up = nullptr;
    a.peer.dn = nullptr;
  }
  dim3 grid((d.W + kTmaCols - 1) / kTmaCols, ty1 - ty0, d.C);
  if (peer) {
    if (d.stencil == 5) return tma_launch_one<5, 4, 4, true, true>(st, 7, a, grid, s);
    return tma_launch_one<9, 4, 4, true, true>(st, 7, a, grid, s);
  }
  if (d.stencil == 5) return tma_launch_one<5, 4, 4, false, true>(st, 8, a, grid, s);
  return tma_launch_one<9, 4, 4, false, true>(st, 8, a, grid, s);
}
