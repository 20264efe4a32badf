// pf_deposit_tile.cuh — K2 fused into K1: the NEXT tick's deposits ride on the stencil's output.
//
// A deposit is a read-modify-write of one cell at a random address: 64 B fetched and 32 B written
// back per agent, row activations all over the HBM stacks — as expensive as the rest of the agent
// work of the tick. But the tick order is deposit(t+1) right after stencil(t), the stencil writes
// every cell of the new buffer anyway, and the amounts of deposit(t+1) are known before stencil(t)
// starts (the move kernel of tick t has just produced them). So:
//
//   k_tile_bin    one streaming pass: every depositing agent goes into the bin of the stencil tile
//                 (channel, 64 rows, 1024 columns — one stencil CTA) that will write its cell;
//   k_tile_group  one CTA per tile: group the bin by cell (shared-memory hash, array-order sums —
//                 the same rounding as every other deposit path), then lay the (column, sum) pairs
//                 out by (row, consumer warp) of the stencil CTA;
//   k_stencil_tma<.., DEP> pulls its tile's lists into shared memory with three bulk copies; its
//                 producer warp sends each sum as a reduction (performed by the L2) to the cell
//                 a few microseconds after the consumer warps stored it, while the line is still
//                 in L2: nxt[cell] = fadd(stencil value, sum) — bit for bit what a separate
//                 deposit kernel after the stencil would have written (amounts >= 1e-30, see
//                 PFDev::dep_red), at none of its DRAM cost.
//
// All or nothing per tick: if any bin or list overflows (agents crowding into one tile), *flag is
// set, the stencil leaves its output alone and the next tick's regular deposit kernels run (they
// are launched behind `only_if = flag` and otherwise return at once).
#pragma once
#include "pf_deposit_owner.cuh"

constexpr int kTileRows = 64;    // stencil rows per CTA on this path
constexpr int kTileCols = 1024;  // = kTmaCols
constexpr int kTileWarps = 8;    // = kTmaWarps: column slices of 128
constexpr int kTileLists = kTileRows * kTileWarps;
constexpr int kTileEnt = 1408;   // distinct deposited cells a tile can hand to the stencil
// one tile's block in global memory, copied piecewise into the stencil CTA's shared memory:
//   uint16 start[kTileLists + 1 (+7 pad)]   list i = entries [start[i], start[i+1])
//   uint16 col[kTileEnt]                    row << 10 | column, within the tile
//   float  sum[kTileEnt]
constexpr int kTilePtrBytes = (kTileLists + 8) * 2;
constexpr int kTileColBytes = kTileEnt * 2;
constexpr int kTileBlkBytes = kTilePtrBytes + kTileColBytes + kTileEnt * 4;
static_assert(kTilePtrBytes % 16 == 0 && kTileColBytes % 16 == 0 && kTileBlkBytes % 16 == 0, "bulk copy alignment");

struct TileGeom {
  int nx, ny;      // tiles per row of tiles / per channel plane column
  int W;           // columns
  uint32_t plane;  // cells per channel
};

__global__ void __launch_bounds__(256)
k_tile_bin(long long n, TileGeom g, const uint32_t* __restrict__ keys, const float* __restrict__ vals,
           uint32_t sentinel, uint32_t* __restrict__ bcnt, uint32_t* __restrict__ bkey,
           float* __restrict__ bval, uint32_t* __restrict__ bidx, uint32_t* __restrict__ flag) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < n; i0 += stride) {
    const long long i = i0 + threadIdx.x;
    uint32_t key = sentinel;
    float v = 0.0f;
    if (i < n) {
      key = __ldcs(keys + i);
      v = __ldcs(vals + i);
    }
    int tile = -1;
    uint32_t local = 0u;
    if (key != sentinel) {
      const uint32_t ch = owner_channel(key, g.plane);
      const uint32_t sk = key - ch * g.plane;
      const uint32_t lr = sk / (uint32_t)g.W, col = sk - lr * (uint32_t)g.W;
      const uint32_t ty = lr / kTileRows, tx = col / kTileCols;
      tile = (int)((ch * (uint32_t)g.ny + ty) * (uint32_t)g.nx + tx);
      local = (lr - ty * kTileRows) * kTileCols + (col - tx * kTileCols);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, tile);  // one counter update per (warp, tile)
    if (tile >= 0) {
      const int leader = __ffs(peers) - 1;
      uint32_t base = 0u;
      if (lane == leader) base = atomicAdd(bcnt + tile, (uint32_t)__popc(peers));
      base = __shfl_sync(peers, base, leader);
      const uint32_t pos = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
      if (pos < (uint32_t)kOwnCap) {
        const long long w = (long long)tile * kOwnCap + pos;
        bkey[w] = local;
        bval[w] = v;
        bidx[w] = (uint32_t)i;
      } else {
        *flag = 1u;
      }
    }
  }
}

// per tile: tcount = number of (cell, sum) entries; block = start[] / col[] / sum[] as laid out above,
// list = row * 8 + (column >> 7)
__global__ void __launch_bounds__(kOwnThreads, 4)
k_tile_group(uint32_t* __restrict__ bcnt, const uint32_t* __restrict__ bkey, const float* __restrict__ bval,
             const uint32_t* __restrict__ bidx, uint32_t* __restrict__ flag, uint32_t* __restrict__ tcount,
             unsigned char* __restrict__ tblk) {
  typedef cub::BlockScan<uint32_t, kOwnThreads> Scan;
  __shared__ OwnScratch u;
  __shared__ typename Scan::TempStorage scan_tmp;
  __shared__ uint32_t lk[kOwnCap], lv[kOwnCap], le[kOwnCap];
  __shared__ uint32_t lcnt[kTileLists], lpos[kTileLists];
  __shared__ int ncl;
  const int tile = blockIdx.x, t = threadIdx.x;
  const uint32_t raw = bcnt[tile];
  __syncthreads();
  if (t == 0) {
    bcnt[tile] = 0u;
    ncl = 0;
    if (raw == 0u) tcount[tile] = 0u;
  }
  if (raw == 0u || raw > (uint32_t)kOwnCap) return;  // empty, or overflowed in k_tile_bin (flag is set)
  const int count = (int)raw;
  const long long base = (long long)tile * kOwnCap;
#pragma unroll
  for (int q = 0; q < kOwnItems; ++q) {
    const int e = t + q * kOwnThreads;
    if (e < count) {
      lk[e] = __ldcs(bkey + base + e);
      lv[e] = __float_as_uint(__ldcs(bval + base + e));
      le[e] = __ldcs(bidx + base + e);
    }
  }
  for (int i = t; i < kOwnSlots; i += kOwnThreads) u.h.ht[i] = 0xffffffffu;
  if (t < kOwnSlots / 32) u.h.coll[t] = 0u;
  for (int i = t; i < kTileLists; i += kOwnThreads) { lcnt[i] = 0u; lpos[i] = 0u; }
  __syncthreads();
  const unsigned shared8 = owner_hash_build(count, lk, &u, &ncl);
  const int nshared = ncl;
  if (nshared > kOwnShared) {  // crowded tile: leave this tick to the regular deposit kernels
    if (t == 0) *flag = 1u;
    return;
  }
  // heads: the cell's ordered sum replaces the head's own amount; count them per list
  unsigned head8 = 0u;
#pragma unroll
  for (int q = 0; q < kOwnItems; ++q) {
    const int e = t + q * kOwnThreads;
    if (e < count) {
      float s = __uint_as_float(lv[e]);
      bool head = true;
      if ((shared8 >> q) & 1u) head = owner_shared_sum(e, nshared, lk, lv, le, &u, &s);
      if (head) {
        head8 |= 1u << q;
        if ((shared8 >> q) & 1u) lv[e] = __float_as_uint(s);  // read by nobody else: members skip heads
        const uint32_t key = lk[e];
        atomicAdd(&lcnt[(key >> 10) * kTileWarps + ((key & 1023u) >> 7)], 1u);
      }
    }
  }
  __syncthreads();
  // exclusive scan of the list sizes (2 lists per thread)
  uint32_t c2[2] = {lcnt[2 * t], lcnt[2 * t + 1]}, o2[2], total;
  Scan(scan_tmp).ExclusiveSum(c2, o2, total);
  if (total > (uint32_t)kTileEnt) {  // block-uniform
    if (t == 0) *flag = 1u;
    return;
  }
  lpos[2 * t] = o2[0];
  lpos[2 * t + 1] = o2[1];
  unsigned char* blk = tblk + (long long)tile * kTileBlkBytes;
  uint16_t* start = reinterpret_cast<uint16_t*>(blk);
  uint16_t* ecol = reinterpret_cast<uint16_t*>(blk + kTilePtrBytes);
  float* esum = reinterpret_cast<float*>(blk + kTilePtrBytes + kTileColBytes);
  reinterpret_cast<uint32_t*>(start)[t] = o2[0] | (o2[1] << 16);
  if (t == 0) {
    start[kTileLists] = (uint16_t)total;
    tcount[tile] = total;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < kOwnItems; ++q) {
    const int e = t + q * kOwnThreads;
    if ((head8 >> q) & 1u) {
      const uint32_t key = lk[e];
      const uint32_t w = atomicAdd(&lpos[(key >> 10) * kTileWarps + ((key & 1023u) >> 7)], 1u);
      ecol[w] = (uint16_t)key;  // row << 10 | column
      esum[w] = __uint_as_float(lv[e]);
    }
  }
}
