// pf_deposit_tile.cuh — K2 fused into K1: the NEXT tick's deposits ride on the stencil's output.
//
// A deposit is a read-modify-write of one cell at a random address: 64 B fetched and 32 B written
// back per agent, row activations all over the HBM stacks — as expensive as the rest of the agent
// work of the tick. But the tick order is deposit(t+1) right after stencil(t), the stencil writes
// every cell of the new buffer anyway, and the amounts of deposit(t+1) are known before stencil(t)
// starts (the move kernel of tick t has just produced them). So:
//
//   move kernel   (k_agents_step, which produces the next tick's deposits anyway) puts every
//                 depositing agent into the bin of the stencil tile (channel, 64 rows, 1024
//                 columns — one stencil CTA) that will write its cell;
//   k_tile_group  one CTA per tile: group the bin by cell (shared-memory hash, array-order sums —
//                 the same rounding as every other deposit path) into a list of (cell, sum) pairs;
//   k_stencil_tma<.., DEP> pulls its tile's list into shared memory with two bulk copies; its
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
constexpr int kTileEnt = 1408;   // distinct deposited cells a tile can hand to the stencil
// The stencil streams its rows in stages of 4 (ring shape 4 x 4); stage j produces the output rows
// [4j - 2, 4j + 2) of the tile. Entries are grouped by stage so that the stencil finds the ones for
// the rows it has just stored without searching.
constexpr int kTileStageRows = 4;
constexpr int kTileGroups = (kTileRows + 2 + kTileStageRows - 1) / kTileStageRows;  // 17
// one tile's block in global memory, copied into the stencil CTA's shared memory (as much as is used):
//   uint16 start[24]        group g = entries [start[g], start[g+1]), g < kTileGroups
//   uint16 cell[kTileEnt]   row << 10 | column, within the tile
//   float  sum[kTileEnt]
constexpr int kTilePtrBytes = 48;
constexpr int kTileColBytes = kTileEnt * 2;
constexpr int kTileBlkBytes = kTilePtrBytes + kTileColBytes + kTileEnt * 4;
static_assert(kTileGroups + 1 <= kTilePtrBytes / 2 && kTileGroups <= 32, "group starts");
static_assert(kTileColBytes % 16 == 0 && kTileBlkBytes % 16 == 0, "bulk copy alignment");

// bins of the stencil tiles: kOwnCap records (cell within the tile, amount, array index) each
struct TileBins {
  uint32_t* bcnt;  // null: no binning
  uint32_t* bkey;
  float* bval;
  uint32_t* bidx;
  uint32_t* flag;
  int nx, ny;      // tiles along a row / along a column of one channel plane
  int ty_lo, ty_hi;  // tile rows whose deposits the stencil takes (all of them on one GPU; on a strip
                     // the interior ones that the stencil writes after the move kernel)
};

// Called by every thread of `act` (converged); tile < 0 = nothing to put. One counter update per
// (warp, tile): agents that are neighbours in the array are mostly neighbours on the grid.
__device__ __forceinline__ void tile_bin_put(const TileBins& tb, unsigned act, int tile, uint32_t local, float v,
                                             uint32_t index) {
  const int lane = threadIdx.x & 31;
  const unsigned peers = __match_any_sync(act, tile);
  if (tile >= 0) {
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0u;
    if (lane == leader) base = atomicAdd(tb.bcnt + tile, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const uint32_t pos = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
    if (pos < (uint32_t)kOwnCap) {
      const long long w = (long long)tile * kOwnCap + pos;
      tb.bkey[w] = local;
      tb.bval[w] = v;
      tb.bidx[w] = index;
    } else {
      *tb.flag = 1u;
    }
  }
}

// tile and cell-within-tile of (channel, strip-local row, column); -1 if the stencil does not take
// that tile row's deposits
__device__ __forceinline__ int tile_of(const TileBins& tb, int ch, int lr, int col, uint32_t* local) {
  const int ty = lr / kTileRows, tx = col / kTileCols;
  if (ty < tb.ty_lo || ty >= tb.ty_hi) return -1;
  *local = (uint32_t)((lr - ty * kTileRows) * kTileCols + (col - tx * kTileCols));
  return (ch * tb.ny + ty) * tb.nx + tx;
}

// per tile: tcount = number of (cell, sum) entries, block = start[] / cell[] / sum[] as laid out above
__global__ void __launch_bounds__(kOwnThreads, 5)
k_tile_group(uint32_t* __restrict__ bcnt, const uint32_t* __restrict__ bkey, const float* __restrict__ bval,
             const uint32_t* __restrict__ bidx, uint32_t* __restrict__ flag, uint32_t* __restrict__ tcount,
             unsigned char* __restrict__ tblk) {
  __shared__ OwnHash hs;
  __shared__ uint32_t lk[kOwnCap], lv[kOwnCap], le[kOwnCap];
  __shared__ int ncl;
  __shared__ uint32_t nent, gcnt[32], gpos[32];
  const int tile = blockIdx.x, t = threadIdx.x;
  const long long base = (long long)tile * kOwnCap;
  // the first half of the bin is fetched together with its fill count (one round trip less)
  constexpr int kEarly = kOwnItems / 2;
  uint32_t k_[kEarly], v_[kEarly], e_[kEarly];
  const uint32_t raw = bcnt[tile];
#pragma unroll
  for (int q = 0; q < kEarly; ++q) {
    const long long e = base + t + q * kOwnThreads;
    k_[q] = __ldcs(bkey + e);
    v_[q] = __float_as_uint(__ldcs(bval + e));
    e_[q] = __ldcs(bidx + e);
  }
  for (int i = t; i < kOwnSlots; i += kOwnThreads) hs.ht[i] = 0xffffffffu;
  if (t < kOwnSlots / 32) hs.coll[t] = 0u;
  if (t < 32) gcnt[t] = 0u;
  __syncthreads();  // every thread has read the count; the hash set is clear
  if (t == 0) {
    bcnt[tile] = 0u;
    ncl = 0;
    if (raw == 0u) tcount[tile] = 0u;
  }
  if (raw == 0u || raw > (uint32_t)kOwnCap) return;  // empty, or overflowed while binning (flag is set)
  const int count = (int)raw;
  // records -> shared memory, keys -> hash set, in one phase (the set holds keys, not indices)
#pragma unroll
  for (int q = 0; q < kOwnItems; ++q) {
    const int e = t + q * kOwnThreads;
    if (e < count) {
      uint32_t k, v, x;
      if (q < kEarly) {
        k = k_[q]; v = v_[q]; x = e_[q];
      } else {
        k = __ldcs(bkey + base + e);
        v = __float_as_uint(__ldcs(bval + base + e));
        x = __ldcs(bidx + base + e);
      }
      lk[e] = k; lv[e] = v; le[e] = x;
      owner_hash_insert(k, &hs);
    }
  }
  __syncthreads();
  // pass 1a: unshared cells are heads with their own amount: count them per stage group; agents
  // of shared cells go on the short list
  const int lane = t & 31;
  unsigned head8 = 0u, shared8 = 0u;
#pragma unroll
  for (int q = 0; q < kOwnItems; ++q) {
    const int e = t + q * kOwnThreads;
    int g = -1;
    if (e < count) {
      const uint32_t key = lk[e];
      if (owner_hash_shared(key, &hs)) {
        shared8 |= 1u << q;
        const int pos = atomicAdd(&ncl, 1);
        if (pos < kOwnShared) hs.clist[pos] = (uint16_t)e;
      } else {
        head8 |= 1u << q;
        g = (int)((key >> 10) + 2u) / kTileStageRows;
      }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, g);
    if (g >= 0 && lane == __ffs(peers) - 1) atomicAdd(&gcnt[g], (uint32_t)__popc(peers));
  }
  __syncthreads();
  const int nshared = ncl;
  if (nshared > kOwnShared) {  // crowded tile: leave this tick to the regular deposit kernels
    if (t == 0) *flag = 1u;
    return;
  }
  // pass 1b: the head of every shared cell gets the cell's ordered sum
  if (nshared > 0) {  // block-uniform
#pragma unroll
    for (int q = 0; q < kOwnItems; ++q) {
      if ((shared8 >> q) & 1u) {
        const int e = t + q * kOwnThreads;
        float s = __uint_as_float(lv[e]);
        if (owner_shared_sum(e, nshared, lk, lv, le, &hs, &s)) {
          lv[e] = __float_as_uint(s);  // read by nobody else: members skip heads
          head8 |= 1u << q;
          atomicAdd(&gcnt[(int)((lk[e] >> 10) + 2u) / kTileStageRows], 1u);
        }
      }
    }
  }
  __syncthreads();
  unsigned char* blk = tblk + (long long)tile * kTileBlkBytes;
  if (t < 32) {  // exclusive scan of the group sizes
    const uint32_t c = t < kTileGroups ? gcnt[t] : 0u;
    uint32_t inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (t >= d) inc += o;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    gpos[t] = inc - c;
    if (total <= (uint32_t)kTileEnt) {
      if (t <= kTileGroups) reinterpret_cast<uint16_t*>(blk)[t] = (uint16_t)(inc - c);  // start[kTileGroups] = total
    }
    if (t == 0) {
      nent = total;
      if (total > (uint32_t)kTileEnt) *flag = 1u;
      else tcount[tile] = total;
    }
  }
  __syncthreads();
  if (nent > (uint32_t)kTileEnt) return;
  // pass 2: place the heads
  uint16_t* ecell = reinterpret_cast<uint16_t*>(blk + kTilePtrBytes);
  float* esum = reinterpret_cast<float*>(blk + kTilePtrBytes + kTileColBytes);
#pragma unroll
  for (int q = 0; q < kOwnItems; ++q) {
    const int e = t + q * kOwnThreads;
    const bool head = (head8 >> q) & 1u;
    const uint32_t key = head ? lk[e] : 0u;
    const int g = head ? (int)((key >> 10) + 2u) / kTileStageRows : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, g);
    if (head) {
      const int leader = __ffs(peers) - 1;
      uint32_t w = 0u;
      if (lane == leader) w = atomicAdd(&gpos[g], (uint32_t)__popc(peers));
      w = __shfl_sync(peers, w, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
      ecell[w] = (uint16_t)key;  // row << 10 | column
      esum[w] = __uint_as_float(lv[e]);
    }
  }
}
