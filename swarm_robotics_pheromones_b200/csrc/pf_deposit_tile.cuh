// pf_deposit_tile.cuh — K2 fused into K1: the NEXT tick's deposits ride on the stencil's output.
//
// A deposit is a read-modify-write of one cell at a random address: 64 B fetched and 32 B written
// back per agent, row activations all over the HBM stacks — as expensive as the rest of the agent
// work of the tick. But the tick order is deposit(t+1) right after stencil(t), the stencil writes
// every cell of the new buffer anyway, and the amounts of deposit(t+1) are known before stencil(t)
// starts (the move kernel of tick t has just produced them). So:
//
//   move kernel   (k_agents_step, which produces the next tick's deposits anyway) drops every
//                 depositing agent as ONE 32-bit record (cell within the tile, which of the three
//                 reference amounts) into the bin of the stencil tile (channel, 64 rows, 1024 columns
//                 — one stencil CTA) AND stage group (the 4 output rows one ring stage produces) that
//                 will write its cell;
//   k_stencil_tma<.., DEP> the spare lanes of its producer warp read the bin of stage group g once the
//                 consumer warps have stored that group's rows, add up the records per cell in a small
//                 shared-memory hash table, and send each cell's total as one reduction (performed by
//                 the L2) to the cell a few microseconds after it was stored, while the line is still
//                 in L2: nxt[cell] = fadd(stencil value, total) — bit for bit what a separate deposit
//                 kernel after the stencil would have written, at none of its DRAM cost.
//
// No grouping kernel in between (round 1 had one: 0.36 ms of barriers per tick at config 3). What makes
// that possible: the reference's amounts are small integers (1, 20, 5: sup:232-234), so the total of a
// cell is EXACT in fp32 whatever the order of the additions (a bin holds at most kTileGroupCap records
// of at most 2^15 each: every partial sum is an integer below 2^24) and equals the oracle's in-order
// sum. Configurations with other amounts (PFDev::dep_exact == 0) never take this path: they keep the
// separate deposit kernels with their array-order sums.
//
// All or nothing per tick: if any bin overflows (agents crowding into one tile), *flag is set, the
// stencil leaves its output alone and the next tick's regular deposit kernels run (they are launched
// behind `only_if = flag` and otherwise return at once).
#pragma once
#include "pf_deposit_owner.cuh"

constexpr int kTileRows = 64;    // stencil rows per CTA on this path
constexpr int kTileCols = 1024;  // = kTmaCols
// The stencil streams its rows in stages of 4 (ring shape 4 x 4); stage j produces the output rows
// [4j - 2, 4j + 2) of the tile: records are binned by that stage.
constexpr int kTileStageRows = 4;
constexpr int kTileGroups = (kTileRows + 2 + kTileStageRows - 1) / kTileStageRows;  // 17
constexpr int kTileGroupCap = 160;  // records per (tile, stage group); mean 64 at one robot per 64 cells
constexpr int kTileHash = 256;      // shared-memory hash slots of the stencil's deposit lanes (>= cap / 0.63)
constexpr uint32_t kTileNoRec = 0xffffffffu;
static_assert(kTileHash >= kTileGroupCap + kTileGroupCap / 2 && (kTileHash & (kTileHash - 1)) == 0, "hash size");

// record: bits 0-9 column within the tile, bits 10-15 row within the tile, bits 16-17 amount
// (0 = intensity_explore, 1 = intensity_high, 2 = intensity_homing)
__device__ __forceinline__ uint32_t tile_record(uint32_t local_cell, uint32_t amount_code) {
  return local_cell | (amount_code << 16);
}

struct TileBins {
  uint32_t* bcnt;  // [tile * kTileGroups + g] records in the bin; null: no binning
  uint32_t* brec;  // [(tile * kTileGroups + g) * kTileGroupCap + k]
  uint32_t* flag;
  int nx, ny;      // tiles along a row / along a column of one channel plane
  int ty_lo, ty_hi;  // tile rows whose deposits the stencil takes (all of them on one GPU; on a strip
                     // the interior ones that the stencil writes after the move kernel)
};

// Called by every thread of `act` (converged); bin < 0 = nothing to put. One counter update per
// (warp, bin): agents that are neighbours in the array are mostly neighbours on the grid.
__device__ __forceinline__ void tile_bin_put(const TileBins& tb, unsigned act, int bin, uint32_t rec) {
  const int lane = threadIdx.x & 31;
  const unsigned peers = __match_any_sync(act, bin);
  if (bin >= 0) {
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0u;
    if (lane == leader) base = atomicAdd(tb.bcnt + bin, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const uint32_t pos = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
    if (pos < (uint32_t)kTileGroupCap) tb.brec[(long long)bin * kTileGroupCap + pos] = rec;
    else *tb.flag = 1u;
  }
}

// bin (tile and stage group) and cell-within-tile of (channel, strip-local row, column); -1 if the
// stencil does not take that tile row's deposits
__device__ __forceinline__ int tile_of(const TileBins& tb, int ch, int lr, int col, uint32_t* local) {
  const int ty = lr / kTileRows, tx = col / kTileCols;
  if (ty < tb.ty_lo || ty >= tb.ty_hi) return -1;
  const int r = lr - ty * kTileRows;
  *local = (uint32_t)(r * kTileCols + (col - tx * kTileCols));
  return ((ch * tb.ny + ty) * tb.nx + tx) * kTileGroups + (r + 2) / kTileStageRows;
}
