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
//   k_tile_group  one CTA per tile, one warp per two stage groups: the warp adds up its bins' records per
//                 cell in a warp-private shared-memory hash table (no block barrier on the way; round 1
//                 grouped a tile's records in one block-wide hash set behind five barriers: 0.36 ms per
//                 tick at config 3), ONE barrier for the offsets of the groups, and the (cell, total)
//                 entries go out grouped by stage;
//   k_stencil_tma<.., DEP> pulls its tile's list into shared memory with two bulk copies; its
//                 producer warp sends each total as a reduction (performed by the L2) to the cell
//                 a few microseconds after the consumer warps stored it, while the line is still
//                 in L2: nxt[cell] = fadd(stencil value, total) — bit for bit what a separate
//                 deposit kernel after the stencil would have written (amounts >= 1e-30, see
//                 PFDev::dep_red), at none of its DRAM cost. (Adding the totals up inside the stencil
//                 kernel as well was tried in round 2: 31 lanes or a tenth warp doing the hashing cost
//                 the stencil 0.6 ms — profiles/r02_notes.md.)
//
// What makes the warp-private grouping possible: the reference's amounts are small integers (1, 20, 5:
// sup:232-234), so the total of a cell is EXACT in fp32 whatever the order of the additions (a bin holds
// at most kTileGroupCap records of at most 2^15 each: every partial sum is an integer below 2^24) and
// equals the oracle's in-order sum. Configurations with other amounts (PFDev::dep_exact == 0) never take
// this path: they keep the separate deposit kernels with their array-order sums.
//
// All or nothing per tick: if any bin or list overflows (agents crowding into one tile), *flag is
// set, the stencil leaves its output alone and the next tick's regular deposit kernels run (they
// are launched behind `only_if = flag` and otherwise return at once).
#pragma once
#include "pf_deposit_owner.cuh"

constexpr int kTileRows = 64;    // stencil rows per CTA on this path
constexpr int kTileCols = 1024;  // = kTmaCols
// The stencil streams its rows in stages of 4 (ring shape 4 x 4); stage j produces the output rows
// [4j - 2, 4j + 2) of the tile: records are binned by that stage.
constexpr int kTileStageRows = 4;
constexpr int kTileGroups = (kTileRows + 2 + kTileStageRows - 1) / kTileStageRows;  // 17
constexpr int kTileGroupCap = 160;  // records per (tile, stage group); mean 64 at one robot per 64 cells
constexpr int kTileHash = 256;      // hash slots per stage group in k_tile_group (>= cap / 0.63)
constexpr int kTileEnt = 1408;      // distinct deposited cells a tile can hand to the stencil
// one tile's block in global memory, copied into the stencil CTA's shared memory (as much as is used):
//   uint16 start[24]        group g = entries [start[g], start[g+1]), g < kTileGroups
//   uint16 cell[kTileEnt]   row << 10 | column, within the tile
//   float  sum[kTileEnt]
constexpr int kTilePtrBytes = 48;
constexpr int kTileColBytes = kTileEnt * 2;
constexpr int kTileBlkBytes = kTilePtrBytes + kTileColBytes + kTileEnt * 4;
static_assert(kTileGroups + 1 <= kTilePtrBytes / 2 && kTileGroups <= 32, "group starts");
static_assert(kTileColBytes % 16 == 0 && kTileBlkBytes % 16 == 0, "bulk copy alignment");
constexpr uint32_t kTileNoRec = 0xffffffffu;
static_assert(kTileHash >= kTileGroupCap + kTileGroupCap / 2 && (kTileHash & (kTileHash - 1)) == 0, "hash size");

// record: bits 0-9 column within the tile, bits 10-15 row within the tile, bits 16-17 amount
// (0 = intensity_explore, 1 = intensity_high, 2 = intensity_homing)
__device__ __forceinline__ uint32_t tile_record(uint32_t local_cell, uint32_t amount_code) {
  return local_cell | (amount_code << 16);
}

// Crowded tiles (config 4's nest: more than one robot per cell): a tile whose bin overflowed is marked DENSE and
// gets a side buffer of one fp32 total per cell (64 x 1024 x 4 B = 256 KB out of a small pool). From the next tick
// on the move kernel adds that tile's amounts straight into the side buffer (reductions performed by the L2: exact,
// the amounts are small integers), and the tile's stencil CTA adds the buffer to its output rows as it stores them
// and clears it. No sort, no grouping for those tiles; marks are reset by pf_agents_sort.
constexpr int kTileCells = kTileRows * kTileCols;

struct TileBins {
  uint32_t* bcnt;  // [tile * kTileGroups + g] records in the bin; null: no binning
  uint32_t* brec;  // [(tile * kTileGroups + g) * kTileGroupCap + k]
  uint32_t* flag;
  int* dense;      // [tile] side buffer slot, -1 = sparse (bins), -2 = being marked
  float* side;     // [slot][kTileCells]
  uint32_t* pool;  // [0] slots handed out so far
  int pool_cap;
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
    if (pos < (uint32_t)kTileGroupCap) {
      tb.brec[(long long)bin * kTileGroupCap + pos] = rec;
    } else {  // this tick falls back to the regular deposit kernels; from the next one on the tile is dense
      *tb.flag = 1u;
      const int tile = bin / kTileGroups;
      if (tb.dense && atomicCAS(tb.dense + tile, -1, -2) == -1) {
        const uint32_t slot = atomicAdd(tb.pool, 1u);
        tb.dense[tile] = slot < (uint32_t)tb.pool_cap ? (int)slot : -1;  // pool exhausted: stays sparse (and falls back)
      }
    }
  }
}

// dense tile: the amount goes straight into the tile's side buffer; returns true if it did
__device__ __forceinline__ bool tile_side_put(const TileBins& tb, int bin, uint32_t local, float amt) {
  if (bin < 0 || !tb.dense) return false;
  const int slot = tb.dense[bin / kTileGroups];
  if (slot < 0) return false;
  atomicAdd(tb.side + (long long)slot * kTileCells + local, amt);
  return true;
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

// per tile: tcount = number of (cell, total) entries, block = start[] / cell[] / sum[] as laid out above
constexpr int kGroupWarps = (kTileGroups + 1) / 2;  // 9: warp w takes stage groups 2w and 2w + 1
constexpr int kGroupThreads = kGroupWarps * 32;
constexpr int kGroupPerLane = (kTileGroupCap + 31) / 32;

__global__ void __launch_bounds__(kGroupThreads, 6)
k_tile_group(uint32_t* __restrict__ bcnt, const uint32_t* __restrict__ brec, uint32_t* __restrict__ flag,
             uint32_t* __restrict__ tcount, unsigned char* __restrict__ tblk, float amt0, float amt1, float amt2,
             int* __restrict__ dense, uint32_t* __restrict__ pool, int pool_cap) {
  __shared__ uint32_t hkey[kTileGroups + 1][kTileHash];
  __shared__ float hsum[kTileGroups + 1][kTileHash];
  __shared__ uint32_t gcnt[32], gstart[32], total_s;
  const int tile = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  uint32_t n[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int g = 2 * warp + j;
    n[j] = g < kTileGroups ? bcnt[(long long)tile * kTileGroups + g] : 0u;
  }
  if (t < 32) gcnt[t] = 0u;
  if (!__syncthreads_or((n[0] | n[1]) != 0u)) {  // nothing deposited into this tile
    if (t == 0) tcount[tile] = 0u;
    return;
  }
  uint32_t slot_of[2][kGroupPerLane];  // hash slots this lane claimed (= distinct cells it reports)
  uint32_t mine[2] = {0u, 0u};
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int g = 2 * warp + j;
#pragma unroll
    for (int k = 0; k < kGroupPerLane; ++k) slot_of[j][k] = kTileNoRec;
    if (n[j] == 0u) continue;                    // warp-uniform
    if (lane == 0) bcnt[(long long)tile * kTileGroups + g] = 0u;  // the next move kernel starts from zero
    if (n[j] > (uint32_t)kTileGroupCap) {        // overflowed while binning (the flag is set already)
      if (lane == 0) *flag = 1u;
      n[j] = (uint32_t)kTileGroupCap;
    }
    uint32_t rec[kGroupPerLane];
#pragma unroll
    for (int k = 0; k < kGroupPerLane; ++k) {
      const uint32_t e = (uint32_t)lane + 32u * k;
      rec[k] = e < n[j] ? __ldcs(brec + ((long long)tile * kTileGroups + g) * kTileGroupCap + e) : kTileNoRec;
    }
    for (int s = lane; s < kTileHash; s += 32) { hkey[g][s] = kTileNoRec; hsum[g][s] = 0.0f; }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kGroupPerLane; ++k) {
      if (rec[k] == kTileNoRec) continue;
      const uint32_t cell = rec[k] & 0xffffu, code = rec[k] >> 16;
      const float amt = code == 0u ? amt0 : (code == 1u ? amt1 : amt2);
      uint32_t slot = (cell * 2654435761u) >> 24;
      while (true) {
        const uint32_t old = atomicCAS(&hkey[g][slot], kTileNoRec, cell);
        if (old == kTileNoRec) { slot_of[j][k] = slot; mine[j] += 1u; }
        if (old == kTileNoRec || old == cell) { atomicAdd(&hsum[g][slot], amt); break; }
        slot = (slot + 1u) & (uint32_t)(kTileHash - 1);
      }
    }
  }
  // rank of this lane's cells inside each of its groups; the groups' sizes
  uint32_t rank[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    uint32_t inc = mine[j];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    rank[j] = inc - mine[j];
    if (lane == 31 && 2 * warp + j < kTileGroups) gcnt[2 * warp + j] = inc;
  }
  __syncthreads();  // the only block barrier behind the binning: group sizes -> offsets
  unsigned char* blk = tblk + (long long)tile * kTileBlkBytes;
  if (t < 32) {
    const uint32_t c = t < kTileGroups ? gcnt[t] : 0u;
    uint32_t inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (t >= d) inc += o;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    gstart[t] = inc - c;
    if (total <= (uint32_t)kTileEnt && t <= kTileGroups)
      reinterpret_cast<uint16_t*>(blk)[t] = (uint16_t)(inc - c);  // start[kTileGroups] = total
    if (t == 0) {
      total_s = total;
      if (total > (uint32_t)kTileEnt) {  // too many distinct cells for one list: fall back this tick, dense from the next
        *flag = 1u;
        if (dense && atomicCAS(dense + tile, -1, -2) == -1) {
          const uint32_t slot = atomicAdd(pool, 1u);
          dense[tile] = slot < (uint32_t)pool_cap ? (int)slot : -1;
        }
      } else {
        tcount[tile] = total;
        // about to outgrow the list: dense from the next tick on, without a fall-back in between. (The margin is
        // small on purpose: a tile of the uniform configs holds 1024 +- 32 cells, and a dense tile costs the move
        // kernel an atomic per robot and the stencil CTA a second read stream.)
        if (dense && total > (uint32_t)(kTileEnt - kTileEnt / 16) && atomicCAS(dense + tile, -1, -2) == -1) {
          const uint32_t slot = atomicAdd(pool, 1u);
          dense[tile] = slot < (uint32_t)pool_cap ? (int)slot : -1;
        }
      }
    }
  }
  __syncthreads();
  if (total_s > (uint32_t)kTileEnt) return;
  uint16_t* ecell = reinterpret_cast<uint16_t*>(blk + kTilePtrBytes);
  float* esum = reinterpret_cast<float*>(blk + kTilePtrBytes + kTileColBytes);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int g = 2 * warp + j;
    if (g >= kTileGroups) continue;
    uint32_t w = gstart[g] + rank[j];
#pragma unroll
    for (int k = 0; k < kGroupPerLane; ++k) {
      if (slot_of[j][k] == kTileNoRec) continue;
      ecell[w] = (uint16_t)hkey[g][slot_of[j][k]];  // row << 10 | column
      esum[w] = hsum[g][slot_of[j][k]];
      ++w;
    }
  }
}
