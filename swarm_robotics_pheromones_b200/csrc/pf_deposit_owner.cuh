// pf_deposit_owner.cuh — K2 variant "owner computes": deterministic deposit WITHOUT a global sort.
//
// After pf_agents_sort the agent arrays are ordered by cell, and agents move well under one cell per
// tick, so the arrays stay nearly ordered for dozens of ticks. At sort time the sorted key sequence is
// cut into splitters every kOwnChunk agents; CTA b owns the cells whose key lies in
// [split[b], split[b+1]) — about kOwnChunk agents, whatever the local density. Every tick
//   1. the key producer (move kernel or key kernel) also records, per array chunk of kOwnChunk
//      agents, the min and max cell key it wrote (two atomics per warp);
//   2. k_owner_bounds turns them into a prefix-max / suffix-min pair, so that "every chunk that can
//      hold a key in [lo, hi)" is a contiguous, binary-searchable chunk range — a guaranteed
//      superset, however far agents have drifted;
//   3. k_deposit_owner: each CTA scans its candidate chunks (L2-resident re-reads), collects the
//      agents that fall into its key range (any order; the array index travels with the agent),
//      groups them by cell in a shared-memory hash set, sums each shared cell's amounts in ARRAY
//      ORDER and adds the total to the field — one writer per cell and kernel, same summation
//      order as the global-sort path and the oracle. Ranges holding more than kOwnCap agents are
//      bisected (work stack); a range with more than kOwnShared agents in shared cells is ordered
//      by one block radix sort instead; a single cell holding more than kOwnCap agents is summed
//      sequentially.
// pf_step on one GPU goes one step further and lets the stencil apply the deposits
// (pf_deposit_tile.cuh); this kernel then only stands by for ticks whose tiles are too crowded.
#pragma once
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>

#include "pf_device.cuh"

constexpr int kOwnChunk = 1024;
constexpr int kOwnThreads = 256;
constexpr int kOwnItems = 8;
constexpr int kOwnCap = kOwnThreads * kOwnItems;  // 2048 agents per pass
constexpr int kOwnSlotBits = 12;
constexpr int kOwnSlots = 1 << kOwnSlotBits;  // hash slots: load factor <= 0.5
constexpr int kOwnShared = 512;               // short list of agents sharing a cell; longer -> full scan

// record the per-chunk key bounds (called by every thread of a warp whose lanes cover consecutive
// agents of ONE chunk; `sk` = sort key, or 0xffffffff when the agent does not deposit)
__device__ __forceinline__ void owner_note_bounds(uint32_t* __restrict__ cmin, uint32_t* __restrict__ cmax,
                                                  long long chunk, uint32_t sk, bool deposits) {
  uint32_t lo = deposits ? sk : 0xffffffffu;
  uint32_t hi = deposits ? sk : 0u;
  unsigned act = __activemask();
  lo = __reduce_min_sync(act, lo);
  hi = __reduce_max_sync(act, hi);
  if ((threadIdx.x & 31) == (__ffs(act) - 1) && lo != 0xffffffffu) {
    atomicMin(&cmin[chunk], lo);
    atomicMax(&cmax[chunk], hi);
  }
}

// prefix max of cmax, suffix min of cmin; one CTA; also resets cmin/cmax for the next tick
// Chunks [0, body) were full of live agents at sort time (the nearly ordered body); chunks
// [body, all) are the tail: the partly filled chunk and the empty slots behind it, where strips put
// arriving agents. Tail chunks keep their own raw bounds (an arrival's key is unrelated to its slot
// and must not poison the body's prefix/suffix bounds); every CTA tests them one by one.
__global__ void __launch_bounds__(1024)
k_owner_bounds(int all_chunks, uint32_t* __restrict__ body_chunks, uint32_t* __restrict__ cmin,
               uint32_t* __restrict__ cmax, uint32_t* __restrict__ pmax, uint32_t* __restrict__ smin,
               uint32_t* __restrict__ tail_list, uint32_t* __restrict__ only_if) {
  typedef cub::BlockScan<uint32_t, 1024> Scan;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ uint32_t carry;
  if (only_if && *only_if == 0u) {  // deposits already applied by the stencil: only reset the bounds
    for (int j = threadIdx.x; j < all_chunks; j += 1024) {
      cmin[j] = 0xffffffffu;
      cmax[j] = 0u;
    }
    return;
  }
  if (only_if && threadIdx.x == 0) only_if[1] += 1u;  // ticks whose fused deposit fell back (one CTA)
  const int nchunks = min((int)*body_chunks, all_chunks);
  if (threadIdx.x == 0) carry = 0u;
  __syncthreads();
  for (int base = 0; base < nchunks; base += 1024) {
    int j = base + threadIdx.x;
    uint32_t v = j < nchunks ? cmax[j] : 0u, out;
    Scan(tmp).InclusiveScan(v, out, cub::Max());
    uint32_t c = carry;
    out = max(out, c);
    if (j < nchunks) pmax[j] = out;
    __syncthreads();
    if (threadIdx.x == 1023) carry = out;
    __syncthreads();
  }
  if (threadIdx.x == 0) carry = 0xffffffffu;
  __syncthreads();
  for (int base = 0; base < nchunks; base += 1024) {  // reversed index
    int j = nchunks - 1 - (base + threadIdx.x);
    uint32_t v = j >= 0 ? cmin[j] : 0xffffffffu, out;
    Scan(tmp).InclusiveScan(v, out, cub::Min());
    uint32_t c = carry;
    out = min(out, c);
    if (j >= 0) smin[j] = out;
    __syncthreads();
    if (threadIdx.x == 1023) carry = out;
    __syncthreads();
  }
  // tail: raw per-chunk bounds, and the ascending list of tail chunks that hold any key at all
  // (body_chunks[1] = its length) so that the deposit CTAs never walk the empty rest of the arrays
  __shared__ uint32_t ntail;
  if (threadIdx.x == 0) ntail = 0u;
  __syncthreads();
  for (int base = nchunks; base < all_chunks; base += 1024) {
    const int j = base + threadIdx.x;
    uint32_t has = 0u;
    if (j < all_chunks) {
      pmax[j] = cmax[j];
      smin[j] = cmin[j];
      has = cmin[j] != 0xffffffffu ? 1u : 0u;
    }
    uint32_t off, total;
    Scan(tmp).ExclusiveSum(has, off, total);
    if (has) tail_list[ntail + off] = (uint32_t)j;
    __syncthreads();
    if (threadIdx.x == 0) ntail += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) body_chunks[1] = ntail;
  __syncthreads();
  for (int j = threadIdx.x; j < all_chunks; j += 1024) {
    cmin[j] = 0xffffffffu;
    cmax[j] = 0u;
  }
}

// splitters from the sorted keys of pf_agents_sort: split[b] = sorted_key[b * kOwnChunk] (offset into
// the strip's local key space), split[0] = 0, split[nb] = plane
__global__ void k_owner_splitters(const uint32_t* __restrict__ sorted_keys, long long n, int nb,
                                  uint32_t key_offset, uint32_t plane, uint32_t empty_key,
                                  uint32_t* __restrict__ split, uint32_t* __restrict__ body_chunks) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nb) return;
  if (b < nb) {  // body = the prefix of chunks whose 1024 slots all hold live agents (*body_chunks preset to 0)
    const long long last = (long long)b * kOwnChunk + kOwnChunk - 1;
    const bool full = last < n && sorted_keys[last] != empty_key;
    const long long nlast = last + kOwnChunk;
    const bool next_full = (b + 1 < nb) && nlast < n && sorted_keys[nlast] != empty_key;
    if (full && !next_full) {
      *body_chunks = (uint32_t)(b + 1);
      body_chunks[3] = sorted_keys[last];  // (about) the largest live key: density estimate for the host
    }
    if (b == 0) body_chunks[2] = sorted_keys[0];
  }
  uint32_t v;
  if (b == 0) v = 0u;
  else if (b == nb) v = plane;
  else {
    uint32_t k = sorted_keys[(long long)b * kOwnChunk];
    // empty slots sort last with a sentinel beyond every cell; keys above/below the strip clamp
    v = (k < key_offset) ? 0u : (k - key_offset > plane ? plane : k - key_offset);
  }
  split[b] = v;
}

// channel of a cell key without an integer division (C <= PF_MAX_CHANNELS = 4)
__device__ __forceinline__ uint32_t owner_channel(uint32_t key, uint32_t plane) {
  const unsigned long long k = key, pl = plane;  // 3 * plane may not fit 32 bits
  return (k >= pl ? 1u : 0u) + (k >= 2ull * pl ? 1u : 0u) + (k >= 3ull * pl ? 1u : 0u);
}

typedef cub::BlockRadixSort<unsigned long long, kOwnThreads, kOwnItems, uint32_t> OwnSort;
struct OwnHash {
  uint32_t ht[kOwnSlots];         // cell keys (0xffffffff = free slot)
  uint32_t coll[kOwnSlots / 32];  // slot holds more than one agent
  uint16_t clist[kOwnShared];     // indices of all agents of shared cells
};
union OwnScratch {
  OwnHash h;
  typename OwnSort::TempStorage sort;  // crowded ranges only
};

// Shared-memory hash set of the cell keys of one CTA's agents (open addressing, linear probing; a
// key is never 0xffffffff). A slot holds the key that claimed it; its collision bit says that more
// than one agent has that key, i.e. the cell is shared.
__device__ __forceinline__ uint32_t owner_hash_slot(uint32_t key) { return (key * 2654435761u) >> (32 - kOwnSlotBits); }

__device__ __forceinline__ void owner_hash_insert(uint32_t key, OwnHash* h) {
  uint32_t slot = owner_hash_slot(key);
  while (true) {
    const uint32_t old = atomicCAS(&h->ht[slot], 0xffffffffu, key);
    if (old == 0xffffffffu) break;
    if (old == key) {  // the slot belongs to this cell already
      atomicOr(&h->coll[slot >> 5], 1u << (slot & 31u));
      break;
    }
    slot = (slot + 1u) & (kOwnSlots - 1);
  }
}

// after all inserts (and a barrier)
__device__ __forceinline__ bool owner_hash_shared(uint32_t key, const OwnHash* h) {
  uint32_t slot = owner_hash_slot(key);
  while (h->ht[slot] != key) slot = (slot + 1u) & (kOwnSlots - 1);
  return (h->coll[slot >> 5] >> (slot & 31u)) & 1u;
}

// Hash phase of the sort-free deposit kernel. lk[0..count) = cell keys (any order). All agents of
// shared cells go on the short list h->clist (*ncl entries; more than kOwnShared means "crowded":
// the list is then incomplete). Returns this thread's mask of items (e = t + q * kOwnThreads) that
// sit in shared cells. The caller has cleared h->ht / h->coll / *ncl and synchronised; ends with a
// barrier.
__device__ __forceinline__ unsigned owner_hash_build(int count, const uint32_t* lk, OwnHash* h, int* ncl) {
  const int t = threadIdx.x;
#pragma unroll
  for (int q = 0; q < kOwnItems; ++q) {
    const int e = t + q * kOwnThreads;
    if (e < count) owner_hash_insert(lk[e], h);
  }
  __syncthreads();
  unsigned shared8 = 0u;
#pragma unroll
  for (int q = 0; q < kOwnItems; ++q) {
    const int e = t + q * kOwnThreads;
    if (e < count && owner_hash_shared(lk[e], h)) {
      shared8 |= 1u << q;
      const int pos = atomicAdd(ncl, 1);
      if (pos < kOwnShared) h->clist[pos] = (uint16_t)e;
    }
  }
  __syncthreads();
  return shared8;
}

// Item e sits in a shared cell: is it the cell's head (smallest array index)? If so, *s (preset to
// its own amount) becomes the cell's amounts summed in array order — the summation order of the
// global-sort path and of the oracle. nshared <= kOwnShared.
__device__ __forceinline__ bool owner_shared_sum(int e, int nshared, const uint32_t* lk, const uint32_t* lv,
                                                 const uint32_t* le, const OwnHash* h, float* s) {
  const uint32_t key = lk[e], mine = le[e];
  for (int i = 0; i < nshared; ++i) {
    const int c = (int)h->clist[i];
    if (lk[c] == key && le[c] < mine) return false;
  }
  uint32_t last = mine;  // next larger array index with this cell, until there is none
  while (true) {
    uint32_t best = 0xffffffffu;
    int bi = -1;
    for (int i = 0; i < nshared; ++i) {
      const int c = (int)h->clist[i];
      const uint32_t ec = le[c];
      if (lk[c] == key && ec > last && ec < best) { best = ec; bi = c; }
    }
    if (bi < 0) break;
    *s = fadd(*s, __uint_as_float(lv[bi]));
    last = best;
  }
  return true;
}

// Group `count` (0 < count <= kOwnCap) selected agents by cell and add each cell's ordered sum to the
// field. lk = keys local to [lo, lo + span) (channel * span + cell - lo), lv = amounts, le = array
// indices, in any order. The caller has cleared u->h.ht / u->h.coll and *ncl, and synchronised.
// A cell with one depositor (almost all of them at bench densities) is updated directly.
__device__ __forceinline__ void owner_group_add(const PFDev& p, uint32_t lo, uint32_t span, int count,
                                                uint32_t* lk, uint32_t* lv, const uint32_t* le,
                                                OwnScratch* u, int* ncl, float* __restrict__ field,
                                                const PeerRows& peer) {
  const int t = threadIdx.x;
  const unsigned shared8 = owner_hash_build(count, lk, &u->h, ncl);
    const int nshared = *ncl;

    if (nshared > kOwnShared) {
      // crowded range: stable order by (cell, array index) from one block radix sort of the
      // composite key, then every segment head sums its run
      __syncthreads();
      unsigned long long ck8[kOwnItems];
      uint32_t sv8[kOwnItems];
#pragma unroll
      for (int q = 0; q < kOwnItems; ++q) {
        const int e = t * kOwnItems + q;
        ck8[q] = e < count ? (((unsigned long long)lk[e] << 32) | le[e]) : ~0ull;
        sv8[q] = e < count ? lv[e] : 0u;
      }
      __syncthreads();
      const unsigned long long maxkey = (unsigned long long)p.C * span;  // one past the largest local key
      int end_bit = 1;
      while ((1ull << end_bit) <= maxkey && end_bit < 32) ++end_bit;  // all-ones padding sorts last
      OwnSort(u->sort).Sort(ck8, sv8, 0, 32 + end_bit);
      __syncthreads();
#pragma unroll
      for (int q = 0; q < kOwnItems; ++q) {
        const int e = t * kOwnItems + q;
        lk[e] = (uint32_t)(ck8[q] >> 32);
        lv[e] = sv8[q];
      }
      __syncthreads();
      for (int q = 0; q < kOwnItems; ++q) {
        const int e = t + q * kOwnThreads;
        if (e >= count) break;
        const uint32_t key = lk[e];
        if (e != 0 && lk[e - 1] == key) continue;  // not a segment head
        float s = __uint_as_float(lv[e]);
        for (int f = e + 1; f < count && lk[f] == key; ++f) s = fadd(s, __uint_as_float(lv[f]));
        const uint32_t ch = key / span;
        const uint32_t sk = lo + (key - ch * span);
        const int lr = (int)(sk / (uint32_t)p.W), col = (int)(sk - (uint32_t)lr * (uint32_t)p.W);
        float* cell = field + (long long)ch * p.chan_stride + (long long)(lr + 1) * p.pitch + col;
        const float nv = fadd(__ldcg(cell), s);
        *cell = nv;
        if (lr == 0 && peer.up) peer.up[ch * peer.up_cs + col] = nv;
        if (lr == p.rows - 1 && peer.dn) peer.dn[ch * peer.dn_cs + col] = nv;
      }
      return;
    }

    // read-modify-write of the touched cells: one writer per cell; the cell loads of a batch are
    // issued before its adds and stores
    constexpr int kBatch = 4;
#pragma unroll
    for (int q0 = 0; q0 < kOwnItems; q0 += kBatch) {
      float* cellp[kBatch];
      float ssum[kBatch], cval[kBatch];
      int erow[kBatch];
#pragma unroll
      for (int q = 0; q < kBatch; ++q) {
        const int e = t + (q0 + q) * kOwnThreads;
        cellp[q] = nullptr;
        if (e < count) {
          const uint32_t key = lk[e];
          float s = __uint_as_float(lv[e]);
          bool head = true;
          if ((shared8 >> (q0 + q)) & 1u) head = owner_shared_sum(e, nshared, lk, lv, le, &u->h, &s);
          if (head) {
            const uint32_t ch = key / span;
            const uint32_t sk = lo + (key - ch * span);
            const int lr = (int)(sk / (uint32_t)p.W), col = (int)(sk - (uint32_t)lr * (uint32_t)p.W);
            cellp[q] = field + (long long)ch * p.chan_stride + (long long)(lr + 1) * p.pitch + col;
            ssum[q] = s;
            erow[q] = lr;
          }
        }
      }
      if (p.dep_red) {
        // one writer per cell and kernel, so a reduction performed by the L2 (no round trip to the
        // SM, no partially written sector) rounds exactly like the load-add-store below. The L2
        // flushes subnormal operands; with every amount >= 1e-30 (p.dep_red) a subnormal cell
        // value is below half an ulp of the sum either way. Edge-row cells of an attached strip
        // need the new value for the neighbour's mailbox and take the load-add-store path.
#pragma unroll
        for (int q = 0; q < kBatch; ++q)
          if (cellp[q] && !((erow[q] == 0 && peer.up) || (erow[q] == p.rows - 1 && peer.dn))) {
            atomicAdd(cellp[q], ssum[q]);
            cellp[q] = nullptr;
          }
      }
#pragma unroll
      for (int q = 0; q < kBatch; ++q)
        if (cellp[q]) cval[q] = __ldcg(cellp[q]);
#pragma unroll
      for (int q = 0; q < kBatch; ++q)
        if (cellp[q]) {
          const float nv = fadd(cval[q], ssum[q]);
          *cellp[q] = nv;
          if (peer.up || peer.dn) {  // strips only: mirror edge-row cells into the neighbour's mailbox
            const long long off = cellp[q] - field;
            const long long ch = off / p.chan_stride;
            const int col = (int)((off - ch * p.chan_stride) % p.pitch);
            if (erow[q] == 0 && peer.up) peer.up[ch * peer.up_cs + col] = nv;
            if (erow[q] == p.rows - 1 && peer.dn) peer.dn[ch * peer.dn_cs + col] = nv;
          }
        }
    }
}

__global__ void __launch_bounds__(kOwnThreads, 4)
k_deposit_owner(PFDev p, long long n, int all_chunks, const uint32_t* __restrict__ body_chunks,
                const uint32_t* __restrict__ keys, const float* __restrict__ vals, uint32_t sentinel,
                const uint32_t* __restrict__ split, const uint32_t* __restrict__ pmax,
                const uint32_t* __restrict__ smin, const uint32_t* __restrict__ tail_list,
                float* __restrict__ field, PeerRows peer, const uint32_t* __restrict__ only_if,
                uint32_t handed_lo, uint32_t handed_hi, const uint32_t* __restrict__ handed_flag) {
  if (only_if && *only_if == 0u) return;  // the stencil already applied this tick's deposits (pf_deposit_tile.cuh)
  // strips: the cells [handed_lo, handed_hi) of every channel got their deposits from the stencil
  // (unless its flag says otherwise); key ranges inside them have nothing to do
  const bool handed = handed_flag && *handed_flag == 0u;
  typedef cub::BlockScan<int, kOwnThreads> Scan;
  __shared__ typename Scan::TempStorage scan_tmp;
  __shared__ OwnScratch u;
  __shared__ uint32_t lk[kOwnCap];  // local keys of the selected agents (any order)
  __shared__ uint32_t lv[kOwnCap];  // their amounts (bit patterns)
  __shared__ uint32_t le[kOwnCap];  // their array indices: the summation order inside a cell
  __shared__ uint32_t st_lo[34], st_hi[34];
  __shared__ int st_n;
  __shared__ int tcand[kOwnThreads];  // tail chunks whose bounds meet the current key range
  __shared__ int jrange[2];
  __shared__ int cnt, ncl;

  const uint32_t plane = (uint32_t)((long long)p.rows * p.W);
  const int t = threadIdx.x;
  const int lane = t & 31;
  const int nchunks = min((int)*body_chunks, all_chunks);  // body chunks: binary-searchable bounds
  const bool vec_ok = ((((uintptr_t)keys) | ((uintptr_t)vals)) & 15u) == 0u;
  // one key range per CTA (grid = all_chunks), or a small grid striding over the ranges (the
  // stand-by launch behind the fused deposit, which nearly always returns at the first line)
  for (int b = blockIdx.x; b < all_chunks; b += gridDim.x) {
  __syncthreads();
  if (t == 0) {
    st_lo[0] = split[b];
    st_hi[0] = split[b + 1];
    st_n = (split[b] < split[b + 1]) ? 1 : 0;
    if (handed && split[b] >= handed_lo && split[b + 1] <= handed_hi) st_n = 0;
  }
  __syncthreads();

  while (true) {
    __syncthreads();
    if (st_n == 0) break;
    const uint32_t lo = st_lo[st_n - 1], hi = st_hi[st_n - 1];
    __syncthreads();
    if (t == 0) { st_n -= 1; cnt = 0; ncl = 0; }
    // candidate chunks: first j with pmax[j] >= lo ... last j with smin[j] < hi. A 32-ary search by
    // one warp each (two dependent load rounds for 16 K chunks instead of fourteen) run side by side.
    if (t < 64) {
      const int w = t >> 5, l = t & 31;
      const uint32_t* arr = w == 0 ? pmax : smin;
      const uint32_t bound = w == 0 ? lo : hi;
      int a = 0, z = nchunks;  // invariant: answer (first index with arr[idx] >= bound) lies in [a, z]
      while (z - a > 0) {
        const int step = (z - a + 31) / 32;
        const int idx = a + l * step;           // probe 32 equally spaced positions
        const bool ge = idx < z ? (arr[idx] >= bound) : true;
        const unsigned m = __ballot_sync(0xffffffffu, ge);
        if (m == 0u) { a = a + 31 * step + 1; continue; }  // every probe below the bound
        const int first = __ffs(m) - 1;         // first probe that is >= bound (probes at idx >= z count)
        const int na = first == 0 ? a : a + (first - 1) * step + 1;
        const int nz = min(z, a + first * step);
        if (first == 0) { z = a; break; }
        a = na; z = nz;
      }
      if (l == 0) jrange[w] = z;
    } else {  // the other warps clear the hash table meanwhile
      for (int i = t - 64; i < kOwnSlots; i += kOwnThreads - 64) u.h.ht[i] = 0xffffffffu;
      if (t - 64 < kOwnSlots / 32) u.h.coll[t - 64] = 0u;
    }
    __syncthreads();
    const int j0 = jrange[0], j1 = jrange[1] - 1;
    const uint32_t span = hi - lo;
    // candidates: body chunks j0..j1 (phase 0), then, 256 at a time, the non-empty tail chunks whose
    // own bounds meet [lo, hi). The agents of [lo, hi) are appended to lk/lv/le in any order (one
    // shared-memory atomic per warp and chunk, no barrier): their array index travels with them.
    const int nt = (int)body_chunks[1];
    const int nphase = 1 + (nt + kOwnThreads - 1) / kOwnThreads;
    const int nbody = (j1 - j0 + 1 > 0 ? j1 - j0 + 1 : 0);
    for (int phase = 0; phase < nphase; ++phase) {
      int ncand = nbody;
      if (phase > 0) {
        const int it = (phase - 1) * kOwnThreads + t;
        int has = 0, jt = 0;
        if (it < nt) {
          jt = (int)tail_list[it];
          has = (pmax[jt] >= lo && smin[jt] < hi) ? 1 : 0;
        }
        int off, total;
        __syncthreads();
        Scan(scan_tmp).ExclusiveSum(has, off, total);
        if (has) tcand[off] = jt;
        __syncthreads();
        ncand = total;
      }
      for (int c = 0; c < ncand; ++c) {
        const int j = phase == 0 ? j0 + c : tcand[c];
        const long long e0 = (long long)j * kOwnChunk + 4 * t;
        uint32_t k4[4], v4[4];
        if (vec_ok && e0 + 3 < n) {
          const uint4 kk = __ldg(reinterpret_cast<const uint4*>(keys + e0));
          const uint4 vv = __ldg(reinterpret_cast<const uint4*>(vals + e0));
          k4[0] = kk.x; k4[1] = kk.y; k4[2] = kk.z; k4[3] = kk.w;
          v4[0] = vv.x; v4[1] = vv.y; v4[2] = vv.z; v4[3] = vv.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const long long e = e0 + q;
            k4[q] = e < n ? keys[e] : sentinel;
            v4[q] = e < n ? __float_as_uint(vals[e]) : 0u;
          }
        }
        bool sel[4];
        unsigned m4[4];
        int tot = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t key = k4[q];
          const uint32_t ch = owner_channel(key, plane);
          const uint32_t sk = key - ch * plane;
          sel[q] = key != sentinel && sk >= lo && sk < hi;
          k4[q] = ch * span + (sk - lo);
          m4[q] = __ballot_sync(0xffffffffu, sel[q]);
          tot += __popc(m4[q]);
        }
        if (tot) {  // warp-uniform
          int base = 0;
          if (lane == 0) base = atomicAdd(&cnt, tot);
          base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int w = base + __popc(m4[q] & ((1u << lane) - 1u));
            if (sel[q] && w < kOwnCap) {
              lk[w] = k4[q];
              lv[w] = v4[q];
              le[w] = (uint32_t)(e0 + q);
            }
            base += __popc(m4[q]);
          }
        }
      }
    }  // phases
    __syncthreads();
    const int count = cnt;

    if (count == 0) continue;
    if (count > kOwnCap) {
      if (span > 1u) {  // bisect the key range; both halves are handled by this CTA
        if (t == 0) {
          const uint32_t mid = lo + span / 2u;
          st_lo[st_n] = mid; st_hi[st_n] = hi; st_n += 1;
          st_lo[st_n] = lo;  st_hi[st_n] = mid; st_n += 1;
        }
        continue;
      }
      // one cell holds more than kOwnCap agents: sum them sequentially, in array order
      for (int ch = 0; ch < p.C; ++ch) {
        float s = 0.0f;
        bool any = false;
        const uint32_t want = (uint32_t)ch * plane + lo;
        for (int phase = 0; phase < nphase; ++phase) {
          int ncand = nbody;
          if (phase > 0) {
            const int it = (phase - 1) * kOwnThreads + t;
            int has = 0, jt = 0;
            if (it < nt) {
              jt = (int)tail_list[it];
              has = (pmax[jt] >= lo && smin[jt] < hi) ? 1 : 0;
            }
            int off, total;
            __syncthreads();
            Scan(scan_tmp).ExclusiveSum(has, off, total);
            if (has) tcand[off] = jt;
            __syncthreads();
            ncand = total;
          }
          for (int c = 0; c < ncand; ++c) {
            const int j = phase == 0 ? j0 + c : tcand[c];
            const long long e0 = (long long)j * kOwnChunk + 4 * t;
            int nsel = 0;
            uint32_t v4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const long long e = e0 + q;
              if (e < n && keys[e] == want) v4[nsel++] = __float_as_uint(vals[e]);
            }
            int off, total;
            Scan(scan_tmp).ExclusiveSum(nsel, off, total);
            __syncthreads();
            for (int q = 0; q < nsel; ++q) lv[off + q] = v4[q];
            __syncthreads();
            if (t == 0)
              for (int q = 0; q < total; ++q) {
                float a = __uint_as_float(lv[q]);
                s = any ? fadd(s, a) : a;
                any = true;
              }
            __syncthreads();
          }
        }  // phases
        if (t == 0 && any) {
          const int lr = (int)(lo / (uint32_t)p.W), col = (int)(lo - (uint32_t)lr * (uint32_t)p.W);
          float* cell = field + (long long)ch * p.chan_stride + (long long)(lr + 1) * p.pitch + col;
          const float nv = fadd(*cell, s);
          *cell = nv;
          if (lr == 0 && peer.up) peer.up[ch * peer.up_cs + col] = nv;
          if (lr == p.rows - 1 && peer.dn) peer.dn[ch * peer.dn_cs + col] = nv;
        }
      }
      continue;
    }

    owner_group_add(p, lo, span, count, lk, lv, le, &u, &ncl, field, peer);
  }
  }  // key ranges of this CTA
}

