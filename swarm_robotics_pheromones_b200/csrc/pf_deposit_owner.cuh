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
//   3. k_deposit_owner: each CTA scans its candidate chunks (L2-resident re-reads), compacts the
//      agents that fall into its key range in ARRAY ORDER, block-radix-sorts them by cell (stable),
//      sums each cell's amounts in that order and adds the total to the field — one writer per cell,
//      no atomics on the field, same summation order as the global-sort path and the oracle.
//      Ranges holding more than kOwnCap agents are bisected (work stack); a single cell holding more
//      than kOwnCap agents is summed sequentially.
#pragma once
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>

#include "pf_device.cuh"

constexpr int kOwnChunk = 1024;
constexpr int kOwnThreads = 256;
constexpr int kOwnItems = 8;
constexpr int kOwnCap = kOwnThreads * kOwnItems;  // 2048 agents per sort

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
               uint32_t* __restrict__ tail_list) {
  typedef cub::BlockScan<uint32_t, 1024> Scan;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ uint32_t carry;
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

__global__ void __launch_bounds__(kOwnThreads, 4)
k_deposit_owner(PFDev p, long long n, int all_chunks, const uint32_t* __restrict__ body_chunks,
                const uint32_t* __restrict__ keys, const float* __restrict__ vals, uint32_t sentinel,
                const uint32_t* __restrict__ split, const uint32_t* __restrict__ pmax,
                const uint32_t* __restrict__ smin, const uint32_t* __restrict__ tail_list,
                float* __restrict__ field, PeerRows peer) {
  typedef cub::BlockRadixSort<uint32_t, kOwnThreads, kOwnItems, uint32_t> Sort;
  typedef cub::BlockScan<int, kOwnThreads> Scan;
  __shared__ union {
    typename Sort::TempStorage sort;
    typename Scan::TempStorage scan;
  } tmp;
  __shared__ uint32_t lk[kOwnCap];  // local keys of the selected agents, array order / sorted order
  __shared__ uint32_t lv[kOwnCap];  // their amounts (bit patterns)
  __shared__ uint32_t st_lo[34], st_hi[34];
  __shared__ int st_n;
  __shared__ int tcand[kOwnThreads];  // tail chunks whose bounds meet the current key range
  __shared__ int jrange[2];

  const uint32_t plane = (uint32_t)((long long)p.rows * p.W);
  const int b = blockIdx.x;
  const int t = threadIdx.x;
  const int nchunks = min((int)*body_chunks, all_chunks);  // body chunks: binary-searchable bounds
  if (t == 0) {
    st_lo[0] = split[b];
    st_hi[0] = split[b + 1];
    st_n = (split[b] < split[b + 1]) ? 1 : 0;
  }
  __syncthreads();

  while (true) {
    __syncthreads();
    if (st_n == 0) break;
    const uint32_t lo = st_lo[st_n - 1], hi = st_hi[st_n - 1];
    __syncthreads();
    if (t == 0) st_n -= 1;
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
    }
    __syncthreads();
    const int j0 = jrange[0], j1 = jrange[1] - 1;
    __syncthreads();
    const uint32_t span = hi - lo;
    int count = 0;
    // candidates: body chunks j0..j1 (phase 0), then, 256 at a time, the non-empty tail chunks whose
    // own bounds meet [lo, hi) (ascending: array order is preserved)
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
      Scan(tmp.scan).ExclusiveSum(has, off, total);
      if (has) tcand[off] = jt;
      __syncthreads();
      ncand = total;
    }
    for (int c = 0; c < ncand; ++c) {
      const int j = phase == 0 ? j0 + c : tcand[c];
      const long long e0 = (long long)j * kOwnChunk + 4 * t;
      uint32_t k4[4], v4[4];
      int sel[4], nsel = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const long long e = e0 + q;
        uint32_t key = sentinel;
        if (e < n) key = keys[e];
        k4[q] = key;
        sel[q] = 0;
        if (key != sentinel) {
          const uint32_t ch = key / plane;
          const uint32_t sk = key - ch * plane;
          if (sk >= lo && sk < hi) {
            sel[q] = 1;
            k4[q] = ch * span + (sk - lo);
            v4[q] = __float_as_uint(vals[e]);
            ++nsel;
          }
        }
      }
      int off, total;
      Scan(tmp.scan).ExclusiveSum(nsel, off, total);
      __syncthreads();
      if (count + total <= kOwnCap) {
        int w = count + off;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (sel[q]) { lk[w] = k4[q]; lv[w] = v4[q]; ++w; }
      }
      count += total;
    }
    }  // phases
    __syncthreads();

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
          Scan(tmp.scan).ExclusiveSum(has, off, total);
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
          Scan(tmp.scan).ExclusiveSum(nsel, off, total);
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

    // ---- stable block radix sort of the selected agents by cell, then ordered segmented sum
    uint32_t sk8[kOwnItems], sv8[kOwnItems];
#pragma unroll
    for (int q = 0; q < kOwnItems; ++q) {
      const int e = t * kOwnItems + q;
      sk8[q] = e < count ? lk[e] : 0xffffffffu;
      sv8[q] = e < count ? lv[e] : 0u;
    }
    __syncthreads();
    unsigned long long maxkey = (unsigned long long)p.C * span;  // one past the largest local key
    int end_bit = 1;
    while ((1ull << end_bit) <= maxkey && end_bit < 32) ++end_bit;  // all-ones padding sorts last
    Sort(tmp.sort).Sort(sk8, sv8, 0, end_bit);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kOwnItems; ++q) {
      const int e = t * kOwnItems + q;
      lk[e] = sk8[q];
      lv[e] = sv8[q];
    }
    __syncthreads();
    // read-modify-write of the touched cells: every thread owns up to kOwnItems segment heads; all
    // their cell loads are issued first (distinct cells: one writer each), then the adds and stores
    constexpr int kBatch = 4;
    for (int q0 = 0; q0 < kOwnItems; q0 += kBatch) {
      float* cellp[kBatch];
      float ssum[kBatch], cval[kBatch];
#pragma unroll
      for (int q = 0; q < kBatch; ++q) {
        const int e = t + (q0 + q) * kOwnThreads;
        cellp[q] = nullptr;
        if (e < count) {
          const uint32_t key = lk[e];
          if (e == 0 || lk[e - 1] != key) {  // segment head
            float s = __uint_as_float(lv[e]);
            for (int f = e + 1; f < count && lk[f] == key; ++f) s = fadd(s, __uint_as_float(lv[f]));
            const uint32_t ch = key / span;
            const uint32_t sk = lo + (key - ch * span);
            const int lr = (int)(sk / (uint32_t)p.W), col = (int)(sk - (uint32_t)lr * (uint32_t)p.W);
            cellp[q] = field + (long long)ch * p.chan_stride + (long long)(lr + 1) * p.pitch + col;
            ssum[q] = s;
          }
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
            const uint32_t key = lk[t + (q0 + q) * kOwnThreads];
            const uint32_t ch = key / span;
            const uint32_t sk = lo + (key - ch * span);
            const int lr = (int)(sk / (uint32_t)p.W), col = (int)(sk - (uint32_t)lr * (uint32_t)p.W);
            if (lr == 0 && peer.up) peer.up[ch * peer.up_cs + col] = nv;
            if (lr == p.rows - 1 && peer.dn) peer.dn[ch * peer.dn_cs + col] = nv;
          }
        }
    }
  }
}
