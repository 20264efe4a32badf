// pf_agents.cuh — K2 (deterministic deposit) and K3 (sense -> decide -> move) kernels.
#pragma once
#include "pf_device.cuh"
#include "pf_deposit_tile.cuh"

// ------------------------------------------------------------------------------------------
// K2a: sort keys. key = local linear cell index ((ch*rows + lr)*W + col); agents that do not
// deposit get the sentinel (= number of cells), which sorts behind every real key.
// ------------------------------------------------------------------------------------------

// Deposit rule of deposit_pheromone (reference sup:291-358): gate on has_moved (sup:308-310),
// SEARCHING: counter+1, every `high_every`-th deposit is 20.0 else 1.0 (sup:312-318);
// HOMING: 5.0 (sup:346-348). The counter lives in meta bits 8..31.
__global__ void __launch_bounds__(256, 8)
k_deposit_keys_agents(PFDev p, long long n, const float* __restrict__ x,
                      const float* __restrict__ y, uint32_t* __restrict__ meta,
                      uint32_t* __restrict__ keys, float* __restrict__ vals, uint32_t sentinel,
                      unsigned long long* __restrict__ counters, uint32_t* __restrict__ cmin,
                      uint32_t* __restrict__ cmax, TileBins handed, uint32_t* __restrict__ handed_flag, int mode,
                      int late_may_skip) {
  // handed.bcnt != null: the last move kernel handed the deposits of the tile rows [ty_lo, ty_hi)
  // to the stencil (counter already bumped). *handed_flag == 0: the stencil applied them, nothing
  // left to do for those agents; != 0: it could not (a crowded tile) and they are deposited here
  // after all, with the amount their bumped counter says.
  // mode (pf_strip_tick's deposit in two parts, see there): 0 = everybody in one pass, as described;
  // 1 = EARLY, at the end of the tick that handed over, beside its stencil launch: only the robots that were NOT
  // handed over (their cells, the edge bands, are stepped already); the handed ones are left alone, mark and all;
  // 2 = LATE, at the start of the next tick: only the handed robots — nothing at all if the stencil applied their
  // deposits and the coming move kernel clears the marks (late_may_skip), else marks cleared and, on a raised flag,
  // deposited after all. The two parts never share a cell (pf_tile_rows), so two passes round like one.
  const bool handed_done = handed.bcnt && *handed_flag == 0u;
  if (mode == 2 && handed_done && late_may_skip) return;
  if (mode == 0 && handed.bcnt && !handed_done && blockIdx.x == 0 && threadIdx.x == 0) handed_flag[1] += 1u;  // pf_tile_stats
  // grid-stride over a fixed grid (a multiple of the SM count): the deposit counter costs one
  // global atomic per CTA instead of one per warp (same-address atomics serialise in L2)
  __shared__ unsigned int s_dep;
  if (threadIdx.x == 0) s_dep = 0;
  __syncthreads();
  unsigned int ndep = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    bool dep = false;
    uint32_t mt = meta[i];
    uint32_t st = mt & PF_META_STATE_MASK;
    uint32_t key = sentinel;
    float amt = 0.0f;
    // the move kernel marked the robots whose deposit it handed to the stencil (PF_META_HANDED): the
    // mark, not the position, says who they are — poses may have been replaced since (pf_agents_observe)
    const bool was_handed = (mt & PF_META_HANDED) != 0u;
    const bool skip = (mode == 1 && was_handed) || (mode == 2 && !was_handed);  // the other pass's robot
    if (was_handed && !skip) {
      mt &= ~PF_META_HANDED;
      meta[i] = mt;
    }
    if (!skip && (st == PF_STATE_SEARCHING || st == PF_STATE_HOMING) && ((mt & PF_META_MOVED) || was_handed)) {
      int r, q;
      bool inside = pf_cell(p, x[i], y[i], r, q);
      int ch = p.dep_ch[st];
      const bool mine = inside && r >= p.row0 && r < p.row1;
      if (!(was_handed && handed_done)) {
        uint32_t cnt = ((mt >> PF_META_COUNT_SHIFT) + (was_handed ? 0u : 1u)) & PF_META_COUNT_MASK;
        amt = (st == PF_STATE_SEARCHING)
                  ? ((cnt % p.high_every == 0u) ? p.i_high : p.i_explore)
                  : p.i_homing;
        if (!was_handed) {
          meta[i] = (mt & 0xFFu) | (cnt << PF_META_COUNT_SHIFT);
          dep = true;
        }
        if (mine) key = (uint32_t)(((long long)ch * p.rows + (r - p.row0)) * p.W + q);
      }
    }
    keys[i] = key;
    vals[i] = amt;
    ndep += dep ? 1u : 0u;
    if (cmin) {
      const uint32_t plane = (uint32_t)((long long)p.rows * p.W);
      owner_note_bounds(cmin, cmax, i / kOwnChunk, key % plane, key != sentinel);
    }
  }
  ndep = __reduce_add_sync(0xffffffffu, ndep);
  if ((threadIdx.x & 31) == 0 && ndep) atomicAdd(&s_dep, ndep);
  __syncthreads();
  if (threadIdx.x == 0 && s_dep) atomicAdd(&counters[1], (unsigned long long)s_dep);
}

// generic positions/amounts (pf_deposit)
__global__ void __launch_bounds__(256)
k_deposit_keys_generic(PFDev p, long long n, const float* __restrict__ x,
                       const float* __restrict__ y, const float* __restrict__ amount,
                       const uint8_t* __restrict__ channel, uint32_t* __restrict__ keys,
                       float* __restrict__ vals, uint32_t sentinel) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r, q;
  bool inside = pf_cell(p, x[i], y[i], r, q);
  int ch = channel ? (int)channel[i] : 0;
  uint32_t key = sentinel;
  if (inside && r >= p.row0 && r < p.row1 && ch < p.C)
    key = (uint32_t)(((long long)ch * p.rows + (r - p.row0)) * p.W + q);
  keys[i] = key;
  vals[i] = amount[i];
}

// ------------------------------------------------------------------------------------------
// K2b: warp-aggregated segmented sum over the sorted (key, amount) pairs, in sorted (= agent)
// order, one store per touched cell. A lane whose key differs from its predecessor's is a segment
// head; heads add the following lanes' amounts one by one through shuffles (fixed order), and keep
// reading global memory when the run continues past the warp. No atomics, no races: exactly one
// thread writes each touched cell.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_deposit_segsum(PFDev p, long long n, const uint32_t* __restrict__ keys,
                 const float* __restrict__ vals, uint32_t sentinel, float* __restrict__ field,
                 PeerRows peer) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  uint32_t key = sentinel;
  float val = 0.0f;
  if (i < n) {
    key = keys[i];
    val = vals[i];
  }
  uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
  if (lane == 0) prev = (i > 0 && i < n) ? keys[i - 1] : sentinel;
  const bool head = (key != sentinel) && (i == 0 || key != prev);
  float s = val;
  bool open = head;  // still extending this head's run
  for (int d = 1; d < 32; ++d) {
    uint32_t nk = __shfl_down_sync(0xffffffffu, key, d);
    float nv = __shfl_down_sync(0xffffffffu, val, d);
    open = open && (lane + d < 32) && (nk == key);
    if (open) s = fadd(s, nv);
    if (!__any_sync(0xffffffffu, open)) break;
  }
  // the run reaches the end of the warp iff the last lane holds the same key: continue from
  // global memory, still in sorted order
  const uint32_t last_key = __shfl_sync(0xffffffffu, key, 31);
  if (head && last_key == key) {
    long long j = i - lane + 32;
    while (j < n && keys[j] == key) {
      s = fadd(s, vals[j]);
      ++j;
    }
  }
  if (head) {
    const long long plane = (long long)p.rows * p.W;
    const int ch = (int)(key / plane);
    const long long rem = key - (long long)ch * plane;
    const int lr = (int)(rem / p.W);
    const int col = (int)(rem - (long long)lr * p.W);
    float* cell = field + (long long)ch * p.chan_stride + (long long)(lr + 1) * p.pitch + col;
    const float nv = fadd(*cell, s);
    *cell = nv;
    if (lr == 0 && peer.up) peer.up[ch * peer.up_cs + col] = nv;  // keep the neighbour's ghost row current
    if (lr == p.rows - 1 && peer.dn) peer.dn[ch * peer.dn_cs + col] = nv;
  }
}

// ------------------------------------------------------------------------------------------
// K3: sense -> ordered arg-max -> PID -> wheel speeds -> diff-drive move, one thread per agent.
// Mirrors oracle/pf_oracle.c po_agent_step_one op for op.
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ int pf_nearest_food(const PFDev& p, float x, float y, float& d2_out) {
  int best = 0;
  float bd = 0.0f;
  for (int k = 0; k < p.n_food; ++k) {
    float dx = fsub(p.food_x[k], x);
    float dy = fsub(p.food_y[k], y);
    float d2 = fadd(fmul(dx, dx), fmul(dy, dy));
    if (k == 0 || d2 < bd) {
      bd = d2;
      best = k;
    }
  }
  d2_out = bd;
  return best;
}

// the five DRUL values (reference sup:379-417 in dict order sup:390); NaN = direction absent
__device__ __forceinline__ void pf_sense5(const PFDev& p, const float* __restrict__ f, float x,
                                          float y, float theta, uint32_t st, float v[5]) {
  const int ch = p.sen_ch[st & 3u];
  float s = 0.0f, co = 1.0f;
  if (p.sense_mode == PF_SENSE_HEADING) pf_sincos(theta, s, co);
  int r0, q0;
  pf_cell(p, x, y, r0, q0);
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    int r, q;
    bool inside;
    if (p.sense_mode == PF_SENSE_WORLD) {
      r = r0 + p.drow[d];
      q = q0 + p.dcol[d];
      inside = (r >= 0 && r < p.H && q >= 0 && q < p.W);
      if (d == PF_DIR_CURRENT) inside = true;
    } else if (d == PF_DIR_CURRENT) {
      r = r0;
      q = q0;
      inside = true;
    } else {
      float ux, uy;
      if (d == PF_DIR_FRONT) { ux = co; uy = s; }
      else if (d == PF_DIR_BACK) { ux = -co; uy = -s; }
      else if (d == PF_DIR_LEFT) { ux = -s; uy = co; }
      else { ux = s; uy = -co; }
      float px = fadd(x, fmul(p.sense_dist, ux));
      float py = fadd(y, fmul(p.sense_dist, uy));
      inside = pf_cell(p, px, py, r, q);
    }
    if (inside && r >= p.row0 - 1 && r <= p.row1)
      v[d] = __ldg(f + pf_addr(p, ch, r, q));
    else
      v[d] = __int_as_float(0x7fc00000);
  }
}

__device__ __forceinline__ int pf_argmax5(const float v[5]) {
  int dir = -1;
  float best = 0.0f;
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    if (v[d] != v[d]) continue;
    if (dir < 0 || v[d] > best) {
      best = v[d];
      dir = d;
    }
  }
  return dir;
}

__global__ void __launch_bounds__(256)
k_sense(PFDev p, const float* __restrict__ f, long long n, const float* __restrict__ x,
        const float* __restrict__ y, const float* __restrict__ theta,
        const uint32_t* __restrict__ meta, float* __restrict__ out5) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v[5];
  uint32_t st = meta ? (meta[i] & PF_META_STATE_MASK) : PF_STATE_SEARCHING;
  pf_sense5(p, f, x[i], y[i], theta ? theta[i] : 0.0f, st, v);
#pragma unroll
  for (int d = 0; d < 5; ++d) out5[i * 5 + d] = v[d];
}

// counters layout (unsigned long long): [0] ticks [1] deposits [2..7] decisions [8] grabbed
// [9] delivered [10] migrated_out [11] migrated_in [12] migrate_dropped
__global__ void __launch_bounds__(256)
k_agents_step(PFDev p, const float* __restrict__ f, uint32_t flags, long long n,
              float* __restrict__ ax, float* __restrict__ ay, float* __restrict__ ath,
              float* __restrict__ awl, float* __restrict__ awr, uint32_t* __restrict__ ameta,
              uint8_t* __restrict__ dec_out, unsigned long long* __restrict__ counters,
              uint32_t* __restrict__ mig_cnt, uint32_t* __restrict__ next_keys,
              float* __restrict__ next_vals, uint32_t sentinel, uint32_t* __restrict__ cmin,
              uint32_t* __restrict__ cmax, TileBins tb) {
  __shared__ unsigned int hist[9];  // 0..5 decisions, 6 grabbed, 7 delivered, 8 deposits (next tick)
  if (threadIdx.x < 9) hist[threadIdx.x] = 0;
  __syncthreads();
  // grid-stride over a fixed grid: the decision histogram is flushed once per CTA
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    uint32_t mt = ameta[i];
    uint32_t st = mt & PF_META_STATE_MASK;
    // a mark left by the LATE deposit pass that had nothing to do (pf_strip_tick): its tick's deposit is over
    mt &= ~PF_META_HANDED;
    if (st == PF_STATE_EMPTY) {
      if (dec_out) dec_out[i] = PF_DEC_NONE;
      if (next_keys) { next_keys[i] = sentinel; next_vals[i] = 0.0f; }
    } else {
      float x = ax[i], y = ay[i], th = ath[i], wl = awl[i], wr = awr[i];
      int dec = PF_DEC_NONE;
      bool delivered = false;
      if (st == PF_STATE_IDLE) {
        wl = 0.0f;
        wr = 0.0f;
      } else if (!(flags & PF_STEP_SENSE)) {
        wl = p.bl;  // Braitenberg base speeds only (reference sup:721-725, 736-744)
        wr = p.br;
      } else if (st == PF_STATE_SEARCHING) {
        float v5[5];
        pf_sense5(p, f, x, y, th, st, v5);
        int dir = pf_argmax5(v5);
        float d2;
        int k = pf_nearest_food(p, x, y, d2);
        float turn = pf_pid_turn(p, x, y, p.food_x[k], p.food_y[k]);
        if (dir == PF_DIR_LEFT) {  // sup:455-458
          float m = fsub(p.vmax, fabsf(turn));
          wl = m > 0.0f ? m : 0.0f;
          wr = p.vmax;
          dec = PF_DEC_LEFT;
        } else if (dir == PF_DIR_RIGHT) {  // sup:459-462
          float m = fsub(p.vmax, fabsf(turn));
          wr = m > 0.0f ? m : 0.0f;
          wl = p.vmax;
          dec = PF_DEC_RIGHT;
        } else {  // sup:463-466
          wl = p.cruise;
          wr = p.cruise;
          dec = PF_DEC_STRAIGHT;
        }
        wl = fadd(wl, p.bl);  // sup:468-470
        wr = fadd(wr, p.br);
      } else {  // HOMING, sup:475-514
        float dx = fsub(p.nest_x, x);
        float dy = fsub(p.nest_y, y);
        float d = fsqrt(fadd(fmul(dx, dx), fmul(dy, dy)));
        float dcm = rintf(fmul(d, 100.0f));
        bool reached = (x == p.nest_x && y == p.nest_y) || (dcm >= p.home_lo && dcm <= p.home_hi);
        if (!reached) {
          float turn = pf_pid_turn(p, x, y, p.nest_x, p.nest_y);
          if (d <= p.home_radius) {
            float rt = rintf(fmul(turn, 100.0f));
            if (rt > -200.0f && rt < 0.0f) {
              wl = 0.0f;
              wr = p.vmax;
              dec = PF_DEC_LEFT;
            } else if (rt < -200.0f) {
              wr = 0.0f;
              wl = p.vmax;
              dec = PF_DEC_RIGHT;
            } else {
              dec = PF_DEC_KEEP;
            }
          } else {
            wl = p.cruise;
            wr = p.cruise;
            dec = PF_DEC_STRAIGHT;
          }
          wl = fadd(wl, p.bl);
          wr = fadd(wr, p.br);
        } else {
          wl = 0.0f;
          wr = 0.0f;
          dec = PF_DEC_STOP;
          delivered = true;
        }
      }

      uint32_t moved = mt & PF_META_MOVED;
      if (flags & PF_STEP_MOVE) {
        float v = fmul(fmul(p.wheel_r, fadd(wl, wr)), 0.5f);
        float om = fdiv(fmul(p.wheel_r, fsub(wr, wl)), p.axle);
        float nth = pf_wrap(fadd(th, fmul(om, p.dt)));
        float sn, cs;
        pf_sincos(nth, sn, cs);
        float step = fmul(v, p.dt);
        float nx = fadd(x, fmul(step, cs));
        float ny = fadd(y, fmul(step, sn));
        if (nx < p.xmin) {
          nx = fadd(p.xmin, fsub(p.xmin, nx));
          nth = pf_wrap(fsub(PF_PI_F, nth));
        } else if (nx > p.xmax) {
          nx = fsub(p.xmax, fsub(nx, p.xmax));
          nth = pf_wrap(fsub(PF_PI_F, nth));
        }
        if (ny < p.ymin) {
          ny = fadd(p.ymin, fsub(p.ymin, ny));
          nth = -nth;
        } else if (ny > p.ymax) {
          ny = fsub(p.ymax, fsub(ny, p.ymax));
          nth = -nth;
        }
        nx = fminf(fmaxf(nx, p.xmin), p.xmax);
        ny = fminf(fmaxf(ny, p.ymin), p.ymax);
        float ddx = fsub(nx, x), ddy = fsub(ny, y);
        float dist = fsqrt(fadd(fmul(ddx, ddx), fmul(ddy, ddy)));
        moved = (dist >= p.move_thr) ? PF_META_MOVED : 0u;  // has_moved, sup:263-273
        x = nx;
        y = ny;
        th = nth;
      }

      uint32_t carry = mt & PF_META_CARRY;
      if (p.fsm && (flags & PF_STEP_SENSE)) {
        if (st == PF_STATE_SEARCHING) {
          float d2;
          pf_nearest_food(p, x, y, d2);
          if (d2 <= p.food_r2) {
            st = PF_STATE_HOMING;
            carry = PF_META_CARRY;
            atomicAdd(&hist[6], 1u);
          }
        } else if (st == PF_STATE_HOMING && delivered) {
          st = PF_STATE_SEARCHING;
          carry = 0;
          atomicAdd(&hist[7], 1u);
        }
      }
      {  // warp-aggregated histogram update: one shared atomic per distinct decision per warp
        unsigned act = __activemask();
        unsigned peers = __match_any_sync(act, dec);
        if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[dec], (unsigned)__popc(peers));
      }
      ax[i] = x;
      ay[i] = y;
      ath[i] = th;
      awl[i] = wl;
      awr[i] = wr;
      uint32_t nm = (mt & ~(PF_META_STATE_MASK | PF_META_MOVED | PF_META_DEC_MASK | PF_META_CARRY)) |
                    st | moved | ((uint32_t)dec << PF_META_DEC_SHIFT) | carry;
      if (next_keys || tb.bcnt) {
        // the next tick's deposit rule, evaluated here on the state it would read (identical to
        // k_deposit_keys_agents; pf_step only asks for it when that deposit is certain to run).
        // next_keys: every agent's key is written. tb.bcnt: agents whose cell lies in a tile the
        // stencil takes care of are handed over through the tile's bin (pf_deposit_tile.cuh); on a
        // strip (no next_keys) the others are left to the next tick's key kernel, untouched.
        uint32_t key = sentinel, local = 0u, code = 0u;
        float amt = 0.0f;
        int tile = -1;  // bin of the stencil tile and stage group that writes the agent's cell
        if ((st == PF_STATE_SEARCHING || st == PF_STATE_HOMING) && moved) {
          int r, q;
          const bool inside = pf_cell(p, x, y, r, q);
          const bool mine = inside && r >= p.row0 && r < p.row1;
          if (mine && tb.bcnt) tile = tile_of(tb, p.dep_ch[st], r - p.row0, q, &local);
          if (next_keys || tile >= 0) {
            uint32_t cnt = ((nm >> PF_META_COUNT_SHIFT) + 1u) & PF_META_COUNT_MASK;
            code = (st == PF_STATE_SEARCHING) ? ((cnt % p.high_every == 0u) ? 1u : 0u) : 2u;
            amt = code == 0u ? p.i_explore : (code == 1u ? p.i_high : p.i_homing);
            nm = (nm & 0xFFu) | (cnt << PF_META_COUNT_SHIFT);
            if (!next_keys) nm |= PF_META_HANDED;  // strips: the next key kernel skips this robot (or deposits it after all)
            if (mine) key = (uint32_t)(((long long)p.dep_ch[st] * p.rows + (r - p.row0)) * p.W + q);
            atomicAdd(&hist[8], 1u);
          }
        }
        if (next_keys) {
          next_keys[i] = key;
          next_vals[i] = amt;
          if (cmin) {
            const uint32_t plane = (uint32_t)((long long)p.rows * p.W);
            owner_note_bounds(cmin, cmax, i / kOwnChunk, key % plane, key != sentinel);
          }
        }
        if (tb.bcnt) {
          if (tile_side_put(tb, tile, local, amt)) tile = -1;  // dense tile: already added to its side buffer
          tile_bin_put(tb, __activemask(), tile, tile_record(local, code));
        }
      }
      ameta[i] = nm;
      if (dec_out) dec_out[i] = (uint8_t)dec;
      if (mig_cnt) {  // same rule as mig_leaves: did the (new) cell row leave the strip? (rare: one atomic each)
        int r, q;
        pf_cell(p, x, y, r, q);
        if (r < p.row0) atomicAdd(&mig_cnt[2 * (i / kOwnChunk)], 1u);
        else if (r >= p.row1) atomicAdd(&mig_cnt[2 * (i / kOwnChunk) + 1], 1u);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 8 && hist[threadIdx.x])
    atomicAdd(&counters[2 + threadIdx.x], (unsigned long long)hist[threadIdx.x]);
  if (threadIdx.x == 8 && hist[8]) atomicAdd(&counters[1], (unsigned long long)hist[8]);
}

// storingPosition + has_moved (reference sup:279-289, 260-273) for externally measured poses
__global__ void __launch_bounds__(256)
k_agents_observe(PFDev p, long long n, float* __restrict__ ax, float* __restrict__ ay,
                 float* __restrict__ ath, uint32_t* __restrict__ ameta,
                 const float* __restrict__ nx, const float* __restrict__ ny,
                 const float* __restrict__ nth) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t mt = ameta[i];
  if ((mt & PF_META_STATE_MASK) == PF_STATE_EMPTY) return;
  float x = nx[i], y = ny[i];
  uint32_t moved = PF_META_MOVED;  // old_pos is None -> True (sup:261-262)
  if (mt & PF_META_HASPOS) {
    float dx = fsub(x, ax[i]), dy = fsub(y, ay[i]);
    float dist = fsqrt(fadd(fmul(dx, dx), fmul(dy, dy)));
    moved = (dist >= p.move_thr) ? PF_META_MOVED : 0u;
  }
  ax[i] = x;
  ay[i] = y;
  if (nth) ath[i] = nth[i];
  ameta[i] = (mt & ~PF_META_MOVED) | moved | PF_META_HASPOS;
}

// ------------------------------------------------------------------------------------------
// migration between row strips
// ------------------------------------------------------------------------------------------
constexpr int kMigWords = 8;  // u32 words per record and per header

// Two-level migration. Leavers and free slots are a few thousand out of millions, so nothing here
// makes a full-array scan: per 1024-slot chunk counts (the move kernel counts leavers with one
// atomic each), one small scan over the chunks, and detail work only in the chunks that need it.
// Order is deterministic: leavers are packed by slot; arrivals take free slots by candidate rank.

// 0 = stays, 1 = leaves upwards, 2 = leaves downwards
__device__ __forceinline__ int mig_leaves(const PFDev& p, float x, float y, uint32_t meta) {
  if ((meta & PF_META_STATE_MASK) == PF_STATE_EMPTY) return 0;
  int r, q;
  pf_cell(p, x, y, r, q);
  return (r < p.row0) ? 1 : ((r >= p.row1) ? 2 : 0);
}

// leaver counts per chunk, for callers that moved the agents themselves (the move kernel counts on its own)
__global__ void __launch_bounds__(256)
k_migrate_count(PFDev p, long long n, const float* __restrict__ x, const float* __restrict__ y,
                const uint32_t* __restrict__ meta, uint32_t* __restrict__ mig_cnt) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int lv = mig_leaves(p, x[i], y[i], meta[i]);
  if (lv) atomicAdd(&mig_cnt[2 * (i / kOwnChunk) + (lv - 1)], 1u);
}

// exclusive scan of `cnt[stride * c + off]` over c < nc into base[]; one CTA of 1024 threads; returns the total
__device__ __forceinline__ uint32_t mig_chunk_scan(const uint32_t* __restrict__ cnt, int stride, int off, int nc,
                                                   uint32_t* __restrict__ base, uint32_t* carry_smem, void* tmp_) {
  typedef cub::BlockScan<uint32_t, 1024> Scan;
  typename Scan::TempStorage& tmp = *reinterpret_cast<typename Scan::TempStorage*>(tmp_);
  if (threadIdx.x == 0) *carry_smem = 0u;
  __syncthreads();
  for (int b0 = 0; b0 < nc; b0 += 1024) {
    const int c = b0 + threadIdx.x;
    const uint32_t v = c < nc ? cnt[stride * c + off] : 0u;
    uint32_t ex, total;
    Scan(tmp).ExclusiveSum(v, ex, total);
    const uint32_t carry = *carry_smem;
    if (c < nc) base[stride * c + off] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) *carry_smem = carry + total;
    __syncthreads();
  }
  return *carry_smem;
}

// headers of the two record buffers + counters; mig_base = first record index of each chunk's leavers
__global__ void __launch_bounds__(1024)
k_migrate_offsets(int nc, long long cap, const uint32_t* __restrict__ mig_cnt, uint32_t* __restrict__ mig_base,
                  uint32_t* __restrict__ up, uint32_t* __restrict__ down,
                  unsigned long long* __restrict__ counters) {
  __shared__ typename cub::BlockScan<uint32_t, 1024>::TempStorage tmp;
  __shared__ uint32_t carry;
  const unsigned long long nu = mig_chunk_scan(mig_cnt, 2, 0, nc, mig_base, &carry, &tmp);
  __syncthreads();
  const unsigned long long nd = mig_chunk_scan(mig_cnt, 2, 1, nc, mig_base, &carry, &tmp);
  if (threadIdx.x == 0) {
    const unsigned long long su = nu < (unsigned long long)cap ? nu : cap;
    const unsigned long long sd = nd < (unsigned long long)cap ? nd : cap;
    up[0] = (uint32_t)su;
    down[0] = (uint32_t)sd;
    atomicAdd(&counters[10], su + sd);
    if (nu + nd > su + sd) atomicAdd(&counters[12], nu + nd - su - sd);
  }
}

// one CTA per chunk; chunks without leavers return at once. Records are written in slot order.
__global__ void __launch_bounds__(256)
k_migrate_scatter(PFDev p, long long n, long long cap, const uint32_t* __restrict__ mig_cnt,
                  const uint32_t* __restrict__ mig_base, float* ax, float* ay, float* ath, float* awl, float* awr,
                  uint32_t* ameta, const uint32_t* __restrict__ agid, uint32_t* __restrict__ up,
                  uint32_t* __restrict__ down) {
  const int c = blockIdx.x;
  if (mig_cnt[2 * c] == 0u && mig_cnt[2 * c + 1] == 0u) return;
  typedef cub::BlockScan<unsigned long long, 256> Scan;
  __shared__ typename Scan::TempStorage tmp;
  const long long i0 = (long long)c * kOwnChunk + 4 * threadIdx.x;
  int lv[4];
  unsigned long long mine = 0ull;  // leavers of this thread: up in the low word, down in the high word
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const long long i = i0 + q;
    lv[q] = i < n ? mig_leaves(p, ax[i], ay[i], ameta[i]) : 0;
    mine += lv[q] == 1 ? 1ull : (lv[q] == 2 ? (1ull << 32) : 0ull);
  }
  unsigned long long ex;
  Scan(tmp).ExclusiveSum(mine, ex);
  unsigned long long ru = mig_base[2 * c] + (ex & 0xffffffffull), rd = mig_base[2 * c + 1] + (ex >> 32);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (!lv[q]) continue;
    const long long i = i0 + q;
    const unsigned long long r = lv[q] == 1 ? ru++ : rd++;
    if (r >= (unsigned long long)cap) continue;  // overflow: stays here, retried next tick
    uint32_t* rec = (lv[q] == 1 ? up : down) + (r + 1) * kMigWords;
    rec[0] = __float_as_uint(ax[i]);
    rec[1] = __float_as_uint(ay[i]);
    rec[2] = __float_as_uint(ath[i]);
    rec[3] = __float_as_uint(awl[i]);
    rec[4] = __float_as_uint(awr[i]);
    rec[5] = ameta[i];
    rec[6] = agid ? agid[i] : 0u;
    rec[7] = 0u;
    ameta[i] = PF_STATE_EMPTY;
  }
}

// Placement order of arriving agents. After pf_agents_sort the arrays are [body (ordered by cell) |
// tail (empty)]. Agents only ever leave from the two ends of the body, so the holes sit there.
// Candidate slots are ranked in the order  [first half of the body | tail | second half of the body]:
// arrivals from the strip above (small keys) take ranks ascending — holes at the start of the body,
// then the tail; arrivals from the strip below (large keys) take ranks descending — holes at the end
// of the body, then the tail from its far end. An arrival therefore never lands among agents whose
// keys are far from its own, which would widen the chunk bounds of the owner-computes deposit.
__device__ __forceinline__ long long mig_slot_of_rank_index(long long ip, long long n, long long T, long long M) {
  if (ip < M) return ip;
  if (ip < M + (n - T)) return T + (ip - M);
  return M + (ip - M - (n - T));
}
__device__ __forceinline__ void mig_body_split(const uint32_t* body_chunks, long long n, long long& T, long long& M) {
  T = body_chunks ? (long long)body_chunks[0] * kOwnChunk : 0;
  if (T > n) T = n;
  M = (T / 2) / kOwnChunk * kOwnChunk;
}

// free slots per chunk of 1024 candidate rank indices
__global__ void __launch_bounds__(256)
k_empty_count(long long n, const uint32_t* __restrict__ meta, const uint32_t* __restrict__ body_chunks,
              uint32_t* __restrict__ emp_cnt) {
  long long T, M;
  mig_body_split(body_chunks, n, T, M);
  const long long ip0 = (long long)blockIdx.x * kOwnChunk + 4 * threadIdx.x;
  unsigned cnt = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const long long ip = ip0 + q;
    if (ip < n && (meta[mig_slot_of_rank_index(ip, n, T, M)] & PF_META_STATE_MASK) == PF_STATE_EMPTY) ++cnt;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  __shared__ unsigned part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned s = 0;
    for (int w = 0; w < 8; ++w) s += part[w];
    emp_cnt[blockIdx.x] = s;
  }
}

// emp_base[c] = free slots before chunk c, emp_base[nc] = all of them; arrival counters
__global__ void __launch_bounds__(1024)
k_empty_offsets(int nc, long long cap, const uint32_t* __restrict__ emp_cnt, uint32_t* __restrict__ emp_base,
                const uint32_t* __restrict__ buf, const uint32_t* __restrict__ buf2,
                unsigned long long* __restrict__ counters) {
  __shared__ typename cub::BlockScan<uint32_t, 1024>::TempStorage tmp;
  __shared__ uint32_t carry;
  const unsigned long long empties = mig_chunk_scan(emp_cnt, 1, 0, nc, emp_base, &carry, &tmp);
  if (threadIdx.x == 0) {
    emp_base[nc] = (uint32_t)empties;
    unsigned long long m1 = buf ? buf[0] : 0ull, m2 = buf2 ? buf2[0] : 0ull;
    if (m1 > (unsigned long long)cap) m1 = cap;
    if (m2 > (unsigned long long)cap) m2 = cap;
    const unsigned long long m = m1 + m2, placed = m < empties ? m : empties;
    atomicAdd(&counters[11], placed);
    if (m > placed) atomicAdd(&counters[12], m - placed);
  }
}

// One warp per arriving record: records of `buf` (from the strip above) take the free slots in
// candidate rank order ascending, records of `buf2` (from the strip below) descending. The warp finds
// its rank's chunk by bisection of emp_base and the slot inside the chunk with 32 ballots at most.
// Slots are only looked up here (slot_of[record] = slot or -1) and filled by k_migrate_fill, so that
// every warp sees the same free slots the counts were taken from.
__global__ void __launch_bounds__(256)
k_migrate_find(long long n, int nc, const uint32_t* __restrict__ emp_base, const uint32_t* __restrict__ buf,
               const uint32_t* __restrict__ buf2, long long cap, const uint32_t* __restrict__ ameta,
               const uint32_t* __restrict__ body_chunks, long long* __restrict__ slot_of) {
  const int lane = threadIdx.x & 31;
  unsigned long long m1 = buf ? buf[0] : 0ull, m2 = buf2 ? buf2[0] : 0ull;
  if (m1 > (unsigned long long)cap) m1 = cap;
  if (m2 > (unsigned long long)cap) m2 = cap;
  const unsigned long long empties = emp_base[nc];
  long long T, M;
  mig_body_split(body_chunks, n, T, M);
  // a small grid strides over the records that actually arrived (a few thousand out of 2 * cap)
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; (unsigned long long)w < m1 + m2; w += nwarps) {
    if (lane == 0) slot_of[w] = -1;
    unsigned long long r;  // candidate rank among the free slots
    if ((unsigned long long)w < m1) {
      r = (unsigned long long)w;
      if (r >= empties) continue;  // no room: dropped (counted in k_empty_offsets)
    } else {
      const unsigned long long back = (unsigned long long)w - m1;  // 0 takes the highest free slot
      if (back >= empties || empties - 1 - back < m1) continue;   // the ascending records keep theirs
      r = empties - 1 - back;
    }
    int lo = 0, hi = nc;  // last chunk with emp_base[c] <= r
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((unsigned long long)emp_base[mid] <= r) lo = mid; else hi = mid;
    }
    unsigned need = (unsigned)(r - emp_base[lo]);  // free slots of this chunk to skip
    long long slot = -1;
    for (int k = 0; k < kOwnChunk / 32; ++k) {
      const long long ip = (long long)lo * kOwnChunk + 32 * k + lane;
      const long long i = ip < n ? mig_slot_of_rank_index(ip, n, T, M) : -1;
      const bool empty = i >= 0 && (ameta[i] & PF_META_STATE_MASK) == PF_STATE_EMPTY;
      const unsigned m = __ballot_sync(0xffffffffu, empty);
      const unsigned c = (unsigned)__popc(m);
      if (need < c) {  // the need-th set bit of m
        unsigned mm = m;
        for (unsigned z = 0; z < need; ++z) mm &= mm - 1u;
        const int src = __ffs(mm) - 1;
        slot = __shfl_sync(0xffffffffu, i, src);
        break;
      }
      need -= c;
    }
    if (lane == 0) slot_of[w] = slot;
  }
}

// one thread per arriving record
__global__ void __launch_bounds__(256)
k_migrate_fill(const uint32_t* __restrict__ buf, const uint32_t* __restrict__ buf2, long long cap,
               const long long* __restrict__ slot_of, float* ax, float* ay, float* ath, float* awl, float* awr,
               uint32_t* ameta, uint32_t* agid) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long m1 = buf ? buf[0] : 0ull, m2 = buf2 ? buf2[0] : 0ull;
  if (m1 > (unsigned long long)cap) m1 = cap;
  if (m2 > (unsigned long long)cap) m2 = cap;
  if ((unsigned long long)w >= m1 + m2) return;
  const long long slot = slot_of[w];
  if (slot < 0) return;
  const uint32_t* rec = (unsigned long long)w < m1 ? buf + (w + 1) * kMigWords : buf2 + (w - m1 + 1) * kMigWords;
  ax[slot] = __uint_as_float(rec[0]);
  ay[slot] = __uint_as_float(rec[1]);
  ath[slot] = __uint_as_float(rec[2]);
  awl[slot] = __uint_as_float(rec[3]);
  awr[slot] = __uint_as_float(rec[4]);
  ameta[slot] = rec[5];
  if (agid) agid[slot] = rec[6];
}

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
// explicit halo push: copy this strip's first / last interior row of `f` into the neighbours' ghost rows
__global__ void k_peer_push_edges(PFDev p, const float* __restrict__ f, PeerRows peer) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  int ch = blockIdx.y;
  if (col >= p.W) return;
  const float* plane = f + (long long)ch * p.chan_stride;
  if (peer.up) peer.up[ch * peer.up_cs + col] = plane[p.pitch + col];
  if (peer.dn) peer.dn[ch * peer.dn_cs + col] = plane[(long long)p.rows * p.pitch + col];
}

// Epoch flags in the RECEIVER's memory: a strip stores `epoch` into its neighbours' flag words after
// its deposit (halo) / after packing its leavers (migration); each strip polls its own words before it
// reads what the neighbours stored. The poll is bounded: a neighbour that has not signalled after
// `limit_ns` (dead rank, strips out of lockstep) is counted in counters[13] (pf_counters.peer_timeouts,
// pf_peer_status) and the kernel goes on instead of wedging the stream.
__device__ __forceinline__ unsigned long long pf_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool peer_spin(volatile const unsigned int* w, unsigned int epoch, unsigned long long limit_ns) {
  if ((int)(*w - epoch) >= 0) return true;
  const unsigned long long t0 = pf_globaltimer();
  unsigned ns = 32;
  while ((int)(*w - epoch) < 0) {
    __nanosleep(ns);
    if (ns < 2048) ns *= 2;
    if (pf_globaltimer() - t0 > limit_ns) return false;
  }
  return true;
}
__global__ void k_peer_signal(volatile unsigned int* up_flag, volatile unsigned int* dn_flag, unsigned int epoch) {
  __threadfence_system();
  if (up_flag) *up_flag = epoch;
  if (dn_flag) *dn_flag = epoch;
  __threadfence_system();
}
__global__ void k_peer_wait(volatile const unsigned int* from_up, volatile const unsigned int* from_dn,
                            unsigned int epoch, unsigned long long limit_ns, unsigned long long* counters) {
  bool ok = true;
  if (from_up) ok = peer_spin(from_up, epoch, limit_ns) && ok;
  if (from_dn) ok = peer_spin(from_dn, epoch, limit_ns) && ok;
  if (!ok) atomicAdd(&counters[13], 1ull);
  __threadfence_system();
}
// halo wait + the rows the neighbours stored into this strip's mailbox become its ghost rows
__global__ void __launch_bounds__(256)
k_peer_wait_ghosts(PFDev p, float* f, const float* mailbox, volatile const unsigned int* from_up,
                   volatile const unsigned int* from_dn, unsigned int epoch, unsigned long long limit_ns,
                   unsigned long long* counters) {
  if (threadIdx.x == 0) {
    bool ok = true;
    if (from_up) ok = peer_spin(from_up, epoch, limit_ns) && ok;
    if (from_dn) ok = peer_spin(from_dn, epoch, limit_ns) && ok;
    if (!ok && blockIdx.x == 0 && blockIdx.y == 0) atomicAdd(&counters[13], 1ull);
    __threadfence_system();
  }
  __syncthreads();
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  int ch = blockIdx.y;
  if (col >= p.W) return;
  float* plane = f + (long long)ch * p.chan_stride;
  if (from_up) plane[col] = __ldcg(mailbox + (long long)ch * p.pitch + col);
  if (from_dn) plane[(long long)(p.rows + 1) * p.pitch + col] = __ldcg(mailbox + (long long)(p.C + ch) * p.pitch + col);
}

__global__ void k_fill_ghosts(PFDev p, float* f, int top, int bottom) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  int ch = blockIdx.y;
  if (col >= p.W) return;
  float* plane = f + (long long)ch * p.chan_stride;
  if (top) plane[col] = plane[p.pitch + col];
  if (bottom) plane[(long long)(p.rows + 1) * p.pitch + col] = plane[(long long)p.rows * p.pitch + col];
}

__global__ void k_fill(float* f, long long n, float v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) f[i] = v;
}

// I * exp(-rate * elapsed), fp64 (reference sup:519-525)
__global__ void k_evaporate_trail(long long n, const double* __restrict__ elapsed,
                                  const double* __restrict__ intensity, double rate,
                                  double* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = intensity[i] * exp(-rate * elapsed[i]);
}

__global__ void k_tick_counter(unsigned long long* counters) { counters[0] += 1ull; }

// ------------------------------------------------------------------------------------------
// spatial re-ordering of the agent arrays (locality for the sense gathers and the deposit scatter)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_sort_keys(PFDev p, long long n, const float* __restrict__ x, const float* __restrict__ y,
            const uint32_t* __restrict__ meta, uint32_t* __restrict__ keys,
            uint32_t* __restrict__ idx, uint32_t sentinel) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t key = sentinel;  // empty slots go to the end
  if ((meta[i] & PF_META_STATE_MASK) != PF_STATE_EMPTY) {
    int r, q;
    pf_cell(p, x[i], y[i], r, q);
    key = (uint32_t)((long long)r * p.W + q);
  }
  keys[i] = key;
  idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
k_permute_u32(long long n, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ src,
              uint32_t* __restrict__ dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[i] = src[perm[i]];
}
