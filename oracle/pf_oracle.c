/*
 * pf_oracle.c — CPU restatement (plain C, scalar, single thread) of the dense pheromone tick.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product path (swarm_robotics_pheromones_b200/, the
 * C-ABI library) may link or call this file. It is used by tests/, by __graft_entry__.smoke()
 * and by bench.py's cpu_baseline / --impl reference legs as the checker and the CPU baseline.
 *
 * What it follows (reference file:line, relative to the reference checkout;
 * sup = Phase-3/pheromone_algo_supervisor.py, algo = Phase-3/pheromone_algo.py):
 *   - deposit  : algo:351 (cell = int(x), int(y); +=), sup:291-358 (moved gate, 1/20/5 amounts,
 *                every-3rd rule on len(global grid)), sup:260-273 (has_moved)
 *   - evaporate: algo:354-356 (grid *= (1 - rate), AFTER the deposit of the same tick)
 *   - sense    : sup:379-417 (five directions in dict order current, front, back, left, right)
 *   - decide   : sup:440-450 (ordered arg-max, strict >, first wins), sup:185-226 (PID),
 *                sup:455-470 (SEARCHING speeds), sup:475-514 (HOMING speeds), sup:665-670
 *   - tick order: sup:575-607 + 753-754: deposit -> sense/decide -> evaporate -> move
 *
 * PARITY PINNING. The pieces above are pinned against the real reference through the golden
 * traces in tests/golden/ (made by tests/golden/make_golden.py, which imports the unmodified
 * reference headless) and the closed form of algo:351-356. The DIFFUSION stencil, the two-channel
 * routing and the diff-drive motion model have NO reference arithmetic:
 * for those this file is the definition and parity is "unpinned"; they are cross-checked against
 * scipy.ndimage.correlate(mode='reflect') and mass conservation in tests/test_oracle_dense.py.
 *
 * Arithmetic contract shared with the CUDA kernels: IEEE fp32, round-to-nearest-even, every
 * operation rounded separately (no FMA contraction; build with -ffp-contract=off), no libm
 * transcendentals on the agent path (po_sincosf / po_atan2f below are fixed polynomial
 * sequences), sqrtf and division correctly rounded.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pherofield.h"

#if defined(__FMA__)
#warning "compile the oracle without FMA contraction (-ffp-contract=off)"
#endif

#define PO_PI 3.14159274101257324f
#define PO_TWO_PI 6.28318548202514648f
#define PO_HALF_PI 1.57079637050628662f
#define PO_QUARTER_PI 0.785398185253143311f

/* ---------------------------------------------------------------- deterministic math */

/* sin and cos of t (|t| a few pi at most): Cody-Waite reduction by pi/2 in three parts, then the
 * Cephes single-precision minimax polynomials on [-pi/4, pi/4]. */
void po_sincosf(float t, float* s_out, float* c_out) {
  float k = rintf(t * 0.636619747f);
  float r = t - k * 1.5703125f;
  r = r - k * 4.83751296997070312e-4f;
  r = r - k * 7.54978995489188e-8f;
  float z = r * r;
  float sp = -1.9515295891e-4f * z;
  sp = sp + 8.3321608736e-3f;
  sp = sp * z;
  sp = sp - 1.6666654611e-1f;
  sp = sp * z;
  sp = sp * r;
  float s = sp + r;
  float cp = 2.443315711809948e-5f * z;
  cp = cp - 1.388731625493765e-3f;
  cp = cp * z;
  cp = cp + 4.166664568298827e-2f;
  cp = cp * z;
  cp = cp * z;
  float hz = 0.5f * z;
  cp = cp - hz;
  float c = cp + 1.0f;
  int q = ((int)k) & 3;
  switch (q) {
    case 0: *s_out = s; *c_out = c; break;
    case 1: *s_out = c; *c_out = -s; break;
    case 2: *s_out = -s; *c_out = -c; break;
    default: *s_out = -c; *c_out = s; break;
  }
}

/* atan2(y, x) in (-pi, pi]: octant reduction to t in [0,1], Cephes atanf polynomial. */
float po_atan2f(float y, float x) {
  float ax = fabsf(x), ay = fabsf(y);
  if (ax == 0.0f && ay == 0.0f) return 0.0f;
  int swap = ay > ax;
  float t = swap ? ax / ay : ay / ax;
  float a0 = 0.0f;
  if (t > 0.414213568f) {
    float num = t - 1.0f;
    float den = t + 1.0f;
    t = num / den;
    a0 = PO_QUARTER_PI;
  }
  float z = t * t;
  float p = 8.05374449538e-2f * z;
  p = p - 1.38776856032e-1f;
  p = p * z;
  p = p + 1.99777106478e-1f;
  p = p * z;
  p = p - 3.33329491539e-1f;
  p = p * z;
  p = p * t;
  p = p + t;
  float a = a0 + p;
  if (swap) a = PO_HALF_PI - a;
  if (x < 0.0f) a = PO_PI - a;
  if (y < 0.0f) a = -a;
  return a;
}

static float po_wrap(float a) {
  if (a > PO_PI) a = a - PO_TWO_PI;
  if (a < -PO_PI) a = a + PO_TWO_PI;
  return a;
}

/* PIDController.update, sup:199-226, with old_e = 0 and the ki term unused:
 * e = wrap(atan2(dy,dx) - pi/2); out = wrap(kp*e + kd*e). */
float po_pid_turn(const pf_config* c, float cx, float cy, float gx, float gy) {
  float dx = gx - cx;
  float dy = gy - cy;
  float a = po_atan2f(dy, dx);
  float e = po_wrap(a - PO_HALF_PI);
  float pe = c->pid_kp * e;
  float de = c->pid_kd * e;
  return po_wrap(pe + de);
}

/* ---------------------------------------------------------------- geometry */

typedef struct po_geom {
  float inv_cs, xmin, xmax, ymin, ymax;
  int rows;     /* local rows */
  int64_t pitch;    /* floats per row (== W in the oracle) */
  int64_t chan_stride; /* floats per channel = (rows+2)*pitch */
} po_geom;

static void po_geometry(const pf_config* c, po_geom* g) {
  g->inv_cs = 1.0f / c->cell_size;
  g->xmin = c->origin_x;
  g->ymin = c->origin_y;
  g->xmax = c->origin_x + (float)c->H * c->cell_size;
  g->ymax = c->origin_y + (float)c->W * c->cell_size;
  g->rows = c->row1 - c->row0;
  g->pitch = c->W;
  g->chan_stride = (int64_t)(g->rows + 2) * g->pitch;
}

/* number of floats of one buffer: C channels x (rows + 2 ghost rows) x W */
int64_t po_field_elems(const pf_config* c) {
  po_geom g;
  po_geometry(c, &g);
  return (int64_t)c->C * g.chan_stride;
}

/* cell of a world position: truncation toward zero (Python int(), algo:351), then clamped into
 * the arena. Returns 0 when the raw index is outside the arena (before clamping); a point exactly
 * on the upper wall (index == H) belongs to the last cell. */
static int po_cell(const pf_config* c, const po_geom* g, float x, float y, int* row, int* col) {
  float fr = (x - c->origin_x) * g->inv_cs;
  float fc = (y - c->origin_y) * g->inv_cs;
  int inside = (fr >= 0.0f) && (fc >= 0.0f) && (fr <= (float)c->H) && (fc <= (float)c->W);
  int r = (int)fr, q = (int)fc; /* C conversion truncates toward zero */
  if (!(fr >= 0.0f)) r = 0;
  if (!(fc >= 0.0f)) q = 0;
  if (fr >= (float)c->H) r = c->H - 1;
  if (fc >= (float)c->W) q = c->W - 1;
  *row = r;
  *col = q;
  return inside;
}

static inline float* po_at(const pf_config* c, const po_geom* g, float* f, int ch, int grow,
                           int col) {
  /* grow is a GLOBAL row; local row -1 and `rows` are the ghosts */
  return f + (int64_t)ch * g->chan_stride + (int64_t)(grow - c->row0 + 1) * g->pitch + col;
}

/* reflect (zero-flux) ghost rows at the arena's outer edges: ghost = duplicate of the edge row
 * (scipy 'reflect' d c b a | a b c d for radius 1; graph.py:24 default mode). */
void po_fill_ghosts(const pf_config* c, float* f) {
  po_geom g;
  po_geometry(c, &g);
  for (int ch = 0; ch < c->C; ++ch) {
    if (c->row0 == 0)
      memcpy(po_at(c, &g, f, ch, -1, 0), po_at(c, &g, f, ch, 0, 0), sizeof(float) * c->W);
    if (c->row1 == c->H)
      memcpy(po_at(c, &g, f, ch, c->H, 0), po_at(c, &g, f, ch, c->H - 1, 0),
             sizeof(float) * c->W);
  }
}

/* ---------------------------------------------------------------- evaporate + diffuse */

/* One step cur -> nxt over the owned rows; ghost rows of cur must be valid.
 *   5-point: lap = ((n + s) + (w + e)) - 4*c
 *   9-point: lap = (4*((n + s) + (w + e)) + ((nw + ne) + (sw + se))) - 20*c, D folded with 1/6
 *   g = c + Dk*lap ; f' = keep*g ; f' <= delete_threshold -> 0 (sup:603) when enabled.
 * Column neighbours clamp at the arena edge (reflect). stencil == 0: f' = keep*c (algo:356). */
void po_step_field(const pf_config* c, const float* cur, float* nxt) {
  po_geom g;
  po_geometry(c, &g);
  const int W = c->W;
  for (int ch = 0; ch < c->C; ++ch) {
    const float keep = c->evap_keep[ch];
    const float D = c->diffusion[ch];
    const float Dk = (c->stencil == 9) ? D / 6.0f : D;
    for (int lr = 0; lr < g.rows; ++lr) {
      const float* rc = cur + (int64_t)ch * g.chan_stride + (int64_t)(lr + 1) * g.pitch;
      const float* rn = rc - g.pitch; /* north = row - 1 */
      const float* rs = rc + g.pitch;
      float* out = nxt + (int64_t)ch * g.chan_stride + (int64_t)(lr + 1) * g.pitch;
      for (int j = 0; j < W; ++j) {
        int jw = j > 0 ? j - 1 : 0;
        int je = j < W - 1 ? j + 1 : W - 1;
        float cc = rc[j];
        float v;
        if (c->stencil == 0) {
          v = keep * cc;
        } else {
          float ns = rn[j] + rs[j];
          float we = rc[jw] + rc[je];
          float cross = ns + we;
          float lap;
          if (c->stencil == 9) {
            float d1 = rn[jw] + rn[je];
            float d2 = rs[jw] + rs[je];
            float diag = d1 + d2;
            float c4 = 4.0f * cross;
            float sum = c4 + diag;
            float c20 = 20.0f * cc;
            lap = sum - c20;
          } else {
            float c4 = 4.0f * cc;
            lap = cross - c4;
          }
          float dl = Dk * lap;
          float gsum = cc + dl;
          v = keep * gsum;
        }
        if (c->delete_threshold > 0.0f && v <= c->delete_threshold) v = 0.0f;
        out[j] = v;
      }
    }
  }
}

/* ---------------------------------------------------------------- deposit */

/* Generic deposit of n amounts. Per touched cell the amounts are summed in input order starting
 * from 0, then the total is added to the cell once: f = f + ((a1 + a2) + a3 ...). Positions
 * outside the arena or outside rows [row0,row1) are skipped. channel may be NULL.
 * Two equivalent implementations (tests/test_oracle_dense.py compares them): a dense accumulator
 * over all cells, and — for the full-size configs, where a per-call accumulator of 2^31 cells is
 * not affordable — a stable sort of the deposits by cell followed by the same in-order sums. */
void po_deposit_dense(const pf_config* c, float* f, int64_t n, const float* x, const float* y,
                      const float* amount, const uint8_t* channel) {
  po_geom g;
  po_geometry(c, &g);
  int64_t cells = (int64_t)c->C * g.rows * c->W;
  float* acc = (float*)calloc((size_t)cells, sizeof(float));
  uint8_t* touched = (uint8_t*)calloc((size_t)cells, 1);
  for (int64_t i = 0; i < n; ++i) {
    int r, q;
    if (!po_cell(c, &g, x[i], y[i], &r, &q)) continue;
    if (r < c->row0 || r >= c->row1) continue;
    int ch = channel ? channel[i] : 0;
    if (ch < 0 || ch >= c->C) continue;
    int64_t k = ((int64_t)ch * g.rows + (r - c->row0)) * c->W + q;
    acc[k] = acc[k] + amount[i];
    touched[k] = 1;
  }
  for (int ch = 0; ch < c->C; ++ch)
    for (int lr = 0; lr < g.rows; ++lr)
      for (int q = 0; q < c->W; ++q) {
        int64_t k = ((int64_t)ch * g.rows + lr) * c->W + q;
        if (touched[k]) {
          float* p = po_at(c, &g, f, ch, lr + c->row0, q);
          *p = *p + acc[k];
        }
      }
  free(acc);
  free(touched);
}

void po_deposit_sorted(const pf_config* c, float* f, int64_t n, const float* x, const float* y,
                       const float* amount, const uint8_t* channel) {
  po_geom g;
  po_geometry(c, &g);
  size_t cap = (size_t)(n > 0 ? n : 1);
  uint64_t* key = (uint64_t*)malloc(sizeof(uint64_t) * cap);
  uint64_t* key2 = (uint64_t*)malloc(sizeof(uint64_t) * cap);
  uint32_t* idx = (uint32_t*)malloc(sizeof(uint32_t) * cap);
  uint32_t* idx2 = (uint32_t*)malloc(sizeof(uint32_t) * cap);
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    int r, q;
    if (!po_cell(c, &g, x[i], y[i], &r, &q)) continue;
    if (r < c->row0 || r >= c->row1) continue;
    int ch = channel ? channel[i] : 0;
    if (ch < 0 || ch >= c->C) continue;
    key[m] = (uint64_t)(((int64_t)ch * g.rows + (r - c->row0)) * c->W + q);
    idx[m] = (uint32_t)i;
    ++m;
  }
  /* stable LSD radix sort by cell, 16 bits a pass: equal cells keep input order */
  uint64_t maxk = (uint64_t)c->C * (uint64_t)g.rows * (uint64_t)c->W;
  size_t* cnt = (size_t*)malloc(sizeof(size_t) * 65537);
  for (int shift = 0; shift < 64 && (maxk >> shift) != 0; shift += 16) {
    memset(cnt, 0, sizeof(size_t) * 65537);
    for (int64_t i = 0; i < m; ++i) cnt[((key[i] >> shift) & 0xffffu) + 1]++;
    for (int b = 0; b < 65536; ++b) cnt[b + 1] += cnt[b];
    for (int64_t i = 0; i < m; ++i) {
      size_t d = cnt[(key[i] >> shift) & 0xffffu]++;
      key2[d] = key[i];
      idx2[d] = idx[i];
    }
    uint64_t* tk = key; key = key2; key2 = tk;
    uint32_t* ti = idx; idx = idx2; idx2 = ti;
  }
  for (int64_t i = 0; i < m;) {
    float acc = 0.0f;
    int64_t j = i;
    while (j < m && key[j] == key[i]) {
      acc = acc + amount[idx[j]];
      ++j;
    }
    int64_t k = (int64_t)key[i];
    int64_t plane = (int64_t)g.rows * c->W;
    int ch = (int)(k / plane);
    int64_t rem = k - (int64_t)ch * plane;
    int lr = (int)(rem / c->W);
    int q = (int)(rem - (int64_t)lr * c->W);
    float* p = po_at(c, &g, f, ch, lr + c->row0, q);
    *p = *p + acc;
    i = j;
  }
  free(cnt);
  free(key); free(key2); free(idx); free(idx2);
}

void po_deposit(const pf_config* c, float* f, int64_t n, const float* x, const float* y,
                const float* amount, const uint8_t* channel) {
  int64_t cells = (int64_t)c->C * (c->row1 - c->row0) * c->W;
  if (cells > ((int64_t)1 << 26)) po_deposit_sorted(c, f, n, x, y, amount, channel);
  else po_deposit_dense(c, f, n, x, y, amount, channel);
}

/* deposit_pheromone over all agents, sup:291-358: gate on the moved flag (sup:308-310); SEARCHING:
 * counter += 1, amount = 20 when counter % high_every == 0 else 1 (sup:312-318); HOMING: counter
 * += 1, amount 5 (sup:346-348); IDLE/empty: nothing. Returns number of depositing agents.
 * x/y -> cell, channel = deposit_channel[state]. */
int64_t po_agents_deposit(const pf_config* c, float* f, int64_t n, const float* x, const float* y,
                          uint32_t* meta) {
  float* amt = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  uint8_t* chn = (uint8_t*)malloc((size_t)(n > 0 ? n : 1));
  float* xs = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  float* ys = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    uint32_t mt = meta[i];
    uint32_t st = mt & PF_META_STATE_MASK;
    if (st != PF_STATE_SEARCHING && st != PF_STATE_HOMING) continue;
    if (!(mt & PF_META_MOVED)) continue;
    uint32_t cnt = (mt >> PF_META_COUNT_SHIFT) + 1u;
    float a;
    if (st == PF_STATE_SEARCHING)
      a = (cnt % (uint32_t)c->high_every == 0u) ? c->intensity_high : c->intensity_explore;
    else
      a = c->intensity_homing;
    meta[i] = (mt & 0xFFu) | (cnt << PF_META_COUNT_SHIFT);
    xs[m] = x[i];
    ys[m] = y[i];
    amt[m] = a;
    chn[m] = (uint8_t)c->deposit_channel[st];
    ++m;
  }
  po_deposit(c, f, m, xs, ys, amt, chn);
  free(amt);
  free(chn);
  free(xs);
  free(ys);
  return m;
}

/* ---------------------------------------------------------------- sense */

/* Five values in dict order (sup:390). A direction whose cell is outside the arena yields NaN
 * (the reference's empty list). Rows outside [row0-1, row1] cannot be read by this strip: the
 * caller guarantees ghost depth >= sensing reach (world offsets of +-1 row, sense_dist <= cell). */
void po_sense_one(const pf_config* c, const float* f, float x, float y, float theta, uint32_t st,
                  float out5[5]) {
  po_geom g;
  po_geometry(c, &g);
  int ch = c->sense_channel[st & 3u];
  float s = 0.0f, co = 1.0f;
  if (c->sense_mode == PF_SENSE_HEADING) po_sincosf(theta, &s, &co);
  int r0, q0;
  po_cell(c, &g, x, y, &r0, &q0);
  for (int d = 0; d < 5; ++d) {
    int r, q, inside;
    if (c->sense_mode == PF_SENSE_WORLD) {
      r = r0 + c->sense_drow[d];
      q = q0 + c->sense_dcol[d];
      inside = (r >= 0 && r < c->H && q >= 0 && q < c->W);
    } else {
      float ux, uy; /* unit vector of the sensor direction */
      switch (d) {
        case PF_DIR_FRONT: ux = co; uy = s; break;
        case PF_DIR_BACK: ux = -co; uy = -s; break;
        case PF_DIR_LEFT: ux = -s; uy = co; break;
        case PF_DIR_RIGHT: ux = s; uy = -co; break;
        default: ux = 0.0f; uy = 0.0f; break;
      }
      float ox = c->sense_dist * ux;
      float oy = c->sense_dist * uy;
      float px = x + ox;
      float py = y + oy;
      if (d == PF_DIR_CURRENT) { px = x; py = y; }
      inside = po_cell(c, &g, px, py, &r, &q);
      if (d == PF_DIR_CURRENT) inside = 1; /* the agent's own (clamped) cell always exists */
    }
    if (d == PF_DIR_CURRENT && c->sense_mode == PF_SENSE_WORLD) inside = 1;
    if (inside && r >= c->row0 - 1 && r <= c->row1)
      out5[d] = *po_at(c, &g, (float*)f, ch, r, q);
    else
      out5[d] = NAN;
  }
}

void po_sense(const pf_config* c, const float* f, int64_t n, const float* x, const float* y,
              const float* theta, const uint32_t* meta, float* out5) {
  for (int64_t i = 0; i < n; ++i)
    po_sense_one(c, f, x[i], y[i], theta ? theta[i] : 0.0f,
                 meta ? (meta[i] & PF_META_STATE_MASK) : PF_STATE_SEARCHING, out5 + 5 * i);
}

/* ordered arg-max with strict >, first maximum wins, absent (NaN) directions skipped: sup:440-450 */
int po_argmax5(const float v[5]) {
  int dir = -1;
  float best = 0.0f;
  for (int d = 0; d < 5; ++d) {
    if (v[d] != v[d]) continue;
    if (dir < 0 || v[d] > best) {
      best = v[d];
      dir = d;
    }
  }
  return dir;
}

/* ---------------------------------------------------------------- decide + move */

static int po_nearest_food(const pf_config* c, float x, float y, float* d2_out) {
  int best = 0;
  float bd = 0.0f;
  for (int k = 0; k < c->n_food; ++k) {
    float dx = c->food_x[k] - x;
    float dy = c->food_y[k] - y;
    float dx2 = dx * dx;
    float dy2 = dy * dy;
    float d2 = dx2 + dy2;
    if (k == 0 || d2 < bd) {
      bd = d2;
      best = k;
    }
  }
  *d2_out = bd;
  return best;
}

/* One agent, one tick of sense -> decide -> move. Returns the PF_DEC_* decision.
 * counters: [0..5] decisions, [6] food_grabbed, [7] food_delivered (may be NULL). */
int po_agent_step_one(const pf_config* c, const float* f, uint32_t flags, float* px, float* py,
                      float* pth, float* pwl, float* pwr, uint32_t* pmeta, uint64_t* counters) {
  po_geom g;
  po_geometry(c, &g);
  uint32_t mt = *pmeta;
  uint32_t st = mt & PF_META_STATE_MASK;
  if (st == PF_STATE_EMPTY) return PF_DEC_NONE;
  float x = *px, y = *py, th = *pth, wl = *pwl, wr = *pwr;
  int dec = PF_DEC_NONE;
  int delivered = 0;

  if (st == PF_STATE_IDLE) {
    wl = 0.0f;
    wr = 0.0f;
  } else if (!(flags & PF_STEP_SENSE)) {
    /* random walk before t = 5 s: Braitenberg base speeds only, sup:721-725, 736-744 */
    wl = c->braitenberg_l;
    wr = c->braitenberg_r;
  } else if (st == PF_STATE_SEARCHING) {
    float v5[5];
    po_sense_one(c, f, x, y, th, st, v5);
    int dir = po_argmax5(v5);
    float d2;
    int k = po_nearest_food(c, x, y, &d2);
    float turn = po_pid_turn(c, x, y, c->food_x[k], c->food_y[k]);
    if (dir == PF_DIR_LEFT) {
      float m = c->max_speed - fabsf(turn);
      wl = m > 0.0f ? m : 0.0f;
      wr = c->max_speed;
      dec = PF_DEC_LEFT;
    } else if (dir == PF_DIR_RIGHT) {
      float m = c->max_speed - fabsf(turn);
      wr = m > 0.0f ? m : 0.0f;
      wl = c->max_speed;
      dec = PF_DEC_RIGHT;
    } else {
      wl = c->cruise_speed;
      wr = c->cruise_speed;
      dec = PF_DEC_STRAIGHT;
    }
    wl = wl + c->braitenberg_l;
    wr = wr + c->braitenberg_r;
  } else { /* HOMING, sup:475-514 */
    float dx = c->nest_x - x;
    float dy = c->nest_y - y;
    float dx2 = dx * dx;
    float dy2 = dy * dy;
    float d = sqrtf(dx2 + dy2);
    float dcm = rintf(d * 100.0f);
    int reached = (x == c->nest_x && y == c->nest_y) ||
                  (dcm >= c->home_lo_cm && dcm <= c->home_hi_cm);
    if (!reached) {
      float turn = po_pid_turn(c, x, y, c->nest_x, c->nest_y);
      if (d <= c->home_radius) {
        float rt = rintf(turn * 100.0f);
        if (rt > -200.0f && rt < 0.0f) {
          wl = 0.0f;
          wr = c->max_speed;
          dec = PF_DEC_LEFT;
        } else if (rt < -200.0f) {
          wr = 0.0f;
          wl = c->max_speed;
          dec = PF_DEC_RIGHT;
        } else {
          dec = PF_DEC_KEEP;
        }
      } else {
        wl = c->cruise_speed;
        wr = c->cruise_speed;
        dec = PF_DEC_STRAIGHT;
      }
      wl = wl + c->braitenberg_l;
      wr = wr + c->braitenberg_r;
    } else {
      wl = 0.0f;
      wr = 0.0f;
      dec = PF_DEC_STOP;
      delivered = 1;
    }
  }

  uint32_t moved = mt & PF_META_MOVED;
  if (flags & PF_STEP_MOVE) {
    float sum = wl + wr;
    float rs = c->wheel_radius * sum;
    float v = rs * 0.5f;
    float dif = wr - wl;
    float rd = c->wheel_radius * dif;
    float om = rd / c->axle_length;
    float dth = om * c->dt;
    float nth = po_wrap(th + dth);
    float sn, cs;
    po_sincosf(nth, &sn, &cs);
    float step = v * c->dt;
    float mx = step * cs;
    float my = step * sn;
    float nx = x + mx;
    float ny = y + my;
    if (nx < g.xmin) {
      float over = g.xmin - nx;
      nx = g.xmin + over;
      nth = po_wrap(PO_PI - nth);
    } else if (nx > g.xmax) {
      float over = nx - g.xmax;
      nx = g.xmax - over;
      nth = po_wrap(PO_PI - nth);
    }
    if (ny < g.ymin) {
      float over = g.ymin - ny;
      ny = g.ymin + over;
      nth = -nth;
    } else if (ny > g.ymax) {
      float over = ny - g.ymax;
      ny = g.ymax - over;
      nth = -nth;
    }
    if (nx < g.xmin) nx = g.xmin;
    if (nx > g.xmax) nx = g.xmax;
    if (ny < g.ymin) ny = g.ymin;
    if (ny > g.ymax) ny = g.ymax;
    float ddx = nx - x;
    float ddy = ny - y;
    float sx = ddx * ddx;
    float sy = ddy * ddy;
    float dist = sqrtf(sx + sy);
    moved = (dist >= c->move_threshold) ? PF_META_MOVED : 0u; /* sup:263-273 */
    x = nx;
    y = ny;
    th = nth;
  }

  uint32_t carry = mt & PF_META_CARRY;
  if (c->fsm && (flags & PF_STEP_SENSE)) {
    if (st == PF_STATE_SEARCHING) {
      float d2;
      po_nearest_food(c, x, y, &d2);
      float fr2 = c->food_radius * c->food_radius;
      if (d2 <= fr2) {
        st = PF_STATE_HOMING;
        carry = PF_META_CARRY;
        if (counters) counters[6] += 1;
      }
    } else if (st == PF_STATE_HOMING && delivered) {
      st = PF_STATE_SEARCHING;
      carry = 0;
      if (counters) counters[7] += 1;
    }
  }
  if (counters) counters[dec] += 1;

  *px = x;
  *py = y;
  *pth = th;
  *pwl = wl;
  *pwr = wr;
  *pmeta = (mt & ~(PF_META_STATE_MASK | PF_META_MOVED | PF_META_DEC_MASK | PF_META_CARRY)) | st |
           moved | ((uint32_t)dec << PF_META_DEC_SHIFT) | carry;
  return dec;
}

void po_agents_step(const pf_config* c, const float* f, uint32_t flags, int64_t n, float* x,
                    float* y, float* th, float* wl, float* wr, uint32_t* meta,
                    uint8_t* decisions_out, uint64_t* counters) {
  for (int64_t i = 0; i < n; ++i) {
    int d = po_agent_step_one(c, f, flags, x + i, y + i, th + i, wl + i, wr + i, meta + i, counters);
    if (decisions_out) decisions_out[i] = (uint8_t)d;
  }
}

/* ---------------------------------------------------------------- whole tick (not sharded) */

/* bufs[0], bufs[1]: the two field buffers; *cur_idx selects the current one and is flipped.
 * Order: deposit -> sense/decide -> evaporate(+diffuse) -> move (sup:575-607, 753-754). Before the
 * start delay only the field step and the Braitenberg walk run (sup:731-744).
 * counters: [0..5] decisions, [6] grabbed, [7] delivered, [8] deposits. */
void po_tick(const pf_config* c, float* buf0, float* buf1, int* cur_idx, int sense_enabled,
             int64_t n, float* x, float* y, float* th, float* wl, float* wr, uint32_t* meta,
             uint8_t* decisions_out, uint64_t* counters) {
  float* cur = *cur_idx ? buf1 : buf0;
  float* nxt = *cur_idx ? buf0 : buf1;
  uint32_t flags = PF_STEP_MOVE | (sense_enabled ? PF_STEP_SENSE : 0u);
  if (sense_enabled) {
    int64_t m = po_agents_deposit(c, cur, n, x, y, meta);
    if (counters) counters[8] += (uint64_t)m;
  }
  /* decide and move read the post-deposit, pre-evaporation field; move does not read the field,
   * so running it before the field step is equivalent to the reference order */
  po_agents_step(c, cur, flags, n, x, y, th, wl, wr, meta, decisions_out, counters);
  po_fill_ghosts(c, cur);
  po_step_field(c, cur, nxt);
  *cur_idx ^= 1;
}

/* ---------------------------------------------------------------- literal algo:351-356 */

/* float64 literal restatement of the dense 10x10 loop for the closed-form fixture
 * (sum after n ticks = 9*(1 - 0.9^n)): deposit 1.0 at (r,q) then multiply every cell
 * by (1 - 0.1), n times. */
double po_algo_dense_f64(int N, int r, int q, int nticks, double* grid /* N*N, zeroed by caller */) {
  const double intensity = 1.0, rate = 0.1;
  for (int t = 0; t < nticks; ++t) {
    grid[r * N + q] += intensity;
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) grid[i * N + j] *= (1 - rate);
  }
  double s = 0.0;
  for (int i = 0; i < N * N; ++i) s += grid[i];
  return s;
}
