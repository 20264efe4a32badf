// pf_trail.cu — the reference's SPARSE timestamped trail (the shipped supervisor) for N robots on the
// GPU: one thread per robot, fp64, per-robot list of live deposits in insertion order.
//
// Follows Phase-3/pheromone_algo_supervisor.py of the reference (`sup`):
//   storingPosition / has_moved   sup:279-289, 260-273      deposit_pheromone   sup:291-358
//   findNeighbours                sup:366-376               sense_pheromone     sup:379-417
//   PIDController.update          sup:199-226               adjust_robot_behavior sup:429-516
//   evaporate_pheromone           sup:519-525               tick order + age window sup:565-607, 627-663
//   reachedHome                   sup:665-670
// Arithmetic: fp64, every + and * separately rounded (__dadd_rn / __dmul_rn) in the reference's
// evaluation order; atan2/sin/cos/exp come from CUDA's fp64 math library (<= 2 ulp from glibc), so
// discrete outputs are exact and real-valued ones agree to ~1e-15 relative.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/pherofield.h"

struct pf_trail {
  int device;
  long long n;
  int cap;
  // entry arrays, [cap][n] (entry-major: coalesced across robots)
  double *et, *ei, *ex, *ey, *ez;
  int* n_live;
  int* n_global;
  double *cur, *prev;  // [3][n]
  unsigned char* have;  // bit0 have current, bit1 have previous
  double *wl, *wr;
  unsigned char* state;
  double* start;  // [3][n]
  double food[3];
  unsigned long long* dropped;  // deposits that found the robot's live list full (capacity too small)
  unsigned long long launches;
  char err[256];
};

static char g_trail_err[256] = "";

__device__ __forceinline__ double t_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double t_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double t_mul(double a, double b) { return __dmul_rn(a, b); }

__constant__ double kCoeff[8][2] = {{0.942, -0.22}, {0.63, -0.1}, {0.5, -0.06}, {-0.06, -0.06},
                                    {-0.06, -0.06}, {-0.06, 0.5}, {-0.19, 0.63}, {-0.13, 0.942}};  // sup:77

// PIDController.update with old_e = 0 (fresh controller per call, sup:437,485)
__device__ double trail_pid(double cx, double cy, double gx, double gy) {
  double dx = t_sub(gx, cx), dy = t_sub(gy, cy);
  double alpha = t_sub(atan2(dy, dx), 1.5707963267948966);  // np.radians(90)
  double e = atan2(sin(alpha), cos(alpha));
  double out = t_add(t_mul(1.0, e), t_mul(0.01, t_sub(e, 0.0)));  // kp*ep + kd*ed
  return atan2(sin(out), cos(out));
}

struct TrailOut {
  double* amount;        // NaN = "nothing deposited"
  int* n_global;
  signed char* direction;  // arg-max direction PF_DIR_*, -1 when sensing did not run
  signed char* decision;   // PF_DEC_*
  double *wl, *wr;
  int* n_live;
  double* sum_live;
  unsigned long long* masks;  // [5][n]: bit j = live entry j is in the direction's list (pre-evaporation)
  double *pre_time, *pre_intensity;  // [cap][n] snapshot of the entries the sense step saw
  int* n_pre;
};

__global__ void __launch_bounds__(128)
k_trail_tick(long long n, int cap, double t_now, double start_delay, const double* __restrict__ pos,
             const double* __restrict__ ps, double* et, double* ei, double* ex, double* ey, double* ez,
             int* n_live_a, int* n_global_a, double* cur, double* prev, unsigned char* have,
             double* wl_a, double* wr_a, const unsigned char* __restrict__ state_a,
             const double* __restrict__ start, double fx, double fy, TrailOut o,
             unsigned long long* __restrict__ dropped) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int st = state_a[r];
  int nl = n_live_a[r];
  // Controller.main only enters startPheromone once t >= 5 s (sup:731-735); before that, and for
  // IDLE / empty robots, nothing is stored, deposited or evaporated
  const bool active = (st == PF_STATE_HOMING) || (st == PF_STATE_SEARCHING && t_now >= start_delay);
  if (!active) {
    if (o.amount) o.amount[r] = __longlong_as_double(0x7ff8000000000000ll);
    if (o.n_global) o.n_global[r] = n_global_a[r];
    if (o.direction) o.direction[r] = -1;
    if (o.decision) o.decision[r] = PF_DEC_NONE;
    if (o.wl) o.wl[r] = wl_a[r];
    if (o.wr) o.wr[r] = wr_a[r];
    if (o.n_live) o.n_live[r] = nl;
    if (o.sum_live) o.sum_live[r] = 0.0;
    if (o.masks) for (int d = 0; d < 5; ++d) o.masks[d * n + r] = 0ull;
    if (o.n_pre) o.n_pre[r] = 0;
    return;
  }
  const double px = pos[3 * r], py = pos[3 * r + 1], pz = pos[3 * r + 2];
#define E(a, k) a[(long long)(k) * n + r]

  // ---- storingPosition + has_moved (xy only), sup:279-289, 260-273
  unsigned char hv = have[r];
  double opx = 0, opy = 0;
  bool has_prev = false;
  if (hv & 1) {
    opx = cur[r]; opy = cur[n + r];
    prev[r] = opx; prev[n + r] = opy; prev[2 * n + r] = cur[2 * n + r];
    has_prev = true;
    hv |= 2;
  }
  cur[r] = px; cur[n + r] = py; cur[2 * n + r] = pz;
  hv |= 1;
  have[r] = hv;
  bool moved = true;
  if (has_prev) {
    double dx = t_sub(px, opx), dy = t_sub(py, opy);
    moved = sqrt(t_add(t_mul(dx, dx), t_mul(dy, dy))) >= 0.0015;
  }

  // ---- deposit_pheromone, sup:291-358
  double amount = __longlong_as_double(0x7ff8000000000000ll);
  int ng = n_global_a[r];
  if (moved && (st == PF_STATE_SEARCHING || st == PF_STATE_HOMING)) {
    const bool same_key = nl > 0 && E(et, nl - 1) == t_now;  // dict overwrite of an existing time key
    if (!same_key) ng += 1;
    amount = (st == PF_STATE_SEARCHING) ? ((ng % 3 == 0) ? 20.0 : 1.0) : 5.0;
    int slot = same_key ? nl - 1 : nl;
    if (slot < cap) {
      E(et, slot) = t_now; E(ei, slot) = amount; E(ex, slot) = px; E(ey, slot) = py; E(ez, slot) = pz;
      if (!same_key) nl += 1;
    } else {
      atomicAdd(dropped, 1ull);  // live list full: the reference's dict would have grown (pf_trail_dropped)
    }
    n_global_a[r] = ng;
  }

  // ---- sense + adjust, sup:379-417, 429-516 (SEARCHING only after the start delay, sup:581)
  int direction = -1, decision = PF_DEC_NONE;
  double wl = wl_a[r], wr = wr_a[r];
  unsigned long long m_cur = 0, m_front = 0, m_back = 0, m_left = 0, m_right = 0;
  if (nl > 0) {
    const int c = nl - 1;  // the scan leaves the LAST inserted live entry, sup:382-384
    const double ct = E(et, c), cx = E(ex, c), cy = E(ey, c), cz = E(ez, c);
    m_cur = 1ull << c;
    m_back = 1ull << (c > 0 ? c - 1 : nl - 1);  // timestaps[index-1], negative index wraps, sup:394-397
    for (int k = 0; k < nl; ++k) {  // findNeighbours, sup:366-376
      double ax = t_sub(cx, E(ex, k)), ay = t_sub(cy, E(ey, k)), az = t_sub(cz, E(ez, k));
      double d = sqrt(t_add(t_add(t_add(0.0, t_mul(ax, ax)), t_mul(ay, ay)), t_mul(az, az)));
      if (d <= 0.1 && fabs(t_sub(E(et, k), ct)) <= 0.2) {
        if (t_sub(E(ey, k), py) >= 0) m_front |= 1ull << k;                          // sup:406-408
        if (E(ex, k) == px && E(ey, k) == t_sub(py, 1.0)) m_left |= 1ull << k;      // sup:409-411
        if (E(ex, k) == px && E(ey, k) == t_add(py, 1.0)) m_right |= 1ull << k;     // sup:412-414
      }
    }
    // ordered arg-max with strict >, first maximum wins, sup:440-450
    const unsigned long long masks[5] = {m_cur, m_front, m_back, m_left, m_right};
    bool any = false;
    double best = 0.0;
    for (int d = 0; d < 5; ++d)
      for (int k = 0; k < nl; ++k)
        if (masks[d] >> k & 1ull) {
          double v = E(ei, k);
          if (!any || v > best) { any = true; best = v; direction = d; }
        }
    if (st == PF_STATE_SEARCHING) {
      double turn = trail_pid(px, py, fx, fy);
      if (direction == PF_DIR_LEFT) {
        wl = fmax(0.0, t_sub(3.14, fabs(turn))); wr = 3.14; decision = PF_DEC_LEFT;
      } else if (direction == PF_DIR_RIGHT) {
        wr = fmax(0.0, t_sub(3.14, fabs(turn))); wl = 3.14; decision = PF_DEC_RIGHT;
      } else {
        wr = 1.0; wl = 1.0; decision = PF_DEC_STRAIGHT;
      }
      for (int j = 0; j < 8; ++j) {  // sup:468-470, same accumulation order
        double f = t_sub(1.0, ps[8 * r + j] / 512.0);
        wl = t_add(wl, t_mul(kCoeff[j][0], f));
        wr = t_add(wr, t_mul(kCoeff[j][1], f));
      }
    } else {  // HOMING, sup:475-514
      const double sx = start[r], sy = start[n + r], sz = start[2 * n + r];
      double ddx = t_sub(sx, px), ddy = t_sub(sy, py);
      double dist = sqrt(t_add(t_mul(ddx, ddx), t_mul(ddy, ddy)));
      double d2 = rint(t_mul(dist, 100.0));  // round(d, 2) in hundredths
      bool reached = (px == sx && py == sy && pz == sz) || (d2 >= 5.0 && d2 <= 6.0);
      if (!reached) {
        double turn = trail_pid(px, py, sx, sy);
        if (dist <= 0.7) {
          double rt = rint(t_mul(turn, 100.0));
          if (rt > -200.0 && rt < 0.0) { wl = 0.0; wr = 3.14; decision = PF_DEC_LEFT; }
          else if (rt < -200.0) { wr = 0.0; wl = 3.14; decision = PF_DEC_RIGHT; }
          else decision = PF_DEC_KEEP;
        } else {
          wr = 1.0; wl = 1.0; decision = PF_DEC_STRAIGHT;
        }
        for (int j = 0; j < 8; ++j) {
          double f = t_sub(1.0, ps[8 * r + j] / 512.0);
          wl = t_add(wl, t_mul(kCoeff[j][0], f));
          wr = t_add(wr, t_mul(kCoeff[j][1], f));
        }
      } else {
        wl = 0.0; wr = 0.0; decision = PF_DEC_STOP;
      }
    }
  }

  if (o.pre_time && o.pre_intensity) {
    for (int k = 0; k < nl; ++k) {
      o.pre_time[(long long)k * n + r] = E(et, k);
      o.pre_intensity[(long long)k * n + r] = E(ei, k);
    }
  }
  if (o.n_pre) o.n_pre[r] = nl;

  // ---- evaporation of entries aged 2..4 s, compounding every tick, delete <= 0.1, sup:594-607
  int w = 0;
  double sum = 0.0;
  for (int k = 0; k < nl; ++k) {
    double tk = E(et, k), ik = E(ei, k);
    double age = t_sub(t_now, tk);
    bool dead = false;
    if (age >= 2.0 && age <= 4.0) {
      ik = t_mul(ik, exp(t_mul(-0.1, age)));  // sup:521-524
      dead = ik <= 0.1;
    }
    if (!dead) {
      if (w != k) {
        E(et, w) = tk; E(ex, w) = E(ex, k); E(ey, w) = E(ey, k); E(ez, w) = E(ez, k);
      }
      E(ei, w) = ik;
      sum = t_add(sum, ik);
      ++w;
    }
  }
  n_live_a[r] = w;
  wl_a[r] = wl;
  wr_a[r] = wr;
#undef E
  if (o.amount) o.amount[r] = amount;
  if (o.n_global) o.n_global[r] = ng;
  if (o.direction) o.direction[r] = (signed char)direction;
  if (o.decision) o.decision[r] = (signed char)decision;
  if (o.wl) o.wl[r] = wl;
  if (o.wr) o.wr[r] = wr;
  if (o.n_live) o.n_live[r] = w;
  if (o.sum_live) o.sum_live[r] = sum;
  if (o.masks) {
    o.masks[r] = m_cur; o.masks[n + r] = m_front; o.masks[2 * n + r] = m_back;
    o.masks[3 * n + r] = m_left; o.masks[4 * n + r] = m_right;
  }
}

#define TR_CUDA(t, call)                                                                       \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      snprintf((t) ? (t)->err : g_trail_err, 256, "%s: %s", #call, cudaGetErrorString(e__));   \
      return e__ == cudaErrorMemoryAllocation ? PF_ENOMEM : PF_ECUDA;                          \
    }                                                                                          \
  } while (0)

extern "C" const char* pf_trail_last_error(const pf_trail* t) { return t ? t->err : g_trail_err; }

extern "C" int pf_trail_destroy(pf_trail* t) {
  if (!t) return PF_OK;
  cudaSetDevice(t->device);
  cudaDeviceSynchronize();
  void* ptrs[] = {t->et, t->ei, t->ex, t->ey, t->ez, t->n_live, t->n_global, t->cur, t->prev,
                  t->have, t->wl, t->wr, t->state, t->start, t->dropped};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete t;
  return PF_OK;
}

extern "C" int pf_trail_create(int device, int64_t n_robots, int capacity, pf_trail** out) {
  if (!out || n_robots <= 0 || capacity < 2 || capacity > 64) {
    snprintf(g_trail_err, 256, "pf_trail_create: need n_robots > 0 and 2 <= capacity <= 64");
    return PF_EINVAL;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    snprintf(g_trail_err, 256, "no CUDA device: pherofield has no CPU fallback");
    return PF_ECUDA;
  }
  pf_trail* t = new (std::nothrow) pf_trail();
  if (!t) return PF_ENOMEM;
  memset(t, 0, sizeof(*t));
  t->device = device; t->n = n_robots; t->cap = capacity;
  const size_t n = (size_t)n_robots, ce = n * capacity * sizeof(double);
#define TR_ALLOC(field, bytes)                                          \
  do {                                                                  \
    cudaError_t e__ = cudaMalloc((void**)&t->field, (bytes));           \
    if (e__ == cudaSuccess) e__ = cudaMemset(t->field, 0, (bytes));     \
    if (e__ != cudaSuccess) {                                           \
      snprintf(g_trail_err, 256, "alloc %s: %s", #field, cudaGetErrorString(e__)); \
      pf_trail_destroy(t);                                              \
      return PF_ENOMEM;                                                 \
    }                                                                   \
  } while (0)
  TR_CUDA((pf_trail*)nullptr, cudaSetDevice(device));
  TR_ALLOC(et, ce); TR_ALLOC(ei, ce); TR_ALLOC(ex, ce); TR_ALLOC(ey, ce); TR_ALLOC(ez, ce);
  TR_ALLOC(n_live, n * sizeof(int)); TR_ALLOC(n_global, n * sizeof(int));
  TR_ALLOC(cur, 3 * n * sizeof(double)); TR_ALLOC(prev, 3 * n * sizeof(double));
  TR_ALLOC(have, n); TR_ALLOC(wl, n * sizeof(double)); TR_ALLOC(wr, n * sizeof(double));
  TR_ALLOC(state, n); TR_ALLOC(start, 3 * n * sizeof(double));
  TR_ALLOC(dropped, sizeof(unsigned long long));
#undef TR_ALLOC
  TR_CUDA(t, cudaMemset(t->state, PF_STATE_SEARCHING, n));
  *out = t;
  return PF_OK;
}

// start_host: [n][3] fp64 (start_position = the robot's pose at construction, sup:248);
// food_host: [3]; state_host: [n] PF_STATE_* or NULL (all SEARCHING)
extern "C" int pf_trail_configure(pf_trail* t, const double* start_host, const double* food_host,
                                  const uint8_t* state_host) {
  if (!t || !start_host || !food_host) return PF_EINVAL;
  TR_CUDA(t, cudaSetDevice(t->device));
  const size_t n = (size_t)t->n;
  double* soa = new (std::nothrow) double[3 * n];
  if (!soa) return PF_ENOMEM;
  for (size_t r = 0; r < n; ++r)
    for (int k = 0; k < 3; ++k) soa[k * n + r] = start_host[3 * r + k];
  cudaError_t e = cudaMemcpy(t->start, soa, 3 * n * sizeof(double), cudaMemcpyHostToDevice);
  delete[] soa;
  TR_CUDA(t, e);
  memcpy(t->food, food_host, sizeof(t->food));
  if (state_host) TR_CUDA(t, cudaMemcpy(t->state, state_host, n, cudaMemcpyHostToDevice));
  return PF_OK;
}

extern "C" int pf_trail_set_state(pf_trail* t, int64_t robot, int state) {
  if (!t || robot < 0 || robot >= t->n || state < 0 || state > 3) return PF_EINVAL;
  TR_CUDA(t, cudaSetDevice(t->device));
  unsigned char s = (unsigned char)state;
  TR_CUDA(t, cudaMemcpy(t->state + robot, &s, 1, cudaMemcpyHostToDevice));
  return PF_OK;
}

extern "C" int pf_trail_tick(pf_trail* t, double t_now, double start_delay, const double* pos_xyz,
                             const double* ps, const pf_trail_out* out, void* stream) {
  if (!t || !pos_xyz || !ps) return PF_EINVAL;
  TR_CUDA(t, cudaSetDevice(t->device));
  TrailOut o;
  memset(&o, 0, sizeof(o));
  if (out) {
    o.amount = out->amount; o.n_global = out->n_global;
    o.direction = (signed char*)out->direction; o.decision = (signed char*)out->decision;
    o.wl = out->wl; o.wr = out->wr; o.n_live = out->n_live; o.sum_live = out->sum_live;
    o.masks = (unsigned long long*)out->masks;
    o.pre_time = out->pre_time; o.pre_intensity = out->pre_intensity; o.n_pre = out->n_pre;
  }
  const unsigned grid = (unsigned)((t->n + 127) / 128);
  k_trail_tick<<<grid, 128, 0, (cudaStream_t)stream>>>(t->n, t->cap, t_now, start_delay, pos_xyz, ps, t->et,
                                                      t->ei, t->ex, t->ey, t->ez, t->n_live, t->n_global,
                                                      t->cur, t->prev, t->have, t->wl, t->wr, t->state,
                                                      t->start, t->food[0], t->food[1], o, t->dropped);
  TR_CUDA(t, cudaGetLastError());
  t->launches++;
  return PF_OK;
}

// Deposits lost so far because a robot's live list was full (capacity <= 64 entries; the shipped supervisor
// peaks near 51 at dt = 0.064 s, a 32 ms tick or slower evaporation needs more): > 0 means sense results no
// longer match the reference's dict. Synchronises the stream.
extern "C" int pf_trail_dropped(pf_trail* t, uint64_t* n_dropped, void* stream) {
  if (!t || !n_dropped) return PF_EINVAL;
  TR_CUDA(t, cudaSetDevice(t->device));
  unsigned long long v = 0;
  TR_CUDA(t, cudaMemcpyAsync(&v, t->dropped, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  TR_CUDA(t, cudaStreamSynchronize((cudaStream_t)stream));
  *n_dropped = v;
  if (v) {
    snprintf(t->err, sizeof(t->err), "%llu deposits did not fit a robot's live list (capacity %d): results differ "
             "from the reference's dict from that tick on", v, t->cap);
    return PF_ESTATE;
  }
  return PF_OK;
}

// device pointers to the live entries ([capacity][n], entry-major) for read-out
extern "C" int pf_trail_entries(pf_trail* t, double** time, double** intensity, double** x, double** y,
                                double** z, int** n_live, int* capacity) {
  if (!t) return PF_EINVAL;
  if (time) *time = t->et;
  if (intensity) *intensity = t->ei;
  if (x) *x = t->ex;
  if (y) *y = t->ey;
  if (z) *z = t->ez;
  if (n_live) *n_live = t->n_live;
  if (capacity) *capacity = t->cap;
  return PF_OK;
}
