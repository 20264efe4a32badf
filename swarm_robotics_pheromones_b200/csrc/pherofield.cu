// pherofield.cu — the C-ABI library (include/pherofield.h) over the sm_100a kernels.
//
// Layout in HBM: two buffers (current / next), each [C][rows + 2][pitch] fp32, pitch = W rounded up
// to 32 floats. Local row r of channel c lives at buf + c*chan_stride + (r+1)*pitch; rows -1 and
// `rows` are ghost rows (neighbour strip's edge row, filled by the halo exchange; unused at the
// arena's outer edges where the stencil clamps = reflect).
#include <cub/cub.cuh>
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/pherofield.h"
#include "pf_agents.cuh"
#include "pf_device.cuh"
#include "pf_stencil.cuh"
#include "pf_stencil_tma.cuh"

static char g_create_err[512] = "";

struct GraphKey {
  const void *x, *y, *theta, *wl, *wr, *meta, *scratch;
  long long n, cap;
  uint32_t mask;
  int parity, variant, rpc, tile;
};

struct pf_handle {
  pf_config cfg;
  PFDev dev;
  float* buf[2];
  int cur;
  long long buf_elems;
  // scratch for sort / scan
  long long cap;
  char* scratch;  // 16 B per agent: keys_in, keys_out, vals_in, vals_out (or 2 x u64 for scans)
  void* cub_temp;
  size_t cub_temp_bytes;
  int key_bits;
  uint32_t sentinel;
  unsigned long long* counters;  // device, 16 entries
  double* d_sum;
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_join;
  uint64_t launches;
  double t0, start_delay;
  uint64_t ticks;
  int variant, rows_per_cta;
  int overlap_agents;
  int no_fused_keys;
  int use_graph;
  cudaGraph_t g_graph;
  cudaGraphExec_t g_exec;
  GraphKey g_key;
  uint64_t g_nodes;
  int n_sm, wave_keys, wave_agents;
  // owner-computes deposit (pf_deposit_owner.cuh): valid after pf_agents_sort for these arrays
  uint32_t *own_split, *own_cmin, *own_cmax, *own_pmax, *own_smin, *own_body, *own_tail;
  // deposit fused into the stencil (pf_deposit_tile.cuh): per-tile bins, per-tile lists, status flag
  uint32_t *tile_bcnt, *tile_brec, *tile_flag, *tile_count, *tile_pool;
  unsigned char* tile_blk;
  int* tile_dense;        // [tile] side-buffer slot of a crowded tile (pf_deposit_tile.cuh), -1 = none
  float* tile_side;       // [tile_pool_cap][64 * 1024]
  int tile_pool_cap;
  long long tile_alloc;   // tiles allocated
  int no_tile_deposit;
  bool tile_pending;      // lists are ready for (pf_step: were given to) the stencil launch over tile_ty0..ty1
  bool tile_handed;       // strips: that launch has happened; the next pf_agents_deposit skips those agents
  uint32_t own_handed_lo, own_handed_hi;  // ... and the sort-free deposit kernel skips key ranges inside these cells
  int tile_ty0, tile_ty1; // tile rows handed over
  unsigned long long tile_attempts, g_tile_attempts;
  long long own_alloc;      // chunks allocated
  long long own_n;          // agent slots the splitters were built for (0 = invalid)
  const float* own_x;
  int no_owner_deposit;
  int force_owner_deposit;
  int own_worth;              // host heuristic from the last pf_agents_sort (see there)
  double own_scan_estimate;
  // peer-store halo (NVLink P2P): neighbours' buffers / flag words, mapped into this process
  struct PeerLink { float* buf[2]; unsigned int* flags; long long rows, pitch, chan_stride; int set;
                    uint32_t* mig; long long mig_cap; } peer_up, peer_dn;
  unsigned int* peer_flags;  // device: halo epochs [0] written by the strip above, [1] by the strip below; migration [2], [3]
  unsigned int peer_epoch, mig_epoch;
  long long mig_cap;         // records per direction the migration mailbox (and the dummy send buffers) take
  uint32_t* mig_dummy;       // send target on a side without a neighbour (walls reflect: stays empty)
  void* ipc_opened[4];       // cudaIpcOpenMemHandle mappings to close in pf_destroy
  int n_ipc;
  unsigned long long peer_timeout_ns;
  int strip_dep0, strip_dep1;  // pf_strip_tick: rows handed over this tick (phase A decides, B and C use)
  bool strip_active;
  bool strip_one_launch;      // set while pf_strip_tick drives the strip in one-launch mode (see strip_tile_range)
  int strip_one_launch_opt;   // PF_OPT_STRIP_ONE_LAUNCH: 0 off, 1 the whole strip, 2 the handed rows and everything below them
  int strip_part_a_tiles;     // PF_OPT_STRIP_PART_A_TILES (0 = an eighth of the strip)
  int strip_mig_inline;       // PF_OPT_STRIP_MIG_INLINE: migration on the caller's stream instead of beside the stencil
  int strip_no_predeposit;    // PF_OPT_STRIP_NO_PREDEPOSIT
  bool strip_predep;          // the EARLY part of the next tick's deposit is enqueued (behind this tick's unpack)
  cudaEvent_t ev_edges;       // the strip's edge rows of this tick are stepped
  cudaEvent_t ev_a, ev_halo;  // phase B has begun (deposit done) / the ghost rows are in
  int strip_edges_side_opt;   // PF_OPT_STRIP_EDGES_SIDE
  bool strip_edges_side;      // this tick's small stencil launches (first rows, band below, edge rows) went to the side stream
  // pf_strip_tick phase marks (profiling enabled): kStripSlots ticks x kStripMarks events, folded into strip_ms
  cudaEvent_t* strip_ev;
  int strip_used, strip_mark;
  double strip_ms[16];
  unsigned long long strip_ticks;
  long long mig_flags_n;   // > 0: mig_cnt holds the per-chunk leaver counts the last move kernel made
  const float* mig_flags_x;
  uint32_t *mig_cnt, *mig_base, *emp_cnt, *emp_base;  // per chunk (two-level migration, pf_agents.cuh)
  long long* mig_slot_of;
  long long mig_slot_cap;
  TmaStencilState tma;
  // profiling: ring of per-tick events
  int prof_on;
  int prof_used;
  cudaEvent_t* prof_ev;  // kProfSlots x 5: main t0,t1,t2 ; side a0,a1
  unsigned char prof_has_agents[256];
  pf_profile prof;
  char err[512];
};

constexpr int kProfSlots = 256;
constexpr int kStripSlots = 64, kStripMarks = 12;

static int set_err(pf_handle* h, int code, const char* fmt, ...) {
  char* dst = h ? h->err : g_create_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

#define PF_CUDA(h, call)                                                                  \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess)                                                               \
      return set_err(h, PF_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                     __FILE__, __LINE__);                                                 \
  } while (0)

#define PF_LAUNCH_CHECK(h)                                                          \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess)                                                         \
      return set_err(h, PF_ECUDA, "kernel launch failed: %s (%s:%d)",               \
                     cudaGetErrorString(e__), __FILE__, __LINE__);                  \
    (h)->launches++;                                                                \
  } while (0)

static bool is_sharded(const pf_handle* h) { return h->cfg.row0 > 0 || h->cfg.row1 < h->cfg.H; }
static int strip_fold(pf_handle* h);
static long long mig_cap_of(const pf_config* c) { return c->migrate_cap > 0 ? c->migrate_cap : 16384; }
// byte offset of the migration mailbox inside the peer block: behind the flag words and the halo rows
static size_t mig_offset(int C, long long pitch) {
  return (256 + sizeof(float) * 4 * (size_t)C * (size_t)pitch + 255) & ~(size_t)255;
}
static size_t mig_box_words(long long cap) { return (size_t)(cap + 1) * PF_MIGRATE_WORDS; }
static void graph_drop(pf_handle* h);

static void tile_free(pf_handle* h) {
  void** arrs[8] = {(void**)&h->tile_bcnt, (void**)&h->tile_brec, (void**)&h->tile_flag, (void**)&h->tile_count,
                    (void**)&h->tile_blk, (void**)&h->tile_pool, (void**)&h->tile_dense, (void**)&h->tile_side};
  for (auto pp : arrs) { if (*pp) cudaFree(*pp); *pp = nullptr; }
  h->tile_alloc = 0;
}

static inline unsigned blocks_for(long long n, int threads) {
  return (unsigned)((n + threads - 1) / threads);
}
// grid for the grid-stride agent kernels: exactly one resident wave (SMs x CTAs/SM from the
// occupancy calculator), so no partial second wave
static inline unsigned wave_grid(long long n, int wave) {
  unsigned b = blocks_for(n, 256);
  return b < (unsigned)wave ? b : (unsigned)wave;
}

static int ceil_log2_u64(unsigned long long v) {
  int b = 0;
  while ((1ull << b) < v) ++b;
  return b < 1 ? 1 : b;
}

extern "C" int pf_abi_version(void) { return PF_ABI_VERSION; }

extern "C" int pf_config_default(pf_config* c, int32_t H, int32_t W, int32_t C) {
  if (!c || H <= 0 || W <= 0 || C <= 0 || C > PF_MAX_CHANNELS) return PF_EINVAL;
  memset(c, 0, sizeof(*c));
  c->abi_version = PF_ABI_VERSION;
  c->device = 0;
  c->H = H; c->W = W; c->C = C;
  c->row0 = 0; c->row1 = H;
  c->cell_size = 0.01f;
  c->origin_x = 0.0f; c->origin_y = 0.0f;
  c->stencil = 0;
  for (int k = 0; k < PF_MAX_CHANNELS; ++k) {
    c->evap_keep[k] = (float)(1.0 - 0.1);  // reference algo:248,356
    c->diffusion[k] = 0.0f;
  }
  c->delete_threshold = 0.0f;
  c->intensity_explore = 1.0f;   // sup:232
  c->intensity_high = 20.0f;     // sup:233
  c->intensity_homing = 5.0f;    // sup:234
  c->high_every = 3;             // sup:316
  c->move_threshold = 0.0015f;   // sup:260
  if (C >= 2) {
    c->deposit_channel[PF_STATE_SEARCHING] = 0;
    c->deposit_channel[PF_STATE_HOMING] = 1;
    c->sense_channel[PF_STATE_SEARCHING] = 1;
    c->sense_channel[PF_STATE_HOMING] = 0;
  }
  c->sense_mode = PF_SENSE_HEADING;
  const int drow[5] = {0, 0, 0, -1, 1}, dcol[5] = {0, 1, -1, 0, 0};
  for (int d = 0; d < 5; ++d) { c->sense_drow[d] = drow[d]; c->sense_dcol[d] = dcol[d]; }
  c->sense_dist = c->cell_size;
  c->pid_kp = 1.0f; c->pid_kd = 0.01f;  // sup:437
  c->cruise_speed = 1.0f;               // sup:197
  c->max_speed = 3.14f;                 // sup:74
  // sum_j coeff[j][i] * (1 - 0/512), accumulated in the order of sup:468-470
  const double coeff[8][2] = {{0.942, -0.22}, {0.63, -0.1}, {0.5, -0.06}, {-0.06, -0.06},
                              {-0.06, -0.06}, {-0.06, 0.5}, {-0.19, 0.63}, {-0.13, 0.942}};
  double bl = 0.0, br = 0.0;
  for (int j = 0; j < 8; ++j) { bl += coeff[j][0] * (1.0 - (0.0 / (1024 / 2))); br += coeff[j][1] * (1.0 - (0.0 / (1024 / 2))); }
  c->braitenberg_l = (float)bl; c->braitenberg_r = (float)br;
  c->home_radius = 0.7f; c->home_lo_cm = 5.0f; c->home_hi_cm = 6.0f;  // sup:489,667
  c->nest_x = (float)(0.5 * H * 0.01); c->nest_y = (float)(0.5 * W * 0.01);
  c->n_food = 1;
  c->food_x[0] = (float)(0.75 * H * 0.01); c->food_y[0] = (float)(0.75 * W * 0.01);
  c->food_radius = 0.05f;
  c->fsm = 0;
  c->dt = 0.064f;
  c->wheel_radius = 0.02f; // E-puck.proto:100
  c->axle_length = 0.052f; // E-puck.proto:86
  return PF_OK;
}

static int validate(const pf_config* c) {
  if (c->abi_version != PF_ABI_VERSION) return set_err(nullptr, PF_EINVAL, "abi_version %d != %d", c->abi_version, PF_ABI_VERSION);
  if (c->H <= 0 || c->W <= 0 || c->C <= 0 || c->C > PF_MAX_CHANNELS) return set_err(nullptr, PF_EINVAL, "bad H/W/C");
  if (c->row0 < 0 || c->row1 > c->H || c->row0 >= c->row1) return set_err(nullptr, PF_EINVAL, "bad strip [%d,%d)", c->row0, c->row1);
  if (!(c->cell_size > 0.0f)) return set_err(nullptr, PF_EINVAL, "cell_size must be > 0");
  if (c->stencil != 0 && c->stencil != 5 && c->stencil != 9) return set_err(nullptr, PF_EINVAL, "stencil must be 0, 5 or 9");
  if (c->high_every <= 0) return set_err(nullptr, PF_EINVAL, "high_every must be > 0");
  if (c->n_food < 1 || c->n_food > PF_MAX_FOOD) return set_err(nullptr, PF_EINVAL, "n_food out of range");
  for (int s = 0; s < 4; ++s)
    if (c->deposit_channel[s] < 0 || c->deposit_channel[s] >= c->C || c->sense_channel[s] < 0 || c->sense_channel[s] >= c->C)
      return set_err(nullptr, PF_EINVAL, "channel routing out of range");
  if (c->sense_mode != PF_SENSE_WORLD && c->sense_mode != PF_SENSE_HEADING) return set_err(nullptr, PF_EINVAL, "bad sense_mode");
  if (c->sense_drow[0] != 0 || c->sense_dcol[0] != 0) return set_err(nullptr, PF_EINVAL, "the 'current' sense offset must be (0,0)");
  bool sharded = c->row0 > 0 || c->row1 < c->H;
  if (sharded) {  // ghost depth is one row
    if (c->sense_mode == PF_SENSE_WORLD) {
      for (int d = 0; d < 5; ++d)
        if (c->sense_drow[d] < -1 || c->sense_drow[d] > 1) return set_err(nullptr, PF_EINVAL, "sharded field: |sense_drow| must be <= 1");
    } else if (c->sense_dist > c->cell_size) {
      return set_err(nullptr, PF_EINVAL, "sharded field: sense_dist must be <= cell_size");
    }
  }
  if (c->migrate_cap < 0 || c->migrate_cap > (1 << 24)) return set_err(nullptr, PF_EINVAL, "migrate_cap out of range");
  unsigned long long cells = (unsigned long long)c->C * (unsigned long long)(c->row1 - c->row0) * (unsigned long long)c->W;
  if (cells >= 0xffffffffull) return set_err(nullptr, PF_EINVAL, "more than 2^32-1 cells per handle: shard the field");
  return PF_OK;
}

static void derive(pf_handle* h) {
  const pf_config& c = h->cfg;
  PFDev& d = h->dev;
  d.H = c.H; d.W = c.W; d.C = c.C; d.row0 = c.row0; d.row1 = c.row1; d.rows = c.row1 - c.row0;
  d.pitch = ((long long)c.W + 31) / 32 * 32;
  d.chan_stride = (long long)(d.rows + 2) * d.pitch;
  d.ox = c.origin_x; d.oy = c.origin_y;
  d.inv_cs = 1.0f / c.cell_size;
  d.xmin = c.origin_x; d.ymin = c.origin_y;
  d.xmax = c.origin_x + (float)c.H * c.cell_size;
  d.ymax = c.origin_y + (float)c.W * c.cell_size;
  d.stencil = c.stencil;
  for (int k = 0; k < PF_MAX_CHANNELS; ++k) {
    d.keep[k] = c.evap_keep[k];
    d.dk[k] = (c.stencil == 9) ? c.diffusion[k] / 6.0f : c.diffusion[k];
  }
  d.del_thr = c.delete_threshold;
  d.i_explore = c.intensity_explore; d.i_high = c.intensity_high; d.i_homing = c.intensity_homing;
  d.dep_red = (c.intensity_explore >= 1e-30f && c.intensity_high >= 1e-30f && c.intensity_homing >= 1e-30f) ? 1 : 0;
  {  // the reference's amounts are small integers (1, 20, 5): per-cell totals are then exact in any order
    const float v[3] = {c.intensity_explore, c.intensity_high, c.intensity_homing};
    d.dep_exact = 1;
    for (float x : v)
      if (!(x >= 1.0f && x <= 32768.0f && x == (float)(int)x)) d.dep_exact = 0;
  }
  d.high_every = (unsigned)c.high_every;
  d.move_thr = c.move_threshold;
  for (int s = 0; s < 4; ++s) { d.dep_ch[s] = c.deposit_channel[s]; d.sen_ch[s] = c.sense_channel[s]; }
  d.sense_mode = c.sense_mode;
  for (int k = 0; k < 5; ++k) { d.drow[k] = c.sense_drow[k]; d.dcol[k] = c.sense_dcol[k]; }
  d.sense_dist = c.sense_dist;
  d.kp = c.pid_kp; d.kd = c.pid_kd; d.cruise = c.cruise_speed; d.vmax = c.max_speed;
  d.bl = c.braitenberg_l; d.br = c.braitenberg_r;
  d.home_radius = c.home_radius; d.home_lo = c.home_lo_cm; d.home_hi = c.home_hi_cm;
  d.nest_x = c.nest_x; d.nest_y = c.nest_y;
  d.n_food = c.n_food;
  for (int k = 0; k < PF_MAX_FOOD; ++k) { d.food_x[k] = c.food_x[k]; d.food_y[k] = c.food_y[k]; }
  d.food_r2 = c.food_radius * c.food_radius;
  d.fsm = c.fsm;
  d.dt = c.dt; d.wheel_r = c.wheel_radius; d.axle = c.axle_length;
  h->buf_elems = (long long)c.C * d.chan_stride;
  unsigned long long cells = (unsigned long long)c.C * d.rows * c.W;
  h->sentinel = (uint32_t)cells;
  h->key_bits = ceil_log2_u64(cells + 1);
}

extern "C" int pf_create(const pf_config* cfg, pf_handle** out) {
  if (!cfg || !out) return set_err(nullptr, PF_EINVAL, "null argument");
  int rc = validate(cfg);
  if (rc) return rc;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return set_err(nullptr, PF_ECUDA, "no CUDA device (%s): pherofield has no CPU fallback",
                   e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return set_err(nullptr, PF_EINVAL, "device %d of %d", cfg->device, ndev);
  pf_handle* h = new (std::nothrow) pf_handle();
  if (!h) return set_err(nullptr, PF_ENOMEM, "host allocation failed");
  memset(h, 0, sizeof(*h));
  h->cfg = *cfg;
  derive(h);
#define CREATE_CUDA(call)                                                                  \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      set_err(nullptr, e__ == cudaErrorMemoryAllocation ? PF_ENOMEM : PF_ECUDA, "%s: %s", #call, \
              cudaGetErrorString(e__));                                                    \
      pf_destroy(h);                                                                       \
      return e__ == cudaErrorMemoryAllocation ? PF_ENOMEM : PF_ECUDA;                      \
    }                                                                                      \
  } while (0)
  CREATE_CUDA(cudaSetDevice(cfg->device));
  {  // L2 fetch granularity: the agent kernels touch 4 B per 32 B sector at random; the default 64 B
     // granularity doubles their DRAM traffic. PF_L2_FETCH=32|64|128 overrides (device-wide limit).
    const char* e = getenv("PF_L2_FETCH");
    if (e && *e) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));
    cudaGetLastError();
  }
  CREATE_CUDA(cudaMalloc(&h->buf[0], sizeof(float) * h->buf_elems));
  CREATE_CUDA(cudaMalloc(&h->buf[1], sizeof(float) * h->buf_elems));
  CREATE_CUDA(cudaMemset(h->buf[0], 0, sizeof(float) * h->buf_elems));
  CREATE_CUDA(cudaMemset(h->buf[1], 0, sizeof(float) * h->buf_elems));
  CREATE_CUDA(cudaMalloc(&h->counters, sizeof(unsigned long long) * 16));
  CREATE_CUDA(cudaMemset(h->counters, 0, sizeof(unsigned long long) * 16));
  CREATE_CUDA(cudaMalloc(&h->d_sum, sizeof(double)));
  {  // peer mailbox block: 256 B of flag words + [2 buffer parities][2 sides][C][pitch] floats. Neighbour strips
     // store edge rows here (small, IPC-mapped) rather than into the field allocation itself: mapping the
     // multi-GB field for peer access cost the stencil 40 % (measured), presumably through smaller pages
    // ... followed by the migration mailbox: two record buffers the neighbours pack their leavers into
    h->mig_cap = mig_cap_of(cfg);
    const size_t bytes = mig_offset(cfg->C, h->dev.pitch) + 2 * sizeof(uint32_t) * mig_box_words(h->mig_cap);
    CREATE_CUDA(cudaMalloc(&h->peer_flags, bytes));
    CREATE_CUDA(cudaMemset(h->peer_flags, 0, bytes));
    CREATE_CUDA(cudaMalloc(&h->mig_dummy, sizeof(uint32_t) * mig_box_words(h->mig_cap)));
    CREATE_CUDA(cudaMemset(h->mig_dummy, 0, sizeof(uint32_t) * mig_box_words(h->mig_cap)));
    h->peer_timeout_ns = 5000000000ull;  // a neighbour that does not signal within 5 s is counted as dead
  }
  {
    int lo = 0, hi = 0;  // the side stream outranks the caller's: its small kernels are not starved
    CREATE_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CREATE_CUDA(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi));
  }
  CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_edges, cudaEventDisableTiming));
  CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_a, cudaEventDisableTiming));
  CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
  {
    int occ = 0;
    CREATE_CUDA(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, cfg->device));
    CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_deposit_keys_agents, 256, 0));
    h->wave_keys = h->n_sm * (occ > 0 ? occ : 1);
    CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_agents_step, 256, 0));
    h->wave_agents = h->n_sm * (occ > 0 ? occ : 1);
  }
  CREATE_CUDA(cudaDeviceSynchronize());
#undef CREATE_CUDA
  h->variant = 0;
  h->rows_per_cta = 0;
  // launch-bound problems (field + agents small enough that a tick is a few tens of microseconds)
  h->use_graph = (h->buf_elems <= (1ll << 24)) ? 1 : 0;
  tma_stencil_init(&h->tma);
  *out = h;
  return PF_OK;
}

extern "C" int pf_destroy(pf_handle* h) {
  if (!h) return PF_OK;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  if (h->buf[0]) cudaFree(h->buf[0]);
  if (h->buf[1]) cudaFree(h->buf[1]);
  if (h->scratch) cudaFree(h->scratch);
  if (h->cub_temp) cudaFree(h->cub_temp);
  if (h->own_split) cudaFree(h->own_split);
  if (h->own_cmin) cudaFree(h->own_cmin);
  if (h->own_cmax) cudaFree(h->own_cmax);
  if (h->own_pmax) cudaFree(h->own_pmax);
  if (h->own_smin) cudaFree(h->own_smin);
  if (h->own_body) cudaFree(h->own_body);
  if (h->own_tail) cudaFree(h->own_tail);
  if (h->mig_cnt) cudaFree(h->mig_cnt);
  if (h->mig_base) cudaFree(h->mig_base);
  if (h->emp_cnt) cudaFree(h->emp_cnt);
  if (h->emp_base) cudaFree(h->emp_base);
  if (h->mig_slot_of) cudaFree(h->mig_slot_of);
  tile_free(h);
  if (h->counters) cudaFree(h->counters);
  if (h->d_sum) cudaFree(h->d_sum);
  for (int k = 0; k < h->n_ipc; ++k) cudaIpcCloseMemHandle(h->ipc_opened[k]);
  if (h->peer_flags) cudaFree(h->peer_flags);
  if (h->mig_dummy) cudaFree(h->mig_dummy);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->ev_edges) cudaEventDestroy(h->ev_edges);
  if (h->ev_a) cudaEventDestroy(h->ev_a);
  if (h->ev_halo) cudaEventDestroy(h->ev_halo);
  if (h->g_exec) cudaGraphExecDestroy(h->g_exec);
  if (h->g_graph) cudaGraphDestroy(h->g_graph);
  if (h->prof_ev) {
    for (int i = 0; i < kProfSlots * 6; ++i) cudaEventDestroy(h->prof_ev[i]);
    delete[] h->prof_ev;
  }
  if (h->strip_ev) {
    for (int i = 0; i < kStripSlots * kStripMarks; ++i) cudaEventDestroy(h->strip_ev[i]);
    delete[] h->strip_ev;
  }
  delete h;
  return PF_OK;
}

extern "C" const char* pf_last_error(const pf_handle* h) { return h ? h->err : g_create_err; }

extern "C" int pf_get_config(const pf_handle* h, pf_config* out) {
  if (!h || !out) return PF_EINVAL;
  *out = h->cfg;
  return PF_OK;
}

extern "C" int pf_reserve_agents(pf_handle* h, int64_t max_agents) {
  if (!h || max_agents < 0) return PF_EINVAL;
  if (max_agents <= h->cap) return PF_OK;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  PF_CUDA(h, cudaDeviceSynchronize());
  if (h->scratch) cudaFree(h->scratch);
  if (h->cub_temp) cudaFree(h->cub_temp);
  h->scratch = nullptr; h->cub_temp = nullptr; h->cap = 0; h->mig_flags_n = 0;
  long long cap = (max_agents + 1023) / 1024 * 1024;
  PF_CUDA(h, cudaMalloc(&h->scratch, (size_t)cap * 20));  // + 4 B/agent for the permute pass
  size_t sort_bytes = 0, scan_bytes = 0;
  PF_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                             (const float*)nullptr, (float*)nullptr, (int)cap, 0, 32, (cudaStream_t)0));
  PF_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const unsigned long long*)nullptr,
                                           (unsigned long long*)nullptr, (int)cap, (cudaStream_t)0));
  h->cub_temp_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  PF_CUDA(h, cudaMalloc(&h->cub_temp, h->cub_temp_bytes));
  h->cap = cap;
  {  // per-chunk bounds and splitters of the owner-computes deposit
    uint32_t** arrs[7] = {&h->own_split, &h->own_cmin, &h->own_cmax, &h->own_pmax, &h->own_smin, &h->own_body,
                          &h->own_tail};
    for (auto pp : arrs) { if (*pp) cudaFree(*pp); *pp = nullptr; }
    h->own_alloc = cap / kOwnChunk + 2;
    for (auto pp : arrs) PF_CUDA(h, cudaMalloc(pp, sizeof(uint32_t) * h->own_alloc));
    h->own_n = 0;
    uint32_t** migs[4] = {&h->mig_cnt, &h->mig_base, &h->emp_cnt, &h->emp_base};
    for (auto pp : migs) {
      if (*pp) cudaFree(*pp);
      *pp = nullptr;
      PF_CUDA(h, cudaMalloc(pp, sizeof(uint32_t) * 2 * h->own_alloc));
    }
  }
  return PF_OK;
}

// ---------------------------------------------------------------------------------------------
// field access
// ---------------------------------------------------------------------------------------------
extern "C" int pf_field_ptr(pf_handle* h, int which, int channel, float** ptr, int64_t* pitch) {
  if (!h || !ptr || channel < 0 || channel >= h->cfg.C || (which != 0 && which != 1)) return PF_EINVAL;
  float* b = h->buf[which ? (h->cur ^ 1) : h->cur];
  *ptr = b + (long long)channel * h->dev.chan_stride + h->dev.pitch;
  if (pitch) *pitch = h->dev.pitch;
  return PF_OK;
}

extern "C" int pf_field_fill(pf_handle* h, int channel, float value, void* stream) {
  if (!h || channel < -1 || channel >= h->cfg.C) return PF_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  float* b = h->buf[h->cur];
  long long n = channel < 0 ? h->buf_elems : h->dev.chan_stride;
  if (channel >= 0) b += (long long)channel * h->dev.chan_stride;
  k_fill<<<592, 256, 0, s>>>(b, n, value);
  PF_LAUNCH_CHECK(h);
  return PF_OK;
}

extern "C" int pf_field_upload_host(pf_handle* h, int channel, const float* src, void* stream) {
  if (!h || !src || channel < 0 || channel >= h->cfg.C) return PF_EINVAL;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  float* dst = h->buf[h->cur] + (long long)channel * h->dev.chan_stride + h->dev.pitch;
  PF_CUDA(h, cudaMemcpy2DAsync(dst, sizeof(float) * h->dev.pitch, src, sizeof(float) * h->cfg.W,
                               sizeof(float) * h->cfg.W, h->dev.rows, cudaMemcpyHostToDevice,
                               (cudaStream_t)stream));
  return PF_OK;
}

extern "C" int pf_snapshot_host(pf_handle* h, int channel, float* dst, void* stream) {
  if (!h || !dst || channel < 0 || channel >= h->cfg.C) return PF_EINVAL;
  if (h->tile_handed || h->tile_pending)
    return set_err(h, PF_ESTATE, "%s after a hand-over tick (PF_STEP_HANDOVER): the interior rows and the deposit "
                                 "counters already hold the NEXT tick's deposits; run the last tick before reading "
                                 "without PF_STEP_HANDOVER", "pf_snapshot_host");
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  const float* src = h->buf[h->cur] + (long long)channel * h->dev.chan_stride + h->dev.pitch;
  PF_CUDA(h, cudaMemcpy2DAsync(dst, sizeof(float) * h->cfg.W, src, sizeof(float) * h->dev.pitch,
                               sizeof(float) * h->cfg.W, h->dev.rows, cudaMemcpyDeviceToHost,
                               (cudaStream_t)stream));
  PF_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
  return PF_OK;
}

extern "C" int pf_fill_ghosts(pf_handle* h, void* stream) {
  if (!h) return PF_EINVAL;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  int top = h->cfg.row0 == 0, bottom = h->cfg.row1 == h->cfg.H;
  if (!top && !bottom) return PF_OK;
  dim3 grid(blocks_for(h->cfg.W, 256), h->cfg.C);
  k_fill_ghosts<<<grid, 256, 0, (cudaStream_t)stream>>>(h->dev, h->buf[h->cur], top, bottom);
  PF_LAUNCH_CHECK(h);
  return PF_OK;
}

struct RowSumOp {
  const float* base;
  long long pitch;
  int W;
  __device__ double operator()(long long k) const {
    long long r = k / W;
    int c = (int)(k - r * W);
    return (double)base[r * pitch + c];
  }
};

extern "C" int pf_field_sum_host(pf_handle* h, int channel, double* sum_host, void* stream) {
  if (!h || !sum_host || channel < 0 || channel >= h->cfg.C) return PF_EINVAL;
  if (h->tile_handed || h->tile_pending)
    return set_err(h, PF_ESTATE, "%s after a hand-over tick (PF_STEP_HANDOVER): the interior rows and the deposit "
                                 "counters already hold the NEXT tick's deposits; run the last tick before reading "
                                 "without PF_STEP_HANDOVER", "pf_field_sum_host");
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  RowSumOp op{h->buf[h->cur] + (long long)channel * h->dev.chan_stride + h->dev.pitch, h->dev.pitch, h->cfg.W};
  cub::CountingInputIterator<long long> cnt(0);
  cub::TransformInputIterator<double, RowSumOp, cub::CountingInputIterator<long long>> it(cnt, op);
  long long n = (long long)h->dev.rows * h->cfg.W;
  if (n > 0x7fffffffll) return set_err(h, PF_EINVAL, "pf_field_sum_host: more than 2^31 cells per channel");
  size_t bytes = 0;
  PF_CUDA(h, cub::DeviceReduce::Sum(nullptr, bytes, it, h->d_sum, (int)n, s));
  void* tmp = nullptr;
  PF_CUDA(h, cudaMallocAsync(&tmp, bytes, s));
  PF_CUDA(h, cub::DeviceReduce::Sum(tmp, bytes, it, h->d_sum, (int)n, s));
  PF_CUDA(h, cudaFreeAsync(tmp, s));
  PF_CUDA(h, cudaMemcpyAsync(sum_host, h->d_sum, sizeof(double), cudaMemcpyDeviceToHost, s));
  PF_CUDA(h, cudaStreamSynchronize(s));
  return PF_OK;
}

// ---------------------------------------------------------------------------------------------
// K1 launch
// ---------------------------------------------------------------------------------------------
// neighbour ghost rows of physical buffer `k` (see PeerRows)
// this strip's mailbox for buffer parity k: side 0 = rows arriving from the strip above, side 1 = from below
static float* mailbox_of(unsigned int* flags_base, int k, int C, long long pitch) {
  return (float*)((char*)flags_base + 256) + (size_t)k * 2 * C * pitch;
}

// where this strip's edge rows of physical buffer `k` go in the neighbours' mailboxes (see PeerRows)
static PeerRows peer_rows(const pf_handle* h, int k) {
  PeerRows r;
  r.up = nullptr; r.dn = nullptr; r.up_cs = 0; r.dn_cs = 0;
  if (h->peer_up.set) {  // the strip above receives it as "from below" (side 1)
    r.up = h->peer_up.buf[k] + (size_t)h->cfg.C * h->peer_up.pitch;
    r.up_cs = h->peer_up.pitch;
  }
  if (h->peer_dn.set) {  // the strip below receives it as "from above" (side 0)
    r.dn = h->peer_dn.buf[k];
    r.dn_cs = h->peer_dn.pitch;
  }
  return r;
}

// with_deposits: the next tick's deposits ride along (the lists of the tile rows tile_ty0..ty1 are ready).
// whole_strip (with_deposits only): the launch covers every tile row of the strip, edge rows with their peer
// stores included — the tiles outside tile_ty0..ty1 have empty lists.
// to_end (with_deposits only): from the first handed tile row to the strip's last row, its peer stores included.
static int launch_stencil(pf_handle* h, int r_begin, int r_end, cudaStream_t s, bool with_deposits = false,
                          bool whole_strip = false, bool to_end = false) {
  if (r_begin >= r_end) return PF_OK;
  const PFDev& d = h->dev;
  StencilArgs a;
  a.cur = h->buf[h->cur];
  a.nxt = h->buf[h->cur ^ 1];
  a.W = d.W; a.rows = d.rows; a.pitch = d.pitch; a.chan_stride = d.chan_stride;
  a.r_begin = r_begin; a.r_end = r_end;
  a.row_lo = (h->cfg.row0 == 0) ? 0 : -1;
  a.row_hi = (h->cfg.row1 == h->cfg.H) ? d.rows - 1 : d.rows;
  for (int k = 0; k < PF_MAX_CHANNELS; ++k) { a.keep[k] = d.keep[k]; a.dk[k] = d.dk[k]; }
  a.del_thr = d.del_thr;
  a.peer = peer_rows(h, h->cur ^ 1);
  a.dep_flag = nullptr; a.dep_count = nullptr; a.dep_blk = nullptr; a.dep_ty0 = 0; a.dep_ny = 0;
  a.dep_dense = nullptr; a.dep_side = nullptr;
  if (with_deposits) {  // tile_ok() / strip_tile_range() hold: whole tile rows, TMA ring, CTA = deposit tile
    a.dep_flag = h->tile_flag; a.dep_count = h->tile_count; a.dep_blk = h->tile_blk;
    a.dep_dense = h->tile_dense; a.dep_side = h->tile_side;
    const int ny = (d.rows + kTileRows - 1) / kTileRows;
    int rc = whole_strip ? tma_stencil_launch_dep(&h->tma, d, a, 0, ny, true, s)
             : to_end    ? tma_stencil_launch_dep(&h->tma, d, a, h->tile_ty0, ny, true, s)
                         : tma_stencil_launch_dep(&h->tma, d, a, h->tile_ty0, h->tile_ty1, false, s);
    if (rc != PF_OK) return set_err(h, rc, "TMA stencil launch failed: %s", h->tma.err);
    h->launches++;
    return PF_OK;
  }
  const int nrows = r_end - r_begin;
  if (d.W % 4 != 0) {
    dim3 grid(blocks_for(d.W, 256), nrows, d.C);
    if (d.stencil == 0) k_stencil_scalar<0><<<grid, 256, 0, s>>>(a);
    else if (d.stencil == 5) k_stencil_scalar<5><<<grid, 256, 0, s>>>(a);
    else k_stencil_scalar<9><<<grid, 256, 0, s>>>(a);
    PF_LAUNCH_CHECK(h);
    return PF_OK;
  }
  int variant = h->variant;
  // auto: measured on B200 at 32768^2 x 2 (profiles/r01_k1_sweep.md): the TMA ring reaches 6535 GB/s
  // (5-point) / 6345 GB/s (9-point) against 5236 / 5055 for the register window; pure evaporation
  // (no neighbours) is a plain stream and the register window wins (6814 GB/s).
  if (variant == 0) variant = (d.stencil != 0 && d.W >= 512) ? 2 : 1;
  if (variant >= 2) {
    int rc = tma_stencil_launch(&h->tma, variant - 2, h->cfg, d, a, h->rows_per_cta, s);
    if (rc == PF_OK) { h->launches++; return PF_OK; }
    if (rc != PF_ESTATE) return set_err(h, rc, "TMA stencil launch failed: %s", h->tma.err);
    // PF_ESTATE: shape not supported by the TMA variant -> register sliding window
  }
  int rpc = h->rows_per_cta;
  const int xblocks = (d.W / 4 + kStencilThreads - 1) / kStencilThreads;
  if (rpc <= 0) {
    // aim for >= 8 waves of 148 SMs x 8 CTAs, at least 8 rows per CTA, at most 128
    long long want = 148ll * 8 * 8;
    long long chunks = want / ((long long)xblocks * d.C);
    if (chunks < 1) chunks = 1;
    rpc = (int)((nrows + chunks - 1) / chunks);
    if (rpc < 8) rpc = 8;
    if (rpc > 128) rpc = 128;
  }
  a.rows_per_cta = rpc;
  dim3 grid(xblocks, (nrows + rpc - 1) / rpc, d.C);
  const bool peer = (a.peer.up && a.r_begin == 0) || (a.peer.dn && a.r_end == a.rows);
  if (peer) {
    if (d.stencil == 0) k_stencil_rows<0, true><<<grid, kStencilThreads, 0, s>>>(a);
    else if (d.stencil == 5) k_stencil_rows<5, true><<<grid, kStencilThreads, 0, s>>>(a);
    else k_stencil_rows<9, true><<<grid, kStencilThreads, 0, s>>>(a);
  } else {
    if (d.stencil == 0) k_stencil_rows<0, false><<<grid, kStencilThreads, 0, s>>>(a);
    else if (d.stencil == 5) k_stencil_rows<5, false><<<grid, kStencilThreads, 0, s>>>(a);
    else k_stencil_rows<9, false><<<grid, kStencilThreads, 0, s>>>(a);
  }
  PF_LAUNCH_CHECK(h);
  return PF_OK;
}


extern "C" int pf_step_field(pf_handle* h, int nsteps, void* stream) {
  if (!h || nsteps < 0) return PF_EINVAL;
  if (is_sharded(h) && nsteps > 1)
    return set_err(h, PF_ESTATE, "sharded handle: one step per halo exchange (nsteps must be 1)");
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  for (int i = 0; i < nsteps; ++i) {
    int rc = launch_stencil(h, 0, h->dev.rows, (cudaStream_t)stream);
    if (rc) return rc;
    h->cur ^= 1;
  }
  return PF_OK;
}

extern "C" int pf_step_field_rows(pf_handle* h, int row_begin, int row_end, void* stream) {
  if (!h || row_begin < 0 || row_end > h->dev.rows) return PF_EINVAL;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  if (h->tile_pending && is_sharded(h) && row_begin == h->tile_ty0 * kTileRows &&
      row_end == (h->tile_ty1 * kTileRows < h->dev.rows ? h->tile_ty1 * kTileRows : h->dev.rows)) {
    // the rows whose next-tick deposits pf_agents_step(PF_STEP_HANDOVER) has prepared: they ride along
    int rc = launch_stencil(h, row_begin, row_end, (cudaStream_t)stream, true);
    if (rc) return rc;
    h->tile_pending = false;
    h->tile_handed = true;
    return PF_OK;
  }
  return launch_stencil(h, row_begin, row_end, (cudaStream_t)stream);
}

extern "C" int pf_swap(pf_handle* h) {
  if (!h) return PF_EINVAL;
  if (h->tile_pending)
    return set_err(h, PF_ESTATE, "pf_swap: deposits were handed over (PF_STEP_HANDOVER) but the rows of pf_tile_rows "
                                 "were not stepped as one pf_step_field_rows call");
  h->cur ^= 1;
  return PF_OK;
}

extern "C" int pf_set_option(pf_handle* h, int option, int value) {
  if (!h) return PF_EINVAL;
  if (option == PF_OPT_OVERLAP_AGENTS) { h->overlap_agents = value ? 1 : 0; return PF_OK; }
  if (option == PF_OPT_NO_FUSED_KEYS) { h->no_fused_keys = value ? 1 : 0; return PF_OK; }
  if (option == PF_OPT_NO_TILE_DEPOSIT) { h->no_tile_deposit = value ? 1 : 0; return PF_OK; }
  if (option == PF_OPT_NO_OWNER_DEPOSIT) { h->no_owner_deposit = value ? 1 : 0; return PF_OK; }
  if (option == PF_OPT_FORCE_OWNER_DEPOSIT) { h->force_owner_deposit = value ? 1 : 0; return PF_OK; }
  if (option == PF_OPT_STRIP_PART_A_TILES) { h->strip_part_a_tiles = value > 0 ? value : 0; return PF_OK; }
  if (option == PF_OPT_STRIP_ONE_LAUNCH) { h->strip_one_launch_opt = value < 0 ? 0 : (value > 2 ? 2 : value); return PF_OK; }
  if (option == PF_OPT_STRIP_MIG_INLINE) { h->strip_mig_inline = value ? 1 : 0; return PF_OK; }
  if (option == PF_OPT_STRIP_EDGES_SIDE) { h->strip_edges_side_opt = value ? 1 : 0; return PF_OK; }
  if (option == PF_OPT_STRIP_NO_PREDEPOSIT) { h->strip_no_predeposit = value ? 1 : 0; return PF_OK; }
  if (option == PF_OPT_PEER_TIMEOUT_MS) { h->peer_timeout_ns = (unsigned long long)(value > 0 ? value : 1) * 1000000ull; return PF_OK; }
  if (option == PF_OPT_USE_GRAPH) { h->use_graph = value ? 1 : 0; if (!value) graph_drop(h); return PF_OK; }
  return set_err(h, PF_EINVAL, "unknown option %d", option);
}

extern "C" int pf_set_stencil_variant(pf_handle* h, int variant, int rows_per_cta) {
  if (!h || variant < 0 || variant > 5 || rows_per_cta < 0) return PF_EINVAL;
  h->variant = variant;
  h->rows_per_cta = rows_per_cta;
  return PF_OK;
}

// ---------------------------------------------------------------------------------------------
// K2 launch
// ---------------------------------------------------------------------------------------------
static bool owner_ready(const pf_handle* h, const pf_agents* a);
static int check_agents(pf_handle* h, const pf_agents* a);

extern "C" int pf_tile_stats(pf_handle* h, uint64_t* attempted, uint64_t* fell_back) {
  if (!h) return PF_EINVAL;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  uint32_t fb = 0;
  if (h->tile_flag) {
    PF_CUDA(h, cudaDeviceSynchronize());
    PF_CUDA(h, cudaMemcpy(&fb, h->tile_flag + 1, sizeof(fb), cudaMemcpyDeviceToHost));
  }
  if (attempted) *attempted = h->tile_attempts;
  if (fell_back) *fell_back = fb;
  return PF_OK;
}

extern "C" int pf_tile_dense(pf_handle* h, uint64_t* n_dense, uint64_t* pool_slots) {
  if (!h) return PF_EINVAL;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  uint32_t n = 0;
  if (h->tile_pool) {
    PF_CUDA(h, cudaDeviceSynchronize());
    PF_CUDA(h, cudaMemcpy(&n, h->tile_pool, sizeof(n), cudaMemcpyDeviceToHost));
  }
  if (n_dense) *n_dense = n < (uint32_t)h->tile_pool_cap ? n : (uint32_t)h->tile_pool_cap;
  if (pool_slots) *pool_slots = (uint64_t)h->tile_pool_cap;
  return PF_OK;
}

// Can this tick's stencil apply the next tick's deposits (pf_deposit_tile.cuh)? Needs the sort-free
// deposit (its kernels are the device-side fallback), the TMA stencil in its default shape on the
// whole field, and no second stream.
static bool tile_common_ok(const pf_handle* h, const pf_agents* a) {
  const PFDev& d = h->dev;
  return !h->no_tile_deposit && !h->overlap_agents && owner_ready(h, a) && d.stencil != 0 && d.dep_red && d.dep_exact &&
         d.W % 4 == 0 && d.W >= 512 && (h->variant == 0 || h->variant == 2) && h->rows_per_cta <= 0 &&
         kTileRows == kDepRows && kTileEnt == kDepEnt && kTileCols == kTmaCols && kTileBlkBytes == kDepBlkBytes &&
         kTilePtrBytes == kDepPtrBytes;
}
static bool tile_ok(const pf_handle* h, const pf_agents* a) { return !is_sharded(h) && tile_common_ok(h, a); }

// Strips: the tile rows [ty0, ty1) whose next-tick deposits the stencil can take. Not the first
// tile rows (the driver steps them before the move kernel, to hide the halo exchange: an eighth of
// the strip, at least 64 rows) and not the last 64+ rows next to a neighbour: arrivals land in the
// edge bands (a robot moves well under 32 cells per tick — checked), so no handed-over cell is
// shared with an agent the regular deposit kernel handles.
static bool strip_tile_range(const pf_handle* h, const pf_agents* a, int* ty0, int* ty1) {
  *ty0 = *ty1 = 0;
  if (!is_sharded(h) || !tile_common_ok(h, a)) return false;
  const PFDev& d = h->dev;
  const double omega = fabs((double)d.vmax) + fabs((double)d.cruise) + fabs((double)d.bl) + fabs((double)d.br);
  const double step_cells = omega * h->cfg.wheel_radius * h->cfg.dt / h->cfg.cell_size;
  if (!(step_cells < 32.0)) return false;
  const int rows = h->dev.rows, ny = (rows + kTileRows - 1) / kTileRows;
  // The rows above the handed ones are stepped before the move kernel, while the halo is in flight; arrivals from
  // the strip above land in the first rows, so the first tile row never hands over. One tile row is enough to
  // cover the halo flag's flight (measured at 8 x 4096 rows: 0.773 ms per tick against 0.788 with an eighth of the
  // strip, round 1's choice; PF_OPT_STRIP_PART_A_TILES changes it). The arena's top strip has no neighbour above
  // and, in one-launch mode, no rows stepped early at all.
  int lo = 1;
  if (h->strip_part_a_tiles > 0) lo = h->strip_part_a_tiles < ny ? h->strip_part_a_tiles : ny;
  if (h->cfg.row0 == 0 && h->strip_one_launch) lo = 0;
  int hi = (h->cfg.row1 < h->cfg.H) ? (rows - kTileRows) / kTileRows : ny;
  if (hi <= lo) return false;
  *ty0 = lo;
  *ty1 = hi;
  return true;
}

extern "C" int pf_tile_rows(pf_handle* h, const pf_agents* a, int* row_begin, int* row_end) {
  int rc = check_agents(h, a);
  if (rc) return rc;
  int ty0, ty1;
  strip_tile_range(h, a, &ty0, &ty1);
  if (row_begin) *row_begin = ty0 * kTileRows;
  if (row_end) *row_end = ty1 * kTileRows < h->dev.rows ? ty1 * kTileRows : h->dev.rows;
  return PF_OK;
}

// before the move kernel: storage, cleared status flag, the bins the kernel fills
static int tile_begin(pf_handle* h, cudaStream_t s, TileBins* tb, int ty0, int ty1) {
  const int nx = (h->dev.W + kTileCols - 1) / kTileCols, ny = (h->dev.rows + kTileRows - 1) / kTileRows;
  const long long ntiles = (long long)nx * ny * h->dev.C;
  if (h->tile_alloc < ntiles) {
    tile_free(h);
    PF_CUDA(h, cudaMalloc(&h->tile_bcnt, sizeof(uint32_t) * ntiles * kTileGroups));
    PF_CUDA(h, cudaMemset(h->tile_bcnt, 0, sizeof(uint32_t) * ntiles * kTileGroups));
    PF_CUDA(h, cudaMalloc(&h->tile_flag, 2 * sizeof(uint32_t)));  // [0] status of this tick, [1] fallbacks so far
    PF_CUDA(h, cudaMemset(h->tile_flag, 0, 2 * sizeof(uint32_t)));
    PF_CUDA(h, cudaMalloc(&h->tile_brec, sizeof(uint32_t) * ntiles * kTileGroups * kTileGroupCap));
    PF_CUDA(h, cudaMalloc(&h->tile_count, sizeof(uint32_t) * ntiles));
    PF_CUDA(h, cudaMalloc(&h->tile_blk, (size_t)ntiles * kTileBlkBytes));
    // side buffers for crowded tiles: a pool of 256 KB blocks, at most 512 (128 MB) or one per tile
    h->tile_pool_cap = (int)(ntiles < 512 ? ntiles : 512);
    PF_CUDA(h, cudaMalloc(&h->tile_pool, sizeof(uint32_t)));
    PF_CUDA(h, cudaMemset(h->tile_pool, 0, sizeof(uint32_t)));
    PF_CUDA(h, cudaMalloc(&h->tile_dense, sizeof(int) * ntiles));
    PF_CUDA(h, cudaMemset(h->tile_dense, 0xff, sizeof(int) * ntiles));
    PF_CUDA(h, cudaMalloc(&h->tile_side, sizeof(float) * (size_t)h->tile_pool_cap * kTileCells));
    PF_CUDA(h, cudaMemset(h->tile_side, 0, sizeof(float) * (size_t)h->tile_pool_cap * kTileCells));
    h->tile_alloc = ntiles;
  }
  PF_CUDA(h, cudaMemsetAsync(h->tile_flag, 0, sizeof(uint32_t), s));
  tb->bcnt = h->tile_bcnt; tb->brec = h->tile_brec;
  tb->dense = h->tile_dense; tb->side = h->tile_side; tb->pool = h->tile_pool; tb->pool_cap = h->tile_pool_cap;
  tb->flag = h->tile_flag; tb->nx = nx; tb->ny = ny; tb->ty_lo = ty0; tb->ty_hi = ty1;
  h->tile_ty0 = ty0;
  h->tile_ty1 = ty1;
  return PF_OK;
}

// after the move kernel: bins -> per-tile lists for the stencil
static int tile_finish(pf_handle* h, cudaStream_t s) {
  const int nx = (h->dev.W + kTileCols - 1) / kTileCols, ny = (h->dev.rows + kTileRows - 1) / kTileRows;
  const long long ntiles = (long long)nx * ny * h->dev.C;
  k_tile_group<<<(unsigned)ntiles, kGroupThreads, 0, s>>>(h->tile_bcnt, h->tile_brec, h->tile_flag, h->tile_count,
                                                          h->tile_blk, h->dev.i_explore, h->dev.i_high, h->dev.i_homing,
                                                          h->tile_dense, h->tile_pool, h->tile_pool_cap);
  PF_LAUNCH_CHECK(h);
  return PF_OK;
}

// is the owner-computes deposit usable for these agent arrays? (splitters from the last pf_agents_sort)
static bool owner_ready(const pf_handle* h, const pf_agents* a) {
  return !h->no_owner_deposit && (h->own_worth || h->force_owner_deposit) && h->own_n == a->n &&
         h->own_x == a->x && a->n >= 4 * kOwnChunk && (a->n >= 64 * kOwnChunk || h->force_owner_deposit);
}

// keys_in / vals_in (array order) -> field, without a global sort
// only_if (device flag, may be null): do the work only if *only_if != 0
static int owner_sum(pf_handle* h, long long n, cudaStream_t s, uint32_t* only_if = nullptr) {
  h->mig_flags_n = 0;
  const uint32_t* keys_in = (const uint32_t*)h->scratch;
  const float* vals_in = (const float*)(keys_in + 2 * h->cap);
  const int nchunks = (int)((n + kOwnChunk - 1) / kOwnChunk);
  k_owner_bounds<<<1, 1024, 0, s>>>(nchunks, h->own_body, h->own_cmin, h->own_cmax, h->own_pmax, h->own_smin,
                                    h->own_tail, only_if);
  PF_LAUNCH_CHECK(h);
  // stand-by behind the fused deposit: one wave of CTAs striding over the key ranges
  const int grid = only_if ? (nchunks < 4 * h->n_sm ? nchunks : 4 * h->n_sm) : nchunks;
  k_deposit_owner<<<grid, kOwnThreads, 0, s>>>(h->dev, n, nchunks, h->own_body, keys_in, vals_in, h->sentinel,
                                               h->own_split, h->own_pmax, h->own_smin, h->own_tail,
                                               h->buf[h->cur], peer_rows(h, h->cur), only_if, h->own_handed_lo,
                                               h->own_handed_hi, h->own_handed_hi ? h->tile_flag : nullptr);
  h->own_handed_lo = h->own_handed_hi = 0;
  PF_LAUNCH_CHECK(h);
  return PF_OK;
}

static int sort_and_sum(pf_handle* h, long long n, cudaStream_t s) {
  h->mig_flags_n = 0;  // the sort reuses the scratch the migration flags live in
  uint32_t* keys_in = (uint32_t*)h->scratch;
  uint32_t* keys_out = keys_in + h->cap;
  float* vals_in = (float*)(keys_out + h->cap);
  float* vals_out = vals_in + h->cap;
  size_t bytes = h->cub_temp_bytes;
  PF_CUDA(h, cub::DeviceRadixSort::SortPairs(h->cub_temp, bytes, keys_in, keys_out, vals_in, vals_out,
                                             (int)n, 0, h->key_bits, s));
  k_deposit_segsum<<<blocks_for(n, 256), 256, 0, s>>>(h->dev, n, keys_out, vals_out, h->sentinel,
                                                     h->buf[h->cur], peer_rows(h, h->cur));
  PF_LAUNCH_CHECK(h);
  return PF_OK;
}

extern "C" int pf_deposit(pf_handle* h, int64_t n, const float* x, const float* y, const float* amount,
                          const uint8_t* channel, void* stream) {
  if (!h || n < 0 || (n > 0 && (!x || !y || !amount))) return PF_EINVAL;
  if (n == 0) return PF_OK;
  if (n > 0x7fffffffll) return set_err(h, PF_EINVAL, "more than 2^31-1 deposits per call");
  int rc = pf_reserve_agents(h, n);
  if (rc) return rc;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  uint32_t* keys_in = (uint32_t*)h->scratch;
  float* vals_in = (float*)(keys_in + 2 * h->cap);
  k_deposit_keys_generic<<<blocks_for(n, 256), 256, 0, s>>>(h->dev, n, x, y, amount, channel, keys_in,
                                                           vals_in, h->sentinel);
  PF_LAUNCH_CHECK(h);
  return sort_and_sum(h, n, s);
}

static int check_agents(pf_handle* h, const pf_agents* a) {
  if (!h || !a || a->n < 0) return PF_EINVAL;
  if (a->n > 0 && (!a->x || !a->y || !a->theta || !a->wl || !a->wr || !a->meta))
    return set_err(h, PF_EINVAL, "pf_agents has null arrays");
  if (a->n > 0x7fffffffll) return set_err(h, PF_EINVAL, "more than 2^31-1 agent slots per handle");
  return PF_OK;
}

// mode 0: the whole deposit. mode 1 (EARLY) / the call after it (LATE): pf_strip_tick's deposit in two parts — the
// robots that were not handed over at the end of the hand-over tick (on the side stream, beside the stencil launch
// that carries the handed ones), the handed ones at the start of the next tick, which is nothing but three
// launches that return at once unless the stencil could not apply them.
static int agents_deposit(pf_handle* h, const pf_agents* a, cudaStream_t s, int mode, int late_may_skip) {
  uint32_t* keys_in = (uint32_t*)h->scratch;
  float* vals_in = (float*)(keys_in + 2 * h->cap);
  const bool own = owner_ready(h, a);
  if (mode != 0 && (!own || !h->tile_handed))
    return set_err(h, PF_ESTATE, "two-part deposit without a hand-over tick behind it");
  TileBins handed;
  memset(&handed, 0, sizeof(handed));
  h->own_handed_lo = h->own_handed_hi = 0;
  if (h->tile_handed) {  // the last stencil took the interior tile rows' deposits (unless its flag says otherwise)
    handed.bcnt = h->tile_bcnt;
    handed.nx = (h->dev.W + kTileCols - 1) / kTileCols;
    handed.ny = (h->dev.rows + kTileRows - 1) / kTileRows;
    handed.ty_lo = h->tile_ty0;
    handed.ty_hi = h->tile_ty1;
    if (mode != 1) h->tile_handed = false;   // EARLY: the tick is still "one ahead" until the LATE part has run
    if (mode != 2) {  // (LATE deposits exactly the handed rows, if anything)
      const long long r1 = (long long)h->tile_ty1 * kTileRows < h->dev.rows ? (long long)h->tile_ty1 * kTileRows : h->dev.rows;
      h->own_handed_lo = (uint32_t)((long long)h->tile_ty0 * kTileRows * h->dev.W);
      h->own_handed_hi = (uint32_t)(r1 * h->dev.W);
    }
  }
  k_deposit_keys_agents<<<wave_grid(a->n, h->wave_keys), 256, 0, s>>>(h->dev, a->n, a->x, a->y, a->meta, keys_in,
                                                             vals_in, h->sentinel, h->counters,
                                                             own ? h->own_cmin : nullptr,
                                                             own ? h->own_cmax : nullptr, handed, h->tile_flag, mode,
                                                             late_may_skip);
  PF_LAUNCH_CHECK(h);
  if (mode == 2) return owner_sum(h, a->n, s, h->tile_flag);   // only if the stencil could not apply them
  return own ? owner_sum(h, a->n, s) : sort_and_sum(h, a->n, s);
}

extern "C" int pf_agents_deposit(pf_handle* h, const pf_agents* a, void* stream) {
  int rc = check_agents(h, a);
  if (rc) return rc;
  if (a->n == 0) return PF_OK;
  rc = pf_reserve_agents(h, a->n);
  if (rc) return rc;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  if (h->strip_predep) {  // the other robots were deposited at the end of the last pf_strip_tick
    h->strip_predep = false;
    return agents_deposit(h, a, (cudaStream_t)stream, 2, 0);
  }
  return agents_deposit(h, a, (cudaStream_t)stream, 0, 0);
}

// ---------------------------------------------------------------------------------------------
// K3 launch
// ---------------------------------------------------------------------------------------------
extern "C" int pf_sense(pf_handle* h, int64_t n, const float* x, const float* y, const float* theta,
                        const uint32_t* meta, float* out5, void* stream) {
  if (!h || n < 0 || (n > 0 && (!x || !y || !out5))) return PF_EINVAL;
  if (h->cfg.sense_mode == PF_SENSE_HEADING && n > 0 && !theta)
    return set_err(h, PF_EINVAL, "heading-relative sensing needs theta");
  if (n == 0) return PF_OK;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  k_sense<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(h->dev, h->buf[h->cur], n, x, y, theta, meta, out5);
  PF_LAUNCH_CHECK(h);
  return PF_OK;
}

static int launch_agents_step(pf_handle* h, const pf_agents* a, uint32_t flags, uint8_t* dec, cudaStream_t s,
                              bool emit_next_keys = false, const TileBins* bins = nullptr) {
  // on a strip, the move kernel also counts the leavers per chunk for pf_migrate_pack
  uint32_t* mig = nullptr;
  if (is_sharded(h) && (flags & PF_STEP_MOVE) && h->mig_cnt && a->n <= h->cap) {
    mig = h->mig_cnt;
    const long long nc = (a->n + kOwnChunk - 1) / kOwnChunk;
    PF_CUDA(h, cudaMemsetAsync(mig, 0, sizeof(uint32_t) * 2 * nc, s));
    h->mig_flags_n = a->n;
    h->mig_flags_x = a->x;
  }
  uint32_t* nk = nullptr;
  float* nv = nullptr;
  TileBins no_bins;
  memset(&no_bins, 0, sizeof(no_bins));
  if (emit_next_keys) {  // keys_in / vals_in of the sort scratch, for the next tick's deposit
    nk = (uint32_t*)h->scratch;
    nv = (float*)(nk + 2 * h->cap);
  }
  k_agents_step<<<wave_grid(a->n, h->wave_agents), 256, 0, s>>>(h->dev, h->buf[h->cur], flags, a->n, a->x, a->y,
                                                     a->theta, a->wl, a->wr, a->meta, dec, h->counters, mig,
                                                     nk, nv, h->sentinel,
                                                     (nk && owner_ready(h, a)) ? h->own_cmin : nullptr,
                                                     (nk && owner_ready(h, a)) ? h->own_cmax : nullptr,
                                                     bins ? *bins : no_bins);
  PF_LAUNCH_CHECK(h);
  return PF_OK;
}

extern "C" int pf_agents_step(pf_handle* h, const pf_agents* a, uint32_t flags, uint8_t* decisions_out,
                              void* stream) {
  int rc = check_agents(h, a);
  if (rc) return rc;
  if (a->n == 0) return PF_OK;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  int ty0, ty1;
  if ((flags & PF_STEP_HANDOVER) && (flags & PF_STEP_MOVE) && strip_tile_range(h, a, &ty0, &ty1)) {
    // strips: the agents that end up in the interior tile rows hand their next deposit to the stencil
    rc = pf_reserve_agents(h, a->n);
    if (rc) return rc;
    TileBins bins;
    rc = tile_begin(h, s, &bins, ty0, ty1);
    if (rc) return rc;
    rc = launch_agents_step(h, a, flags & ~PF_STEP_HANDOVER, decisions_out, s, false, &bins);
    if (rc) return rc;
    rc = tile_finish(h, s);
    if (rc) return rc;
    h->tile_pending = true;
    h->tile_attempts++;
    return PF_OK;
  }
  return launch_agents_step(h, a, flags & ~PF_STEP_HANDOVER, decisions_out, s);
}

extern "C" int pf_set_time(pf_handle* h, double time_s, double start_delay_s) {
  if (!h) return PF_EINVAL;
  h->t0 = time_s;
  h->ticks = 0;
  h->start_delay = start_delay_s;
  return PF_OK;
}

extern "C" int pf_get_time(const pf_handle* h, double* time_s, uint64_t* ticks) {
  if (!h) return PF_EINVAL;
  if (time_s) *time_s = h->t0 + (double)h->ticks * (double)h->cfg.dt;
  if (ticks) *ticks = h->ticks;
  return PF_OK;
}

// fold the recorded event pairs into the running sums (synchronises the device)
static int prof_fold(pf_handle* h) {
  if (!h->prof_used) return PF_OK;
  PF_CUDA(h, cudaDeviceSynchronize());
  for (int i = 0; i < h->prof_used; ++i) {
    cudaEvent_t* pe = h->prof_ev + 6 * i;
    float ms = 0.f;
    PF_CUDA(h, cudaEventElapsedTime(&ms, pe[0], pe[1]));
    h->prof.deposit_ms += ms;
    PF_CUDA(h, cudaEventElapsedTime(&ms, pe[5], pe[2]));
    h->prof.stencil_ms += ms;
    if (h->prof_has_agents[i]) {
      PF_CUDA(h, cudaEventElapsedTime(&ms, pe[3], pe[4]));
      h->prof.agents_ms += ms;
    }
    h->prof.ticks++;
  }
  h->prof_used = 0;
  return PF_OK;
}

extern "C" int pf_profile_enable(pf_handle* h, int on) {
  if (!h) return PF_EINVAL;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  if (on && !h->prof_ev) {
    h->prof_ev = new (std::nothrow) cudaEvent_t[kProfSlots * 6];
    if (!h->prof_ev) return set_err(h, PF_ENOMEM, "host allocation failed");
    for (int i = 0; i < kProfSlots * 6; ++i) PF_CUDA(h, cudaEventCreate(&h->prof_ev[i]));
  }
  if (!on) { int rc = prof_fold(h); if (rc) return rc; rc = strip_fold(h); if (rc) return rc; }
  h->prof_on = on ? 1 : 0;
  return PF_OK;
}

extern "C" int pf_profile_read(pf_handle* h, pf_profile* out, int reset) {
  if (!h || !out) return PF_EINVAL;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  int rc = prof_fold(h);
  if (rc) return rc;
  *out = h->prof;
  if (reset) memset(&h->prof, 0, sizeof(h->prof));
  return PF_OK;
}

extern "C" int pf_step_ex(pf_handle* h, const pf_agents* a, int nsteps, uint32_t mask, void* stream);
extern "C" int pf_migrate_unpack2(pf_handle* h, const pf_agents* a, const uint32_t* buf, const uint32_t* buf2,
                                  int64_t cap, void* stream);

extern "C" int pf_agents_observe(pf_handle* h, const pf_agents* a, const float* x_new, const float* y_new,
                                 const float* theta_new, void* stream) {
  int rc = check_agents(h, a);
  if (rc) return rc;
  if (a->n == 0) return PF_OK;
  if (!x_new || !y_new) return PF_EINVAL;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  k_agents_observe<<<blocks_for(a->n, 256), 256, 0, (cudaStream_t)stream>>>(h->dev, a->n, a->x, a->y, a->theta,
                                                                          a->meta, x_new, y_new, theta_new);
  PF_LAUNCH_CHECK(h);
  return PF_OK;
}

extern "C" int pf_step(pf_handle* h, const pf_agents* a, int nsteps, void* stream) {
  return pf_step_ex(h, a, nsteps, PF_STEP_ALL, stream);
}

// One tick enqueued on `s`. keys_ready: the previous tick's move kernel already produced this tick's
// deposit keys; emit_next: produce the next tick's keys in this tick's move kernel.
static int tick_once(pf_handle* h, const pf_agents* a, uint32_t mask, cudaStream_t s, bool keys_ready,
                     bool emit_next, bool* keys_ready_out) {
  int rc;
  const double now = h->t0 + (double)h->ticks * (double)h->cfg.dt;
  const bool sense = (now >= h->start_delay) && (mask & PF_STEP_SENSE);
  const uint32_t flags = (mask & PF_STEP_MOVE) | (sense ? PF_STEP_SENSE : 0u);
  cudaEvent_t* pe = nullptr;
  if (h->prof_on) {
    if (h->prof_used == kProfSlots) { rc = prof_fold(h); if (rc) return rc; }
    pe = h->prof_ev + 6 * h->prof_used;
    h->prof_has_agents[h->prof_used] = a->n > 0 && (mask & (PF_STEP_SENSE | PF_STEP_MOVE));
    PF_CUDA(h, cudaEventRecord(pe[0], s));
  }
  if (now >= h->start_delay && (mask & PF_STEP_DEPOSIT) && a->n > 0) {  // deposit_pheromone, sup:577
    if (keys_ready && h->tile_pending)  // the last stencil was asked to apply them: only_if it could not
      rc = owner_sum(h, a->n, s, h->tile_flag);
    else if (keys_ready)  // keys came from the previous tick's move kernel
      rc = owner_ready(h, a) ? owner_sum(h, a->n, s) : sort_and_sum(h, a->n, s);
    else rc = pf_agents_deposit(h, a, (void*)s);
    if (rc) return rc;
  }
  h->tile_pending = false;
  *keys_ready_out = false;
  if (pe) PF_CUDA(h, cudaEventRecord(pe[1], s));
  const bool run_agents = a->n > 0 && (mask & (PF_STEP_SENSE | PF_STEP_MOVE));
  // Measured on B200 (profiles/r01a_*): overlapped, the agent kernel is starved behind the
  // stencil's CTAs and the pair takes 4.55 ms against 0.63 + 2.65 ms back to back, so the
  // default is back to back on the caller's stream; PF_OPT_OVERLAP_AGENTS restores the fork.
  const bool fork = run_agents && h->overlap_agents;
  cudaStream_t as = fork ? h->side : s;
  if (run_agents) {
    if (fork) {
      PF_CUDA(h, cudaEventRecord(h->ev_fork, s));
      PF_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    }
    if (pe) PF_CUDA(h, cudaEventRecord(pe[3], as));
    // fuse the next tick's key generation into this tick's move kernel when that deposit is certain
    const double next = h->t0 + (double)(h->ticks + 1) * (double)h->cfg.dt;
    const bool emit = emit_next && !h->no_fused_keys && (mask & PF_STEP_DEPOSIT) && next >= h->start_delay &&
                      !is_sharded(h);
    // hand the next tick's deposits to this tick's stencil? (pf_deposit_tile.cuh)
    const bool tile_try = emit && (mask & PF_STEP_FIELD) && tile_ok(h, a);
    TileBins bins;
    if (tile_try) {
      rc = tile_begin(h, as, &bins, 0, (h->dev.rows + kTileRows - 1) / kTileRows);
      if (rc) return rc;
    }
    rc = launch_agents_step(h, a, flags, nullptr, as, emit, tile_try ? &bins : nullptr);
    if (rc) return rc;
    *keys_ready_out = emit;
    if (tile_try) {
      rc = tile_finish(h, as);
      if (rc) return rc;
      h->tile_pending = true;
      h->tile_attempts++;
    }
    if (pe) PF_CUDA(h, cudaEventRecord(pe[4], as));
    if (fork) PF_CUDA(h, cudaEventRecord(h->ev_join, h->side));
  }
  if (pe) PF_CUDA(h, cudaEventRecord(pe[5], s));
  if (mask & PF_STEP_FIELD) {
    rc = launch_stencil(h, 0, h->dev.rows, s, h->tile_pending);
    if (rc) return rc;
  }
  if (pe) {
    PF_CUDA(h, cudaEventRecord(pe[2], s));
    h->prof_used++;
  }
  if (fork) PF_CUDA(h, cudaStreamWaitEvent(s, h->ev_join, 0));
  k_tick_counter<<<1, 1, 0, s>>>(h->counters);
  PF_LAUNCH_CHECK(h);
  if (mask & PF_STEP_FIELD) h->cur ^= 1;
  h->ticks++;
  return PF_OK;
}

static void graph_drop(pf_handle* h) {
  if (h->g_exec) cudaGraphExecDestroy(h->g_exec);
  if (h->g_graph) cudaGraphDestroy(h->g_graph);
  h->g_exec = nullptr;
  h->g_graph = nullptr;
}

// Capture two steady-state ticks (the double buffer returns to its parity) into a CUDA graph.
static int graph_prepare(pf_handle* h, const pf_agents* a, uint32_t mask, cudaStream_t s) {
  GraphKey k;
  memset(&k, 0, sizeof(k));
  k.x = a->x; k.y = a->y; k.theta = a->theta; k.wl = a->wl; k.wr = a->wr; k.meta = a->meta;
  k.n = a->n; k.mask = mask; k.parity = h->cur; k.variant = h->variant; k.rpc = h->rows_per_cta;
  k.tile = (h->tile_pending ? 1 : 0) | (tile_ok(h, a) ? 2 : 0) | (owner_ready(h, a) ? 4 : 0);
  k.scratch = h->scratch; k.cap = h->cap;
  if (h->g_exec && memcmp(&k, &h->g_key, sizeof(k)) == 0) return PF_OK;
  graph_drop(h);
  const uint64_t ticks0 = h->ticks, launches0 = h->launches;
  const unsigned long long attempts0 = h->tile_attempts;
  // capture on the handle's own stream (the caller's may be the legacy default stream, which cannot
  // be captured); the instantiated graph is then launched into the caller's stream
  (void)s;
  cudaStream_t cs = h->side;
  PF_CUDA(h, cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
  bool kr = true;
  int rc = tick_once(h, a, mask, cs, true, true, &kr);
  if (rc == PF_OK) rc = tick_once(h, a, mask, cs, kr, true, &kr);
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(cs, &g);
  h->g_nodes = h->launches - launches0;
  h->g_tile_attempts = h->tile_attempts - attempts0;
  h->tile_attempts = attempts0;
  h->ticks = ticks0;       // capturing enqueues nothing: undo the host-side bookkeeping
  h->launches = launches0;  // (two buffer flips cancel)
  if (rc != PF_OK || e != cudaSuccess || !g || !kr) {
    if (g) cudaGraphDestroy(g);
    if (rc == PF_OK) rc = set_err(h, PF_ECUDA, "graph capture failed: %s", cudaGetErrorString(e));
    return rc;
  }
  h->g_graph = g;
  PF_CUDA(h, cudaGraphInstantiate(&h->g_exec, g, 0));
  h->g_key = k;
  return PF_OK;
}

extern "C" int pf_step_ex(pf_handle* h, const pf_agents* a, int nsteps, uint32_t mask, void* stream) {
  int rc = check_agents(h, a);
  if (rc) return rc;
  if (nsteps < 0) return PF_EINVAL;
  if (is_sharded(h)) return set_err(h, PF_ESTATE, "pf_step on a sharded handle: drive the phases from the host");
  rc = pf_reserve_agents(h, a->n);
  if (rc) return rc;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  bool keys_ready = false;
  h->tile_pending = false;
  int t = 0;
  // CUDA-graph path for launch-bound (small) problems: tick 0 normally (builds keys), then pairs of
  // steady-state ticks replayed from one graph, the last tick normally (must not pre-bump the
  // deposit counters for a tick that this call does not run).
  const double now0 = h->t0 + (double)h->ticks * (double)h->cfg.dt;
  const bool graph_ok = h->use_graph && !h->prof_on && !h->overlap_agents && !h->no_fused_keys && a->n > 0 &&
                        nsteps >= 6 && mask == PF_STEP_ALL && now0 >= h->start_delay;
  if (graph_ok) {
    rc = tick_once(h, a, mask, s, false, true, &keys_ready);
    if (rc) return rc;
    t = 1;
    if (keys_ready) {
      const int pairs = (nsteps - 2) / 2;
      rc = graph_prepare(h, a, mask, s);
      if (rc == PF_OK) {
        for (int p = 0; p < pairs; ++p) {
          PF_CUDA(h, cudaGraphLaunch(h->g_exec, s));
          h->ticks += 2;
          h->launches += h->g_nodes;
          h->tile_attempts += h->g_tile_attempts;
        }
        t += 2 * pairs;
      } else {
        h->use_graph = 0;  // fall back to plain launches for good; the error text stays readable
        cudaGetLastError();  // clear the pending error so the plain path starts clean
      }
    }
  }
  for (; t < nsteps; ++t) {
    rc = tick_once(h, a, mask, s, keys_ready, t + 1 < nsteps, &keys_ready);
    if (rc) return rc;
  }
  return PF_OK;
}

extern "C" int pf_evaporate_trail(pf_handle* h, int64_t n, const double* elapsed, const double* intensity,
                                  double rate, double* out, void* stream) {
  if (!h || n < 0 || (n > 0 && (!elapsed || !intensity || !out))) return PF_EINVAL;
  if (n == 0) return PF_OK;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  k_evaporate_trail<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, elapsed, intensity, rate, out);
  PF_LAUNCH_CHECK(h);
  return PF_OK;
}

// ---------------------------------------------------------------------------------------------
// spatial sort of the agent arrays
// ---------------------------------------------------------------------------------------------
extern "C" int pf_agents_sort(pf_handle* h, const pf_agents* a, void* stream) {
  int rc = check_agents(h, a);
  if (rc) return rc;
  if (a->n <= 1) return PF_OK;
  if (h->tile_handed || h->tile_pending)
    return set_err(h, PF_ESTATE, "pf_agents_sort between a hand-over tick (PF_STEP_HANDOVER) and the next deposit: "
                                 "shared cells were summed in the old array order; do not hand over on the tick "
                                 "before a re-sort");
  rc = pf_reserve_agents(h, a->n);
  if (rc) return rc;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = a->n;
  h->mig_flags_n = 0;
  if (h->tile_dense) {  // crowded-tile marks are re-learnt after every re-sort (their side buffers are empty here:
                        // every stencil launch that follows a binning move kernel drains them)
    PF_CUDA(h, cudaMemsetAsync(h->tile_dense, 0xff, sizeof(int) * h->tile_alloc, s));
    PF_CUDA(h, cudaMemsetAsync(h->tile_pool, 0, sizeof(uint32_t), s));
  }
  uint32_t* keys_in = (uint32_t*)h->scratch;
  uint32_t* keys_out = keys_in + h->cap;
  uint32_t* idx_in = keys_out + h->cap;
  uint32_t* perm = idx_in + h->cap;
  uint32_t* tmp = perm + h->cap;
  const unsigned long long cells = (unsigned long long)h->cfg.H * h->cfg.W;
  if (cells >= 0xffffffffull) return set_err(h, PF_EINVAL, "pf_agents_sort: more than 2^32-1 cells");
  const uint32_t sentinel = (uint32_t)cells;
  k_sort_keys<<<blocks_for(n, 256), 256, 0, s>>>(h->dev, n, a->x, a->y, a->meta, keys_in, idx_in, sentinel);
  PF_LAUNCH_CHECK(h);
  size_t bytes = h->cub_temp_bytes;
  PF_CUDA(h, cub::DeviceRadixSort::SortPairs(h->cub_temp, bytes, keys_in, keys_out, idx_in, perm, (int)n, 0,
                                             ceil_log2_u64(cells + 1), s));
  if (n >= 4 * kOwnChunk) {  // splitters + fresh chunk bounds for the owner-computes deposit
    const int nb = (int)((n + kOwnChunk - 1) / kOwnChunk);
    const uint32_t plane = (uint32_t)((long long)h->dev.rows * h->dev.W);
    PF_CUDA(h, cudaMemsetAsync(h->own_body, 0, 4 * sizeof(uint32_t), s));
    k_owner_splitters<<<blocks_for(nb + 1, 256), 256, 0, s>>>(keys_out, n, nb, (uint32_t)((long long)h->cfg.row0 * h->cfg.W),
                                                             plane, sentinel, h->own_split, h->own_body);
    PF_LAUNCH_CHECK(h);
    PF_CUDA(h, cudaMemsetAsync(h->own_cmin, 0xff, sizeof(uint32_t) * nb, s));
    PF_CUDA(h, cudaMemsetAsync(h->own_cmax, 0, sizeof(uint32_t) * nb, s));
    h->own_n = n;
    h->own_x = a->x;
    {  // Is the sort-free deposit worth it here? A CTA re-scans about 1 + 2*drift_rows*agents_per_row/1024
       // array chunks; in dense clusters (thousands of agents per populated row) the global radix sort
       // is cheaper. One small synchronous read-back per pf_agents_sort.
      uint32_t st[4] = {0, 0, 0, 0};
      PF_CUDA(h, cudaMemcpyAsync(st, h->own_body, sizeof(st), cudaMemcpyDeviceToHost, s));
      PF_CUDA(h, cudaStreamSynchronize(s));
      const double alive = (double)st[0] * kOwnChunk;
      const double rows_pop = (double)(st[3] - st[2]) / (double)h->cfg.W + 1.0;
      const double per_row = alive > 0 ? alive / rows_pop : 0.0;
      const double drift_rows = 2.7;  // ~0.33 cells per tick over half of a 16-tick re-sort interval
      h->own_scan_estimate = 1.0 + 2.0 * drift_rows * per_row / kOwnChunk;
      h->own_worth = (st[0] > 0 && h->own_scan_estimate <= 8.0) ? 1 : 0;
    }
  }
  uint32_t* arrays[7] = {(uint32_t*)a->x, (uint32_t*)a->y, (uint32_t*)a->theta, (uint32_t*)a->wl,
                         (uint32_t*)a->wr, a->meta, a->gid};
  for (int k = 0; k < 7; ++k) {
    if (!arrays[k]) continue;
    k_permute_u32<<<blocks_for(n, 256), 256, 0, s>>>(n, perm, arrays[k], tmp);
    PF_LAUNCH_CHECK(h);
    PF_CUDA(h, cudaMemcpyAsync(arrays[k], tmp, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, s));
  }
  return PF_OK;
}

// ---------------------------------------------------------------------------------------------
// migration
// ---------------------------------------------------------------------------------------------
extern "C" int pf_migrate_pack(pf_handle* h, const pf_agents* a, uint32_t* up_buf, uint32_t* down_buf,
                               int64_t cap, void* stream) {
  int rc = check_agents(h, a);
  if (rc) return rc;
  if (!up_buf || !down_buf || cap < 0) return PF_EINVAL;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (a->n == 0) {
    PF_CUDA(h, cudaMemsetAsync(up_buf, 0, sizeof(uint32_t) * kMigWords, s));
    PF_CUDA(h, cudaMemsetAsync(down_buf, 0, sizeof(uint32_t) * kMigWords, s));
    return PF_OK;
  }
  rc = pf_reserve_agents(h, a->n);
  if (rc) return rc;
  const int nc = (int)((a->n + kOwnChunk - 1) / kOwnChunk);
  if (h->mig_flags_n != a->n || h->mig_flags_x != a->x) {  // the agents were moved by somebody else: count now
    PF_CUDA(h, cudaMemsetAsync(h->mig_cnt, 0, sizeof(uint32_t) * 2 * nc, s));
    k_migrate_count<<<blocks_for(a->n, 256), 256, 0, s>>>(h->dev, a->n, a->x, a->y, a->meta, h->mig_cnt);
    PF_LAUNCH_CHECK(h);
  }
  h->mig_flags_n = 0;
  k_migrate_offsets<<<1, 1024, 0, s>>>(nc, cap, h->mig_cnt, h->mig_base, up_buf, down_buf, h->counters);
  PF_LAUNCH_CHECK(h);
  k_migrate_scatter<<<nc, 256, 0, s>>>(h->dev, a->n, cap, h->mig_cnt, h->mig_base, a->x, a->y, a->theta, a->wl,
                                       a->wr, a->meta, a->gid, up_buf, down_buf);
  PF_LAUNCH_CHECK(h);
  return PF_OK;
}

extern "C" int pf_migrate_unpack(pf_handle* h, const pf_agents* a, const uint32_t* buf, int64_t cap,
                                 void* stream) {
  return pf_migrate_unpack2(h, a, buf, nullptr, cap, stream);
}

extern "C" int pf_migrate_unpack2(pf_handle* h, const pf_agents* a, const uint32_t* buf, const uint32_t* buf2,
                                  int64_t cap, void* stream) {
  int rc = check_agents(h, a);
  if (rc) return rc;
  if ((!buf && !buf2) || cap < 0) return PF_EINVAL;
  h->mig_flags_n = 0;
  if (a->n == 0) return PF_OK;
  rc = pf_reserve_agents(h, a->n);
  if (rc) return rc;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  // the body/tail split of the last pf_agents_sort of these arrays steers where arrivals are put
  const uint32_t* body = (h->own_n == a->n && h->own_x == a->x) ? h->own_body : nullptr;
  const int nc = (int)((a->n + kOwnChunk - 1) / kOwnChunk);
  if (h->mig_slot_cap < 2 * cap) {
    if (h->mig_slot_of) cudaFree(h->mig_slot_of);
    h->mig_slot_of = nullptr;
    PF_CUDA(h, cudaMalloc(&h->mig_slot_of, sizeof(long long) * (size_t)(2 * cap > 0 ? 2 * cap : 1)));
    h->mig_slot_cap = 2 * cap;
  }
  k_empty_count<<<nc, 256, 0, s>>>(a->n, a->meta, body, h->emp_cnt);
  PF_LAUNCH_CHECK(h);
  k_empty_offsets<<<1, 1024, 0, s>>>(nc, cap, h->emp_cnt, h->emp_base, buf, buf2, h->counters);
  PF_LAUNCH_CHECK(h);
  // one warp per record that may arrive (the headers stay on the device: the grid covers 2 * cap)
  {
    const unsigned want = blocks_for(2 * cap * 32, 256), wave = (unsigned)(2 * h->n_sm);
  k_migrate_find<<<want < wave ? want : wave, 256, 0, s>>>(a->n, nc, h->emp_base, buf, buf2, cap, a->meta, body,
                                                               h->mig_slot_of);
  }
  PF_LAUNCH_CHECK(h);
  k_migrate_fill<<<blocks_for(2 * cap, 256), 256, 0, s>>>(buf, buf2, cap, h->mig_slot_of, a->x, a->y, a->theta,
                                                          a->wl, a->wr, a->meta, a->gid);
  PF_LAUNCH_CHECK(h);
  return PF_OK;
}

// ---------------------------------------------------------------------------------------------
// peer-store halo (NVLink P2P)
// ---------------------------------------------------------------------------------------------
static uint32_t* mig_box_of(unsigned int* flags_base, int side, int C, long long pitch, long long cap) {
  return (uint32_t*)((char*)flags_base + mig_offset(C, pitch)) + (size_t)side * mig_box_words(cap);
}

extern "C" int pf_peer_export(pf_handle* h, pf_peer_info* out) {
  if (!h || !out) return PF_EINVAL;
  out->buf0 = mailbox_of(h->peer_flags, 0, h->cfg.C, h->dev.pitch);
  out->buf1 = mailbox_of(h->peer_flags, 1, h->cfg.C, h->dev.pitch);
  out->flags = h->peer_flags;
  out->rows = h->dev.rows; out->pitch = h->dev.pitch; out->chan_stride = h->dev.chan_stride;
  out->mig = mig_box_of(h->peer_flags, 0, h->cfg.C, h->dev.pitch, h->mig_cap);
  out->mig_cap = h->mig_cap;
  return PF_OK;
}

extern "C" int pf_peer_ipc_handles(pf_handle* h, uint8_t* out192) {
  if (!h || !out192) return PF_EINVAL;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memset(out192, 0, 192);
  cudaIpcMemHandle_t mh;  // one small allocation: flag words + mailboxes (the field itself is never exported)
  PF_CUDA(h, cudaIpcGetMemHandle(&mh, h->peer_flags));
  memcpy(out192, &mh, 64);
  return PF_OK;
}

// map a neighbour process's buffers into this process; `dims` = that neighbour's pf_peer_export
extern "C" int pf_peer_ipc_open(pf_handle* h, const uint8_t* in192, const pf_peer_info* dims, pf_peer_info* out) {
  if (!h || !in192 || !dims || !out) return PF_EINVAL;
  if (h->n_ipc >= 4) return set_err(h, PF_ESTATE, "pf_peer_ipc_open: more than 4 mappings on one handle");
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  void* base = nullptr;
  cudaIpcMemHandle_t mh;
  memcpy(&mh, in192, 64);
  PF_CUDA(h, cudaIpcOpenMemHandle(&base, mh, cudaIpcMemLazyEnablePeerAccess));
  h->ipc_opened[h->n_ipc++] = base;
  *out = *dims;
  out->flags = base;
  out->buf0 = mailbox_of((unsigned int*)base, 0, h->cfg.C, dims->pitch);
  out->buf1 = mailbox_of((unsigned int*)base, 1, h->cfg.C, dims->pitch);
  out->mig = dims->mig_cap > 0 ? mig_box_of((unsigned int*)base, 0, h->cfg.C, dims->pitch, dims->mig_cap) : nullptr;
  return PF_OK;
}

extern "C" int pf_peer_attach(pf_handle* h, const pf_peer_info* up, const pf_peer_info* down) {
  if (!h) return PF_EINVAL;
  if ((up && h->cfg.row0 == 0) || (down && h->cfg.row1 == h->cfg.H))
    return set_err(h, PF_EINVAL, "pf_peer_attach: no neighbour on that side of this strip");
  auto set = [](pf_handle::PeerLink& l, const pf_peer_info* i) {
    memset(&l, 0, sizeof(l));
    if (!i) return;
    l.buf[0] = (float*)i->buf0; l.buf[1] = (float*)i->buf1; l.flags = (unsigned int*)i->flags;
    l.rows = i->rows; l.pitch = i->pitch; l.chan_stride = i->chan_stride; l.set = 1;
    l.mig = (uint32_t*)i->mig; l.mig_cap = i->mig_cap;
  };
  set(h->peer_up, up);
  set(h->peer_dn, down);
  graph_drop(h);
  return PF_OK;
}

extern "C" int pf_peer_push_edges(pf_handle* h, void* stream) {
  if (!h) return PF_EINVAL;
  if (!h->peer_up.set && !h->peer_dn.set) return PF_OK;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  dim3 grid(blocks_for(h->cfg.W, 256), h->cfg.C);
  k_peer_push_edges<<<grid, 256, 0, (cudaStream_t)stream>>>(h->dev, h->buf[h->cur], peer_rows(h, h->cur));
  PF_LAUNCH_CHECK(h);
  return PF_OK;
}

extern "C" int pf_peer_signal(pf_handle* h, void* stream) {
  if (!h) return PF_EINVAL;
  h->peer_epoch += 1;
  if (!h->peer_up.set && !h->peer_dn.set) return PF_OK;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  k_peer_signal<<<1, 1, 0, (cudaStream_t)stream>>>(h->peer_up.set ? h->peer_up.flags + 1 : nullptr,
                                                  h->peer_dn.set ? h->peer_dn.flags + 0 : nullptr, h->peer_epoch);
  PF_LAUNCH_CHECK(h);
  return PF_OK;
}

extern "C" int pf_peer_wait(pf_handle* h, void* stream) {
  if (!h) return PF_EINVAL;
  if (!h->peer_up.set && !h->peer_dn.set) return PF_OK;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  // every CTA polls this strip's own two epoch words (bounded), then copies its columns of the
  // mailbox rows of the current buffer into the ghost rows
  dim3 grid(blocks_for(h->cfg.W, 256), h->cfg.C);
  k_peer_wait_ghosts<<<grid, 256, 0, (cudaStream_t)stream>>>(h->dev, h->buf[h->cur],
                                                           mailbox_of(h->peer_flags, h->cur, h->cfg.C, h->dev.pitch),
                                                           h->peer_up.set ? h->peer_flags + 0 : nullptr,
                                                           h->peer_dn.set ? h->peer_flags + 1 : nullptr,
                                                           h->peer_epoch, h->peer_timeout_ns, h->counters);
  PF_LAUNCH_CHECK(h);
  return PF_OK;
}

extern "C" int pf_peer_status(pf_handle* h, void* stream) {
  if (!h) return PF_EINVAL;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  unsigned long long t = 0;
  PF_CUDA(h, cudaMemcpyAsync(&t, h->counters + 13, sizeof(t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  PF_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
  if (t)
    return set_err(h, PF_ESTATE, "%llu peer wait(s) timed out: a neighbour strip did not signal (dead rank, or the "
                                 "strips are not ticking in lockstep); field and agents are not valid", t);
  return PF_OK;
}

// ---------------------------------------------------------------------------------------------
// one strip tick by one call (no host round trips between the phases)
// ---------------------------------------------------------------------------------------------
// Leavers are packed straight into the neighbours' migration mailboxes (peer stores), followed by an
// epoch flag — the halo's protocol once more. One mailbox per direction is enough: the neighbour's
// next pack comes behind its halo wait of the next tick, which needs this strip's halo signal, which
// this strip enqueues behind its unpack (same stream).
static int strip_pack(pf_handle* h, const pf_agents* a, cudaStream_t s) {
  uint32_t* up = h->peer_up.set && h->peer_up.mig ? h->peer_up.mig + mig_box_words(h->peer_up.mig_cap) : h->mig_dummy;
  uint32_t* dn = h->peer_dn.set && h->peer_dn.mig ? h->peer_dn.mig : h->mig_dummy;
  long long cap = h->mig_cap;
  if (h->peer_up.set && h->peer_up.mig_cap < cap) cap = h->peer_up.mig_cap;
  if (h->peer_dn.set && h->peer_dn.mig_cap < cap) cap = h->peer_dn.mig_cap;
  int rc = pf_migrate_pack(h, a, up, dn, cap, (void*)s);
  if (rc) return rc;
  h->mig_epoch += 1;
  if (h->peer_up.set || h->peer_dn.set) {
    k_peer_signal<<<1, 1, 0, s>>>(h->peer_up.set ? h->peer_up.flags + 3 : nullptr,
                                  h->peer_dn.set ? h->peer_dn.flags + 2 : nullptr, h->mig_epoch);
    PF_LAUNCH_CHECK(h);
  }
  return PF_OK;
}

static int strip_unpack(pf_handle* h, const pf_agents* a, cudaStream_t s) {
  if (!h->peer_up.set && !h->peer_dn.set) return PF_OK;
  uint32_t* from_up = h->peer_up.set ? mig_box_of(h->peer_flags, 0, h->cfg.C, h->dev.pitch, h->mig_cap) : nullptr;
  uint32_t* from_dn = h->peer_dn.set ? mig_box_of(h->peer_flags, 1, h->cfg.C, h->dev.pitch, h->mig_cap) : nullptr;
  return pf_migrate_unpack2(h, a, from_up, from_dn, h->mig_cap, (void*)s);
}

// fold the recorded marks of pf_strip_tick into per-phase sums (synchronises the device)
static int strip_fold(pf_handle* h) {
  if (!h->strip_used) return PF_OK;
  PF_CUDA(h, cudaDeviceSynchronize());
  for (int i = 0; i < h->strip_used; ++i) {
    cudaEvent_t* e = h->strip_ev + (size_t)i * kStripMarks;
    for (int k = 0; k + 1 < kStripMarks; ++k) {
      float ms = 0.f;
      PF_CUDA(h, cudaEventElapsedTime(&ms, e[k], e[k + 1]));
      h->strip_ms[k] += ms;
    }
    h->strip_ticks++;
  }
  h->strip_used = 0;
  return PF_OK;
}

extern "C" int pf_strip_profile(pf_handle* h, double* ms_out11, uint64_t* ticks, int reset) {
  if (!h || !ms_out11) return PF_EINVAL;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  int rc = strip_fold(h);
  if (rc) return rc;
  for (int k = 0; k < kStripMarks - 1; ++k) ms_out11[k] = h->strip_ms[k];
  if (ticks) *ticks = h->strip_ticks;
  if (reset) { memset(h->strip_ms, 0, sizeof(h->strip_ms)); h->strip_ticks = 0; }
  return PF_OK;
}

#define STRIP_MARK()                                                                                  \
  do {                                                                                                \
    if (h->prof_on && h->strip_ev && h->strip_mark < kStripMarks)                                     \
      PF_CUDA(h, cudaEventRecord(h->strip_ev[(size_t)h->strip_used * kStripMarks + h->strip_mark++], s)); \
  } while (0)

extern "C" int pf_strip_tick(pf_handle* h, const pf_agents* a, uint32_t flags, uint32_t phases, void* stream) {
  int rc = check_agents(h, a);
  if (rc) return rc;
  if (!(phases & PF_TICK_ALL)) return set_err(h, PF_EINVAL, "pf_strip_tick: no phase selected");
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const bool dep = (flags & PF_STEP_DEPOSIT) != 0, move = (flags & PF_STEP_MOVE) != 0;
  const int rows = h->dev.rows;
  if (a->n > 0) { rc = pf_reserve_agents(h, a->n); if (rc) return rc; }
  if (phases & PF_TICK_A) {  // deposit, halo signal
    if (h->strip_active) return set_err(h, PF_ESTATE, "pf_strip_tick: phase A twice (phases must run A, B, C)");
    if (h->prof_on) {
      if (!h->strip_ev) {
        h->strip_ev = new (std::nothrow) cudaEvent_t[kStripSlots * kStripMarks];
        if (!h->strip_ev) return set_err(h, PF_ENOMEM, "host allocation failed");
        for (int i = 0; i < kStripSlots * kStripMarks; ++i) PF_CUDA(h, cudaEventCreate(&h->strip_ev[i]));
      }
      if (h->strip_used == kStripSlots) { rc = strip_fold(h); if (rc) return rc; }
      h->strip_mark = 0;
    }
    STRIP_MARK();  // 0: start
    h->strip_one_launch = h->strip_one_launch_opt == 1;
    h->strip_dep0 = h->strip_dep1 = 0;
    if ((flags & PF_STEP_HANDOVER) && dep && move && (flags & PF_STEP_SENSE) && a->n > 0) {
      int ty0, ty1;
      if (strip_tile_range(h, a, &ty0, &ty1)) {
        h->strip_dep0 = ty0 * kTileRows;
        h->strip_dep1 = ty1 * kTileRows < rows ? ty1 * kTileRows : rows;
      }
    }
    if (h->strip_predep) {  // the robots that were not handed over deposited at the end of the last tick (phase C)
      if (!dep || a->n == 0)
        return set_err(h, PF_ESTATE, "pf_strip_tick: the tick after a hand-over tick must deposit (PF_STEP_DEPOSIT)");
      h->strip_predep = false;
      rc = agents_deposit(h, a, s, 2, (flags & (PF_STEP_SENSE | PF_STEP_MOVE)) ? 1 : 0);
      if (rc) return rc;
    } else if (dep && a->n > 0) { rc = pf_agents_deposit(h, a, stream); if (rc) return rc; }
    STRIP_MARK();  // 1: deposit
    rc = pf_peer_signal(h, stream);
    if (rc) return rc;
    h->strip_active = true;
  }
  const bool handing = h->strip_dep1 > h->strip_dep0;
  const bool one_launch = handing && h->strip_one_launch;
  // Default: the rows above the handed ones run while the halo is in flight (part A), the handed rows with their
  // passengers and the band below follow behind the move kernel. PF_OPT_STRIP_ONE_LAUNCH steps the whole strip in ONE
  // launch behind the move kernel instead (edge rows and their peer stores included; only the first tile row stays
  // with the separate deposit): measured at 8 x 4096 rows it is 3 % slower (0.439 against 0.427 ms of stencil per
  // tick, and the halo wait is exposed), profiles/r02_notes.md.
  const int lo = 1, hi = rows - 1 > 1 ? rows - 1 : 1;
  const int mid = one_launch ? lo : (handing ? h->strip_dep0 : lo + (hi - lo) / 2);
  if (phases & PF_TICK_B) {  // interior rows while the halo is in flight, halo wait, agents, leavers out
    if (!h->strip_active) return set_err(h, PF_ESTATE, "pf_strip_tick: phase B without phase A");
    // PF_OPT_STRIP_EDGES_SIDE (off by default): on a hand-over tick with the deposit in two parts the small stencil
    // launches leave the caller's stream altogether — the first rows start on the side stream at once (they need
    // neither the halo nor the agents), the band below the handed rows and the two edge rows follow there as soon as
    // the ghost rows are in. Measured at 8 x 4096 rows: 0.679 ms per strip tick against 0.567 in line — the
    // high-priority launches take SM slots from the move kernel's single resident wave and from each other.
    const bool edges_side = handing && !one_launch && move && !h->strip_mig_inline && dep && a->n > 0 &&
                            !h->strip_no_predeposit && h->strip_edges_side_opt && h->strip_one_launch_opt != 2 &&
                            (h->peer_up.set || h->peer_dn.set);
    h->strip_edges_side = edges_side;
    if (edges_side) {
      PF_CUDA(h, cudaEventRecord(h->ev_a, s));
      PF_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_a, 0));
    }
    if (mid > lo) { rc = launch_stencil(h, lo, mid, edges_side ? h->side : s); if (rc) return rc; }
    STRIP_MARK();  // 2: signal + stencil over the first rows
    rc = pf_peer_wait(h, stream);
    if (rc) return rc;
    if (edges_side) {
      PF_CUDA(h, cudaEventRecord(h->ev_halo, s));
      PF_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_halo, 0));
      if (h->strip_dep1 < rows) { rc = launch_stencil(h, h->strip_dep1, rows, h->side); if (rc) return rc; }
      rc = launch_stencil(h, 0, rows < 1 ? rows : 1, h->side);
      if (rc) return rc;
    }
    STRIP_MARK();  // 3: halo wait + ghost rows
    if (a->n > 0 && (flags & (PF_STEP_SENSE | PF_STEP_MOVE))) {
      const uint32_t f = (flags & (PF_STEP_SENSE | PF_STEP_MOVE)) | (handing ? PF_STEP_HANDOVER : 0u);
      rc = pf_agents_step(h, a, f, nullptr, stream);
      if (rc) return rc;
      if (handing && !h->tile_pending) return set_err(h, PF_ESTATE, "pf_strip_tick: hand-over range changed inside a tick");
    }
    STRIP_MARK();  // 4: sense/decide/move (+ bins + grouping)
    // Migration (pack -> signal -> wait -> unpack: six small latency-bound kernels that only touch the agent arrays
    // and the mailboxes) runs on the handle's side stream, beside the stencil launches of phase C, which only touch
    // the field and the tile lists; the caller's stream joins it at the end of the tick, so the next tick's deposit
    // and halo signal still come behind this tick's unpack (what the one-mailbox protocol above relies on).
    const bool mig_side = move && !h->strip_mig_inline;
    if (mig_side) {
      PF_CUDA(h, cudaEventRecord(h->ev_fork, s));
      PF_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    }
    if (move) { rc = strip_pack(h, a, mig_side ? h->side : s); if (rc) return rc; }
    STRIP_MARK();  // 5: pack + migration signal
  }
  if (phases & PF_TICK_C) {  // the remaining rows (handed deposits ride along), swap, arrivals in
    if (!h->strip_active) return set_err(h, PF_ESTATE, "pf_strip_tick: phase C without phase A");
    const bool mig_side_c = move && !h->strip_mig_inline;
    bool predep_c = false;
    if (one_launch) {  // every row of the strip, edge rows and their peer stores included, in one launch
      rc = launch_stencil(h, 0, rows, s, true, true);
      if (rc) return rc;
      h->tile_pending = false;
      h->tile_handed = true;
    } else {
      bool last_done = false;
      // Deposit in two parts: a hand-over tick promises that a depositing tick follows, and that tick's deposit of
      // the robots NOT handed over (edge bands, arrivals) only needs this tick's edge rows and its unpack. So the edge
      // rows are stepped first, and behind the unpack the side stream runs that deposit beside the big launch that
      // carries the handed robots' deposits; the next tick starts with its halo signal almost at once.
      const bool predep = handing && h->tile_pending && mig_side_c && dep && a->n > 0 && !h->strip_no_predeposit &&
                          h->strip_one_launch_opt != 2 && (h->peer_up.set || h->peer_dn.set);
      if (h->strip_edges_side && !predep)
        return set_err(h, PF_ESTATE, "pf_strip_tick: the tick's options changed between phases B and C");
      if (predep) {
        if (!h->strip_edges_side) {  // (else they are on the side stream already, in front of the migration kernels)
          if (h->strip_dep1 < rows) { rc = launch_stencil(h, h->strip_dep1, rows, s); if (rc) return rc; }
          rc = launch_stencil(h, 0, rows < 1 ? rows : 1, s);
          if (rc) return rc;
          PF_CUDA(h, cudaEventRecord(h->ev_edges, s));
        }
        rc = pf_step_field_rows(h, h->strip_dep0, h->strip_dep1, stream);
        if (rc) return rc;
        predep_c = true;
      } else if (handing && h->strip_one_launch_opt == 2 && h->tile_pending) {
        // the handed rows, the band below them and the last row (peer stores included) in one launch: the tiles of
        // the band simply have no passengers
        rc = launch_stencil(h, h->strip_dep0, rows, s, true, false, true);
        if (rc) return rc;
        h->tile_pending = false;
        h->tile_handed = true;
        last_done = true;
      } else if (handing) {
        rc = pf_step_field_rows(h, h->strip_dep0, h->strip_dep1, stream);  // exactly these rows: the deposits ride along
        if (rc) return rc;
        last_done = true;
        if (h->strip_dep1 < rows) { rc = launch_stencil(h, h->strip_dep1, rows, s); if (rc) return rc; }
      } else if (hi > mid) {
        rc = launch_stencil(h, mid, hi, s);
        if (rc) return rc;
      }
      if (!predep) {
        rc = launch_stencil(h, 0, rows < 1 ? rows : 1, s);
        if (rc) return rc;
        if (rows > 1 && !last_done) { rc = launch_stencil(h, rows - 1, rows, s); if (rc) return rc; }
      }
    }
    rc = pf_swap(h);
    if (rc) return rc;
    STRIP_MARK();  // 6: stencil
    const bool mig_side = mig_side_c;
    cudaStream_t ms = mig_side ? h->side : s;
    if (move && (h->peer_up.set || h->peer_dn.set)) {
      k_peer_wait<<<1, 1, 0, ms>>>(h->peer_up.set ? h->peer_flags + 2 : nullptr, h->peer_dn.set ? h->peer_flags + 3 : nullptr,
                                   h->mig_epoch, h->peer_timeout_ns, h->counters);
      PF_LAUNCH_CHECK(h);
    }
    STRIP_MARK();  // 7: migration wait
    if (move) { rc = strip_unpack(h, a, ms); if (rc) return rc; }
    if (predep_c) {  // the EARLY part of the next tick's deposit: behind the unpack and this tick's edge rows
      if (!h->strip_edges_side) PF_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_edges, 0));
      rc = agents_deposit(h, a, h->side, 1, 0);
      if (rc) return rc;
      h->strip_predep = true;
    }
    if (mig_side) {  // join: everything the caller enqueues next is ordered behind the unpack
      PF_CUDA(h, cudaEventRecord(h->ev_join, h->side));
      PF_CUDA(h, cudaStreamWaitEvent(s, h->ev_join, 0));
    }
    k_tick_counter<<<1, 1, 0, s>>>(h->counters);
    PF_LAUNCH_CHECK(h);
    STRIP_MARK();  // 8: unpack
    if (h->prof_on && h->strip_ev) {
      while (h->strip_mark < kStripMarks) STRIP_MARK();  // unused marks: zero-length intervals
      h->strip_used++;
    }
    h->strip_active = false;
    h->strip_edges_side = false;
    h->strip_one_launch = false;
    h->strip_dep0 = h->strip_dep1 = 0;
  }
  return PF_OK;
}

// ---------------------------------------------------------------------------------------------
// counters
// ---------------------------------------------------------------------------------------------
extern "C" int pf_counters_host(pf_handle* h, pf_counters* out, void* stream) {
  if (!h || !out) return PF_EINVAL;
  if (h->tile_handed || h->tile_pending)
    return set_err(h, PF_ESTATE, "%s after a hand-over tick (PF_STEP_HANDOVER): the interior rows and the deposit "
                                 "counters already hold the NEXT tick's deposits; run the last tick before reading "
                                 "without PF_STEP_HANDOVER", "pf_counters_host");
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  unsigned long long c[16];
  PF_CUDA(h, cudaMemcpyAsync(c, h->counters, sizeof(c), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  PF_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
  out->ticks = c[0];
  out->deposits = c[1];
  for (int k = 0; k < 6; ++k) out->decisions[k] = c[2 + k];
  out->food_grabbed = c[8];
  out->food_delivered = c[9];
  out->migrated_out = c[10];
  out->migrated_in = c[11];
  out->migrate_dropped = c[12];
  out->peer_timeouts = c[13];
  return PF_OK;
}

extern "C" int pf_counters_reset(pf_handle* h, void* stream) {
  if (!h) return PF_EINVAL;
  PF_CUDA(h, cudaSetDevice(h->cfg.device));
  PF_CUDA(h, cudaMemsetAsync(h->counters, 0, sizeof(unsigned long long) * 16, (cudaStream_t)stream));
  return PF_OK;
}

extern "C" int pf_launch_count(const pf_handle* h, uint64_t* n) {
  if (!h || !n) return PF_EINVAL;
  *n = h->launches;
  return PF_OK;
}
