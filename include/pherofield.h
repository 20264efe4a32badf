/*
 * pherofield.h — C ABI of the B200-native pheromone-field hot path.
 *
 * This is the drop-in boundary for the per-tick pheromone update (deposit, evaporation,
 * diffusion) and the agent loop (sense pheromone -> pick a heading -> move) of
 * ChitteshKumar/swarm_robotics_pheromones. The reference has no FFI of its own: its boundary is
 * five Python methods and two tick drivers. Each entry point below names the
 * reference interface (file:line, relative to the reference checkout) that it replaces.
 * `sup`  = Phase-3/pheromone_algo_supervisor.py, `algo` = Phase-3/pheromone_algo.py,
 * `graph` = Phase-3/graph.py.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types, no exceptions across the boundary.
 *   - every function returns 0 (PF_OK) or a negative PF_E* code; pf_last_error(h) has the text.
 *   - array arguments are DEVICE pointers owned by the caller (torch tensors as the buffer
 *     carrier) unless the name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it.
 *   - a handle is not thread-safe; distinct handles are independent.
 *   - there is no CPU fallback: without a CUDA device pf_create fails with PF_ECUDA.
 *
 * Grid convention (algo:351 `pheromone_grid[int(x)][int(y)] += intensity`): the FIRST world
 * coordinate (x) selects the row, the second (y) the column; indices are truncated toward zero.
 * Multi-GPU row strips therefore partition the x axis.
 */
#ifndef PHEROFIELD_H
#define PHEROFIELD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_ABI_VERSION 2

#define PF_MAX_CHANNELS 4
#define PF_MAX_FOOD 16

/* error codes */
#define PF_OK 0
#define PF_EINVAL (-1)   /* bad argument / config */
#define PF_ECUDA (-2)    /* CUDA runtime error (text in pf_last_error) */
#define PF_ENOMEM (-3)   /* device allocation failed */
#define PF_ESTATE (-4)   /* call not valid in this state (e.g. sharded handle, missing ghosts) */

/* robot states, sup:180-183 (RobotState). 0 marks an empty agent slot (multi-GPU holes). */
#define PF_STATE_EMPTY 0
#define PF_STATE_SEARCHING 1
#define PF_STATE_HOMING 2
#define PF_STATE_IDLE 3

/* the five sensed directions in the reference's dict order, sup:390 */
#define PF_DIR_CURRENT 0
#define PF_DIR_FRONT 1
#define PF_DIR_BACK 2
#define PF_DIR_LEFT 3
#define PF_DIR_RIGHT 4

/* discrete decision of adjust_robot_behavior, sup:429-516 */
#define PF_DEC_STRAIGHT 0 /* "no change needed": both wheels = v            sup:463-466,501-504 */
#define PF_DEC_LEFT 1     /* SEARCHING sup:455-458; HOMING sup:491-494                          */
#define PF_DEC_RIGHT 2    /* SEARCHING sup:459-462; HOMING sup:496-500                          */
#define PF_DEC_KEEP 3     /* HOMING inside home radius, no branch taken: speeds persist sup:489-500 */
#define PF_DEC_STOP 4     /* reached home sup:510-513                                          */
#define PF_DEC_NONE 5     /* no decision this tick (pre-start random walk sup:736-744, IDLE, empty) */

/* sense geometry */
#define PF_SENSE_WORLD 0   /* fixed world-frame cell offsets (mirrors sup:399-414, "ref-DRUL") */
#define PF_SENSE_HEADING 1 /* offsets rotated with the agent heading, sense_dist metres away   */

/* agent meta word: one u32 per agent */
#define PF_META_STATE_MASK 0x3u        /* bits 0-1: PF_STATE_*                                  */
#define PF_META_MOVED 0x4u             /* bit 2: has_moved() result for the NEXT deposit sup:260 */
#define PF_META_DEC_SHIFT 3            /* bits 3-5: last PF_DEC_*                               */
#define PF_META_DEC_MASK 0x38u
#define PF_META_CARRY 0x40u            /* bit 6: carrying food (fsm mode)                       */
#define PF_META_HASPOS 0x80u           /* bit 7: a previous observed position exists (sup:281)  */
#define PF_META_COUNT_SHIFT 8          /* bits 8-30: deposits made so far (len(global grid))    */
#define PF_META_COUNT_MASK 0x7FFFFFu   /* 23 bits                                               */
#define PF_META_HANDED 0x80000000u     /* bit 31: this robot's NEXT deposit was handed to the stencil of the
                                          tick that just ran (strips, PF_STEP_HANDOVER); set by the move
                                          kernel, cleared by the next deposit. Never set between pf_step calls */

/* step flags */
#define PF_STEP_DEPOSIT 0x1u /* run deposit_pheromone                                   sup:577 */
#define PF_STEP_SENSE 0x2u   /* run sense + adjust (off before t = 5 s)                 sup:581 */
#define PF_STEP_FIELD 0x4u   /* run evaporate(+diffuse)                       algo:354, sup:594 */
#define PF_STEP_MOVE 0x8u    /* integrate the diff-drive stand-in (Webots physics)              */
#define PF_STEP_ALL 0xFu
#define PF_STEP_HANDOVER 0x10u /* pf_agents_step on a strip: robots that end up in the rows of pf_tile_rows hand
                                  their NEXT deposit to the stencil (see pf_tile_rows)                    */

typedef struct pf_config {
  int32_t abi_version;  /* must be PF_ABI_VERSION */
  int32_t device;       /* CUDA device ordinal */

  /* ---- field storage (algo:244 grid, generalised) ---- */
  int32_t H, W, C;      /* global rows (x), cols (y), channels */
  int32_t row0, row1;   /* rows [row0,row1) owned by this handle; 0,H when not sharded */
  float cell_size;      /* metres per cell */
  float origin_x;       /* world x of the low edge of row 0 */
  float origin_y;       /* world y of the low edge of col 0 */

  /* ---- evaporate + diffuse: f' = keep * (f + D * L(f)) ---- */
  int32_t stencil;                  /* 0 = evaporate only (algo:354-356), 5 or 9 point Laplacian */
  float evap_keep[PF_MAX_CHANNELS]; /* (1 - evaporation_rate); reference 1-0.1      algo:248,356 */
  float diffusion[PF_MAX_CHANNELS]; /* D; 0 = off. 9-point uses D/6 folded on the host           */
  float delete_threshold;           /* values <= this become 0 (sup:603); <= 0 disables          */

  /* ---- deposit rule (sup:230-234, 291-358) ---- */
  float intensity_explore; /* 1.0  sup:232 */
  float intensity_high;    /* 20.0 sup:233, every `high_every`-th SEARCHING deposit sup:316-318 */
  float intensity_homing;  /* 5.0  sup:234 */
  int32_t high_every;      /* 3 */
  float move_threshold;    /* 0.0015 m, sup:260 */
  int32_t deposit_channel[4]; /* channel written by an agent in state s (index PF_STATE_*) */
  int32_t sense_channel[4];   /* channel read by an agent in state s */

  /* ---- sense geometry (sup:379-417) ---- */
  int32_t sense_mode;       /* PF_SENSE_* */
  int32_t sense_drow[5];    /* PF_SENSE_WORLD: row offset (x axis) per PF_DIR_* */
  int32_t sense_dcol[5];    /* PF_SENSE_WORLD: col offset (y axis) per PF_DIR_* */
  float sense_dist;         /* PF_SENSE_HEADING: metres from the agent centre */

  /* ---- decide (sup:185-226 PID, sup:429-516) ---- */
  float pid_kp, pid_kd;     /* 1.0, 0.01 (ki is never applied, sup:217-218) */
  float cruise_speed;       /* desiredV = 1.0 rad/s, sup:197 */
  float max_speed;          /* MAX_SPEED = 3.14 rad/s, sup:74 */
  float braitenberg_l;      /* sum_j coeff[j][0]*(1 - ps_j/512) precomputed, sup:468-470 */
  float braitenberg_r;      /* sum_j coeff[j][1]*(1 - ps_j/512) */
  float home_radius;        /* 0.7 m, sup:489 */
  float home_lo_cm, home_hi_cm; /* reachedHome window 5..6 (0.05 <= round(d,2) <= 0.06) sup:667 */
  float nest_x, nest_y;     /* start_position, sup:248 */
  int32_t n_food;           /* >= 1 */
  float food_x[PF_MAX_FOOD], food_y[PF_MAX_FOOD]; /* get_food_position, sup:419-422 */
  float food_radius;        /* fsm: SEARCHING -> HOMING inside this distance (camera stand-in) */
  int32_t fsm;              /* 0: states never change (supervisor without camera);
                               1: SEARCHING->HOMING at food, HOMING->SEARCHING at nest
                                  (experiment2.py:615-616, 746-764) */

  /* ---- move: diff-drive stand-in for Webots physics ---- */
  float dt;                 /* 0.064 s: two 32 ms Webots steps per controller tick */
  float wheel_radius;       /* 0.02 m,  E-puck.proto:100 */
  float axle_length;        /* 0.052 m, E-puck.proto:86  */

  /* ---- multi-GPU strips: records per direction per tick the migration mailbox can take; 0 = default (16384) */
  int32_t migrate_cap;
} pf_config;

/* Agents, structure-of-arrays, caller-owned device memory. `n` slots are processed; slots whose
 * state is PF_STATE_EMPTY are skipped. gid is only read by migration and may be NULL otherwise. */
typedef struct pf_agents {
  int64_t n;
  float* x;       /* world x (row axis) */
  float* y;       /* world y (col axis) */
  float* theta;   /* heading, radians in [-pi, pi] */
  float* wl;      /* adjusted_left_speed, persists between ticks  sup:249 */
  float* wr;      /* adjusted_right_speed                          sup:250 */
  uint32_t* meta; /* PF_META_* */
  uint32_t* gid;  /* global agent id (stable across migration) */
} pf_agents;

typedef struct pf_counters {
  uint64_t ticks;
  uint64_t deposits;        /* agents that deposited */
  uint64_t decisions[6];    /* histogram over PF_DEC_* */
  uint64_t food_grabbed;    /* fsm: SEARCHING->HOMING transitions  (experiment2.py:752-755) */
  uint64_t food_delivered;  /* fsm: HOMING->SEARCHING at nest      (experiment2.py:615-616) */
  uint64_t migrated_out;    /* agents packed for a neighbour strip */
  uint64_t migrated_in;
  uint64_t migrate_dropped; /* records that did not fit a send buffer or found no empty slot */
  uint64_t peer_timeouts;   /* pf_peer_wait / migration waits that gave up (neighbour strip dead or out of lockstep) */
} pf_counters;

typedef struct pf_handle pf_handle;

/* -------------------------------------------------------------------------------------------
 * lifecycle
 * ------------------------------------------------------------------------------------------- */
int pf_abi_version(void);

/* Fill `cfg` with the reference constants (sup:230-250, sup:74-77, algo:244-249) for an H x W x C
 * field. Replaces the hard-coded constants of PheromoneAlgorithm.__init__. */
int pf_config_default(pf_config* cfg, int32_t H, int32_t W, int32_t C);

/* Allocate the double-buffered fp32 SoA field [2][C][rows+2 ghost][pitch] in HBM, zeroed, plus the
 * sort scratch. Replaces `pheromone_grid = [[0.0]*10 ...]` (algo:244) and
 * `self.pheromone_grids = {}` (sup:237). */
int pf_create(const pf_config* cfg, pf_handle** out);
int pf_destroy(pf_handle* h);
const char* pf_last_error(const pf_handle* h); /* h may be NULL: last create error */
int pf_get_config(const pf_handle* h, pf_config* out);

/* Reserve scratch for up to `max_agents` per deposit call (sort keys, values, cub temp).
 * Called implicitly (and synchronously) on first use when not called up front. */
int pf_reserve_agents(pf_handle* h, int64_t max_agents);

/* -------------------------------------------------------------------------------------------
 * field access
 * ------------------------------------------------------------------------------------------- */
/* Device pointer to local row 0 (global row cfg.row0) of `channel` in the current (which=0) or
 * next (which=1) buffer; the ghost rows are at ptr - pitch and ptr + rows*pitch. pitch in floats. */
int pf_field_ptr(pf_handle* h, int which, int channel, float** ptr, int64_t* pitch);
int pf_field_fill(pf_handle* h, int channel, float value, void* stream);
/* copy rows [row0,row1) x W of one channel from/to a dense host array [rows][W] */
int pf_field_upload_host(pf_handle* h, int channel, const float* src_host, void* stream);
int pf_snapshot_host(pf_handle* h, int channel, float* dst_host, void* stream);
/* Write the reflect (zero-flux) ghost rows of the CURRENT buffer at the arena's outer edges
 * (row0 == 0 and/or row1 == H). Interior strip edges are filled by the halo exchange. */
int pf_fill_ghosts(pf_handle* h, void* stream);
/* fp64 sum of one channel over the owned rows, written to *sum_host after a stream sync. */
int pf_field_sum_host(pf_handle* h, int channel, double* sum_host, void* stream);

/* -------------------------------------------------------------------------------------------
 * the hot path
 * ------------------------------------------------------------------------------------------- */
/* Deterministic deposit: cell = (int((x-ox)/cs), int((y-oy)/cs)); agents are radix-sorted by
 * (channel, cell) (stable => agent order inside a cell), each cell's amounts are summed in that
 * fixed order by a warp-aggregated segmented sum, and the total is added to the field with one
 * store per touched cell. No atomics. `channel` may be NULL (all channel 0). Positions outside
 * rows [row0,row1) or outside the arena are skipped.
 * Replaces `pheromone_grid[int(x)][int(y)] += intensity` (algo:351), the dict insert of
 * deposit_pheromone (sup:312-318, 346-348) and `heatmap[..] += 20` (graph:23). */
int pf_deposit(pf_handle* h, int64_t n, const float* x, const float* y, const float* amount,
               const uint8_t* channel, void* stream);

/* Deposit driven by agent state: gate on has_moved (sup:308-310), amount by state and the
 * every-3rd rule on the agent's own deposit counter (sup:312-318, 346-348), channel by state.
 * Updates meta (counter). Replaces PheromoneAlgorithm.deposit_pheromone (sup:291-358). */
int pf_agents_deposit(pf_handle* h, const pf_agents* a, void* stream);

/* Fused evaporate + diffuse, `nsteps` times, swapping the double buffer each step. Reflect
 * boundary at the arena edge. On a sharded handle (row0 > 0 or row1 < H) only nsteps == 1 is
 * accepted and the ghost rows of the current buffer must have been exchanged by the caller.
 * Replaces `grid[i][j] *= (1-evaporation_rate)` (algo:354-356); the diffusion term has no
 * reference arithmetic and follows oracle/pf_oracle.c. */
int pf_step_field(pf_handle* h, int nsteps, void* stream);
/* Sharded driver pieces: compute local rows [row_begin,row_end) of the next buffer without
 * swapping (interior rows first while the halo is in flight, edge rows after), then pf_swap. */
int pf_step_field_rows(pf_handle* h, int row_begin, int row_end, void* stream);
int pf_swap(pf_handle* h);

/* Gather the five DRUL values per agent: out5[i*5 + PF_DIR_*]; NaN marks a direction whose cell
 * lies outside the arena (the reference's empty list). `state` may be NULL (channel of
 * SEARCHING). Replaces PheromoneAlgorithm.sense_pheromone (sup:379-417). */
int pf_sense(pf_handle* h, int64_t n, const float* x, const float* y, const float* theta,
             const uint32_t* meta, float* out5, void* stream);

/* sense -> ordered arg-max -> PID -> wheel speeds -> diff-drive move, one thread per agent.
 * flags: PF_STEP_SENSE and/or PF_STEP_MOVE. `decisions_out` (u8 per agent) may be NULL.
 * Replaces sense_pheromone + adjust_robot_behavior + PIDController.update
 * (sup:379-417, 429-516, 185-226) and stands in for Webots physics. */
int pf_agents_step(pf_handle* h, const pf_agents* a, uint32_t flags, uint8_t* decisions_out,
                   void* stream);

/* Whole ticks in the reference order deposit -> sense/decide -> evaporate/diffuse -> move
 * (sup:575-607, 753-754). Sensing is enabled once time >= start_delay (5 s, sup:581,732);
 * before that no deposit happens either (startPheromone is not called, sup:731-744).
 * Not valid on a sharded handle (the host drives the exchange between the phases).
 * Replaces Controller.startPheromone / Controller.homing (sup:565-607, 627-663). */
int pf_step(pf_handle* h, const pf_agents* a, int nsteps, void* stream);
/* pf_step with a phase mask (PF_STEP_*): e.g. without PF_STEP_MOVE when the caller (Webots in the
 * reference) owns the positions and feeds them through pf_agents_observe every tick. */
int pf_step_ex(pf_handle* h, const pf_agents* a, int nsteps, uint32_t flags, void* stream);
/* Take externally measured poses (supervisor.getSelf().getPosition(), sup:255-258): shifts
 * current -> previous and sets the moved flag = sqrt(dx^2+dy^2) >= move_threshold, True on the
 * first observation. theta_new may be NULL (heading kept).
 * Replaces storingPosition + has_moved (sup:279-289, 260-273). */
int pf_agents_observe(pf_handle* h, const pf_agents* a, const float* x_new, const float* y_new,
                      const float* theta_new, void* stream);
/* Physically re-order the agent arrays (all seven, in place) by cell, row-major, empty slots last.
 * Agents move well under one cell per tick, so calling this every few dozen ticks keeps the sense
 * gathers and the deposit scatter local in HBM. Identity is carried by gid. Results of every other
 * call are independent of the order (per-cell deposit sums are order-independent for the
 * reference's exactly representable amounts 1, 5, 20). No reference counterpart. */
int pf_agents_sort(pf_handle* h, const pf_agents* a, void* stream);
int pf_set_time(pf_handle* h, double time_s, double start_delay_s);
int pf_get_time(const pf_handle* h, double* time_s, uint64_t* ticks);

/* I * exp(-rate * elapsed) elementwise in fp64. Replaces evaporate_pheromone (sup:519-525). */
int pf_evaporate_trail(pf_handle* h, int64_t n, const double* elapsed, const double* intensity,
                       double rate, double* out, void* stream);

/* -------------------------------------------------------------------------------------------
 * multi-GPU row strips (no reference counterpart)
 * ------------------------------------------------------------------------------------------- */
/* Agents whose row left [row0,row1) are appended (ordered by slot) to the up (row < row0) or
 * down (row >= row1) send buffer as 8-word (32 B) records {x,y,theta,wl,wr,meta,gid,0}; their slot
 * becomes PF_STATE_EMPTY. Each buffer starts with an 8-word header whose first u32 is the record
 * count. `cap` = record capacity of each buffer (excluding the header): (cap+1)*8 u32 words. */
#define PF_MIGRATE_WORDS 8
int pf_migrate_pack(pf_handle* h, const pf_agents* a, uint32_t* up_buf, uint32_t* down_buf,
                    int64_t cap, void* stream);
/* Append the records of a received buffer into empty slots (lowest slot first, deterministic). */
int pf_migrate_unpack(pf_handle* h, const pf_agents* a, const uint32_t* buf, int64_t cap,
                      void* stream);
/* Two received buffers in one pass: the records of `buf` (from the strip above: small cell keys)
 * take the lowest empty slots in ascending order, those of `buf2` (from the strip below) the
 * highest empty slots in descending order, so arrivals sit next to agents with similar keys and the
 * arrays stay nearly ordered between two pf_agents_sort calls. Either buffer may be NULL. */
int pf_migrate_unpack2(pf_handle* h, const pf_agents* a, const uint32_t* buf, const uint32_t* buf2,
                       int64_t cap, void* stream);

/* Peer-store halo over NVLink P2P — the halo exchange fused into the producing kernels. Once the
 * neighbours are attached, every kernel that writes this strip's first / last interior row (the
 * stencil's edge rows into the next buffer, the deposit into the current buffer) also stores the
 * value into the neighbour strip's mailbox (a small IPC-mapped block) through its mapped address;
 * no halo message is sent, and pf_peer_wait moves the mailbox rows into the ghost rows.
 * Per tick: deposit -> pf_peer_signal (epoch flag into the neighbours' memory) -> interior stencil
 * -> pf_peer_wait (own flag words) -> agents, edge rows. All strips must swap buffers in lockstep.
 * After the field was changed from outside (upload, fill) call pf_peer_push_edges + signal/wait. */
typedef struct pf_peer_info {
  void* buf0;  /* the strip's mailboxes for buffer parity 0 / 1 (device pointers valid in THIS process) */
  void* buf1;
  void* flags; /* u32 epoch words: halo [0] written by the strip above, [1] by the strip below; migration [2], [3] */
  int64_t rows, pitch, chan_stride;
  void* mig;   /* migration mailbox: two record buffers of (mig_cap + 1) * PF_MIGRATE_WORDS u32 — [0] filled by the
                  strip above, [1] by the strip below (pf_strip_tick packs leavers straight into the neighbour's) */
  int64_t mig_cap;
} pf_peer_info;
int pf_peer_export(pf_handle* h, pf_peer_info* out);                /* this strip, raw pointers */
int pf_peer_ipc_handles(pf_handle* h, uint8_t* out192);             /* cudaIpcMemHandle_t + pad  */
int pf_peer_ipc_open(pf_handle* h, const uint8_t* in192, const pf_peer_info* dims, pf_peer_info* out);
int pf_peer_attach(pf_handle* h, const pf_peer_info* up, const pf_peer_info* down); /* NULL = none */
int pf_peer_push_edges(pf_handle* h, void* stream);
int pf_peer_signal(pf_handle* h, void* stream);
int pf_peer_wait(pf_handle* h, void* stream);

/* PF_ESTATE when a bounded peer wait gave up since the last pf_counters_reset (a neighbour strip is dead or the
 * strips do not tick in lockstep): the stream is not wedged, but field and agents are no longer valid.
 * Synchronises the stream. */
int pf_peer_status(pf_handle* h, void* stream);

/* One whole tick of a row strip enqueued by ONE call — no host round trips between the phases, leavers packed
 * straight into the neighbours' migration mailboxes over NVLink (peer stores + an epoch flag, like the halo; no
 * NCCL message). Needs pf_peer_attach on the sides that have a neighbour (pf_peer_info.mig set). `flags`:
 * PF_STEP_DEPOSIT (deposit_pheromone), PF_STEP_SENSE, PF_STEP_MOVE (without it nothing migrates), and
 * PF_STEP_HANDOVER = the interior rows' next deposits may ride on this tick's stencil (see pf_tile_rows: not on
 * the last tick before the field is read nor on the tick before pf_agents_sort). The field step always runs.
 * `phases`: PF_TICK_ALL for one process per GPU. Several strips driven by ONE host thread on ONE stream (tests)
 * must interleave: phase A for every strip, then B for every strip, then C — a strip's wait can only be
 * satisfied by kernels enqueued before it.
 *   A: deposit -> halo signal
 *   B: stencil over the first rows -> halo wait -> sense/decide/move (+ hand-over) -> pack leavers -> migration signal
 *   C: stencil over the remaining rows (handed deposits ride along) -> swap -> migration wait -> unpack
 * With PF_STEP_MOVE the migration kernels (pack, signal, wait, unpack) run on the handle's own side stream, beside
 * the stencil launches of phase C, and `stream` joins them before the call's last kernel: for the caller the tick
 * remains one ordered unit of work on `stream`. On a PF_STEP_HANDOVER tick the side stream goes one step further
 * behind the unpack: the NEXT tick's deposit of the robots that were not handed over (edge bands, arrivals) — that
 * tick's phase A then only deposits the handed robots, i.e. nothing unless the stencil could not apply them. A
 * hand-over tick must therefore be followed by a tick with PF_STEP_DEPOSIT (PF_ESTATE otherwise), and poses replaced
 * in between (pf_agents_observe) do not change where that tick's deposits went: all of them were taken at the poses
 * this tick's move kernel produced. PF_OPT_STRIP_MIG_INLINE / PF_OPT_STRIP_NO_PREDEPOSIT keep everything in line.
 * Replaces Controller.startPheromone / homing (sup:565-607, 627-663) on a strip; order as pf_step. */
#define PF_TICK_A 0x1u
#define PF_TICK_B 0x2u
#define PF_TICK_C 0x4u
#define PF_TICK_ALL 0x7u
int pf_strip_tick(pf_handle* h, const pf_agents* a, uint32_t flags, uint32_t phases, void* stream);

/* Per-phase device times of pf_strip_tick while profiling is enabled (pf_profile_enable), summed over `ticks`:
 * [0] deposit, [1] halo signal + stencil over the first rows, [2] halo wait + ghost rows, [3] sense/decide/move
 * (+ hand-over bins and grouping), [4] pack + migration signal, [5] stencil over the remaining rows, [6] migration
 * wait, [7] unpack, [8..10] unused. The marks are events on the caller's stream: what runs on the side stream
 * (migration, the early part of the next deposit) shows up as the wait for it at the end of [7], and [4] / [6] are
 * then enqueue times only. Synchronises the device. */
int pf_strip_profile(pf_handle* h, double* ms_out11, uint64_t* ticks, int reset);

/* pf_snapshot_host, pf_field_sum_host and pf_counters_host return PF_ESTATE between a hand-over tick and the next
 * deposit (the interior rows and the deposit counters are one tick ahead). */
int pf_counters_host(pf_handle* h, pf_counters* out_host, void* stream);
int pf_counters_reset(pf_handle* h, void* stream);

/* number of kernels this handle has launched (bench.py reports it as gpu_launches) */
int pf_launch_count(const pf_handle* h, uint64_t* n);

/* Row strips: deposits riding on the stencil (what pf_step does by itself on one GPU).
 * [row_begin, row_end) = the local rows whose next-tick deposits the stencil can apply, empty when the
 * feature is off for this handle (not a strip, no pf_agents_sort yet, stencil variant, amounts that are not small
 * integers, ...). They never include the first 64 rows nor the last 64 next to a neighbour. Protocol per tick, driven by the host:
 *   pf_agents_deposit            (skips the robots whose deposit the previous tick's stencil applied)
 *   pf_step_field_rows(1, row_begin)                       before the move kernel, as before
 *   pf_agents_step(flags | PF_STEP_HANDOVER)               bins + groups the interior rows' deposits
 *   pf_step_field_rows(row_begin, row_end)                 exactly this range, ONE call: they ride along
 *   pf_step_field_rows(...) for the remaining rows, pf_swap
 * After such a tick the rows [row_begin, row_end) already hold the NEXT tick's deposits: do not hand over on
 * the last tick before reading the field, on the tick before pf_agents_sort (refused), nor if the next tick
 * will not deposit. */
int pf_tile_rows(pf_handle* h, const pf_agents* a, int* row_begin, int* row_end);

/* pf_step diagnostics: ticks whose stencil was handed the next tick's deposits, and how many of
 * those fell back on the device to the separate deposit kernels (a tile too crowded).
 * Synchronises the device. */
int pf_tile_stats(pf_handle* h, uint64_t* attempted, uint64_t* fell_back);
/* Crowded tiles since the last pf_agents_sort: tiles (channel x 64 rows x 1024 columns) whose deposits no longer fit the
 * per-tile lists (more than one robot per few cells, e.g. a nest) and go through a per-cell side buffer instead, and
 * the number of such buffers available. Synchronises the device. */
int pf_tile_dense(pf_handle* h, uint64_t* n_dense, uint64_t* pool_slots);

/* Per-phase device times of pf_step, from CUDA events recorded on the launching streams while
 * profiling is enabled (bench.py's live roofline numbers). Times are sums over `ticks` ticks. */
typedef struct pf_profile {
  uint64_t ticks;
  double deposit_ms;  /* key build + radix sort + segmented sum */
  double stencil_ms;  /* evaporate + diffuse kernel (runs concurrently with the agent kernel) */
  double agents_ms;   /* sense/decide/move kernel, on the side stream */
} pf_profile;
int pf_profile_enable(pf_handle* h, int on);
int pf_profile_read(pf_handle* h, pf_profile* out_host, int reset);

/* -------------------------------------------------------------------------------------------
 * sparse timestamped trail: the shipped supervisor's per-robot trail dict for
 * N robots, fp64, one thread per robot. Discrete outputs (deposit amounts, counts, sense lists,
 * arg-max direction, decision) are exact; real-valued ones agree with CPython/glibc to ~1e-15.
 * Replaces PheromoneAlgorithm's dict state + deposit/sense/adjust/evaporate and the tick drivers
 * Controller.startPheromone / homing (sup:237-250, 291-525, 565-663).
 * ------------------------------------------------------------------------------------------- */
typedef struct pf_trail pf_trail;
typedef struct pf_trail_out { /* device arrays of n_robots entries each; any may be NULL */
  double* amount;       /* deposited intensity, NaN = "nothing deposited" (sup:321-322) */
  int32_t* n_global;    /* len(global_pheromone_grid[name]) */
  int8_t* direction;    /* arg-max PF_DIR_* of sup:440-450, -1 if the robot did not sense */
  int8_t* decision;     /* PF_DEC_* */
  double* wl;           /* adjusted_left_speed  */
  double* wr;           /* adjusted_right_speed */
  int32_t* n_live;      /* len(pheromone_grids[name]) after this tick's evaporation */
  double* sum_live;     /* sum of live intensities after evaporation */
  uint64_t* masks;      /* [5][n]: bit j set = live entry j (insertion order, before this tick's
                           evaporation) is in the sensed list of direction PF_DIR_* */
  double* pre_time;     /* [capacity][n]: times of the live entries as the sense step saw them */
  double* pre_intensity;/* [capacity][n]: their intensities (before this tick's evaporation) */
  int32_t* n_pre;       /* number of such entries */
} pf_trail_out;
int pf_trail_create(int device, int64_t n_robots, int capacity /* live entries per robot, <= 64 */,
                    pf_trail** out);
int pf_trail_destroy(pf_trail* t);
const char* pf_trail_last_error(const pf_trail* t);
/* start_host [n][3] (start_position, sup:248), food_host [3] (sup:419-422), state_host [n] or NULL */
int pf_trail_configure(pf_trail* t, const double* start_host, const double* food_host,
                       const uint8_t* state_host);
int pf_trail_set_state(pf_trail* t, int64_t robot, int state);
/* one supervisor tick for every robot: pos_xyz [n][3] and ps [n][8] are device fp64 arrays */
int pf_trail_tick(pf_trail* t, double t_now, double start_delay, const double* pos_xyz,
                  const double* ps, const pf_trail_out* out, void* stream);
/* PF_ESTATE (and *n_dropped > 0) once a deposit found a robot's live list full: from then on the sense results no
 * longer match the reference's dict. Synchronises the stream. */
int pf_trail_dropped(pf_trail* t, uint64_t* n_dropped, void* stream);
int pf_trail_entries(pf_trail* t, double** time, double** intensity, double** x, double** y,
                     double** z, int** n_live, int* capacity);

/* -------------------------------------------------------------------------------------------
 * offline heat-map: graph.py:17-24 — per trajectory point `heatmap[|int(y*H)|,
 * |int(x*W)|] += 20` followed by scipy.ndimage.gaussian_filter(heatmap, sigma) (mode 'reflect',
 * truncate 4.0), fp64. weights_host = the normalised 1-D kernel's centre and right half
 * (radius + 1 values), computed on the host the way scipy's _gaussian_kernel1d does.
 * ------------------------------------------------------------------------------------------- */
typedef struct pf_heatmap pf_heatmap;
int pf_heatmap_create(int device, int H, int W, int radius, const double* weights_host, pf_heatmap** out);
int pf_heatmap_destroy(pf_heatmap* h);
const char* pf_heatmap_last_error(const pf_heatmap* h);
int pf_heatmap_accumulate(pf_heatmap* h, int64_t n, const double* x_host, const double* y_host,
                          double amount, void* stream);
int pf_heatmap_snapshot_host(pf_heatmap* h, double* dst_host, void* stream);

/* tuning options */
#define PF_OPT_OVERLAP_AGENTS 1 /* pf_step: run sense/decide/move on a side stream beside the stencil */
#define PF_OPT_NO_FUSED_KEYS 2  /* pf_step: always build deposit keys in their own kernel */
#define PF_OPT_NO_OWNER_DEPOSIT 4 /* always use the global radix sort for the agent deposit */
#define PF_OPT_FORCE_OWNER_DEPOSIT 5 /* use the sort-free deposit after pf_agents_sort even where the
                                        density heuristic prefers the radix sort (dense clusters) */
#define PF_OPT_NO_TILE_DEPOSIT 6 /* pf_step: never hand the next tick's deposits to the stencil kernel
                                   (default: done where the sort-free deposit is in use and the TMA
                                   stencil runs; falls back on the device when a tile is too crowded) */
#define PF_OPT_STRIP_ONE_LAUNCH 8 /* pf_strip_tick: 1 = step the whole strip in one stencil launch behind the move kernel;
                                    2 = the rows above the handed ones still run while the halo is in flight, the
                                    handed rows, the band below them and the last row (peer stores included) go in one
                                    launch; 0 = separate launches */
#define PF_OPT_STRIP_PART_A_TILES 9 /* strips: tile rows (64 rows each) stepped before the move kernel while the halo is
                                      in flight = the rows that cannot hand their deposits over; 0 = one tile row */
#define PF_OPT_STRIP_MIG_INLINE 11 /* pf_strip_tick: keep the migration kernels (pack, signal, wait, unpack) on the caller's
                                     stream, in line with the stencil launches (default 0: they run beside the stencil
                                     on the handle's side stream and the caller's stream joins them at the end of the tick) */
#define PF_OPT_STRIP_NO_PREDEPOSIT 12 /* pf_strip_tick: keep the whole deposit at the start of the tick (default 0: on a hand-over
                                        tick the NEXT tick's deposit of the robots that were not handed over — edge bands and
                                        arrivals — runs at the end of the tick on the side stream, behind the unpack and beside
                                        the stencil launch that carries the handed robots' deposits) */
#define PF_OPT_STRIP_EDGES_SIDE 13 /* pf_strip_tick: the small stencil launches of a hand-over tick (first rows, band below the
                                     handed rows, the two edge rows) on the side stream as well (default 0: measured slower) */
#define PF_OPT_PEER_TIMEOUT_MS 7 /* bound of the peer waits in milliseconds (default 5000) */
#define PF_OPT_USE_GRAPH 3      /* pf_step: replay pairs of ticks from a CUDA graph (default: on for
                                   fields of <= 2^24 floats per buffer, where a tick is launch-bound) */
int pf_set_option(pf_handle* h, int option, int value);

/* which stencil kernel variant pf_step_field uses: 0 = auto, 1 = register sliding window (LDG),
 * 2..5 = TMA + mbarrier pipelined shared-memory ring (4x4, 8x3, 2x8, 4x6 rows x stages). */
int pf_set_stencil_variant(pf_handle* h, int variant, int rows_per_cta);

#ifdef __cplusplus
}
#endif
#endif /* PHEROFIELD_H */
