/* flint_b200.h -- C ABI of the B200-native FLINT ensemble integrator.
 *
 * This is the drop-in boundary: every entry point replaces one binding of the
 * reference's object API (Fortran; paths below are into the reference tree) and
 * is what an ISO_C_BINDING shim binds (fortran/flint_gpu.f90).
 * Plain pointers and sizes only.  All enum VALUES are ABI and equal the
 * reference's.
 *
 *   reference                                         this header
 *   ------------------------------------------------  -----------------------------
 *   type, extends(DiffEqSys)  src/FLINT_base.f90:220  flint_sys_create / _destroy
 *   ERK_class%Init            src/FLINT_base.f90:318  flint_erk_init
 *                             src/ERKInit.f90:32-342
 *   ERK_class%Integrate       src/FLINT_base.f90:400  flint_erk_integrate (scalar)
 *                             src/ERKIntegrate.f90:29 flint_erk_integrate_batch (ensemble)
 *   ERK_class%Interpolate     src/FLINT_base.f90:479  flint_erk_interpolate (scalar)
 *                             src/ERKInterp.f90:28    flint_erk_interpolate_batch
 *   ERK_class%Info            src/FLINT_base.f90:507  flint_erk_info
 *                             src/ERKInit.f90:346
 *   FLINT_ErrorStrings        src/FLINT_base.f90:64   flint_status_string
 *   final :: erk_destroy      src/ERKInit.f90:402     flint_erk_destroy
 *
 * F and G are not function pointers: the GPU build selects __device__ functors
 * compiled into the library by (system_id, eventset_id).  There is no CPU
 * fallback: every compute entry point returns FLINT_ERROR (detail in
 * flint_last_error()) when no CUDA device is usable.
 */
#ifndef FLINT_B200_H
#define FLINT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes: src/FLINT_base.f90:43-61 ---- */
enum {
    FLINT_SUCCESS = 0, FLINT_EVENT_TERM = 1, FLINT_STIFF_PROBLEM = 2, FLINT_ERROR = 3,
    FLINT_ERROR_PARAMS = 4, FLINT_ERROR_MAXSTEPS = 5, FLINT_ERROR_MEMALLOC = 6, FLINT_ERROR_MEMDEALLOC = 7,
    FLINT_ERROR_DIMENSION = 8, FLINT_ERROR_STEPSZ_TOOSMALL = 9, FLINT_ERROR_INTPSTATES = 10,
    FLINT_ERROR_INTP_ARRAY = 11, FLINT_ERROR_INTP_OFF = 12, FLINT_ERROR_INIT_REQD = 13,
    FLINT_ERROR_EVENTPARAMS = 14, FLINT_ERROR_EVENTROOT = 15, FLINT_ERROR_CONSTSTEPSZ = 16
};
/* ---- event actions (OR-able mask bit): src/FLINT_base.f90:86-102 ---- */
enum {
    FLINT_EVENTACTION_CONTINUE = 0, FLINT_EVENTACTION_TERMINATE = 1, FLINT_EVENTACTION_CHANGESOLY = 2,
    FLINT_EVENTACTION_RESTARTINT = 4, FLINT_EVENTACTION_MASK = 128
};
/* ---- event options: src/FLINT_base.f90:107-120 ---- */
enum { FLINT_EVENTOPTION_ROOTFINDING = 0, FLINT_EVENTOPTION_STEPBEGIN = 1, FLINT_EVENTOPTION_STEPEND = 2 };
/* ---- methods: src/ERK.f90:56-66 ---- */
enum { ERK_DOP853 = 17, ERK_DOP54 = 18, ERK_VERNER98R = 19, ERK_VERNER65E = 20 };

/* ---- DiffEqSys functors compiled into the library (flint_b200/csrc/systems.cuh) ---- */
enum {
    FLINT_SYS_CR3BP_PLANAR = 1,  /* tests/test_CR3BP.f90:167-186  sysparams = {mu}              n = 6 */
    FLINT_SYS_CR3BP_SPATIAL = 2, /* tests/test_WP.f90:88-109      sysparams = {mu}              n = 6 */
    FLINT_SYS_LORENZ = 3,        /* tests/test_Lorenz.f90:45-57   sysparams = {sigma,rho,beta}  n = 3 */
    FLINT_SYS_TWOBODY = 4,       /* tests/test_TBP.f90:46-59      sysparams = {GM}              n = 6 */
    FLINT_SYS_USER_BASE = 100    /* ids >= 100: user functors, registered at build time (csrc/systems.cuh "user systems") or at
                                    run time from their source (flint_sys_register_source);
                                    their event sets are numbered by the functor itself, m <= its M_MAX */
};
enum {
    FLINT_EVSET_NONE = 0,
    FLINT_EVSET_CR3BP_SAMPLE3 = 1,        /* tests/test_CR3BP.f90:97-164, m = 3 (CR3BP_PLANAR) */
    FLINT_EVSET_XY_CROSSING = 2,          /* README.md:121-149,           m = 2 (CR3BP_PLANAR) */
    FLINT_EVSET_APSIS = 3,                /* tests/test_TBP.f90:72-90,    m = 1 (TWOBODY)      */
    FLINT_EVSET_CR3BP_SAMPLE3_TERM = 4,   /* SAMPLE3 detectors, event 2 -> TERMINATE (exercises the action path) */
    FLINT_EVSET_CR3BP_SAMPLE3_RESTART = 5 /* SAMPLE3 detectors, event 2 -> RESTARTINT|MASK */
};

/* ---- arithmetic contract (extension; DESIGN.md "Arithmetic contract") ---- */
enum {
    FLINT_ARITH_STRICT = 0, /* separate roundings for * and +: gfortran -O3 x86-64 semantics; the parity mode */
    FLINT_ARITH_FMA = 1     /* a*b+c sites fused into DFMA: the throughput mode (mirrored by the oracle)      */
};

#define FLINT_MAX_N 8
#define FLINT_MAX_M 4
#define FLINT_MAX_DEVICES 16
#define FLINT_SHARD_CHUNK 4096         /* trajectories per chunk of the block-cyclic deal over devices / ranks */
#define FLINT_MAXNUMSTEPSZEVENTS 4     /* src/ERK.f90:48 */

typedef struct flint_sys flint_sys_t; /* DiffEqSys: n, m, functor ids, system parameters */
typedef struct flint_erk flint_erk_t; /* ERK_class: options + mutable state + device storage */

/* `present` bits mirror Fortran `optional` dummies of Init (src/FLINT_base.f90:318-392) */
enum {
    FLINT_INIT_METHOD = 1 << 0, FLINT_INIT_INTERP_ON = 1 << 1, FLINT_INIT_INTERP_STATES = 1 << 2,
    FLINT_INIT_MIN_STEP = 1 << 3, FLINT_INIT_MAX_STEP = 1 << 4, FLINT_INIT_STEPSZ_PARAMS = 1 << 5,
    FLINT_INIT_EVENTS_ON = 1 << 6, FLINT_INIT_EVENT_STEPSZ = 1 << 7, FLINT_INIT_EVENT_OPTIONS = 1 << 8,
    FLINT_INIT_EVENT_TOL = 1 << 9, FLINT_INIT_ARITH = 1 << 10
};
typedef struct {
    unsigned present;
    int method;                                   /* Method        (default keeps me%method, DOP853 initially) */
    int n_atol; const double* atol;               /* ATol(:)  size 1 or n -- mandatory (reference quirk Q12)   */
    int n_rtol; const double* rtol;               /* RTol(:)                                                    */
    int interp_on;                                /* InterpOn                                                   */
    int n_interp_states; const int* interp_states;/* InterpStates(:), 1-based                                   */
    double min_step_size, max_step_size;          /* MinStepSize (parsed, unused), MaxStepSize (0 = |Xf-X0|)    */
    double stepsz_params[6];                      /* StepSzParams(6): SF, SFmin, SFmax, beta, beta_mult, hbyhoptOld */
    int events_on;                                /* EventsOn                                                   */
    int n_event_stepsz; const double* event_stepsz;   /* EventStepSz(m)                                         */
    int n_event_options; const int* event_options;    /* EventOptions(m)                                        */
    int n_event_tol; const double* event_tol;         /* EventTol(m)                                            */
    int arith;                                    /* FLINT_ARITH_* (extension)                                  */
} flint_init_opts;

/* `present` bits for the optionals of Integrate (src/FLINT_base.f90:400-468).  Xint/Yint, EventStates and StiffTest are
 * present when their pointer arguments are non-NULL: the three bits named after them are accepted and not consulted. */
enum {
    FLINT_INT_CONST_STEPSZ = 1 << 0, FLINT_INT_STEP_ADVANCE = 1 << 1, FLINT_INT_INTSTEPS_ON = 1 << 2,
    FLINT_INT_XINT_YINT = 1 << 3, FLINT_INT_EVENT_MASK = 1 << 4, FLINT_INT_EVENT_STATES = 1 << 5,
    FLINT_INT_STIFF_TEST = 1 << 6, FLINT_INT_PARAMS = 1 << 7
};
typedef struct {
    unsigned present;
    int use_const_stepsz;          /* UseConstStepSz */
    int step_advance;              /* StepAdvance    */
    int int_steps_on;              /* IntStepsOn     */
    int event_mask[FLINT_MAX_M];   /* EventMask(m)   */
    int n_params; const double* params;   /* params(:): at most 8 values, forwarded to user functors that declare load_params()
                                             (csrc/systems.cuh); the built-in systems ignore them, as the reference's own F do.
                                             More than 8 values: FLINT_ERROR_PARAMS */
} flint_int_opts;

/* Natural-step output (Xint/Yint, `allocatable, intent(out)` in the reference: caller-allocated
 * with a capacity here; src/ERKIntegrate.f90:129-140,585-588,642-656).
 * scalar : xint[cap], yint[cap][n] (Fortran Yint(n,cap) column-major), *count = accepted+1
 * batch  : xint[cap][N], yint[n][cap][N], count[N]                                              */
typedef struct { long cap; double* xint; double* yint; int* count; } flint_steps_out;

/* EventStates (src/ERKIntegrate.f90:493,679-681): records [Xev, Yev(1:n), id].
 * scalar : rec[cap][n+2] (Fortran EventStates(n+2,cap)), *count = number of events (may exceed cap)
 * batch  : rec[n+2][cap][N], count[N].  FLINT_MEM_HOST: slots not filled read back as NaN;
 *          FLINT_MEM_DEVICE: slots not filled keep the caller's contents                          */
typedef struct { long cap; double* rec; int* count; } flint_events_out;

/* Where the batch buffers live and how to run */
enum { FLINT_MEM_HOST = 0, FLINT_MEM_DEVICE = 1 };
typedef struct {
    int mem;            /* FLINT_MEM_HOST: pointers are host memory, the call stages H2D/D2H itself;
                           FLINT_MEM_DEVICE: pointers are device memory on `device`, nothing is copied */
    int device;         /* CUDA device ordinal */
    void* stream;       /* cudaStream_t or NULL (default stream); the call is synchronous unless async != 0 */
    int async;          /* FLINT_MEM_DEVICE only: return after enqueueing; caller synchronises the stream */
    long dense_cap;     /* InterpOn: accepted steps the first level of the dense store holds per trajectory (0 = what a third of
                           the free device memory holds, at most MaxSteps).  Not a limit: trajectories that need more are
                           suspended, given a deeper level and resumed (synchronous calls; an async call keeps what fits) */
    int block_threads;  /* 0 = the kernel's own CTA size (a property of the system, 128..512); a smaller multiple of 32 lowers it */
    int blocks_per_sm;  /* 0 = as many CTAs per SM as are resident; a smaller number lowers it (tuning experiments) */
    int event_mode;     /* how accepted steps with an event to locate are handled: 0 = automatic, 1 = inside the
                           stepping kernel, 2 = parked and located in batches by a separate kernel (the default
                           choice unless an unmasked event has an EventStepSz).  Results are identical. */
    int plane_variant;  /* 0 = automatic: ensembles whose trajectories all lie in an invariant coordinate plane of the
                           system (planar CR3BP with z = vz = +0) run a kernel that drops the identically-zero
                           components; 2 = never.  Results are identical. */
    int pipeline;       /* FLINT_MEM_HOST: k >= 2 = the call runs as k chunks of trajectories on two streams, so that the
                           host<->device copies of one chunk travel while another computes (pays for short integrations,
                           where the copies are a large part of the call); 0 / 1 = one chunk.  Results are identical. */
    int round_len;      /* small shards (a few trajectories per GPU lane): the step kernel time-shares its lanes in rounds of k step
                           attempts, each round on the trajectories furthest behind in x, so that all lanes stay busy until the
                           ensemble finishes together.  0 = automatic (between 1 and 8 trajectories per lane),
                           1 = never, k >= 2 = rounds of k attempts.  Results are identical. */
    int n_devices;      /* FLINT_MEM_HOST: > 1 = the ensemble is sharded over devices[0 .. n_devices-1]: chunks of shard_chunk
                           trajectories are dealt round-robin (block-cyclic), each device runs its shard on its own
                           stream from its own host thread, results land at the caller's offsets.  No collective exists on this
                           path.  0 / 1 = the single `device`.  Results are identical to the one-device call. */
    int devices[FLINT_MAX_DEVICES];
    long shard_chunk;   /* trajectories per dealt chunk; 0 = FLINT_SHARD_CHUNK */
    int dense_mode;     /* InterpOn: what Integrate keeps per accepted step.  0 / 2 = the step's log (2n+2 doubles): Interpolate
                           rebuilds the coefficients in registers and evaluates the grid from there, Bip never exists in memory
                           (the fused path); 1 = the coefficient record Bip(nI, 0:pstar) as the reference stores it
                           (src/ERKIntegrate.f90:555-566), built once at the end of Integrate: the cheaper choice when one
                           dense output is interpolated many times.  Results are identical. */
    int dense_spill;    /* InterpOn, FLINT_MEM_HOST: the dense output of an ensemble that does not fit in device memory (the
                           reference keeps Xint / Bip in host memory, src/ERKInit.f90:321-332).  k >= 2: the call runs as k parts per
                           device, one after the other; after a part's Integrate its dense store is moved to host memory, and
                           Interpolate brings the parts back one at a time (so Yarr needs device memory for one part only).
                           1 = automatic: as many parts as keep a part's store within a third of the free device memory (one
                           part = no spill).  0 = never.  Results are identical. */
} flint_exec;

/* ---- the partition of an ensemble over shards (devices of one call, or ranks of a multi-process job): trajectory index space
 * [0, N) in chunks of `chunk`, chunk c -> shard c mod n_shards.  Pure host arithmetic (no CUDA call). ---- */
long flint_b200_shard_count(long N, int n_shards, int shard, long chunk);                 /* trajectories of `shard` */
long flint_b200_shard_global_index(long N, int n_shards, int shard, long chunk, long local_index);   /* -1 when out of range */
/* out[0 .. count-1] = global indices of the shard's trajectories in local order; returns count */
long flint_b200_shard_indices(long N, int n_shards, int shard, long chunk, long* out);

/* ---- DiffEqSys ---- */
flint_sys_t* flint_sys_create(int system_id, int eventset_id, int n, int m, const double* sysparams, int nsysparams);
void flint_sys_destroy(flint_sys_t*);
/* Run-time registration of a user DiffEqSys (the reference's user extends DiffEqSys with F and G, src/FLINT_base.f90:220-309; here
 * they are device functors).  `source`: the text of a header that defines `struct <struct_name> : flint::UserSystem<n, m_max>` in
 * namespace flint with `static constexpr int ID = FLINT_SYS_USER_BASE + k`, `load`, `accel<FMA>` (F) and, if has_events, `events`
 * (G) -- the same text a build-time user header holds (flint_b200/csrc/systems.cuh; tests/user_systems/).  NVRTC compiles it
 * against the kernel headers embedded in the library, with the library's own options; the integrate kernels of a (method,
 * arithmetic, events) combination are built on first use and cached on disk ($FLINT_B200_RTC_CACHE, default
 * $HOME/.cache/flint_b200).  Needs libnvrtc at run time (dlopen; $FLINT_B200_NVRTC overrides the search), no GPU for the
 * registration itself.  Returns FLINT_SUCCESS and the id for flint_sys_create, else FLINT_ERROR_PARAMS / FLINT_ERROR_DIMENSION /
 * FLINT_ERROR with the compiler's log in flint_last_error(). */
int flint_sys_register_source(const char* struct_name, const char* source, int has_events, int* system_id_out);
/* Compiles and caches the integrate kernels of one (method, arithmetic, events) combination of a system registered from source,
 * without loading them (no GPU needed): warms the cache ahead of the first Integrate. */
int flint_sys_precompile(int system_id, int method, int arith, int events_on);

/* ---- ERK_class ---- */
flint_erk_t* flint_erk_create(void);
void flint_erk_destroy(flint_erk_t*);

/* Init: returns me%status.  Validation and error codes as src/ERKInit.f90:68-190,296-309. */
int flint_erk_init(flint_erk_t*, const flint_sys_t*, int max_steps, const flint_init_opts*);

/* Integrate, FLINT-shaped (one trajectory, host pointers).  xf, stepsz in/out; stifftest in/out
 * (NULL = absent); steps/events NULL = absent.  Runs the same kernel with N = 1. */
int flint_erk_integrate(flint_erk_t*, double x0, const double* y0, double* xf, double* yf, double* stepsz,
                        const flint_int_opts*, flint_steps_out*, flint_events_out*, int* stifftest);

/* Integrate, ensemble.  SoA layouts: y0[n][N], yf[n][N], counters[4][N] = TotalSteps, AcceptedSteps,
 * RejectedSteps, FCalls.  x0: NULL -> x0_scalar for all.  xf[N], stepsz[N] in/out (stepsz NULL -> 0,
 * not returned).  status[N] per trajectory.  stifftest[N] in/out or NULL (then opts' StiffTest is
 * absent = 0).  Returns the object-level status (first non-success per-trajectory code is NOT
 * promoted: FLINT_SUCCESS means the batch ran; inspect status[]). */
int flint_erk_integrate_batch(flint_erk_t*, long N, double x0_scalar, const double* x0, const double* y0,
                              double* xf, double* yf, double* stepsz, const flint_int_opts*,
                              int* counters, int* status, flint_steps_out*, flint_events_out*, int* stifftest,
                              const flint_exec*);

/* Interpolate (scalar object state of the last scalar Integrate): yarr[G][nI] (Fortran Yarr(nI,G)). */
int flint_erk_interpolate(flint_erk_t*, const double* xarr, long G, double* yarr, int has_save, int save_coeffs);

/* Interpolate the dense output of the last batch Integrate onto a grid shared by all trajectories.
 * layout 0: yarr[N][G][nI] (each trajectory Fortran-compatible Yarr(nI,G)); layout 1: yarr[nI][G][N] (SoA).
 * status[N] (nullable) receives per-trajectory FLINT_ERROR_INTP_ARRAY where xarr leaves [Xint(1),Xint(acc+1)]. */
int flint_erk_interpolate_batch(flint_erk_t*, const double* xarr, long G, double* yarr, int layout, int* status,
                                int has_save, int save_coeffs, const flint_exec*);
/* The same with one grid per trajectory, xarr[N][G] (each row strictly monotone in that trajectory's direction of integration and
 * inside its span: ERK_class%Interpolate called once per trajectory with its own Xarr, src/ERKInterp.f90:28-76). */
int flint_erk_interpolate_batch_grids(flint_erk_t*, const double* xarr, long G, double* yarr, int layout, int* status,
                                      int has_save, int save_coeffs, const flint_exec*);

/* Info (scalar state).  Any pointer may be NULL (= optional absent).  Unlike the reference (quirk Q8)
 * hf/xf/yf are written to the arguments they are named after. */
int flint_erk_info(const flint_erk_t*, int* last_status, int* nsteps, int* naccept, int* nreject, int* nfcalls,
                   int* interp_ready, double* x0, double* y0, double* hf, double* xf, double* yf);

const char* flint_status_string(int status);   /* the 17 strings of src/FLINT_base.f90:64-82, verbatim */
const char* flint_last_error(void);            /* detail of the last FLINT_ERROR (CUDA error text etc.) */

/* ---- device utilities used by bench.py (not part of the reference API) ---- */
int flint_b200_device_count(void);
/* Measures the fp64 FMA peak of `device` with a register-resident DFMA microbenchmark
 * (8 independent chains per thread, full occupancy).  Returns TFLOP/s (FMA = 2 flop), < 0 on error. */
double flint_b200_measure_fp64_peak(int device, double* sm_clock_mhz_out);
/* Number of kernel launches this library has issued in this process (bench.py's gpu_launches). */
long flint_b200_launch_count(void);
/* 1 when the library was built with -D FLINT_BOUNDS_CHECK=1 (every device-side index tested, trap on the first one out of range) */
int flint_b200_bounds_check(void);
/* Last integrate kernel's duration in ms measured with CUDA events on the launch stream (batch calls). */
double flint_b200_last_kernel_ms(void);
/* Device self-test of the branch-free division / square-root helpers the right-hand sides use
 * (csrc/arith.cuh): n operand sets drawn from `seed`; result3 = {results that differ from the plain
 * IEEE operator although the helper's range test passed (must be 0), divisions that passed the range
 * test, square roots that passed}.  Returns FLINT_SUCCESS or FLINT_ERROR. */
int flint_b200_selftest_arith(long n, unsigned long long seed, unsigned long long* result3);

#ifdef __cplusplus
}
#endif
#endif /* FLINT_B200_H */
