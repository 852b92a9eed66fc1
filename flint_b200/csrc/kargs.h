/* kargs.h -- plain structs shared by the host side of the C ABI and the kernels. */
#ifndef FLINT_B200_KARGS_H
#define FLINT_B200_KARGS_H

#ifndef __CUDACC_RTC__              /* NVRTC (rtc.cu) declares the runtime's device-side types itself */
#include <cuda_runtime.h>
#endif
#include "../../include/flint_b200.h"

#ifndef FLINT_BLOCK_THREADS
#define FLINT_BLOCK_THREADS 128      /* __launch_bounds__ of the integrate kernels */
#endif
#ifndef FLINT_MIN_BLOCKS
#define FLINT_MIN_BLOCKS 4
#endif
/* stage-vector slots of each method that the hot kernel keeps in shared memory (bit i = slot i); tuning knobs */
#ifndef FLINT_SMEM_SLOTS_DOP853
#define FLINT_SMEM_SLOTS_DOP853 0u
#endif
#ifndef FLINT_SMEM_SLOTS_DOP54
#define FLINT_SMEM_SLOTS_DOP54 0u
#endif
#ifndef FLINT_SMEM_SLOTS_VERNER98R
#define FLINT_SMEM_SLOTS_VERNER98R 0u
#endif
#ifndef FLINT_SMEM_SLOTS_VERNER65E
#define FLINT_SMEM_SLOTS_VERNER65E 0u
#endif
/* which out-of-line copy of the right-hand side stage-iteration `it` of `n` calls (arith.cuh: accel_call) */
#ifndef FLINT_RHS_SPLIT
#define FLINT_RHS_SPLIT 0            /* stage iterations below this index call copy 1, the rest copy 0 */
#endif
#define FLINT_RHS_COPY(it, n) ((it) < FLINT_RHS_SPLIT ? 1 : 0)
#ifndef FLINT_SYNC_BLOCK
#define FLINT_SYNC_BLOCK 1
#endif
#define FLINT_DEFAULT_BLOCK FLINT_BLOCK_THREADS
#define FLINT_MAX_BLOCK FLINT_BLOCK_THREADS

#define FLINT_NCOEF 256                /* capacity of KArgs::coef (largest pool: Verner98R, 254 values) */

#define FLINT_DENSE_LEVELS 6          /* capacity levels of the dense-output store (each 4x the one before) */

/* FLINT_BOUNDS_CHECK=1 (python -m flint_b200.build --tag _chk -D FLINT_BOUNDS_CHECK=1): every index into the work lists, the
 * state block, the event records, the dense store and the interpolation kernels' shared-memory buffers is tested on the device
 * and the kernel traps on the first one out of range -- the stand-in for compute-sanitizer's memcheck, which is closed on the
 * GPU pool this was developed on.  The GPU test-suite runs clean under that build (profiles/r02_bounds_check.log). */
#ifndef FLINT_BOUNDS_CHECK
#define FLINT_BOUNDS_CHECK 0
#endif
#if FLINT_BOUNDS_CHECK && defined(__CUDA_ARCH__)
#define FLINT_CHK(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define FLINT_CHK(cond) ((void)0)
#endif

namespace flint {

/* ---- dense-output storage (InterpOn; src/ERKIntegrate.f90:555-566,660-676; src/ERKInit.f90:321-332).
 * The reference sizes Xint/Bip for MaxSteps + 1 per object; for an ensemble that is N * MaxSteps records, so the store grows on
 * demand instead: level 0 holds cap_0 steps for every trajectory; a trajectory that fills its levels is suspended (state block),
 * the host adds a level 4x as deep for exactly the suspended ones and resumes them (capi.cu).  No trajectory ever loses a step.
 * One RECORD per accepted step, `rstride` doubles (even), rec[rstride-1] = its kind:
 *   kDenseLog    the step as the hot loop left it: X, h, Y(n), F0(n)                        -- 2n+2 doubles
 *   kDenseEvLog  a step committed by the event code (possibly truncated, quirk Q6): the above + h_stage, Y1(n)  -- 3n+3 doubles
 *   kDenseCoef   X0, X1, Bip(ii, j) as [j][ii]: (pstar+1)*nI + 2 doubles (erk_dense_kernel turns logs into this, or evaluates
 *                the grid points of the step straight from registers and never stores it: the fused Interpolate)
 * Xint(j+1) of the reference is X0 of record j (and X1 = X + h of the last one): there is no separate Xint array. ---- */
constexpr double kDenseCoef = 0.0, kDenseLog = 1.0, kDenseEvLog = 2.0;
struct DenseView {
    int nlev;                                   /* levels in use; 0 = InterpOn is off */
    int rstride;                                /* doubles per record */
    long cum[FLINT_DENSE_LEVELS + 1];           /* steps held below level l: cum[0] = 0, cum[l+1] = cum[l] + cap_l */
    double* base[FLINT_DENSE_LEVELS];           /* level l: [rows_l][cap_l][rstride] */
    const int* slot[FLINT_DENSE_LEVELS];        /* level l >= 1: row of trajectory t in that level (-1: none); level 0: row = t */
    /* record of accepted step a (0-based) of trajectory t; a < cum[nlev] */
    __host__ __device__ __forceinline__ double* rec(long t, long a) const
    {
        int l = 0;
        FLINT_CHK(t >= 0 && a >= 0 && nlev > 0 && a < cum[nlev]);
        while (l + 1 < nlev && a >= cum[l + 1]) ++l;
        const long row = l == 0 ? t : (long)slot[l][t];
        FLINT_CHK(row >= 0);
        return base[l] + (row * (cum[l + 1] - cum[l]) + (a - cum[l])) * (long)rstride;
    }
    __host__ __device__ __forceinline__ long capacity() const { return nlev > 0 ? cum[nlev] : 0; }
};
/* end abscissa of a record, whatever its kind */
__host__ __device__ __forceinline__ double dense_x1(const double* r, int rstride) { return r[rstride - 1] == kDenseCoef ? r[1] : r[0] + r[1]; }

/* ATol/RTol as Init stores them (scalar or per-component; src/ERKInit.f90:76-87) */
struct Tol {
    double a[FLINT_MAX_N], r[FLINT_MAX_N];
    int scalar;
    __host__ __device__ __forceinline__ double at(int c) const { return scalar ? a[0] : a[c]; }
    __host__ __device__ __forceinline__ double rt(int c) const { return scalar ? r[0] : r[c]; }
};

/* Everything an integrate launch needs; passed by value (constant bank, warp-uniform reads). */
struct KArgs {
    long N;
    int n, m;                    /* DiffEqSys n, m (m <= SYS::M_MAX) */
    int evset;
    double sp[8];                /* system parameters */
    int max_steps;
    Tol tol;
    double hmax_opt;             /* MaxStepSize, 0 => |Xf - X0| (src/ERKIntegrate.f90:86,237) */
    double szp[6];               /* StepSzParams */
    int const_step, step_once;
    int stiff_present, stiff_mode;
    /* events */
    int events_on;
    unsigned evmask_bits;        /* static EventMask */
    double ev_stepsz[FLINT_MAX_M];
    int ev_opt[FLINT_MAX_M];
    double etol[FLINT_MAX_M];
    /* inputs / outputs (device pointers, SoA over trajectories) */
    double x0_scalar; const double* x0;
    const double* y0;            /* [n][N] */
    double* xf;                  /* [N] in/out */
    double* yf;                  /* [n][N] */
    double* stepsz;              /* [N] in/out or NULL */
    int* counters;               /* [4][N] or NULL */
    int* status;                 /* [N] */
    int* stifftest;              /* [N] in/out or NULL */
    long ev_cap; int* ev_count; double* ev_rec;          /* [(n+2)][cap][N] */
    int steps_on; long steps_cap; double* xint; double* yint; int* steps_count;   /* IntStepsOn output */
    /* InterpOn storage */
    DenseView dv;                                /* where accepted steps are logged */
    long dense_avail;                            /* a trajectory is suspended when it has this many accepted steps (= dv.capacity(), or
                                                    huge for calls that cannot grow the store: those drop the records beyond it) */
    int* dense_count;                            /* [N] accepted steps logged */
    int* ovf_list; unsigned long long* n_ovf;    /* trajectories suspended because their levels are full */
    int nI; int istates[FLINT_MAX_N];
    /* erk_dense_kernel: coefficient records are written to dvo (may equal dv: in place); fused Interpolate (dense_eval != 0):
     * the grid ix[G] (per trajectory: ix[N][G]) is evaluated into iy[N][G][nI] straight from the coefficients in registers */
    DenseView dvo;
    int dense_eval, ix_per_traj; long iG; const double* ix; double* iy; const int* istatus;
    int dense_bulk;                              /* dense_eval == 0: the block's records are staged in shared memory and written by one bulk copy */
    /* trajectory state block + work lists (erk_kernel.cuh) */
    double* sd; int* si; long scap;              /* [fields][scap] */
    const int* list_in; const unsigned long long* n_in;   /* trajectories of this pass (NULL: all of [0, N)) */
    int* list_out; unsigned long long* n_out;    /* EVMODE 2: trajectories parked by this pass */
    int* zflag;                                  /* NULL, or: 1 = every trajectory lies in the system's invariant plane (plane-variant kernel runs) */
    unsigned long long* work_counter;
    /* small shards (erk_kernel.cuh: GANG): round_req 0 = automatic, 1 = never, k >= 2 = gang-scheduled rounds of k step attempts;
     * n_items_exact: the host knows the pass's item count (synchronous calls); round_len is filled by launch_step */
    int round_req, n_items_exact, round_len;
    double params[8]; int n_params;              /* Integrate(params=...) forwarded to user functors (systems.cuh: load_params) */
    double coef[FLINT_NCOEF];    /* the method's Butcher / error / dense coefficients (filled by the host, capi.cu) */
    double powc[32];             /* constants of the controller's pow (arith.cuh: det_pow_constants) */
};

/* erk_dense_kernel: CTA = 128 consecutive steps of one trajectory; shared memory of its fused evaluation (sB, sX, sG) in doubles */
constexpr int kDenseBlock = 128;
__host__ __device__ constexpr int dense_smem_doubles_ps(int pstar, int nI)
{
    const int rs = (pstar + 1) * nI;
    return kDenseBlock * (rs + (rs & 1)) + 2 * (kDenseBlock + 2);
}

/* erk_dense_kernel: a CTA walks trajectories blockIdx.x, blockIdx.x + gridDim.x, ... and prefetches the next one's step logs
 * (cp.async) into a per-thread slot of kDensePfSlot doubles at the end of its dynamic shared memory, + two step counts */
constexpr int kDensePfSlot = 26;                       /* 22 doubles of log (3n+3 <= 21 for n <= 6), the kind, padding: 13 x 16 bytes */
__host__ __device__ constexpr int dense_prefetch_doubles() { return kDenseBlock * kDensePfSlot + 2; }

struct InterpArgs {
    long N, G;
    int nI, pstar;
    DenseView dv;                    /* coefficient records (kDenseCoef) */
    const int* dcount;
    const double* xarr; int per_traj; /* grid: xarr[G] shared, or xarr[N][G] */
    double* yarr; int layout;        /* 0: [N][G][nI], 1: [nI][G][N] */
    int* status;                     /* [N] */
    int fma;
    long chunk; int nchunks;         /* filled by launch_interp */
};
#ifndef __CUDACC_RTC__             /* ---- host side: launchers and the table of integrate kernels ---- */
cudaError_t launch_interp(const InterpArgs&, int grid_up, int grid_down, int sm_count, cudaStream_t, int* launches = nullptr);   /* *launches = kernels launched */
/* per-trajectory validation of the grid (src/ERKInterp.f90:60-76) against any kind of records */
cudaError_t launch_interp_validate(const InterpArgs&, int grid_up, int grid_down, cudaStream_t);
/* slot[list[j]] = j for j < n: rows of a new level of the dense store (registry.cu) */
cudaError_t launch_dense_slots(int* slot, const int* list, long n, cudaStream_t);
cudaError_t launch_set_counters(unsigned long long* ctr, unsigned long long c0, unsigned long long c1, unsigned long long c2,
                                unsigned long long c3, cudaStream_t);   /* registry.cu */

/* The integrate kernels of one (method, system, arithmetic, events) combination, as the launchers see them: entry points that
 * cudaLaunchKernel accepts -- the host stub of a kernel compiled into the library (inst_*.cu) or the cudaKernel_t of a module
 * NVRTC built at run time from a user's source (rtc.cu) -- plus the compile-time facts the launch needs.  One code path launches
 * both. */
struct StepKernel {              /* one erk_step_kernel variant (event mode, plane / general) */
    const void* classic;         /* persistent lanes, one trajectory per lane at a time */
    const void* gang;            /* gang-scheduled rounds (small shards); NULL: not instantiated (EVMODE 1) */
    int block;                   /* CTA size the kernel was compiled for (__launch_bounds__; the shared-memory KStore is laid out for it) */
    int smem;                    /* bytes of dynamic shared memory of the KStore */
};
struct KernelEntry {
    int method, system;
    bool fma, events;
    const void* init;            /* erk_init_kernel */
    StepKernel step;             /* EVMODE 0 (events == false) or 1 (events located in the lane loop) */
    StepKernel step_park;        /* EVMODE 2 (classic == NULL when events == false) */
    const void* event;           /* erk_event_kernel (NULL when events == false) */
    const void* dense_coeffs;    /* erk_dense_kernel: dense-output coefficients from the step logs */
    StepKernel step_plane;       /* plane variant (SYS::zc) of `step` for EVMODE 0, of `step_park` for events; classic == NULL if none */
    const void* dense_plane;     /* plane variant of `dense_coeffs`; NULL if none */
    int pstar;                   /* METHOD::PSTAR (shared memory of the fused dense kernel) */
    int n_state_doubles, n_state_ints;
    int sys_n, sys_m_max;        /* SYS::N, SYS::M_MAX (validation of user systems in flint_sys_create) */
};
void register_kernel(const KernelEntry&);
const KernelEntry* find_kernel(int method, int system, bool fma, bool events);
const KernelEntry* find_system(int system);          /* any entry of that system, or NULL */
int kernel_count();
void load_method_coefficients(int method, double* dst);   /* Butcher / error / dense coefficients into KArgs::coef */

/* grid x block threads of kernel k with the parameter block by value */
cudaError_t launch_kernel(const void* k, const KArgs&, dim3 grid, int block, size_t smem, cudaStream_t);
/* the persistent step kernels size themselves: one resident wave of the device (SM count x resident CTAs of the kernel's own
 * CTA size), never more lanes than items; block_req / bps_req > 0 lower the CTA size / the CTAs per SM.  Chooses the
 * gang-scheduled variant for small shards (erk_kernel.cuh). */
cudaError_t launch_step(const StepKernel&, const KArgs&, int sm_count, long n_items, int block_req, int bps_req, cudaStream_t);
/* erk_dense_kernel: grid = (trajectories, blocks of 128 steps up to the store's capacity) */
cudaError_t launch_dense(const KernelEntry&, bool plane, const KArgs&, cudaStream_t);
#endif

}  // namespace flint
#endif
