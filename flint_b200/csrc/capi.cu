/* capi.cu -- host side of the C ABI (include/flint_b200.h).
 *
 * Mirrors the object semantics of the reference: flint_erk_t holds what
 * ERK_class holds (options parsed by Init, counters/status of the last scalar
 * Integrate, dense-output storage), validation and error codes follow
 * src/ERKInit.f90:68-190,296-309 and src/ERKIntegrate.f90:94-199.  All compute
 * is the CUDA kernels; there is no host fallback.
 */
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/flint_b200.h"
#include "kargs.h"
#include "arith.cuh"

using flint::KArgs;

static thread_local std::string g_last_error;
static std::atomic<long> g_launches{0};                 /* handles on different devices launch from different host threads */
static std::atomic<double> g_last_kernel_ms{0.0};

static int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
namespace flint {
void set_last_error(const std::string& msg) { g_last_error = msg; }
/* rtc.cu: systems registered from source at run time */
const KernelEntry* build_rt_kernel(int method, int system, bool fma, bool events);
bool rt_system_facts(int system, int* n, int* m_max, int* has_events);
/* the kernels of a combination: compiled into the library, already built at run time, or built now */
static const KernelEntry* get_kernel(int method, int system, bool fma, bool events)
{
    const KernelEntry* ke = find_kernel(method, system, fma, events);
    if (!ke && system >= FLINT_SYS_USER_BASE) ke = build_rt_kernel(method, system, fma, events);
    return ke;
}
}  // namespace flint
static bool cuda_ok(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return true;
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}

struct flint_sys {
    int system_id, eventset_id, n, m;
    double sp[8];
};

struct DevCache {                      /* device staging buffers kept between calls (see DevBuf) */
    int device = -1;
    std::vector<void*> ptr;
    std::vector<size_t> size;
    void release() { for (void* p : ptr) if (p) cudaFree(p); ptr.clear(); size.clear(); device = -1; }
};

struct ScopedCache : DevCache { ~ScopedCache() { release(); } };

struct Workspace {                     /* trajectory state block + work lists of Integrate; grows, never shrinks */
    int device = -1;
    size_t bytes = 0;
    void* ptr = nullptr;
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; bytes = 0; device = -1; }
};

/* Device-resident dense output of the last Integrate with InterpOn (kargs.h: DenseView): capacity levels of step records.
 * kind_coef: the records are (to be) coefficient records, rstride holds them; else step logs only (the fused Interpolate). */
struct DenseStore {
    int device = -1;
    long N = 0;
    int n = 0, nI = 0, pstar = 0, rstride = 0;
    bool kind_coef = false, events = false;
    int nlev = 0;
    long cap[FLINT_DENSE_LEVELS] = {0}, rows[FLINT_DENSE_LEVELS] = {0};
    double* base[FLINT_DENSE_LEVELS] = {nullptr};
    int* slot[FLINT_DENSE_LEVELS] = {nullptr};
    int* dcount = nullptr;
    int* dzflag = nullptr;             /* device word: 1 = every trajectory lies in the system's invariant plane (plane-variant kernels) */
    /* flint_exec.dense_spill: between Integrate and Interpolate the store of a part of the ensemble waits in host memory (the
     * reference's Xint / Bip are host arrays, src/ERKInit.f90:321-332); spilled = the device copy is gone, hbase etc. hold it */
    bool spilled = false;
    double* hbase[FLINT_DENSE_LEVELS] = {nullptr};
    int* hslot[FLINT_DENSE_LEVELS] = {nullptr};
    int* hdcount = nullptr;
    int hzflag[2] = {0, 0};
    bool hpinned = false;
    void* host_alloc(size_t bytes)
    {
        void* p = nullptr;
        if (hpinned) { if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) return p; cudaGetLastError(); return nullptr; }
        return malloc(bytes);
    }
    void host_free(void* p) { if (!p) return; if (hpinned) cudaFreeHost(p); else free(p); }
    void release_host()
    {
        for (int l = 0; l < FLINT_DENSE_LEVELS; ++l) { host_free(hbase[l]); host_free(hslot[l]); hbase[l] = nullptr; hslot[l] = nullptr; }
        host_free(hdcount); hdcount = nullptr; spilled = false;
    }
    void release_device()
    {
        if (dzflag) cudaFree(dzflag);
        dzflag = nullptr;
        for (int l = 0; l < FLINT_DENSE_LEVELS; ++l) {
            if (base[l]) cudaFree(base[l]);
            if (slot[l]) cudaFree(slot[l]);
            base[l] = nullptr; slot[l] = nullptr;
        }
        if (dcount) cudaFree(dcount);
        dcount = nullptr;
    }
    size_t level_bytes(int l) const { return (size_t)rows[l] * (size_t)cap[l] * (size_t)rstride * 8; }
    /* device -> host, then the device copy is freed.  Page-locked host memory when it can be had (the copies then run at the
     * link's rate), pageable otherwise.  false + nothing changed on failure. */
    bool spill(cudaStream_t s)
    {
        if (spilled || !allocated()) return spilled;
        hpinned = getenv("FLINT_B200_SPILL_PAGEABLE") == nullptr;
        bool ok = true;
        for (int pass = 0; pass < 2; ++pass) {             /* second pass: pageable, if page-locked memory ran out */
            ok = true;
            for (int l = 0; l < nlev && ok; ++l) {
                hbase[l] = (double*)host_alloc(level_bytes(l));
                ok = hbase[l] != nullptr;
                if (ok && l >= 1) { hslot[l] = (int*)host_alloc((size_t)N * 4); ok = hslot[l] != nullptr; }
            }
            if (ok) { hdcount = (int*)host_alloc((size_t)N * 4); ok = hdcount != nullptr; }
            if (ok || !hpinned) break;
            release_host(); hpinned = false;
        }
        if (!ok) { release_host(); return false; }
        for (int l = 0; l < nlev && ok; ++l) {
            ok = cudaMemcpyAsync(hbase[l], base[l], level_bytes(l), cudaMemcpyDeviceToHost, s) == cudaSuccess;
            if (ok && l >= 1) ok = cudaMemcpyAsync(hslot[l], slot[l], (size_t)N * 4, cudaMemcpyDeviceToHost, s) == cudaSuccess;
        }
        ok = ok && cudaMemcpyAsync(hdcount, dcount, (size_t)N * 4, cudaMemcpyDeviceToHost, s) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(hzflag, dzflag, 8, cudaMemcpyDeviceToHost, s) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(s) == cudaSuccess;
        if (!ok) { cudaGetLastError(); release_host(); return false; }
        release_device();
        spilled = true;
        return true;
    }
    /* host -> device (the host copy stays: an Interpolate with SaveInterpCoeffs leaves the part spilled) */
    bool restore(cudaStream_t s)
    {
        if (!spilled) return allocated();
        bool ok = true;
        for (int l = 0; l < nlev && ok; ++l) {
            ok = cudaMalloc(&base[l], level_bytes(l)) == cudaSuccess;
            if (ok && l >= 1) ok = cudaMalloc(&slot[l], (size_t)N * 4) == cudaSuccess;
        }
        ok = ok && cudaMalloc(&dcount, (size_t)N * 4) == cudaSuccess && cudaMalloc(&dzflag, 8) == cudaSuccess;
        for (int l = 0; l < nlev && ok; ++l) {
            ok = cudaMemcpyAsync(base[l], hbase[l], level_bytes(l), cudaMemcpyHostToDevice, s) == cudaSuccess;
            if (ok && l >= 1) ok = cudaMemcpyAsync(slot[l], hslot[l], (size_t)N * 4, cudaMemcpyHostToDevice, s) == cudaSuccess;
        }
        ok = ok && cudaMemcpyAsync(dcount, hdcount, (size_t)N * 4, cudaMemcpyHostToDevice, s) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(dzflag, hzflag, 8, cudaMemcpyHostToDevice, s) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(s) == cudaSuccess;
        if (!ok) { cudaGetLastError(); release_device(); return false; }
        return true;
    }
    void release()
    {
        release_host();
        if (dzflag) cudaFree(dzflag);
        dzflag = nullptr;
        for (int l = 0; l < FLINT_DENSE_LEVELS; ++l) {
            if (base[l]) cudaFree(base[l]);
            if (slot[l]) cudaFree(slot[l]);
            base[l] = nullptr; slot[l] = nullptr; cap[l] = rows[l] = 0;
        }
        if (dcount) cudaFree(dcount);
        dcount = nullptr; N = 0; nlev = 0;
    }
    bool allocated() const { return nlev > 0 && base[0] != nullptr; }
    flint::DenseView view() const
    {
        flint::DenseView v{};
        v.nlev = nlev; v.rstride = rstride; v.cum[0] = 0;
        for (int l = 0; l < nlev; ++l) { v.cum[l + 1] = v.cum[l] + cap[l]; v.base[l] = base[l]; v.slot[l] = slot[l]; }
        return v;
    }
    size_t bytes() const
    {
        size_t b = 0;
        for (int l = 0; l < nlev; ++l) b += (size_t)rows[l] * (size_t)cap[l] * (size_t)rstride * 8;
        return b;
    }
};

struct flint_erk {
    /* FLINT_class / ERK_class state (src/FLINT_base.f90:125-211, src/ERK.f90:72-117) */
    int status = FLINT_SUCCESS;
    bool inited = false;
    const flint_sys* sys = nullptr;
    flint_sys sys_copy{};
    int method = ERK_DOP853;
    int max_steps = 0;
    bool scalar_tol = true;
    double atol[FLINT_MAX_N]{}, rtol[FLINT_MAX_N]{};
    bool interp_on = false;
    int nI = 0; int istates[FLINT_MAX_N]{};
    double min_step = 0, max_step = 0;
    double szp[6]{};
    bool events_on = false;
    double ev_stepsz[FLINT_MAX_M]{}; int ev_opt[FLINT_MAX_M]{}; double etol[FLINT_MAX_M]{};
    int arith = FLINT_ARITH_STRICT;
    int s = 0, sint = 0, p = 8, q = 7, pstar = 7;
    /* last scalar Integrate (Info) */
    int total = 0, accepted = 0, rejected = 0, fcalls = 0;
    double X0 = 0, Xf = 0, hf = 0;
    double Y0[FLINT_MAX_N]{}, Yf[FLINT_MAX_N]{};
    /* dense storage */
    DenseStore dense;
    /* two pipeline slots (host-buffer calls on large ensembles run as chunks on two streams, below); slot 0 serves every other call */
    Workspace ws[2];
    DevCache stage_int[2];
    cudaStream_t pipe_stream[2] = {nullptr, nullptr};
    int pipe_device = -1;
    bool dense_is_scalar = false;
    bool dense_plane_ok = false;       /* the plane-variant kernels were eligible for the call that filled the dense store */
    bool dense_allocated = false;      /* allocated(me%Xint) */
    double last_kernel_ms = 0.0;       /* device time of the last batch Integrate of this handle (CUDA events on its stream) */
    unsigned long long* pinned = nullptr;   /* page-locked words the kernels' counters are read back into */
    /* one sub-handle per device of a sharded call (flint_exec.n_devices > 1): same options, own stream / workspace / dense store */
    std::vector<flint_erk*> shard;
    std::vector<cudaStream_t> shard_stream;
    std::vector<int> shard_device;
    long shard_N = 0, shard_chunk = 0;      /* the partition of the last sharded Integrate (Interpolate follows it) */
    int shard_devices = 0;                  /* distinct device slots of that call: shard s ran on slot s % shard_devices, the parts
                                               of a slot one after the other (flint_exec.dense_spill) */
    bool shard_spill = false;               /* the parts' dense stores wait in host memory */
};

/* ------------------------------------------------------------------ misc */
static const char* const kErrorStrings[17] = {   /* src/FLINT_base.f90:64-82 */
    "Successfully Completed", "Termination Event Occurred", "Problem appears to be stiff", "Unknown error",
    "Incorrect user-provided parameters", "Maximum integration steps reached", "Memory allocation error",
    "Memory deallocation error", "Incorrect dimensions", "Step size reduced to very small value",
    "Incorrect interpolation states", "Incorrect interpolation array provided", "Interpolation is not enabled",
    "Init is required", "Incorrect event-related parameters ", "Event root finding failed",
    "Incorrect step size for the non-adaptive option" };

extern "C" const char* flint_status_string(int s) { return (s >= 0 && s <= 16) ? kErrorStrings[s] : kErrorStrings[3]; }
extern "C" const char* flint_last_error(void) { return g_last_error.c_str(); }
extern "C" long flint_b200_launch_count(void) { return g_launches.load(); }
extern "C" int flint_b200_bounds_check(void) { return FLINT_BOUNDS_CHECK; }
extern "C" double flint_b200_last_kernel_ms(void) { return g_last_kernel_ms.load(); }
extern "C" int flint_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

/* ------------------------------------------------------------------ DiffEqSys */
extern "C" flint_sys_t* flint_sys_create(int system_id, int eventset_id, int n, int m, const double* sysparams, int nsysparams)
{
    int need_n = 0, max_m = 0;
    switch (system_id) {
    case FLINT_SYS_CR3BP_PLANAR: need_n = 6; max_m = 3; break;
    case FLINT_SYS_CR3BP_SPATIAL: need_n = 6; max_m = 0; break;
    case FLINT_SYS_LORENZ: need_n = 3; max_m = 0; break;
    case FLINT_SYS_TWOBODY: need_n = 6; max_m = 1; break;
    default:
        if (system_id >= FLINT_SYS_USER_BASE) {            /* registered at build time (csrc/systems.cuh) or at run time (rtc.cu) */
            int sys_n = 0, sys_m_max = 0, has_ev = 1;
            if (const flint::KernelEntry* ke = flint::find_system(system_id)) { sys_n = ke->sys_n; sys_m_max = ke->sys_m_max; }
            else if (!flint::rt_system_facts(system_id, &sys_n, &sys_m_max, &has_ev)) {
                g_last_error = "flint_sys_create: no user system with this id is compiled into the library or registered (flint_sys_register_source)";
                return nullptr;
            }
            if (n != sys_n || m < 0 || m > sys_m_max || m > FLINT_MAX_M || (eventset_id == FLINT_EVSET_NONE && m != 0) ||
                (m > 0 && !has_ev) || eventset_id < 0 || nsysparams < 0 || nsysparams > 7) {
                g_last_error = "flint_sys_create: n/m/eventset do not match the registered user functor";
                return nullptr;
            }
            flint_sys* s = new flint_sys();
            s->system_id = system_id; s->eventset_id = eventset_id; s->n = n; s->m = m;
            for (int i = 0; i < 8; ++i) s->sp[i] = (sysparams && i < nsysparams) ? sysparams[i] : 0.0;
            return s;
        }
        g_last_error = "unknown system_id"; return nullptr;
    }
    int ev_m = 0;
    switch (eventset_id) {
    case FLINT_EVSET_NONE: ev_m = 0; break;
    case FLINT_EVSET_CR3BP_SAMPLE3: case FLINT_EVSET_CR3BP_SAMPLE3_TERM: case FLINT_EVSET_CR3BP_SAMPLE3_RESTART:
        ev_m = (system_id == FLINT_SYS_CR3BP_PLANAR) ? 3 : -1; break;
    case FLINT_EVSET_XY_CROSSING: ev_m = (system_id == FLINT_SYS_CR3BP_PLANAR) ? 2 : -1; break;
    case FLINT_EVSET_APSIS: ev_m = (system_id == FLINT_SYS_TWOBODY) ? 1 : -1; break;
    default: ev_m = -1;
    }
    if (n != need_n || ev_m < 0 || m != ev_m || m > max_m || nsysparams < 0 || nsysparams > 8) {
        g_last_error = "flint_sys_create: n/m/eventset do not match the selected functor";
        return nullptr;
    }
    flint_sys* s = new flint_sys();
    s->system_id = system_id; s->eventset_id = eventset_id; s->n = n; s->m = m;
    for (int i = 0; i < 8; ++i) s->sp[i] = (sysparams && i < nsysparams) ? sysparams[i] : 0.0;
    return s;
}
extern "C" void flint_sys_destroy(flint_sys_t* s) { delete s; }

/* ------------------------------------------------------------------ ERK_class */
extern "C" flint_erk_t* flint_erk_create(void) { return new flint_erk(); }
extern "C" void flint_erk_destroy(flint_erk_t* e)
{
    if (!e) return;
    for (size_t k = 0; k < e->shard.size(); ++k) {
        if (e->shard_stream[k]) { cudaSetDevice(e->shard_device[k]); cudaStreamDestroy(e->shard_stream[k]); }
        flint_erk_destroy(e->shard[k]);
    }
    e->shard.clear();
    if (e->pinned) cudaFreeHost(e->pinned);
    e->dense.release();
    for (int k = 0; k < 2; ++k) { e->stage_int[k].release(); e->ws[k].release(); if (e->pipe_stream[k]) cudaStreamDestroy(e->pipe_stream[k]); e->pipe_stream[k] = nullptr; }
    delete e;
}

static void method_consts(int method, int& s, int& sint, int& p, int& q, int& pstar)
{
    switch (method) {   /* src/ButcherTableaus.f90:81-91,207-217,519-529,972-982 */
    case ERK_DOP54: s = 7; sint = 7; p = 5; q = 4; pstar = 4; break;
    case ERK_VERNER98R: s = 21; sint = 16; p = 9; q = 8; pstar = 8; break;
    case ERK_VERNER65E: s = 10; sint = 9; p = 6; q = 5; pstar = 5; break;
    default: s = 16; sint = 13; p = 8; q = 7; pstar = 7; break;
    }
}

extern "C" int flint_erk_init(flint_erk_t* me, const flint_sys_t* DE, int max_steps, const flint_init_opts* o)
{
    if (!me || !DE || !o) return fail(FLINT_ERROR_PARAMS, "flint_erk_init: null argument");
    me->status = FLINT_SUCCESS;                                          /* src/ERKInit.f90:57 */
    if (o->present & FLINT_INIT_METHOD) me->method = o->method;          /* :60 */
    if (me->method < ERK_DOP853 || me->method > ERK_VERNER65E) return me->status = FLINT_ERROR_PARAMS;
    me->sys_copy = *DE; me->sys = &me->sys_copy;                         /* :63 */
    const int n = DE->n, m = DE->m;
    if (max_steps <= 0) return me->status = FLINT_ERROR_PARAMS;          /* :68-73 */
    me->max_steps = max_steps;
    if (!o->atol || !o->rtol) return me->status = FLINT_ERROR_PARAMS;    /* quirk Q12: mandatory */
    if (o->n_atol == 1 && o->n_rtol == 1) {                              /* :76-87 */
        me->scalar_tol = true; me->atol[0] = o->atol[0]; me->rtol[0] = o->rtol[0];
    } else if (o->n_atol == n && o->n_rtol == n) {
        me->scalar_tol = false;
        for (int i = 0; i < n; ++i) { me->atol[i] = o->atol[i]; me->rtol[i] = o->rtol[i]; }
    } else return me->status = FLINT_ERROR_PARAMS;
    me->min_step = (o->present & FLINT_INIT_MIN_STEP) ? std::fabs(o->min_step_size) : 2.220446049250313e-16;   /* :92-96 */
    me->max_step = (o->present & FLINT_INIT_MAX_STEP) ? std::fabs(o->max_step_size) : 0.0;                     /* :100-104 */
    if (o->present & FLINT_INIT_STEPSZ_PARAMS) {                         /* :108-120 */
        for (int i = 0; i < 6; ++i) me->szp[i] = o->stepsz_params[i];
    } else {
        double beta = 0.0, bm = 0.0;
        if (me->method == ERK_DOP54) { beta = 0.04; bm = 0.75; }         /* src/ButcherTableaus.f90:95-96 */
        else if (me->method == ERK_DOP853) { beta = 0.0; bm = 0.2; }     /* :221-222 */
        me->szp[0] = 0.9; me->szp[1] = 0.333; me->szp[2] = 6.0;          /* src/StepSz.f90:35-44 */
        me->szp[3] = beta; me->szp[4] = bm; me->szp[5] = 1.0e-4;
    }
    me->dense.release(); me->dense_allocated = false;                    /* :123 */
    if (o->present & FLINT_INIT_INTERP_ON) me->interp_on = o->interp_on != 0;   /* :124-130 */
    me->events_on = false;                                               /* :135-136 */
    if (o->present & FLINT_INIT_EVENTS_ON) me->events_on = o->events_on != 0;
    if (me->events_on) {
        if (DE->eventset_id == FLINT_EVSET_NONE || m <= 0) me->status = FLINT_ERROR_EVENTPARAMS;   /* :141 */
        me->nI = n; for (int i = 0; i < n; ++i) me->istates[i] = i + 1;                            /* :144 */
        if (o->present & FLINT_INIT_EVENT_STEPSZ) {                      /* :147-158 */
            if (o->n_event_stepsz != m) return me->status = FLINT_ERROR_DIMENSION;
            for (int i = 0; i < m; ++i) me->ev_stepsz[i] = std::fabs(o->event_stepsz[i]);
        } else for (int i = 0; i < FLINT_MAX_M; ++i) me->ev_stepsz[i] = 0.0;
        if (o->present & FLINT_INIT_EVENT_OPTIONS) {                     /* :161-176 */
            if (o->n_event_options != m) return me->status = FLINT_ERROR_DIMENSION;
            for (int i = 0; i < m; ++i)
                if (o->event_options[i] < FLINT_EVENTOPTION_ROOTFINDING || o->event_options[i] > FLINT_EVENTOPTION_STEPEND)
                    return me->status = FLINT_ERROR_EVENTPARAMS;
            for (int i = 0; i < m; ++i) me->ev_opt[i] = o->event_options[i];
        } else for (int i = 0; i < FLINT_MAX_M; ++i) me->ev_opt[i] = FLINT_EVENTOPTION_ROOTFINDING;
        if (o->present & FLINT_INIT_EVENT_TOL) {                         /* :179-188 */
            if (o->n_event_tol != m) return me->status = FLINT_ERROR_DIMENSION;
            for (int i = 0; i < m; ++i) me->etol[i] = std::fabs(o->event_tol[i]);
        } else for (int i = 0; i < FLINT_MAX_M; ++i) me->etol[i] = 1.0e-6;   /* DEFAULT_EVENTTOL src/ERK.f90:50 */
    }
    method_consts(me->method, me->s, me->sint, me->p, me->q, me->pstar);    /* :198-274 */
    if (o->present & FLINT_INIT_ARITH) {
        if (o->arith != FLINT_ARITH_STRICT && o->arith != FLINT_ARITH_FMA) return me->status = FLINT_ERROR_PARAMS;
        me->arith = o->arith;
    }
    if (me->interp_on) {                                                 /* :291-335 */
        if ((o->present & FLINT_INIT_INTERP_STATES) && !me->events_on) {
            bool bad = o->n_interp_states < 1 || o->n_interp_states > n || !o->interp_states;
            for (int i = 0; !bad && i < o->n_interp_states; ++i)
                if (o->interp_states[i] < 1 || o->interp_states[i] > n) bad = true;
            if (bad) me->status = FLINT_ERROR_INTPSTATES;
            else { me->nI = o->n_interp_states; for (int i = 0; i < me->nI; ++i) me->istates[i] = o->interp_states[i]; me->status = FLINT_SUCCESS; }
        } else {
            me->nI = n; for (int i = 0; i < n; ++i) me->istates[i] = i + 1;
            me->status = FLINT_SUCCESS;
        }
        if (me->status != FLINT_SUCCESS) return me->status;
        me->dense_allocated = true;     /* IntpMemAllocate: device storage is sized at Integrate time */
    }
    me->inited = true;
    return me->status;
}

/* ------------------------------------------------------------------ partition of an ensemble over shards */
static long shard_count(long N, int D, int d, long chunk)
{
    if (N <= 0 || D < 1 || d < 0 || d >= D) return 0;
    if (chunk <= 0) chunk = FLINT_SHARD_CHUNK;
    const long nch = (N + chunk - 1) / chunk;               /* chunks c = d, d + D, ... < nch; only chunk nch-1 may be short */
    if (d >= nch) return 0;
    const long mine = (nch - 1 - d) / D + 1;
    long cnt = mine * chunk;
    if ((nch - 1) % D == d) cnt -= nch * chunk - N;
    return cnt;
}
static long shard_global(long N, int D, int d, long chunk, long i)
{
    if (chunk <= 0) chunk = FLINT_SHARD_CHUNK;
    if (i < 0 || i >= shard_count(N, D, d, chunk)) return -1;
    return ((i / chunk) * D + d) * chunk + i % chunk;
}
extern "C" long flint_b200_shard_count(long N, int n_shards, int shard, long chunk) { return shard_count(N, n_shards, shard, chunk); }
extern "C" long flint_b200_shard_global_index(long N, int n_shards, int shard, long chunk, long i) { return shard_global(N, n_shards, shard, chunk, i); }
extern "C" long flint_b200_shard_indices(long N, int n_shards, int shard, long chunk, long* out)
{
    const long cnt = shard_count(N, n_shards, shard, chunk);
    if (chunk <= 0) chunk = FLINT_SHARD_CHUNK;
    if (out) for (long i = 0; i < cnt; ++i) out[i] = ((i / chunk) * n_shards + shard) * chunk + i % chunk;
    return cnt;
}

/* ------------------------------------------------------------------ Integrate */
namespace {

/* Device staging buffers of the FLINT_MEM_HOST path, cached in the handle: the k-th allocation of a call reuses the
 * k-th buffer of the previous call when it is large enough (cudaMalloc/cudaFree of ~100 MB buffers cost milliseconds
 * and cudaFree synchronises the device). */
struct DevBuf {
    DevCache& c;
    size_t k = 0;
    DevBuf(DevCache& cache, int device) : c(cache) { if (c.device != device) { c.release(); c.device = device; } }
    template <class T> T* alloc(size_t count)
    {
        const size_t bytes = (count ? count : 1) * sizeof(T);
        if (k >= c.ptr.size()) { c.ptr.push_back(nullptr); c.size.push_back(0); }
        if (c.size[k] < bytes) {
            if (c.ptr[k]) cudaFree(c.ptr[k]);
            c.ptr[k] = nullptr; c.size[k] = 0;
            if (cudaMalloc(&c.ptr[k], bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            c.size[k] = bytes;
        }
        return (T*)c.ptr[k++];
    }
};

struct ExecCtx {
    int device = 0; cudaStream_t stream = nullptr; bool host = true; bool async = false;
    long dense_cap = 0; int block = 0, blocks_per_sm = 0, event_mode = 0, plane_variant = 0, pipeline = 0, round_len = 0, dense_mode = 0;
    int n_devices = 0; int devices[FLINT_MAX_DEVICES] = {0}; long shard_chunk = 0; int dense_spill = 0;
};

/* Which columns of the caller's SoA host arrays (row pitch Ntot) a call works on: shard d of D in the block-cyclic deal of
 * `chunk` trajectories (include/flint_b200.h: flint_b200_shard_*).  D == 1: all of them, in order. */
struct Shard {
    long Ntot = 0; int D = 1, d = 0; long chunk = FLINT_SHARD_CHUNK;
};

}  // namespace

/* rows x cnt elements between a compact device array [rows][cnt] and the shard's columns of a host array of row pitch Ntot:
 * columns [off, off+cnt) of an unsharded call; the shard's chunks (one strided 2-D copy per row, or one for all rows when the
 * rows line up with the deal) of a sharded one.  N = the call's local trajectory count. */
static bool shard_copy(const Shard& sh, long N, long off, long cnt, void* dst, const void* src, size_t elem, size_t rows,
                       bool to_device, const char* what, cudaStream_t s)
{
    const size_t szN = (size_t)cnt, szT = (size_t)sh.Ntot;
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    if (sh.D == 1) {
        if (cnt == N && (size_t)N == szT) return cuda_ok(cudaMemcpyAsync(dst, src, rows * szN * elem, kind, s), what);
        if (to_device)
            return cuda_ok(cudaMemcpy2DAsync(dst, szN * elem, (const char*)src + (size_t)off * elem, szT * elem, szN * elem, rows, kind, s), what);
        return cuda_ok(cudaMemcpy2DAsync((char*)dst + (size_t)off * elem, szT * elem, src, szN * elem, szN * elem, rows, kind, s), what);
    }
    /* sharded: local trajectory i = k * chunk + r is global ((k * D + d) * chunk + r); off == 0, cnt == N (no pipeline) */
    char* hbase = to_device ? (char*)src : (char*)dst;       /* host side */
    char* dbase = to_device ? (char*)dst : (char*)src;       /* device side, compact [rows][cnt] */
    const size_t ch = (size_t)sh.chunk, K = szN / ch, tail = szN % ch, stride = (size_t)sh.D * ch;
    bool ok = true;
    auto one = [&](char* h, size_t hpitch, char* d, size_t dpitch, size_t width, size_t height) {
        if (!ok || width == 0 || height == 0) return;
        if (height == 1 || hpitch > (size_t)1 << 30) {       /* plain copies: no pitch limit (a "column" may be a whole trajectory's Yarr) */
            for (size_t r = 0; r < height && ok; ++r)
                ok = to_device ? cuda_ok(cudaMemcpyAsync(d + r * dpitch, h + r * hpitch, width, kind, s), what)
                               : cuda_ok(cudaMemcpyAsync(h + r * hpitch, d + r * dpitch, width, kind, s), what);
            return;
        }
        ok = to_device ? cuda_ok(cudaMemcpy2DAsync(d, dpitch, h, hpitch, width, height, kind, s), what)
                       : cuda_ok(cudaMemcpy2DAsync(h, hpitch, d, dpitch, width, height, kind, s), what);
    };
    if (tail == 0 && szT % stride == 0 && szT / stride == K) {
        /* every row holds exactly K chunks of this shard, one deal period apart: rows * K lines of one copy */
        one(hbase + (size_t)sh.d * ch * elem, stride * elem, dbase, ch * elem, ch * elem, rows * K);
        return ok;
    }
    for (size_t r = 0; r < rows && ok; ++r) {
        char* hrow = hbase + (r * szT + (size_t)sh.d * ch) * elem;
        char* drow = dbase + r * szN * elem;
        one(hrow, stride * elem, drow, ch * elem, ch * elem, K);
        if (tail) one(hrow + K * stride * elem, tail * elem, drow + K * ch * elem, tail * elem, tail * elem, 1);
    }
    return ok;
}

/* erk_dense_kernel writes a block's coefficient records with one bulk copy from shared memory (default) or thread by thread
 * (FLINT_B200_DENSE_BULK=0; kept for comparison, profiles/r02_notes.md) */
static int dense_bulk_default()
{
    const char* e = getenv("FLINT_B200_DENSE_BULK");
    return e ? atoi(e) : 1;
}

/* what Init parsed, as the kernels' parameter block (everything that does not depend on one Integrate call) */
static void fill_static_kargs(const flint_erk* me, KArgs& a)
{
    const flint_sys& sys = *me->sys;
    a.n = sys.n; a.m = sys.m; a.evset = sys.eventset_id;
    for (int i = 0; i < 8; ++i) a.sp[i] = sys.sp[i];
    /* derived parameter, computed once here with the same IEEE operation the reference's F performs on every call
     * (`1.0_WP - mu`, tests/test_CR3BP.f90:178): the kernels read it from the parameter block */
    if (sys.system_id == FLINT_SYS_CR3BP_PLANAR || sys.system_id == FLINT_SYS_CR3BP_SPATIAL) a.sp[7] = 1.0 - a.sp[0];
    a.max_steps = me->max_steps;
    for (int i = 0; i < FLINT_MAX_N; ++i) { a.tol.a[i] = me->atol[i]; a.tol.r[i] = me->rtol[i]; }
    a.tol.scalar = me->scalar_tol ? 1 : 0;
    a.hmax_opt = me->max_step;
    for (int i = 0; i < 6; ++i) a.szp[i] = me->szp[i];
    for (int i = 0; i < FLINT_MAX_M; ++i) { a.ev_stepsz[i] = me->ev_stepsz[i]; a.ev_opt[i] = me->ev_opt[i]; a.etol[i] = me->etol[i]; }
    a.nI = me->nI; for (int i = 0; i < FLINT_MAX_N; ++i) a.istates[i] = me->istates[i];
}

/* record size of the dense store: step logs (2n+2 doubles; 3n+3 with events: the event code's steps carry h_stage and Y1), or
 * room for the coefficient record the log is turned into, + the kind, rounded to an even count (16-byte alignment) */
static int dense_rstride(const flint_erk* me, bool coef)
{
    const int n = me->sys->n;
    int r = me->events_on ? 3 * n + 3 : 2 * n + 2;
    if (coef && me->nI * (me->pstar + 1) + 2 > r) r = me->nI * (me->pstar + 1) + 2;
    r += 1;
    r += r & 1;
    if (coef) r += (r % 4 == 0) ? 2 : 0;      /* an odd number of 16-byte units: records of different steps start in different banks of
                                                 shared memory when a tile of them arrives as one bulk copy (interp_items_kernel) */
    return r;
}

static int integrate_impl(flint_erk* me, long N, double x0_scalar, const double* x0, const double* y0, double* xf,
                          double* yf, double* stepsz, const flint_int_opts* io, int* counters, int* status,
                          flint_steps_out* steps, flint_events_out* events, int* stifftest, const ExecCtx& ex,
                          bool scalar_call, const Shard* shard_in = nullptr)
{
    if (!me->inited || !me->sys) return me->status = FLINT_ERROR_INIT_REQD;
    g_last_error.clear();
    const flint_sys& sys = *me->sys;
    const int n = sys.n, m = sys.m;
    static const flint_int_opts kNoOpts{};
    if (!io) io = &kNoOpts;
    int st = FLINT_SUCCESS;
    const bool const_step = (io->present & FLINT_INT_CONST_STEPSZ) && io->use_const_stepsz;   /* :99-103 (sign check per trajectory, in-kernel) */
    const bool step_once = (io->present & FLINT_INT_STEP_ADVANCE) && io->step_advance;        /* :113-114 */
    bool int_steps = false;
    if (io->present & FLINT_INT_INTSTEPS_ON) {                                                /* :120-127 */
        if (io->int_steps_on && me->interp_on) st = FLINT_ERROR_PARAMS;
        else if (io->int_steps_on) int_steps = true;
    }
    if (int_steps && !(steps && steps->xint && steps->yint && steps->cap > 0)) st = FLINT_ERROR_PARAMS;   /* :129-140 */
    const bool stiff_present = stifftest != nullptr;                                          /* :143-154 (range check per trajectory) */
    if (me->interp_on && !me->dense_allocated) return me->status = FLINT_ERROR_INIT_REQD;    /* :157-161 */
    bool events_on = me->events_on;
    unsigned evmask = 0;
    if (events_on) {                                                                          /* :175-193 */
        for (int i = 0; i < m; ++i) {
            const bool on = (io->present & FLINT_INT_EVENT_MASK) ? (io->event_mask[i] != 0) : true;
            if (on) evmask |= (1u << i);
        }
        if (evmask == 0) events_on = false;
        else if (!events || !events->rec || !events->count) return me->status = FLINT_ERROR_EVENTPARAMS;
    }
    if ((io->present & FLINT_INT_PARAMS) && (io->n_params < 0 || io->n_params > 8 || (io->n_params > 0 && !io->params)))
        return me->status = fail(FLINT_ERROR_PARAMS, "params(:): at most 8 values are forwarded to the device functors");
    if (st != FLINT_SUCCESS) return me->status = st;                                          /* :196-199 */
    if (N <= 0) return me->status = FLINT_SUCCESS;
    Shard sh;
    if (shard_in) sh = *shard_in; else sh.Ntot = N;

    if (!cuda_ok(cudaSetDevice(ex.device), "cudaSetDevice")) return me->status = FLINT_ERROR;
    const bool dense = me->interp_on;
    const flint::KernelEntry* ke = flint::get_kernel(me->method, sys.system_id, me->arith == FLINT_ARITH_FMA, events_on);
    if (!ke && !g_last_error.empty()) return me->status = FLINT_ERROR;    /* a run-time system whose kernels failed to build: the compiler's log */
    if (!ke) {
        char buf[160];
        snprintf(buf, sizeof buf, "no kernel compiled for method=%d system=%d arith=%d events=%d dense=%d", me->method,
                 sys.system_id, me->arith, (int)events_on, (int)dense);
        return me->status = fail(FLINT_ERROR, buf);
    }

    KArgs a{};
    fill_static_kargs(me, a);
    a.N = N;
    a.const_step = const_step; a.step_once = step_once;
    a.stiff_present = stiff_present; a.stiff_mode = 0;
    a.events_on = events_on; a.evmask_bits = evmask;
    a.x0_scalar = x0_scalar;
    a.steps_on = int_steps;
    a.steps_cap = int_steps ? steps->cap : 0;
    a.ev_cap = events_on ? events->cap : 0;
    a.round_req = ex.round_len;
    if (io->present & FLINT_INT_PARAMS) { a.n_params = io->n_params; for (int i = 0; i < io->n_params; ++i) a.params[i] = io->params[i]; }

    flint::load_method_coefficients(me->method, a.coef);
    flint::det_pow_constants(a.powc);

    /* events that are located on (nearly) every step keep the in-loop handling; otherwise park + batch */
    bool park = events_on && ke->step_park.classic && ke->event;
    if (park)
        for (int i = 0; i < m; ++i)
            if (((evmask >> i) & 1u) && me->ev_stepsz[i] != 0.0) park = false;
    if (ex.event_mode == 1) park = false;
    if (ex.event_mode == 2 && events_on && ke->step_park.classic) park = true;

    /* persistent lanes: the step kernels size themselves (one resident wave of their own CTA size, erk_kernel.cuh: StepCfg) */
    int dev_sms = 0;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, ex.device);
    /* plane variant: trajectories with identically zero components (csrc/systems.cuh: SysCR3BPPlanarZ).  The dropped
     * components contribute (0/Sc)^2 to the error norms, so Sc = ATol must be positive there. */
    bool try_plane = ke->step_plane.classic != nullptr && (!events_on || park) && ex.plane_variant != 2;
    if (try_plane) {
        unsigned zp = sys.system_id == FLINT_SYS_CR3BP_PLANAR ? ((1u << 2) | (1u << 5)) : 0u;
        for (int c = 0; c < n; ++c)
            if (((zp >> c) & 1u) && !(me->atol[me->scalar_tol ? 0 : c] > 0.0)) try_plane = false;
    }

    /* ---- host buffers, flint_exec.pipeline = k >= 2: the call runs as k chunks of trajectories on two streams, so that the
     * H2D copy of chunk j+1 and the D2H copy of chunk j-1 travel while chunk j computes (no trajectory depends on another; a
     * chunk of the SoA arrays is a pitched 2-D copy).  Chunks use the fixed launch schedule of asynchronous calls: no
     * parked-count read-back, so nothing stalls the host between them.  Off by default: it pays when the copies are a large
     * part of the call (short integrations); for the north-star configuration (141 ms of compute against 6 ms of copies per
     * 2^20 orbits) every chunk adds the tails of its two step passes and the call gets slower -- measured, 2^20 orbits:
     * 151 ms (1 chunk), 152 (2), 162 (4), 182 (8); profiles/r01_notes.md. ---- */
    long n_chunks = 1;
    if (ex.host && !dense && !int_steps && !scalar_call && ex.pipeline != 1 && sh.D == 1) {
        n_chunks = ex.pipeline >= 2 ? ex.pipeline : 1;
        if (const char* e = getenv("FLINT_B200_PIPE_CHUNKS")) { const long v = atol(e); if (v >= 1) n_chunks = v; }
        if (n_chunks > N) n_chunks = N;
        if (n_chunks > 64) n_chunks = 64;
    }
    const bool piped = n_chunks > 1;
    /* Host buffers with parked events: the event records are append-only, and in the common case (every event masked after its
     * first trigger) they are complete after the first event kernel.  They are copied back on a second stream while the
     * resume pass computes; if any count changed afterwards they are copied again at the end (profiles/r01_notes.md). */
    const bool early_ev = ex.host && events_on && park && !piped && !ex.async && getenv("FLINT_B200_NO_EARLY_EVENTS") == nullptr;
    if (piped || early_ev) {
        if (me->pipe_device != ex.device)
            for (int k = 0; k < 2; ++k) { if (me->pipe_stream[k]) cudaStreamDestroy(me->pipe_stream[k]); me->pipe_stream[k] = nullptr; }
        for (int k = 0; k < 2; ++k)
            if (!me->pipe_stream[k] && !cuda_ok(cudaStreamCreateWithFlags(&me->pipe_stream[k], cudaStreamNonBlocking), "cudaStreamCreate"))
                return me->status = FLINT_ERROR;
        me->pipe_device = ex.device;
    }
    if (!me->pinned && !cuda_ok(cudaHostAlloc((void**)&me->pinned, 64, cudaHostAllocDefault), "cudaHostAlloc")) return me->status = FLINT_ERROR;

    /* one chunk [off, off + cnt) of the ensemble on stream s with the buffers of pipeline slot `slot` */
    auto run_chunk = [&](long off, long cnt, int slot, cudaStream_t s, bool fixed_schedule, cudaEvent_t e0, cudaEvent_t e1) -> int {
        KArgs c = a;
        c.N = cnt;
        const size_t szN = (size_t)cnt;
        bool early_done = false, reparked = false;  /* early copy of the event records issued / a trajectory parked again after it */
        cudaEvent_t ev_ready = nullptr;
        auto copy2d = [&](void* dst, const void* src, size_t elem, size_t rows, bool to_device, const char* what, cudaStream_t st) -> bool {
            return shard_copy(sh, N, off, cnt, dst, src, elem, rows, to_device, what, st);
        };
        /* ---- device views of the caller's buffers ---- */
        if (ex.host) {
            DevBuf scratch(me->stage_int[slot], ex.device);
            double* d_y0 = scratch.alloc<double>(szN * n);
            double* d_xf = scratch.alloc<double>(szN);
            double* d_yf = scratch.alloc<double>(szN * n);
            int* d_status = scratch.alloc<int>(szN);
            if (!d_y0 || !d_xf || !d_yf || !d_status) return fail(FLINT_ERROR_MEMALLOC, "device allocation failed");
            bool ok = copy2d(d_y0, y0, 8, n, true, "H2D y0", s) && copy2d(d_xf, xf, 8, 1, true, "H2D xf", s);
            c.y0 = d_y0; c.xf = d_xf; c.yf = d_yf; c.status = d_status;
            if (x0) { double* d = scratch.alloc<double>(szN); ok = ok && d && copy2d(d, x0, 8, 1, true, "H2D x0", s); c.x0 = d; }
            if (stepsz) { double* d = scratch.alloc<double>(szN); ok = ok && d && copy2d(d, stepsz, 8, 1, true, "H2D stepsz", s); c.stepsz = d; }
            if (counters) { c.counters = scratch.alloc<int>(szN * 4); ok = ok && c.counters; }
            if (stifftest) { int* d = scratch.alloc<int>(szN); ok = ok && d && copy2d(d, stifftest, 4, 1, true, "H2D stifftest", s); c.stifftest = d; }
            if (events_on) {
                c.ev_count = scratch.alloc<int>(szN);
                c.ev_rec = scratch.alloc<double>(szN * (size_t)(n + 2) * (size_t)c.ev_cap);
                ok = ok && c.ev_count && c.ev_rec;
                /* slots the run does not fill read back as NaN (all-ones pattern) */
                ok = ok && cuda_ok(cudaMemsetAsync(c.ev_rec, 0xFF, szN * (size_t)(n + 2) * (size_t)c.ev_cap * 8, s), "memset ev_rec");
            }
            if (int_steps) {
                c.xint = scratch.alloc<double>(szN * (size_t)c.steps_cap);
                c.yint = scratch.alloc<double>(szN * (size_t)c.steps_cap * n);
                c.steps_count = steps->count ? scratch.alloc<int>(szN) : nullptr;
                ok = ok && c.xint && c.yint;
            }
            if (!ok) return g_last_error.empty() ? fail(FLINT_ERROR_MEMALLOC, "device allocation failed") : FLINT_ERROR;
        } else {
            c.x0 = x0; c.y0 = y0; c.xf = xf; c.yf = yf; c.stepsz = stepsz; c.counters = counters; c.status = status;
            c.stifftest = stifftest;
            if (events_on) { c.ev_count = events->count; c.ev_rec = events->rec; }
            if (int_steps) { c.xint = steps->xint; c.yint = steps->yint; c.steps_count = steps->count; }
            if (!status) return fail(FLINT_ERROR_PARAMS, "status[] is required");
        }
        /* ---- dense storage owned by the object.  The reference sizes it for MaxSteps + 1 (src/ERKInit.f90:321-332); here level 0
         * holds cap0 steps per trajectory and the store grows by levels for the trajectories that need more (below): nothing is
         * ever dropped.  cap0: flint_exec.dense_cap, else what a third of the free device memory holds, at most MaxSteps. ---- */
        if (dense) {
            DenseStore& D = me->dense;
            if (D.spilled) D.release();                                /* a previous call's parts wait in host memory: gone with that call */
            const bool coef = ex.dense_mode == 1;
            const int rstride = dense_rstride(me, coef);
            long cap = ex.dense_cap > 0 ? ex.dense_cap : 0;
            if (cap <= 0) {
                size_t fr = 0, tot = 0;
                cudaMemGetInfo(&fr, &tot);
                fr += D.bytes();                                       /* the previous call's store is reused or released */
                cap = (long)(fr / 3 / (szN * (size_t)rstride * 8));
                if (cap > 4096) cap = 4096;
            }
            if (cap > me->max_steps) cap = me->max_steps;
            cap = (cap + 127) / 128 * 128;                             /* a CTA of the dense kernels never straddles two levels */
            if (cap < 128) cap = 128;
            if (D.device != ex.device || D.N != cnt || D.cap[0] != cap || D.nI != me->nI || D.pstar != me->pstar || D.rstride != rstride ||
                D.kind_coef != coef || !D.allocated()) {
                D.release();
                if (cudaMalloc(&D.base[0], szN * (size_t)cap * (size_t)rstride * 8) != cudaSuccess || cudaMalloc(&D.dcount, szN * 4) != cudaSuccess ||
                    cudaMalloc(&D.dzflag, 8) != cudaSuccess) {
                    cudaGetLastError(); D.release();
                    return fail(FLINT_ERROR_MEMALLOC, "dense-output storage does not fit in device memory; lower flint_exec.dense_cap");
                }
                D.device = ex.device; D.N = cnt; D.n = n; D.nI = me->nI; D.pstar = me->pstar; D.rstride = rstride; D.kind_coef = coef;
                D.cap[0] = cap; D.rows[0] = cnt;
            }
            for (int l = 1; l < FLINT_DENSE_LEVELS; ++l) {            /* levels a previous call added */
                if (D.base[l]) cudaFree(D.base[l]);
                if (D.slot[l]) cudaFree(D.slot[l]);
                D.base[l] = nullptr; D.slot[l] = nullptr; D.cap[l] = D.rows[l] = 0;
            }
            D.nlev = 1; D.events = me->events_on;
            c.dv = D.view(); c.dense_count = D.dcount;
            me->dense_is_scalar = scalar_call;
        }
        /* ---- trajectory state block, work lists and counters (erk_kernel.cuh) ---- */
        const size_t sd_b = szN * (size_t)ke->n_state_doubles * 8, si_b = szN * (size_t)ke->n_state_ints * 4, li_b = (szN + (szN & 1)) * 4;
        const int n_lists = dense ? 4 : 2;                     /* parked lists A, B; suspended-for-capacity lists C, D (InterpOn) */
        {
            const size_t need_b = sd_b + si_b + (size_t)n_lists * li_b + 72;
            Workspace& W = me->ws[slot];
            if (W.device != ex.device || W.bytes < need_b) {
                W.release();
                if (cudaMalloc(&W.ptr, need_b) != cudaSuccess) { cudaGetLastError(); return fail(FLINT_ERROR_MEMALLOC, "device allocation failed (trajectory state)"); }
                W.bytes = need_b; W.device = ex.device;
            }
            char* base = (char*)W.ptr;
            c.sd = (double*)base; c.si = (int*)(base + sd_b); c.scap = cnt;
        }
        int* listA = (int*)((char*)me->ws[slot].ptr + sd_b + si_b);
        int* listB = listA + li_b / 4;
        int* listCD[2] = {listB + li_b / 4, listB + 2 * (li_b / 4)};
        unsigned long long* ctr = (unsigned long long*)((char*)me->ws[slot].ptr + (sd_b + si_b + (size_t)n_lists * li_b + 7) / 8 * 8);

        /* [0] work counter (every item is handed out through it), [1] |A|, [2] |B|, [3] low word: the plane flag, cleared by
         * erk_init_kernel, [4] |C|, [5] |D|.  Set on the device: a pageable H2D copy would make the host wait for everything
         * queued on this stream */
        if (!cuda_ok(flint::launch_set_counters(ctr, 0ULL, 0ULL, 0ULL, 1ULL, s), "work counters") ||
            !cuda_ok(cudaMemsetAsync(ctr + 4, 0, 16, s), "work counters")) return FLINT_ERROR;
        g_launches += 1;
        c.work_counter = ctr;
        c.zflag = try_plane ? (int*)(ctr + 3) : nullptr;
        const int grid_all = (int)((cnt + 127) / 128);
        /* InterpOn: a trajectory whose levels of the dense store are full is suspended on list C / D; synchronous calls add a
         * level and resume it (below).  Calls that never read a count back cannot: they keep what fits (Interpolate then
         * reports FLINT_ERROR_INTP_ARRAY for those trajectories, the reference's answer to a full Xint) */
        const bool can_grow = dense && !fixed_schedule;
        int ovf = 0;                                           /* index of the list being filled */
        if (dense) { c.dense_avail = can_grow ? c.dv.capacity() : (1L << 62); c.ovf_list = listCD[ovf]; c.n_ovf = ctr + 4 + ovf; }

        /* one sweep over a set of trajectories -- all cnt of them (in == nullptr) or a list -- to their end, to a parked event
         * (located in rounds of {event kernel, resume}) or to a full dense store */
        auto sweep = [&](const int* in, const unsigned long long* n_in, long n_items) -> cudaError_t {
            cudaError_t le = cudaSuccess;
            c.n_items_exact = 1;                               /* the host knows how many trajectories this pass runs (gang-scheduled kernel) */
            if (!park) {
                c.list_in = in; c.n_in = n_in; c.list_out = nullptr; c.n_out = nullptr;
                if (try_plane) { le = flint::launch_step(ke->step_plane, c, dev_sms, n_items, ex.block, ex.blocks_per_sm, s); g_launches += 1; }   /* returns at once unless the flag is set */
                if (le == cudaSuccess) { le = flint::launch_step(ke->step, c, dev_sms, n_items, ex.block, ex.blocks_per_sm, s); g_launches += 1; }
                return le;
            }
            /* pass 0; then rounds of {locate the parked events, resume}.  Synchronous calls read the parked count and stop when
             * it is zero; the fixed schedule runs a set number of rounds (empty kernels return at once).  The last round
             * resumes with in-loop event handling, so it always finishes. */
            const int max_rounds = fixed_schedule ? 3 : 64;
            int cur = 1;                                               /* index of the counter of the list being filled */
            c.list_in = in; c.n_in = n_in; c.list_out = listA; c.n_out = ctr + 1;
            if (try_plane) { le = flint::launch_step(ke->step_plane, c, dev_sms, n_items, ex.block, ex.blocks_per_sm, s); g_launches += 1; }
            if (le == cudaSuccess) { le = flint::launch_step(ke->step_park, c, dev_sms, n_items, ex.block, ex.blocks_per_sm, s); g_launches += 1; }
            for (int round = 0; round < max_rounds && le == cudaSuccess; ++round) {
                long n_parked = n_items;
                if (!fixed_schedule) {
                    /* the parked count through a page-locked word: a pageable destination would be staged by the driver */
                    if (!cuda_ok(cudaMemcpyAsync(me->pinned, ctr + cur, 8, cudaMemcpyDeviceToHost, s), "D2H parked count") ||
                        !cuda_ok(cudaStreamSynchronize(s), "kernel execution")) return cudaErrorUnknown;
                    n_parked = (long)me->pinned[0];
                    if (n_parked == 0) break;
                    if (round >= 1 || early_done) reparked = true;     /* events located after the early copy of the records */
                }
                int* lcur = cur == 1 ? listA : listB;
                int* loth = cur == 1 ? listB : listA;
                const int oth = cur == 1 ? 2 : 1;
                c.list_in = lcur; c.n_in = ctr + cur; c.list_out = loth; c.n_out = ctr + oth;
                c.n_items_exact = fixed_schedule ? 0 : 1;      /* synchronous calls know every pass's item count */
                le = flint::launch_kernel(ke->event, c, dim3((unsigned)((n_parked + 127) / 128)), 128, 0, s);   /* also re-arms work_counter and zeroes |other| */
                g_launches += 1;
                if (le != cudaSuccess) break;
                bool issue_early = false;
                if (early_ev && round == 0 && !early_done) {   /* stream order: the marker sits between the event kernel and the resume pass */
                    const bool ok = cuda_ok(cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming), "cudaEventCreate") &&
                                    cuda_ok(cudaEventRecord(ev_ready, s), "cudaEventRecord");
                    if (!ok) return cudaErrorUnknown;
                    issue_early = true;
                }
                /* the event kernel re-armed the work counter; a pass never launches more lanes than it has items */
                const bool last = round == max_rounds - 1;
                if (last) { int* zf = c.zflag; c.zflag = nullptr; le = flint::launch_step(ke->step, c, dev_sms, n_parked, ex.block, ex.blocks_per_sm, s); g_launches += 1; c.zflag = zf; }   /* in-loop events, general kernel */
                else {
                    if (try_plane) { le = flint::launch_step(ke->step_plane, c, dev_sms, n_parked, ex.block, ex.blocks_per_sm, s); g_launches += 1; }
                    if (le == cudaSuccess) { le = flint::launch_step(ke->step_park, c, dev_sms, n_parked, ex.block, ex.blocks_per_sm, s); g_launches += 1; }
                }
                if (issue_early && le == cudaSuccess) {        /* issued after the launch: a pageable destination blocks the host, not the GPU */
                    cudaStream_t s2 = me->pipe_stream[1];
                    const bool ok = cuda_ok(cudaStreamWaitEvent(s2, ev_ready, 0), "cudaStreamWaitEvent") &&
                                    copy2d(events->rec, c.ev_rec, 8, (size_t)(n + 2) * (size_t)c.ev_cap, false, "early D2H ev_rec", s2);
                    if (!ok) return cudaErrorUnknown;
                    early_done = true;
                }
                cur = oth;
            }
            return le;
        };

        if (e0) cudaEventRecord(e0, s);
        cudaError_t le = flint::launch_kernel(ke->init, c, dim3((unsigned)grid_all), 128, 0, s);
        g_launches += 1;
        if (le == cudaSuccess) le = sweep(nullptr, nullptr, cnt);
        /* ---- InterpOn: trajectories suspended with a full dense store get a new level, 4x as deep, and are resumed ---- */
        while (le == cudaSuccess && can_grow) {
            if (!cuda_ok(cudaMemcpyAsync(me->pinned, ctr + 4 + ovf, 8, cudaMemcpyDeviceToHost, s), "D2H suspended count") ||
                !cuda_ok(cudaStreamSynchronize(s), "kernel execution")) { le = cudaErrorUnknown; break; }
            const long n_susp = (long)me->pinned[0];
            if (n_susp == 0) break;
            DenseStore& D = me->dense;
            if (D.nlev >= FLINT_DENSE_LEVELS) { g_last_error = "dense-output storage: level limit reached"; le = cudaErrorUnknown; break; }
            const int l = D.nlev;
            long capl = 4 * D.cap[l - 1];
            const long held = c.dv.capacity();
            if (held + capl > me->max_steps + 127) capl = (me->max_steps - held + 127) / 128 * 128;
            if (capl < 128) capl = 128;
            if (cudaMalloc(&D.base[l], (size_t)n_susp * (size_t)capl * (size_t)D.rstride * 8) != cudaSuccess ||
                cudaMalloc(&D.slot[l], szN * 4) != cudaSuccess) {
                cudaGetLastError();
                if (ev_ready) cudaEventDestroy(ev_ready);
                return fail(FLINT_ERROR_MEMALLOC, "dense-output storage does not fit in device memory");
            }
            D.cap[l] = capl; D.rows[l] = n_susp; D.nlev = l + 1;
            if (!cuda_ok(cudaMemsetAsync(D.slot[l], 0xFF, szN * 4, s), "dense slots") ||
                !cuda_ok(flint::launch_dense_slots(D.slot[l], listCD[ovf], n_susp, s), "dense slots")) { le = cudaErrorUnknown; break; }
            g_launches += 1;
            c.dv = D.view(); c.dense_avail = c.dv.capacity();
            const int* in = listCD[ovf];
            const unsigned long long* n_in = ctr + 4 + ovf;
            ovf ^= 1;
            c.ovf_list = listCD[ovf]; c.n_ovf = ctr + 4 + ovf;
            /* re-arm: work counter, both parked counts, the other suspended count (the plane flag keeps its value) */
            if (!cuda_ok(cudaMemsetAsync(ctr, 0, 24, s), "work counters") || !cuda_ok(cudaMemsetAsync(ctr + 4 + ovf, 0, 8, s), "work counters")) { le = cudaErrorUnknown; break; }
            le = sweep(in, n_in, n_susp);
        }
        if (le == cudaSuccess && dense && me->dense.kind_coef) {   /* coefficient records from the logged steps (erk_kernel.cuh: erk_dense_kernel) */
            c.dvo = c.dv; c.dense_eval = 0; c.dense_bulk = dense_bulk_default();
            int* zf = c.zflag;
            c.zflag = (try_plane && ke->dense_plane) ? (int*)(ctr + 3) : nullptr;
            if (c.zflag) { le = flint::launch_dense(*ke, true, c, s); g_launches += 1; }
            if (le == cudaSuccess) { le = flint::launch_dense(*ke, false, c, s); g_launches += 1; }
            c.zflag = zf;
        }
        if (dense) {                                           /* Interpolate runs the dense kernels again (fused path): keep the plane flag */
            me->dense_plane_ok = try_plane && ke->dense_plane;
            if (le == cudaSuccess && me->dense_plane_ok && !cuda_ok(cudaMemcpyAsync(me->dense.dzflag, ctr + 3, 4, cudaMemcpyDeviceToDevice, s), "plane flag")) le = cudaErrorUnknown;
        }
        if (e1) cudaEventRecord(e1, s);
        if (le != cudaSuccess) {                               /* cudaErrorUnknown: a step above failed and left its own message */
            if (le != cudaErrorUnknown || g_last_error.empty()) cuda_ok(le, "kernel launch");
            if (ev_ready) cudaEventDestroy(ev_ready);
            return FLINT_ERROR;
        }

        if (ex.host) {
            bool ok = copy2d(xf, c.xf, 8, 1, false, "D2H xf", s) && copy2d(yf, c.yf, 8, n, false, "D2H yf", s);
            if (status) ok = ok && copy2d(status, c.status, 4, 1, false, "D2H status", s);
            if (stepsz && !const_step) ok = ok && copy2d(stepsz, c.stepsz, 8, 1, false, "D2H stepsz", s);
            if (counters) ok = ok && copy2d(counters, c.counters, 4, 4, false, "D2H counters", s);
            if (stifftest) ok = ok && copy2d(stifftest, c.stifftest, 4, 1, false, "D2H stifftest", s);
            /* the counts are final only now (a trajectory writes its count when it finishes); the records went out during
             * the resume pass and are complete unless a trajectory parked again after that copy */
            if (events_on) ok = ok && copy2d(events->count, c.ev_count, 4, 1, false, "D2H ev_count", s);
            if (events_on && !early_done) ok = ok && copy2d(events->rec, c.ev_rec, 8, (size_t)(n + 2) * (size_t)c.ev_cap, false, "D2H ev_rec", s);
            if (int_steps) {
                ok = ok && copy2d(steps->xint, c.xint, 8, (size_t)c.steps_cap, false, "D2H xint", s) &&
                     copy2d(steps->yint, c.yint, 8, (size_t)c.steps_cap * n, false, "D2H yint", s);
                if (steps->count) ok = ok && copy2d(steps->count, c.steps_count, 4, 1, false, "D2H steps_count", s);
            }
            if (ok && early_done) {
                ok = cuda_ok(cudaStreamSynchronize(me->pipe_stream[1]), "early event copy");     /* never two copies into the same host lines */
                if (ok && reparked) ok = copy2d(events->rec, c.ev_rec, 8, (size_t)(n + 2) * (size_t)c.ev_cap, false, "D2H ev_rec", s);
            }
            if (ev_ready) cudaEventDestroy(ev_ready);
            if (!ok) return FLINT_ERROR;
        } else if (ev_ready) cudaEventDestroy(ev_ready);
        return FLINT_SUCCESS;
    };

    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool timed = !ex.async;
    if (timed) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
    auto drop_events = [&]() { if (timed) { cudaEventDestroy(e0); cudaEventDestroy(e1); } };
    if (!piped) {
        const int rc = run_chunk(0, N, 0, ex.stream, ex.async, e0, e1);
        if (rc != FLINT_SUCCESS) { drop_events(); return me->status = rc; }
        if (!ex.async || ex.host) {
            if (!cuda_ok(cudaStreamSynchronize(ex.stream), "kernel execution")) { drop_events(); return me->status = FLINT_ERROR; }
        }
    } else {
        /* chunk boundaries on multiples of 128 trajectories; the caller's stream is joined on both sides */
        cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
        cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
        cudaEventRecord(fork, ex.stream);
        for (int k = 0; k < 2; ++k) { cudaStreamWaitEvent(me->pipe_stream[k], fork, 0); cudaEventCreateWithFlags(&join[k], cudaEventDisableTiming); }
        long per = ((N + n_chunks - 1) / n_chunks + 127) / 128 * 128;
        int rc = FLINT_SUCCESS, k = 0;
        for (long off = 0; off < N && rc == FLINT_SUCCESS; off += per, ++k) {
            const long cnt = off + per <= N ? per : N - off;
            const bool first = off == 0, last = off + per >= N;
            rc = run_chunk(off, cnt, k & 1, me->pipe_stream[k & 1], true, first ? e0 : nullptr, nullptr);
            if (last && rc == FLINT_SUCCESS && timed) {          /* the pipeline's span: first chunk's first kernel .. everything copied back */
                cudaEventRecord(join[(k & 1) ^ 1], me->pipe_stream[(k & 1) ^ 1]);
                cudaStreamWaitEvent(me->pipe_stream[k & 1], join[(k & 1) ^ 1], 0);
                cudaEventRecord(e1, me->pipe_stream[k & 1]);
            }
        }
        bool ok = true;
        for (int q = 0; q < 2; ++q) ok = cuda_ok(cudaStreamSynchronize(me->pipe_stream[q]), "kernel execution") && ok;
        cudaEventDestroy(fork); cudaEventDestroy(join[0]); cudaEventDestroy(join[1]);
        if (rc != FLINT_SUCCESS || !ok) { drop_events(); return me->status = (rc != FLINT_SUCCESS ? rc : FLINT_ERROR); }
    }
    if (timed) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        me->last_kernel_ms = ms;
        g_last_kernel_ms = ms;
        drop_events();
    }
    return me->status = FLINT_SUCCESS;
}

static ExecCtx make_exec(const flint_exec* x)
{
    ExecCtx e;
    if (x) {
        e.device = x->device; e.stream = (cudaStream_t)x->stream; e.host = (x->mem == FLINT_MEM_HOST);
        e.async = x->async != 0 && !e.host; e.dense_cap = x->dense_cap; e.block = x->block_threads; e.blocks_per_sm = x->blocks_per_sm;
        e.event_mode = x->event_mode; e.plane_variant = x->plane_variant; e.pipeline = x->pipeline; e.round_len = x->round_len; e.dense_mode = x->dense_mode;
        e.n_devices = x->n_devices; e.shard_chunk = x->shard_chunk > 0 ? x->shard_chunk : FLINT_SHARD_CHUNK; e.dense_spill = x->dense_spill;
        for (int k = 0; k < FLINT_MAX_DEVICES; ++k) e.devices[k] = x->devices[k];
    }
    return e;
}

/* Init's results copied into the sub-handle of one device of a sharded call */
static void sync_options(flint_erk* dst, const flint_erk* src)
{
    dst->status = src->status; dst->inited = src->inited;
    dst->sys_copy = src->sys_copy; dst->sys = src->sys ? &dst->sys_copy : nullptr;
    dst->method = src->method; dst->max_steps = src->max_steps; dst->scalar_tol = src->scalar_tol;
    for (int i = 0; i < FLINT_MAX_N; ++i) { dst->atol[i] = src->atol[i]; dst->rtol[i] = src->rtol[i]; dst->istates[i] = src->istates[i]; }
    dst->interp_on = src->interp_on; dst->nI = src->nI; dst->min_step = src->min_step; dst->max_step = src->max_step;
    for (int i = 0; i < 6; ++i) dst->szp[i] = src->szp[i];
    dst->events_on = src->events_on;
    for (int i = 0; i < FLINT_MAX_M; ++i) { dst->ev_stepsz[i] = src->ev_stepsz[i]; dst->ev_opt[i] = src->ev_opt[i]; dst->etol[i] = src->etol[i]; }
    dst->arith = src->arith; dst->s = src->s; dst->sint = src->sint; dst->p = src->p; dst->q = src->q; dst->pstar = src->pstar;
    dst->dense_allocated = src->dense_allocated;
}

/* The sub-handles and streams of a sharded call, created on first use and kept (staging buffers, workspaces, dense stores).
 * S shards over the D device slots of `devs`: shard s runs on devs[s % D]. */
static int prepare_shards(flint_erk* me, const int* devs, int D, int S)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); ndev = 0; }
    for (int k = 0; k < D; ++k) {
        /* a device may be listed more than once: its shards then share the GPU (two streams); the tests of a one-GPU box do that */
        if (devs[k] < 0 || devs[k] >= ndev) return fail(FLINT_ERROR_PARAMS, "flint_exec.devices: no such CUDA device");
    }
    bool same = (int)me->shard.size() == S;
    for (int k = 0; same && k < S; ++k) same = me->shard_device[k] == devs[k % D];
    if (!same) {
        for (size_t k = 0; k < me->shard.size(); ++k) {
            if (me->shard_stream[k]) { cudaSetDevice(me->shard_device[k]); cudaStreamDestroy(me->shard_stream[k]); }
            flint_erk_destroy(me->shard[k]);
        }
        me->shard.clear(); me->shard_stream.clear(); me->shard_device.clear();
        for (int k = 0; k < S; ++k) {
            cudaStream_t st = nullptr;
            if (!cuda_ok(cudaSetDevice(devs[k % D]), "cudaSetDevice") ||
                !cuda_ok(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate")) return FLINT_ERROR;
            me->shard.push_back(new flint_erk()); me->shard_stream.push_back(st); me->shard_device.push_back(devs[k % D]);
        }
    }
    for (int k = 0; k < S; ++k) sync_options(me->shard[k], me);
    return FLINT_SUCCESS;
}

/* flint_exec.dense_spill = 1: parts per device slot such that a part's dense store (first level at the depth an unspilled call
 * would choose, at most 1024 steps; deeper levels only hold the trajectories that need them) stays within a third of the free
 * memory of the emptiest device */
static int auto_spill_parts(const flint_erk* me, long N, const int* devs, int D, const ExecCtx& ex)
{
    size_t fr_min = ~(size_t)0;
    for (int k = 0; k < D; ++k) {
        size_t fr = 0, tot = 0;
        if (cudaSetDevice(devs[k]) != cudaSuccess || cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); return 1; }
        if (fr < fr_min) fr_min = fr;
    }
    long cap = ex.dense_cap > 0 ? ex.dense_cap : 1024;
    if (cap > me->max_steps) cap = me->max_steps;
    cap = (cap + 127) / 128 * 128;
    const double per_traj = (double)dense_rstride(me, ex.dense_mode == 1) * 8.0 * (double)cap;
    const double need = per_traj * (double)((N + D - 1) / D);
    if (const char* e = getenv("FLINT_B200_SPILL_BUDGET")) { const double b = atof(e); if (b > 0) fr_min = (size_t)b * 3; }   /* tests: bytes a part may use */
    long parts = (long)(need / ((double)fr_min / 3.0)) + 1;
    if (parts > 4096) parts = 4096;
    return (int)parts;
}

extern "C" int flint_erk_integrate_batch(flint_erk_t* me, long N, double x0_scalar, const double* x0, const double* y0,
                                         double* xf, double* yf, double* stepsz, const flint_int_opts* io, int* counters,
                                         int* status, flint_steps_out* steps, flint_events_out* events, int* stifftest,
                                         const flint_exec* exec)
{
    if (!me) return FLINT_ERROR_PARAMS;
    if (!y0 || !xf || !yf) return me->status = fail(FLINT_ERROR_PARAMS, "y0, xf and yf are required");
    const ExecCtx ex = make_exec(exec);
    if (ex.n_devices > FLINT_MAX_DEVICES) return me->status = fail(FLINT_ERROR_PARAMS, "flint_exec.n_devices > FLINT_MAX_DEVICES");
    int devs[FLINT_MAX_DEVICES];
    const int D = ex.n_devices >= 1 ? ex.n_devices : 1;
    for (int k = 0; k < D; ++k) devs[k] = ex.n_devices >= 1 ? ex.devices[k] : ex.device;
    /* ---- InterpOn with a dense output larger than device memory: parts whose stores wait in host memory (dense_spill) ---- */
    int parts = 1;
    if (ex.dense_spill != 0 && ex.host && me->inited && me->sys && me->interp_on && N > 0) {
        parts = ex.dense_spill >= 2 ? ex.dense_spill : auto_spill_parts(me, N, devs, D, ex);
        const long max_parts = (N / D + 127) / 128;          /* a part is at least 128 trajectories */
        if (parts > max_parts) parts = (int)(max_parts < 1 ? 1 : max_parts);
    }
    const int S = D * parts;
    if (S <= 1) {
        ExecCtx e1 = ex;
        e1.device = devs[0];
        me->shard_N = 0; me->shard_spill = false;
        return integrate_impl(me, N, x0_scalar, x0, y0, xf, yf, stepsz, io, counters, status, steps, events, stifftest, e1, false);
    }
    /* ---- the ensemble sharded over the devices of one box: chunks of shard_chunk trajectories dealt round-robin;
     * one host thread, one stream, one set of staging buffers per device; every trajectory is independent, so there is no
     * exchange step and nothing to reduce -- results are copied to the caller's arrays at the trajectories' own offsets.
     * With dense_spill every device slot has `parts` shards and runs them one after the other ---- */
    if (!ex.host) return me->status = fail(FLINT_ERROR_PARAMS, "a sharded call takes host buffers (FLINT_MEM_HOST)");
    if (!me->inited || !me->sys) return me->status = FLINT_ERROR_INIT_REQD;
    if (N <= 0) return me->status = FLINT_SUCCESS;
    { const int rc = prepare_shards(me, devs, D, S); if (rc != FLINT_SUCCESS) return me->status = rc; }
    long chunk = ex.shard_chunk;
    if (parts > 1 && chunk * S > N) { chunk = (N / S + 127) / 128 * 128; if (chunk < 128) chunk = 128; }   /* every part gets trajectories */
    std::vector<int> rc(S, FLINT_SUCCESS);
    std::vector<std::string> err(S);
    std::vector<std::thread> th;
    const bool spill = parts > 1;
    for (int d = 0; d < D; ++d) {
        th.emplace_back([&, d]() {
            for (int k = d; k < S; k += D) {
                Shard sh; sh.Ntot = N; sh.D = S; sh.d = k; sh.chunk = chunk;
                ExecCtx e1 = ex;
                e1.device = devs[d]; e1.stream = me->shard_stream[k]; e1.n_devices = 0; e1.pipeline = 1;
                const long nk = shard_count(N, S, k, sh.chunk);
                flint_erk* sub = me->shard[k];
                rc[k] = integrate_impl(sub, nk, x0_scalar, x0, y0, xf, yf, stepsz, io, counters, status, steps, events, stifftest,
                                       e1, false, &sh);
                if (rc[k] == FLINT_SUCCESS && spill && nk > 0) {
                    /* the part's dense output to host memory; its staging buffers and workspace make room for the next part */
                    if (!sub->dense.spill(e1.stream)) rc[k] = fail(FLINT_ERROR_MEMALLOC, "dense_spill: no host memory for a part's dense output");
                    for (int q = 0; q < 2; ++q) { sub->stage_int[q].release(); sub->ws[q].release(); }
                }
                if (rc[k] != FLINT_SUCCESS) { err[k] = g_last_error; break; }
            }
        });
    }
    for (auto& t : th) t.join();
    double ms = 0.0;
    int st = FLINT_SUCCESS;
    for (int d = 0; d < D; ++d) {                              /* the call's device time: the slowest device slot (its parts add up) */
        double md = 0.0;
        for (int k = d; k < S; k += D) md += me->shard[k]->last_kernel_ms;
        if (md > ms) ms = md;
    }
    for (int k = 0; k < S; ++k)
        if (rc[k] != FLINT_SUCCESS && st == FLINT_SUCCESS) { st = rc[k]; g_last_error = err[k]; }
    me->last_kernel_ms = ms; g_last_kernel_ms = ms;
    me->shard_N = N; me->shard_chunk = chunk; me->shard_devices = D; me->shard_spill = spill;
    me->dense.release();                                       /* the dense output of this call lives in the shards */
    return me->status = st;
}

extern "C" int flint_erk_integrate(flint_erk_t* me, double x0, const double* y0, double* xf, double* yf, double* stepsz,
                                   const flint_int_opts* io, flint_steps_out* steps, flint_events_out* events, int* stifftest)
{
    if (!me) return FLINT_ERROR_PARAMS;
    if (!y0 || !xf || !yf || !stepsz) return me->status = fail(FLINT_ERROR_PARAMS, "y0, xf, yf and stepsz are required");
    if (!me->sys) return me->status = FLINT_ERROR_INIT_REQD;
    const int n = me->sys->n;
    int counters[4] = {0, 0, 0, 0};
    int st1 = FLINT_SUCCESS;
    /* N = 1: SoA [c][j][1] is the transpose of the Fortran-shaped [j][c] outputs */
    std::vector<double> yint_t, ev_t;
    flint_steps_out s2{}; flint_events_out e2{};
    int scount = 0, ecount = 0;
    if (steps) { yint_t.resize((size_t)steps->cap * n); s2.cap = steps->cap; s2.xint = steps->xint; s2.yint = yint_t.data(); s2.count = &scount; }
    if (events) { ev_t.resize((size_t)events->cap * (n + 2)); e2.cap = events->cap; e2.rec = ev_t.data(); e2.count = &ecount; }
    ExecCtx ex;   /* device 0, host memory, synchronous */
    const double x0s = x0, xf_in = *xf;
    (void)xf_in;
    int rc = integrate_impl(me, 1, x0, nullptr, y0, xf, yf, stepsz, io, counters, &st1, steps ? &s2 : nullptr,
                            events ? &e2 : nullptr, stifftest, ex, true);
    if (rc != FLINT_SUCCESS) return rc;   /* argument error or CUDA failure: me%status already set */
    if (st1 == FLINT_ERROR_CONSTSTEPSZ || (st1 == FLINT_ERROR_PARAMS && counters[0] == 0)) return me->status = st1;
    if (steps) {
        const long cnt = scount < steps->cap ? scount : steps->cap;
        for (long j = 0; j < cnt; ++j) for (int c = 0; c < n; ++c) steps->yint[j * n + c] = yint_t[(size_t)c * steps->cap + j];
        if (steps->count) *steps->count = scount;
    }
    if (events) {
        const long cnt = ecount < events->cap ? ecount : events->cap;
        for (long j = 0; j < cnt; ++j) for (int c = 0; c < n + 2; ++c) events->rec[j * (n + 2) + c] = ev_t[(size_t)c * events->cap + j];
        if (events->count) *events->count = ecount;
    }
    me->total = counters[0]; me->accepted = counters[1]; me->rejected = counters[2]; me->fcalls = counters[3];
    me->X0 = x0s; for (int c = 0; c < n; ++c) { me->Y0[c] = y0[c]; me->Yf[c] = yf[c]; }
    me->Xf = *xf; me->hf = *stepsz;
    return me->status = st1;
}

/* ------------------------------------------------------------------ Interpolate */
static int interpolate_impl(flint_erk* me, const double* xarr, long G, double* yarr, int layout, int* status,
                            bool has_save, bool save, const ExecCtx& ex, bool scalar_call, const Shard* shard_in = nullptr,
                            bool per_traj = false)
{
    if (!me->interp_on) return me->status = FLINT_ERROR_INTP_OFF;                       /* src/ERKInterp.f90:45-48 */
    const bool dealloc = has_save ? !save : true;                                        /* :52-53 */
    if (G < 1) return me->status;                                                        /* :56-57 */
    DenseStore& D = me->dense;
    if (!D.allocated() || D.N <= 0) return me->status = FLINT_ERROR_INTP_OFF;            /* storage already freed */
    if (scalar_call && !(me->dense_is_scalar && D.N == 1))     /* the scalar Interpolate reads the dense output of a scalar Integrate */
        return fail(FLINT_ERROR_INTP_OFF, "flint_erk_interpolate: the stored dense output belongs to a batch Integrate; use flint_erk_interpolate_batch");   /* the object's status is untouched */
    if (!xarr || !yarr) return me->status = fail(FLINT_ERROR_PARAMS, "xarr and yarr are required");
    if (!cuda_ok(cudaSetDevice(D.device), "cudaSetDevice")) return me->status = FLINT_ERROR;
    const long N = D.N;
    Shard sh;
    if (shard_in) sh = *shard_in; else sh.Ntot = N;
    cudaStream_t s = ex.stream;
    ScopedCache scoped;                      /* Yarr can be most of the device memory: not kept between calls */
    DevBuf scratch(scoped, D.device);
    const size_t nx = per_traj ? (size_t)N * (size_t)G : (size_t)G;
    /* monotonicity of a shared grid is a host-side check (:60-68); per-trajectory grids are checked on the device */
    int up = 1, down = 1;
    if (!per_traj) {
        std::vector<double> hx;
        const double* xh = xarr;
        if (!ex.host) { hx.resize((size_t)G); if (!cuda_ok(cudaMemcpy(hx.data(), xarr, (size_t)G * 8, cudaMemcpyDeviceToHost), "D2H xarr")) return me->status = FLINT_ERROR; xh = hx.data(); }
        for (long i = 0; i + 1 < G; ++i) { if (!(xh[i] < xh[i + 1])) up = 0; if (!(xh[i] > xh[i + 1])) down = 0; }
    }
    flint::InterpArgs a{};
    a.N = N; a.G = G; a.nI = D.nI; a.pstar = D.pstar; a.dv = D.view(); a.dcount = D.dcount; a.per_traj = per_traj ? 1 : 0;
    a.layout = layout; a.fma = me->arith == FLINT_ARITH_FMA;
    const size_t ny = (size_t)N * (size_t)G * (size_t)D.nI;
    double* d_y = nullptr; int* d_st = nullptr;
    if (ex.host) {
        double* d_x = scratch.alloc<double>(nx);
        d_y = scratch.alloc<double>(ny); d_st = scratch.alloc<int>((size_t)N);
        if (!d_x || !d_y || !d_st) return me->status = fail(FLINT_ERROR_MEMALLOC, "device allocation failed");
        /* a per-trajectory grid of a sharded call: this shard's rows of xarr[Ntot][G] */
        const bool okx = per_traj ? shard_copy(sh, N, 0, N, d_x, xarr, (size_t)G * 8, 1, true, "H2D xarr", s)
                                  : cuda_ok(cudaMemcpyAsync(d_x, xarr, (size_t)G * 8, cudaMemcpyHostToDevice, s), "H2D xarr");
        if (!okx) return me->status = FLINT_ERROR;
        a.xarr = d_x; a.yarr = d_y; a.status = d_st;
    } else {
        a.xarr = xarr; a.yarr = yarr; a.status = status;
        if (!a.status) { d_st = scratch.alloc<int>((size_t)N); a.status = d_st; }
    }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, D.device);
    if (D.kind_coef) {                                       /* coefficient records: the evaluation kernels (interp_kernel.cu) */
        int nl = 0;
        if (!cuda_ok(flint::launch_interp(a, up, down, sms, s, &nl), "interp launch")) return me->status = FLINT_ERROR;
        g_launches += nl;
    } else {
        /* step logs: erk_dense_kernel rebuilds every step's coefficients in registers and (layout 0) evaluates the grid from
         * there -- the fused path; for [nI][G][N] it writes them to a temporary coefficient store first */
        const flint_sys& sys = *me->sys;
        const flint::KernelEntry* ke = flint::find_kernel(me->method, sys.system_id, me->arith == FLINT_ARITH_FMA, me->events_on);
        if (!ke) ke = flint::find_kernel(me->method, sys.system_id, me->arith == FLINT_ARITH_FMA, false);
        if (!ke) ke = flint::get_kernel(me->method, sys.system_id, me->arith == FLINT_ARITH_FMA, false);
        if (!ke) return me->status = fail(FLINT_ERROR, "no kernel compiled for this method / system");
        if (!cuda_ok(flint::launch_interp_validate(a, up, down, s), "interp validate")) return me->status = FLINT_ERROR;
        g_launches += 1;
        KArgs c{};
        fill_static_kargs(me, c);
        c.N = N;
        flint::load_method_coefficients(me->method, c.coef);
        flint::det_pow_constants(c.powc);
        c.dv = a.dv; c.dense_count = D.dcount;
        c.zflag = me->dense_plane_ok ? D.dzflag : nullptr;
        DenseStore T;                                        /* temporary coefficient records ([nI][G][N] output) */
        struct Guard { DenseStore& t; ~Guard() { t.release(); } } guard{T};
        /* [N][G][nI] from step logs, two ways.  STAGED (default where the thread-per-item evaluation kernel serves, i.e. an even
         * nI): the ensemble goes through a small temporary coefficient store in parts of a few thousand trajectories -- the dense
         * kernel turns a part's logs into coefficient records, the evaluation kernel reads them back (mostly from L2) and writes
         * Yarr -- so both phases run their best kernel and the store the caller keeps is still the 4x smaller one of step logs.
         * FUSED (FLINT_B200_DENSE_EVAL=fused, odd nI, or no memory for the temporary): one kernel, the coefficients go from
         * registers through shared memory straight into the evaluation, Bip never exists in device memory. */
        bool staged = false;
        if (layout == 0 && D.nI % 2 == 0 && (reinterpret_cast<uintptr_t>(a.yarr) & 15) == 0) {
            const char* e = getenv("FLINT_B200_DENSE_EVAL");
            staged = !(e && e[0] == 'f');
        }
        double* tb[FLINT_DENSE_LEVELS] = {nullptr};          /* the staged path's temporary records: stream-ordered allocations */
        int trs = 0; long part = 0;
        if (staged) {
            /* freed blocks stay in the device's pool between calls (the default hands them back to the driver at the next
             * synchronisation, and mapping 2 GB again costs every call several milliseconds) */
            static std::atomic<unsigned long long> pool_kept{0};
            if (D.device >= 0 && D.device < 64 && !((pool_kept.load() >> D.device) & 1ull)) {
                cudaMemPool_t pool = nullptr;
                if (cudaDeviceGetDefaultMemPool(&pool, D.device) == cudaSuccess) {
                    unsigned long long keep = 3ull << 30;
                    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
                }
                cudaGetLastError();
                pool_kept.fetch_or(1ull << D.device);
            }
            trs = D.nI * (D.pstar + 1) + 3; trs += trs & 1; trs += (trs % 4 == 0) ? 2 : 0;
            const size_t per_traj = (size_t)D.cap[0] * (size_t)trs * 8;
            part = (long)(((size_t)1 << 30) / per_traj);                      /* ~1 GB of records per part */
            { const char* e = getenv("FLINT_B200_DENSE_PART"); if (e && atol(e) > 0) part = atol(e); }
            if (part < 256) part = 256;
            if (part > N) part = N;
            bool ok = true;
            for (int l = 0; l < D.nlev && ok; ++l)
                ok = cudaMallocAsync(&tb[l], (size_t)(l == 0 ? part : D.rows[l]) * (size_t)D.cap[l] * (size_t)trs * 8, s) == cudaSuccess;
            if (!ok) {                                                        /* no room: the fused kernel needs none */
                cudaGetLastError();
                for (int l = 0; l < D.nlev; ++l) if (tb[l]) { cudaFreeAsync(tb[l], s); tb[l] = nullptr; }
                staged = false;
            }
        }
        if (staged) {
            struct TmpGuard { double** b; int n; cudaStream_t st; ~TmpGuard() { for (int l = 0; l < n; ++l) if (b[l]) cudaFreeAsync(b[l], st); } } tguard{tb, D.nlev, s};
            c.dense_eval = 0; c.dense_bulk = dense_bulk_default();
            for (long t0 = 0; t0 < N; t0 += part) {
                const long nc = N - t0 < part ? N - t0 : part;
                /* the part's view of the logs and of the temporary store: level 0 is indexed by trajectory, deeper levels through
                 * the slot tables (rows of the whole ensemble) */
                flint::DenseView lv = a.dv, tv = a.dv;
                lv.base[0] += (size_t)t0 * (size_t)D.cap[0] * (size_t)D.rstride;
                tv.rstride = trs;
                for (int l = 0; l < D.nlev; ++l) tv.base[l] = tb[l];
                for (int l = 1; l < D.nlev; ++l) { lv.slot[l] = D.slot[l] + t0; tv.slot[l] = D.slot[l] + t0; }
                c.N = nc; c.dv = lv; c.dvo = tv; c.dense_count = D.dcount + t0;
                cudaError_t le = cudaSuccess;
                if (c.zflag) { le = flint::launch_dense(*ke, true, c, s); g_launches += 1; }
                if (le == cudaSuccess) { le = flint::launch_dense(*ke, false, c, s); g_launches += 1; }
                if (!cuda_ok(le, "dense kernel launch")) return me->status = FLINT_ERROR;
                flint::InterpArgs b = a;
                b.N = nc; b.dv = tv; b.dcount = D.dcount + t0; b.status = a.status + t0;
                b.yarr = a.yarr + (size_t)t0 * (size_t)G * (size_t)D.nI;
                if (a.per_traj) b.xarr = a.xarr + (size_t)t0 * (size_t)G;
                int nl = 0;
                if (!cuda_ok(flint::launch_interp(b, up, down, sms, s, &nl), "interp launch")) return me->status = FLINT_ERROR;
                g_launches += nl;
            }
        } else {
        if (layout == 0) {
            c.dense_eval = 1; c.dvo = c.dv;
            c.ix = a.xarr; c.iy = a.yarr; c.iG = G; c.ix_per_traj = a.per_traj; c.istatus = a.status;
        } else {
            T.device = D.device; T.N = N; T.n = D.n; T.nI = D.nI; T.pstar = D.pstar; T.kind_coef = true; T.nlev = D.nlev;
            T.rstride = D.nI * (D.pstar + 1) + 3; T.rstride += T.rstride & 1; T.rstride += (T.rstride % 4 == 0) ? 2 : 0;
            for (int l = 0; l < D.nlev; ++l) {
                T.cap[l] = D.cap[l]; T.rows[l] = D.rows[l];
                if (cudaMalloc(&T.base[l], (size_t)T.rows[l] * (size_t)T.cap[l] * (size_t)T.rstride * 8) != cudaSuccess) {
                    cudaGetLastError();
                    return me->status = fail(FLINT_ERROR_MEMALLOC, "no device memory for the coefficient records of a [nI][G][N] Interpolate; use layout 0 or flint_exec.dense_mode = 1");
                }
            }
            flint::DenseView tv = T.view();
            for (int l = 1; l < D.nlev; ++l) tv.slot[l] = D.slot[l];
            c.dense_eval = 0; c.dvo = tv; c.dense_bulk = dense_bulk_default();
        }
        cudaError_t le = cudaSuccess;
        if (c.zflag) { le = flint::launch_dense(*ke, true, c, s); g_launches += 1; }
        if (le == cudaSuccess) { le = flint::launch_dense(*ke, false, c, s); g_launches += 1; }
        if (!cuda_ok(le, "dense kernel launch")) return me->status = FLINT_ERROR;
        if (layout != 0) {
            flint::InterpArgs b = a;
            b.dv = c.dvo;
            int nl = 0;
            if (!cuda_ok(flint::launch_interp(b, up, down, sms, s, &nl), "interp launch")) return me->status = FLINT_ERROR;
            g_launches += nl;
            if (!cuda_ok(cudaStreamSynchronize(s), "interp execution")) return me->status = FLINT_ERROR;   /* before the temporary store goes */
        }
        }
    }
    std::vector<int> hst;
    if (ex.host) {
        hst.resize((size_t)N);
        /* layout 0: a trajectory's Yarr(nI,G) is one "element" of a one-row array; layout 1: nI*G rows of N doubles */
        const bool okc = cuda_ok(cudaMemcpyAsync(hst.data(), d_st, (size_t)N * 4, cudaMemcpyDeviceToHost, s), "D2H status") &&
                         (layout == 0 ? shard_copy(sh, N, 0, N, yarr, d_y, (size_t)G * (size_t)D.nI * 8, 1, false, "D2H yarr", s)
                                      : shard_copy(sh, N, 0, N, yarr, d_y, 8, (size_t)G * (size_t)D.nI, false, "D2H yarr", s));
        if (!okc) return me->status = FLINT_ERROR;
    }
    if (!cuda_ok(cudaStreamSynchronize(s), "interp execution")) return me->status = FLINT_ERROR;
    int rc = me->status;
    if (ex.host) {
        if (status) {
            if (sh.D == 1) for (long i = 0; i < N; ++i) status[i] = hst[(size_t)i];
            else for (long i = 0; i < N; ++i) status[((i / sh.chunk) * sh.D + sh.d) * sh.chunk + i % sh.chunk] = hst[(size_t)i];
        }
        if (scalar_call && hst[0] != FLINT_SUCCESS) { me->status = hst[0]; return me->status; }   /* :62-76: error return keeps storage */
    }
    if (dealloc) { D.release(); me->dense_allocated = false; }                                     /* :114-118 */
    return rc;
}

extern "C" int flint_erk_interpolate(flint_erk_t* me, const double* xarr, long G, double* yarr, int has_save, int save_coeffs)
{
    if (!me) return FLINT_ERROR_PARAMS;
    ExecCtx ex;
    return interpolate_impl(me, xarr, G, yarr, 0, nullptr, has_save != 0, save_coeffs != 0, ex, true);
}

static int interpolate_batch(flint_erk_t* me, const double* xarr, long G, double* yarr, int layout, int* status,
                             int has_save, int save_coeffs, const flint_exec* exec, bool per_traj)
{
    if (!me) return FLINT_ERROR_PARAMS;
    if (layout != 0 && layout != 1) return me->status = fail(FLINT_ERROR_PARAMS, "layout must be 0 or 1");
    const ExecCtx ex = make_exec(exec);
    if (me->shard_N > 0 && !me->shard.empty() && !me->dense.allocated()) {
        /* the last Integrate was sharded over devices: every shard interpolates where its dense output lives and writes the
         * caller's (host) Yarr at its trajectories' own offsets */
        if (!ex.host) return me->status = fail(FLINT_ERROR_PARAMS, "the dense output of a sharded Integrate is interpolated into host buffers");
        if (!me->interp_on) return me->status = FLINT_ERROR_INTP_OFF;
        const int S = (int)me->shard.size(), D = me->shard_devices > 0 ? me->shard_devices : S;
        const bool dealloc = has_save ? !save_coeffs : true;
        std::vector<int> rc(S, FLINT_SUCCESS);
        std::vector<std::string> err(S);
        std::vector<std::thread> th;
        for (int d = 0; d < D; ++d)
            th.emplace_back([&, d]() {
                for (int k = d; k < S; k += D) {
                    Shard sh; sh.Ntot = me->shard_N; sh.D = S; sh.d = k; sh.chunk = me->shard_chunk;
                    ExecCtx e1 = ex;
                    e1.device = me->shard_device[k]; e1.stream = me->shard_stream[k]; e1.n_devices = 0;
                    if (shard_count(sh.Ntot, S, k, sh.chunk) == 0) continue;
                    flint_erk* sub = me->shard[k];
                    const bool was_spilled = sub->dense.spilled;
                    if (was_spilled) {                       /* the part's dense output back from host memory */
                        if (!cuda_ok(cudaSetDevice(e1.device), "cudaSetDevice") || !sub->dense.restore(e1.stream)) {
                            rc[k] = fail(FLINT_ERROR_MEMALLOC, "dense_spill: a part's dense output does not fit in device memory");
                            err[k] = g_last_error;
                            break;
                        }
                        sub->dense.spilled = false;          /* interpolate_impl sees an ordinary device store */
                    }
                    /* a spilled part keeps its host copy when the caller saves the coefficients: only the device copy goes */
                    rc[k] = interpolate_impl(sub, xarr, G, yarr, layout, status, true, was_spilled ? true : !dealloc, e1, false, &sh, per_traj);
                    if (was_spilled) {
                        cudaSetDevice(e1.device);
                        sub->dense.release_device();
                        if (dealloc && rc[k] == FLINT_SUCCESS) { sub->dense.release(); sub->dense_allocated = false; }
                        else sub->dense.spilled = true;
                    }
                    if (rc[k] != FLINT_SUCCESS) { err[k] = g_last_error; break; }
                }
            });
        for (auto& t : th) t.join();
        int st = FLINT_SUCCESS;
        for (int k = 0; k < S; ++k) if (rc[k] != FLINT_SUCCESS && st == FLINT_SUCCESS) { st = rc[k]; g_last_error = err[k]; }
        if (dealloc && st == FLINT_SUCCESS) { me->dense_allocated = false; me->shard_N = 0; }
        return me->status = st;
    }
    return interpolate_impl(me, xarr, G, yarr, layout, status, has_save != 0, save_coeffs != 0, ex, false, nullptr, per_traj);
}

extern "C" int flint_erk_interpolate_batch(flint_erk_t* me, const double* xarr, long G, double* yarr, int layout, int* status,
                                           int has_save, int save_coeffs, const flint_exec* exec)
{
    return interpolate_batch(me, xarr, G, yarr, layout, status, has_save, save_coeffs, exec, false);
}
extern "C" int flint_erk_interpolate_batch_grids(flint_erk_t* me, const double* xarr, long G, double* yarr, int layout, int* status,
                                                 int has_save, int save_coeffs, const flint_exec* exec)
{
    return interpolate_batch(me, xarr, G, yarr, layout, status, has_save, save_coeffs, exec, true);
}

/* ------------------------------------------------------------------ Info (src/ERKInit.f90:346-394) */
extern "C" int flint_erk_info(const flint_erk_t* me, int* last_status, int* nsteps, int* naccept, int* nreject, int* nfcalls,
                              int* interp_ready, double* x0, double* y0, double* hf, double* xf, double* yf)
{
    if (!me) return FLINT_ERROR_PARAMS;
    if (last_status) *last_status = me->status;
    if (nsteps) *nsteps = me->total;
    if (naccept) *naccept = me->accepted;
    if (nreject) *nreject = me->rejected;
    if (nfcalls) *nfcalls = me->fcalls;
    if (interp_ready) *interp_ready = me->interp_on ? 1 : 0;
    const int n = me->sys ? me->sys->n : 0;
    if (x0) *x0 = me->X0;
    if (y0) for (int c = 0; c < n; ++c) y0[c] = me->Y0[c];
    if (hf) *hf = me->hf;
    if (xf) *xf = me->Xf;
    if (yf) for (int c = 0; c < n; ++c) yf[c] = me->Yf[c];
    return me->status;
}

/* ------------------------------------------------------------------ fp64 peak microbenchmark */
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-3, x2 = x0 + 2e-3, x3 = x0 + 3e-3, x4 = x0 + 4e-3, x5 = x0 + 5e-3,
           x6 = x0 + 6e-3, x7 = x0 + 7e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
            x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

extern "C" double flint_b200_measure_fp64_peak(int device, double* sm_clock_mhz_out)
{
    if (cudaSetDevice(device) != cudaSuccess) { g_last_error = "cudaSetDevice failed"; cudaGetLastError(); return -1.0; }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int block = 256, grid = sms * 8, iters = 4096;
    double* d = nullptr;
    if (cudaMalloc(&d, (size_t)grid * block * 8) != cudaSuccess) { cudaGetLastError(); return -1.0; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        dfma_peak_kernel<<<grid, block>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaGetLastError(); best = -1.0; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        g_launches += 1;
        const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)grid * (double)block;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    if (sm_clock_mhz_out) { int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device); *sm_clock_mhz_out = khz / 1000.0; }
    return best;
}
