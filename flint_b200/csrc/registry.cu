/* registry.cu -- table of the integrate kernels (compiled into the library by inst_*.cu, or built at run time by rtc.cu) and
 * the launchers they share. */
#include <mutex>
#include <vector>
#include "kargs.h"
#include "erk_methods_gen.cuh"

namespace flint {

/* entries are added by static initialisers and, later, by run-time registration from any host thread; they are never removed
 * and never move (heap nodes), so the pointers find_kernel returns stay valid */
static std::mutex& table_mutex()
{
    static std::mutex m;
    return m;
}
static std::vector<KernelEntry*>& table()
{
    static std::vector<KernelEntry*> t;
    return t;
}
void register_kernel(const KernelEntry& e)
{
    std::lock_guard<std::mutex> g(table_mutex());
    table().push_back(new KernelEntry(e));
}
const KernelEntry* find_kernel(int method, int system, bool fma, bool events)
{
    std::lock_guard<std::mutex> g(table_mutex());
    for (const KernelEntry* e : table())
        if (e->method == method && e->system == system && e->fma == fma && e->events == events) return e;
    return nullptr;
}
const KernelEntry* find_system(int system)
{
    std::lock_guard<std::mutex> g(table_mutex());
    for (const KernelEntry* e : table())
        if (e->system == system) return e;
    return nullptr;
}
int kernel_count()
{
    std::lock_guard<std::mutex> g(table_mutex());
    return (int)table().size();
}
void load_method_coefficients(int method, double* dst)
{
    double (&d)[FLINT_NCOEF] = *(double (*)[FLINT_NCOEF])dst;
    switch (method) {
    case ERK_DOP54: MethodDOP54::load_coefficients(d); break;
    case ERK_VERNER98R: MethodVerner98R::load_coefficients(d); break;
    case ERK_VERNER65E: MethodVerner65E::load_coefficients(d); break;
    default: MethodDOP853::load_coefficients(d); break;
    }
}

/* ---- launchers (kargs.h) ---- */
cudaError_t launch_kernel(const void* k, const KArgs& a, dim3 grid, int block, size_t smem, cudaStream_t s)
{
    void* args[1] = {const_cast<KArgs*>(&a)};
    return cudaLaunchKernel(k, grid, dim3((unsigned)block), args, smem, s);
}

/* resident CTAs per SM of kernel k at this CTA size: one occupancy query per (kernel, CTA size, device) and host thread --
 * handles on different devices launch from different threads */
static cudaError_t resident_ctas(const void* k, int block, int smem, int* out)
{
    struct Key { const void* k; int block, smem, dev, bps; };
    static thread_local std::vector<Key> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    for (const Key& c : cache)
        if (c.k == k && c.block == block && c.smem == smem && c.dev == dev) { *out = c.bps; return cudaSuccess; }
    int q = 0;
    const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, k, block, (size_t)smem);
    if (e != cudaSuccess) return e;
    if (q < 1) q = 1;
    cache.push_back(Key{k, block, smem, dev, q});
    *out = q;
    return cudaSuccess;
}

cudaError_t launch_step(const StepKernel& k, const KArgs& a, int sm_count, long n_items, int block_req, int bps_req, cudaStream_t s)
{
    const int smem = k.smem, BLOCK = k.block;
    int block = (block_req > 0 && block_req < BLOCK) ? (block_req + 31) / 32 * 32 : BLOCK;
    if (smem > 0) block = BLOCK;                            /* the shared-memory layout is compiled for this block size */
    if (smem > 48 * 1024) {                                 /* per device context, and cheap: not cached */
        const cudaError_t e = cudaFuncSetAttribute(k.classic, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    int bps = 1;
    { const cudaError_t e = resident_ctas(k.classic, block, smem, &bps); if (e != cudaSuccess) return e; }
    if (bps_req > 0 && bps_req < bps) bps = bps_req;
    long grid = (long)sm_count * bps;                       /* persistent lanes: one resident wave */
    /* a pass smaller than the device keeps its lanes PACKED (full warps on few SMs): a warp costs the same issue slots whatever
     * the number of its active lanes, and one warp per scheduler is as fast as a warp gets -- spreading 4096 trajectories as four
     * per CTA was measured 1.7x slower (profiles/r02_notes.md) */
    const long need = (n_items + block - 1) / block;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (k.gang) {
        /* small shards: more than one but fewer than ~8 trajectories per lane -> the gang-scheduled kernel (see erk_step_kernel) */
        const long lanes = grid * block;
        int S = 0;
        if (a.round_req >= 2) S = a.round_req;
        else if (a.round_req == 0 && a.n_items_exact && n_items > lanes && n_items < 8 * lanes) S = n_items < 4 * lanes ? 64 : 128;
        if (S > 4096) S = 4096;
        if (S >= 2 && a.n_items_exact) {
            const long per = (n_items + grid - 1) / grid;
            const long smem2 = smem + ((per + 1) & ~1L) * 4 + (long)block * 4;
            if (smem2 <= 200 * 1024) {
                KArgs b = a;
                b.round_len = S;
                if (smem2 > 48 * 1024) {
                    const cudaError_t e = cudaFuncSetAttribute(k.gang, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
                    if (e != cudaSuccess) return e;
                }
                return launch_kernel(k.gang, b, dim3((unsigned)grid), block, (size_t)smem2, s);
            }
        }
    }
    return launch_kernel(k.classic, a, dim3((unsigned)grid), block, (size_t)smem, s);
}

cudaError_t launch_dense(const KernelEntry& e, bool plane, const KArgs& a, cudaStream_t s)
{
    const void* k = plane ? e.dense_plane : e.dense_coeffs;
    const long cap = a.dv.capacity();
    long gy = (cap + kDenseBlock - 1) / kDenseBlock;
    if (gy < 1) gy = 1;
    {   /* blocks of level 0 only: deeper levels hold a few trajectories, whose CTAs walk their further blocks (erk_dense_kernel) */
        const long gy0 = (a.dv.nlev > 0 ? a.dv.cum[1] : cap) / kDenseBlock;
        if (gy0 >= 1 && gy > gy0) gy = gy0;
    }
    if (gy > 65535) return cudaErrorInvalidConfiguration;
    /* fused evaluation: coefficients + ranges; otherwise the block's records, staged for one bulk store (erk_dense_kernel) */
    const int smem_main = a.dense_eval ? dense_smem_doubles_ps(e.pstar, a.nI) * 8 : (a.dense_bulk ? kDenseBlock * a.dvo.rstride * 8 : 0);
    const int smem = (smem_main + 15) / 16 * 16 + dense_prefetch_doubles() * 8;       /* + the prefetch slots (erk_dense_kernel) */
    if (smem > 48 * 1024) {
        const cudaError_t er = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (er != cudaSuccess) return er;
    }
    /* a CTA serves several trajectories in turn (the next one's logs travel while it computes): ~8 per CTA, at least one
     * resident wave */
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long gx = (a.N + 7) / 8;
    const long wave = ((long)sms * 3 + gy - 1) / gy;
    if (gx < wave) gx = wave;
    if (gx > a.N) gx = a.N;
    return launch_kernel(k, a, dim3((unsigned)gx, (unsigned)gy), kDenseBlock, (size_t)smem, s);
}

/* work counters of one Integrate chunk, written in stream order (a pageable host-to-device copy would make the host wait
 * for everything already queued on the stream) */
__global__ void set_counters_kernel(unsigned long long* ctr, unsigned long long c0, unsigned long long c1, unsigned long long c2,
                                    unsigned long long c3)
{
    ctr[0] = c0; ctr[1] = c1; ctr[2] = c2; ctr[3] = c3;
}
cudaError_t launch_set_counters(unsigned long long* ctr, unsigned long long c0, unsigned long long c1, unsigned long long c2,
                                unsigned long long c3, cudaStream_t s)
{
    set_counters_kernel<<<1, 1, 0, s>>>(ctr, c0, c1, c2, c3);
    return cudaGetLastError();
}

__global__ void dense_slots_kernel(int* slot, const int* __restrict__ list, long n)
{
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) slot[list[j]] = (int)j;
}
cudaError_t launch_dense_slots(int* slot, const int* list, long n, cudaStream_t s)
{
    dense_slots_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(slot, list, n);
    return cudaGetLastError();
}

}  // namespace flint

extern "C" int flint_b200_kernel_count(void) { return flint::kernel_count(); }
