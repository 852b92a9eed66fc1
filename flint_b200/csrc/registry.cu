/* registry.cu -- table of the integrate kernels compiled into the library. */
#include <vector>
#include "kargs.h"

namespace flint {

static std::vector<KernelEntry>& table()
{
    static std::vector<KernelEntry> t;
    return t;
}
void register_kernel(const KernelEntry& e) { table().push_back(e); }
const KernelEntry* find_kernel(int method, int system, bool fma, bool events)
{
    for (const KernelEntry& e : table())
        if (e.method == method && e.system == system && e.fma == fma && e.events == events) return &e;
    return nullptr;
}
const KernelEntry* find_system(int system)
{
    for (const KernelEntry& e : table())
        if (e.system == system) return &e;
    return nullptr;
}
int kernel_count() { return (int)table().size(); }

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
