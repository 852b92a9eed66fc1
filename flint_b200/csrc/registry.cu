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

}  // namespace flint

extern "C" int flint_b200_kernel_count(void) { return flint::kernel_count(); }
