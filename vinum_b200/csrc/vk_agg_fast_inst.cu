// Instantiations of agg_fast_kernel, one slice per VK_FAST_PART (see build.py) so that the
// 60-odd kernel variants compile in parallel.
#include "vk_agg_fast.cuh"

namespace vk {

#define VK_FAST_LEAN(name, PK, MODE)                                                   \
    VK_FAST_DECL(name) {                                                               \
        return l.direct ? launch_fast_lean<PK, MODE, true>(p, l, s) : launch_fast_lean<PK, MODE, false>(p, l, s); \
    }
#define VK_FAST_RT(name, PK) \
    VK_FAST_DECL(name) { return launch_fast_generic<PK>(p, l, s); }

#if VK_FAST_PART == 0
VK_FAST_LEAN(launch_fast_none_all8, PK_NONE, FM_ALL8)
#elif VK_FAST_PART == 1
VK_FAST_LEAN(launch_fast_none_key4, PK_NONE, FM_KEY4)
VK_FAST_RT(launch_fast_none_rt, PK_NONE)
#elif VK_FAST_PART == 2
VK_FAST_LEAN(launch_fast_f64_all8, PK_F64_VEC, FM_ALL8)
#elif VK_FAST_PART == 3
VK_FAST_LEAN(launch_fast_f64_key4, PK_F64_VEC, FM_KEY4)
VK_FAST_RT(launch_fast_f64_rt, PK_F64_VEC)
#elif VK_FAST_PART == 4
VK_FAST_RT(launch_fast_mask_rt, PK_MASK)
VK_FAST_RT(launch_fast_i64_rt, PK_I64_VEC)
VK_FAST_RT(launch_fast_gen_rt, PK_GENERIC)
#else
#error "VK_FAST_PART must be 0..4"
#endif

}  // namespace vk
