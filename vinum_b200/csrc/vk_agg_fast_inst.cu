// Instantiations of agg_fast_kernel, one slice per VK_FAST_PART (see build.py) so that the
// 90-odd kernel variants compile in parallel (each lean slice is one predicate kind x one key
// width x direct / table group ids; ptxas time is dominated by these kernels).
#include "vk_agg_fast.cuh"

namespace vk {

#define VK_FAST_LEAN_ONE(name, PK, MODE, DIRECT) \
    VK_FAST_DECL(name) { return launch_fast_lean<PK, MODE, DIRECT>(p, l, s); }
#define VK_FAST_RT(name, PK) \
    VK_FAST_DECL(name) { return launch_fast_generic<PK>(p, l, s); }

#define VK_FAST_GEN_ONE(name, NV, NW) \
    VK_FAST_DECL(name) { VK_FAST_GO(PK_GENERIC, NV, NW, FM_RUNTIME, false, false); }

#if VK_FAST_PART == 0
VK_FAST_LEAN_ONE(launch_fast_none_all8_d, PK_NONE, FM_ALL8, true)
#elif VK_FAST_PART == 1
VK_FAST_LEAN_ONE(launch_fast_none_all8_t, PK_NONE, FM_ALL8, false)
#elif VK_FAST_PART == 2
VK_FAST_LEAN_ONE(launch_fast_none_key4_d, PK_NONE, FM_KEY4, true)
#elif VK_FAST_PART == 3
VK_FAST_LEAN_ONE(launch_fast_none_key4_t, PK_NONE, FM_KEY4, false)
#elif VK_FAST_PART == 4
VK_FAST_LEAN_ONE(launch_fast_f64_all8_d, PK_F64_VEC, FM_ALL8, true)
#elif VK_FAST_PART == 5
VK_FAST_LEAN_ONE(launch_fast_f64_all8_t, PK_F64_VEC, FM_ALL8, false)
#elif VK_FAST_PART == 6
VK_FAST_LEAN_ONE(launch_fast_f64_key4_d, PK_F64_VEC, FM_KEY4, true)
#elif VK_FAST_PART == 7
VK_FAST_LEAN_ONE(launch_fast_f64_key4_t, PK_F64_VEC, FM_KEY4, false)
#elif VK_FAST_PART == 8
VK_FAST_RT(launch_fast_none_rt, PK_NONE)
VK_FAST_RT(launch_fast_f64_rt, PK_F64_VEC)
#elif VK_FAST_PART == 9
VK_FAST_RT(launch_fast_mask_rt, PK_MASK)
VK_FAST_RT(launch_fast_i64_rt, PK_I64_VEC)
#elif VK_FAST_PART == 10
VK_FAST_GEN_ONE(launch_fast_gen_rt_c0, 0, 2)
#elif VK_FAST_PART == 11
VK_FAST_GEN_ONE(launch_fast_gen_rt_c1n, 1, 2)
#elif VK_FAST_PART == 12
VK_FAST_GEN_ONE(launch_fast_gen_rt_c1, 1, 4)
#elif VK_FAST_PART == 13
VK_FAST_GEN_ONE(launch_fast_gen_rt_c2, 2, 4)
#elif VK_FAST_PART == 14
VK_FAST_GEN_ONE(launch_fast_gen_rt_c3, 3, 4)
#else
#error "VK_FAST_PART must be 0..14"
#endif

}  // namespace vk
