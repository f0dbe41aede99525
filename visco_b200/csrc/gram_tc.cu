// placeholder until the tcgen05 kernel lands (next commit)
#include "common.cuh"
bool vk_gram_tc_supported(int, int, int) { return false; }
int vk_launch_gram_tc(vk_context* h, const float2*, int, int, int, float2*) {
    return vk_fail(h, VK_EINVAL, "tcgen05 Gram not built");
}
