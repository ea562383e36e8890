// placeholder: matcher entry points (implemented next)
#include "fbe_internal.cuh"
extern "C" {
int fbe_matcher_create(float, int32_t, int32_t, fbe_matcher**) { return FBE_E_UNSUPPORTED; }
}
