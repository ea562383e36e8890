// Shared by the drop-in host shims (ORBextractor.cc, ORBmatcher_fbe.cc, Frame_fbe.cc): which GPU they use.
// FBE_DEVICE (environment, read once) selects the CUDA device of EVERY shim, so the extractor, the matchers and the Frame
// helpers of one process always land on the same GPU; unset = device 0.
#ifndef FBE_HOST_H
#define FBE_HOST_H
#include <cstdlib>

inline int fbe_host_device() {
    static const int dev = [] { const char* d = std::getenv("FBE_DEVICE"); return d ? std::atoi(d) : 0; }();
    return dev;
}

#endif
