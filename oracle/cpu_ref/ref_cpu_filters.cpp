// Exposes the reference's OWN host-side CPU filters (main_aux_functions.h:1323-1773,
// parallelOptFilterCpuInt_3x3 / _5x5) through a C ABI, compiled from the reference checkout where it
// lies (-I$(REF)); nothing of the reference is copied here.  They are an independent pin -- runnable
// without any GPU -- for the oracle's filterFrame_2d_int_quarterCtu / filterFrame_2d_int_5x5_quarterCtu
// (same zero-padding + renormalisation rule).  Test infrastructure only.
#ifndef USE_ALTERNATIVE_SAMPLES
#define USE_ALTERNATIVE_SAMPLES 1   // main.cpp:10 defines it before including the header
#endif
#include "main_aux_functions.h"

extern "C" __attribute__((visibility("default")))
void mipref_cpu_filter_int(const unsigned short* in, unsigned short* out, int width, int height, int taps, int kernel_idx, int threads) {
    unsigned short* src = const_cast<unsigned short*>(in);
    if (taps == 3) parallelOptFilterCpuInt_3x3(src, out, width, height, kernel_idx, threads);
    else parallelOptFilterCpuInt_5x5(src, out, width, height, kernel_idx, threads);
}
