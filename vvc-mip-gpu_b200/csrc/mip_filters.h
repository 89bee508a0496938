// Low-pass coefficient library of the "alternative samples" filters.
// Values follow the reference's convKernelLib / convKernelLib_5x5 (constants.cl:91-183,
// constants.h:63-128): five 3x3 tables and three 5x5 tables selected by --KernelIdx.
// Every table is symmetric, so it is stored by its generating rule where one exists.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MIP_HD __host__ __device__
#else
#define MIP_HD
#endif

#define MIP_NUM_K3 5
#define MIP_NUM_K5 3

// 3x3 tables: all are {corner, edge, corner; edge, centre, edge; corner, edge, corner}
//   idx 0: box (1,1,1)   1: (1,2,3)   2: (1,2,12)   3: (1,1,8)   4: binomial (1,2,4)
MIP_HD static inline int mip_k3(int idx, int dy, int dx) {  // dy,dx in -1..1
    const int cls = (dy == 0) + (dx == 0);  // 0 corner, 1 edge, 2 centre
    if (cls == 0) return 1;
    if (cls == 1) return (idx == 0 || idx == 3) ? 1 : 2;
    return idx == 0 ? 1 : (idx == 1 ? 3 : (idx == 2 ? 12 : (idx == 3 ? 8 : 4)));
}

// 5x5 tables: 0 = box, 1 = box with centre 5, 2 = (1,2,3,2,1) outer product
MIP_HD static inline int mip_k5(int idx, int dy, int dx) {  // dy,dx in -2..2
    int ay = dy < 0 ? -dy : dy, ax = dx < 0 ? -dx : dx;
    if (idx == 0) return 1;
    if (idx == 1) return (ay == 0 && ax == 0) ? 5 : 1;
    return (3 - ay) * (3 - ax);
}
