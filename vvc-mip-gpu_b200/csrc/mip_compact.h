// mip_compact.h -- the compact cost table (MIPB200_EMIT_COSTS_COMPACT): same entries, same order as the int32 table
// (reference order, constants.h:1558-1631), but the costs of CU types of at most 32 samples -- 4x4, 8x4, 4x8: 57 344 of the
// 97 840 entries of a CTU -- are uint16: min(2*SAD, SATD) <= 2 * 32 * 1023 = 65 472 for samples of up to 10 bits, so nothing
// is lost (0xFFFF = CU not inside the frame); all other types keep int32 (-1).  276 672 instead of 391 360 bytes per CTU:
// 71 % of the bytes over the host link, which is what bounds the full-table rate.  Host-side layout helpers, shared by the
// engine (mipb200_expand_costs), the CLI and the tests.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "mip_tables.h"

#define MIP_COMPACT_BYTES_PER_CTU 276672

static inline int mip_compact_narrow(int t) { return (int)MIP_TYPES[t].w * MIP_TYPES[t].h <= 32; }

// byte offset of type t's block inside a CTU's compact record
static inline size_t mip_compact_type_offset(int t) {
    size_t o = 0;
    for (int i = 0; i < t; ++i) o += (size_t)MIP_TYPES[i].n * MIP_TYPES[i].modes * (mip_compact_narrow(i) ? 2 : 4);
    return o;
}

// compact[n_ctus][MIP_COMPACT_BYTES_PER_CTU] -> cost[n_ctus][MIP_COSTS_PER_CTU], CTUs [ctu0, ctu1)
static inline void mip_compact_expand(const void* compact, int32_t* cost, int ctu0, int ctu1) {
    size_t off[MIP_NUM_TYPES];
    for (int t = 0; t < MIP_NUM_TYPES; ++t) off[t] = mip_compact_type_offset(t);
    for (int ctu = ctu0; ctu < ctu1; ++ctu) {
        const uint8_t* rec = static_cast<const uint8_t*>(compact) + (size_t)ctu * MIP_COMPACT_BYTES_PER_CTU;
        int32_t* out = cost + (size_t)ctu * MIP_COSTS_PER_CTU;
        for (int t = 0; t < MIP_NUM_TYPES; ++t) {
            const size_t n = (size_t)MIP_TYPES[t].n * MIP_TYPES[t].modes;
            int32_t* o = out + MIP_TYPES[t].cost_off;
            if (mip_compact_narrow(t)) {
                const uint16_t* s = reinterpret_cast<const uint16_t*>(rec + off[t]);
                for (size_t i = 0; i < n; ++i) o[i] = s[i] == 0xFFFFu ? -1 : (int32_t)s[i];
            } else {
                memcpy(o, rec + off[t], n * sizeof(int32_t));
            }
        }
    }
}

// the inverse (tests, the CLI's CPU test double): 0 on success, -1 if a narrow entry does not fit 16 bits
static inline int mip_compact_pack(const int32_t* cost, void* compact, int n_ctus) {
    size_t off[MIP_NUM_TYPES];
    for (int t = 0; t < MIP_NUM_TYPES; ++t) off[t] = mip_compact_type_offset(t);
    for (int ctu = 0; ctu < n_ctus; ++ctu) {
        uint8_t* rec = static_cast<uint8_t*>(compact) + (size_t)ctu * MIP_COMPACT_BYTES_PER_CTU;
        const int32_t* in = cost + (size_t)ctu * MIP_COSTS_PER_CTU;
        for (int t = 0; t < MIP_NUM_TYPES; ++t) {
            const size_t n = (size_t)MIP_TYPES[t].n * MIP_TYPES[t].modes;
            const int32_t* s = in + MIP_TYPES[t].cost_off;
            if (mip_compact_narrow(t)) {
                uint16_t* o = reinterpret_cast<uint16_t*>(rec + off[t]);
                for (size_t i = 0; i < n; ++i) {
                    if (s[i] < -1 || s[i] >= 0xFFFF) return -1;
                    o[i] = s[i] < 0 ? (uint16_t)0xFFFFu : (uint16_t)s[i];
                }
            } else {
                memcpy(rec + off[t], s, n * sizeof(int32_t));
            }
        }
    }
    return 0;
}
