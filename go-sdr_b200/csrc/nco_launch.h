// nco_launch.h -- cut one buffer's NCO segment list into kernel launches.
// A launch carries at most kMaxSegsPerLaunch segments in its parameters and indexes samples with
// 32 bits.  Launch boundaries fall on multiples of `align` samples (the FFT block for the fused
// chain, 2 for the pair-vectorised elementwise kernels); a linear segment that straddles a
// boundary is split exactly (base + d*step is exact inside a binade).  Steady state: one launch.
#pragma once
#include <vector>

#include "nco.cuh"

namespace hz {

struct NcoLaunch {
    size_t first = 0, count = 0;  // sample range [first, first+count) of the buffer
    NcoTable table;
};

inline int plan_nco_launches(const std::vector<HostSeg> &segs, size_t n, size_t align, double shift_hz,
                             std::vector<NcoLaunch> &out, size_t max_count = (size_t)1 << 30) {
    out.clear();
    size_t s = 0, pos = 0;
    while (pos < n) {
        // 1. how far can this launch reach?
        size_t end = pos, k = s;
        int cnt = 0;
        while (k < segs.size() && cnt < kMaxSegsPerLaunch) {
            end = (size_t)(segs[k].j0 + segs[k].count);
            cnt++;
            k++;
            if (end - pos >= max_count) break;
        }
        if (end - pos > max_count) end = pos + max_count;
        if (end > n) end = n;
        // 2. align the cut (the buffer end itself needs no alignment)
        if (end < n) end = pos + ((end - pos) / align) * align;
        if (end <= pos)
            return fail(HZSDR_ERR_UNSUPPORTED, "NCO: more than %d accumulator segments inside one %zu-sample unit", kMaxSegsPerLaunch, align);
        // 3. clip the overlapping segments into the launch table
        NcoLaunch L;
        L.first = pos;
        L.count = end - pos;
        L.table.count = 0;
        size_t k2 = s;
        for (; k2 < segs.size() && (size_t)segs[k2].j0 < end; k2++) {
            HostSeg h = segs[k2];
            const size_t h_end = (size_t)(h.j0 + h.count);
            if (h_end <= pos) continue;
            if ((size_t)h.j0 < pos) {  // head already consumed by the previous launch
                const size_t d = pos - (size_t)h.j0;
                h.base += (double)d * h.step;
                h.j0 = pos;
                h.count -= d;
            }
            if (h_end > end) h.count = end - (size_t)h.j0;
            if (L.table.count >= kMaxSegsPerLaunch) return fail(HZSDR_ERR_UNSUPPORTED, "NCO: launch table overflow");
            L.table.seg[L.table.count++] = to_device_segment(h, pos, shift_hz);
        }
        out.push_back(L);
        // 4. next launch starts at the first segment that reaches past `end`
        while (s < segs.size() && (size_t)(segs[s].j0 + segs[s].count) <= end) s++;
        pos = end;
    }
    return HZSDR_OK;
}

}  // namespace hz
