// batch_host.h -- host-side logic of the batched launches (chain1024.cu BATCH kernels, k_shift_batch): filling a
// stream descriptor, cutting a list of streams into launches whose descriptors fit the kernel parameters, and the
// check that no stream of a launch writes what another stream reads or writes (inside one kernel nothing is
// ordered).  Header-only and CUDA-free apart from the types of nco.cuh, so that tests/host_logic can run it on the CPU.
#pragma once
#include <algorithm>
#include <cstdint>
#include <utility>
#include <vector>

#include "nco.cuh"

namespace hz {

// descriptor k of a launch: source, destination, the table's segments appended to `pool` at seg_off, and the phase
// step of the DOMINANT (longest) linear segment -- split launches build the stream's multiplier table for it
inline void fill_desc(StreamDesc &d, NcoSegment *pool, uint32_t seg_off, const void *src, void *dst, const NcoTable &table) {
    d.src = (const uint8_t *)src;
    d.dst = dst;
    d.seg_off = seg_off;
    d.count = table.count;
    d.dp_nom = 0;
    uint32_t longest = 0;
    for (int k = 0; k < table.count; k++) {
        pool[seg_off + k] = table.seg[k];
        if (table.seg[k].count > longest && table.seg[k].dp) longest = table.seg[k].count, d.dp_nom = table.seg[k].dp;
    }
}

// Greedy cut of streams 0..n-1 (segment counts seg_count[k]) into consecutive launches of at most max_streams streams
// and max_segs segments: (first, count) pairs that cover every stream exactly once.  A stream with more than max_segs
// segments gets a launch of its own (the caller routes such streams elsewhere first).
inline std::vector<std::pair<uint32_t, uint32_t>> plan_param_launches(const std::vector<int> &seg_count, uint32_t max_streams, uint32_t max_segs) {
    std::vector<std::pair<uint32_t, uint32_t>> out;
    uint32_t k0 = 0;
    const uint32_t n = (uint32_t)seg_count.size();
    while (k0 < n) {
        uint32_t nb = 0, ns = 0;
        while (k0 + nb < n && nb < max_streams && (nb == 0 || ns + (uint32_t)seg_count[k0 + nb] <= max_segs)) {
            ns += (uint32_t)seg_count[k0 + nb];
            nb++;
        }
        out.emplace_back(k0, nb);
        k0 += nb;
    }
    return out;
}

struct BufSpan {
    uintptr_t lo, hi;  // [lo, hi)
    bool write;
};
// true when some WRITE span overlaps any other span (read-read overlap is fine).  Sweep over the spans sorted by
// address: O(n log n) -- the channelizer calls this with a thousand spans per step.
inline bool write_conflict(std::vector<BufSpan> spans) {
    std::sort(spans.begin(), spans.end(), [](const BufSpan &a, const BufSpan &b) { return a.lo < b.lo; });
    uintptr_t hi_any = 0, hi_write = 0;
    for (const BufSpan &v : spans) {
        if (v.hi <= v.lo) continue;  // empty
        if (v.lo < (v.write ? hi_any : hi_write)) return true;
        if (v.hi > hi_any) hi_any = v.hi;
        if (v.write && v.hi > hi_write) hi_write = v.hi;
    }
    return false;
}

// Same-kind spans that touch or overlap, merged (reads first, each kind sorted by address).  K buffers that sit side by
// side in memory -- ring slots, one allocation cut in K -- become one span: what a batched launch records in the
// context's OverlapWindow, so that a 64-buffer launch does not use up a third of the window.
inline std::vector<BufSpan> coalesce_spans(std::vector<BufSpan> spans) {
    std::sort(spans.begin(), spans.end(), [](const BufSpan &a, const BufSpan &b) { return a.write != b.write ? !a.write : a.lo < b.lo; });
    std::vector<BufSpan> out;
    for (const BufSpan &v : spans) {
        if (v.hi <= v.lo) continue;
        if (!out.empty() && out.back().write == v.write && v.lo <= out.back().hi) {
            if (v.hi > out.back().hi) out.back().hi = v.hi;
        } else {
            out.push_back(v);
        }
    }
    return out;
}

}  // namespace hz
