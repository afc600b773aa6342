// poly_host.h -- the index plan of the fused polyphase decimator (polyphase.cu), shared by the kernel's host side and
// by tests/host_logic, which replays the kernel's data flow (scatter by phase, register windows) on the CPU against a
// direct FIR for many (taps, decimate) pairs.
//
//   tap k = D q + p      ->  taps_t[p][q]                 q < qpad = 8 npairs (zero beyond the filter)
//   tile input a         ->  phase p = D - 1 - a % D, slot jj = a / D, stored at row p, position pos(jj + joff)
//   output 8 lane + r, tap 8 m + c  ->  window slot x = r + 7 - c of pair m = position 9 (lane + k0 - m) + x + x / 8
// joff is chosen so that Q + joff is a multiple of 8: then the window base of every pair sits on a padding period and
// all offsets inside a pair are compile-time constants.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace hz {

constexpr int kPolyR = 8;              // outputs per lane
constexpr int kPolyOT = 32 * kPolyR;   // outputs per tile (a pair of warps)

struct PolyPlan {
    uint32_t D, Q, qpad, npairs, half_last, joff, k0, row;
};

inline PolyPlan poly_plan(size_t ntaps, uint32_t D) {
    PolyPlan p{};
    p.D = D;
    p.Q = (uint32_t)((ntaps + D - 1) / D);        // taps per phase
    p.npairs = (p.Q + 7) / 8;                     // stage B consumes 8 taps of a phase at a time
    p.qpad = 8 * p.npairs;
    p.half_last = (p.Q % 8 != 0 && p.Q % 8 <= 4) ? 1u : 0u;  // the last group of 8 holds at most 4 real taps
    p.joff = (8 - p.Q % 8) % 8;
    p.k0 = p.npairs - 1;
    p.row = 9 * (32 + p.k0) + 8;                  // float2 per phase row: 8 (32 + k0) slots, one pad every 8, the last window's reach
    return p;
}

// [D][qpad] floats: taps_t[p][q] = h[D q + p]
inline std::vector<float> poly_taps_layout(const float *taps, size_t ntaps, const PolyPlan &p) {
    std::vector<float> t((size_t)p.D * p.qpad, 0.0f);
    for (size_t k = 0; k < ntaps; k++) t[(k % p.D) * p.qpad + k / p.D] = taps[k];
    return t;
}

#ifdef __CUDACC__
__host__ __device__
#endif
inline int poly_pos(int x) { return x + (x >> 3); }

}  // namespace hz
