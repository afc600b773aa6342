// CPU-only checks of the library's host-side logic, compiled straight from the headers the CUDA
// sources include (nothing is launched): the NCO segment table against the reference's
// literal serial accumulator loop (stream/shifter.go:73-79), the cut of a segment list into kernel
// launches, the fixed-point phase conversion, and the window that decides which launches may overlap.
// Run by tests/test_host_logic.py.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <vector>

#include "../../go-sdr_b200/csrc/nco_launch.h"

namespace hz {  // the two error helpers common.cuh declares (api.cu defines them in the library)
void set_error(const char *, ...) {}
int fail(int status, const char *, ...) { return status; }
}  // namespace hz

using namespace hz;

static int g_fail = 0, g_checks = 0;
#define CHECK(cond)                                                     \
    do {                                                                \
        g_checks++;                                                     \
        if (!(cond)) {                                                  \
            g_fail++;                                                   \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); \
        }                                                               \
    } while (0)

// stream/shifter.go:73-79, literally
static void serial_ts(unsigned fs, size_t n, double ts, std::vector<double> &out) {
    const double inc = 1.0 / (double)fs, tau = M_PI * 2;
    out.resize(n);
    for (size_t i = 0; i < n; i++) {
        ts += inc;
        if (ts > tau) ts -= tau;
        out[i] = ts;
    }
}

static void test_segments(unsigned fs, size_t n, double ts0) {
    std::vector<double> want;
    serial_ts(fs, n, ts0, want);
    std::vector<HostSeg> segs;
    double ts = ts0;
    build_segments(fs, n, &ts, segs);
    CHECK(ts == want.back());  // carried accumulator: bit-equal
    size_t covered = 0;
    bool ok = true;
    for (const HostSeg &s : segs) {
        ok = ok && s.j0 == covered && s.count > 0;
        for (size_t k = 0; k < s.count && ok; k++) {
            const double v = s.step == 0.0 ? s.base : s.base + (double)(k + 1) * s.step;
            ok = v == want[s.j0 + k];  // every sample's time: bit-equal
        }
        covered += s.count;
    }
    CHECK(ok && covered == n);
    // steady state needs few segments; a fresh stream a few dozen
    CHECK(segs.size() < 200);

    // the launch plan covers [0, n) once, on block boundaries, and re-bases segments exactly
    std::vector<NcoLaunch> launches;
    const size_t align = 1024;
    CHECK(plan_nco_launches(segs, n, align, -2.5e6, launches) == HZSDR_OK);
    size_t pos = 0;
    for (const NcoLaunch &L : launches) {
        CHECK(L.first == pos && L.count > 0 && (L.first % align) == 0 && L.table.count <= kMaxSegsPerLaunch);
        uint32_t j = 0;
        for (int k = 0; k < L.table.count; k++) {
            CHECK(L.table.seg[k].j0 == j);
            j += L.table.seg[k].count;
        }
        CHECK(j == L.count);
        // first sample of the launch: fixed-point phase == frac(shift * ts) to 2^-40 turns
        const double turns = -2.5e6 * want[pos];
        const double frac = turns - std::floor(turns);
        const NcoSegment &s0 = L.table.seg[0];
        const uint64_t ph = s0.dp ? s0.p0 + s0.dp : s0.p0;
        double d = std::ldexp((double)ph, -64) - frac;
        d -= std::round(d);
        CHECK(std::fabs(d) < 1e-9);  // fp64 evaluation of `turns` itself is only good to ~1e-10 here
        pos += L.count;
    }
    CHECK(pos == n);
}

static void test_turns_fix() {
    CHECK(turns_fix(0.25, 1.0) == (1ull << 62));
    CHECK(turns_fix(1.0, 3.5) == (1ull << 63));
    CHECK(turns_fix(-0.25, 1.0) == (3ull << 62));  // frac of a negative number: in [0, 1)
    CHECK(turns_fix(1e6, 1e-6) == 0 || turns_fix(1e6, 1e-6) > (~0ull - (1ull << 20)) || turns_fix(1e6, 1e-6) < (1ull << 20));
}

static void test_overlap_window() {
    OverlapWindow w;
    auto sp = [](uintptr_t lo, uintptr_t n) { return OverlapWindow::Span{lo, lo + n}; };
    CHECK(w.admit(sp(0x1000, 0x100), sp(0x2000, 0x100)));   // first launch
    CHECK(w.admit(sp(0x1100, 0x100), sp(0x2100, 0x100)));   // disjoint: may overlap
    CHECK(w.admit(sp(0x1000, 0x100), sp(0x2200, 0x100)));   // reading what another launch reads: fine
    CHECK(!w.admit(sp(0x2000, 0x10), sp(0x3000, 0x10)));    // reads what launch 1 writes: serialised ...
    CHECK(w.n == 1);                                        // ... and the window restarts with it
    CHECK(w.admit(sp(0x5000, 0x10), sp(0x6000, 0x10)));
    CHECK(!w.admit(sp(0x7000, 0x10), sp(0x6008, 0x10)));    // write after write
    CHECK(!w.admit(sp(0x8000, 0x10), sp(0x7000, 0x10)));    // write after read
    CHECK(!w.admit(sp(0x9000, 0x10), sp(0x9000, 0x10)) == false);  // in place on fresh memory: fine
    CHECK(!w.admit(sp(0x9000, 0x10), sp(0x9000, 0x10)));    // the same buffer in place again: ordered
    // adjacent but not overlapping spans do not conflict
    CHECK(w.admit(sp(0xa000, 0x10), sp(0xa010, 0x10)));
    CHECK(w.admit(sp(0xa020, 0x10), sp(0xa030, 0x10)));
    // a full window forces a serialised launch
    OverlapWindow f;
    int admitted = 0;
    for (int i = 0; i < OverlapWindow::kMax + 1; i++) admitted += f.admit(sp(0x100000 + 64 * i, 16), sp(0x900000 + 64 * i, 16));
    CHECK(admitted == OverlapWindow::kMax && f.n == 1);
    // counter slots: distinct for every launch that can still be in flight
    OverlapWindow g;
    bool distinct = true;
    std::vector<int> seen(OverlapWindow::kSlots, -1);
    for (int i = 0; i < 3 * OverlapWindow::kSlots; i++) {
        g.admit(sp(0x100000 + 64 * (uintptr_t)i, 16), sp(0x9000000 + 64 * (uintptr_t)i, 16));
        const int s = g.slot();
        if (seen[s] >= 0 && i - seen[s] < OverlapWindow::kMax) distinct = false;
        seen[s] = i;
    }
    CHECK(distinct);
    static_assert(OverlapWindow::kSlots > OverlapWindow::kMax, "a slot must outlive the window");
}

int main() {
    test_turns_fix();
    test_overlap_window();
    test_segments(20000000u, 1u << 22, 0.0);          // C2: stream start
    test_segments(20000000u, 1u << 22, 3.9999);       // across the binade edge at 4
    test_segments(20000000u, 1u << 22, 6.2);          // across the 2*pi wrap
    test_segments(2400000u, 1u << 20, 0.0);           // C1
    test_segments(61440000u, 1u << 22, 1.999999);     // C3 / C5 rate, binade edge at 2
    test_segments(8000000u, 300000, 0.49999);
    std::printf("host logic: %d checks, %d failed\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
