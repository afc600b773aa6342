// CPU-only checks of the library's host-side logic, compiled straight from the headers the CUDA
// sources include (nothing is launched): the NCO segment table against the reference's
// literal serial accumulator loop (stream/shifter.go:73-79), the cut of a segment list into kernel
// launches, the fixed-point phase conversion, and the window that decides which launches may overlap.
// Run by tests/test_host_logic.py.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <vector>

#include "../../go-sdr_b200/csrc/batch_admit.h"
#include "../../go-sdr_b200/csrc/batch_host.h"
#include "../../go-sdr_b200/csrc/nco_launch.h"
#include "../../go-sdr_b200/csrc/poly_host.h"

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

// host side of the batched launches (batch_host.h): descriptors, the cut into parameter-sized launches, the hazard sweep
static void test_batch_host() {
    // fill_desc: segments land in the pool at seg_off, dp_nom is the step of the LONGEST linear segment
    NcoTable t{};
    t.count = 4;
    t.seg[0] = NcoSegment{0, 1, 11, 0};        // single-step segments carry dp = 0 and never dominate
    t.seg[1] = NcoSegment{1, 1000, 22, 7};
    t.seg[2] = NcoSegment{1001, 5000, 33, 9};
    t.seg[3] = NcoSegment{6001, 4000, 44, 8};
    std::vector<NcoSegment> pool(16);
    StreamDesc d{};
    int a = 0, b = 0;
    fill_desc(d, pool.data(), 5, &a, &b, t);
    CHECK(d.src == (const uint8_t *)&a && d.dst == (void *)&b && d.seg_off == 5 && d.count == 4 && d.dp_nom == 9);
    CHECK(pool[5].p0 == 11 && pool[8].p0 == 44 && pool[4].p0 == 0 && pool[9].p0 == 0);
    NcoTable only_steps{};
    only_steps.count = 2;
    only_steps.seg[0] = NcoSegment{0, 1, 1, 0};
    only_steps.seg[1] = NcoSegment{1, 1, 2, 0};
    fill_desc(d, pool.data(), 0, &a, &b, only_steps);
    CHECK(d.dp_nom == 0 && d.count == 2);

    // plan_param_launches: every stream exactly once, in order, within both limits
    auto check_plan = [&](const std::vector<int> &segs, uint32_t ms, uint32_t mg) {
        const auto plan = plan_param_launches(segs, ms, mg);
        uint32_t next = 0;
        bool ok = true;
        for (const auto &c : plan) {
            ok &= c.first == next && c.second >= 1 && c.second <= ms;
            uint32_t ns = 0;
            for (uint32_t i = 0; i < c.second; i++) ns += (uint32_t)segs[c.first + i];
            ok &= ns <= mg || c.second == 1;
            next = c.first + c.second;
        }
        return ok && next == segs.size();
    };
    CHECK(plan_param_launches({}, 64, 416).empty());
    CHECK(check_plan(std::vector<int>(512, 1), 64, 416) && plan_param_launches(std::vector<int>(512, 1), 64, 416).size() == 8);
    CHECK(check_plan(std::vector<int>(64, 3), 64, 416) && plan_param_launches(std::vector<int>(64, 3), 64, 416).size() == 1);
    CHECK(check_plan(std::vector<int>(64, 88), 64, 416) && plan_param_launches(std::vector<int>(64, 88), 64, 416).size() == 16);  // stream start: 4 per launch
    std::vector<int> mixed(200, 1);
    mixed[0] = 88, mixed[17] = 79, mixed[18] = 13, mixed[199] = 112;
    CHECK(check_plan(mixed, 64, 416));
    CHECK(check_plan({500}, 64, 416));  // over the limit on its own: still one launch (the callers never pass such a table)

    // write_conflict: a write span may touch nothing else; reads may share
    auto S = [](uintptr_t lo, uintptr_t len, bool w) { return BufSpan{lo, lo + len, w}; };
    CHECK(!write_conflict({}));
    CHECK(!write_conflict({S(0x1000, 0x100, false), S(0x1000, 0x100, false), S(0x1080, 0x100, false)}));  // reads overlap: fine
    CHECK(!write_conflict({S(0x1000, 0x100, false), S(0x1100, 0x100, true), S(0x1200, 0x100, true)}));   // adjacent: fine
    CHECK(write_conflict({S(0x1000, 0x100, true), S(0x10ff, 0x10, true)}));                                // write / write
    CHECK(write_conflict({S(0x1000, 0x100, false), S(0x10ff, 0x10, true)}));                               // write into a read
    CHECK(write_conflict({S(0x2000, 0x10, true), S(0x1000, 0x2000, false)}));                              // read that covers a write
    CHECK(write_conflict({S(0x1000, 0x10, false), S(0x3000, 0x10, false), S(0x1000, 0x3000, true)}));      // one big write over everything
    CHECK(!write_conflict({S(0x1000, 0, true), S(0x1000, 0x10, true)}));                                   // empty spans do not count
    std::vector<BufSpan> many;  // 512 streams: src and dst regions interleaved, no conflict; then one destination aliased
    for (uintptr_t k = 0; k < 512; k++) many.push_back(S(0x10000000 + k * 0x5000, 0x4000, false)), many.push_back(S(0x10000000 + k * 0x5000 + 0x4000, 0x1000, true));
    CHECK(!write_conflict(many));
    many[2 * 300 + 1] = many[2 * 7 + 1];
    CHECK(write_conflict(many));
}

// batched launches in the OverlapWindow (batch_admit.h): coalesced spans, nothing of a launch forgotten
static void test_admit_spans() {
    auto S = [](uintptr_t lo, uintptr_t len, bool w) { return BufSpan{lo, lo + len, w}; };
    // coalesce: neighbours of the same kind merge, kinds stay apart, empties vanish
    {
        const auto m = coalesce_spans({S(0x3000, 0x1000, false), S(0x1000, 0x1000, false), S(0x2000, 0x1000, false), S(0x9000, 0x100, true),
                                       S(0x9100, 0x100, true), S(0x9300, 0x100, true), S(0x5000, 0, false), S(0x4000, 0x10, true)});
        CHECK(m.size() == 4);
        CHECK(!m[0].write && m[0].lo == 0x1000 && m[0].hi == 0x4000);  // three adjacent reads
        CHECK(m[1].write && m[1].lo == 0x4000 && m[1].hi == 0x4010);   // a write that touches the reads stays a write span of its own
        CHECK(m[2].write && m[2].lo == 0x9000 && m[2].hi == 0x9200 && m[3].lo == 0x9300);
    }
    hzsdr_ctx ctx{};
    ctx.api_seq = 1;
    ctx.scheme_seq = 1;  // the previous operation on the stream was an overlappable launch of this call
    // 64 buffers side by side: two window entries, launch may overlap
    std::vector<BufSpan> a;
    for (uintptr_t k = 0; k < 64; k++) a.push_back(S(0x10000000 + k * 0x800000, 0x800000, false)), a.push_back(S(0x40000000 + k * 0x100000, 0x100000, true));
    CHECK(admit_spans(&ctx, a) && ctx.overlap.n == 2);
    // the same sources, other destinations: read-read is no hazard
    std::vector<BufSpan> b;
    for (uintptr_t k = 0; k < 64; k++) b.push_back(S(0x10000000 + k * 0x800000, 0x800000, false)), b.push_back(S(0x50000000 + k * 0x100000, 0x100000, true));
    CHECK(admit_spans(&ctx, b) && ctx.overlap.n == 4);
    // writes into the first launch's destinations: serialised, and the window holds exactly this launch (2 entries)
    CHECK(!admit_spans(&ctx, a) && ctx.overlap.n == 2);
    CHECK(ctx.overlap.reads[0].lo == 0x10000000 && ctx.overlap.writes[1].lo == 0x40000000);
    // the next launch still sees ALL of it: a write into the LAST buffer's destination is caught
    CHECK(!admit_spans(&ctx, {S(0x70000000, 0x10, false), S(0x40000000 + 63 * 0x100000, 0x10, true)}));
    // scattered buffers, more than the window holds: the launch is serialised and so is whatever comes next
    std::vector<BufSpan> big;
    for (uintptr_t k = 0; k < (uintptr_t)OverlapWindow::kMax + 8; k++) big.push_back(S(0x80000000 + k * 0x2000, 0x1000, true));
    CHECK(!admit_spans(&ctx, big));
    CHECK(!admit_spans(&ctx, {S(0x1000, 0x10, false), S(0x2000, 0x10, true)}));  // forced: the window could not hold the big launch
    CHECK(admit_spans(&ctx, {S(0x3000, 0x10, false), S(0x4000, 0x10, true)}));   // and then life goes on
}

// The polyphase decimator's index plan (poly_host.h), replayed on the CPU exactly as k_polyphase_chain moves data: stage A
// scatters a tile's inputs by phase into padded rows, stage B walks 15-slot windows down each row 8 taps at a time.
// Every kept output must equal the direct FIR sum_k h[k] y[D i - k] -- for filter lengths and decimation factors that
// exercise every remainder of Q mod 8, D < 8 (several wraps per group), D > 32, a single tap, and no decimation.
static void test_polyphase_plan(size_t ntaps, uint32_t D) {
    const PolyPlan P = poly_plan(ntaps, D);
    CHECK((P.Q + P.joff) % 8 == 0 && P.qpad >= P.Q && P.qpad % 8 == 0 && P.k0 + 1 == P.npairs);
    std::vector<float> h(ntaps);
    for (size_t k = 0; k < ntaps; k++) h[k] = 0.25f + 0.001f * (float)((k * 37) % 101);
    const std::vector<float> tt = poly_taps_layout(h.data(), ntaps, P);
    // a tile in the middle of a stream: y[n] for n in [n_lo, n_hi), outputs i0 .. i0 + 255
    const long i0 = 1000;
    const long a_len = (long)D * (kPolyOT + P.Q - 1);
    const long n_lo = (long)D * (i0 - (long)(P.Q - 1)) - (long)(D - 1);  // stream index of tile input a = 0
    auto y = [&](long n) { return (double)((n * 2654435761u) % 1009) / 1009.0 - 0.5; };
    std::vector<double> U((size_t)D * P.row, 0.0);
    std::vector<char> written(U.size(), 0);
    bool in_row = true, once = true;
    for (long a = 0; a < a_len; a++) {  // stage A
        const uint32_t p = D - 1 - (uint32_t)(a % D), jj = (uint32_t)(a / D);
        const size_t pos = (size_t)poly_pos((int)(jj + P.joff));
        in_row &= pos < P.row;
        const size_t at = (size_t)p * P.row + pos;
        once &= !written[at];
        written[at] = 1;
        U[at] = y(n_lo + a);
    }
    CHECK(in_row && once);
    double worst = 0.0;
    bool reads_ok = true;
    for (int lane = 0; lane < 32; lane++)
        for (int r = 0; r < kPolyR; r++) {
            double acc = 0.0;
            for (uint32_t p = 0; p < D; p++)
                for (uint32_t m = 0; m < P.npairs; m++) {
                    const bool half = P.half_last && m + 1 == P.npairs;
                    for (int c = 0; c < (half ? 4 : 8); c++) {
                        const int x = r + 7 - c;  // (c >= 4: x = r + 3 - (c - 4), the same number)
                        const size_t at = (size_t)p * P.row + 9 * ((size_t)lane + P.k0 - m) + (size_t)x + (size_t)(x >> 3);
                        reads_ok &= 9 * ((size_t)lane + P.k0 - m) + (size_t)x + (size_t)(x >> 3) < P.row;
                        acc += (double)tt[(size_t)p * P.qpad + 8 * m + c] * U[at];
                    }
                }
            double want = 0.0;
            const long i = i0 + 8 * lane + r;
            for (size_t k = 0; k < ntaps; k++) want += (double)h[k] * y((long)D * i - (long)k);
            worst = std::fmax(worst, std::fabs(acc - want));
        }
    CHECK(reads_ok);
    CHECK(worst < 1e-9);
    if (worst >= 1e-9) std::printf("  polyphase plan taps %zu D %u: worst |diff| %.3g\n", ntaps, D, worst);
}

int main() {
    test_turns_fix();
    for (auto td : std::vector<std::pair<size_t, uint32_t>>{{255, 10}, {255, 16}, {127, 10}, {63, 8}, {1, 7}, {31, 1}, {64, 3}, {1023, 48}, {2047, 64},
                                                             {9, 2}, {17, 2}, {25, 2}, {33, 2}, {41, 5}, {49, 6}, {57, 7}, {65, 8}, {100, 33}, {4095, 16}})
        test_polyphase_plan(td.first, td.second);
    test_batch_host();
    test_admit_spans();
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
