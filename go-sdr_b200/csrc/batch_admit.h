// batch_admit.h -- the hazard check of a batched launch against the context's OverlapWindow (common.cuh).
#pragma once
#include <utility>
#include <vector>

#include "batch_host.h"
#include "common.cuh"

namespace hz {

// A batched launch (K buffers / streams in one kernel): may it start while earlier overlappable launches drain?  The
// buffers' spans are coalesced (batch_host.h) and go through the context's OverlapWindow like a single launch's.  If
// the window was restarted somewhere in the middle (a conflict, or it was full), it is made to hold exactly this
// launch's spans -- all of them, or the next launch is forced to be serialised -- so nothing of this launch is forgotten.
inline bool admit_spans(hzsdr_ctx *ctx, std::vector<BufSpan> spans) {
    const std::vector<BufSpan> merged = coalesce_spans(std::move(spans));
    const OverlapWindow::Span none{0, 0};
    bool may = true;
    for (const BufSpan &v : merged) {
        const OverlapWindow::Span s{v.lo, v.hi};
        may &= ctx->overlap.admit(v.write ? none : s, v.write ? s : none, ctx->overlap_pred_ok());
        ctx->overlap_launched();
    }
    if (!may) {
        ctx->overlap.n = 0;
        bool fits = true;
        for (const BufSpan &v : merged) {
            const OverlapWindow::Span s{v.lo, v.hi};
            fits = fits && ctx->overlap.push(v.write ? none : s, v.write ? s : none);
        }
        if (!fits) ctx->overlap.n = OverlapWindow::kMax;  // "full": whatever is admitted next restarts the window and goes out serialised
    }
    return may;
}

}  // namespace hz
