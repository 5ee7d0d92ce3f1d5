// rig_timeline.h -- optional GPU timeline of the library's kernels (RIG_TIMELINE=1): every launch is bracketed by two CUDA events,
// rig_timeline_dump writes "name,stream,start_ms,end_ms" relative to the first event.  A debugging aid for the pipelined end-to-end
// loops (which kernel waits for which), off by default: two cudaEventRecord calls per launch.
#pragma once
#include <cuda_runtime.h>

namespace rig {
bool timeline_on();
void timeline_mark(const char *name, cudaStream_t st, bool begin);

struct TimelineScope {
    const char *name;
    cudaStream_t st;
    bool on;
    TimelineScope(const char *n, cudaStream_t s) : name(n), st(s), on(timeline_on())
    {
        if (on) timeline_mark(name, st, true);
    }
    ~TimelineScope()
    {
        if (on) timeline_mark(name, st, false);
    }
};
}  // namespace rig
