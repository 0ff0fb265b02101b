// Shared by the drop-in translation units: one process-wide ebvo context (one GPU), grown on demand, handed out under a lock.
//
// The reference classes have no room for a context handle (their members are fixed by the reference headers), and
// Pipeline constructs exactly one of each (Pipeline.cpp:15-21), so a process-wide context created on first use
// mirrors the reference's lifetime.  The detector, the matcher and the quad tracker share it (a context holds the buffers
// of all three; one per class instance would triple the device memory).  A caller holds a Lease for the duration of its
// C-ABI calls: growing the context destroys the old one, which must not happen under another caller's feet.
// No CPU fallback: when the context cannot be created the caller reports the error the reference way (a LOG_ERROR-style
// print) and returns an empty result.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "ebvo_b200.h"

namespace ebvo_dropin {

// Grow-only host buffer for per-frame results, page-locked through the library (ebvo_host_alloc) so that the device -> host
// copies of a frame's mates, patches and descriptors are direct DMA transfers; plain memory when that allocation fails.
// Kept for the life of the process like the context (a fresh 90 MB of zero-filled vectors per frame cost 8 ms).
struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    void* ensure(size_t bytes)
    {
        if (bytes <= cap) return p;
        if (p) { if (pinned) ebvo_host_free(p); else std::free(p); }
        const size_t want = bytes + bytes / 4;
        p = ebvo_host_alloc(want);
        pinned = p != nullptr;
        if (!p) p = std::malloc(want);
        cap = p ? want : 0;
        return p;
    }
    float* floats(size_t n) { return static_cast<float*>(ensure(n * sizeof(float))); }
};

// what get_Stereo_Edge_Pairs already computed for finalize_stereo_edge_mates (same call, same upload): right patches and
// right descriptor pairs of the mates (in Shared::r_plus / r_minus / r_desc), valid while `frame` / `n` / the first and
// last mate still identify the same result
struct FinalizeCache {
    const void* frame = nullptr;
    size_t n = 0;
    double x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    bool has_desc = false;
};

struct Shared {
    std::recursive_mutex mu;
    ebvo_ctx* ctx = nullptr;
    int w = 0, h = 0, edges = 0;
    FinalizeCache fin;
    HostBuf mates, l_plus, l_minus, l_desc, r_plus, r_minus, r_desc;      // stereo drop-in: results of a frame
    HostBuf quads, tq_desc[4];                                             // quad-tracking drop-in: result records, descriptor staging
    size_t tq_last = 0;                                                    // quads of the previous call (sizes the next result buffer)
};
inline Shared& shared()
{
    static Shared s;
    return s;
}

// SIFT gate / BNB-SIFT / finalisation descriptors (Stereo_Matches.cpp:655-787,1452,1627-1635) run on the device as in the
// reference's default flow; EBVO_DROPIN_SIFT=0 in the environment selects the "SIFT-off" parity configuration.
inline bool sift_enabled()
{
    const char* e = std::getenv("EBVO_DROPIN_SIFT");
    return !(e && e[0] == '0');
}

// A context able to hold w x h images and `edges` edges per image, locked for the lifetime of the object
// (ctx == nullptr + message on failure).
struct Lease {
    std::unique_lock<std::recursive_mutex> lk;
    ebvo_ctx* ctx = nullptr;
    Lease(int w, int h, int edges) : lk(shared().mu)
    {
        Shared& s = shared();
        if (s.ctx && w <= s.w && h <= s.h && edges <= s.edges) { ctx = s.ctx; return; }
        if (s.ctx) { ebvo_destroy(s.ctx); s.ctx = nullptr; s.fin = FinalizeCache(); }
        const int W = w > s.w ? w : s.w, H = h > s.h ? h : s.h;
        int E = edges > s.edges ? edges : s.edges;
        if (E < 1 << 16) E = 1 << 16;
        ebvo_ctx* c = nullptr;
        ebvo_params prm;
        ebvo_params_default(&prm);
        prm.sift_mode = sift_enabled() ? 1 : 0;
        const int rc = ebvo_create(&c, 0, W, H, 1, E, &prm);
        if (rc != EBVO_OK) {
            std::printf("\033[1;31m[ERROR] ebvo_create failed (%d): %s\033[0m\n", rc, c ? ebvo_last_error(c) : "no CUDA device");
            if (c) ebvo_destroy(c);
            return;
        }
        ebvo_set_profiling(c, 1);      // per-kernel CUDA events: time_conv / time_nms and Timing_Statistics are filled from them
        s.ctx = c; s.w = W; s.h = H; s.edges = E;
        ctx = c;
    }
};

// milliseconds the named kernels took in the last profiled call (prefix match)
inline double kernel_ms(ebvo_ctx* c, const char* const* prefixes, int np)
{
    const char** names = nullptr; const float* ms = nullptr; const int* launches = nullptr; int n = 0;
    if (ebvo_get_kernel_times(c, &names, &ms, &launches, &n) != EBVO_OK) return 0.0;
    double t = 0.0;
    for (int k = 0; k < n; ++k)
        for (int p = 0; p < np; ++p)
            if (std::strncmp(names[k], prefixes[p], std::strlen(prefixes[p])) == 0) { t += ms[k]; break; }
    return t;
}

// EBVO_DROPIN_TRACE=1: wall-clock milliseconds of the drop-in's own sections on stderr (where a member's time goes)
struct Trace {
    bool on;
    const char* who;
    std::chrono::steady_clock::time_point t0;
    explicit Trace(const char* w) : on(std::getenv("EBVO_DROPIN_TRACE") != nullptr), who(w), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what)
    {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[ebvo trace] %s: %s %.3f ms\n", who, what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// Tightly packed copy of an 8-bit single-channel image given (data, rows, cols, step in bytes).
inline std::vector<unsigned char> packed_u8(const unsigned char* data, int rows, int cols, size_t step)
{
    std::vector<unsigned char> out((size_t)rows * cols);
    for (int r = 0; r < rows; ++r) std::memcpy(out.data() + (size_t)r * cols, data + (size_t)r * step, (size_t)cols);
    return out;
}

}  // namespace ebvo_dropin
