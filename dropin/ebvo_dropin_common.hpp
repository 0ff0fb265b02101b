// Shared by the drop-in translation units: one process-wide ebvo context (one GPU), grown on demand.
//
// The reference classes have no room for a context handle (their members are fixed by the reference headers), and
// Pipeline constructs exactly one of each (Pipeline.cpp:15-21), so a process-wide context created on first use
// mirrors the reference's lifetime.  No CPU fallback: when the context cannot be created the caller reports the
// error the reference way (a LOG_ERROR-style print) and returns an empty result.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "ebvo_b200.h"

namespace ebvo_dropin {

struct Shared {
    std::mutex mu;
    ebvo_ctx* ctx = nullptr;
    int w = 0, h = 0, edges = 0;
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

// Returns a context able to hold w x h images and `edges` edges per image (nullptr + message on failure).
inline ebvo_ctx* context(int w, int h, int edges)
{
    Shared& s = shared();
    std::lock_guard<std::mutex> lk(s.mu);
    if (s.ctx && w <= s.w && h <= s.h && edges <= s.edges) return s.ctx;
    if (s.ctx) { ebvo_destroy(s.ctx); s.ctx = nullptr; }
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
        return nullptr;
    }
    s.ctx = c; s.w = W; s.h = H; s.edges = E;
    return c;
}

// Tightly packed copy of an 8-bit single-channel image given (data, rows, cols, step in bytes).
inline std::vector<unsigned char> packed_u8(const unsigned char* data, int rows, int cols, size_t step)
{
    std::vector<unsigned char> out((size_t)rows * cols);
    for (int r = 0; r < rows; ++r) std::memcpy(out.data() + (size_t)r * cols, data + (size_t)r * step, (size_t)cols);
    return out;
}

}  // namespace ebvo_dropin
