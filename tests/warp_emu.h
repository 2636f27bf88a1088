// warp_emu.h — a 32-lane warp on the CPU for the host harnesses under tests/.
//
// Each lane is a ucontext fiber running the same callable; __any_sync / __ballot_sync / __shfl_sync / __syncwarp
// suspend the lane until every lane of the warp has arrived at a collective, then all resume with the combined
// result.  Between collectives the lanes run one after the other, which is a legal schedule of independent threads.
// The emulation also does what compute-sanitizer's synccheck does for these call sites: it is an error if the lanes
// wait at DIFFERENT collectives (source lines), or if a lane has returned while others still wait at a full-mask
// collective (on the GPU that is undefined behaviour / a hang).  Only full masks are supported.
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstring>
#include <functional>
#include <vector>

namespace warp_emu {

constexpr int W = 32;
constexpr size_t STACK = 256 * 1024;

struct Warp {
    enum State { READY, WAITING, DONE };
    ucontext_t main_ctx, lane_ctx[W];
    State st[W];
    int cur = -1;
    // pending collective of each waiting lane
    const void *site[W];
    int op[W];
    uint64_t val[W], res[W];
    int src[W];
    const char *error = nullptr;
    std::function<void(int)> body;
    unsigned long long collectives = 0;
};

inline Warp *&current() {
    static thread_local Warp *w = nullptr;
    return w;
}

inline void lane_entry() {
    Warp *w = current();
    const int l = w->cur;
    w->body(l);
    w->st[l] = Warp::DONE;
    swapcontext(&w->lane_ctx[l], &w->main_ctx);
}

// Run `body(lane)` on 32 lanes.  Returns nullptr, or a description of the convergence error.
inline const char *run(const std::function<void(int)> &body, unsigned long long *n_collectives = nullptr) {
    static thread_local std::vector<char> stacks(W * STACK);  // reused between runs
    Warp w;
    w.body = body;
    current() = &w;
    for (int l = 0; l < W; ++l) {
        getcontext(&w.lane_ctx[l]);
        w.lane_ctx[l].uc_stack.ss_sp = stacks.data() + (size_t)l * STACK;
        w.lane_ctx[l].uc_stack.ss_size = STACK;
        w.lane_ctx[l].uc_link = &w.main_ctx;
        makecontext(&w.lane_ctx[l], lane_entry, 0);
        w.st[l] = Warp::READY;
    }
    for (;;) {
        for (int l = 0; l < W; ++l)
            if (w.st[l] == Warp::READY) {
                w.cur = l;
                swapcontext(&w.main_ctx, &w.lane_ctx[l]);
            }
        int waiting = 0, done = 0, first = -1;
        for (int l = 0; l < W; ++l) {
            if (w.st[l] == Warp::WAITING) { ++waiting; if (first < 0) first = l; }
            else if (w.st[l] == Warp::DONE) ++done;
        }
        if (waiting == 0) break;
        if (done) { w.error = "a lane returned while others wait at a full-mask collective"; break; }
        uint64_t bits = 0;
        for (int l = 0; l < W; ++l) {
            if (w.site[l] != w.site[first] || w.op[l] != w.op[first]) { w.error = "lanes wait at different collectives (divergent call sites)"; break; }
            if (w.val[l]) bits |= 1ull << l;
        }
        if (w.error) break;
        for (int l = 0; l < W; ++l) {
            switch (w.op[l]) {
            case 0: w.res[l] = bits != 0; break;             // any
            case 1: w.res[l] = bits; break;                  // ballot
            case 2: w.res[l] = w.val[w.src[l] & (W - 1)]; break;  // shfl (idx)
            default: w.res[l] = 0; break;                    // syncwarp
            }
            w.st[l] = Warp::READY;
        }
        ++w.collectives;
    }
    current() = nullptr;
    if (n_collectives) *n_collectives = w.collectives;
    return w.error;
}

__attribute__((noinline)) inline uint64_t collective(int op, uint64_t v, int src, const void *site) {
    Warp *w = current();
    const int l = w->cur;
    w->op[l] = op; w->val[l] = v; w->src[l] = src; w->site[l] = site;
    w->st[l] = Warp::WAITING;
    swapcontext(&w->lane_ctx[l], &w->main_ctx);
    return w->res[l];
}

}  // namespace warp_emu

// ---- the CUDA spellings ---------------------------------------------------------------------------------------
// The call site is the source line of the collective (macros below), not a code address: the compiler may clone a call
// for different predecessors, which would make converged lanes look divergent.
namespace warp_emu {
inline const void *site_of(int line) { return reinterpret_cast<const void *>(static_cast<uintptr_t>(line)); }
inline int any_at(int line, unsigned, int pred) { return (int)collective(0, pred ? 1 : 0, 0, site_of(line)); }
inline unsigned ballot_at(int line, unsigned, int pred) { return (unsigned)collective(1, pred ? 1 : 0, 0, site_of(line)); }
inline void syncwarp_at(int line) { collective(3, 0, 0, site_of(line)); }
inline unsigned shfl_at(int line, unsigned, unsigned v, int src) { return (unsigned)collective(2, v, src, site_of(line)); }
inline int shfl_at(int line, unsigned, int v, int src) { return (int)(unsigned)collective(2, (unsigned)v, src, site_of(line)); }
inline double shfl_at(int line, unsigned, double v, int src) {
    uint64_t b;
    std::memcpy(&b, &v, 8);
    b = collective(2, b, src, site_of(line));
    std::memcpy(&v, &b, 8);
    return v;
}
inline int lane_id() { return current()->cur; }
template <typename T> inline T shfl_xor_at(int line, unsigned m, T v, int lane_mask) { return shfl_at(line, m, v, lane_id() ^ lane_mask); }
}  // namespace warp_emu
#define __shfl_xor_sync(mask, v, lm) warp_emu::shfl_xor_at(__LINE__, mask, v, lm)
#define __any_sync(mask, pred) warp_emu::any_at(__LINE__, mask, pred)
#define __ballot_sync(mask, pred) warp_emu::ballot_at(__LINE__, mask, pred)
#define __shfl_sync(mask, v, src) warp_emu::shfl_at(__LINE__, mask, v, src)
#define __syncwarp(...) warp_emu::syncwarp_at(__LINE__)
